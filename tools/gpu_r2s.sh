#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
CER_L2_PERSIST=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-reference-gpu --no-whole-forward --no-cpu-baseline > gpurun_out/r2s_bench_l2_$v.json 2> gpurun_out/r2s_bench_l2_$v.err
echo "bench L2=$v rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2s_bench_l2_$v.json").read().strip().splitlines()[-1])
print("L2_PERSIST=$v value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms/step", round(d["ms_per_step"],3))
for k,v in d["kernels"].items(): print(" ", k, round(v["ms_per_step"],3), round(v["avg_us"],1))
print(d["build_roofline"].get("coherent"))
PY
done
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_update.py -x -q -m gpu 2>&1 | tail -2
