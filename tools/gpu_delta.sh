#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_update.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_reference.py -x -q -m gpu > gpurun_out/delta_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/delta_pytest.log
for v in half full; do
CER_DELTA=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-reference-gpu --no-whole-forward --no-cpu-baseline > gpurun_out/delta_bench_$v.json 2> gpurun_out/delta_bench_$v.err
echo "bench $v rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/delta_bench_$v.json").read().strip().splitlines()[-1])
print("CER_DELTA=$v value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms/step", round(d["ms_per_step"],3), "delta", round(d["kernels"]["conv_delta"]["avg_us"],1), "lookup", round(d["kernels"]["lookup"]["avg_us"],1))
PY
done
