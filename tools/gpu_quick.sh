#!/bin/bash
# quick check after a conv-kernel change: conv / loop / full-size tests + the headline bench without the slow legs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_update.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-reference-gpu --no-whole-forward --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/quick_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms/step", round(d["ms_per_step"],3))
for k,v in d["kernels"].items(): print(" ", k, round(v["ms_per_step"],3), round(v["avg_us"],1))
PY
