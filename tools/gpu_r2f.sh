#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_build_staged.py -x -q -s -m gpu > gpurun_out/r2f_pytest_build.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest_build.log
grep -E "cfg|noise|passed|failed|rc=|Error" gpurun_out/r2f_pytest_build.log | tail -12
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest_all.log 2>&1
echo "pytest all rc=$?" >> gpurun_out/r2f_pytest_all.log
tail -5 gpurun_out/r2f_pytest_all.log
for sp in 4 0; do
CER_BUILD_SPREAD=$sp timeout 600 python bench.py --steps 20 --warmup 5 --no-reference-gpu --no-cpu-baseline > gpurun_out/r2f_bench_spread$sp.json 2> gpurun_out/r2f_bench_spread$sp.err
echo "bench spread $sp rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-reference-gpu --no-cpu-baseline --build-variant 1 > gpurun_out/r2f_bench_gather.json 2> gpurun_out/r2f_bench_gather.err
python - <<'PY'
import json
for n in ("spread4","spread0","gather"):
    try:
        d=json.loads(open(f"gpurun_out/r2f_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"],2), "depth-maps/s; build ms/step", round(d["kernels"]["volume_build"]["ms_per_step"],3), "lookup", round(d["kernels"]["lookup"]["ms_per_step"],3))
    except Exception as e: print(n, "ERR", e)
PY
