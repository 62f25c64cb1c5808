#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/build_profile.py 2>&1 | tail -24
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_pytest_all.log 2>&1
echo "pytest all rc=$?" >> gpurun_out/r2e_pytest_all.log
tail -5 gpurun_out/r2e_pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-reference-gpu > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"
