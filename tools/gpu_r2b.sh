#!/bin/bash
# round 2, run B: first correctness run of the staged tcgen05 build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_build_staged.py -x -q -s -m gpu > gpurun_out/r2b_pytest_build.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest_build.log
tail -c 2500 gpurun_out/r2b_pytest_build.log
