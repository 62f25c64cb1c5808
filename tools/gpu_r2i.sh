#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_fusion.py -x -q -s -m gpu > gpurun_out/r2i_pytest_enc.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest_enc.log
grep -E "rel L1|fusion\(\)|passed|failed|rc=|Error|error" gpurun_out/r2i_pytest_enc.log | tail -20
