#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raft.py tests/test_gpu_encoder.py -x -q -s -m gpu > gpurun_out/r2j_pytest_raft.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest_raft.log
grep -E "rel L1|passed|failed|rc=|Error|error" gpurun_out/r2j_pytest_raft.log | tail -12
timeout 300 python tools/encoder_timing.py 2>&1 | tail -4
