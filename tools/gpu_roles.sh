#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_update.py tests/test_gpu_e2e.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/conv_roles.py > gpurun_out/conv_roles.txt 2>&1; cat gpurun_out/conv_roles.txt
