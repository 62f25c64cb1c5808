#!/bin/bash
# round 2, run A: reference-on-GPU parity tests + bench with the reference_gpu leg
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests/test_gpu_reference.py -x -q -s -m gpu > gpurun_out/r2a_pytest_ref.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest_ref.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_pytest_ref.log
