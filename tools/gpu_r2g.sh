#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_lookup3.py 2>&1 | tail -8
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2g_pytest_all.log 2>&1
echo "pytest all rc=$?" >> gpurun_out/r2g_pytest_all.log
tail -6 gpurun_out/r2g_pytest_all.log
