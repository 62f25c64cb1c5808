#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"enc_conv3x3" -s 8 -c 8 -f -o gpurun_out/r2r_enc python tools/encoder_timing.py > gpurun_out/r2r_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2r_ncu.log
