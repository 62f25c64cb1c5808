#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over small GPU tests of the kernels that alias / hand over shared memory
mkdir -p gpurun_out
for t in "tests/test_gpu_lookup_encode.py -k smooth" "tests/test_gpu_update.py -k vs_autocast" "tests/test_gpu_encoder.py"; do
  name=$(echo $t | sed 's/[^a-z_]/_/g' | cut -c1-40)
  timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --launch-timeout 0 python -m pytest $t -x -q -m gpu -p no:cacheprovider > gpurun_out/race_$name.log 2>&1
  echo "$t rc=$? ; $(grep -c 'Race reported' gpurun_out/race_$name.log) races; $(tail -1 gpurun_out/race_$name.log)"
done
grep -h -A3 "Race reported" gpurun_out/race_*.log | grep -v "Host Frame" | sort | uniq -c | head -20
