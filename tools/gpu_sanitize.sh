#!/bin/bash
# compute-sanitizer memcheck over the small-size GPU tests of the kernels touched this round
mkdir -p gpurun_out
for t in "tests/test_gpu_lookup_encode.py" "tests/test_gpu_update.py" "tests/test_gpu_e2e.py -k golden" "tests/test_gpu_encoder.py" "tests/test_gpu_build_staged.py -k staged_equals_gather"; do
  name=$(echo $t | sed 's/[^a-z_]/_/g' | cut -c1-40)
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python -m pytest $t -x -q -m gpu -p no:cacheprovider > gpurun_out/san_$name.log 2>&1
  echo "$t rc=$? $(grep -c 'Invalid\|out of bounds\|misaligned' gpurun_out/san_$name.log) findings; $(tail -1 gpurun_out/san_$name.log)"
done
grep -h -A6 "Invalid\|misaligned" gpurun_out/san_*.log | head -60
