#!/bin/bash
# lookup v4: op + loop tests, bench without the slow legs, ncu full of the lookup kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lookup_encode.py tests/test_gpu_e2e.py tests/test_gpu_update.py tests/test_gpu_reference.py -x -q -s -m gpu > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2n_pytest.log; grep -n "warp-autonomous\|lookup variants\|rel L1\|FAILED\|Error" gpurun_out/r2n_pytest.log | head -40
timeout 600 python bench.py --steps 20 --warmup 5 --no-reference-gpu --no-whole-forward --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lookup_enc1" -s 4 -c 2 -f -o gpurun_out/r2n_lookup python bench.py --profile-step --warmup 1 > gpurun_out/r2n_ncu_lookup.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2n_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms/step", round(d["ms_per_step"],3))
for k,v in d["kernels"].items(): print(" ", k, round(v["ms_per_step"],3), round(v["avg_us"],1))
PY
