#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lookup_encode.py tests/test_gpu_e2e.py -x -q -m gpu > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2o_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-reference-gpu --no-whole-forward --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lookup_enc1" -c 40 --csv --log-file gpurun_out/r2o_lookup_times.csv python bench.py --profile-step --warmup 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open("gpurun_out/r2o_lookup_times.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[h]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(list)
for r in rows[h+1:]:
    if len(r)>vi: agg[r[ki][:60]].append(float(r[vi].replace(",","")))
for k,v in agg.items(): print(k, len(v), "mean us", sum(v)/len(v)/1000 if max(v)>1000 else sum(v)/len(v), "min", min(v))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lookup_enc1" -s 4 -c 2 -f -o gpurun_out/r2n_lookup python bench.py --profile-step --warmup 1 > gpurun_out/r2n_ncu_lookup.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2n_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms/step", round(d["ms_per_step"],3))
for k,v in d["kernels"].items(): print(" ", k, round(v["ms_per_step"],3), round(v["avg_us"],1))
PY
