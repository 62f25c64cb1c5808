#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lookup_encode.py tests/test_gpu_e2e.py -x -q -m gpu > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2p_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-reference-gpu --no-whole-forward --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2p_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms/step", round(d["ms_per_step"],3), "lookup", round(d["kernels"]["lookup"]["avg_us"],1))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lookup_enc1" -c 40 --csv --log-file gpurun_out/r2p_lookup_times.csv python bench.py --profile-step --warmup 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"enc_" -c 400 --csv --log-file gpurun_out/r2p_enc_times.csv python tools/encoder_timing.py > gpurun_out/r2p_enc.log 2>&1
python - <<'PY'
import csv, collections
for f in ("gpurun_out/r2p_lookup_times.csv","gpurun_out/r2p_enc_times.csv"):
    rows=list(csv.reader(open(f)))
    h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
    hdr=rows[h]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); mi=hdr.index("Metric Name"); ii=hdr.index("ID")
    agg=collections.OrderedDict()
    for r in rows[h+1:]:
        if len(r)>vi and r[mi]=="gpu__time_duration.sum":
            agg.setdefault(r[ki][:70],[]).append(float(r[vi].replace(",","")))
    for k,v in agg.items(): print(k, len(v), "mean", round(sum(v)/len(v)), "min", min(v))
PY
