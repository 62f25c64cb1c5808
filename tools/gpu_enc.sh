#!/bin/bash
# encoder: parity tests, per-image timing, ncu launch list of the encoder kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_raft.py -x -q -m gpu > gpurun_out/enc_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/enc_pytest.log
timeout 600 python tools/encoder_timing.py > gpurun_out/enc_timing.log 2>&1; cat gpurun_out/enc_timing.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"enc_" -c 400 --csv --log-file gpurun_out/enc_times.csv python tools/encoder_timing.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open("gpurun_out/enc_times.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[h]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); mi=hdr.index("Metric Name")
agg=collections.OrderedDict()
for r in rows[h+1:]:
    if len(r)>vi and r[mi]=="gpu__time_duration.sum":
        agg.setdefault(r[ki][:70],[]).append(float(r[vi].replace(",","")))
for k,v in agg.items(): print(k, len(v), "mean", round(sum(v)/len(v)), "min", min(v))
PY
