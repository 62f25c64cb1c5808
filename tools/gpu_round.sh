#!/bin/bash
# One gpurun call: GPU tests, bench A/B of the switchable kernels, ncu launch list + --set full captures.
# Everything lands in gpurun_out/ (tag = $1).  Numbers printed under ncu are never bench values.
TAG=${1:-r1b}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi
timeout 400 python bench.py --steps 30 --warmup 3 > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
for v in ${VARIANTS:-5 1}; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --conv-variant $v > $OUT/${TAG}_bench_conv$v.json 2> $OUT/${TAG}_bench_conv$v.err
done
CER_LOOKUP=v2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_lookupv2.json 2> $OUT/${TAG}_bench_lookupv2.err
for f in $OUT/${TAG}_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["e2e"]["value"], {k:round(v["avg_us"],1) for k,v in d["kernels"].items()})
except Exception as e: print("bad", e)
PY
done
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 513 -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --profile-step --warmup 3 > $OUT/${TAG}_launches.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:lookup_enc1_v3 -s 96 -c 1 -f -o $OUT/${TAG}_full_lookup64 \
    python bench.py --profile-step --warmup 3 > $OUT/${TAG}_full_lookup64.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:lookup_enc1_v3 -s 112 -c 1 -f -o $OUT/${TAG}_full_lookup44 \
    python bench.py --profile-step --warmup 3 > $OUT/${TAG}_full_lookup44.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:build_volume_h16 -s 6 -c 2 -f -o $OUT/${TAG}_full_build \
    python bench.py --profile-step --warmup 3 > $OUT/${TAG}_full_build.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 384 -c 4 -f -o $OUT/${TAG}_full_convs \
    python bench.py --profile-step --warmup 3 > $OUT/${TAG}_full_convs.log 2>&1
fi
ls -la $OUT | tail -30
