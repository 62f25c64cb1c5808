#!/bin/bash
# tile-flag A/B: flag tests first (bounded), then the whole suite, then bench with and without flags
TAG=${1:-r1g}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k tile_flag > $OUT/${TAG}_pytest_flags.log 2>&1; echo "flags pytest exit $?" >> $OUT/${TAG}_pytest_flags.log; tail -5 $OUT/${TAG}_pytest_flags.log
if grep -q "flags pytest exit 0" $OUT/${TAG}_pytest_flags.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
  timeout 400 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_flags.json 2> $OUT/${TAG}_bench_flags.err
fi
CER_TILE_FLAGS=0 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_noflags.json 2> $OUT/${TAG}_bench_noflags.err
for f in $OUT/${TAG}_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["e2e"]["value"], {k:round(v["avg_us"],1) for k,v in d["kernels"].items()})
except Exception as e: print("bad", e)
PY
done
