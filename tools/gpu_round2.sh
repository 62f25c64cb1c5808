#!/bin/bash
# Second-kind GPU round: tests, default bench, build-reuse A/B, role profile + per-kernel scaling.
TAG=${1:-r1e}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
timeout 400 python bench.py --steps 30 --warmup 3 > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
CER_BUILD_REUSE=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_noreuse.json 2> $OUT/${TAG}_bench_noreuse.err
CER_BUILD_REUSE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_reuseall.json 2> $OUT/${TAG}_bench_reuseall.err
for f in $OUT/${TAG}_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["e2e"]["value"], {k:round(v["avg_us"],1) for k,v in d["kernels"].items()})
except Exception as e: print("bad", e)
PY
done
for v in 1 2; do timeout 200 python tools/conv_roles.py $v > $OUT/${TAG}_roles_v$v.txt 2>&1; cat $OUT/${TAG}_roles_v$v.txt; done
timeout 300 python tools/kernel_scaling.py > $OUT/${TAG}_scaling.txt 2>&1; cat $OUT/${TAG}_scaling.txt
