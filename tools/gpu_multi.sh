#!/bin/bash
# multi-GPU: sharded-build parity tests + bench legs. usage: gpu_r2h_multi.sh N
N=$1
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multigpu.py -q -s -m gpu > gpurun_out/r2h_pytest_multigpu_n$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest_multigpu_n$N.log
tail -4 gpurun_out/r2h_pytest_multigpu_n$N.log
for cfg in cfg2 cfg4 cfg5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --config $cfg > gpurun_out/r2h_bench_${cfg}_n$N.json 2> gpurun_out/r2h_bench_${cfg}_n$N.err
  echo "bench $cfg N=$N rc=$?"
done
python - <<PY
import json
for cfg in ("cfg2","cfg4","cfg5"):
    try:
        d=json.loads(open(f"gpurun_out/r2h_bench_{cfg}_n$N.json").read().strip().splitlines()[-1])
        print(cfg, "N=$N", d["config"]["parallelism"], "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), {k:d[k] for k in ("sharded","replicas") if k in d})
    except Exception as e: print(cfg, "ERR", e)
PY
