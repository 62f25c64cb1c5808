#!/bin/bash
# round-end evidence: full GPU suite, smoke, headline bench (all legs), other BASELINE configs, ncu launch list + full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench rc=$?"
for cfg in cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --steps 10 --warmup 3 --config $cfg --no-reference-gpu --no-whole-forward --no-cpu-baseline > gpurun_out/final_bench_$cfg.json 2> gpurun_out/final_bench_$cfg.err
  echo "bench $cfg rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --profile-step --warmup 1 > gpurun_out/final_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"build_volume|lookup_enc1" -s 6 -c 6 -f -o gpurun_out/final_full python bench.py --profile-step --warmup 1 > gpurun_out/final_ncu_full.log 2>&1
echo "ncu full rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms/step", round(d["ms_per_step"],3))
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "| build", d["build_roofline"]["avg_launch_us"], d["build_roofline"]["frac"])
print("coherent", {k:(round(v,3) if isinstance(v,float) else '') for k,v in d["build_roofline"].get("coherent",{}).items() if k!="what"})
print("lookup in plan", d["lookup_roofline"]["in_plan_kernel"]["avg_launch_us"], d["lookup_roofline"]["in_plan_kernel"]["frac"])
print("ref_gpu", {k:v for k,v in (d.get("reference_gpu") or {}).items() if k!="corr_kernels" and k!="what"})
print("whole", d.get("whole_forward"))
print("cpu", d.get("cpu_baseline",{}).get("value"))
for k,v in d["kernels"].items(): print(" ", k, round(v["ms_per_step"],3), round(v["avg_us"],1))
for cfg in ("cfg3","cfg4","cfg5"):
    try:
        e=json.loads(open(f"gpurun_out/final_bench_{cfg}.json").read().strip().splitlines()[-1])
        print(cfg, e["config"]["workload"][:60], "value", round(e["value"],2), "ms/step", round(e["ms_per_step"],3), "e2e", round(e["e2e"]["value"],2))
    except Exception as ex: print(cfg, "ERR", ex)
PY
