#!/bin/bash
# headline bench + ncu evidence (launch list of one step, full captures of the build kernels and one encoder conv)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --profile-step --warmup 1 > gpurun_out/r2k_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"build_volume" -s 4 -c 4 -f -o gpurun_out/r2k_build python bench.py --profile-step --warmup 1 > gpurun_out/r2k_ncu_build.log 2>&1
echo "ncu build rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2k_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms/step", round(d["ms_per_step"],3))
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "| build", d["build_roofline"]["avg_launch_us"], d["build_roofline"]["frac"])
print("ref_gpu", {k:v for k,v in (d.get("reference_gpu") or {}).items() if k!="corr_kernels" and k!="what"})
print("whole", d.get("whole_forward"))
print("cpu", d.get("cpu_baseline",{}).get("value"))
PY
