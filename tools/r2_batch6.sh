#!/bin/bash
# round 2 measurement set: progress test, default bench (C2), launch list, full capture of one steady-state bounce iteration
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "progress or several" 2>&1 | tail -5
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err || tail -5 gpurun_out/bench_c2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_c2.json")); r = d["roofline"]
print("C2:", round(d["value"],1), "Msamples/s e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "frac", round(r["frac"],3), "breakdown", {k: (round(v,1) if isinstance(v,float) else v) for k,v in d["breakdown"].items() if k!="source"}, "setup", d["scene"]["scene_ctor_s"])
print("adaptive:", d.get("adaptive"))
PY
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --no-cpu-baseline --no-adaptive-line > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"traceClosestKernelILi2ELb0|traceShadowKernelILb0|shadeKernel|accumulateKernel" -s 28 -c 4 -f -o gpurun_out/r02_iteration python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-adaptive-line > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/*.ncu-rep
