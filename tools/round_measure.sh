#!/bin/bash
# End-of-milestone measurement set (run on the GPU box): default bench, reference arm, ncu launch list of the default
# command, one ncu --set full capture of one steady-state bounce iteration.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout -s KILL 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_default.csv python bench.py --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"traceClosestKernelILi2ELb0|traceShadowKernelILb0|shadeKernelINS_4RngTILb0|accumulateKernelINS_4RngTILb0" -s 12 -c 4 -f -o gpurun_out/prof_iteration_r1 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
python - <<'PY'
import json
for f in ("bench_default", "bench_reference"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 2), d["unit"], "e2e", round(d["e2e"]["value"], 2), "frac", d.get("roofline", {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
