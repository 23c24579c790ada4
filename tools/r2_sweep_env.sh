#!/bin/bash
# usage: r2_sweep_env.sh VAR v1 v2 ...   -- quick C2 bench per value of an environment variable
VAR=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $VAR=$v timeout -s KILL 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-adaptive-line > gpurun_out/sweep_$v.json 2> gpurun_out/sweep.err || tail -5 gpurun_out/sweep.err
  python - <<PY
import json
d = json.load(open("gpurun_out/sweep_$v.json")); b = d["breakdown"]
print("$VAR=$v", round(d["value"],1), "Msamples/s e2e", round(d["e2e"]["value"],1), "frame", round(b["frame_ms"],1), "closest", round(b["closest_trace_ms"],1), "shadow", round(b["shadow_trace_ms"],1), "shade+acc", round(b["generate_shade_accumulate_resolve_ms"],1))
PY
done
