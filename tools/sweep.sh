#!/bin/bash
# usage: tools/sweep.sh "ENV1=a ENV2=b [-- bench args]" "ENV1=c" ...   — one 32 spp bench per setting, each under a hard timeout
mkdir -p gpurun_out
for cfg in "$@"; do
  envs="${cfg%% -- *}"; [[ "$cfg" == " -- "* ]] && envs=""
  args=""
  [[ "$cfg" == *" -- "* ]] && args="${cfg#* -- }"
  env $envs timeout -s KILL ${T:-90} python bench.py --spp ${SPP:-32} --steps 2 --warmup 1 --no-cpu-baseline $args > gpurun_out/sweep.json 2> gpurun_out/sweep.err || { echo "$cfg: FAILED/timeout"; tail -2 gpurun_out/sweep.err; continue; }
  python - "$cfg" <<'PY'
import json, sys
d = json.load(open("gpurun_out/sweep.json")); r = d["roofline"]
print(f"{sys.argv[1]:50s} Msamples/s {d['value']:7.1f} trace-only Mrays/s {r['mrays_per_s_trace_only']:6.0f} trace ms {r['trace_ms_per_step']:6.1f} shade ms {r['shade_ms_per_step']:5.1f} | closest {r['closest_mrays_per_s']:5.0f} Mr/s ({r['closest_inner_per_ray']:.1f}+{r['closest_leaf_per_ray']:.1f}) shadow {r['shadow_mrays_per_s']:5.0f} Mr/s ({r['shadow_inner_per_ray']:.1f}+{r['shadow_leaf_per_ray']:.1f}) shadow ms {r['shadow_trace_ms_per_step']:.1f}")
PY
done
