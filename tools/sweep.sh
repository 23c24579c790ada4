#!/bin/bash
# usage: tools/sweep.sh "ENV1=a ENV2=b" "ENV1=c" ...   — one 32 spp bench per environment setting
mkdir -p gpurun_out
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --spp 32 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/sweep.json 2> gpurun_out/sweep.err || tail -3 gpurun_out/sweep.err
  python - "$cfg" <<'PY'
import json, sys
d = json.load(open("gpurun_out/sweep.json")); r = d["roofline"]
print(f"{sys.argv[1]:50s} Msamples/s {d['value']:7.1f} trace-only Mrays/s {r['mrays_per_s_trace_only']:6.0f} trace ms {r['trace_ms_per_step']:6.1f} shade ms {r['shade_ms_per_step']:5.1f} | closest {r['closest_mrays_per_s']:5.0f} Mr/s ({r['closest_inner_per_ray']:.1f} nodes) shadow {r['shadow_mrays_per_s']:5.0f} Mr/s ({r['shadow_inner_per_ray']:.1f} nodes) shadow ms {r['shadow_trace_ms_per_step']:.1f}")
PY
done
