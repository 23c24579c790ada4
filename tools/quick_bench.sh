#!/bin/bash
# quick GPU check used during kernel iteration: parity tests, then a short bench (no CPU baseline); everything under hard timeouts
# usage: tools/quick_bench.sh [spp] [extra bench.py args...]
mkdir -p gpurun_out
spp=${1:-32}
shift
timeout -s KILL 240 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout -s KILL 120 python bench.py --spp $spp --steps 2 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
r = d["roofline"]
print(f"Msamples/s {d['value']:.1f}  e2e {d['e2e']['value']:.1f}  Mrays/s {d['mrays_per_s']:.0f}  trace-only Mrays/s {r['mrays_per_s_trace_only']:.0f}  closest {r['closest_mrays_per_s']:.0f}  shadow {r['shadow_mrays_per_s']:.0f}  trace ms {r['trace_ms_per_step']:.1f}  shade ms {r['shade_ms_per_step']:.1f}  step ms {d['ms_per_step']:.1f}  frac {r['frac']:.3f}")
print({k: round(v, 2) for k, v in r.items() if k.endswith('_per_ray')}, d.get('closest_hit'))
PY
