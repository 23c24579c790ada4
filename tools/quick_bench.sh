#!/bin/bash
# quick GPU check used during kernel iteration: parity tests, then a 16 spp bench (no CPU baseline)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --spp ${1:-32} --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
r = d["roofline"]
print(f"Msamples/s {d['value']:.1f}  Mrays/s {d['mrays_per_s']:.0f}  trace-only Mrays/s {r['mrays_per_s_trace_only']:.0f}  trace ms {r['trace_ms_per_step']:.1f}  shade ms {r['shade_ms_per_step']:.1f}  step ms {d['ms_per_step']:.1f}  frac {r['frac']:.3f}")
PY
