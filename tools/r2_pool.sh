#!/bin/bash
mkdir -p gpurun_out
run() {
  env "$@" timeout -s KILL 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-adaptive-line > gpurun_out/var.json 2> gpurun_out/var.err || tail -3 gpurun_out/var.err
  python - "$*" <<'PY'
import json, sys
d = json.load(open("gpurun_out/var.json")); b = d["breakdown"]
print(f"{sys.argv[1]:40s} value {d['value']:6.1f} steps {d['ms_steps']} e2e {d['e2e']['ms_steps']} closest {b['closest_trace_ms']:6.1f} shadow {b['shadow_trace_ms']:6.1f} shade+acc {b['generate_shade_accumulate_resolve_ms']:6.1f} frame {b['frame_ms']:6.1f} iterations {d['bounce_iterations_per_step']}")
PY
}
run PTB_POOL_PATHS=167772160
run PTB_POOL_PATHS=268435456
run PTB_POOL_PATHS=402653184
