#!/bin/bash
mkdir -p gpurun_out
run() {
  env "$@" timeout -s KILL 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-adaptive-line > gpurun_out/var.json 2> gpurun_out/var.err || tail -3 gpurun_out/var.err
  python - "$*" <<'PY'
import json, sys
d = json.load(open("gpurun_out/var.json")); b = d["breakdown"]
print(f"{sys.argv[1]:70s} best step {min(d['ms_steps']):6.1f} closest {b['closest_trace_ms']:6.1f} shadow {b['shadow_trace_ms']:6.1f} shade+acc {b['generate_shade_accumulate_resolve_ms']:6.1f} frame {b['frame_ms']:6.1f}")
PY
}
run PTB_X=base
run PTB_REFILL_VOTE=8
run PTB_REFILL_VOTE=16
run PTB_LEAF_VOTE=8
run PTB_LEAF_VOTE=16
run PTB_INNER_BURST=3
run PTB_INNER_BURST=6
run PTB_LEAF_BURST=3
run PTB_SHADOW_REFILL_VOTE=12 PTB_SHADOW_LEAF_VOTE=8
run PTB_SHADOW_REFILL_VOTE=20 PTB_SHADOW_LEAF_VOTE=16
run PTB_POOL_PATHS=201326592
