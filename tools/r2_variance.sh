#!/bin/bash
mkdir -p gpurun_out
run() {
  env "$@" timeout -s KILL 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-adaptive-line > gpurun_out/var.json 2> gpurun_out/var.err || tail -3 gpurun_out/var.err
  python - "$*" <<'PY'
import json, sys
d = json.load(open("gpurun_out/var.json"))
print(f"{sys.argv[1]:60s} value {d['value']:6.1f} steps {d['ms_steps']} e2e {d['e2e']['ms_steps']} profile frame {d['breakdown']['frame_ms']:.1f}")
PY
}
run PTB_X=1
run PTB_BENCH_NO_CLOCK_SAMPLER=1
run PTB_PIPELINED_BATCHES=0
run PTB_PIPELINED_BATCHES=0 PTB_BENCH_NO_CLOCK_SAMPLER=1
run PTB_ITERATIONS_PER_SYNC=8
