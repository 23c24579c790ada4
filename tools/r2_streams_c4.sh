#!/bin/bash
mkdir -p gpurun_out
for cfg in "2 c4" "1 c4" "2 c2" "1 c2"; do
  set -- $cfg
  PTB_STREAMS=$1 timeout -s KILL 600 python bench.py --config $2 --steps 2 --warmup 2 --no-cpu-baseline --no-adaptive-line > gpurun_out/sweep_s.json 2> gpurun_out/sweep.err || tail -5 gpurun_out/sweep.err
  python - <<PY
import json
d = json.load(open("gpurun_out/sweep_s.json"))
print("streams $1 $2:", round(d["value"],1), "Msamples/s e2e", round(d["e2e"]["value"],1), d["e2e"]["ms_steps"], "ms/step", round(d["ms_per_step"],1), "iterations", d["bounce_iterations_per_step"])
PY
done
PTB_STREAMS=2 PTB_POOL_PATHS=67108864 timeout -s KILL 600 python bench.py --config c4 --steps 2 --warmup 2 --no-cpu-baseline --no-adaptive-line > gpurun_out/sweep_s.json 2> gpurun_out/sweep.err
python -c "
import json; d=json.load(open('gpurun_out/sweep_s.json')); print('streams 2 c4 pool 64Mi:', round(d['value'],1), round(d['ms_per_step'],1))"
