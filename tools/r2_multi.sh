#!/bin/bash
# usage: r2_multi.sh N [config]   -- bench.py on N GPUs of one box (C4 by default), JSON into gpurun_out/
N=$1; CFG=${2:-auto}
mkdir -p gpurun_out
PTB_LOG_MULTI=1 timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $CFG --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu_$CFG.json 2> gpurun_out/bench_${N}gpu_$CFG.err || tail -20 gpurun_out/bench_${N}gpu_$CFG.err
grep "render_multi" gpurun_out/bench_${N}gpu_$CFG.err | tail -40
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${N}gpu_$CFG.json"))
print("N=$N", d["config"]["workload"][:40], round(d["value"],1), "Msamples/s e2e", round(d["e2e"]["value"],1), d["e2e"].get("ms_steps"), "ms/step", round(d["ms_per_step"],1), "per_rank", d["per_rank"])
PY
