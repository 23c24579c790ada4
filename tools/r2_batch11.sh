#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
PTB_LOG_BUILD=1 timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_quick.json 2> gpurun_out/bench_c2_quick.err || tail -5 gpurun_out/bench_c2_quick.err
grep "scene setup" gpurun_out/bench_c2_quick.err | tail -9
python -c "
import json; d=json.load(open('gpurun_out/bench_c2_quick.json')); print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['scene'], d['adaptive'])"
PTB_LOG_BUILD=1 timeout -s KILL 900 python bench.py --workload soup --tris 16 --rays incoherent --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/soup_16_incoherent.json 2> gpurun_out/soup.err || tail -5 gpurun_out/soup.err
grep "scene setup" gpurun_out/soup.err | tail -9
python -c "
import json; d=json.load(open('gpurun_out/soup_16_incoherent.json')); print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['frac'], d['scene'])"
