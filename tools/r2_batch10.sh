#!/bin/bash
mkdir -p gpurun_out
PTB_LOG_BUILD=1 python - <<'PY'
import time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from cpupathtrace_b200 import capi, pth, scenes
b200 = pth.load_b200()
verts, normals = scenes.standin_triangles(1000, 500, scenes.DEMO_DRAGON_TRANSFORM)
spec = scenes.cornell_demo(("triangles", verts, normals))
for k in range(3):
    t0 = time.perf_counter(); b = spec.replay(b200); t1 = time.perf_counter(); sc = b.scene(); t2 = time.perf_counter(); b.close()
    print(f"round {k}: objects {t1 - t0:.3f} s, Scene::Scene {t2 - t1:.3f} s", flush=True)
    del sc
PY
timeout -s KILL 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-adaptive-line > gpurun_out/bench_c2_quick.json 2> gpurun_out/bench_c2_quick.err || tail -5 gpurun_out/bench_c2_quick.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c2_quick.json')); print(round(d['value'],1), d['scene'])"
