#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from cpupathtrace_b200 import capi, pth, scenes
b200 = pth.load_b200()
verts, normals = scenes.standin_triangles(1000, 500, scenes.DEMO_DRAGON_TRANSFORM)
spec = scenes.cornell_demo(("triangles", verts, normals))
sc = spec.build(b200)
cam = scenes.demo_camera(b200, 1920, 1080)
b200.set_fast_queries(True, True, True)
b200.set_render_control(max_depth=16, relaxed_guard=True)
for devices in (1, 2):
    got = b200.set_devices(devices)
    for spp in (8, 256, 256, 256, 1024, 1024):
        t0 = time.perf_counter()
        img, info = sc.process_job(cam, 1920, 1080, spp, spp, 1e-3, 0)
        dt = time.perf_counter() - t0
        print(f"devices {got} spp {spp:5d}: {dt*1e3:8.1f} ms = {1920*1080*spp/dt/1e6:7.1f} Msamples/s", flush=True)
PY
