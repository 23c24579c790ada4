#!/bin/bash
mkdir -p gpurun_out
for div in 2 3 4 6 8 16; do
PTB_ADAPTIVE_ROUND_DIVISOR=$div python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import ctypes as C
import numpy as np, torch
from cpupathtrace_b200 import capi, pth, scenes
b200 = pth.load_b200()
verts, normals = scenes.standin_triangles(1000, 500, scenes.DEMO_DRAGON_TRANSFORM)
spec = scenes.cornell_demo(("triangles", verts, normals))
sc = spec.build(b200)
handle = sc.device_handle()
kw = scenes.demo_camera(None, 1920, 1080)
camera = capi.camera_init(kw["origin"], kw["look_at"], kw["up"], kw["focal_length"], kw["height"], kw["aspect_ratio"], kw["aperture_width"], kw["aperture_height"], kw["sampler"], 0.0, kw["focal_plane_dist"])
lib = capi.load()
flags = capi.PTB_FLAG_ANY_HIT_SHADOWS | capi.PTB_FLAG_SKIP_NULL_SHADOWS | capi.PTB_FLAG_CERTIFIED_CLOSEST | capi.PTB_FLAG_CERTIFIED_RELAXED | capi.PTB_FLAG_DEVICE_IO
image = torch.zeros(1080, 1920, 4, device="cuda")
for lo, hi in ((32, 256), (16, 1024)):
    for rep in range(2):
        st = capi.RenderStats()
        o = capi.render_opts(1920, 1080, lo, hi, 1e-3, 0, capi.PTB_RNG_COUNTER, flags, 5, 0, 0, 1)
        capi.check(lib.ptb_render(handle, C.byref(camera), C.byref(o), 0, 0, 1920, 1080, C.c_void_p(image.data_ptr()), C.byref(st)))
    print(f"divisor {os.environ['PTB_ADAPTIVE_ROUND_DIVISOR']:>2s}  {lo}..{hi}: traced {st.samples/1e6:8.1f} M consumed {st.samples_used/1e6:8.1f} M of {1920*1080*hi/1e6:.0f} M, rounds {st.adaptive_rounds}, {st.device_ms_total:8.1f} ms = {st.samples_used/st.device_ms_total/1e3:.1f} M useful samples/s", flush=True)
PY
done
