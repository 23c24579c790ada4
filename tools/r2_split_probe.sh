#!/bin/bash
PTB_LOG_SPLIT=1 python - <<'PY'
import os, sys, time, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from cpupathtrace_b200 import capi, pth, scenes
b200 = pth.load_b200()
verts, normals = scenes.standin_triangles(1000, 500, scenes.DEMO_DRAGON_TRANSFORM)
spec = scenes.cornell_demo(("triangles", verts, normals))
sc = spec.build(b200)
w, h, spp = 1920, 1080, 256
cam = scenes.demo_camera(b200, w, h)
b200.set_fast_queries(True, True, True)
b200.set_render_control(max_depth=0, relaxed_guard=True)
handle = sc.device_handle()
kw = scenes.demo_camera(None, w, h)
camera = capi.camera_init(kw["origin"], kw["look_at"], kw["up"], kw["focal_length"], kw["height"], kw["aspect_ratio"], kw["aperture_width"], kw["aperture_height"], kw["sampler"], 0.0, kw["focal_plane_dist"])
lib = capi.load()
flags = capi.PTB_FLAG_ANY_HIT_SHADOWS | capi.PTB_FLAG_SKIP_NULL_SHADOWS | capi.PTB_FLAG_CERTIFIED_CLOSEST | capi.PTB_FLAG_CERTIFIED_RELAXED
image = torch.zeros(h, w, 4, device="cuda")
host = np.zeros((h, w, 4), np.float32)
for rep in range(3):
    st = capi.RenderStats()
    o = capi.render_opts(w, h, spp, spp, 1e-3, 0, capi.PTB_RNG_COUNTER, flags | capi.PTB_FLAG_DEVICE_IO, 5 + rep, 0, 0, 1)
    t0 = time.perf_counter(); capi.check(lib.ptb_render(handle, C.byref(camera), C.byref(o), 0, 0, w, h, C.c_void_p(image.data_ptr()), C.byref(st))); dt = time.perf_counter() - t0
    print(f"== C-ABI device-resident: {dt*1e3:.1f} ms wall, device_ms_total {st.device_ms_total:.1f}", file=sys.stderr, flush=True)
for rep in range(3):
    st = capi.RenderStats()
    o = capi.render_opts(w, h, spp, spp, 1e-3, 0, capi.PTB_RNG_COUNTER, flags, 15 + rep, 0, 0, 1)
    t0 = time.perf_counter(); capi.check(lib.ptb_render(handle, C.byref(camera), C.byref(o), 0, 0, w, h, host.ctypes.data_as(C.c_void_p), C.byref(st))); dt = time.perf_counter() - t0
    print(f"== C-ABI host result: {dt*1e3:.1f} ms wall, device_ms_total {st.device_ms_total:.1f}", file=sys.stderr, flush=True)
for rep in range(3):
    t0 = time.perf_counter(); img, info = sc.process_job(cam, w, h, spp, spp, 1e-3, 0); dt = time.perf_counter() - t0
    print(f"== processJob: {dt*1e3:.1f} ms wall, call {info['seconds']*1e3:.1f} ms", file=sys.stderr, flush=True)
PY
