#!/bin/bash
PTB_LOG_BATCHES=1 python - 2> gpurun_out/fb.err <<'PY'
import os, sys, time, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from cpupathtrace_b200 import capi, pth, scenes
b200 = pth.load_b200()
verts, normals = scenes.standin_triangles(1000, 500, scenes.DEMO_DRAGON_TRANSFORM)
sc = scenes.cornell_demo(("triangles", verts, normals)).build(b200)
w, h, spp = 1920, 1080, 256
handle = sc.device_handle()
kw = scenes.demo_camera(None, w, h)
camera = capi.camera_init(kw["origin"], kw["look_at"], kw["up"], kw["focal_length"], kw["height"], kw["aspect_ratio"], kw["aperture_width"], kw["aperture_height"], kw["sampler"], 0.0, kw["focal_plane_dist"])
lib = capi.load()
flags = capi.PTB_FLAG_ANY_HIT_SHADOWS | capi.PTB_FLAG_SKIP_NULL_SHADOWS | capi.PTB_FLAG_CERTIFIED_CLOSEST | capi.PTB_FLAG_CERTIFIED_RELAXED | capi.PTB_FLAG_DEVICE_IO
image = torch.zeros(h, w, 4, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def frame(seed):
    st = capi.RenderStats()
    o = capi.render_opts(w, h, spp, spp, 1e-3, 0, capi.PTB_RNG_COUNTER, flags, seed, 0, 0, 1)
    capi.check(lib.ptb_render(handle, C.byref(camera), C.byref(o), 0, 0, w, h, C.c_void_p(image.data_ptr()), C.byref(st)))
    return st.device_ms_total
for mode in ("warm", "back-to-back", "flush between", "sleep 0.2 s between", "same seed"):
    ts = []
    for i in range(8):
        if mode == "flush between":
            flush.zero_(); torch.cuda.synchronize()
        if mode == "sleep 0.2 s between":
            time.sleep(0.2)
        ts.append(round(frame(7 if mode == "same seed" else 100 + i), 1))
    print(mode, ts, flush=True)
PY
python - <<'PY'
import re
firsts, cur_prev = [], None
for line in open("gpurun_out/fb.err"):
    m = re.search(r"done ([0-9.]+) ms after the call started \((\d+) paths", line)
    if m and int(m.group(2)) > 100_000_000 and (cur_prev is None or float(m.group(1)) < cur_prev):
        firsts.append(round(float(m.group(1))))
    if m:
        cur_prev = float(m.group(1))
print("first-batch times:", firsts)
PY
