#!/usr/bin/env python
"""Experiment: two halves of the frame rendered concurrently on ONE GPU by two contexts (two streams, two workspaces)
through ptb_render_multi, with the kernels' grids reduced so that the halves' kernels can be co-resident on an SM
(PTB_TRACE_BLOCKS_PER_SM / PTB_SHADE_BLOCKS_PER_SM).  argv: n_streams [spp]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cpupathtrace_b200 import capi, pth, scenes

n_streams = int(sys.argv[1])
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 256
w, h = 1920, 1080
b200 = pth.load_b200()
verts, normals = scenes.standin_triangles(1000, 500, scenes.DEMO_DRAGON_TRANSFORM)
spec = scenes.cornell_demo(("triangles", verts, normals))
sc = spec.build(b200)
handle = sc.device_handle()
kw = scenes.demo_camera(None, w, h)
camera = capi.camera_init(kw["origin"], kw["look_at"], kw["up"], kw["focal_length"], kw["height"], kw["aspect_ratio"], kw["aperture_width"], kw["aperture_height"], kw["sampler"], 0.0,
                          kw["focal_plane_dist"])
lib = capi.load()
flags = capi.PTB_FLAG_ANY_HIT_SHADOWS | capi.PTB_FLAG_SKIP_NULL_SHADOWS | capi.PTB_FLAG_CERTIFIED_CLOSEST | capi.PTB_FLAG_CERTIFIED_RELAXED | capi.PTB_FLAG_DEVICE_IO
image = torch.zeros(h, w, 4, device="cuda")
handles = [C.c_void_p(handle)]
contexts = []
for k in range(1, n_streams):
    ctx = capi.Context(0)
    contexts.append(ctx)
    clone = C.c_void_p()
    capi.check(lib.ptb_scene_clone(C.c_void_p(handle), ctx.handle, C.byref(clone)))
    handles.append(clone)
array = (C.c_void_p * n_streams)(*handles)
stats = (capi.RenderStats * n_streams)()
times = []
for rep in range(4):
    o = capi.render_opts(w, h, spp, spp, 1e-3, 0, capi.PTB_RNG_COUNTER, flags, 5 + rep, 0, 0, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    capi.check(lib.ptb_render_multi(array, n_streams, C.byref(camera), C.byref(o), 0, 0, w, h, C.c_void_p(image.data_ptr()), stats, None, None))
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
best = min(times[1:])
print(f"streams {n_streams} trace/SM {os.environ.get('PTB_TRACE_BLOCKS_PER_SM', '16')} shade/SM {os.environ.get('PTB_SHADE_BLOCKS_PER_SM', '6')} pool {os.environ.get('PTB_POOL_PATHS', 'auto')}: "
      f"{best * 1e3:7.1f} ms = {w * h * spp / best / 1e6:6.1f} Msamples/s  (all: {[round(t * 1e3) for t in times]}) mean {float(image[..., :3].mean()):.5f}", flush=True)
