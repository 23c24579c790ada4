"""Probe: closest-hit throughput of incoherent rays inside the bench scene (Cornell + stand-in-1M), sorted vs unsorted."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from cpupathtrace_b200 import capi, scenes, pth

ref = pth.load_reference()
verts, normals = scenes.standin_triangles(1000, 500, scenes.DEMO_DRAGON_TRANSFORM)
spec = scenes.cornell_demo(("triangles", verts, normals))
prims, mats, lights = spec.to_pod(ref)
ctx = capi.Context(-1)
scene = capi.Scene(ctx, prims, mats, lights)
n = 1 << 24
g = torch.Generator(device="cuda"); g.manual_seed(1)
# origins near the mesh (where path vertices concentrate) and in the room
o = torch.rand((n, 3), generator=g, device="cuda") * 1.9 - 0.95
near = torch.rand(n, generator=g, device="cuda") < 0.6
centre = torch.tensor([0.4, -0.7, -0.75], device="cuda")
o[near] = centre + (torch.rand((int(near.sum()), 3), generator=g, device="cuda") - 0.5) * torch.tensor([0.8, 0.3, 0.8], device="cuda")
d = torch.randn((n, 3), generator=g, device="cuda"); d = d * (1.0 / torch.sqrt((d * d).sum(1, keepdim=True)))
rays = torch.cat([o, d], 1).contiguous()
t = torch.empty(n, device="cuda"); p = torch.empty(n, dtype=torch.int32, device="cuda")
flags = capi.PTB_FLAG_CERTIFIED_CLOSEST | capi.PTB_FLAG_CERTIFIED_RELAXED
for _ in range(3):
    st = scene.intersect_device(rays.data_ptr(), n, t.data_ptr(), p.data_ptr(), flags)
ms = [scene.intersect_device(rays.data_ptr(), n, t.data_ptr(), p.data_ptr(), flags) for _ in range(5)]
print("PTB_SORT_RAYS", os.environ.get("PTB_SORT_RAYS", "1"), "trace ms", np.median([m.device_ms_trace for m in ms]), "total ms", np.median([m.device_ms_total for m in ms]),
      "Mrays/s (trace only)", n / np.median([m.device_ms_trace for m in ms]) / 1e3, "hit", float((t >= 0).float().mean()))
