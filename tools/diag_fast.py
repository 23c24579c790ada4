"""Diagnostic: per-sample radiance of the production generator under two builds of ptb_fast.cu (same keys)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from cpupathtrace_b200 import capi, scenes, pth

out = sys.argv[1]
ref = pth.load_reference()
if len(sys.argv) > 2 and sys.argv[2] == "nomesh":
    spec = scenes.cornell_demo(None)
else:
    verts, normals = scenes.standin_triangles(400, 200, scenes.DEMO_DRAGON_TRANSFORM)
    spec = scenes.cornell_demo(("triangles", verts, normals))
prims, mats, lights = spec.to_pod(ref)
ctx = capi.Context(-1)
scene = capi.Scene(ctx, prims, mats, lights)
kw = scenes.demo_camera(None, 256, 144)
camera = capi.camera_init(kw["origin"], kw["look_at"], kw["up"], kw["focal_length"], kw["height"], kw["aspect_ratio"], kw["aperture_width"], kw["aperture_height"], kw["sampler"], 0.0, kw["focal_plane_dist"])
rng = np.random.Generator(np.random.PCG64(3))
n = 3_000_000
pixels = np.stack([rng.integers(0, 256, n), rng.integers(0, 144, n)], axis=1).astype(np.int32)
seeds = rng.integers(1, 2**63 - 1, n, dtype=np.int64).astype(np.uint64)
opts = capi.render_opts(256, 144, 1, 1, 1e-3, rng_mode=capi.PTB_RNG_COUNTER)
rgba, stats = scene.render_samples(camera, opts, pixels, seeds)
np.save(out, rgba)
print(out, "mean", rgba[:, :3].mean(), "clipped mean", np.minimum(rgba[:, :3], 4).mean(), "vertices/sample", stats.path_vertices / n, "shadow rays/sample", stats.shadow_rays / n)
