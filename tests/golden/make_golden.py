#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libpth_ref.so, built by `make -C oracle ref`
from /root/reference) — run in the dev container, where the reference tree is mounted:

    python tests/golden/make_golden.py

The fixtures pin the oracle restatement (oracle/pt_oracle.c) and the CUDA path on machines where neither
/root/reference nor oracle/_ref exists.  Each .npz stores inputs (POD scene, camera parameters, rays / pixels / seeds)
and the reference's outputs:
  hits_<scene>.npz     Scene::getIntersection: t, primitive index for 4096 rays (a third aimed at vertices/edges)
  samples_<scene>.npz  impl::getSample via processItem 1x1 @ 1 spp with RandomEngine(seed): RGBA for 2048 samples
  tile_<scene>.npz     processItem on a tile with one sequential engine: fixed spp and adaptive (min != max)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from conftest import random_rays  # noqa: E402

from cpupathtrace_b200 import pth, scenes  # noqa: E402

MIXED_CAMERA = dict(origin=(0.0, 0.0, -1.9), look_at=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), focal_length=0.7, height=1.0, aspect_ratio=-1.5,
                    aperture_width=0.04, aperture_height=0.03, sampler=2, hex_ratio=0.5, focal_plane_dist=2.0)
ADVANCED_CAMERA = dict(origin=(0.0, 0.0, 0.0), look_at=(0.0, 0.0, 1.0), up=(0.0, 1.0, 0.0), focal_length=0.2, height=0.5, aspect_ratio=1.94)


def cases():
    yield "cornell_mesh", scenes.cornell_demo(("obj", scenes.standin_obj(40, 30))), scenes.demo_camera(None, 64, 64), 64, 64
    yield "mixed", scenes.mixed_materials(n_tris=240), MIXED_CAMERA, 96, 64
    yield "advanced", scenes.advanced_render(), ADVANCED_CAMERA, 132, 68


def camera_array(kw):
    keys = ["focal_length", "height", "aspect_ratio", "aperture_width", "aperture_height", "sampler", "hex_ratio", "focal_plane_dist"]
    return np.array(list(kw["origin"]) + list(kw["look_at"]) + list(kw["up"]) + [float(kw.get(k, 0.0)) for k in keys], np.float64)


def main():
    ref = pth.load_reference()
    for name, spec, cam_kw, width, height in cases():
        prims, materials, lights = spec.to_pod(ref)
        scene = spec.build(ref)
        camera = ref.camera(**cam_kw)
        common = dict(prims=prims, materials=materials, lights=lights, camera=camera_array(cam_kw), size=np.array([width, height]))

        tri = prims[prims["kind"] == 0]["p"]
        aim = np.concatenate([tri[:, :9].reshape(-1, 3), 0.5 * (tri[:, 0:3] + tri[:, 3:6])])
        aim = aim[np.random.Generator(np.random.PCG64(5)).permutation(len(aim))][:1400]
        rays = random_rays(4096, seed=17, box=1.2, aim=aim)
        t, ids = scene.intersect(rays)
        np.savez_compressed(os.path.join(HERE, f"hits_{name}.npz"), rays=rays, t=t, prim=ids, **common)

        rng = np.random.Generator(np.random.PCG64(29))
        n = 2048
        pixels = np.stack([rng.integers(0, width, n), rng.integers(0, height, n)], axis=1).astype(np.int32)
        seeds = rng.integers(1, 2**63 - 1, n, dtype=np.int64).astype(np.uint64)
        rgba = scene.render_samples(camera, width, height, 1e-3, pixels, seeds)
        np.savez_compressed(os.path.join(HERE, f"samples_{name}.npz"), pixels=pixels, seeds=seeds, rgba=rgba, epsilon=np.float32(1e-3), **common)

        rect = (5, 7, 12, 9)
        fixed = scene.process_item(camera, width, height, 16, 16, 1e-3, rect, 4242)
        adaptive = scene.process_item(camera, width, height, 5, 10, 1e-3, rect, 4243)
        np.savez_compressed(os.path.join(HERE, f"tile_{name}.npz"), rect=np.array(rect), fixed=fixed, adaptive=adaptive, fixed_spp=np.array([16, 16]),
                            adaptive_spp=np.array([5, 10]), seeds=np.array([4242, 4243], np.uint64), epsilon=np.float32(1e-3), **common)
        print(name, "prims", len(prims), "hit rate", float((t >= 0).mean()), "mean radiance", float(rgba[:, :3].mean()))


if __name__ == "__main__":
    main()
