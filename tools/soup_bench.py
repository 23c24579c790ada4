#!/usr/bin/env python
"""BASELINE.json configs[2]: synthetic random-triangle soup, 1M-16M triangles, primary + shadow ray intersection
microbench of the traversal kernels alone (ptb_intersect / ptb_occluded with device-resident rays).

For every soup size N (centres uniform in [-1,1]^3, vertices = centre + uniform[-s,s]^3, s = 0.5 N^(-1/3),
no culling) three ray sets of 2^24 rays:
  coherent    pinhole at (0,0,-3) through a 4096 x 4096 grid over [-1,1]^2 at z = -1
  incoherent  uniform origins in [-1,1]^3, uniform directions
  shadow      from every coherent hit towards a point light at (0,0.99,0), built as worker.cpp:80-86 (any-hit, limit = dist - eps)
Prints one JSON object per (N, ray set) with Mrays/s, node / primitive fetches per ray (counting pass), algorithmic
bytes per ray and the fraction of the measured HBM bandwidth they amount to.  At 4M triangles and above the node and
geometry arrays (128 B per triangle) exceed the 126 MB L2, so this is the regime where the HBM roofline applies.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from cpupathtrace_b200 import capi, scenes

    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1,4,16", help="soup sizes in Mi triangles")
    ap.add_argument("--rays", type=int, default=1 << 24)
    ap.add_argument("--repeats", type=int, default=5)
    args = ap.parse_args()

    peak = 6650.0
    peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(peaks):
        peak = float(json.load(open(peaks))["hbm_gbs"])

    ctx = capi.Context(-1)
    dev = torch.device("cuda", ctx.device())
    n = args.rays
    side = int(round(n ** 0.5))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    # ray sets on the device (float32, normalised like rt_vector::normalize)
    g = torch.Generator(device=dev)
    g.manual_seed(0x7A750001)
    u = (torch.arange(side, device=dev, dtype=torch.float32) + 0.5) / side * 2 - 1
    tx, ty = torch.meshgrid(u, u, indexing="xy")
    target = torch.stack([tx.reshape(-1), ty.reshape(-1), torch.full((side * side,), -1.0, device=dev)], dim=1)
    origin = torch.tensor([0.0, 0.0, -3.0], device=dev).expand_as(target)
    d = target - origin
    d = d * (1.0 / torch.sqrt((d * d).sum(dim=1, keepdim=True)))
    coherent = torch.cat([origin, d], dim=1).contiguous()
    o = torch.rand((n, 3), generator=g, device=dev) * 2 - 1
    d = torch.randn((n, 3), generator=g, device=dev)
    d = d * (1.0 / torch.sqrt((d * d).sum(dim=1, keepdim=True)))
    incoherent = torch.cat([o, d], dim=1).contiguous()

    for mi in [int(v) for v in args.sizes.split(",")]:
        n_tris = mi << 20
        t0 = time.perf_counter()
        verts = scenes.soup_triangles(n_tris, 0x5EED0000 + int(np.log2(n_tris)))
        prims = np.zeros(n_tris, capi.PRIM_DTYPE)
        prims["kind"] = capi.PTB_PRIM_TRIANGLE
        prims["p"][:, :9] = verts
        a, b, c = verts[:, 0:3], verts[:, 3:6], verts[:, 6:9]
        nrm = np.cross(b - a, c - a)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        prims["p"][:, 9:12] = prims["p"][:, 12:15] = prims["p"][:, 15:18] = nrm
        mats = np.zeros(1, capi.MATERIAL_DTYPE)
        mats[0] = ((1, 1, 1, 1), (0, 0, 0, 0), 1.0, 0, 0, 0)
        t_gen = time.perf_counter() - t0
        scene = capi.Scene(ctx, prims, mats)
        info = scene.info()
        del prims, verts

        t_out = torch.empty(n, dtype=torch.float32, device=dev)
        prim_out = torch.empty(n, dtype=torch.int32, device=dev)
        occ_out = torch.empty(n, dtype=torch.uint8, device=dev)

        def run_closest(rays, flags=0):
            return scene.intersect_device(rays.data_ptr(), len(rays), t_out.data_ptr(), prim_out.data_ptr(), flags)

        # shadow set from the coherent hits
        run_closest(coherent)
        torch.cuda.synchronize()
        hit = t_out >= 0
        pos = coherent[:, :3] + coherent[:, 3:] * t_out[:, None]
        to_light = torch.tensor([0.0, 0.99, 0.0], device=dev) - pos
        dist = torch.sqrt((to_light * to_light).sum(dim=1))
        ldir = to_light * (1.0 / dist)[:, None]
        eps = 1e-3
        shadow = torch.cat([pos + ldir * eps, ldir, (dist - eps)[:, None]], dim=1)[hit].contiguous()

        cases = []
        for name, rays in (("coherent", coherent), ("incoherent", incoherent)):
            cases.append((name, rays, False, 0, "reference-topology tree"))
            cases.append((name, rays, False, capi.PTB_FLAG_CERTIFIED_CLOSEST, "certified SAH walk + re-trace"))
        cases.append(("shadow", shadow, True, 0, "any-hit on the SAH hierarchy"))
        reference_result = {}
        for name, rays, any_hit, mode_flags, mode in cases:
            def run(flags=0):
                if any_hit:
                    return scene.occluded_device(rays.data_ptr(), len(rays), occ_out.data_ptr(), flags)
                return run_closest(rays, flags | mode_flags)

            counted = run(capi.PTB_FLAG_COUNT_VISITS)
            inner = counted.inner_visits / len(rays)
            leaf = counted.leaf_visits / len(rays)
            for _ in range(3):
                run()
            ms = []
            for _ in range(args.repeats):
                flush.zero_()
                torch.cuda.synchronize()
                ms.append(run().device_ms_trace)
            ms = float(np.median(ms))
            bytes_per_ray = 64 * inner + 48 * leaf + (28 + 1 if any_hit else 24 + 8)
            achieved = bytes_per_ray * len(rays) / (ms / 1e3) / 1e9
            result = (t_out >= 0).float().mean().item() if not any_hit else occ_out[: len(rays)].float().mean().item()
            identical = None
            if not any_hit:
                # size-independent parity property: both closest-hit modes return the same primitive and distance for every ray
                if mode_flags == 0:
                    reference_result[name] = (t_out.clone(), prim_out.clone())
                else:
                    t_ref, prim_ref = reference_result[name]
                    hit_mask = t_ref >= 0
                    identical = bool(torch.equal(prim_ref, prim_out) and torch.equal(t_ref[hit_mask], t_out[hit_mask]) and bool((t_out[~hit_mask] < 0).all()))
            print(json.dumps({
                "config": f"soup-{mi}Mi", "rays": name, "closest_hit": mode, "identical_to_reference_tree": identical,
                "retraced": int(counted.closest_rays_retraced), "n_rays": len(rays), "mrays_per_s": len(rays) / ms / 1e3, "ms": ms,
                "inner_fetches_per_ray": inner, "leaf_fetches_per_ray": leaf, "bytes_per_ray": bytes_per_ray,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak},
                "hit_or_occluded_fraction": result, "bvh_depth": info.bvh_depth, "scene_mb": info.device_bytes / 2**20,
                "bvh_build_s": info.build_seconds, "upload_s": info.upload_seconds, "soup_gen_s": t_gen,
            }), flush=True)
        scene.close()
        del t_out, prim_out, occ_out


if __name__ == "__main__":
    main()
