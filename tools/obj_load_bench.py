#!/usr/bin/env python
"""io::loadMesh (reference src/scene/mesh.cpp) on the stand-in mesh written as Wavefront OBJ text: the reference's parser
(oracle/_ref, unmodified sources) against this repository's host/mesh.cpp, same text, same transform, smooth normals.
CPU only.  Prints one JSON line; the triangles of both loaders are compared bit for bit."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpupathtrace_b200 import pth, scenes  # noqa: E402


def main():
    nu, nv = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1000x500").split("x"))
    text = scenes.standin_obj(nu, nv).encode()
    out = {"mesh": f"stand-in {nu}x{nv}", "triangles": 2 * nu * nv, "obj_megabytes": round(len(text) / 1e6, 1)}
    tris = {}
    # "reference": the timing build (-O3 -march=x86-64-v3, the one bench.py's CPU arm uses); "reference_parity": the build
    # without FMA contraction that value parity is defined against (oracle/Makefile)
    libs = (("reference", pth.load_reference(fast=True) if os.path.exists(pth.REF_FAST) else None),
            ("reference_parity", pth.load_reference() if os.path.exists(pth.REF_PARITY) else None), ("b200", pth.load_b200()))
    for name, lib in libs:
        if lib is None:
            continue
        best = None
        for _ in range(3):
            b = lib.builder()
            t0 = time.perf_counter()
            b.mesh_obj(text, scenes.DEMO_DRAGON_TRANSFORM, cull=False, smooth=True, material=-1)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            tris[name] = b.get_triangles()
            b.close()
        out[f"{name}_load_s"] = round(best, 3)
        out[f"{name}_s_per_million_triangles"] = round(best / (2 * nu * nv) * 1e6, 3)
    if "reference_parity" in tris:
        out["identical_to_reference_parity_build"] = bool(np.array_equal(tris["reference_parity"].view(np.uint32), tris["b200"].view(np.uint32)))
    if "reference" in tris:
        out["speedup"] = round(out["reference_load_s"] / out["b200_load_s"], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
