"""Inputs shared by the unit-entry parity tests (Object::getSurfaceNormal / sampleSurface, BSDF::propagateRay / getSpectrum):
one smooth-shaded triangle with backface culling, one flat triangle, one sphere, one NullObject-like default, and the
four BSDF configurations of the reference (Lambertian, Glass, Mirror, one-way Mirror)."""
import numpy as np

from cpupathtrace_b200 import capi

N = 4000


def _unit(v):
    v = np.asarray(v, np.float32)
    l2 = (v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2]
    return (v * (np.float32(1.0) / np.sqrt(l2))[..., None]).astype(np.float32)


def prims():
    out = np.zeros(3, capi.PRIM_DTYPE)
    a, b, c = (0.1, -0.2, 0.3), (0.9, 0.1, 0.2), (0.3, 0.8, -0.4)
    na, nb, nc = _unit([[0.1, 0.2, 1.0], [-0.3, 0.1, 0.9], [0.2, -0.2, 0.8]])
    out[0]["kind"], out[0]["cull_backface"] = capi.PTB_PRIM_TRIANGLE, 1
    out[0]["p"][:] = np.concatenate([a, b, c, na, nb, nc]).astype(np.float32)
    a, b, c = np.float32([-20, -1, -20]), np.float32([20, -1, -20]), np.float32([20, -1, 20])
    n = _unit(np.cross(b - a, c - a)[None])[0]
    out[1]["kind"] = capi.PTB_PRIM_TRIANGLE
    out[1]["p"][:] = np.concatenate([a, b, c, n, n, n]).astype(np.float32)
    out[2]["kind"] = capi.PTB_PRIM_SPHERE
    out[2]["p"][:4] = (0.5, -0.5, 0.5, 0.5)
    return out


def add_to_builder(builder, prim_array):
    """The same primitives through the public C++ API (harness builder); returns their object indices."""
    index = []
    for prim in prim_array:
        index.append(builder.object_count())
        p = prim["p"]
        if prim["kind"] == capi.PTB_PRIM_TRIANGLE:
            builder.triangles(p[:9], p[9:18], cull=bool(prim["cull_backface"]))
        else:
            builder.spheres(p[:4])
    return index


MATERIALS = [
    dict(diffuse=(0.8, 0.6, 0.3, 1.0), ior=1.0, emission=(0, 0, 0, 0), bsdf=capi.PTB_BSDF_LAMBERT, one_way=False),
    dict(diffuse=(0.9, 0.95, 1.0, 1.0), ior=1.5, emission=(0, 0, 0, 0), bsdf=capi.PTB_BSDF_GLASS, one_way=False),
    dict(diffuse=(1.0, 1.0, 1.0, 1.0), ior=1.0, emission=(0, 0, 0, 0), bsdf=capi.PTB_BSDF_MIRROR, one_way=False),
    dict(diffuse=(0.5, 0.5, 0.5, 1.0), ior=1.0, emission=(0.2, 0.1, 0, 1), bsdf=capi.PTB_BSDF_MIRROR, one_way=True),
]


def materials():
    out = np.zeros(len(MATERIALS), capi.MATERIAL_DTYPE)
    for m, spec in zip(out, MATERIALS):
        m["diffuse"], m["emission"] = spec["diffuse"], spec["emission"]
        m["refractive_index"], m["bsdf"], m["one_way"] = spec["ior"], spec["bsdf"], 1 if spec["one_way"] else 0
    return out


def inputs(seed=2024):
    """positions on / near the primitives, incoming directions from both sides of the normal, light spectra, engine seeds"""
    rng = np.random.Generator(np.random.PCG64(seed))
    positions = rng.uniform(-1.0, 1.0, size=(N, 3)).astype(np.float32)
    normals = _unit(rng.normal(size=(N, 3)))
    normals[: N // 4] = np.float32([0.0, 1.0, 0.0])  # axis-aligned normals take localToGlobal's special branches
    normals[N // 4: N // 3] = np.float32([0.0, 0.0, -1.0])
    normals[N // 3: N // 3 + 50] = np.float32([1.0, 0.0, 0.0])
    d_in = _unit(rng.normal(size=(N, 3)))
    grazing = slice(N // 2, N // 2 + 200)  # total internal reflection and near-grazing incidence
    tangent = _unit(np.cross(normals[grazing], d_in[grazing]))
    d_in[grazing] = _unit(tangent + normals[grazing] * rng.uniform(-0.05, 0.05, size=(200, 1)).astype(np.float32))
    d_out = _unit(rng.normal(size=(N, 3)))
    light = rng.uniform(0.0, 2.0, size=(N, 4)).astype(np.float32)
    seeds = rng.integers(1, 2**62, N, dtype=np.int64).astype(np.uint64)
    return dict(positions=positions, normals=normals, d_in=d_in, d_out=d_out, light=light, seeds=seeds,
                propagate=np.concatenate([d_in, positions, normals], axis=1), spectrum=np.concatenate([d_in, d_out, normals, light], axis=1))


def next_draw_from_state(states):
    """Output of the reference engine's next call given its raw state (base.h:28-35): hi32(state * 0xD989BCACC137DCD5)."""
    with np.errstate(over="ignore"):
        return ((np.asarray(states, np.uint64) * np.uint64(0xD989BCACC137DCD5)) >> np.uint64(32)).astype(np.uint32)
