"""Adversarial geometry for the certified closest-hit walk (csrc/traverse.cuh): very large and very thin triangles met at
grazing incidence, where Moeller-Trumbore's |det| sits around the 1e-6 rejection threshold of
Triangle::getIntersection (reference src/scene/object.cpp:146-182) and the reported distance can leave the triangle's
own bounding box.  Used by the CPU test of the certificate theorem and by the -m gpu parity test."""
import numpy as np

from cpupathtrace_b200 import capi


def _unit(v):
    v = np.asarray(v, np.float64)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def sliver_scene(seed=7, n_small=300):
    """Cornell-sized planes (+-20), fans of slivers 40 units long and 1e-4..1e-2 wide, needles, and a cloud of small
    triangles that supplies the competing hits."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tris = []
    # large axis-aligned planes as two triangles each, like makePlane
    for y in (-1.0, 1.0):
        a, b, c, d = (-20, y, -20), (20, y, -20), (20, y, 20), (-20, y, 20)
        tris += [(a, b, c), (c, d, a)]
    for x in (-1.0, 1.0):
        a, b, c, d = (x, -20, -20), (x, 20, -20), (x, 20, 20), (x, -20, 20)
        tris += [(a, b, c), (c, d, a)]
    # slivers: long in one direction, almost degenerate in the other
    for k in range(60):
        o = rng.uniform(-1, 1, 3)
        u = _unit(rng.normal(size=3)) * rng.choice([5.0, 20.0, 40.0])
        v = _unit(np.cross(u, rng.normal(size=3))) * rng.choice([1e-4, 1e-3, 1e-2])
        tris.append((o - u / 2, o + u / 2, o + v))
    # slightly tilted copies of a big triangle (near-coplanar stack: ties and near-ties at grazing angles)
    base = np.array([(-15.0, -0.5, -15.0), (15.0, -0.5, -15.0), (0.0, -0.5, 15.0)])
    for k in range(12):
        tilt = base.copy()
        tilt[:, 1] += rng.uniform(-2e-3, 2e-3, 3)
        tris.append(tuple(tilt))
    for k in range(n_small):
        o = rng.uniform(-1, 1, 3)
        tris.append((o, o + rng.normal(size=3) * 0.08, o + rng.normal(size=3) * 0.08))

    prims = np.zeros(len(tris), capi.PRIM_DTYPE)
    for prim, (a, b, c) in zip(prims, tris):
        a, b, c = (np.asarray(p, np.float32) for p in (a, b, c))
        n = np.cross((b - a).astype(np.float64), (c - a).astype(np.float64))
        ln = np.linalg.norm(n)
        n = (n / ln if ln > 0 else np.array([0.0, 1.0, 0.0])).astype(np.float32)
        prim["kind"] = capi.PTB_PRIM_TRIANGLE
        prim["cull_backface"] = 0
        prim["p"][:] = np.concatenate([a, b, c, n, n, n])
    mats = np.zeros(1, capi.MATERIAL_DTYPE)
    mats[0]["diffuse"] = (1, 1, 1, 1)
    mats[0]["refractive_index"] = 1.0
    return prims, mats, np.zeros(0, capi.LIGHT_DTYPE)


def fine_scene(seed=9, grid=140):
    """What the guard is built for: a few very large triangles (Cornell-style planes of +-20 and +-1, a tall box) around a
    finely tessellated bumpy sheet (edge products below the safe threshold) and a sphere."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tris = []
    for y in (-1.0, 1.0):
        a, b, c, d = (-20, y, -20), (20, y, -20), (20, y, 20), (-20, y, 20)
        tris += [(a, b, c), (c, d, a)]
    for x in (-1.0, 1.0):
        a, b, c, d = (x, -1, -1), (x, 1, -1), (x, 1, 1), (x, -1, 1)
        tris += [(a, b, c), (c, d, a)]
    a, b, c, d = (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)
    tris += [(a, b, c), (c, d, a)]
    # fine sheet: grid x grid cells over [-0.6, 0.6]^2, height field with bumps -> edges ~ 1.2 / grid
    u = np.linspace(-0.6, 0.6, grid + 1)
    xx, zz = np.meshgrid(u, u, indexing="ij")
    yy = -0.3 + 0.15 * np.sin(7 * xx) * np.cos(5 * zz) + 0.001 * rng.normal(size=xx.shape)
    pts = np.stack([xx, yy, zz], axis=-1)
    for i in range(grid):
        for j in range(grid):
            p00, p10, p11, p01 = pts[i, j], pts[i + 1, j], pts[i + 1, j + 1], pts[i, j + 1]
            tris += [(p00, p10, p11), (p00, p11, p01)]
    prims = np.zeros(len(tris) + 1, capi.PRIM_DTYPE)
    for prim, (a, b, c) in zip(prims, tris):
        a, b, c = (np.asarray(p, np.float32) for p in (a, b, c))
        n = np.cross((b - a).astype(np.float64), (c - a).astype(np.float64))
        n = (n / np.linalg.norm(n)).astype(np.float32)
        prim["kind"] = capi.PTB_PRIM_TRIANGLE
        prim["p"][:] = np.concatenate([a, b, c, n, n, n])
    prims[-1]["kind"] = capi.PTB_PRIM_SPHERE
    prims[-1]["p"][:4] = (0.5, -0.5, 0.5, 0.3)
    mats = np.zeros(1, capi.MATERIAL_DTYPE)
    mats[0]["diffuse"] = (1, 1, 1, 1)
    mats[0]["refractive_index"] = 1.0
    return prims, mats, np.zeros(0, capi.LIGHT_DTYPE)


def mixed_rays(prims, n, seed=13):
    """Half grazing rays (see grazing_rays, aimed at the large triangles in particular), half uniformly random rays
    inside the unit box."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tri = prims[prims["kind"] == capi.PTB_PRIM_TRIANGLE]
    p = tri["p"].astype(np.float64)
    m = np.linalg.norm(p[:, 3:6] - p[:, 0:3], axis=1) * np.linalg.norm(p[:, 6:9] - p[:, 0:3], axis=1)
    large = tri[m > 1e-2]
    parts = [grazing_rays(large, n // 4, seed), grazing_rays(tri, n // 4, seed + 1)]
    k = n - 2 * (n // 4)
    o = rng.uniform(-0.98, 0.98, size=(k, 3)).astype(np.float32)
    d = rng.normal(size=(k, 3)).astype(np.float32)
    l2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    d = d * (np.float32(1.0) / np.sqrt(l2))[:, None]
    parts.append(np.concatenate([o, d.astype(np.float32)], axis=1))
    return np.concatenate(parts)


def grazing_rays(prims, n, seed=11):
    """Rays that graze the triangles: origin near a random triangle's plane, direction in the plane plus a normal
    component drawn log-uniformly from 1e-8..1e-2 (so |det| = |n . d| * 2 area sweeps across the 1e-6 threshold),
    a third starting exactly on a vertex / on the plane (the self-intersection regime of bounce rays)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = prims["p"].astype(np.float64)
    idx = rng.integers(0, len(prims), n)
    a, b, c = p[idx, 0:3], p[idx, 3:6], p[idx, 6:9]
    normal = np.cross(b - a, c - a)
    area2 = np.linalg.norm(normal, axis=1, keepdims=True)
    normal = normal / np.maximum(area2, 1e-30)
    u, v = rng.uniform(0, 1, (n, 1)), rng.uniform(0, 1, (n, 1))
    flip = (u + v) > 1
    u, v = np.where(flip, 1 - u, u), np.where(flip, 1 - v, v)
    target = a + (b - a) * u + (c - a) * v
    in_plane = _unit(np.cross(normal, rng.normal(size=(n, 3))))
    # det = dot(ab, cross(d, ac)) = -(d . (ab x ac)): pick the normal component so that |det| straddles 1e-6
    want_det = 10.0 ** rng.uniform(-8, -4, (n, 1)) * rng.choice([-1.0, 1.0], (n, 1))
    eps_n = np.clip(want_det / np.maximum(area2, 1e-12), -0.05, 0.05)
    wide = rng.uniform(0, 1, (n, 1)) < 0.3
    eps_n = np.where(wide, 10.0 ** rng.uniform(-8, -2, (n, 1)) * rng.choice([-1.0, 1.0], (n, 1)), eps_n)
    d = _unit(in_plane + normal * eps_n)
    dist = rng.uniform(0.05, 30.0, (n, 1))
    o = target - d * dist
    third = n // 3
    o[:third] = a[:third]  # start on a vertex
    o[third:2 * third] = target[third:2 * third] + d[third:2 * third] * 1e-3  # just past a surface point, like a bounce ray
    d32 = d.astype(np.float32)
    l2 = (d32[:, 0] * d32[:, 0] + d32[:, 1] * d32[:, 1]) + d32[:, 2] * d32[:, 2]
    d32 = d32 * (np.float32(1.0) / np.sqrt(l2))[:, None]
    return np.concatenate([o.astype(np.float32), d32.astype(np.float32)], axis=1)
