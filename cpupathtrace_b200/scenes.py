"""Scene descriptions shared by the tests, the smoke test and bench.py.

A :class:`SceneSpec` is an ordered list of construction steps expressed in the vocabulary of the reference's public
API (makePlane, makeBox, io::loadMesh, Triangle, Sphere, ConstantMaterial + BSDF, PointLightSource).  It can be
replayed into any harness library (reference or b200 build), which guarantees that both sides see the same scene.

The scene definitions restate the reference's own callers:
  cornell_demo     reference demo/main.cpp:36-203          (BASELINE.json configs C1, C2, C4, C5)
  benchmark_box    reference benchmark/main.cpp:34-57      (renderSceneBox)
  two_spheres      reference test/scene/scene_test.cpp:8-19
  simple_render    reference test/render_test.cpp:31-44
  advanced_render  reference test/render_test.cpp:54-82
The XYZ RGB dragon asset is not shipped with the reference checkout (.MISSING_LARGE_BLOBS); standin_mesh() is the
deterministic stand-in defined in SURVEY.md section 8d and every report names it "stand-in-N".
"""
import math
from dataclasses import dataclass, field

import numpy as np

LAMBERT, GLASS, MIRROR = 0, 1, 2


@dataclass
class SceneSpec:
    steps: list = field(default_factory=list)
    n_materials: int = 0

    def material(self, diffuse=(1, 1, 1, 1), ior=1.0, emission=(0, 0, 0, 0), bsdf=LAMBERT, one_way=False):
        self.steps.append(("material", tuple(diffuse), float(ior), tuple(emission), int(bsdf), bool(one_way)))
        self.n_materials += 1
        return self.n_materials - 1

    def plane(self, a, b, cull=False, material=-1):
        self.steps.append(("plane", tuple(a), tuple(b), cull, material))

    def box(self, a, b, cull=False, transform=None, material=-1):
        self.steps.append(("box", tuple(a), tuple(b), cull, transform, material))

    def triangles(self, verts, normals=None, cull=False, material=-1):
        self.steps.append(("triangles", np.asarray(verts, np.float32), None if normals is None else np.asarray(normals, np.float32), cull, material))

    def spheres(self, spheres, material=-1):
        self.steps.append(("spheres", np.asarray(spheres, np.float32), material))

    def mesh_obj(self, text, transform=None, cull=True, smooth=True, material=-1):
        self.steps.append(("mesh_obj", text, transform, cull, smooth, material))

    def point_light(self, pos, rgba):
        self.steps.append(("point_light", tuple(pos), tuple(rgba)))

    def replay(self, pth):
        """Replays the steps into a new builder of the given harness library; returns the builder."""
        b = pth.builder()
        for step in self.steps:
            kind = step[0]
            if kind == "material":
                b.material(*step[1:])
            elif kind == "plane":
                b.plane(step[1], step[2], step[3], step[4])
            elif kind == "box":
                b.box(step[1], step[2], step[3], step[4], step[5])
            elif kind == "triangles":
                b.triangles(step[1], step[2], step[3], step[4])
            elif kind == "spheres":
                b.spheres(step[1], step[2])
            elif kind == "mesh_obj":
                b.mesh_obj(step[1], step[2], step[3], step[4], step[5])
            elif kind == "point_light":
                b.point_light(step[1], step[2])
        return b

    def build(self, pth):
        b = self.replay(pth)
        scene = b.scene()
        b.close()
        return scene

    def to_pod(self, pth):
        """Lowers the spec to the C-ABI's POD arrays (capi.PRIM_DTYPE / MATERIAL_DTYPE / LIGHT_DTYPE).

        Geometry is produced by the given harness library's own mesh code (makePlane, makeBox, io::loadMesh) and read
        back, so the POD scene is identical to what that library's Scene would contain, in the same object order."""
        from . import capi

        materials, lights, kinds, mats, spheres = [], [], [], [], []
        b = pth.builder()
        try:
            for step in self.steps:
                kind = step[0]
                before = b.object_count()
                material = -1
                if kind == "material":
                    b.material(*step[1:])
                    materials.append(step[1:])
                    continue
                if kind == "point_light":
                    lights.append((step[1], step[2]))
                    continue
                if kind == "plane":
                    b.plane(step[1], step[2], step[3], step[4])
                    material, cull = step[4], step[3]
                elif kind == "box":
                    b.box(step[1], step[2], step[3], step[4], step[5])
                    material, cull = step[5], step[3]
                elif kind == "triangles":
                    b.triangles(step[1], step[2], step[3], step[4])
                    material, cull = step[4], step[3]
                elif kind == "mesh_obj":
                    b.mesh_obj(step[1], step[2], step[3], step[4], step[5])
                    material, cull = step[5], step[3]
                elif kind == "spheres":
                    b.spheres(step[1], step[2])
                    material, cull = step[2], False
                    spheres.extend(np.asarray(step[1], np.float32).reshape(-1, 4))
                added = b.object_count() - before
                kinds.extend([(capi.PTB_PRIM_SPHERE if kind == "spheres" else capi.PTB_PRIM_TRIANGLE, bool(cull))] * added)
                mats.extend([material] * added)
            tris = b.get_triangles()
        finally:
            b.close()

        default_index = len(materials)
        mat_arr = np.zeros(len(materials) + 1, capi.MATERIAL_DTYPE)
        for i, (diffuse, ior, emission, bsdf, one_way) in enumerate(materials):
            mat_arr[i] = (diffuse, emission, ior, bsdf, 1 if one_way else 0, 0)
        mat_arr[default_index] = ((1, 1, 1, 1), (0, 0, 0, 0), 1.0, capi.PTB_BSDF_LAMBERT, 0, 0)

        prims = np.zeros(len(kinds), capi.PRIM_DTYPE)
        sphere_iter = iter(spheres)
        for i, ((kind, cull), material) in enumerate(zip(kinds, mats)):
            prims[i]["kind"] = kind
            prims[i]["material"] = material if material >= 0 else default_index
            prims[i]["cull_backface"] = 1 if cull else 0
            if kind == capi.PTB_PRIM_TRIANGLE:
                prims[i]["p"] = tris[i]
            else:
                prims[i]["p"][:4] = next(sphere_iter)
        light_arr = np.zeros(len(lights), capi.LIGHT_DTYPE)
        for i, (pos, rgba) in enumerate(lights):
            light_arr[i] = (pos, rgba)
        return prims, mat_arr, light_arr


# ---------------------------------------------------------------------------------------------- stand-in mesh / soup


def standin_vertices(nu, nv):
    """Vertices (float64) of the bumpy torus of SURVEY.md 8d in raw dragon-like units, i-major."""
    i = np.arange(nu, dtype=np.float64)[:, None]
    j = np.arange(nv, dtype=np.float64)[None, :]
    u = 2.0 * math.pi * i / nu
    v = 2.0 * math.pi * j / nv
    big_r, small_r = 55.0, 18.0
    bump = 1.0 + 0.25 * np.sin(7 * u) * np.sin(5 * v) + 0.1 * np.sin(23 * u + 3 * v)
    ring = big_r + small_r * bump * np.cos(v)
    x = ring * np.cos(u)
    z = ring * np.sin(u)
    y = small_r * bump * np.sin(v) + 20.0 + 10.0 * np.sin(3 * u)
    return np.stack([x, y, z], axis=-1).reshape(-1, 3)


def standin_faces(nu, nv):
    """0-based vertex indices, two triangles per quad: (a b c) then (a c d) with wrap-around in both directions."""
    i = np.arange(nu)[:, None]
    j = np.arange(nv)[None, :]
    a = i * nv + j
    b = ((i + 1) % nu) * nv + j
    c = ((i + 1) % nu) * nv + (j + 1) % nv
    d = i * nv + (j + 1) % nv
    f1 = np.stack([a, b, c], axis=-1).reshape(-1, 3)
    f2 = np.stack([a, c, d], axis=-1).reshape(-1, 3)
    return np.stack([f1, f2], axis=1).reshape(-1, 3)


def standin_obj(nu, nv):
    """The stand-in as Wavefront OBJ text ("v %.6f %.6f %.6f" / "f a b c"), to go through io::loadMesh."""
    verts = standin_vertices(nu, nv)
    faces = standin_faces(nu, nv) + 1
    lines = ["v %.6f %.6f %.6f" % tuple(p) for p in verts]
    lines += ["f %d %d %d" % tuple(f) for f in faces]
    return "\n".join(lines) + "\n"


def standin_triangles(nu, nv, transform):
    """The stand-in as explicit triangles with smooth vertex normals, bypassing OBJ text (for large benchmark scenes).

    Geometry matches loadMesh only up to fp32 rounding of the text round trip; both benchmark arms are fed these same
    arrays, parity of the OBJ path itself is tested separately on standin_obj().
    """
    m = np.asarray(transform, np.float64).reshape(4, 4)
    raw = standin_vertices(nu, nv)
    verts = (raw @ m[:3, :3].T + m[:3, 3]).astype(np.float32)
    faces = standin_faces(nu, nv)
    tri = verts[faces]  # (F, 3, 3)
    fn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]).astype(np.float64)
    fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    vn = np.zeros((len(verts), 3), np.float64)
    for k in range(3):
        np.add.at(vn, faces[:, k], fn)
    vn /= np.linalg.norm(vn, axis=1, keepdims=True)
    normals = vn[faces].astype(np.float32)
    return tri.reshape(-1, 9), normals.reshape(-1, 9)


def soup_triangles(n, seed):
    """C3: n random triangles, centres uniform in [-1,1]^3, vertices = centre + uniform[-s,s]^3, s = 0.5 n^(-1/3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    s = 0.5 * n ** (-1.0 / 3.0)
    centres = rng.uniform(-1.0, 1.0, size=(n, 1, 3))
    offsets = rng.uniform(-s, s, size=(n, 3, 3))
    return (centres + offsets).astype(np.float32).reshape(n, 9)


# ---------------------------------------------------------------------------------------------- reference scenes

DEMO_DRAGON_TRANSFORM = (0.005, 0.0, 0.0, 0.4, 0.0, 0.005, 0.0, -0.8, 0.0, 0.0, 0.005, -0.75, 0.0, 0.0, 0.0, 1.0)


def cornell_demo(mesh=None, epsilon=1e-3):
    """demo/main.cpp:52-203.  mesh: None (no dragon), ("obj", text) or ("triangles", verts, normals)."""
    s = SceneSpec()
    ground_y, ceiling_y, walls_x, walls_z = -1.0, 1.0, 1.0, 1.0
    white = (1.0, 1.0, 1.0, 1.0)

    m_ground = s.material(white)
    m_ceiling = s.material(white)
    m_light = s.material(white, 1.0, (1.0, 1.0, 1.0, 1.0))
    m_back = s.material((0.0, 0.0, 1.0, 1.0))
    m_left = s.material((1.0, 0.0, 0.0, 1.0))
    m_front = s.material(white)
    m_right = s.material((0.0, 1.0, 0.0, 1.0))

    light_y = float(np.float32(ceiling_y) - np.float32(epsilon))
    s.plane((20.0, ground_y, -20.0), (-20.0, ground_y, 20.0), True, m_ground)
    s.plane((-20.0, ceiling_y, -20.0), (20.0, ceiling_y, 20.0), True, m_ceiling)
    s.plane((-0.25, light_y, -0.25), (0.25, light_y, 0.25), True, m_light)
    s.plane((-walls_x, ground_y, -walls_z), (walls_x, ceiling_y, -walls_z), True, m_back)
    s.plane((-walls_x, ground_y, -walls_z), (-walls_x, ceiling_y, walls_z), True, m_left)
    s.plane((walls_x, ground_y, walls_z), (-walls_x, ceiling_y, walls_z), True, m_front)
    s.plane((walls_x, ground_y, walls_z), (walls_x, ceiling_y, -walls_z), True, m_right)

    if mesh is not None:
        m_dragon = s.material(white, 1.5, (0, 0, 0, 0), GLASS)
        if mesh[0] == "obj":
            s.mesh_obj(mesh[1], DEMO_DRAGON_TRANSFORM, False, True, m_dragon)
        else:
            s.triangles(mesh[1], mesh[2], False, m_dragon)

    m_sphere = s.material((0.0, 0.0, 1.0, 1.0), 1.0, (0, 0, 0, 0), MIRROR, False)
    s.spheres([[0.5, -1.0 + 0.5, 0.5, 0.5]], m_sphere)

    rot_y = np.float32(0.25)
    c, sn = float(np.cos(rot_y)), float(np.sin(rot_y))
    box_transform = (c, 0.0, sn, -0.5, 0.0, 3.0, 0.0, -0.25, -sn, 0.0, c, 0.5, 0.0, 0.0, 0.0, 1.0)
    m_box = s.material(white)
    lo = tuple(float(np.float32(-1.0) * np.float32(0.3)) for _ in range(3))
    hi = tuple(float(np.float32(1.0) * np.float32(0.3)) for _ in range(3))
    s.box(lo, hi, False, box_transform, m_box)
    return s


def demo_camera(pth_or_none, width, height):
    """demo/main.cpp:46-50: thin lens, circular aperture 0.05, focal plane 3.5, NEGATIVE aspect ratio."""
    aspect = float(np.float32(width) / np.float32(height))
    args = dict(origin=(0.0, 0.0, -3.0), look_at=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), focal_length=1.0, height=1.0, aspect_ratio=-aspect,
                aperture_width=0.05, aperture_height=0.05, sampler=1, hex_ratio=0.0, focal_plane_dist=3.5)
    if pth_or_none is None:
        return args
    return pth_or_none.camera(**args)


def benchmark_box():
    """benchmark/main.cpp:34-57 (renderSceneBox): default-material box + emissive ceiling patch."""
    s = SceneSpec()
    s.box((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), False, None, -1)
    m_light = s.material((1, 1, 1, 1), 1.0, (1.0, 1.0, 1.0, 1.0))
    y = float(np.float32(1.0) - np.float32(0.01))
    s.plane((-0.25, y, -0.25), (0.25, y, 0.25), False, m_light)
    return s


def benchmark_camera(pth):
    return pth.camera((0.0, 0.0, -3.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 1.0, 1.0, -1.0)


def two_spheres():
    s = SceneSpec()
    s.spheres([[-1.0, -1.0, -1.0, 1.0], [1.0, 1.0, 1.0, 1.0]], -1)
    return s


def simple_render():
    s = SceneSpec()
    s.point_light((0.0, 1.0, 0.0), (1.0, 1.0, 1.0, 1.0))
    s.spheres([[0.0, 0.0, 0.6, 0.5]], -1)
    return s


def advanced_render():
    s = SceneSpec()
    s.point_light((0.0, 1.0, 0.0), (1.0, 1.0, 1.0, 1.0))
    m_glass = s.material((1.0, 1.0, 1.0, 1.5), 1.0, (0, 0, 0, 0), GLASS)
    m_emit = s.material((0.8, 0.4, 0.6, 1.0), 1.0, (0.2, 0.1, 0.3, 1.0), LAMBERT)
    m_ground = s.material((0.4, 0.6, 0.4, 1.0), 1.0, (0, 0, 0, 0), LAMBERT)
    s.spheres([[0.1, 0.1, 1.0, 0.5]], m_glass)
    s.spheres([[-0.1, 0.2, 2.0, 0.6]], m_emit)
    s.triangles([[5.0, -1.0, 5.0, 0.0, -1.0, -5.0, -5.0, -1.0, 5.0]], None, False, m_ground)
    return s


def mixed_materials(seed=7, n_tris=400):
    """Synthetic parity scene exercising every primitive / BSDF / light kind at once: a lit room with random glass,
    mirror (one-way and two-way) and diffuse triangles, emissive spheres and triangles, and two point lights."""
    rng = np.random.Generator(np.random.PCG64(seed))
    s = SceneSpec()
    mats = [
        s.material((0.8, 0.8, 0.8, 1.0)),
        s.material((0.9, 0.3, 0.2, 1.0)),
        s.material((1.0, 1.0, 1.0, 1.0), 1.45, (0, 0, 0, 0), GLASS),
        s.material((0.6, 0.9, 0.7, 1.0), 1.7, (0, 0, 0, 0), GLASS),
        s.material((1.0, 1.0, 1.0, 1.0), 1.0, (0, 0, 0, 0), MIRROR, False),
        s.material((1.0, 1.0, 1.0, 1.0), 1.0, (0, 0, 0, 0), MIRROR, True),
    ]
    m_emit_a = s.material((1, 1, 1, 1), 1.0, (2.0, 1.5, 1.0, 1.0))
    m_emit_b = s.material((1, 1, 1, 1), 1.0, (0.2, 0.6, 1.2, 1.0))
    s.box((-2.0, -2.0, -2.0), (2.0, 2.0, 2.0), False, None, mats[0])
    s.plane((-0.5, 1.95, -0.5), (0.5, 1.95, 0.5), True, m_emit_a)
    s.spheres([[-1.2, -1.2, 0.8, 0.35]], m_emit_b)
    s.spheres([[0.9, -1.3, 0.2, 0.6]], mats[4])
    s.spheres([[-0.2, 0.4, 0.9, 0.45]], mats[2])
    centres = rng.uniform(-1.6, 1.6, size=(n_tris, 1, 3))
    tris = (centres + rng.uniform(-0.35, 0.35, size=(n_tris, 3, 3))).astype(np.float32).reshape(n_tris, 9)
    per = n_tris // len(mats)
    for k, m in enumerate(mats):
        chunk = tris[k * per:(k + 1) * per]
        s.triangles(chunk, None, bool(k % 2), m)
    s.point_light((1.5, 1.5, -1.5), (0.6, 0.6, 0.6, 1.0))
    s.point_light((-1.5, 0.5, -1.0), (0.3, 0.2, 0.5, 1.0))
    return s
