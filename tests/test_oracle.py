"""CPU tests (-m "not gpu") of the oracle: the plain-C restatement (oracle/pt_oracle.c) is pinned against

  * the golden fixtures generated from the unmodified reference (tests/golden/, script make_golden.py),
  * the reference itself when oracle/_ref is present (bit for bit: same libm, same host),
  * the reference's own known-answer tests,

and the libm restatement used by the device code is pinned against the system libm.
"""
import ctypes as C
import os

import numpy as np
import pytest

from cpupathtrace_b200 import REPO_ROOT, scenes
from helpers import GOLDEN_SCENES, camera_kwargs, load_golden, pod_camera
from oracle import pto


def _oracle_scene(g):
    return pto.OracleScene(g["prims"], g["materials"], g["lights"])


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_restatement_matches_golden_hits(name):
    g = load_golden("hits", name)
    t, prim = _oracle_scene(g).intersect(g["rays"])
    assert np.array_equal(prim, g["prim"])
    assert np.array_equal(t, g["t"])  # bit-exact, including the negative "miss" values of the root box test


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_restatement_matches_golden_samples(name):
    g = load_golden("samples", name)
    w, h = (int(v) for v in g["size"])
    rgba, counters = _oracle_scene(g).render_samples(pod_camera(camera_kwargs(g["camera"])), w, h, float(g["epsilon"]), g["pixels"], g["seeds"])
    assert np.array_equal(rgba, g["rgba"])
    assert counters["samples"] == len(g["seeds"]) and counters["closest_rays"] >= counters["vertices"]


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_restatement_matches_golden_tiles(name):
    """processItem with one sequential engine: fixed spp and the adaptive path (min != max, candidate merge)."""
    g = load_golden("tile", name)
    w, h = (int(v) for v in g["size"])
    scene = _oracle_scene(g)
    camera = pod_camera(camera_kwargs(g["camera"]))
    rect = tuple(int(v) for v in g["rect"])
    for key, spp_key, seed in (("fixed", "fixed_spp", g["seeds"][0]), ("adaptive", "adaptive_spp", g["seeds"][1])):
        lo, hi = (int(v) for v in g[spp_key])
        got = scene.process_item(camera, w, h, lo, hi, float(g["epsilon"]), rect, int(seed))
        assert np.array_equal(got, g[key]), key


def test_reference_kats_on_restatement():
    """AABB slab KATs (reference test/scene/boundig_box_test.cpp:24-46), two-sphere identity KAT
    (test/scene/scene_test.cpp:21-46), empty scene (test/render_test.cpp:14-29)."""
    s = np.float32(np.sqrt(np.float32(2.0)) / 2)
    inv = np.float32(1.0) / np.sqrt(np.float32(2.0))
    rays = np.array([
        [5, 0, 0, -1, 0, 0],            # axis ray from 5 -> 4.0
        [1.5, 0, 0, -inv, -inv, 0],     # diagonal ray from 1.5 -> sqrt(2)/2
        [0.5, 0, 0, -1, 0, 0],          # origin inside -> 0
        [5, 0, 0, 1, 0, 0],             # pointing away -> miss
        [5, -2, -2, -1, 0, 0],          # offset -> miss
    ], np.float32)
    t = pto.aabb_intersect((-1, -1, -1), (1, 1, 1), rays)
    assert t[0] == 4.0 and abs(t[1] - s) <= 1e-6 and t[2] == 0.0 and t[3] < 0 and t[4] < 0

    from cpupathtrace_b200 import capi

    prims = np.zeros(2, capi.PRIM_DTYPE)
    prims["kind"] = capi.PTB_PRIM_SPHERE
    prims[0]["p"][:4] = (-1, -1, -1, 1)
    prims[1]["p"][:4] = (1, 1, 1, 1)
    mats = np.zeros(1, capi.MATERIAL_DTYPE)
    mats[0] = ((1, 1, 1, 1), (0, 0, 0, 0), 1.0, 0, 0, 0)
    scene = pto.OracleScene(prims, mats, np.zeros(0, capi.LIGHT_DTYPE))
    t, prim = scene.intersect(np.array([[-0.5, -0.5, -5, 0, 0, 1], [0.5, 0.5, -5, 0, 0, 1], [0, 0, 0, 0, 0, 1]], np.float32))
    assert t[0] >= 0 and prim[0] == 0 and t[1] >= 0 and prim[1] == 1 and t[2] < 0 and prim[2] == -1

    empty = pto.OracleScene(np.zeros(0, capi.PRIM_DTYPE), mats, np.zeros(0, capi.LIGHT_DTYPE))
    cam = pod_camera(dict(origin=(0, 0, 0), look_at=(0, 0, 1), up=(0, 1, 0), focal_length=1.0, height=1.0, aspect_ratio=1.0))
    assert (empty.process_item(cam, 1, 1, 1, 1, 1e-3, (0, 0, 1, 1), 5) == 0).all()


def test_restatement_equals_reference_bit_for_bit(ref):
    """With oracle/_ref present: fresh random inputs, not just the committed fixtures."""
    spec = scenes.mixed_materials(seed=99, n_tris=300)
    prims, materials, lights = spec.to_pod(ref)
    oracle = pto.OracleScene(prims, materials, lights)
    scene = spec.build(ref)
    from conftest import random_rays

    rays = random_rays(20000, seed=123, box=2.2)
    t_ref, id_ref = scene.intersect(rays)
    t, prim = oracle.intersect(rays)
    assert np.array_equal(t_ref, t) and np.array_equal(id_ref, prim)

    kw = dict(origin=(0.1, 0.2, -1.8), look_at=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), focal_length=0.8, height=1.0, aspect_ratio=-1.2,
              aperture_width=0.03, aperture_height=0.05, sampler=1, hex_ratio=0.0, focal_plane_dist=1.7)
    rng = np.random.Generator(np.random.PCG64(77))
    n = 4000
    pixels = np.stack([rng.integers(0, 80, n), rng.integers(0, 60, n)], axis=1).astype(np.int32)
    seeds = rng.integers(1, 2**63 - 1, n, dtype=np.int64).astype(np.uint64)
    want = scene.render_samples(ref.camera(**kw), 80, 60, 2e-3, pixels, seeds)
    got, _ = oracle.render_samples(pod_camera(kw), 80, 60, 2e-3, pixels, seeds)
    assert np.array_equal(want, got)

    # light sampling on its own, including the engine's draw order
    for seed in (1, 2, 3, 1234567):
        a, na = scene.sample_lights((0.1, -0.4, 0.2), (0, 1, 0), seed)
        b, nb = oracle.sample_lights((0.1, -0.4, 0.2), seed)
        assert na == nb and np.array_equal(a, b)

    # camera rays
    xy = rng.uniform(-1, 1, size=(500, 2)).astype(np.float32)
    cam_seeds = rng.integers(1, 2**62, 500, dtype=np.int64).astype(np.uint64)
    assert np.array_equal(ref.camera(**kw).shoot(xy, 1 / 80, 1 / 60, cam_seeds), pto.camera_shoot(pod_camera(kw), xy, 1 / 80, 1 / 60, cam_seeds))


def test_golden_fixtures_are_current(ref):
    """The committed fixtures equal what the reference produces today (guards against stale files)."""
    for name in GOLDEN_SCENES:
        g = load_golden("hits", name)
        if name == "cornell_mesh":
            spec = scenes.cornell_demo(("obj", scenes.standin_obj(40, 30)))
        elif name == "mixed":
            spec = scenes.mixed_materials(n_tris=240)
        else:
            spec = scenes.advanced_render()
        t, ids = spec.build(ref).intersect(g["rays"])
        assert np.array_equal(t, g["t"]) and np.array_equal(ids, g["prim"])


@pytest.mark.parametrize("lib", ["libm_check.so", "libm_check_fma.so"])
def test_libm_restatement_matches_system_libm(lib):
    """glibc sinf/cosf/powf(.,0.5)/acosf restated operation for operation (the device code uses the same constants
    from csrc/glibc_libm_tables.h): zero disagreements with the system libm, fused and unfused fp64 evaluation."""
    path = os.path.join(REPO_ROOT, "oracle", lib)
    if not os.path.exists(path):
        pytest.skip(f"{lib} not built (run __graft_entry__.build())")
    check = C.CDLL(path).pto_libm_check
    check.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    out = (C.c_uint64 * 4)()
    check(3_000_000, 20261018, out)
    assert list(out) == [0, 0, 0, 0]


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_certificate_theorem_on_arbitrary_hierarchies(name):
    """The claim behind PTB_FLAG_CERTIFIED_CLOSEST (csrc/traverse.cuh), checked on the CPU independently of the CUDA code:
    walking ANY hierarchy with exact union boxes -- random splits of a random permutation, a chain, a balanced tree --
    and keeping a hit only when it carries the certificate gives exactly what the reference walk returns on the
    reference tree (here: the golden hits of the unmodified reference, a third of them aimed at shared vertices and edges).
    Rays without a certificate are the ones the kernels re-trace on the reference tree; they must be the minority."""
    g = load_golden("hits", name)
    scene = _oracle_scene(g)
    rays, want_t, want_prim = g["rays"], g["t"], g["prim"]
    for seed, shape in ((1, 0), (2, 0), (5, 0), (6, 0), (3, 1), (4, 2)):
        if shape == 1 and len(g["prims"]) > 3000:
            continue  # the chain is quadratic in the primitive count
        t, prim, certain = scene.intersect_certified(rays, seed, shape)
        hit = want_t >= 0
        assert certain.mean() > 0.5
        assert np.array_equal(prim[certain], want_prim[certain]), (seed, shape)
        assert np.array_equal(t[certain & hit], want_t[certain & hit]), (seed, shape)
        assert (t[certain & ~hit] < 0).all()
        # and the certificate is not vacuous: uncertified rays do differ from the reference now and then on a foreign tree
        if name != "advanced":
            assert (~certain).sum() > 0

    # fresh rays, a third of them aimed at triangle vertices and edge midpoints (the tie cases), against the restatement's
    # own reference-order walk (pinned bit for bit to the reference above)
    from conftest import random_rays

    tris = g["prims"]["p"][g["prims"]["kind"] == 0][:, :9]
    aim = np.concatenate([tris.reshape(-1, 3), 0.5 * (tris[:, 0:3] + tris[:, 3:6])]) if len(tris) else None
    if aim is not None:
        aim = aim[np.random.Generator(np.random.PCG64(9)).permutation(len(aim))][:20000]
    rays = random_rays(60000, seed=13, box=1.3, aim=aim)
    want_t, want_prim = scene.intersect(rays)
    differing = 0
    for seed in (11, 12):
        t, prim, certain = scene.intersect_certified(rays, seed, 0)
        hit = want_t >= 0
        assert np.array_equal(prim[certain], want_prim[certain]) and np.array_equal(t[certain & hit], want_t[certain & hit])
        assert (t[certain & ~hit] < 0).all() and certain.mean() > 0.6
        differing += int((prim[~certain] != want_prim[~certain]).sum())
    if name == "cornell_mesh":
        assert differing > 0  # ties on a foreign tree do come out differently: that is what the re-trace is for


def test_unit_functions_match_reference(ref):
    """Object::getSurfaceNormal / sampleSurface and BSDF::propagateRay / getSpectrum of the plain-C restatement against
    the unmodified reference's virtual methods (object.cpp:86-144, 101-116, 192-207; propagation.cpp:89-217), bit for
    bit, including how many engine draws each call consumes."""
    import unit_cases as uc
    from cpupathtrace_b200 import capi

    prims, mats, x = uc.prims(), uc.materials(), uc.inputs()
    builder = ref.builder()
    handles = [builder.material(**spec) for spec in uc.MATERIALS]
    index = uc.add_to_builder(builder, prims)
    states = capi.xorshift_state(x["seeds"])
    for prim, obj in zip(prims, index):
        assert np.array_equal(pto.prim_normal(prim, x["positions"]), builder.object_normal(obj, x["positions"]), equal_nan=True)
        got, after = pto.prim_sample(prim, states)
        want, next_draw = builder.object_sample(obj, x["seeds"])
        assert np.array_equal(got, want) and np.array_equal(uc.next_draw_from_state(after), next_draw)
    for material, handle in zip(mats, handles):
        got, after = pto.bsdf_propagate(material, 1e-3, x["propagate"], states)
        want, next_draw = builder.bsdf_propagate(handle, 1e-3, x["propagate"], x["seeds"])
        assert np.array_equal(got, want, equal_nan=True) and np.array_equal(uc.next_draw_from_state(after), next_draw)
        for synthetic in (False, True):
            assert np.array_equal(pto.bsdf_spectrum(material, synthetic, x["spectrum"]), builder.bsdf_spectrum(handle, synthetic, x["spectrum"]))
    builder.close()


def test_certificate_guard_on_adversarial_geometry():
    """VERDICT r1 W1 on the CPU.  The certificate is exact about what the walk tests; about what it prunes it assumes
    that a primitive cannot report a hit in front of its own box.  tests/stress_cases.py breaks that assumption (slivers,
    grazing rays with |det| around the 1e-6 threshold of object.cpp:146-182): the unguarded walk then certifies answers
    the reference does not give.  With the guard table (csrc/cert_guard.h, restated in pt_oracle.c) such a scene is not
    certifiable at all, and on a scene the table does cover every certified answer equals the reference's on three
    hierarchy shapes."""
    import stress_cases as sc

    prims, mats, lights = sc.sliver_scene()
    scene = pto.OracleScene(prims, mats, lights)
    rays = sc.grazing_rays(prims, 120_000)
    want_t, want_prim = scene.intersect(rays)
    hit = want_t >= 0
    assert not scene.certifiable()
    t, prim, certain = scene.intersect_certified(rays, 1, 0, guarded=False)
    wrong = certain & np.where(hit, (prim != want_prim) | (t != want_t), prim != -1)
    assert wrong.sum() > 0, "the stress scene no longer exposes the hazard it was built for"
    t, prim, certain = scene.intersect_certified(rays, 1, 0, guarded=True)
    assert not certain.any()

    prims, mats, lights = sc.fine_scene()
    scene = pto.OracleScene(prims, mats, lights)
    assert scene.certifiable()
    rays = sc.mixed_rays(prims, 90_000)
    want_t, want_prim = scene.intersect(rays)
    hit = want_t >= 0
    for seed, shape in ((1, 0), (2, 2)):
        t, prim, certain = scene.intersect_certified(rays, seed, shape, guarded=True)
        assert 0.3 < certain.mean() < 0.95
        assert np.array_equal(prim[certain], want_prim[certain]) and np.array_equal(t[certain & hit], want_t[certain & hit])
        assert (t[certain & ~hit] < 0).all()
