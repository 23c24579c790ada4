"""Parity of the CUDA path against the oracle, through the reference's own API (harness) — the -m gpu tests proper.

Bars (BASELINE.json north_star): closest-hit primitive index bit-exact, hit distance within 1 ulp (we expect 0);
per-sample radiance within 1e-4 relative in validation mode (identical random numbers: RandomEngine(seed) per
sample on both sides).  Since the device restates glibc's sinf/cosf/powf/acosf bit for bit (csrc/glibc_libm.cuh) and
contracts no FMA, the measured state is stronger than the bar: every sample is bit-identical to the reference's, so
the tests demand >= 99.99 % bit-exact samples and none outside 1e-4.
"""
import numpy as np
import pytest

from cpupathtrace_b200 import scenes
from conftest import random_rays, ulp_distance

pytestmark = pytest.mark.gpu

RADIANCE_RTOL = 1e-4  # north_star: per-pixel radiance within 1e-4 relative error
RADIANCE_ATOL = 1e-6
MAX_DIVERGED_FRACTION = 0.0
MIN_BIT_EXACT_FRACTION = 0.9999


def _pair(spec, ref, b200):
    return spec.build(ref), spec.build(b200)


def _compare_hits(scene_ref, scene_gpu, rays):
    t_ref, id_ref = scene_ref.intersect(rays)
    t_gpu, id_gpu = scene_gpu.intersect(rays)
    hit_ref = t_ref >= 0
    hit_gpu = t_gpu >= 0
    assert np.array_equal(hit_ref, hit_gpu), f"hit/miss differs on {(hit_ref != hit_gpu).sum()} of {len(rays)} rays"
    assert np.array_equal(id_ref, id_gpu), f"primitive index differs on {(id_ref != id_gpu).sum()} of {len(rays)} rays"
    ulps = ulp_distance(t_ref[hit_ref], t_gpu[hit_ref])
    assert ulps.max(initial=0) <= 1, f"hit distance differs by up to {ulps.max()} ulp"
    # we build without FMA contraction on both sides, so equality is expected, not just 1 ulp
    assert (t_ref[hit_ref] == t_gpu[hit_ref]).all()
    return hit_ref.mean()


def test_scene_kat_two_spheres(ref, b200):
    """reference test/scene/scene_test.cpp:21-46"""
    sr, sg = _pair(scenes.two_spheres(), ref, b200)
    rays = np.array([[-0.5, -0.5, -5, 0, 0, 1], [0.5, 0.5, -5, 0, 0, 1], [0, 0, 0, 0, 0, 1]], np.float32)
    t, ids = sg.intersect(rays)
    assert t[0] >= 0 and ids[0] == 0
    assert t[1] >= 0 and ids[1] == 1
    assert t[2] < 0 and ids[2] == -1
    _compare_hits(sr, sg, rays)
    # the single-ray API goes through the same kernel
    assert sg.intersect_one(rays[0]) == (pytest.approx(float(t[0]), abs=0), 0)


def test_aabb_kat(ref, b200):
    """reference test/scene/boundig_box_test.cpp:8-49: 4.0, sqrt(2)/2, 0, miss, miss for all three axes"""
    rays, want = [], []
    for dim in range(3):
        e = np.eye(3, dtype=np.float32)[dim]
        f = np.float32(-1.0)
        rays.append(np.concatenate([e * f * 5, e * f * -1]))
        want.append(4.0)
        for dim2 in range(3):
            if dim2 == dim:
                continue
            d = (e + np.eye(3, dtype=np.float32)[dim2]) * f * np.float32(-1.0)
            d = d * (np.float32(1.0) / np.sqrt(np.float32((d * d).sum())))
            rays.append(np.concatenate([e * f * np.float32(1.5), d]))
            want.append(np.sqrt(np.float32(2.0)) / 2)
        rays.append(np.concatenate([e * f * np.float32(0.5), e * f * -1]))
        want.append(0.0)
        rays.append(np.concatenate([e * f * 5, e * f]))
        want.append(-1.0)
        rays.append(np.concatenate([(7 * e - 2) * f, e * f * -1]))
        want.append(-1.0)
    rays = np.array(rays, np.float32)
    got = b200.aabb_intersect((-1, -1, -1), (1, 1, 1), rays)
    oracle = ref.aabb_intersect((-1, -1, -1), (1, 1, 1), rays)
    for g, w in zip(got, want):
        if w < 0:
            assert g < 0
        else:
            assert g == pytest.approx(w, rel=4e-7)
    assert np.array_equal(got, oracle)


@pytest.mark.parametrize("name", ["cornell", "cornell_mesh", "mixed", "advanced"])
def test_closest_hit_parity(ref, b200, name):
    if name == "cornell":
        spec = scenes.cornell_demo()
    elif name == "cornell_mesh":
        spec = scenes.cornell_demo(("obj", scenes.standin_obj(120, 80)))
    elif name == "mixed":
        spec = scenes.mixed_materials()
    else:
        spec = scenes.advanced_render()
    builder = spec.replay(ref)
    tris = builder.get_triangles()
    builder.close()
    tris = tris[~np.isnan(tris).any(axis=1)]
    # aim a third of the rays at vertices and edge midpoints: shared edges/vertices produce equal-distance hits on
    # several primitives and exercise the traversal's tie-break rules
    verts = tris[:, :9].reshape(-1, 3)
    mids = 0.5 * (tris[:, 0:3] + tris[:, 3:6])
    aim = np.concatenate([verts, mids])
    rng = np.random.Generator(np.random.PCG64(3))
    aim = aim[rng.permutation(len(aim))][:20000]
    rays = random_rays(60000, seed=11, box=1.1, aim=aim)
    sr, sg = _pair(spec, ref, b200)
    hit_rate = _compare_hits(sr, sg, rays)
    assert hit_rate > 0.3


def test_empty_and_single_primitive_scenes(ref, b200):
    rays = random_rays(2000, seed=5, box=3.0)
    # empty scene: everything misses (reference test/render_test.cpp:14-29)
    sg = scenes.SceneSpec().build(b200)
    t, ids = sg.intersect(rays)
    assert (t < 0).all() and (ids == -1).all()
    # one primitive: the root is a leaf
    for spec in (scenes.simple_render(),):
        sr, sg = _pair(spec, ref, b200)
        _compare_hits(sr, sg, rays)
    one = scenes.SceneSpec()
    one.triangles([[5.0, -1.0, 5.0, 0.0, -1.0, -5.0, -5.0, -1.0, 5.0]], None, True, -1)
    sr, sg = _pair(one, ref, b200)
    _compare_hits(sr, sg, rays)


def _sample_parity(ref, b200, spec, cam_kwargs, width, height, n, seed, epsilon=1e-3):
    sr, sg = _pair(spec, ref, b200)
    cr, cg = ref.camera(**cam_kwargs), b200.camera(**cam_kwargs)
    rng = np.random.Generator(np.random.PCG64(seed))
    pixels = np.stack([rng.integers(0, width, n), rng.integers(0, height, n)], axis=1).astype(np.int32)
    seeds = rng.integers(1, 2**63 - 1, n, dtype=np.int64).astype(np.uint64)
    want = sr.render_samples(cr, width, height, epsilon, pixels, seeds)
    got = sg.render_samples(cg, width, height, epsilon, pixels, seeds)
    assert np.array_equal(want[:, 3], got[:, 3]), "collected flag (alpha) differs"
    err = np.abs(got[:, :3] - want[:, :3])
    tol = RADIANCE_ATOL + RADIANCE_RTOL * np.abs(want[:, :3])
    bad = (err > tol).any(axis=1)
    exact = (got == want).all(axis=1)
    return bad.mean(), exact.mean(), want, got


def test_sample_radiance_parity_cornell(ref, b200):
    cam = scenes.demo_camera(None, 64, 64)
    bad, exact, want, got = _sample_parity(ref, b200, scenes.cornell_demo(("obj", scenes.standin_obj(60, 40))), cam, 64, 64, 20000, seed=21)
    print(f"cornell+mesh: diverged {bad:.5f}, bit-exact {exact:.4f}, mean radiance {want[:, :3].mean():.4f}")
    assert bad <= MAX_DIVERGED_FRACTION
    assert exact >= MIN_BIT_EXACT_FRACTION


def test_sample_radiance_parity_mixed(ref, b200):
    cam = dict(origin=(0.0, 0.0, -1.9), look_at=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), focal_length=0.7, height=1.0, aspect_ratio=-1.5,
               aperture_width=0.04, aperture_height=0.03, sampler=2, hex_ratio=0.5, focal_plane_dist=2.0)
    bad, exact, want, got = _sample_parity(ref, b200, scenes.mixed_materials(), cam, 96, 64, 20000, seed=22)
    print(f"mixed: diverged {bad:.5f}, bit-exact {exact:.4f}")
    assert bad <= MAX_DIVERGED_FRACTION
    assert exact >= MIN_BIT_EXACT_FRACTION


def test_sample_radiance_parity_point_light_pinhole(ref, b200):
    cam = dict(origin=(0.0, 0.0, 0.0), look_at=(0.0, 0.0, 1.0), up=(0.0, 1.0, 0.0), focal_length=0.2, height=0.5, aspect_ratio=1.94)
    bad, exact, want, got = _sample_parity(ref, b200, scenes.advanced_render(), cam, 132, 68, 20000, seed=23)
    print(f"advanced: diverged {bad:.5f}, bit-exact {exact:.4f}")
    assert bad <= MAX_DIVERGED_FRACTION
    assert exact >= MIN_BIT_EXACT_FRACTION


@pytest.mark.parametrize("fast", [False, True], ids=["reference_order", "default_fast_queries"])
def test_query_options_through_the_cpp_api_change_nothing(ref, b200, fast):
    """ptb::RenderControl's query options (certified closest hits on the SAH hierarchy, any-hit shadow rays, zero-weight
    shadow rays skipped) are ON by default; with them and without them (exactly the reference's rays in the reference's
    order) Scene::getIntersection results and every validation-mode sample must be bit-identical to the unmodified
    reference, through the reference's own C++ API."""
    b200.set_fast_queries(fast, fast, fast)
    try:
        spec = scenes.cornell_demo(("obj", scenes.standin_obj(120, 80)))
        builder = spec.replay(ref)
        tris = builder.get_triangles()
        builder.close()
        tris = tris[~np.isnan(tris).any(axis=1)]
        aim = np.concatenate([tris[:, :9].reshape(-1, 3), 0.5 * (tris[:, 0:3] + tris[:, 3:6])])
        rng = np.random.Generator(np.random.PCG64(4))
        rays = random_rays(60000, seed=12, box=1.1, aim=aim[rng.permutation(len(aim))][:20000])
        sr, sg = _pair(spec, ref, b200)
        assert _compare_hits(sr, sg, rays) > 0.3

        cam = scenes.demo_camera(None, 64, 64)
        bad, exact, want, got = _sample_parity(ref, b200, scenes.cornell_demo(("obj", scenes.standin_obj(60, 40))), cam, 64, 64, 20000, seed=24)
        assert bad == 0 and exact == 1.0
        cam = dict(origin=(0.0, 0.0, -1.9), look_at=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), focal_length=0.7, height=1.0, aspect_ratio=-1.5,
                   aperture_width=0.04, aperture_height=0.03, sampler=2, hex_ratio=0.5, focal_plane_dist=2.0)
        bad, exact, want, got = _sample_parity(ref, b200, scenes.mixed_materials(), cam, 96, 64, 20000, seed=25)
        assert bad == 0 and exact == 1.0
    finally:
        b200.set_fast_queries(True, True, True)


def test_unit_virtual_methods_through_the_cpp_api(ref, b200):
    """Object::getSurfaceNormal / sampleSurface and BSDF::propagateRay / getSpectrum called one element at a time through
    the public virtual interface (the host layer answers them with ptb_prim_normal / ptb_prim_sample /
    ptb_bsdf_propagate / ptb_bsdf_spectrum): values and engine consumption equal the unmodified reference's."""
    import unit_cases as uc

    prims, x = uc.prims(), uc.inputs()
    n = 300  # one GPU round trip per element through this API
    builders = []
    for lib in (ref, b200):
        builder = lib.builder()
        handles = [builder.material(**spec) for spec in uc.MATERIALS]
        builders.append((builder, handles, uc.add_to_builder(builder, prims)))
    (br, hr, ir), (bg, hg, ig) = builders
    for obj_r, obj_g in zip(ir, ig):
        assert np.array_equal(bg.object_normal(obj_g, x["positions"][:n]), br.object_normal(obj_r, x["positions"][:n]), equal_nan=True)
        got, next_got = bg.object_sample(obj_g, x["seeds"][:n])
        want, next_want = br.object_sample(obj_r, x["seeds"][:n])
        assert np.array_equal(got, want) and np.array_equal(next_got, next_want)
    for mat_r, mat_g in zip(hr, hg):
        got, next_got = bg.bsdf_propagate(mat_g, 1e-3, x["propagate"][:n], x["seeds"][:n])
        want, next_want = br.bsdf_propagate(mat_r, 1e-3, x["propagate"][:n], x["seeds"][:n])
        assert np.array_equal(got, want, equal_nan=True) and np.array_equal(next_got, next_want)
        for synthetic in (False, True):
            assert np.array_equal(bg.bsdf_spectrum(mat_g, synthetic, x["spectrum"][:n]), br.bsdf_spectrum(mat_r, synthetic, x["spectrum"][:n]))
    br.close()
    bg.close()


def test_progress_callbacks_fire_while_the_frame_renders(b200):
    """worker.cpp:354-360 fires progress_callback as tiles complete.  Here the render core reports retired samples between
    batches of bounce iterations (and between pixel groups when the per-sample buffer is small) and processJob turns them
    into per-tile callbacks, in order; half of them must have fired well before the call returned."""
    import os

    spec = scenes.cornell_demo(("obj", scenes.standin_obj(120, 80)))
    scene = spec.build(b200)
    w, h, spp = 640, 360, 256  # with the 2 Mi-path pool below: ~150 bounce iterations = ~19 batches of 8
    camera = scenes.demo_camera(b200, w, h)
    scene.process_job(camera, w, h, spp, spp, 1e-3)  # warm-up at the same size: the workspace (pool, per-sample buffer) is allocated here, not in the timed call
    for budget_mb in ("0", "64"):  # one pixel group / several pixel groups
        os.environ["PTB_SAMPLE_BUFFER_MB"] = budget_mb
        os.environ["PTB_POOL_PATHS"] = str(1 << 21)  # a pool much smaller than the frame: many batches of bounce iterations, as in a long render
        try:
            image, info = scene.process_job(camera, w, h, spp, spp, 1e-3)
        finally:
            os.environ.pop("PTB_SAMPLE_BUFFER_MB", None)
            os.environ.pop("PTB_POOL_PATHS", None)
        assert info["callbacks"] == info["total_tiles"] == 20 * 12 and info["monotonic"]
        assert 0 <= info["first_callback_s"] < 0.6 * info["seconds"], info
        assert info["half_callbacks_s"] < 0.9 * info["seconds"], info
        assert image[..., 3].max() == 1.0


def test_process_job_on_several_gpus_equals_one_gpu(b200):
    """One call -> the whole image on every GPU of the process (ptb::RenderControl::devices, ptb_render_multi: a host
    thread and a scene replica per GPU, interleaved tiles, one gather kernel reading the peers' tiles over NVLink): the
    frame must equal the one-GPU frame bit for bit.  Needs two devices."""
    if b200.set_devices(2) < 2:
        b200.set_devices(1)
        pytest.skip("needs two CUDA devices")
    try:
        spec = scenes.cornell_demo(("obj", scenes.standin_obj(60, 40)))
        scene = spec.build(b200)
        w, h = 200, 136
        camera = scenes.demo_camera(b200, w, h)
        b200.set_sharding(0, 1, 777)
        multi, info = scene.process_job(camera, w, h, 16, 16, 1e-3)
        assert info["callbacks"] == info["total_tiles"] and info["monotonic"]
        b200.set_devices(1)
        single, _ = scene.process_job(camera, w, h, 16, 16, 1e-3)
        assert np.array_equal(multi, single)
        adaptive_single, _ = scene.process_job(camera, w, h, 5, 12, 1e-3)
        b200.set_devices(2)
        adaptive_multi, _ = scene.process_job(camera, w, h, 5, 12, 1e-3)
        assert np.array_equal(adaptive_multi, adaptive_single)
    finally:
        b200.set_devices(1)
        b200.set_sharding(0, 1, 0)


def test_process_job_shards_sum_to_the_frame(b200):
    """Multi-GPU through the C++ API, one process per GPU: with ptb::RenderControl::shard_index / shard_count (env
    PTB_SHARD_INDEX / PTB_SHARD_COUNT) processJob renders its interleaved share of the reference's tile grid and leaves
    the rest 0, so the sum of the shards (what the NCCL reduce computes) is the unsharded frame, bit for bit; processItem
    is not affected."""
    cam = scenes.demo_camera(None, 96, 80)
    sg = scenes.cornell_demo(("obj", scenes.standin_obj(40, 30))).build(b200)
    camera = b200.camera(**cam)
    try:
        b200.set_sharding(0, 1, 77)
        full, info = sg.process_job(camera, 96, 80, 4, 4, 1e-3)
        total = np.zeros_like(full)
        for rank in range(3):
            b200.set_sharding(rank, 3, 77)
            part, _ = sg.process_job(camera, 96, 80, 4, 4, 1e-3)
            assert (part != 0).any() and (part == 0).all(axis=2).mean() > 0.5
            total += part
            tile = sg.process_item(camera, 96, 80, 4, 4, 1e-3, (8, 8, 16, 16), 5)
            assert (tile[..., 3] > 0).any()  # a tile rendered on request is never somebody else's
        assert np.array_equal(total, full)
    finally:
        b200.set_sharding(0, 1, 0)


def test_render_kats(ref, b200):
    """reference test/render_test.cpp: empty scene -> (0,0,0,0); lit sphere: corner exactly 0, centre alpha > 0."""
    cam = dict(origin=(0.0, 0.0, 0.0), look_at=(0.0, 0.0, 1.0), up=(0.0, 1.0, 0.0), focal_length=1.0, height=1.0, aspect_ratio=1.0)
    sg = scenes.SceneSpec().build(b200)
    image, info = sg.process_job(b200.camera(**cam), 1, 1, 1, 1, 1e-3)
    assert (image == 0).all() and info["callbacks"] == info["total_tiles"] == 1

    cam["focal_length"] = 0.1
    sg = scenes.simple_render().build(b200)
    image, info = sg.process_job(b200.camera(**cam), 16, 16, 2, 2, 1e-3)
    assert (image[0, 0] == 0).all() and image[8, 8, 3] > 0
    assert info["monotonic"] and info["callbacks"] == info["total_tiles"] == 16

    cam = dict(origin=(0.0, 0.0, 0.0), look_at=(0.0, 0.0, 1.0), up=(0.0, 1.0, 0.0), focal_length=0.2, height=0.5, aspect_ratio=1.94)
    sg = scenes.advanced_render().build(b200)
    image, info = sg.process_job(b200.camera(**cam), 132, 68, 5, 10, 1e-3)
    assert (image[0, 0] == 0).all() and image[32, 64, 3] > 0
    assert info["total_tiles"] == 8 * 4 and info["callbacks"] == 32 and info["monotonic"]  # tile size clamp(min(132, 68) / 4, 1, 32) = 17
    # zero-sized image returns an empty image (worker.cpp:390-396)
    image, info = sg.process_job(b200.camera(**cam), 0, 7, 1, 1, 1e-3)
    assert image.size == 0 and info["callbacks"] == 0


def test_image_statistics_match_reference(ref, b200):
    """Image level: production RNG vs the reference at equal spp.  The two renders use unrelated random numbers, so
    the comparison is statistical: the difference of the image means must be within the Monte-Carlo noise estimated
    from two independent reference renders, and trimmed per-pixel errors must be comparable."""
    w = h = 48
    spp = 64
    spec = scenes.cornell_demo(("obj", scenes.standin_obj(60, 40)))
    sr, sg = _pair(spec, ref, b200)
    cr, cg = scenes.demo_camera(ref, w, h), scenes.demo_camera(b200, w, h)
    a = sr.process_item(cr, w, h, spp, spp, 1e-3, (0, 0, w, h), 101)
    b = sr.process_item(cr, w, h, spp, spp, 1e-3, (0, 0, w, h), 202)
    g = sg.process_item(cg, w, h, spp, spp, 1e-3, (0, 0, w, h), 303)
    assert np.array_equal(a[..., 3] > 0, g[..., 3] > 0) or np.mean((a[..., 3] > 0) != (g[..., 3] > 0)) < 0.02

    def trimmed_rmse(x, y):
        e = np.sort(((x[..., :3] - y[..., :3]) ** 2).sum(axis=-1).ravel())
        return np.sqrt(e[: int(0.98 * len(e))].mean())

    noise = trimmed_rmse(a, b)
    ours = trimmed_rmse(a, g)
    print(f"trimmed RMSE ref-vs-ref {noise:.5f}, ref-vs-gpu {ours:.5f}; means {a[..., :3].mean():.5f} {b[..., :3].mean():.5f} {g[..., :3].mean():.5f}")
    assert ours <= 1.35 * noise + 1e-4
    med = [np.median(x[..., :3]) for x in (a, b, g)]
    assert abs(med[2] - med[0]) <= 3 * abs(med[1] - med[0]) + 0.01


def test_bench_scene_image_rmse_within_noise(ref, b200):
    """north_star: "the converged image's RMSE against the reference render falls within the Monte Carlo noise bound at
    equal spp on Cornell + dragon".  The bench's own configuration -- Cornell box + glass stand-in mesh, demo camera,
    production generator, production-math kernels, certified (relaxed) closest hits, any-hit shadows, null shadows
    skipped -- at 256 x 144 and 64 spp against TWO independent renders of the unmodified reference (processJob on all
    host cores), which give the noise bound empirically.  Checked: trimmed RMSE, the clipped image mean, and the means
    of 16 x 9 pixel blocks (z-scores against the reference pair's own scatter) -- a bias of a few percent in any region
    of the image would show."""
    import os

    w, h, spp = 256, 144, 64
    verts, normals = scenes.standin_triangles(400, 200, scenes.DEMO_DRAGON_TRANSFORM)
    spec = scenes.cornell_demo(("triangles", verts, normals))
    sr, sg = _pair(spec, ref, b200)
    cr, cg = scenes.demo_camera(ref, w, h), scenes.demo_camera(b200, w, h)
    cores = os.cpu_count() or 1
    a, _ = sr.process_job(cr, w, h, spp, spp, 1e-3, cores)
    b, _ = sr.process_job(cr, w, h, spp, spp, 1e-3, cores)
    b200.set_sharding(0, 1, 424242)
    try:
        g, _ = sg.process_job(cg, w, h, spp, spp, 1e-3)
        b200.set_sharding(0, 1, 434343)
        g2, _ = sg.process_job(cg, w, h, spp, spp, 1e-3)
    finally:
        b200.set_sharding(0, 1, 0)
    assert np.mean((a[..., 3] > 0) != (g[..., 3] > 0)) < 0.01

    def trimmed_rmse(x, y):
        e = np.sort(((x[..., :3] - y[..., :3]) ** 2).sum(axis=-1).ravel())
        return np.sqrt(e[: int(0.98 * len(e))].mean())

    noise = trimmed_rmse(a, b)
    ours = trimmed_rmse(a, g)
    ours_pair = trimmed_rmse(g, g2)
    report = [f"trimmed RMSE ref-vs-ref {noise:.5f}, ref-vs-gpu {ours:.5f}, gpu-vs-gpu {ours_pair:.5f}"]
    worst = 0.0
    for level in (4.0, 0.5):  # fireflies clipped identically on both sides, at two levels
        clip = lambda x: np.minimum(x[..., :3], level).mean(axis=-1)
        ca, cb, c1, c2 = clip(a), clip(b), clip(g), clip(g2)
        # per pixel, (a - b) scatters with sqrt(2) sigma_pixel and the difference of two pair-means with sigma_pixel:
        # a mean over n pixels of that difference scatters with std(a - b) / sqrt(2 n)
        sigma = np.std(ca - cb) / np.sqrt(2.0 * ca.size)
        diff = 0.5 * (c1.mean() + c2.mean()) - 0.5 * (ca.mean() + cb.mean())
        tiles = lambda x: x.reshape(4, h // 4, 4, w // 4).transpose(0, 2, 1, 3).reshape(16, -1)
        block_sigma = tiles(ca - cb).std(axis=-1) / np.sqrt(2.0 * (ca.size // 16)) + 1e-5
        z = (0.5 * (tiles(c1) + tiles(c2)).mean(axis=-1) - 0.5 * (tiles(ca) + tiles(cb)).mean(axis=-1)) / block_sigma
        k = int(np.abs(z).argmax())
        report.append(f"clip {level}: mean ref {0.5 * (ca.mean() + cb.mean()):.5f} gpu {0.5 * (c1.mean() + c2.mean()):.5f} = {diff / sigma:+.2f} sigma; "
                      f"16 image blocks: max |z| {np.abs(z).max():.2f} (block {k}: ref {tiles(ca)[k].mean():.5f} {tiles(cb)[k].mean():.5f} gpu {tiles(c1)[k].mean():.5f} {tiles(c2)[k].mean():.5f})")
        # (block statistics only at the low clip level: single fireflies dominate a block's scatter estimate at the high one)
        worst = max(worst, abs(diff) / sigma, np.abs(z).max() / 1.2 if level < 1.0 else 0.0)
    print("; ".join(report))
    assert ours <= 1.10 * noise + 1e-4 and ours_pair <= 1.10 * noise + 1e-4
    # image mean within 4 sigma at both clip levels, every sixteenth of the image within 4.8 sigma
    assert worst <= 4.0, report


def test_axis_aligned_and_boundary_rays(ref, b200):
    """Edge cases of the slab test and of Moeller-Trumbore: directions with exact zero components (the 1/d -> FLT_MAX
    path, bounding_box.cpp:43-45), origins exactly on box planes and on triangle planes, rays along shared edges and
    through vertices, rays starting inside spheres (near root only, object.cpp:77-83), back-face-culled hits."""
    spec = scenes.mixed_materials(seed=5, n_tris=180)
    builder = spec.replay(ref)
    tris = builder.get_triangles()
    builder.close()
    tris = tris[~np.isnan(tris).any(axis=1)]
    rng = np.random.Generator(np.random.PCG64(9))
    rays = []
    axes = np.eye(3, dtype=np.float32)
    for _ in range(6000):
        o = rng.uniform(-2.2, 2.2, 3).astype(np.float32)
        d = axes[rng.integers(0, 3)] * np.float32(rng.choice([-1.0, 1.0]))
        rays.append(np.concatenate([o, d]))
    for _ in range(6000):  # two zero components replaced by one: directions inside a coordinate plane
        o = rng.uniform(-2.0, 2.0, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        d[rng.integers(0, 3)] = 0.0
        d = d * (np.float32(1.0) / np.sqrt(np.float32((d * d).sum())))
        rays.append(np.concatenate([o, d]))
    for t in tris[rng.permutation(len(tris))[:3000]]:  # origins on a vertex / edge midpoint / centroid, towards another
        a, b, c = t[0:3], t[3:6], t[6:9]
        start = [a, 0.5 * (a + b), (a + b + c) / 3][rng.integers(0, 3)].astype(np.float32)
        other = tris[rng.integers(0, len(tris))]
        target = [other[0:3], 0.5 * (other[3:6] + other[6:9])][rng.integers(0, 2)].astype(np.float32)
        d = target - start
        n2 = np.float32((d * d).sum())
        if n2 > 0:
            rays.append(np.concatenate([start, d * (np.float32(1.0) / np.sqrt(n2))]))
    for _ in range(3000):  # on the walls of the enclosing box (x, y or z = +-2 exactly)
        o = rng.uniform(-2.0, 2.0, 3).astype(np.float32)
        o[rng.integers(0, 3)] = np.float32(rng.choice([-2.0, 2.0]))
        d = rng.normal(size=3).astype(np.float32)
        d = d * (np.float32(1.0) / np.sqrt(np.float32((d * d).sum())))
        rays.append(np.concatenate([o, d]))
    for _ in range(3000):  # inside the spheres
        centre, radius = [((-1.2, -1.2, 0.8), 0.35), ((0.9, -1.3, 0.2), 0.6), ((-0.2, 0.4, 0.9), 0.45)][rng.integers(0, 3)]
        o = (np.array(centre) + rng.uniform(-0.5, 0.5, 3) * radius).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        d = d * (np.float32(1.0) / np.sqrt(np.float32((d * d).sum())))
        rays.append(np.concatenate([o, d]))
    rays = np.array(rays, np.float32)
    sr, sg = _pair(spec, ref, b200)
    _compare_hits(sr, sg, rays)


def test_large_sample_batch_is_bit_exact(ref, b200):
    """250 k samples on the demo scene (more than the wavefront's refill granularity, fewer than one pool)."""
    cam = scenes.demo_camera(None, 160, 90)
    bad, exact, want, got = _sample_parity(ref, b200, scenes.cornell_demo(("obj", scenes.standin_obj(90, 60))), cam, 160, 90, 250000, seed=77)
    print(f"250k: diverged {bad:.6f}, bit-exact {exact:.6f}")
    assert bad == 0.0 and exact == 1.0


def test_post_processing_is_bit_exact(ref, b200, ctx):
    """toneMap / gammaCorrect / postProcess run on the device behind the reference's API (row f3) and must reproduce the
    reference's values bit for bit (reference test/post_processing_test.cpp only checks dimensions and gamma 1.0)."""
    from cpupathtrace_b200 import capi

    rng = np.random.Generator(np.random.PCG64(1234))
    for (h, w) in [(37, 53), (1, 1), (3, 300), (270, 480)]:
        image = rng.gamma(0.6, 0.4, size=(h, w, 4)).astype(np.float32)
        image[..., 3] = (rng.uniform(size=(h, w)) > 0.1).astype(np.float32)
        if h > 8:
            image[5:9, 7:20] = 0.0  # black pixels: gamma gives 0 ** negative = inf, inf * 0 = NaN on both sides
        for mode, gamma in ((0, 1.8), (1, 1.8), (1, 2.2), (1, 1.0), (2, 1.8)):
            want = ref.post_process(mode, image, gamma)
            got = b200.post_process(mode, image, gamma)
            assert np.array_equal(want, got, equal_nan=True), (h, w, mode, gamma)
            assert np.array_equal(got[..., 3], image[..., 3])
            assert np.array_equal(ctx.post_process(image, mode, gamma), want, equal_nan=True)
    # a rendered frame, processed where it lives
    import torch

    frame = torch.from_numpy(rng.gamma(0.5, 0.2, size=(64, 96, 4)).astype(np.float32)).cuda()
    host = frame.cpu().numpy()
    ctx.post_process_device(frame.data_ptr(), 96, 64, 2, 1.8)
    torch.cuda.synchronize()
    assert np.array_equal(frame.cpu().numpy(), ref.post_process(2, host), equal_nan=True)
