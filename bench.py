#!/usr/bin/env python
"""bench.py — Msamples/s and Mrays/s of the render hot path on BASELINE.json's configs[1]:
Cornell box + dragon at 1920x1080, 256 spp, on N B200 (one process per GPU).

The XYZ RGB dragon asset is absent from the reference checkout, so the scene uses the deterministic stand-in mesh of
SURVEY.md section 8d ("stand-in-1M": 1000 x 500 bumpy torus, 1 M triangles, the demo's glass material and transform).

One step = one full render of the frame (every pixel, every sample, resolve included):
  value   device-resident: ptb_render writes the image into HBM; tiles are interleaved over the ranks and, for N > 1,
          one image-sized NCCL reduce to rank 0 follows (scene + BVH replicated per GPU).  Timed with CUDA events on
          the launching streams (ptb_render_stats.device_ms_total + torch events around the reduce), max over ranks.
  e2e     the same render through the reference-facing API with HOST buffers: processJob (N = 1) resp. the C-ABI with
          a host result (N > 1); the device->host copy of the image is inside the timed region.
  roofline  traversal kernels only: algorithmic bytes per ray (64 B per inner record fetched + 48 B per primitive
          fetched + 48 B ray/hit record, fetch counts measured by a counting pass of the same traversal) x rays traced
          / summed CUDA-event time of the trace kernels, against the measured HBM copy bandwidth.
  cpu_baseline  the reference's own multithreaded CPU path (oracle/_ref timing build) on this host, bounded sample.

`--impl reference` prints the CPU reference arm alone (same metric, config and units).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO_ROOT = os.path.dirname(os.path.abspath(__file__))
if REPO_ROOT not in sys.path:
    sys.path.insert(0, REPO_ROOT)

METRIC = "Msamples/s"
RAY_RECORD_BYTES = 48  # 32 B ray read + 16 B hit written
INNER_BYTES = 64
LEAF_BYTES = 48


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--width", type=int, default=1920)
    p.add_argument("--height", type=int, default=1080)
    p.add_argument("--spp", type=int, default=256)
    p.add_argument("--mesh", default="1000x500", help="stand-in mesh grid nu x nv (2 triangles per cell); 'none' = Cornell only")
    p.add_argument("--max-depth", type=int, default=0, help="0 = unlimited, as the reference")
    p.add_argument("--reference-closest", action="store_true",
                   help="closest-hit queries walk the reference-topology tree only (default: certified SAH walk + re-trace of uncertified rays)")
    p.add_argument("--reference-shadows", action="store_true",
                   help="trace shadow rays exactly like the reference (closest-hit queries, also for glass/mirror vertices whose result is discarded)")
    p.add_argument("--guarded", action="store_true",
                   help="certified closest hits with the guard table (the exact mode validation uses); default: relaxed, as production renders")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of one reference-arm sample")
    return p.parse_args()


def mesh_arg(args):
    if args.mesh == "none":
        return None, "cornell-only"
    nu, nv = (int(v) for v in args.mesh.lower().split("x"))
    return (nu, nv), f"stand-in-{2 * nu * nv}"


def build_spec(args):
    from cpupathtrace_b200 import scenes

    grid, label = mesh_arg(args)
    mesh = None
    if grid is not None:
        verts, normals = scenes.standin_triangles(grid[0], grid[1], scenes.DEMO_DRAGON_TRANSFORM)
        mesh = ("triangles", verts, normals)
    return scenes.cornell_demo(mesh), label


def config_dict(args, label, n_gpus):
    return {
        "workload": f"Cornell box + {label} glass mesh (xyzrgb_dragon.obj is absent from the reference checkout), "
                    f"{args.width}x{args.height}, {args.spp} spp (min=max), demo camera (thin lens), eps 1e-3",
        "image": [args.width, args.height],
        "spp": args.spp,
        "max_depth": args.max_depth,
        "scene": label,
        "parallelism": f"interleaved 32x32 tiles over {n_gpus} GPU(s), scene+BVH replicated, one NCCL image reduce" if n_gpus > 1 else "1 GPU",
        "l2": "working set (path pool of up to 128 Mi paths ~30 GB + per-sample buffer, BVH ~230 MB) exceeds the 126 MB L2; a 512 MB buffer is also written between steps",
    }


# ------------------------------------------------------------------------------------------------ clocks


class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device_index):
        self.device_index = device_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.strip().split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, sm_max, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            if len(row) < 7:
                continue
            try:
                sm.append(float(row[0]))
                sm_max.append(float(row[1]))
                power.append(float(row[2]))
            except ValueError:
                continue
            for name, value in zip(names, row[3:7]):
                if value.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(sm_max)) if sm_max else None,
            "power_w_max": float(max(power)) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------------------------ CPU reference arm


def reference_sample(args, spec, label, steps, warmup):
    """Times the unmodified reference (oracle/_ref timing build) with all host threads on a bounded sample of the
    workload: the full frame at a reduced spp chosen so that one step takes about --cpu-seconds."""
    from cpupathtrace_b200 import pth, scenes

    ref = pth.load_reference(fast=True)
    cores = os.cpu_count() or 1
    scene = spec.build(ref)
    camera = scenes.demo_camera(ref, args.width, args.height)

    # calibrate on a quarter-resolution 1 spp pass of the same scene and camera
    cw, ch = max(args.width // 4, 16), max(args.height // 4, 16)
    cal_camera = scenes.demo_camera(ref, cw, ch)
    t0 = time.perf_counter()
    scene.process_job(cal_camera, cw, ch, 1, 1, 1e-3, cores)
    per_sample = (time.perf_counter() - t0) / (cw * ch)
    spp = int(max(1, min(args.spp, round(args.cpu_seconds / max(per_sample * args.width * args.height, 1e-9)))))

    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        scene.process_job(camera, args.width, args.height, spp, spp, 1e-3, cores)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    samples = args.width * args.height * spp
    seconds = float(np.mean(times))
    return {
        "value": samples / seconds / 1e6,
        "unit": METRIC,
        "cores": cores,
        "kind": "reference",
        "sample": f"full {args.width}x{args.height} frame of the same scene at {spp} spp (min=max) instead of {args.spp}; processJob with worker_count={cores}; "
                  f"reference sources built -O3 -march=x86-64-v3; {len(times)} timed pass(es), {seconds:.2f} s each",
        "seconds_per_step": seconds,
        "spp_sampled": spp,
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec, label = build_spec(args)
    steps, warmup = max(args.steps, 1), max(min(args.warmup, 1), 0)
    base = reference_sample(args, spec, label, steps, warmup)
    line = {
        "metric": METRIC,
        "value": base["value"],
        "unit": METRIC,
        "impl": "reference",
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": warmup,
        "ms_per_step": base["seconds_per_step"] * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": config_dict(args, label, args.gpus),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm


def run_b200_arm(args):
    import ctypes as C

    import torch

    from cpupathtrace_b200 import capi, pth, scenes, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    os.environ["PTB_DEVICE"] = str(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist = None

    spec, label = build_spec(args)
    b200 = pth.load_b200()
    t_build = time.perf_counter()
    scene_cpp = spec.build(b200)  # Scene::Scene through the public C++ API: lowering + BVH + upload
    t_build = time.perf_counter() - t_build
    handle = scene_cpp.device_handle()
    camera_cpp = scenes.demo_camera(b200, args.width, args.height)
    kw = scenes.demo_camera(None, args.width, args.height)
    camera = capi.camera_init(kw["origin"], kw["look_at"], kw["up"], kw["focal_length"], kw["height"], kw["aspect_ratio"], kw["aperture_width"],
                              kw["aperture_height"], kw["sampler"], 0.0, kw["focal_plane_dist"])
    lib = capi.load()
    info = capi.SceneInfo()
    capi.check(lib.ptb_scene_get_info(handle, C.byref(info)))

    flags = 0 if args.reference_shadows else (capi.PTB_FLAG_ANY_HIT_SHADOWS | capi.PTB_FLAG_SKIP_NULL_SHADOWS)
    if not args.reference_closest:
        flags |= capi.PTB_FLAG_CERTIFIED_CLOSEST | (0 if args.guarded else capi.PTB_FLAG_CERTIFIED_RELAXED)
    # the same options for the C++ API (processJob): through ptb::RenderControl, not the environment (which is read once)
    b200.set_fast_queries(not args.reference_closest, not args.reference_shadows, not args.reference_shadows)
    b200.set_render_control(max_depth=args.max_depth, relaxed_guard=not args.guarded)

    def opts(spp, extra_flags=0, seed=1):
        return capi.render_opts(args.width, args.height, spp, spp, 1e-3, args.max_depth, capi.PTB_RNG_COUNTER, flags | extra_flags, seed, 0, rank, world)

    image = torch.zeros(args.height, args.width, 4, device="cuda", dtype=torch.float32)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    host_image = torch.empty(args.height, args.width, 4, dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_image():
        """NCCL sum-reduce of the per-rank images (disjoint tiles, zeros elsewhere) to rank 0; returns device ms."""
        if dist is None:
            return 0.0
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        sharding.reduce_image(image, dist, dst=0)
        stop.record()
        stop.synchronize()
        return start.elapsed_time(stop)

    rank_detail = {"render_ms": [], "reduce_ms": []}

    def device_step(seed):
        stats = capi.RenderStats()
        o = opts(args.spp, capi.PTB_FLAG_DEVICE_IO, seed)
        capi.check(lib.ptb_render(handle, C.byref(camera), C.byref(o), 0, 0, args.width, args.height, C.c_void_p(image.data_ptr()), C.byref(stats)))
        reduce_ms = reduce_image()
        rank_detail["render_ms"].append(stats.device_ms_total)
        rank_detail["reduce_ms"].append(reduce_ms)
        ms = stats.device_ms_total + reduce_ms
        return ms, stats

    def max_over_ranks(value):
        if dist is None:
            return value
        t = torch.tensor([value], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(value):
        if dist is None:
            return value
        t = torch.tensor([value], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- warm-up.  The clock sampler (one nvidia-smi process polling every 500 ms) is started here rather than at the
    # first timed step: its NVML start-up takes driver locks for a few hundred milliseconds, which showed up as idle gaps
    # between launches of the first timed step; the samples it reports cover the warm-up and the timed steps, all under
    # the same load.
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        flush.zero_()
        barrier()
        device_step(1000 + i)

    # ---- timed steps, device-resident
    step_ms, wall_ms = [], []
    rank_detail["render_ms"].clear()
    rank_detail["reduce_ms"].clear()
    totals = {"samples": 0, "closest": 0, "shadow": 0, "skipped": 0, "vertices": 0, "launches": 0, "retraced": 0, "trace_ms": 0.0, "shade_ms": 0.0, "iterations": 0,
              "shadow_ms": 0.0}
    for i in range(args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        ms, stats = device_step(2000 + i)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        step_ms.append(max_over_ranks(ms))
        wall_ms.append(max_over_ranks(wall))
        totals["samples"] += stats.samples
        totals["closest"] += stats.closest_rays
        totals["shadow"] += stats.shadow_rays
        totals["skipped"] += stats.shadow_rays_skipped
        totals["retraced"] += stats.closest_rays_retraced
        totals["vertices"] += stats.path_vertices
        totals["launches"] += stats.kernel_launches + (1 if dist is not None else 0)
        totals["trace_ms"] += stats.device_ms_trace
        totals["shade_ms"] += stats.device_ms_shade
        totals["shadow_ms"] += stats.device_ms_trace_shadow
        totals["iterations"] += stats.bounce_iterations
    clocks = sampler.stop()
    # this rank's own render time (its tiles) and the time it spent in the image reduce, which includes waiting for the
    # slowest rank; gathered so that imbalance between ranks is visible in the report
    per_rank = None
    if dist is not None:
        mine = torch.tensor([float(np.mean(rank_detail["render_ms"])), float(np.mean(rank_detail["reduce_ms"]))], dtype=torch.float64, device="cuda")
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        per_rank = {"render_ms": [round(float(g[0]), 2) for g in gathered], "reduce_ms_incl_wait": [round(float(g[1]), 2) for g in gathered]}

    # ---- counting pass (untimed, after the timed steps so that a profiler attached to this command meets steady-state
    # launches first): inner / leaf fetches per ray of the same traversal at reduced spp
    count_spp = max(1, min(args.spp, 2))
    cstats = capi.RenderStats()
    co = opts(count_spp, capi.PTB_FLAG_DEVICE_IO | capi.PTB_FLAG_COUNT_VISITS, 99)
    capi.check(lib.ptb_render(handle, C.byref(camera), C.byref(co), 0, 0, args.width, args.height, C.c_void_p(image.data_ptr()), C.byref(cstats)))
    count_detail = {
        "closest_inner_per_ray": (cstats.inner_visits - cstats.shadow_inner_visits) / max(cstats.closest_rays, 1),
        "closest_leaf_per_ray": (cstats.leaf_visits - cstats.shadow_leaf_visits) / max(cstats.closest_rays, 1),
        "shadow_inner_per_ray": cstats.shadow_inner_visits / max(cstats.shadow_rays, 1),
        "shadow_leaf_per_ray": cstats.shadow_leaf_visits / max(cstats.shadow_rays, 1),
    }
    certificate_audit = {"primitive_tests": int(cstats.leaf_visits - cstats.shadow_leaf_visits), "suspect_hits": int(cstats.certified_suspect_hits),
                         "closest_rays": int(cstats.closest_rays), "retraced": int(cstats.closest_rays_retraced)}
    count_rays = sum_over_ranks(float(cstats.closest_rays + cstats.shadow_rays))
    inner_per_ray = sum_over_ranks(float(cstats.inner_visits)) / max(count_rays, 1.0)
    leaf_per_ray = sum_over_ranks(float(cstats.leaf_visits)) / max(count_rays, 1.0)
    bytes_per_ray = INNER_BYTES * inner_per_ray + LEAF_BYTES * leaf_per_ray + RAY_RECORD_BYTES

    job_samples = sum_over_ranks(float(totals["samples"]))
    job_rays = sum_over_ranks(float(totals["closest"] + totals["shadow"]))
    job_skipped = sum_over_ranks(float(totals["skipped"]))
    total_s = sum(step_ms) / 1e3
    value = job_samples / total_s / 1e6
    mrays = job_rays / total_s / 1e6

    # roofline of the traversal kernels on this rank (every rank runs the same kernels on its own tiles).  The dominant
    # kernel is the closest-hit trace (certified SAH walk + its re-trace launch, timed together); `achieved` is its
    # algorithmic bytes per ray x the rays it traced / its CUDA-event time.  The north star's figure over ALL rays
    # (closest + shadow kernels) is reported next to it as frac_all_rays.
    rank_rays = float(totals["closest"] + totals["shadow"])
    trace_s = totals["trace_ms"] / 1e3
    achieved_all = bytes_per_ray * rank_rays / max(trace_s, 1e-12) / 1e9
    closest_s = (totals["trace_ms"] - totals["shadow_ms"]) / 1e3
    closest_bytes_per_ray = INNER_BYTES * count_detail["closest_inner_per_ray"] + LEAF_BYTES * count_detail["closest_leaf_per_ray"] + RAY_RECORD_BYTES
    achieved = closest_bytes_per_ray * float(totals["closest"]) / max(closest_s, 1e-12) / 1e9
    closest_launches = max(int(totals["iterations"]), 1)
    peak, peak_source = 6650.0, "fallback"
    peaks_path = os.path.join(REPO_ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        try:
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_source = "measured"
        except (KeyError, ValueError):
            pass
    traffic = None
    traffic_note = None
    profile_path = os.path.join(REPO_ROOT, "profiles", "traffic.json")
    if os.path.exists(profile_path):
        try:
            profile = json.load(open(profile_path))
            traffic = profile.get("dram_bytes_per_launch")
            traffic_note = profile.get("source")
        except ValueError:
            traffic = None

    # ---- e2e through the reference-facing API with host buffers
    e2e_ms = []
    for i in range(args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            b200.set_sharding(0, 1, 3000 + i)
            scene_cpp.process_job(camera_cpp, args.width, args.height, args.spp, args.spp, 1e-3, 0)
        else:
            device_step(3000 + i)
            if rank == 0:
                host_image.copy_(image, non_blocking=False)
        torch.cuda.synchronize()
        e2e_ms.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
        barrier()
    e2e_value = (job_samples / args.steps) / (np.mean(e2e_ms) / 1e3) / 1e6

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and os.path.exists(pth.REF_FAST):
        base = reference_sample(args, spec, label, 1, 0)
        cpu_baseline = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": METRIC,
        "value": value,
        "unit": METRIC,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": float(np.mean(step_ms)),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": config_dict(args, label, world),
        "mrays_per_s": mrays,
        "rays_per_sample": job_rays / max(job_samples, 1.0),
        "ray_convention": "rays actually traced on the GPU (closest-hit + shadow)" + (
            "" if args.reference_shadows else f"; {job_skipped / max(job_samples, 1.0):.2f} zero-weight shadow rays per sample that the reference traces are skipped, shadow rays are any-hit"),
        "wall_ms_per_step": float(np.mean(wall_ms)),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": C.sizeof(capi.Camera) + C.sizeof(capi.RenderOpts),
                "d2h_bytes_per_step": args.width * args.height * 16, "ms_per_step": float(np.mean(e2e_ms)),
                "api": "processJob (C++ host API via harness)" if world == 1 else "ptb_render + NCCL reduce + D2H"},
        "gpu_launches": int(totals["launches"]),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_note, "peak_source": f"{peak_source} HBM copy bandwidth",
                     "kernel": "traceClosestKernel" + ("<reference tree>" if args.reference_closest else "<certified SAH walk> + re-trace launch"),
                     "bytes_per_ray": closest_bytes_per_ray, "rays_per_launch": float(totals["closest"]) / closest_launches,
                     "algorithmic_bytes_per_launch": closest_bytes_per_ray * float(totals["closest"]) / closest_launches,
                     "ms_per_launch": closest_s * 1e3 / closest_launches,
                     "frac_all_rays": achieved_all / peak, "achieved_all_rays": achieved_all, "bytes_per_ray_all_rays": bytes_per_ray,
                     "inner_fetches_per_ray": inner_per_ray, "leaf_fetches_per_ray": leaf_per_ray,
                     "trace_ms_per_step": totals["trace_ms"] / args.steps, "shade_ms_per_step": totals["shade_ms"] / args.steps,
                     "trace_share_of_step": totals["trace_ms"] / max(sum(step_ms), 1e-9), "mrays_per_s_trace_only": rank_rays / max(trace_s, 1e-12) / 1e6,
                     "shadow_trace_ms_per_step": totals["shadow_ms"] / args.steps,
                     "closest_mrays_per_s": totals["closest"] / max(totals["trace_ms"] - totals["shadow_ms"], 1e-9) / 1e3,
                     "shadow_mrays_per_s": totals["shadow"] / max(totals["shadow_ms"], 1e-9) / 1e3, **count_detail},
        "cpu_baseline": cpu_baseline,
        "scene": {"prims": int(info.n_prims), "inner_nodes": int(info.n_inner_nodes), "bvh_depth": int(info.bvh_depth),
                  "device_mb": info.device_bytes / 2**20, "build_s": info.build_seconds, "scene_ctor_s": t_build,
                  "query_tree": "device LBVH" if info.query_tree_on_device else "host binned SAH", "query_tree_device_ms": info.query_tree_device_ms},
        "bounce_iterations_per_step": totals["iterations"] / args.steps,
        "per_rank": per_rank,
        "certificate_audit": None if args.reference_closest else certificate_audit,
        "closest_hit": ("reference-topology tree" if args.reference_closest else
                        f"certified SAH walk; {totals['retraced']} of {totals['closest']} closest-hit rays had no certificate and were re-traced on the reference tree"),
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries exactly one JSON line (rank 0).  Libraries write there too (NCCL prints its version banner to
    # stdout when NCCL_DEBUG=VERSION is set in the environment), so everything else is sent to stderr: file descriptor 1
    # is pointed at stderr for the duration of the run and the JSON line goes to the saved original stdout.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved_stdout, "w")
    global print
    builtin_print = print

    def print(*a, **k):  # noqa: A001 - only the final JSON line is printed in this module
        k.setdefault("file", real_stdout)
        builtin_print(*a, **k)
        real_stdout.flush()

    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
