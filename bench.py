#!/usr/bin/env python
"""bench.py — Msamples/s and Mrays/s of the render hot path on BASELINE.json's configs:

  --gpus 1 (default)   configs[1] "C2": Cornell box + dragon at 1920x1080, 256 spp, 1 B200
  --gpus 2/4/8         configs[3] "C4": the same scene at 1024 spp, max depth 16, image-tile split over the GPUs
  --config c5          configs[4] "C5": 3840x2160 at 4096 spp (meant for 8 GPUs)
  --workload soup      configs[2] "C3": random-triangle soup (--tris Mi triangles), 2^24 coherent / incoherent / shadow rays
                       through ptb_intersect / ptb_occluded (metric: Mrays/s)

The XYZ RGB dragon asset is absent from the reference checkout, so the scene uses the deterministic stand-in mesh of
SURVEY.md section 8d ("stand-in-1M": 1000 x 500 bumpy torus, 1 M triangles, the demo's glass material and transform).

One step = one full render of the frame (every pixel, every sample, resolve included):
  value   device-resident: ptb_render writes the image into HBM; tiles are interleaved over the ranks and, for N > 1,
          one image-sized NCCL reduce to rank 0 follows (scene + BVH replicated per GPU).  Timed with CUDA events on
          the launching streams (ptb_render_stats.device_ms_total + torch events around the reduce), max over ranks.
  e2e     the same render through the reference-facing API with HOST buffers: one processJob call per step (the duration
          of the call itself, device->host copy of the image included).  For N > 1 rank 0 makes that ONE call render on
          all N GPUs in-library (ptb::RenderControl::devices -> ptb_render_multi) after the other ranks have left.
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
    p.add_argument("--workload", default="render", choices=["render", "soup"])
    p.add_argument("--config", default="auto", choices=["auto", "c2", "c4", "c5"],
                   help="render workload: auto = c2 on one GPU, c4 on several (BASELINE.json configs[1] / configs[3]); --width/--height/--spp/--max-depth override")
    p.add_argument("--width", type=int, default=None)
    p.add_argument("--height", type=int, default=None)
    p.add_argument("--spp", type=int, default=None)
    p.add_argument("--tris", type=int, default=4, help="soup workload: Mi triangles (1..16)")
    p.add_argument("--rays", default="incoherent", choices=["coherent", "incoherent", "shadow"], help="soup workload: ray set (2^24 rays)")
    p.add_argument("--n-rays", type=int, default=1 << 24)
    p.add_argument("--mesh", default="1000x500", help="stand-in mesh grid nu x nv (2 triangles per cell); 'none' = Cornell only")
    p.add_argument("--max-depth", type=int, default=None, help="0 = unlimited, as the reference (default: the config's)")
    p.add_argument("--reference-closest", action="store_true",
                   help="closest-hit queries walk the reference-topology tree only (default: certified SAH walk + re-trace of uncertified rays)")
    p.add_argument("--reference-shadows", action="store_true",
                   help="trace shadow rays exactly like the reference (closest-hit queries, also for glass/mirror vertices whose result is discarded)")
    p.add_argument("--guarded", action="store_true",
                   help="certified closest hits with the guard table (the exact mode validation uses); default: relaxed, as production renders")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-adaptive-line", action="store_true", help="skip the untimed adaptive-sampling frame (min = spp / 8)")
    p.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of one reference-arm sample")
    args = p.parse_args()
    world = int(os.environ.get("WORLD_SIZE", str(max(args.gpus, 1))))
    name = args.config if args.config != "auto" else ("c2" if world <= 1 else "c4")
    width, height, spp, depth = {"c2": (1920, 1080, 256, 0), "c4": (1920, 1080, 1024, 16), "c5": (3840, 2160, 4096, 16)}[name]
    args.config_name = name
    args.width = args.width if args.width is not None else width
    args.height = args.height if args.height is not None else height
    args.spp = args.spp if args.spp is not None else spp
    args.max_depth = args.max_depth if args.max_depth is not None else depth
    return args


def mesh_arg(args):
    """--mesh: 'none', a stand-in grid 'NUxNV', or the path of an OBJ file (the real xyzrgb_dragon.obj drops in here)."""
    if args.mesh == "none":
        return None, "cornell-only"
    if os.path.isfile(args.mesh):
        return args.mesh, "obj:" + os.path.basename(args.mesh)
    nu, nv = (int(v) for v in args.mesh.lower().split("x"))
    return (nu, nv), f"stand-in-{2 * nu * nv}"


def build_spec(args):
    from cpupathtrace_b200 import scenes

    grid, label = mesh_arg(args)
    mesh = None
    if isinstance(grid, str):
        mesh = ("obj", open(grid, "rb").read())  # io::loadMesh with the demo's transform, cull = false, smooth = true
    elif grid is not None:
        verts, normals = scenes.standin_triangles(grid[0], grid[1], scenes.DEMO_DRAGON_TRANSFORM)
        mesh = ("triangles", verts, normals)
    return scenes.cornell_demo(mesh), label


def config_dict(args, label, n_gpus):
    return {
        "workload": f"BASELINE configs[{ {'c2': 1, 'c4': 3, 'c5': 4}[args.config_name] }]: Cornell box + {label} glass mesh (xyzrgb_dragon.obj is absent from the reference "
                    f"checkout), {args.width}x{args.height}, {args.spp} spp (min=max), max depth {args.max_depth or 'unlimited'}, demo camera (thin lens), eps 1e-3",
        "image": [args.width, args.height],
        "spp": args.spp,
        "max_depth": args.max_depth,
        "scene": label,
        "parallelism": f"interleaved 32x32 tiles over {n_gpus} GPU(s), scene+BVH replicated, one NCCL image reduce" if n_gpus > 1 else "1 GPU",
        "l2": "working set (path pool of up to 256 Mi paths ~63 GB + per-sample buffer, BVH ~230 MB) exceeds the 126 MB L2; a 512 MB buffer is also written between steps",
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
        # the duration of the processJob call itself, taken inside the harness (steady_clock around the call): the same
        # definition as the GPU arm's e2e, without the harness's own copy of the image into numpy
        _, info = scene.process_job(camera, args.width, args.height, spp, spp, 1e-3, cores)
        if i >= warmup:
            times.append(info["seconds"])
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
        "scaling": "strong",
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
    control = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # host-side barrier for the phases in which only rank 0 works (a NCCL barrier would keep a spinning kernel on every
        # waiting GPU, and rank 0's in-library multi-GPU render uses those GPUs)
        control = dist.new_group(backend="gloo")
    else:
        dist = None

    spec, label = build_spec(args)
    b200 = pth.load_b200()
    # CUDA initialisation, context creation and the first load of the setup kernels are paid by a small scene first, so that
    # scene_ctor_s below is the setup time of the bench scene itself
    warm = scenes.cornell_demo(None).build(b200)
    del warm
    t_objects = time.perf_counter()
    builder = spec.replay(b200)  # the caller's side: a million Triangle objects, materials, lights through the public C++ API
    t_objects = time.perf_counter() - t_objects
    t_build_first = time.perf_counter()
    scene_cpp = builder.scene()  # Scene::Scene: lowering of the object graph + ptb_scene_create (boxes, both trees, leaf records on the GPU)
    t_build_first = time.perf_counter() - t_build_first
    builder.close()
    # the first large scene of a process also pays the driver's first large allocations (0.1-0.3 s, varies from box to box):
    # the same scene is set up a second time and both times are reported
    scene_cpp.close()
    builder = spec.replay(b200)
    t_build = time.perf_counter()
    scene_cpp = builder.scene()
    t_build = time.perf_counter() - t_build
    builder.close()
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
    if os.environ.get("PTB_BENCH_NO_CLOCK_SAMPLER", "") == "":  # (experiments only: is the sampler visible in the step times?)
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

    # ---- the kernel times behind `roofline`: the per-launch CUDA events of the timed steps themselves.  Only when the
    # opt-in PTB_STREAMS > 1 splits each frame between concurrent contexts of the GPU (event brackets around co-scheduled
    # launches overlap and do not measure a kernel's own duration) are they taken from the same number of frames rendered
    # on ONE stream right after the timed steps (PTB_FLAG_SINGLE_STREAM; same scene, same options, next seeds).
    split_streams = int(os.environ.get("PTB_STREAMS", "1") or "1") > 1
    roof = {"closest": totals["closest"], "shadow": totals["shadow"], "trace_ms": totals["trace_ms"], "iterations": totals["iterations"],
            "frame_ms": float(sum(rank_detail["render_ms"]))}
    if split_streams:
        roof = {"closest": 0, "shadow": 0, "trace_ms": 0.0, "iterations": 0, "frame_ms": 0.0}
        for i in range(args.steps):
            flush.zero_()
            rstats = capi.RenderStats()
            ro = opts(args.spp, capi.PTB_FLAG_DEVICE_IO | capi.PTB_FLAG_SINGLE_STREAM, 2500 + i)
            capi.check(lib.ptb_render(handle, C.byref(camera), C.byref(ro), 0, 0, args.width, args.height, C.c_void_p(image.data_ptr()), C.byref(rstats)))
            roof["closest"] += rstats.closest_rays
            roof["shadow"] += rstats.shadow_rays
            roof["trace_ms"] += rstats.device_ms_trace
            roof["iterations"] += rstats.bounce_iterations
            roof["frame_ms"] += rstats.device_ms_total
    clocks = sampler.stop()
    # this rank's own render time (its tiles) and the time it spent in the image reduce, which includes waiting for the
    # slowest rank; gathered so that imbalance between ranks is visible in the report
    per_rank = None
    if dist is not None:
        mine = torch.tensor([float(np.mean(rank_detail["render_ms"])), float(np.mean(rank_detail["reduce_ms"]))], dtype=torch.float64, device="cuda")
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        per_rank = {"render_ms": [round(float(g[0]), 2) for g in gathered], "reduce_ms_incl_wait": [round(float(g[1]), 2) for g in gathered]}

    # ---- profile pass (untimed): one more frame with CUDA events around EVERY launch for the per-kernel breakdown (the
    # timed steps time only the closest-hit trace per launch: the ~300 extra event pairs cost 4-5 % of a frame)
    pstats = capi.RenderStats()
    po = opts(args.spp, capi.PTB_FLAG_DEVICE_IO | capi.PTB_FLAG_PROFILE_ALL | capi.PTB_FLAG_SINGLE_STREAM, 98)
    capi.check(lib.ptb_render(handle, C.byref(camera), C.byref(po), 0, 0, args.width, args.height, C.c_void_p(image.data_ptr()), C.byref(pstats)))
    breakdown = {
        "source": "one untimed frame with PTB_FLAG_PROFILE_ALL (events around every launch)",
        "frame_ms": pstats.device_ms_total,
        "closest_trace_ms": pstats.device_ms_trace - pstats.device_ms_trace_shadow,
        "shadow_trace_ms": pstats.device_ms_trace_shadow,
        "generate_shade_accumulate_resolve_ms": pstats.device_ms_shade,
        "shadow_mrays_per_s": pstats.shadow_rays / max(pstats.device_ms_trace_shadow, 1e-9) / 1e3,
    }

    # ---- counting pass (untimed, after the timed steps so that a profiler attached to this command meets steady-state
    # launches first): inner / leaf fetches per ray of the same traversal at reduced spp
    count_spp = max(1, min(args.spp, 2))
    cstats = capi.RenderStats()
    co = opts(count_spp, capi.PTB_FLAG_DEVICE_IO | capi.PTB_FLAG_COUNT_VISITS | capi.PTB_FLAG_SINGLE_STREAM, 99)
    capi.check(lib.ptb_render(handle, C.byref(camera), C.byref(co), 0, 0, args.width, args.height, C.c_void_p(image.data_ptr()), C.byref(cstats)))
    count_detail = {
        "closest_inner_per_ray": (cstats.inner_visits - cstats.shadow_inner_visits) / max(cstats.closest_rays, 1),
        "closest_leaf_per_ray": (cstats.leaf_visits - cstats.shadow_leaf_visits) / max(cstats.closest_rays, 1),
        "shadow_inner_per_ray": cstats.shadow_inner_visits / max(cstats.shadow_rays, 1),
        "shadow_leaf_per_ray": cstats.shadow_leaf_visits / max(cstats.shadow_rays, 1),
    }
    certificate_audit = {"primitive_tests": int(cstats.leaf_visits - cstats.shadow_leaf_visits), "suspect_hits": int(cstats.certified_suspect_hits),
                         "closest_rays": int(cstats.closest_rays), "retraced": int(cstats.closest_rays_retraced)}
    # ---- adaptive sampling (untimed, reported next to the headline): the same frame with min = spp / 8 < max = spp.  The
    # per-pixel loops of processItem end early where the acceptance test fires (worker.cpp:236-260); the device traces
    # rounds of samples for the pixels still sampling, so a frame whose loops end early costs less.
    adaptive = None
    if args.spp >= 16 and not args.no_adaptive_line:
        astats = capi.RenderStats()
        ao = capi.render_opts(args.width, args.height, max(args.spp // 8, 1), args.spp, 1e-3, args.max_depth, capi.PTB_RNG_COUNTER, flags | capi.PTB_FLAG_DEVICE_IO, 97, 0, rank, world)
        for _ in range(2):  # the first adaptive frame allocates the parked per-pixel state; the second one is reported
            capi.check(lib.ptb_render(handle, C.byref(camera), C.byref(ao), 0, 0, args.width, args.height, C.c_void_p(image.data_ptr()), C.byref(astats)))
        adaptive = {"min_spp": max(args.spp // 8, 1), "max_spp": args.spp, "samples_max": int(sharding.owned_pixels(args.width, args.height, rank, world).sum()) * args.spp, "samples_traced": int(astats.samples),
                    "samples_the_reference_loops_consume": int(astats.samples_used), "rounds": int(astats.adaptive_rounds), "frame_ms": astats.device_ms_total,
                    "fixed_spp_frame_ms": breakdown["frame_ms"], "note": "this rank's tiles; second of two untimed frames; fixed_spp_frame_ms = the profile pass at max_spp"}
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
    rank_rays = float(roof["closest"] + roof["shadow"])
    closest_s = roof["trace_ms"] / 1e3  # the roofline steps time the closest-hit trace only (class 0 events)
    # all traversal kernels: the closest-hit time of the timed steps plus the shadow-trace time of the profile pass
    trace_s = closest_s + breakdown["shadow_trace_ms"] / 1e3 * args.steps
    achieved_all = bytes_per_ray * rank_rays / max(trace_s, 1e-12) / 1e9
    closest_bytes_per_ray = INNER_BYTES * count_detail["closest_inner_per_ray"] + LEAF_BYTES * count_detail["closest_leaf_per_ray"] + RAY_RECORD_BYTES
    achieved = closest_bytes_per_ray * float(roof["closest"]) / max(closest_s, 1e-12) / 1e9
    closest_launches = max(int(roof["iterations"]), 1)
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

    # ---- e2e through the reference-facing API with host buffers: processJob (one call -> the whole image in host memory).
    # On several GPUs rank 0 alone calls it, with ptb::RenderControl::devices = N: the library itself renders on all N GPUs
    # of the node (ptb_render_multi: a host thread and a scene replica per GPU, NVLink gather of the tiles, one D2H) after
    # the other ranks have left.
    def host_barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(group=control)

    # On several GPUs the other ranks have done their part (every collective is behind us): they leave now, so that the
    # one-call path below owns the GPUs the way a single-process caller of processJob does.  (Measured with the ranks
    # merely idling on their GPUs instead: the replica sharing a GPU with another process's context ran 1.5-2x slower.)
    if dist is not None:
        dist.barrier(group=control)
        dist.destroy_process_group()
        dist = None
        if rank != 0:
            return
        deadline = time.time() + 120.0
        for k in range(1, world):
            device = (local_rank + k) % torch.cuda.device_count()  # where Scene::deviceScenes puts replica k
            while time.time() < deadline:
                free_bytes, total_bytes = torch.cuda.mem_get_info(device)
                if total_bytes - free_bytes < (3 << 30):  # the other rank's pool and buffers are gone
                    break
                time.sleep(0.25)

    e2e_ms = []
    e2e_devices = 1
    if rank == 0:
        e2e_devices = b200.set_devices(world)
        b200.set_sharding(0, 1, 2999)
        # untimed warm-up of this path: scene replicas on the other GPUs (first call), then one frame at full size so that the
        # replicas' workspaces (path pool, per-sample buffer) have their final size before the timed steps
        scene_cpp.process_job(camera_cpp, args.width, args.height, min(args.spp, 8), min(args.spp, 8), 1e-3, 0)
        scene_cpp.process_job(camera_cpp, args.width, args.height, args.spp, args.spp, 1e-3, 0)
    for i in range(args.steps):
        flush.zero_()
        host_barrier()
        if rank == 0:
            b200.set_sharding(0, 1, 3000 + i)
            _, job_info = scene_cpp.process_job(camera_cpp, args.width, args.height, args.spp, args.spp, 1e-3, 0)
            e2e_ms.append(job_info["seconds"] * 1e3)  # the processJob call (render + D2H into the caller's Image), timed inside the harness
        host_barrier()
    if rank == 0:
        b200.set_devices(1)
    e2e_value = (job_samples / args.steps) / (np.mean(e2e_ms) / 1e3) / 1e6 if rank == 0 else 0.0

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
        "ms_steps": [round(float(v), 1) for v in step_ms],
        "higher_is_better": True,
        "scaling": "strong",
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
                "d2h_bytes_per_step": args.width * args.height * 16, "ms_per_step": float(np.mean(e2e_ms)), "ms_steps": [round(float(v), 1) for v in e2e_ms],
                "api": "processJob (C++ host API via harness; the duration of the call itself, device->host copy of the image included)" + ("" if world == 1 else f", one call on rank 0 rendering on {e2e_devices} GPUs in-library (ptb_render_multi)")},
        "gpu_launches": int(totals["launches"]),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_note, "peak_source": f"{peak_source} HBM copy bandwidth",
                     "kernel": "traceClosestKernel" + ("<reference tree>" if args.reference_closest else "<certified SAH walk> + re-trace launch"),
                     "bytes_per_ray": closest_bytes_per_ray, "rays_per_launch": float(roof["closest"]) / closest_launches,
                     "algorithmic_bytes_per_launch": closest_bytes_per_ray * float(roof["closest"]) / closest_launches,
                     "ms_per_launch": closest_s * 1e3 / closest_launches,
                     "frac_all_rays": achieved_all / peak, "achieved_all_rays": achieved_all, "bytes_per_ray_all_rays": bytes_per_ray,
                     "inner_fetches_per_ray": inner_per_ray, "leaf_fetches_per_ray": leaf_per_ray,
                     "closest_trace_ms_per_step": roof["trace_ms"] / args.steps, "closest_trace_share_of_step": roof["trace_ms"] / max(roof["frame_ms"], 1e-9),
                     "measured_on": (f"{args.steps} frames rendered on one stream (PTB_FLAG_SINGLE_STREAM) right after the timed steps, {roof['frame_ms'] / args.steps:.1f} ms each: "
                                     "PTB_STREAMS splits the timed frames between concurrent contexts of the GPU, where per-launch event brackets overlap"
                                     if split_streams else "the timed steps (CUDA events around every closest-hit trace launch, on the launching stream)"),
                     "trace_ms_per_step": trace_s * 1e3 / args.steps, "shade_ms_per_step": breakdown["generate_shade_accumulate_resolve_ms"],
                     "shadow_trace_ms_per_step": breakdown["shadow_trace_ms"], "mrays_per_s_trace_only": rank_rays / max(trace_s, 1e-12) / 1e6,
                     "closest_mrays_per_s": roof["closest"] / max(roof["trace_ms"], 1e-9) / 1e3,
                     "shadow_mrays_per_s": breakdown["shadow_mrays_per_s"], **count_detail},
        "breakdown": breakdown,
        "cpu_baseline": cpu_baseline,
        "scene": {"prims": int(info.n_prims), "inner_nodes": int(info.n_inner_nodes), "bvh_depth": int(info.bvh_depth),
                  "device_mb": info.device_bytes / 2**20, "build_s": info.build_seconds, "scene_ctor_s": t_build, "scene_ctor_first_s": t_build_first, "caller_objects_s": t_objects,
                  "upload_s": info.upload_seconds, "built_on_device": bool(info.built_on_device), "reference_tree_device_ms": info.reference_tree_device_ms,
                  "query_tree": ["none", "host binned SAH", "device LBVH", "device full-sweep SAH"][info.query_tree_kind], "query_tree_device_ms": info.query_tree_device_ms,
                  "note": "caller_objects_s = the caller creating its Triangle objects; scene_ctor_first_s / scene_ctor_s = Scene::Scene, first and second time in this process (scene.cpp:153-181: lowering of the object graph + ptb_scene_create); build_s + upload_s = ptb_scene_create"},
        "bounce_iterations_per_step": totals["iterations"] / args.steps,
        "per_rank": per_rank,
        "adaptive": adaptive,
        "certificate_audit": None if args.reference_closest else certificate_audit,
        "closest_hit": ("reference-topology tree" if args.reference_closest else
                        f"certified SAH walk; {totals['retraced']} of {totals['closest']} closest-hit rays had no certificate and were re-traced on the reference tree"),
    }
    print(json.dumps(line))



# ------------------------------------------------------------------------------------------------ C3: triangle soup


def soup_scene_arrays(args):
    """BASELINE configs[2] / SURVEY 8d: N random triangles, centres uniform in [-1,1]^3, vertices = centre + uniform[-s,s]^3,
    s = 0.5 N^(-1/3), no culling; seed 0x5EED0000 + log2 N."""
    from cpupathtrace_b200 import capi, scenes

    n_tris = args.tris << 20
    verts = scenes.soup_triangles(n_tris, 0x5EED0000 + int(np.log2(n_tris)))
    prims = np.zeros(n_tris, capi.PRIM_DTYPE)
    prims["kind"] = capi.PTB_PRIM_TRIANGLE
    prims["p"][:, :9] = verts
    a, b, c = verts[:, 0:3], verts[:, 3:6], verts[:, 6:9]
    nrm = np.cross(b - a, c - a)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
    prims["p"][:, 9:12] = prims["p"][:, 12:15] = prims["p"][:, 15:18] = nrm
    mats = np.zeros(1, capi.MATERIAL_DTYPE)
    mats[0] = ((1, 1, 1, 1), (0, 0, 0, 0), 1.0, 0, 0, 0)
    return prims, mats


def soup_rays_host(kind, n, seed=0x7A750001):
    """coherent: pinhole at (0,0,-3) through a sqrt(n) x sqrt(n) grid over [-1,1]^2 at z = -1; incoherent: uniform
    origins in [-1,1]^3, uniform directions (normalised like rt_vector::normalize: multiply by the reciprocal length)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == "incoherent":
        o = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32)
    else:
        side = int(round(n ** 0.5))
        u = ((np.arange(side, dtype=np.float32) + np.float32(0.5)) / np.float32(side) * 2 - 1).astype(np.float32)
        tx, ty = np.meshgrid(u, u, indexing="xy")
        target = np.stack([tx.ravel(), ty.ravel(), np.full(side * side, -1.0, np.float32)], axis=1)
        o = np.broadcast_to(np.float32([0.0, 0.0, -3.0]), target.shape).copy()
        d = (target - o).astype(np.float32)
    l2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    d = d * (np.float32(1.0) / np.sqrt(l2))[:, None]
    return np.ascontiguousarray(np.concatenate([o, d.astype(np.float32)], axis=1))


def soup_shadow_rays(rays, t, eps=1e-3):
    """worker.cpp:80-86 from every primary hit towards a point light at (0, 0.99, 0): origin pos + dir * eps, limit |to_light| - eps."""
    hit = t >= 0
    pos = rays[hit, :3] + rays[hit, 3:] * t[hit, None]
    to_light = np.float32([0.0, 0.99, 0.0]) - pos
    dist = np.sqrt((to_light * to_light).sum(axis=1)).astype(np.float32)
    ldir = (to_light * (np.float32(1.0) / dist)[:, None]).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([pos + ldir * np.float32(eps), ldir, (dist - np.float32(eps))[:, None]], axis=1).astype(np.float32))


def soup_config(args, n_rays):
    return {"workload": f"BASELINE configs[2]: random-triangle soup, {args.tris} Mi triangles, {n_rays} {args.rays} rays "
                        + ("(any-hit, towards a point light from the primary hits)" if args.rays == "shadow" else "(closest hit)"),
            "tris": args.tris << 20, "rays": args.rays, "n_rays": int(n_rays),
            "l2": f"scene arrays ({(args.tris << 20) * 128 / 2**20:.0f} MiB of nodes + triangles for the query tree) exceed the 126 MB L2 from 1 Mi triangles up; "
                  "a 512 MB buffer is written between steps"}


def soup_reference(args, prims, rays, shadow, seconds):
    """The reference's own Scene::getIntersection on this host's cores: a bounded sample of the ray set on the same soup
    (for the shadow set: the closest-hit query + distance compare of worker.cpp:84-86, which is what the reference runs)."""
    import concurrent.futures

    from cpupathtrace_b200 import pth

    ref = pth.load_reference(fast=True)
    cores = os.cpu_count() or 1
    builder = ref.builder()
    t0 = time.perf_counter()
    builder.triangles(prims["p"][:, :9], prims["p"][:, 9:18], cull=False)
    scene = builder.scene()
    build_s = time.perf_counter() - t0
    query = rays if shadow is None else np.ascontiguousarray(shadow[:, :6])
    chunk = 4096

    def trace(lo):
        scene.intersect(query[lo:lo + chunk])
        return min(chunk, len(query) - lo)

    done, t0 = 0, time.perf_counter()
    with concurrent.futures.ThreadPoolExecutor(cores) as pool:
        starts = list(range(0, len(query), chunk))
        pending = []
        for lo in starts:
            pending.append(pool.submit(trace, lo))
            if len(pending) >= 4 * cores:
                done += pending.pop(0).result()
                if time.perf_counter() - t0 > seconds:
                    break
        for f in pending:
            done += f.result()
    elapsed = time.perf_counter() - t0
    return {"value": done / elapsed / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference",
            "sample": f"{done} of the {len(query)} {args.rays} rays on the same {args.tris} Mi-triangle soup, Scene::getIntersection per ray from {cores} threads "
                      f"({elapsed:.1f} s; reference -O3 -march=x86-64-v3; its BVH build took {build_s:.1f} s and is not counted)",
            "seconds_per_step": elapsed}


def run_soup_arm(args, reference_only=False):
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if reference_only and rank != 0:
        return
    prims, mats = soup_scene_arrays(args)
    primary = soup_rays_host("incoherent" if args.rays == "incoherent" else "coherent", args.n_rays)

    if reference_only:
        # the shadow set needs the primary hits: taken from the reference itself on a subset
        shadow = None
        if args.rays == "shadow":
            from cpupathtrace_b200 import pth

            ref = pth.load_reference(fast=True)
            b = ref.builder()
            b.triangles(prims["p"][:, :9], prims["p"][:, 9:18], cull=False)
            sub = primary[:: max(1, len(primary) // 200000)]
            t_sub, _ = b.scene().intersect(sub)
            shadow = soup_shadow_rays(sub, t_sub)
        base = soup_reference(args, prims, primary, shadow, args.cpu_seconds)
        line = {"metric": "Mrays/s", "value": base["value"], "unit": "Mrays/s", "impl": "reference", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
                "ms_per_step": base["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": soup_config(args, len(primary)), "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch

    from cpupathtrace_b200 import capi

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    os.environ["PTB_DEVICE"] = str(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = capi.Context(-1)
    t0 = time.perf_counter()
    scene = capi.Scene(ctx, prims, mats)
    t_scene = time.perf_counter() - t0
    info = scene.info()
    lib = capi.load()
    dev = torch.device("cuda", local_rank)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    # every rank traces its own contiguous share of the ray set (independent rays: no data-path collective)
    closest_flags = 0 if args.reference_closest else (capi.PTB_FLAG_CERTIFIED_CLOSEST | (0 if args.guarded else capi.PTB_FLAG_CERTIFIED_RELAXED))
    if args.rays == "shadow":
        t_primary, _, _ = scene.intersect(primary, closest_flags)
        rays_host = soup_shadow_rays(primary, t_primary)
    else:
        rays_host = primary
    share = (len(rays_host) + world - 1) // world
    rays_host = np.ascontiguousarray(rays_host[rank * share:(rank + 1) * share])
    n = len(rays_host)
    any_hit = args.rays == "shadow"
    width = rays_host.shape[1]
    pinned_in = torch.from_numpy(rays_host).pin_memory()
    d_rays = pinned_in.to(dev)
    d_t = torch.empty(n, dtype=torch.float32, device=dev)
    d_prim = torch.empty(n, dtype=torch.int32, device=dev)
    d_occ = torch.empty(n, dtype=torch.uint8, device=dev)
    h_t = torch.empty(n, dtype=torch.float32).pin_memory()
    h_prim = torch.empty(n, dtype=torch.int32).pin_memory()
    h_occ = torch.empty(n, dtype=torch.uint8).pin_memory()

    def device_step(extra=0):
        if any_hit:
            return scene.occluded_device(d_rays.data_ptr(), n, d_occ.data_ptr(), extra)
        return scene.intersect_device(d_rays.data_ptr(), n, d_t.data_ptr(), d_prim.data_ptr(), closest_flags | extra)

    def host_step():
        stats = capi.RenderStats()
        if any_hit:
            capi.check(lib.ptb_occluded(scene._h, C.c_void_p(pinned_in.data_ptr()), n, C.c_void_p(h_occ.data_ptr()), 0, C.byref(stats)))
        else:
            capi.check(lib.ptb_intersect(scene._h, C.c_void_p(pinned_in.data_ptr()), n, C.c_void_p(h_t.data_ptr()), C.c_void_p(h_prim.data_ptr()), closest_flags, C.byref(stats)))
        return stats

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def over_ranks(value, op):
        if dist is None:
            return value
        t = torch.tensor([value], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        barrier()
        device_step()
    step_ms, trace_ms = [], 0.0
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        stats = device_step()
        step_ms.append(over_ranks(stats.device_ms_total, torch.distributed.ReduceOp.MAX if dist is not None else None))
        trace_ms += stats.device_ms_trace
    clocks = sampler.stop()
    counted = device_step(capi.PTB_FLAG_COUNT_VISITS)
    inner, leaf = counted.inner_visits / n, counted.leaf_visits / n
    ray_bytes = (28 + 1) if any_hit else (24 + 8)
    bytes_per_ray = INNER_BYTES * inner + LEAF_BYTES * leaf + ray_bytes

    # size-independent parity: the walk the bench times returns what the reference-order walk returns, ray for ray
    identical = None
    if not any_hit and not args.reference_closest:
        t_fast, prim_fast = d_t.clone(), d_prim.clone()
        scene.intersect_device(d_rays.data_ptr(), n, d_t.data_ptr(), d_prim.data_ptr(), 0)
        hit = d_t >= 0
        identical = bool(torch.equal(d_prim, prim_fast) and torch.equal(d_t[hit], t_fast[hit]) and bool((t_fast[~hit] < 0).all()))

    e2e_ms = []
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        host_step()
        e2e_ms.append(over_ranks((time.perf_counter() - t0) * 1e3, torch.distributed.ReduceOp.MAX if dist is not None else None))
    total_rays = over_ranks(float(n), torch.distributed.ReduceOp.SUM if dist is not None else None)
    if rank != 0:
        dist.destroy_process_group()
        return

    value = total_rays / (np.mean(step_ms) / 1e3) / 1e6
    peak, peak_source = 6650.0, "fallback"
    peaks_path = os.path.join(REPO_ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        try:
            peak, peak_source = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        except (KeyError, ValueError):
            pass
    achieved = bytes_per_ray * n * args.steps / max(trace_ms / 1e3, 1e-12) / 1e9
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from cpupathtrace_b200 import pth

        if os.path.exists(pth.REF_FAST):
            base = soup_reference(args, prims, primary, rays_host if any_hit else None, args.cpu_seconds)
            cpu_baseline = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    out_bytes = n * (1 if any_hit else 8)
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": float(np.mean(step_ms)),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": soup_config(args, total_rays),
        "clocks": clocks,
        "e2e": {"value": total_rays / (np.mean(e2e_ms) / 1e3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(n * width * 4), "d2h_bytes_per_step": int(out_bytes),
                "ms_per_step": float(np.mean(e2e_ms)), "api": "ptb_occluded" if any_hit else "ptb_intersect", "note": "pinned host ray buffer in, pinned host results out, per rank"},
        "gpu_launches": int(args.steps * (3 + (0 if any_hit or args.reference_closest else 1))),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": f"{peak_source} HBM copy bandwidth",
                     "kernel": "occludedKernel" if any_hit else ("intersectKernel<reference tree>" if args.reference_closest else "intersectKernel<certified SAH walk> + re-trace launch"),
                     "bytes_per_ray": bytes_per_ray, "inner_fetches_per_ray": inner, "leaf_fetches_per_ray": leaf, "rays_per_launch": n,
                     "algorithmic_bytes_per_launch": bytes_per_ray * n, "ms_per_launch": trace_ms / args.steps, "sort_ms_per_launch": float(np.mean(step_ms)) - trace_ms / args.steps,
                     "note": "rays are traced in Morton order of (direction octant, origin); ms_per_launch is the traversal kernel(s), ms_per_step adds the key + radix-sort pass"},
        "cpu_baseline": cpu_baseline,
        "scene": {"prims": int(info.n_prims), "bvh_depth": int(info.bvh_depth), "device_mb": info.device_bytes / 2**20, "build_s": info.build_seconds, "scene_ctor_s": t_scene,
                  "certifiable": int(info.certifiable)},
        "identical_to_reference_walk": identical, "retraced": int(counted.closest_rays_retraced),
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

    if args.workload == "soup":
        run_soup_arm(args, reference_only=args.impl == "reference")
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
