"""CPU test (-m "not gpu") of the multi-GPU data path with two gloo ranks: every rank produces the frame restricted to
its interleaved tiles (zeros elsewhere), one sum-reduce to rank 0 reassembles the full frame.  The per-pixel values
come from the oracle restatement keyed per (pixel, sample) exactly like the device's reference-engine mode, so the
reassembled frame must equal the single-rank frame bit for bit (adding zeros is exact)."""
import os
import socket
import sys

import numpy as np
import pytest

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _frame(rank, world, width, height, spp):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpupathtrace_b200 import sharding
    from helpers import camera_kwargs, counter_key, load_golden, pod_camera
    from oracle import pto

    g = load_golden("samples", "mixed")
    scene = pto.OracleScene(g["prims"], g["materials"], g["lights"])
    camera = pod_camera(camera_kwargs(g["camera"]))
    mask = sharding.owned_pixels(width, height, rank, world)
    ys, xs = np.nonzero(mask)
    frame = np.zeros((height, width, 4), np.float32)
    samples = np.zeros((spp, len(xs), 4), np.float32)
    for s in range(spp):
        seeds = counter_key(7, xs, ys, np.full(len(xs), s))
        samples[s], _ = scene.render_samples(camera, width, height, 1e-3, np.stack([xs, ys], axis=1).astype(np.int32), seeds)
    frame[ys, xs] = pto.resolve(spp, spp, samples)
    return frame


def _worker(rank, world, port, width, height, spp, out_path):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from cpupathtrace_b200 import sharding

    image = torch.from_numpy(_frame(rank, world, width, height, spp))
    sharding.reduce_image(image, dist, dst=0)
    if rank == 0:
        np.save(out_path, image.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tile_sharding_reassembles_the_frame(tmp_path):
    import torch.multiprocessing as mp

    width, height, spp = 40, 24, 4
    out_path = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(WORLD, _free_port(), width, height, spp, out_path), nprocs=WORLD, join=True)
    reassembled = np.load(out_path)
    full = _frame(0, 1, width, height, spp)
    assert np.array_equal(reassembled, full)
    assert reassembled[..., 3].max() == 1.0
