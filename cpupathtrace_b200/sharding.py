"""Multi-GPU work split of a frame: interleaved tiles + one image-sized sum-reduce.

The reference parallelises one frame over CPU threads by tiles (src/worker.cpp:398-414).  Across GPUs the same tile
grid is dealt round-robin: tile k (row-major) belongs to rank k % world.  ptb_render implements the rule on the device
side (ptb_render_opts.shard_index / shard_count: pixels of foreign tiles are written as 0); this module restates it for
the host so that tests can check the split and so that bench.py and the tests share one reduce routine.
"""
import numpy as np


def reference_tile_size(width, height):
    """processJob's tile edge (worker.cpp:398): clamp(min(W, H) / 4, 1, 32)."""
    return max(min(min(width, height) // 4, 32), 1)


def tile_grid(width, height, tile=None):
    tile = tile or reference_tile_size(width, height)
    return (width + tile - 1) // tile, (height + tile - 1) // tile, tile


def owner_map(width, height, world, tile=None):
    """[height, width] int array: rank that renders each pixel."""
    tiles_x, tiles_y, tile = tile_grid(width, height, tile)
    ty = np.arange(height) // tile
    tx = np.arange(width) // tile
    return ((ty[:, None] * tiles_x + tx[None, :]) % max(world, 1)).astype(np.int32)


def owned_pixels(width, height, rank, world, tile=None):
    return owner_map(width, height, world, tile) == rank


def reduce_image(image, dist=None, dst=0):
    """Sum-reduces per-rank images (disjoint tiles, zeros elsewhere) onto rank `dst` in place.

    `image` is a torch tensor (CUDA with the nccl backend, CPU with gloo); with dist None (single process) it is a
    no-op.  This is the only collective of the render path: scene and BVH are replicated, paths never communicate."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return image
    dist.reduce(image, dst=dst, op=dist.ReduceOp.SUM)
    return image
