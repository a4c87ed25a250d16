"""Spatial sharding of a datacube across the GPUs of one box (one process per GPU).

Every spatial pixel of `evaluate_ybar` is independent, so the cube is
partitioned into contiguous blocks of x1 rows, one per rank, with no exchange during compute.
The only collective is the final gather of the per-rank cube tiles onto rank 0 (NCCL over
NVLink on the GPU box; gloo in the CPU tests): ranks produce spaxel-major tiles
[rows*nx2, n_fft] so the gathered buffer is one contiguous [nx1*nx2, n_fft] array that a
single transpose turns into the reference layout [n_fft, nx1, nx2].
"""
import numpy as np


def row_partition(nx1, world_size):
    """Contiguous x1-row blocks: rank r owns rows [b[r], b[r+1]); sizes differ by at most 1."""
    if world_size < 1 or nx1 < 0:
        raise ValueError("need world_size >= 1 and nx1 >= 0")
    return [(r * nx1) // world_size for r in range(world_size + 1)]


def my_rows(nx1, world_size, rank):
    b = row_partition(nx1, world_size)
    return b[rank], b[rank + 1]


def gather_pixel_major(tile, nx1, nx2, dist=None, dst=0):
    """Gather per-rank spaxel-major tiles [rows_r*nx2, n] on `dst`.

    Returns the [nx1*nx2, n] array on `dst` (None elsewhere).  `tile` is a torch tensor on the
    backend's device (CUDA for nccl, CPU for gloo).  With `dist=None` (single process) the
    tile is returned unchanged."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return tile
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = row_partition(nx1, world)
    n = tile.shape[1]
    if tile.shape[0] != (bounds[rank + 1] - bounds[rank]) * nx2:
        raise ValueError("tile does not match this rank's row block")
    if rank == dst:
        full = torch.empty((nx1 * nx2, n), dtype=tile.dtype, device=tile.device)
        parts = [full[bounds[r] * nx2:bounds[r + 1] * nx2] for r in range(world)]
        # ranks own different row counts: point-to-point instead of a padded gather
        reqs = [dist.irecv(parts[r], src=r) for r in range(world) if r != dst and parts[r].numel()]
        parts[dst].copy_(tile)
        for q in reqs:
            q.wait()
        return full
    if tile.numel():
        dist.send(tile.contiguous(), dst=dst)
    return None


def pixel_major_to_cube(full, nx1, nx2):
    """[nx1*nx2, n] -> [n, nx1, nx2] (host/NumPy or torch, any device) - the reference layout."""
    if hasattr(full, "numpy") or type(full).__module__.startswith("torch"):
        return full.t().contiguous().reshape(full.shape[1], nx1, nx2)
    return np.ascontiguousarray(full.T).reshape(full.shape[1], nx1, nx2)
