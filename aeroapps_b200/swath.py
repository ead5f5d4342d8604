"""Sharding of a swath / cube-face slab over the GPUs of one box: one process per GPU, each
owns a contiguous block of observations (columns, pixels); there is no data-path collective -- the only exchange is
the final gather of the per-rank outputs (NCCL on the GPUs, gloo in the CPU tests).

The reference shards the same way across MPI ranks / python multiprocessing workers: contiguous pixel blocks per
process, results written back by block (geo_vlidort/geo_vlidort.F90 pixel partition, leo_vlidort/mp_lidar_vlidort.py)."""
import numpy as np


def slab_bounds(nobs, world, rank):
    """[begin, end) of rank's contiguous block; blocks differ by at most one observation, earlier ranks take the extras"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("slab_bounds: bad world/rank")
    base, extra = divmod(int(nobs), int(world))
    b = rank * base + min(rank, extra)
    return b, b + base + (1 if rank < extra else 0)


def slab_of(arr, axis, world, rank):
    """view of the rank's block of `arr` along the observation axis"""
    b, e = slab_bounds(arr.shape[axis], world, rank)
    idx = [slice(None)] * arr.ndim
    idx[axis] = slice(b, e)
    return arr[tuple(idx)]


def gather_swath(local, nobs, axis=0, dist=None, dst=0, device=None):
    """final gather: every rank passes its block (numpy array, observation axis `axis`); rank `dst` gets the whole
    swath back (others get None).  dist = torch.distributed (initialised) or None for a single process."""
    if dist is None or dist.get_world_size() == 1:
        return np.asarray(local)
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    loc = np.moveaxis(np.ascontiguousarray(local), axis, 0)
    rest = loc.shape[1:]
    sizes = [slab_bounds(nobs, world, r)[1] - slab_bounds(nobs, world, r)[0] for r in range(world)]
    mx = max(sizes)
    buf = torch.zeros((mx,) + rest, dtype=torch.from_numpy(loc[:0].copy()).dtype, device=device)
    buf[:loc.shape[0]] = torch.from_numpy(np.ascontiguousarray(loc)).to(buf.device)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)                      # NCCL has no gather-to-one with ragged sizes; blocks are small
    if rank != dst:
        return None
    full = np.concatenate([o[:s].cpu().numpy() for o, s in zip(outs, sizes)], axis=0)
    return np.moveaxis(full, 0, axis)
