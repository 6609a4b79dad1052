"""Multi-GPU plumbing for the one place the path shards: independent clips.

One process per GPU (torch.distributed: NCCL on the GPU box, gloo in CPU tests).  Exactly two
collectives exist, both outside the forward pass:
  * one broadcast of the weights blob from rank 0 at start-up (~240 MB over NVLink),
  * one gather of the greedy ids ([clips_per_rank, T'] int32) per batch.
No kernel is followed by a collective inside the forward, so there is nothing to fuse.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_clips: int, rank: int, world: int) -> tuple[int, int]:
    """Static block partition of the clip index (last shards may be one shorter)."""
    base, rem = divmod(n_clips, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_blob(blob_tensor, src: int = 0):
    """In-place broadcast of the u8 weights blob (torch tensor on this rank's device)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(blob_tensor, src=src)
    return blob_tensor


def gather_ids(local_ids, n_clips_total: int, dst: int = 0):
    """Gathers per-rank ids [n_local, T] (torch int32 tensor) to `dst`; returns [n_clips_total, T]
    on dst and None elsewhere.  Shards may differ by one clip, so they are padded to the
    largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_ids
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_clips_total, r, world) for r in range(world)]
    max_n = max(e - s for s, e in sizes)
    T = local_ids.shape[1]
    padded = torch.zeros((max_n, T), dtype=local_ids.dtype, device=local_ids.device)
    padded[: local_ids.shape[0]] = local_ids
    if rank == dst:
        bufs = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded, bufs, dst=dst)
        return torch.cat([bufs[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)
    dist.gather(padded, None, dst=dst)
    return None


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
