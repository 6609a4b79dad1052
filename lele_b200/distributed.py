"""Multi-GPU plumbing for the one place the path shards: independent clips.

One process per GPU (torch.distributed: NCCL on the GPU box, gloo in CPU tests).  Exactly two
collectives exist, both outside the forward pass:
  * one broadcast of the weights blob from rank 0 at start-up (~240 MB over NVLink),
  * one gather of the greedy ids ([clips_per_rank, T'] int32) per batch.
No kernel is followed by a collective inside the forward, so there is nothing to fuse.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


def nccl_library_path() -> str | None:
    """The NCCL a Python host should bind: the pip-installed one next to torch (nvidia/nccl/lib/libnccl.so.2) when it exists.  A
    process has ONE libnccl.so.2 (the loader de-duplicates by soname): if the library bound the system NCCL first, a later
    `import torch` would be handed that copy instead of the version it was built against."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    return None


def ensure_nccl_env():
    if not os.environ.get("LELE_B200_NCCL_LIB"):
        p = nccl_library_path()
        if p:
            os.environ["LELE_B200_NCCL_LIB"] = p


ensure_nccl_env()


def nccl_version() -> int:
    from ._lib import lib
    return int(lib.lele_b200_comm_nccl_version())


class Comm:
    """The library's own communicator (include/lele_b200.h `lele_b200_comm_*`: NCCL bound at run time): what a host without
    torch.distributed uses.  `Comm.create(ctx, rank, world, exchange)`: rank 0 makes the 128-byte NCCL id, `exchange(bytes | None)
    -> bytes` is the host's own rendez-vous (returns rank 0's bytes on every rank; bench.py passes a torch.distributed object
    broadcast, a Rust host would use a file or a socket)."""

    def __init__(self, ctx, handle, rank, world):
        self.ctx, self.h, self.rank, self.world = ctx, handle, rank, world

    @classmethod
    def create(cls, ctx, rank: int, world: int, exchange):
        from ._lib import call
        uid = (C.c_char * 128)()
        if rank == 0:
            call("lele_b200_comm_unique_id", uid)
        raw = exchange(bytes(uid.raw) if rank == 0 else None)
        uid = (C.c_char * 128).from_buffer_copy(raw)
        h = C.c_void_p()
        call("lele_b200_comm_create", ctx.h, uid, C.c_int(world), C.c_int(rank), C.byref(h))
        return cls(ctx, h, rank, world)

    def broadcast(self, dev_ptr: int, nbytes: int, root: int = 0):
        from ._lib import call
        call("lele_b200_comm_broadcast", self.ctx.h, self.h, C.c_void_p(dev_ptr), C.c_size_t(nbytes), C.c_int(root))

    def gather(self, send_ptr: int, recv_ptr: int | None, nbytes: int, root: int = 0):
        from ._lib import call
        call("lele_b200_comm_gather", self.ctx.h, self.h, C.c_void_p(send_ptr), C.c_void_p(recv_ptr), C.c_size_t(nbytes), C.c_int(root))

    def close(self):
        if self.h:
            from ._lib import lib
            lib.lele_b200_comm_destroy(self.h); self.h = None


def bind_to_gpu_numa_node(gpu_index: int) -> str:
    """Pins this process to the CPUs local to its GPU (NVML's affinity mask intersected with what the container allows): eight ranks
    feeding pinned H2D copies from one NUMA node was the measured limiter of the 8-GPU end-to-end curve.  Returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        pick = cpus & allowed
        if not pick:
            return f"gpu {gpu_index}: NVML affinity {len(cpus)} cpus, none allowed here; unchanged ({len(allowed)} cpus)"
        os.sched_setaffinity(0, pick)
        return f"gpu {gpu_index}: bound to {len(pick)} of {len(allowed)} cpus (NVML affinity, NUMA-local)"
    except Exception as e:   # no NVML / not permitted: leave the affinity alone
        return f"gpu {gpu_index}: affinity unchanged ({type(e).__name__})"


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
