"""One process per GPU: Gaussian slices and the K-means exchange step.

Lifting shards by Gaussian with every view replicated and needs no collective; K-means
all-reduces K x (D+1) float64 partial sums per iteration (k_means.lloyd does it when a process
group is initialised).  NCCL on GPUs, gloo in the CPU tests of the slicing logic.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def slice_bounds(n: int, rank: int, world: int):
    """Rows [lo, hi) owned by `rank`: contiguous, index order preserved, sizes differ by <= 1."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_labels(local_labels: torch.Tensor, n_total: int, rank: int, world: int, dst: int = 0):
    """Concatenate per-rank label slices on `dst` (for the PLY writer).  Returns the full
    tensor on dst, None elsewhere."""
    if world == 1:
        return local_labels
    sizes = [slice_bounds(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(pad, dtype=local_labels.dtype, device=local_labels.device)
    buf[: local_labels.numel()] = local_labels
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)])


def barrier_max_ms(ms: float, device) -> float:
    """Max over ranks of a per-rank duration."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
