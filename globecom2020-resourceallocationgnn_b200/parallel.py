"""Data-parallel plumbing: every batch row is an independent graph (``batch_dot`` is per sample,
BS_brain.py:73), so the batch is split contiguously over ranks and the only exchange is ONE sum
all-reduce of the flat gradient buffer per step (NCCL over NVLink on the GPU box, gloo in CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int):
    """Contiguous, balanced split of ``batch`` rows: returns (lo, hi) for ``rank``."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradients_(flat: torch.Tensor, weight: float = 1.0, group=None):
    """In-place SUM all-reduce of a flat gradient buffer.  Each rank's gradient is that of a *local*
    mean loss, so the caller divides by world size afterwards; ``weight`` = local_rows / (B / world)
    re-weights uneven shards so that the result is exactly the full-batch mean gradient."""
    if weight != 1.0:
        flat.mul_(weight)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat
