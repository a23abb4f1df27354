"""Batch sharding across GPUs: states are independent, so rank g of G owns one contiguous slice
and nothing is exchanged on the compute path.  The only collectives are the barrier and the max-over-ranks of
the per-rank device times that bench.py reports."""
from __future__ import annotations


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of `total` items for `rank`; the first total % world ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """All-reduce(MAX) of a per-rank scalar (device time in ms).  No-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
