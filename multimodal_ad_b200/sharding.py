"""Subject sharding for N ranks (one process per GPU).

ROI extraction shards by subject and ResNet training shards by batch; neither
needs a data-path collective for the forward pass.  These helpers hold the
host-side logic so it can be tested with gloo on CPU.
"""
from __future__ import annotations

import torch


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced (sizes differ by at most 1) share of `n_items` for `rank`."""
    if not 0 <= rank < world:
        raise ValueError("rank outside [0, world)")
    lo = n_items * rank // world
    hi = n_items * (rank + 1) // world
    return range(lo, hi)


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (step time) across the process group; identity without one."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_subject_rows(local_rows: torch.Tensor, n_total: int):
    """All-gather per-subject result rows (ragged by at most one row) in subject order; returns
    the (n_total, ...) tensor on every rank.  Only used off the timed path (results collection)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_rows
    world, rank = dist.get_world_size(), dist.get_rank()
    mx = (n_total + world - 1) // world
    pad = torch.zeros((mx,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([bufs[r][: len(shard_range(n_total, r, world))] for r in range(world)], dim=0)
