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


class GradReducer:
    """Data-parallel gradient averaging overlapped with the backward pass (one process per GPU, NCCL over NVLink).

    The accelerated ResNet hands over each convolution's weight gradient the moment its kernels are enqueued
    (`push`), in reverse layer order; large tensors start an asynchronous all-reduce immediately, so the collective
    runs under the remaining backward kernels.  Small tensors (BatchNorm scales / shifts, the classifier head) are
    coalesced into one flat all-reduce in `finish`, which also waits for everything.  Replaces the reference's
    nn.DataParallel replication (Resnet3D.py:89-99) with the standard all-reduce formulation.

    Works with any backend: SUM + divide (gloo has no AVG), tensors stay where they are.
    """

    def __init__(self, group=None, large_numel: int = 1 << 16):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.large = large_numel
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.active else 1
        self._handles, self._small, self._seen = [], [], set()

    def push(self, grad: torch.Tensor) -> None:
        if not self.active or grad is None or id(grad) in self._seen:
            return
        self._seen.add(id(grad))
        if grad.numel() >= self.large:
            self._handles.append((self.dist.all_reduce(grad, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True), grad))
        else:
            self._small.append(grad)

    def finish(self, params=None) -> None:
        """Reduce what is still pending (optionally every `.grad` of `params` not pushed yet), wait, average."""
        if not self.active:
            self._handles, self._small, self._seen = [], [], set()
            return
        if params is not None:
            for p in params:
                if p.grad is not None:
                    self.push(p.grad)
        if self._small:
            flat = torch.cat([g.reshape(-1) for g in self._small])
            self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)
            off = 0
            for g in self._small:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        for h, g in self._handles:
            h.wait()
            g.div_(self.world)
        self._handles, self._small, self._seen = [], [], set()
