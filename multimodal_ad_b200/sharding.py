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

    The accelerated ResNet asks the reducer for the storage of each convolution's weight gradient (`alloc_like`): large
    gradients become views of flat BUCKETS (`bucket_numel` elements), filled in reverse layer order.  A bucket is
    all-reduced asynchronously as soon as it is closed (the next gradient does not fit) and all its members have been
    handed over (`push`, called with the producing stream current), on a helper stream that waits for the producers'
    events - a handful of large collectives running under the remaining backward kernels instead of one per layer
    (every collective launch delays the persistent convolution kernels that share the SMs with it).  Small tensors
    (BatchNorm scales / shifts, the classifier head) are coalesced into one flat all-reduce in `finish`, which also
    waits for everything.  Replaces the reference's nn.DataParallel replication (Resnet3D.py:89-99) with the standard
    all-reduce formulation.

    Works with any backend: NCCL averages in the collective (AVG), others SUM + divide; tensors stay where they are.
    """

    def __init__(self, group=None, large_numel: int = 1 << 16, bucket_numel: int = 1 << 23):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.large = large_numel
        self.bucket_numel = bucket_numel
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.active else 1
        self.avg = self.active and dist.get_backend(group) == "nccl"
        self._handles, self._small, self._seen = [], [], set()
        self._buckets, self._view_bucket, self._comm_stream = [], {}, None

    # ---- bucket storage -------------------------------------------------------------------------------------------
    def alloc_like(self, param: torch.Tensor) -> torch.Tensor:
        """Uninitialised gradient storage shaped like `param` (fp32 or param's dtype); a bucket view when bucketing applies."""
        n = param.numel()
        if not self.active or n < self.large or self.bucket_numel <= 0:
            return torch.empty_like(param)
        b = self._buckets[-1] if self._buckets and not self._buckets[-1]["closed"] else None
        if b is not None and (b["used"] + n > b["flat"].numel() or b["flat"].dtype != param.dtype or b["flat"].device != param.device):
            self._close(b)
            b = None
        if b is None:
            b = dict(flat=torch.empty(max(self.bucket_numel, n), dtype=param.dtype, device=param.device), used=0, pending=set(),
                     events=[], closed=False, launched=False)
            self._buckets.append(b)
        view = b["flat"][b["used"]:b["used"] + n].view(param.shape)
        b["used"] += n
        b["pending"].add(id(view))
        self._view_bucket[id(view)] = (b, view)          # keeps the view (and so its id) alive
        return view

    def flush(self) -> None:
        """Close the bucket that is being filled: it is all-reduced as soon as its members have been handed over, instead of
        waiting for `finish`.  The backward pass calls this before its last (stem) stage, so that the final bucket's collective runs
        under that stage's kernels and only the small-tensor all-reduce is left for the end of the step."""
        if self.active and self._buckets and not self._buckets[-1]["closed"]:
            self._close(self._buckets[-1])

    def _close(self, b) -> None:
        b["closed"] = True
        if not b["pending"]:
            self._launch(b)

    def _launch(self, b) -> None:
        if b["launched"] or b["used"] == 0:
            return
        b["launched"] = True
        flat = b["flat"][: b["used"]]
        op = self.dist.ReduceOp.AVG if self.avg else self.dist.ReduceOp.SUM
        if flat.is_cuda:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=flat.device)
            for ev in b["events"]:
                self._comm_stream.wait_event(ev)
            with torch.cuda.stream(self._comm_stream):
                h = self.dist.all_reduce(flat, op=op, group=self.group, async_op=True)
        else:
            h = self.dist.all_reduce(flat, op=op, group=self.group, async_op=True)
        self._handles.append((h, flat))

    # ---- hand-over ------------------------------------------------------------------------------------------------
    def push(self, grad: torch.Tensor) -> None:
        if not self.active or grad is None or id(grad) in self._seen:
            return
        self._seen.add(id(grad))
        if id(grad) in self._view_bucket:
            b = self._view_bucket[id(grad)][0]
            b["pending"].discard(id(grad))
            if grad.is_cuda:
                ev = torch.cuda.Event()
                ev.record()                               # on the stream that produces this gradient
                b["events"].append(ev)
            if b["closed"] and not b["pending"]:
                self._launch(b)
        elif grad.numel() >= self.large:
            op = self.dist.ReduceOp.AVG if self.avg else self.dist.ReduceOp.SUM
            self._handles.append((self.dist.all_reduce(grad, op=op, group=self.group, async_op=True), grad))
        else:
            self._small.append(grad)

    def finish(self, params=None) -> None:
        """Reduce what is still pending (optionally every `.grad` of `params` not pushed yet), wait, average."""
        if not self.active:
            self._handles, self._small, self._seen = [], [], set()
            self._buckets, self._view_bucket = [], {}
            return
        if params is not None:
            for p in params:
                if p.grad is not None:
                    self.push(p.grad)
        for b in self._buckets:                            # open or incomplete buckets: everything is enqueued by now
            if not b["launched"]:
                if b["flat"].is_cuda:
                    ev = torch.cuda.Event()
                    ev.record()
                    b["events"].append(ev)
                b["closed"] = True
                self._launch(b)
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
            if not self.avg:
                g.div_(self.world)
        self._handles, self._small, self._seen = [], [], set()
        self._buckets, self._view_bucket = [], {}
