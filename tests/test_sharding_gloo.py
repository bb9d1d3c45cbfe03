"""N>1 host logic on CPU: world_size-2 gloo process group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_ad_b200.sharding import gather_subject_rows, max_over_ranks, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in shard_range(n, r, world)]
            assert got == list(range(n))
            sizes = [len(shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_subjects = 7
        mine = shard_range(n_subjects, rank, world)
        rows = torch.tensor([[float(i), float(i) * 2] for i in mine]).reshape(len(mine), 2)
        full = gather_subject_rows(rows, n_subjects)
        assert torch.equal(full, torch.tensor([[float(i), float(i) * 2] for i in range(n_subjects)]))
        t = max_over_ranks(1.0 + rank)
        assert t == float(world)
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"


def _reducer_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multimodal_ad_b200.sharding import GradReducer

        red = GradReducer(large_numel=8)
        big = torch.full((16,), float(rank + 1))
        small = [torch.full((3,), float(10 * (rank + 1))), torch.full((2, 2), float(rank))]
        lin = torch.nn.Linear(2, 1)
        lin.weight.grad = torch.full((1, 2), float(rank + 5))
        lin.bias.grad = torch.full((1,), float(rank))
        red.push(big)
        for t in small:
            red.push(t)
        red.push(big)                                     # pushing the same tensor twice must not reduce it twice
        # bucketed storage: views of flat buckets, reduced per bucket once closed and complete (or at finish)
        red.bucket_numel = 40
        w1, w2, w3 = torch.nn.Parameter(torch.zeros(4, 6)), torch.nn.Parameter(torch.zeros(30)), torch.nn.Parameter(torch.zeros(2))
        g1 = red.alloc_like(w1); g1.fill_(float(rank + 1))
        red.push(g1)
        g2 = red.alloc_like(w2); g2.fill_(float(3 * rank))   # does not fit behind g1: closes (and launches) the first bucket
        g3 = red.alloc_like(w3); g3.fill_(7.0)                # below large_numel: plain tensor
        assert g1.shape == (4, 6) and g2.shape == (30,) and g1.untyped_storage().data_ptr() != g2.untyped_storage().data_ptr()
        red.push(g2)
        red.push(g3)
        # flush(): the open bucket (g2's) is closed and reduced NOW (the backward pass does this before its last stage); a
        # gradient allocated afterwards starts a new bucket
        assert not red._buckets[-1]["launched"]
        red.flush()
        assert red._buckets[-1]["closed"] and red._buckets[-1]["launched"]
        w4 = torch.nn.Parameter(torch.zeros(12))
        g4 = red.alloc_like(w4); g4.fill_(float(2 * rank))
        assert g4.untyped_storage().data_ptr() != g2.untyped_storage().data_ptr()
        red.push(g4)
        red.finish(lin.parameters())
        assert torch.allclose(g4, torch.full((12,), 1.0))
        assert torch.allclose(g1, torch.full((4, 6), 1.5)) and torch.allclose(g2, torch.full((30,), 1.5)) and torch.allclose(g3, torch.full((2,), 7.0))
        assert torch.allclose(big, torch.full((16,), 1.5))
        assert torch.allclose(small[0], torch.full((3,), 15.0)) and torch.allclose(small[1], torch.full((2, 2), 0.5))
        assert torch.allclose(lin.weight.grad, torch.full((1, 2), 5.5)) and torch.allclose(lin.bias.grad, torch.full((1,), 0.5))
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_grad_reducer_world_size_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"


def test_grad_reducer_is_a_no_op_without_a_process_group():
    from multimodal_ad_b200.sharding import GradReducer

    red = GradReducer()
    t = torch.ones(4)
    red.push(t)
    red.finish()
    assert not red.active and torch.equal(t, torch.ones(4))
