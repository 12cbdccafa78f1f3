"""Host logic of the data-parallel path on CPU: flat buffers, batch sharding and the one
gradient all-reduce, with world_size 2 over gloo (the GPU path uses the same code over NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hotrack_b200.flat import FlatParams, shard_batch


def test_shard_batch_partitions_exactly():
    for n in (0, 1, 7, 32, 256, 257):
        for w in (1, 2, 3, 8):
            spans = [shard_batch(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_flat_params_alias_parameters_and_grads():
    m = torch.nn.Sequential(torch.nn.Conv1d(3, 5, 1), torch.nn.BatchNorm1d(5), torch.nn.Conv1d(5, 2, 1))
    before = [p.detach().clone() for p in m.parameters()]
    flat = FlatParams(m)
    for p, b, o in zip(flat.params, before, flat.offsets):
        assert torch.equal(p, b) and o % 4 == 0
        assert p.data_ptr() == flat.data.data_ptr() + 4 * o
    m(torch.randn(4, 3, 9)).sum().backward()
    assert flat.grad.abs().sum() > 0
    g = flat.grad.clone()
    flat.zero_grad()
    assert flat.grad.abs().sum() == 0
    m(torch.randn(4, 3, 9)).sum().backward()  # accumulates into the same flat buffer again
    assert flat.grad.abs().sum() > 0 and g.shape == flat.grad.shape
    with torch.no_grad():
        flat.data.add_(1.0)
    for p, b in zip(flat.params, before):
        assert torch.allclose(p, b + 1.0)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.Linear(6, 3))
    flat = FlatParams(m)
    with torch.no_grad():
        flat.data.add_(float(rank))  # replicas start different ...
    flat.broadcast(0)                # ... and are made identical
    data = torch.arange(8 * 4, dtype=torch.float32).reshape(8, 4) / 10.0
    lo, hi = shard_batch(8, rank, world)
    flat.zero_grad()
    (m(data[lo:hi]).square().sum() / 8).backward()
    scale = flat.allreduce_grads()
    q.put((rank, flat.data.clone(), flat.grad.clone(), scale))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, d0, g0, s0), (_, d1, g1, s1) = res
    assert torch.equal(d0, d1) and torch.equal(g0, g1) and s0 == s1 == 0.5
    # single-process full-batch gradient of the same loss
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.Linear(6, 3))
    flat = FlatParams(m)
    data = torch.arange(8 * 4, dtype=torch.float32).reshape(8, 4) / 10.0
    (m(data).square().sum() / 8).backward()
    torch.testing.assert_close(g0, flat.grad, rtol=1e-5, atol=1e-6)  # SUM of shard grads == full-batch grad


def test_zero_arena_bump_allocation_and_fallback():
    """fused.ZeroArena (TrainStep's one-memset-per-step accumulator buffer): slices are disjoint, 16-byte granular, zero
    after begin(), and a stack that finds the arena exhausted falls back to a fresh allocation."""
    import torch

    from hotrack_b200 import fused

    dev = torch.device("cpu")
    arena = fused.ZeroArena(dev, floats=64)
    arena.begin()
    a, b = arena.take(5), arena.take(9)
    assert a.numel() == 8 and b.numel() == 12  # rounded up to 4-float granules
    a.fill_(1.0)
    b.fill_(2.0)
    assert a.data_ptr() + 8 * 4 == b.data_ptr() and float(arena.buf.sum()) == 8 + 24
    assert arena.take(64) is None  # exhausted: callers fall back
    arena.begin()
    assert float(arena.buf.abs().sum()) == 0.0 and arena.off == 0
    fused.ACTIVE_ARENA = arena
    try:
        z = fused._zeros(10, dev)
        assert z.numel() == 10 and z.data_ptr() == arena.buf.data_ptr()
        big = fused._zeros(1000, dev)  # does not fit -> ordinary zeros
        assert big.numel() == 1000 and float(big.abs().sum()) == 0.0
    finally:
        fused.ACTIVE_ARENA = None
    assert fused._zeros(3, dev).numel() == 3
