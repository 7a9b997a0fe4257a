"""CPU tests of the N>1 path: world_size-2 gloo processes exercise the scatter → per-rank solve →
gather plumbing of stabletts_b200.shard with a stand-in per-utterance "solve" (the real solve is
CUDA-only), plus the cost-balanced partition used for bucketed variable-length batches."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stabletts_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _standin_solve(mu, mask, c, z):
    # per-utterance, batch-independent function (like the real path: no cross-sample op)
    return (z + 2.0 * mu) * mask + c.mean(dim=1)[:, None, None]


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        M, T, G = 8, 13, 16
        if rank == 0:
            mu, z = torch.randn(B, M, T), torch.randn(B, M, T)
            lens = torch.randint(1, T + 1, (B,))
            mask = (torch.arange(T)[None] < lens[:, None]).float().unsqueeze(1)
            c = torch.randn(B, G)
            out = shard.sharded_solve(_standin_solve, mu, mask, c, z, device=torch.device("cpu"))
            ok = torch.equal(out, _standin_solve(mu, mask, c, z))
            q.put(bool(ok))
        else:
            out = shard.sharded_solve(_standin_solve, None, None, None, None, device=torch.device("cpu"))
            assert out is None
    finally:
        dist.destroy_process_group()


def _run(world, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_sharded_solve_gloo_even():
    _run(2, 6)


def test_sharded_solve_gloo_ragged_and_tiny():
    _run(2, 5)      # 3 + 2
    _run(2, 1)      # one rank gets an empty slice


def test_split_counts_and_cost_partition():
    assert shard.split_counts(10, 4) == [3, 3, 2, 2]
    assert shard.split_counts(2, 8) == [1, 1, 0, 0, 0, 0, 0, 0]
    g = torch.Generator().manual_seed(3)
    lens = torch.randint(200, 2001, (128,), generator=g).tolist()
    parts = shard.partition_by_cost(lens, 8)
    assert sorted(i for p in parts for i in p) == list(range(128))
    loads = [sum(shard.utterance_cost(lens[i]) for i in p) for p in parts]
    assert max(loads) / min(loads) < 1.05
    for p in parts:
        assert [lens[i] for i in p] == sorted(lens[i] for i in p)


def test_length_buckets_and_bucketed_solve():
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(200, 2001, (128,), generator=g).tolist()
    bk = shard.length_buckets(lens, 4)
    assert sorted(i for b in bk for i in b) == list(range(128)) and len(bk) == 4
    mx = [max(lens[i] for i in b) for b in bk]
    assert mx == sorted(mx)
    padded = sum(len(b) * m for b, m in zip(bk, mx))
    assert padded < 0.75 * 128 * max(lens)                 # bucketing removes most of the padding waste
    # plumbing with a per-utterance stand-in solve
    B, M, T = 10, 4, 50
    ls = [50, 7, 33, 20, 45, 12, 50, 9, 27, 41]
    mask = (torch.arange(T)[None] < torch.tensor(ls)[:, None]).float().unsqueeze(1)
    mu, z, c = torch.randn(B, M, T) * mask, torch.randn(B, M, T), torch.randn(B, 6)
    # like the real solve, the stand-in leaves padded frames at the (unmasked) initial noise z
    f = lambda mu_, mask_, c_, z_: z_ + (2 * mu_ + c_.mean(1)[:, None, None]) * mask_
    out = shard.bucketed_solve(f, mu, mask, c, z, ls, n_buckets=3)
    assert torch.allclose(out, f(mu, mask, c, z))


# ---- property tests of the host-side partition logic (hypothesis) ------------------------------------------
from hypothesis import given, settings, strategies as hst   # noqa: E402

_lengths = hst.lists(hst.integers(min_value=1, max_value=4000), min_size=0, max_size=64)


@settings(max_examples=200, deadline=None)
@given(n=hst.integers(min_value=0, max_value=5000), world=hst.integers(min_value=1, max_value=16))
def test_split_counts_properties(n, world):
    from stabletts_b200 import shard
    c = shard.split_counts(n, world)
    assert len(c) == world and sum(c) == n
    assert max(c) - min(c) <= 1 and c == sorted(c, reverse=True)          # near-equal, larger slices first


@settings(max_examples=200, deadline=None)
@given(lengths=_lengths, world=hst.integers(min_value=1, max_value=8))
def test_partition_by_cost_properties(lengths, world):
    """every utterance lands in exactly one shard; shards are length-sorted; the greedy LPT bound holds:
    max load <= mean load + the largest single cost"""
    from stabletts_b200 import shard
    shards = shard.partition_by_cost(lengths, world)
    assert len(shards) == world
    flat = sorted(i for s in shards for i in s)
    assert flat == list(range(len(lengths)))
    for s in shards:
        assert [lengths[i] for i in s] == sorted(lengths[i] for i in s)
    if lengths:
        costs = [shard.utterance_cost(x) for x in lengths]
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) <= sum(costs) / world + max(costs) + 1e-6 * sum(costs)


@settings(max_examples=200, deadline=None)
@given(lengths=_lengths, n_buckets=hst.integers(min_value=1, max_value=8))
def test_length_buckets_properties(lengths, n_buckets):
    """buckets partition the indices, are contiguous in sorted-length order, and there are at most n_buckets"""
    from stabletts_b200 import shard
    b = shard.length_buckets(lengths, n_buckets)
    assert len(b) <= n_buckets
    flat = [i for g in b for i in g]
    assert sorted(flat) == list(range(len(lengths)))
    assert [lengths[i] for i in flat] == sorted(lengths)
    assert all(g for g in b)
