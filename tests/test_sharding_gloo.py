"""CPU, world_size 2 over gloo: the N>1 path is "shard the index range, no collective".

Each rank computes its block with shard_range + first_index (here through the CPU oracle, which
implements the same (seed, offset, index) stream contract as the kernels); the blocks gathered
over gloo must equal the single-process result bit for bit, for 2 ranks and for an uneven
3-way split."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 5001


def _inputs():
    rng = np.random.default_rng(42)
    return rng.uniform(0.1, 120.0, N), rng.uniform(-30.0, 30.0, N)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from polyagamma_b200.sharding import shard_range

    dist.init_process_group("gloo", rank=rank, world_size=world)
    h, z = _inputs()
    lo, hi = shard_range(N, rank, world)
    mine = O.fill2(2024, 5, h[lo:hi], z[lo:hi], first_index=lo)
    # pad to a common length for all_gather
    width = (N + world - 1) // world
    buf = torch.zeros(width, dtype=torch.float64)
    buf[: hi - lo] = torch.from_numpy(mine)
    gathered = [torch.zeros(width, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, buf)
    # a reduced statistic, the only thing a collective is ever needed for on this path
    s = torch.tensor([mine.sum(), float(hi - lo)], dtype=torch.float64)
    dist.all_reduce(s)
    if rank == 0:
        parts = []
        for r in range(world):
            a, b = shard_range(N, r, world)
            parts.append(gathered[r][: b - a].numpy())
        q.put((np.concatenate(parts), s.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_two_rank_shards_equal_single_process(oracle):
    h, z = _inputs()
    whole = oracle.fill2(2024, 5, h, z)
    got, stats = _run(2)
    assert np.array_equal(got, whole)
    assert stats[1] == N and np.isclose(stats[0], whole.sum(), rtol=1e-12)


def test_uneven_three_rank_split(oracle):
    h, z = _inputs()
    whole = oracle.fill2(2024, 5, h, z)
    got, _ = _run(3)
    assert np.array_equal(got, whole)


def _gram_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    from polyagamma_b200.gibbs import allreduce_gram
    from polyagamma_b200.sharding import shard_range

    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    X = torch.from_numpy(rng.normal(size=(1003, 5)))
    w = torch.from_numpy(rng.uniform(0.1, 2.0, 1003))
    lo, hi = shard_range(1003, rank, world)
    part = X[lo:hi].T @ (w[lo:hi, None] * X[lo:hi])
    total = allreduce_gram(part.clone())
    if rank == 0:
        q.put((total.numpy(), (X.T @ (w[:, None] * X)).numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_gram_sums_row_shards():
    """The one collective next to the path: the p x p partial Gram matrices of a row-sharded Gibbs sweep."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gram_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    got, want = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)


def test_allreduce_gram_is_identity_without_a_process_group():
    sys.path.insert(0, ROOT)
    from polyagamma_b200.gibbs import allreduce_gram

    g = torch.arange(9.0, dtype=torch.float64).reshape(3, 3)
    assert torch.equal(allreduce_gram(g.clone()), g)
