"""Row-sharded Gibbs omega-step over NCCL: run under torchrun with N ranks (one per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/gibbs_sharded_check.py

Every rank builds the same global X (seeded), takes its block of rows, calls sharded_logistic_omega_step and
checks: the all-reduced Gram matrix equals the single-device one to 1e-12 relative, every rank holds the same
bits, and the concatenated omegas are bit-identical to the single-device draw.  Rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from polyagamma_b200.gibbs import logistic_omega_step, sharded_logistic_omega_step  # noqa: E402
from polyagamma_b200.sharding import shard_range  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
out = {"world": world, "cases": []}
for n, p in ((1_000_003, 8), (4_000_001, 32), (2_000_000, 64)):
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    X = torch.randn(n, p, device=dev, dtype=torch.float64, generator=g)
    beta = torch.randn(p, device=dev, dtype=torch.float64, generator=g) * (2.0 / p**0.5)
    G1, om1 = logistic_omega_step(X, beta, 1.0, random_state=(42, 3))
    lo, hi = shard_range(n, rank, world)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    G, om = sharded_logistic_omega_step(X[lo:hi], beta, 1.0, row_offset=lo, random_state=(42, 3))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0.record()
    for k in range(5):
        G, om = sharded_logistic_omega_step(X[lo:hi], beta, 1.0, row_offset=lo, random_state=(42, 3))
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1) / 5], device=dev)
    rel = ((G - G1).abs().max() / G1.abs().max()).reshape(1)
    same_omega = torch.tensor([float(torch.equal(om, om1[lo:hi]))], device=dev)
    gbits = G.view(torch.int64).sum().reshape(1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rel, op=dist.ReduceOp.MAX)
        dist.all_reduce(same_omega, op=dist.ReduceOp.MIN)
        parts = [torch.zeros_like(gbits) for _ in range(world)]
        dist.all_gather(parts, gbits)
        same_bits = all(int(q.item()) == int(parts[0].item()) for q in parts)
    else:
        same_bits = True
    ok = bool(rel.item() < 1e-12) and bool(same_omega.item() == 1.0) and same_bits
    out["cases"].append({"n": n, "p": p, "gram_rel_err_vs_single_device": float(rel.item()),
                         "omega_bit_identical": bool(same_omega.item() == 1.0), "gram_same_bits_on_all_ranks": same_bits,
                         "ms_per_sweep_max_over_ranks": float(ms.item()), "rows_per_s": n / (float(ms.item()) * 1e-3), "ok": ok})
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if all(c["ok"] for c in out["cases"]) else 1)
