"""Host-link roof of the end-to-end path: concurrent pinned H2D + D2H per GPU when N ranks of one box copy at
once (the pattern of pgm_cuda_random_polyagamma_fill2_host: 16 B in + 8 B out per element, 64 MiB chunks on
three streams).  Run under torchrun with N ranks, or plain for N = 1; rank 0 merges the result into
profiles/r2_h2d_peak.json under the key str(N).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_peak.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
CH = 1 << 23  # doubles per chunk
NCH = 12
hin = [torch.empty(2 * CH, dtype=torch.float64, pin_memory=True).fill_(1.0) for _ in range(3)]
hout = [torch.empty(CH, dtype=torch.float64, pin_memory=True) for _ in range(3)]
din = [torch.empty(2 * CH, dtype=torch.float64, device=dev) for _ in range(3)]
dout = [torch.zeros(CH, dtype=torch.float64, device=dev) for _ in range(3)]
streams = [torch.cuda.Stream(dev) for _ in range(3)]


def run(h2d, d2h):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])
    t0 = time.perf_counter()
    for c in range(NCH):
        s = streams[c % 3]
        with torch.cuda.stream(s):
            if h2d:
                din[c % 3].copy_(hin[c % 3], non_blocking=True)
            if d2h:
                hout[c % 3].copy_(dout[c % 3], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt


res = {}
for name, h2d, d2h, nbytes in (("h2d_gbs_per_gpu", True, False, 16 * CH), ("d2h_gbs_per_gpu", False, True, 8 * CH),
                               ("bidir_gbs_per_gpu", True, True, 24 * CH)):
    run(h2d, d2h)
    best = min(run(h2d, d2h) for _ in range(3))
    res[name] = nbytes * NCH / best / 1e9
if rank == 0:
    res["ranks"] = world
    res["host_cores"] = len(os.sched_getaffinity(0))
    res["how"] = "12 chunks of 128 MiB H2D + 64 MiB D2H on three streams per rank, pinned memory, best of 3, max over ranks"
    path = os.path.join(ROOT, "gpurun_out", "r2_h2d_peak.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    try:
        allres = json.load(open(path))
    except Exception:
        allres = {}
    allres[str(world)] = res
    json.dump(allres, open(path, "w"), indent=1)
    print(json.dumps({str(world): res}))
if world > 1:
    dist.destroy_process_group()
