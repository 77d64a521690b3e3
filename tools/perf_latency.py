"""Per-call latency of small fills (the Gibbs-sampler regime: one call per sweep on 1e3..1e6 elements).
Host wall-clock per call, device-resident inputs and output, averaged over 200 calls after warm-up."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import polyagamma_b200 as pg  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)
for n in (1_000, 10_000, 100_000, 250_000, 1_000_000):
    h = torch.rand(n, device=dev, dtype=torch.float64, generator=g) * 199.9 + 0.1
    z = torch.randn(n, device=dev, dtype=torch.float64, generator=g) * 3
    ones = torch.ones(n, device=dev, dtype=torch.float64)
    out = torch.empty(n, device=dev, dtype=torch.float64)
    cases = {"hybrid h~U[0.1,200]": lambda k: pg.random_polyagamma(h, z, out=out, random_state=(1, k), disable_checks=True),
             "hybrid h=1 (cfg2)": lambda k: pg.random_polyagamma(ones, z, out=out, random_state=(1, k), disable_checks=True),
             "hybrid scalar h=1, z array": lambda k: pg.random_polyagamma(1.0, z, out=out, random_state=(1, k), disable_checks=True),
             "scalar h=1 z=0 (cfg1)": lambda k: pg.random_polyagamma(1.0, 0.0, out=out, random_state=(1, k))}
    for name, f in cases.items():
        for k in range(20):
            f(k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(200):
            f(k)
        torch.cuda.synchronize()
        us = (time.perf_counter() - t0) / 200 * 1e6
        print(f"n={n:>9,d}  {name:27s} {us:8.1f} us/call  {n / us / 1e3:8.3f} G samples/s", flush=True)
