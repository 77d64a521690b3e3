"""Per-call latency of small fills (the Gibbs-sampler regime: one call per sweep on 1e3..1e6 elements).
Host wall-clock per call, device-resident inputs and output, averaged over 200 calls after warm-up."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import polyagamma_b200 as pg  # noqa: E402

import ctypes  # noqa: E402

import numpy as np  # noqa: E402

from polyagamma_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)

# ---- n = 1: the reference quotes 1.69 us per scalar draw with a passed generator (README.md:224-232) ----
L = _lib.lib()
for name, f in (("C-ABI pgm_cuda_random_polyagamma(seed, k, 1.0, 0.0, hybrid)", lambda k: L.pgm_cuda_random_polyagamma(7, k, 1.0, 0.0, 4)),
                ("Python random_polyagamma(1.0, 0.0, random_state=(7, k))", lambda k: pg.random_polyagamma(1.0, 0.0, random_state=(7, k))),
                ("Python polyagamma_pdf(0.3, 2.0, 1.0) (scalar)", lambda k: pg.polyagamma_pdf(0.3, 2.0, 1.0))):
    for k in range(200):
        f(k)
    t0 = time.perf_counter()
    for k in range(2000):
        f(k)
    print(f"n=        1  {name:62s} {(time.perf_counter() - t0) / 2000 * 1e6:8.2f} us/call", flush=True)
x1, h1, z1, o1 = (np.array([v]) for v in (0.3, 2.0, 1.0, 0.0))
for which, nm in ((0, "pdf"), (3, "logcdf")):
    for k in range(200):
        L.pgm_cuda_polyagamma_density_host(which, x1.ctypes.data, 0, h1.ctypes.data, 0, z1.ctypes.data, 0, 1, o1.ctypes.data, -1)
    t0 = time.perf_counter()
    for k in range(2000):
        L.pgm_cuda_polyagamma_density_host(which, x1.ctypes.data, 0, h1.ctypes.data, 0, z1.ctypes.data, 0, 1, o1.ctypes.data, -1)
    print(f"n=        1  C-ABI pgm_cuda_polyagamma_density_host({nm}, one point){' ' * 22} {(time.perf_counter() - t0) / 2000 * 1e6:8.2f} us/call",
          flush=True)

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
