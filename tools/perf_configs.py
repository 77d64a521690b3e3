"""Device-resident throughput of every BASELINE.json config on one B200 (CUDA events, best of 3
after a warm-up), with the per-kernel time breakdown of each.  Writes gpurun_out/perf_configs.json.

    python tools/perf_configs.py [--n 1e8]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import polyagamma_b200 as pg  # noqa: E402
from polyagamma_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=float, default=1e8)
args = ap.parse_args()
N = int(args.n)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(20240)
u1 = torch.rand(N, device=dev, dtype=torch.float64, generator=g)
zU = torch.rand(N, device=dev, dtype=torch.float64, generator=g) * 100 - 50
zN = torch.randn(N, device=dev, dtype=torch.float64, generator=g) * 3
one = torch.ones(N, device=dev, dtype=torch.float64)
out = torch.empty(N, device=dev, dtype=torch.float64)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    kern = None
    for _ in range(reps):
        _lib.profile_begin()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        k = _lib.profile_end()
        t = e0.elapsed_time(e1)
        if t < best:
            best, kern = t, {a: round(b[0], 3) for a, b in k.items() if b[1]}
    return best, kern


res = {}


def run(name, n, fn, reps=3):
    ms, kern = timed(fn, reps)
    res[name] = {"n": n, "ms": ms, "per_s": n / (ms * 1e-3), "kernel_ms": kern}
    print(f"{name:28s} {n / (ms * 1e-3) / 1e9:8.3f} G/s  {ms:9.3f} ms  {kern}", flush=True)


rs = (12345, 0)
run("cfg1 fill h=1 z=0 (1e6)", 10**6, lambda: pg.random_polyagamma(1.0, 0.0, size=10**6, device=dev, random_state=rs))
run("cfg1 fill h=1 z=0", N, lambda: pg.random_polyagamma(1.0, 0.0, size=N, device=dev, random_state=rs))
run("cfg2 fill2 h=1(scalar) z~N(0,3)", N, lambda: pg.random_polyagamma(1.0, zN, out=out, random_state=rs, disable_checks=True))
run("cfg2 fill2 h=ones z~N(0,3)", N, lambda: pg.random_polyagamma(one, zN, out=out, random_state=rs, disable_checks=True))
run("cfg3 devroye h=1 z~U", N, lambda: pg.random_polyagamma(one, zU, out=out, method="devroye", random_state=rs, disable_checks=True))
hA = 1 + 3 * u1
run("cfg3 alternate h~U[1,4]", N, lambda: pg.random_polyagamma(hA, zU, out=out, method="alternate", random_state=rs, disable_checks=True))
hS = 10 + 90 * u1
run("cfg3 saddle h~U[10,100]", N, lambda: pg.random_polyagamma(hS, zU, out=out, method="saddle", random_state=rs, disable_checks=True))
M = N // 16
hG = 0.1 + 0.9 * u1[:M]
run("cfg3 gamma h~U[0.1,1]", M, lambda: pg.random_polyagamma(hG, zU[:M], out=out[:M], method="gamma", random_state=rs, disable_checks=True), reps=2)
hH = 0.1 + 199.9 * u1
run("cfg4 hybrid h~U[0.1,200]", N, lambda: pg.random_polyagamma(hH, zU, out=out, random_state=rs, disable_checks=True))
# cfg5a densities: 1e4 x 1e4 broadcast grid (x along rows; h, z along columns)
m = int(round(N ** 0.5))
x = torch.linspace(1e-3, 12.5, m, device=dev, dtype=torch.float64)[:, None]
hd = (0.5 + 24.5 * u1[:m])[None, :]
zd = (zU[:m] / 5)[None, :]
for name, f, log, frac in (("pdf", pg.polyagamma_pdf, False, 1), ("logpdf", pg.polyagamma_pdf, True, 1),
                           ("cdf", pg.polyagamma_cdf, False, 1), ("logcdf", pg.polyagamma_cdf, True, 8)):
    xx = x[:: frac]
    run(f"cfg5a {name} grid", xx.numel() * m, lambda: f(xx, hd, zd, return_log=log), reps=2)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "perf_configs.json"), "w"), indent=1)
