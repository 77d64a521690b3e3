"""Run ONE BASELINE config a few times (for ncu: `ncu -k regex:... python tools/profile_case.py cfg2`).

cases: cfg1 cfg2 cfg3_devroye cfg3_alternate cfg3_saddle cfg3_gamma cfg4 pdf logpdf cdf logcdf gibbs<p>
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import polyagamma_b200 as pg  # noqa: E402

case = sys.argv[1]
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1 << 25
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(20240)
u1 = torch.rand(n, device=dev, dtype=torch.float64, generator=g)
zU = torch.rand(n, device=dev, dtype=torch.float64, generator=g) * 100 - 50
zN = torch.randn(n, device=dev, dtype=torch.float64, generator=g) * 3
out = torch.empty(n, device=dev, dtype=torch.float64)
rs = (12345, 0)
kw = dict(out=out, random_state=rs, disable_checks=True)
if case.startswith("gibbs"):
    from polyagamma_b200.gibbs import logistic_omega_step

    p = int(case[5:])
    rows = n * 8 // p
    X = torch.randn(rows, p, device=dev, dtype=torch.float64, generator=g)
    beta = torch.randn(p, device=dev, dtype=torch.float64, generator=g) * (3.0 / p**0.5)
    for _ in range(reps):
        logistic_omega_step(X, beta, 1.0, random_state=rs)
    torch.cuda.synchronize()
    print("done", case, rows)
    sys.exit(0)
m = int(round(n ** 0.5))
gx = torch.linspace(1e-3, 12.5, m, device=dev, dtype=torch.float64)[:, None]
gh = (0.5 + 24.5 * u1[:m])[None, :]
gz = (zU[:m] / 5)[None, :]
run = {
    "cfg1": lambda: pg.random_polyagamma(1.0, 0.0, size=n, device=dev, random_state=rs),
    "cfg2": lambda: pg.random_polyagamma(1.0, zN, **kw),
    "cfg2_ones": lambda: pg.random_polyagamma(torch.ones_like(zN), zN, **kw),
    "cfg3_devroye": lambda: pg.random_polyagamma(torch.ones_like(zU), zU, method="devroye", **kw),
    "cfg3_alternate": lambda: pg.random_polyagamma(1 + 3 * u1, zU, method="alternate", **kw),
    "cfg3_saddle": lambda: pg.random_polyagamma(10 + 90 * u1, zU, method="saddle", **kw),
    "cfg3_gamma": lambda: pg.random_polyagamma(0.1 + 0.9 * u1[: n // 16], zU[: n // 16], method="gamma",
                                               out=out[: n // 16], random_state=rs, disable_checks=True),
    "cfg4": lambda: pg.random_polyagamma(0.1 + 199.9 * u1, zU, **kw),
    # cfg5a: x along rows, (h, z) along columns of a sqrt(n) x sqrt(n) broadcast grid
    "pdf": lambda: pg.polyagamma_pdf(gx, gh, gz),
    "logpdf": lambda: pg.polyagamma_pdf(gx, gh, gz, return_log=True),
    "cdf": lambda: pg.polyagamma_cdf(gx, gh, gz),
    "logcdf": lambda: pg.polyagamma_cdf(gx, gh, gz, return_log=True),
}[case]
for _ in range(reps):
    run()
torch.cuda.synchronize()
print("done", case, n)
