"""Per-kernel time of the alternate sampler by class of shape parameter (which elements make cfg4's lists slow?)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PGM_CUDA_NO_OVERLAP"] = "1"
import torch  # noqa: E402

import polyagamma_b200 as pg  # noqa: E402
from polyagamma_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(3)
n = 1 << 24
u = torch.rand(n, device=dev, dtype=torch.float64, generator=g)
v = torch.rand(n, device=dev, dtype=torch.float64, generator=g)
out = torch.empty(n, device=dev, dtype=torch.float64)
classes = {"h in [0.1, 0.3)": 0.1 + 0.2 * u, "h in [0.3, 1)": 0.3 + 0.7 * u, "h in [1, 2)": 1 + u, "h in [2, 4]": 2 + 2 * u,
           "h in (4, 5) chunk 3": 4.0001 + 0.9998 * u, "h in [5, 8) chunk 4": 5 + 3 * u}
for zname, z in (("|z| < 4", 8 * v - 4), ("4 < z < 14", 4 + 10 * v), ("14 < z < 50", 14 + 36 * v)):
    for name, h in classes.items():
        pg.random_polyagamma(h, z, out=out, method="alternate", random_state=(1, 1), disable_checks=True)
        torch.cuda.synchronize()
        _lib.profile_begin()
        pg.random_polyagamma(h, z, out=out, method="alternate", random_state=(1, 2), disable_checks=True)
        torch.cuda.synchronize()
        k = {a: b for a, b in _lib.profile_end().items() if b[1]}
        tot = sum(b[0] for b in k.values())
        print(f"{zname:12s} {name:22s} total {tot:7.2f} ms  {n / tot / 1e6:7.2f} G/s  " +
              "  ".join(f"{a} {b[0]:.2f}" for a, b in k.items() if b[0] > 0.02), flush=True)
