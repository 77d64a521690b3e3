"""Call latency of fills between 5e4 and 4e6 elements through the three launch paths:
one-launch kernel (small), tile kernel + drain kernel (mid), bucketed pipeline (pipe).
Host wall-clock per call over 200 calls after warm-up, operands and output resident on the device."""
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
MODES = {"small": {"PGM_CUDA_SMALL_MAX": str(1 << 22)}, "mid": {"PGM_CUDA_SMALL_MAX": "0", "PGM_CUDA_MID_MAX": str(1 << 25)},
         "pipe": {"PGM_CUDA_SMALL_MAX": "0", "PGM_CUDA_MID_MAX": "0"}}
print(f"{'n':>10s}  {'case':28s}" + "".join(f"{m:>10s}" for m in MODES) + "   (us per call)")
for n in (50_000, 100_000, 250_000, 500_000, 1_000_000, 2_000_000, 4_000_000):
    h = torch.rand(n, device=dev, dtype=torch.float64, generator=g) * 199.9 + 0.1
    h14 = torch.rand(n, device=dev, dtype=torch.float64, generator=g) * 3 + 1
    z = torch.randn(n, device=dev, dtype=torch.float64, generator=g) * 3
    zu = torch.rand(n, device=dev, dtype=torch.float64, generator=g) * 100 - 50
    ones = torch.ones(n, device=dev, dtype=torch.float64)
    out = torch.empty(n, device=dev, dtype=torch.float64)
    cases = {"hybrid h~U[0.1,200] z~U[-50,50]": lambda k: pg.random_polyagamma(h, zu, out=out, random_state=(1, k), disable_checks=True),
             "hybrid h~U[0.1,200] z~N(0,3)": lambda k: pg.random_polyagamma(h, z, out=out, random_state=(1, k), disable_checks=True),
             "hybrid h=1 z~N(0,3) (cfg2)": lambda k: pg.random_polyagamma(ones, z, out=out, random_state=(1, k), disable_checks=True),
             "hybrid h~U[1,4] z~N(0,3)": lambda k: pg.random_polyagamma(h14, z, out=out, random_state=(1, k), disable_checks=True)}
    for name, f in cases.items():
        row = []
        for mode, env in MODES.items():
            if mode == "small" and n > 1_000_000:
                row.append(float("nan"))
                continue
            for k in ("PGM_CUDA_SMALL_MAX", "PGM_CUDA_MID_MAX"):
                os.environ.pop(k, None)
            os.environ.update(env)
            for k in range(20):
                f(k)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(200):
                f(k)
            torch.cuda.synchronize()
            row.append((time.perf_counter() - t0) / 200 * 1e6)
        print(f"{n:>10,d}  {name:28s}" + "".join(f"{v:10.1f}" for v in row), flush=True)
