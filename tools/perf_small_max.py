"""Where does the one-launch small_kernel stop paying?  Mixed-method fills of n elements with the threshold
forced above / below n (PGM_CUDA_SMALL_MAX)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import polyagamma_b200 as pg
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
for n in (50_000, 100_000, 200_000, 400_000, 1_000_000):
    h = torch.rand(n, device=dev, dtype=torch.float64, generator=g) * 199.9 + 0.1
    z = torch.rand(n, device=dev, dtype=torch.float64, generator=g) * 100 - 50
    out = torch.empty(n, device=dev, dtype=torch.float64)
    for label, env in (("small_kernel", str(1 << 22)), ("pipeline", "0")):
        os.environ["PGM_CUDA_SMALL_MAX"] = env
        for k in range(20):
            pg.random_polyagamma(h, z, out=out, random_state=(1, k), disable_checks=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(200):
            pg.random_polyagamma(h, z, out=out, random_state=(1, k), disable_checks=True)
        torch.cuda.synchronize()
        print(f"n={n:>9,d} cfg4 mix  {label:12s} {(time.perf_counter() - t0) / 200 * 1e6:8.1f} us/call", flush=True)
