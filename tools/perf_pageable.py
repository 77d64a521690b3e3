"""End-to-end throughput of random_polyagamma with PAGEABLE numpy operands (what a user of the reference has)
against pinned ones (what bench.py's e2e uses).  cfg4-like operands, n elements."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import polyagamma_b200 as pg  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
rng = np.random.default_rng(1)
h = rng.uniform(0.1, 200.0, n)
z = rng.uniform(-50.0, 50.0, n)
out = np.empty(n)
out[:] = 0.0  # touched: no first-touch cost inside the timing
hp, zp, op = (torch.from_numpy(a.copy()).pin_memory() for a in (h, z, out))


def timed(f, reps=3):
    f()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        best = min(best, time.perf_counter() - t0)
    return best


rows = {
    "pinned h, z, out=": lambda: pg.random_polyagamma(hp.numpy(), zp.numpy(), out=op.numpy(), random_state=(1, 2), disable_checks=True),
    "pageable h, z, out= (touched)": lambda: pg.random_polyagamma(h, z, out=out, random_state=(1, 2), disable_checks=True),
    "pageable h, z, new result array": lambda: pg.random_polyagamma(h, z, random_state=(1, 2), disable_checks=True),
}
for name, f in rows.items():
    t = timed(f)
    print(f"n={n:.0e}  {name:34s} {t * 1e3:8.1f} ms  {n / t / 1e9:6.3f} G samples/s  ({24 * n / t / 1e9:5.1f} GB/s over the link)", flush=True)
a = pg.random_polyagamma(h, z, random_state=(1, 2), disable_checks=True)
b = pg.random_polyagamma(hp.numpy(), zp.numpy(), out=op.numpy(), random_state=(1, 2), disable_checks=True)
print("pageable == pinned:", bool(np.array_equal(a, b)))
