"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled in oracle/_ref.

Run in the build container (needs /root/reference to have been compiled with
`make -C oracle ref`):   python tests/golden/make_golden.py

density_golden.npz   x, h, z grids and the reference's pdf / cdf / logpdf / logcdf on them
sampler_golden.npz   per (method, h, z) cell: 1001 quantiles of 4e6 reference draws
                     (numpy PCG64 seeds recorded), for two-sample checks without the reference
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref  # noqa: E402

assert ref.available(), "build the reference first: make -C oracle ref"

# ---------------------------------------------------------------- densities
rng = np.random.default_rng(20240)
n = 6000
x = np.concatenate([rng.uniform(1e-3, 12.5, n - 40), 10.0 ** rng.uniform(-6, 2, 40)])
h = np.concatenate([rng.uniform(0.1, 25.0, n - 400), rng.integers(1, 30, 400).astype(float)])
z = rng.uniform(-25.0, 25.0, n)
z[::11] = 0.0
rng.shuffle(h)
dens = {"x": x, "h": h, "z": z}
for name in ("pdf", "cdf", "logpdf", "logcdf"):
    with np.errstate(all="ignore"):
        dens[name] = ref.c_density(name, x, h, z)
np.savez_compressed(os.path.join(HERE, "density_golden.npz"), **dens)
print("density_golden.npz", {k: v.shape for k, v in dens.items()})

# ---------------------------------------------------------------- sampler quantiles
CELLS = [
    ("devroye", 1.0, 0.0), ("devroye", 1.0, 2.5), ("devroye", 2.0, -5.0), ("devroye", 4.0, 1.0),
    ("alternate", 1.0, 0.0), ("alternate", 2.5, 0.0), ("alternate", 3.5, 1.0), ("alternate", 0.5, 2.0),
    ("alternate", 2.7, 30.0), ("alternate", 6.5, 10.0),
    ("saddle", 4.5, 0.0), ("saddle", 10.0, 2.0), ("saddle", 30.0, 5.0), ("saddle", 10.0, 20.0),
    ("saddle", 80.0, 1.0),
    ("hybrid", 60.0, 5.0), ("hybrid", 120.0, 0.0), ("hybrid", 3.0, 0.5), ("hybrid", 3.0, 2.5),
    ("gamma", 1.0, 1.0), ("gamma", 0.5, 3.0),
]
NREF = 4_000_000
Q = 1001
probs = (np.arange(Q) + 0.5) / Q
cells = np.zeros((len(CELLS), 3))
quant = np.zeros((len(CELLS), Q))
nref = np.zeros(len(CELLS), dtype=np.int64)
names = []
for i, (method, hh, zz) in enumerate(CELLS):
    nn = NREF if method != "gamma" else 200_000
    draws = ref.random_polyagamma(hh, zz, size=nn, method=None if method == "hybrid" else method,
                                  random_state=np.random.Generator(np.random.PCG64(1000 + i)))
    draws.sort()
    # empirical quantile at p: order statistic ceil(p n) - 1
    idx = np.minimum(np.ceil(probs * nn).astype(np.int64) - 1, nn - 1)
    quant[i] = draws[idx]
    cells[i] = (ref.METHODS[method], hh, zz)
    nref[i] = nn
    names.append(method)
    print(method, hh, zz, draws.mean(), flush=True)
np.savez_compressed(os.path.join(HERE, "sampler_golden.npz"), cells=cells, quantiles=quant, probs=probs,
                    nref=nref, methods=np.array(names))
print("sampler_golden.npz written")
