import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _cuda_available():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref); skipped when it has not been built."""
    import ref as R

    if not R.available():
        pytest.skip("oracle/_ref not built (make -C oracle ref)")
    return R


@pytest.fixture(scope="session")
def density_golden():
    return dict(np.load(os.path.join(GOLDEN, "density_golden.npz")))


@pytest.fixture(scope="session")
def sampler_golden():
    return dict(np.load(os.path.join(GOLDEN, "sampler_golden.npz")))


# ---- shared statistical helpers -------------------------------------------------------------
def pg_moments(h, z):
    """Exact mean and variance of PG(h, z) (same formulas as src/pgm_random.c:53-61)."""
    z = abs(float(z))
    if z == 0.0:
        return h / 4.0, h / 24.0
    m = h * np.tanh(z / 2.0) / (2.0 * z)
    if z < 1e-2:
        v = h / 24.0 * (1.0 - z * z / 5.0)
    else:
        v = h * (np.sinh(z) - z) / np.cosh(z / 2.0) ** 2 / (4.0 * z**3) if z < 300 else h * 0.5 / z**3
    return m, v


def gamma_sum_moments(h, z):
    """Moments of the 200-term truncated gamma convolution (src/pgm_random.c:79-97)."""
    k = np.arange(200)
    d = np.pi**2 * (k + 0.5) ** 2 + z * z / 4.0
    return 0.5 * h * np.sum(1.0 / d), 0.25 * h * np.sum(1.0 / d**2)


def moment_zscores(x, mean, var):
    n = x.size
    zm = (x.mean() - mean) / np.sqrt(var / n)
    m4 = np.mean((x - x.mean()) ** 4)
    zv = (x.var() - var) / np.sqrt(max(m4 - var * var, 1e-300) / n)
    return zm, zv


def ks_vs_quantiles(x, quantiles, probs, nref):
    """sqrt(N_eff) * sup |F_x(q_k) - p_k| against stored reference quantiles."""
    xs = np.sort(x)
    F = np.searchsorted(xs, quantiles, side="right") / xs.size
    d = np.max(np.abs(F - probs))
    neff = xs.size * nref / (xs.size + nref)
    return d * np.sqrt(neff)
