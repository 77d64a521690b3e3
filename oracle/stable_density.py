"""oracle/stable_density.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy, complex128) of the cancellation-free pdf / cdf of PG(h, z) that the CUDA density kernels
offer as which = 4..7 (csrc/pgm_density.cu pg_contour): the Laplace inversion integral of
E exp(-tX) = [cosh(z/2) / cosh sqrt(z^2/4 + t/2)]^h along the vertical line through its saddle point, by the
trapezoid rule.  It is pinned by tests/golden/stable_density_golden.npz -- the reference's own alternating series
(src/pgm_density.c:67-92,236-280) evaluated in 300-digit arithmetic, tests/golden/make_stable_golden.py.
"""
import numpy as np


def _logcosh(r):
    r = np.where(r.real < 0, -r, r)
    return r - np.log(2.0) + np.log1p(np.exp(-2 * r))


def _saddle(xp):
    """w = r^2 with tanh(sqrt w)/sqrt w = xp (tan for w < 0), by bisection."""
    def g(w):
        if abs(w) < 1e-8:
            return 1 - w / 3
        if w > 0:
            r = np.sqrt(w)
            return np.tanh(r) / r
        r = np.sqrt(-w)
        return np.tan(r) / r
    lo, hi = -np.pi ** 2 / 4 + 1e-14, (1.0 / xp + 1.0) ** 2
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if g(mid) > xp:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def pdf_cdf(x, h, z, n=None, step=0.2):
    """(pdf, cdf, 1 - cdf) of PG(h, z) at x > 0, h >= 8."""
    z = abs(float(z))
    if n is None:
        n = 192 if h >= 16 else 1024  # (the integrand decays like exp(-h sqrt(y) / 2): slowly for small h)
    w0 = _saddle(4.0 * x / h)
    t0 = 2 * w0 - z * z / 2

    def psi(t):
        r = np.sqrt(z * z / 4 + t / 2 + 0j)
        return t * x + h * (_logcosh(np.array(z / 2 + 0j)) - _logcosh(r))

    e = 1e-4 * max(1.0, abs(t0))
    sig = np.sqrt(((psi(t0 + e) - 2 * psi(t0) + psi(t0 - e)) / e ** 2).real)
    dy = step / sig
    y = dy * np.arange(0, n)
    wts = np.ones(n)
    wts[0] = 0.5
    pdf = np.exp(psi(t0).real) * dy / np.pi * np.sum(wts * np.exp(psi(t0 + 1j * y) - psi(t0).real).real)
    cmin = 1.1 / sig
    c = t0 if abs(t0) >= cmin else cmin
    t = c + 1j * y
    integral = np.exp(psi(c).real) * dy / np.pi * np.sum(wts * (np.exp(psi(t) - psi(c).real) / t).real)
    if c > 0:
        return float(pdf), float(integral), float(1.0 - integral)
    return float(pdf), float(1.0 + integral), float(-integral)
