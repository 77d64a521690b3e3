"""Golden values for the cancellation-free ("stable") pdf / cdf of PG(h, z) at large h.

The oracle is the reference's own alternating series (src/pgm_density.c:67-92, 236-280)
evaluated with mpmath in 300-digit arithmetic, where its cancellation (cond up to 1e240 at h = 200) is harmless.
Points: x = mean + k sd for k in K, over a grid of (h, z).  Writes tests/golden/stable_density_golden.npz
(x, h, z, pdf, cdf, sf = 1 - cdf, all float64).  Takes a few minutes.

    python tests/golden/make_stable_golden.py
"""
import math
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 300
EPS = mp.mpf(10) ** (-290)


def series(x, h, z):
    x, h, z = mp.mpf(x), mp.mpf(h), abs(mp.mpf(z))
    gh = mp.gamma(h)
    s_pdf = s_cdf = mp.mpf(0)
    done_pdf = done_cdf = False
    for n in range(0, 6000):
        c = mp.gamma(n + h) / (gh * mp.factorial(n))
        a = 2 * n + h
        sign = -1 if n & 1 else 1
        if not done_pdf:
            t = c * a * mp.exp(-a * a / (8 * x))
            s_pdf += sign * t
            done_pdf = n > 50 and t < EPS * abs(s_pdf)
        if not done_cdf:
            if z == 0:
                t = c * mp.erfc(a / (2 * mp.sqrt(2 * x)))
            else:
                mu, lam = a / (2 * z), a * a / 4
                sq = mp.sqrt(lam / x)
                t = c * mp.exp(-n * z) * (mp.ncdf(sq * (x / mu - 1)) + mp.exp(2 * lam / mu) * mp.ncdf(-sq * (x / mu + 1)))
            s_cdf += sign * t
            done_cdf = n > 50 and t < EPS * abs(s_cdf)
        if done_pdf and done_cdf:
            break
    pdf = mp.cosh(z / 2) ** h * mp.exp(-z * z * x / 2) * 2 ** (h - 1) / mp.sqrt(2 * mp.pi * x ** 3) * s_pdf
    cdf = ((1 + mp.exp(-z)) ** h if z != 0 else mp.mpf(2) ** h) * s_cdf
    return pdf, cdf


rows = []
for h in (8, 16, 40, 100, 200):
    for z in (0.0, 1.0, 5.0, 20.0):
        m = h * (math.tanh(z / 2) / (2 * z) if z else 0.25)
        v = h / 24 if z == 0 else h * (math.sinh(z) - z) / math.cosh(z / 2) ** 2 / (4 * z ** 3)
        for k in (-3.0, -1.0, 0.0, 0.5, 2.0, 4.0):
            x = m + k * math.sqrt(v)
            if x <= 0:
                continue
            pdf, cdf = series(x, h, z)
            rows.append((x, h, z, float(pdf), float(cdf), float(1 - cdf)))
            print(rows[-1], flush=True)
a = np.array(rows)
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "stable_density_golden.npz"), x=a[:, 0], h=a[:, 1], z=a[:, 2],
         pdf=a[:, 3], cdf=a[:, 4], sf=a[:, 5])
print(len(rows), "points")
