"""Small invocation of every kernel family of libpgm_cuda.so, meant to run under compute-sanitizer
(tools/sanitize.sh: memcheck, racecheck, synccheck, initcheck).  Sizes are tiny because the tools slow a
kernel down by one to two orders of magnitude; every code path of the launcher is still taken:
the one-launch small-fill kernel, the bucketed pipeline on the caller's stream and on the helper
streams (PGM_CUDA_MULTI_MIN_LOG2), every explicit method, the gamma convolution, the host-buffer entry
point, the density kernels and the Gibbs omega-step (HBM-bound and DMMA Gram kernels)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import polyagamma_b200 as pg  # noqa: E402
from polyagamma_b200.gibbs import logistic_omega_step  # noqa: E402


def main():
    rng = np.random.default_rng(7)
    n = int(os.environ.get("SANITIZE_N", "40000"))
    h = rng.uniform(0.1, 200.0, n)
    h[::5] = rng.integers(1, 5, h[::5].shape)
    h[7] = -1.0  # invalid shape -> NaN written by the classify pass
    h[9] = np.inf
    z = rng.uniform(-50.0, 50.0, n)
    z[::11] = 0.0
    hd, zd = torch.from_numpy(h).cuda(), torch.from_numpy(z).cuda()
    # bucketed pipeline (n > small-fill limit is forced by PGM_CUDA_SMALL_MAX=0 in sanitize.sh's second pass)
    out = pg.random_polyagamma(hd, zd, random_state=(1, 0), disable_checks=True)
    assert torch.isnan(out[7]) and torch.isinf(out[9])
    assert torch.isfinite(out[torch.isfinite(hd) & (hd > 0)]).all()
    # the one-launch small-fill kernel, also when sanitize.sh disables it for the other calls (read per call)
    saved = os.environ.pop("PGM_CUDA_SMALL_MAX", None)
    pg.random_polyagamma(hd[:5000], zd[:5000], random_state=(9, 0), disable_checks=True)
    if saved is not None:
        os.environ["PGM_CUDA_SMALL_MAX"] = saved
    # ragged tail + unaligned views (scalar-load classify path) + broadcasting
    pg.random_polyagamma(hd[1:1001], zd[3:1003], random_state=(2, 0), disable_checks=True)
    pg.random_polyagamma(hd[:64].reshape(64, 1).abs() + 0.1, zd[:33].reshape(1, 33), random_state=(3, 0))
    pg.random_polyagamma(2.5, 1.25, size=(17, 19), random_state=(4, 0))
    # every explicit method (devroye needs integer h)
    small = slice(0, 6000)
    hh = hd[small].abs().clamp(0.1, 60.0)
    hh[7] = 1.0
    hh[9] = 2.0
    for m in ("alternate", "saddle", "gamma"):
        r = pg.random_polyagamma(hh, zd[small], method=m, random_state=(5, 0))
        assert torch.isfinite(r).all(), m
    r = pg.random_polyagamma(torch.ceil(hh.clamp(max=8.0)), zd[small] * 0.2, method="devroye", random_state=(6, 0))
    assert torch.isfinite(r).all()
    r = pg.random_polyagamma(hh, zd[small], ref_compat=True, random_state=(6, 1))
    assert torch.isfinite(r).all()
    # host buffers -> pgm_cuda_random_polyagamma_fill2_host (slots on three streams)
    ho = np.empty(n)
    hv = np.abs(h) + 0.1
    hv[9] = 3.0
    pg.random_polyagamma(hv, z, out=ho, random_state=(7, 0))
    assert np.isfinite(ho).all()
    # densities
    x = rng.uniform(0.01, 5.0, 2048)
    dh = rng.uniform(0.5, 8.0, 2048)
    dz = rng.uniform(-8.0, 8.0, 2048)
    for f in (pg.polyagamma_pdf, pg.polyagamma_cdf):
        for log in (False, True):
            v = f(x, dh, dz, return_log=log)
            assert not np.isnan(v).any()
    # Gibbs omega-step: HBM-bound Gram kernels (p <= 32) and the DMMA kernel (p = 64), ragged p
    for p in (3, 8, 32, 50, 64):
        rows = 3001
        X = torch.from_numpy(rng.normal(size=(rows, p))).cuda()
        beta = torch.from_numpy(rng.normal(size=p) * 0.3).cuda()
        G, om = logistic_omega_step(X, beta, 1.0, random_state=(8, p))
        want = X.T @ (om[:, None] * X)
        err = ((G - want).abs().max() / want.abs().max()).item()
        assert err < 1e-12, (p, err)
        assert torch.equal(G, G.T)
    torch.cuda.synchronize()
    pg.release_scratch()
    print(f"sanitize case ok: {pg.launch_count()} kernel launches")


if __name__ == "__main__":
    main()
