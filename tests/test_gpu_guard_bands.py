"""Guard bands: no entry point of the C-ABI writes outside the n doubles it was given.

compute-sanitizer is not available on the GPU pool, so out-of-bounds stores are looked for directly:
every output pointer handed to libpgm_cuda.so points into the middle of a buffer whose other
elements hold a sentinel bit pattern (a NaN payload no kernel produces); after the call the bands on
both sides must be bit-identical and the n elements in between must all have been written.  Sizes sit on
either side of every tiling constant of the launcher (8 elements per thread of the classify pass, its
2048-element tiles, 64-element warp reservations, the 2^16-element one-launch limit, pass boundaries).
Inputs get the same treatment against out-of-bounds *reads* changing a result: h and z are views into
larger buffers whose surroundings hold values that would be drawn very differently (the outputs are
compared with a call on exact-size copies)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SENTINEL = np.array([0x7FF8DEADBEEF1234], dtype=np.uint64).view(np.float64)[0]
BAND = 4096
SIZES = [1, 7, 8, 9, 63, 65, 2047, 2049, 4097, 65535, 65537, 100_003]


def _torch():
    import torch

    return torch


def _banded(n, dev="cuda", shift=0):
    """shift = 1 makes the output 8- but not 16-byte aligned"""
    torch = _torch()
    buf = torch.full((n + 2 * BAND,), float(SENTINEL), dtype=torch.float64, device=dev)
    return buf, buf[BAND + shift:BAND + shift + n]


def _bits(t):
    torch = _torch()
    return t.view(torch.int64)


def _check(buf, n, what, shift=0):
    want = int(np.array([SENTINEL]).view(np.int64)[0])
    b = _bits(buf)
    lo = BAND + shift
    assert bool((b[:lo] == want).all()), f"{what}: wrote below the output (n = {n})"
    assert bool((b[lo + n:] == want).all()), f"{what}: wrote past the output (n = {n})"
    assert not bool((b[lo:lo + n] == want).any()), f"{what}: left output elements unwritten (n = {n})"


def _inputs(n, seed, shift=0):
    """h, z as views into larger buffers; the surroundings hold very different parameters."""
    torch = _torch()
    rng = np.random.default_rng(seed)
    h = rng.uniform(0.1, 200.0, n)
    h[::5] = rng.integers(1, 5, h[::5].shape)
    z = rng.uniform(-50.0, 50.0, n)
    hb = torch.full((n + 2 * BAND,), 1e6, dtype=torch.float64, device="cuda")
    zb = torch.full((n + 2 * BAND,), 1e-3, dtype=torch.float64, device="cuda")
    lo = BAND + shift
    hb[lo:lo + n] = torch.from_numpy(h).cuda()
    zb[lo:lo + n] = torch.from_numpy(z).cuda()
    return hb[lo:lo + n], zb[lo:lo + n]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(_torch().cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("small_max", [None, "0", "mid"])
@pytest.mark.parametrize("shift", [0, 1])
def test_fill2_stays_inside_its_output(n, small_max, shift, monkeypatch):
    """hybrid fill2 through the one-launch kernel (default, n <= 2^17), the bucketed pipeline
    (PGM_CUDA_SMALL_MAX=0, 2^13-element passes so that several passes and a ragged last one run) and the
    two-launch path of mid-size fills ("mid": tile kernel + drain kernel), on 16-byte aligned operands (vector
    loads in the classify pass) and on 8-byte aligned ones."""
    torch = _torch()
    from polyagamma_b200 import _lib

    if small_max is not None:
        monkeypatch.setenv("PGM_CUDA_SMALL_MAX", "0")
        if small_max == "0":
            monkeypatch.setenv("PGM_CUDA_CHUNK_LOG2", "13")
    L = _lib.lib()
    h, z = _inputs(n, n, shift)
    buf, out = _banded(n, shift=shift)
    _lib.check(L.pgm_cuda_random_polyagamma_fill2(11, 0, _ptr(h), _ptr(z), _lib.HYBRID, n, _ptr(out), 0, _stream()),
               "fill2")
    torch.cuda.synchronize()
    _check(buf, n, "fill2", shift)
    # the same call on exact-size copies of the inputs draws the same values
    ref = torch.empty(n, dtype=torch.float64, device="cuda")
    hc, zc = h.clone(), z.clone()  # named: a temporary's block would be reused by the next allocation
    _lib.check(L.pgm_cuda_random_polyagamma_fill2(11, 0, _ptr(hc), _ptr(zc), _lib.HYBRID, n, _ptr(ref), 0, _stream()),
               "fill2")
    torch.cuda.synchronize()
    assert torch.equal(_bits(out.contiguous()), _bits(ref))


@pytest.mark.parametrize("method", ["GAMMA", "DEVROYE", "ALTERNATE", "SADDLE"])
@pytest.mark.parametrize("n", [1, 9, 2049, 70_001])
def test_explicit_methods_stay_inside_their_output(method, n):
    torch = _torch()
    from polyagamma_b200 import _lib

    L = _lib.lib()
    mid = getattr(_lib, method)
    h, z = _inputs(n, 3 * n + mid)
    h = h.clamp(min=1.0, max=12.0)
    if method == "DEVROYE":
        h = torch.ceil(h)
    if method == "GAMMA":
        n = min(n, 9000)  # 200 gamma draws per element
        h, z = h[:n], z[:n]
    h = h.contiguous()
    buf, out = _banded(n)
    _lib.check(L.pgm_cuda_random_polyagamma_fill2(5, 1, _ptr(h), _ptr(z), mid, n, _ptr(out), 0, _stream()), method)
    torch.cuda.synchronize()
    _check(buf, n, method)
    assert bool(torch.isfinite(out).all())
    # scalar-parameter entry point
    buf, out = _banded(n)
    _lib.check(L.pgm_cuda_random_polyagamma_fill(5, 2, 2.0, 1.5, mid, n, _ptr(out), 0, _stream()), method + " fill")
    torch.cuda.synchronize()
    _check(buf, n, method + " fill")


@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (257, 33), (64, 1025)])
def test_strided_fill_stays_inside_its_output(shape):
    """broadcast index space: h is a column, z a row"""
    torch = _torch()
    from polyagamma_b200 import _lib

    L = _lib.lib()
    r, c = shape
    n = r * c
    h, _ = _inputs(r, 17 + r)
    _, z = _inputs(c, 19 + c)
    buf, out = _banded(n)
    shp = _lib.int64_array([r, c])
    hs = _lib.int64_array([1, 0])
    zs = _lib.int64_array([0, 1])
    _lib.check(L.pgm_cuda_random_polyagamma_fill2_strided(3, 0, _ptr(h), _ptr(z), _lib.HYBRID, 2, shp, hs, zs,
                                                          _ptr(out), 0, _stream()), "fill2_strided")
    torch.cuda.synchronize()
    _check(buf, n, "fill2_strided")
    dense = torch.empty(n, dtype=torch.float64, device="cuda")
    hh = h[:, None].expand(r, c).contiguous().view(-1)
    zz = z[None, :].expand(r, c).contiguous().view(-1)
    _lib.check(L.pgm_cuda_random_polyagamma_fill2(3, 0, _ptr(hh), _ptr(zz), _lib.HYBRID, n, _ptr(dense), 0, _stream()),
               "fill2")
    torch.cuda.synchronize()
    assert torch.equal(_bits(out.contiguous()), _bits(dense))


@pytest.mark.parametrize("n", [1, 31, 33, 1025, 50_001])
def test_densities_stay_inside_their_output(n):
    torch = _torch()
    from polyagamma_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(n)
    x = torch.from_numpy(rng.uniform(0.01, 5.0, n)).cuda()
    h = torch.from_numpy(rng.uniform(0.5, 8.0, n)).cuda()
    z = torch.from_numpy(rng.uniform(-8.0, 8.0, n)).cuda()
    for name in ("pdf", "cdf", "logpdf", "logcdf"):
        buf, out = _banded(n)
        f = getattr(L, "pgm_cuda_polyagamma_" + name)
        _lib.check(f(_ptr(x), _ptr(h), _ptr(z), n, _ptr(out), _stream()), name)
        torch.cuda.synchronize()
        _check(buf, n, name)


@pytest.mark.parametrize("n", [1, 4097, 300_001])
def test_host_fill_stays_inside_its_output(n):
    """host-buffer entry point: slots on three streams; the host output has bands too"""
    from polyagamma_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(n)
    h = rng.uniform(0.1, 200.0, n)
    z = rng.uniform(-50.0, 50.0, n)
    buf = np.full(n + 2 * BAND, SENTINEL)
    out = buf[BAND:BAND + n]
    _lib.check(L.pgm_cuda_random_polyagamma_fill2_host(9, 0, h.ctypes.data, 1, z.ctypes.data, 1, _lib.HYBRID, n,
                                                       out.ctypes.data, 0, 0), "fill2_host")
    want = np.array([SENTINEL]).view(np.int64)[0]
    b = buf.view(np.int64)
    assert (b[:BAND] == want).all() and (b[BAND + n:] == want).all()
    assert not (b[BAND:BAND + n] == want).any()
    assert np.isfinite(out).all()


@pytest.mark.parametrize("p", [1, 3, 8, 17, 32, 50, 64])
@pytest.mark.parametrize("n", [1, 1000, 30_001])
def test_gibbs_step_stays_inside_its_outputs(n, p):
    """omega, psi and the p x p Gram matrix each sit between bands; X has a padded leading dimension whose
    padding holds huge values that must not reach psi or G"""
    torch = _torch()
    from polyagamma_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(100 * p + n % 97)
    ldx = p + 3
    Xp = torch.full((n, ldx), 1e30, dtype=torch.float64, device="cuda")
    Xp[:, :p] = torch.from_numpy(rng.normal(size=(n, p))).cuda()
    beta = torch.from_numpy(rng.normal(size=p) * 0.3).cuda()
    gbuf, G = _banded(p * p)
    obuf, om = _banded(n)
    pbuf, psi = _banded(n)
    _lib.check(L.pgm_cuda_logistic_omega_step(21, 0, _ptr(Xp), n, p, ldx, _ptr(beta), None, 1.0, _lib.HYBRID, _ptr(G),
                                              _ptr(om), _ptr(psi), 0, _stream()), "logistic_omega_step")
    torch.cuda.synchronize()
    _check(gbuf, p * p, "gram")
    _check(obuf, n, "omega")
    _check(pbuf, n, "psi")
    X = Xp[:, :p]
    want_psi = X @ beta
    assert float((psi - want_psi).abs().max()) < 1e-12 * max(1.0, float(want_psi.abs().max()))
    want = X.T @ (om[:, None] * X)
    assert float((G.view(p, p) - want).abs().max()) <= 1e-12 * max(float(want.abs().max()), 1e-300)


@pytest.mark.parametrize("n", [1, 4097, (1 << 22) + 3, 3 * (1 << 22), 13_000_007])
def test_download_fills_exactly_its_destination(n):
    """pgm_cuda_download (device result -> pageable host array through pinned bounce buffers and several host
    threads): every element arrives, in order, on both sides of the 2^22-element chunk and of the three-slot
    ring, and nothing outside [0, n) is written."""
    torch = _torch()
    from polyagamma_b200 import _lib

    src = torch.arange(n, dtype=torch.float64, device="cuda") * 0.5 + 1.0
    host = np.full(n + 2 * 512, -7.0)
    dst = host[512:512 + n]
    _lib.check(_lib.lib().pgm_cuda_download(ctypes.c_void_p(dst.ctypes.data), _ptr(src), n, _stream()), "download")
    assert np.array_equal(dst, src.cpu().numpy())
    assert (host[:512] == -7.0).all() and (host[512 + n:] == -7.0).all()
    # a second call reuses the ring
    src.mul_(-1.0)
    _lib.check(_lib.lib().pgm_cuda_download(ctypes.c_void_p(dst.ctypes.data), _ptr(src), n, _stream()), "download")
    assert np.array_equal(dst, src.cpu().numpy())


def test_host_results_equal_device_results():
    """The numpy results of the public functions (which travel through pgm_cuda_download from 2^18 elements on)
    are the device results, bit for bit."""
    import polyagamma_b200 as pg

    torch = _torch()
    x = np.linspace(0.01, 9.0, 700)[:, None]
    rng = np.random.default_rng(3)
    h = rng.uniform(0.5, 20.0, 600)[None, :]
    z = rng.uniform(-8.0, 8.0, 600)[None, :]
    for f in (pg.polyagamma_pdf, pg.polyagamma_cdf):
        host = f(x, h, z)
        dev = f(x, h, z, device="cuda")
        assert isinstance(host, np.ndarray) and host.shape == (700, 600)
        assert np.array_equal(host, dev.cpu().numpy(), equal_nan=True)
    a = pg.random_polyagamma(1.0, 0.5, size=1_000_003, random_state=(5, 6))
    b = pg.random_polyagamma(1.0, 0.5, size=1_000_003, random_state=(5, 6), device="cuda")
    assert isinstance(a, np.ndarray) and np.array_equal(a, b.cpu().numpy())
    hh = rng.uniform(0.2, 90.0, 400_001)
    zz = rng.uniform(-20.0, 20.0, 400_001)
    a = pg.random_polyagamma(hh, zz, random_state=(5, 7))
    b = pg.random_polyagamma(torch.from_numpy(hh).cuda(), torch.from_numpy(zz).cuda(), random_state=(5, 7))
    assert np.array_equal(a, b.cpu().numpy())


@pytest.mark.parametrize("n", [65_536, (1 << 21) - 1, (1 << 21) + 1, 3 * (1 << 21) + 5, 7 * (1 << 21) + 1])
@pytest.mark.parametrize("scalar_h", [False, True])
def test_pageable_host_fill_is_bounced_correctly(n, scalar_h):
    """fill2_host with plain numpy (pageable) arrays goes through pinned bounce buffers in 2^21-element chunks
    on a three-slot ring: every element equals the device-resident call's, on both sides of the chunk size and
    of the ring length, and the sentinels around `out` stay."""
    torch = _torch()
    from polyagamma_b200 import _lib

    rng = np.random.default_rng(n % 1000)
    h = np.array([2.5]) if scalar_h else rng.uniform(0.2, 150.0, n)
    z = rng.uniform(-30.0, 30.0, n)
    host = np.full(n + 2 * 512, -7.0)
    out = host[512:512 + n]
    rc = _lib.lib().pgm_cuda_random_polyagamma_fill2_host(31, 4, h.ctypes.data, 0 if scalar_h else 1, z.ctypes.data, 1,
                                                          _lib.HYBRID, n, out.ctypes.data, 17, -1)
    _lib.check(rc, "fill2_host")
    assert (host[:512] == -7.0).all() and (host[512 + n:] == -7.0).all()
    hd = torch.full((n,), 2.5, dtype=torch.float64, device="cuda") if scalar_h else torch.from_numpy(h).cuda()
    zd = torch.from_numpy(z).cuda()
    ref = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().pgm_cuda_random_polyagamma_fill2(31, 4, _ptr(hd), _ptr(zd), _lib.HYBRID, n, _ptr(ref), 17, _stream()),
               "fill2")
    torch.cuda.synchronize()
    assert np.array_equal(out, ref.cpu().numpy())
    # pinned arrays take the direct path and give the same bits
    hp, zp = torch.from_numpy(h).pin_memory(), torch.from_numpy(z).pin_memory()
    op = torch.empty(n, dtype=torch.float64).pin_memory()
    rc = _lib.lib().pgm_cuda_random_polyagamma_fill2_host(31, 4, hp.data_ptr(), 0 if scalar_h else 1, zp.data_ptr(), 1,
                                                          _lib.HYBRID, n, op.data_ptr(), 17, -1)
    _lib.check(rc, "fill2_host pinned")
    assert np.array_equal(op.numpy(), out)
