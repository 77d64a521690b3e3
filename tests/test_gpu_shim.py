"""GPU (-m gpu): the host-interop shim.  libpgm_cuda_shim.so exports the reference's
symbols (include/pgm_random.h:47-74, include/pgm_density.h:4-14); driven with a numpy BitGenerator
it must draw exactly what the C-ABI draws for the (seed, offset) the generator hands out, and the
reference's own Cython front-end relinked against it (oracle/_ref_shim, built by
`make -C oracle ref_shim`) must behave like the reference's Python API."""
import ctypes

import numpy as np
import pytest

from test_shim_host import _load

pytestmark = pytest.mark.gpu

GAMMA, DEVROYE, ALTERNATE, SADDLE, HYBRID = range(5)


def _ptr(bg):
    return ctypes.cast(bg.ctypes.bit_generator, ctypes.c_void_p).value


@pytest.mark.parametrize("method,name", [(HYBRID, None), (DEVROYE, "devroye"), (SADDLE, "saddle")])
def test_shim_draws_what_the_c_abi_draws(method, name):
    import polyagamma_b200 as pg

    lib = _load()
    n = 300_001
    rng = np.random.default_rng(1)
    h = (rng.integers(1, 5, n).astype(float) if method == DEVROYE else rng.uniform(0.5, 80, n))
    z = rng.uniform(-20, 20, n)
    bg = np.random.PCG64(42)
    seed, offset = (int(v) for v in np.random.PCG64(42).random_raw(2))
    out = np.empty(n)
    lib.pgm_random_polyagamma_fill2(_ptr(bg), h.ctypes.data, z.ctypes.data, method, n, out.ctypes.data)
    assert lib.pgm_cuda_shim_last_error() == 0
    want = pg.random_polyagamma(h, z, method=name, random_state=(seed, offset), disable_checks=True)
    assert np.array_equal(out, want)
    # second call: the generator moved on by two words
    seed2, offset2 = (int(v) for v in np.random.PCG64(42).random_raw(4)[2:])
    out2 = np.empty(1000)
    lib.pgm_random_polyagamma_fill(_ptr(bg), 2.5, 1.25, method if method != DEVROYE else HYBRID, 1000,
                                   out2.ctypes.data)
    want2 = pg.random_polyagamma(2.5, 1.25, size=1000, method=name if method != DEVROYE else None,
                                 random_state=(seed2, offset2))
    assert np.array_equal(out2, want2)
    one = lib.pgm_random_polyagamma(_ptr(bg), 1.0, 0.0, HYBRID)
    assert np.isfinite(one) and one > 0


def test_shim_invalid_h_semantics():
    """h <= 0 or NaN -> NaN, h = +inf -> +inf, in the data (src/pgm_random.c:137-155)."""
    lib = _load()
    bg = np.random.PCG64(0)
    h = np.array([1.0, -1.0, np.nan, np.inf, 0.0, 3.0])
    z = np.zeros(6)
    out = np.empty(6)
    lib.pgm_random_polyagamma_fill2(_ptr(bg), h.ctypes.data, z.ctypes.data, HYBRID, 6, out.ctypes.data)
    assert np.isfinite(out[[0, 5]]).all() and np.isnan(out[[1, 2, 4]]).all() and np.isposinf(out[3])


def test_shim_density_scalars(oracle):
    lib = _load()
    for x, h, z in [(0.25, 1.0, 0.0), (1.5, 4.0, -2.0), (0.02, 0.5, 7.0), (6.0, 20.0, 1.0)]:
        for name in ("pdf", "cdf", "logpdf", "logcdf"):
            got = getattr(lib, "pgm_polyagamma_" + name)(x, h, z)
            want = float(np.asarray(oracle.density(name, x, h, z)).reshape(-1)[0])
            assert got == pytest.approx(want, rel=1e-10, abs=1e-300), (name, x, h, z)
    assert lib.pgm_polyagamma_pdf(-1.0, 1.0, 0.0) == 0.0 and lib.pgm_polyagamma_cdf(np.inf, 1.0, 0.0) == 1.0


@pytest.fixture(scope="module")
def relinked():
    import ref  # oracle/ref.py (tests/conftest.py puts oracle/ on sys.path)

    if not ref.shim_module_available():
        pytest.skip("oracle/_ref_shim not built (make -C oracle ref_shim needs /root/reference)")
    return ref.shim_module()


def test_relinked_reference_module_runs_on_the_gpu(relinked):
    """The reference's unmodified Cython module, relinked: same call contract, GPU draws."""
    import polyagamma_b200 as pg

    launches = pg._lib.lib().pgm_cuda_launch_count()
    n = 200_000
    rng = np.random.default_rng(3)
    h = rng.uniform(0.2, 120, n)
    z = rng.normal(0, 4, n)
    got = relinked.rvs(h, z, random_state=77, disable_checks=True)
    assert pg._lib.lib().pgm_cuda_launch_count() > launches  # our kernels ran
    seed, offset = (int(v) for v in np.random.PCG64(77).random_raw(2))
    want = pg.random_polyagamma(h, z, random_state=(seed, offset), disable_checks=True)
    assert np.array_equal(got, want)
    # scalar parameters + size -> pgm_random_polyagamma_fill
    got = relinked.rvs(3.0, -1.0, size=(50, 40), random_state=np.random.Generator(np.random.PCG64(5)))
    seed, offset = (int(v) for v in np.random.PCG64(5).random_raw(2))
    want = pg.random_polyagamma(3.0, -1.0, size=(50, 40), random_state=(seed, offset))
    assert got.shape == (50, 40) and np.array_equal(got, want)
    # the reference's own argument checks still run in front of the GPU call (_polyagamma.pyx:147-166)
    with pytest.raises(ValueError):
        relinked.rvs(1.0, 0.0, method="nope")
    with pytest.raises(ValueError):
        relinked.rvs(-1.0, 0.0)
    x = rng.uniform(0.05, 3, 500)
    assert np.allclose(relinked.pdf(x, 2.0, 1.0), pg.polyagamma_pdf(x, 2.0, 1.0), rtol=1e-12, atol=0)
    assert np.allclose(relinked.cdf(x, 2.0, 1.0, return_log=True), pg.polyagamma_cdf(x, 2.0, 1.0, return_log=True),
                       rtol=1e-12, atol=1e-14)


def test_relinked_module_moments(relinked):
    """1e6 hybrid draws through the relinked module at a devroye, an alternate and a saddle cell:
    exact PG mean/variance within 4 standard errors (default, corrected constants)."""
    n = 1_000_000
    for h, z in [(1.0, 1.5), (2.5, 6.0), (20.0, 2.0)]:
        x = relinked.rvs(h, z, size=n, random_state=11)
        mean = h * np.tanh(z / 2) / (2 * z)
        var = h * (np.sinh(z) - z) / (np.cosh(z / 2) ** 2) / (4 * z ** 3)
        assert abs(x.mean() - mean) <= 4 * np.sqrt(var / n), (h, z)
        assert abs(x.var() - var) <= 6 * var * np.sqrt(2.0 / n) * 2, (h, z)
