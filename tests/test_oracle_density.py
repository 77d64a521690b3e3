"""CPU: the oracle's density restatement against the reference's own fixtures and build.

Pins oracle/pgm_oracle.c (density part) to: the five known-answer values and the edge values
of tests/test_polyagamma.py:189-221, the doc-string values of polyagamma/_polyagamma.pyx:309-320
that still reproduce, the committed golden grid generated from oracle/_ref, and -- when the
compiled reference is present -- oracle/_ref itself on a fresh grid.
"""
import numpy as np
import pytest

NAMES = ("pdf", "cdf", "logpdf", "logcdf")


def _agree(a, b, rtol):
    both_nan = np.isnan(a) & np.isnan(b)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    fin = np.isfinite(a) & np.isfinite(b)
    ok = both_nan | same_inf
    ok[fin] = np.abs(a[fin] - b[fin]) <= rtol * np.maximum(np.abs(b[fin]), 1e-300)
    return ok


def test_reference_known_answers(oracle):
    # tests/test_polyagamma.py:213-221 (np.isclose, rtol 1e-5)
    assert np.isclose(oracle.density("logcdf", 0.01, 10, 3), -1238.6998500970105)
    assert np.isclose(oracle.density("logcdf", 1e-3, 10, 3), -12489.8793293567)
    assert np.isclose(oracle.density("logpdf", 1e-3, 10, 3), -12473.46649418656)
    assert np.isclose(oracle.density("logcdf", 1e-16), -1250000000000017.2)
    assert np.isclose(oracle.density("logpdf", 1e-16), -1249999999999945.8)
    # polyagamma/_polyagamma.pyx:309-320 doc-string values that the reference reproduces exactly
    assert oracle.density("pdf", 0.2) == pytest.approx(2.339176537265802, rel=1e-14)
    assert oracle.density("pdf", 3.0, 6.0, 1.0) == pytest.approx(0.012537773310487178, rel=1e-13)
    assert oracle.density("pdf", 0.1) == pytest.approx(3.613955566329298, rel=1e-14)


def test_reference_edge_values(oracle):
    # tests/test_polyagamma.py:189-200
    for x, pdf, cdf in ((0.0, 0.0, 0.0), (-1.0, 0.0, 0.0), (np.inf, 0.0, 1.0)):
        assert oracle.density("pdf", x) == pdf
        assert oracle.density("cdf", x) == cdf
        assert oracle.density("logpdf", x) == -np.inf
        assert oracle.density("logcdf", x) == (0.0 if cdf == 1.0 else -np.inf)


@pytest.mark.parametrize("name", NAMES)
def test_golden_grid(oracle, density_golden, name):
    g = density_golden
    with np.errstate(all="ignore"):
        got = oracle.density(name, g["x"], g["h"], g["z"])
    ok = _agree(got, g[name], 1e-13)
    assert ok.all(), f"{name}: {np.sum(~ok)} of {ok.size} golden points differ"


@pytest.mark.parametrize("name", NAMES)
def test_against_compiled_reference(oracle, ref, name):
    rng = np.random.default_rng(5)
    n = 3000
    x = rng.uniform(1e-3, 10, n)
    h = rng.uniform(0.1, 40, n)
    z = rng.uniform(-30, 30, n)
    z[::7] = 0.0
    with np.errstate(all="ignore"):
        got = oracle.density(name, x, h, z)
        want = ref.c_density(name, x, h, z)
    assert _agree(got, want, 1e-13).all()


def test_even_in_z(oracle):
    # pdf/cdf are even in z (fabs at src/pgm_density.c:74,246,301)
    x = np.linspace(0.05, 4, 50)
    for name in NAMES:
        assert np.array_equal(oracle.density(name, x, 2.5, 3.0), oracle.density(name, x, 2.5, -3.0))


def test_stable_density_restatement_against_300_digit_series():
    """oracle/stable_density.py (the contour-integration pdf / cdf the CUDA kernels offer as which = 4..7, SURVEY
    8f-3) against tests/golden/stable_density_golden.npz: the reference's alternating series in 300-digit
    arithmetic (tests/golden/make_stable_golden.py), on h in {8 .. 200}, z in {0 .. 20}, x = mean + k sd."""
    import os
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "oracle"))
    import stable_density as S

    g = np.load(os.path.join(here, "golden", "stable_density_golden.npz"))
    assert g["x"].size >= 100 and set(np.unique(g["h"])) >= {8.0, 40.0, 200.0}
    worst = 0.0
    for x, h, z, pdf, cdf, sf in zip(g["x"], g["h"], g["z"], g["pdf"], g["cdf"], g["sf"]):
        p, c, s = S.pdf_cdf(x, h, z)
        for got, want in ((p, pdf), (c, cdf), (s, sf)):
            if want > 1e-290:
                worst = max(worst, abs(got - want) / want)
    assert worst < 2e-12, worst
