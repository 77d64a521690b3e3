"""CPU: libpgm_cuda_shim.so exports the reference's own symbols and, without a GPU,
fails loudly instead of falling back to a CPU sampler."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "polyagamma_b200", "libpgm_cuda_shim.so")


def _load():
    lib = ctypes.CDLL(SHIM)
    d, vp, sz = ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t
    lib.pgm_random_polyagamma.restype = d
    lib.pgm_random_polyagamma.argtypes = [vp, d, d, ctypes.c_int]
    lib.pgm_random_polyagamma_fill.restype = None
    lib.pgm_random_polyagamma_fill.argtypes = [vp, d, d, ctypes.c_int, sz, vp]
    lib.pgm_random_polyagamma_fill2.restype = None
    lib.pgm_random_polyagamma_fill2.argtypes = [vp, vp, vp, ctypes.c_int, sz, vp]
    for name in ("pdf", "cdf", "logpdf", "logcdf"):
        f = getattr(lib, "pgm_polyagamma_" + name)
        f.restype = d
        f.argtypes = [d, d, d]
    lib.pgm_cuda_shim_last_error.restype = ctypes.c_int
    return lib


def test_shim_exports_the_reference_symbols():
    text = open(os.path.join(ROOT, "include", "pgm_cuda_shim.h")).read()
    assert "include/pgm_random.h:47-74" in text and "include/pgm_density.h:4-14" in text
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(pgm_[a-z0-9_]+)\s*\(", code)))
    # exactly the seven functions of the reference's public headers, plus the error getter
    assert declared == sorted(["pgm_random_polyagamma", "pgm_random_polyagamma_fill", "pgm_random_polyagamma_fill2",
                               "pgm_polyagamma_pdf", "pgm_polyagamma_cdf", "pgm_polyagamma_logpdf",
                               "pgm_polyagamma_logcdf", "pgm_cuda_shim_last_error"])
    lib = _load()
    for name in declared:
        assert hasattr(lib, name), name


def test_shim_without_gpu_fails_loudly(capfd):
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present: covered by tests/test_gpu_shim.py")
    lib = _load()
    bg = np.random.PCG64(3)
    twin = np.random.PCG64(3)
    ptr = ctypes.cast(bg.ctypes.bit_generator, ctypes.c_void_p).value
    out = np.zeros(5)
    lib.pgm_random_polyagamma_fill(ptr, 1.0, 0.5, 4, out.size, out.ctypes.data)
    assert np.isnan(out).all() and lib.pgm_cuda_shim_last_error() != 0
    assert "libpgm_cuda_shim" in capfd.readouterr().err
    # the caller's generator advanced by exactly two 64-bit words
    twin.random_raw(2)
    assert bg.random_raw() == twin.random_raw()
    assert np.isnan(lib.pgm_polyagamma_pdf(0.5, 1.0, 0.0))
