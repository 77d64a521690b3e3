"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY.

Loader for the UNMODIFIED reference (zoj613/polyagamma) compiled by ``make -C oracle ref`` into
``oracle/_ref/``.  Two views of the same compiled C sources:

* ``random_polyagamma / polyagamma_pdf / polyagamma_cdf`` -- the reference's own Python API
  (its Cython module ``polyagamma/_polyagamma.pyx``: ``rvs``/``pdf``/``cdf``), importable
  without the reference's ``__init__.py`` (which needs a generated ``_version.py``).
* ``clib()`` -- ctypes handle on ``libpgmref.so`` exporting the C API of
  ``include/pgm_random.h`` / ``include/pgm_density.h``; drive it with a numpy BitGenerator
  through ``bitgen_ptr(bitgen)`` (numpy exposes its ``bitgen_t*`` via ``.ctypes``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl reference
legs may import this module.  Nothing under ``polyagamma_b200/`` does.
"""
from __future__ import annotations

import ctypes
import glob
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")

GAMMA, DEVROYE, ALTERNATE, SADDLE, HYBRID = range(5)  # include/pgm_random.h:13
METHODS = {None: HYBRID, "gamma": GAMMA, "devroye": DEVROYE, "alternate": ALTERNATE,
           "saddle": SADDLE, "hybrid": HYBRID}


def available() -> bool:
    return bool(glob.glob(os.path.join(_REF_DIR, "_polyagamma*.so"))) and \
        os.path.exists(os.path.join(_REF_DIR, "libpgmref.so"))


_mod = None


def module():
    """The reference's compiled Cython module (rvs, pdf, cdf)."""
    global _mod
    if _mod is None:
        paths = glob.glob(os.path.join(_REF_DIR, "_polyagamma*.so"))
        if not paths:
            raise ImportError("oracle/_ref/_polyagamma*.so missing: run `make -C oracle ref`")
        spec = importlib.util.spec_from_file_location("_polyagamma", paths[0])
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def random_polyagamma(*a, **k):
    return module().rvs(*a, **k)


def polyagamma_pdf(*a, **k):
    return module().pdf(*a, **k)


def polyagamma_cdf(*a, **k):
    return module().cdf(*a, **k)


_clib = None


def clib():
    global _clib
    if _clib is None:
        path = os.path.join(_REF_DIR, "libpgmref.so")
        if not os.path.exists(path):
            raise ImportError("oracle/_ref/libpgmref.so missing: run `make -C oracle ref`")
        lib = ctypes.CDLL(path)
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
        _clib = lib
    return _clib


def bitgen_ptr(bit_generator) -> int:
    """Address of the numpy ``bitgen_t`` behind a BitGenerator (e.g. ``np.random.PCG64(1)``)."""
    return ctypes.cast(bit_generator.ctypes.bit_generator, ctypes.c_void_p).value


def c_fill2(bit_generator, h, z, method=HYBRID):
    """pgm_random_polyagamma_fill2 (src/pgm_random.c:173-180) on contiguous double arrays."""
    h = np.ascontiguousarray(h, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    assert h.shape == z.shape
    out = np.empty(h.shape, dtype=np.float64)
    clib().pgm_random_polyagamma_fill2(bitgen_ptr(bit_generator), h.ctypes.data, z.ctypes.data,
                                       int(method), h.size, out.ctypes.data)
    return out


def c_fill(bit_generator, h, z, n, method=HYBRID):
    """pgm_random_polyagamma_fill (src/pgm_random.c:158-170)."""
    out = np.empty(int(n), dtype=np.float64)
    clib().pgm_random_polyagamma_fill(bitgen_ptr(bit_generator), float(h), float(z),
                                      int(method), out.size, out.ctypes.data)
    return out


def c_density(name, x, h, z):
    """Elementwise pgm_polyagamma_{pdf,cdf,logpdf,logcdf} (src/pgm_density.c) with broadcasting."""
    f = getattr(clib(), "pgm_polyagamma_" + name)
    xb, hb, zb = np.broadcast_arrays(np.asarray(x, float), np.asarray(h, float), np.asarray(z, float))
    out = np.empty(xb.shape, dtype=np.float64)
    of = out.reshape(-1)
    for i, (a, b, c) in enumerate(zip(xb.reshape(-1), hb.reshape(-1), zb.reshape(-1))):
        of[i] = f(a, b, c)
    return out


_shim_mod = None


def shim_module_available() -> bool:
    return bool(glob.glob(os.path.join(_HERE, "_ref_shim", "_polyagamma*.so")))


def shim_module():
    """The reference's Cython front-end linked against polyagamma_b200/libpgm_cuda_shim.so instead
    of the reference's C sources (``make -C oracle ref_shim``): its ``rvs``/``pdf``/``cdf`` run on
    the GPU.  Used by tests/test_gpu_shim.py only ("switch by relinking")."""
    global _shim_mod
    if _shim_mod is None:
        paths = glob.glob(os.path.join(_HERE, "_ref_shim", "_polyagamma*.so"))
        if not paths:
            raise ImportError("oracle/_ref_shim/_polyagamma*.so missing: run `make -C oracle ref_shim`")
        # a distinct module name keeps it apart from the CPU build loaded by module(); the init
        # symbol is still PyInit__polyagamma, which spec_from_file_location derives from the name
        # after the last dot
        spec = importlib.util.spec_from_file_location("pgm_shim_linked._polyagamma", paths[0])
        _shim_mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_shim_mod)
    return _shim_mod
