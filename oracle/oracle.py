"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``libpgm_oracle.so`` (``pgm_oracle.c``: the scalar FP64/Philox CPU
restatement of the reference's hot path).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline leg import this; nothing under ``polyagamma_b200/`` does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libpgm_oracle.so")

GAMMA, DEVROYE, ALTERNATE, SADDLE, HYBRID, NORMAL = range(6)
REF_COMPAT = 0x100
METHODS = {None: HYBRID, "hybrid": HYBRID, "gamma": GAMMA, "devroye": DEVROYE,
           "alternate": ALTERNATE, "saddle": SADDLE}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pgm_oracle.c")
    if force or not os.path.exists(_PATH) or os.path.getmtime(_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_PATH)
        d, vp, sz, u64, i = (ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64,
                             ctypes.c_int)
        L.pgo_philox4x32_10.argtypes = [vp, vp, vp]
        L.pgo_call_key.argtypes = [u64, u64, vp]
        L.pgo_stream_words.argtypes = [u64, u64, u64, sz, vp]
        L.pgo_stream_variates.argtypes = [u64, u64, u64, i, d, sz, vp]
        L.pgo_hybrid_select.argtypes = [d, d]
        L.pgo_hybrid_select.restype = i
        L.pgo_random_polyagamma_fill.argtypes = [u64, u64, d, d, i, sz, vp, u64]
        L.pgo_random_polyagamma_fill2.argtypes = [u64, u64, vp, vp, i, sz, vp, u64]
        L.pgo_last_words_consumed.restype = u64
        for name in ("pgo_lgamma", "pgo_erfcx", "pgo_alternate_trunc_point"):
            getattr(L, name).argtypes = [d]
            getattr(L, name).restype = d
        L.pgo_log_upper_gamma.argtypes = [d, d, i]
        L.pgo_log_upper_gamma.restype = d
        L.pgo_saddle_newton.argtypes = [d, vp]
        L.pgo_saddle_newton.restype = d
        L.pgo_setup_probe.argtypes = [i, d, d, vp]
        for name in ("pdf", "cdf", "logpdf", "logcdf"):
            f = getattr(L, "pgo_polyagamma_" + name)
            f.argtypes = [d, d, d]
            f.restype = d
        L.pgo_density_array.argtypes = [i, vp, vp, vp, sz, vp, vp]
        _lib = L
    return _lib


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.empty(4, dtype=np.uint32)
    lib().pgo_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def call_key(seed, offset):
    k = np.empty(2, dtype=np.uint32)
    lib().pgo_call_key(seed, offset, k.ctypes.data)
    return k


def stream_words(seed, offset, index, n):
    o = np.empty(n, dtype=np.uint32)
    lib().pgo_stream_words(seed, offset, index, n, o.ctypes.data)
    return o


def stream_variates(seed, offset, index, kind, n, param=0.0):
    o = np.empty(n, dtype=np.float64)
    lib().pgo_stream_variates(seed, offset, index, kind, param, n, o.ctypes.data)
    return o


def method_id(method, ref_compat=False):
    m = METHODS[method] if not isinstance(method, int) else method
    return m | (REF_COMPAT if ref_compat else 0)


def fill(seed, offset, h, z, n, method=None, first_index=0, ref_compat=False):
    out = np.empty(int(n), dtype=np.float64)
    lib().pgo_random_polyagamma_fill(seed, offset, float(h), float(z), method_id(method, ref_compat),
                                     out.size, out.ctypes.data, first_index)
    return out


def fill2(seed, offset, h, z, method=None, first_index=0, ref_compat=False):
    hb, zb = np.broadcast_arrays(np.asarray(h, dtype=np.float64), np.asarray(z, dtype=np.float64))
    hb = np.ascontiguousarray(hb)
    zb = np.ascontiguousarray(zb)
    out = np.empty(hb.shape, dtype=np.float64)
    lib().pgo_random_polyagamma_fill2(seed, offset, hb.ctypes.data, zb.ctypes.data,
                                      method_id(method, ref_compat), out.size, out.ctypes.data,
                                      first_index)
    return out


def hybrid_select(h, z):
    return lib().pgo_hybrid_select(float(h), float(z))


def setup_probe(method, h, z, ref_compat=False):
    o = np.full(8, np.nan)
    lib().pgo_setup_probe(method_id(method, ref_compat), float(h), float(z), o.ctypes.data)
    return o


def saddle_newton(x):
    fp = ctypes.c_double()
    u = lib().pgo_saddle_newton(float(x), ctypes.byref(fp))
    return u, fp.value


_WHICH = {"pdf": 0, "cdf": 1, "logpdf": 2, "logcdf": 3}


def density(name, x, h=1.0, z=0.0, return_cond=False):
    xb, hb, zb = np.broadcast_arrays(np.asarray(x, dtype=np.float64), np.asarray(h, dtype=np.float64),
                                     np.asarray(z, dtype=np.float64))
    xb, hb, zb = (np.ascontiguousarray(a) for a in (xb, hb, zb))
    out = np.empty(xb.shape, dtype=np.float64)
    cond = np.empty(xb.shape, dtype=np.float64)
    lib().pgo_density_array(_WHICH[name], xb.ctypes.data, hb.ctypes.data, zb.ctypes.data, out.size,
                            out.ctypes.data, cond.ctypes.data)
    return (out, cond) if return_cond else out
