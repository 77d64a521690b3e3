"""ctypes binding of ``libpgm_cuda.so`` (the C-ABI declared in ``include/pgm_cuda.h``).

There is no CPU fallback anywhere in this package: if the CUDA library has not been built the
import of any compute entry point raises, loudly.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpgm_cuda.so")

# include/pgm_cuda.h: pgm_cuda_sampler_t (same values as the reference's sampler_t,
# include/pgm_random.h:13)
GAMMA, DEVROYE, ALTERNATE, SADDLE, HYBRID = range(5)
EXACT = 6  # Python layer only: drawn as ALTERNATE with the consistent envelope, for h >= 1
REF_COMPAT = 0x100
EINVAL = -1

# Every symbol include/pgm_cuda.h declares; tests check the library exports all of them.
EXPORTS = (
    "pgm_cuda_random_polyagamma",
    "pgm_cuda_random_polyagamma_fill",
    "pgm_cuda_random_polyagamma_fill2",
    "pgm_cuda_random_polyagamma_fill2_strided",
    "pgm_cuda_random_polyagamma_fill2_host",
    "pgm_cuda_random_polyagamma_fill2_host_multi",
    "pgm_cuda_last_host_fill_ms",
    "pgm_cuda_any_invalid_h",
    "pgm_cuda_polyagamma_pdf",
    "pgm_cuda_polyagamma_cdf",
    "pgm_cuda_polyagamma_logpdf",
    "pgm_cuda_polyagamma_logcdf",
    "pgm_cuda_polyagamma_density_strided",
    "pgm_cuda_polyagamma_density_host",
    "pgm_cuda_download",
    "pgm_cuda_logistic_omega_step",
    "pgm_cuda_launch_count",
    "pgm_cuda_profile_begin",
    "pgm_cuda_profile_end",
    "pgm_cuda_fp64_peak_tflops",
    "pgm_cuda_test_math",
    "pgm_cuda_release_scratch",
    "pgm_cuda_init",
    "pgm_cuda_debug_report",
    "pgm_cuda_version",
    "pgm_cuda_hybrid_select",
    "pgm_cuda_call_key",
)

_lib = None


class ExtensionMissing(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load libpgm_cuda.so (once) and declare its prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ExtensionMissing(
            f"polyagamma_b200: {LIB_PATH} is missing and there is no CPU fallback. Build it with "
            "`make -C polyagamma_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`.")
    L = ctypes.CDLL(LIB_PATH)
    d, vp, sz, u64, i = ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int
    p64 = ctypes.POINTER(ctypes.c_int64)
    L.pgm_cuda_random_polyagamma.restype = d
    L.pgm_cuda_random_polyagamma.argtypes = [u64, u64, d, d, i]
    L.pgm_cuda_random_polyagamma_fill.restype = i
    L.pgm_cuda_random_polyagamma_fill.argtypes = [u64, u64, d, d, i, sz, vp, u64, vp]
    L.pgm_cuda_random_polyagamma_fill2.restype = i
    L.pgm_cuda_random_polyagamma_fill2.argtypes = [u64, u64, vp, vp, i, sz, vp, u64, vp]
    L.pgm_cuda_random_polyagamma_fill2_strided.restype = i
    L.pgm_cuda_random_polyagamma_fill2_strided.argtypes = [u64, u64, vp, vp, i, i, p64, p64, p64, vp, u64, vp]
    L.pgm_cuda_random_polyagamma_fill2_host.restype = i
    L.pgm_cuda_random_polyagamma_fill2_host.argtypes = [u64, u64, vp, i, vp, i, i, sz, vp, u64, i]
    L.pgm_cuda_random_polyagamma_fill2_host_multi.restype = i
    L.pgm_cuda_random_polyagamma_fill2_host_multi.argtypes = [u64, u64, vp, i, vp, i, i, sz, vp, u64,
                                                              ctypes.POINTER(ctypes.c_int), i]
    L.pgm_cuda_last_host_fill_ms.restype = d
    L.pgm_cuda_last_host_fill_ms.argtypes = []
    L.pgm_cuda_any_invalid_h.restype = i
    L.pgm_cuda_any_invalid_h.argtypes = [vp, sz, i, ctypes.POINTER(ctypes.c_int), vp]
    for name in ("pdf", "cdf", "logpdf", "logcdf"):
        f = getattr(L, "pgm_cuda_polyagamma_" + name)
        f.restype = i
        f.argtypes = [vp, vp, vp, sz, vp, vp]
    L.pgm_cuda_polyagamma_density_strided.restype = i
    L.pgm_cuda_polyagamma_density_strided.argtypes = [i, vp, vp, vp, i, p64, p64, p64, p64, vp, vp]
    L.pgm_cuda_polyagamma_density_host.restype = i
    L.pgm_cuda_polyagamma_density_host.argtypes = [i, vp, i, vp, i, vp, i, sz, vp, i]
    L.pgm_cuda_download.restype = i
    L.pgm_cuda_download.argtypes = [vp, vp, sz, vp]
    L.pgm_cuda_logistic_omega_step.restype = i
    L.pgm_cuda_logistic_omega_step.argtypes = [u64, u64, vp, sz, i, sz, vp, vp, d, i, vp, vp, vp, u64, vp]
    L.pgm_cuda_launch_count.restype = u64
    L.pgm_cuda_launch_count.argtypes = []
    L.pgm_cuda_profile_begin.restype = None
    L.pgm_cuda_profile_begin.argtypes = []
    L.pgm_cuda_profile_end.restype = i
    L.pgm_cuda_profile_end.argtypes = [ctypes.POINTER(d), ctypes.POINTER(u64), i]
    L.pgm_cuda_fp64_peak_tflops.restype = i
    L.pgm_cuda_fp64_peak_tflops.argtypes = [ctypes.POINTER(d), vp]
    L.pgm_cuda_release_scratch.restype = i
    L.pgm_cuda_release_scratch.argtypes = []
    L.pgm_cuda_init.restype = i
    L.pgm_cuda_init.argtypes = []
    L.pgm_cuda_debug_report.restype = None
    L.pgm_cuda_debug_report.argtypes = []
    L.pgm_cuda_test_math.restype = i
    L.pgm_cuda_test_math.argtypes = [i, vp, sz, vp, vp]
    L.pgm_cuda_version.restype = ctypes.c_char_p
    L.pgm_cuda_version.argtypes = []
    L.pgm_cuda_hybrid_select.restype = i
    L.pgm_cuda_hybrid_select.argtypes = [d, d]
    L.pgm_cuda_call_key.restype = None
    L.pgm_cuda_call_key.argtypes = [u64, u64, ctypes.POINTER(ctypes.c_uint32)]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc == EINVAL:
        raise ValueError(f"{what}: invalid argument")
    raise RuntimeError(f"{what}: CUDA error {rc}")


def int64_array(values):
    arr = (ctypes.c_int64 * max(len(values), 1))(*[int(v) for v in values])
    return arr


def release_scratch() -> None:
    """Return the library's cached per-pass scratch of the current device to the driver."""
    check(lib().pgm_cuda_release_scratch(), "pgm_cuda_release_scratch")


def launch_count() -> int:
    return int(lib().pgm_cuda_launch_count())


_METHOD_KERNELS = ("devroye", "alternate", "saddle", "saddle_left", "saddle_far", "alternate_left", "saddle_mid")
KERNEL_KINDS = (("drain", "small", "tile") + _METHOD_KERNELS + ("normal", "gamma", "density") +
                tuple("setup_" + k for k in _METHOD_KERNELS) + ("gibbs_xbeta", "gibbs_gram"))


def profile_begin() -> None:
    lib().pgm_cuda_profile_begin()


def profile_end() -> dict:
    """{kind: (milliseconds, launches)} accumulated since profile_begin()."""
    n = len(KERNEL_KINDS)
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_uint64 * n)()
    check(lib().pgm_cuda_profile_end(ms, cnt, n), "pgm_cuda_profile_end")
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(KERNEL_KINDS)}


def fp64_peak_tflops() -> float:
    v = ctypes.c_double(0.0)
    check(lib().pgm_cuda_fp64_peak_tflops(ctypes.byref(v), None), "pgm_cuda_fp64_peak_tflops")
    return float(v.value)
