"""polyagamma_b200 -- B200-native (sm_100a) Polya-Gamma sampling and densities.

Drop-in for the hot path of zoj613/polyagamma: ``random_polyagamma``, ``polyagamma_pdf`` and
``polyagamma_cdf`` with the reference's call contract (``polyagamma/__init__.py:1-5``), running
as hand-written FP64 CUDA kernels behind the C-ABI of ``include/pgm_cuda.h``.
"""
from ._api import PhiloxGenerator, polyagamma_cdf, polyagamma_pdf, random_polyagamma
from ._lib import ExtensionMissing, launch_count, release_scratch

__version__ = "0.1.0"
__all__ = [
    "random_polyagamma",
    "polyagamma_pdf",
    "polyagamma_cdf",
    "PhiloxGenerator",
    "ExtensionMissing",
    "launch_count",
    "release_scratch",
]
