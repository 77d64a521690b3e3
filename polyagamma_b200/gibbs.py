"""The logistic-regression Gibbs omega-step on the GPU.

The reference's users draw ``omega_i ~ PG(n_i, x_i . beta)`` with ``random_polyagamma`` and then form
``X' diag(omega) X`` themselves (README.md:146-148 of the reference; Polson, Scott & Windle 2013,
section 3).  :func:`logistic_omega_step` does the three steps on the device with one call into
``pgm_cuda_logistic_omega_step``; the draw is exactly ``random_polyagamma(n_trials, X @ beta)`` with the
same ``random_state`` (same Philox streams, same hybrid dispatch).
"""
from __future__ import annotations

import ctypes

from . import _lib
from ._api import _device_of, _method_id, _resolve_random_state, _stream_ptr, _torch

__all__ = ["logistic_omega_step"]


def logistic_omega_step(X, beta, n_trials=1.0, *, random_state=None, method=None, ref_compat=False,
                        return_omega=True, return_psi=False, first_index=0):
    """``psi = X @ beta``; ``omega ~ PG(n_trials, psi)``; ``G = X.T @ diag(omega) @ X``.

    X: (n, p) float64 CUDA tensor, row-major (``stride(1) == 1``), ``p <= 64``; beta: (p,);
    n_trials: scalar or (n,) tensor of shape parameters.  Returns ``(G, omega)`` (``omega`` is None
    when ``return_omega=False``), plus ``psi`` when ``return_psi=True``.  All outputs are CUDA tensors.
    """
    torch = _torch()
    if not (isinstance(X, torch.Tensor) and X.is_cuda and X.dtype == torch.float64 and X.dim() == 2):
        raise TypeError("X must be a 2-d float64 CUDA tensor")
    if X.stride(1) != 1 and X.shape[1] > 1:
        X = X.contiguous()
    n, p = X.shape
    if not 1 <= p <= 64:
        raise ValueError("logistic_omega_step supports 1 <= p <= 64 columns")
    dev = _device_of(X)
    beta = torch.as_tensor(beta, dtype=torch.float64, device=dev).contiguous()
    if beta.shape != (p,):
        raise ValueError(f"beta must have shape ({p},)")
    nt_ptr, nt_scalar = None, 1.0
    if isinstance(n_trials, (int, float)):
        nt_scalar = float(n_trials)
    else:
        nt = torch.as_tensor(n_trials, dtype=torch.float64, device=dev).contiguous()
        if nt.shape != (n,):
            raise ValueError(f"n_trials must be a scalar or have shape ({n},)")
        nt_ptr = ctypes.c_void_p(nt.data_ptr())
    seed, offset = _resolve_random_state(random_state)
    gram = torch.empty((p, p), dtype=torch.float64, device=dev)
    omega = torch.empty(n, dtype=torch.float64, device=dev) if return_omega else None
    psi = torch.empty(n, dtype=torch.float64, device=dev) if return_psi else None
    ldx = X.stride(0) if n > 1 else p
    with torch.cuda.device(dev):
        rc = _lib.lib().pgm_cuda_logistic_omega_step(
            seed, offset, ctypes.c_void_p(X.data_ptr()), n, p, ldx, ctypes.c_void_p(beta.data_ptr()), nt_ptr,
            nt_scalar, _method_id(method, ref_compat), ctypes.c_void_p(gram.data_ptr()),
            ctypes.c_void_p(omega.data_ptr()) if omega is not None else None,
            ctypes.c_void_p(psi.data_ptr()) if psi is not None else None, int(first_index), _stream_ptr(dev))
    _lib.check(rc, "pgm_cuda_logistic_omega_step")
    return (gram, omega, psi) if return_psi else (gram, omega)
