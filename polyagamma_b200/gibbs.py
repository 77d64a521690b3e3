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

__all__ = ["logistic_omega_step", "sharded_logistic_omega_step", "allreduce_gram"]


def logistic_omega_step(X, beta, n_trials=1.0, *, random_state=None, method=None, ref_compat=False,
                        return_omega=True, return_psi=False, first_index=0):
    """``psi = X @ beta``; ``omega ~ PG(n_trials, psi)``; ``G = X.T @ diag(omega) @ X``.

    X: (n, p) float64 CUDA tensor, row-major (``stride(1) == 1``); beta: (p,).  ``p <= 64`` runs the fused
    kernels (one C-ABI call); wider X falls back to two cuBLAS GEMMs around the same draw.
    n_trials: scalar or (n,) tensor of shape parameters.  Returns ``(G, omega)`` (``omega`` is None
    when ``return_omega=False``), plus ``psi`` when ``return_psi=True``.  All outputs are CUDA tensors.
    """
    torch = _torch()
    if not (isinstance(X, torch.Tensor) and X.is_cuda and X.dtype == torch.float64 and X.dim() == 2):
        raise TypeError("X must be a 2-d float64 CUDA tensor")
    if X.stride(1) != 1 and X.shape[1] > 1:
        X = X.contiguous()
    n, p = X.shape
    if p < 1:
        raise ValueError("logistic_omega_step needs at least one column")
    dev = _device_of(X)
    if p > 64:
        return _wide_omega_step(X, beta, n_trials, random_state, method, ref_compat, return_omega, return_psi,
                                first_index, dev)
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


def _wide_omega_step(X, beta, n_trials, random_state, method, ref_compat, return_omega, return_psi, first_index, dev):
    """p > 64: the fused kernels hold the upper triangle of G in registers (72 accumulators per lane at p = 64)
    and do not go wider.  psi and G become two library GEMMs (cuBLAS FP64 through torch); the draw in between
    is the same call, with the same streams, as for narrow X."""
    from ._api import random_polyagamma

    torch = _torch()
    beta = torch.as_tensor(beta, dtype=torch.float64, device=dev).contiguous()
    if beta.shape != (X.shape[1],):
        raise ValueError(f"beta must have shape ({X.shape[1]},)")
    psi = X @ beta
    if isinstance(n_trials, (int, float)):
        nt = float(n_trials)
    else:
        nt = torch.as_tensor(n_trials, dtype=torch.float64, device=dev).contiguous()
        if nt.shape != (X.shape[0],):
            raise ValueError(f"n_trials must be a scalar or have shape ({X.shape[0]},)")
    omega = random_polyagamma(nt, psi, method=method, random_state=random_state, disable_checks=True,
                              ref_compat=ref_compat, first_index=int(first_index))
    gram = X.T @ (omega[:, None] * X)
    gram = 0.5 * (gram + gram.T)  # (exactly symmetric, like the fused kernel's result)
    out = (gram, omega if return_omega else None)
    return out + (psi,) if return_psi else out


def allreduce_gram(gram, group=None):
    """Sum the p x p partial Gram matrices of the ranks of ``group`` in place (``torch.distributed``; NCCL for
    CUDA tensors).  The one exchange of a row-sharded Gibbs sweep: p * p doubles per sweep, whatever n is.  The
    sum is taken in the collective's order, so it agrees with the single-device matrix to rounding (1e-13
    relative), not bit for bit; every rank receives the same bits."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(gram, op=dist.ReduceOp.SUM, group=group)
    return gram


def sharded_logistic_omega_step(X_rows, beta, n_trials=1.0, *, row_offset, random_state, group=None, method=None,
                                ref_compat=False, return_omega=True):
    """One rank's part of the omega-step when the rows of X are sharded over the ranks of ``group``.

    ``X_rows`` holds rows ``row_offset .. row_offset + len(X_rows)`` of the global design matrix and ``beta`` the
    full coefficient vector; every rank passes the same ``random_state``.  The rank draws its rows' omegas with
    ``first_index = row_offset`` -- the Philox streams of the global rows, so the concatenated omegas are
    bit-identical to the single-device draw -- forms its partial Gram matrix and all-reduces it.  Returns
    ``(G_global, omega_rows)``."""
    if random_state is None:
        raise ValueError("sharded_logistic_omega_step needs an explicit random_state shared by all ranks")
    gram, omega = logistic_omega_step(X_rows, beta, n_trials, random_state=random_state, method=method,
                                      ref_compat=ref_compat, return_omega=return_omega, first_index=int(row_offset))
    return allreduce_gram(gram, group), omega
