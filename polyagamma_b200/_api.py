"""Python face of the B200 Polya-Gamma path: the reference's call contract over GPU kernels.

Mirrors ``polyagamma/_polyagamma.pyx`` (``rvs`` :22-256, ``pdf`` :259-324, ``cdf`` :327-384,
``dispatch`` :433-472): same names, argument meaning, broadcasting, ``size`` / ``out``
semantics and exceptions.  The arithmetic happens in ``libpgm_cuda.so`` (``include/pgm_cuda.h``);
torch is used only to own device memory and streams.  There is no CPU fallback.

Differences a reference user should know about:

* randomness is a Philox stream named by ``(seed, offset)``; ``random_state`` accepts what the
  reference accepts (None / int / sequence / SeedSequence / BitGenerator / Generator -- one
  64-bit draw is taken from a numpy generator to seed the call) plus ``(seed, offset)`` tuples,
  ``torch.Generator`` and :class:`PhiloxGenerator`;
* inputs may be torch CUDA tensors, ``__cuda_array_interface__`` or DLPack objects; the result
  is then a torch CUDA tensor.  Host inputs give numpy results as in the reference, unless
  ``device=`` asks for the result to stay on the GPU;
* ``ref_compat=True`` reproduces two envelope constants of the reference that bias its
  ``alternate`` and ``saddle`` samplers (SURVEY.md 8-DEFECTS a, i); the default is the
  mathematically consistent envelope.
"""
from __future__ import annotations

import ctypes
import numbers
import operator
import os

import numpy as np

from . import _lib

__all__ = ["random_polyagamma", "polyagamma_pdf", "polyagamma_cdf", "PhiloxGenerator"]

# polyagamma/_polyagamma.pyx:392-398
METHODS = {
    None: _lib.HYBRID,
    "gamma": _lib.GAMMA,
    "saddle": _lib.SADDLE,
    "devroye": _lib.DEVROYE,
    "alternate": _lib.ALTERNATE,
    # not in the reference: an EXACT sampler for every h >= 1 -- the alternate method's pieces,
    # PG(h, z) = sum of PG(b_i, z) with b_i <= 4 (src/pgm_alternate.c:294-323), with the mathematically consistent
    # envelope whatever ref_compat says; the hybrid rule's saddle / normal methods are approximations
    "exact": _lib.EXACT,
}
ZERO = 1e-04  # polyagamma/_polyagamma.pyx:399
PARAM_CHECK_ERROR_MSG = "Values of `h` must be positive, and also integer if `method` is devroye."
EXACT_ERROR_MSG = "`method='exact'` needs h >= 1 (the alternate sampler's pieces are exact from h = 1 on)."

_MASK64 = (1 << 64) - 1


class PhiloxGenerator:
    """Resumable random state of this package: ``(seed, offset)``; ``offset`` advances by one
    per sampling call (the reference keeps its state inside a numpy BitGenerator)."""

    def __init__(self, seed=None, offset=0):
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        self.seed = int(seed) & _MASK64
        self.offset = int(offset) & _MASK64

    def next_call(self):
        s = (self.seed, self.offset)
        self.offset = (self.offset + 1) & _MASK64
        return s

    def __repr__(self):
        return f"PhiloxGenerator(seed={self.seed}, offset={self.offset})"


def _torch():
    import torch

    return torch


def _resolve_random_state(random_state):
    """(seed, offset) for one call.  polyagamma/_polyagamma.pyx:92-99,158."""
    if random_state is None:
        return int.from_bytes(os.urandom(8), "little"), 0
    if isinstance(random_state, PhiloxGenerator):
        return random_state.next_call()
    if isinstance(random_state, tuple) and len(random_state) == 2 and all(
            isinstance(v, numbers.Integral) for v in random_state):
        return int(random_state[0]) & _MASK64, int(random_state[1]) & _MASK64
    if isinstance(random_state, numbers.Integral):
        v = int(random_state)
        if v < 0:
            raise ValueError("expected non-negative integer")
        if v <= _MASK64:
            return v, 0
        return int(np.random.SeedSequence(v).generate_state(1, np.uint64)[0]), 0
    if isinstance(random_state, np.random.Generator):
        return int(random_state.bit_generator.random_raw()), 0
    if isinstance(random_state, np.random.BitGenerator):
        return int(random_state.random_raw()), 0
    if isinstance(random_state, np.random.SeedSequence):
        return int(random_state.generate_state(1, np.uint64)[0]), 0
    try:
        torch = _torch()
        if isinstance(random_state, torch.Generator):
            v = torch.randint(0, 2**62, (2,), generator=random_state,
                              device=random_state.device, dtype=torch.int64)
            return int(v[0].item()), int(v[1].item())
    except ImportError:  # pragma: no cover
        pass
    # sequence of ints etc.: whatever numpy's default_rng accepts
    return int(np.random.default_rng(random_state).bit_generator.random_raw()), 0


def _is_number(o) -> bool:
    """is_number (polyagamma/_polyagamma.pyx:406-408): Python float / int only."""
    return isinstance(o, (float, int))


def _is_device_array(o) -> bool:
    if hasattr(o, "__cuda_array_interface__"):
        return True
    try:
        torch = _torch()
    except ImportError:  # pragma: no cover
        return False
    if isinstance(o, torch.Tensor):
        return o.is_cuda
    if hasattr(o, "__dlpack_device__") and not isinstance(o, np.ndarray):
        try:
            return int(o.__dlpack_device__()[0]) == 2  # kDLCUDA
        except Exception:
            return False
    return False


def _as_device_tensor(o, device):
    """float64 torch CUDA tensor viewing / copying ``o`` (device array, host array or scalar)."""
    torch = _torch()
    if isinstance(o, torch.Tensor):
        t = o
    elif hasattr(o, "__cuda_array_interface__"):
        t = torch.as_tensor(o, device=device)
    elif hasattr(o, "__dlpack__") and not isinstance(o, np.ndarray):
        t = torch.from_dlpack(o)
    else:
        t = torch.from_numpy(np.asarray(o, dtype=np.float64, order="C"))
    return t.to(device=device, dtype=torch.float64)


def _require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("polyagamma_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def _device_of(*objs, device=None):
    torch = _require_cuda()
    if device is not None:
        return torch.device(device)
    for o in objs:
        if isinstance(o, torch.Tensor) and o.is_cuda:
            return o.device
    for o in objs:
        if hasattr(o, "__cuda_array_interface__") and not isinstance(o, torch.Tensor):
            return torch.as_tensor(o).device
    return torch.device("cuda", torch.cuda.current_device())


def _stream_ptr(device):
    torch = _torch()
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _shape_from_size(size):
    """np.PyArray_IntpConverter: int or sequence of ints; TypeError otherwise."""
    if isinstance(size, numbers.Integral):
        return (operator.index(size),)
    if isinstance(size, (str, bytes, set, frozenset, dict)):
        raise TypeError("expected a sequence of integers or a single integer")
    try:
        return tuple(operator.index(s) for s in size)
    except TypeError:
        if hasattr(size, "__iter__"):
            raise
        return (operator.index(size),)  # raises "... cannot be interpreted as an integer"


def _broadcast_strides(t, shape):
    """Element strides of tensor ``t`` broadcast to ``shape`` (0 on broadcast axes)."""
    nd = len(shape)
    tshape = (1,) * (nd - t.dim()) + tuple(t.shape)
    tstride = (0,) * (nd - t.dim()) + tuple(t.stride())
    out = []
    for want, have, st in zip(shape, tshape, tstride):
        if have == want:
            out.append(st if want != 1 else 0)
        elif have == 1:
            out.append(0)
        else:
            raise ValueError("operands could not be broadcast together with shapes "
                             f"{tuple(t.shape)} {tuple(shape)}")
    return out


def _broadcast_shape(*shapes):
    try:
        return tuple(np.broadcast_shapes(*shapes))
    except ValueError:
        raise ValueError("operands could not be broadcast together with shapes " +
                         " ".join(str(tuple(s)) for s in shapes)) from None


def _method_id(method, ref_compat):
    try:
        known = method in METHODS
    except TypeError:
        known = False
    if not known:
        names = {k for k in METHODS}
        raise ValueError(f"`method` must be one of {names}")
    return METHODS[method] | (_lib.REF_COMPAT if ref_compat else 0)


# ------------------------------------------------------------------------------------------
# sampling
# ------------------------------------------------------------------------------------------
def _fill_scalar_device(seed, offset, h, z, mid, n, device, first_index=0, out=None):
    torch = _torch()
    if out is None:
        out = torch.empty(n, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        rc = _lib.lib().pgm_cuda_random_polyagamma_fill(seed, offset, float(h), float(z), mid, n,
                                                        ctypes.c_void_p(out.data_ptr()), first_index,
                                                        _stream_ptr(device))
    _lib.check(rc, "pgm_cuda_random_polyagamma_fill")
    return out


def _fill_arrays_device(seed, offset, ht, zt, shape, mid, device, out=None, first_index=0):
    """fill2 over the broadcast of device tensors ht, zt to ``shape`` (no materialisation)."""
    torch = _torch()
    if out is None:
        out = torch.empty(shape, dtype=torch.float64, device=device)
    hs = _broadcast_strides(ht, shape)
    zs = _broadcast_strides(zt, shape)
    with torch.cuda.device(device):
        rc = _lib.lib().pgm_cuda_random_polyagamma_fill2_strided(
            seed, offset, ctypes.c_void_p(ht.data_ptr()), ctypes.c_void_p(zt.data_ptr()), mid, len(shape),
            _lib.int64_array(shape), _lib.int64_array(hs), _lib.int64_array(zs),
            ctypes.c_void_p(out.data_ptr()), first_index, _stream_ptr(device))
    _lib.check(rc, "pgm_cuda_random_polyagamma_fill2_strided")
    return out


def _check_h_device(ht, mid, device):
    torch = _torch()
    flag = ctypes.c_int(0)
    hc = ht.contiguous()
    with torch.cuda.device(device):
        rc = _lib.lib().pgm_cuda_any_invalid_h(ctypes.c_void_p(hc.data_ptr()), hc.numel(), mid & 0xff,
                                               ctypes.byref(flag), _stream_ptr(device))
    _lib.check(rc, "pgm_cuda_any_invalid_h")
    return bool(flag.value)


def _check_h_host(h, mid):
    """any_nonpositive (polyagamma/_polyagamma.pyx:411-430)."""
    h = np.asarray(h, dtype=np.float64)
    bad = h <= ZERO
    if (mid & 0xff) == _lib.DEVROYE:
        with np.errstate(invalid="ignore"):
            bad = bad | (h != np.floor(h))
    return bool(np.any(bad))


def _host_out_array(out):
    """The double array behind a host ``out`` object and whether a cast-back is needed
    (np.PyArray_FROM_OTF(out, NPY_DOUBLE, NPY_ARRAY_WRITEBACKIFCOPY),
    polyagamma/_polyagamma.pyx:183,231)."""
    arr = np.asarray(out)
    if not np.can_cast(arr.dtype, np.float64, "safe"):
        raise TypeError(f"Cannot cast array data from {arr.dtype!r} to dtype('float64') "
                        "according to the rule 'safe'")
    return arr


def random_polyagamma(h=1.0, z=0.0, *, size=None, out=None, method=None, disable_checks=False,
                      random_state=None, device=None, ref_compat=False, first_index=0, devices=None):
    """Draw from PG(h, z).  Same contract as the reference's ``random_polyagamma``
    (``polyagamma/_polyagamma.pyx:22-256``); see the module docstring for the extensions.

    ``first_index`` is the global index of the first output element: a shard that passes the
    same ``random_state`` and its own ``first_index`` draws exactly what a single call over the
    whole index range would draw for those elements (see ``polyagamma_b200.sharding``).

    ``devices`` (host operands only): a sequence of CUDA device indices, or ``"all"``.  The index range is
    cut into one contiguous shard per device and all devices copy and compute at once, from this one call
    and host thread; the result is bit-identical to the single-device one."""
    mid = _method_id(method, ref_compat)
    exact = (mid & 0xff) == _lib.EXACT
    if exact:
        mid = _lib.ALTERNATE  # (default envelope: ref_compat is ignored)
    fi = int(first_index)
    seed, offset = _resolve_random_state(random_state)
    has_size = size is not None
    has_out = out is not None

    # ---- scalar fast path (polyagamma/_polyagamma.pyx:163-193) ----
    if _is_number(h) and _is_number(z):
        ch, cz = float(h), float(z)
        if not disable_checks and (ch <= ZERO or ((mid & 0xff) == _lib.DEVROYE and ch != np.floor(ch))):
            raise ValueError(PARAM_CHECK_ERROR_MSG)
        if exact and not ch >= 1.0:
            raise ValueError(EXACT_ERROR_MSG)
        if has_size:
            shape = _shape_from_size(size)
            n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
            if any(s < 0 for s in shape):
                raise ValueError("negative dimensions are not allowed")
            dev = _device_of(device=device)
            res = _fill_scalar_device(seed, offset, ch, cz, mid, n, dev, fi).reshape(shape)
            return res if device is not None else _to_numpy(res, dev)
        if has_out:
            if _is_device_array(out):
                dev = _device_of(out, device=device)
                ot = _as_device_tensor_view(out, dev)
                torch = _torch()
                if ot.dtype == torch.float64 and ot.is_contiguous():  # drawn in place
                    _fill_scalar_device(seed, offset, ch, cz, mid, ot.numel(), dev, fi, out=ot)
                else:
                    res = _fill_scalar_device(seed, offset, ch, cz, mid, ot.numel(), dev, fi)
                    ot.copy_(res.reshape(ot.shape))
                return out
            arr = _host_out_array(out)
            # the pipelined host-buffer entry point with both parameters as one-element operands: the
            # device->host copies of the result overlap the kernels
            res = _fill_arrays_host(seed, offset, np.asarray(ch, dtype=np.float64), np.asarray(cz, dtype=np.float64),
                                    tuple(arr.shape), mid, arr, fi)
            if res is not arr:
                arr[...] = res
            return out
        torch = _require_cuda()
        with torch.cuda.device(_device_of(device=device)):
            if fi:
                return float(_fill_scalar_device(seed, offset, ch, cz, mid, 1, _device_of(device=device), fi).item())
            return float(_lib.lib().pgm_cuda_random_polyagamma(seed, offset, ch, cz, mid))

    # ---- array path (polyagamma/_polyagamma.pyx:195-256) ----
    on_device = _is_device_array(h) or _is_device_array(z) or (has_out and _is_device_array(out))
    if on_device or device is not None:
        dev = _device_of(h, z, out, device=device)
        ht = _as_device_tensor(h, dev)
        zt = _as_device_tensor(z, dev)
        if not disable_checks and _check_h_device(ht, mid, dev):
            raise ValueError(PARAM_CHECK_ERROR_MSG)
        if exact and not disable_checks and bool((ht < 1.0).any()):
            raise ValueError(EXACT_ERROR_MSG)
        if has_size:
            shape = _broadcast_shape(tuple(ht.shape), tuple(zt.shape), _shape_from_size(size))
            if shape != _shape_from_size(size):
                raise ValueError("operands could not be broadcast together with remapped shapes "
                                 f"{tuple(ht.shape)} {tuple(zt.shape)} and requested shape {_shape_from_size(size)}")
            return _fill_arrays_device(seed, offset, ht, zt, shape, mid, dev, first_index=fi)
        if has_out:
            if _is_device_array(out):
                ot = _as_device_tensor_view(out, dev)
                shape = tuple(ot.shape)
                torch = _torch()
                direct = ot.dtype == torch.float64 and ot.is_contiguous()
                res = _fill_arrays_device(seed, offset, ht, zt, shape, mid, dev, out=ot if direct else None, first_index=fi)
                if not direct:
                    ot.copy_(res)
                return out
            arr = _host_out_array(out)
            res = _fill_arrays_device(seed, offset, ht, zt, tuple(arr.shape), mid, dev, first_index=fi)
            arr[...] = res.cpu().numpy()
            return out
        shape = _broadcast_shape(tuple(ht.shape), tuple(zt.shape))
        return _fill_arrays_device(seed, offset, ht, zt, shape, mid, dev, first_index=fi)

    # host operands -> host result
    harr = np.asarray(h, dtype=np.float64, order="C")
    if not disable_checks and _check_h_host(harr, mid):
        raise ValueError(PARAM_CHECK_ERROR_MSG)
    zarr = np.asarray(z, dtype=np.float64, order="C")
    if exact and not disable_checks and bool(np.any(harr < 1.0)):
        raise ValueError(EXACT_ERROR_MSG)
    arr = None
    if has_size:
        want = _shape_from_size(size)
        shape = _broadcast_shape(harr.shape, zarr.shape, want)
        if shape != want:
            raise ValueError("operands could not be broadcast together with remapped shapes "
                             f"{harr.shape} {zarr.shape} and requested shape {want}")
    elif has_out:
        arr = _host_out_array(out)
        shape = tuple(arr.shape)
        if _broadcast_shape(harr.shape, zarr.shape, shape) != shape:
            raise ValueError("operands could not be broadcast together with shapes "
                             f"{harr.shape} {zarr.shape} {shape}")
    else:
        shape = _broadcast_shape(harr.shape, zarr.shape)
    res = _fill_arrays_host(seed, offset, harr, zarr, shape, mid, arr, fi, devices)
    if arr is not None:
        if res is not arr:
            arr[...] = res
        return out
    return res


def _as_device_tensor_view(o, dev):
    """A torch tensor sharing memory with the device array ``o`` (no dtype conversion)."""
    torch = _torch()
    if isinstance(o, torch.Tensor):
        return o
    if hasattr(o, "__cuda_array_interface__"):
        return torch.as_tensor(o, device=dev)
    return torch.from_dlpack(o)


def _fill_arrays_host(seed, offset, harr, zarr, shape, mid, arr=None, first_index=0, devices=None):
    """Host operands.  When h and z are each a scalar or already of the full output shape the
    pipelined host-buffer entry point is used (copies overlap the kernels); other broadcasts
    upload the un-broadcast operands and read the result back."""
    torch = _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    if n == 0:
        return np.empty(shape, dtype=np.float64)

    def linear(a):
        return a.size == 1 or tuple(a.shape) == tuple(shape)

    if linear(harr) and linear(zarr):
        direct = arr is not None and arr.dtype == np.float64 and arr.flags.c_contiguous and arr.flags.writeable
        res = arr if direct else np.empty(shape, dtype=np.float64)
        if devices is not None:
            devs = list(range(torch.cuda.device_count())) if devices == "all" else [int(d) for d in devices]
            carr = (ctypes.c_int * len(devs))(*devs)
            rc = _lib.lib().pgm_cuda_random_polyagamma_fill2_host_multi(
                seed, offset, harr.ctypes.data, 0 if harr.size == 1 else 1, zarr.ctypes.data,
                0 if zarr.size == 1 else 1, mid, n, res.ctypes.data, first_index, carr, len(devs))
            _lib.check(rc, "pgm_cuda_random_polyagamma_fill2_host_multi")
            return res
        rc = _lib.lib().pgm_cuda_random_polyagamma_fill2_host(
            seed, offset, harr.ctypes.data, 0 if harr.size == 1 else 1, zarr.ctypes.data,
            0 if zarr.size == 1 else 1, mid, n, res.ctypes.data, first_index, dev.index)
        _lib.check(rc, "pgm_cuda_random_polyagamma_fill2_host")
        return res
    ht = torch.from_numpy(harr).to(dev)
    zt = torch.from_numpy(zarr).to(dev)
    return _to_numpy(_fill_arrays_device(seed, offset, ht, zt, tuple(shape), mid, dev, first_index=first_index), dev)


# ------------------------------------------------------------------------------------------
# densities
# ------------------------------------------------------------------------------------------
_WHICH = {("pdf", False): 0, ("cdf", False): 1, ("pdf", True): 2, ("cdf", True): 3}


def _density(kind, x, h, z, return_log, device, stable=False):
    """dispatch (polyagamma/_polyagamma.pyx:433-472)."""
    which = _WHICH[(kind, bool(return_log))] + (4 if stable else 0)
    all_scalar = _is_number(x) and _is_number(h) and _is_number(z)
    if all_scalar and device is None:
        # the scalar fast path of dispatch() (polyagamma/_polyagamma.pyx:437-441): one point, one launch that
        # reads and writes mapped pinned memory
        _require_cuda()
        buf = (ctypes.c_double * 4)(float(x), float(h), float(z), 0.0)
        base = ctypes.addressof(buf)
        rc = _lib.lib().pgm_cuda_polyagamma_density_host(which, base, 0, base + 8, 0, base + 16, 0, 1, base + 24, -1)
        _lib.check(rc, "pgm_cuda_polyagamma_density_host")
        return float(buf[3])
    on_device = any(_is_device_array(o) for o in (x, h, z)) or device is not None
    dev = _device_of(x, h, z, device=device)
    torch = _torch()
    xt = _as_device_tensor(x, dev)
    ht = _as_device_tensor(h, dev)
    zt = _as_device_tensor(z, dev)
    shape = _broadcast_shape(tuple(xt.shape), tuple(ht.shape), tuple(zt.shape))
    out = torch.empty(shape, dtype=torch.float64, device=dev)
    if out.numel():
        with torch.cuda.device(dev):
            rc = _lib.lib().pgm_cuda_polyagamma_density_strided(
                which, ctypes.c_void_p(xt.data_ptr()), ctypes.c_void_p(ht.data_ptr()),
                ctypes.c_void_p(zt.data_ptr()), len(shape), _lib.int64_array(shape),
                _lib.int64_array(_broadcast_strides(xt, shape)), _lib.int64_array(_broadcast_strides(ht, shape)),
                _lib.int64_array(_broadcast_strides(zt, shape)), ctypes.c_void_p(out.data_ptr()),
                _stream_ptr(dev))
        _lib.check(rc, "pgm_cuda_polyagamma_density_strided")
    if all_scalar and device is None:
        return float(out.item())
    if on_device:
        return out
    return _to_numpy(out, dev)


def _to_numpy(t, dev):
    """A device result as a new numpy array.  Large results go through the library's pinned bounce buffers and
    several host threads (``pgm_cuda_download``): a plain ``.cpu()`` into fresh pageable memory runs at the
    page-fault rate of one thread."""
    if t.numel() < (1 << 18) or not t.is_contiguous():
        return t.cpu().numpy()
    res = np.empty(tuple(t.shape), dtype=np.float64)
    with _torch().cuda.device(dev):
        rc = _lib.lib().pgm_cuda_download(ctypes.c_void_p(res.ctypes.data), ctypes.c_void_p(t.data_ptr()), t.numel(),
                                          _stream_ptr(dev))
    _lib.check(rc, "pgm_cuda_download")
    return res


def polyagamma_pdf(x, h=1.0, z=0.0, return_log=False, *, device=None, stable=False):
    """Density of PG(h, z) at x (``polyagamma/_polyagamma.pyx:259-324``).

    ``stable=True`` (not in the reference): for h >= 8 the value comes from a saddle-point contour integral
    instead of the reference's alternating series, which loses every digit for h >~ 40 at small |z|
    (SURVEY 8-DEFECTS g).  The default reproduces the reference's numbers, cancellation included."""
    return _density("pdf", x, h, z, return_log, device, stable)


def polyagamma_cdf(x, h=1.0, z=0.0, return_log=False, *, device=None, stable=False):
    """Distribution function of PG(h, z) at x (``polyagamma/_polyagamma.pyx:327-384``); ``stable`` as in
    :func:`polyagamma_pdf`."""
    return _density("cdf", x, h, z, return_log, device, stable)
