"""CPU: the C-ABI library loads without a GPU and exports what include/pgm_cuda.h declares;
host-side logic of the Python layer (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pgm_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pgm_cuda_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from polyagamma_b200 import _lib

    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    assert sorted(_lib.EXPORTS) == declared
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.pgm_cuda_version()


def test_header_cites_reference_and_has_no_torch():
    text = open(os.path.join(ROOT, "include", "pgm_cuda.h")).read()
    assert "include/pgm_random.h:47-74" in text and "include/pgm_density.h:4-14" in text
    assert "torch" not in text.split("*/", 1)[1].lower()


def test_sampler_enum_values_match_reference():
    # include/pgm_random.h:13  typedef enum {GAMMA, DEVROYE, ALTERNATE, SADDLE, HYBRID}
    from polyagamma_b200 import _lib

    assert (_lib.GAMMA, _lib.DEVROYE, _lib.ALTERNATE, _lib.SADDLE, _lib.HYBRID) == (0, 1, 2, 3, 4)


def test_host_side_helpers_match_oracle(oracle):
    from polyagamma_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(3)
    for h, z in zip(np.r_[rng.uniform(0.1, 80, 200), [1, 2, 3, 4, 8, 50, 51]], rng.uniform(-10, 10, 207)):
        assert L.pgm_cuda_hybrid_select(float(h), float(z)) == oracle.hybrid_select(h, z)
    for seed, off in ((0, 0), (12345, 0), (2**64 - 1, 2**40 + 3), (7, 1)):
        key = (ctypes.c_uint32 * 2)()
        L.pgm_cuda_call_key(seed, off, key)
        assert list(key) == [int(v) for v in oracle.call_key(seed, off)]


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "polyagamma_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+(oracle|ref)\b", text, flags=re.M), f
                assert "libpgm_oracle" not in text and "pgm_oracle.h" not in text and "libpgmref" not in text, f


def test_missing_extension_fails_loudly(monkeypatch):
    from polyagamma_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpgm_cuda.so")
    with pytest.raises(_lib.ExtensionMissing, match="no CPU fallback"):
        _lib.lib()


def test_no_gpu_means_error_not_fallback():
    import torch

    import polyagamma_b200 as pg

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pg.random_polyagamma(1.0, 0.0, size=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pg.polyagamma_pdf(np.ones(3))


def test_argument_validation_precedes_compute():
    """The reference's Python-level errors (tests/test_polyagamma.py:101-118,138-145) are
    raised before any device work, so they can be checked without a GPU."""
    import polyagamma_b200 as pg

    with pytest.raises(ValueError, match="`method` must be one of"):
        pg.random_polyagamma(method="unknown method")
    for bad in (0, -1.3232):
        with pytest.raises(ValueError, match="`h` must be positive"):
            pg.random_polyagamma(bad)
    with pytest.raises(ValueError, match="`h` must be positive"):
        pg.random_polyagamma([1, 2, 3, 0.00001])
    with pytest.raises(ValueError, match="`h` must be positive"):
        pg.random_polyagamma([-1, 4.0], method="devroye")
    with pytest.raises(ValueError, match="must be positive, and also integer"):
        pg.random_polyagamma(2.0000000001, method="devroye")
    with pytest.raises(TypeError, match="positional argument"):
        pg.random_polyagamma(1, 0, 5)


def test_size_and_broadcast_helpers():
    from polyagamma_b200 import _api

    assert _api._shape_from_size(5) == (5,)
    assert _api._shape_from_size((5, 2)) == (5, 2)
    assert _api._shape_from_size(()) == ()
    with pytest.raises(TypeError, match="object cannot be interpreted as an integer"):
        _api._shape_from_size((5.1, 2.9))
    with pytest.raises(TypeError):
        _api._shape_from_size(10.4)
    with pytest.raises(TypeError):
        _api._shape_from_size({10, 1, 4})
    import torch

    h = torch.zeros(4, 5, 1, dtype=torch.float64)
    z = torch.zeros(4, dtype=torch.float64)
    assert _api._broadcast_shape(tuple(h.shape), tuple(z.shape)) == (4, 5, 4)
    assert _api._broadcast_strides(h, (4, 5, 4)) == [5, 1, 0]
    assert _api._broadcast_strides(z, (4, 5, 4)) == [0, 0, 1]
    with pytest.raises(ValueError, match="operands could not be broadcast together"):
        _api._broadcast_shape((5,), (10, 1, 4))


def test_random_state_resolution():
    from polyagamma_b200 import PhiloxGenerator, _api

    assert _api._resolve_random_state(12345) == (12345, 0)
    assert _api._resolve_random_state((9, 4)) == (9, 4)
    a = _api._resolve_random_state(np.random.default_rng(12345))
    b = _api._resolve_random_state(np.random.default_rng(12345))
    assert a == b and a != _api._resolve_random_state(np.random.default_rng(12346))
    rng = np.random.default_rng(1)
    assert _api._resolve_random_state(rng) != _api._resolve_random_state(rng)  # generator advances
    g = PhiloxGenerator(5)
    assert [_api._resolve_random_state(g) for _ in range(3)] == [(5, 0), (5, 1), (5, 2)]
    assert _api._resolve_random_state(None) != _api._resolve_random_state(None)
    assert _api._resolve_random_state(np.random.SeedSequence(3)) == _api._resolve_random_state(np.random.SeedSequence(3))
    assert _api._resolve_random_state([1, 2, 3]) == _api._resolve_random_state([1, 2, 3])


def test_shard_range_partitions():
    from polyagamma_b200.sharding import shard_range

    for n in (0, 1, 7, 10**10, 1_000_003):
        for world in (1, 2, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_argument_checks_need_no_gpu():
    """Bad arguments are rejected before anything touches the device: NULL shape / strides, an element count
    that overflows, an unknown `which`."""
    import ctypes

    from polyagamma_b200 import _lib

    L = _lib.lib()
    i64 = ctypes.c_int64
    shape = (i64 * 2)(1 << 40, 1 << 40)
    st = (i64 * 2)(0, 0)
    one = ctypes.c_void_p(8)  # never dereferenced
    assert L.pgm_cuda_random_polyagamma_fill2_strided(1, 0, one, one, _lib.HYBRID, 2, None, st, st, one, 0, None) == _lib.EINVAL
    assert L.pgm_cuda_random_polyagamma_fill2_strided(1, 0, one, one, _lib.HYBRID, 2, shape, None, st, one, 0, None) == _lib.EINVAL
    assert L.pgm_cuda_random_polyagamma_fill2_strided(1, 0, one, one, _lib.HYBRID, 2, shape, st, st, one, 0, None) == _lib.EINVAL
    assert L.pgm_cuda_polyagamma_density_strided(0, one, one, one, 2, shape, st, st, st, one, None) == _lib.EINVAL
    assert L.pgm_cuda_polyagamma_density_strided(0, one, one, one, 1, None, st, st, st, one, None) == _lib.EINVAL
    assert L.pgm_cuda_polyagamma_density_strided(8, one, one, one, 0, shape, st, st, st, one, None) == _lib.EINVAL
    neg = (i64 * 1)(-3)
    assert L.pgm_cuda_random_polyagamma_fill2_strided(1, 0, one, one, _lib.HYBRID, 1, neg, st, st, one, 0, None) == _lib.EINVAL
    # host-pointer entry points: NULL arrays, bad strides, duplicate devices; an empty call is a no-op
    assert L.pgm_cuda_download(None, one, 5, None) == _lib.EINVAL
    assert L.pgm_cuda_download(one, None, 5, None) == _lib.EINVAL
    assert L.pgm_cuda_download(None, None, 0, None) == 0
    assert L.pgm_cuda_random_polyagamma_fill2_host(1, 0, None, 1, one, 1, _lib.HYBRID, 4, one, 0, 0) == _lib.EINVAL
    assert L.pgm_cuda_random_polyagamma_fill2_host(1, 0, one, 2, one, 1, _lib.HYBRID, 4, one, 0, 0) == _lib.EINVAL
    assert L.pgm_cuda_random_polyagamma_fill2_host(1, 0, one, 1, one, 1, _lib.HYBRID, 0, None, 0, 0) == 0
    devs = (ctypes.c_int * 2)(0, 0)
    assert L.pgm_cuda_random_polyagamma_fill2_host_multi(1, 0, one, 1, one, 1, _lib.HYBRID, 4, one, 0, devs, 2) == _lib.EINVAL
    assert L.pgm_cuda_polyagamma_density_host(0, None, 1, one, 1, one, 1, 3, one, -1) == _lib.EINVAL
    assert L.pgm_cuda_polyagamma_density_host(9, one, 1, one, 1, one, 1, 3, one, -1) == _lib.EINVAL
