"""GPU (-m gpu): the reference's own API-contract test (tests/test_polyagamma.py:14-206) run
against polyagamma_b200 -- same calls, same expectations."""
import array
import functools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pg():
    import polyagamma_b200 as pg

    return pg


def test_polyagamma(pg):
    random_polyagamma = pg.random_polyagamma
    rng = np.random.default_rng(1)
    rng_polyagamma = functools.partial(random_polyagamma, random_state=rng)

    assert len(rng_polyagamma(size=5)) == 5
    assert rng_polyagamma(size=(5, 2)).shape == (5, 2)
    with pytest.raises(TypeError, match="object cannot be interpreted as an integer"):
        assert rng_polyagamma(size=(5.1, 2.9))
    with pytest.raises(TypeError):
        assert rng_polyagamma(size=10.4)

    h = [[[0.59103028], [0.15228518], [0.53494081], [0.85875483], [0.05796053]],
         [[0.63934113], [0.1341983], [0.73854957], [0.76351291], [0.38431413]],
         [[0.90747263], [0.19601777], [0.04178972], [0.08220703], [0.55739387]],
         [[0.83603371], [0.04006968], [0.36563885], [0.96736881], [0.61390783]]]
    z = [-8.2654694, 4.72289344, -1.07163445, 8.00492261]
    assert rng_polyagamma(h, z).shape == (4, 5, 4)
    z = np.array([1, 2, 3, 4, 6])
    assert rng_polyagamma(h, z).shape == (4, 5, 5)
    assert rng_polyagamma(h, 0.12345).shape == (4, 5, 1)

    assert rng_polyagamma(z, size=(10, 1, 5)).shape == (10, 1, 5)
    assert rng_polyagamma(h, z, size=(2, 4, 5, 5)).shape == (2, 4, 5, 5)
    with pytest.raises(ValueError, match="operands could not be broadcast together"):
        rng_polyagamma(z, size=(10, 1, 4))
    with pytest.raises(TypeError):
        rng_polyagamma(z, size={10, 1, 4})

    h = [[1000, 2, 3], [3, 3, 1]]
    out = rng_polyagamma(h)
    assert out.shape == (2, 3)
    assert not np.allclose(out, 0)
    z = (1000, 2, -10, 0)
    out = rng_polyagamma(z=z)
    assert out.shape == (4,)
    assert not np.allclose(out, 0)

    h = (1, 2, 3, 4, 5, 6)
    z = np.ones(6)
    out = rng_polyagamma(h, z)
    assert out.shape == (6,)

    out = np.zeros(6)
    h = (1, 1, 1, 1, 1, 1)
    ret = rng_polyagamma(h, z, out=out)
    assert ret is out
    assert not np.allclose(out, 0)
    with pytest.raises(ValueError, match="operands could not be broadcast together"):
        rng_polyagamma(h, z, out=out[1:])
    out = np.array([0., 0., 0., 0., 0.])
    rng_polyagamma(out=out)
    assert not np.allclose(out, 0)
    out2 = np.array([0., 0., 0., 0., 0., 0.])
    rng_polyagamma(h, out=out2)
    assert not np.allclose(out2, 0)
    with pytest.raises(ValueError, match="operands could not be broadcast together"):
        rng_polyagamma(h, out=out)
    out = np.zeros((2, 1))
    rng_polyagamma(out=out)
    assert out.all()
    out3 = array.array('d', [0] * 5)
    assert not all(out3)
    rng_polyagamma(out=out3)
    assert all(out3)
    out4 = array.array('f', [0] * 2)
    rng_polyagamma(out=out4)
    assert all(out4)
    with pytest.raises(TypeError, match=r"Cannot cast array data from dtype\('complex64'\)*"):
        rng_polyagamma(out=np.empty(2, dtype=np.complex64))

    with pytest.raises(ValueError, match="`h` must be positive"):
        rng_polyagamma(0)
    with pytest.raises(ValueError, match="`h` must be positive"):
        rng_polyagamma(-1.3232)
    with pytest.raises(ValueError, match="`h` must be positive"):
        rng_polyagamma([1, 2, 3, 0.00001])
    with pytest.raises(ValueError, match="`h` must be positive"):
        rng_polyagamma([-1, 4.0], method="devroye")

    rng_polyagamma(z=-10.5)

    with pytest.raises(ValueError, match="`method` must be one of"):
        rng_polyagamma(method="unknown method")
    rng_polyagamma(method="gamma")
    rng_polyagamma(method="devroye")
    rng_polyagamma(method="alternate")
    rng_polyagamma(method="saddle")
    h = (1, 2, 3)
    rng_polyagamma(h, method="devroye")
    rng_polyagamma(h, method="alternate")
    random_polyagamma(np.ones(1), method="devroye")
    random_polyagamma(np.array([1., 2, 3.0]), method="devroye")

    with pytest.raises(ValueError, match="must be positive, and also integer"):
        rng_polyagamma(2.0000000001, method="devroye")
    rng_polyagamma(2.0000000000, method="devroye")

    with pytest.raises(TypeError, match="positional argument"):
        rng_polyagamma(1, 0, 5)

    rng_polyagamma(1.5, method="devroye", disable_checks=True)
    assert np.isnan(rng_polyagamma(h=-np.inf, disable_checks=True))
    assert np.isinf(rng_polyagamma(h=np.inf, disable_checks=True))
    assert np.isnan(rng_polyagamma(h=np.nan, disable_checks=True))
    assert np.isnan(rng_polyagamma(h=-12.5, disable_checks=True))
    assert np.isnan(rng_polyagamma(h=0, disable_checks=True))

    rng = np.random.default_rng(12345)
    expected = random_polyagamma(size=5, random_state=rng)
    rng2 = np.random.default_rng(12345)
    assert np.allclose(expected, random_polyagamma(size=5, random_state=rng2))
    assert not np.allclose(expected, random_polyagamma(size=5))


def test_return_types_and_out_semantics(pg):
    """SURVEY 8b [measured] facts about the reference's call contract."""
    import torch

    assert isinstance(pg.random_polyagamma(), float)
    assert isinstance(pg.random_polyagamma(2, 1.5), float)
    r = pg.random_polyagamma(np.int64(2), 1.0)
    assert isinstance(r, np.ndarray) and r.shape == ()
    assert pg.random_polyagamma(size=()).shape == ()
    assert pg.random_polyagamma(size=0).shape == (0,)
    out = np.zeros(4)
    res = pg.random_polyagamma(size=3, out=out)        # `size` wins over `out`
    assert res.shape == (3,) and not out.any()
    out_i = np.zeros(5, dtype=np.int64)
    pg.random_polyagamma(400.0, 0.0, out=out_i)          # cast-back into int64
    assert (out_i > 50).all()
    out_f = np.zeros((3, 4), order="F")
    pg.random_polyagamma(out=out_f)
    assert out_f.all()
    strided = np.zeros(10)
    pg.random_polyagamma(out=strided[::2])               # the reference mis-fills this; we fill it
    assert strided[::2].all() and not strided[1::2].any()
    # device arrays in -> device tensor out; __cuda_array_interface__ and DLPack accepted
    h = torch.full((6,), 2.0, dtype=torch.float64, device="cuda")
    res = pg.random_polyagamma(h, 1.0, random_state=(1, 1))
    assert isinstance(res, torch.Tensor) and res.is_cuda and res.dtype == torch.float64

    class CAI:
        def __init__(self, t):
            self.t = t
            self.__cuda_array_interface__ = t.__cuda_array_interface__

    res2 = pg.random_polyagamma(CAI(h), 1.0, random_state=(1, 1))
    assert torch.equal(res, res2)

    class DLP:  # a foreign array that only speaks DLPack
        def __init__(self, t):
            self.t = t

        def __dlpack__(self, **kw):
            return self.t.__dlpack__(**kw)

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()

    res3 = pg.random_polyagamma(DLP(h), 1.0, random_state=(1, 1))
    assert isinstance(res3, torch.Tensor) and torch.equal(res, res3)
    d1 = pg.polyagamma_pdf(DLP(res), CAI(h), 1.0)
    assert d1.is_cuda and torch.equal(d1, pg.polyagamma_pdf(res, h, 1.0))
    dout = torch.zeros(6, dtype=torch.float64, device="cuda")
    assert pg.random_polyagamma(h, 1.0, out=dout, random_state=(1, 1)) is dout
    assert torch.equal(dout, res)
    dout32 = torch.zeros(6, dtype=torch.float32, device="cuda")
    pg.random_polyagamma(h, 1.0, out=dout32, random_state=(1, 1))
    assert torch.allclose(dout32.double(), res, rtol=1e-6)
    with pytest.raises(ValueError, match="`h` must be positive"):
        pg.random_polyagamma(torch.tensor([1.0, -2.0], device="cuda", dtype=torch.float64))
    pg.release_scratch()  # cached pass scratch goes back to the driver; the next call re-allocates
    assert pg.random_polyagamma(h, 1.0, random_state=(1, 1)).equal(res)
    g = pg.PhiloxGenerator(11)
    a = pg.random_polyagamma(size=4, random_state=g)
    b = pg.random_polyagamma(size=4, random_state=g)
    assert not np.allclose(a, b) and g.offset == 2
    assert np.array_equal(a, pg.random_polyagamma(size=4, random_state=(11, 0)))


# "devroye" is not included because it does not play well with non-integer h
@pytest.mark.parametrize("method", ("alternate", "saddle", "gamma"))
@pytest.mark.parametrize("h", (0.5, 1, 4, 7, 15, 25))
@pytest.mark.parametrize("z", (0, 1, -4, 7, -15, 25))
def test_polyagamma_pdf_cdf(pg, method, h, z):
    # tests/test_polyagamma.py:166-206
    rng = np.random.default_rng(1)
    x = pg.random_polyagamma(h, z, size=5000, method=method, random_state=rng)
    x.sort()
    d = pg.polyagamma_pdf(x, h=h, z=z)
    assert np.isclose(1.0, np.trapezoid(d, x), rtol=1e-2)
    xx = x.mean()
    mask = x <= xx
    cdf = pg.polyagamma_cdf(xx, h=h, z=z)
    assert np.allclose(np.trapezoid(d[mask], x[mask]), cdf, rtol=1e-2)
    ld = np.exp(pg.polyagamma_pdf(x, h=h, z=z, return_log=True))
    assert np.allclose(ld, d)
    lc = np.exp(pg.polyagamma_cdf(xx, h=h, z=z, return_log=True))
    assert np.allclose(lc, cdf)
