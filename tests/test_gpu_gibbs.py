"""GPU (-m gpu): the logistic-regression Gibbs omega-step against numpy + the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch

    return torch


@pytest.mark.parametrize("n,p", [(1, 1), (257, 3), (1_001, 5), (10_000, 8), (7_777, 12), (20_001, 17),
                                 (30_000, 32), (9_999, 33), (12_345, 47), (5_000, 64)])
def test_omega_step_matches_numpy_and_oracle(torch, oracle, n, p):
    from polyagamma_b200.gibbs import logistic_omega_step

    rng = np.random.default_rng(n + p)
    X = rng.normal(size=(n, p))
    beta = rng.normal(scale=0.7, size=p)
    nt = rng.integers(1, 6, n).astype(float)
    G, omega, psi = logistic_omega_step(torch.from_numpy(X).cuda(), torch.from_numpy(beta).cuda(),
                                        torch.from_numpy(nt).cuda(), random_state=(2024, 3), return_psi=True)
    psi_ref = X @ beta
    assert np.allclose(psi.cpu().numpy(), psi_ref, rtol=1e-13, atol=1e-13)
    # the draw is the library's own fill2 on (n_trials, psi): element for element against the oracle
    want = oracle.fill2(2024, 3, nt, psi.cpu().numpy())
    got = omega.cpu().numpy()
    agree = np.abs(got - want) <= 1e-9 * np.abs(want)
    assert agree.mean() >= 1 - 1e-3
    G_ref = X.T @ (got[:, None] * X)
    assert np.allclose(G.cpu().numpy(), G_ref, rtol=1e-11, atol=1e-11 * np.abs(G_ref).max())
    Gn = G.cpu().numpy()
    assert np.array_equal(Gn, Gn.T)  # only the upper triangle is computed, then mirrored


@pytest.mark.parametrize("p,pad", [(6, 0), (7, 1), (20, 1), (40, 0), (40, 3), (56, 2), (64, 1)])
def test_gram_variants(torch, p, pad):
    """Both staging granularities (16-byte copies need an even leading dimension and a 16-byte aligned
    base) for every column-block count, against numpy."""
    from polyagamma_b200.gibbs import logistic_omega_step

    n = 40_001
    rng = np.random.default_rng(p * 7 + pad)
    wide = torch.from_numpy(rng.normal(size=(n, p + pad))).cuda()
    X = wide[:, :p]
    beta = torch.from_numpy(rng.normal(scale=0.3, size=p)).cuda()
    G, omega = logistic_omega_step(X, beta, 2.0, random_state=(1, 2))
    Xn, w = X.cpu().numpy(), omega.cpu().numpy()
    G_ref = Xn.T @ (w[:, None] * Xn)
    Gn = G.cpu().numpy()
    assert np.allclose(Gn, G_ref, rtol=1e-11, atol=1e-11 * np.abs(G_ref).max())
    assert np.array_equal(Gn, Gn.T)


def test_omega_step_equals_random_polyagamma(torch):
    """Same (seed, offset): omega is bit-identical to random_polyagamma(n_trials, psi), scalar n_trials
    included, and the Gram matrix is reproducible run to run (fixed reduction order)."""
    import polyagamma_b200 as pg
    from polyagamma_b200.gibbs import logistic_omega_step

    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    n, p = 1_000_003, 24
    X = torch.randn(n, p, device="cuda", dtype=torch.float64, generator=g)
    beta = torch.randn(p, device="cuda", dtype=torch.float64, generator=g) * 0.5
    G1, om1, psi = logistic_omega_step(X, beta, 1.0, random_state=(9, 9), return_psi=True)
    om2 = pg.random_polyagamma(1.0, psi, random_state=(9, 9), disable_checks=True)
    assert torch.equal(om1, om2)
    G2, _ = logistic_omega_step(X, beta, 1.0, random_state=(9, 9), return_omega=False)
    assert torch.equal(G1, G2)
    # a strided view of a wider matrix works through ldx
    wide = torch.randn(n, p + 5, device="cuda", dtype=torch.float64, generator=g)
    G3, om3 = logistic_omega_step(wide[:, :p], beta, 1.0, random_state=(9, 9))
    G4, om4 = logistic_omega_step(wide[:, :p].contiguous(), beta, 1.0, random_state=(9, 9))
    assert torch.equal(om3, om4) and torch.equal(G3, G4)
    with pytest.raises(ValueError):
        logistic_omega_step(torch.zeros(4, 0, device="cuda", dtype=torch.float64), torch.zeros(0))


def test_wide_design_matrix_falls_back_to_library_gemms(torch):
    """p > 64: psi and the Gram matrix come from cuBLAS, the draw is the same call with the same streams."""
    import polyagamma_b200 as pg
    from polyagamma_b200.gibbs import logistic_omega_step

    g = torch.Generator(device="cuda")
    g.manual_seed(8)
    n, p = 50_003, 100
    X = torch.randn(n, p, device="cuda", dtype=torch.float64, generator=g)
    beta = torch.randn(p, device="cuda", dtype=torch.float64, generator=g) * 0.2
    G, om, psi = logistic_omega_step(X, beta, 2.0, random_state=(4, 4), return_psi=True)
    assert torch.equal(om, pg.random_polyagamma(2.0, psi, random_state=(4, 4), disable_checks=True))
    want = X.T @ (om[:, None] * X)
    assert float((G - want).abs().max() / want.abs().max()) < 1e-13
    assert torch.equal(G, G.T) and G.shape == (p, p)
    # the first 64 columns through the fused kernels give the same omegas for the same psi
    b64 = torch.zeros(p, device="cuda", dtype=torch.float64)
    b64[:64] = beta[:64]
    _, om_wide = logistic_omega_step(X, b64, 2.0, random_state=(4, 5))
    _, om_fused = logistic_omega_step(X[:, :64], beta[:64], 2.0, random_state=(4, 5))
    assert float((om_wide - om_fused).abs().max()) < 1e-9  # (psi differs in the last bits between the two GEMVs)


@pytest.mark.parametrize("world", [2, 3])
def test_row_shards_reproduce_the_single_device_step(torch, world):
    """Row-sharded sweep (what sharded_logistic_omega_step does per rank, here one shard after the other on
    one device, and on every visible device in turn): omegas concatenate to the single-call draw bit for bit,
    the partial Gram matrices add up to the single-call matrix."""
    from polyagamma_b200.gibbs import logistic_omega_step, sharded_logistic_omega_step
    from polyagamma_b200.sharding import shard_range

    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    n, p = 300_007, 20
    X = torch.randn(n, p, device="cuda", dtype=torch.float64, generator=g)
    beta = torch.randn(p, device="cuda", dtype=torch.float64, generator=g) * 0.6
    nt = torch.randint(1, 4, (n,), device="cuda", generator=g).double()
    G, om = logistic_omega_step(X, beta, nt, random_state=(3, 1))
    ndev = torch.cuda.device_count()
    total = torch.zeros_like(G)
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        dev = torch.device("cuda", r % ndev)
        # (no process group here: the all-reduce inside is the identity, the sum is taken below)
        Gr, omr = sharded_logistic_omega_step(X[lo:hi].to(dev), beta.to(dev), nt[lo:hi].to(dev), row_offset=lo,
                                              random_state=(3, 1))
        assert torch.equal(omr.to("cuda:0"), om[lo:hi])
        total += Gr.to("cuda:0")
    assert float((total - G).abs().max() / G.abs().max()) < 1e-13
    with pytest.raises(ValueError):
        sharded_logistic_omega_step(X[:10], beta, 1.0, row_offset=0, random_state=None)
