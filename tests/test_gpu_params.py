"""GPU: the parameter update "done right" (npb_params.cu).

Not in the reference (its UpdateClusters result is sliced away), so there is no reference arithmetic to match:
the sufficient statistics and the posterior-mean mode are checked against numpy in double precision, the posterior draw
through its moments, and the full loop (sweep + update) by recovering a 16-D mixture from a poor start.
"""
import numpy as np
import pytest

from noparama_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def niw_posterior(Xk, pr):
    n, D = Xk.shape
    xbar = Xk.mean(0)
    S = (Xk - xbar).T @ (Xk - xbar)
    kn, nun = pr["kappa"] + n, pr["nu"] + n
    mun = (pr["kappa"] * pr["mu0"] + n * xbar) / kn
    dm = (xbar - pr["mu0"])[:, None]
    Ln = pr["Lambda"] + S + pr["kappa"] * n / kn * (dm @ dm.T)
    return kn, nun, mun, Ln


@pytest.mark.parametrize("D", [2, 3, 16])
def test_posterior_mean_mode_matches_numpy(npb, ctx, D):
    X, y = syn.gmm(600, D, 4, 300 + D)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(D)).bind(ctx)
    ch = npb.Chains(ctx, ds, 6, Kmax=32, K0=8, seed=5)
    means = np.stack([X[y == k].mean(0) for k in range(4)])
    ch.init_from_params(means + 0.5, np.tile(np.eye(D), (4, 1, 1)))
    ch.sweep(npb.ALG8, 2)
    pr = dict(mu0=X.mean(0), kappa=0.01, nu=D + 2.0, Lambda=np.eye(D))
    ch.update_params(npb.UPDATE_POSTERIOR_MEAN, pr)
    z = ch.assignments()
    for c in (0, 5):
        slots, counts, mu, Sigma = ch.params(c)
        assert counts.sum() == len(X)
        for j, s in enumerate(slots):
            Xk = X[z[c] == s]
            assert len(Xk) == counts[j]
            kn, nun, mun, Ln = niw_posterior(Xk, pr)
            want = Ln / (nun - D - 1)
            assert np.allclose(mu[j], mun, rtol=1e-5, atol=1e-5), (c, s)
            assert np.allclose(Sigma[j], want, rtol=2e-4, atol=2e-5 * np.abs(want).max()), (c, s, np.abs(Sigma[j] - want).max())
    # the refreshed parameters are what the next sweep uses: densities evaluate and the state stays consistent
    st = ch.sweep(npb.ALG8, 1)
    assert st.overflow_chains == 0
    ch.close()
    ds.close()


def test_posterior_draw_moments(npb, ctx):
    D, C = 3, 768
    rng = np.random.default_rng(3)
    A = rng.standard_normal((D, D))
    X = rng.standard_normal((80, D)) @ A.T + np.array([2.0, -1.0, 0.5])
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(D)).bind(ctx)
    ch = npb.Chains(ctx, ds, C, Kmax=32, K0=4, seed=9)
    ch.init_from_params(X.mean(0)[None, :], np.eye(D)[None])   # one cluster holding every item, in every chain
    pr = dict(mu0=np.zeros(D), kappa=0.5, nu=D + 3.0, Lambda=2.0 * np.eye(D))
    ch.update_params(npb.UPDATE_POSTERIOR_DRAW, pr)
    kn, nun, mun, Ln = niw_posterior(X, pr)
    mus, Sig = [], []
    for c in range(C):
        slots, counts, mu, Sigma = ch.params(c)
        assert len(slots) == 1 and counts[0] == len(X)
        mus.append(mu[0])
        Sig.append(Sigma[0])
    mus, Sig = np.array(mus), np.array(Sig)
    ESig = Ln / (nun - D - 1)                       # mean of IW(nu_n, Lambda_n)
    assert np.allclose(Sig.mean(0), ESig, rtol=0.05, atol=0.02 * np.abs(ESig).max())
    assert np.allclose(mus.mean(0), mun, atol=4 * np.sqrt(np.diag(ESig) / kn / C) + 1e-3)
    assert np.allclose(np.cov(mus.T), ESig / kn, rtol=0.25, atol=0.1 * np.abs(ESig / kn).max())
    # variance of a diagonal element of IW: 2 L_ii^2 / ((nu-D-1)^2 (nu-D-3))
    v00 = 2 * Ln[0, 0] ** 2 / ((nun - D - 1) ** 2 * (nun - D - 3))
    assert abs(Sig[:, 0, 0].var() - v00) < 0.3 * v00
    # draws differ between chains and between calls
    ch.update_params(npb.UPDATE_POSTERIOR_DRAW, pr)
    assert not np.allclose(ch.params(0)[2], mus[0])
    ch.close()
    ds.close()


def test_sweep_plus_update_recovers_16d_mixture(npb, ctx):
    """The `fixed` regime of SURVEY 8d: chains start from K_true clusters with poor parameters (means off by one sigma
    per coordinate, covariance 4 I) and alternate an Alg. 8 sweep with a posterior draw of the parameters."""
    X, y = syn.gmm(4000, 16, 8, 77)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(16)).bind(ctx)
    ch = npb.Chains(ctx, ds, 16, Kmax=32, seed=2)
    rng = np.random.default_rng(0)
    means = np.stack([X[y == k].mean(0) for k in range(8)])
    ch.init_from_params(means + rng.standard_normal(means.shape), np.tile(4.0 * np.eye(16), (8, 1, 1)))
    pr = dict(mu0=X.mean(0), kappa=0.01, nu=18.0, Lambda=np.eye(16))
    for _ in range(6):
        st = ch.sweep(npb.ALG8, 1)
        assert st.overflow_chains == 0
        ch.update_params(npb.UPDATE_POSTERIOR_DRAW, pr)
    m = ch.metrics(y)
    assert m["purity"].mean() > 0.99 and np.all(m["K"] == 8)
    slots, counts, mu, Sigma = ch.params(3)
    z = ch.assignments(3, 1)[0]
    for j, s in enumerate(slots):
        k = np.bincount(y[z == s]).argmax()
        assert np.abs(mu[j] - means[k]).max() < 0.3
        assert np.abs(Sigma[j] - np.eye(16)).max() < 0.35
    ch.close()
    ds.close()


def test_update_params_argument_checks(npb, ctx):
    X, _ = syn.config(1)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(2)).bind(ctx)
    ch = npb.Chains(ctx, ds, 4, Kmax=32, K0=8, seed=1)
    with pytest.raises(npb.NpbError):
        ch.update_params(7)
    with pytest.raises(npb.NpbError):
        ch.update_params(1, dict(mu0=np.zeros(2), kappa=-1.0, nu=4.0, Lambda=np.eye(2)))
    ch.update_params(npb.UPDATE_POSTERIOR_DRAW)   # the bound (reference) prior
    assert ch.sweep(npb.ALG8, 1).overflow_chains == 0
    ch.close()
    ds.close()
