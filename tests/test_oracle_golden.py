"""CPU: the oracle against every golden vector the reference's own tests hold for this path
and against independent implementations (numpy / scipy / sklearn)."""
import numpy as np
import pytest
from scipy import stats as sps

from noparama_b200 import synthetic as syn

KAT_MU = [1.0, 1.0]
KAT_SIGMA = [[2.0, 0.0], [1.0, 2.0]]  # non-symmetric on purpose, test/test_mvn_likelihood.cpp:21-22
KAT_X = [1.0, 2.0]


def test_kat_probability(oracle):
    # test/test_mvn_likelihood.cpp:31-33: assert(prob - 0.061974 < 1e-5)
    p = oracle.mvn_pdf(KAT_MU, KAT_SIGMA, KAT_X)
    assert p - 0.061974 < 0.00001
    assert abs(p - 0.0619749972) < 1e-9  # full-precision value
    assert abs(oracle.mvn_logpdf(KAT_MU, KAT_SIGMA, KAT_X) - (-2.7810242470)) < 1e-9


def test_kat_probability_dataset(oracle):
    # test/test_mvn_likelihood.cpp:35-44: two copies of the item, assert(prob_dataset - 0.0038409 < 1e-5)
    p = oracle.mvn_pdf_dataset(KAT_MU, KAT_SIGMA, [KAT_X, KAT_X])
    assert p - 0.0038409 < 0.00001
    assert abs(p - 0.0038409003) < 1e-9
    assert abs(oracle.mvn_logpdf_dataset(KAT_MU, KAT_SIGMA, [KAT_X, KAT_X]) - 2 * (-2.7810242470)) < 1e-9


def test_weighted_pick_frequencies(oracle):
    # test/test_weighted_vector.cpp:10-36: weights 0,10,..,70; 280000 draws; frequencies ~ weights, bin 0 never hit
    w = np.arange(8) * 10.0
    f = oracle.weighted_pick_freq(w, 280000, 12345)
    assert f[0] == 0 and f.sum() == 280000
    expect = 280000 * w / w.sum()
    assert np.all(np.abs(f[1:] - expect[1:]) < 5 * np.sqrt(expect[1:]))


def test_weighted_pick_semantics(oracle):
    # dim1algebra.hpp:2078-2104: first index with cumsum >= u*total; all-zero weights give index 0 (Q6)
    assert oracle.weighted_pick_u([1, 1, 2], 0.0) == 0
    assert oracle.weighted_pick_u([1, 1, 2], 0.25) == 0
    assert oracle.weighted_pick_u([1, 1, 2], 0.2500001) == 1
    assert oracle.weighted_pick_u([1, 1, 2], 0.5) == 1
    assert oracle.weighted_pick_u([1, 1, 2], 0.9999) == 2
    assert oracle.weighted_pick_u([0, 0, 0], 0.7) == 0
    assert oracle.weighted_pick_u([0, 3, 0], 0.7) == 1
    assert oracle.weighted_pick_u([-5.0, 2.0], 0.5) == 2  # Q8: decreasing cumsum -> end


def test_membertrix_invariant(oracle):
    # test/test_membertrix.cpp:75-91
    for seed in range(50):
        for dense in (False, True):
            assert oracle.membertrix_selftest(seed, dense) == 0


@pytest.mark.parametrize("n", [1, 2, 3, 5, 16, 33])
def test_lu_against_numpy(oracle, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)) + n * np.eye(n)
    assert np.allclose(oracle.lu_inverse(A), np.linalg.inv(A), rtol=1e-10, atol=1e-12)
    assert np.isclose(oracle.lu_determinant(A), np.linalg.det(A), rtol=1e-10)


@pytest.mark.parametrize("D", [2, 3, 16, 64])
def test_logpdf_against_scipy(oracle, D):
    rng = np.random.default_rng(D)
    B = rng.standard_normal((D, D))
    Sigma = B @ B.T / D + np.eye(D)
    mu = rng.standard_normal(D)
    X = rng.standard_normal((50, D)) * 1.5
    want = sps.multivariate_normal(mu, Sigma).logpdf(X)
    got = oracle.mvn_logpdf_batch(mu[None], Sigma[None], X)[:, 0]
    assert np.allclose(got, want, rtol=1e-10, atol=1e-9)
    assert np.isclose(oracle.mvn_logpdf_dataset(mu, Sigma, X), want.sum(), rtol=1e-10)


def test_sample_base_structure(oracle):
    # Q2: Sigma = v^2 L^T L (= 0.01 v^2 I here), v ~ N(D, nu^2); mu ~ N(mu0, Sigma/kappa)
    pr = syn.reference_prior(2)
    mu, Sigma = oracle.sample_base(oracle.make_prior(**pr), 99, 20000)
    assert np.allclose(Sigma[:, 0, 1], 0) and np.allclose(Sigma[:, 0, 0], Sigma[:, 1, 1])
    v = np.sqrt(Sigma[:, 0, 0] / 0.01)  # |v|
    # |v| with v ~ N(2, 4^2): compare a few moments of v^2
    assert abs(np.mean(v ** 2) - (4.0 + 16.0)) < 0.5
    zs = (mu - 6.0) / (0.1 * v[:, None] * np.sqrt(500.0))
    assert abs(zs.mean()) < 0.03 and abs(zs.std() - 1.0) < 0.03
    assert sps.kstest(zs[:, 0], "norm").pvalue > 1e-3


def test_sample_base_nondiagonal_lambda(oracle):
    # general Lambda: Sigma must equal v^2 L^T L with L = chol(Lambda) (invwishart.h:40-43)
    Lam = np.array([[2.0, 0.6, 0.1], [0.6, 1.5, 0.3], [0.1, 0.3, 1.0]])
    p = oracle.make_prior(np.zeros(3), 0.5, 5.0, Lam, 1.0)
    mu, Sigma = oracle.sample_base(p, 5, 200)
    L = np.linalg.cholesky(Lam)
    A = L.T @ L
    ratio = Sigma / A
    assert np.allclose(ratio, ratio[:, :1, :1], rtol=1e-9)


def test_metrics_against_sklearn(oracle):
    from sklearn.metrics import adjusted_rand_score, rand_score
    rng = np.random.default_rng(3)
    for _ in range(5):
        a = rng.integers(0, 4, 300)
        b = (a + (rng.random(300) < 0.3) * rng.integers(0, 6, 300)) % 7
        pur, ri, ari = oracle.metrics(a, b)
        assert np.isclose(ri, rand_score(a, b))
        assert np.isclose(ari, adjusted_rand_score(a, b))
        cont = np.zeros((4, 7))
        np.add.at(cont, (a, b), 1)
        assert np.isclose(pur, cont.max(0).sum() / 300)


def test_metrics_no_int_overflow(oracle):
    # Q12: the reference's int arithmetic overflows for N > 46340; the oracle uses int64
    n = 100000
    a = np.arange(n) % 10
    pur, ri, ari = oracle.metrics(a, a)
    assert pur == 1.0 and np.isclose(ri, 1.0) and np.isclose(ari, 1.0)


def test_alg8_run_config1(oracle):
    """Config 1 end to end: the qualitative expectation of README.rst:55 and the survey's emulation band."""
    X, y = syn.config(1)
    p = oracle.make_prior(**syn.reference_prior(2))
    r = oracle.Run(p, X, T=300, seed_main=11, seed_shuffle=12, flags=oracle.FAITHFUL)
    s = r.stats()
    assert s.updates == 300 * 200
    assert s.candidates == s.updates * 3 + round(s.mean_K * s.updates)
    pur, ri, ari = oracle.metrics(y, r.assignments(0))
    assert pur > 0.95 and ri > 0.5 and ri < pur
    assert 3 <= s.K_final <= 20
    z = r.assignments(0)
    mu, Sigma, counts = r.params()
    assert counts.sum() == 200 and len(counts) == s.K_final == len(np.unique(z))


def test_alg8_cost_profile_flags_do_not_change_results(oracle):
    """Dense matrix / per-call LU only change the cost profile, never the trajectory."""
    X, y = syn.config(1)
    p = oracle.make_prior(**syn.reference_prior(2))
    base = oracle.UPDATE_CLUSTERS | oracle.MAX_LIKELIHOOD
    a = oracle.Run(p, X, T=40, seed_main=3, seed_shuffle=4, flags=base)
    b = oracle.Run(p, X, T=40, seed_main=3, seed_shuffle=4, flags=base | oracle.DENSE_MATRIX | oracle.PER_CALL_LU)
    assert np.array_equal(a.assignments(0), b.assignments(0))
    assert np.array_equal(a.assignments(1), b.assignments(1))
    assert a.stats().max_loglik == b.stats().max_loglik


def test_alg8_trace_is_self_consistent(oracle):
    X, y = syn.config(1)
    p = oracle.make_prior(**syn.reference_prior(2))
    r = oracle.Run(p, X, T=5, seed_main=3, seed_shuffle=4, flags=oracle.RECORD_TRACE)
    t = r.trace()
    z0, slots, mu, Sigma = r.init_state()
    assert len(t["item"]) == 5 * 200 and t["order_off"][-1] == len(t["order"])
    # replay the trace with the oracle's own density: picks must be reproduced
    z = z0.copy()
    theta = {int(s): (mu[i], Sigma[i]) for i, s in enumerate(slots)}
    for s in range(len(t["item"])):
        i = t["item"][s]
        z[i] = -1
        order = t["order"][t["order_off"][s]:t["order_off"][s + 1]]
        w = [oracle.mvn_pdf(*theta[int(k)], X[i]) * np.sum(z == k) for k in order]
        w += [oracle.mvn_pdf(t["aux_mu"][s, m], t["aux_Sigma"][s, m], X[i]) / 3.0 for m in range(3)]
        j = oracle.weighted_pick_u(w, t["u"][s])
        assert j == t["picked"][s]
        if j >= len(order):
            theta[int(t["new_slot"][s])] = (t["aux_mu"][s, j - len(order)], t["aux_Sigma"][s, j - len(order)])
            z[i] = t["new_slot"][s]
        else:
            z[i] = order[j]
        if (s + 1) % 200 == 0:
            assert np.array_equal(z, t["z_after"][(s + 1) // 200 - 1])
