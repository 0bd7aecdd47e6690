"""CPU: the conjugate Algorithm 2 oracle (oracle/np_oracle_alg2.inc).  The reference has no executable conjugate path
(np_neal_algorithm2.cpp is dead code, include/statistics/conjugate/*.h are empty): parity against the reference is UNPINNED and
the oracle is pinned against scipy.stats.multivariate_t with the textbook NIW posterior instead."""
import numpy as np
import pytest
from scipy import stats as sps

from noparama_b200 import synthetic as syn


def niw_posterior(pr, Xm):
    n = len(Xm)
    D = len(pr["mu0"])
    if n == 0:
        return pr["mu0"], pr["kappa"], pr["nu"], pr["Lambda"]
    xbar = Xm.mean(0)
    S = (Xm - xbar).T @ (Xm - xbar)
    kn, nn = pr["kappa"] + n, pr["nu"] + n
    mun = (pr["kappa"] * pr["mu0"] + n * xbar) / kn
    Ln = pr["Lambda"] + S + pr["kappa"] * n / kn * np.outer(xbar - pr["mu0"], xbar - pr["mu0"])
    return mun, kn, nn, Ln


@pytest.mark.parametrize("D", [2, 16, 64])
def test_logpred_matches_scipy_multivariate_t(oracle, D):
    rng = np.random.default_rng(D)
    pr = dict(mu0=rng.standard_normal(D), kappa=0.3, nu=D + 2.0, Lambda=np.eye(D) + 0.1 * np.ones((D, D)), alpha=1.0)
    P = oracle.make_prior(**pr)
    for n in (0, 1, 5, 300):
        Xm = rng.standard_normal((n, D)) * 1.5 + 2.0
        for _ in range(3):
            x = rng.standard_normal(D) * 2.0 + 1.0
            mun, kn, nn, Ln = niw_posterior(pr, Xm)
            df = nn - D + 1
            want = sps.multivariate_t(loc=mun, shape=Ln * (kn + 1) / (kn * df), df=df).logpdf(x)
            got = oracle.niw_logpred(P, Xm, x)
            assert abs(got - want) < 1e-9 * max(1.0, abs(want)), (D, n, got, want)
            inc = oracle.niw_logpred(P, Xm, x, incremental=True)
            assert abs(inc - want) < 1e-8 * max(1.0, abs(want))
    # rank-1 removal: insert 40 rows, remove the first 15 again == the cluster of the last 25
    Xm = rng.standard_normal((40, D)) + 1.0
    x = rng.standard_normal(D)
    a = oracle.niw_logpred(P, Xm, x, incremental=True, remove_first=15)
    b = oracle.niw_logpred(P, Xm[15:], x)
    assert abs(a - b) < 1e-8 * max(1.0, abs(b))


def test_alg2_run_recovers_two_gaussians(oracle):
    X, y = syn.config(1)
    pr = dict(mu0=X.mean(0), kappa=0.01, nu=4.0, Lambda=np.eye(2), alpha=1.0)
    P = oracle.make_prior(**pr)
    z, Kt, moved, births = oracle.alg2_run(P, X, 60, 4, 7)
    pur, ri, ari = oracle.metrics(y, z)
    assert pur > 0.97 and ari > 0.9 and 2 <= Kt[-1] <= 4, (pur, ari, Kt[-5:])
    assert moved > 0 and births >= 0
