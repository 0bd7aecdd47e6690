"""CPU: the scalar-noise likelihood families of `-c regression` / `-c angular` (scalarnoise_multivariatenormal.cpp,
normalinvgamma.h, gamma.h) in the oracle, pinned to the reference's OWN sources (oracle/_ref np_ref_run ... FAMILY): same
partition after every sweep, same labels at the end."""
import math

import numpy as np
import pytest

from noparama_b200 import synthetic as syn
from oracle import refrun
from test_ref_pin import same_partition


def data(family, N, seed):
    return syn.regression_lines(N, 2, seed) if family == "regression" else syn.angular_lines(N, 2, seed)


def test_density_formulas(oracle):
    # scalarnoise_multivariatenormal.cpp:86-101 / :190-207 (regression) and :120-140 / :225-249 (angular), by hand
    mu, sigma = np.array([0.5, -2.0]), 0.3
    x = np.array([1.0, 1.5, -2.0])
    r = x[2] - (mu[0] * x[0] + mu[1] * x[1])
    want = -0.5 * r * r / sigma ** 2 - 0.5 * math.log(2 * math.pi * sigma ** 2)
    assert abs(oracle.scalarnoise_logpdf(oracle.REGRESSION, mu, sigma, x) - want) < 1e-14
    assert abs(oracle.scalarnoise_logpdf(oracle.REGRESSION, mu, sigma, x, log=False) - math.exp(want)) < 1e-15
    # angular: prepare() makes (d, theta) canonical -- through the C abs(int): both are truncated to integers first (Q12)
    mu = np.array([-3.9, -(1.7 + 2 * math.pi)])   # -> d = 3, theta = fmod(7, 2 pi)
    p = np.array([1.2, 2.5])
    th = math.fmod(7.0, 2 * math.pi)
    q1 = -math.sin(th) * p[0] + math.cos(th) * p[1]
    want = -0.5 * (3.0 - q1) ** 2 / sigma ** 2 - 0.5 * math.log(2 * math.pi * sigma ** 2)
    assert abs(oracle.scalarnoise_logpdf(oracle.ANGULAR, mu, sigma, p) - want) < 1e-12


def test_nig_draws_have_the_inverse_gamma_law(oracle):
    from scipy import stats as sps
    pr = syn.reference_nig_prior()
    mu, sg = oracle.sample_base_nig(pr, 11, 4000)
    # 1 / sigma^2 ~ Gamma(shape alpha, scale beta) (gamma.h:41); mu | sigma ~ N(0, sigma^2 Lambda^-1)
    assert sps.kstest(1.0 / sg ** 2, "gamma", args=(pr["nig_alpha"], 0, pr["nig_beta"])).pvalue > 1e-3
    zs = mu / (sg[:, None] * 10.0)
    assert sps.kstest(zs.ravel(), "norm").pvalue > 1e-3


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("family,seed", [("regression", 1), ("regression", 2), ("angular", 1), ("angular", 2)])
def test_alg8_trajectory_matches_reference(oracle, family, seed):
    X, _ = data(family, 150, 40 + seed)
    pr = syn.reference_nig_prior()
    T = 60
    ref = refrun.run(X, pr, 8, T=T, seed_main=seed, seed_shuffle=700 + seed, record=True, family=family)
    fam = oracle.REGRESSION if family == "regression" else oracle.ANGULAR
    run = oracle.ScalarNoiseRun(fam, pr, X, oracle.ALG8, T=T, seed_main=seed, seed_shuffle=700 + seed, flags=oracle.FAITHFUL)
    assert ref["calls"] == T * len(X) == run.stats().updates
    assert np.array_equal(run.assignments(0), ref["z_final"])
    assert same_partition(run.assignments(1), ref["z_maxlik"])
    assert run.stats().K_final == ref["K_final"]


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("family,alg", [("regression", 2), ("angular", 3)])
def test_split_merge_trajectory_matches_reference(oracle, family, alg):
    X, _ = data(family, 120, 77)
    pr = syn.reference_nig_prior()
    T = 15
    ref = refrun.run(X, pr, alg, T=T, seed_main=5, seed_shuffle=55, record=True, family=family)
    fam = oracle.REGRESSION if family == "regression" else oracle.ANGULAR
    run = oracle.ScalarNoiseRun(fam, pr, X, {2: oracle.JAIN_NEAL, 3: oracle.TRIADIC}[alg], T=T, seed_main=5, seed_shuffle=55,
                                flags=oracle.FAITHFUL)
    assert ref["calls"] == run.stats().updates
    assert np.array_equal(run.assignments(0), ref["z_final"])
    assert run.stats().K_final == ref["K_final"]
