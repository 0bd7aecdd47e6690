"""GPU (pytest -m gpu): the LAW of the throughput races.  The sweep kernels replace random_weighted_pick (cumulative sum +
lower_bound, include/helper/dim1algebra.hpp:2078-2104) by an exponential race with cheap noise (a one-instruction LCG at D <= 3,
a hash-seeded LCG in the FP32-pipe kernels, a counter hash capped at 20 and floored at -6 on the tensor paths, speculative and
sequential passes).  Here 2^17 chains start from the SAME state (npb_chains_broadcast_state) and make the same first
reassignment of a sweep independently; the frequencies of the chosen cluster are compared with the exact categorical
p(x | theta_k) n_k / sum (fp64, oracle density) by a chi-square pooled over 16 first steps (2.1e6 draws per kernel: a relative
bias of 1e-3 in a pick probability is a 3-sigma event), including a cluster that should win about once in a thousand draws.
Steps won by an auxiliary draw (a birth) are set aside: the ratios among the existing clusters do not depend on them."""
import os

import numpy as np
import pytest
from scipy import stats as sps

import noparama_b200
from noparama_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
KEYS = ("NPB_D16_PATH", "NPB_D16_BLOCK", "NPB_D64_SPEC", "NPB_D16_AUX", "NPB_TILE_KERNEL", "NPB_D64_BLOCK")


@pytest.fixture
def env():
    saved = {k: os.environ.get(k) for k in KEYS}
    for k in KEYS:
        os.environ.pop(k, None)
    yield os.environ
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def problem(D, seed):
    """96 items of four overlapping clusters of different sizes (a heavy one, two middling ones 2 apart, a light one 3 away
    whose weight for a typical item of its neighbour is ~1e-3), full covariances at D >= 4."""
    rng = np.random.default_rng(seed)
    sizes = np.array([48, 28, 16, 4])
    means = np.zeros((4, D))
    means[:, 0] = [0.0, 2.0, 4.0, 7.0]
    if D > 1:
        means[:, 1] = [0.0, 1.0, -1.0, 0.5]
    y = np.repeat(np.arange(4), sizes).astype(np.int32)
    Sigma = np.empty((4, D, D))
    for k in range(4):
        B = rng.standard_normal((D, D)) * (0.25 if D >= 4 else 0.1)
        Sigma[k] = np.eye(D) * (0.8 + 0.2 * k) + B @ B.T
    X = means[y] + np.stack([np.linalg.cholesky(Sigma[k]) @ rng.standard_normal(D) for k in y])
    perm = rng.permutation(len(y))
    return X[perm] + 6.0, y[perm], means + 6.0, Sigma


def law_check(npb, ctx, oracle, D, kmax, m_aux, chains, rounds, seed, sampler=None):
    X, y, means, Sigma = problem(D, 100 + D)
    N = len(y)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(D)).bind(ctx)
    ch = npb.Chains(ctx, ds, chains, Kmax=kmax, m_aux=m_aux, K0=4, seed=seed)
    logp = oracle.mvn_logpdf_batch(means, Sigma, X)  # [N, 4]
    chi2, dof, n_used, births, low_obs, low_exp = 0.0, 0, 0, 0, 0.0, 0.0
    for r in range(rounds):
        ch.set_state(0, y, [0, 1, 2, 3], means, Sigma)
        ch.broadcast_state(0)
        i0 = int(noparama_b200.scan_order(seed, r, N)[0])  # the first item of sweep r (the chains' scan order is keyed by (seed, sweep))
        ch.sweep(sampler if sampler is not None else npb.ALG8, 1)
        z = ch.assignments()[:, i0]
        n = np.bincount(y, minlength=4).astype(np.float64)
        n[y[i0]] -= 1.0
        w = np.exp(logp[i0] - logp[i0].max()) * n
        p = w / w.sum()
        existing = z < 4
        births += int((~existing).sum())
        obs = np.bincount(z[existing], minlength=4)[:4].astype(np.float64)
        exp = p * existing.sum()
        keep = exp >= 5.0
        # (cells with a tiny expectation are pooled into the largest cell's complement by dropping them from both sides)
        scale = obs[keep].sum() / exp[keep].sum()
        chi2 += float((((obs[keep] - exp[keep] * scale) ** 2) / (exp[keep] * scale)).sum())
        dof += int(keep.sum()) - 1
        n_used += int(existing.sum())
        low = np.argmin(np.where(keep, p, 1.0))
        if p[low] < 0.02:
            low_obs += obs[low]
            low_exp += exp[low]
    pval = float(sps.chi2.sf(chi2, dof))
    ch.close()
    ds.close()
    return dict(chi2=chi2, dof=dof, p=pval, draws=n_used, births=births, low_obs=low_obs, low_exp=low_exp)


def assert_law(res, name):
    print("%-28s chi2 %.1f / %d dof, p = %.3g, %d draws, %d births; low-weight cells: observed %.0f, expected %.1f" % (
        name, res["chi2"], res["dof"], res["p"], res["draws"], res["births"], res["low_obs"], res["low_exp"]))
    assert res["p"] > 1e-3, (name, res)
    if res["low_exp"] > 200:
        assert abs(res["low_obs"] - res["low_exp"]) < 4.0 * np.sqrt(res["low_exp"]) + 0.002 * res["low_exp"], (name, res)


@pytest.mark.parametrize("D,kmax", [(2, 32), (3, 32), (2, 256)])
def test_law_register_kernel(npb, ctx, oracle, env, D, kmax):
    assert_law(law_check(npb, ctx, oracle, D, kmax, 3, 1 << 17, 16, 4242 + D), "k_alg8_sweep_reg D=%d Kmax=%d" % (D, kmax))


@pytest.mark.parametrize("D,kmax", [(4, 32), (8, 32), (4, 64)])
def test_law_fp32_tile_kernels(npb, ctx, oracle, env, D, kmax):
    assert_law(law_check(npb, ctx, oracle, D, kmax, 3, 1 << 17, 16, 777 + D + kmax), "tile kernels D=%d Kmax=%d" % (D, kmax))


@pytest.mark.parametrize("path,spec,aux", [("tc", "1", None), ("tc", "0", None), ("tc2", "1", None), ("tc2", "0", None),
                                           ("tc2", "1", "lazy"), ("fp32", "1", None)])
def test_law_d16_paths(npb, ctx, oracle, env, path, spec, aux):
    env["NPB_D16_PATH"], env["NPB_D64_SPEC"] = path, spec
    if aux:
        env["NPB_D16_AUX"] = aux
    assert_law(law_check(npb, ctx, oracle, 16, 32, 3, 1 << 17, 16, 1616), "D=16 path=%s spec=%s aux=%s" % (path, spec, aux))


@pytest.mark.parametrize("spec", ["1", "0"])
def test_law_d64_fused_race(npb, ctx, oracle, env, spec):
    env["NPB_D64_SPEC"] = spec
    assert_law(law_check(npb, ctx, oracle, 64, 32, 1, 1 << 14, 32, 6464, sampler=npb.ALG2), "D=64 k_density_tc spec=%s" % spec)


@pytest.mark.parametrize("D", [2, 16])
def test_base_measure_draws_follow_the_reference_law(npb, ctx, oracle, D):
    """a5: the device's draws from the base measure (npb_draw_theta: the K0 initial clusters of every chain) against the
    oracle's restatement of dirichlet_process::sample_base (dirichlet.h:91-93 -> normalinvwishart.h:44-64 -> invwishart.h:34-46,
    bug-compatible Q2: Sigma = v^2 L^T L with one scalar v ~ N(D, nu)): two-sample KS on the scale v^2 = Sigma_00 / Lambda_00, on
    the standardised mean coordinates, and the exact structure Sigma = s Lambda of the degenerate inverse-Wishart draw."""
    from scipy import stats as sps
    pr = syn.reference_prior(D)
    X, _ = syn.gmm(400, D, 2, 5)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**pr).bind(ctx)
    ch = npb.Chains(ctx, ds, 256, Kmax=32, K0=20, seed=31)
    mus, sig = [], []
    for c in range(256):
        _, _, mu, Sigma = ch.params(c)  # the clusters that kept a member: all 20 at N = 400 with overwhelming probability
        mus.append(mu)
        sig.append(Sigma)
    mu, Sigma = np.concatenate(mus), np.concatenate(sig)
    assert len(mu) > 4000
    s = Sigma[:, 0, 0] / pr["Lambda"][0, 0]
    assert np.allclose(Sigma, s[:, None, None] * pr["Lambda"][None], rtol=2e-4, atol=1e-9)
    omu, oSigma = oracle.sample_base(oracle.make_prior(**pr), 77, 6000)
    os_ = oSigma[:, 0, 0] / pr["Lambda"][0, 0]
    p = sps.ks_2samp(s, os_).pvalue
    assert p > 1e-3, ("scale", p, s.mean(), os_.mean())
    # mu | Sigma ~ N(mu0, Sigma / kappa): standardised coordinates are N(0, 1) on both sides
    z = (mu - pr["mu0"]) / np.sqrt(Sigma[:, 0, 0] / pr["kappa"])[:, None]
    oz = (omu - pr["mu0"]) / np.sqrt(oSigma[:, 0, 0] / pr["kappa"])[:, None]
    for d in (0, D - 1):
        assert sps.ks_2samp(z[:, d], oz[:, d]).pvalue > 1e-3, d
    assert sps.kstest(z.ravel(), "norm").pvalue > 1e-3
    ch.close()
    ds.close()
