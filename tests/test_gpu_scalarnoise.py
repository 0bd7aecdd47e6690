"""GPU (pytest -m gpu): the scalar-noise likelihood families of `-c regression` / `-c angular` (npb_scalarnoise.cu) against the
oracle, which tests/test_oracle_scalarnoise.py pins to the reference's own compiled sources: densities, the base measure's law,
state invariants, and the chains' distributions (K, purity, ARI, birth and move rates) against oracle seeds."""
import numpy as np
import pytest

from noparama_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
FAM = {"regression": 1, "angular": 2}


def data(family, N, seed, K=2):
    return syn.regression_lines(N, K, seed) if family == "regression" else syn.angular_lines(N, K, seed)


@pytest.mark.parametrize("family", ["regression", "angular"])
def test_logdensity_matches_oracle(npb, ctx, oracle, family):
    rng = np.random.default_rng(3)
    X, _ = data(family, 500, 9)
    ds = npb.Dataset(ctx, X)
    K = 40
    mu = rng.normal(0.0, 4.0, (K, 2))
    sigma = rng.uniform(0.05, 2.0, K)
    got = npb.ScalarNoiseNormal(ctx, ds, FAM[family]).logprobability(mu, sigma)
    want = oracle.scalarnoise_logpdf_batch(FAM[family], mu, sigma, X)
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    print(family, "max relative error", err.max())
    assert err.max() < 1e-12
    rows = np.array([7, 3, 499], dtype=np.int64)
    assert np.allclose(npb.ScalarNoiseNormal(ctx, ds, FAM[family]).logprobability(mu, sigma, rows=rows), want[rows], rtol=1e-12, atol=1e-12)
    ds.close()


def test_base_measure_law(npb, ctx):
    from scipy import stats as sps
    X, _ = data("regression", 64, 1)
    ds = npb.Dataset(ctx, X)
    pr = syn.reference_nig_prior()
    npb.NormalInverseGamma(npb.FAMILY_REGRESSION, pr["mu0"], pr["Lambda"], pr["nig_alpha"], pr["nig_beta"], pr["alpha"]).bind(ctx)
    ch = npb.Chains(ctx, ds, 4, Kmax=64, K0=20, seed=5)
    d = ch.sample_base_nig(2, 200_000).astype(np.float64)
    # 1 / sigma^2 ~ Gamma(shape alpha, scale beta) (gamma.h:41); mu | sigma ~ N(mu0, sigma^2 Lambda^-1) (normalinvgamma.h:72-76)
    assert sps.kstest(1.0 / d[:, 2] ** 2, "gamma", args=(pr["nig_alpha"], 0, pr["nig_beta"])).pvalue > 1e-3
    zs = d[:, :2] / (d[:, 2:3] * 10.0)
    assert sps.kstest(zs[:, 0], "norm").pvalue > 1e-3 and sps.kstest(zs[:, 1], "norm").pvalue > 1e-3
    assert abs(np.corrcoef(zs.T)[0, 1]) < 0.01
    ch.close()
    ds.close()


@pytest.mark.parametrize("family", ["regression", "angular"])
def test_invariants_and_unsupported_samplers(npb, ctx, oracle, family):
    N = 700 + 13
    X, y = data(family, N, 21, K=3)
    ds = npb.Dataset(ctx, X)
    pr = syn.reference_nig_prior()
    prior = npb.NormalInverseGamma(FAM[family], pr["mu0"], pr["Lambda"], pr["nig_alpha"], pr["nig_beta"], pr["alpha"])
    mc = npb.MCMC(ctx, ds, prior, chains=37, Kmax=64, K0=20, seed=8)
    births = 0
    for _ in range(4):
        st = mc.chains.sweep(npb.ALG8, 5)
        births += st.new_clusters
        assert st.overflow_chains == 0 and st.reassignments == 37 * N * 5
    z = mc.getMembershipMatrix()
    m = mc.chains.metrics(y)
    for c in (0, 11, 36):
        slots, counts, mu, Sigma = mc.chains.params(c)
        assert counts.sum() == N and np.array_equal(np.bincount(z[c], minlength=64)[slots], counts)
        assert len(slots) == m["K"][c]
        assert np.allclose([m["purity"][c], m["rand_index"][c], m["adjusted_rand"][c]], oracle.metrics(y, z[c]), atol=1e-12)
        # the joint log-likelihood of the metrics kernel against the oracle density of the reported parameters
        sg = np.sqrt(Sigma[:, 0, 0])
        if family == "regression":
            lp = oracle.scalarnoise_logpdf_batch(FAM[family], mu[:, :2], sg, X)
        else:  # (the reported (d, theta) are the canonical values prepare() produced, not something to truncate again)
            q1 = -np.sin(mu[None, :, 1]) * X[:, None, 0] + np.cos(mu[None, :, 1]) * X[:, None, 1]
            lp = -0.5 * (mu[None, :, 0] - q1) ** 2 / sg[None] ** 2 - 0.5 * np.log(2 * np.pi * sg[None] ** 2)
        inv = {s: i for i, s in enumerate(slots)}
        want = sum(lp[i, inv[z[c][i]]] for i in range(N))
        assert abs(m["joint_loglik"][c] - want) < 1e-3 * max(1.0, abs(want)), (m["joint_loglik"][c], want)
    assert births > 0
    print(family, "births", births, "mean K", m["K"].mean(), "mean purity", m["purity"].mean())
    with pytest.raises(npb.NpbError):
        mc.chains.sweep(npb.JAIN_NEAL, 1)
    ds.close()
    npb.NormalInverseWishart(**syn.reference_prior(2)).bind(ctx)


def _oracle_seed(args):
    family, seed, T = args
    from oracle import binding as orc
    X, y = data(family, 150, 31)
    pr = syn.reference_nig_prior()
    r = orc.ScalarNoiseRun(FAM[family], pr, X, orc.ALG8, T=T, seed_main=100 + seed, seed_shuffle=900 + seed, flags=0)
    s = r.stats()
    pur, ri, ari = orc.metrics(y, r.assignments(0))
    return s.K_final, pur, ari, s.new_cluster_events / s.updates, s.moved / s.updates


@pytest.mark.parametrize("family", ["regression", "angular"])
def test_distribution_against_oracle(npb, ctx, family):
    """192 oracle seeds (the reference's sampler, pinned) against 1024 device chains, T = 200 sweeps from the reference's own
    initialisation: K, purity, ARI in distribution (two-sample KS), birth and move rates over the whole run."""
    from multiprocessing import Pool
    from scipy import stats as sps
    T = 200
    with Pool(8) as pool:
        res = np.array(pool.map(_oracle_seed, [(family, s, T) for s in range(192)]))
    X, y = data(family, 150, 31)
    ds = npb.Dataset(ctx, X)
    pr = syn.reference_nig_prior()
    prior = npb.NormalInverseGamma(FAM[family], pr["mu0"], pr["Lambda"], pr["nig_alpha"], pr["nig_beta"], pr["alpha"])
    mc = npb.MCMC(ctx, ds, prior, chains=1024, Kmax=64, K0=20, seed=23)
    stats = mc.run(T, sweeps_per_launch=50)
    assert all(s.overflow_chains == 0 for s in stats)
    m = mc.chains.metrics(y)
    n = sum(s.reassignments for s in stats)
    births = sum(s.new_clusters for s in stats) / n
    moved = sum(s.moved for s in stats) / n
    print("%s: K gpu %.2f oracle %.2f; purity gpu %.3f oracle %.3f; births/step gpu %.5f oracle %.5f; moved gpu %.4f oracle %.4f" % (
        family, m["K"].mean(), res[:, 0].mean(), m["purity"].mean(), res[:, 1].mean(), births, res[:, 3].mean(), moved, res[:, 4].mean()))
    for name, got, want in (("K", m["K"].astype(float), res[:, 0]), ("purity", m["purity"], res[:, 1]), ("ari", m["adjusted_rand"], res[:, 2])):
        p = sps.ks_2samp(got, want).pvalue
        assert p > 0.005, "%s: KS p=%.2e (gpu %.4f vs oracle %.4f)" % (name, p, got.mean(), want.mean())
    assert abs(births - res[:, 3].mean()) < 0.1 * res[:, 3].mean() + 2e-5
    assert abs(moved - res[:, 4].mean()) < 0.1 * res[:, 4].mean() + 2e-4
    ds.close()
    npb.NormalInverseWishart(**syn.reference_prior(2)).bind(ctx)
