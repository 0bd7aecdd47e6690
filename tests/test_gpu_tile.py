"""GPU (pytest -m gpu): the D = 4, 8, 16 sweep kernel (shared-memory slot table, producer/consumer warps)."""
import numpy as np
import pytest
from scipy import stats as sps

from noparama_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def invariants(chains, z, c, N):
    slots, counts, mu, Sigma = chains.params(c)
    assert counts.sum() == N
    bc = np.bincount(z, minlength=chains.Kmax)
    assert np.array_equal(np.nonzero(bc)[0], slots) and np.array_equal(bc[slots], counts)
    assert np.all(np.isfinite(mu)) and np.all(np.isfinite(Sigma))
    return len(slots)


@pytest.mark.parametrize("D,kmax", [(4, 32), (4, 64), (8, 32), (16, 32), (16, 64)])
def test_tile_kernel_invariants(npb, ctx, oracle, D, kmax):
    N = (200 if kmax == 32 else 1000) + D  # not a multiple of the 32-step tile; small enough for 32 slots
    X, y = syn.gmm(N, D, 4, 100 + D)
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=24, Kmax=kmax,
                  K0=20 if kmax > 32 else 8, seed=D)
    st = None
    for _ in range(3):
        st = mc.chains.sweep(npb.ALG8, 3)
        assert st.overflow_chains == 0 and st.reassignments == 24 * N * 3
        assert 4 * st.reassignments <= st.candidates <= (kmax + 3) * st.reassignments
    z = mc.getMembershipMatrix()
    m = mc.chains.metrics(y)
    for c in range(0, 24, 5):
        k = invariants(mc.chains, z[c], c, N)
        assert k == m["K"][c]
        want = oracle.metrics(y, z[c])
        assert np.allclose([m["purity"][c], m["rand_index"][c], m["adjusted_rand"][c]], want, atol=1e-12)
    assert abs(st.mean_K - m["K"].mean()) < 1e-9
    # determinism and launch splitting
    k0 = 20 if kmax > 32 else 8  # 20 initial clusters plus the first sweeps' births can exceed 32 slots
    a = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=6, Kmax=kmax, K0=k0, seed=77)
    b = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=6, Kmax=kmax, K0=k0, seed=77)
    a.run(4)
    b.run(4, sweeps_per_launch=1)
    assert np.array_equal(a.getMembershipMatrix(), b.getMembershipMatrix())
    ds.close()


def test_tile_kernel_recovers_given_clusters_16d(npb, ctx):
    """Config-5 regime: chains start from K_true known clusters (init_from_params) and a random assignment; a few
    sweeps of Alg. 8 with frozen parameters must put (almost) every item into its own component."""
    X, y = syn.gmm(4000, 16, 8, 5)
    means = np.stack([X[y == k].mean(0) for k in range(8)])
    Sigma = np.tile(np.eye(16), (8, 1, 1))
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(16)), chains=32, Kmax=32, seed=3)
    mc.chains.init_from_params(means, Sigma)
    m0 = mc.chains.metrics(y)
    assert m0["purity"].mean() < 0.3 and np.all(m0["K"] == 8)
    st = mc.run(3)[0]
    assert st.overflow_chains == 0
    m = mc.chains.metrics(y)
    assert m["purity"].mean() > 0.99 and m["adjusted_rand"].mean() > 0.98
    assert np.all(m["K"] >= 8)
    z = mc.getMembershipMatrix(0, 2)
    for c in range(2):
        invariants(mc.chains, z[c], c, 4000)
    ds.close()


def test_tile_kernel_distribution_matches_log_domain_oracle_4d(npb, ctx, oracle):
    """D = 4: 96 device chains against 96 oracle seeds (oracle in its log-domain mode: the same categorical without the
    reference's double underflow), same prior, T = 150 sweeps: K, purity and adjusted Rand must agree in distribution."""
    from multiprocessing import Pool
    X, y = syn.gmm(240, 4, 3, 42, min_dist=5.0)
    pr = syn.reference_prior(4)
    T = 150
    with Pool(8) as pool:
        res = np.array(pool.map(_oracle_seed, [(s, T) for s in range(96)]))
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**pr), chains=96, Kmax=64, seed=11)
    stats = mc.run(T, sweeps_per_launch=50)
    assert all(s.overflow_chains == 0 for s in stats)
    m = mc.chains.metrics(y)
    for name, got, want in (("K", m["K"].astype(float), res[:, 0]), ("purity", m["purity"], res[:, 1]),
                            ("ari", m["adjusted_rand"], res[:, 2])):
        p = sps.ks_2samp(got, want).pvalue
        assert p > 0.01, "%s: KS p=%.2e (gpu %.4f vs oracle %.4f)" % (name, p, got.mean(), want.mean())
    ds.close()


def _oracle_seed(args):
    seed, T = args
    from oracle import binding as orc
    X, y = syn.gmm(240, 4, 3, 42, min_dist=5.0)
    p = orc.make_prior(**syn.reference_prior(4))
    r = orc.Run(p, X, T=T, seed_main=500 + seed, seed_shuffle=900 + seed, flags=orc.LOG_DOMAIN)
    s = r.stats()
    pur, ri, ari = orc.metrics(y, r.assignments(0))
    return s.K_final, pur, ari


def _oracle_seed_k0(args):
    seed, T, K0 = args
    from oracle import binding as orc
    X, y = syn.gmm(240, 4, 3, 42, min_dist=5.0)
    p = orc.make_prior(**syn.reference_prior(4))
    r = orc.Run(p, X, T=T, K0=K0, seed_main=700 + seed, seed_shuffle=1900 + seed, flags=orc.LOG_DOMAIN)
    s = r.stats()
    pur, ri, ari = orc.metrics(y, r.assignments(0))
    return s.K_final, pur, ari, s.new_cluster_events / s.updates, s.moved / s.updates


def test_tile4_kernel_distribution_and_birth_rate_4d(npb, ctx, oracle):
    """The four-chains-per-CTA kernel (Kmax = 32) draws an auxiliary candidate's race key from the non-central
    chi-square form and materialises theta' only at a birth (npb_alg8_tile4.cuh).  128 device chains against 128 oracle
    seeds (1024 device chains) from the same K0 = 8 start: K, purity, ARI in distribution, and the birth and move rates over the whole run --
    a newborn cluster built inconsistently with its key would die at a different rate."""
    from multiprocessing import Pool
    X, y = syn.gmm(240, 4, 3, 42, min_dist=5.0)
    pr = syn.reference_prior(4)
    T, K0 = 150, 8
    with Pool(8) as pool:
        res = np.array(pool.map(_oracle_seed_k0, [(s, T, K0) for s in range(128)]))
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**pr), chains=1024, Kmax=32, K0=K0, seed=13)
    stats = mc.run(T, sweeps_per_launch=50)
    assert all(s.overflow_chains == 0 for s in stats)
    m = mc.chains.metrics(y)
    for name, got, want in (("K", m["K"].astype(float), res[:, 0]), ("purity", m["purity"], res[:, 1]),
                            ("ari", m["adjusted_rand"], res[:, 2])):
        p = sps.ks_2samp(got, want).pvalue
        assert p > 0.01, "%s: KS p=%.2e (gpu %.4f vs oracle %.4f)" % (name, p, got.mean(), want.mean())
    n = sum(s.reassignments for s in stats)
    births = sum(s.new_clusters for s in stats) / n
    moved = sum(s.moved for s in stats) / n
    print("births/step gpu %.5f oracle %.5f; moved gpu %.4f oracle %.4f" % (births, res[:, 3].mean(), moved, res[:, 4].mean()))
    assert abs(births - res[:, 3].mean()) < 0.05 * res[:, 3].mean() + 1e-5
    assert abs(moved - res[:, 4].mean()) < 0.05 * res[:, 4].mean() + 1e-4
    ds.close()


def test_full_size_properties_config5_shard(npb, ctx):
    """BASELINE configs[4] per-GPU shard shape (16-D, N = 100 000, Kmax 32) at a reduced chain count: size-independent
    properties -- every chain keeps its bookkeeping consistent, the given well-separated clusters are recovered exactly,
    counters add up, readback and metrics are idempotent, a second sweep moves (almost) nothing."""
    X, y = syn.config(5)
    K = int(y.max()) + 1
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(16)), chains=44, Kmax=32, seed=12)
    means = np.stack([X[y == k].mean(0) for k in range(K)])
    mc.chains.init_from_params(means, np.tile(np.eye(16), (K, 1, 1)))
    s1 = mc.chains.sweep(npb.ALG8, 1)
    s2 = mc.chains.sweep(npb.ALG8, 1)
    assert s1.overflow_chains == 0 and s1.reassignments == 44 * 100000 == s2.reassignments
    assert s1.candidates == 35 * s1.reassignments == s2.candidates     # 32 occupied clusters + 3 auxiliary draws, every step
    assert s1.moved > 0.9 * s1.reassignments and s2.moved < 1e-4 * s2.reassignments
    z = mc.getMembershipMatrix(0, 3)
    for c in range(3):
        invariants(mc.chains, z[c], c, ds.N)
    m = mc.chains.metrics(y)
    assert np.all(m["K"] == K) and m["purity"].min() > 0.9999 and m["adjusted_rand"].min() > 0.9999
    assert np.array_equal(mc.getMembershipMatrix(0, 3), z)
    m2 = mc.chains.metrics(y)
    assert np.array_equal(m["joint_loglik"], m2["joint_loglik"])
    ds.close()


def test_tile4_algorithm2_single_auxiliary_draw(npb, ctx):
    """m_aux = 1 (the sampler np_neal_algorithm2.cpp describes) through the D >= 4 kernels: bookkeeping invariants and the
    candidate count of K + 1 per reassignment."""
    X, y = syn.gmm(1000, 8, 4, 55)
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(8)), npb.NealAlgorithm2, chains=20, Kmax=32, K0=8,
                  m_aux=1, seed=3)
    means = np.stack([X[y == k].mean(0) for k in range(4)])
    mc.chains.init_from_params(means, np.tile(np.eye(8), (4, 1, 1)))
    st = mc.run(3)[0]
    assert st.overflow_chains == 0 and st.reassignments == 20 * 1000 * 3
    assert 4 * st.reassignments < st.candidates <= 6 * st.reassignments   # K = 4 (+ rare births) + 1 auxiliary draw
    z = mc.getMembershipMatrix()
    for c in (0, 19):
        invariants(mc.chains, z[c], c, ds.N)
    assert mc.chains.metrics(y)["purity"].mean() > 0.99
    ds.close()


@pytest.mark.parametrize("D", [4, 8, 16])
def test_tile4_producer_logdensity_within_1e5_of_oracle(npb, ctx, oracle, D):
    """The log-density tile of the sweep kernel's producer warp (packed FP32, y = T2 x - T2 mu with the mean folded into a
    per-row offset) against the oracle's double-precision density: 1e-5 relative (north_star tolerance), for full
    covariances, near and far clusters."""
    rng = np.random.default_rng(D)
    K = 32
    X, y = syn.gmm(4000, D, 8, 200 + D)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(D)).bind(ctx)
    ch = npb.Chains(ctx, ds, 3, Kmax=32, K0=8, seed=1)
    mu = X[rng.integers(0, len(X), K)] + 0.5 * rng.standard_normal((K, D))
    B = rng.standard_normal((K, D, D)) / np.sqrt(D)
    Sigma = B @ np.transpose(B, (0, 2, 1)) + 0.3 * np.eye(D)
    ch.init_from_params(mu, Sigma)
    items = rng.integers(0, len(X), 32)
    got = ch.probe_tile_logdensity(1, items).astype(np.float64)     # [slot, item]
    slots, counts, _, _ = ch.params(1)
    want = oracle.mvn_logpdf_batch(mu, Sigma, X[items]).T           # [slot, item]
    occ = np.zeros(32, bool)
    occ[slots] = True
    assert occ.sum() >= 30 and np.all(np.isnan(got[~occ]))
    err = np.abs(got[occ] - want[occ]) / np.maximum(1.0, np.abs(want[occ]))
    assert err.max() < 1e-5, err.max()
    ch.close()
    ds.close()
