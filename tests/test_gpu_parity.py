"""GPU (pytest -m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the committed fixtures.

Bars: log-densities within 1e-5 relative (north_star) -- 1e-12 for the double path; integer work (counts, slot
bookkeeping, contingency metrics) exact; sampler statistics distributionally indistinguishable from the oracle
over 256 seeds/chains.  Nothing here reads /root/reference."""
import os

import numpy as np
import pytest
from scipy import stats as sps

from noparama_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(got, want):
    return np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want)))


# ------------------------------------------------------------------------------------------------ density ----
def test_density_kat(npb, ctx):
    """test/test_mvn_likelihood.cpp:18-44 through npb_logdensity_batch (non-symmetric Sigma)."""
    ds = npb.Dataset(ctx, np.array([[1.0, 2.0], [1.0, 2.0]]))
    mvn = npb.MultivariateNormal(ctx, ds)
    mu, Sigma = np.array([[1.0, 1.0]]), np.array([[[2.0, 0.0], [1.0, 2.0]]])
    for precision, tol in ((64, 1e-12), (32, 1e-6)):
        p = mvn.probability(mu, Sigma, precision=precision)[:, 0]
        assert p[0] - 0.061974 < 0.00001 and abs(p[0] - 0.0619749972) < max(tol * 10, 1e-9)
        assert p[0] * p[1] - 0.0038409 < 0.00001
    assert abs(np.exp(mvn.logprobability_dataset(mu, Sigma)[0]) - 0.0038409003) < 1e-9
    ds.close()


@pytest.mark.parametrize("D", [2, 16, 64])
def test_density_golden_and_oracle(npb, ctx, oracle, D):
    g = np.load(os.path.join(GOLDEN, "density_cases.npz"))
    X, mu, Sigma, want = g["X%d" % D], g["mu%d" % D], g["Sigma%d" % D], g["logp%d" % D]
    assert np.allclose(oracle.mvn_logpdf_batch(mu, Sigma, X), want, rtol=1e-12, atol=1e-9)  # fixture == live oracle
    ds = npb.Dataset(ctx, X)
    mvn = npb.MultivariateNormal(ctx, ds)
    got64 = mvn.logprobability(mu, Sigma, precision=64)
    assert rel_err(got64, want) < 1e-10
    got32 = mvn.logprobability(mu, Sigma, precision=32)
    assert rel_err(got32, want) < 1e-5  # north_star tolerance on log-densities
    rows = np.array([5, 0, 47, 5], dtype=np.int64)
    assert np.array_equal(mvn.logprobability(mu, Sigma, rows=rows), got64[rows])
    sums = mvn.logprobability_dataset(mu, Sigma, rows=rows)
    assert np.allclose(sums, want[rows].sum(0), rtol=1e-10)
    ds.close()


def test_density_large_random(npb, ctx, oracle):
    """10^5 (x, theta) pairs per D against the oracle, thetas drawn from the reference prior."""
    for D in (2, 16):
        X, _ = syn.gmm(2000, D, 8, 5 + D)
        pr = syn.reference_prior(D)
        mu, Sigma = oracle.sample_base(oracle.make_prior(**pr), 17, 50)
        want = oracle.mvn_logpdf_batch(mu, Sigma, X)
        ds = npb.Dataset(ctx, X)
        mvn = npb.MultivariateNormal(ctx, ds)
        assert rel_err(mvn.logprobability(mu, Sigma, precision=64), want) < 1e-10
        assert rel_err(mvn.logprobability(mu, Sigma, precision=32), want) < 1e-5
        ds.close()


def test_density_rejects_singular(npb, ctx):
    ds = npb.Dataset(ctx, np.zeros((4, 2)))
    mvn = npb.MultivariateNormal(ctx, ds)
    with pytest.raises(npb.NpbError) as e:
        mvn.logprobability(np.zeros((1, 2)), np.array([[[1.0, 1.0], [1.0, 1.0]]]))
    assert e.value.status == -4
    ds.close()


# ------------------------------------------------------------------------------------------------ sweeps -----
def check_chain_invariants(chains, z, c, N):
    slots, counts, mu, Sigma = chains.params(c)
    assert counts.sum() == N
    bc = np.bincount(z, minlength=chains.Kmax)
    assert np.array_equal(np.nonzero(bc)[0], slots)          # every used slot is occupied and vice versa
    assert np.array_equal(bc[slots], counts)                 # member counts == vector sizes (membertrix.cpp:328-330)
    assert np.all(np.isfinite(mu)) and np.all(np.isfinite(Sigma))
    return len(slots)


def test_alg8_state_invariants_config1(npb, ctx, oracle):
    X, y = syn.config(1)
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(2)), chains=96, Kmax=64, seed=5)
    z0 = mc.getMembershipMatrix()
    for c in (0, 50, 95):
        assert 1 <= check_chain_invariants(mc.chains, z0[c], c, ds.N) <= 20  # K0 = 20 minus empties (np_mcmc.cpp:49-91)
    total = None
    for _ in range(3):
        st = mc.chains.sweep(npb.ALG8, 10)
        assert st.reassignments == 96 * 200 * 10 and st.overflow_chains == 0
        # candidates = sum (K_i + m): between (1+3) and (Kmax+3) per step, and consistent with mean K
        assert 4 * st.reassignments <= st.candidates <= 67 * st.reassignments
        assert 0 < st.new_clusters < st.moved < st.reassignments
        total = st
    z = mc.getMembershipMatrix()
    m = mc.chains.metrics(y)
    for c in range(0, 96, 7):
        k = check_chain_invariants(mc.chains, z[c], c, ds.N)
        assert k == m["K"][c]
        want = oracle.metrics(y, z[c])
        assert np.allclose([m["purity"][c], m["rand_index"][c], m["adjusted_rand"][c]], want, rtol=0, atol=1e-12)
        # joint log-likelihood of np_mcmc.cpp:187-203 against the oracle's density on the chain's own parameters
        slots, counts, mu, Sigma = mc.chains.params(c)
        lp = oracle.mvn_logpdf_batch(mu, Sigma, X)
        idx = np.searchsorted(slots, z[c])
        want_jll = lp[np.arange(ds.N), idx].sum()
        assert abs(m["joint_loglik"][c] - want_jll) < 1e-3 * max(1.0, abs(want_jll))
    assert abs(total.mean_K - m["K"].mean()) < 1e-9
    mc.chains.close()
    ds.close()


def test_alg8_is_deterministic_and_seed_sensitive(npb, ctx):
    X, _ = syn.config(1)
    ds = npb.Dataset(ctx, X)
    prior = npb.NormalInverseWishart(**syn.reference_prior(2))
    runs = []
    for seed in (1, 1, 2):
        mc = npb.MCMC(ctx, ds, prior, chains=8, Kmax=64, seed=seed)
        mc.run(7)
        runs.append(mc.getMembershipMatrix().copy())
        mc.chains.close()
    assert np.array_equal(runs[0], runs[1])
    assert not np.array_equal(runs[0], runs[2])
    # sweeps in one launch == the same sweeps in several launches (counter-based RNG, restartable)
    a = npb.MCMC(ctx, ds, prior, chains=8, Kmax=64, seed=9)
    a.run(6)
    b = npb.MCMC(ctx, ds, prior, chains=8, Kmax=64, seed=9)
    b.run(6, sweeps_per_launch=2)
    assert np.array_equal(a.getMembershipMatrix(), b.getMembershipMatrix())
    ds.close()


def test_alg8_ragged_sizes(npb, ctx):
    """N not a multiple of the 32-step tile, tiny N, Kmax levels 1..16, D = 3."""
    prior2 = npb.NormalInverseWishart(**syn.reference_prior(2))
    for N, kmax in ((2, 32), (31, 32), (33, 64), (257, 128), (1000, 512)):
        X, _ = syn.gmm(N, 2, 2, N, min_dist=3.0)
        ds = npb.Dataset(ctx, X)
        mc = npb.MCMC(ctx, ds, prior2, chains=5, Kmax=kmax, K0=min(20, kmax), seed=N)
        mc.run(4)
        z = mc.getMembershipMatrix()
        for c in range(5):
            check_chain_invariants(mc.chains, z[c], c, N)
        mc.chains.close()
        ds.close()
    X, _ = syn.gmm(500, 3, 4, 11)
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(3)), chains=7, Kmax=128, seed=3)
    mc.run(5)
    z = mc.getMembershipMatrix()
    for c in range(7):
        check_chain_invariants(mc.chains, z[c], c, 500)
    ds.close()


def test_kmax_overflow_is_reported(npb, ctx):
    X, _ = syn.gmm(20000, 2, 10, 3, min_dist=3.0)
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(2)), chains=4, Kmax=32, K0=20, seed=1)
    with pytest.raises(npb.NpbError) as e:
        for _ in range(10):
            mc.chains.sweep(npb.ALG8, 1)
    assert e.value.status == -3
    ds.close()


def test_set_state_roundtrip_and_first_step_weights(npb, ctx, oracle):
    """State injected through the ABI comes back unchanged; Sigma <-> triangular precision factor round trip."""
    X, y = syn.config(1)
    p = oracle.make_prior(**syn.reference_prior(2))
    r = oracle.Run(p, X, T=1, seed_main=3, seed_shuffle=4, flags=0)
    z0, slots, mu, Sigma = r.init_state()
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(2)), chains=3, Kmax=32, seed=1)
    mc.chains.set_state(1, z0, slots, mu, Sigma)
    assert np.array_equal(mc.getMembershipMatrix(1, 1)[0], z0)
    s2, counts, mu2, Sigma2 = mc.chains.params(1)
    occ = np.unique(z0)
    assert np.array_equal(s2, occ)
    sel = np.searchsorted(slots, occ)
    assert np.allclose(mu2, mu[sel], rtol=1e-6) and np.allclose(Sigma2, Sigma[sel], rtol=1e-5)
    assert np.array_equal(counts, np.bincount(z0, minlength=32)[occ])
    ds.close()


def test_sweep_host_end_to_end(npb, ctx):
    """npb_chains_sweep_host: X up, sweep, every assignment down (item-major uint16) == the device state."""
    X, _ = syn.config(1)
    ds = npb.Dataset(ctx, X)
    prior = npb.NormalInverseWishart(**syn.reference_prior(2))
    a = npb.MCMC(ctx, ds, prior, chains=16, Kmax=64, seed=4)
    b = npb.MCMC(ctx, ds, prior, chains=16, Kmax=64, seed=4)
    zh = np.empty((ds.N, 16), dtype=np.uint16)
    for _ in range(3):
        a.chains.sweep_host(X, npb.ALG8, 1, z_out=zh)
        b.chains.sweep(npb.ALG8, 1)
    assert np.array_equal(zh.T.astype(np.int32), a.getMembershipMatrix())
    assert np.array_equal(a.getMembershipMatrix(), b.getMembershipMatrix())
    ds.close()


# ------------------------------------------------------------------------------- distributional parity ------
def ks_ok(a, b, name, alpha=0.01):
    p = sps.ks_2samp(a, b).pvalue
    assert p > alpha, "%s: KS p=%.2e (gpu mean %.4f, oracle mean %.4f)" % (name, p, np.mean(a), np.mean(b))


def test_alg8_distribution_matches_oracle_256_seeds(npb, ctx):
    """Parity gate 4: purity / Rand / adjusted Rand / K after T=1000 sweeps of config 1,
    256 GPU chains against 256 oracle seeds (fixture made by tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, "oracle_cfg1_alg8_256seeds.npz"))
    X, y = syn.config(1)
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(2)), chains=256, Kmax=64, seed=20261018)
    stats = mc.run(int(g["T"]), sweeps_per_launch=250)
    assert all(s.overflow_chains == 0 for s in stats)
    m = mc.chains.metrics(y)
    ks_ok(m["purity"], g["purity"], "purity")
    ks_ok(m["rand_index"], g["rand"], "rand index")
    ks_ok(m["adjusted_rand"], g["ari"], "adjusted rand index")
    ks_ok(m["K"].astype(float), g["K_final"], "K")
    # rates over the whole run: moved fraction and births per reassignment
    moved = sum(s.moved for s in stats) / sum(s.reassignments for s in stats)
    births = sum(s.new_clusters for s in stats) / sum(s.reassignments for s in stats)
    assert abs(moved - g["moved"].mean()) < 0.01
    assert abs(births - g["births"].mean()) < 1e-4
    # the reference's own qualitative claim (README.rst:55)
    assert m["purity"].mean() > 0.98 and m["rand_index"].mean() < m["purity"].mean()
    assert m["adjusted_rand"].mean() < m["rand_index"].mean()
    ds.close()


# ------------------------------------------------------------------------------- full-size properties -------
def test_full_size_properties_config2(npb, ctx):
    """BASELINE config 2 shape (N = 100k, D = 2) at reduced chain count: size-independent properties."""
    X, y = syn.config(2)
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(2)), chains=64, Kmax=256, seed=8)
    st = mc.run(4)[0]
    assert st.overflow_chains == 0 and st.reassignments == 64 * 100000 * 4
    z = mc.getMembershipMatrix(0, 4)
    for c in range(4):
        check_chain_invariants(mc.chains, z[c], c, ds.N)
    m = mc.chains.metrics(y)
    assert np.all(m["K"] == np.array([len(np.unique(z[c])) for c in range(4)] + list(m["K"][4:])))
    assert 0.6 < m["purity"].mean() <= 1.0
    # idempotence of the readback path and of the metrics
    assert np.array_equal(mc.getMembershipMatrix(0, 4), z)
    m2 = mc.chains.metrics(y)
    assert np.array_equal(m["purity"], m2["purity"]) and np.array_equal(m["joint_loglik"], m2["joint_loglik"])
    ds.close()


def test_cocluster_counts(npb, ctx):
    import torch
    X, _ = syn.config(1)
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(2)), chains=40, Kmax=64, seed=2)
    mc.run(20)
    anchors = np.arange(0, 200, 5)
    S = torch.zeros((40, 40), dtype=torch.float32, device="cuda")
    mc.chains.cocluster_into(anchors, S.data_ptr())
    z = mc.getMembershipMatrix()[:, anchors]
    want = (z[:, :, None] == z[:, None, :]).sum(0)
    assert np.array_equal(S.cpu().numpy(), want.astype(np.float32))
    mc.chains.cocluster_into(anchors, S.data_ptr(), accumulate=True)
    assert np.array_equal(S.cpu().numpy(), 2 * want.astype(np.float32))
    ds.close()


# ------------------------------------------------------------------------------- bit-exact replay ------------
def test_alg8_replay_bit_exact_config1(npb, ctx, oracle):
    """Parity gate 3: config 1 (200 items x 1000 sweeps) with the oracle's recorded draws replayed in
    double precision on the device: every pick and the assignments after every sweep must be identical."""
    X, y = syn.config(1)
    p = oracle.make_prior(**syn.reference_prior(2))
    T = 1000
    r = oracle.Run(p, X, T=T, seed_main=31, seed_shuffle=32, flags=oracle.RECORD_TRACE | oracle.UPDATE_CLUSTERS)
    t = r.trace()
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(2)).bind(ctx)
    picked, z_after = npb.replay_alg8(ctx, ds, t, r.init_state(), m_aux=3, z_every=ds.N)
    assert len(picked) == T * 200
    assert np.array_equal(picked, t["picked"])
    assert np.array_equal(z_after, t["z_after"])
    # and the final state carries the same clustering metrics as the oracle's own snapshot
    assert np.allclose(oracle.metrics(y, z_after[-1]), oracle.metrics(y, r.assignments(0)))
    ds.close()


def test_alg8_replay_general_covariance_3d(npb, ctx, oracle):
    """Replay with a non-diagonal Lambda in 3-D (full covariances through the triangular-factor path)."""
    X, _ = syn.gmm(150, 3, 3, 21)
    Lam = np.array([[0.02, 0.004, 0.0], [0.004, 0.015, 0.003], [0.0, 0.003, 0.01]])
    prior = dict(mu0=np.full(3, 6.0), kappa=1.0 / 500, nu=5.0, Lambda=Lam, alpha=1.0)
    r = oracle.Run(oracle.make_prior(**prior), X, T=60, seed_main=5, seed_shuffle=6, flags=oracle.RECORD_TRACE)
    t = r.trace()
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**prior).bind(ctx)
    picked, z_after = npb.replay_alg8(ctx, ds, t, r.init_state(), m_aux=3, z_every=ds.N)
    assert np.array_equal(picked, t["picked"]) and np.array_equal(z_after, t["z_after"])
    ds.close()


def test_max_likelihood_snapshot(npb, ctx, oracle):
    """MCMC::considerMaxLikelihood (np_mcmc.cpp:187-203) on the device: the kept state of every chain is the one with
    the highest joint log-likelihood among the states it was shown, chain by chain."""
    X, y = syn.config(1)
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(2)), chains=48, Kmax=64, seed=21)
    best_j = np.full(48, -np.inf)
    best_z = np.zeros((48, ds.N), np.int32)
    for it in range(12):
        mc.chains.sweep(npb.ALG8, 5)
        z = mc.getMembershipMatrix()
        cur, best = mc.chains.consider_max_likelihood()
        want = mc.chains.metrics(None)["joint_loglik"]
        assert np.array_equal(cur, want)
        better = cur > best_j
        best_j = np.where(better, cur, best_j)
        best_z[better] = z[better]
        assert np.array_equal(best, best_j)
        assert np.array_equal(mc.getMaxLikelihoodMatrix(), best_z)
    assert (best_j >= cur).all() and not np.array_equal(best_z, z)   # some chain kept an earlier state
    ds.close()


def _oracle_alg2_seed(args):
    seed, T = args
    from oracle import binding as orc
    X, y = syn.config(1)
    r = orc.Run(orc.make_prior(**syn.reference_prior(2)), X, T=T, M_aux=1, seed_main=800 + seed, seed_shuffle=2800 + seed,
                flags=orc.UPDATE_CLUSTERS)
    s = r.stats()
    pur, ri, ari = orc.metrics(y, r.assignments(0))
    return s.K_final, pur, ri, ari, s.new_cluster_events / s.updates, s.moved / s.updates


def test_algorithm2_distribution_matches_oracle(npb, ctx):
    """NPB_ALG2 = the sampler np_neal_algorithm2.cpp:32-120 describes (dead code in the reference): K occupied clusters +
    ONE prior draw weighted alpha.  The oracle runs NealAlgorithm8's code path with M = 1 (the same categorical); 512
    device chains against 128 oracle seeds on config 1."""
    from multiprocessing import Pool
    T = 300
    with Pool(8) as pool:
        res = np.array(pool.map(_oracle_alg2_seed, [(s, T) for s in range(128)]))
    X, y = syn.config(1)
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(2)), npb.NealAlgorithm2, chains=512, Kmax=64, m_aux=1,
                  seed=31)
    stats = mc.run(T, sweeps_per_launch=100)
    assert all(s.overflow_chains == 0 for s in stats)
    m = mc.chains.metrics(y)
    for name, got, want in (("K", m["K"].astype(float), res[:, 0]), ("purity", m["purity"], res[:, 1]),
                            ("rand", m["rand_index"], res[:, 2]), ("ari", m["adjusted_rand"], res[:, 3])):
        p = sps.ks_2samp(got, want).pvalue
        assert p > 0.01, "%s: KS p=%.2e (gpu %.4f vs oracle %.4f)" % (name, p, got.mean(), want.mean())
    n = sum(s.reassignments for s in stats)
    assert sum(s.candidates for s in stats) < sum(s.reassignments for s in stats) * (m["K"].max() + 40)
    births = sum(s.new_clusters for s in stats) / n
    assert abs(births - res[:, 4].mean()) < 0.08 * res[:, 4].mean() + 2e-5
    # an ALG8 handle (m_aux = 3) refuses the Algorithm-2 sweep
    mc3 = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(2)), chains=4, Kmax=64, seed=1)
    with pytest.raises(npb.NpbError):
        mc3.chains.sweep(npb.ALG2, 1)
    ds.close()


def test_single_item_seam_update(npb, ctx, oracle):
    """npb_chain_update_alg8: one NealAlgorithm8::update(membertrix&, {item}) per call.  Every chain is put in the same
    state; the empirical distribution of the cluster the item goes to must match the reference's categorical
    p(x|theta_k) n_k / sum (np_neal_algorithm8.cpp:93-130, the item itself retracted) evaluated with the oracle's density."""
    X, y = syn.config(1)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(2)).bind(ctx)
    C = 4096
    ch = npb.Chains(ctx, ds, C, Kmax=32, K0=8, seed=17)
    # three overlapping clusters so that the item has a real choice
    mu = np.array([[0.0, 0.0], [1.5, 1.0], [5.0, 5.0]])
    Sigma = np.array([np.eye(2), [[1.5, 0.3], [0.3, 0.8]], np.eye(2)])
    ch.init_from_params(mu, Sigma)
    z0 = ch.assignments(0, 1)[0]
    for c in range(1, C):
        ch.set_state(c, z0, [0, 1, 2], mu, Sigma)
    item = int(np.argmin(np.abs(X - np.array([0.9, 0.6])).sum(1)))
    ch.update_item(item)
    z = ch.assignments()
    assert np.array_equal(np.delete(z, item, axis=1), np.tile(np.delete(z0, item), (C, 1)))   # only that item may move
    for c in (0, 1, C - 1):
        slots, counts, _, _ = ch.params(c)
        assert counts.sum() == ds.N and np.array_equal(np.bincount(z[c], minlength=32)[slots], counts)
    n = np.bincount(np.delete(z0, item), minlength=3).astype(float)
    w = np.exp(oracle.mvn_logpdf_batch(mu, Sigma, X[item:item + 1])[0]) * n
    p = w / w.sum()
    chosen = z[:, item]
    born = chosen > 2
    assert born.mean() < 0.02            # the reference prior's auxiliary draws are far from the data
    freq = np.bincount(chosen[~born], minlength=3) / (~born).sum()
    assert p[:2].min() > 0.2 and np.all(np.abs(freq - p) < 4 * np.sqrt(p * (1 - p) / (~born).sum()) + 0.005), (freq, p)
    # a second call draws afresh
    ch.update_item(item)
    assert not np.array_equal(ch.assignments()[:, item], chosen)
    ch.close()
    ds.close()
