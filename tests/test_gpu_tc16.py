"""GPU (pytest -m gpu): D = 16 on the kernel pair of npb_alg8_gemm.cu (NPB_D16_PATH=tc2): FP16x3 density tables + race kernel
(round 1's tensor path; the handle's default now also uses the fused kernel of npb_alg8_fused16.cu, tests/test_gpu_fused16.py)."""
import os

import numpy as np
import pytest

from noparama_b200 import synthetic as syn
from test_gpu_tile import invariants

pytestmark = pytest.mark.gpu
D = 16


@pytest.fixture
def env():
    keys = ("NPB_D16_PATH", "NPB_D16_BLOCK", "NPB_D16_EPI", "NPB_D64_SPEC", "NPB_D16_AUX")
    saved = {k: os.environ.get(k) for k in keys}
    os.environ["NPB_D16_PATH"] = "tc2"
    yield os.environ
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("epi", ["8", "16"])
def test_tc16_density_table_within_1e5_of_oracle(npb, ctx, oracle, env, epi):
    env["NPB_D16_EPI"] = epi
    rng = np.random.default_rng(16)
    K = 32
    X, y = syn.gmm(4000, D, 8, 216)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(D)).bind(ctx)
    ch = npb.Chains(ctx, ds, 3, Kmax=32, K0=8, seed=1)
    mu = X[rng.integers(0, len(X), K)] + 0.5 * rng.standard_normal((K, D))
    B = rng.standard_normal((K, D, D)) / np.sqrt(D)
    Sigma = B @ np.transpose(B, (0, 2, 1)) + 0.3 * np.eye(D)
    ch.init_from_params(mu, Sigma)
    for chain in (0, 2):
        items = rng.integers(0, len(X), 32)
        got = ch.probe_tile_logdensity(chain, items).astype(np.float64)
        slots, counts, _, _ = ch.params(chain)
        want = oracle.mvn_logpdf_batch(mu, Sigma, X[items]).T
        occ = np.zeros(32, bool)
        occ[slots] = True
        assert occ.sum() >= 30 and np.all(np.isnan(got[~occ]))
        err = np.abs(got[occ] - want[occ]) / np.maximum(1.0, np.abs(want[occ]))
        print("epilogue warps", epi, "max relative error of the table", err.max())
        assert err.max() < 1e-5, err.max()
    ch.close()
    ds.close()


@pytest.mark.parametrize("block", [128, 1024])
def test_tc16_invariants(npb, ctx, oracle, env, block):
    env["NPB_D16_BLOCK"] = str(block)
    N = 1000 + 13
    X, y = syn.gmm(N, D, 4, 116)
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=24, Kmax=32, K0=8, seed=D)
    births = 0
    for _ in range(3):
        st = mc.chains.sweep(npb.ALG8, 3)
        births += st.new_clusters
        assert st.overflow_chains == 0 and st.reassignments == 24 * N * 3
        assert 4 * st.reassignments <= st.candidates <= 35 * st.reassignments
    z = mc.getMembershipMatrix()
    m = mc.chains.metrics(y)
    for c in range(0, 24, 5):
        k = invariants(mc.chains, z[c], c, N)
        assert k == m["K"][c]
        want = oracle.metrics(y, z[c])
        assert np.allclose([m["purity"][c], m["rand_index"][c], m["adjusted_rand"][c]], want, atol=1e-12)
    print("births", births, "mean K", st.mean_K)
    zs = []
    # speculation on/off, batching of sweeps, auxiliary keys lazily in the race, from the k_aux_keys pre-pass, or bounded per 32-step
    # group by k_aux_bound and drawn on demand (the default): same result
    for spec, per_launch, aux in (("1", None, "lazy"), ("0", None, "lazy"), ("1", 1, "lazy"), ("1", None, "pre"), ("0", 2, "pre"),
                                  ("1", None, "bound"), ("0", 3, "bound")):
        env["NPB_D64_SPEC"], env["NPB_D16_AUX"] = spec, aux
        a = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=6, Kmax=32, K0=8, seed=77)
        a.run(4, sweeps_per_launch=per_launch)
        zs.append(a.getMembershipMatrix().copy())
    assert all(np.array_equal(zs[0], z) for z in zs[1:])
    ds.close()


def test_tc16_recovers_given_clusters(npb, ctx, env):
    X, y = syn.gmm(4000, D, 8, 5)
    means = np.stack([X[y == k].mean(0) for k in range(8)])
    Sigma = np.tile(np.eye(D), (8, 1, 1))
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=32, Kmax=32, seed=3)
    mc.chains.init_from_params(means, Sigma)
    s1 = mc.chains.sweep(npb.ALG8, 1)
    s2 = mc.chains.sweep(npb.ALG8, 1)
    assert s1.overflow_chains == 0 and s1.candidates == 11 * s1.reassignments == s2.candidates
    assert s1.moved > 0.8 * s1.reassignments and s2.moved < 1e-3 * s2.reassignments
    m = mc.chains.metrics(y)
    assert np.all(m["K"] == 8) and m["purity"].min() > 0.999
    ds.close()


def test_fp32_pipe_kernel_still_serves_16d(npb, ctx, oracle, env):
    """NPB_D16_PATH=fp32: k_alg8_sweep_tile4 at D = 16 (the default at D = 4, 8): recovery, candidates, and its producer's
    density tile against the oracle -- the same checks the tensor path passes above."""
    env["NPB_D16_PATH"] = "fp32"
    X, y = syn.gmm(4000, D, 8, 5)
    means = np.stack([X[y == k].mean(0) for k in range(8)])
    Sigma = np.tile(np.eye(D), (8, 1, 1))
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=32, Kmax=32, seed=3)
    mc.chains.init_from_params(means, Sigma)
    s1 = mc.chains.sweep(npb.ALG8, 1)
    s2 = mc.chains.sweep(npb.ALG8, 1)
    assert s1.overflow_chains == 0 and s1.candidates == 11 * s1.reassignments == s2.candidates
    assert s2.moved < 1e-3 * s2.reassignments
    m = mc.chains.metrics(y)
    assert np.all(m["K"] == 8) and m["purity"].min() > 0.999
    items = np.arange(32, dtype=np.int32) * 100
    got = mc.chains.probe_tile_logdensity(1, items).astype(np.float64)[:8]
    want = oracle.mvn_logpdf_batch(means, Sigma, X[items]).T
    assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) < 1e-5
    ds.close()


def _oracle_seed16(args):
    seed, T, K0 = args
    from oracle import binding as orc
    X, y = syn.gmm(160, D, 3, 43, min_dist=6.0)
    p = orc.make_prior(**syn.reference_prior(D))
    r = orc.Run(p, X, T=T, K0=K0, seed_main=300 + seed, seed_shuffle=2900 + seed, flags=orc.LOG_DOMAIN)
    s = r.stats()
    pur, ri, ari = orc.metrics(y, r.assignments(0))
    return s.K_final, pur, ari, s.new_cluster_events / s.updates, s.moved / s.updates


def test_tc16_distribution_and_birth_rate_against_oracle(npb, ctx, oracle, env):
    """The tensor path's race (counter-hash noise capped at 20, speculative step-parallel pass, births repaired in the
    density table) against the oracle (log-domain mode) at D = 16: 96 oracle seeds against 768 device chains from the same
    K0 = 8 start, T = 60 sweeps: K, purity, ARI in distribution, birth and move rates over the whole run."""
    from multiprocessing import Pool
    from scipy import stats as sps
    env["NPB_D16_BLOCK"] = "128"
    X, y = syn.gmm(160, D, 3, 43, min_dist=6.0)
    T, K0 = 60, 8
    with Pool(8) as pool:
        res = np.array(pool.map(_oracle_seed16, [(s, T, K0) for s in range(96)]))
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=768, Kmax=32, K0=K0, seed=17)
    stats = mc.run(T, sweeps_per_launch=20)
    assert all(s.overflow_chains == 0 for s in stats)
    m = mc.chains.metrics(y)
    for name, got, want in (("K", m["K"].astype(float), res[:, 0]), ("purity", m["purity"], res[:, 1]),
                            ("ari", m["adjusted_rand"], res[:, 2])):
        p = sps.ks_2samp(got, want).pvalue
        assert p > 0.01, "%s: KS p=%.2e (gpu %.4f vs oracle %.4f)" % (name, p, got.mean(), want.mean())
    n = sum(s.reassignments for s in stats)
    births = sum(s.new_clusters for s in stats) / n
    moved = sum(s.moved for s in stats) / n
    print("births/step gpu %.5f oracle %.5f; moved gpu %.4f oracle %.4f; K gpu %.2f oracle %.2f" % (
        births, res[:, 3].mean(), moved, res[:, 4].mean(), m["K"].mean(), res[:, 0].mean()))
    assert abs(births - res[:, 3].mean()) < 0.1 * res[:, 3].mean() + 2e-5
    assert abs(moved - res[:, 4].mean()) < 0.1 * res[:, 4].mean() + 2e-4
    ds.close()
