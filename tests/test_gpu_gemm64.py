"""GPU (pytest -m gpu): the D = 64 sweep path (npb_alg8_gemm.cu): tcgen05 FP16x3 density table + warp-per-chain race."""
import os

import numpy as np
import pytest

from noparama_b200 import synthetic as syn
from test_gpu_tile import invariants

pytestmark = pytest.mark.gpu
D = 64


@pytest.fixture
def env():
    saved = {k: os.environ.get(k) for k in ("NPB_D64_BLOCK", "NPB_D64_DENSITY", "NPB_D64_SPEC", "NPB_D64_OVERLAP")}
    yield os.environ
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _full_cov_params(rng, X, K):
    mu = X[rng.integers(0, len(X), K)] + 0.5 * rng.standard_normal((K, D))
    B = rng.standard_normal((K, D, D)) / np.sqrt(D)
    Sigma = B @ np.transpose(B, (0, 2, 1)) + 0.3 * np.eye(D)
    return mu, Sigma


@pytest.mark.parametrize("density", ["tc", "fp32"])
def test_gemm64_density_table_within_1e5_of_oracle(npb, ctx, oracle, env, density):
    """The log-density table the D = 64 sweep reads (tcgen05 kind::f16, three FP16 products per FP32 product, operands
    centred on the dataset mean and scaled by powers of two; or the FP32 kernel) against the oracle's double-precision density: 1e-5 relative
    (north_star tolerance), full covariances, near and far clusters."""
    env["NPB_D64_DENSITY"] = density
    rng = np.random.default_rng(64)
    K = 32
    X, y = syn.gmm(3000, D, 8, 264)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(D)).bind(ctx)
    ch = npb.Chains(ctx, ds, 3, Kmax=32, K0=8, seed=1)
    mu, Sigma = _full_cov_params(rng, X, K)
    ch.init_from_params(mu, Sigma)
    for chain in (0, 2):
        items = rng.integers(0, len(X), 32)
        got = ch.probe_tile_logdensity(chain, items).astype(np.float64)  # [slot, item]
        slots, counts, _, _ = ch.params(chain)
        want = oracle.mvn_logpdf_batch(mu, Sigma, X[items]).T
        occ = np.zeros(32, bool)
        occ[slots] = True
        assert occ.sum() >= 30 and np.all(np.isnan(got[~occ]))
        err = np.abs(got[occ] - want[occ]) / np.maximum(1.0, np.abs(want[occ]))
        print(density, "max relative error of the table", err.max())
        assert err.max() < 1e-5, (density, err.max())
    ch.close()
    ds.close()


@pytest.mark.parametrize("block", [128, 256, 4096])
def test_gemm64_invariants(npb, ctx, oracle, env, block):
    """Reference-style start (K0 prior draws, uniform assignment) at D = 64: bookkeeping invariants after every launch,
    metrics against the oracle, candidates per step, determinism across launch splitting -- over blocks of steps that
    do and do not divide N, with births inside a block."""
    env["NPB_D64_BLOCK"] = str(block)
    N = 700 + 13
    X, y = syn.gmm(N, D, 4, 164)
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=24, Kmax=32, K0=8, seed=D)
    st = None
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
    assert abs(st.mean_K - m["K"].mean()) < 1e-9
    print("births", births, "mean K", st.mean_K)
    a = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=6, Kmax=32, K0=8, seed=77)
    b = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=6, Kmax=32, K0=8, seed=77)
    a.run(4)
    b.run(4, sweeps_per_launch=1)
    assert np.array_equal(a.getMembershipMatrix(), b.getMembershipMatrix())
    ds.close()


@pytest.mark.parametrize("m_aux", [3, 1])
def test_gemm64_recovers_given_clusters(npb, ctx, env, m_aux):
    """Config-4 regime (Algorithm 8 with m = 3, Algorithm 2 with one auxiliary draw): chains start from K_true known
    clusters and a random assignment; sweeps with frozen parameters must put every item into its own component.  The
    tensor-core table and the FP32 table must lead the same chains (same seed) to the same assignments."""
    env["NPB_D64_BLOCK"] = "1024"
    K = 8
    X, y = syn.gmm(5000, D, K, 5)
    means = np.stack([X[y == k].mean(0) for k in range(K)])
    Sigma = np.tile(np.eye(D), (K, 1, 1))
    ds = npb.Dataset(ctx, X)
    zs = {}
    for density in ("tc", "fp32"):
        env["NPB_D64_DENSITY"] = density
        sampler = npb.NealAlgorithm8 if m_aux == 3 else npb.NealAlgorithm2
        mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), sampler, chains=20, Kmax=32, m_aux=m_aux, seed=3)
        mc.chains.init_from_params(means, Sigma)
        m0 = mc.chains.metrics(y)
        assert m0["purity"].mean() < 0.3 and np.all(m0["K"] == K)
        s1 = mc.chains.sweep(npb.ALG8 if m_aux == 3 else npb.ALG2, 1)
        s2 = mc.chains.sweep(npb.ALG8 if m_aux == 3 else npb.ALG2, 1)
        assert s1.overflow_chains == 0 and s1.reassignments == 20 * 5000 == s2.reassignments
        assert s1.candidates == (K + m_aux) * s1.reassignments == s2.candidates
        assert s1.moved > 0.8 * s1.reassignments and s2.moved < 1e-3 * s2.reassignments
        z = mc.getMembershipMatrix()
        for c in (0, 19):
            invariants(mc.chains, z[c], c, ds.N)
        m = mc.chains.metrics(y)
        assert np.all(m["K"] == K) and m["purity"].min() > 0.9999 and m["adjusted_rand"].min() > 0.9999
        zs[density] = z
    assert (zs["tc"] == zs["fp32"]).mean() > 0.9999
    ds.close()


def test_gemm64_schedule_does_not_change_the_result(npb, ctx, env):
    """The step-parallel speculative pass, the sequential pass, the two-stream overlap and the batching of sweeps into launches are schedules of
    the same computation: same seed, same assignments, bit for bit -- from a reference-style start (births, deaths, many
    moves) into the converged regime (tiles without a move)."""
    X, y = syn.gmm(3000, D, 6, 77)
    ds = npb.Dataset(ctx, X)
    zs = []
    env["NPB_D64_BLOCK"] = "512"
    for spec, overlap, per_launch in (("1", "1", None), ("0", "1", None), ("1", "0", None), ("1", "1", 1)):
        env["NPB_D64_SPEC"], env["NPB_D64_OVERLAP"] = spec, overlap
        mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=10, Kmax=32, K0=8, seed=11)
        sts = mc.run(6, sweeps_per_launch=per_launch)
        assert all(st.overflow_chains == 0 for st in sts)
        zs.append((mc.getMembershipMatrix().copy(), sum(st.candidates for st in sts), sum(st.moved for st in sts),
                   sum(st.new_clusters for st in sts)))
    print("moved", zs[0][2], "births", zs[0][3])
    for z, cand, moved, born in zs[1:]:
        assert np.array_equal(z, zs[0][0]) and (cand, moved, born) == zs[0][1:]
    ds.close()
