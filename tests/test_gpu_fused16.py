"""GPU (pytest -m gpu): k_sweep_tc16 (npb_alg8_fused16.cu), the fused D = 16 sweep kernel -- tcgen05 densities, epilogue and
race in one CTA -- against the round-1 kernel pair (NPB_D16_PATH=tc2: k_density_tc16 + k_race, same keys, same noise), with and
without its speculation, and its bookkeeping with births, deaths, ragged blocks and odd chain counts."""
import os

import numpy as np
import pytest

from noparama_b200 import synthetic as syn
from test_gpu_tile import invariants

pytestmark = pytest.mark.gpu
D = 16
KEYS = ("NPB_D16_PATH", "NPB_D16_BLOCK", "NPB_D64_SPEC", "NPB_D16_AUX")


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


def overlapping(N, K, seed, dist=2.5):
    """K unit-covariance components on a line, neighbours `dist` apart: every item has one or two real contenders."""
    rng = np.random.default_rng(seed)
    means = np.zeros((K, D))
    means[:, 0] = dist * np.arange(K)
    means[:, 1] = 0.5 * dist * (np.arange(K) % 2)
    y = rng.integers(0, K, N).astype(np.int32)
    X = means[y] + rng.standard_normal((N, D)) + 6.0
    return X, y, means + 6.0


def run(npb, ctx, ds, means, chains, sweeps, seed, per_launch=None):
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=chains, Kmax=32, seed=seed)
    mc.chains.init_from_params(means, np.tile(np.eye(D), (len(means), 1, 1)))
    stats = mc.run(sweeps, sweeps_per_launch=per_launch)
    z = mc.getMembershipMatrix().copy()
    tot = dict(moved=sum(s.moved for s in stats), births=sum(s.new_clusters for s in stats),
               cand=sum(s.candidates for s in stats), n=sum(s.reassignments for s in stats))
    return mc, z, tot


@pytest.mark.parametrize("chains,N,block", [(7, 1000 + 13, 256), (32, 3000, 1024)])
def test_fused_equals_kernel_pair_and_sequential(npb, ctx, env, chains, N, block):
    """Mixing regime without births (overlapping given clusters, births are hopeless under the diffuse prior): the fused
    kernel, the fused kernel without speculation, and the round-1 kernel pair produce the same assignments bit for bit."""
    env["NPB_D16_BLOCK"] = str(block)
    X, y, means = overlapping(N, 6, 5)
    ds = npb.Dataset(ctx, X)
    out = {}
    for name, path, spec in (("fused", None, "1"), ("fused_seq", None, "0"), ("pair", "tc2", "1")):
        env["NPB_D16_PATH"] = path or "tc"  # (unset = auto: by the moved fraction of the last sweep; here each path on its own)
        env["NPB_D64_SPEC"] = spec
        mc, z, tot = run(npb, ctx, ds, means, chains, 3, seed=11)
        out[name] = (z, tot)
        for c in range(0, chains, 3):
            invariants(mc.chains, z[c], c, N)
    z0, t0 = out["fused"]
    print("moved fraction %.3f births %d candidates/step %.2f" % (t0["moved"] / t0["n"], t0["births"], t0["cand"] / t0["n"]))
    assert t0["moved"] > 0.05 * t0["n"]
    for name in ("fused_seq", "pair"):
        z1, t1 = out[name]
        assert np.array_equal(z0, z1), name
        assert (t0["moved"], t0["births"], t0["cand"]) == (t1["moved"], t1["births"], t1["cand"]), name
    ds.close()


def test_fused_births_deaths_and_batching(npb, ctx, oracle, env):
    """Reference regime (K0 = 8 prior clusters, births and deaths): invariants, metric parity, and identical results with
    the speculation on or off and however the sweeps are batched into launches."""
    env["NPB_D16_BLOCK"] = "128"
    env["NPB_D16_PATH"] = "tc"
    N = 1000 + 13
    X, y = syn.gmm(N, D, 4, 116)
    ds = npb.Dataset(ctx, X)
    mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=25, Kmax=32, K0=8, seed=D)
    births = 0
    for _ in range(3):
        st = mc.chains.sweep(npb.ALG8, 3)
        births += st.new_clusters
        assert st.overflow_chains == 0 and st.reassignments == 25 * N * 3
        assert 4 * st.reassignments <= st.candidates <= 35 * st.reassignments
    assert births > 0
    z = mc.getMembershipMatrix()
    m = mc.chains.metrics(y)
    for c in range(0, 25, 4):
        k = invariants(mc.chains, z[c], c, N)
        assert k == m["K"][c]
        want = oracle.metrics(y, z[c])
        assert np.allclose([m["purity"][c], m["rand_index"][c], m["adjusted_rand"][c]], want, atol=1e-12)
    zs = []
    for spec, per_launch in (("1", None), ("0", None), ("1", 1), ("0", 2)):
        env["NPB_D64_SPEC"] = spec
        a = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=6, Kmax=32, K0=8, seed=77)
        st = a.run(4, sweeps_per_launch=per_launch)
        zs.append((a.getMembershipMatrix().copy(), sum(s.new_clusters for s in st), sum(s.moved for s in st)))
    print("births", births, [(b, mv) for _, b, mv in zs])
    assert all(np.array_equal(zs[0][0], z) and (b, mv) == zs[0][1:] for z, b, mv in zs[1:])
    ds.close()


def test_fused_probe_unfolded_epilogue(npb, ctx, oracle, env):
    """The density table of the fused kernel against the oracle, including slots whose mean lies far outside the data (a
    prior-born cluster): their offsets do not fit the folded FP16 columns and take the long epilogue."""
    env["NPB_D16_PATH"] = "tc"
    rng = np.random.default_rng(3)
    X, y = syn.gmm(4000, D, 8, 216)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(D)).bind(ctx)
    ch = npb.Chains(ctx, ds, 3, Kmax=32, K0=8, seed=1)
    K = 32
    mu = X[rng.integers(0, len(X), K)] + 0.5 * rng.standard_normal((K, D))
    far = np.arange(K) % 3 == 0
    mu[far] = 6.0 + 400.0 * rng.standard_normal((far.sum(), D))  # prior-scale means (std of mu under G0 ~ 2.2 |v| per dimension)
    B = rng.standard_normal((K, D, D)) / np.sqrt(D)
    Sigma = B @ np.transpose(B, (0, 2, 1)) + 0.3 * np.eye(D)
    Sigma[far] *= 50.0
    ch.init_from_params(mu, Sigma)
    items = rng.integers(0, len(X), 32)
    got = ch.probe_tile_logdensity(1, items).astype(np.float64)
    slots, counts, _, _ = ch.params(1)
    want = oracle.mvn_logpdf_batch(mu, Sigma, X[items]).T
    occ = np.zeros(32, bool)
    occ[slots] = True
    err = np.abs(got[occ] - want[occ]) / np.maximum(1.0, np.abs(want[occ]))
    print("max relative error: near slots %.2e, far slots %.2e" % (err[~far[occ]].max(), err[far[occ]].max()))
    assert occ.sum() >= 30 and err.max() < 1e-5
    ch.close()
    ds.close()


def test_auto_path_switches_without_changing_results(npb, ctx, env):
    """NPB_D16_PATH unset: the handle picks the fused kernel or the kernel pair by the moved fraction of its last sweep;
    without births the assignments are the same whichever it took."""
    X, y, means = overlapping(2000, 6, 9)
    ds = npb.Dataset(ctx, X)
    out = []
    for path in (None, "tc", "tc2"):
        if path:
            env["NPB_D16_PATH"] = path
        else:
            env.pop("NPB_D16_PATH", None)
        mc, z, tot = run(npb, ctx, ds, means, 9, 4, seed=5, per_launch=1)
        out.append((z, tot["moved"]))
    assert out[0][1] > 0 and all(np.array_equal(out[0][0], z) and mv == out[0][1] for z, mv in out[1:])
    ds.close()
