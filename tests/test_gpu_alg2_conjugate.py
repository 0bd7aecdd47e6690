"""GPU (pytest -m gpu): CONJUGATE Algorithm 2 (npb_alg2.cu; BASELINE configs[3]) against oracle/np_oracle_alg2.inc.  The
reference has no executable conjugate path (np_neal_algorithm2.cpp:32-120 is dead code): parity is against the textbook model
the oracle restates and pins to scipy.stats.multivariate_t -- UNPINNED against the reference by construction.
  (1) posterior-predictive log-densities within 1e-5 relative of the oracle at D = 2, 16, 64 (the tolerance north_star states);
  (2) the sufficient statistics after sweeps (rank-1 insert / remove per move) against a recount from the assignments: counts
      exact, sum x and sum x x^T to 1e-10 relative;
  (3) 256 chains against 256 oracle seeds: K, purity, ARI in distribution, move rate."""
import numpy as np
import pytest
from scipy import stats as sps

from noparama_b200 import synthetic as syn
from noparama_b200.api import NpbError

pytestmark = pytest.mark.gpu


def conj_prior(X):
    D = X.shape[1]
    return dict(mu0=X.mean(0), kappa=0.01, nu=D + 2.0, Lambda=np.eye(D), alpha=1.0)


@pytest.mark.parametrize("D,N", [(2, 400), (16, 600), (64, 900)])
def test_predictive_logdensity_within_1e5_of_oracle(npb, ctx, oracle, D, N):
    X, y = syn.gmm(N, D, 4, 300 + D)
    pr = conj_prior(X)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**pr).bind(ctx)
    ch = npb.Chains(ctx, ds, 3, Kmax=32, K0=4, seed=5)
    rng = np.random.default_rng(D)
    z = rng.integers(0, 5, N).astype(np.int32)  # five clusters of mixed content, slots 1, 3, 4, 7, 30
    slots = np.array([1, 3, 4, 7, 30], np.int32)
    ch.set_state(1, slots[z], slots, np.zeros((5, D)), np.tile(np.eye(D), (5, 1, 1)))
    items = rng.integers(0, N, 40).astype(np.int32)
    got = ch.alg2_logpred(1, items).astype(np.float64)
    P = oracle.make_prior(**pr)
    worst = 0.0
    for j, it in enumerate(items):
        for k, s in enumerate(slots):
            want = oracle.niw_logpred(P, X[z == k], X[it])
            worst = max(worst, abs(got[j, s] - want) / max(1.0, abs(want)))
        want0 = oracle.niw_logpred(P, X[:0], X[it])
        worst = max(worst, abs(got[j, 32] - want0) / max(1.0, abs(want0)))
        occ = np.zeros(32, bool)
        occ[slots] = True
        assert np.all(np.isnan(got[j, :32][~occ]))
    print("D = %d: max relative error of the predictive log-density %.2e" % (D, worst))
    assert worst < 1e-5
    ch.close()
    ds.close()


@pytest.mark.parametrize("D", [2, 16, 64])
def test_suffstats_follow_the_moves(npb, ctx, D):
    N = 700 if D < 64 else 300
    X, y = syn.gmm(N, D, 3, 500 + D, min_dist=3.0)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**conj_prior(X)).bind(ctx)
    ch = npb.Chains(ctx, ds, 5, Kmax=32, K0=12, seed=8)
    tot_moved = 0
    for _ in range(3):
        st = ch.sweep(npb.ALG2_CONJUGATE, 2)
        assert st.overflow_chains == 0 and st.reassignments == 5 * N * 2
        tot_moved += st.moved
    assert tot_moved > N  # the chains did move items: the incremental updates were exercised
    z = ch.assignments()
    for c in (0, 4):
        n, sx, sxx = ch.alg2_suffstats(c)
        assert np.array_equal(n, np.bincount(z[c], minlength=32))
        for k in np.nonzero(n)[0]:
            M = X[z[c] == k]
            assert np.allclose(sx[k], M.sum(0), rtol=1e-10, atol=1e-9)
            assert np.allclose(sxx[k], M.T @ M, rtol=1e-10, atol=1e-8)
    ch.close()
    ds.close()


def _problem(which):
    if which == "cfg1":
        return syn.config(1)
    D, N = which
    return syn.gmm(N, D, 3, 900 + D, min_dist=5.0)


def _oracle_seed(args):
    which, seed, T, K0 = args
    from oracle import binding as orc
    X, y = _problem(which)
    P = orc.make_prior(**conj_prior(X))
    z, Kt, moved, births = orc.alg2_run(P, X, T, K0, 1000 + seed)
    pur, ri, ari = orc.metrics(y, z)
    return Kt[-1], pur, ari, moved / (T * len(X)), births / (T * len(X))


@pytest.mark.parametrize("which,T,seeds", [("cfg1", 60, 256), ((16, 240), 30, 128), ((64, 150), 12, 64)])
def test_distribution_against_oracle(npb, ctx, oracle, which, T, seeds):
    """config 1 (2-D, one lane per slot), a 16-D problem (two lanes per slot, P in registers) and a 64-D one (a warp per slot, P
    read from L2): the device chains against independent oracle runs"""
    from multiprocessing import Pool
    K0 = 4
    with Pool(8) as pool:
        res = np.array(pool.map(_oracle_seed, [(which, s, T, K0) for s in range(seeds)]))
    X, y = _problem(which)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**conj_prior(X)).bind(ctx)
    ch = npb.Chains(ctx, ds, 256, Kmax=32, K0=K0, seed=77)
    moved = births = 0
    for _ in range(T // 6):
        st = ch.sweep(npb.ALG2_CONJUGATE, 6)
        assert st.overflow_chains == 0
        moved += st.moved
        births += st.new_clusters
    m = ch.metrics(y)
    for name, got, want in (("K", m["K"].astype(float), res[:, 0]), ("purity", m["purity"], res[:, 1]), ("ari", m["adjusted_rand"], res[:, 2])):
        p = sps.ks_2samp(got, want).pvalue
        print("%s: gpu %.4f oracle %.4f KS p = %.3f" % (name, got.mean(), want.mean(), p))
        assert p > 0.01, name
    mr, br = moved / (256 * T * len(X)), births / (256 * T * len(X))
    print("moved per step gpu %.4f oracle %.4f; births per step gpu %.5f oracle %.5f" % (mr, res[:, 3].mean(), br, res[:, 4].mean()))
    assert abs(mr - res[:, 3].mean()) < 0.1 * res[:, 3].mean() + 1e-3
    assert abs(br - res[:, 4].mean()) < 0.15 * res[:, 4].mean() + 2e-4
    ch.close()
    ds.close()


@pytest.mark.parametrize("D,N,K0,tc", [(16, 900, 10, 0), (16, 900, 10, 1), (64, 500, 8, 1), (64, 500, 8, 0)])
def test_tile_length_does_not_change_the_chain(npb, ctx, D, N, K0, tc):
    """k_a2_tile (npb_alg2_tile.cu) evaluates up to 64 steps ahead of the chain and corrects the two changed columns after every
    move: the assignments, counts and FP64 statistics must equal those of the strictly sequential schedule (a2_tile = 1) bit for
    bit, from a start where most items move (K0 random clusters) to the settled chain"""
    X, y = syn.gmm(N, D, 4, 700 + D, min_dist=3.0)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**conj_prior(X)).bind(ctx)
    res = {}
    for tile in (1, 8, 64, 128):
        ch = npb.Chains(ctx, ds, 7, Kmax=32, K0=K0, seed=21)
        ch.set_option("a2_tile", str(tile))
        ch.set_option("a2_tc", str(tc))  # D = 64: the tcgen05 kernel k_a2_tc (tiles of up to 128 steps) or the FP32 tile kernel (up to 64)
        ch.set_option("a2_tc16", str(tc))  # D = 16: k_a2_tc16 or the FP32 tile kernel
        moved = births = 0
        zs = []
        for _ in range(3):
            st = ch.sweep(npb.ALG2_CONJUGATE, 2)
            assert st.overflow_chains == 0
            moved += st.moved
            births += st.new_clusters
            zs.append(ch.assignments().copy())
        res[tile] = (zs, moved, births, st.candidates, [ch.alg2_suffstats(c) for c in (0, 6)])
        ch.close()
    assert res[1][1] > 2 * N  # items moved: the correction path ran
    for tile in (8, 64, 128):
        for a, b in zip(res[1][0], res[tile][0]):
            assert np.array_equal(a, b), tile
        assert res[1][1:4] == res[tile][1:4], tile
        for (n0, sx0, sxx0), (n1, sx1, sxx1) in zip(res[1][4], res[tile][4]):
            assert np.array_equal(n0, n1) and np.array_equal(sx0[n0 > 0], sx1[n1 > 0]) and np.array_equal(sxx0[n0 > 0], sxx1[n1 > 0])
    print("D = %d: tiles of 1, 8, 64, 128 steps give the same chain (%d moves, %d births over 6 sweeps of 7 chains)" % (D, res[1][1], res[1][2]))
    ds.close()


@pytest.mark.parametrize("D,N", [(16, 600), (64, 300)])
def test_step_at_a_time_kernel_still_follows_the_moves(npb, ctx, D, N):
    """the round-2 first path (k_a2_sweep, option a2_tile = 0) stays available for A/B measurements: statistics against a recount"""
    X, y = syn.gmm(N, D, 3, 500 + D, min_dist=3.0)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**conj_prior(X)).bind(ctx)
    ch = npb.Chains(ctx, ds, 4, Kmax=32, K0=12, seed=8)
    ch.set_option("a2_tile", "0")
    st = ch.sweep(npb.ALG2_CONJUGATE, 3)
    assert st.overflow_chains == 0 and st.moved > N
    z = ch.assignments()
    n, sx, sxx = ch.alg2_suffstats(2)
    assert np.array_equal(n, np.bincount(z[2], minlength=32))
    for k in np.nonzero(n)[0]:
        M = X[z[2] == k]
        assert np.allclose(sx[k], M.sum(0), rtol=1e-10, atol=1e-9) and np.allclose(sxx[k], M.T @ M, rtol=1e-10, atol=1e-8)
    ch.close()
    ds.close()


@pytest.mark.parametrize("D,N,tc", [(16, 800, 0), (16, 800, 1), (64, 500, 1), (64, 500, 0)])
def test_tile_kernel_decides_like_the_step_at_a_time_kernel(npb, ctx, D, N, tc):
    """k_a2_tile's quadratic forms (register-blocked products) against k_a2_sweep's row dots: the two
    kernels draw the same race noise, so from the same state one sweep must make the same decisions except where two keys lie within
    the FP32 rounding of the forms (~1e-6 relative); an error of 1e-4 in a form would flip a visible share of the steps"""
    X, y = syn.gmm(N, D, 4, 800 + D, min_dist=3.0)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**conj_prior(X)).bind(ctx)
    base = npb.Chains(ctx, ds, 6, Kmax=32, K0=6, seed=31)
    base.sweep(npb.ALG2_CONJUGATE, 4 if D == 16 else 1)  # past the first reshuffle, still moving
    zs = base.assignments()
    out = {}
    for tile in (0, 64):
        ch = npb.Chains(ctx, ds, 6, Kmax=32, K0=6, seed=31)
        ch.set_option("a2_tile", str(tile))
        ch.set_option("a2_tc", str(tc))
        ch.set_option("a2_tc16", str(tc))
        for c in range(6):
            slots = np.unique(zs[c]).astype(np.int32)
            ch.set_state(c, zs[c], slots, np.zeros((len(slots), D)), np.tile(np.eye(D), (len(slots), 1, 1)))
        st = ch.sweep(npb.ALG2_CONJUGATE, 1)
        out[tile] = (ch.assignments().copy(), st.moved)
        ch.close()
    same = (out[0][0] == out[64][0]).mean()
    print("D = %d: %.5f of the assignments equal after one sweep; moves %d (step at a time) / %d (tile)" % (D, same, out[0][1], out[64][1]))
    assert out[0][1] > 0
    assert same > 0.99
    base.close()
    ds.close()


@pytest.mark.parametrize("D,N", [(16, 400), (64, 400)])
def test_full_slot_table_is_handled_alike_by_every_tile_length(npb, ctx, D, N):
    """all 32 slots occupied at the start (K0 = 32): a step that draws a new cluster while no slot is free keeps its item and reports the
    chain (np_error_t's overflow), slots that empty are taken again by later births -- same chain for tiles of 1 and 128 steps, counts
    consistent with the assignments"""
    X, y = syn.gmm(N, D, 4, 900 + D, min_dist=2.0)
    ds = npb.Dataset(ctx, X)
    pr = conj_prior(X)
    pr["alpha"] = 50.0  # births are frequent
    npb.NormalInverseWishart(**pr).bind(ctx)
    res = {}
    for tile in (1, 128):
        ch = npb.Chains(ctx, ds, 5, Kmax=32, K0=32, seed=3)
        ch.set_option("a2_tile", str(tile))
        full = 0
        for _ in range(2):
            try:
                ch.sweep(npb.ALG2_CONJUGATE, 1)
            except NpbError as e:  # NPB_E_KMAX_OVERFLOW: reported, the chains stay valid
                assert "Kmax" in str(e)
                full += 1
        z = ch.assignments().copy()
        n, _, _ = ch.alg2_suffstats(2)
        assert np.array_equal(n, np.bincount(z[2], minlength=32))
        res[tile] = (z, full, ch.metrics(y)["K"].copy())
        ch.close()
    assert np.array_equal(res[1][0], res[128][0]) and res[1][1] == res[128][1] and np.array_equal(res[1][2], res[128][2])
    print("D = %d: %d of 2 sweeps met a full slot table; K after: %s" % (D, res[1][1], res[1][2]))
    ds.close()


def test_tensor_core_forms_hold_on_shifted_wide_data(npb, ctx):
    """k_a2_tc centres the items on the data mean and splits FP16 x 3: data far from the origin (+500) with widely separated components
    (x 3) must still make the step-at-a-time FP32 kernel's decisions from the same state"""
    D, N = 64, 500
    X, y = syn.gmm(N, D, 4, 864, min_dist=3.0)
    X = 3.0 * X + 500.0
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**conj_prior(X)).bind(ctx)
    base = npb.Chains(ctx, ds, 6, Kmax=32, K0=6, seed=31)
    base.sweep(npb.ALG2_CONJUGATE, 1)
    zs = base.assignments()
    out = {}
    for tile in (0, 128):
        ch = npb.Chains(ctx, ds, 6, Kmax=32, K0=6, seed=31)
        ch.set_option("a2_tile", str(tile))
        for c in range(6):
            slots = np.unique(zs[c]).astype(np.int32)
            ch.set_state(c, zs[c], slots, np.zeros((len(slots), D)), np.tile(np.eye(D), (len(slots), 1, 1)))
        st = ch.sweep(npb.ALG2_CONJUGATE, 1)
        out[tile] = (ch.assignments().copy(), st.moved)
        ch.close()
    same = (out[0][0] == out[128][0]).mean()
    print("shifted, wide 64-D data: %.5f of the assignments equal after one sweep; moves %d / %d" % (same, out[0][1], out[128][1]))
    assert out[0][1] > 0 and same > 0.99
    base.close()
    ds.close()
