"""GPU: the split-merge samplers (npb_splitmerge.cu) against the oracle.

The oracle's Jain-Neal / triadic restatements (oracle/np_oracle_sm.inc) are pinned proposal by proposal to the
reference's own code (tests/test_ref_pin.py).  Here:
  * invariants of the device state after sweeps (counts == histogram of z, occupied count, no overflow);
  * the acceptance arithmetic of single proposals recomputed in double precision with the oracle's density
    (np_jain_neal_algorithm.cpp:243-296,339-392; np_triadic_algorithm.cpp:370-437,529-590);
  * the distribution of K / purity / Rand / ARI and the per-move-type acceptance rates over 256 chains against the
    committed 256-seed oracle fixture (tests/golden/oracle_cfg1_{jain_neal,triadic}_256seeds.npz).
"""
import math
import os

import numpy as np
import pytest
from scipy import stats as sps

from noparama_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MASK = (1 << 64) - 1


def picks_of(npb, seed, sweep, N, nsub, s=0):
    # k_scan_order3: one keyed permutation per subset position
    return [int(npb.scan_order((seed ^ (0x9E3779B97F4A7C15 * (j + 1))) & MASK, sweep, N)[s]) for j in range(nsub)]


def check_state(ch, z, c, N):
    slots, counts, _, _ = ch.params(c)
    assert counts.sum() == N
    assert np.array_equal(np.bincount(z, minlength=ch.Kmax)[slots], counts)
    assert set(np.unique(z).tolist()) == set(slots.tolist())


@pytest.mark.parametrize("sampler_name", ["jain_neal", "triadic"])
def test_split_merge_state_invariants(npb, ctx, sampler_name):
    sampler = {"jain_neal": npb.JAIN_NEAL, "triadic": npb.TRIADIC}[sampler_name]
    X, y = syn.config(1)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(2)).bind(ctx)
    ch = npb.Chains(ctx, ds, 96, Kmax=64, seed=3)
    total = 0
    for _ in range(5):
        st = ch.sweep(sampler, 4)
        assert st.overflow_chains == 0
        total += st.reassignments
        assert sum(st.sm_attempts) == st.reassignments  # every proposal made is one of the four move types
        z = ch.assignments()
        for c in (0, 17, 95):
            check_state(ch, z[c], c, ds.N)
        m = ch.metrics(y)
        assert np.array_equal(m["K"], [len(np.unique(z[c])) for c in range(96)])
    # subsets with a repeated item are skipped (np_mcmc.cpp:155-158): a little under N per sweep
    assert 0.97 * 96 * 200 * 20 < total < 96 * 200 * 20
    ch.close()
    ds.close()


def ref_logA(kind, alpha, lp, zb, za, picks, plan, det):
    """log acceptance ratio in double from the state before (zb), after (za, accepted moves only) and log-densities
    lp[item, slot]."""
    lg = math.lgamma
    la = math.log(alpha)
    if kind == "jn_split":
        cur, new = plan["src"][0], int(det["new_slot"])
        pool = np.flatnonzero(zb == cur)
        move = pool[za[pool] == new]
        n0, n1 = len(move), len(pool) - len(move)
        return la + lg(n0) + lg(n1) - lg(n0 + n1) + (lp[move, new] - lp[move, cur]).sum()
    if kind == "jn_merge":
        c0, c1 = plan["src"]
        m0 = np.flatnonzero(zb == c0)
        n0, n1 = len(m0), int((zb == c1).sum())
        return -(la + lg(n0) + lg(n1) - lg(n0 + n1)) + (lp[m0, c1] - lp[m0, c0]).sum()
    src, tgt = plan["src"], plan["tgt"]
    pool = np.flatnonzero(np.isin(zb, src))
    parts = [pool[za[pool] == t] for t in tgt] if len(tgt) > 1 else [pool]
    assert sum(len(p) for p in parts) == len(pool)
    rL = sum(lp[p, t].sum() for p, t in zip(parts, tgt)) - sum(lp[zb == c, c].sum() for c in src)
    lgt = sum(lg(len(p)) for p in parts) - sum(lg(int((zb == c).sum())) for c in src)
    if kind == "tri_split":
        rR = math.log(0.5) if len(src) == 1 else -math.log(0.5)
        return la + lgt + rR + rL
    rR = -math.log(0.5) if len(src) == 2 else math.log(0.5)
    return -la + lgt + rR + rL


@pytest.mark.parametrize("sampler_name", ["jain_neal", "triadic"])
def test_split_merge_acceptance_arithmetic(npb, ctx, oracle, sampler_name):
    sampler = {"jain_neal": npb.JAIN_NEAL, "triadic": npb.TRIADIC}[sampler_name]
    nsub = 2 if sampler == npb.JAIN_NEAL else 3
    X, _ = syn.config(1)
    N = len(X)
    ds = npb.Dataset(ctx, X)
    pr = syn.reference_prior(2)
    npb.NormalInverseWishart(**pr).bind(ctx)
    seed, C = 77, 192
    ch = npb.Chains(ctx, ds, C, Kmax=64, seed=seed)
    if sampler == npb.TRIADIC:
        ch.sweep(npb.ALG8, 3)  # a few Gibbs sweeps first so that picks share clusters often enough to see splits
    sweep0 = 3 if sampler == npb.TRIADIC else 0
    checked = {}
    for rep in range(12):
        zb = ch.assignments()
        pk = picks_of(npb, seed, sweep0 + rep, N, nsub)
        params_b = [ch.params(c) for c in range(C)]
        st = ch.split_merge(sampler, 1)
        det = ch.last_proposal()
        za = ch.assignments()
        if len(set(pk)) != nsub:
            assert st.reassignments == 0 and np.array_equal(za, zb)
            continue
        assert st.reassignments == C
        for c in range(C):
            d = {k: float(v[c]) for k, v in det.items()}
            cl = [int(zb[c][i]) for i in pk]
            accepted = d["accept"] > 0.5
            if not accepted:
                assert np.array_equal(za[c], zb[c])
            else:
                check_state(ch, za[c], c, N)
            typ, stat = int(d["type"]), int(d["stat"])
            # the plan, restated from update() (np_jain_neal_algorithm.cpp:424-502, np_triadic_algorithm.cpp:633-795)
            if sampler == npb.JAIN_NEAL:
                if cl[0] == cl[1]:
                    kind, plan = "jn_split", dict(src=[cl[0]])
                    assert typ == 0 and stat == 1
                else:
                    kind, plan = "jn_merge", dict(src=[cl[0], cl[1]])
                    assert typ == 1 and stat == 0
            else:
                uniq = len(set(cl))
                dup = 1 if cl[1] == cl[0] else 2
                if uniq == 1:
                    assert typ == 2 and stat == 1
                    kind, plan = "tri_split", dict(src=[cl[0]], tgt=[cl[0], int(d["new_slot"])])
                elif stat == 0:
                    assert typ == 3
                    keep = [0, 2] if dup == 1 else [0, 1]
                    kind, plan = "tri_merge", dict(src=[cl[keep[0]], cl[keep[1]]], tgt=[cl[keep[0]]])
                elif uniq == 2:
                    assert typ == 2 and stat == 3
                    o = [0, 2, 1] if dup == 1 else [0, 1, 2]
                    kind, plan = "tri_split", dict(src=[cl[o[0]], cl[o[1]]], tgt=[cl[o[0]], cl[o[1]], int(d["new_slot"])])
                else:
                    assert typ == 3 and stat == 2
                    kind, plan = "tri_merge", dict(src=cl, tgt=cl[:2])
            deterministic = kind == "jn_merge" or (kind == "tri_merge" and len(plan["tgt"]) == 1)
            if not (deterministic or accepted):
                continue  # a rejected randomised allocation leaves no trace to recompute from
            slots, _, mu, Sigma = ch.params(c) if accepted else params_b[c]
            if not accepted and kind in ("jn_split", "tri_split"):
                continue
            lp = np.full((N, ch.Kmax), np.nan)
            lp[:, slots] = oracle.mvn_logpdf_batch(mu, Sigma, X)
            if accepted and plan.get("src"):
                # parameters of a cluster removed by an accepted merge come from the state before
                sb, _, mub, Sigb = params_b[c]
                missing = [s for s in plan["src"] if s not in slots]
                if missing:
                    lpb = oracle.mvn_logpdf_batch(mub, Sigb, X)
                    for s in missing:
                        lp[:, s] = lpb[:, list(sb).index(s)]
            want = ref_logA(kind, pr["alpha"], lp, zb[c], za[c] if accepted else zb[c], pk, plan, d)
            got = d["logA"]
            assert abs(got - want) <= 2e-4 * max(1.0, abs(want)) + 2e-3, (kind, c, rep, got, want)
            # the Metropolis-Hastings decision itself: accept iff !(exp(logA) < u)
            assert accepted == (not (math.exp(min(want, 50.0)) < d["u"])) or abs(math.exp(min(want, 50.0)) - d["u"]) < 1e-3
            checked[kind + ("+" if accepted else "-")] = checked.get(kind + ("+" if accepted else "-"), 0) + 1
    print("checked", checked)
    if sampler == npb.JAIN_NEAL:
        assert checked.get("jn_merge+", 0) + checked.get("jn_merge-", 0) > 100
    else:
        assert sum(v for k, v in checked.items() if k.startswith("tri_merge")) > 100
    ch.close()
    ds.close()


def ks_ok(a, b, name, pmin=0.01):
    p = sps.ks_2samp(a, b).pvalue
    assert p > pmin, "%s: KS p = %.4g (device mean %.4f, oracle mean %.4f)" % (name, p, np.mean(a), np.mean(b))


@pytest.mark.parametrize("sampler_name", ["jain_neal", "triadic"])
def test_split_merge_distribution_matches_oracle_256_seeds(npb, ctx, sampler_name):
    sampler = {"jain_neal": npb.JAIN_NEAL, "triadic": npb.TRIADIC}[sampler_name]
    g = np.load(os.path.join(GOLDEN, "oracle_cfg1_%s_256seeds.npz" % sampler_name))
    T = int(g["T"])
    X, y = syn.config(1)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(2)).bind(ctx)
    ch = npb.Chains(ctx, ds, 256, Kmax=64, seed=4242)
    att, acc, sams, made = np.zeros(4), np.zeros(4), 0, 0
    for _ in range(T // 50):
        st = ch.sweep(sampler, 50)
        assert st.overflow_chains == 0
        att += np.array(st.sm_attempts[:])
        acc += np.array(st.sm_accepts[:])
        sams += st.sams_allocations
        made += st.reassignments
    m = ch.metrics(y)
    gatt, gacc = g["attempts"].sum(0), g["accepts"].sum(0)
    print(sampler_name, "device attempts", att, "accepts", acc, "oracle attempts", gatt, "accepts", gacc,
          "K", m["K"].mean(), g["K_final"].mean())
    # proposals made and the mix of move types
    assert abs(made - g["updates"].sum()) < 0.002 * g["updates"].sum()
    for j in range(4):
        if gatt[j] > 0:
            assert abs(att[j] - gatt[j]) < 0.05 * gatt[j] + 200, (j, att[j], gatt[j])
            # acceptance rate per move type: binomial tolerance (5 sigma) plus 10 % relative
            ra, rg = acc[j] / max(att[j], 1), gacc[j] / gatt[j]
            tol = 5 * math.sqrt(rg * (1 - rg) / gatt[j] + rg * (1 - rg) / max(att[j], 1)) + 0.10 * rg
            assert abs(ra - rg) < tol, (j, ra, rg, tol)
    assert abs(sams / made - g["sams"].sum() / g["updates"].sum()) < 0.05 * g["sams"].sum() / g["updates"].sum()
    ks_ok(m["K"].astype(float), g["K_final"], "K")
    if sampler == npb.TRIADIC:
        ks_ok(m["purity"], g["purity"], "purity")
        ks_ok(m["rand_index"], g["rand"], "rand")
        ks_ok(m["adjusted_rand"], g["ari"], "ari")
    else:
        # the bug-compatible Jain-Neal sampler collapses to one cluster: purity 0.5 on two equal classes
        assert abs(m["purity"].mean() - g["purity"].mean()) < 0.01
    ch.close()
    ds.close()


@pytest.mark.parametrize("sampler_name", ["jain_neal", "triadic"])
def test_split_merge_full_size_properties_config3(npb, ctx, sampler_name):
    """BASELINE configs[2] shape (16-D, 32 components, N = 100 000) at a reduced chain count: size-independent
    properties of a prefix of a sweep -- bookkeeping invariants, conservation of the items, rejected proposals leave the
    state untouched, idempotent readback."""
    sampler = {"jain_neal": npb.JAIN_NEAL, "triadic": npb.TRIADIC}[sampler_name]
    X, y = syn.config(3)
    K = int(y.max()) + 1
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(16)).bind(ctx)
    ch = npb.Chains(ctx, ds, 24, Kmax=64, seed=6)
    means = np.stack([X[y == k].mean(0) for k in range(K)])
    ch.init_from_params(means, np.tile(np.eye(16), (K, 1, 1)))
    ch.sweep(npb.ALG8, 1)
    zb = ch.assignments()
    st = ch.split_merge(sampler, 12)
    assert st.overflow_chains == 0 and 0 < st.reassignments <= 24 * 12
    assert sum(st.sm_attempts) == st.reassignments
    za = ch.assignments()
    for c in range(24):
        check_state(ch, za[c], c, ds.N)
    if sum(st.sm_accepts) == 0:
        # well separated unit-variance clusters: merging two of them or splitting one with a prior draw is rejected
        assert np.array_equal(za, zb)
    m = ch.metrics(y)
    assert np.array_equal(m["K"], [len(np.unique(za[c])) for c in range(24)])
    assert np.array_equal(ch.assignments(), za)
    ch.close()
    ds.close()


@pytest.mark.parametrize("sampler_name,seed", [("jain_neal", 1), ("jain_neal", 2), ("triadic", 1), ("triadic", 2)])
def test_split_merge_replay_bit_exact_config1(npb, ctx, oracle, sampler_name, seed):
    """Parity level 2: config 1 with the oracle's recorded draws (subsets, prior draws, shuffled pools, every uniform)
    replayed in double precision on the device.  The oracle run itself is pinned proposal by proposal to the reference's
    own code (tests/test_ref_pin.py).  The device must derive the same move types, allocate every pool member to the same
    part, accept and reject the same proposals and end with identical assignments."""
    alg = {"jain_neal": oracle.JAIN_NEAL, "triadic": oracle.TRIADIC}[sampler_name]
    sampler = {"jain_neal": npb.JAIN_NEAL, "triadic": npb.TRIADIC}[sampler_name]
    X, _ = syn.config(1)
    T = 60
    run = oracle.Run(oracle.make_prior(**syn.reference_prior(2)), X, alg, T=T, seed_main=40 + seed, seed_shuffle=50 + seed,
                     flags=oracle.RECORD_TRACE | oracle.UPDATE_CLUSTERS)
    t = run.sm_trace()
    n = len(t["type"])
    assert n == run.stats().updates and t["accept"].sum() > 20
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(2)).bind(ctx)
    out = npb.replay_split_merge(ctx, ds, sampler, t, run.init_state())
    assert np.array_equal(out["type"], t["type"])
    assert np.array_equal(out["dec"], t["dec"])
    assert np.array_equal(out["accept"], t["accept"])
    assert np.array_equal(out["z_final"], t["z_after"][-1])
    fin = np.isfinite(t["logA"])
    assert np.allclose(out["logA"][fin], t["logA"][fin], rtol=1e-9, atol=1e-7)
    ds.close()


@pytest.mark.parametrize("D,kmax", [(2, 64), (16, 32)])
def test_mixed_schedule_keeps_the_state_consistent(npb, ctx, D, kmax):
    """One handle driven through every entry point in turn (Gibbs sweeps, triadic and Jain-Neal proposals, single-item
    updates, parameter refresh, max-likelihood snapshot): the bookkeeping invariants must hold after each."""
    X, y = syn.gmm(800, D, 5, 400 + D)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(D)).bind(ctx)
    ch = npb.Chains(ctx, ds, 10, Kmax=kmax, K0=8, seed=8)
    means = np.stack([X[y == k].mean(0) for k in range(5)])
    ch.init_from_params(means + 0.3, np.tile(np.eye(D), (5, 1, 1)))
    pr = dict(mu0=X.mean(0), kappa=0.01, nu=D + 2.0, Lambda=np.eye(D))

    def check():
        z = ch.assignments()
        for c in (0, 9):
            check_state(ch, z[c], c, ds.N)
        m = ch.metrics(y)
        assert np.array_equal(m["K"], [len(np.unique(z[c])) for c in range(10)])
        return z

    for rnd in range(2):
        assert ch.sweep(npb.ALG8, 2).overflow_chains == 0
        check()
        ch.split_merge(npb.TRIADIC, 150)
        check()
        ch.update_params(npb.UPDATE_POSTERIOR_DRAW, pr)
        check()
        ch.sweep(npb.JAIN_NEAL, 1)
        check()
        if D == 2:
            ch.update_item(17)
            check()
        cur, best = ch.consider_max_likelihood()
        assert np.all(best >= cur - 1e-9)
        ch.sweep(npb.ALG8, 1)
        z = check()
    assert ch.best_assignments().shape == z.shape
    ch.close()
    ds.close()


@pytest.mark.parametrize("sampler_name,which", [("jain_neal", 1), ("triadic", 1), ("triadic", 3)])
def test_scan_32_members_at_a_time_equals_the_sequential_scan(npb, ctx, sampler_name, which):
    """the restricted (SAMS) scan of k_split_merge decides 32 pool members per round and repeats the round until no decision changes
    (every lane uses the part sizes implied by the earlier lanes' decisions): that fixed point IS the sequential scan's result --
    option spec = 0 runs the member-at-a-time loop; assignments, attempts and accepts must be identical"""
    sampler = {"jain_neal": npb.JAIN_NEAL, "triadic": npb.TRIADIC}[sampler_name]
    X, y = syn.config(which)
    if which == 3:
        X, y = X[:6000], y[:6000]
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(X.shape[1])).bind(ctx)
    res = {}
    for spec in ("1", "0"):
        ch = npb.Chains(ctx, ds, 48, Kmax=64, seed=11)
        ch.set_option("spec", spec)
        att, acc = np.zeros(4, np.int64), np.zeros(4, np.int64)
        for _ in range(3):
            st = ch.sweep(sampler, 2)
            att += np.array(list(st.sm_attempts))
            acc += np.array(list(st.sm_accepts))
        res[spec] = (ch.assignments().copy(), att, acc)
        ch.close()
    assert np.array_equal(res["1"][0], res["0"][0])
    assert np.array_equal(res["1"][1], res["0"][1]) and np.array_equal(res["1"][2], res["0"][2])
    assert res["1"][1].sum() > 0
    print("%s on config %d: identical; attempts %s accepts %s" % (sampler_name, which, res["1"][1], res["1"][2]))
    ds.close()
