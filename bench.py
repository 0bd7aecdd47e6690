#!/usr/bin/env python
"""bench.py -- item-reassignments/sec of the Gibbs reassignment path (BASELINE.json metric).

Headline workload (config.workload): the per-GPU shard of BASELINE.json configs[4] -- the "16-D N=100k config" the
metric and its target are quoted on: 8192 lockstep Algorithm-8 chains per GPU (65536 over 8 GPUs) over one synthetic
32-component 16-D GMM with N = 100 000 items (noparama_b200/synthetic.py, seed 20261003), reference NIW prior
generalised to 16-D (np_main.cpp:164,367-371), m = 3 auxiliary draws, Kmax = 32.  Regime (SURVEY 8d honesty rule): under
the reference's bug-compatible prior a 16-D run collapses to one cluster and a reassignment becomes trivially cheap, so
the chains start from K0 = K_true = 32 given clusters (class means, identity covariance) and every reassignment weighs
32 occupied clusters + 3 auxiliary draws; cluster parameters stay frozen between births exactly as in the reference (Q1).
A "step" is one sweep: every chain reassigns every item once (np_mcmc.cpp:146-163) = chains * N item-reassignments, one
sweep-kernel launch per GPU.  `--config cfg2` measures BASELINE configs[1] (1024 chains, 2-D, reference prior and
initialisation) instead; the default run also reports it, and the split-merge samplers at configs[2]'s shape, under "also".

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
  python bench.py --impl reference [--gpus N] ...              the reference's own CPU sampler on the host cores

Under torchrun every rank owns one GPU and its own chains (chains are independent: weak scaling, no collective on the
data path); timing is barrier + synchronize on both sides, CUDA events on the library's stream, max over ranks.  After
the timed region the ranks all-reduce (NCCL) the posterior co-clustering matrix of an anchor subset and the per-chain
diagnostics, the only exchange the path has.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

M_AUX, K0_REF = 3, 20
SEED = 20261018
METRIC = "item-reassignments/sec"
UNIT = "reassignments/s"

CONFIGS = {
    # per-GPU shard of BASELINE configs[4] (also the data of configs[2]): the headline
    "cfg5": dict(synthetic=5, N=100_000, D=16, K_true=32, chains=8192, kmax=32, given=True,
                 workload="BASELINE configs[4] per-GPU shard (the 16-D N=100k config): 8192 lockstep Algorithm-8 chains per "
                          "GPU (65536 over 8), synthetic 32-component 16-D GMM, N=100000, m=3 aux, Kmax=32, reference NIW "
                          "prior generalised to 16-D; chains start from K0=K_true=32 given clusters (class means, identity "
                          "covariance), parameters frozen between births as in the reference"),
    # BASELINE configs[1]
    "cfg2": dict(synthetic=2, N=100_000, D=2, K_true=10, chains=1024, kmax=256, given=False,
                 workload="BASELINE configs[1]: 1024 lockstep Algorithm-8 chains per GPU, synthetic 10-component 2-D GMM, "
                          "N=100000, m=3 aux, K0=20, reference NIW prior and initialisation (bug-compatible)"),
}


def f_eval(D):
    """algorithmic flops of one density evaluation, SURVEY 8d: D^2 + 4D + 3"""
    return D * D + 4 * D + 3


def given_clusters(X, y):
    K = int(y.max()) + 1
    return np.stack([X[y == k].mean(0) for k in range(K)]), np.tile(np.eye(X.shape[1]), (K, 1, 1))


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own sampler (oracle/_ref/np_ref_run = its unmodified sources built against oracle/eigen_shim)
# when that binary travelled with the snapshot, else the oracle port.  One independent chain per host core (the
# reference is single threaded), on a bounded prefix of the workload's (already shuffled) items.
# ---------------------------------------------------------------------------------------------------------------
def ref_chains(cfg, cores, sweeps, n_items):
    """`cores` concurrent runs of the reference binary -> list of per-sweep update()-loop seconds, K_final, seconds_run"""
    from oracle import refrun
    from noparama_b200 import synthetic as syn
    X, _ = syn.config(cfg["synthetic"])
    X = X[:n_items]
    pr = syn.reference_prior(cfg["D"])
    with tempfile.TemporaryDirectory() as d:
        req = os.path.join(d, "req.bin")
        refrun.write_request(req, X, pr)
        procs = []
        for c in range(cores):
            res = os.path.join(d, "res%d.bin" % c)
            procs.append((res, subprocess.Popen(refrun.command(8, sweeps, 1000 + c, 2000 + c, req, res),
                                                stdout=subprocess.DEVNULL)))
        out = []
        for res, p in procs:
            if p.wait() != 0:
                raise RuntimeError("np_ref_run failed")
            r = refrun.read_result(res)
            out.append((r["sweep_update_seconds"], r["K_final"], r["seconds_run"]))
    return out


def port_chain_worker(args):
    cfgname, seed, sweeps, n_items, given = args[:5]
    optimised = len(args) > 5 and args[5]
    from oracle import binding as orc
    from noparama_b200 import synthetic as syn
    cfg = CONFIGS[cfgname]
    X, y = syn.config(cfg["synthetic"])
    g = given_clusters(X, y) if given else None
    X = X[:n_items]
    pr = orc.make_prior(**syn.reference_prior(cfg["D"]))
    # FAITHFUL = the reference's cost profile (LU inverse + determinant per density call, dense membership matrix, UpdateClusters,
    # max-likelihood bookkeeping); optimised = inverse and determinant cached per cluster, log-domain weights, none of the rest
    r = orc.Run(pr, X, T=sweeps, K0=K0_REF, M_aux=M_AUX, seed_main=1000 + seed, seed_shuffle=2000 + seed,
                flags=orc.LOG_DOMAIN if optimised else orc.FAITHFUL, given=g)
    re, tot = r.sweep_seconds()
    s = r.stats()
    return re.tolist(), tot.tolist(), s.candidates / max(1, s.updates), s.K_final


def port_chains(cfgname, cores, sweeps, n_items, given, optimised=False):
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(cores) as pool:
        return pool.map(port_chain_worker, [(cfgname, c, sweeps, n_items, given, optimised) for c in range(cores)])


def cpu_sample_items(cfg):
    # about 10-30 s of CPU work per core: the reference spends ~0.3 ms (16-D, collapsed) .. 20 us (2-D) per reassignment
    return 4000 if cfg["D"] >= 8 else 100_000


def cpu_baseline_sample(cfgname, cores):
    """bounded sample of the same workload on the host cores: 1 warm-up + 2 timed sweeps per chain"""
    from oracle import refrun
    cfg = CONFIGS[cfgname]
    n_items = cpu_sample_items(cfg)
    out = {"unit": UNIT, "cores": cores}
    if refrun.available():
        res = ref_chains(cfg, cores, 3, n_items)
        t = max(float(np.sum(r[0][1:])) for r in res)
        out.update(value=cores * n_items * 2 / t, kind="reference",
                   sample="%d independent chains (one per host core) of the reference's own sampler (oracle/_ref: its unmodified "
                          "sources built against the Eigen stand-in), sweeps 2-3 over the first %d items of the workload, "
                          "update() loop only (np_mcmc.cpp:146-163); reference initialisation (K0=20 prior draws): final "
                          "K = %.1f clusters" % (cores, n_items, float(np.mean([r[1] for r in res]))),
                   single_core_value=n_items * 2 / float(np.sum(res[0][0][1:])))
    # the oracle port with the reference cost profile, in the very regime the GPU number is taken in
    n_port = 1000 if cfg["given"] else n_items
    pres = port_chains(cfgname, cores, 3, n_port, cfg["given"])
    tp = max(sum(r[0][1:]) for r in pres)
    port = {"value": cores * n_port * 2 / tp, "full_sweep_value": cores * n_port * 2 / max(sum(r[1][1:]) for r in pres),
            "candidates_per_reassignment": float(np.mean([r[2] for r in pres])), "items": n_port,
            "regime": "same as the GPU run (%s)" % ("K0=K_true given clusters" if cfg["given"] else "reference initialisation")}
    if "value" not in out:
        out.update(value=port["value"], kind="port", sample="%d chains x sweeps 2-3 x %d items, oracle port, update() loop only"
                   % (cores, n_port))
    out["port_same_regime"] = port
    # an OPTIMISED CPU figure for honesty -- NOT the reference: the same sampler with every cluster's inverse covariance
    # and normaliser cached (parameters are frozen between births, Q1), log-domain weights, no dense membership matrix
    n_opt = 20_000 if cfg["given"] else n_items
    ores = port_chains(cfgname, cores, 3, n_opt, cfg["given"], optimised=True)
    out["optimised_cpu_not_the_reference"] = {
        "value": cores * n_opt * 2 / max(sum(r[0][1:]) for r in ores), "unit": UNIT, "cores": cores, "items": n_opt,
        "candidates_per_reassignment": float(np.mean([r[2] for r in ores])),
        "what": "oracle port with the inverse covariance and normaliser cached per cluster and log-domain weights, one chain per core, "
                "same regime as the GPU run; not the reference's cost profile"}
    return out


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on all host cores, one independent chain per
    core, same config / metric / unit; a step = one sweep of every chain over a bounded prefix of the items."""
    if rank != 0:
        return
    from oracle import refrun
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    sweeps = args.warmup + args.steps
    n_items = cpu_sample_items(cfg)
    if sweeps > 12:  # keep the whole run within a few minutes
        n_items = max(500, n_items * 12 // sweeps)
    t0 = time.time()
    if refrun.available():
        res = ref_chains(cfg, cores, sweeps, n_items)
        t = max(float(np.sum(r[0][args.warmup:])) for r in res)
        kind, how = "reference", "oracle/_ref/np_ref_run (the reference's unmodified sampler sources built against oracle/eigen_shim)"
        extra = {"K_final_mean": float(np.mean([r[1] for r in res]))}
    else:
        res = port_chains(args.config, cores, sweeps, n_items, False)
        t = max(sum(r[0][args.warmup:]) for r in res)
        kind, how = "port", "oracle port (oracle/_ref was not built)"
        extra = {"candidates_per_reassignment": float(np.mean([r[2] for r in res]))}
    wall = time.time() - t0
    n = cores * n_items * args.steps
    val = n / t
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict({"workload": cfg["workload"], "chains": cores, "items_per_sweep": n_items, "how": how,
                            "note": "one reference chain per host core with the reference's own initialisation (K0=20 prior "
                                    "draws); a step = one sweep of every chain over items_per_sweep items, update() loop only"},
                           **extra),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d chains x %d sweeps x %d items" % (cores, args.steps, n_items)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
def build_chains(npb, syn, ctx, cfg, chains, kmax, seed):
    X, y = syn.config(cfg["synthetic"])
    ds = npb.Dataset(ctx, X)
    prior = npb.NormalInverseWishart(**syn.reference_prior(cfg["D"]))
    mc = npb.MCMC(ctx, ds, prior, chains=chains, Kmax=kmax, K0=K0_REF, m_aux=M_AUX, seed=seed)
    if cfg["given"]:
        mc.chains.init_from_params(*given_clusters(X, y))
    return X, y, ds, mc


def timed_sweeps(npb, chains, steps):
    ms, cand, moved, births, last = [], 0, 0, 0, None
    for _ in range(steps):
        st = chains.sweep(npb.ALG8, 1)
        ms.append(st.kernel_ms)
        cand += st.candidates
        moved += st.moved
        births += st.new_clusters
        last = st
    return ms, cand, moved, births, last


def also_measure(npb, syn, ctx, fp32_peak, rank, cpu=True):
    """secondary figures of the default run: BASELINE configs[1] (2-D) and the split-merge samplers at configs[2]'s shape"""
    out = {}
    cfg = CONFIGS["cfg2"]
    X, y, ds, mc = build_chains(npb, syn, ctx, cfg, cfg["chains"], cfg["kmax"], SEED + 17 * rank)
    for _ in range(5):
        mc.chains.sweep(npb.ALG8, 1, want_stats=False)
    ms, cand, moved, births, last = timed_sweeps(npb, mc.chains, 10)
    k_ms = float(np.mean(ms))
    n_step = cfg["chains"] * ds.N
    fl = (cand / 10) * (f_eval(cfg["D"]) + 6)
    out["cfg2"] = {"workload": cfg["workload"], "value": n_step / (k_ms * 1e-3), "unit": UNIT, "kernel_ms": k_ms, "steps": 10,
                   "warmup": 5, "mean_K": last.mean_K, "candidates_per_reassignment": cand / (n_step * 10),
                   "moved_fraction": moved / (n_step * 10),
                   "roofline": {"bound": "fp32", "achieved": fl / (k_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                                "frac": fl / (k_ms * 1e-3) / 1e12 / fp32_peak if fp32_peak else None,
                                "kernel": "k_alg8_sweep_reg<2,8>"}}
    mc.chains.close()
    ds.close()
    # split-merge proposals (BASELINE configs[2] shape: 16-D, N = 100k, 4096 chains; SURVEY 8d: proposals/s and
    # SAMS item-allocations/s)
    cfg = CONFIGS["cfg5"]
    X, y = syn.config(3)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(cfg["D"])).bind(ctx)
    # the `fixed` regime of SURVEY 8d on the headline shape at reduced chain count: every sweep is followed by a draw of all
    # cluster parameters from their conjugate NIW posterior (npb_chains_update_params)
    pr_fixed = dict(mu0=X.mean(0), kappa=0.01, nu=cfg["D"] + 2.0, Lambda=np.eye(cfg["D"]))
    chf = npb.Chains(ctx, ds, 4096, Kmax=32, K0=K0_REF, seed=SEED + 37 * rank)
    chf.init_from_params(*given_clusters(X, y))
    chf.sweep(npb.ALG8, 1, want_stats=False)
    t_sw, t_up = [], []
    for it in range(4):
        st = chf.sweep(npb.ALG8, 1)
        ctx.synchronize()
        t0 = time.perf_counter()
        chf.update_params(npb.UPDATE_POSTERIOR_DRAW, pr_fixed)
        ctx.synchronize()
        if it:
            t_sw.append(st.kernel_ms)
            t_up.append((time.perf_counter() - t0) * 1e3)
    mfix = chf.metrics(y)
    chf.close()
    out["fixed_mode"] = {"workload": "4096 chains x N=100000 x 16-D, K=32: one Algorithm-8 sweep + one posterior draw of every cluster's "
                                     "(mu, Sigma) per step", "sweep_ms": float(np.mean(t_sw)), "update_params_ms": float(np.mean(t_up)),
                         "value": 4096 * ds.N / ((np.mean(t_sw) + np.mean(t_up)) * 1e-3), "unit": UNIT,
                         "mean_purity": float(mfix["purity"].mean()), "mean_K": float(mfix["K"].mean())}
    out["split_merge"] = split_merge_measure(npb, syn, ctx, ds, X, y, fp32_peak, rank, cpu=cpu)
    ds.close()
    out["cfg4"] = cfg4_measure(npb, syn, ctx, fp32_peak, rank)
    out["fp32_pipe_kernel"] = fp32_path_measure(npb, syn, ctx, fp32_peak, rank)
    out["mixing"] = mixing_measure(npb, syn, ctx, rank)
    out["conjugate_alg2"] = conjugate_measure(npb, syn, ctx, rank)
    return out


def split_merge_measure(npb, syn, ctx, ds, X, y, fp32_peak, rank, proposals=10_000, cpu=True):
    """Jain-Neal and triadic proposals, 10 000 lockstep proposals per chain (np_mcmc.cpp:146-163: a sweep is N of them), in two
    regimes.  (1) BASELINE configs[2]'s shape (16-D, N = 100 000, 32 components), the chains started from 16 clusters that each
    merge two true components: a split draws its new cluster's parameters from the PRIOR (np_jain_neal_algorithm.cpp:148,
    np_triadic_algorithm.cpp:260), which in 16 dimensions never lands near the data, so the reference's samplers accept nothing
    here -- the figure is the cost of the proposals.  (2) BASELINE configs[1]'s data (2-D, N = 100 000, 10 components), where the
    reference's own tests run these samplers, from the reference's initialisation (K0 = 20 prior draws): merges and splits ARE
    accepted.  Reported: proposals/s, SAMS item-allocations/s, attempts / accepts, the FP32 figure of the counted
    density evaluations.  CPU: the reference's own samplers (oracle/_ref np_ref_run), one chain per core, on a bounded prefix."""
    K = int(y.max()) + 1
    pairs = [(2 * i, 2 * i + 1) for i in range(K // 2)]
    mu = np.stack([X[(y == a) | (y == b)].mean(0) for a, b in pairs])
    Sig = np.stack([np.cov(X[(y == a) | (y == b)].T) for a, b in pairs])
    res = {}

    def one(ch, sampler, q_eval, D, truth, kernel):
        ch.split_merge(sampler, 256)
        st = ch.split_merge(sampler, proposals)
        sec = st.kernel_ms * 1e-3
        m = ch.metrics(truth)
        # a SAMS allocation weighs its member under the Q parameter sets of the move (2 for Jain-Neal, 2 or 3 for the triadic
        # sampler), the acceptance sums evaluate every member once more per set: ~2 Q evaluations per allocated member
        fl = st.sams_allocations * 2.0 * q_eval * f_eval(D)
        return {"proposals_per_s": st.reassignments / sec, "sams_allocations_per_s": st.sams_allocations / sec,
                "kernel_ms": st.kernel_ms, "chains": ch.C, "proposals": int(st.reassignments), "proposals_per_chain": proposals,
                "attempts": list(st.sm_attempts), "accepts": list(st.sm_accepts),
                "accept_rate": float(sum(st.sm_accepts)) / max(1, sum(st.sm_attempts)), "mean_K_after": float(m["K"].mean()),
                "mean_purity_after": float(m["purity"].mean()),
                "roofline": {"bound": "fp32", "achieved": fl / sec / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                             "frac": fl / sec / 1e12 / fp32_peak if fp32_peak else None, "kernel": kernel,
                             "note": "density evaluations estimated as 2 Q per SAMS allocation (Q = parameter sets of the move) x (D^2 + 4D + 3) flops"}}

    D = X.shape[1]
    for name, sampler, q_eval, chains in (("jain_neal", npb.JAIN_NEAL, 2.0, 592), ("triadic", npb.TRIADIC, 2.5, 148)):
        ch = npb.Chains(ctx, ds, chains, Kmax=64, K0=K0_REF, seed=SEED + 31 * rank)
        ch.init_from_params(mu, Sig)
        ch.sweep(npb.ALG8, 1, want_stats=False)  # the items settle on the merged clusters
        res[name] = one(ch, sampler, q_eval, D, y, "k_split_merge<16>")
        ch.close()
    res["workload"] = ("BASELINE configs[2] shape (of its 4096 chains: 592 Jain-Neal, 148 triadic), 16-D 32-component GMM, N=100000, Kmax=64; start = "
                       "16 clusters, each two true components merged; %d lockstep proposals per chain; a split's parameters are a prior "
                       "draw, hopeless at 16-D: nothing is accepted, in the reference either" % proposals)
    # regime 2: the reference's own territory (2-D), accepted moves
    X2, y2 = syn.config(2)
    X2, y2 = X2[:4000], y2[:4000]  # (from this start the chains soon hold a few clusters of ~N items: a proposal re-allocates ~N of them)
    ds2 = npb.Dataset(ctx, X2)
    acc = {}
    for name, sampler, q_eval in (("jain_neal", npb.JAIN_NEAL, 2.0), ("triadic", npb.TRIADIC, 2.5)):
        mc = npb.MCMC(ctx, ds2, npb.NormalInverseWishart(**syn.reference_prior(2)), chains=148, Kmax=256, K0=K0_REF, seed=SEED + 41 * rank)
        acc[name] = one(mc.chains, sampler, q_eval, 2, y2, "k_split_merge<2>")
        mc.chains.close()
    ds2.close()
    acc["workload"] = ("BASELINE configs[1] data, its first 4000 items (2-D 10-component GMM), 148 chains, Kmax=256, reference initialisation (K0 = 20 "
                       "prior draws), %d lockstep proposals per chain" % proposals)
    res["accepting_regime_2d"] = acc
    npb.NormalInverseWishart(**syn.reference_prior(D)).bind(ctx)
    if cpu:
        try:
            res["cpu_baseline"] = split_merge_cpu(X, D)
        except Exception as e:
            res["cpu_baseline"] = {"failed": repr(e)}
    return res


def split_merge_cpu(X, D, n_items=1000, sweeps=1):
    """the reference's own split-merge samplers on the host cores: one chain per core over the first n_items items, `sweeps`
    sweeps (= n_items proposals each, np_mcmc.cpp:146-163) after its own initialisation"""
    from oracle import refrun
    from noparama_b200 import synthetic as syn
    if not refrun.available():
        return {"unavailable": "oracle/_ref/np_ref_run was not built"}
    cores = os.cpu_count() or 1
    out = {"cores": cores, "kind": "reference", "items": n_items}
    pr = syn.reference_prior(D)
    with tempfile.TemporaryDirectory() as d:
        req = os.path.join(d, "req.bin")
        refrun.write_request(req, X[:n_items], pr)
        for alg, name in ((2, "jain_neal"), (3, "triadic")):
            procs = []
            for c in range(cores):
                r = os.path.join(d, "res%d_%d.bin" % (alg, c))
                procs.append((r, subprocess.Popen(refrun.command(alg, sweeps, 3000 + c, 4000 + c, req, r), stdout=subprocess.DEVNULL)))
            t, calls = 0.0, 0
            for r, pp in procs:
                if pp.wait() != 0:
                    raise RuntimeError("np_ref_run failed")
                rr = refrun.read_result(r)
                t = max(t, rr["seconds_update"])
                calls += rr["calls"]
            out[name] = {"proposals_per_s": calls / t, "proposals": int(calls), "seconds": t}
    out["sample"] = "%d chains (one per core) x %d sweeps over the first %d items, reference initialisation (K0 = 20 prior draws)" % (cores, sweeps, n_items)
    return out


def conjugate_measure(npb, syn, ctx, rank):
    """BASELINE configs[3]: CONJUGATE Algorithm 2 (collapsed Gibbs, NIW posterior predictive, statistics up- and down-dated per
    move).  `cfg4_64d` is configs[3]'s own shape, 256 chains x N = 1 000 000 x 64-D, on k_a2_tc (npb_alg2_tc.cu: a tile of 128
    steps evaluated ahead of the chain with the quadratic forms on tcgen05, the two changed clusters evaluated again after every
    move; same chain as the sequential schedule bit for bit); `fp32_tile_kernel_64d` is the FP32 tile kernel k_a2_tile on a prefix
    of the same data; the chains start from the true partition (one state broadcast to all chains, their random streams
    their own) so that the warm-up does not have to move a million items per chain.  `step_at_a_time_kernel` is round 2's
    first path (k_a2_sweep) on a short prefix of the same data, for the ratio."""
    out = {}
    for name, D, N, chains, K, tile, truth, tc in (("cfg4_64d", 64, 1_000_000, 256, 16, 128, True, 1), ("headline_shape_16d", 16, 100_000, 1024, 20, 64, False, 1),
                                                     ("fp32_tile_kernel_64d", 64, 100_000, 256, 16, 64, True, 0), ("step_at_a_time_kernel_64d", 64, 20_000, 256, 16, 0, True, 0)):
        X, y = syn.gmm(N, D, K, 20261004)
        ds = npb.Dataset(ctx, X)
        npb.NormalInverseWishart(mu0=X.mean(0), kappa=0.01, nu=D + 2.0, Lambda=np.eye(D), alpha=1.0).bind(ctx)
        ch = npb.Chains(ctx, ds, chains, Kmax=32, K0=K, m_aux=M_AUX, seed=SEED + 59 * rank)
        ch.set_option("a2_tile", str(tile))
        ch.set_option("a2_tc", str(tc))
        mu, Sig = given_clusters(X, y)
        if truth:
            ch.set_state(0, y.astype(np.int32), np.arange(K, dtype=np.int32), mu, Sig)
            ch.broadcast_state(0)
        else:
            ch.init_from_params(mu, Sig)
        warm = 2 if truth else 3
        for _ in range(warm):
            ch.sweep(npb.ALG2_CONJUGATE, 1)
        ms, cand, moved, births = [], 0, 0, 0
        for _ in range(2):
            st = ch.sweep(npb.ALG2_CONJUGATE, 1)
            ms.append(st.kernel_ms)
            cand, moved, births = cand + st.candidates, moved + st.moved, births + st.new_clusters
        k_ms = float(np.mean(ms))
        m = ch.metrics(y)
        # an Algorithm 2 step = (K_i + 1) (F_eval(D) + 12) flops + 2 (2 D^2 + 6 D) per rank-1 down- / up-date (moves only here)
        fl = (cand / 2) * (f_eval(D) + 12) + (moved / 2) * 2 * (2 * D * D + 6 * D)
        out[name] = {"workload": "%d chains x N=%d x %d-D, %d components, conjugate NIW (mu0 = data mean, kappa0 = 0.01, nu0 = D + 2, Lambda0 = I), "
                                 "Kmax=32; start: %s" % (chains, N, D, K, "the true partition" if truth else "the true clusters' parameters, random assignment"),
                     "kernel": ("k_a2_tc (tcgen05 kind::f16, FP16x3 split; tile of %d steps)" % tile if (tc and D == 64) else
                                "k_a2_tile<%d> (FP32, tile of %d steps)" % (D, tile)) if tile else "k_a2_sweep<%d> (one step at a time)" % D,
                     "value": chains * N / (k_ms * 1e-3), "unit": UNIT, "kernel_ms": k_ms, "steps": 2,
                     "warmup": warm, "moved_fraction": moved / (2 * chains * N), "candidates_per_reassignment": cand / (2 * chains * N),
                     "mean_K": float(m["K"].mean()), "mean_purity": float(m["purity"].mean()), "algorithmic_tflops": fl / (k_ms * 1e-3) / 1e12}
        out[name]["roofline"] = {"bound": "fp32", "achieved": out[name]["algorithmic_tflops"], "unit": "TFLOP/s", "kernel": out[name]["kernel"],
                                 "note": "SURVEY 8(d) algorithmic flops of the sweep / the sweep time; FP32 FMA peak: roofline_fp32_equivalent.peak"}
        if tc and D == 64 and tile:
            # issued kind::f16 flops: 12 MMAs of 128 x 64 x 16 per (tile of 128 steps, cluster with members), clusters = candidates - 1 per step
            issued = (cand / 2 - chains * N) / 128.0 * 12 * 2 * 128 * 64 * 16
            out[name]["roofline"].update({"bound": "tensor", "issued_tflops": issued / (k_ms * 1e-3) / 1e12,
                                          "note": "achieved = SURVEY 8(d) algorithmic flops / sweep time; issued_tflops = kind::f16 MMA flops "
                                                  "(three FP16 products per FP32 product) / sweep time; peak: roofline.peak (measured bf16)"})
        ch.close()
        ds.close()
    return out


def mixing_measure(npb, syn, ctx, rank, chains=8192, timed=3):
    """The headline shape in a regime where items keep moving: the 32 components in pairs 2.5 apart (synthetic.gmm_mixing),
    the true parameters given, so a reassignment changes the item's cluster with probability ~0.15 for ever.  This is what the
    sequential part of the sweep kernels costs; the handle picks its kernels by the moved fraction of its previous sweep
    (fused kernel while few items move, the table + race kernel pair in a mixing chain), and both are timed here."""
    D, K, N = 16, 32, 100_000
    X, y = syn.gmm_mixing(N, D, K, 20261005)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(D)).bind(ctx)
    out = {"workload": "BASELINE configs[4] per-GPU shape (%d chains, N=%d, 16-D, 32 components, Kmax=32, m=3) with the components in "
                       "pairs 2.5 apart (synthetic.gmm_mixing, seed 20261005) and their true parameters given: the chains mix for ever" % (chains, N)}
    for name, path in (("auto", "auto"), ("fused_kernel", "tc"), ("kernel_pair", "tc2")):
        ch = npb.Chains(ctx, ds, chains, Kmax=32, K0=K0_REF, m_aux=M_AUX, seed=SEED + 53 * rank)
        ch.set_option("d16_path", path)
        ch.init_from_params(*given_clusters(X, y))
        for _ in range(2):
            ch.sweep(npb.ALG8, 1)
        ms, cand, moved, births, last = timed_sweeps(npb, ch, timed)
        k_ms = float(np.mean(ms))
        st = ch.sweep(npb.ALG8, 1)
        out[name] = {"value": chains * N / (k_ms * 1e-3), "unit": UNIT, "kernel_ms": k_ms, "steps": timed, "warmup": 2,
                     "moved_fraction": moved / (chains * N * timed), "new_clusters_per_step": births / timed,
                     "candidates_per_reassignment": cand / (chains * N * timed), "mean_K": last.mean_K}
        if name == "auto":
            m = ch.metrics(y)
            out["mean_purity"], out["mean_ari"] = float(m["purity"].mean()), float(m["adjusted_rand"].mean())
        ch.close()
    out["value"], out["unit"] = out["auto"]["value"], UNIT
    out["moved_fraction"] = out["auto"]["moved_fraction"]
    ds.close()
    return out


def fp32_path_measure(npb, syn, ctx, fp32_peak, rank):
    """the headline shape on the FP32-pipe kernel (k_alg8_sweep_tile4, NPB_D16_PATH=fp32): the figure the tensor path replaced"""
    cfg = CONFIGS["cfg5"]
    old = os.environ.get("NPB_D16_PATH")
    os.environ["NPB_D16_PATH"] = "fp32"
    try:
        X, y, ds, mc = build_chains(npb, syn, ctx, cfg, cfg["chains"], cfg["kmax"], SEED + 43 * rank)
        for _ in range(3):
            mc.chains.sweep(npb.ALG8, 1, want_stats=False)
        ms, cand, moved, births, last = timed_sweeps(npb, mc.chains, 3)
        k_ms = float(np.mean(ms))
        n_step = cfg["chains"] * ds.N
        fl = (cand / 3) * (f_eval(cfg["D"]) + 6)
        mc.chains.close()
        ds.close()
    finally:
        if old is None:
            os.environ.pop("NPB_D16_PATH", None)
        else:
            os.environ["NPB_D16_PATH"] = old
    return {"workload": cfg["workload"], "value": n_step / (k_ms * 1e-3), "unit": UNIT, "kernel_ms": k_ms, "steps": 3, "warmup": 3,
            "roofline": {"bound": "fp32", "achieved": fl / (k_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": fl / (k_ms * 1e-3) / 1e12 / fp32_peak if fp32_peak else None, "kernel": "k_alg8_sweep_tile4<16,3>"}}


def cfg4_measure(npb, syn, ctx, fp32_peak, rank, chains=256, N=1_000_000, timed=3):
    """BASELINE configs[3] shape: 256 chains, 64-D, N = 1M, 32 given clusters, Algorithm 2 (one auxiliary draw) through the
    D = 64 path (npb_alg8_gemm.cu): tcgen05 kind::f16 density tables (three FP16 products per FP32 product) fused with the
    warp-per-chain race of the previous block."""
    D, K = 64, 32
    X, y = syn.gmm(N, D, K, syn.SEEDS[4])
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(D)).bind(ctx)
    ch = npb.Chains(ctx, ds, chains, Kmax=32, K0=K0_REF, m_aux=1, seed=SEED + 41 * rank)
    ch.init_from_params(*given_clusters(X, y))
    for _ in range(3):  # the first sweep moves every item to its cluster
        ch.sweep(npb.ALG2, 1, want_stats=False)
    ms, cand = [], 0
    for _ in range(timed):
        st = ch.sweep(npb.ALG2, 1)
        ms.append(st.kernel_ms)
        cand += st.candidates
    k_ms = float(np.mean(ms))
    n_step = chains * N
    m = ch.metrics(y)
    bf16_peak = None
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        bf16_peak = json.load(open(pk)).get("bf16_tflops")
    f16_peak = bf16_peak if bf16_peak else 2250.0  # kind::f16 runs at the bf16 rate (nominal dense figure as fallback)
    mma = n_step * 32 * 3 * 0.75 * 2 * D * D  # issued: 3 FP16 products, 3 of the 4 32x32 blocks of the triangular factor, all 32 slots
    alg = (cand / timed) * (f_eval(D) + 6)
    sec = k_ms * 1e-3
    out = {"workload": "BASELINE configs[3] shape: %d chains, synthetic 32-component 64-D GMM, N=%d, Algorithm 2 (one auxiliary "
                       "draw), Kmax=32, K0=K_true=32 given clusters" % (chains, N),
           "value": n_step / sec, "unit": UNIT, "kernel_ms": k_ms, "steps": timed, "warmup": 3,
           "candidates_per_reassignment": cand / (n_step * timed), "mean_purity": float(m["purity"].mean()),
           "mean_K": float(m["K"].mean()), "algorithmic_tflops": alg / sec / 1e12,
           "algorithmic_over_fp32_peak": alg / sec / 1e12 / fp32_peak if fp32_peak else None,
           "roofline": {"bound": "tensor", "achieved": mma / sec / 1e12, "peak": f16_peak, "unit": "TFLOP/s",
                        "frac": mma / sec / 1e12 / f16_peak, "kernel": "k_density_tc",
                        "note": "achieved = kind::f16 MMA flops issued per sweep (three FP16 products per FP32 product, 3 of the 4 "
                                "32x32 blocks of the triangular 64x64 factor, every (step, slot)) / sweep time, the race of the "
                                "previous block running in two more warps of the same kernel; peak = measured bf16 burst (MEASURED_PEAKS.json)"}}
    ch.close()
    ds.close()
    return out


_REAL_STDOUT = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner at the first
    communicator), so the process's fd 1 is pointed at stderr for the whole run and the line goes out through a saved copy."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="npb200", choices=["npb200", "reference"])
    ap.add_argument("--config", default="cfg5", choices=sorted(CONFIGS))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the config's)")
    ap.add_argument("--kmax", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    cfg = CONFIGS[args.config]
    n_chains = args.chains or cfg["chains"]
    kmax = args.kmax or cfg["kmax"]
    DIM = cfg["D"]

    import torch
    import torch.distributed as dist
    import noparama_b200 as npb
    from noparama_b200 import synthetic as syn
    from noparama_b200 import diagnostics as dg

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = npb.Context(local_rank)
    X, y, ds, mc = build_chains(npb, syn, ctx, cfg, n_chains, kmax, dg.rank_seed(SEED, rank))
    chains = mc.chains
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    d16_path = os.environ.get("NPB_D16_PATH", "auto")
    tc_path = DIM == 16 and kmax == 32 and not d16_path.startswith("f")
    if tc_path:
        chains.set_option("time_kernels", "1")

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- warm-up (also moves the items from the random initial assignment to their clusters) ----
    for _ in range(args.warmup):
        chains.sweep(npb.ALG8, 1)  # (with statistics: the handle picks its D = 16 kernels by the moved fraction of its last sweep)
    if tc_path:
        chains.kernel_time()  # reset
    # ---- timed region: exactly K sweeps, inputs resident in HBM ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    kernel_ms, cand, moved, births, last = timed_sweeps(npb, chains, args.steps)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    kt_ms, kt_n = chains.kernel_time() if tc_path else (0.0, 0)

    # ---- end to end through the public call with host buffers: every step uploads X from pinned host memory and brings the
    # caller's page-locked copy of all assignments up to date (npb_chains_sweep_host_delta: the entries that changed travel as a
    # list compacted on the device, everything when more than a quarter changed).  The full copy of every assignment every step
    # (npb_chains_sweep_host, round 1's figure) is timed next to it. ----
    Xh = torch.from_numpy(np.ascontiguousarray(X)).pin_memory().numpy()
    z_host = torch.empty((ds.N, n_chains), dtype=torch.uint16, pin_memory=True).numpy()
    chains.sweep_host_delta(Xh, npb.ALG8, 1, z_mirror=z_host)  # first call: everything travels, staging buffers warm up
    chains.sweep_host_delta(Xh, npb.ALG8, 1, z_mirror=z_host)
    barrier()
    t0 = time.perf_counter()
    changed = 0
    for _ in range(args.e2e_steps):
        changed += chains.sweep_host_delta(Xh, npb.ALG8, 1, z_mirror=z_host)[1]
    ctx.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    mirror_ok = bool(np.array_equal(z_host[:, :4].T.astype(np.int32), chains.assignments(0, 4)))
    chains.sweep_host(Xh, npb.ALG8, 1, z_out=z_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        chains.sweep_host(Xh, npb.ALG8, 1, z_out=z_host)
    ctx.synchronize()
    e2e_full_s = time.perf_counter() - t0
    barrier()

    # ---- diagnostics exchange (outside the timed region): a short traced phase for R-hat, then one all-reduce of
    # the co-clustering counts, the score sums and the R-hat partials over all ranks (NCCL when world > 1) ----
    k_trace, jll_trace = [], []
    for _ in range(4):
        chains.sweep(npb.ALG8, 1, want_stats=False)
        mm = chains.metrics(None)
        k_trace.append(mm["K"].astype(np.float64))
        jll_trace.append(mm["joint_loglik"])
    m = chains.metrics(y)
    cocl = cocluster_measure(npb, chains, ds, torch, dist, world, rank)
    diag = dg.combine(dg.score_partial(m), {"K": dg.rhat_partial(np.stack(k_trace, 1)),
                                            "joint_loglik": dg.rhat_partial(np.stack(jll_trace, 1))},
                      cocluster=None, device="cuda")
    t = torch.tensor([elapsed_ms, e2e_s, e2e_full_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_s, e2e_full_s = t.tolist()
    n_items = ds.N
    chains.close()
    ds.close()

    if rank == 0:
        n_step = n_chains * n_items  # reassignments per step per GPU
        total = world * n_step * args.steps
        value = total / (elapsed_ms * 1e-3)
        e2e_value = world * n_step * args.e2e_steps / e2e_s
        k_ms = float(np.mean(kernel_ms))
        fp32_peak = mc_fp32_peak(ctx)
        hbm_peak, hbm_src = measured_peaks()
        # algorithmic work of a step = sum over its reassignments of (K_i + m) (D^2 + 4D + 3 + 6) flops (the kernels
        # keep the exact sum (K_i + m)) and 2 sizeof(z) + 4 D / 32 bytes of HBM per reassignment
        alg_flops_step = (cand / args.steps) * (f_eval(DIM) + 6)
        alg_bytes_step = n_step * (2 * 2 + 4 * DIM / 32.0)
        fp32_roofline = {"bound": "fp32", "achieved": alg_flops_step / (k_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": alg_flops_step / (k_ms * 1e-3) / 1e12 / fp32_peak if fp32_peak else None,
                         "peak_source": "FP32 FMA peak measured in this run by npb_fp32_peak (better of the scalar FFMA and the packed "
                                        "FFMA2 instruction streams; MEASURED_PEAKS.json has no FP32 figure)",
                         "note": "algorithmic flops of a step / the whole step time"}
        kernel = "k_alg8_sweep_tile4<%d,3>" % DIM if (DIM >= 4 and kmax == 32) else ("k_alg8_sweep_tile" if DIM >= 4 else "k_alg8_sweep_reg")
        launches = args.steps * (3 if (DIM >= 4 and kmax == 32) else 2)
        if tc_path:
            # D = 16, Kmax = 32: per block of steps k_pre_aimg16, k_pre_bimg16, k_gather_z and k_sweep_tc16 (tcgen05 kind::f16
            # quadratic forms, epilogue and race fused; npb_alg8_fused16.cu), once per sweep k_scan_order and k_aux_bound.
            # `roofline` is the dominant kernel's: ALGORITHMIC flops per launch / its average launch duration (CUDA events on the
            # library's stream around every launch of it inside the timed region) against the measured bf16 peak -- the FP16x3
            # split issues 5.7 flops per algorithmic flop, reported as `issued`.
            bs = int(os.environ.get("NPB_D16_BLOCK", "8192"))
            blocks = (n_items + bs - 1) // bs
            launches = args.steps * (2 + 4 * blocks)
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            f16_peak = peaks.get("bf16_tflops")
            peak_src = "measured bf16 burst peak (MEASURED_PEAKS.json; kind::f16 runs at the bf16 rate)"
            if not f16_peak:
                f16_peak, peak_src = 2250.0, "fallback: nominal dense bf16 peak of B200 (B200_PROFILING.md); MEASURED_PEAKS.json missing"
            traffic, traffic_src = None, "missing: profiles/r2_traffic.json has no entry for this configuration"
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).get(args.config)
                if tj and tj["chains"] == n_chains and tj["block"] == bs:
                    traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
            except Exception:
                pass
            if kt_n:
                lps = kt_n / args.steps
                launch_ms = kt_ms / kt_n
                issued = n_step * 32 * 4 * 2 * DIM * DIM / lps  # kind::f16 flops per launch: 32 slots x 4 K-steps x 2 x 16 x 16 per reassignment
                roofline = {"bound": "tensor", "achieved": alg_flops_step / lps / (launch_ms * 1e-3) / 1e12, "peak": f16_peak, "unit": "TFLOP/s",
                            "frac": alg_flops_step / lps / (launch_ms * 1e-3) / 1e12 / f16_peak, "traffic": traffic,
                            "traffic_unit": "bytes per launch (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)",
                            "traffic_source": traffic_src, "kernel": "k_sweep_tc16<3,0>", "launches_per_step": lps,
                            "launch_ms": launch_ms, "share_of_step": kt_ms / elapsed_ms,
                            "algorithmic_flops_per_launch": alg_flops_step / lps, "algorithmic_bytes_per_launch": alg_bytes_step / lps,
                            "issued": {"flops_per_launch": issued, "achieved": issued / (launch_ms * 1e-3) / 1e12,
                                       "frac": issued / (launch_ms * 1e-3) / 1e12 / f16_peak,
                                       "note": "kind::f16 MMA flops the kernel issues: three FP16 products per FP32 product plus the folded "
                                               "offsets = 4 K-steps of 16 per (step, slot row)"},
                            "peak_source": peak_src,
                            "note": "achieved = SURVEY 8(d) algorithmic flops, sum(K_i + m) (D^2 + 4D + 3 + 6) from the kernel's own counter, per "
                                    "launch / the kernel's average launch duration measured with CUDA events on its stream"}
            else:  # the handle ran another kernel set (NPB_D16_PATH=tc2): whole-step figure
                roofline = dict(fp32_roofline, bound="tensor", peak=f16_peak, frac=alg_flops_step / (k_ms * 1e-3) / 1e12 / f16_peak,
                                kernel="k_density_tc16 + k_race (round 1 kernel pair)", traffic=None, peak_source=peak_src)
            fp32_roofline = dict(fp32_roofline, kernel="k_sweep_tc16<3,0> (+ pre-pass kernels)")
        else:
            roofline = dict(fp32_roofline, kernel=kernel, traffic=None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "chains_per_gpu": n_chains, "N": n_items, "D": DIM, "Kmax": kmax,
                       "l2": "per-GPU assignment state %d MB (+ %d MB of slot tables) exceeds the 126 MB L2; no flush needed"
                             % (n_items * n_chains * 2 // 2 ** 20, n_chains * kmax * (DIM + DIM * (DIM + 1) // 2 + 1) * 4 // 2 ** 20),
                       "mean_K": last.mean_K, "max_K": last.max_K, "candidates_per_reassignment": cand / (n_step * args.steps),
                       "moved_fraction": moved / (n_step * args.steps), "new_clusters_per_step": births / args.steps,
                       "regime": "stationary: the chains sit at the given clusters and next to no item moves (the fused kernel's best case); "
                                 "the same shape with chains that keep mixing is also.mixing",
                       "path": ("tensor, NPB_D16_PATH=%s (auto: fused kernel k_sweep_tc16 while few items move, table + race kernel pair "
                                "in a mixing chain; same assignments either way)" % d16_path) if tc_path else "fp32 pipe"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_items * DIM * 8),
                    "d2h_bytes_per_step": int(8 + 6 * changed / args.e2e_steps), "steps": args.e2e_steps,
                    "call": "npb_chains_sweep_host_delta (X up from pinned host memory; the caller's page-locked copy of every chain's "
                            "assignments brought up to date: the changed entries travel as an (index, slot) list compacted on the device)",
                    "changed_entries_per_step": changed / args.e2e_steps, "mirror_checked": mirror_ok,
                    "full_copy": {"value": world * n_step * args.e2e_steps / e2e_full_s, "unit": UNIT,
                                  "d2h_bytes_per_step": int(n_items * n_chains * 2),
                                  "call": "npb_chains_sweep_host (every assignment down every step: round 1's figure)"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes_step / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes_step / (k_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                             "note": "SURVEY 8(d) algorithmic bytes of a step (6 B per reassignment) / the step time: the path is not HBM-bound"},
            "diagnostics": {"mean_purity": diag["mean_purity"], "mean_rand": diag["mean_rand"], "mean_ari": diag["mean_ari"],
                            "mean_K": diag["mean_K"], "chains": diag["chains"], "rhat": diag["rhat"], "cocluster": cocl,
                            "allreduce": "nccl" if world > 1 else "none (1 rank)"},
        }
        if tc_path:
            line["roofline_fp32_equivalent"] = fp32_roofline
        if world == 1 and not args.no_also and args.config == "cfg5":
            try:
                line["also"] = also_measure(npb, syn, ctx, fp32_peak, rank, cpu=not args.no_cpu_baseline)
            except Exception as e:
                line["also"] = {"failed": repr(e)}
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline_sample(args.config, os.cpu_count() or 1)
            except Exception as e:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def cocluster_measure(npb, chains, ds, torch, dist, world, rank=0, n_anchor=4096):
    """posterior co-clustering counts of an anchor subset over this rank's chains (k_cc_gather + k_cc_tile), summed over the ranks
    by ncclAllReduce INSIDE the library (npb_cocluster_allreduce; the communicator is the library's own, its unique id handed round
    through torch.distributed) -- the path's only exchange; returns the figures for the JSON line"""
    anchors = np.arange(0, ds.N, max(1, ds.N // n_anchor))[:n_anchor]
    A = len(anchors)
    comm = None
    if world > 1:
        uid = npb.Comm.unique_id(chains.ctx) if rank == 0 else bytes(128)
        t = torch.tensor(list(uid), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        comm = npb.Comm(chains.ctx, bytes(t.cpu().tolist()), rank, world)
    S = torch.zeros((A, A), dtype=torch.float32, device="cuda")
    chains.cocluster_allreduce(anchors, S.data_ptr(), comm)  # warm-up (NCCL sets its rings up on the first call)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    chains.cocluster_into(anchors, S.data_ptr())
    chains.ctx.synchronize()
    t_k = time.perf_counter() - t0
    t0 = time.perf_counter()
    chains.cocluster_allreduce(anchors, S.data_ptr(), comm)
    chains.ctx.synchronize()
    t_all = time.perf_counter() - t0
    if comm is not None:
        comm.close()
    pairs = A * (A + 1) / 2
    return {"anchors": int(A), "chains_per_gpu": int(chains.C), "kernel_ms": 1e3 * t_k, "with_allreduce_ms": 1e3 * t_all,
            "allreduce_ms": max(0.0, 1e3 * (t_all - t_k)), "allreduce_bytes": int(A * A * 4),
            "byte_compares_per_s": pairs * chains.C / t_k, "anchor_diag_mean": float(S.diag().mean().item()),
            "expected_diag": int(chains.C) * world}


def mc_fp32_peak(ctx):
    """FP32 FMA peak of this GPU, TFLOP/s (register-resident FMA loops, scalar and packed, best of 10, CUDA events)"""
    import ctypes as C
    lib = ctx._lib
    if not hasattr(lib, "npb_fp32_peak"):
        return None
    lib.npb_fp32_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    out = C.c_double()
    st = lib.npb_fp32_peak(ctx._h, C.byref(out))
    return out.value if st == 0 else None


if __name__ == "__main__":
    main()
