#!/usr/bin/env python
"""bench.py -- item-reassignments/sec of the Gibbs reassignment path (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1] -- 1024 lockstep Algorithm-8 chains per GPU over one synthetic
10-component 2-D GMM with N = 100 000 items (noparama_b200/synthetic.py, seed 20261002), reference prior
(np_main.cpp:164,367-371), m = 3 auxiliary draws, K0 = 20.  A "step" is one sweep: every chain reassigns every
item once (np_mcmc.cpp:146-163) = chains * N item-reassignments, one kernel launch per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
  python bench.py --impl reference [--gpus N] ...              the reference algorithm on the host cores (CPU oracle)

Under torchrun every rank owns one GPU and 1024 chains of its own (chains are independent: weak scaling, no
collective on the data path); timing is barrier + synchronize on both sides, CUDA events on the library's stream,
max over ranks.  After the timed region the ranks all-reduce (NCCL) the posterior co-clustering matrix of an anchor
subset and the per-chain diagnostics, the only exchange the path has.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

N_ITEMS, DIM, K_TRUE = 100_000, 2, 10
CHAINS_PER_GPU = 1024
KMAX, M_AUX, K0 = 256, 3, 20
SEED = 20261018
METRIC = "item-reassignments/sec"
UNIT = "reassignments/s"
WORKLOAD = ("BASELINE configs[1]: %d lockstep Algorithm-8 chains per GPU, synthetic %d-component %d-D GMM, N=%d, "
            "m=%d aux, K0=%d, reference NIW prior (bug-compatible)" % (CHAINS_PER_GPU, K_TRUE, DIM, N_ITEMS, M_AUX, K0))


def f_eval(D):
    """algorithmic flops of one density evaluation, SURVEY 8d: D^2 + 4D + 3"""
    return D * D + 4 * D + 3


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
def oracle_chain_worker(args):
    """one independent reference chain on one host core (the reference is single threaded)"""
    seed, sweeps, faithful, n_items = args
    from oracle import binding as orc
    from noparama_b200 import synthetic as syn
    X, _ = syn.config(2)
    X = X[:n_items]
    pr = orc.make_prior(**syn.reference_prior(DIM))
    flags = orc.FAITHFUL if faithful else 0
    r = orc.Run(pr, X, T=sweeps, K0=K0, M_aux=M_AUX, seed_main=1000 + seed, seed_shuffle=2000 + seed, flags=flags)
    re, tot = r.sweep_seconds()
    s = r.stats()
    return re.tolist(), tot.tolist(), s.candidates / max(1, s.updates), s.mean_K


def run_oracle(cores, sweeps, faithful=True, n_items=N_ITEMS):
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(cores) as pool:
        return pool.map(oracle_chain_worker, [(c, sweeps, faithful, n_items) for c in range(cores)])


def cpu_baseline_sample(cores):
    """bounded sample of the same workload: `cores` independent chains, 1 warm-up + 2 timed sweeps each
    (the first sweep starts from the random initial assignment and is not representative)"""
    res = run_oracle(cores, 3, True)
    t_reassign = max(sum(r[0][1:]) for r in res)
    t_total = max(sum(r[1][1:]) for r in res)
    n = cores * N_ITEMS * 2
    return {"value": n / t_reassign, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d independent chains (one per host core), sweeps 2-3 of the same N=%d config, update() loop only "
                      "(np_mcmc.cpp:146-163) with the reference cost profile (per-call LU inverse+determinant, density "
                      "evaluated twice, dense bool matrix)" % (cores, N_ITEMS),
            "full_sweep_value": n / t_total, "single_core_value": N_ITEMS * 2 / max(sum(r[0][1:]) for r in res[:1]),
            "candidates_per_reassignment": float(np.mean([r[2] for r in res]))}


def reference_arm(args, rank, world):
    """--impl reference: the reference algorithm (CPU oracle port; the reference itself needs Eigen, which this image
    lacks) on all host cores, one independent chain per core, same config / metric / unit."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sweeps = args.warmup + args.steps
    # bounded sample: a faithful sweep of N = 100k costs ~3 s per chain; keep the whole run within a few minutes by
    # sweeping a prefix of the (shuffled) items when many steps are asked for
    n_items = N_ITEMS if sweeps <= 40 else max(2000, N_ITEMS * 40 // sweeps)
    t0 = time.time()
    res = run_oracle(cores, sweeps, True, n_items)
    wall = time.time() - t0
    t = max(sum(r[0][args.warmup:]) for r in res)
    t_full = max(sum(r[1][args.warmup:]) for r in res)
    n = cores * n_items * args.steps
    val = n / t
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chains": cores, "items_per_sweep": n_items,
                       "note": "one reference chain per host core; a step = one sweep of every chain over "
                               "items_per_sweep items of the config (a prefix when W+K > 40)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d chains x %d sweeps x N=%d, update() loop only" % (cores, args.steps, n_items),
                             "full_sweep_value": n / t_full},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="npb200", choices=["npb200", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU, help="chains per GPU")
    ap.add_argument("--kmax", type=int, default=KMAX)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import noparama_b200 as npb
    from noparama_b200 import synthetic as syn

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    X, y = syn.config(2)
    ctx = npb.Context(local_rank)
    ds = npb.Dataset(ctx, X)
    prior = npb.NormalInverseWishart(**syn.reference_prior(DIM))
    mc = npb.MCMC(ctx, ds, prior, chains=args.chains, Kmax=args.kmax, K0=K0, m_aux=M_AUX, seed=__import__('noparama_b200').diagnostics.rank_seed(SEED, rank))
    chains = mc.chains
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- warm-up (also burns the chains in past the random initial assignment) ----
    for _ in range(args.warmup):
        chains.sweep(npb.ALG8, 1, want_stats=False)
    # ---- timed region: exactly K sweeps, inputs resident in HBM ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, cand, moved, births = [], 0, 0, 0
    ev0.record(stream)
    for _ in range(args.steps):
        st = chains.sweep(npb.ALG8, 1)
        kernel_ms.append(st.kernel_ms)
        cand += st.candidates
        moved += st.moved
        births += st.new_clusters
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    last = st

    # ---- end to end through the public call with host buffers (H2D of X, D2H of every assignment) ----
    Xh = np.ascontiguousarray(X)
    z_host = np.empty((ds.N, args.chains), dtype=np.uint16)
    chains.sweep_host(Xh, npb.ALG8, 1, z_out=z_host)  # warm the staging buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        chains.sweep_host(Xh, npb.ALG8, 1, z_out=z_host)
    ctx.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---- diagnostics exchange (outside the timed region): a short traced phase for R-hat, then one all-reduce of
    # the co-clustering counts, the score sums and the R-hat partials over all ranks (NCCL when world > 1) ----
    from noparama_b200 import diagnostics as dg
    k_trace, jll_trace = [], []
    for _ in range(8):
        chains.sweep(npb.ALG8, 1, want_stats=False)
        mm = chains.metrics(None)
        k_trace.append(mm["K"].astype(np.float64))
        jll_trace.append(mm["joint_loglik"])
    m = chains.metrics(y)
    anchors = np.arange(0, ds.N, ds.N // 256)[:256]
    S = torch.zeros((len(anchors), len(anchors)), dtype=torch.float32, device="cuda")
    chains.cocluster_into(anchors, S.data_ptr())
    diag = dg.combine(dg.score_partial(m), {"K": dg.rhat_partial(np.stack(k_trace, 1)),
                                            "joint_loglik": dg.rhat_partial(np.stack(jll_trace, 1))},
                      cocluster=S, device="cuda")
    t = torch.tensor([elapsed_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_s = t.tolist()

    if rank == 0:
        n_step = args.chains * ds.N  # reassignments per step per GPU
        total = world * n_step * args.steps
        value = total / (elapsed_ms * 1e-3)
        e2e_value = world * n_step * args.e2e_steps / e2e_s
        # roofline of the dominant (only) kernel, per launch, from this rank's counters and CUDA-event times
        flops_per_launch = (cand / args.steps) * (f_eval(DIM) + 6)
        k_ms = float(np.mean(kernel_ms))
        fp32_peak = mc_fp32_peak(ctx)
        hbm_peak, hbm_src = measured_peaks()
        bytes_per_launch = n_step * (2 * 2 + 4 * DIM / 32.0)  # z read+write (u16) + x shared by the 32 lanes' prefetch
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chains_per_gpu": args.chains, "N": ds.N, "D": DIM, "Kmax": args.kmax,
                       "l2": "per-GPU assignment state %d MB + slot tables exceed the 126 MB L2; no flush needed"
                             % (ds.N * args.chains * 2 // 2 ** 20),
                       "mean_K": last.mean_K, "max_K": last.max_K, "candidates_per_reassignment": cand / (n_step * args.steps),
                       "moved_fraction": moved / (n_step * args.steps), "new_clusters_per_step": births / args.steps},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ds.N * DIM * 8),
                    "d2h_bytes_per_step": int(ds.N * args.chains * 2), "steps": args.e2e_steps,
                    "call": "npb_chains_sweep_host (pinned staging; X up, all assignments down)"},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": flops_per_launch / (k_ms * 1e-3) / 1e12, "peak": fp32_peak,
                         "unit": "TFLOP/s", "frac": flops_per_launch / (k_ms * 1e-3) / 1e12 / fp32_peak if fp32_peak else None,
                         "traffic": None, "kernel": "k_alg8_sweep", "kernel_ms": k_ms,
                         "peak_source": "FP32 FFMA peak measured in this run by npb_fp32_peak (MEASURED_PEAKS.json has no "
                                        "FP32 figure); algorithmic flops = sum(K_i+m) * (D^2+4D+3+6), counter kept by the kernel"},
            "roofline_hbm": {"bound": "hbm", "achieved": bytes_per_launch / (k_ms * 1e-3) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": bytes_per_launch / (k_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                             "peak_source": hbm_src},
            "diagnostics": {"mean_purity": diag["mean_purity"], "mean_rand": diag["mean_rand"], "mean_ari": diag["mean_ari"],
                            "mean_K": diag["mean_K"], "chains": diag["chains"], "rhat": diag["rhat"],
                            "cocluster_anchor_diag_mean": float(S.diag().mean().item()),
                            "allreduce": "nccl" if world > 1 else "none (1 rank)"},
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline_sample(os.cpu_count() or 1)
            except Exception as e:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def mc_fp32_peak(ctx):
    """FP32 FFMA peak of this GPU, TFLOP/s (register-resident FMA loop, best of 10, CUDA events)"""
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
