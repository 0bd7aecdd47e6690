"""Cross-rank diagnostics: the only exchange the reassignment path has.

Chains are sharded over ranks (one process per GPU) and never communicate while sweeping.  What is combined after
(or between) sweeps: the posterior co-clustering counts of an anchor subset, sums of the per-chain clustering scores
(purity / Rand / adjusted Rand of clustering_performance.cpp:38-82) and the sufficient statistics of the Gelman-Rubin
R-hat of per-chain scalar traces.  All of it reduces with SUM, so one `all_reduce` per tensor over NCCL (GPU) or gloo
(CPU tests) is the whole protocol.  torch.distributed is plumbing here: it must be initialised by the caller.
"""
import numpy as np


def shard_chains(total_chains, world_size, rank):
    """Contiguous block partition of chain ids: rank r owns [lo, hi).  Sizes differ by at most one."""
    base, rem = divmod(int(total_chains), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(seed, rank):
    """Philox key stride between ranks: chain c of rank r uses key (seed + 7919 r, c) -- disjoint streams."""
    return int(seed) + 7919 * int(rank)


def rhat_partial(traces):
    """Per-rank sufficient statistics of R-hat for traces [chains, T]: (m, sum mean, sum mean^2, sum var, T)."""
    traces = np.asarray(traces, dtype=np.float64)
    m, T = traces.shape
    means = traces.mean(axis=1)
    var = traces.var(axis=1, ddof=1) if T > 1 else np.zeros(m)
    return np.array([m, means.sum(), (means ** 2).sum(), var.sum(), T], dtype=np.float64)


def rhat_from_partial(p):
    """Gelman-Rubin potential scale reduction from summed partials (the T entry must be the common length)."""
    m, s1, s2, sv, T = p
    if m < 2 or T < 2:
        return float("nan")
    W = sv / m
    B_over_T = (s2 - s1 * s1 / m) / (m - 1)     # variance of the chain means
    var_plus = (T - 1) / T * W + B_over_T
    return float(np.sqrt(var_plus / W)) if W > 0 else float("nan")


def score_partial(metrics):
    """sums of purity / Rand / adjusted Rand / K over this rank's chains + chain count (dict from Chains.metrics)."""
    return np.array([metrics["purity"].sum(), metrics["rand_index"].sum(), metrics["adjusted_rand"].sum(),
                     float(metrics["K"].sum()), float(len(metrics["K"]))], dtype=np.float64)


def all_reduce_sum(tensors, group=None):
    """SUM-all-reduce a list of torch tensors in place over the default (or given) process group; no-op when
    torch.distributed is not initialised (single process)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors


def combine(score_sums, rhat_partials, cocluster=None, device="cpu", group=None):
    """All-reduce the diagnostics of this rank with every other rank.

    score_sums: array from score_partial; rhat_partials: dict name -> array from rhat_partial (T entry is kept, not
    summed); cocluster: optional torch tensor [A, A] of co-clustering counts living on `device` (reduced in place).
    Returns dict(mean_purity, mean_rand, mean_ari, mean_K, chains, rhat={name: value})."""
    import torch
    names = sorted(rhat_partials)
    T = {k: rhat_partials[k][4] for k in names}
    flat = np.concatenate([np.asarray(score_sums, dtype=np.float64)] + [rhat_partials[k][:4] for k in names])
    t = torch.from_numpy(flat.copy()).to(device)
    tensors = [t] + ([cocluster] if cocluster is not None else [])
    all_reduce_sum(tensors, group)
    flat = t.cpu().numpy()
    s = flat[:5]
    out = dict(mean_purity=s[0] / s[4], mean_rand=s[1] / s[4], mean_ari=s[2] / s[4], mean_K=s[3] / s[4], chains=int(s[4]),
               rhat={})
    for i, k in enumerate(names):
        p = flat[5 + 4 * i: 9 + 4 * i]
        out["rhat"][k] = rhat_from_partial(np.append(p, T[k]))
    return out
