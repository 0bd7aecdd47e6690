"""CPU: the N>1 host logic (chain sharding, per-rank seeds, all-reduce of diagnostics) under gloo, world size 2."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from noparama_b200 import diagnostics as dg  # noqa: E402

TOTAL_CHAINS, T, ANCHORS = 37, 50, 12


def fake_rank_data(lo, hi):
    """deterministic per-chain diagnostics for chains [lo, hi): what a rank would read back from its device"""
    rng = [np.random.default_rng(1000 + c) for c in range(lo, hi)]
    purity = np.array([r.uniform(0.8, 1.0) for r in rng])
    ri = np.array([r.uniform(0.5, 0.9) for r in rng])
    ari = np.array([r.uniform(0.1, 0.6) for r in rng])
    K = np.array([r.integers(4, 15) for r in rng], dtype=np.int32)
    traces = np.stack([r.normal(c % 3, 1.0, T) for r, c in zip(rng, range(lo, hi))])
    z = np.stack([r.integers(0, 4, ANCHORS) for r in rng])
    S = (z[:, :, None] == z[:, None, :]).sum(0).astype(np.float32)
    return dict(purity=purity, rand_index=ri, adjusted_rand=ari, K=K), traces, S


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = dg.shard_chains(TOTAL_CHAINS, world, rank)
    m, traces, S = fake_rank_data(lo, hi)
    St = torch.from_numpy(S.copy())
    out = dg.combine(dg.score_partial(m), {"K_trace": dg.rhat_partial(traces)}, cocluster=St)
    q.put((rank, lo, hi, out, St.numpy(), dg.rank_seed(5, rank)))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_chains_partition():
    for total in (1, 7, 8, 1024, 65536):
        for world in (1, 2, 3, 8):
            spans = [dg.shard_chains(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert len({dg.rank_seed(1, r) for r in range(8)}) == 8


def test_rhat_matches_direct_formula():
    rng = np.random.default_rng(0)
    tr = rng.normal(0, 1, (8, 200)) + rng.normal(0, 0.5, (8, 1))
    W = tr.var(axis=1, ddof=1).mean()
    B = 200 * tr.mean(axis=1).var(ddof=1)
    want = np.sqrt(((199 / 200) * W + B / 200) / W)
    assert np.isclose(dg.rhat_from_partial(dg.rhat_partial(tr)), want)
    same = np.tile(rng.normal(0, 1, 200), (4, 1))
    assert dg.rhat_from_partial(dg.rhat_partial(same)) < 1.0 + 1e-9


@pytest.mark.timeout(120)
def test_gloo_world2_diagnostics_match_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # single-process ground truth over all chains
    m, traces, S = fake_rank_data(0, TOTAL_CHAINS)
    want = dg.combine(dg.score_partial(m), {"K_trace": dg.rhat_partial(traces)})
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == TOTAL_CHAINS
    assert res[0][5] != res[1][5]
    for _, _, _, out, St, _ in res:
        assert out["chains"] == TOTAL_CHAINS
        for k in ("mean_purity", "mean_rand", "mean_ari", "mean_K"):
            assert np.isclose(out[k], want[k], rtol=1e-12)
        assert np.isclose(out["rhat"]["K_trace"], want["rhat"]["K_trace"], rtol=1e-10)
        assert np.array_equal(St, S)  # co-clustering counts add up exactly across ranks
