"""Quick A/B of the conjugate Algorithm 2 kernels (k_a2_tile against k_a2_sweep) at the bench shapes; GPU box only."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noparama_b200 as npb
from noparama_b200 import synthetic as syn

def given(X, y):
    K = int(y.max()) + 1
    return np.stack([X[y == k].mean(0) for k in range(K)]), np.tile(np.eye(X.shape[1]), (K, 1, 1))

ctx = npb.Context(0)
shapes = [("64d", 64, int(os.environ.get("A2_N64", 100000)), 256, 16), ("16d", 16, 100000, 1024, 20)]
for name, D, N, chains, K in [sh for sh in shapes if sh[0] in os.environ.get("A2_SHAPES", "64d,16d").split(",")]:
    X, y = syn.gmm(N, D, K, 20261004)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(mu0=X.mean(0), kappa=0.01, nu=D + 2.0, Lambda=np.eye(D), alpha=1.0).bind(ctx)
    for tile in [int(t) for t in os.environ.get("A2_TILES", "64,16").split(",")]:
        if tile == 0 and N > 20000 and D == 64:
            continue
        ch = npb.Chains(ctx, ds, chains, Kmax=32, K0=K, seed=1234)
        ch.set_option("a2_tile", str(tile))
        ch.set_option("a2_tc", os.environ.get("A2_TC", "1"))
        ch.set_option("a2_tc16", os.environ.get("A2_TC16", "0"))
        if os.environ.get("A2_TRUTH"):
            ch.set_state(0, y.astype(np.int32), np.arange(K, dtype=np.int32), *given(X, y))
            ch.broadcast_state(0)
        else:
            ch.init_from_params(*given(X, y))
        out = []
        for i in range(4):
            st = ch.sweep(npb.ALG2_CONJUGATE, 1)
            out.append((round(st.kernel_ms, 1), round(st.moved / (chains * N), 4), round(st.mean_K, 2)))
        m = ch.metrics(y)
        print(json.dumps({"shape": name, "N": N, "tile": tile, "sweeps(ms, moved, K)": out, "rate": chains * N / (out[-1][0] * 1e-3),
                          "purity": float(m["purity"].mean())}), flush=True)
        ch.close()
    ds.close()
