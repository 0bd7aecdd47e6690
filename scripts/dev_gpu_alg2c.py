"""Developer check under gpurun: throughput of the conjugate Algorithm 2 path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import noparama_b200 as npb
from noparama_b200 import synthetic as syn
D = int(sys.argv[1]); N = int(sys.argv[2]); chains = int(sys.argv[3]); K = int(sys.argv[4]) if len(sys.argv) > 4 else 32
X, y = syn.gmm(N, D, K, 20261004)
ctx = npb.Context(0)
ds = npb.Dataset(ctx, X)
npb.NormalInverseWishart(mu0=X.mean(0), kappa=0.01, nu=D + 2.0, Lambda=np.eye(D), alpha=1.0).bind(ctx)
ch = npb.Chains(ctx, ds, chains, Kmax=32, K0=K, seed=3)
means = np.stack([X[y == k].mean(0) for k in range(K)])
ch.init_from_params(means, np.tile(np.eye(D), (K, 1, 1)))
ch.sweep(npb.ALG8 if D != 64 else npb.ALG2, 1) if False else None
for it in range(4):
    st = ch.sweep(npb.ALG2_CONJUGATE, 1)
    print(D, N, chains, it, "ms %.1f rate %.3e meanK %.1f cand/step %.1f moved %.4f births %d" % (
        st.kernel_ms, st.reassignments / (st.kernel_ms * 1e-3), st.mean_K, st.candidates / st.reassignments, st.moved / st.reassignments, st.new_clusters), flush=True)
m = ch.metrics(y)
print("purity %.4f K %.2f" % (m["purity"].mean(), m["K"].mean()))
