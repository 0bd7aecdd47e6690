"""Developer check under gpurun: the D = 16 sweep paths at the headline shape, stationary and mixing regimes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import noparama_b200 as npb
from noparama_b200 import synthetic as syn

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
regime = sys.argv[2] if len(sys.argv) > 2 else "stationary"
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
D = 16
if regime == "stationary":
    X, y = syn.config(5)
else:
    X, y = syn.gmm_mixing(100_000, D, 32, 20261005)
K = int(y.max()) + 1
means = np.stack([X[y == k].mean(0) for k in range(K)])
Sigma = np.tile(np.eye(D), (K, 1, 1))
ctx = npb.Context(0)
ds = npb.Dataset(ctx, X)
mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(D)), chains=chains, Kmax=32, K0=20, seed=3)
mc.chains.init_from_params(means, Sigma)
if os.environ.get("TK"):
    mc.chains.set_option("time_kernels", "1")
for it in range(sweeps):
    st = mc.chains.sweep(npb.ALG8, 1)
    if os.environ.get("TK"):
        print("   sweep-kernel launches: ms, n =", mc.chains.kernel_time())
    print(os.environ.get("NPB_D16_PATH", "fused"), regime, chains, it, "ms %.2f rate %.3e meanK %.1f cand/step %.1f moved %.6f births %d" % (
        st.kernel_ms, st.reassignments / (st.kernel_ms * 1e-3), st.mean_K, st.candidates / st.reassignments,
        st.moved / st.reassignments, st.new_clusters), flush=True)
m = mc.chains.metrics(y)
print("purity %.4f ari %.4f K %.2f" % (m["purity"].mean(), m["adjusted_rand"].mean(), m["K"].mean()))
