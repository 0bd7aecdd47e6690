"""Developer check under gpurun: throughput of the D=16 tile kernel in the config-5 regime (K = 32 given clusters)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import noparama_b200 as npb
from noparama_b200 import synthetic as syn
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
kmax = int(sys.argv[2]) if len(sys.argv) > 2 else 32
X, y = syn.config(5)
K = int(y.max()) + 1
means = np.stack([X[y == k].mean(0) for k in range(K)])
Sigma = np.tile(np.eye(X.shape[1]), (K, 1, 1))
ctx = npb.Context(0)
ds = npb.Dataset(ctx, X)
mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(X.shape[1])), chains=chains, Kmax=kmax, K0=20, seed=3)
mc.chains.init_from_params(means, Sigma)
for it in range(5):
    st = mc.chains.sweep(npb.ALG8, 1)
    D = X.shape[1]
    fl = st.candidates * (D * D + 4 * D + 3 + 6)
    print(chains, kmax, it, "ms %.2f rate %.3e meanK %.1f cand/step %.1f moved %.6f births %d TFLOPs %.2f" % (
        st.kernel_ms, st.reassignments / (st.kernel_ms * 1e-3), st.mean_K, st.candidates / st.reassignments,
        st.moved / st.reassignments, st.new_clusters, fl / (st.kernel_ms * 1e-3) / 1e12), flush=True)
m = mc.chains.metrics(y)
print("purity %.4f ari %.4f K %.2f" % (m["purity"].mean(), m["adjusted_rand"].mean(), m["K"].mean()))
