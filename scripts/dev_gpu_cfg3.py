"""Developer check under gpurun: split-merge proposal throughput in the config-3 shape (16-D, N = 100k, K = 32 given clusters)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import noparama_b200 as npb
from noparama_b200 import synthetic as syn
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nprop = int(sys.argv[2]) if len(sys.argv) > 2 else 16
X, y = syn.config(3)
K = int(y.max()) + 1
means = np.stack([X[y == k].mean(0) for k in range(K)])
Sigma = np.tile(np.eye(X.shape[1]), (K, 1, 1))
ctx = npb.Context(0)
ds = npb.Dataset(ctx, X)
npb.NormalInverseWishart(**syn.reference_prior(X.shape[1])).bind(ctx)
ch = npb.Chains(ctx, ds, chains, Kmax=64, K0=20, seed=3)
ch.init_from_params(means, Sigma)
ch.sweep(npb.ALG8, 1)  # items to their clusters
for name, sampler in (("jain_neal", npb.JAIN_NEAL), ("triadic", npb.TRIADIC)):
    for it in range(3):
        st = ch.split_merge(sampler, nprop)
        s = st.kernel_ms * 1e-3
        print(name, it, "ms %.1f proposals %d (%.3e/s) SAMS allocations %.3e/s attempts %s accepts %s meanK %.2f" % (
            st.kernel_ms, st.reassignments, st.reassignments / s, st.sams_allocations / s, list(st.sm_attempts), list(st.sm_accepts), st.mean_K), flush=True)
