"""Short run of the bench workload for ncu: config 2, 1024 chains, 3 warm-up sweeps + 2 profiled sweeps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noparama_b200 as npb
from noparama_b200 import synthetic as syn
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
kmax = int(sys.argv[2]) if len(sys.argv) > 2 else 256
X, y = syn.config(2)
ctx = npb.Context(0)
ds = npb.Dataset(ctx, X)
mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(2)), chains=chains, Kmax=kmax, seed=3)
for it in range(5):
    st = mc.chains.sweep(npb.ALG8, 1)
    print(it, "ms %.2f rate %.3e meanK %.1f maxK %d" % (st.kernel_ms, st.reassignments / (st.kernel_ms * 1e-3), st.mean_K, st.max_K), flush=True)
