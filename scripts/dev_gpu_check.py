"""Developer check run under gpurun: smoke + a first throughput look at config 2."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.smoke()
import noparama_b200 as npb
from noparama_b200 import synthetic as syn
X, y = syn.config(2)
ctx = npb.Context(0)
ds = npb.Dataset(ctx, X)
prior = npb.NormalInverseWishart(**syn.reference_prior(2))
for chains, kmax in ((1024, 256), (1024, 512), (8192, 256)):
    mc = npb.MCMC(ctx, ds, prior, chains=chains, Kmax=kmax, seed=3)
    for it in range(6):
        try:
            st = mc.chains.sweep(npb.ALG8, 1)
        except npb.NpbError as e:
            print("ERR", e); break
        print(chains, kmax, it, "ms %.2f" % st.kernel_ms, "rate %.3e" % (st.reassignments / (st.kernel_ms * 1e-3)),
              "meanK %.1f maxK %d cand/step %.1f moved %.3f births %d" % (st.mean_K, st.max_K, st.candidates / st.reassignments,
              st.moved / st.reassignments, st.new_clusters), flush=True)
    m = mc.chains.metrics(y)
    print("purity %.3f ri %.3f ari %.3f K %.1f jll %.1f" % (m["purity"].mean(), m["rand_index"].mean(), m["adjusted_rand"].mean(), m["K"].mean(), m["joint_loglik"].mean()))
    mc.chains.close()
