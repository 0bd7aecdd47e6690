import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, noparama_b200 as npb
from noparama_b200 import synthetic as syn
X, y = syn.config(5)
K = 32
means = np.stack([X[y == k].mean(0) for k in range(K)])
ctx = npb.Context(0); ds = npb.Dataset(ctx, X)
npb.NormalInverseWishart(**syn.reference_prior(16)).bind(ctx)
ch = npb.Chains(ctx, ds, 8192, Kmax=32, seed=3)
ch.init_from_params(means, np.tile(np.eye(16), (K, 1, 1)))
ch.sweep(npb.ALG8, 1)
for it in range(3):
    ctx.synchronize(); t0 = time.perf_counter(); m = ch.metrics(y); ctx.synchronize()
    print("metrics ms %.1f purity %.4f jll %.6e" % ((time.perf_counter() - t0) * 1e3, m["purity"].mean(), m["joint_loglik"].mean()))
