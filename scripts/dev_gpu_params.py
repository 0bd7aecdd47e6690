"""Developer check under gpurun: cost of the parameter update at the headline shape, and a sweep+update loop."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import noparama_b200 as npb
from noparama_b200 import synthetic as syn
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
X, y = syn.config(5)
K = int(y.max()) + 1
means = np.stack([X[y == k].mean(0) for k in range(K)])
ctx = npb.Context(0)
ds = npb.Dataset(ctx, X)
npb.NormalInverseWishart(**syn.reference_prior(16)).bind(ctx)
ch = npb.Chains(ctx, ds, chains, Kmax=32, seed=3)
ch.init_from_params(means + 0.5, np.tile(2.0 * np.eye(16), (K, 1, 1)))
pr = dict(mu0=X.mean(0), kappa=0.01, nu=18.0, Lambda=np.eye(16))
for it in range(4):
    st = ch.sweep(npb.ALG8, 1)
    ctx.synchronize(); t0 = time.perf_counter()
    ch.update_params(npb.UPDATE_POSTERIOR_DRAW, pr)
    ctx.synchronize(); dt = time.perf_counter() - t0
    m = ch.metrics(y)
    print(it, "sweep ms %.1f moved %.4f update ms %.1f purity %.4f K %.2f" % (st.kernel_ms, st.moved / st.reassignments, dt * 1e3, m["purity"].mean(), m["K"].mean()), flush=True)
