"""Developer check under gpurun: BASELINE configs[3] shape (256 chains x N = 1M, D = 64, 32 given clusters, one auxiliary
draw = Algorithm 2; or m = 3) through the D = 64 path, for a few block sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import noparama_b200 as npb
from noparama_b200 import synthetic as syn
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
blocks = [int(b) for b in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4096]
m_aux = int(sys.argv[4]) if len(sys.argv) > 4 else 1
nsweeps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
t0 = time.time()
X, y = syn.gmm(N, 64, 32, syn.SEEDS[4])
K = int(y.max()) + 1
means = np.stack([X[y == k].mean(0) for k in range(K)])
Sigma = np.tile(np.eye(64), (K, 1, 1))
print("data %.1fs" % (time.time() - t0), flush=True)
ctx = npb.Context(0)
ds = npb.Dataset(ctx, X)
npb.NormalInverseWishart(**syn.reference_prior(64)).bind(ctx)
for bs in blocks:
    os.environ["NPB_D64_BLOCK"] = str(bs)
    ch = npb.Chains(ctx, ds, chains, Kmax=32, K0=20, m_aux=m_aux, seed=3)
    ch.init_from_params(means, Sigma)
    for it in range(nsweeps):
        st = ch.sweep(npb.ALG2 if m_aux == 1 else npb.ALG8, 1)
        s = st.kernel_ms * 1e-3
        print("block %d sweep %d: %.1f ms, %.3e reassignments/s, candidates/step %.2f moved %.4f births %d meanK %.2f" % (
            bs, it, st.kernel_ms, st.reassignments / s, st.candidates / st.reassignments, st.moved / st.reassignments,
            st.new_clusters, st.mean_K), flush=True)
    m = ch.metrics(y)
    print("purity %.5f ARI %.5f K %.2f" % (m["purity"].mean(), m["adjusted_rand"].mean(), m["K"].mean()), flush=True)
    ch.close()
