"""Quick timing of the mixing regime at the headline shape (bench.py's also.mixing) on the kernel pair; GPU box only."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import noparama_b200 as npb
from noparama_b200 import synthetic as syn
ctx = npb.Context(0)
chains = int(os.environ.get("MIX_CHAINS", 8192))
D, K, N = 16, 32, 100_000
X, y = syn.gmm_mixing(N, D, K, 20261005)
ds = npb.Dataset(ctx, X)
npb.NormalInverseWishart(**syn.reference_prior(D)).bind(ctx)
for path, spec in (("tc2", "1"), ("tc2", "0")) if os.environ.get("MIX_SEQ") else (("tc2", "1"),):
    ch = npb.Chains(ctx, ds, chains, Kmax=32, K0=bench.K0_REF, m_aux=bench.M_AUX, seed=bench.SEED)
    ch.set_option("d16_path", path)
    ch.set_option("spec", spec)
    ch.set_option("time_kernels", "1")
    ch.init_from_params(*bench.given_clusters(X, y))
    for _ in range(2):
        ch.sweep(npb.ALG8, 1)
    ch.kernel_time()
    ms = []
    for _ in range(3):
        st = ch.sweep(npb.ALG8, 1)
        ms.append(st.kernel_ms)
    kt = ch.kernel_time()
    print(json.dumps({"path": path, "spec": spec, "ms": ms, "moved": st.moved / (chains * N), "dominant_kernel_ms_launches": kt}), flush=True)
    ch.close()
