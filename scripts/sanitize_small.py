"""Small invocation of every kernel family, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py all
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py fused
(logs kept under profiles/).  Sizes are tiny: the sanitizer slows kernels by 10-100x."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noparama_b200 as npb
from noparama_b200 import synthetic as syn


def given(X, y):
    K = int(y.max()) + 1
    return np.stack([X[y == k].mean(0) for k in range(K)]), np.tile(np.eye(X.shape[1]), (K, 1, 1))


def main(which):
    ctx = npb.Context(0)
    if which in ("all", "d2"):
        X, y = syn.config(1)
        ds = npb.Dataset(ctx, X)
        mc = npb.MCMC(ctx, ds, npb.NormalInverseWishart(**syn.reference_prior(2)), chains=8, Kmax=64, seed=1)
        mc.run(3)
        for s in (npb.JAIN_NEAL, npb.TRIADIC):
            mc.chains.sweep(s, 1)
        mc.chains.metrics(y)
        mc.chains.update_params(npb.UPDATE_POSTERIOR_DRAW, syn.reference_prior(2))
        print("d2 ok")
        mc.chains.close()
        ds.close()
    if which in ("all", "fused", "pair", "fp32"):
        X, y = syn.gmm(600, 16, 4, 3)
        ds = npb.Dataset(ctx, X)
        npb.NormalInverseWishart(**syn.reference_prior(16)).bind(ctx)
        for path in (("tc", "tc2", "fp32") if which == "all" else ({"fused": "tc", "pair": "tc2", "fp32": "fp32"}[which],)):
            ch = npb.Chains(ctx, ds, 5, Kmax=32, K0=6, seed=2)
            ch.set_option("d16_path", path)
            for _ in range(2):
                ch.sweep(npb.ALG8, 1)  # reference initialisation: births and deaths
            ch.init_from_params(*given(X, y))
            st = ch.sweep(npb.ALG8, 2)
            mirror = np.zeros((ds.N, 5), np.uint16)
            ch.sweep_host_delta(None, npb.ALG8, 1, mirror)
            ch.cocluster(np.arange(0, 600, 7))
            print("16-D path", path, "ok, moved", st.moved)
            ch.close()
        npb.NormalInverseWishart(mu0=X.mean(0), kappa=0.01, nu=18.0, Lambda=np.eye(16), alpha=1.0).bind(ctx)
        if which == "all":
            ch = npb.Chains(ctx, ds, 3, Kmax=32, K0=6, seed=4)
            ch.init_from_params(*given(X, y))
            ch.sweep(npb.ALG2_CONJUGATE, 2)
            ch.alg2_logpred(1, np.arange(8, dtype=np.int32))
            print("conjugate 16-D ok")
            ch.close()
        ds.close()
    if which in ("all", "d64"):
        X, y = syn.gmm(400, 64, 3, 5)
        ds = npb.Dataset(ctx, X)
        npb.NormalInverseWishart(**syn.reference_prior(64)).bind(ctx)
        ch = npb.Chains(ctx, ds, 3, Kmax=32, K0=4, m_aux=1, seed=6)
        ch.init_from_params(*given(X, y))
        ch.sweep(npb.ALG8, 2)
        print("64-D ok")
        ch.close()
        npb.NormalInverseWishart(mu0=X.mean(0), kappa=0.01, nu=66.0, Lambda=np.eye(64), alpha=1.0).bind(ctx)
        ch = npb.Chains(ctx, ds, 2, Kmax=32, K0=4, seed=7)
        ch.init_from_params(*given(X, y))
        ch.sweep(npb.ALG2_CONJUGATE, 1)
        print("conjugate 64-D ok")
        ch.close()
        ds.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "all")
