"""Generates the committed fixtures under tests/golden/ with the CPU oracle (oracle/np_oracle.cpp).

The reference itself cannot be run in this image (Eigen is absent), so the fixtures are outputs of the
oracle, which is pinned to the reference's known-answer test.  Run from the repo root:

    python tests/golden/make_golden.py
"""
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from noparama_b200 import synthetic as syn  # noqa: E402
from oracle import binding as orc  # noqa: E402

N_SEEDS = 256
T_CFG1 = 1000


def cfg1_seed(seed):
    X, y = syn.config(1)
    pr = orc.make_prior(**syn.reference_prior(2))
    r = orc.Run(pr, X, T=T_CFG1, seed_main=10_000 + seed, seed_shuffle=20_000 + seed,
                flags=orc.UPDATE_CLUSTERS | orc.MAX_LIKELIHOOD)
    s = r.stats()
    fin = orc.metrics(y, r.assignments(0))
    best = orc.metrics(y, r.assignments(1))
    return fin + best + (s.K_final, s.mean_K, s.moved / s.updates, s.new_cluster_events / s.updates, s.max_loglik)


T_SM = {orc.JAIN_NEAL: 100, orc.TRIADIC: 300}


def cfg1_sm_seed(args):
    """split-merge samplers on config 1 (oracle/np_oracle_sm.inc, pinned to the reference's own code by
    tests/test_ref_pin.py)"""
    alg, seed = args
    X, y = syn.config(1)
    pr = orc.make_prior(**syn.reference_prior(2))
    r = orc.Run(pr, X, alg, T=T_SM[alg], seed_main=30_000 + seed, seed_shuffle=40_000 + seed,
                flags=orc.UPDATE_CLUSTERS | orc.MAX_LIKELIHOOD)
    s = r.stats()
    return orc.metrics(y, r.assignments(0)) + (s.K_final, s.mean_K, s.updates, s.sams_allocations) + \
        tuple(s.sm_attempts) + tuple(s.sm_accepts)


def density_cases():
    out = {}
    for D in (2, 16, 64):
        rng = np.random.default_rng(900 + D)
        X, _ = syn.gmm(48, D, 4, 77 + D)
        pr = syn.reference_prior(D)
        mu_p, Sig_p = orc.sample_base(orc.make_prior(**pr), 5 + D, 6)
        # general full covariances as well: a parameter update "done right" would produce them
        mu_g = X[rng.integers(0, len(X), 6)] + 0.3 * rng.standard_normal((6, D))
        B = rng.standard_normal((6, D, D)) / np.sqrt(D)
        Sig_g = B @ np.transpose(B, (0, 2, 1)) + 0.5 * np.eye(D)
        mu = np.concatenate([mu_p, mu_g])
        Sigma = np.concatenate([Sig_p, Sig_g])
        out["X%d" % D] = X
        out["mu%d" % D] = mu
        out["Sigma%d" % D] = Sigma
        out["logp%d" % D] = orc.mvn_logpdf_batch(mu, Sigma, X)
    return out


def main():
    with Pool(os.cpu_count()) as pool:
        res = np.array(pool.map(cfg1_seed, range(N_SEEDS)))
    np.savez_compressed(os.path.join(HERE, "oracle_cfg1_alg8_256seeds.npz"),
                        purity=res[:, 0], rand=res[:, 1], ari=res[:, 2], purity_maxlik=res[:, 3], rand_maxlik=res[:, 4],
                        ari_maxlik=res[:, 5], K_final=res[:, 6], mean_K=res[:, 7], moved=res[:, 8], births=res[:, 9],
                        max_loglik=res[:, 10], T=T_CFG1)
    np.savez_compressed(os.path.join(HERE, "density_cases.npz"), **density_cases())
    for alg, name in ((orc.JAIN_NEAL, "jain_neal"), (orc.TRIADIC, "triadic")):
        with Pool(os.cpu_count()) as pool:
            r = np.array(pool.map(cfg1_sm_seed, [(alg, s) for s in range(N_SEEDS)]), dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, "oracle_cfg1_%s_256seeds.npz" % name), purity=r[:, 0], rand=r[:, 1],
                            ari=r[:, 2], K_final=r[:, 3], mean_K=r[:, 4], updates=r[:, 5], sams=r[:, 6],
                            attempts=r[:, 7:11], accepts=r[:, 11:15], T=T_SM[alg])
        print("%s: K %.2f purity %.4f ARI %.4f attempts %s accepts %s" % (name, r[:, 3].mean(), r[:, 0].mean(), r[:, 2].mean(),
              r[:, 7:11].sum(0), r[:, 11:15].sum(0)))
    print("cfg1 over %d seeds: purity %.4f  RI %.4f  ARI %.4f  K %.2f  moved %.3f  births/step %.5f" %
          (N_SEEDS, res[:, 0].mean(), res[:, 1].mean(), res[:, 2].mean(), res[:, 6].mean(), res[:, 8].mean(), res[:, 9].mean()))


if __name__ == "__main__":
    main()
