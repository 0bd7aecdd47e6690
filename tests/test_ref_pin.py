"""CPU: pins the Eigen-free oracle (oracle/np_oracle.cpp) to the reference's OWN sampler code.

oracle/_ref/np_ref_run is the reference's unmodified translation units (np_mcmc.cpp, np_neal_algorithm8.cpp,
membertrix.cpp, multivariatenormal.cpp, ...) compiled where they lie against oracle/eigen_shim (oracle/Makefile, target
_ref).  Both sides use the same libstdc++ <random> objects, so with the two random_device seeds fixed the oracle must
reproduce the reference's partition after EVERY sweep, label for label at the end (final and max-likelihood states).
"""
import numpy as np
import pytest

from noparama_b200 import synthetic as syn
from oracle import refrun

pytestmark = pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built (needs /root/reference at build time)")


def same_partition(a, b):
    pairs = set(zip(a.tolist(), b.tolist()))
    return len(pairs) == len(set(a.tolist())) == len(set(b.tolist()))


def test_reference_unit_tests_pass_on_the_shim():
    # the reference's own tests (test/*.cpp), unmodified, built by the same recipe
    import os
    import subprocess
    import tempfile
    d = os.path.dirname(refrun.BINARY)
    with tempfile.TemporaryDirectory() as cwd:
        out = subprocess.run([os.path.join(d, "test_mvn_likelihood")], capture_output=True, text=True, cwd=cwd)
        assert out.returncode == 0 and "0.061975" in out.stdout and "0.0038409" in out.stdout  # KAT, test_mvn_likelihood.cpp:33,44
        for t in ("test_membertrix", "test_weighted_vector", "test_multivariate_normal_distribution"):
            assert subprocess.run([os.path.join(d, t)], capture_output=True, cwd=cwd).returncode == 0, t


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_alg8_trajectory_matches_reference(oracle, seed):
    X, _ = syn.twogaussians()
    pr = syn.reference_prior(2)
    T = 120
    ref = refrun.run(X, pr, 8, T=T, seed_main=seed, seed_shuffle=1000 + seed, record=True)
    run = oracle.Run(oracle.make_prior(**pr), X, oracle.ALG8, T=T, seed_main=seed, seed_shuffle=1000 + seed,
                     flags=oracle.FAITHFUL | oracle.RECORD_TRACE)
    tr = run.trace()
    assert ref["calls"] == T * len(X) == run.stats().updates
    assert ref["z_snaps"].shape == tr["z_after"].shape
    for t in range(T):
        assert same_partition(tr["z_after"][t], ref["z_snaps"][t]), t
    assert np.array_equal(run.assignments(0), ref["z_final"])   # same labels: same hash-map iteration order (Q9)
    # considerMaxLikelihood picks the same sweep (labels differ: the driver's extra copy relabels once more)
    assert same_partition(run.assignments(1), ref["z_maxlik"])
    assert run.stats().K_final == ref["K_final"]


def test_alg8_trajectory_matches_reference_3d(oracle):
    # generalised D (the reference's reader is 2-D only, its classes are not): 3-D, other prior scale
    rng = np.random.default_rng(5)
    X = np.concatenate([rng.standard_normal((60, 3)), rng.standard_normal((60, 3)) + 4.0])
    pr = syn.reference_prior(3)
    ref = refrun.run(X, pr, 8, T=40, seed_main=9, seed_shuffle=10, record=True)
    run = oracle.Run(oracle.make_prior(**pr), X, oracle.ALG8, T=40, seed_main=9, seed_shuffle=10,
                     flags=oracle.FAITHFUL | oracle.RECORD_TRACE)
    assert np.array_equal(run.assignments(0), ref["z_final"])
    assert same_partition(run.assignments(1), ref["z_maxlik"])


@pytest.mark.parametrize("alg,seed", [(2, 1), (2, 2), (3, 1), (3, 2), (3, 3)])
def test_split_merge_trajectory_matches_reference(oracle, alg, seed):
    """Jain-Neal (np_jain_neal_algorithm.cpp) and triadic (np_triadic_algorithm.cpp) restatements in
    oracle/np_oracle_sm.inc: same number of proposals (Q11 collisions skipped alike), same cluster count after EVERY
    proposal, same final and max-likelihood partitions as the reference's own code."""
    X, _ = syn.twogaussians()
    pr = syn.reference_prior(2)
    T = 30
    ref = refrun.run(X, pr, alg, T=T, seed_main=seed, seed_shuffle=500 + seed, record=True)
    run = oracle.Run(oracle.make_prior(**pr), X, {2: oracle.JAIN_NEAL, 3: oracle.TRIADIC}[alg], T=T, seed_main=seed,
                     seed_shuffle=500 + seed, flags=oracle.FAITHFUL | oracle.RECORD_TRACE)
    st = run.stats()
    assert ref["calls"] == st.updates and st.updates < T * len(X)  # some subsets collide (np_mcmc.cpp:155-158)
    assert np.array_equal(run.K_after(), ref["K_after"])
    assert sum(st.sm_accepts) > 0
    assert same_partition(run.assignments(0), ref["z_final"])
    assert same_partition(run.assignments(1), ref["z_maxlik"])
    assert st.K_final == ref["K_final"]
