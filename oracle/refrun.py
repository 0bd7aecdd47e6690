"""Runs oracle/_ref/np_ref_run: the reference's OWN sampler sources (mrquincle/noparama) compiled against
oracle/eigen_shim by oracle/Makefile (target _ref).

TEST INFRASTRUCTURE ONLY: used by tests/ (to pin oracle/np_oracle.cpp to the real reference code) and by bench.py's
cpu_baseline / --impl reference legs.  The product package (noparama_b200/) never imports it.
The binary is built in the dev container (where /root/reference exists) and travels to the GPU box with the snapshot;
nothing here reads /root/reference at run time.
"""
import os
import struct
import subprocess
import tempfile
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BINARY = os.path.join(_HERE, "_ref", "np_ref_run")
ALGORITHMS = {8: "algorithm8", 2: "jain_neal_split", 3: "triadic"}


def available():
    return os.path.exists(BINARY) and os.access(BINARY, os.X_OK)


def build():
    """(Re)build oracle/_ref when the reference tree is present; returns availability."""
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", _HERE, "_ref"], stdout=subprocess.DEVNULL)
    return available()


def write_request(path, X, prior):
    X = np.ascontiguousarray(X, dtype=np.float64)
    N, D = X.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", N, D))
        # the scalar-noise families (regression / angular) send the inverse-gamma's (alpha, beta) in the kappa, nu fields
        k, n = (prior["kappa"], prior["nu"]) if "kappa" in prior else (prior["nig_alpha"], prior["nig_beta"])
        f.write(struct.pack("<ddd", float(k), float(n), float(prior["alpha"])))
        f.write(np.ascontiguousarray(prior["mu0"], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(prior["Lambda"], dtype=np.float64).tobytes())
        f.write(X.tobytes())


def read_result(path):
    with open(path, "rb") as f:
        b = f.read()
    N, T, calls, s_run, s_upd, K_final, n_snap = struct.unpack_from("<iiqddii", b, 0)
    off = struct.calcsize("<iiqddii")
    z_final = np.frombuffer(b, np.int32, N, off); off += 4 * N
    z_maxlik = np.frombuffer(b, np.int32, N, off); off += 4 * N
    (nk,) = struct.unpack_from("<q", b, off); off += 8
    K_after = np.frombuffer(b, np.int32, nk, off); off += 4 * nk
    snaps = np.frombuffer(b, np.int32, n_snap * N, off).reshape(n_snap, N)
    off += 4 * n_snap * N
    (ns,) = struct.unpack_from("<q", b, off); off += 8
    sweep_cum = np.frombuffer(b, np.float64, ns, off)
    return dict(sweep_update_seconds=np.diff(np.concatenate([[0.0], sweep_cum])), N=N, T=T, calls=calls, seconds_run=s_run, seconds_update=s_upd, K_final=K_final, z_final=z_final.copy(),
                z_maxlik=z_maxlik.copy(), K_after=K_after.copy(), z_snaps=snaps.copy())


def command(algorithm, T, seed_main, seed_shuffle, request, result, record=False, family="clustering"):
    return [BINARY, ALGORITHMS[algorithm], str(T), str(seed_main), str(seed_shuffle), request, result, "1" if record else "0", family]


def run(X, prior, algorithm=8, T=1000, seed_main=1, seed_shuffle=2, record=False, timeout=None, family="clustering"):
    """One run of the reference's MCMC::run (np_mcmc.cpp:48-175) in a fresh process (its static distributions are
    process-wide state).  prior: dict(mu0, kappa, nu, Lambda, alpha); for family "regression" / "angular" (np_main.cpp -c):
    dict(mu0[2], Lambda[2,2], nig_alpha, nig_beta, alpha) and X the rows read_data builds ((1, a, b) resp. (a, b))."""
    if not available():
        raise RuntimeError("oracle/_ref/np_ref_run is not built (make -C oracle _ref needs /root/reference)")
    with tempfile.TemporaryDirectory() as d:
        req, res = os.path.join(d, "req.bin"), os.path.join(d, "res.bin")
        write_request(req, X, prior)
        # (the reference's binary is not instrumented: a sanitizer runtime preloaded for the oracle, scripts/oracle_sanitize.sh, stays out of it)
        env = {k: v for k, v in os.environ.items() if k != "LD_PRELOAD"}
        subprocess.run(command(algorithm, T, seed_main, seed_shuffle, req, res, record, family), check=True,
                       stdout=subprocess.DEVNULL, timeout=timeout, env=env)
        return read_result(res)
