"""ctypes binding of the CPU oracle (oracle/np_oracle.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (noparama_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ALG8, JAIN_NEAL, TRIADIC = 8, 2, 3
DENSE_MATRIX, PER_CALL_LU, UPDATE_CLUSTERS, MAX_LIKELIHOOD, RECORD_TRACE, LOG_DOMAIN = (1 << i for i in range(6))
FAITHFUL = DENSE_MATRIX | PER_CALL_LU | UPDATE_CLUSTERS | MAX_LIKELIHOOD


class Prior(C.Structure):
    _fields_ = [("D", C.c_int), ("mu0", C.POINTER(C.c_double)), ("kappa", C.c_double), ("nu", C.c_double),
                ("Lambda", C.POINTER(C.c_double)), ("alpha", C.c_double)]


class Options(C.Structure):
    _fields_ = [("algorithm", C.c_int), ("T", C.c_int), ("K0", C.c_int), ("M_aux", C.c_int), ("mh_steps", C.c_int),
                ("seed_main", C.c_uint32), ("seed_shuffle", C.c_uint32), ("flags", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("seconds_total", C.c_double), ("seconds_reassign", C.c_double), ("updates", C.c_int64),
                ("density_evals", C.c_int64), ("candidates", C.c_int64), ("new_cluster_events", C.c_int64),
                ("moved", C.c_int64), ("sm_attempts", C.c_int64 * 4), ("sm_accepts", C.c_int64 * 4),
                ("sams_allocations", C.c_int64), ("mean_K", C.c_double), ("K_final", C.c_int),
                ("max_loglik", C.c_double)]


def build(force=False):
    if os.environ.get("NP_ORACLE_LIB"):  # e.g. the sanitizer build (make -C oracle sanitize; scripts/oracle_sanitize.sh)
        return os.environ["NP_ORACLE_LIB"]
    so = os.path.join(_HERE, "libnp_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("np_oracle.cpp", "np_oracle.h", "np_oracle_sm.inc", "np_oracle_alg2.inc")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "libnp_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        for name in ("npo_mvn_pdf", "npo_mvn_logpdf"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 3
        for name in ("npo_mvn_pdf_dataset", "npo_mvn_logpdf_dataset"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 3 + [C.c_int]
        L.npo_mvn_logpdf_batch.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                                           C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        L.npo_weighted_pick_u.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double]
        L.npo_weighted_pick_freq.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_int64)]
        L.npo_sample_base.argtypes = [C.POINTER(Prior), C.c_uint32, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.npo_lu_determinant.restype = C.c_double
        L.npo_lu_determinant.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.npo_lu_inverse.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.npo_metrics.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double)]
        L.npo_membertrix_selftest.argtypes = [C.c_uint32, C.c_int]
        L.npo_mcmc_run.restype = C.c_void_p
        L.npo_mcmc_run.argtypes = [C.POINTER(Prior), C.POINTER(Options), C.POINTER(C.c_double), C.c_int]
        L.npo_mcmc_run_given.restype = C.c_void_p
        L.npo_mcmc_run_given.argtypes = [C.POINTER(Prior), C.POINTER(Options), C.POINTER(C.c_double), C.c_int, C.c_int,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.npo_run_free.argtypes = [C.c_void_p]
        L.npo_run_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.npo_run_assignments.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.npo_run_params.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_int64), C.c_int]
        L.npo_run_sweep_seconds.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.npo_run_init_K.argtypes = [C.c_void_p]
        L.npo_run_init_state.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double)]
        L.npo_run_K_after_len.restype = C.c_int64
        L.npo_run_K_after_len.argtypes = [C.c_void_p]
        L.npo_run_K_after.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.npo_sm_trace_proposals.restype = C.c_int64
        L.npo_sm_trace_proposals.argtypes = [C.c_void_p]
        L.npo_sm_trace_pool_len.restype = C.c_int64
        L.npo_sm_trace_pool_len.argtypes = [C.c_void_p]
        L.npo_sm_trace_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 13
        L.npo_trace_steps.restype = C.c_int64
        L.npo_trace_steps.argtypes = [C.c_void_p]
        L.npo_trace_order_len.restype = C.c_int64
        L.npo_trace_order_len.argtypes = [C.c_void_p]
        L.npo_trace_max_slot.argtypes = [C.c_void_p]
        L.npo_trace_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 10
        _LIB = L
    return _LIB


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def mvn_pdf(mu, Sigma, x):
    mu, Sigma, x = _f64(mu), _f64(Sigma), _f64(x)
    return lib().npo_mvn_pdf(len(mu), _dp(mu), _dp(Sigma), _dp(x))


def mvn_logpdf(mu, Sigma, x):
    mu, Sigma, x = _f64(mu), _f64(Sigma), _f64(x)
    return lib().npo_mvn_logpdf(len(mu), _dp(mu), _dp(Sigma), _dp(x))


def mvn_pdf_dataset(mu, Sigma, X):
    mu, Sigma, X = _f64(mu), _f64(Sigma), _f64(X)
    return lib().npo_mvn_pdf_dataset(len(mu), _dp(mu), _dp(Sigma), _dp(X), X.shape[0])


def mvn_logpdf_dataset(mu, Sigma, X):
    mu, Sigma, X = _f64(mu), _f64(Sigma), _f64(X)
    return lib().npo_mvn_logpdf_dataset(len(mu), _dp(mu), _dp(Sigma), _dp(X), X.shape[0])


def mvn_logpdf_batch(mu, Sigma, X):
    """mu [K,D], Sigma [K,D,D], X [n,D] -> [n,K]"""
    mu, Sigma, X = _f64(mu), _f64(Sigma), _f64(X)
    K, D = mu.shape
    out = np.empty((X.shape[0], K), dtype=np.float64)
    lib().npo_mvn_logpdf_batch(D, _dp(mu), _dp(Sigma), K, _dp(X), X.shape[0], _dp(out))
    return out


def weighted_pick_u(w, u):
    w = _f64(w)
    return lib().npo_weighted_pick_u(_dp(w), len(w), float(u))


def weighted_pick_freq(w, draws, seed):
    w = _f64(w)
    freq = np.zeros(len(w), dtype=np.int64)
    lib().npo_weighted_pick_freq(_dp(w), len(w), draws, seed, freq.ctypes.data_as(C.POINTER(C.c_int64)))
    return freq


def make_prior(mu0, kappa, nu, Lambda, alpha):
    mu0, Lambda = _f64(mu0), _f64(Lambda)
    p = Prior(len(mu0), _dp(mu0), kappa, nu, _dp(Lambda), alpha)
    p._keep = (mu0, Lambda)
    return p


def sample_base(prior, seed, count):
    D = prior.D
    mu = np.empty((count, D))
    Sigma = np.empty((count, D, D))
    lib().npo_sample_base(C.byref(prior), seed, count, _dp(mu), _dp(Sigma))
    return mu, Sigma


def lu_determinant(A):
    A = _f64(A)
    return lib().npo_lu_determinant(A.shape[0], _dp(A))


def lu_inverse(A):
    A = _f64(A)
    out = np.empty_like(A)
    lib().npo_lu_inverse(A.shape[0], _dp(A), _dp(out))
    return out


def metrics(truth, result):
    truth = np.ascontiguousarray(truth, dtype=np.int32)
    result = np.ascontiguousarray(result, dtype=np.int32)
    out = np.zeros(3)
    lib().npo_metrics(_ip(truth), _ip(result), len(truth), _dp(out))
    return tuple(out)


def membertrix_selftest(seed, dense=True):
    return lib().npo_membertrix_selftest(seed, int(dense))


class Run:
    """One oracle run of MCMC::run (np_mcmc.cpp:48-175)."""

    def __init__(self, prior, X, algorithm=ALG8, T=1000, K0=20, M_aux=3, mh_steps=20, seed_main=1, seed_shuffle=2,
                 flags=FAITHFUL, given=None):
        X = _f64(X)
        self.N, self.D = X.shape
        self.M_aux = M_aux
        self.T = T
        opt = Options(algorithm, T, K0, M_aux, mh_steps, seed_main, seed_shuffle, flags)
        if given is None:
            self._h = lib().npo_mcmc_run(C.byref(prior), C.byref(opt), _dp(X), self.N)
        else:  # NOT the reference: start from caller-supplied clusters (mu [K,D], Sigma [K,D,D])
            mu, Sigma = _f64(given[0]), _f64(given[1])
            self._h = lib().npo_mcmc_run_given(C.byref(prior), C.byref(opt), _dp(X), self.N, mu.shape[0], _dp(mu), _dp(Sigma))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().npo_run_free(self._h)
            self._h = None

    def stats(self):
        s = Stats()
        lib().npo_run_stats(self._h, C.byref(s))
        return s

    def sweep_seconds(self):
        a, b = np.empty(self.T), np.empty(self.T)
        lib().npo_run_sweep_seconds(self._h, _dp(a), _dp(b))
        return a, b

    def K_after(self):
        """cluster count after every sampler.update() call (RECORD_TRACE runs)"""
        k = np.empty(lib().npo_run_K_after_len(self._h), dtype=np.int32)
        lib().npo_run_K_after(self._h, _ip(k))
        return k

    def assignments(self, which=0):
        z = np.empty(self.N, dtype=np.int32)
        lib().npo_run_assignments(self._h, which, _ip(z))
        return z

    def params(self, cap=4096):
        K = C.c_int()
        mu = np.empty((cap, self.D))
        Sigma = np.empty((cap, self.D, self.D))
        counts = np.empty(cap, dtype=np.int64)
        rc = lib().npo_run_params(self._h, C.byref(K), _dp(mu), _dp(Sigma), counts.ctypes.data_as(C.POINTER(C.c_int64)), cap)
        assert rc == 0
        return mu[:K.value].copy(), Sigma[:K.value].copy(), counts[:K.value].copy()

    def init_state(self):
        K = lib().npo_run_init_K(self._h)
        z0 = np.empty(self.N, dtype=np.int32)
        slots = np.empty(K, dtype=np.int32)
        mu = np.empty((K, self.D))
        Sigma = np.empty((K, self.D, self.D))
        lib().npo_run_init_state(self._h, _ip(z0), _ip(slots), _dp(mu), _dp(Sigma))
        return z0, slots, mu, Sigma

    def sm_trace(self):
        """split-merge replay trace (RECORD_TRACE runs of JAIN_NEAL / TRIADIC), see np_oracle.h"""
        n = lib().npo_sm_trace_proposals(self._h)
        L = lib().npo_sm_trace_pool_len(self._h)
        D = self.D
        t = dict(picks=np.empty((n, 3), np.int32), u0=np.empty(n), type=np.empty(n, np.int32), th_mu=np.empty((n, D)),
                 th_sigma=np.empty((n, D, D)), pool_off=np.empty(n + 1, np.int64), pool=np.empty(L, np.int32),
                 us=np.empty(L), dec=np.empty(L, np.int32), logA=np.empty(n), uacc=np.empty(n),
                 accept=np.empty(n, np.int32), new_slot=np.empty(n, np.int32))
        lib().npo_sm_trace_copy(self._h, *[t[k].ctypes.data for k in
                                           ("picks", "u0", "type", "th_mu", "th_sigma", "pool_off", "pool", "us", "dec",
                                            "logA", "uacc", "accept", "new_slot")])
        z = np.empty((self.T, self.N), np.int32)
        tr = self.trace()
        t["z_after"] = tr["z_after"]
        t["max_slot"] = tr["max_slot"]
        return t

    def trace(self):
        S = lib().npo_trace_steps(self._h)
        OL = lib().npo_trace_order_len(self._h)
        M, D = self.M_aux, self.D
        t = dict(item=np.empty(S, np.int32), K=np.empty(S, np.int32), order_off=np.empty(S + 1, np.int64),
                 order=np.empty(OL, np.int32), aux_mu=np.empty((S, M, D)), aux_Sigma=np.empty((S, M, D, D)),
                 u=np.empty(S), picked=np.empty(S, np.int32), new_slot=np.empty(S, np.int32),
                 z_after=np.empty((self.T, self.N), np.int32))
        lib().npo_trace_copy(self._h, *[t[k].ctypes.data for k in
                                        ("item", "K", "order_off", "order", "aux_mu", "aux_Sigma", "u", "picked",
                                         "new_slot", "z_after")])
        t["max_slot"] = lib().npo_trace_max_slot(self._h)
        return t


# ---- conjugate Algorithm 2 (np_oracle_alg2.inc; not in the reference: pinned against scipy.stats.multivariate_t) ----
def niw_logpred(prior, Xm, x, incremental=False, remove_first=0):
    """log posterior-predictive density of x for the cluster made of the rows of Xm (none: the prior predictive)"""
    L = lib()
    Xm = _f64(np.asarray(Xm, dtype=np.float64).reshape(-1, prior.D))
    x = _f64(x)
    if incremental:
        L.npo_niw_logpred_incremental.restype = C.c_double
        L.npo_niw_logpred_incremental.argtypes = [C.POINTER(Prior), C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        return L.npo_niw_logpred_incremental(C.byref(prior), len(Xm), _dp(Xm), remove_first, _dp(x))
    L.npo_niw_logpred.restype = C.c_double
    L.npo_niw_logpred.argtypes = [C.POINTER(Prior), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return L.npo_niw_logpred(C.byref(prior), len(Xm), _dp(Xm), _dp(x))


def alg2_run(prior, X, T, K0, seed):
    """T sweeps of conjugate Algorithm 2 -> (z [N] compact labels, K after every sweep [T], moved, births)"""
    L = lib()
    X = _f64(X)
    N = len(X)
    z = np.empty(N, dtype=np.int32)
    Kt = np.empty(T, dtype=np.int32)
    moved, births = C.c_int64(), C.c_int64()
    L.npo_alg2_run.argtypes = [C.POINTER(Prior), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_int),
                               C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.npo_alg2_run(C.byref(prior), _dp(X), N, T, K0, seed, _ip(z), _ip(Kt), C.byref(moved), C.byref(births))
    return z, Kt, moved.value, births.value


# ---- scalar-noise families (`-c regression` / `-c angular`): scalarnoise_multivariatenormal.cpp + normalinvgamma.h ----
REGRESSION, ANGULAR = 1, 2


def _sn_setup():
    L = lib()
    if getattr(L, "_sn_ready", False):
        return L
    dp = C.POINTER(C.c_double)
    for name in ("npo_scalarnoise_logpdf", "npo_scalarnoise_pdf"):
        getattr(L, name).restype = C.c_double
        getattr(L, name).argtypes = [C.c_int, dp, C.c_double, dp]
    L.npo_sample_base_nig.argtypes = [dp, dp, C.c_double, C.c_double, C.c_uint32, C.c_int, dp, dp]
    L.npo_mcmc_run_scalarnoise.restype = C.c_void_p
    L.npo_mcmc_run_scalarnoise.argtypes = [C.c_int, dp, dp, C.c_double, C.c_double, C.c_double, C.POINTER(Options), dp, C.c_int]
    L.npo_run_params_scalarnoise.argtypes = [C.c_void_p, C.POINTER(C.c_int), dp, dp, C.POINTER(C.c_int64), C.c_int]
    L._sn_ready = True
    return L


def scalarnoise_logpdf(family, mu, sigma, x, log=True):
    L = _sn_setup()
    mu, x = _f64(mu), _f64(x)
    return (L.npo_scalarnoise_logpdf if log else L.npo_scalarnoise_pdf)(family, _dp(mu), float(sigma), _dp(x))


def scalarnoise_logpdf_batch(family, mu, sigma, X):
    """[n, K] log-densities of the rows of X under the K parameter sets (mu [K,2], sigma [K])"""
    mu, X = _f64(mu), _f64(X)
    return np.array([[scalarnoise_logpdf(family, mu[k], sigma[k], X[i]) for k in range(len(mu))] for i in range(len(X))])


def sample_base_nig(prior, seed, count):
    """count draws of (mu [2], sigma) from the normal-inverse-gamma base measure; prior: dict(mu0, Lambda, nig_alpha, nig_beta)"""
    L = _sn_setup()
    mu0, Lam = _f64(prior["mu0"]), _f64(prior["Lambda"])
    mu, sg = np.empty((count, 2)), np.empty(count)
    L.npo_sample_base_nig(_dp(mu0), _dp(Lam), float(prior["nig_alpha"]), float(prior["nig_beta"]), seed, count, _dp(mu), _dp(sg))
    return mu, sg


class ScalarNoiseRun(Run):
    """One oracle run of MCMC::run with the scalar-noise likelihood; X: the rows read_data builds ((1, a, b) | (a, b))."""

    def __init__(self, family, prior, X, algorithm=ALG8, T=1000, K0=20, M_aux=3, mh_steps=20, seed_main=1, seed_shuffle=2,
                 flags=FAITHFUL):
        L = _sn_setup()
        X = _f64(X)
        self.N, self.D = X.shape
        assert self.D == (3 if family == REGRESSION else 2)
        self.M_aux, self.T = M_aux, T
        opt = Options(algorithm, T, K0, M_aux, mh_steps, seed_main, seed_shuffle, flags)
        mu0, Lam = _f64(prior["mu0"]), _f64(prior["Lambda"])
        self._h = L.npo_mcmc_run_scalarnoise(family, _dp(mu0), _dp(Lam), float(prior["nig_alpha"]), float(prior["nig_beta"]),
                                             float(prior["alpha"]), C.byref(opt), _dp(X), self.N)
        assert self._h

    def params(self, cap=4096):
        L = _sn_setup()
        K = C.c_int()
        mu, sg, cnt = np.empty((cap, 2)), np.empty(cap), np.empty(cap, dtype=np.int64)
        assert L.npo_run_params_scalarnoise(self._h, C.byref(K), _dp(mu), _dp(sg), cnt.ctypes.data_as(C.POINTER(C.c_int64)), cap) == 0
        return mu[:K.value].copy(), sg[:K.value].copy(), cnt[:K.value].copy()
