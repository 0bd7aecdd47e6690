"""Host-side mirror (Python) of the reference's sampler seam on top of the C ABI of libnpb200.so.

The reference wires `MCMC(generator, InitClusters, UpdateClusters, UpdateClusterPopulation, subset_count, likelihood)`
and calls `run(dataset, T)` (include/np_mcmc.h:71-95, src/np_main.cpp:391-471).  The classes below keep those names
and argument meanings; all arithmetic happens in the CUDA library.  There is no CPU fallback: importing works
without a GPU (so that CPU-only checks can load the library and inspect its symbols), creating a Context does not.
"""
import ctypes as C
import weakref
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnpb200.so")

ALG8, ALG2, JAIN_NEAL, TRIADIC = 8, 2, 20, 30
FAMILY_MVN, FAMILY_REGRESSION, FAMILY_ANGULAR = 0, 1, 2  # likelihood families of -c (np_main.cpp:196-205)
ALG2_CONJUGATE = 22  # collapsed Gibbs with the NIW posterior predictive (BASELINE configs[3]; not in the reference)
UPDATE_POSTERIOR_DRAW, UPDATE_POSTERIOR_MEAN = 1, 2
BUGCOMPAT_DEGENERATE_IW, BUGCOMPAT_UNDERFLOW = 1, 2
BUGCOMPAT_DEFAULT = BUGCOMPAT_DEGENERATE_IW

OK = 0
E_KMAX_OVERFLOW = -3
E_UNSUPPORTED = -5

EXPORTS = [
    "npb_ctx_create", "npb_ctx_destroy", "npb_ctx_stream", "npb_ctx_synchronize", "npb_status_str",
    "npb_ctx_last_error", "npb_dataset_upload", "npb_dataset_update", "npb_dataset_destroy", "npb_prior_set_niw",
    "npb_prior_set_nig", "npb_scalarnoise_logdensity_batch", "npb_chains_sample_base_nig",
    "npb_logdensity_batch", "npb_logdensity_sum", "npb_chains_create", "npb_chains_destroy", "npb_chains_set_state",
    "npb_chains_sweep", "npb_chains_sweep_host", "npb_chain_update_alg8", "npb_replay_alg8",
    "npb_chains_get_assignments", "npb_chains_get_params", "npb_chains_metrics", "npb_cocluster",
    "npb_chains_count", "npb_chains_kmax", "npb_scan_order_host", "npb_fp32_peak", "npb_chains_init_from_params",
    "npb_chains_split_merge", "npb_chains_last_proposal", "npb_chains_update_params", "npb_replay_split_merge", "npb_chains_consider_max_likelihood", "npb_chains_get_best_assignments",
    "npb_chains_probe_tile_logdensity", "npb_chains_set_option", "npb_chains_sweep_host_delta", "npb_chains_get_best_params",
    "npb_chain_move_item", "npb_chain_move_item_new", "npb_chain_remove_cluster", "npb_chains_kernel_time", "npb_chains_broadcast_state", "npb_chains_alg2_logpred", "npb_chains_alg2_suffstats",
    "npb_comm_unique_id", "npb_comm_create", "npb_comm_create_all", "npb_comm_destroy", "npb_comm_allreduce_sum", "npb_comm_group",
    "npb_cocluster_allreduce", "npb_cocluster_host",
]


class NpbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("npb200: %s (status %d)" % (msg, status))
        self.status = status


class SweepStats(C.Structure):
    _fields_ = [("reassignments", C.c_int64), ("candidates", C.c_int64), ("moved", C.c_int64),
                ("new_clusters", C.c_int64), ("sm_attempts", C.c_int64 * 4), ("sm_accepts", C.c_int64 * 4),
                ("sams_allocations", C.c_int64), ("mean_K", C.c_double), ("max_K", C.c_int32),
                ("overflow_chains", C.c_int32), ("kernel_ms", C.c_float)]


_lib = None


def load_library():
    """dlopen libnpb200.so; raises if the CUDA extension has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libnpb200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(noparama_b200 has no CPU path)")
    L = C.CDLL(LIB_PATH)
    vp, i64, dp = C.c_void_p, C.c_int64, C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int32)
    L.npb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.npb_ctx_destroy.argtypes = [vp]
    L.npb_ctx_stream.restype = vp
    L.npb_ctx_stream.argtypes = [vp]
    L.npb_ctx_synchronize.argtypes = [vp]
    L.npb_status_str.restype = C.c_char_p
    L.npb_status_str.argtypes = [C.c_int]
    L.npb_ctx_last_error.restype = C.c_char_p
    L.npb_ctx_last_error.argtypes = [vp]
    L.npb_dataset_upload.argtypes = [vp, dp, i64, C.c_int, C.POINTER(vp)]
    L.npb_dataset_update.argtypes = [vp, dp]
    L.npb_dataset_destroy.argtypes = [vp]
    L.npb_prior_set_niw.argtypes = [vp, C.c_int, dp, C.c_double, C.c_double, dp, C.c_double, C.c_int]
    L.npb_logdensity_batch.argtypes = [vp, vp, C.POINTER(i64), i64, dp, dp, C.c_int, C.c_int, dp]
    L.npb_logdensity_sum.argtypes = [vp, vp, C.POINTER(i64), i64, dp, dp, C.c_int, dp]
    L.npb_chains_create.argtypes = [vp, vp, i64, C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(vp)]
    L.npb_chains_destroy.argtypes = [vp]
    L.npb_chains_set_state.argtypes = [vp, i64, ip, C.c_int, ip, dp, dp]
    L.npb_chains_init_from_params.argtypes = [vp, C.c_int, dp, dp]
    L.npb_chains_sweep.argtypes = [vp, C.c_int, C.c_int, C.POINTER(SweepStats)]
    L.npb_chains_split_merge.argtypes = [vp, C.c_int, i64, C.POINTER(SweepStats)]
    L.npb_chains_last_proposal.argtypes = [vp, C.POINTER(C.c_float)]
    L.npb_chains_update_params.argtypes = [vp, C.c_int, dp, C.c_double, C.c_double, dp]
    L.npb_chains_sweep_host.argtypes = [vp, dp, C.c_int, C.c_int, C.POINTER(C.c_uint16), C.POINTER(SweepStats)]
    L.npb_chain_update_alg8.argtypes = [vp, i64, i64]
    L.npb_replay_alg8.argtypes = [vp, vp, C.c_int, C.c_int, ip, C.c_int, ip, dp, dp, i64, ip, C.POINTER(i64), ip, dp, dp,
                                  dp, ip, ip, i64, ip]
    L.npb_replay_split_merge.argtypes = [vp, vp, C.c_int, C.c_int, ip, C.c_int, ip, dp, dp, i64, ip, dp, dp, dp,
                                         C.POINTER(i64), ip, dp, dp, ip, ip, ip, ip, dp, ip]
    L.npb_chains_get_assignments.argtypes = [vp, i64, i64, ip]
    L.npb_chains_consider_max_likelihood.argtypes = [vp, dp, dp]
    L.npb_chains_probe_tile_logdensity.argtypes = [vp, i64, ip, C.POINTER(C.c_float)]
    L.npb_chains_set_option.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.npb_chains_sweep_host_delta.argtypes = [vp, dp, C.c_int, C.c_int, C.POINTER(C.c_uint16), C.POINTER(SweepStats), C.POINTER(i64)]
    L.npb_chains_get_best_params.argtypes = [vp, i64, C.c_int, C.POINTER(C.c_int), ip, C.POINTER(i64), dp, dp]
    L.npb_chain_move_item.argtypes = [vp, i64, i64, C.c_int]
    L.npb_chain_move_item_new.argtypes = [vp, i64, i64, dp, dp, C.POINTER(C.c_int)]
    L.npb_chain_remove_cluster.argtypes = [vp, i64, C.c_int]
    L.npb_chains_kernel_time.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
    L.npb_chains_broadcast_state.argtypes = [vp, i64]
    L.npb_chains_alg2_logpred.argtypes = [vp, i64, ip, C.c_int, C.POINTER(C.c_float)]
    L.npb_chains_alg2_suffstats.argtypes = [vp, i64, ip, dp, dp]
    L.npb_comm_unique_id.argtypes = [C.c_char_p]
    L.npb_comm_create.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    L.npb_comm_destroy.argtypes = [vp]
    L.npb_comm_allreduce_sum.argtypes = [vp, vp, i64, C.c_int]
    L.npb_comm_group.argtypes = [C.c_int]
    L.npb_cocluster_allreduce.argtypes = [vp, C.POINTER(i64), i64, vp, vp]
    L.npb_cocluster_host.argtypes = [vp, C.POINTER(i64), i64, vp, C.POINTER(C.c_float)]
    L.npb_chains_get_best_assignments.argtypes = [vp, i64, i64, ip]
    L.npb_chains_get_params.argtypes = [vp, i64, C.c_int, C.POINTER(C.c_int), ip, C.POINTER(i64), dp, dp]
    L.npb_chains_metrics.argtypes = [vp, ip, dp, dp, dp, dp, ip]
    L.npb_cocluster.argtypes = [vp, C.POINTER(i64), i64, vp, C.c_int]
    L.npb_chains_count.restype = i64
    L.npb_chains_count.argtypes = [vp]
    L.npb_chains_kmax.argtypes = [vp]
    L.npb_scan_order_host.argtypes = [C.c_uint64, C.c_uint32, i64, ip]
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def scan_order(seed, sweep, N):
    """Item visited at each step of sweep `sweep` (host evaluation of the kernels' keyed permutation)."""
    out = np.empty(N, dtype=np.int32)
    st = load_library().npb_scan_order_host(seed, sweep, N, _ip(out))
    if st != OK:
        raise NpbError(st, "bad argument")
    return out


class Context:
    """One CUDA device + stream (npb_ctx)."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = C.c_void_p()
        st = self._lib.npb_ctx_create(device, C.byref(h))
        if st != OK:
            raise NpbError(st, "cannot create a context on CUDA device %d (no GPU? there is no CPU path)" % device)
        self._h = h
        self.device = device

    def check(self, st):
        if st != OK:
            detail = self._lib.npb_ctx_last_error(self._h).decode()
            raise NpbError(st, self._lib.npb_status_str(st).decode() + (": " + detail if detail else ""))

    @property
    def stream(self):
        return self._lib.npb_ctx_stream(self._h)

    def synchronize(self):
        self.check(self._lib.npb_ctx_synchronize(self._h))

    def close(self):
        if self._h:
            self._lib.npb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """NCCL communicator of one context (npb_comm): rank 0 makes the unique id, every rank creates its communicator with it."""

    @staticmethod
    def unique_id(ctx):
        buf = C.create_string_buffer(128)
        ctx.check(ctx._lib.npb_comm_unique_id(buf))
        return buf.raw

    def __init__(self, ctx, uid, rank, world):
        self.ctx, self.rank, self.world = ctx, rank, world
        h = C.c_void_p()
        ctx.check(ctx._lib.npb_comm_create(ctx._h, uid, rank, world, C.byref(h)))
        self._h = h

    def allreduce_sum(self, dev_ptr, count, bits=32):
        self.ctx.check(self.ctx._lib.npb_comm_allreduce_sum(self._h, C.c_void_p(dev_ptr), count, bits))

    def close(self):
        if self._h:
            self.ctx._lib.npb_comm_destroy(self._h)
            self._h = None


class Dataset:
    """dataset_t (include/np_data.h:9-15) resident in HBM."""

    def __init__(self, ctx, X):
        X = _f64(X)
        assert X.ndim == 2
        self.ctx, self.N, self.D = ctx, X.shape[0], X.shape[1]
        h = C.c_void_p()
        ctx.check(ctx._lib.npb_dataset_upload(ctx._h, _dp(X), self.N, self.D, C.byref(h)))
        self._h = h
        self._chains = weakref.WeakSet()  # the library refuses to destroy a dataset under live chain handles

    def update(self, X):
        X = _f64(X)
        assert X.shape == (self.N, self.D)
        self.ctx.check(self.ctx._lib.npb_dataset_update(self._h, _dp(X)))

    def close(self):
        if self._h:
            for ch in list(self._chains):
                ch.close()
            self.ctx.check(self.ctx._lib.npb_dataset_destroy(self._h))
            self._h = None


class NormalInverseWishart:
    """Suffies_NormalInvWishart + Suffies_Dirichlet (np_suffies.h:80-95; constants np_main.cpp:164,367-371)."""

    def __init__(self, mu0, kappa, nu, Lambda, alpha=1.0, flags=BUGCOMPAT_DEFAULT):
        self.mu0, self.Lambda = _f64(mu0), _f64(Lambda)
        self.kappa, self.nu, self.alpha, self.flags = float(kappa), float(nu), float(alpha), int(flags)
        self.D = len(self.mu0)

    def bind(self, ctx):
        ctx.check(ctx._lib.npb_prior_set_niw(ctx._h, self.D, _dp(self.mu0), self.kappa, self.nu, _dp(self.Lambda),
                                             self.alpha, self.flags))


class NormalInverseGamma:
    """Suffies_NormalInvGamma + Suffies_Dirichlet (np_suffies.h:112-128; constants np_main.cpp:357-364): the base measure of the
    scalar-noise families `-c regression` (family=REGRESSION, rows (1, a, b)) and `-c angular` (family=ANGULAR, rows (a, b))."""

    def __init__(self, family, mu0=(0.0, 0.0), Lambda=((0.01, 0.0), (0.0, 0.01)), nig_alpha=10.0, nig_beta=0.1, alpha=1.0):
        self.family = int(family)
        self.mu0, self.Lambda = _f64(mu0), _f64(Lambda)
        self.nig_alpha, self.nig_beta, self.alpha = float(nig_alpha), float(nig_beta), float(alpha)
        self.D = 3 if self.family == FAMILY_REGRESSION else 2  # width of a data row

    def bind(self, ctx):
        ctx.check(ctx._lib.npb_prior_set_nig(ctx._h, self.family, _dp(self.mu0), _dp(self.Lambda), C.c_double(self.nig_alpha),
                                             C.c_double(self.nig_beta), C.c_double(self.alpha)))


class ScalarNoiseNormal:
    """scalarnoise_multivariate_normal_distribution (statistics/scalarnoise_multivariatenormal.h) in regression or angular mode:
    batched log-densities on the device."""

    def __init__(self, ctx, dataset, family):
        self.ctx, self.ds, self.family = ctx, dataset, int(family)

    def logprobability(self, mu, sigma, rows=None):
        """log p(X[rows] | mu[k], sigma[k]) -> [n_rows, K]  (scalarnoise_multivariatenormal.cpp:182-250)"""
        mu, sigma = _f64(mu), _f64(sigma)
        K = len(sigma)
        assert mu.shape == (K, 2)
        if rows is None:
            n, rp = self.ds.N, None
        else:
            rows = np.ascontiguousarray(rows, dtype=np.int64)
            n, rp = len(rows), rows.ctypes.data_as(C.POINTER(C.c_int64))
        out = np.empty((n, K), dtype=np.float64)
        self.ctx.check(self.ctx._lib.npb_scalarnoise_logdensity_batch(self.ctx._h, self.ds._h, self.family, rp, C.c_int64(n), _dp(mu),
                                                                      _dp(sigma), K, _dp(out)))
        return out


class MultivariateNormal:
    """multivariate_normal_distribution (statistics/multivariatenormal.h): batched log-densities on the device."""

    def __init__(self, ctx, dataset):
        self.ctx, self.ds = ctx, dataset

    def logprobability(self, mu, Sigma, rows=None, precision=64):
        """log N(X[rows] | mu[k], Sigma[k]) -> [n_rows, K]  (multivariatenormal.cpp:106-136)"""
        mu, Sigma = _f64(mu), _f64(Sigma)
        K, D = mu.shape
        assert D == self.ds.D and Sigma.shape == (K, D, D)
        if rows is None:
            n, rp = self.ds.N, None
        else:
            rows = np.ascontiguousarray(rows, dtype=np.int64)
            n, rp = len(rows), rows.ctypes.data_as(C.POINTER(C.c_int64))
        out = np.empty((n, K), dtype=np.float64)
        self.ctx.check(self.ctx._lib.npb_logdensity_batch(self.ctx._h, self.ds._h, rp, n, _dp(mu), _dp(Sigma), K,
                                                          precision, _dp(out)))
        return out

    def probability(self, mu, Sigma, rows=None, precision=64):
        """multivariatenormal.cpp:64-94"""
        return np.exp(self.logprobability(mu, Sigma, rows, precision))

    def logprobability_dataset(self, mu, Sigma, rows=None):
        """sum over a member list (multivariatenormal.cpp:138-146) -> [K]"""
        mu, Sigma = _f64(mu), _f64(Sigma)
        K, D = mu.shape
        if rows is None:
            n, rp = self.ds.N, None
        else:
            rows = np.ascontiguousarray(rows, dtype=np.int64)
            n, rp = len(rows), rows.ctypes.data_as(C.POINTER(C.c_int64))
        out = np.empty(K, dtype=np.float64)
        self.ctx.check(self.ctx._lib.npb_logdensity_sum(self.ctx._h, self.ds._h, rp, n, _dp(mu), _dp(Sigma), K, _dp(out)))
        return out


def replay_alg8(ctx, dataset, trace, init_state, m_aux=3, z_every=None):
    """Parity level 2: replay a recorded NealAlgorithm8 trace in double precision (npb_replay_alg8).

    trace: dict with item, order_off, order, aux_mu, aux_Sigma, u, new_slot, max_slot (oracle/binding.py Run.trace());
    init_state: (z0, slots, mu, Sigma).  Returns (picked [S], z_after [S // z_every, N])."""
    z0, slots, mu, Sigma = init_state
    z0 = np.ascontiguousarray(z0, dtype=np.int32)
    slots = np.ascontiguousarray(slots, dtype=np.int32)
    mu, Sigma = _f64(mu), _f64(Sigma)
    S = len(trace["item"])
    N = dataset.N
    z_every = N if z_every is None else z_every
    item = np.ascontiguousarray(trace["item"], dtype=np.int32)
    off = np.ascontiguousarray(trace["order_off"], dtype=np.int64)
    order = np.ascontiguousarray(trace["order"], dtype=np.int32)
    aux_mu, aux_Sigma, u = _f64(trace["aux_mu"]), _f64(trace["aux_Sigma"]), _f64(trace["u"])
    new_slot = np.ascontiguousarray(trace["new_slot"], dtype=np.int32)
    picked = np.empty(S, dtype=np.int32)
    z_after = np.empty((S // z_every, N), dtype=np.int32)
    nslots = int(max(trace.get("max_slot", 0), slots.max() + 1, new_slot.max() + 1))
    ctx.check(ctx._lib.npb_replay_alg8(ctx._h, dataset._h, m_aux, nslots, _ip(z0), len(slots), _ip(slots), _dp(mu),
                                       _dp(Sigma), S, _ip(item), off.ctypes.data_as(C.POINTER(C.c_int64)), _ip(order),
                                       _dp(aux_mu), _dp(aux_Sigma), _dp(u), _ip(new_slot), _ip(picked), z_every,
                                       _ip(z_after)))
    return picked, z_after


def replay_split_merge(ctx, dataset, sampler, trace, init_state):
    """Parity level 2 for the split-merge samplers (npb_replay_split_merge): replay the oracle's recorded run
    (oracle/binding.py Run.sm_trace()) in double precision.  Returns dict(type, dec, accept, logA, z_final)."""
    z0, slots, mu, Sigma = init_state
    z0 = np.ascontiguousarray(z0, dtype=np.int32)
    slots = np.ascontiguousarray(slots, dtype=np.int32)
    mu, Sigma = _f64(mu), _f64(Sigma)
    n = len(trace["type"])
    picks = np.ascontiguousarray(trace["picks"], dtype=np.int32)
    u0, th_mu, th_sigma = _f64(trace["u0"]), _f64(trace["th_mu"]), _f64(trace["th_sigma"])
    off = np.ascontiguousarray(trace["pool_off"], dtype=np.int64)
    pool = np.ascontiguousarray(trace["pool"], dtype=np.int32)
    us, uacc = _f64(trace["us"]), _f64(trace["uacc"])
    new_slot = np.ascontiguousarray(trace["new_slot"], dtype=np.int32)
    nslots = int(max(trace.get("max_slot", 0), slots.max() + 1, new_slot.max() + 1))
    out = dict(type=np.empty(n, np.int32), dec=np.empty(len(pool), np.int32), accept=np.empty(n, np.int32),
               logA=np.empty(n), z_final=np.empty(dataset.N, np.int32))
    ctx.check(ctx._lib.npb_replay_split_merge(ctx._h, dataset._h, sampler, nslots, _ip(z0), len(slots), _ip(slots), _dp(mu),
                                              _dp(Sigma), n, _ip(picks), _dp(u0), _dp(th_mu), _dp(th_sigma),
                                              off.ctypes.data_as(C.POINTER(C.c_int64)), _ip(pool), _dp(us), _dp(uacc),
                                              _ip(new_slot), _ip(out["type"]), _ip(out["dec"]), _ip(out["accept"]),
                                              _dp(out["logA"]), _ip(out["z_final"])))
    return out


class Chains:
    """The membertrix state (include/membertrix.h) of n_chains independent chains, device resident."""

    def __init__(self, ctx, dataset, n_chains, Kmax=256, m_aux=3, K0=20, seed=20261018):
        self.ctx, self.ds = ctx, dataset
        self.C, self.Kmax, self.m_aux, self.K0 = int(n_chains), int(Kmax), int(m_aux), int(K0)
        h = C.c_void_p()
        ctx.check(ctx._lib.npb_chains_create(ctx._h, dataset._h, self.C, self.Kmax, self.m_aux, self.K0, seed, C.byref(h)))
        self._h = h
        dataset._chains.add(self)

    def sweep(self, sampler=ALG8, n_sweeps=1, want_stats=True):
        st = SweepStats()
        self.ctx.check(self.ctx._lib.npb_chains_sweep(self._h, sampler, n_sweeps, C.byref(st) if want_stats else None))
        return st

    def probe_tile_logdensity(self, chain, items32):
        """[32 slots, 32 items] log-densities as the D >= 4 sweep kernel's producer warp computes them (parity probe)"""
        items32 = np.ascontiguousarray(items32, dtype=np.int32)
        assert len(items32) == 32
        out = np.empty((32, 32), dtype=np.float32)
        self.ctx.check(self.ctx._lib.npb_chains_probe_tile_logdensity(self._h, chain, _ip(items32),
                                                                       out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def sample_base_nig(self, chain, count):
        """[count, 3] raw draws (mu_0, mu_1, sigma) of the normal-inverse-gamma base measure from `chain`'s generator"""
        out = np.empty((count, 3), dtype=np.float32)
        self.ctx.check(self.ctx._lib.npb_chains_sample_base_nig(self._h, C.c_int64(chain), count, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def set_option(self, name, value):
        """behaviour switch of this handle, e.g. ("d16_path", "auto" | "tc" | "tc2" | "fp32")"""
        self.ctx.check(self.ctx._lib.npb_chains_set_option(self._h, name.encode(), value.encode()))

    def kernel_time(self):
        """(ms, launches) of the dominant sweep kernel since the last call (after set_option("time_kernels", "1"))"""
        ms, n = C.c_double(), C.c_int64()
        self.ctx.check(self.ctx._lib.npb_chains_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def update_item(self, item, chain=-1):
        """one NealAlgorithm8::update(membertrix&, {item}) on `chain` (every chain if negative): the single-item seam"""
        self.ctx.check(self.ctx._lib.npb_chain_update_alg8(self._h, chain, item))

    def split_merge(self, sampler, n_proposals):
        """the first n_proposals subsets of the next sweep of a split-merge sampler (np_mcmc.cpp:146-163)"""
        st = SweepStats()
        self.ctx.check(self.ctx._lib.npb_chains_split_merge(self._h, sampler, n_proposals, C.byref(st)))
        return st

    def update_params(self, mode=1, prior=None):
        """refresh (mu, Sigma) of every occupied cluster from the conjugate NIW posterior of its members: mode 1 = draw,
        2 = posterior mean; prior = dict(mu0, kappa, nu, Lambda) or None for the bound prior"""
        if prior is None:
            self.ctx.check(self.ctx._lib.npb_chains_update_params(self._h, mode, None, 0.0, 0.0, None))
        else:
            mu0, Lam = _f64(prior["mu0"]), _f64(prior["Lambda"])
            self.ctx.check(self.ctx._lib.npb_chains_update_params(self._h, mode, _dp(mu0), float(prior["kappa"]),
                                                                  float(prior["nu"]), _dp(Lam)))

    PROPOSAL_FIELDS = ("type", "stat", "logA", "accept", "n0", "n1", "n2", "pool", "new_slot", "dying", "u", "Q")

    def last_proposal(self):
        """detail of the last split-merge proposal of every chain -> dict of [C] arrays (npb_chains_last_proposal)"""
        out = np.zeros((self.C, 16), dtype=np.float32)
        self.ctx.check(self.ctx._lib.npb_chains_last_proposal(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return {k: out[:, i].copy() for i, k in enumerate(self.PROPOSAL_FIELDS)}

    def sweep_host(self, X, sampler=ALG8, n_sweeps=1, z_out=None, want_stats=False):
        """end-to-end step with host buffers: H2D of X, sweep, D2H of all assignments [N, C] uint16"""
        st = SweepStats()
        zp = z_out.ctypes.data_as(C.POINTER(C.c_uint16)) if z_out is not None else None
        self.ctx.check(self.ctx._lib.npb_chains_sweep_host(self._h, _dp(X), sampler, n_sweeps, zp,
                                                           C.byref(st) if want_stats else None))
        return st

    def sweep_host_delta(self, X, sampler=ALG8, n_sweeps=1, z_mirror=None, want_stats=False):
        """end-to-end step with an incremental result: z_mirror [N, C] uint16 is the caller's copy of the assignments; only the
        entries that changed since the previous call travel.  Returns (stats, entries written)."""
        st = SweepStats()
        n = C.c_int64()
        assert z_mirror is not None and z_mirror.dtype == np.uint16 and z_mirror.flags.c_contiguous
        self.ctx.check(self.ctx._lib.npb_chains_sweep_host_delta(self._h, _dp(X) if X is not None else None, sampler, n_sweeps,
                                                                 z_mirror.ctypes.data_as(C.POINTER(C.c_uint16)),
                                                                 C.byref(st) if want_stats else None, C.byref(n)))
        return st, n.value

    def move_item(self, chain, item, slot):
        """membertrix::retract + assign of one item (status codes mirror np_error_t)"""
        self.ctx.check(self.ctx._lib.npb_chain_move_item(self._h, chain, item, slot))

    def move_item_new(self, chain, item, mu, Sigma):
        """membertrix::addCluster + assign: the item founds a cluster; returns its slot"""
        mu, Sigma = _f64(mu), _f64(Sigma)
        slot = C.c_int()
        self.ctx.check(self.ctx._lib.npb_chain_move_item_new(self._h, chain, item, _dp(mu), _dp(Sigma), C.byref(slot)))
        return slot.value

    def set_state(self, chain, z, slots, mu, Sigma):
        z = np.ascontiguousarray(z, dtype=np.int32)
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        mu, Sigma = _f64(mu), _f64(Sigma)
        self.ctx.check(self.ctx._lib.npb_chains_set_state(self._h, chain, _ip(z), len(slots), _ip(slots), _dp(mu), _dp(Sigma)))

    def alg2_logpred(self, chain, items):
        """[n, 33] NIW posterior-predictive log-densities of `items` under the clusters of `chain` (column 32: the prior's)"""
        items = np.ascontiguousarray(items, dtype=np.int32)
        out = np.empty((len(items), 33), dtype=np.float32)
        self.ctx.check(self.ctx._lib.npb_chains_alg2_logpred(self._h, chain, _ip(items), len(items), out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def alg2_suffstats(self, chain):
        """(counts [32], sum x [32, D], sum x x^T [32, D, D]) the conjugate path keeps for `chain`"""
        D = self.ds.D
        n = np.empty(32, np.int32)
        sx, sxx = np.empty((32, D)), np.empty((32, D, D))
        self.ctx.check(self.ctx._lib.npb_chains_alg2_suffstats(self._h, chain, _ip(n), _dp(sx), _dp(sxx)))
        return n, sx, sxx

    def broadcast_state(self, src=0):
        """every chain takes chain src's assignments and clusters (their random streams stay their own)"""
        self.ctx.check(self.ctx._lib.npb_chains_broadcast_state(self._h, src))

    def init_from_params(self, mu, Sigma):
        """every chain restarts from the same K clusters (given parameters) and a uniform random assignment"""
        mu, Sigma = _f64(mu), _f64(Sigma)
        self.ctx.check(self.ctx._lib.npb_chains_init_from_params(self._h, mu.shape[0], _dp(mu), _dp(Sigma)))

    def assignments(self, chain0=0, n=None):
        n = self.C - chain0 if n is None else n
        out = np.empty((n, self.ds.N), dtype=np.int32)
        self.ctx.check(self.ctx._lib.npb_chains_get_assignments(self._h, chain0, n, _ip(out)))
        return out

    def consider_max_likelihood(self):
        """MCMC::considerMaxLikelihood (np_mcmc.cpp:187-203) for every chain -> (joint log-lik now, best so far)"""
        cur, best = np.empty(self.C), np.empty(self.C)
        self.ctx.check(self.ctx._lib.npb_chains_consider_max_likelihood(self._h, _dp(cur), _dp(best)))
        return cur, best

    def best_assignments(self, chain0=0, n=None):
        """assignments of the max-likelihood state kept by consider_max_likelihood (MCMC::getMaxLikelihoodMatrix)"""
        n = self.C - chain0 if n is None else n
        out = np.empty((n, self.ds.N), dtype=np.int32)
        self.ctx.check(self.ctx._lib.npb_chains_get_best_assignments(self._h, chain0, n, _ip(out)))
        return out

    def params(self, chain):
        cap, D = self.Kmax, self.ds.D
        K = C.c_int()
        slots = np.empty(cap, np.int32)
        counts = np.empty(cap, np.int64)
        mu = np.empty((cap, D))
        Sigma = np.empty((cap, D, D))
        self.ctx.check(self.ctx._lib.npb_chains_get_params(self._h, chain, cap, C.byref(K), _ip(slots),
                                                           counts.ctypes.data_as(C.POINTER(C.c_int64)), _dp(mu), _dp(Sigma)))
        k = K.value
        return slots[:k].copy(), counts[:k].copy(), mu[:k].copy(), Sigma[:k].copy()

    def best_params(self, chain):
        """the clusters of the state kept by consider_max_likelihood (not of the current state)"""
        cap, D = self.Kmax, self.ds.D
        K = C.c_int()
        slots = np.empty(cap, np.int32)
        counts = np.empty(cap, np.int64)
        mu = np.empty((cap, D))
        Sigma = np.empty((cap, D, D))
        self.ctx.check(self.ctx._lib.npb_chains_get_best_params(self._h, chain, cap, C.byref(K), _ip(slots),
                                                                counts.ctypes.data_as(C.POINTER(C.c_int64)), _dp(mu), _dp(Sigma)))
        k = K.value
        return slots[:k].copy(), counts[:k].copy(), mu[:k].copy(), Sigma[:k].copy()

    def metrics(self, truth=None, joint_loglik=True):
        """purity / Rand / adjusted Rand per chain (clustering_performance.cpp:38-82), joint log-likelihood
        (np_mcmc.cpp:187-203) and occupied cluster count."""
        Cn = self.C
        pur, ri, ari, jll = (np.zeros(Cn) for _ in range(4))
        K = np.zeros(Cn, np.int32)
        tp = None
        if truth is not None:
            truth = np.ascontiguousarray(truth, dtype=np.int32)
            tp = _ip(truth)
        self.ctx.check(self.ctx._lib.npb_chains_metrics(self._h, tp, _dp(pur), _dp(ri), _dp(ari),
                                                        _dp(jll) if joint_loglik else None, _ip(K)))
        return dict(purity=pur, rand_index=ri, adjusted_rand=ari, joint_loglik=jll, K=K)

    def cocluster_into(self, anchors, S_dev_ptr, accumulate=False):
        anchors = np.ascontiguousarray(anchors, dtype=np.int64)
        self.ctx.check(self.ctx._lib.npb_cocluster(self._h, anchors.ctypes.data_as(C.POINTER(C.c_int64)), len(anchors),
                                                   C.c_void_p(S_dev_ptr), int(accumulate)))

    def cocluster(self, anchors, comm=None):
        """[A, A] float32 co-clustering counts of the anchors (summed over the communicator's ranks if one is given)"""
        anchors = np.ascontiguousarray(anchors, dtype=np.int64)
        out = np.empty((len(anchors), len(anchors)), dtype=np.float32)
        self.ctx.check(self.ctx._lib.npb_cocluster_host(self._h, anchors.ctypes.data_as(C.POINTER(C.c_int64)), len(anchors),
                                                        comm._h if comm is not None else None, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def cocluster_allreduce(self, anchors, S_dev_ptr, comm=None):
        """co-clustering counts of the anchors over this handle's chains, summed over the communicator's ranks (in the library:
        k_cc_gather + k_cc_tile, then ncclAllReduce on the context's stream)"""
        anchors = np.ascontiguousarray(anchors, dtype=np.int64)
        self.ctx.check(self.ctx._lib.npb_cocluster_allreduce(self._h, anchors.ctypes.data_as(C.POINTER(C.c_int64)), len(anchors),
                                                             comm._h if comm is not None else None, C.c_void_p(S_dev_ptr)))

    def close(self):
        if self._h:
            if self.ctx._h:  # (a context closed first took its stream with it: nothing left to release the handle on)
                self.ctx._lib.npb_chains_destroy(self._h)
            self._h = None

    def __del__(self):
        # a handle dropped without close() must not pin its dataset (npb_dataset_destroy refuses under live chain handles)
        try:
            self.close()
        except Exception:
            pass


class NealAlgorithm8:
    """UpdateClusterPopulation implementation selected by `-a algorithm8` (np_main.cpp:433-439)."""
    sampler = ALG8
    subset_count = 1


class NealAlgorithm2:
    """The sampler src/np_neal_algorithm2.cpp:32-120 describes (not compiled by the reference): one prior draw weighted
    alpha beside the K occupied clusters.  Use with MCMC(..., m_aux=1)."""
    sampler = ALG2
    subset_count = 1


class NealAlgorithm2Conjugate:
    """The CONJUGATE form of the same file's intent (updateSuffies, np_neal_algorithm2.cpp:54): collapsed Gibbs with the NIW posterior
    predictive (npb_alg2*.cu).  Kmax = 32; the prior needs nu > D - 1."""
    sampler = ALG2_CONJUGATE
    subset_count = 1


class JainNealAlgorithm:
    """UpdateClusterPopulation implementation selected by `-a jain_neal_split` (np_main.cpp:440-446)."""
    sampler = JAIN_NEAL
    subset_count = 2


class TriadicAlgorithm:
    """UpdateClusterPopulation implementation selected by `-a triadic` (np_main.cpp:447-458)."""
    sampler = TRIADIC
    subset_count = 3


class MCMC:
    """MCMC::run (np_mcmc.cpp:48-175) for `chains` lockstep chains on one device.

    run(T) performs T sweeps; getMembershipMatrix() mirrors MCMC::getMembershipMatrix (np_mcmc.h:88)."""

    def __init__(self, ctx, dataset, prior, update_cluster_population=NealAlgorithm8, chains=1, Kmax=256, K0=20,
                 m_aux=3, seed=20261018):
        prior.bind(ctx)
        self.algorithm = update_cluster_population
        self.chains = Chains(ctx, dataset, chains, Kmax=Kmax, m_aux=m_aux, K0=K0, seed=seed)

    def run(self, T, sweeps_per_launch=None, cycle_max_likelihood=0):
        """T sweeps; cycle_max_likelihood = 5 keeps the max-likelihood state like np_mcmc.cpp:172-174 (every 5th sweep)"""
        stats = []
        step = T if not sweeps_per_launch else sweeps_per_launch
        if cycle_max_likelihood:
            step = cycle_max_likelihood
        done = 0
        while done < T:
            n = min(step, T - done)
            stats.append(self.chains.sweep(self.algorithm.sampler, n))
            done += n
            if cycle_max_likelihood:
                self.chains.consider_max_likelihood()
        return stats

    def getMaxLikelihoodMatrix(self, chain0=0, n=None):
        return self.chains.best_assignments(chain0, n)

    def getMembershipMatrix(self, chain0=0, n=None):
        return self.chains.assignments(chain0, n)
