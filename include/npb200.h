/*
 * npb200.h -- C ABI of libnpb200.so: the B200 (sm_100a) Gibbs-reassignment path for
 * Dirichlet-process mixtures, drop-in behind the sampler seam of mrquincle/noparama.
 *
 * The reference has no FFI: its "operator API" is constructor injection of three C++
 * interfaces (citations relative to the reference tree):
 *     UpdateClusterPopulation::update(membertrix&, const data_ids_t&)   include/np_update_cluster_population.h:35-43
 *     distribution_t::{init, probability, logprobability, operator()}    include/statistics/distribution.h:47-87
 *     membertrix::{assign, retract, addCluster, getClusterId, count ...} include/membertrix.h:106-290
 * driven by MCMC::run (src/np_mcmc.cpp:48-175).  Each entry point below names the reference
 * function(s) it replaces.  Conventions: plain pointers and sizes, no C++/torch types; every
 * function returns npb_status (0 = OK, negative = error) and never aborts; host buffers are caller
 * owned and are copied during the call; every *_create pairs with a *_destroy; one context = one
 * CUDA device and one stream; a context is thread-compatible (one caller at a time).
 *
 * There is NO CPU fallback: without a CUDA device npb_ctx_create fails with NPB_E_CUDA.
 */
#ifndef NPB200_H
#define NPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int npb_status;
enum {
	NPB_OK = 0,
	NPB_E_BAD_ARG = -1,
	NPB_E_CUDA = -2,
	NPB_E_KMAX_OVERFLOW = -3,     /* a chain needed more than Kmax cluster slots */
	NPB_E_NOT_POSITIVE = -4,      /* covariance not invertible / precision not positive definite */
	NPB_E_UNSUPPORTED = -5,
	NPB_E_REPLAY_MISMATCH = -6,   /* replay picked a different candidate than the recorded trace */
	NPB_E_NOMEM = -7,
	NPB_E_NCCL = -8,              /* NCCL missing or a collective failed */
	/* mirrors of np_error_t (include/membertrix.h:16-23) for the single-item seam */
	NPB_E_ALREADY_ASSIGNED = -16,
	NPB_E_ASSIGNMENT_REMAINING = -17,
	NPB_E_ASSIGNMENT_ABSENT = -18
};

/* samplers selectable by -a (src/np_main.cpp:212-236, 424-459).  NPB_ALG2 is the sampler src/np_neal_algorithm2.cpp:32-120
 * describes (the reference does not compile it): K weights p(x|theta_k) n_k plus ONE prior draw weighted p(x|theta') alpha --
 * the Algorithm 8 kernels with a single auxiliary draw; it needs chains created with m_aux = 1. */
enum { NPB_ALG8 = 8, NPB_ALG2 = 2, NPB_JAIN_NEAL = 20, NPB_TRIADIC = 30,
       /* the CONJUGATE (collapsed) Algorithm 2 BASELINE configs[3] names: NIW posterior-predictive weights n_k pred_k(x) and
        * alpha pred_0(x), sufficient statistics up- and down-dated on insert and removal.  The reference has no executable
        * form of it (np_neal_algorithm2.cpp:32-120 is dead code, normalinvwishart.h:66-75 asserts false): textbook model,
        * parity against the reference unpinned.  Kmax = 32; D = 2, 4, 8, 16, 64; the context's NIW prior with nu > D - 1. */
       NPB_ALG2_CONJUGATE = 22 };

/* behaviour switches of npb_prior_set_niw; the default (all bug-compatible flags set) reproduces the
 * reference, quirk numbers refer to SURVEY.md 7.4 */
enum {
	NPB_BUGCOMPAT_DEGENERATE_IW = 1 << 0, /* Q2: Sigma = v^2 L^T L with scalar v ~ N(D, nu^2) (invwishart.h:34-46) */
	NPB_BUGCOMPAT_UNDERFLOW     = 1 << 1, /* Q6: linear-domain double weights underflow to 0; all-zero => first candidate */
	NPB_BUGCOMPAT_DEFAULT       = NPB_BUGCOMPAT_DEGENERATE_IW
};

typedef struct npb_ctx npb_ctx;
typedef struct npb_dataset npb_dataset;
typedef struct npb_chains npb_chains;

typedef struct npb_sweep_stats {
	int64_t reassignments;      /* item updates performed = chains * N * sweeps (Alg. 8) */
	int64_t candidates;         /* sum over updates of (K_i + m): density evaluations that enter a pick */
	int64_t moved;              /* updates that changed the item's cluster */
	int64_t new_clusters;       /* updates that picked an auxiliary draw (np_neal_algorithm8.cpp:136-145) */
	int64_t sm_attempts[4];     /* split/merge samplers: merge 2->1, split 1->2, merge 3->2, split 2->3 */
	int64_t sm_accepts[4];
	int64_t sams_allocations;   /* items allocated by restricted (SAMS) scans over all proposals */
	double  mean_K;             /* mean occupied clusters over chains after the last sweep */
	int32_t max_K;              /* largest occupied-cluster count over chains after the last sweep */
	int32_t overflow_chains;    /* chains that hit Kmax (their results are invalid) */
	float   kernel_ms;          /* device time of the sweep kernel(s), CUDA events on the context stream */
} npb_sweep_stats;

/* ---- context ------------------------------------------------------------------------------------------- */
npb_status npb_ctx_create(int device, npb_ctx **out);
npb_status npb_ctx_destroy(npb_ctx *ctx);
/* the cudaStream_t all work of this context is enqueued on (for event timing by the caller) */
void *npb_ctx_stream(npb_ctx *ctx);
npb_status npb_ctx_synchronize(npb_ctx *ctx);
const char *npb_status_str(npb_status s);
/* last CUDA error text seen by this context ("" if none) */
const char *npb_ctx_last_error(npb_ctx *ctx);

/* ---- data: replaces membertrix::addData / dataset_t (membertrix.cpp:124-138, np_data.h:9-15) ------------ */
/* X row-major [N,D] doubles; stored on the device as fp32 (sweeps) and fp64 (replay / precision=64) */
npb_status npb_dataset_upload(npb_ctx *ctx, const double *X, int64_t N, int D, npb_dataset **out);
/* re-upload new values into an existing dataset of the same shape (end-to-end timing path) */
npb_status npb_dataset_update(npb_dataset *ds, const double *X);
npb_status npb_dataset_destroy(npb_dataset *ds);

/* ---- prior: replaces dirichlet_process(alpha, normal_inverse_wishart_distribution) ----------------------
 * (dirichlet.h:20-93, normalinvwishart.h:27-64, constants np_main.cpp:164,367-371) */
npb_status npb_prior_set_niw(npb_ctx *ctx, int D, const double *mu0, double kappa, double nu, const double *Lambda,
		double alpha, int flags);

/* ---- the other two likelihood families of `-c` (np_main.cpp:196-205): scalar-noise normal likelihood + normal-inverse-gamma
 * base measure.  Replaces scalarnoise_multivariate_normal_distribution::{init, prepare, probability, logprobability}
 * (scalarnoise_multivariatenormal.cpp:16-47, 77-250) and normal_inverse_gamma_distribution::operator() (normalinvgamma.h:58-84,
 * gamma.h:37-46); constants of np_main.cpp:322-328, 357-364: mu0 = (0, 0), alpha = 10, beta = 0.1, Lambda = 0.01 I.
 * A cluster's theta = (mu [2], sigma).  Dataset rows are what read_data builds (np_main.cpp:83-101): (1, a, b) for regression
 * (D = 3: residual b - mu . (1, a)), (a, b) for angular (D = 2: residual |d - (-sin(t) a + cos(t) b)| with (d, t) made canonical
 * through abs(int), the reference's prepare(): SURVEY quirk list, Q12).  With such a prior bound, npb_chains_create draws the K0
 * initial clusters from it, npb_chains_sweep(NPB_ALG8) runs k_sn_sweep (npb_scalarnoise.cu), npb_chains_get_params returns
 * mu[k, 0..1] = (mu_0, mu_1) and Sigma[k, 0, 0] = sigma^2 (other entries 0); the split-merge samplers, Algorithm 2 and the
 * parameter update are multivariate-normal only (NPB_E_UNSUPPORTED). */
enum { NPB_FAMILY_MVN = 0, NPB_FAMILY_REGRESSION = 1, NPB_FAMILY_ANGULAR = 2 };
npb_status npb_prior_set_nig(npb_ctx *ctx, int family, const double *mu0 /*[2]*/, const double *Lambda /*[2,2]*/, double ig_alpha,
		double ig_beta, double alpha);
/* out[r*K+k] = log p(X[rows[r]] | mu[k], sigma[k]) in double (rows == NULL: rows 0..n_rows-1) */
npb_status npb_scalarnoise_logdensity_batch(npb_ctx *ctx, npb_dataset *ds, int family, const int64_t *rows, int64_t n_rows,
		const double *mu /*[K,2]*/, const double *sigma /*[K]*/, int K, double *out);
/* parity probe: `count` raw draws (mu_0, mu_1, sigma) of the base measure from chain `chain`'s generator; out [count, 3] */
npb_status npb_chains_sample_base_nig(npb_chains *ch, int64_t chain, int count, float *out);

/* ---- density: replaces multivariate_normal_distribution::{init, probability, logprobability} -----------
 * (multivariatenormal.cpp:16-35, 64-146).  out[r*K+k] = log N(X[rows[r]] | mu[k], Sigma[k]); rows == NULL
 * means rows 0..n_rows-1.  Sigma[k] is a general D x D matrix (row-major; the reference's known-answer test
 * uses a non-symmetric one).  precision = 64 evaluates in double, 32 in float with double parameters prepared
 * on the host. */
npb_status npb_logdensity_batch(npb_ctx *ctx, npb_dataset *ds, const int64_t *rows, int64_t n_rows, const double *mu,
		const double *Sigma, int K, int precision, double *out);
/* sum over a member list, multivariatenormal.cpp:138-146: out[k] = sum_r log N(X[rows[r]] | theta_k) */
npb_status npb_logdensity_sum(npb_ctx *ctx, npb_dataset *ds, const int64_t *rows, int64_t n_rows, const double *mu,
		const double *Sigma, int K, double *out);

/* ---- chains: replaces MCMC::run's state (np_mcmc.cpp:48-91: K0 prior clusters, uniform assignment,
 * cleanup) for n_chains independent chains run in lockstep ------------------------------------------------ */
npb_status npb_chains_create(npb_ctx *ctx, npb_dataset *ds, int64_t n_chains, int Kmax, int m_aux, int K0,
		uint64_t seed, npb_chains **out);
npb_status npb_chains_destroy(npb_chains *ch);
/* behaviour switches of a handle (no reference counterpart; the library reads its NPB_* environment defaults once, when a
 * handle is created).  "d16_path": "auto" | "tc" | "tc2" | "fp32" -- which kernels sweep D = 16, Kmax = 32 chains. */
npb_status npb_chains_set_option(npb_chains *ch, const char *name, const char *value);
/* with option "time_kernels" = "1": accumulated duration (CUDA events on the library's stream) and number of launches of the
 * dominant sweep kernel (k_sweep_tc16) since the previous call -- what a roofline of that kernel divides by */
npb_status npb_chains_kernel_time(npb_chains *ch, double *ms, int64_t *launches);
/* overwrite the state of one chain (used by parity tests and by the single-item seam):
 * z [N] slot ids, K clusters with slot ids, means [K,D], covariances [K,D,D] */
npb_status npb_chains_set_state(npb_chains *ch, int64_t chain, const int32_t *z, int K, const int32_t *slots,
		const double *mu, const double *Sigma);

/* every chain takes the state of chain `src` (assignments, clusters, counts); their random streams stay their own.  For tests
 * of the sampler's law: many chains making the same reassignment independently. */
npb_status npb_chains_broadcast_state(npb_chains *ch, int64_t src);

/* re-initialise EVERY chain with the same K given clusters (slots 0..K-1) and a fresh uniform assignment of the items
 * to them: the InitClusters step (np_init_clusters.cpp:24-41) with caller-supplied instead of prior-drawn parameters */
npb_status npb_chains_init_from_params(npb_chains *ch, int K, const double *mu /* [K,D] */, const double *Sigma /* [K,D,D] */);

/* n_sweeps sweeps of the chosen sampler over all chains: replaces the doubly nested loop of
 * np_mcmc.cpp:109-163 with NealAlgorithm8::update (np_neal_algorithm8.cpp:49-167),
 * JainNealAlgorithm::update (np_jain_neal_algorithm.cpp:424-502) or TriadicAlgorithm::update
 * (np_triadic_algorithm.cpp:633-795) inside. */
npb_status npb_chains_sweep(npb_chains *ch, int sampler, int n_sweeps, npb_sweep_stats *stats);
/* split-merge samplers only: the first n_proposals subsets of the next sweep(s) (np_mcmc.cpp:146-163; subsets with a
 * repeated item are skipped like the reference does, :155-158).  n_proposals == N is one sweep.  Reports
 * stats->reassignments = proposals actually made, sm_attempts/sm_accepts per move type, sams_allocations. */
npb_status npb_chains_split_merge(npb_chains *ch, int sampler, int64_t n_proposals, npb_sweep_stats *stats);
/* detail of the LAST proposal of every chain, detail_out [n_chains,16] floats: move type (0 JN split, 1 JN merge,
 * 2 triadic split, 3 triadic merge), statistics index, log acceptance ratio, accepted, part sizes |P_0..2|, pool size,
 * new slot, removed slot, the uniform, Q.  For parity tests of the acceptance arithmetic
 * (np_jain_neal_algorithm.cpp:243-296,339-392; np_triadic_algorithm.cpp:370-437,529-590). */
npb_status npb_chains_last_proposal(npb_chains *ch, float *detail_out);
/* The parameter update of a sweep done right.  The reference calls UpdateClusters::update after every sweep
 * (np_mcmc.cpp:170, np_update_clusters.cpp:71-141) but its result is sliced away (np_cluster.h:49-51), so parameters
 * never change after birth; this opt-in call refreshes (mu, Sigma) of every occupied cluster of every chain from the
 * conjugate normal-inverse-Wishart posterior of its members: a draw (NPB_UPDATE_POSTERIOR_DRAW) or the posterior mean
 * (NPB_UPDATE_POSTERIOR_MEAN).  mu0 [D], Lambda0 [D,D] row-major; pass mu0 = Lambda0 = NULL to use the prior bound with
 * npb_prior_set_niw (kappa0, nu0 are then ignored). */
enum { NPB_UPDATE_POSTERIOR_DRAW = 1, NPB_UPDATE_POSTERIOR_MEAN = 2 };
npb_status npb_chains_update_params(npb_chains *ch, int mode, const double *mu0, double kappa0, double nu0, const double *Lambda0);
/* end-to-end form with host buffers: upload X (as npb_dataset_update), sweep, download the assignments of
 * all chains; z_out [N, n_chains] uint16 slot ids (item-major), may be NULL */
npb_status npb_chains_sweep_host(npb_chains *ch, const double *X, int sampler, int n_sweeps, uint16_t *z_out,
		npb_sweep_stats *stats);
/* the same with an incremental result: z_mirror [N, n_chains] is the CALLER's copy of the assignments, which the call brings
 * up to date -- only the entries that changed since the previous call on this handle travel (an (index, slot) list compacted
 * on the device), the whole array on the first call or when more than a quarter of the entries changed.  The caller passes
 * the same array, unmodified in between, every time; *n_changed (may be NULL) receives the number of entries written.
 * (The reference's host reads every assignment after every sweep through membertrix::getAssignments, membertrix.cpp:315-322;
 * a converged chain changes a handful of them.) */
npb_status npb_chains_sweep_host_delta(npb_chains *ch, const double *X, int sampler, int n_sweeps, uint16_t *z_mirror,
		npb_sweep_stats *stats, int64_t *n_changed);

/* conjugate Algorithm 2 parity probes: the posterior-predictive log-densities of `items` under every cluster of `chain` as the
 * sweep kernel evaluates them (out [n_items, 33]; NaN without members; column 32 = prior predictive), and the sufficient
 * statistics the path keeps (counts [32], sum x [32, D], sum x x^T [32, D, D]) */
npb_status npb_chains_alg2_logpred(npb_chains *ch, int64_t chain, const int32_t *items, int n_items, float *out);
npb_status npb_chains_alg2_suffstats(npb_chains *ch, int64_t chain, int32_t *counts, double *sx, double *sxx);

/* membertrix::retract + membertrix::assign of one item of one chain (membertrix.cpp:147-233), for a host that drives single
 * reassignments itself: the item moves to the occupied cluster `slot`; a cluster left without members disappears
 * (membertrix.cpp:200-203).  NPB_E_ASSIGNMENT_ABSENT: no such cluster; NPB_E_ALREADY_ASSIGNED: the item is in it already.
 * _new: addCluster + assign (membertrix.cpp:87-118): the item founds a cluster with the given parameters in the lowest free
 * slot (*slot_out); NPB_E_KMAX_OVERFLOW without one. */
npb_status npb_chain_move_item(npb_chains *ch, int64_t chain, int64_t item, int slot);
npb_status npb_chain_move_item_new(npb_chains *ch, int64_t chain, int64_t item, const double *mu, const double *Sigma, int *slot_out);
/* membertrix::remove (membertrix.cpp:213-228): NPB_E_ASSIGNMENT_REMAINING while the cluster has members */
npb_status npb_chain_remove_cluster(npb_chains *ch, int64_t chain, int slot);

/* one NealAlgorithm8::update(membertrix&, {item}) (np_neal_algorithm8.cpp:49-167) on one chain -- the reference's
 * single-item seam (np_mcmc.cpp:162); chain < 0 applies it to every chain of the handle.  Any D and Kmax. */
npb_status npb_chain_update_alg8(npb_chains *ch, int64_t chain, int64_t item);

/* parity level 2: replay a recorded trace of NealAlgorithm8::update in double precision with
 * the reference's linear-domain weights, for one chain over `ds`, using the bound prior's alpha.
 * Initial state: z0 [N] slot ids, K0 clusters (slots0, mu0 [K0,D], Sigma0 [K0,D,D]); nslots = slot capacity.
 * Per step s < n_steps: item[s]; the K candidates order[order_off[s] .. order_off[s+1]) as slot ids in the order
 * the reference iterated them; m_aux auxiliary thetas aux_mu [S,m,D], aux_Sigma [S,m,D,D]; the uniform u[s]; the
 * slot new_slot[s] a picked auxiliary is born into.  picked_out [n_steps] receives the candidate index chosen on
 * the device; z_after_out (may be NULL) receives z after every `z_every` steps ([n_steps / z_every, N]). */
npb_status npb_replay_alg8(npb_ctx *ctx, npb_dataset *ds, int m_aux, int nslots, const int32_t *z0, int K0,
		const int32_t *slots0, const double *mu0, const double *Sigma0, int64_t n_steps, const int32_t *item,
		const int64_t *order_off, const int32_t *order, const double *aux_mu, const double *aux_Sigma,
		const double *u, const int32_t *new_slot, int32_t *picked_out, int64_t z_every, int32_t *z_after_out);

/* parity level 2 for the split-merge samplers: replay a recorded run of JainNealAlgorithm::update
 * (np_jain_neal_algorithm.cpp:424-502) or TriadicAlgorithm::update (np_triadic_algorithm.cpp:633-795) in double precision
 * with the reference's own weights and picks.  Initial state as in npb_replay_alg8.  Per proposal p < n_prop: the subset
 * picks [n_prop,3] (-1 pad) MCMC::run handed over, u0 (the triadic sampler's first uniform), the prior draw of a split
 * (new_mu [n_prop,D], new_Sigma [n_prop,D,D]), the pool pool[pool_off[p] .. pool_off[p+1]) in the visiting order the
 * reference shuffled it into with the uniform each allocation consumed (us, < 0 for a pick that seeded its part), the
 * acceptance uniform uacc and the slot new_slot a cluster born by an accepted split takes.  Outputs: the move type the
 * device derived (0 JN split, 1 JN merge, 2 triadic split, 3 triadic merge), the part every pool member was allocated to,
 * accept, log acceptance ratio, and the assignments after the last proposal. */
npb_status npb_replay_split_merge(npb_ctx *ctx, npb_dataset *ds, int sampler, int nslots, const int32_t *z0, int K0,
		const int32_t *slots0, const double *mu0, const double *Sigma0, int64_t n_prop, const int32_t *picks, const double *u0,
		const double *new_mu, const double *new_Sigma, const int64_t *pool_off, const int32_t *pool, const double *us,
		const double *uacc, const int32_t *new_slot, int32_t *type_out, int32_t *dec_out, int32_t *accept_out, double *logA_out,
		int32_t *z_final_out);

/* parity probe (Kmax = 32, D = 4 / 8 / 16): the [32 slots x 32 items] log-density tile exactly as the sweep kernel's producer
 * warp computes it (packed FP32, the mean folded into a per-row offset; at D = 64 the tcgen05 density table of
 * npb_alg8_gemm.cu exactly as the race reads it), natural-log units, out[slot * 32 + j] for the 32
 * given items; NaN for a slot without members.  For the 1e-5 relative bar on log-densities
 * (multivariatenormal.cpp:106-136). */
npb_status npb_chains_probe_tile_logdensity(npb_chains *ch, int64_t chain, const int32_t *items32, float *out);

/* state readback: replaces membertrix::getClusterId / getClusters / count (membertrix.cpp:235-257,328-330) */
npb_status npb_chains_get_assignments(npb_chains *ch, int64_t chain0, int64_t n, int32_t *z_out /* [n,N] slot ids */);
npb_status npb_chains_get_params(npb_chains *ch, int64_t chain, int cap, int *K, int32_t *slots, int64_t *counts,
		double *mu /* [cap,D] */, double *Sigma /* [cap,D,D] */);

/* metrics for every chain: replaces clustering_performance::{calculateContingencyMatrix, calculateSimilarity}
 * (clustering_performance.cpp:14-82) and MCMC::considerMaxLikelihood's joint log-likelihood
 * (np_mcmc.cpp:187-203).  truth [N] labels >= 0; outputs [n_chains] each, any may be NULL. */
npb_status npb_chains_metrics(npb_chains *ch, const int32_t *truth, double *purity, double *rand_index,
		double *adjusted_rand, double *joint_loglik, int32_t *K);

/* MCMC::considerMaxLikelihood (np_mcmc.cpp:187-203) for every chain: the joint log-likelihood of the current state is
 * compared with that of the state kept so far, and chains that improved copy their assignments, parameters and counts into
 * the kept state (the reference deep-clones its membertrix).  joint_loglik_out / best_out [n_chains] may be NULL.
 * npb_chains_get_best_assignments reads the kept assignments back like npb_chains_get_assignments
 * (MCMC::getMaxLikelihoodMatrix, np_mcmc.h:90). */
npb_status npb_chains_consider_max_likelihood(npb_chains *ch, double *joint_loglik_out, double *best_out);
/* the clusters of the kept state, as npb_chains_get_params reports the current ones (slot ids are re-used after a death, so
 * the current slot table does not describe a state kept earlier) */
npb_status npb_chains_get_best_params(npb_chains *ch, int64_t chain, int cap, int *K, int32_t *slots, int64_t *counts, double *mu,
		double *Sigma);
npb_status npb_chains_get_best_assignments(npb_chains *ch, int64_t chain0, int64_t n, int32_t *z_out /* [n,N] slot ids */);

/* posterior co-clustering counts over this context's chains for an anchor subset: S[a,b] = #chains with
 * z[anchors[a]] == z[anchors[b]].  S_dev is a DEVICE pointer to n_anchor*n_anchor floats (so that the caller can
 * all-reduce it over NCCL without a host round trip); accumulate != 0 adds to S_dev. */
npb_status npb_cocluster(npb_chains *ch, const int64_t *anchors, int64_t n_anchor, float *S_dev, int accumulate);

/* ---- the path's only exchange: all-reduce of the co-clustering matrix and the diagnostics over the GPUs of a node.
 * NCCL is bound at run time (dlopen of libnccl.so.2; NPB_E_NCCL without it).  A communicator belongs to one context (device).
 * Multi-process (one rank per GPU): rank 0 calls npb_comm_unique_id and hands the 128 bytes to the others by any means, every
 * rank calls npb_comm_create.  Single process (the CLI's --gpus): npb_comm_create_all, one communicator per context.
 * npb_cocluster_allreduce = npb_cocluster followed by the in-place sum of S_dev over the communicator's ranks, on the
 * context's stream; npb_comm_allreduce_sum reduces any device buffer of 32- or 64-bit floats in place (R-hat partial sums,
 * score sums).  In a single process the collective calls of all ranks go between npb_comm_group(1) and npb_comm_group(0). */
typedef struct npb_comm npb_comm;
npb_status npb_comm_unique_id(char out[128]);
npb_status npb_comm_create(npb_ctx *ctx, const char id[128], int rank, int world, npb_comm **out);
npb_status npb_comm_create_all(npb_ctx *const *ctxs, int n, npb_comm **out /* [n] */);
npb_status npb_comm_destroy(npb_comm *c);
npb_status npb_comm_allreduce_sum(npb_comm *c, void *dev_buf, int64_t count, int bits);
npb_status npb_comm_group(int start);
npb_status npb_cocluster_allreduce(npb_chains *ch, const int64_t *anchors, int64_t n_anchor, npb_comm *comm_or_null, float *S_dev);
npb_status npb_cocluster_host(npb_chains *ch, const int64_t *anchors, int64_t n_anchor, npb_comm *comm_or_null, float *S_host);

/* The scan order of sweep number `sweep` (replaces the per-sweep std::shuffle of np_mcmc.cpp:120-125): a keyed
 * permutation of 0..N-1 shared by all chains of a run; order_out[s] = item visited at step s.  Pure host
 * function (no device needed); the kernels evaluate the same function point-wise. */
npb_status npb_scan_order_host(uint64_t seed, uint32_t sweep, int64_t N, int32_t *order_out);

/* measured FP32 FFMA peak of the context's device in TFLOP/s (register-resident FMA loop, best of 10): the
 * roofline denominator of the sweep kernels, which are bound by the FP32 pipe */
npb_status npb_fp32_peak(npb_ctx *ctx, double *tflops);

int64_t npb_chains_count(npb_chains *ch);
int npb_chains_kmax(npb_chains *ch);

#ifdef __cplusplus
}
#endif
#endif /* NPB200_H */
