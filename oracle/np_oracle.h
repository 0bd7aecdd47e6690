/*
 * np_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement (no Eigen) of the
 * Gibbs-reassignment hot path of mrquincle/noparama, used as the checker for
 * the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product
 * (noparama_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   density (a4)            : PINNED by the reference KAT test/test_mvn_likelihood.cpp:33,44
 *   pick (a6)               : semantics pinned by test/test_weighted_vector.cpp:10-36
 *   membertrix (a7)         : invariant pinned by test/test_membertrix.cpp:75-91
 *   control flow + RNG order: PINNED against the reference's own sources compiled
 *                             against oracle/eigen_shim (oracle/_ref, see oracle/Makefile)
 *   Eigen numerics (LU inverse/determinant, LLT, eigensolver): restated from
 *                             Eigen 3's published algorithms; Eigen itself is
 *                             absent from the image => ulp-level agreement unpinned.
 */
#ifndef NP_ORACLE_H
#define NP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { NPO_ALG8 = 8, NPO_JAIN_NEAL = 2, NPO_TRIADIC = 3 }; /* JN / triadic: np_oracle_sm.inc */

/* option flags */
enum {
	NPO_DENSE_MATRIX      = 1 << 0, /* keep the N x cols bool matrix of membertrix.h:30 (reference cost profile) */
	NPO_PER_CALL_LU       = 1 << 1, /* LU inverse + determinant on every density call (multivariatenormal.cpp:87,90) */
	NPO_UPDATE_CLUSTERS   = 1 << 2, /* run UpdateClusters::update (np_mcmc.cpp:170); no effect on theta (Q1) but consumes RNG */
	NPO_MAX_LIKELIHOOD    = 1 << 3, /* run considerMaxLikelihood every 5 sweeps (np_mcmc.cpp:172-174) */
	NPO_RECORD_TRACE      = 1 << 4, /* record the Alg. 8 replay trace */
	NPO_LOG_DOMAIN        = 1 << 5, /* NOT the reference: log-domain weights (no underflow); used for high-D statistical parity */
	NPO_FAITHFUL          = NPO_DENSE_MATRIX | NPO_PER_CALL_LU | NPO_UPDATE_CLUSTERS | NPO_MAX_LIKELIHOOD
};

typedef struct npo_prior {
	int D;
	const double *mu0;    /* [D]   np_main.cpp:368 */
	double kappa;         /*       np_main.cpp:369 */
	double nu;            /*       np_main.cpp:370 */
	const double *Lambda; /* [D,D] np_main.cpp:371, row-major */
	double alpha;         /*       np_main.cpp:164 */
} npo_prior;

typedef struct npo_options {
	int algorithm;      /* NPO_ALG8 | NPO_JAIN_NEAL | NPO_TRIADIC  (np_main.cpp:424-459) */
	int T;              /* sweeps (np_main.cpp:242) */
	int K0;             /* initial clusters, 20 in np_mcmc.cpp:49 */
	int M_aux;          /* auxiliary draws, 3 in np_neal_algorithm8.cpp:33 */
	int mh_steps;       /* 20 in np_mcmc.cpp:54 */
	uint32_t seed_main; /* replaces random_device at np_main.cpp:180 */
	uint32_t seed_shuffle; /* replaces the random_device seed of the static mt19937, dim1algebra.hpp:2069-2070 */
	int flags;
} npo_options;

typedef struct npo_stats {
	double seconds_total;      /* whole MCMC::run loop */
	double seconds_reassign;   /* only the update() loop, np_mcmc.cpp:146-163 */
	int64_t updates;           /* calls of sampler.update() */
	int64_t density_evals;     /* calls of probability()/logprobability() on a single datum */
	int64_t candidates;        /* Alg. 8: sum over steps of (K_i + M) */
	int64_t new_cluster_events;
	int64_t moved;             /* Alg. 8: steps that changed the item's cluster */
	int64_t sm_attempts[4];    /* split/merge: merge21, split12, merge32, split23 (JN uses [0],[1]) */
	int64_t sm_accepts[4];
	int64_t sams_allocations;  /* items allocated by SAMS over all proposals */
	double  mean_K;            /* mean occupied clusters over all update() calls */
	int K_final;
	double max_loglik;
} npo_stats;

typedef struct npo_run npo_run; /* opaque result of one run */

/* ---- density (multivariatenormal.cpp:64-146) ---- */
double npo_mvn_pdf(int D, const double *mu, const double *Sigma, const double *x);
double npo_mvn_logpdf(int D, const double *mu, const double *Sigma, const double *x);
double npo_mvn_pdf_dataset(int D, const double *mu, const double *Sigma, const double *X, int n);
double npo_mvn_logpdf_dataset(int D, const double *mu, const double *Sigma, const double *X, int n);
/* batch helper for tests: out[i*K+k] = logpdf(X[i] | mu[k], Sigma[k]) */
void npo_mvn_logpdf_batch(int D, const double *mu, const double *Sigma, int K, const double *X, int n, double *out);

/* ---- categorical pick (dim1algebra.hpp:2078-2104) ---- */
int npo_weighted_pick_u(const double *w, int n, double u);           /* explicit uniform */
void npo_weighted_pick_freq(const double *w, int n, int draws, uint32_t seed, int64_t *freq); /* test_weighted_vector.cpp */

/* ---- base measure (dirichlet.h:91-93 -> normalinvwishart.h:44-64) ---- */
/* draws `count` thetas with a fresh process (fresh static distributions) from engine seed `seed`;
 * mu_out [count,D], Sigma_out [count,D,D] */
void npo_sample_base(const npo_prior *prior, uint32_t seed, int count, double *mu_out, double *Sigma_out);

/* ---- linear algebra restated from Eigen (for tests) ---- */
double npo_lu_determinant(int n, const double *A);
void npo_lu_inverse(int n, const double *A, double *Ainv);

/* ---- metrics (clustering_performance.cpp:14-82) ---- */
void npo_metrics(const int *truth, const int *result, int N, double out3[3]); /* purity, RI, ARI */

/* ---- membertrix invariant test (test/test_membertrix.cpp) ---- */
int npo_membertrix_selftest(uint32_t seed, int dense);

/* ---- the sampler (np_mcmc.cpp:48-175 with np_neal_algorithm8.cpp / np_jain_neal_algorithm.cpp / np_triadic_algorithm.cpp) ---- */
npo_run *npo_mcmc_run(const npo_prior *prior, const npo_options *opt, const double *X /* [N,D] */, int N);
/* NOT the reference: the same run started from K caller-supplied clusters (mu [K,D], Sigma [K,D,D]) instead of K0 prior
 * draws; used by bench.py for the CPU figure in the D = 16 regime (the reference's own prior collapses to one cluster there) */
npo_run *npo_mcmc_run_given(const npo_prior *prior, const npo_options *opt, const double *X, int N, int K, const double *mu,
		const double *Sigma);
void npo_run_free(npo_run *r);
void npo_run_stats(const npo_run *r, npo_stats *out);
void npo_run_assignments(const npo_run *r, int which /*0 final, 1 max-likelihood*/, int *z_out /*[N], compact labels*/);
/* per-sweep wall time: the update() loop only (np_mcmc.cpp:146-163) and the whole sweep body; [T] each */
void npo_run_sweep_seconds(const npo_run *r, double *reassign, double *total);
int npo_run_params(const npo_run *r, int *K, double *mu /*[K,D]*/, double *Sigma /*[K,D,D]*/, int64_t *counts, int cap);

/* cluster count after every sampler.update() call (needs NPO_RECORD_TRACE) */
int64_t npo_run_K_after_len(const npo_run *r);
void npo_run_K_after(const npo_run *r, int *out);

/* initial state (after np_mcmc.cpp:49-91), in SLOT numbering: for GPU replay */
int npo_run_init_K(const npo_run *r);
void npo_run_init_state(const npo_run *r, int *z0 /*[N] slot ids*/, int *slots /*[K]*/, double *mu /*[K,D]*/, double *Sigma /*[K,D,D]*/);

/* Split-merge replay trace (needs NPO_RECORD_TRACE; Jain-Neal / triadic), one entry per proposal p < n:
 * picks [n,3] the subset MCMC::run handed to update() (-1 pad), u0 [n] the triadic sampler's first uniform, type [n]
 * (0 JN split, 1 JN merge, 2 triadic split, 3 triadic merge), th_mu [n,D] / th_sigma [n,D,D] the prior draw of a split,
 * the pool pool[pool_off[p] .. pool_off[p+1]) in visiting order with us (the uniform an allocation consumed, -1 for a
 * skipped pick) and dec (the part chosen), logA, the acceptance uniform uacc, accept, and the slot a new cluster got. */
int64_t npo_sm_trace_proposals(const npo_run *r);
int64_t npo_sm_trace_pool_len(const npo_run *r);
void npo_sm_trace_copy(const npo_run *r, int *picks, double *u0, int *type, double *th_mu, double *th_sigma, int64_t *pool_off,
		int *pool, double *us, int *dec, double *logA, double *uacc, int *accept, int *new_slot);

/* Alg. 8 replay trace.  steps = T*N. All cluster references are SLOT ids. */
int64_t npo_trace_steps(const npo_run *r);
int64_t npo_trace_order_len(const npo_run *r);
int npo_trace_max_slot(const npo_run *r); /* number of slots ever used */
void npo_trace_copy(const npo_run *r,
		int *item /*[S]*/, int *K /*[S]*/, int64_t *order_off /*[S+1]*/, int *order /*[order_len]*/,
		double *aux_mu /*[S,M,D]*/, double *aux_Sigma /*[S,M,D,D]*/, double *u /*[S]*/,
		int *picked /*[S]*/, int *new_slot /*[S], -1 if existing*/, int *z_after /*[T,N] slot ids*/);

/* ---- conjugate Algorithm 2 (np_oracle_alg2.inc): NOT in the reference (np_neal_algorithm2.cpp is dead code, the conjugate
 * headers are empty) => PARITY UNPINNED against the reference; pinned against scipy.stats.multivariate_t instead ---- */
double npo_niw_logpred(const npo_prior *prior, int n, const double *Xm /*[n,D]*/, const double *x);
double npo_niw_logpred_incremental(const npo_prior *prior, int n, const double *Xm, int remove_first, const double *x);
void npo_alg2_run(const npo_prior *prior, const double *X, int N, int T, int K0, uint64_t seed, int *z_out, int *K_trace,
		int64_t *moved_out, int64_t *births_out);

/* ---- scalar-noise likelihood families (`-c regression` / `-c angular`, np_main.cpp:196-205): scalarnoise_multivariatenormal.cpp
 * + normalinvgamma.h + gamma.h.  PINNED against the reference's own sources (oracle/_ref np_ref_run ... FAMILY): same assignments
 * after every run of tests/test_oracle_scalarnoise.py ---- */
enum { NPO_FAMILY_MVN = 0, NPO_FAMILY_REGRESSION = 1, NPO_FAMILY_ANGULAR = 2 };
double npo_scalarnoise_logpdf(int family, const double *mu /*[2]*/, double sigma, const double *x /*[3] (1, a, b) | [2]*/);
double npo_scalarnoise_pdf(int family, const double *mu, double sigma, const double *x);
void npo_sample_base_nig(const double *mu0 /*[2]*/, const double *Lambda /*[2,2]*/, double ig_alpha, double ig_beta, uint32_t seed, int count,
		double *mu_out /*[count,2]*/, double *sigma_out /*[count]*/);
npo_run *npo_mcmc_run_scalarnoise(int family, const double *mu0, const double *Lambda, double ig_alpha, double ig_beta, double dp_alpha,
		const npo_options *opt, const double *X /*[N, 3 | 2]*/, int N);
int npo_run_params_scalarnoise(const npo_run *r, int *K, double *mu /*[K,2]*/, double *sigma /*[K]*/, int64_t *counts, int cap);

#ifdef __cplusplus
}
#endif
#endif
