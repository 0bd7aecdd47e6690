// npb_internal.h -- host-side structs behind the opaque handles of include/npb200.h
#pragma once
#include <nvtx3/nvToolsExt.h>
#include "../../include/npb200.h"
#include "npb_common.cuh"
#include <vector>

typedef uint16_t npb_z_t; // slot id of an item; Kmax <= 65535

struct PriorHost {
	bool set = false;
	int D = 0, flags = 0;
	double kappa = 0, nu = 0, alpha = 0;
	std::vector<double> mu0, Lambda;
	std::vector<double> CT;  // C^T packed upper (A^-1 = C C^T)
	std::vector<double> S;   // C^-T packed upper
	double logdetA = 0;
	// scalar-noise families (npb_prior_set_nig): family != 0, D = width of a data row, mu0 [2], Lambda [2, 2]
	int family = 0;
	double ig_alpha = 0, ig_beta = 0;
};

// NVTX range around every public entry point that launches device work (visible in Nsight Systems / ncu --nvtx; free when no
// tool is attached: the header-only NVTX 3 stubs return at once)
struct NpbRange {
	explicit NpbRange(const char *name) { nvtxRangePushA(name); }
	~NpbRange() { nvtxRangePop(); }
	NpbRange(const NpbRange &) = delete;
	NpbRange &operator=(const NpbRange &) = delete;
};

struct npb_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	char err[512] = {0};
	int n_sm = 0;               // multiprocessors of this context's device (grid size of the persistent kernels)
	bool a2_tile_attr_set = false; // the same for k_a2_tile (npb_alg2_tile.cu)
	bool a2_tc_attr_set = false;   // and for k_a2_tc (npb_alg2_tc.cu)
	bool a2_tc16_attr_set = false; // and for k_a2_tc16 (npb_alg2_tc16.cu)
	bool fused_attr_set = false; // per device: the shared-memory opt-in of k_sweep_tc16 has been made on this context's device
	PriorHost prior;
	float *d_CT2 = nullptr, *d_S = nullptr; // device copies of the packed prior factors
	uint64_t prior_epoch = 0;               // bumped by npb_prior_set_niw; datasets re-whiten lazily
};

struct npb_dataset {
	npb_ctx *ctx = nullptr;
	int64_t N = 0;
	int D = 0;
	double *X64 = nullptr;  // [N,D]
	float *X32 = nullptr;   // [N,D]
	float *Xw = nullptr;    // [N,D] whitened against the prior: C^T (x - mu0) * sqrt(log2e/2)
	float *Xwn = nullptr;   // [N] Euclidean norm of the whitened row
	uint64_t whitened_epoch = 0;
	double *h_stage = nullptr; // pinned staging for uploads
	double *Xbar = nullptr;    // [2 D + 1] column means, scale exponent, column maxima (tensor paths: operands are centred before the split)
	bool xbar_valid = false;   // Xbar describes the current contents (npb_dataset_update invalidates it; the buffer is kept)
	int n_chains_alive = 0;    // chain handles that reference this dataset
};

struct npb_chains {
	npb_ctx *ctx = nullptr;
	npb_dataset *ds = nullptr;
	int64_t C = 0;
	int Kmax = 0, m_aux = 0, K0 = 0, D = 0;
	uint64_t seed = 0;
	uint32_t sweep = 0;         // sweeps done so far (Philox counter / scan-order key)
	char opt_d16_path[8] = {0}; // NPB_D16_PATH as read when the handle was created (auto / tc / tc2 / fp32)
	// behaviour / measurement switches: the NPB_* environment variables are read ONCE, in npb_chains_create
	// (npb_switches_from_env), and changed afterwards only through npb_chains_set_option; no launch path reads the environment
	struct Switches {
		int spec = 1;             // NPB_D64_SPEC: 0 sequential evaluation, 1 speculation (default), 2 table only (measurement)
		int f16_flags = 0;        // NPB_F16_FLAGS: A/B measurement bits of k_sweep_tc16 (results unchanged)
		int d16_block = 8192;     // NPB_D16_BLOCK: steps per launch of the D = 16 tensor paths
		int d64_block = 4096;     // NPB_D64_BLOCK: steps per launch of the D = 64 path
		int d16_epi = 16;         // NPB_D16_EPI: epilogue warps of k_density_tc16 (8 | 16)
		int d16_nh = 4;           // NPB_D16_NH: chain halves per unit of k_density_tc16 (1 | 2 | 4)
		bool d16_aux_pre = false; // NPB_D16_AUX=pre: the kernel pair reads exact auxiliary keys from the k_aux_keys pre-pass (round 1)
		bool d16_aux_grp = true;  // NPB_D16_AUX=lazy -> false: no pre-pass at all, bounds per step inside the race; default: k_aux_bound group maxima
		bool d16_aux_auto = true; // no explicit d16_aux: the kernel pair takes the bounds per step inside the race (lazy) in a chain that keeps moving items
		                          // (there the race waits on its table reads and has issue slots to spare: 101 against 106 ms per sweep), group maxima otherwise
		bool d64_overlap = true;  // NPB_D64_OVERLAP=0 -> false
		bool d64_fp32 = false;    // NPB_D64_DENSITY=fp32
		bool two_warp = false;    // NPB_TILE_KERNEL=2warp: round-1 one-chain-per-CTA kernel
		int a2_tile = 128;        // NPB_A2_TILE: conjugate Algorithm 2 at D = 16, 64: steps evaluated ahead of the chain (1 .. 128, same results; k_a2_tile takes at most 64); 0 = the step-at-a-time kernel k_a2_sweep
		bool a2_tc = true;        // NPB_A2_TC=0 -> false: D = 64 on the FP32 tile kernel k_a2_tile instead of the tcgen05 kernel k_a2_tc
		bool a2_tc16 = false;     // NPB_A2_TC16=1 -> true: D = 16 on the tcgen05 kernel k_a2_tc16 instead of the FP32 tile kernel k_a2_tile
	} sw;
	double moved_frac_last = -1.0; // moved / reassignments of the last Algorithm 8 launch whose statistics were read; -1 unknown
	bool time_kernels = false;  // option "time_kernels": CUDA events around every launch of the dominant sweep kernel
	std::vector<cudaEvent_t> kt_ev; // pairs (start, stop) not yet read back
	double kt_ms = 0.0;         // accumulated duration of the launches read back so far
	int64_t kt_launches = 0;
	bool counted = false;       // this handle is counted in ds->n_chains_alive
	uint32_t init_epoch = 0;    // initialisations from given parameters so far (distinct initial assignments per call)
	uint32_t item_calls = 0;    // single-item updates so far (Philox counter of npb_chain_update_alg8)
	npb_z_t *z = nullptr;       // [N, C] item-major
	float *theta = nullptr;     // [C, Kmax, PS]
	int *counts = nullptr;      // [C, Kmax]
	unsigned long long *st = nullptr; // [C, 4]: candidates, moved, births, reserved
	int *kocc = nullptr;        // [C]
	int *overflow = nullptr;    // [C]
	npb_z_t *h_z = nullptr;     // pinned staging for npb_chains_sweep_host
	int32_t *scan_order = nullptr; // [scan_cap, N] item visited at each step of the sweeps of one launch
	int scan_cap = 0;              // sweeps per launch the buffer holds
	uint32_t *aux_keys = nullptr;  // [aux_cap, C, N] packed race key of the best auxiliary draw of every (chain, step)
	int aux_cap = 0;               // sweeps per launch that buffer holds
	float *aux_max = nullptr;      // [aux_cap, C, ceil(N / 32)] (fused D = 16 kernel)
	// split-merge samplers (npb_splitmerge.cu)
	unsigned long long *smst = nullptr; // [C, 12]: attempts[4], accepts[4], SAMS allocations, proposals, 2 reserved
	npb_z_t *sm_zt = nullptr;      // [C, zstride] chain-major working copy of z
	int32_t *sm_pool = nullptr;    // [C, N]
	uint8_t *sm_dec = nullptr;     // [C, N]
	int32_t *sm_order = nullptr;   // [3, N]
	float *sm_detail = nullptr;    // [C, 16] detail of the last proposal of every chain (tests)
	// max-likelihood snapshot (MCMC::considerMaxLikelihood, np_mcmc.cpp:187-203)
	npb_z_t *best_z = nullptr;     // [N, C]
	float *best_theta = nullptr;   // [C, Kmax, PS]
	int *best_counts = nullptr;    // [C, Kmax]
	double *best_jll = nullptr;    // [C] device: joint log-likelihood of the kept state (-inf before the first call)
	double *cur_jll = nullptr;     // [C] device scratch
	// parameter update (npb_params.cu)
	double *pstats = nullptr;      // [C, Kmax, D + D(D+1)/2] sum x, upper triangle of sum x x^T
	double *pLambda0 = nullptr;    // [D, D]
	int *pfail = nullptr;
	uint32_t param_epoch = 0;
	// D = 64 path (npb_alg8_gemm.cu)
	uint8_t *g_aimg = nullptr;     // [g_bs / 128][2][32 KB] A-operand images of the current block of steps
	uint8_t *g_bimg = nullptr;     // [C, 32][24 KB] B-operand images of the slots
	float *g_bconst = nullptr;     // [C, 32][68] nb = -T2 (mu - xbar), c2
	float *g_L = nullptr;          // [2][C][g_bs + 32][32] log2-density tables of the current and the next block
	uint8_t *g_dirty = nullptr;    // [C, 32] slots whose image is out of date
	int g_bs = 0;                  // steps per block
	uint32_t g_k = 0;              // blocks consumed so far (parity selects the table / born-mask buffer)
	uint32_t *g_born = nullptr;    // [2][C] slots born during block k (buffer k & 1)
	npb_z_t *z_prev = nullptr;     // [N, C] npb_chains_sweep_host_delta: the assignments the caller's mirror holds
	uint32_t *dl_idx = nullptr;    // [dl_cap] indices (item * C + chain) of the entries that changed
	npb_z_t *dl_val = nullptr;     // [dl_cap] their new values
	unsigned long long *dl_count = nullptr; // device counter
	size_t dl_cap = 0;
	uint32_t *h_dl_idx = nullptr;  // pinned staging, grown on demand
	npb_z_t *h_dl_val = nullptr;
	size_t h_dl_cap = 0;
	// conjugate Algorithm 2 (npb_alg2.cu)
	uint64_t z_gen = 0;            // bumped by everything that changes the assignments outside npb_alg2.cu
	uint64_t a2_gen = ~0ull;       // z_gen the statistics below correspond to
	uint64_t a2_prior_epoch = ~0ull;
	double *a2_sx = nullptr, *a2_sxx = nullptr, *a2_work = nullptr, *a2_prior = nullptr;
	float *a2_mu = nullptr, *a2_P = nullptr, *a2_ld = nullptr, *a2_G = nullptr, *a2_lp0 = nullptr;
	float a2_ld0 = 0.0f;
	npb_z_t *g_zblk = nullptr;     // [C][g_bs] fused D = 16 kernel: assignments of the current block's items, in step order
};

struct SweepArgs {
	const float *X, *Xw, *Xwn;
	const int32_t *scan_order; // [n_sweeps, N]
	const uint32_t *aux_keys;  // [n_sweeps, C, N] (k_aux_keys) or NULL
	float *aux_max;            // [n_sweeps, C, aux_groups] largest auxiliary key of every 32 consecutive steps, or NULL
	int aux_groups;            // ceil(N / 32)
	npb_z_t *z;
	float *theta;
	int *counts;
	unsigned long long *st;
	int *kocc, *overflow;
	int N, C, Kmax;
	uint32_t sweep0;
	int n_sweeps;
	uint64_t seed;
	PriorDev prior;
};

// arguments of the split-merge replay kernel (npb_replay_sm.cu)
struct RsmArgs {
	const double *X;
	int N, D, sampler, nslots;
	double alpha;
	double *theta;      // [nslots][PSR]
	int *counts;        // [nslots]
	int32_t *z;         // [N]
	int64_t n_prop;
	const int32_t *picks;   // [n,3]
	const double *u0;       // [n]
	const double *th_new;   // [n][PSR]
	const int64_t *pool_off;
	const int32_t *pool;
	const double *us;
	const double *uacc;
	const int32_t *new_slot;
	int32_t *type_out, *dec_out, *accept_out;
	double *logA_out;
	int *status;
};


npb_status npb_fail_cuda(npb_ctx *ctx, cudaError_t e, const char *expr, const char *file, int line);
npb_status npb_fail(npb_ctx *ctx, npb_status s, const char *msg);

// host linear algebra (double), npb_linalg.cpp
bool npb_prepare_theta(int D, const double *mu, const double *Sigma, double *T_packed_upper, double *logdet);
bool npb_prepare_prior(PriorHost &p);
void npb_theta_to_sigma(int D, const double *T_packed_upper, double *Sigma);

// launchers (npb_alg8.cu / npb_density.cu / npb_metrics.cu)
npb_status npb_launch_whiten(npb_dataset *ds);
npb_status npb_launch_chains_init(npb_chains *ch, int K0, const float *d_theta_given);
npb_status npb_launch_alg8_sweep(npb_chains *ch, int n_sweeps);
npb_status npb_launch_tile_probe(npb_chains *ch, int chain, const int32_t *d_items, float *d_out);
struct SweepArgs;
npb_status npb_launch_alg8_gemm64(npb_chains *ch, const SweepArgs &a);
npb_status npb_launch_gemm64_probe(npb_chains *ch, int chain, const int32_t *d_items, float *d_out);
npb_status npb_launch_alg8_tc16(npb_chains *ch, const SweepArgs &a);
npb_status npb_launch_tc16_probe(npb_chains *ch, int chain, const int32_t *d_items, float *d_out);
npb_status npb_launch_alg8_fused16(npb_chains *ch, const SweepArgs &a);
npb_status npb_launch_scan_order(npb_chains *ch, int n_sweeps);
// scalar-noise likelihood families (npb_scalarnoise.cu)
npb_status npb_launch_sn_init(npb_chains *ch, int K0);
npb_status npb_launch_sn_sweep(npb_chains *ch, int n_sweeps);
npb_status npb_launch_sn_sample_base(npb_chains *ch, int chain, int count, float *d_out);
void npb_sn_slot_from_raw(int family, int D, const double *mu, double sigma, double *mu_slot, double *T_packed, double *cst);
npb_status npb_launch_alg2_conjugate(npb_chains *ch, int n_sweeps);
npb_status npb_launch_alg2_probe(npb_chains *ch, int chain, const int32_t *d_items, int n_items, float *d_out);
npb_status npb_launch_fused16_probe(npb_chains *ch, int chain, const int32_t *d_items, float *d_out);
npb_status npb_launch_update_item(npb_chains *ch, int64_t chain0, int64_t n, int64_t item);
npb_status npb_launch_update_params(npb_chains *ch, int mode, const double *mu0, double kappa0, double nu0, const double *Lambda0);
npb_status npb_launch_split_merge(npb_chains *ch, int sampler, int64_t n_proposals, int whole_sweeps, float *d_detail);
PriorDev npb_prior_dev(const npb_ctx *ctx, int m_aux);
