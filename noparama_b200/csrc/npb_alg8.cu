// npb_alg8.cu -- Neal Algorithm 8 sweeps for thousands of lockstep chains (sm_100a).
//
// Replaces, per chain and per item, NealAlgorithm8::update (src/np_neal_algorithm8.cpp:49-167):
//   retract (membertrix.cpp:175-233), M draws from the base measure (np_neal_algorithm8.cpp:79-84 ->
//   normalinvwishart.h:44-64), K weighted densities p(x|theta_k) n_k (:93-109) and M densities
//   p(x|theta'_m) alpha/M (:119-126), random_weighted_pick (dim1algebra.hpp:2078-2104), assign or addCluster
//   (:136-157) -- inside the sweep loop of MCMC::run (np_mcmc.cpp:109-163).
//
// Layout: one warp per chain.  Cluster slot `s` of a chain lives on lane (s & 31), register level (s >> 5):
// its member count, mean, triangular precision factor and log-normaliser stay in registers for the whole
// launch (D <= 3) so a step touches no memory but the item's coordinates and its assignment.  The scan order
// of a sweep is a keyed permutation shared by all chains (npb_common.cuh), evaluated 32 steps at a time:
// lane j of the warp prefetches item, old assignment and coordinates of step s0+j and draws that step's
// auxiliary parameters from Philox, so the sequential part of a step is: weights -> warp scan -> pick ->
// count update.  Weights are formed in the log2 domain (no underflow) and summed in the same left-to-right
// "first cumulative weight >= u * total" rule as the reference.
#include "npb_internal.h"

#define NPB_SWEEP_WARPS 2

__device__ __forceinline__ float warp_max(float v) {
	int i = __float_as_int(v);
	i ^= (i >> 31) & 0x7fffffff; // order-preserving map float -> signed int
	i = __reduce_max_sync(0xffffffffu, i);
	i ^= (i >> 31) & 0x7fffffff;
	return __int_as_float(i);
}
__device__ __forceinline__ float warp_inclusive_sum(float v, int lane) {
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		float t = __shfl_up_sync(0xffffffffu, v, o);
		if (lane >= o) v += t;
	}
	return v;
}
__device__ __forceinline__ float fast_ex2(float x) {
	float y;
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
	float y;
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}

// normal number `f` (flat index) of the Philox stream (c0, *, c2, c3): call f>>2, Box-Muller pair (f>>1)&1
__device__ inline float npb_normal_at(const Philox &ph, uint32_t c0, uint32_t c2, uint32_t c3, int f) {
	uint32_t w[4];
	ph(c0, (uint32_t)(f >> 2), c2, c3, w);
	float a, b;
	if (f & 2) npb_normal2(w[2], w[3], a, b); else npb_normal2(w[0], w[1], a, b);
	return (f & 1) ? b : a;
}

// One draw from the base measure in device form (generic D, used by init and by the generic-D kernels):
// v = D + nu*g0 ; mu = mu0 + (|v|/sqrt(kappa)) S g[1..D] ; T2 = CT2/|v| ; c2 = c0_2 - D log2|v|
// normals are numbers f0 .. f0+D of the stream (c0,*,c2,c3).
__device__ inline void npb_draw_theta(const PriorDev &pr, const Philox &ph, uint32_t c0, uint32_t c2c, uint32_t c3, int f0,
		float *out /* [PS]: mu, T2, c2 */) {
	const int D = pr.D, TRI = npb_tri(D);
	float v = pr.v_mean + pr.nu * npb_normal_at(ph, c0, c2c, c3, f0);
	float av = fmaxf(fabsf(v), 1e-20f);
	float sc = av * pr.inv_sqrt_kappa;
	for (int i = 0; i < D; ++i) out[i] = pr.mu0[i];
	for (int j = 0; j < D; ++j) {
		float g = npb_normal_at(ph, c0, c2c, c3, f0 + 1 + j) * sc;
		for (int i = 0; i <= j; ++i) out[i] += pr.S[npb_tri_off(D, i, j)] * g;
	}
	float inv = 1.0f / av;
	for (int t = 0; t < TRI; ++t) out[D + t] = pr.CT2[t] * inv;
	out[D + TRI] = pr.c0_2 - (float)D * log2f(av);
}

// ---------------------------------------------------------------------------------------------------------
// init: np_mcmc.cpp:49-91 -- K0 clusters from the base measure (np_init_clusters.cpp:24-41), every item to a
// uniformly chosen cluster, clusters that got nothing dropped (cleanup, membertrix.cpp:343-364).
// One warp per chain.
// ---------------------------------------------------------------------------------------------------------
__global__ void k_chains_init(SweepArgs a, int K0) {
	extern __shared__ int sh_counts[]; // [warps][Kmax]
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int chain = blockIdx.x * (blockDim.x >> 5) + warp;
	if (chain >= a.C) return;
	const int D = a.prior.D, PS = npb_ps(D);
	int *cnt = sh_counts + warp * a.Kmax;
	for (int k = lane; k < a.Kmax; k += 32) cnt[k] = 0;
	__syncwarp();
	Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
	float *th = a.theta + (size_t)chain * a.Kmax * PS;
	for (int k = lane; k < a.Kmax; k += 32) {
		float *o = th + (size_t)k * PS;
		if (k < K0) npb_draw_theta(a.prior, ph, (uint32_t)k, 0u, NPB_RNG_INIT_THETA, 0, o);
		else for (int t = 0; t < PS; ++t) o[t] = 0.0f;
	}
	for (int i = lane; i < a.N; i += 32) {
		uint32_t w[4];
		ph((uint32_t)i, 0u, 0u, NPB_RNG_INIT_Z, w);
		int k = (int)__umulhi(w[0], (uint32_t)K0);
		a.z[(size_t)i * a.C + chain] = (npb_z_t)k;
		atomicAdd(&cnt[k], 1);
	}
	__syncwarp();
	int occ = 0;
	for (int k = lane; k < a.Kmax; k += 32) {
		a.counts[(size_t)chain * a.Kmax + k] = cnt[k];
		occ += cnt[k] > 0;
	}
	occ = __reduce_add_sync(0xffffffffu, occ);
	if (lane == 0) {
		a.kocc[chain] = occ;
		a.overflow[chain] = 0;
		a.st[(size_t)chain * 4 + 0] = a.st[(size_t)chain * 4 + 1] = a.st[(size_t)chain * 4 + 2] = a.st[(size_t)chain * 4 + 3] = 0ull;
	}
}

// ---------------------------------------------------------------------------------------------------------
// whitening of the dataset against the prior: Xw = CT2 (x - mu0)  (CT2 upper triangular packed)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_whiten(const float *X, float *Xw, int64_t N, PriorDev pr) {
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	const int D = pr.D;
	const float *x = X + i * D;
	float *o = Xw + i * D;
	for (int r = 0; r < D; ++r) {
		float s = 0.0f;
		for (int c = r; c < D; ++c) s += pr.CT2[npb_tri_off(D, r, c)] * (x[c] - pr.mu0[c]);
		o[r] = s;
	}
}

// ---------------------------------------------------------------------------------------------------------
// The register-resident sweep kernel: D <= 3, Kmax = 32*SPL slots, M auxiliary draws.
// ---------------------------------------------------------------------------------------------------------
template <int D, int SPL, int M>
__global__ void __launch_bounds__(NPB_SWEEP_WARPS * 32) k_alg8_sweep_reg(const SweepArgs a) {
	constexpr int TRI = npb_tri(D), PS = npb_ps(D);
	constexpr int NN = M * (D + 1);   // normals per step
	constexpr int NC = (NN + 3) / 4;  // Philox calls per step for the normals
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int chain = blockIdx.x * NPB_SWEEP_WARPS + warp;
	if (chain >= a.C) return;
	const int N = a.N, C = a.C;

	// ---- chain state into registers ----
	float n[SPL], mu[SPL][D], T[SPL][TRI], c2[SPL];
	{
		const float *th = a.theta + (size_t)chain * a.Kmax * PS;
		const int *cn = a.counts + (size_t)chain * a.Kmax;
#pragma unroll
		for (int s = 0; s < SPL; ++s) {
			const int slot = s * 32 + lane;
			n[s] = (float)cn[slot];
#pragma unroll
			for (int d = 0; d < D; ++d) mu[s][d] = th[(size_t)slot * PS + d];
#pragma unroll
			for (int t = 0; t < TRI; ++t) T[s][t] = th[(size_t)slot * PS + D + t];
			c2[s] = th[(size_t)slot * PS + D + TRI];
		}
	}
	int kocc = 0, nlev = 0;
#pragma unroll
	for (int s = 0; s < SPL; ++s) {
		unsigned b = __ballot_sync(0xffffffffu, n[s] > 0.0f);
		kocc += __popc(b);
		if (b) nlev = s + 1;
	}
	unsigned long long st_cand = 0ull, st_moved = 0ull, st_births = 0ull;
	int overflow = 0;

	const Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
	const float ik2 = a.prior.inv_sqrt_kappa * (float)NPB_HALF_LOG2E_SQRT;
	const float fD = (float)D;

	for (int sw = 0; sw < a.n_sweeps; ++sw) {
		const uint32_t sweep = a.sweep0 + (uint32_t)sw;
		const ScanOrder so = npb_scan_order(a.seed, sweep, (uint32_t)N);
		for (int s0 = 0; s0 < N; s0 += 32) {
			// ---------------- tile prologue: lane j owns step s0 + j ----------------
			const int sj = s0 + lane;
			const bool valid = sj < N;
			const uint32_t item = valid ? npb_scan_item(so, (uint32_t)sj) : 0u;
			int zold_j = valid ? (int)a.z[(size_t)item * C + chain] : 0;
			float x_j[D], xw_j[D];
#pragma unroll
			for (int d = 0; d < D; ++d) {
				x_j[d] = a.X[(size_t)item * D + d];
				xw_j[d] = a.Xw[(size_t)item * D + d];
			}
			float g[NC * 4];
#pragma unroll
			for (int c = 0; c < NC; ++c) {
				uint32_t w[4];
				ph((uint32_t)sj, (uint32_t)c, sweep, NPB_RNG_AUX, w);
				npb_normal2(w[0], w[1], g[4 * c + 0], g[4 * c + 1]);
				npb_normal2(w[2], w[3], g[4 * c + 2], g[4 * c + 3]);
			}
			float u_j;
			{
				uint32_t w[4];
				ph((uint32_t)sj, 0u, sweep, NPB_RNG_PICK, w);
				u_j = npb_u01(w[0]);
			}
			float auxl_j[M], av_j[M];
#pragma unroll
			for (int m = 0; m < M; ++m) {
				float v = a.prior.v_mean + a.prior.nu * g[m * (D + 1)];
				float av = fmaxf(fabsf(v), 1e-20f);
				float inv = __frcp_rn(av);
				float q = 0.0f;
#pragma unroll
				for (int d = 0; d < D; ++d) {
					float y = xw_j[d] * inv - g[m * (D + 1) + 1 + d] * ik2;
					q = fmaf(y, y, q);
				}
				av_j[m] = av;
				auxl_j[m] = a.prior.c0_2 - fD * fast_lg2(av) - q + a.prior.log2_alpha_m;
			}
			int znew_j = zold_j;
			const int cnt = min(32, N - s0);

			// ---------------- the sequential part: one step per iteration ----------------
			for (int j = 0; j < cnt; ++j) {
				const int zo = __shfl_sync(0xffffffffu, zold_j, j);
				float xs[D];
#pragma unroll
				for (int d = 0; d < D; ++d) xs[d] = __shfl_sync(0xffffffffu, x_j[d], j);
				float al[M];
#pragma unroll
				for (int m = 0; m < M; ++m) al[m] = __shfl_sync(0xffffffffu, auxl_j[m], j);
				const float u = __shfl_sync(0xffffffffu, u_j, j);
				const int osub = zo >> 5, olane = zo & 31;

				// log2-weights of the occupied slots (count after removing the item itself)
				float l[SPL], ne[SPL];
				float mx = al[0];
#pragma unroll
				for (int m = 1; m < M; ++m) mx = fmaxf(mx, al[m]);
				bool mine_dead = false;
#pragma unroll
				for (int s = 0; s < SPL; ++s) {
					l[s] = -INFINITY;
					ne[s] = 0.0f;
					if (s < nlev) {
						const bool own = (s == osub) && (lane == olane);
						ne[s] = n[s] - (own ? 1.0f : 0.0f);
						float dd[D];
#pragma unroll
						for (int d = 0; d < D; ++d) dd[d] = xs[d] - mu[s][d];
						float q = 0.0f;
#pragma unroll
						for (int r = 0; r < D; ++r) {
							float y = 0.0f;
#pragma unroll
							for (int c = r; c < D; ++c) y = fmaf(T[s][npb_tri_off(D, r, c)], dd[c], y);
							q = fmaf(y, y, q);
						}
						l[s] = (ne[s] > 0.0f) ? (c2[s] - q) : -INFINITY;
						mine_dead = mine_dead || (own && ne[s] <= 0.0f);
						mx = fmaxf(mx, l[s]);
					}
				}
				const float Mx = fmaxf(warp_max(mx), -1e30f);
				const bool died = __any_sync(0xffffffffu, mine_dead);
				const int kafter = kocc - (died ? 1 : 0);

				float ea[M], A = 0.0f;
#pragma unroll
				for (int m = 0; m < M; ++m) {
					ea[m] = fast_ex2(al[m] - Mx);
					A += ea[m];
				}
				float e[SPL], p[SPL];
#pragma unroll
				for (int s = 0; s < SPL; ++s) {
					e[s] = (s < nlev) ? ne[s] * fast_ex2(l[s] - Mx) : 0.0f;
					p[s] = (s == 0) ? e[0] : p[s - 1] + e[s];
				}
				const float lt = p[SPL - 1];
				const float incl = warp_inclusive_sum(lt, lane);
				float excl = __shfl_up_sync(0xffffffffu, incl, 1);
				if (lane == 0) excl = 0.0f;
				const float Tslots = __shfl_sync(0xffffffffu, incl, 31);
				const float total = A + Tslots;
				const float target = u * total;

				int new_slot = -1, aux_pick = -1;
				bool born = false;
				if (target <= A && A > 0.0f) {
					// auxiliary candidates come first in the cumulative order
					float c = 0.0f;
					int last_pos = -1;
#pragma unroll
					for (int m = 0; m < M; ++m) {
						c += ea[m];
						if (ea[m] > 0.0f) {
							last_pos = m;
							if (aux_pick < 0 && c >= target) aux_pick = m;
						}
					}
					if (aux_pick < 0) aux_pick = last_pos;
				} else {
					unsigned b = __ballot_sync(0xffffffffu, (A + incl >= target) && (lt > 0.0f));
					int wl, sub_w = -1;
					if (b) {
						wl = __ffs(b) - 1;
						int last_pos = -1;
#pragma unroll
						for (int s = 0; s < SPL; ++s) {
							if (e[s] > 0.0f) {
								last_pos = s;
								if (sub_w < 0 && (A + excl + p[s] >= target)) sub_w = s;
							}
						}
						if (sub_w < 0) sub_w = last_pos;
					} else {
						// rounding left the target above the last cumulative weight: take the last positive one
						unsigned b2 = __ballot_sync(0xffffffffu, lt > 0.0f);
						wl = b2 ? (31 - __clz(b2)) : -1;
#pragma unroll
						for (int s = 0; s < SPL; ++s)
							if (e[s] > 0.0f) sub_w = s;
					}
					if (wl >= 0) new_slot = __shfl_sync(0xffffffffu, sub_w * 32 + lane, wl);
					else aux_pick = 0; // nothing has weight at all: open a cluster
				}

				if (aux_pick >= 0) {
					// birth: lowest free slot after the removal (np_neal_algorithm8.cpp:136-145)
					int fsub = -1, fl = 0;
#pragma unroll
					for (int s = 0; s < SPL; ++s) {
						const bool own = (s == osub) && (lane == olane);
						const float cur = n[s] - (own ? 1.0f : 0.0f);
						unsigned fb = __ballot_sync(0xffffffffu, cur <= 0.0f);
						if (fsub < 0 && fb) { fsub = s; fl = __ffs(fb) - 1; }
					}
					if (fsub < 0) {
						overflow = 1;
						new_slot = zo; // no room: the item stays where it was
					} else {
						born = true;
						new_slot = fsub * 32 + fl;
						// the lane that owns this step re-derives theta' of the picked auxiliary draw
						float avp = 1.0f, gz[D];
#pragma unroll
						for (int m = 0; m < M; ++m)
							if (m == aux_pick) {
								avp = av_j[m];
#pragma unroll
								for (int d = 0; d < D; ++d) gz[d] = g[m * (D + 1) + 1 + d];
							}
						avp = __shfl_sync(0xffffffffu, avp, j);
						float munew[D];
#pragma unroll
						for (int d = 0; d < D; ++d) munew[d] = a.prior.mu0[d];
						const float sc = avp * a.prior.inv_sqrt_kappa;
#pragma unroll
						for (int c = 0; c < D; ++c) {
							const float gc = __shfl_sync(0xffffffffu, gz[c], j) * sc;
#pragma unroll
							for (int r = 0; r <= c; ++r) munew[r] = fmaf(a.prior.S[npb_tri_off(D, r, c)], gc, munew[r]);
						}
						const float inv = 1.0f / avp;
						const float c2new = a.prior.c0_2 - fD * log2f(avp);
#pragma unroll
						for (int s = 0; s < SPL; ++s)
							if (s == fsub && lane == fl) {
#pragma unroll
								for (int d = 0; d < D; ++d) mu[s][d] = munew[d];
#pragma unroll
								for (int t = 0; t < TRI; ++t) T[s][t] = a.prior.CT2[t] * inv;
								c2[s] = c2new;
							}
						if (fsub + 1 > nlev) nlev = fsub + 1;
					}
				}

				// a stay (same slot, no birth) changes nothing; everything else moves the item
				if (born || new_slot != zo) {
					if (new_slot != zo) {
						const int nsub = new_slot >> 5, nlane = new_slot & 31;
#pragma unroll
						for (int s = 0; s < SPL; ++s) {
							if (s == nsub && lane == nlane) n[s] += 1.0f;
							if (s == osub && lane == olane) n[s] -= 1.0f;
						}
					}
					if (died) kocc--;
					if (born) { kocc++; st_births++; }
					st_moved++;
				}
				st_cand += (unsigned long long)(kafter + M);
				if (lane == j) znew_j = new_slot;
			}
			// ---------------- tile epilogue ----------------
			if (valid) a.z[(size_t)item * C + chain] = (npb_z_t)znew_j;
		}
		__syncwarp();
	}

	// ---- chain state back to memory ----
	{
		float *th = a.theta + (size_t)chain * a.Kmax * PS;
		int *cn = a.counts + (size_t)chain * a.Kmax;
#pragma unroll
		for (int s = 0; s < SPL; ++s) {
			const int slot = s * 32 + lane;
			cn[slot] = (int)n[s];
#pragma unroll
			for (int d = 0; d < D; ++d) th[(size_t)slot * PS + d] = mu[s][d];
#pragma unroll
			for (int t = 0; t < TRI; ++t) th[(size_t)slot * PS + D + t] = T[s][t];
			th[(size_t)slot * PS + D + TRI] = c2[s];
		}
		if (lane == 0) {
			a.kocc[chain] = kocc;
			if (overflow) a.overflow[chain] = 1;
			a.st[(size_t)chain * 4 + 0] += st_cand;
			a.st[(size_t)chain * 4 + 1] += st_moved;
			a.st[(size_t)chain * 4 + 2] += st_births;
		}
	}
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
PriorDev npb_prior_dev(const npb_ctx *ctx, int m_aux) {
	const PriorHost &p = ctx->prior;
	PriorDev d;
	memset(&d, 0, sizeof(d));
	d.D = p.D;
	d.m_aux = m_aux;
	d.flags = p.flags;
	d.inv_sqrt_kappa = (float)(1.0 / sqrt(p.kappa));
	d.nu = (float)p.nu;
	d.v_mean = (float)p.D;
	d.c0_2 = (float)(-0.5 * (p.D * log2(2.0 * M_PI) + p.logdetA / log(2.0)));
	d.log2_alpha_m = (float)log2(p.alpha / (double)(m_aux > 0 ? m_aux : 1));
	for (int i = 0; i < p.D; ++i) d.mu0[i] = (float)p.mu0[i];
	d.CT2 = ctx->d_CT2;
	d.S = ctx->d_S;
	return d;
}

static SweepArgs make_args(npb_chains *ch, int n_sweeps) {
	SweepArgs a;
	a.X = ch->ds->X32;
	a.Xw = ch->ds->Xw;
	a.z = ch->z;
	a.theta = ch->theta;
	a.counts = ch->counts;
	a.st = ch->st;
	a.kocc = ch->kocc;
	a.overflow = ch->overflow;
	a.N = (int)ch->ds->N;
	a.C = (int)ch->C;
	a.Kmax = ch->Kmax;
	a.sweep0 = ch->sweep;
	a.n_sweeps = n_sweeps;
	a.seed = ch->seed;
	a.prior = npb_prior_dev(ch->ctx, ch->m_aux);
	return a;
}

npb_status npb_launch_whiten(npb_dataset *ds) {
	npb_ctx *ctx = ds->ctx;
	PriorDev pd = npb_prior_dev(ctx, 1);
	int threads = 256;
	int64_t blocks = (ds->N + threads - 1) / threads;
	k_whiten<<<(unsigned)blocks, threads, 0, ctx->stream>>>(ds->X32, ds->Xw, ds->N, pd);
	NPB_CUDA_OK(cudaGetLastError());
	ds->whitened_epoch = ctx->prior_epoch;
	return NPB_OK;
}

npb_status npb_launch_chains_init(npb_chains *ch) {
	npb_ctx *ctx = ch->ctx;
	SweepArgs a = make_args(ch, 0);
	const int warps = 4;
	int64_t blocks = (ch->C + warps - 1) / warps;
	size_t shmem = (size_t)warps * ch->Kmax * sizeof(int);
	k_chains_init<<<(unsigned)blocks, warps * 32, shmem, ctx->stream>>>(a, ch->K0);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

template <int D, int SPL>
static npb_status launch_reg(npb_chains *ch, const SweepArgs &a) {
	npb_ctx *ctx = ch->ctx;
	int64_t blocks = (ch->C + NPB_SWEEP_WARPS - 1) / NPB_SWEEP_WARPS;
	if (ch->m_aux == 3) k_alg8_sweep_reg<D, SPL, 3><<<(unsigned)blocks, NPB_SWEEP_WARPS * 32, 0, ctx->stream>>>(a);
	else if (ch->m_aux == 1) k_alg8_sweep_reg<D, SPL, 1><<<(unsigned)blocks, NPB_SWEEP_WARPS * 32, 0, ctx->stream>>>(a);
	else return npb_fail(ctx, NPB_E_UNSUPPORTED, "m_aux must be 1 or 3 for the register-resident sweep kernel");
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

template <int D>
static npb_status launch_reg_d(npb_chains *ch, const SweepArgs &a) {
	switch (ch->Kmax) {
	case 32: return launch_reg<D, 1>(ch, a);
	case 64: return launch_reg<D, 2>(ch, a);
	case 128: return launch_reg<D, 4>(ch, a);
	case 256: return launch_reg<D, 8>(ch, a);
	case 512: return launch_reg<D, 16>(ch, a);
	default: return npb_fail(ch->ctx, NPB_E_UNSUPPORTED, "Kmax must be 32, 64, 128, 256 or 512");
	}
}

npb_status npb_launch_alg8_sweep(npb_chains *ch, int n_sweeps) {
	SweepArgs a = make_args(ch, n_sweeps);
	npb_status s;
	switch (ch->D) {
	case 2: s = launch_reg_d<2>(ch, a); break;
	case 3: s = launch_reg_d<3>(ch, a); break;
	default: return npb_fail(ch->ctx, NPB_E_UNSUPPORTED, "D not supported by the Alg. 8 sweep kernels yet");
	}
	if (s == NPB_OK) ch->sweep += (uint32_t)n_sweeps;
	return s;
}
