// npb_scalarnoise.cu -- the other two likelihood families of the reference (`-c regression`, `-c angular`, np_main.cpp:196-205,
// :322-328) behind the same sampler seam: Algorithm 8 with the scalar-noise normal likelihood of
// scalarnoise_multivariatenormal.cpp:77-250 and the normal-inverse-gamma base measure of normalinvgamma.h:58-84 (np_main.cpp:357-364).
//
// A cluster's theta = (mu [2], sigma).  Either family's log-density is a one-dimensional normal of a residual that is AFFINE in
// the item's row:  regression (row (x0, x1, y)): r = y - mu . (x0, x1);  angular (row (a, b)): r = d + sin(t) a - cos(t) b with
// (d, t) = (|trunc mu_0|, fmod(|trunc mu_1|, 2 pi)) -- prepare() (scalarnoise_multivariatenormal.cpp:31-47) reaches ::abs(int), the
// truncation is the reference's (quirk Q12, pinned by tests/test_oracle_scalarnoise.py against its compiled sources).  So a slot
// is stored in the SAME layout as a multivariate-normal slot of the row's dimension D -- [mu (D) | T upper packed | c2] with
// only row 0 of T set: log2 p = c2 - (T_0 . (x - mu))^2 -- and everything downstream of the sweep that evaluates a slot
// (k_chain_metrics' joint log-likelihood, npb_logdensity_batch's kernels) works on it unchanged.
//
// k_sn_sweep: one warp per chain, the chain's slots over the lanes (Kmax / 32 per lane, derived coefficients in shared memory),
// the m auxiliary draws on lanes 0 .. m-1, one exponential race per step (arg max of log2 w - log2 E, E ~ Exp(1): the
// categorical pick of dim1algebra.hpp:2078-2104 in law), sequential over the items of the shared scan order exactly like the
// multivariate-normal kernels.  The path is integer/latency work of a few flops per candidate: no tensor-core shape here.
#include "npb_internal.h"
#include <cmath>
#include <cstring>

struct SnArgs {
	const float *X;            // [N, D] rows as read_data builds them
	const int32_t *scan_order; // [n_sweeps, N]
	npb_z_t *z;                // [N, C]
	float *theta;              // [C, Kmax, PS]
	int *counts;               // [C, Kmax]
	unsigned long long *st;
	int *kocc, *overflow;
	int N, C, Kmax, D, M, family;
	uint32_t sweep0;
	int n_sweeps;
	uint64_t seed;
	float mu0[2], A[4];        // mu = mu0 + sigma A z, A A^T = Lambda^-1 (lower triangular)
	float ig_alpha, ig_beta;   // 1 / sigma^2 ~ Gamma(shape alpha, scale beta) (gamma.h:41)
	float log2_alpha_m;        // log2(alpha / m)
};

__device__ __forceinline__ float sn_u_open(uint32_t r) { // uniform in (0, 1), never an end point
	return (__uint2float_rz(r >> 8) + 0.5f) * 5.9604644775390625e-8f;
}

// Gamma(shape a, scale 1) by Marsaglia-Tsang (a >= 1; a < 1 through the a + 1 boost), counter-based: iteration it of draw
// `draw` takes the Philox block (c0, 16 + it, c2, c3)
__device__ float sn_gamma(const Philox &ph, uint32_t c0, uint32_t c2, uint32_t c3, float a) {
	const float a1 = a < 1.0f ? a + 1.0f : a;
	const float d = a1 - 1.0f / 3.0f, c = rsqrtf(9.0f * d);
	float g = d;
	for (uint32_t it = 0; it < 64u; ++it) {
		uint32_t w[4];
		ph(c0, 16u + it, c2, c3, w);
		float n0, n1;
		npb_normal2(w[0], w[1], n0, n1);
		const float t = 1.0f + c * n0;
		if (t <= 0.0f) continue;
		const float v = t * t * t, u = sn_u_open(w[2]);
		if (__logf(u) < 0.5f * n0 * n0 + d - d * v + d * __logf(v)) {
			g = d * v;
			if (a < 1.0f) g *= __powf(sn_u_open(w[3]), 1.0f / a);
			break;
		}
	}
	return g;
}

// one draw from the normal-inverse-gamma base measure (normalinvgamma.h:58-84): raw (mu_0, mu_1, sigma)
__device__ void sn_draw(const SnArgs &a, const Philox &ph, uint32_t c0, uint32_t c2, uint32_t c3, float &m0, float &m1, float &sigma) {
	const float val = a.ig_beta * sn_gamma(ph, c0, c2, c3, a.ig_alpha);
	sigma = rsqrtf(fmaxf(val, 1e-30f));
	uint32_t w[4];
	ph(c0, 8u, c2, c3, w);
	float z0, z1;
	npb_normal2(w[0], w[1], z0, z1);
	m0 = a.mu0[0] + sigma * (a.A[0] * z0);
	m1 = a.mu0[1] + sigma * (a.A[2] * z0 + a.A[3] * z1);
}

// raw theta -> the slot layout [mu (D) | T packed upper | c2] (log2 units)
__device__ void sn_write_slot(int family, int D, float m0, float m1, float sigma, float *o) {
	const int PS = npb_ps(D), TRI = npb_tri(D);
	for (int t = 0; t < PS; ++t) o[t] = 0.0f;
	const float g = (float)NPB_HALF_LOG2E_SQRT / sigma;
	if (family == NPB_FAMILY_REGRESSION) { // rows (x_0 .. x_{D-2}, y): residual y - mu . x
		o[D + 0] = -g * m0;
		o[D + 1] = -g * m1;
		o[D + 2] = g;
	} else { // angular (D = 2): residual d + sin(t) a - cos(t) b, (d, t) canonical through abs(int) (Q12)
		const float d = fabsf(truncf(m0));
		const float th = fmodf(fabsf(truncf(m1)), 6.283185307179586f);
		float s, c;
		sincosf(th, &s, &c);
		o[0] = -d * s;
		o[1] = d * c;
		o[D + 0] = g * s;
		o[D + 1] = -g * c;
	}
	o[D + TRI] = -0.5f * log2f(6.283185307179586f * sigma * sigma);
}

// slot -> the four coefficients of the residual (r' = a0 x0 + a1 x1 + a2 x2 + a3, already scaled to log2 units) and c2
__device__ __forceinline__ void sn_coef(const float *p, int D, float4 &cf, float &c2) {
	const float t0 = p[D], t1 = p[D + 1], t2 = D > 2 ? p[D + 2] : 0.0f;
	cf = make_float4(t0, t1, t2, -(t0 * p[0] + t1 * p[1] + (D > 2 ? t2 * p[2] : 0.0f)));
	c2 = p[D + npb_tri(D)];
}

__global__ void k_sn_init(SnArgs a, int K0, uint32_t epoch) {
	extern __shared__ int sn_counts[]; // [warps][Kmax]
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int chain = blockIdx.x * (blockDim.x >> 5) + warp;
	if (chain >= a.C) return;
	const int PS = npb_ps(a.D);
	int *cnt = sn_counts + warp * a.Kmax;
	for (int k = lane; k < a.Kmax; k += 32) cnt[k] = 0;
	__syncwarp();
	Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
	float *th = a.theta + (size_t)chain * a.Kmax * PS;
	for (int k = lane; k < a.Kmax; k += 32) {
		float *o = th + (size_t)k * PS;
		if (k < K0) { // np_init_clusters.cpp:24-41: K0 draws from the base measure
			float m0, m1, sg;
			sn_draw(a, ph, (uint32_t)k, epoch, NPB_RNG_INIT_THETA, m0, m1, sg);
			sn_write_slot(a.family, a.D, m0, m1, sg, o);
		} else {
			for (int t = 0; t < PS; ++t) o[t] = 0.0f;
		}
	}
	for (int i = lane; i < a.N; i += 32) { // np_mcmc.cpp:69-85: every item to a uniformly chosen cluster
		uint32_t w[4];
		ph((uint32_t)i, 0u, epoch, NPB_RNG_INIT_Z, w);
		const int k = (int)__umulhi(w[0], (uint32_t)K0);
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

constexpr int SN_WARPS = 4;

// NealAlgorithm8::update (np_neal_algorithm8.cpp:49-167) for every item of every sweep, one warp per chain
__global__ void __launch_bounds__(SN_WARPS * 32) k_sn_sweep(SnArgs a) {
	extern __shared__ float4 sn_smem[]; // per warp: [Kmax] float4 coefficients, [Kmax] c2, [Kmax] counts
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int chain = blockIdx.x * SN_WARPS + warp;
	if (chain >= a.C) return;
	const int Kmax = a.Kmax, D = a.D, PS = npb_ps(D), SPL = Kmax >> 5;
	float4 *cf = sn_smem + (size_t)warp * Kmax * 2;
	float *c2 = reinterpret_cast<float *>(cf + Kmax);
	int *cnt = reinterpret_cast<int *>(c2 + Kmax);
	float *th = a.theta + (size_t)chain * Kmax * PS;
	int kocc = 0;
	for (int k = lane; k < Kmax; k += 32) {
		sn_coef(th + (size_t)k * PS, D, cf[k], c2[k]);
		cnt[k] = a.counts[(size_t)chain * Kmax + k];
		kocc += cnt[k] > 0;
	}
	kocc = __reduce_add_sync(0xffffffffu, kocc);
	__syncwarp();
	const Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
	unsigned long long st_cand = 0, st_moved = 0, st_births = 0;
	int overflow = 0;
	for (int sw = 0; sw < a.n_sweeps; ++sw) {
		const uint32_t sweep = a.sweep0 + (uint32_t)sw;
		const int32_t *order = a.scan_order + (size_t)sw * a.N;
		for (int j = 0; j < a.N; ++j) {
			const int item = __ldg(order + j);
			const float *x = a.X + (size_t)item * D;
			const float x0 = __ldg(x), x1 = __ldg(x + 1), x2 = D > 2 ? __ldg(x + 2) : 0.0f;
			const int zold = (int)a.z[(size_t)item * a.C + chain];
			const int n_old = cnt[zold]; // (:63 retract: a cluster left empty is removed)
			const int K_i = kocc - (n_old == 1);
			float best = -INFINITY;
			int bidx = 0;
			// existing clusters: w_k = p(x | theta_k) n_k (np_neal_algorithm8.cpp:100-109)
			for (int s = 0; s < SPL; s += 4) {
				uint32_t w[4];
				ph((uint32_t)j, (uint32_t)(lane * 64 + (s >> 2)), sweep, NPB_RNG_PICK, w);
#pragma unroll
				for (int q = 0; q < 4; ++q) {
					const int k = lane + 32 * (s + q);
					if (s + q < SPL) {
						const int n = cnt[k] - (k == zold);
						if (n > 0) {
							const float4 c = cf[k];
							const float r = fmaf(c.x, x0, fmaf(c.y, x1, fmaf(c.z, x2, c.w)));
							const float key = __log2f((float)n) + c2[k] - r * r - __log2f(-__logf(sn_u_open(w[q])));
							if (key > best) { best = key; bidx = k; }
						}
					}
				}
			}
			// the m auxiliary draws from the base measure: w = p(x | theta') alpha / m (:79-84, :111-119)
			float am0 = 0.0f, am1 = 0.0f, asg = 1.0f;
			if (lane < a.M) {
				sn_draw(a, ph, (uint32_t)j, sweep, NPB_RNG_AUX + 16u * (uint32_t)(lane + 1), am0, am1, asg);
				float slot[16];
				sn_write_slot(a.family, D, am0, am1, asg, slot);
				float4 c;
				float cc;
				sn_coef(slot, D, c, cc);
				const float r = fmaf(c.x, x0, fmaf(c.y, x1, fmaf(c.z, x2, c.w)));
				uint32_t w[4];
				ph((uint32_t)j, 4096u + (uint32_t)lane, sweep, NPB_RNG_PICK, w);
				const float key = a.log2_alpha_m + cc - r * r - __log2f(-__logf(sn_u_open(w[0])));
				if (key > best) { best = key; bidx = Kmax + lane; }
			}
			// arg max over the warp (ties: the lower index, deterministic)
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) {
				const float ob = __shfl_xor_sync(0xffffffffu, best, o);
				const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
				if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
			}
			st_cand += (unsigned long long)(K_i + a.M);
			int znew = bidx;
			if (bidx >= Kmax) { // a new cluster (:139-160): the lowest free slot takes the winning draw's parameters
				const int src = bidx - Kmax;
				const float m0 = __shfl_sync(0xffffffffu, am0, src), m1 = __shfl_sync(0xffffffffu, am1, src), sg = __shfl_sync(0xffffffffu, asg, src);
				int fs = -1;
				for (int s = 0; s < SPL && fs < 0; ++s) {
					const int k = lane + 32 * s;
					const unsigned m = __ballot_sync(0xffffffffu, cnt[k] - (k == zold) == 0);
					if (m) fs = 32 * s + (__ffs(m) - 1);
				}
				if (fs < 0) { // no room: the item stays where it was, the chain is reported (NPB_E_KMAX_OVERFLOW)
					overflow = 1;
					znew = zold;
				} else {
					znew = fs;
					if (lane == 0) {
						float *o = th + (size_t)fs * PS;
						sn_write_slot(a.family, D, m0, m1, sg, o);
						sn_coef(o, D, cf[fs], c2[fs]);
					}
					st_births++;
				}
			}
			__syncwarp();
			const bool born = bidx >= Kmax && !(znew == zold && n_old != 1); // (not born: the overflow case above)
			if (znew != zold) {
				if (lane == 0) {
					cnt[zold] = n_old - 1;
					cnt[znew] += 1;
					a.z[(size_t)item * a.C + chain] = (npb_z_t)znew;
				}
				kocc += (born ? 1 : 0) - (n_old == 1);
			}
			// (a singleton that picks a new cluster gets its own slot back with the new parameters: a move all the same)
			if (znew != zold || born) st_moved++;
			__syncwarp();
		}
	}
	for (int k = lane; k < Kmax; k += 32) a.counts[(size_t)chain * Kmax + k] = cnt[k];
	if (lane == 0) {
		a.kocc[chain] = kocc;
		if (overflow) a.overflow[chain] = 1;
		a.st[(size_t)chain * 4 + 0] += st_cand;
		a.st[(size_t)chain * 4 + 1] += st_moved;
		a.st[(size_t)chain * 4 + 2] += st_births;
	}
}

// parity probe of the base measure: count raw draws (mu_0, mu_1, sigma) of chain `chain`'s auxiliary stream
__global__ void k_sn_sample_base(SnArgs a, int chain, int count, float *out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
	float m0, m1, sg;
	sn_draw(a, ph, (uint32_t)i, 0x5eedu, NPB_RNG_AUX, m0, m1, sg);
	out[3 * i] = m0;
	out[3 * i + 1] = m1;
	out[3 * i + 2] = sg;
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static SnArgs sn_args(npb_chains *ch, int n_sweeps) {
	const PriorHost &p = ch->ctx->prior;
	SnArgs a;
	memset(&a, 0, sizeof(a));
	a.X = ch->ds->X32;
	a.scan_order = ch->scan_order;
	a.z = ch->z;
	a.theta = ch->theta;
	a.counts = ch->counts;
	a.st = ch->st;
	a.kocc = ch->kocc;
	a.overflow = ch->overflow;
	a.N = (int)ch->ds->N;
	a.C = (int)ch->C;
	a.Kmax = ch->Kmax;
	a.D = ch->D;
	a.M = ch->m_aux;
	a.family = p.family;
	a.sweep0 = ch->sweep;
	a.n_sweeps = n_sweeps;
	a.seed = ch->seed;
	// A = lower Cholesky factor of Lambda^-1 (2 x 2)
	const double l00 = p.Lambda[0], l01 = p.Lambda[1], l11 = p.Lambda[3];
	const double det = l00 * l11 - l01 * p.Lambda[2];
	const double i00 = l11 / det, i01 = -l01 / det, i11 = l00 / det;
	const double a00 = sqrt(i00), a10 = i01 / a00, a11 = sqrt(i11 - a10 * a10);
	a.mu0[0] = (float)p.mu0[0];
	a.mu0[1] = (float)p.mu0[1];
	a.A[0] = (float)a00; a.A[1] = 0.0f; a.A[2] = (float)a10; a.A[3] = (float)a11;
	a.ig_alpha = (float)p.ig_alpha;
	a.ig_beta = (float)p.ig_beta;
	a.log2_alpha_m = (float)log2(p.alpha / (double)(ch->m_aux > 0 ? ch->m_aux : 1));
	return a;
}

static npb_status sn_check(npb_chains *ch) {
	npb_ctx *ctx = ch->ctx;
	if (ch->Kmax % 32 || ch->Kmax > 1024) return npb_fail(ctx, NPB_E_UNSUPPORTED, "scalar-noise families: Kmax must be a multiple of 32, at most 1024");
	if (ch->m_aux < 1 || ch->m_aux > 8) return npb_fail(ctx, NPB_E_UNSUPPORTED, "scalar-noise families: 1 <= m_aux <= 8");
	if (ch->D != (ctx->prior.family == NPB_FAMILY_REGRESSION ? 3 : 2))
		return npb_fail(ctx, NPB_E_BAD_ARG, "scalar-noise families: rows are (1, a, b) for regression and (a, b) for angular (np_main.cpp:83-101)");
	return NPB_OK;
}

npb_status npb_launch_sn_init(npb_chains *ch, int K0) {
	npb_ctx *ctx = ch->ctx;
	npb_status s = sn_check(ch);
	if (s != NPB_OK) return s;
	if (K0 < 1 || K0 > ch->Kmax) return npb_fail(ctx, NPB_E_BAD_ARG, "K0 must be in [1, Kmax]");
	SnArgs a = sn_args(ch, 0);
	const int warps = 4;
	k_sn_init<<<(unsigned)((ch->C + warps - 1) / warps), warps * 32, (size_t)warps * ch->Kmax * sizeof(int), ctx->stream>>>(a, K0, ch->init_epoch++);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

npb_status npb_launch_sn_sweep(npb_chains *ch, int n_sweeps) {
	npb_ctx *ctx = ch->ctx;
	npb_status s = sn_check(ch);
	if (s != NPB_OK) return s;
	s = npb_launch_scan_order(ch, n_sweeps);
	if (s != NPB_OK) return s;
	SnArgs a = sn_args(ch, n_sweeps);
	const size_t shmem = (size_t)SN_WARPS * ch->Kmax * (sizeof(float4) + sizeof(float) + sizeof(int) + 8);
	NPB_CUDA_OK(cudaFuncSetAttribute(k_sn_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
	k_sn_sweep<<<(unsigned)((ch->C + SN_WARPS - 1) / SN_WARPS), SN_WARPS * 32, shmem, ctx->stream>>>(a);
	NPB_CUDA_OK(cudaGetLastError());
	ch->sweep += (uint32_t)n_sweeps;
	return NPB_OK;
}

npb_status npb_launch_sn_sample_base(npb_chains *ch, int chain, int count, float *d_out) {
	npb_ctx *ctx = ch->ctx;
	SnArgs a = sn_args(ch, 0);
	k_sn_sample_base<<<(count + 127) / 128, 128, 0, ctx->stream>>>(a, chain, count, d_out);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

// host: raw (mu [2], sigma) <-> the slot layout in double (npb_logdensity paths, get_params)
void npb_sn_slot_from_raw(int family, int D, const double *mu, double sigma, double *mu_slot /*[D]*/, double *T_packed /*[tri]*/, double *cst) {
	const int TRI = npb_tri(D);
	for (int t = 0; t < D; ++t) mu_slot[t] = 0.0;
	for (int t = 0; t < TRI; ++t) T_packed[t] = 0.0;
	const double g = 1.0 / sigma; // natural-log units here: log p = cst - 0.5 |T (x - mu)|^2
	if (family == NPB_FAMILY_REGRESSION) {
		T_packed[0] = -g * mu[0];
		T_packed[1] = -g * mu[1];
		T_packed[2] = g;
	} else {
		const double d = fabs((double)(int)mu[0]);
		const double th = fmod(fabs((double)(int)mu[1]), 2.0 * M_PI);
		mu_slot[0] = -d * sin(th);
		mu_slot[1] = d * cos(th);
		T_packed[0] = g * sin(th);
		T_packed[1] = -g * cos(th);
	}
	*cst = -0.5 * log(2.0 * M_PI * sigma * sigma);
}
