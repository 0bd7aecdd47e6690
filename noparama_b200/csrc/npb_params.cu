// npb_params.cu -- the parameter update of a sweep "done right" (SURVEY 8f-1), for all chains at once.
//
// The reference calls UpdateClusters::update after every sweep (np_mcmc.cpp:170; np_update_clusters.cpp:71-141: twenty
// Metropolis-Hastings proposals from the prior per cluster) but cluster_t::setSuffies slices the new parameters away
// (np_cluster.h:49-51), so cluster parameters never change after birth.  This file is the opt-in fix: every
// occupied cluster's (mu, Sigma) is refreshed from the conjugate normal-inverse-Wishart posterior of its members
//     kappa_n = kappa_0 + n,  nu_n = nu_0 + n,  mu_n = (kappa_0 mu_0 + n xbar) / kappa_n,
//     Lambda_n = Lambda_0 + sum (x - xbar)(x - xbar)^T + kappa_0 n / kappa_n (xbar - mu_0)(xbar - mu_0)^T,
// either by a draw  Sigma ~ IW(nu_n, Lambda_n), mu ~ N(mu_n, Sigma / kappa_n)  or by the posterior mean
// (mu_n, Lambda_n / (nu_n - D - 1)).  There is no reference arithmetic to match (parity unpinned by construction); the
// tests check the sufficient statistics and the posterior-mean mode against numpy in double and the draw's moments.
//
// k_member_lists  one warp per chain: the items of every slot, contiguous and in ascending order (deterministic).
// k_suffstats<D>   one warp per (chain, slot): lanes stride over the slot's members, each keeping sum (x - c) and the upper
//                  triangle of sum (x - c)(x - c)^T about the slot's current mean c in fp32 registers; lanes combined in
//                  fp64 -- no atomics.
// k_niw_update<D>  one thread per (chain, slot), fp64: UL-Cholesky  Lambda_n = G G^T  (G upper), L = G^-T (lower, so that
//                  Lambda_n^-1 = L L^T), Bartlett factor B (lower; B_ii^2 ~ chi^2_{nu_n - i}, B_ij ~ N(0,1)), M = L B, and
//                  Sigma^-1 = M M^T: the kernels' upper-triangular precision factor is M^T, no inverse is ever formed;
//                  mu = mu_n + M^-T z / sqrt(kappa_n) by back-substitution.
#include "npb_internal.h"
#include "npb_alg8_kernel.cuh"

enum { NPB_RNG_PARAMS = 6 };

// Member lists of every chain: perm[chain][off_k .. off_k + n_k) = the items of slot k in ascending item order, off_k the
// prefix sum of the member counts.  One warp per chain, deterministic (no atomics): a batch of 32 items is grouped by slot
// with match.any, the first lane of a group advances that slot's cursor (shared memory, one word per slot).
__global__ void __launch_bounds__(128) k_member_lists(const npb_z_t *z, const int *counts, int N, int C, int Kmax, int32_t *perm) {
	extern __shared__ int s_cursor[]; // [warps][Kmax]
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int chain = blockIdx.x * (blockDim.x >> 5) + warp;
	if (chain >= C) return;
	int *cur = s_cursor + warp * Kmax;
	{
		int run = 0; // exclusive prefix sum of the counts, 32 slots at a time
		for (int k0 = 0; k0 < Kmax; k0 += 32) {
			const int c = counts[(size_t)chain * Kmax + k0 + lane];
			int incl = c;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const int t = __shfl_up_sync(0xffffffffu, incl, o);
				if (lane >= o) incl += t;
			}
			cur[k0 + lane] = run + incl - c;
			run += __shfl_sync(0xffffffffu, incl, 31);
		}
	}
	__syncwarp();
	int32_t *out = perm + (size_t)chain * N;
	int znext = lane < N ? (int)z[(size_t)lane * C + chain] : -1;
	for (int i0 = 0; i0 < N; i0 += 32) {
		const int i = i0 + lane;
		const int zz = znext;
		const int inx = i + 32;
		znext = inx < N ? (int)z[(size_t)inx * C + chain] : -1; // next batch in flight while this one is placed
		const unsigned active = __ballot_sync(0xffffffffu, i < N);
		if (i < N) {
			const unsigned grp = __match_any_sync(active, zz);
			const int rank = __popc(grp & ((1u << lane) - 1u));
			const int base = cur[zz];
			out[base + rank] = i;
			__syncwarp(active);
			if (rank == 0) cur[zz] = base + __popc(grp);
		}
		__syncwarp();
	}
}

// Sufficient statistics of one (chain, slot) per warp from its member list: lanes stride over the members, each lane
// keeps sum (x - c) and the upper triangle of sum (x - c)(x - c)^T about the slot's current mean c in fp32 registers
// (D + D(D+1)/2 of them: the lane-private outer product costs D(D+1)/2 + 2D instructions per MEMBER AND LANE, i.e. 5 warp
// instructions per member at D = 16 against ~65 for a version that spread one member's entries over the lanes), then the
// lanes are combined in fp64.
template <int D>
__global__ void __launch_bounds__(128) k_suffstats(const float *X, const int32_t *perm, const int *counts, const float *theta, int N,
		int C, int Kmax, double *stats) {
	constexpr int NE = D + npb_tri(D);
	const int lane = threadIdx.x & 31;
	const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); // (chain, slot)
	if (idx >= C * Kmax) return;
	const int chain = idx / Kmax, slot = idx - chain * Kmax;
	const int n = counts[idx];
	double *o = stats + (size_t)idx * NE;
	if (n <= 0) return;
	int off = 0;
	for (int k = lane; k < slot; k += 32) off += counts[(size_t)chain * Kmax + k];
#pragma unroll
	for (int s = 16; s > 0; s >>= 1) off += __shfl_xor_sync(0xffffffffu, off, s);
	const int32_t *mem = perm + (size_t)chain * N + off;
	float cen[D], acc[NE];
	{
		const float *th = theta + (size_t)idx * npb_ps(D);
#pragma unroll
		for (int r = 0; r < D; ++r) cen[r] = th[r];
	}
#pragma unroll
	for (int e = 0; e < NE; ++e) acc[e] = 0.0f;
	int item_next = lane < n ? mem[lane] : 0;
	for (int t = lane; t < n; t += 32) {
		const float *row = X + (size_t)item_next * D;
		if (t + 32 < n) item_next = mem[t + 32]; // the next member's index is in flight while this row is used
		float d[D];
		if constexpr (D % 4 == 0) {
#pragma unroll
			for (int r = 0; r < D; r += 4) {
				const float4 v = __ldg(reinterpret_cast<const float4 *>(row + r));
				d[r] = v.x - cen[r]; d[r + 1] = v.y - cen[r + 1]; d[r + 2] = v.z - cen[r + 2]; d[r + 3] = v.w - cen[r + 3];
			}
		} else {
#pragma unroll
			for (int r = 0; r < D; ++r) d[r] = __ldg(row + r) - cen[r];
		}
		int e = D;
#pragma unroll
		for (int r = 0; r < D; ++r) {
			acc[r] += d[r];
#pragma unroll
			for (int c = r; c < D; ++c, ++e) acc[e] = fmaf(d[r], d[c], acc[e]);
		}
	}
#pragma unroll
	for (int e = 0; e < NE; ++e) {
		double v = (double)acc[e];
#pragma unroll
		for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
		if (lane == 0) o[e] = v;
	}
}

struct ParamsPrior {
	double mu0[NPB_MAX_D];
	double kappa0, nu0;
	const double *Lambda0; // device [D, D]
};

__device__ inline double rng_u01(const Philox &ph, uint32_t c0, uint32_t &ctr, uint32_t c2, uint32_t c3, uint32_t (&buf)[4], int &have) {
	if (have == 0) { ph(c0, ctr++, c2, c3, buf); have = 4; }
	const uint32_t a = buf[4 - have];
	have--;
	return ((double)a + 0.5) * 2.3283064365386963e-10;
}
__device__ inline double rng_normal(const Philox &ph, uint32_t c0, uint32_t &ctr, uint32_t c2, uint32_t c3, uint32_t (&buf)[4], int &have) {
	const double u1 = rng_u01(ph, c0, ctr, c2, c3, buf, have), u2 = rng_u01(ph, c0, ctr, c2, c3, buf, have);
	return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
// chi^2_k = 2 Gamma(k / 2), Marsaglia & Tsang (2000); k >= 1 (k / 2 < 1 is boosted with U^(1/a))
__device__ inline double rng_chi2(double k, const Philox &ph, uint32_t c0, uint32_t &ctr, uint32_t c2, uint32_t c3, uint32_t (&buf)[4], int &have) {
	double a = 0.5 * k, boost = 1.0;
	if (a < 1.0) {
		boost = pow(rng_u01(ph, c0, ctr, c2, c3, buf, have), 1.0 / a);
		a += 1.0;
	}
	const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
	for (int it = 0; it < 64; ++it) {
		const double x = rng_normal(ph, c0, ctr, c2, c3, buf, have);
		double v = 1.0 + c * x;
		if (v <= 0.0) continue;
		v = v * v * v;
		const double u = rng_u01(ph, c0, ctr, c2, c3, buf, have);
		if (u < 1.0 - 0.0331 * x * x * x * x || log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return 2.0 * d * v * boost;
	}
	return 2.0 * d * boost;
}

template <int D>
__global__ void k_niw_update(const double *stats, const int *counts, float *theta, int C, int Kmax, ParamsPrior pr, int mode,
		uint64_t seed, uint32_t epoch, int *fail) {
	constexpr int NE = D + npb_tri(D), TRI = npb_tri(D), PS = npb_ps(D);
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= C * Kmax) return;
	const int chain = idx / Kmax, slot = idx - chain * Kmax;
	const int n = counts[idx];
	if (n <= 0) return;
	const double *st = stats + (size_t)idx * NE;
	double xbar[D], A[D][D];
	const double dn = (double)n;
	{
		// the statistics were taken about the slot's current mean (k_suffstats): st[r] = sum (x_r - c_r),
		// st[(r,c)] = sum (x_r - c_r)(x_c - c_c)
		const float *cen = theta + (size_t)idx * PS;
		double off[D];
		for (int r = 0; r < D; ++r) {
			off[r] = st[r] / dn;
			xbar[r] = (double)cen[r] + off[r];
		}
		int e = D;
		for (int r = 0; r < D; ++r)
			for (int c = r; c < D; ++c, ++e) {
				const double s = st[e] - dn * off[r] * off[c]; // scatter about the members' mean
				A[r][c] = s;
				A[c][r] = s;
			}
	}
	const double kn = pr.kappa0 + dn, nun = pr.nu0 + dn;
	double mun[D];
	for (int r = 0; r < D; ++r) mun[r] = (pr.kappa0 * pr.mu0[r] + dn * xbar[r]) / kn;
	const double w = pr.kappa0 * dn / kn;
	for (int r = 0; r < D; ++r)
		for (int c = 0; c < D; ++c) A[r][c] += pr.Lambda0[r * D + c] + w * (xbar[r] - pr.mu0[r]) * (xbar[c] - pr.mu0[c]);
	if (mode == 2) { // posterior mean of Sigma
		const double sc = 1.0 / fmax(nun - D - 1.0, 1e-300);
		for (int r = 0; r < D; ++r)
			for (int c = 0; c < D; ++c) A[r][c] *= sc;
	}
	// UL-Cholesky: A = G G^T, G upper triangular (in place in the upper triangle of A)
	bool ok = true;
	for (int j = D - 1; j >= 0; --j) {
		double s = A[j][j];
		for (int k = j + 1; k < D; ++k) s -= A[j][k] * A[j][k];
		if (!(s > 0.0)) { ok = false; s = 1e-300; }
		const double g = sqrt(s);
		A[j][j] = g;
		for (int i = 0; i < j; ++i) {
			double t = A[i][j];
			for (int k = j + 1; k < D; ++k) t -= A[i][k] * A[j][k];
			A[i][j] = t / g;
		}
	}
	if (!ok) { atomicExch(fail, 1); return; }
	// Ginv = G^-1 (upper), stored in the strict lower + diagonal transposed: Lm[r][c] = Ginv[c][r]  (L = G^-T, lower)
	double Lm[D][D];
	for (int c = 0; c < D; ++c) {
		for (int r = 0; r < D; ++r) Lm[r][c] = 0.0;
	}
	for (int c = 0; c < D; ++c) { // column c of Ginv: solve G y = e_c by back-substitution
		for (int r = c; r >= 0; --r) {
			double s = (r == c) ? 1.0 : 0.0;
			for (int k = r + 1; k <= c; ++k) s -= A[r][k] * Lm[c][k]; // Lm[c][k] = Ginv[k][c]
			Lm[c][r] = s / A[r][r];
		}
	}
	// M (lower) with Sigma^-1 = M M^T
	Philox ph((uint32_t)seed, (uint32_t)(seed >> 32) + (uint32_t)chain);
	uint32_t ctr = 0, buf[4];
	int have = 0;
	const uint32_t c0 = (uint32_t)slot, c2 = epoch, c3 = NPB_RNG_PARAMS;
	double (&M)[D][D] = A; // the Cholesky factor is no longer needed
	if (mode == 2) {
		for (int r = 0; r < D; ++r)
			for (int c = 0; c < D; ++c) M[r][c] = c <= r ? Lm[r][c] : 0.0;
	} else {
		for (int c = 0; c < D; ++c) { // column c of the Bartlett factor, then column c of M = L B
			double bcol[D];
			bcol[c] = sqrt(rng_chi2(nun - c, ph, c0, ctr, c2, c3, buf, have));
			for (int k = c + 1; k < D; ++k) bcol[k] = rng_normal(ph, c0, ctr, c2, c3, buf, have);
			for (int r = 0; r < D; ++r) {
				double s = 0.0;
				for (int k = c; k <= r; ++k) s += Lm[r][k] * bcol[k];
				M[r][c] = s;
			}
		}
	}
	// mu = mu_n + M^-T zeta / sqrt(kappa_n): solve M^T y = zeta (M^T upper)
	double mu[D];
	if (mode == 2) {
		for (int r = 0; r < D; ++r) mu[r] = mun[r];
	} else {
		double zeta[D], y[D];
		for (int r = 0; r < D; ++r) zeta[r] = rng_normal(ph, c0, ctr, c2, c3, buf, have);
		for (int r = D - 1; r >= 0; --r) {
			double s = zeta[r];
			for (int k = r + 1; k < D; ++k) s -= M[k][r] * y[k];
			y[r] = s / M[r][r];
		}
		const double isk = 1.0 / sqrt(kn);
		for (int r = 0; r < D; ++r) mu[r] = mun[r] + y[r] * isk;
	}
	// theta in the kernels' form: mu, T2 = sqrt(log2e / 2) M^T (upper, packed row-wise), c2 = log2 normaliser
	float *o = theta + (size_t)idx * PS;
	for (int r = 0; r < D; ++r) o[r] = (float)mu[r];
	double lg = 0.0;
	for (int r = 0; r < D; ++r) {
		lg += log2(M[r][r]);
		for (int c = r; c < D; ++c) o[D + npb_tri_off(D, r, c)] = (float)(M[c][r] * NPB_HALF_LOG2E_SQRT);
	}
	o[D + TRI] = (float)(-0.5 * D * log2(2.0 * 3.141592653589793) + lg);
}

template <int D>
static npb_status update_params_d(npb_chains *ch, int mode, const ParamsPrior &pr) {
	npb_ctx *ctx = ch->ctx;
	constexpr int NE = D + npb_tri(D);
	const int C = (int)ch->C, Kmax = ch->Kmax, N = (int)ch->ds->N;
	if (!ch->pstats) NPB_CUDA_OK(cudaMalloc((void **)&ch->pstats, (size_t)C * Kmax * NE * sizeof(double)));
	if (!ch->pfail) NPB_CUDA_OK(cudaMalloc((void **)&ch->pfail, sizeof(int)));
	NPB_CUDA_OK(cudaMemsetAsync(ch->pfail, 0, sizeof(int), ctx->stream));
	if (!ch->sm_pool) NPB_CUDA_OK(cudaMalloc((void **)&ch->sm_pool, (size_t)N * C * sizeof(int32_t))); // shared with split-merge
	{
		const int warps = 4;
		k_member_lists<<<(C + warps - 1) / warps, warps * 32, (size_t)warps * Kmax * sizeof(int), ctx->stream>>>(ch->z, ch->counts, N, C,
				Kmax, ch->sm_pool);
		NPB_CUDA_OK(cudaGetLastError());
		const int total_w = C * Kmax;
		k_suffstats<D><<<(total_w + warps - 1) / warps, warps * 32, 0, ctx->stream>>>(ch->ds->X32, ch->sm_pool, ch->counts, ch->theta, N,
				C, Kmax, ch->pstats);
		NPB_CUDA_OK(cudaGetLastError());
	}
	const int total = C * Kmax;
	k_niw_update<D><<<(total + 63) / 64, 64, 0, ctx->stream>>>(ch->pstats, ch->counts, ch->theta, C, Kmax, pr, mode, ch->seed,
			ch->param_epoch++, ch->pfail);
	NPB_CUDA_OK(cudaGetLastError());
	int fail = 0;
	NPB_CUDA_OK(cudaMemcpyAsync(&fail, ch->pfail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	if (fail) return npb_fail(ctx, NPB_E_NOT_POSITIVE, "posterior scale matrix of a cluster is not positive definite");
	return NPB_OK;
}

npb_status npb_launch_update_params(npb_chains *ch, int mode, const double *mu0, double kappa0, double nu0, const double *Lambda0) {
	npb_ctx *ctx = ch->ctx;
	const int D = ch->D;
	ParamsPrior pr;
	for (int i = 0; i < D; ++i) pr.mu0[i] = mu0[i];
	pr.kappa0 = kappa0;
	pr.nu0 = nu0;
	if (!ch->pLambda0) NPB_CUDA_OK(cudaMalloc((void **)&ch->pLambda0, sizeof(double) * D * D));
	NPB_CUDA_OK(cudaMemcpyAsync(ch->pLambda0, Lambda0, sizeof(double) * D * D, cudaMemcpyHostToDevice, ctx->stream));
	pr.Lambda0 = ch->pLambda0;
	switch (D) {
	case 2: return update_params_d<2>(ch, mode, pr);
	case 3: return update_params_d<3>(ch, mode, pr);
	case 4: return update_params_d<4>(ch, mode, pr);
	case 8: return update_params_d<8>(ch, mode, pr);
	case 16: return update_params_d<16>(ch, mode, pr);
	default: return npb_fail(ctx, NPB_E_UNSUPPORTED, "parameter update covers D = 2, 3, 4, 8, 16");
	}
}
