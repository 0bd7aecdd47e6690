// npb_density.cu -- batched multivariate-normal log-densities and per-chain diagnostics (sm_100a).
//
// k_logdensity<T>: replaces multivariate_normal_distribution::logprobability
//   (src/statistics/multivariatenormal.cpp:106-136): log N(x | mu_k, Sigma_k) = -q/2 - log sqrt((2 pi)^D det Sigma),
//   q = |T_k (x - mu_k)|^2 with the triangular factor prepared once per theta (npb_linalg.cpp) instead of an LU
//   inverse + determinant per call.
// k_chain_metrics: replaces clustering_performance::{calculateContingencyMatrix, calculateSimilarity}
//   (src/clustering_performance.cpp:14-82, int64 instead of int: Q12) and the joint log-likelihood of
//   MCMC::considerMaxLikelihood (src/np_mcmc.cpp:187-203), one CTA per chain.
#include "npb_internal.h"
#include <cstdio>
#include <cstdlib>

// T = double: everything in double.  T = float: the quadratic form in float, but the difference x - mu is formed
// in double first (float coordinates would lose |x| * 6e-8, which is 1e-5 of a tight cluster's scale and breaks the
// 1e-5 relative bar on the log-density).
template <typename T>
__global__ void k_logdensity(const double *X, const int64_t *rows, int64_t n_rows, int D, const double *mu, const T *Tf,
		const T *cst, int K, double *out) {
	const int TRI = npb_tri(D);
	int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= n_rows * K) return;
	const int64_t r = idx / K;
	const int k = (int)(idx - r * K);
	const int64_t i = rows ? rows[r] : r;
	const double *x = X + i * D;
	const double *m = mu + (size_t)k * D;
	const T *t = Tf + (size_t)k * TRI;
	T q = 0;
	for (int a = 0; a < D; ++a) {
		T y = 0;
		for (int b = a; b < D; ++b) y += t[npb_tri_off(D, a, b)] * (T)(x[b] - m[b]);
		q += y * y;
	}
	out[idx] = (double)(cst[k] - (T)0.5 * q);
}

// out[k] = sum_r log N(X[rows[r]] | theta_k): one CTA per k, fp64, fixed-order tree reduction (deterministic)
__global__ void k_logdensity_sum(const double *X, const int64_t *rows, int64_t n_rows, int D, const double *mu,
		const double *Tf, const double *cst, double *out) {
	__shared__ double red[256];
	const int TRI = npb_tri(D);
	const int k = blockIdx.x;
	const double *m = mu + (size_t)k * D;
	const double *t = Tf + (size_t)k * TRI;
	double acc = 0.0;
	for (int64_t r = threadIdx.x; r < n_rows; r += blockDim.x) {
		const int64_t i = rows ? rows[r] : r;
		const double *x = X + i * D;
		double q = 0;
		for (int a = 0; a < D; ++a) {
			double y = 0;
			for (int b = a; b < D; ++b) y += t[npb_tri_off(D, a, b)] * (x[b] - m[b]);
			q += y * y;
		}
		acc += cst[k] - 0.5 * q;
	}
	red[threadIdx.x] = acc;
	__syncthreads();
	for (int o = blockDim.x / 2; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) out[k] = red[0];
}

// One CTA per chain.  Contingency n[a][b] (a = truth label, b = slot) in shared memory, then purity / RI / ARI,
// joint log-likelihood from the chain's own slot table (log2 domain, float factors) and the occupied count.
// DT > 0: the dimension as a compile-time constant (straight-line quadratic form); DT = 0: generic (runtime D).
template <int DT>
__global__ void k_chain_metrics(const npb_z_t *z, const float *X, const float *theta, const int *counts, const int32_t *truth,
		int N, int C, int Kmax, int Drt, int Ktrue, int stage_theta, double *purity, double *ri, double *ari, double *jll, int32_t *Kout) {
	const int D = DT > 0 ? DT : Drt;
	extern __shared__ unsigned char smem_raw[];
	double *red = (double *)smem_raw;                        // [blockDim.x]
	int *cont = (int *)(red + blockDim.x);                   // [Ktrue][Kmax]
	float *sth = (float *)(cont + (size_t)Ktrue * Kmax);     // [Kmax][PS] when stage_theta
	const int chain = blockIdx.x;
	const int PS = npb_ps(D), TRI = npb_tri(D);
	const float *th = theta + (size_t)chain * Kmax * PS;
	for (int t = threadIdx.x; t < Ktrue * Kmax; t += blockDim.x) cont[t] = 0;
	if (stage_theta && jll) {
		// the chain's slot table in shared memory: an item reads all PS parameters of ITS slot, a gather that ran at
		// L2 latency from global memory (322 ms for 8192 chains x 100k items at D = 16 before this)
		for (int t = threadIdx.x; t < Kmax * PS; t += blockDim.x) sth[t] = th[t];
		th = sth;
	}
	__syncthreads();
	double acc = 0.0;
	for (int i = threadIdx.x; i < N; i += blockDim.x) {
		const int s = (int)z[(size_t)i * C + chain];
		if (truth) atomicAdd(&cont[truth[i] * Kmax + s], 1);
		if (jll) {
			const float *p = th + (size_t)s * PS;
			const float *x = X + (size_t)i * D;
			float q = 0.0f;
			if (DT > 0) {
				float d[DT > 0 ? DT : 1];
#pragma unroll
				for (int b = 0; b < DT; ++b) d[b] = __ldg(x + b) - p[b];
#pragma unroll
				for (int a = 0; a < DT; ++a) {
					float y = 0.0f;
#pragma unroll
					for (int b = a; b < DT; ++b) y = fmaf(p[DT + npb_tri_off(DT, a, b)], d[b], y);
					q = fmaf(y, y, q);
				}
			} else {
				for (int a = 0; a < D; ++a) {
					float y = 0.0f;
					for (int b = a; b < D; ++b) y = fmaf(p[D + npb_tri_off(D, a, b)], x[b] - p[b], y);
					q = fmaf(y, y, q);
				}
			}
			acc += (double)(p[D + TRI] - q);
		}
	}
	red[threadIdx.x] = acc;
	__syncthreads();
	for (int o = blockDim.x / 2; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		if (jll) jll[chain] = red[0] * (double)NPB_LN2;
		if (Kout) {
			int k = 0;
			for (int s = 0; s < Kmax; ++s) k += counts[(size_t)chain * Kmax + s] > 0;
			Kout[chain] = k;
		}
		if (truth && (purity || ri || ari)) {
			// clustering_performance.cpp:38-82 in int64 / double
			long long a = 0, b = 0, c = 0, colmax = 0, n = 0;
			for (int s = 0; s < Kmax; ++s) {
				long long col = 0, mx = 0;
				for (int t = 0; t < Ktrue; ++t) {
					long long v = cont[t * Kmax + s];
					col += v;
					mx = v > mx ? v : mx;
					a += (v * v - v) / 2;
				}
				colmax += mx;
				c += (col * col - col) / 2;
				n += col;
			}
			for (int t = 0; t < Ktrue; ++t) {
				long long row = 0;
				for (int s = 0; s < Kmax; ++s) row += cont[t * Kmax + s];
				b += (row * row - row) / 2;
			}
			double S = ((double)n * (double)n - (double)n) / 2.0;
			double pu = n ? (double)colmax / (double)n : 0.0, r = 0.0, ar = 0.0;
			if (S != 0.0) {
				r = (2.0 * (double)a - (double)b - (double)c) / S + 1.0;
				double bc = (double)b * (double)c / S, bpc = ((double)b + (double)c) / 2.0;
				if (bc != bpc) ar = ((double)a - bc) / (bpc - bc);
			}
			if (purity) purity[chain] = pu;
			if (ri) ri[chain] = r;
			if (ari) ari[chain] = ar;
		}
	}
}

// ---- launchers used by npb_api.cu -------------------------------------------------------------------------
npb_status npb_launch_logdensity(npb_ctx *ctx, npb_dataset *ds, const int64_t *d_rows, int64_t n_rows, int K,
		const double *d_mu, const double *d_T, const double *d_c, const float *f_mu, const float *f_T, const float *f_c,
		int precision, double *d_out) {
	int64_t total = n_rows * K;
	int threads = 256;
	int64_t blocks = (total + threads - 1) / threads;
	if (precision == 64)
		k_logdensity<double><<<(unsigned)blocks, threads, 0, ctx->stream>>>(ds->X64, d_rows, n_rows, ds->D, d_mu, d_T, d_c, K, d_out);
	else
		k_logdensity<float><<<(unsigned)blocks, threads, 0, ctx->stream>>>(ds->X64, d_rows, n_rows, ds->D, d_mu, f_T, f_c, K, d_out);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

npb_status npb_launch_logdensity_sum(npb_ctx *ctx, npb_dataset *ds, const int64_t *d_rows, int64_t n_rows, int K,
		const double *d_mu, const double *d_T, const double *d_c, double *d_out) {
	k_logdensity_sum<<<K, 256, 0, ctx->stream>>>(ds->X64, d_rows, n_rows, ds->D, d_mu, d_T, d_c, d_out);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

npb_status npb_launch_metrics(npb_chains *ch, const int32_t *d_truth, int Ktrue, double *d_purity, double *d_ri,
		double *d_ari, double *d_jll, int32_t *d_K) {
	npb_ctx *ctx = ch->ctx;
	const int threads = 256;
	size_t shmem = (size_t)Ktrue * ch->Kmax * sizeof(int) + threads * sizeof(double);
	if (shmem > 200 * 1024) return npb_fail(ctx, NPB_E_UNSUPPORTED, "contingency table does not fit shared memory");
	const size_t table = (size_t)ch->Kmax * npb_ps(ch->D) * sizeof(float);
	const int stage = (shmem + table <= 96 * 1024) ? 1 : 0;
	if (stage) shmem += table;
#define NPB_METRICS_LAUNCH(DT)                                                                                              \
	do {                                                                                                               \
		NPB_CUDA_OK(cudaFuncSetAttribute(k_chain_metrics<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem)); \
		k_chain_metrics<DT><<<(unsigned)ch->C, threads, shmem, ctx->stream>>>(ch->z, ch->ds->X32, ch->theta, ch->counts, \
				d_truth, (int)ch->ds->N, (int)ch->C, ch->Kmax, ch->D, Ktrue, stage, d_purity, d_ri, d_ari, d_jll, d_K);      \
	} while (0)
	switch (ch->D) {
	case 2: NPB_METRICS_LAUNCH(2); break;
	case 3: NPB_METRICS_LAUNCH(3); break;
	case 4: NPB_METRICS_LAUNCH(4); break;
	case 8: NPB_METRICS_LAUNCH(8); break;
	case 16: NPB_METRICS_LAUNCH(16); break;
	default: NPB_METRICS_LAUNCH(0); break;
	}
#undef NPB_METRICS_LAUNCH
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

// ---- FP32 FFMA peak of the device (roofline denominator for the sweep kernels; SURVEY 8d asks the builder to
// measure it because MEASURED_PEAKS.json holds only HBM and bf16 tensor figures) -------------------------------
__global__ void __launch_bounds__(256) k_fma_peak(float *out, int iters, float a, float b) {
	float r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int u = 0; u < 16; ++u) {
			r0 = fmaf(r0, a, b); r1 = fmaf(r1, a, b); r2 = fmaf(r2, a, b); r3 = fmaf(r3, a, b);
			r4 = fmaf(r4, a, b); r5 = fmaf(r5, a, b); r6 = fmaf(r6, a, b); r7 = fmaf(r7, a, b);
		}
	}
	float s = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
	if (s == 123.456f) out[0] = s;
}

// the same probe with the packed form (sm_100a FFMA2: two FMAs per instruction on a 64-bit register pair)
__global__ void __launch_bounds__(256) k_fma2_peak(float *out, int iters, float a, float b) {
	unsigned long long r[8], aa, bb;
	asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
	asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
	for (int u = 0; u < 8; ++u) {
		const float lo = threadIdx.x + u, hi = lo + 0.5f;
		asm("mov.b64 %0, {%1, %2};" : "=l"(r[u]) : "f"(lo), "f"(hi));
	}
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int u = 0; u < 16; ++u) {
#pragma unroll
			for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[k]) : "l"(aa), "l"(bb));
		}
	}
	float s = 0.0f;
	for (int u = 0; u < 8; ++u) {
		float lo, hi;
		asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r[u]));
		s += lo + hi;
	}
	if (s == 123.456f) out[0] = s;
}

npb_status npb_launch_fma_peak(npb_ctx *ctx, double *tflops) {
	float *d = nullptr;
	NPB_CUDA_OK(cudaMalloc((void **)&d, sizeof(float)));
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
	const int blocks = sms * 8, threads = 256, iters = 4096;
	double best = 0.0;
	for (int rep = 0; rep < 12; ++rep) {
		cudaEventRecord(ctx->ev0, ctx->stream);
		k_fma_peak<<<blocks, threads, 0, ctx->stream>>>(d, iters, 0.999f, 0.001f);
		cudaEventRecord(ctx->ev1, ctx->stream);
		cudaError_t e = cudaStreamSynchronize(ctx->stream);
		if (e != cudaSuccess) { cudaFree(d); return npb_fail_cuda(ctx, e, "k_fma_peak", __FILE__, __LINE__); }
		float ms = 0;
		cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
		double fl = 2.0 * 8 * 16 * (double)iters * blocks * threads;
		double tf = fl / (ms * 1e-3) / 1e12;
		if (rep >= 2 && tf > best) best = tf;
	}
	// the roofline denominator is the better of the scalar and the packed instruction stream
	for (int rep = 0; rep < 12; ++rep) {
		cudaEventRecord(ctx->ev0, ctx->stream);
		k_fma2_peak<<<blocks, threads, 0, ctx->stream>>>(d, iters, 0.999f, 0.001f);
		cudaEventRecord(ctx->ev1, ctx->stream);
		cudaError_t e = cudaStreamSynchronize(ctx->stream);
		if (e != cudaSuccess) { cudaFree(d); return npb_fail_cuda(ctx, e, "k_fma2_peak", __FILE__, __LINE__); }
		float ms = 0;
		cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
		double fl = 2.0 * 2 * 8 * 16 * (double)iters * blocks * threads;
		double tf = fl / (ms * 1e-3) / 1e12;
		if (getenv("NPB_PEAK_VERBOSE") && rep == 11) fprintf(stderr, "npb_fp32_peak: FFMA %.2f TFLOP/s, FFMA2 %.2f TFLOP/s\n", best, tf);
		if (rep >= 2 && tf > best) best = tf;
	}
	cudaFree(d);
	*tflops = best;
	return NPB_OK;
}
