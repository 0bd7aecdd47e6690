// npb_common.cuh -- shared device helpers: Philox4x32-10, keyed scan-order permutation, normals,
// packed-triangular indexing, internal structs.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NPB_MAX_D 64
#define NPB_MAX_AUX 8
#define NPB_LOG2E 1.4426950408889634f
#define NPB_LN2 0.6931471805599453f
// T2 = T * sqrt(0.5*log2(e)) so that |T2 d|^2 is the quadratic form's contribution in log2 units
#define NPB_HALF_LOG2E_SQRT 0.8493218002880191

#define NPB_CUDA_OK(expr)                                                                                    \
	do {                                                                                                     \
		cudaError_t _e = (expr);                                                                             \
		if (_e != cudaSuccess) return npb_fail_cuda(ctx, _e, #expr, __FILE__, __LINE__);                     \
	} while (0)

__host__ __device__ constexpr int npb_tri(int D) { return D * (D + 1) / 2; }
// params per slot: mu[D], T2 upper-triangular packed row-wise [D(D+1)/2], c2
__host__ __device__ constexpr int npb_ps(int D) { return D + npb_tri(D) + 1; }
// offset of T(i,j), j>=i, in the row-wise packed upper triangle
__host__ __device__ constexpr int npb_tri_off(int D, int i, int j) { return i * D - i * (i - 1) / 2 + (j - i); }

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: key = (seed, chain), counter = (step, draw, sweep, purpose)
// ---------------------------------------------------------------------------------------------------------
struct Philox {
	uint32_t k0, k1;
	__host__ __device__ Philox(uint32_t a, uint32_t b) : k0(a), k1(b) {}
	__host__ __device__ inline void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) const {
		uint32_t ka = k0, kb = k1;
#pragma unroll
		for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
			uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
			uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
			uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
			uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
			uint32_t n0 = hi1 ^ c1 ^ ka, n1 = lo1, n2 = hi0 ^ c3 ^ kb, n3 = lo0;
			c0 = n0; c1 = n1; c2 = n2; c3 = n3;
			ka += 0x9E3779B9u;
			kb += 0xBB67AE85u;
		}
		out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
	}
};

enum { NPB_RNG_INIT_THETA = 1, NPB_RNG_INIT_Z = 2, NPB_RNG_AUX = 3, NPB_RNG_PICK = 4, NPB_RNG_SM = 5 };

// uniform in (0,1]: 24 significant bits of (r+1) * 2^-32, dense near 0 (the auxiliary candidates sit first in
// the cumulative sum, so small weights keep their resolution)
__device__ __forceinline__ float npb_u01(uint32_t r) {
	return __fmul_rn(__uint2float_rn(r) + 1.0f, 2.3283064365386963e-10f);
}
// Box-Muller on two 32-bit words
__device__ __forceinline__ void npb_normal2(uint32_t r0, uint32_t r1, float &n0, float &n1) {
	float u1 = npb_u01(r0);
	float rad = sqrtf(-2.0f * NPB_LN2 * __log2f(u1));
	float ang = __uint2float_rn(r1) * (6.283185307179586f * 2.3283064365386963e-10f);
	float s, c;
	__sincosf(ang, &s, &c);
	n0 = rad * c;
	n1 = rad * s;
}

// ---------------------------------------------------------------------------------------------------------
// Scan order of a sweep (np_mcmc.cpp:120-125 draws a fresh std::shuffle per sweep).  Here: a keyed
// pseudo-random permutation of [0,N) evaluated point-wise -- a 6-round balanced Feistel network over
// 2h >= log2(N) bits with cycle walking -- so that no permutation array has to be generated, stored or read,
// every chain of a launch shares the order (lockstep), and a run is restartable from (seed, sweep).
// ---------------------------------------------------------------------------------------------------------
struct ScanOrder {
	uint32_t key[6];
	uint32_t half_bits, half_mask;
	uint32_t N;
};
__host__ __device__ inline uint32_t npb_mix32(uint32_t x) {
	x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
	return x;
}
__host__ __device__ inline ScanOrder npb_scan_order(uint64_t seed, uint32_t sweep, uint32_t N) {
	ScanOrder so;
	uint32_t bits = 2;
	while ((1ull << bits) < (uint64_t)N) bits++;
	so.half_bits = (bits + 1) / 2;
	so.half_mask = (1u << so.half_bits) - 1u;
	so.N = N;
	Philox ph((uint32_t)seed, (uint32_t)(seed >> 32));
	uint32_t a[4], b[4];
	ph(sweep, 0x5CA10FDEu, 0, 0, a);
	ph(sweep, 0x5CA10FDEu, 1, 0, b);
	so.key[0] = a[0]; so.key[1] = a[1]; so.key[2] = a[2]; so.key[3] = a[3]; so.key[4] = b[0]; so.key[5] = b[1];
	return so;
}
__host__ __device__ inline uint32_t npb_scan_item(const ScanOrder &so, uint32_t s) {
	uint32_t x = s;
	do {
		uint32_t l = x >> so.half_bits, r = x & so.half_mask;
#pragma unroll
		for (int i = 0; i < 6; ++i) {
			uint32_t f = npb_mix32(r ^ so.key[i]) & so.half_mask;
			uint32_t nl = r;
			r = l ^ f;
			l = nl;
		}
		x = (l << so.half_bits) | r;
	} while (x >= so.N);
	return x;
}

// ---------------------------------------------------------------------------------------------------------
// Prior in device form.  Sigma' = v^2 A (A = L^T L, L = chol(Lambda), invwishart.h:38-43); A^-1 = C C^T.
//   xw      = C^T (x - mu0)                         (whitened data, precomputed per item)
//   theta'  : mu' = mu0 + (|v|/sqrt(kappa)) S z, S = C^-T  (any square root of Sigma'/kappa is equivalent to
//             the reference's eigen square root, normalinvwishart.h:56-61), T' = C^T / |v|
//   log2 N(x|theta') = c0_2 - D log2|v| - |xw/|v| - z/sqrt(kappa)|^2 * (log2e/2)
// ---------------------------------------------------------------------------------------------------------
struct PriorDev {
	int D;
	int m_aux;
	int flags;
	float inv_sqrt_kappa;
	float nu;            // used as the standard deviation of v (Q2/Q3: "variance" passed as stddev)
	float v_mean;        // = D
	float c0_2;          // -0.5*(D log2(2 pi) + log2 det A)
	float log2_alpha_m;  // log2(alpha / m)
	float mu0[NPB_MAX_D];
	const float *CT2;    // [tri] C^T packed upper, times sqrt(log2e/2)   (device pointer)
	const float *S;      // [tri] S = C^-T packed upper                  (device pointer)
};
