// npb_replay.cu -- parity level 2: one chain of NealAlgorithm8::update (src/np_neal_algorithm8.cpp:49-167) in
// double precision with the reference's LINEAR-domain weights, consuming a recorded trace (scan order, candidate
// order, auxiliary thetas, uniform draws) instead of Philox.  With the recorded draws replayed the device must
// reproduce the recorded assignments bit for bit.
//
// Semantics kept from the reference:
//   w_k     = probability(x | theta_k) * count(k)        for k in the recorded (hash-map) order   :93-109
//   w_{K+m} = probability(x | theta'_m) * alpha / M                                              :119-126
//   pick    = first j with cumsum_j >= u * cumsum_last, cumsum accumulated left to right in double
//             (dim1algebra.hpp:2078-2104); all-zero weights give index 0 (Q6)
// One warp: lanes evaluate the candidates' densities, lane 0 does the ordered cumulative sum.
#include "npb_internal.h"

#define NPB_REPLAY_MAXCAND 1024

__device__ inline double replay_pdf(const double *x, const double *th, int D) {
	// th: mu[D], T upper packed [tri], c = -0.5 (D log 2pi + log det Sigma)
	const int TRI = npb_tri(D);
	double q = 0.0;
	for (int r = 0; r < D; ++r) {
		double y = 0.0;
		for (int c = r; c < D; ++c) y += th[D + npb_tri_off(D, r, c)] * (x[c] - th[c]);
		q += y * y;
	}
	return exp(th[D + TRI] - 0.5 * q);
}

__global__ void __launch_bounds__(32) k_replay_alg8(const double *X, int N, int D, int M, double alpha, int nslots,
		double *theta /* [nslots][PS] */, int *counts /* [nslots] */, int32_t *z /* [N] */, int64_t n_steps,
		const int32_t *item, const int64_t *order_off, const int32_t *order, const double *aux_theta /* [S][M][PS] */,
		const double *u, const int32_t *new_slot, int32_t *picked_out, int64_t z_every, int32_t *z_after, int *status) {
	__shared__ double w[NPB_REPLAY_MAXCAND];
	__shared__ int sh_pick;
	const int lane = threadIdx.x;
	const int PS = npb_ps(D);
	for (int64_t s = 0; s < n_steps; ++s) {
		const int i = item[s];
		const double *x = X + (size_t)i * D;
		const int64_t off = order_off[s];
		const int K = (int)(order_off[s + 1] - off);
		if (K + M > NPB_REPLAY_MAXCAND) { if (lane == 0) *status = NPB_E_KMAX_OVERFLOW; return; }
		if (lane == 0) counts[z[i]] -= 1; // retract (membertrix.cpp:175-233)
		__syncwarp();
		for (int k = lane; k < K + M; k += 32) {
			if (k < K) {
				const int slot = order[off + k];
				const int c = (slot >= 0 && slot < nslots) ? counts[slot] : 0;
				w[k] = c > 0 ? replay_pdf(x, theta + (size_t)slot * PS, D) * (double)c : 0.0;
			} else {
				w[k] = replay_pdf(x, aux_theta + ((size_t)s * M + (k - K)) * PS, D) * alpha / (double)M;
			}
		}
		__syncwarp();
		if (lane == 0) {
			double total = 0.0;
			for (int k = 0; k < K + M; ++k) total += w[k];
			const double target = u[s] * total;
			double c = 0.0;
			int j = K + M; // lower_bound's "end"
			for (int k = 0; k < K + M; ++k) {
				c += w[k];
				if (c >= target) { j = k; break; }
			}
			sh_pick = j;
			picked_out[s] = j;
		}
		__syncwarp();
		const int j = sh_pick;
		if (j >= K + M) { if (lane == 0) *status = NPB_E_REPLAY_MISMATCH; return; }
		if (j >= K) {
			const int ns = new_slot[s];
			if (ns < 0 || ns >= nslots) { if (lane == 0) *status = NPB_E_REPLAY_MISMATCH; return; }
			const double *src = aux_theta + ((size_t)s * M + (j - K)) * PS;
			for (int t = lane; t < PS; t += 32) theta[(size_t)ns * PS + t] = src[t];
			if (lane == 0) { counts[ns] = 1; z[i] = ns; }
		} else if (lane == 0) {
			const int slot = order[off + j];
			counts[slot] += 1;
			z[i] = slot;
		}
		__syncwarp();
		if (z_after && z_every > 0 && (s + 1) % z_every == 0) {
			int32_t *dst = z_after + ((s + 1) / z_every - 1) * (int64_t)N;
			for (int t = lane; t < N; t += 32) dst[t] = z[t];
			__syncwarp();
		}
	}
}

npb_status npb_launch_replay(npb_ctx *ctx, const double *X, int N, int D, int M, double alpha, int nslots, double *theta,
		int *counts, int32_t *z, int64_t n_steps, const int32_t *item, const int64_t *order_off, const int32_t *order,
		const double *aux_theta, const double *u, const int32_t *new_slot, int32_t *picked, int64_t z_every,
		int32_t *z_after, int *status) {
	k_replay_alg8<<<1, 32, 0, ctx->stream>>>(X, N, D, M, alpha, nslots, theta, counts, z, n_steps, item, order_off, order,
			aux_theta, u, new_slot, picked, z_every, z_after, status);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}
