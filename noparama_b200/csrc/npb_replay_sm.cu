// npb_replay_sm.cu -- parity level 2 for the split-merge samplers: one chain of JainNealAlgorithm::update
// (src/np_jain_neal_algorithm.cpp:424-502) or TriadicAlgorithm::update (src/np_triadic_algorithm.cpp:633-795) in double
// precision with the reference's own arithmetic -- linear-domain SAMS weights p(x|theta_q) |P_q| for the triadic sampler,
// the "log-density + count" rule with its lower_bound behaviour for Jain-Neal, cumulative-sum picks -- consuming
// a recorded trace (subsets, the new cluster's prior draw, the shuffled pool, every uniform) instead of Philox.  With the
// recorded draws replayed the device must reproduce every allocation, every accept/reject and the assignments bit for bit.
// One warp: lanes evaluate the densities of 32 pool members at a time, lane 0 runs the sequential scan.
#include "npb_internal.h"

// theta in double: mu[D], T upper packed [tri], c = -0.5 (D log 2pi + log det Sigma), cst = sqrt((2 pi)^D det Sigma)
__host__ __device__ constexpr int npb_psr(int D) { return npb_ps(D) + 1; }

__device__ inline double rsm_exponent(const double *x, const double *th, int D) {
	double q = 0.0;
	for (int r = 0; r < D; ++r) {
		double y = 0.0;
		for (int c = r; c < D; ++c) y += th[D + npb_tri_off(D, r, c)] * (x[c] - th[c]);
		q += y * y;
	}
	return -0.5 * q;
}
// std::lower_bound on a (possibly non-monotone) array of n <= 3 entries, the way libstdc++ bisects
__device__ inline int rsm_lower_bound(const double *c, int n, double v) {
	int first = 0, len = n;
	while (len > 0) {
		const int half = len >> 1, mid = first + half;
		if (c[mid] < v) { first = mid + 1; len = len - half - 1; }
		else len = half;
	}
	return first;
}

__global__ void __launch_bounds__(32) k_replay_split_merge(RsmArgs a) {
	__shared__ double s_e[3][32];   // exponent of the member under theta k
	__shared__ int s_plan[16];
	__shared__ double s_acc[8];
	const int lane = threadIdx.x, D = a.D, PSR = npb_psr(D), TRI = npb_tri(D);
	for (int64_t p = 0; p < a.n_prop; ++p) {
		// ---- plan: exactly the branches of update() ----
		if (lane == 0) {
			const int nsub = a.sampler == NPB_JAIN_NEAL ? 2 : 3;
			int pk[3], cl[3];
			for (int j = 0; j < nsub; ++j) { pk[j] = a.picks[p * 3 + j]; cl[j] = a.z[pk[j]]; }
			int type, nth, nsrc, Q, nskip = 0, dying = -1, th[3] = {-1, -1, -1}, tg[3] = {-1, -1, -1}, sp[3] = {-1, -1, -1};
			if (a.sampler == NPB_JAIN_NEAL) {
				if (cl[0] == cl[1]) { type = 0; nth = 2; nsrc = 1; Q = 2; nskip = 2; th[0] = cl[0]; tg[0] = cl[0]; sp[0] = pk[1]; sp[1] = pk[0]; }
				else { type = 1; nth = 2; nsrc = 1; Q = 1; th[0] = cl[0]; th[1] = cl[1]; tg[0] = cl[1]; dying = cl[0]; }
			} else {
				const int uniq = 1 + (cl[1] != cl[0]) + (cl[2] != cl[0] && cl[2] != cl[1]);
				const int dup = (cl[1] == cl[0]) ? 1 : 2; // duplicate_pick, dim1algebra.hpp:2115-2137
				if (uniq == 1) { type = 2; nth = 2; nsrc = 1; Q = 2; nskip = 2; th[0] = cl[0]; tg[0] = cl[0]; sp[0] = pk[0]; sp[1] = pk[2]; }
				else if (a.u0[p] < 0.5) {
					const int a1 = dup == 1 ? 2 : 1;
					type = 3; nth = 2; nsrc = 2; Q = 1; nskip = 1; th[0] = cl[0]; th[1] = cl[a1]; tg[0] = cl[0]; sp[0] = pk[0]; dying = cl[a1];
				} else if (uniq == 2) {
					const int o1 = dup == 1 ? 2 : 1, o2 = dup == 1 ? 1 : 2;
					type = 2; nth = 3; nsrc = 2; Q = 3; nskip = 3; th[0] = cl[0]; th[1] = cl[o1]; tg[0] = cl[0]; tg[1] = cl[o1];
					sp[0] = pk[0]; sp[1] = pk[o1]; sp[2] = pk[o2];
				} else { type = 3; nth = 3; nsrc = 3; Q = 2; nskip = 2; th[0] = cl[0]; th[1] = cl[1]; th[2] = cl[2]; tg[0] = cl[0]; tg[1] = cl[1];
					sp[0] = pk[0]; sp[1] = pk[1]; dying = cl[2]; }
			}
			s_plan[0] = type; s_plan[1] = nth; s_plan[2] = nsrc; s_plan[3] = Q; s_plan[4] = nskip; s_plan[5] = dying;
			for (int k = 0; k < 3; ++k) { s_plan[6 + k] = th[k]; s_plan[9 + k] = tg[k]; s_plan[12 + k] = sp[k]; }
			a.type_out[p] = type;
		}
		__syncwarp();
		const int type = s_plan[0], nth = s_plan[1], nsrc = s_plan[2], Q = s_plan[3], nskip = s_plan[4], dying = s_plan[5];
		const bool is_split = type == 0 || type == 2;
		const double *thp[3];
		for (int k = 0; k < nth; ++k) thp[k] = s_plan[6 + k] >= 0 ? a.theta + (size_t)s_plan[6 + k] * PSR : a.th_new + (size_t)p * PSR;
		const int64_t off = a.pool_off[p];
		const int npool = (int)(a.pool_off[p + 1] - off);
		// lane 0 state
		int npart[3] = {0, 0, 0};
		for (int q = 0; q < nskip; ++q) npart[q] = 1;
		double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}; // S[q][k]: sum over part q of log p(x | theta_k), in visiting order
		double own[3] = {0, 0, 0};
		for (int t0 = 0; t0 < npool; t0 += 32) {
			const int t = t0 + lane;
			if (t < npool) {
				const double *x = a.X + (size_t)a.pool[off + t] * D;
				for (int k = 0; k < nth; ++k) s_e[k][lane] = rsm_exponent(x, thp[k], D);
			}
			__syncwarp();
			if (lane == 0) {
				const int lim = min(32, npool - t0);
				for (int j = 0; j < lim; ++j) {
					const int id = a.pool[off + t0 + j];
					const double u = a.us[off + t0 + j];
					double lp[3];
					for (int k = 0; k < nth; ++k) lp[k] = s_e[k][j] - log(thp[k][D + TRI + 1]); // exponent - log(constant)
					int d;
					if (type == 1) d = 0;
					else if (u < 0.0) { // a pick that seeded its part
						d = -1;
						for (int q = 0; q < nskip; ++q)
							if (s_plan[12 + q] == id) d = q;
						if (d < 0) { *a.status = NPB_E_REPLAY_MISMATCH; d = 0; }
					} else if (type == 0) { // np_jain_neal_algorithm.cpp:158,165: log-density + count, then the pick
						double c[2];
						c[0] = lp[0] + (double)npart[0];
						c[1] = c[0] + (lp[1] + (double)npart[1]);
						d = rsm_lower_bound(c, 2, u * c[1]) == 0 ? 0 : 1;
						npart[d]++;
					} else { // np_triadic_algorithm.cpp:189,272: probability * |P_q|, cumulative sum, lower_bound
						double c[3], run = 0.0;
						for (int q = 0; q < Q; ++q) {
							run += exp(s_e[q][j]) / thp[q][D + TRI + 1] * (double)npart[q];
							c[q] = run;
						}
						d = rsm_lower_bound(c, Q, u * c[Q - 1]);
						if (d >= Q) { *a.status = NPB_E_REPLAY_MISMATCH; d = Q - 1; }
						npart[d]++;
					}
					a.dec_out[off + t0 + j] = d;
					for (int k = 0; k < nth; ++k) S[d][k] += lp[k];
					// own-cluster sums (triadic): the member's density under the cluster it currently sits in
					const int zz = a.z[id];
					for (int k = 0; k < nsrc; ++k)
						if (zz == s_plan[6 + k]) own[k] += lp[k];
				}
			}
			__syncwarp();
		}
		if (lane == 0) {
			const double la = log(a.alpha);
			double logA;
			if (type == 0) {
				logA = 0.0 + (la + lgamma((double)npart[1]) + lgamma((double)npart[0]) - lgamma((double)(npart[0] + npart[1]))) + (S[1][1] - S[1][0]);
			} else if (type == 1) {
				const int n0 = a.counts[s_plan[6]], n1 = a.counts[s_plan[7]];
				logA = 0.0 + -(la + lgamma((double)n0) + lgamma((double)n1) - lgamma((double)(n0 + n1))) + (S[0][1] - S[0][0]);
				npart[0] = n0 + n1;
			} else {
				if (Q == 1) npart[0] = npool;
				double lg = 0.0, rLd = 0.0, rLdp = 0.0;
				if (type == 2) {
					for (int q = 0; q < Q; ++q) lg += lgamma((double)npart[q]);
					for (int k = 0; k < nsrc; ++k) lg -= lgamma((double)a.counts[s_plan[6 + k]]);
				} else {
					for (int k = 0; k < nsrc; ++k) lg += lgamma((double)a.counts[s_plan[6 + k]]);
					for (int q = 0; q < Q; ++q) lg -= lgamma((double)npart[q]);
				}
				const double rP = type == 2 ? la + lg : -(la + lg);
				double rR;
				if (type == 2) rR = nsrc == 1 ? log(0.5) : -log(1.0 - 0.5);
				else rR = nsrc == 2 ? -log(0.5) : log(1.0 - 0.5);
				for (int k = 0; k < nsrc; ++k) rLd += own[k];
				for (int q = 0; q < Q; ++q) rLdp += S[q][q];
				logA = 0.0 + rP + rR + (rLdp - rLd);
			}
			const int accept = !(exp(logA) < a.uacc[p]);
			a.logA_out[p] = logA;
			a.accept_out[p] = accept;
			s_plan[15] = accept;
			if (accept) {
				int ns = -1;
				if (is_split) {
					ns = a.new_slot[p];
					if (ns < 0 || ns >= a.nslots || a.counts[ns] != 0) { *a.status = NPB_E_REPLAY_MISMATCH; ns = 0; }
				}
				for (int q = 0; q < Q; ++q) a.counts[s_plan[9 + q] >= 0 ? s_plan[9 + q] : ns] = npart[q];
				if (dying >= 0) a.counts[dying] = 0;
				s_plan[14] = ns; // (overwrites the third seed pick, no longer needed)
			}
		}
		__syncwarp();
		if (s_plan[15]) {
			const int ns = s_plan[14];
			for (int t = lane; t < npool; t += 32) {
				const int q = a.dec_out[off + t];
				a.z[a.pool[off + t]] = s_plan[9 + q] >= 0 ? s_plan[9 + q] : ns;
			}
			if (is_split)
				for (int t = lane; t < PSR; t += 32) a.theta[(size_t)ns * PSR + t] = a.th_new[(size_t)p * PSR + t];
		}
		__syncwarp();
	}
}

npb_status npb_launch_replay_sm(npb_ctx *ctx, const RsmArgs &a) {
	k_replay_split_merge<<<1, 32, 0, ctx->stream>>>(a);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}
