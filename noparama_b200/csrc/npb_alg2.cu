// npb_alg2.cu -- CONJUGATE Algorithm 2: collapsed Gibbs reassignment with the normal-inverse-Wishart posterior predictive
// (BASELINE.json north_star: "NIW-predictive log-likelihood ... np_suffies sufficient-statistic update ... suffstat deltas on
// removal and insert"; configs[3] "Neal Algorithm 2 conjugate NIW").
//
// The reference's intent is src/np_neal_algorithm2.cpp:32-120 -- retract (:46), updateSuffies (:54), weights n_k * p (:79-108)
// and alpha * p for a new cluster (:110-119) -- but that file is not compiled and cannot compile, and the conjugate arithmetic
// (include/statistics/normalinvwishart.h:66-75 `update(data, downdate)`, include/statistics/conjugate/*.h) does not exist.
// What runs here is therefore the textbook model, checked against the test oracle's fp64 restatement (itself pinned
// to scipy.stats.multivariate_t): parity with the reference is UNPINNED by construction.
//
// Per cluster the chain keeps, next to the member count, the posterior mean mu_n, P = Lambda_n^-1 and log det Lambda_n (fp32
// working state), and the sufficient statistics sum x, sum x x^T (fp64, global memory).  For an item x and a cluster:
//     t = (x - mu_n)^T P (x - mu_n),   log pred = G[n] - 1/2 log det Lambda_n - (nu_n + 1)/2 log1p(kappa_n / (kappa_n + 1) t)
// with G[n] = lgamma((nu_n + 1)/2) - lgamma((nu_n + 1 - D)/2) - D/2 log(pi (kappa_n + 1) / kappa_n) tabulated in fp64 per count.
// Removing the item from its own cluster needs no state change to be EVALUATED (Sherman-Morrison in closed form: with
// c = kappa_n / (kappa_n - 1), the quadratic form becomes c^2 t / (1 - c t) and log det gains log(1 - c t)), so a step that
// keeps the item where it is touches nothing -- "stay-moves restore identical suffstats" holds exactly.  Only a real move
// changes state: rank-1 down-date of the old cluster's P (P += c (P u)(P u)^T / (1 - c t)), rank-1 up-date of the new one's
// (P -= c' (P u)(P u)^T / (1 + c' t)), means, log-determinants, and +- x, x x^T on the fp64 statistics.  The derived state is
// rebuilt from the fp64 statistics (Cholesky, inverse) at the start of every launch, which bounds the drift of the fp32
// rank-1 updates to one launch.
//
// Mapping: one CTA per chain, LPS lanes per slot (32 slots): a lane owns D / LPS rows of its slot's P -- in registers
// (D = 16: 8 rows x 16 = 128 floats, two lanes per slot, 64-thread CTA), or read from L2 every step when 32 slots x D^2
// floats exceed the register file (D = 64: 512 KB per chain; 1024-thread CTA).  The categorical draw is the exponential race
// of the other sweep kernels.  The sequential chain is latency-bound; this is the first correct device path of configs[3], not
// yet a tuned one (the tensor-core precompute with rank-1 corrections of the few dirty slots, SURVEY 7.3-5, is the next step).
#include "npb_alg2.cuh"

npb_status npb_launch_a2_tile(npb_chains *ch, const A2Args &a); // npb_alg2_tile.cu
npb_status npb_launch_a2_tc(npb_chains *ch, const A2Args &a);   // npb_alg2_tc.cu
npb_status npb_launch_a2_tc16(npb_chains *ch, const A2Args &a); // npb_alg2_tc16.cu

template <int D, int LPS, bool PREG>
__global__ void __launch_bounds__(32 * LPS) k_a2_sweep(const A2Args a) {
	constexpr int RPL = D / LPS; // rows of P per lane
	static_assert(D % LPS == 0, "rows per lane");
	__shared__ float sd[32][D + 1];     // x - mu_n of every slot
	__shared__ float spd[2][D];         // P u of the cluster that loses / gains the item
	__shared__ float sxv[D];            // the item
	__shared__ float skey[33];
	__shared__ int s_pick[4];           // winner, members the old cluster keeps, -, free slot
	__shared__ float s_upd[2][2];       // (factor, -) of the down-date / up-date
	const int chain = blockIdx.x, tid = threadIdx.x;
	const int slot = tid / LPS, l = tid % LPS;
	const unsigned gmask = LPS >= 32 ? 0xffffffffu : (((1u << LPS) - 1u) << ((tid & 31) / LPS * LPS)); // the lanes of my slot
	const int C = a.C, N = a.N;
	float Prow[PREG ? RPL * D : 1];
	float mu_r[RPL];
	float *Pg = a.P + ((size_t)chain * 32 + slot) * D * D;
	float *mug = a.mu + ((size_t)chain * 32 + slot) * D;
	double *sxg = a.sx + ((size_t)chain * 32 + slot) * D, *sxxg = a.sxx + ((size_t)chain * 32 + slot) * D * D;
	int n = a.counts[(size_t)chain * 32 + slot];
	float logdet = a.ld[(size_t)chain * 32 + slot];
#pragma unroll
	for (int i = 0; i < RPL; ++i) {
		const int r = l + i * LPS;
		mu_r[i] = mug[r];
		if (PREG) {
#pragma unroll
			for (int c = 0; c < D; ++c) Prow[i * D + c] = Pg[r * D + c];
		}
	}
	int kocc = a.kocc[chain];
	unsigned long long st_cand = 0ull, st_moved = 0ull, st_births = 0ull;
	const uint32_t k0 = (uint32_t)a.seed ^ 0xA2A2A2A2u, k1 = (uint32_t)(a.seed >> 32) + (uint32_t)chain;

	for (int sw = 0; sw < a.n_sweeps; ++sw) {
		const int32_t *order = a.order + (size_t)sw * N;
		const uint32_t sweep = a.sweep0 + (uint32_t)sw;
		for (int s = 0; s < N; ++s) {
			const int item = order[s];
			const int zold = (int)a.z[(size_t)item * C + chain];
			if (tid < D) sxv[tid] = __ldg(a.X + (size_t)item * D + tid);
			__syncthreads();
			// ---- t = (x - mu_n)^T P (x - mu_n) of my slot ----
#pragma unroll
			for (int i = 0; i < RPL; ++i) sd[slot][l + i * LPS] = sxv[l + i * LPS] - mu_r[i];
			__syncwarp();
			float pd[RPL], qp = 0.0f;
#pragma unroll
			for (int i = 0; i < RPL; ++i) {
				const int r = l + i * LPS;
				float acc = 0.0f;
				if (n <= 0) { // a slot without members has no candidate (and, out of registers, no P worth 16 KB of L2 traffic)
				} else if (PREG) {
#pragma unroll
					for (int c = 0; c < D; ++c) acc = fmaf(Prow[i * D + c], sd[slot][c], acc);
				} else {
					const float4 *row = reinterpret_cast<const float4 *>(Pg + (size_t)r * D);
#pragma unroll 4
					for (int c4 = 0; c4 < D / 4; ++c4) {
						const float4 p = __ldcg(row + c4);
						acc = fmaf(p.x, sd[slot][4 * c4], acc); acc = fmaf(p.y, sd[slot][4 * c4 + 1], acc);
						acc = fmaf(p.z, sd[slot][4 * c4 + 2], acc); acc = fmaf(p.w, sd[slot][4 * c4 + 3], acc);
					}
				}
				pd[i] = acc;
				qp = fmaf(sd[slot][r], acc, qp);
			}
#pragma unroll
			for (int o = LPS / 2; o > 0; o >>= 1) qp += __shfl_xor_sync(0xffffffffu, qp, o);
			const float t = qp;
			// ---- the slot's key: log2(n_k pred_k(x)) + race noise; the item's own cluster is evaluated with the item removed ----
			const bool own = slot == zold;
			const int n_eff = n - (own ? 1 : 0);
			float q_eff = t, ld_eff = logdet, one_m = 1.0f, cdown = 0.0f;
			if (own && n_eff > 0) {
				const float kp = a.kappa0 + (float)n, km = kp - 1.0f;
				cdown = kp / km;
				one_m = fmaxf(1.0f - cdown * t, 1e-12f);
				q_eff = cdown * cdown * t / one_m;
				ld_eff = logdet + __logf(one_m);
			}
			if (l == 0) {
				float key = -INFINITY;
				if (n_eff > 0) {
					const float kap = a.kappa0 + (float)n_eff;
					const float lp = __ldg(a.G + n_eff) - 0.5f * ld_eff - 0.5f * (a.nu0 + (float)n_eff + 1.0f) * log1pf(kap / (kap + 1.0f) * q_eff);
					key = fast_lg2((float)n_eff) + lp * NPB_LOG2E + a2_noise(k0 ^ (uint32_t)s, k1 ^ (sweep * 0x9E3779B9u), (uint32_t)slot);
				}
				skey[slot] = key;
				if (own) s_pick[1] = n_eff;
			}
			if (tid == 0) skey[32] = a.log2_alpha + __ldg(a.lp0 + item) * NPB_LOG2E + a2_noise(k0 ^ (uint32_t)s, k1 ^ (sweep * 0x9E3779B9u), 32u);
			__syncthreads();
			if (tid < 32) {
				const float key = skey[tid];
				const float top = fmaxf(redux_max_f32(key), skey[32]);
				const unsigned bal = __ballot_sync(0xffffffffu, key == top && key > -INFINITY);
				int w = bal ? __ffs(bal) - 1 : 32;
				int fs = -1;
				if (w == 32) { // a new cluster: the lowest slot without members once the item is retracted
					const unsigned fb = __ballot_sync(0xffffffffu, a.counts[(size_t)chain * 32 + tid] - (tid == zold ? 1 : 0) <= 0);
					fs = fb ? __ffs(fb) - 1 : -1;
					if (fs < 0) { w = zold; if (tid == 0) a.overflow[chain] = 1; } // no room: the item stays, the chain is reported
				}
				if (tid == 0) { s_pick[0] = w; s_pick[3] = fs; }
			}
			__syncthreads();
			const int w = s_pick[0], fs = s_pick[3];
			const bool died = s_pick[1] == 0; // the item was its cluster's only member: the cluster is gone once it is retracted
			st_cand += (unsigned long long)(kocc - (died ? 1 : 0) + 1);
			if (w != zold) {
				const int dst = w == 32 ? fs : w;
				const bool born = w == 32;
				// ---- the old cluster loses the item (rank-1 down-date), the new one gains it (rank-1 up-date) ----
				if (slot == zold) {
#pragma unroll
					for (int i = 0; i < RPL; ++i) spd[0][l + i * LPS] = pd[i];
				}
				if (slot == dst) {
					if (born) { // starts from the prior: u = x - mu0, P = Lambda_0^-1
#pragma unroll
						for (int i = 0; i < RPL; ++i) {
							const int r = l + i * LPS;
							mu_r[i] = a.mu0[r];
							sd[slot][r] = sxv[r] - a.mu0[r];
						}
						__syncwarp(gmask);
						float qb = 0.0f;
#pragma unroll
						for (int i = 0; i < RPL; ++i) {
							const int r = l + i * LPS;
							float acc = 0.0f;
							for (int c = 0; c < D; ++c) {
								const float p = __ldg(a.P0 + r * D + c);
								if (PREG) Prow[i * D + c] = p; else Pg[(size_t)r * D + c] = p;
								acc = fmaf(p, sd[slot][c], acc);
							}
							pd[i] = acc;
							qb = fmaf(sd[slot][r], acc, qb);
						}
#pragma unroll
						for (int o = LPS / 2; o > 0; o >>= 1) qb += __shfl_xor_sync(gmask, qb, o);
						logdet = a.ld0;
						n = 0;
						if (l == 0) s_upd[1][1] = qb;
					} else if (l == 0) {
						s_upd[1][1] = t;
					}
#pragma unroll
					for (int i = 0; i < RPL; ++i) spd[1][l + i * LPS] = pd[i];
				}
				__syncthreads();
				if (slot == zold) {
					if (n_eff > 0) {
						const float f = cdown / one_m;
						const float kp = a.kappa0 + (float)n, km = kp - 1.0f;
#pragma unroll
						for (int i = 0; i < RPL; ++i) {
							const int r = l + i * LPS;
							if (PREG) {
#pragma unroll
								for (int c = 0; c < D; ++c) Prow[i * D + c] = fmaf(f * pd[i], spd[0][c], Prow[i * D + c]);
							} else {
								for (int c = 0; c < D; ++c) Pg[(size_t)r * D + c] = fmaf(f * pd[i], spd[0][c], Pg[(size_t)r * D + c]);
							}
							mu_r[i] = (kp * mu_r[i] - sxv[r]) / km;
						}
						logdet = ld_eff;
					}
					n -= 1;
#pragma unroll
					for (int i = 0; i < RPL; ++i) {
						const int r = l + i * LPS;
						const double *xd = a.X64 + (size_t)item * D;
						sxg[r] -= xd[r];
						for (int c = 0; c < D; ++c) sxxg[(size_t)r * D + c] -= xd[r] * xd[c];
					}
				}
				if (slot == dst) {
					const float tt = s_upd[1][1];
					const float kap = a.kappa0 + (float)n, kap1 = kap + 1.0f;
					const float cc = kap / kap1, den = 1.0f + cc * tt, f = cc / den;
#pragma unroll
					for (int i = 0; i < RPL; ++i) {
						const int r = l + i * LPS;
						if (PREG) {
#pragma unroll
							for (int c = 0; c < D; ++c) Prow[i * D + c] = fmaf(-f * pd[i], spd[1][c], Prow[i * D + c]);
						} else {
							for (int c = 0; c < D; ++c) Pg[(size_t)r * D + c] = fmaf(-f * pd[i], spd[1][c], Pg[(size_t)r * D + c]);
						}
						mu_r[i] = (kap * mu_r[i] + sxv[r]) / kap1;
					}
					logdet += __logf(den);
					n += 1;
#pragma unroll
					for (int i = 0; i < RPL; ++i) {
						const int r = l + i * LPS;
						if (born) { sxg[r] = 0.0; for (int c = 0; c < D; ++c) sxxg[(size_t)r * D + c] = 0.0; }
						const double *xd = a.X64 + (size_t)item * D;
						sxg[r] += xd[r];
						for (int c = 0; c < D; ++c) sxxg[(size_t)r * D + c] += xd[r] * xd[c];
					}
				}
				if (l == 0 && (slot == zold || slot == dst)) a.counts[(size_t)chain * 32 + slot] = n;
				if (tid == 0) a.z[(size_t)item * C + chain] = (npb_z_t)dst;
				kocc += (born ? 1 : 0) - (died ? 1 : 0);
				st_moved++;
				if (born) st_births++;
			}
			__syncthreads();
		}
	}
	// ---- state back to memory ----
#pragma unroll
	for (int i = 0; i < RPL; ++i) {
		const int r = l + i * LPS;
		mug[r] = mu_r[i];
		if (PREG) {
#pragma unroll
			for (int c = 0; c < D; ++c) Pg[r * D + c] = Prow[i * D + c];
		}
	}
	if (l == 0) a.ld[(size_t)chain * 32 + slot] = logdet;
	__syncthreads();
	if (tid < 32) {
		const int occ = __popc(__ballot_sync(0xffffffffu, a.counts[(size_t)chain * 32 + tid] > 0));
		if (tid == 0) {
			a.kocc[chain] = occ;
			a.st[(size_t)chain * 4 + 0] += st_cand;
			a.st[(size_t)chain * 4 + 1] += st_moved;
			a.st[(size_t)chain * 4 + 2] += st_births;
		}
	}
}

// ---------------------------------------------------------------------------------------------------------
// sufficient statistics of every (chain, slot) recounted from the assignments: CTA = (slot, chain)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_a2_recount(const double *X64, const npb_z_t *z, int N, int C, int D, double *sx, double *sxx, int *counts_out) {
	extern __shared__ double a2sm[]; // [D] the current member
	__shared__ int list[256];
	__shared__ int wcnt[8];
	const int slot = blockIdx.x, chain = blockIdx.y, tid = threadIdx.x;
	const int E = D * D + D; // elements: D of sum x, D^2 of sum x x^T
	double acc[17];          // up to (64 * 64 + 64) / 256 = 16.25 elements per thread
	for (int e = 0; e < 17; ++e) acc[e] = 0.0;
	int total = 0;
	for (int i0 = 0; i0 < N; i0 += 256) {
		// members of this block of 256 items, in item order (a fixed summation order: the recount is reproducible bit for bit)
		const int i = i0 + tid;
		const bool hit = i < N && (int)z[(size_t)i * C + chain] == slot;
		const unsigned bal = __ballot_sync(0xffffffffu, hit);
		__syncthreads(); // the previous block's list and warp counts have been read
		if ((tid & 31) == 0) wcnt[tid >> 5] = __popc(bal);
		__syncthreads();
		int base = 0, nl = 0;
		for (int w = 0; w < 8; ++w) { base += w < (tid >> 5) ? wcnt[w] : 0; nl += wcnt[w]; }
		if (hit) list[base + __popc(bal & ((1u << (tid & 31)) - 1u))] = i;
		__syncthreads();
		total += nl;
		for (int j = 0; j < nl; ++j) {
			const double *x = X64 + (size_t)list[j] * D;
			if (tid < D) a2sm[tid] = x[tid];
			__syncthreads();
			for (int e = 0, el = tid; el < E; ++e, el += 256)
				acc[e] += el < D ? a2sm[el] : a2sm[(el - D) / D] * a2sm[(el - D) % D];
			__syncthreads();
		}
	}
	double *osx = sx + ((size_t)chain * 32 + slot) * D, *osxx = sxx + ((size_t)chain * 32 + slot) * D * D;
	for (int e = 0, el = tid; el < E; ++e, el += 256) {
		if (el < D) osx[el] = acc[e]; else osxx[el - D] = acc[e];
	}
	if (tid == 0 && counts_out) counts_out[(size_t)chain * 32 + slot] = total;
}

// ---------------------------------------------------------------------------------------------------------
// derived state of every (chain, slot) from its statistics: posterior mean, Lambda_n -> Cholesky -> log det and inverse
// (fp64; one WARP per cluster, the matrix in shared memory [D][D + 1]; a slot without members is skipped: it is never read, a
// birth starts it from the prior -- except in the one-slot call that computes the prior's own Lambda_0^-1)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_a2_refresh(const int *counts, const double *sx, const double *sxx, const double *mu0, double kappa0,
		const double *Lambda0, int D, int n_slots, float *mu_out, float *P_out, float *ld_out) {
	extern __shared__ double a2S[]; // [D][D + 1]
	const int cs = blockIdx.x, lane = threadIdx.x;
	if (cs >= n_slots) return;
	const int n = counts[cs];
	if (n <= 0 && n_slots > 1) return;
	const int LD = D + 1;
	const double *s1 = sx + (size_t)cs * D, *s2 = sxx + (size_t)cs * D * D;
	const double kn = kappa0 + n, inv_n = n > 0 ? 1.0 / n : 0.0;
	// Lambda_n = Lambda_0 + sum x x^T - n xbar xbar^T + kappa_0 n / kappa_n (xbar - mu0)(xbar - mu0)^T
	for (int e = lane; e < D * D; e += 32) {
		const int a = e / D, b = e % D;
		double v = Lambda0[e];
		if (n > 0) {
			const double xa = s1[a] * inv_n, xb = s1[b] * inv_n;
			v += s2[e] - n * xa * xb + kappa0 * n / kn * (xa - mu0[a]) * (xb - mu0[b]);
		}
		a2S[a * LD + b] = v;
	}
	for (int a = lane; a < D; a += 32) mu_out[(size_t)cs * D + a] = (float)((kappa0 * mu0[a] + (n > 0 ? s1[a] : 0.0)) / kn);
	__syncwarp();
	// Cholesky in place (lower, right-looking): column j is scaled, then the rows below subtract their share of it
	double ld = 0.0;
	for (int j = 0; j < D; ++j) {
		double d = a2S[j * LD + j];
		d = d > 1e-300 ? d : 1e-300;
		const double ljj = sqrt(d);
		ld += log(d);
		__syncwarp();
		for (int i = j + 1 + lane; i < D; i += 32) a2S[i * LD + j] /= ljj;
		if (lane == 0) a2S[j * LD + j] = ljj;
		__syncwarp();
		for (int i = j + 1 + lane; i < D; i += 32) {
			const double lij = a2S[i * LD + j];
			for (int k = j + 1; k <= i; ++k) a2S[i * LD + k] -= lij * a2S[k * LD + j];
		}
		__syncwarp();
	}
	// W = L^-1 (lower), a lane per column j: W[j][j] = 1 / L[j][j], W[i][j] = -(sum_{j <= k < i} L[i][k] W[k][j]) / L[i][i]; the strict
	// lower part of W is stored transposed in the (now unused) strict upper triangle: W[i][j], i > j, lives at S[j][i] -- row j is the lane's own
	for (int j = lane; j < D; j += 32) {
		const double wjj = 1.0 / a2S[j * LD + j];
		for (int i = j + 1; i < D; ++i) {
			double s = a2S[i * LD + j] * wjj;
			for (int k = j + 1; k < i; ++k) s += a2S[i * LD + k] * a2S[j * LD + k];
			a2S[j * LD + i] = -s / a2S[i * LD + i];
		}
	}
	__syncwarp();
	// P = W^T W: P[a][b] = sum_{i >= a} W[i][a] W[i][b], b <= a, the pairs dealt over the lanes
	for (int e = lane; e < D * D; e += 32) {
		const int a = e / D, b = e % D;
		if (b > a) continue;
		double s = 0.0;
		for (int i = a; i < D; ++i) {
			const double wa = i == a ? 1.0 / a2S[a * LD + a] : a2S[a * LD + i];
			const double wb = i == b ? 1.0 / a2S[b * LD + b] : a2S[b * LD + i];
			s += wa * wb;
		}
		P_out[(size_t)cs * D * D + a * D + b] = (float)s;
		P_out[(size_t)cs * D * D + b * D + a] = (float)s;
	}
	if (lane == 0) ld_out[cs] = (float)ld;
}

// G[n] and the prior predictive of every item
__global__ void k_a2_gtable(float *G, int nmax, int D, double kappa0, double nu0) {
	const int n = blockIdx.x * blockDim.x + threadIdx.x;
	if (n > nmax) return;
	const double kn = kappa0 + n, nn = nu0 + n;
	G[n] = (float)(lgamma(0.5 * (nn + 1.0)) - lgamma(0.5 * (nn + 1.0 - D)) - 0.5 * D * log(M_PI * (kn + 1.0) / kn));
}
__global__ void k_a2_prior_pred(const double *X64, int N, int D, const double *mu0, const float *P0, float ld0, double kappa0, double nu0, float G0, float *lp0) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	double q = 0.0;
	for (int r = 0; r < D; ++r) {
		double acc = 0.0;
		for (int c = 0; c < D; ++c) acc += (double)P0[r * D + c] * (X64[(size_t)i * D + c] - mu0[c]);
		q += (X64[(size_t)i * D + r] - mu0[r]) * acc;
	}
	lp0[i] = (float)((double)G0 - 0.5 * (double)ld0 - 0.5 * (nu0 + 1.0) * log1p(kappa0 / (kappa0 + 1.0) * q));
}

// predictive log-density of given items under every cluster of one chain as the sweep kernel evaluates it (no removal), and under
// the prior (column 32): the parity probe.  out[j * 33 + k]; NaN for a slot without members.
__global__ void k_a2_probe(const A2Args a, int chain, const int32_t *items, int n_items, int D, float *out) {
	const int j = blockIdx.x, k = threadIdx.x;
	if (j >= n_items || k > 32) return;
	const int item = items[j];
	if (k == 32) { out[j * 33 + 32] = a.lp0[item]; return; }
	const int n = a.counts[(size_t)chain * 32 + k];
	if (n <= 0) { out[j * 33 + k] = NAN; return; }
	const float *P = a.P + ((size_t)chain * 32 + k) * D * D, *mu = a.mu + ((size_t)chain * 32 + k) * D;
	float t = 0.0f;
	for (int r = 0; r < D; ++r) {
		float acc = 0.0f;
		for (int c = 0; c < D; ++c) acc = fmaf(P[r * D + c], a.X[(size_t)item * D + c] - mu[c], acc);
		t = fmaf(a.X[(size_t)item * D + r] - mu[r], acc, t);
	}
	const float kap = a.kappa0 + (float)n;
	out[j * 33 + k] = a.G[n] - 0.5f * a.ld[(size_t)chain * 32 + k] - 0.5f * (a.nu0 + (float)n + 1.0f) * log1pf(kap / (kap + 1.0f) * t);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static npb_status a2_ensure(npb_chains *ch) {
	npb_ctx *ctx = ch->ctx;
	const PriorHost &p = ctx->prior;
	const int D = ch->D;
	const size_t CS = (size_t)ch->C * 32;
	if (ch->Kmax != 32) return npb_fail(ctx, NPB_E_UNSUPPORTED, "the conjugate Algorithm 2 kernel covers Kmax = 32");
	if (!(p.nu > D - 1.0)) return npb_fail(ctx, NPB_E_BAD_ARG, "the NIW posterior predictive needs nu > D - 1");
	if (D != 2 && D != 4 && D != 8 && D != 16 && D != 64) return npb_fail(ctx, NPB_E_UNSUPPORTED, "the conjugate Algorithm 2 kernel covers D = 2, 4, 8, 16, 64");
	if (!ch->a2_sx) {
		NPB_CUDA_OK(cudaMalloc((void **)&ch->a2_sx, CS * D * sizeof(double)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->a2_sxx, CS * D * D * sizeof(double)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->a2_mu, CS * D * sizeof(float)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->a2_P, CS * D * D * sizeof(float)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->a2_ld, CS * sizeof(float)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->a2_G, (size_t)(ch->ds->N + 2) * sizeof(float)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->a2_lp0, (size_t)ch->ds->N * sizeof(float)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->a2_prior, (size_t)(D + D * D) * sizeof(double) + (size_t)D * D * sizeof(float) + 64));
		ch->a2_prior_epoch = ~0ull;
		ch->a2_gen = ~0ull;
	}
	return NPB_OK;
}

static A2Args a2_args(npb_chains *ch, int n_sweeps) {
	const PriorHost &p = ch->ctx->prior;
	A2Args a;
	memset(&a, 0, sizeof(a));
	a.X = ch->ds->X32;
	a.X64 = ch->ds->X64;
	a.order = ch->scan_order;
	a.z = ch->z;
	a.counts = ch->counts;
	a.kocc = ch->kocc;
	a.overflow = ch->overflow;
	a.st = ch->st;
	a.sx = ch->a2_sx;
	a.sxx = ch->a2_sxx;
	a.mu = ch->a2_mu;
	a.P = ch->a2_P;
	a.ld = ch->a2_ld;
	a.G = ch->a2_G;
	a.lp0 = ch->a2_lp0;
	a.P0 = reinterpret_cast<const float *>(reinterpret_cast<const char *>(ch->a2_prior) + (size_t)(ch->D + ch->D * ch->D) * sizeof(double));
	for (int d = 0; d < ch->D; ++d) a.mu0[d] = (float)p.mu0[d];
	a.ld0 = ch->a2_ld0;
	a.kappa0 = (float)p.kappa;
	a.nu0 = (float)p.nu;
	a.log2_alpha = (float)log2(p.alpha);
	a.N = (int)ch->ds->N;
	a.C = (int)ch->C;
	a.n_sweeps = n_sweeps;
	a.sweep0 = ch->sweep;
	a.seed = ch->seed;
	a.tile = ch->sw.a2_tile;
	return a;
}

// prior-dependent tables (G, prior predictive of the items, Lambda_0^-1) and, when the assignments changed behind this path's
// back, the statistics recounted from z; then the derived state from the statistics
static npb_status a2_sync(npb_chains *ch, bool force_recount) {
	npb_ctx *ctx = ch->ctx;
	const PriorHost &p = ctx->prior;
	const int D = ch->D, N = (int)ch->ds->N;
	const int CS = (int)ch->C * 32;
	double *d_mu0 = ch->a2_prior, *d_L0 = ch->a2_prior + D;
	float *d_P0 = reinterpret_cast<float *>(reinterpret_cast<char *>(ch->a2_prior) + (size_t)(D + D * D) * sizeof(double));
	if (ch->a2_prior_epoch != ctx->prior_epoch) {
		NPB_CUDA_OK(cudaMemcpyAsync(d_mu0, p.mu0.data(), sizeof(double) * D, cudaMemcpyHostToDevice, ctx->stream));
		NPB_CUDA_OK(cudaMemcpyAsync(d_L0, p.Lambda.data(), sizeof(double) * D * D, cudaMemcpyHostToDevice, ctx->stream));
		// Lambda_0^-1 and log det Lambda_0 through the same refresh kernel: one "cluster" without members
		int zero = 0;
		int *d_zero = reinterpret_cast<int *>(ch->a2_ld); // scratch: overwritten by the real refresh below
		NPB_CUDA_OK(cudaMemcpyAsync(d_zero, &zero, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
		float *d_tmp = ch->a2_mu; // scratch for the posterior mean of the empty cluster
		k_a2_refresh<<<1, 32, sizeof(double) * D * (D + 1), ctx->stream>>>(d_zero, ch->a2_sx, ch->a2_sxx, d_mu0, p.kappa, d_L0, D, 1, d_tmp, d_P0, ch->a2_ld + 1);
		NPB_CUDA_OK(cudaGetLastError());
		NPB_CUDA_OK(cudaMemcpyAsync(&ch->a2_ld0, ch->a2_ld + 1, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
		NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
		k_a2_gtable<<<(N + 2 + 255) / 256, 256, 0, ctx->stream>>>(ch->a2_G, N + 1, D, p.kappa, p.nu);
		NPB_CUDA_OK(cudaGetLastError());
		float G0 = 0.0f;
		NPB_CUDA_OK(cudaMemcpyAsync(&G0, ch->a2_G, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
		NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
		k_a2_prior_pred<<<(N + 255) / 256, 256, 0, ctx->stream>>>(ch->ds->X64, N, D, d_mu0, d_P0, ch->a2_ld0, p.kappa, p.nu, G0, ch->a2_lp0);
		NPB_CUDA_OK(cudaGetLastError());
		ch->a2_prior_epoch = ctx->prior_epoch;
		force_recount = true;
	}
	if (force_recount || ch->a2_gen != ch->z_gen) {
		dim3 grid(32, (unsigned)ch->C);
		k_a2_recount<<<grid, 256, sizeof(double) * D, ctx->stream>>>(ch->ds->X64, ch->z, N, (int)ch->C, D, ch->a2_sx, ch->a2_sxx, ch->counts);
		NPB_CUDA_OK(cudaGetLastError());
		ch->a2_gen = ch->z_gen;
	}
	k_a2_refresh<<<CS, 32, sizeof(double) * D * (D + 1), ctx->stream>>>(ch->counts, ch->a2_sx, ch->a2_sxx, d_mu0, p.kappa, d_L0, D, CS, ch->a2_mu, ch->a2_P, ch->a2_ld);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

npb_status npb_launch_alg2_conjugate(npb_chains *ch, int n_sweeps) {
	npb_ctx *ctx = ch->ctx;
	npb_status s = a2_ensure(ch);
	if (s != NPB_OK) return s;
	const int N = (int)ch->ds->N;
	if (!ch->scan_order) {
		const size_t per_sweep = (size_t)N * sizeof(int32_t);
		size_t cap = (64u << 20) / per_sweep;
		ch->scan_cap = (int)(cap < 1 ? 1 : (cap > 1024 ? 1024 : cap));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->scan_order, per_sweep * ch->scan_cap));
	}
	for (int done = 0; done < n_sweeps;) {
		const int n = n_sweeps - done < ch->scan_cap ? n_sweeps - done : ch->scan_cap;
		s = a2_sync(ch, false);
		if (s != NPB_OK) return s;
		s = npb_launch_scan_order(ch, n);
		if (s != NPB_OK) return s;
		const A2Args a = a2_args(ch, n);
		const unsigned C = (unsigned)ch->C;
		switch (ch->D) {
		case 2: k_a2_sweep<2, 1, true><<<C, 32, 0, ctx->stream>>>(a); break;
		case 4: k_a2_sweep<4, 1, true><<<C, 32, 0, ctx->stream>>>(a); break;
		case 8: k_a2_sweep<8, 1, true><<<C, 32, 0, ctx->stream>>>(a); break;
		case 16:
			if (ch->sw.a2_tile > 0) { s = ch->sw.a2_tc16 ? npb_launch_a2_tc16(ch, a) : npb_launch_a2_tile(ch, a); if (s != NPB_OK) return s; }
			else k_a2_sweep<16, 2, true><<<C, 64, 0, ctx->stream>>>(a);
			break;
		case 64:
			if (ch->sw.a2_tile > 0) { s = ch->sw.a2_tc ? npb_launch_a2_tc(ch, a) : npb_launch_a2_tile(ch, a); if (s != NPB_OK) return s; }
			else k_a2_sweep<64, 32, false><<<C, 1024, 0, ctx->stream>>>(a);
			break;
		default: return npb_fail(ctx, NPB_E_UNSUPPORTED, "D");
		}
		NPB_CUDA_OK(cudaGetLastError());
		ch->sweep += (uint32_t)n;
		done += n;
	}
	// the assignments changed, through this path: its own statistics followed them
	ch->z_gen++;
	ch->a2_gen = ch->z_gen;
	return NPB_OK;
}

npb_status npb_launch_alg2_probe(npb_chains *ch, int chain, const int32_t *d_items, int n_items, float *d_out) {
	npb_ctx *ctx = ch->ctx;
	npb_status s = a2_ensure(ch);
	if (s != NPB_OK) return s;
	s = a2_sync(ch, false);
	if (s != NPB_OK) return s;
	const A2Args a = a2_args(ch, 0);
	k_a2_probe<<<n_items, 64, 0, ctx->stream>>>(a, chain, d_items, n_items, ch->D, d_out);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}
