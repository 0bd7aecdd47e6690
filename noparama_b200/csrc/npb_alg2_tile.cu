// npb_alg2_tile.cu -- CONJUGATE Algorithm 2 (npb_alg2.cu has the model and its formulas), D = 16 and 64: a TILE of up to 64
// steps of the scan is evaluated ahead of the chain, and corrected after every move.
//
// The chain itself is sequential, but its state -- per cluster (n, mu_n, P = Lambda_n^-1, log det Lambda_n) -- changes only when
// an item MOVES: a step that puts the item back where it was restores the state exactly (the own cluster is evaluated with the
// item removed in closed form, npb_alg2.cu).  So for a tile of T consecutive steps
//   1. the quadratic forms t[j][k] = (x_j - mu_k)^T P_k (x_j - mu_k) of every step j against every cluster k are computed from the
//      state at the start of the tile, as a register-blocked FP32 product per cluster (a warp takes UI items x D columns of one
//      cluster; P_k is read once per UI items, from L1/L2, instead of once per step);
//   2. every step's race keys and winner follow in parallel (one warp per step);
//   3. the first step whose winner is not its current cluster is THE next event of the sequential chain: all steps before it are
//      final.  The move is applied (rank-1 down-date / up-date of the two P's, means, log-determinants, counts, FP64 statistics),
//      the two changed columns of t and of the keys are recomputed for the steps after it, winners are re-decided, and the search
//      continues behind the move.
// Every (step, cluster) value that is USED was computed by the same instruction sequence from the state the sequential chain has
// at that step, so the assignments do not depend on the tile length: a.tile = 1 is the strictly sequential schedule and gives the
// same bits (tests/test_gpu_alg2_conjugate.py).  The tile length adapts to the recent move rate (a move costs two columns of the
// remaining tile), which changes the cost only.
//
// One CTA of 8 warps per chain.  P [32, D, D] stays in global memory (512 KB per chain at D = 64; the CTA's own stores and loads
// of it are ordered by its barriers); counts, means, log-determinants live in shared memory for the launch.
#include "npb_alg2.cuh"

namespace {

template <int D> struct A2T {
	static constexpr int CG = D / 4;              // lanes across the columns of a row (4 columns each)
	static constexpr int IGW = 32 / CG;           // item groups per warp
	static constexpr int IT = D == 64 ? 8 : 4;    // items per lane
	static constexpr int UI = IGW * IT;           // items per warp unit: 16 (D = 64), 32 (D = 16)
	static constexpr int NCH = UI / 4;            // 16-byte chunks per row of a unit's difference tile
	static constexpr int TMAX = 64;               // steps per tile
	static constexpr int XS = TMAX + 4;           // row stride of the transposed item tile
	static constexpr int WARPS = 8;
	// shared memory, in floats
	static constexpr int O_XT = 0;                          // [D][XS] items of the tile, transposed
	static constexpr int O_DW = O_XT + D * XS;              // [WARPS][D][UI] x - mu of a unit, chunk-swizzled
	static constexpr int O_MU = O_DW + WARPS * D * UI;      // [32][D]
	static constexpr int O_TT = O_MU + 32 * D;              // [TMAX][33] quadratic forms (column 32 unused: the stride keeps a column conflict-free)
	static constexpr int O_KT = O_TT + TMAX * 33 + 1;       // [TMAX][33] race keys
	static constexpr int O_DM = O_KT + TMAX * 33 + 1;       // [2][D] x - mu of a moving item against (old, new) cluster
	static constexpr int O_PU = O_DM + 2 * D;               // [2][D] P (x - mu)
	static constexpr int O_XD = O_PU + 2 * D;               // [D] doubles: the moving item in FP64 (8-byte aligned: all terms even)
	static constexpr int O_LD = O_XD + 2 * D;               // [32] log det Lambda_n
	static constexpr int O_CNT = O_LD + 32;                 // [32] int
	static constexpr int O_ZOLD = O_CNT + 32;               // [TMAX] int
	static constexpr int O_ITEM = O_ZOLD + TMAX;            // [TMAX] int
	static constexpr int O_WIN = O_ITEM + TMAX;             // [TMAX] int
	static constexpr int O_RED = O_WIN + TMAX;              // [8] float
	static constexpr int O_SL = O_RED + 8;                  // [34] int: the occupied slots at the start of the tile, then 32 (a new cluster)
	static constexpr int FLOATS = O_SL + 36;
};

// the quadratic forms of the clusters in `mask` for the tile's items [j_lo, T): ttab[j][k]
template <int D>
__device__ __forceinline__ void a2_columns(float *sm, const float *Pc, unsigned mask, int j_lo, int T) {
	using L = A2T<D>;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int cg = lane % L::CG, ig = lane / L::CG;
	const int iu_lo = j_lo / L::UI, nups = (T + L::UI - 1) / L::UI - iu_lo;
	const int total = __popc(mask) * nups;
	float *dw = sm + L::O_DW + warp * (D * L::UI);
	const float *xT = sm + L::O_XT, *mus = sm + L::O_MU;
	float *ttab = sm + L::O_TT;
	for (int u = warp; u < total; u += L::WARPS) {
		const int k = __fns(mask, 0, u / nups + 1);
		const int u0 = (iu_lo + u % nups) * L::UI;
		// x - mu_k of the unit's items, row r = dimension, 16-byte chunks swizzled by the row's column group
#pragma unroll 4
		for (int e = lane; e < D * L::NCH; e += 32) {
			const int r = e / L::NCH, c = e % L::NCH;
			float4 v = *reinterpret_cast<const float4 *>(xT + r * L::XS + u0 + 4 * c);
			const float m = mus[k * D + r];
			v.x -= m; v.y -= m; v.z -= m; v.w -= m;
			*reinterpret_cast<float4 *>(dw + r * L::UI + (((c ^ (r >> 2)) & (L::NCH - 1)) << 2)) = v;
		}
		__syncwarp();
		float acc[L::IT][4];
#pragma unroll
		for (int it = 0; it < L::IT; ++it) { acc[it][0] = acc[it][1] = acc[it][2] = acc[it][3] = 0.0f; }
		const float *Pk = Pc + (size_t)k * D * D + 4 * cg;
#pragma unroll 16
		for (int r = 0; r < D; ++r) {
			const float4 p = *reinterpret_cast<const float4 *>(Pk + r * D);
			float dv[L::IT];
#pragma unroll
			for (int h = 0; h < L::IT / 4; ++h) {
				const int ch = ((ig * (L::IT / 4) + h) ^ (r >> 2)) & (L::NCH - 1);
				const float4 d4 = *reinterpret_cast<const float4 *>(dw + r * L::UI + 4 * ch);
				dv[4 * h] = d4.x; dv[4 * h + 1] = d4.y; dv[4 * h + 2] = d4.z; dv[4 * h + 3] = d4.w;
			}
#pragma unroll
			for (int it = 0; it < L::IT; ++it) {
				acc[it][0] = fmaf(p.x, dv[it], acc[it][0]);
				acc[it][1] = fmaf(p.y, dv[it], acc[it][1]);
				acc[it][2] = fmaf(p.z, dv[it], acc[it][2]);
				acc[it][3] = fmaf(p.w, dv[it], acc[it][3]);
			}
		}
		// t = sum over the columns of (P d)_c d_c: own four columns, then across the CG lanes of the item group
		float tt[L::IT];
#pragma unroll
		for (int it = 0; it < L::IT; ++it) tt[it] = 0.0f;
#pragma unroll
		for (int cc = 0; cc < 4; ++cc) {
			const int r = 4 * cg + cc;
#pragma unroll
			for (int h = 0; h < L::IT / 4; ++h) {
				const int ch = ((ig * (L::IT / 4) + h) ^ cg) & (L::NCH - 1);
				const float4 d4 = *reinterpret_cast<const float4 *>(dw + r * L::UI + 4 * ch);
				tt[4 * h] = fmaf(acc[4 * h][cc], d4.x, tt[4 * h]);
				tt[4 * h + 1] = fmaf(acc[4 * h + 1][cc], d4.y, tt[4 * h + 1]);
				tt[4 * h + 2] = fmaf(acc[4 * h + 2][cc], d4.z, tt[4 * h + 2]);
				tt[4 * h + 3] = fmaf(acc[4 * h + 3][cc], d4.w, tt[4 * h + 3]);
			}
		}
#pragma unroll
		for (int it = 0; it < L::IT; ++it) {
#pragma unroll
			for (int o = L::CG / 2; o > 0; o >>= 1) tt[it] += __shfl_xor_sync(0xffffffffu, tt[it], o);
			const int j = u0 + ig * L::IT + it;
			if (cg == 0 && j >= j_lo && j < T) ttab[j * 33 + k] = tt[it];
		}
		__syncwarp();
	}
}

// race key of step j against cluster k (k = 32: a new cluster), from the quadratic form and the cluster's count
template <int D>
__device__ __forceinline__ float a2_key(const A2Args &a, const float *sm, int j, int k, uint32_t step, uint32_t ka, uint32_t kb) {
	using L = A2T<D>;
	const int *cnt = reinterpret_cast<const int *>(sm + L::O_CNT), *zold = reinterpret_cast<const int *>(sm + L::O_ZOLD);
	const int *items = reinterpret_cast<const int *>(sm + L::O_ITEM);
	if (k == 32) return a.log2_alpha + __ldg(a.lp0 + items[j]) * NPB_LOG2E + a2_noise(ka ^ step, kb, 32u);
	const int n = cnt[k];
	const bool own = k == zold[j];
	const int n_eff = n - (own ? 1 : 0);
	if (n_eff <= 0) return -INFINITY;
	const float t = sm[L::O_TT + j * 33 + k];
	float q_eff = t, ld_eff = sm[L::O_LD + k];
	if (own) { // the item's own cluster with the item removed, in closed form (Sherman-Morrison)
		const float kp = a.kappa0 + (float)n, cdown = kp / (kp - 1.0f);
		const float one_m = fmaxf(1.0f - cdown * t, 1e-12f);
		q_eff = cdown * cdown * t / one_m;
		ld_eff += __logf(one_m);
	}
	const float kap = a.kappa0 + (float)n_eff;
	const float lp = __ldg(a.G + n_eff) - 0.5f * ld_eff - 0.5f * (a.nu0 + (float)n_eff + 1.0f) * log1pf(kap / (kap + 1.0f) * q_eff);
	return fast_lg2((float)n_eff) + lp * NPB_LOG2E + a2_noise(ka ^ step, kb, (uint32_t)k);
}

template <int D>
__global__ void __launch_bounds__(256, (D == 16 ? 3 : 2)) k_a2_tile(const A2Args a) {
	using L = A2T<D>;
	extern __shared__ __align__(16) float sm[];
	const int chain = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int C = a.C, N = a.N;
	float *xT = sm + L::O_XT, *mus = sm + L::O_MU, *ktab = sm + L::O_KT, *ldv = sm + L::O_LD;
	float *dm = sm + L::O_DM, *pu = sm + L::O_PU, *red = sm + L::O_RED;
	double *xd = reinterpret_cast<double *>(sm + L::O_XD);
	int *cnt = reinterpret_cast<int *>(sm + L::O_CNT), *zold = reinterpret_cast<int *>(sm + L::O_ZOLD);
	int *items = reinterpret_cast<int *>(sm + L::O_ITEM), *win = reinterpret_cast<int *>(sm + L::O_WIN), *slist = reinterpret_cast<int *>(sm + L::O_SL);
	float *Pc = a.P + (size_t)chain * 32 * D * D;
	double *sxc = a.sx + (size_t)chain * 32 * D, *sxxc = a.sxx + (size_t)chain * 32 * D * D;

	for (int e = tid; e < 32 * D; e += 256) mus[e] = a.mu[(size_t)chain * 32 * D + e];
	for (int e = tid; e < D * L::XS; e += 256) xT[e] = 0.0f;
	if (tid < 32) { cnt[tid] = a.counts[(size_t)chain * 32 + tid]; ldv[tid] = a.ld[(size_t)chain * 32 + tid]; }
	int kocc = a.kocc[chain];
	unsigned long long st_cand = 0ull, st_moved = 0ull, st_births = 0ull; // thread 0's are the ones written back
	const uint32_t ka = (uint32_t)a.seed ^ 0xA2A2A2A2u, k1 = (uint32_t)(a.seed >> 32) + (uint32_t)chain;
	const int tile_max = a.tile < 1 ? 1 : (a.tile > L::TMAX ? L::TMAX : a.tile);
	int tile = tile_max;
	__syncthreads();

	for (int sw = 0; sw < a.n_sweeps; ++sw) {
		const int32_t *order = a.order + (size_t)sw * N;
		const uint32_t kb = k1 ^ ((a.sweep0 + (uint32_t)sw) * 0x9E3779B9u);
		for (int s = 0; s < N;) {
			const int T = min(tile, N - s);
			// ---- the tile's items: ids, current clusters, rows (transposed) ----
			if (tid < T) {
				const int it = order[s + tid];
				items[tid] = it;
				zold[tid] = (int)a.z[(size_t)it * C + chain];
			}
			__syncthreads();
			for (int e = tid; e < L::TMAX * (D / 4); e += 256) {
				const int j = e & (L::TMAX - 1), q = e / L::TMAX;
				if (j < T) {
					const float4 v = __ldg(reinterpret_cast<const float4 *>(a.X + (size_t)items[j] * D) + q);
					xT[(4 * q) * L::XS + j] = v.x; xT[(4 * q + 1) * L::XS + j] = v.y;
					xT[(4 * q + 2) * L::XS + j] = v.z; xT[(4 * q + 3) * L::XS + j] = v.w;
				}
			}
			unsigned occ = 0u;
#pragma unroll
			for (int k = 0; k < 32; ++k) occ |= cnt[k] > 0 ? 1u << k : 0u;
			const int nocc = __popc(occ);
			if (tid < 32 && (occ >> tid & 1u)) slist[__popc(occ & ((1u << tid) - 1u))] = tid;
			if (tid == 32) slist[nocc] = 32;
			__syncthreads();
			a2_columns<D>(sm, Pc, occ, 0, T);
			__syncthreads();
			for (int kk = warp; kk <= nocc; kk += L::WARPS) { // a warp per candidate column (occupied slots, then a new cluster), lanes over the steps
				const int k = slist[kk];
				for (int j = lane; j < T; j += 32) ktab[j * 33 + k] = a2_key<D>(a, sm, j, k, (uint32_t)(s + j), ka, kb);
			}
			__syncthreads();
			int j0 = 0, tile_moves = 0;
			while (j0 < T) {
				// ---- winners of the steps not yet final ----
				for (int j = j0 + warp; j < T; j += L::WARPS) {
					const float key = cnt[lane] - (lane == zold[j] ? 1 : 0) > 0 ? ktab[j * 33 + lane] : -INFINITY; // only columns of clusters with (other) members hold keys
					const float top = fmaxf(redux_max_f32(key), ktab[j * 33 + 32]);
					const unsigned bal = __ballot_sync(0xffffffffu, key == top && key > -INFINITY);
					int w = bal ? __ffs(bal) - 1 : 32;
					if (w == 32) { // a new cluster needs a slot without members once the item is retracted; none: the item stays (code 33)
						const unsigned fb = __ballot_sync(0xffffffffu, cnt[lane] - (lane == zold[j] ? 1 : 0) <= 0);
						if (!fb) w = 33;
					}
					if (lane == 0) win[j] = w;
				}
				__syncthreads();
				// ---- the first step that does not simply stay ----
				const int ja = j0 + lane, jb = j0 + lane + 32;
				const unsigned e0 = __ballot_sync(0xffffffffu, ja < T && win[ja] != zold[ja]);
				const unsigned e1 = __ballot_sync(0xffffffffu, jb < T && win[jb] != zold[jb]);
				const int jm = e0 ? j0 + __ffs(e0) - 1 : (e1 ? j0 + 32 + __ffs(e1) - 1 : T);
				if (warp == 0) { // candidates weighed by the steps now final (the event step included)
					int cs = 0;
					if (ja <= jm && ja < T) cs += kocc - (cnt[zold[ja]] == 1 ? 1 : 0) + 1;
					if (jb <= jm && jb < T) cs += kocc - (cnt[zold[jb]] == 1 ? 1 : 0) + 1;
#pragma unroll
					for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
					st_cand += (unsigned long long)cs;
				}
				if (jm >= T) break;
				const int w = win[jm], src = zold[jm], item = items[jm];
				j0 = jm + 1;
				if (w == 33) { // no room for a new cluster: the item stays, the chain is reported
					if (tid == 0) a.overflow[chain] = 1;
					__syncthreads(); // win[] is rewritten next
					continue;
				}
				// ================= the move: src loses the item, dst gains it =================
				const bool born = w == 32;
				int dst = w;
				if (born) {
					dst = 0;
					while (cnt[dst] - (dst == src ? 1 : 0) > 0) ++dst;
				}
				const int n_src = cnt[src], n_eff = n_src - 1;
				const bool died = n_eff == 0;
				const int n_dst = born ? 0 : cnt[dst];
				if (tid < D) dm[tid] = xT[tid * L::XS + jm] - mus[src * D + tid];
				else if (tid < 2 * D) dm[tid] = xT[(tid - D) * L::XS + jm] - (born ? a.mu0[tid - D] : mus[dst * D + tid - D]);
				else if (tid < 3 * D) xd[tid - 2 * D] = a.X64[(size_t)item * D + tid - 2 * D];
				if (born) { // the new cluster starts from the prior
					for (int e = tid; e < D * D; e += 256) { Pc[(size_t)dst * D * D + e] = __ldg(a.P0 + e); sxxc[(size_t)dst * D * D + e] = 0.0; }
					if (tid < D) sxc[dst * D + tid] = 0.0;
				}
				__syncthreads();
				// P u of both clusters (P is symmetric: column sums, coalesced), u^T P u
				float prod = 0.0f;
				if (tid < 2 * D) {
					const int which = tid / D, r = tid % D;
					if (which == 1 || !died) {
						const float *Pk = Pc + (size_t)(which ? dst : src) * D * D + r;
						const float *dv = dm + which * D;
						float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll 4
						for (int c = 0; c < D; c += 4) {
							a0 = fmaf(Pk[(c) * D], dv[c], a0); a1 = fmaf(Pk[(c + 1) * D], dv[c + 1], a1);
							a2 = fmaf(Pk[(c + 2) * D], dv[c + 2], a2); a3 = fmaf(Pk[(c + 3) * D], dv[c + 3], a3);
						}
						const float v = (a0 + a1) + (a2 + a3);
						pu[tid] = v;
						prod = v * dv[r];
					}
				}
				if (warp < 2 * D / 32 || (D < 32 && warp == 0)) {
					if (D >= 32) {
#pragma unroll
						for (int o = 16; o > 0; o >>= 1) prod += __shfl_xor_sync(0xffffffffu, prod, o);
						if (lane == 0) red[warp] = prod;
					} else { // D = 16: lanes 0-15 the old cluster, 16-31 the new one
#pragma unroll
						for (int o = 8; o > 0; o >>= 1) prod += __shfl_xor_sync(0xffffffffu, prod, o);
						if ((lane & 15) == 0) red[lane >> 4] = prod;
					}
				}
				__syncthreads();
				float t_s, t_d;
				if (D >= 32) {
					t_s = 0.0f; t_d = 0.0f;
#pragma unroll
					for (int h = 0; h < D / 32; ++h) { t_s += red[h]; t_d += red[D / 32 + h]; }
				} else { t_s = red[0]; t_d = red[1]; }
				const float kp = a.kappa0 + (float)n_src, km = kp - 1.0f;
				const float cdown = kp / km, one_m = fmaxf(1.0f - cdown * t_s, 1e-12f), f_s = cdown / one_m;
				const float kap = a.kappa0 + (float)n_dst, kap1 = kap + 1.0f;
				const float cc = kap / kap1, den = 1.0f + cc * t_d, f_d = cc / den;
				// rank-1 down-date / up-date (products formed symmetrically: P stays symmetric bit for bit), statistics
				for (int e = tid; e < D * D; e += 256) {
					const int r = e / D, c = e % D;
					if (!died) {
						float *p = Pc + (size_t)src * D * D + e;
						*p = fmaf(f_s, pu[r] * pu[c], *p);
					}
					float *p2 = Pc + (size_t)dst * D * D + e;
					*p2 = fmaf(-f_d, pu[D + r] * pu[D + c], *p2);
					const double xx = xd[r] * xd[c];
					if (!(died && born && dst == src)) sxxc[(size_t)src * D * D + e] -= xx;
					sxxc[(size_t)dst * D * D + e] += xx;
				}
				if (tid < D) {
					const float x = xT[tid * L::XS + jm];
					if (!died) mus[src * D + tid] = (kp * mus[src * D + tid] - x) / km;
					if (!(died && born && dst == src)) sxc[src * D + tid] -= xd[tid];
				}
				__syncthreads(); // the old cluster's mean and first moment are settled before the new one's (the same slot when a lone member is born again)
				if (tid < D) {
					const float x = xT[tid * L::XS + jm];
					const float m0 = born ? a.mu0[tid] : mus[dst * D + tid];
					mus[dst * D + tid] = (kap * m0 + x) / kap1;
					sxc[dst * D + tid] += xd[tid];
				}
				if (tid == 0) {
					if (!died) ldv[src] += __logf(one_m);
					cnt[src] = n_eff;
					ldv[dst] = (born ? a.ld0 : ldv[dst]) + __logf(den);
					cnt[dst] = n_dst + 1;
					a.z[(size_t)item * C + chain] = (npb_z_t)dst;
					st_moved++;
					if (born) st_births++;
				}
				kocc += (born ? 1 : 0) - (died ? 1 : 0);
				++tile_moves;
				__syncthreads();
				if (j0 < T) { // the two changed columns for the steps behind the move
					const unsigned dirty = (died ? 0u : 1u << src) | 1u << dst;
					a2_columns<D>(sm, Pc, dirty, j0, T);
					__syncthreads();
					if (warp < 2 && !(warp == 0 && died)) {
						const int k = warp ? dst : src;
						for (int j = j0 + lane; j < T; j += 32) ktab[j * 33 + k] = a2_key<D>(a, sm, j, k, (uint32_t)(s + j), ka, kb);
					}
					__syncthreads();
				}
			}
			__syncthreads(); // win[], zold[], items[] are rewritten by the next tile
			s += T;
			// a move costs two columns of the rest of its tile: shorter tiles while many items move
			if (tile_moves * 8 > T) tile = max(tile / 2, min(tile_max, L::UI / 2));
			else if (tile_moves * 32 <= T) tile = min(tile * 2, tile_max);
		}
	}
	// ---- state back to memory ----
	__syncthreads();
	for (int e = tid; e < 32 * D; e += 256) a.mu[(size_t)chain * 32 * D + e] = mus[e];
	if (tid < 32) {
		a.counts[(size_t)chain * 32 + tid] = cnt[tid];
		a.ld[(size_t)chain * 32 + tid] = ldv[tid];
		const int o = __popc(__ballot_sync(0xffffffffu, cnt[tid] > 0));
		if (tid == 0) {
			a.kocc[chain] = o;
			a.st[(size_t)chain * 4 + 0] += st_cand;
			a.st[(size_t)chain * 4 + 1] += st_moved;
			a.st[(size_t)chain * 4 + 2] += st_births;
		}
	}
}

template <int D>
npb_status a2_tile_launch(npb_chains *ch, const A2Args &a) {
	npb_ctx *ctx = ch->ctx;
	const size_t bytes = (size_t)A2T<D>::FLOATS * sizeof(float);
	if (!ctx->a2_tile_attr_set) { // per device; both instantiations at once
		NPB_CUDA_OK(cudaFuncSetAttribute(k_a2_tile<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(A2T<16>::FLOATS * sizeof(float))));
		NPB_CUDA_OK(cudaFuncSetAttribute(k_a2_tile<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(A2T<64>::FLOATS * sizeof(float))));
		ctx->a2_tile_attr_set = true;
	}
	k_a2_tile<D><<<(unsigned)ch->C, 256, bytes, ctx->stream>>>(a);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

} // namespace

npb_status npb_launch_a2_tile(npb_chains *ch, const A2Args &a) {
	switch (ch->D) {
	case 16: return a2_tile_launch<16>(ch, a);
	case 64: return a2_tile_launch<64>(ch, a);
	default: return npb_fail(ch->ctx, NPB_E_UNSUPPORTED, "the tiled conjugate Algorithm 2 kernel covers D = 16, 64");
	}
}
