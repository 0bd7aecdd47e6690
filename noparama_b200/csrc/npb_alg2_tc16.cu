// npb_alg2_tc16.cu -- CONJUGATE Algorithm 2 at D = 16 on tcgen05: npb_alg2_tc.cu's schedule (a tile of 128 steps evaluated ahead of the
// chain, the first step that moves is the next event, the two changed clusters evaluated again behind it) with the operand layout of
// the D = 16 Algorithm 8 kernels (npb_alg8_gemm.cu, section 4.8 of DESIGN.md): a 16-D product is ONE K = 16 MMA step, so the three FP16
// products hi hi + hi lo + lo hi are laid ALONG K of one 128-byte swizzled row -- A row = [x_hi | x_hi | x_lo | 0], B row (cluster k,
// column c of P_k) = [P_hi | P_lo | P_hi | 0] -- and a round of up to 8 clusters is N = 128: three `tcgen05.mma kind::f16` of
// M = 128, N = 16 n, K = 16 per (tile, round).  Epilogue as at D = 64: thread = step = TMEM lane, t = sum_c Y_c (x'_c - 2 mu'_c) +
// mu'^T P mu' with the step's x' (16 values, hi + lo read back from the A image once per tile) in registers; warps 0-3 take the
// round's first four clusters, warps 4-7 the other four; the race key follows in the same thread.  The B image of a cluster is built
// from its FP32 P (16 x 16) every time it is used: thread = (cluster of the round, row), the operand scale from the row maxima of the
// diagonal by a half-warp shuffle.  One CTA of 8 warps per chain, three per SM (56 KB of shared memory, 128 TMEM columns each).
#include "npb_alg2.cuh"
#include "npb_tc_common.cuh"

npb_status npb_launch_a2_tc16(npb_chains *ch, const A2Args &a);

namespace {

constexpr int TD = 16;    // dimension
constexpr int TM = 128;   // steps per tile = UMMA M
constexpr int RC = 8;     // clusters per round (N = 16 RC)
constexpr uint32_t S_A = 0;                       // A image: 128 rows x 128 bytes
constexpr uint32_t S_B = 16384;                   // B image of a round: RC x 16 rows x 128 bytes
constexpr uint32_t S_KT = S_B + RC * 16 * 128;    // [TM][33] race keys
constexpr uint32_t S_MU = S_KT + TM * 33 * 4;     // [32][16] centred means
constexpr uint32_t S_M2 = S_MU + 32 * TD * 4;     // [32][16] -2 2^ex mu'
constexpr uint32_t S_MISC = S_M2 + 32 * TD * 4;

struct TcMisc {
	double xd[TD];       // the moving item in FP64
	unsigned long long bar;
	float ldv[32];
	int cnt[32];
	int zold[TM], items[TM], win[TM];
	float mk[RC];        // mu'^T P mu' of the round's clusters
	int ep[RC];          // their operand scale exponents
	float red2[4];
	float xm[TD];        // the moving item, centred
	float dm[2 * TD], pu[2 * TD];
	float xbar[TD];
	uint32_t tmem;
};
constexpr uint32_t TC_SMEM = S_MISC + sizeof(TcMisc) + 1024; // + alignment slack

struct TcArgs {
	A2Args a;
	const double *xbar; // [2 D + 1] column means of the dataset, exponent of the A scale, column maxima (npb_dataset::Xbar)
};

__device__ __forceinline__ float tc_key(const A2Args &a, const TcMisc *m, float t, int j, int k, uint32_t step, uint32_t ka, uint32_t kb) {
	const int n = m->cnt[k];
	const bool own = k == m->zold[j];
	const int n_eff = n - (own ? 1 : 0);
	if (n_eff <= 0) return -INFINITY;
	float q_eff = t, ld_eff = m->ldv[k];
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

// v0, v1 -> packed FP16 pairs (hi, lo) with v = hi + lo
__device__ __forceinline__ void tc_split2(float v0, float v1, uint32_t &hi, uint32_t &lo) {
	const __half2 h = __floats2half2_rn(v0, v1);
	const float2 f = __half22float2(h);
	const __half2 l = __floats2half2_rn(v0 - f.x, v1 - f.y);
	hi = *reinterpret_cast<const uint32_t *>(&h);
	lo = *reinterpret_cast<const uint32_t *>(&l);
}

// The race keys of the clusters in `mask` for the tile's steps [j_lo, T), up to RC clusters per round.
__device__ __forceinline__ void tc_pass(const A2Args &a, uint8_t *gen, TcMisc *m, const float *Pc, unsigned mask, int j_lo, int T, uint32_t s0,
		uint32_t ka, uint32_t kb, float sx_inv, uint32_t &phase) {
	if (!mask) return;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int wq = warp & 3, grp = warp >> 2, j = wq * 32 + lane;
	float *ktab = reinterpret_cast<float *>(gen + S_KT);
	const float *mus = reinterpret_cast<const float *>(gen + S_MU), *mus2 = reinterpret_cast<const float *>(gen + S_M2);
	// the step's own x' 2^ex, read back from its row of the A image (hi: chunks 0, 1; lo: chunks 4, 5)
	float xs[TD];
	{
		const uint8_t *row = gen + S_A + j * 128;
		const uint32_t sw = (uint32_t)j & 7u;
#pragma unroll
		for (int h = 0; h < 2; ++h) {
			const uint4 h4 = *reinterpret_cast<const uint4 *>(row + (((0u + h) ^ sw) << 4)), l4 = *reinterpret_cast<const uint4 *>(row + (((4u + h) ^ sw) << 4));
			const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
			for (int e = 0; e < 4; ++e) {
				xs[8 * h + 2 * e] = __half2float(__ushort_as_half((unsigned short)(hw[e] & 0xffffu))) + __half2float(__ushort_as_half((unsigned short)(lw[e] & 0xffffu)));
				xs[8 * h + 2 * e + 1] = __half2float(__ushort_as_half((unsigned short)(hw[e] >> 16))) + __half2float(__ushort_as_half((unsigned short)(lw[e] >> 16)));
			}
		}
	}
	while (mask) {
		// ---- the round's clusters: thread = (cluster kl of the round, row r of its P) ----
		const int nc = min(RC, __popc(mask));
		const int kl = tid >> 4, r = tid & 15;
		int kmine = -1;
		{
			unsigned mm = mask;
			for (int i = 0; i < RC; ++i) {
				const int k = mm ? __ffs(mm) - 1 : -1;
				mm &= mm - 1;
				if (i == kl) kmine = k;
			}
		}
		if (kl < RC && kmine >= 0) {
			const float4 *src = reinterpret_cast<const float4 *>(Pc + (size_t)kmine * TD * TD + r * TD);
			float v[TD];
#pragma unroll
			for (int i = 0; i < 4; ++i) { const float4 p = src[i]; v[4 * i] = p.x; v[4 * i + 1] = p.y; v[4 * i + 2] = p.z; v[4 * i + 3] = p.w; }
			float dmax = 0.0f; // P is positive definite: its largest magnitude sits on the diagonal
#pragma unroll
			for (int i = 0; i < TD; ++i) dmax = i == r ? v[i] : dmax;
			float mp = 0.0f;
#pragma unroll
			for (int i = 0; i < TD; ++i) mp = fmaf(v[i], mus[kmine * TD + i], mp);
			mp *= mus[kmine * TD + r];
			const unsigned hmask = 0xffffu << (tid & 16); // the 16 lanes of this cluster: the warp's other half may have no cluster this round
#pragma unroll
			for (int o = 8; o > 0; o >>= 1) {
				dmax = fmaxf(dmax, __shfl_xor_sync(hmask, dmax, o));
				mp += __shfl_xor_sync(hmask, mp, o);
			}
			const int ep = g_scale_exp(dmax);
			const float sp = ldexpf(1.0f, ep);
			uint4 hi0, lo0, hi1, lo1;
			tc_split2(v[0] * sp, v[1] * sp, hi0.x, lo0.x); tc_split2(v[2] * sp, v[3] * sp, hi0.y, lo0.y);
			tc_split2(v[4] * sp, v[5] * sp, hi0.z, lo0.z); tc_split2(v[6] * sp, v[7] * sp, hi0.w, lo0.w);
			tc_split2(v[8] * sp, v[9] * sp, hi1.x, lo1.x); tc_split2(v[10] * sp, v[11] * sp, hi1.y, lo1.y);
			tc_split2(v[12] * sp, v[13] * sp, hi1.z, lo1.z); tc_split2(v[14] * sp, v[15] * sp, hi1.w, lo1.w);
			uint8_t *row = gen + S_B + (kl * 16 + r) * 128; // B row n = 16 kl + r: [P_hi | P_lo | P_hi | 0]
			const uint32_t sw = (uint32_t)r & 7u;
			*reinterpret_cast<uint4 *>(row + ((0u ^ sw) << 4)) = hi0; *reinterpret_cast<uint4 *>(row + ((1u ^ sw) << 4)) = hi1;
			*reinterpret_cast<uint4 *>(row + ((2u ^ sw) << 4)) = lo0; *reinterpret_cast<uint4 *>(row + ((3u ^ sw) << 4)) = lo1;
			*reinterpret_cast<uint4 *>(row + ((4u ^ sw) << 4)) = hi0; *reinterpret_cast<uint4 *>(row + ((5u ^ sw) << 4)) = hi1;
			*reinterpret_cast<uint4 *>(row + ((6u ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
			*reinterpret_cast<uint4 *>(row + ((7u ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
			if (r == 0) { m->mk[kl] = mp; m->ep[kl] = ep; }
		}
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // written through the generic proxy, read by the MMA through the async one
		g_tc_fence_before();
		__syncthreads(); // the round's image is complete; the accumulator has been read (previous round)
		if (tid == 0) {
			g_tc_fence_after();
			const uint32_t base = g_smem_u32(gen);
			const uint32_t ID = g_idesc(TM, 16 * nc);
#pragma unroll
			for (int ks = 0; ks < 3; ++ks) // the fourth quarter of the rows is zero
				g_mma_f16(m->tmem, g_desc(base + S_A + ks * 32), g_desc(base + S_B + ks * 32), ID, ks != 0);
			g_tc_commit(g_smem_u32(&m->bar));
		}
		// ---- epilogue: warps 0-3 the round's clusters 0-3, warps 4-7 clusters 4-7; thread = step = TMEM lane ----
		int kq[4];
		{
			unsigned mm = mask;
#pragma unroll
			for (int i = 0; i < RC; ++i) {
				const int k = mm ? __ffs(mm) - 1 : -1;
				mm &= mm - 1;
				if ((i >> 2) == grp) kq[i & 3] = k;
			}
			mask = mm;
		}
		g_mbar_wait(g_smem_u32(&m->bar), phase);
		phase ^= 1u;
		g_tc_fence_after();
		if (kq[0] >= 0) {
			float v[32], w[32];
			g_tmem_ld32(m->tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(64 * grp), v);
			if (kq[2] >= 0) g_tmem_ld32(m->tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(64 * grp + 32), w);
#pragma unroll
			for (int cl = 0; cl < 4; ++cl) {
				const int k = kq[cl];
				if (k < 0) continue;
				const float *m2 = mus2 + k * TD;
				float part = 0.0f;
#pragma unroll
				for (int c = 0; c < TD; ++c) part = fmaf(cl < 2 ? v[16 * (cl & 1) + c] : w[16 * (cl & 1) + c], xs[c] + m2[c], part);
				const float descale = ldexpf(sx_inv * sx_inv, -m->ep[4 * grp + cl]); // Y carries 2^(ex + ep), the item and -2 mu' another 2^ex
				const float t = fmaxf(fmaf(descale, part, m->mk[4 * grp + cl]), 0.0f);
				if (j >= j_lo && j < T) ktab[j * 33 + k] = tc_key(a, m, t, j, k, s0 + (uint32_t)j, ka, kb);
			}
		}
		g_tc_fence_before();
		__syncthreads(); // mk / ep / the image are rewritten by the next round; after the last one the keys are complete
	}
}

__global__ void __launch_bounds__(256, 2) k_a2_tc16(const TcArgs g) {
	extern __shared__ uint8_t tc_raw[];
	uint8_t *gen = tc_raw + ((1024u - (g_smem_u32(tc_raw) & 1023u)) & 1023u); // 1024-aligned, and still known to be shared memory
	const A2Args &a = g.a;
	TcMisc *m = reinterpret_cast<TcMisc *>(gen + S_MISC);
	float *ktab = reinterpret_cast<float *>(gen + S_KT), *mus = reinterpret_cast<float *>(gen + S_MU);
	const int chain = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int C = a.C, N = a.N;
	float *Pc = a.P + (size_t)chain * 32 * TD * TD;
	double *sxc = a.sx + (size_t)chain * 32 * TD, *sxxc = a.sxx + (size_t)chain * 32 * TD * TD;

	if (tid < TD) m->xbar[tid] = (float)g.xbar[tid];
	if (tid < 32) { m->cnt[tid] = a.counts[(size_t)chain * 32 + tid]; m->ldv[tid] = a.ld[(size_t)chain * 32 + tid]; }
	if (tid == 0) {
		g_mbar_init(g_smem_u32(&m->bar), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async;" ::: "memory");
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(g_smem_u32(&m->tmem)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	g_tc_fence_before();
	__syncthreads();
	g_tc_fence_after();
	const float sx2 = -2.0f * ldexpf(1.0f, (int)g.xbar[TD]);
	float *mus2 = reinterpret_cast<float *>(gen + S_M2);
	for (int e = tid; e < 32 * TD; e += 256) {
		const float v = a.mu[(size_t)chain * 32 * TD + e] - m->xbar[e & (TD - 1)];
		mus[e] = v;
		mus2[e] = sx2 * v;
	}
	const int ex = (int)g.xbar[TD];
	const float sx = ldexpf(1.0f, ex), sx_inv = ldexpf(1.0f, -ex);
	int kocc = a.kocc[chain];
	unsigned long long st_cand = 0ull, st_moved = 0ull, st_births = 0ull; // thread 0's are the ones written back
	const uint32_t ka = (uint32_t)a.seed ^ 0xA2A2A2A2u, k1 = (uint32_t)(a.seed >> 32) + (uint32_t)chain;
	const int tile_max = a.tile < 1 ? 1 : (a.tile > TM ? TM : a.tile);
	int tile = tile_max;
	uint32_t phase = 0u;
	__syncthreads();

	for (int sw = 0; sw < a.n_sweeps; ++sw) {
		const int32_t *order = a.order + (size_t)sw * N;
		const uint32_t kb = k1 ^ ((a.sweep0 + (uint32_t)sw) * 0x9E3779B9u);
		for (int s = 0; s < N;) {
			const int T = min(tile, N - s);
			if (tid < T) {
				const int it = order[s + tid];
				m->items[tid] = it;
				m->zold[tid] = (int)a.z[(size_t)it * C + chain];
			}
			__syncthreads();
			// ---- A image: row j = [x_hi | x_hi | x_lo | 0] (16 FP16 each, K-major, 128-byte swizzle): thread = (step, half of the coordinates) ----
			{
				const int j = tid >> 1, h = tid & 1;
				uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
				if (j < T) {
					const float4 *src = reinterpret_cast<const float4 *>(a.X + (size_t)m->items[j] * TD + 8 * h);
					const float4 p0 = __ldg(src), p1 = __ldg(src + 1);
					const float *xb = m->xbar + 8 * h;
					tc_split2((p0.x - xb[0]) * sx, (p0.y - xb[1]) * sx, hi.x, lo.x);
					tc_split2((p0.z - xb[2]) * sx, (p0.w - xb[3]) * sx, hi.y, lo.y);
					tc_split2((p1.x - xb[4]) * sx, (p1.y - xb[5]) * sx, hi.z, lo.z);
					tc_split2((p1.z - xb[6]) * sx, (p1.w - xb[7]) * sx, hi.w, lo.w);
				}
				uint8_t *row = gen + S_A + j * 128;
				const uint32_t sw = (uint32_t)j & 7u;
				*reinterpret_cast<uint4 *>(row + (((0u + h) ^ sw) << 4)) = hi;
				*reinterpret_cast<uint4 *>(row + (((2u + h) ^ sw) << 4)) = hi;
				*reinterpret_cast<uint4 *>(row + (((4u + h) ^ sw) << 4)) = lo;
				*reinterpret_cast<uint4 *>(row + (((6u + h) ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
			}
			unsigned occ = 0u;
#pragma unroll
			for (int k = 0; k < 32; ++k) occ |= m->cnt[k] > 0 ? 1u << k : 0u;
			// the candidate "a new cluster" (prior predictive, tabulated per item)
			if (tid < T) ktab[tid * 33 + 32] = a.log2_alpha + __ldg(a.lp0 + m->items[tid]) * NPB_LOG2E + a2_noise(ka ^ (uint32_t)(s + tid), kb, 32u);
			__syncthreads();
			tc_pass(a, gen, m, Pc, occ, 0, T, (uint32_t)s, ka, kb, sx_inv, phase);
			__syncthreads();
			int j0 = 0, tile_moves = 0;
			while (j0 < T) {
				// ---- winners of the steps not yet final ----
				for (int j = j0 + warp; j < T; j += 8) {
					const float key = m->cnt[lane] - (lane == m->zold[j] ? 1 : 0) > 0 ? ktab[j * 33 + lane] : -INFINITY;
					const float top = fmaxf(redux_max_f32(key), ktab[j * 33 + 32]);
					const unsigned bal = __ballot_sync(0xffffffffu, key == top && key > -INFINITY);
					int w = bal ? __ffs(bal) - 1 : 32;
					if (w == 32) { // a new cluster needs a slot without members once the item is retracted; none: the item stays (code 33)
						const unsigned fb = __ballot_sync(0xffffffffu, m->cnt[lane] - (lane == m->zold[j] ? 1 : 0) <= 0);
						if (!fb) w = 33;
					}
					if (lane == 0) m->win[j] = w;
				}
				__syncthreads();
				// ---- the first step that does not simply stay ----
				int jm = T;
#pragma unroll
				for (int q = 3; q >= 0; --q) {
					const int jq = j0 + lane + 32 * q;
					const unsigned ev = __ballot_sync(0xffffffffu, jq < T && m->win[jq] != m->zold[jq]);
					if (ev) jm = j0 + 32 * q + __ffs(ev) - 1;
				}
				if (warp == 0) { // candidates weighed by the steps now final (the event step included)
					int cs = 0;
#pragma unroll
					for (int q = 0; q < 4; ++q) {
						const int jq = j0 + lane + 32 * q;
						if (jq <= jm && jq < T) cs += kocc - (m->cnt[m->zold[jq]] == 1 ? 1 : 0) + 1;
					}
#pragma unroll
					for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
					st_cand += (unsigned long long)cs;
				}
				if (jm >= T) break;
				const int w = m->win[jm], src = m->zold[jm], item = m->items[jm];
				j0 = jm + 1;
				if (w == 33) { // no room for a new cluster: the item stays, the chain is reported
					if (tid == 0) a.overflow[chain] = 1;
					__syncthreads();
					continue;
				}
				// ================= the move (FP32, as in npb_alg2_tile.cu): src loses the item, dst gains it =================
				const bool born = w == 32;
				int dst = w;
				if (born) {
					dst = 0;
					while (m->cnt[dst] - (dst == src ? 1 : 0) > 0) ++dst;
				}
				const int n_src = m->cnt[src], n_eff = n_src - 1;
				const bool died = n_eff == 0;
				const int n_dst = born ? 0 : m->cnt[dst];
				if (tid < TD) {
					const float x = __ldg(a.X + (size_t)item * TD + tid) - m->xbar[tid];
					m->xm[tid] = x;
					m->dm[tid] = x - mus[src * TD + tid];
				} else if (tid < 2 * TD) m->dm[tid] = (__ldg(a.X + (size_t)item * TD + tid - TD) - m->xbar[tid - TD]) - (born ? a.mu0[tid - TD] - m->xbar[tid - TD] : mus[dst * TD + tid - TD]);
				else if (tid < 3 * TD) m->xd[tid - 2 * TD] = a.X64[(size_t)item * TD + tid - 2 * TD];
				if (born) { // the new cluster starts from the prior
					for (int e = tid; e < TD * TD; e += 256) { Pc[(size_t)dst * TD * TD + e] = __ldg(a.P0 + e); sxxc[(size_t)dst * TD * TD + e] = 0.0; }
					if (tid < TD) sxc[dst * TD + tid] = 0.0;
				}
				__syncthreads();
				float prod = 0.0f;
				if (tid < 2 * TD) {
					const int which = tid / TD, r = tid % TD;
					if (which == 1 || !died) {
						const float *Pk = Pc + (size_t)(which ? dst : src) * TD * TD + r;
						const float *dv = m->dm + which * TD;
						float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll 4
						for (int c = 0; c < TD; c += 4) {
							a0 = fmaf(Pk[(c) * TD], dv[c], a0); a1 = fmaf(Pk[(c + 1) * TD], dv[c + 1], a1);
							a2 = fmaf(Pk[(c + 2) * TD], dv[c + 2], a2); a3 = fmaf(Pk[(c + 3) * TD], dv[c + 3], a3);
						}
						const float v = (a0 + a1) + (a2 + a3);
						m->pu[tid] = v;
						prod = v * dv[r];
					}
				}
				if (warp == 0) { // lanes 0-15 the old cluster, 16-31 the new one
#pragma unroll
					for (int o = 8; o > 0; o >>= 1) prod += __shfl_xor_sync(0xffffffffu, prod, o);
					if ((lane & 15) == 0) m->red2[lane >> 4] = prod;
				}
				__syncthreads();
				const float t_s = m->red2[0], t_d = m->red2[1];
				const float kp = a.kappa0 + (float)n_src, km = kp - 1.0f;
				const float cdown = kp / km, one_m = fmaxf(1.0f - cdown * t_s, 1e-12f), f_s = cdown / one_m;
				const float kap = a.kappa0 + (float)n_dst, kap1 = kap + 1.0f;
				const float cc = kap / kap1, den = 1.0f + cc * t_d, f_d = cc / den;
				for (int e = tid; e < TD * TD; e += 256) {
					const int r = e / TD, c = e % TD;
					if (!died) {
						float *p = Pc + (size_t)src * TD * TD + e;
						*p = fmaf(f_s, m->pu[r] * m->pu[c], *p);
					}
					float *p2 = Pc + (size_t)dst * TD * TD + e;
					*p2 = fmaf(-f_d, m->pu[TD + r] * m->pu[TD + c], *p2);
					const double xx = m->xd[r] * m->xd[c];
					if (!(died && born && dst == src)) sxxc[(size_t)src * TD * TD + e] -= xx;
					sxxc[(size_t)dst * TD * TD + e] += xx;
				}
				if (tid < TD) {
					const float x = m->xm[tid];
					if (!died) {
						const float v = (kp * mus[src * TD + tid] - x) / km;
						mus[src * TD + tid] = v;
						mus2[src * TD + tid] = sx2 * v;
					}
					if (!(died && born && dst == src)) sxc[src * TD + tid] -= m->xd[tid];
				}
				__syncthreads();
				if (tid < TD) {
					const float x = m->xm[tid];
					const float m0 = born ? a.mu0[tid] - m->xbar[tid] : mus[dst * TD + tid];
					const float v = (kap * m0 + x) / kap1;
					mus[dst * TD + tid] = v;
					mus2[dst * TD + tid] = sx2 * v;
					sxc[dst * TD + tid] += m->xd[tid];
				}
				if (tid == 0) {
					if (!died) m->ldv[src] += __logf(one_m);
					m->cnt[src] = n_eff;
					m->ldv[dst] = (born ? a.ld0 : m->ldv[dst]) + __logf(den);
					m->cnt[dst] = n_dst + 1;
					a.z[(size_t)item * C + chain] = (npb_z_t)dst;
					st_moved++;
					if (born) st_births++;
				}
				kocc += (born ? 1 : 0) - (died ? 1 : 0);
				++tile_moves;
				__syncthreads();
				if (j0 < T) { // the two changed clusters again, for the steps behind the move
					tc_pass(a, gen, m, Pc, (died ? 0u : 1u << src) | 1u << dst, j0, T, (uint32_t)s, ka, kb, sx_inv, phase);
					__syncthreads();
				}
			}
			__syncthreads();
			s += T;
			// a move costs two more cluster evaluations of the tile: shorter tiles only where that would no longer pay
			if (tile_moves * 4 > T) tile = max(tile / 2, min(tile_max, 16));
			else if (tile_moves * 16 <= T) tile = min(tile * 2, tile_max);
		}
	}
	__syncthreads();
	for (int e = tid; e < 32 * TD; e += 256) a.mu[(size_t)chain * 32 * TD + e] = mus[e] + m->xbar[e & (TD - 1)];
	if (tid < 32) {
		a.counts[(size_t)chain * 32 + tid] = m->cnt[tid];
		a.ld[(size_t)chain * 32 + tid] = m->ldv[tid];
		const int o = __popc(__ballot_sync(0xffffffffu, m->cnt[tid] > 0));
		if (tid == 0) {
			a.kocc[chain] = o;
			a.st[(size_t)chain * 4 + 0] += st_cand;
			a.st[(size_t)chain * 4 + 1] += st_moved;
			a.st[(size_t)chain * 4 + 2] += st_births;
		}
	}
	g_tc_fence_before();
	__syncthreads();
	if (warp == 0) {
		g_tc_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(m->tmem) : "memory");
	}
}

__global__ void k_a2_colmean16(const double *X, int64_t N, int D, double *out) { // mean and largest centred magnitude of column blockIdx.x
	__shared__ double red[256];
	const int c = blockIdx.x;
	double s = 0.0;
	for (int64_t i = threadIdx.x; i < N; i += 256) s += X[i * D + c];
	red[threadIdx.x] = s;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
		__syncthreads();
	}
	const double mean = red[0] / (double)N;
	__syncthreads();
	double mx = 0.0;
	for (int64_t i = threadIdx.x; i < N; i += 256) mx = fmax(mx, fabs(X[i * D + c] - mean));
	red[threadIdx.x] = mx;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
		__syncthreads();
	}
	if (threadIdx.x == 0) { out[c] = mean; out[D + 1 + c] = red[0]; }
}
__global__ void k_a2_xscale16(double *out, int D) {
	double mx = 0.0;
	for (int c = 0; c < D; ++c) mx = fmax(mx, out[D + 1 + c]);
	out[D] = (double)g_scale_exp((float)mx);
}

} // namespace

npb_status npb_launch_a2_tc16(npb_chains *ch, const A2Args &a) {
	npb_ctx *ctx = ch->ctx;
	npb_dataset *ds = ch->ds;
	if (ch->D != TD) return npb_fail(ctx, NPB_E_UNSUPPORTED, "k_a2_tc16 covers D = 16");
	if (!ds->Xbar) NPB_CUDA_OK(cudaMalloc((void **)&ds->Xbar, sizeof(double) * (2 * TD + 1)));
	if (!ds->xbar_valid) { // the same contents the D = 16 Algorithm 8 tensor paths keep there
		ds->xbar_valid = true;
		k_a2_colmean16<<<TD, 256, 0, ctx->stream>>>(ds->X64, ds->N, TD, ds->Xbar);
		NPB_CUDA_OK(cudaGetLastError());
		k_a2_xscale16<<<1, 1, 0, ctx->stream>>>(ds->Xbar, TD);
		NPB_CUDA_OK(cudaGetLastError());
	}
	if (!ctx->a2_tc16_attr_set) {
		NPB_CUDA_OK(cudaFuncSetAttribute(k_a2_tc16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
		ctx->a2_tc16_attr_set = true;
	}
	TcArgs g;
	g.a = a;
	g.xbar = ds->Xbar;
	k_a2_tc16<<<(unsigned)ch->C, 256, TC_SMEM, ctx->stream>>>(g);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}
