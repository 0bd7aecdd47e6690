// npb_alg2_tc.cu -- CONJUGATE Algorithm 2 at D = 64 (BASELINE configs[3]: "tensor-core whitened quadratic forms"): the tile
// schedule of npb_alg2_tile.cu -- up to 128 steps evaluated ahead of the chain, the two changed columns corrected after every
// move -- with the quadratic forms on tcgen05.
//
// For a tile of 128 steps and one cluster k:  Y = X' P_k  (128 x 64 x 64) is ONE accumulator of 64 TMEM columns, X' the items
// centred on the data mean.  FP32 accuracy from FP16 operands as in npb_alg8_gemm.cu: both operands are scaled by a power of two to
// just below 2^14 and split into two FP16 terms, three products (hi hi + hi lo + lo hi) = 12 `tcgen05.mma kind::f16` of
// M = 128, N = 64, K = 16 per (tile, cluster).  P is symmetric, so with mu' the centred posterior mean
//     t = (x' - mu')^T P (x' - mu') = sum_c Y_c (x'_c - 2 mu'_c) + mu'^T P mu'
// and the epilogue is 64 FMAs per (step, cluster) -- thread = step = TMEM lane, the accumulator's 64 columns split over two warps --
// instead of the 4096 of the FP32 product; the race key follows in the same thread.  A images are built once per tile (the items
// gathered in scan order, centred, scaled, split, stored K-major with the 128-byte swizzle), the B image of a cluster from its FP32
// P in global memory every time it is used -- so there is no second copy of the state to keep consistent: after a move the two
// changed clusters are simply evaluated again.  Same schedule-independence as the FP32 tile kernel: a row of the product depends
// only on its own item and the cluster's state, so a.tile = 1 ... 128 give the same chain bit for bit.
//
// One CTA of 8 warps per chain, two CTAs per SM (96 KB of shared memory, 128 TMEM columns = two accumulators, <= 128 registers).
// Per cluster: all threads build the B image (two buffers: cluster k + 1 is built and issued while k's MMAs run); one thread issues
// the 12 MMAs and commits to an mbarrier; all wait; all read TMEM and their own row of the A image.
#include "npb_alg2.cuh"
#include "npb_tc_common.cuh"

npb_status npb_launch_a2_tc(npb_chains *ch, const A2Args &a);

namespace {

constexpr int TD = 64;    // dimension
constexpr int TM = 128;   // steps per tile = UMMA M
constexpr uint32_t S_AHI = 0, S_ALO = 16384;            // A image of the tile: 128 rows x 128 bytes, hi then lo
constexpr uint32_t S_B = 32768, S_BSZ = 16384;          // two B images: hi of cluster 0, hi of cluster 1 (8 KB each: 128 contiguous rows = ONE N = 128 operand), then the two lo
constexpr uint32_t S_KT = S_B + 2 * S_BSZ;              // [TM][33] race keys
constexpr uint32_t S_MU = S_KT + TM * 33 * 4;           // [32][64] centred means
constexpr uint32_t S_M2 = S_MU + 32 * TD * 4;           // [32][64] -2 2^ex mu': what the epilogue adds to the scaled item
constexpr uint32_t S_MISC = S_M2 + 32 * TD * 4;

struct TcMisc {
	double xd[TD];       // the moving item in FP64
	unsigned long long bar;
	float ldv[32];
	int cnt[32];
	int zold[TM], items[TM], win[TM];
	float red[2][8];     // shares of mu'^T P mu' per warp, per B buffer
	float red2[4];
	float xm[TD];        // the moving item, centred
	float dm[2 * TD], pu[2 * TD];
	float xbar[TD];
	float pmax[32];      // largest diagonal element of every cluster's P (the scale of its B operand)
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

// v0, v1 -> packed FP16 pairs (hi, lo) with v = hi + lo: the same roundings as g_split, two values per conversion instruction
__device__ __forceinline__ void tc_split2(float v0, float v1, uint32_t &hi, uint32_t &lo) {
	const __half2 h = __floats2half2_rn(v0, v1);
	const float2 f = __half22float2(h);
	const __half2 l = __floats2half2_rn(v0 - f.x, v1 - f.y);
	hi = *reinterpret_cast<const uint32_t *>(&h);
	lo = *reinterpret_cast<const uint32_t *>(&l);
}
// this thread's 16 values of P_k, coalesced: float4 number i 256 + tid of the matrix, i = 0..3 = row 16 i + tid / 16, columns 4 (tid % 16) ...
__device__ __forceinline__ void tc_load(const float *Pc, int k, float4 (&pf)[4]) {
	const float4 *src = reinterpret_cast<const float4 *>(Pc + (size_t)k * TD * TD) + threadIdx.x;
#pragma unroll
	for (int i = 0; i < 4; ++i) pf[i] = src[i * 256];
}
// B image of cluster k (from pf) into buffer b, and this warp's share of mu'^T P mu'
__device__ __forceinline__ void tc_build(uint8_t *gen, TcMisc *m, int k, int b, const float4 (&pf)[4]) {
	const int tid = threadIdx.x, r0 = tid >> 4, c0 = 4 * (tid & 15);
	const float *mus = reinterpret_cast<const float *>(gen + S_MU) + k * TD;
	const float sp = ldexpf(1.0f, g_scale_exp(m->pmax[k])); // P is positive definite: its largest magnitude sits on the diagonal
	const float4 mc = *reinterpret_cast<const float4 *>(mus + c0);
	uint8_t *B = gen + S_B + b * 8192; // hi; lo 16 KB further
	float mp = 0.0f;
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const int r = 16 * i + r0;
		const float v[4] = {pf[i].x, pf[i].y, pf[i].z, pf[i].w};
		mp = fmaf(mus[r], fmaf(v[0], mc.x, fmaf(v[1], mc.y, fmaf(v[2], mc.z, v[3] * mc.w))), mp);
		uint2 hi, lo;
		tc_split2(v[0] * sp, v[1] * sp, hi.x, lo.x);
		tc_split2(v[2] * sp, v[3] * sp, hi.y, lo.y);
		*reinterpret_cast<uint2 *>(B + g_sw128(r, c0)) = hi;
		*reinterpret_cast<uint2 *>(B + 16384 + g_sw128(r, c0)) = lo;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) mp += __shfl_xor_sync(0xffffffffu, mp, o);
	if ((tid & 31) == 0) m->red[b][tid >> 5] = mp;
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // written through the generic proxy, read by the MMA through the async one
}
// the 12 MMAs of a round: both clusters at once as N = 128 (accumulator columns 0-63 and 64-127), or the first alone as N = 64;
// completion arrives on the mbarrier
__device__ __forceinline__ void tc_issue(uint8_t *gen, TcMisc *m, bool two) {
	g_tc_fence_after();
	const uint32_t base = g_smem_u32(gen);
	const uint32_t ID = two ? g_idesc(TM, 2 * TD) : g_idesc(TM, TD);
#pragma unroll
	for (int prod = 0; prod < 3; ++prod) {
		const uint32_t A = base + (prod == 2 ? S_ALO : S_AHI), B = base + S_B + (prod == 1 ? 16384u : 0u);
#pragma unroll
		for (int ks = 0; ks < 4; ++ks) g_mma_f16(m->tmem, g_desc(A + ks * 32), g_desc(B + ks * 32), ID, (prod | ks) != 0);
	}
	g_tc_commit(g_smem_u32(&m->bar));
}
// The race keys of the clusters in `mask` for the tile's steps [j_lo, T), two clusters per round: all threads build both B images,
// one thread issues the round's 12 MMAs (N = 128: both clusters at once), then warps 0-3 take the first cluster and warps
// 4-7 the second -- thread = step = TMEM lane, all 64 columns of its accumulator against the step's own row of the A image
// (x' 2^ex = hi + lo: the very operand the MMA saw) and the cluster's -2 2^ex mu' -- and the race key follows in the same thread.
// One CTA barrier per round; the first cluster of the next round is fetched while this round's accumulators are read.
__device__ __forceinline__ void tc_pass(const A2Args &a, uint8_t *gen, TcMisc *m, const float *Pc, unsigned mask, int j_lo, int T, uint32_t s0,
		uint32_t ka, uint32_t kb, float sx_inv, uint32_t &phase) {
	if (!mask) return;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int wq = warp & 3, grp = warp >> 2, j = wq * 32 + lane;
	float *ktab = reinterpret_cast<float *>(gen + S_KT);
	float4 pf[4];
	tc_load(Pc, __ffs(mask) - 1, pf);
	float pend_t = 0.0f; // a round's race keys are computed while the next round's MMAs run
	int pend_k = -1;
	while (mask) {
		const int kA = __ffs(mask) - 1;
		mask &= mask - 1;
		const int kB = mask ? __ffs(mask) - 1 : -1;
		mask &= mask - 1;
		{
			float4 pg[4];
			if (kB >= 0) tc_load(Pc, kB, pg); // in flight while the first image is built
			tc_build(gen, m, kA, 0, pf);
			if (kB >= 0) tc_build(gen, m, kB, 1, pg);
		}
		if (mask) tc_load(Pc, __ffs(mask) - 1, pf);
		g_tc_fence_before();
		__syncthreads(); // both images are complete; both accumulators have been read (previous round)
		if (tid == 0) tc_issue(gen, m, kB >= 0);
		if (pend_k >= 0 && j >= j_lo && j < T) ktab[j * 33 + pend_k] = tc_key(a, m, pend_t, j, pend_k, s0 + (uint32_t)j, ka, kb);
		pend_k = -1;
		const int k = grp ? kB : kA;
		if (k >= 0) {
			float mk = 0.0f;
#pragma unroll
			for (int w = 0; w < 8; ++w) mk += m->red[grp][w];
			const float descale = ldexpf(sx_inv * sx_inv, -g_scale_exp(m->pmax[k])); // Y carries 2^(ex + ep), the item and -2 mu' another 2^ex
			const float *m2 = reinterpret_cast<const float *>(gen + S_M2) + k * TD;
			g_mbar_wait(g_smem_u32(&m->bar), phase);
			g_tc_fence_after();
			float part = 0.0f;
#pragma unroll
			for (int half = 0; half < 2; ++half) {
				float v[32];
				g_tmem_ld32(m->tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(64 * grp + 32 * half), v);
#pragma unroll
				for (int c8 = 0; c8 < 4; ++c8) { // 8 coordinates per 16-byte chunk of the row, chunks swizzled by the row
					const uint32_t off = (uint32_t)j * 128u + ((((uint32_t)(4 * half + c8)) ^ ((uint32_t)j & 7u)) << 4);
					const uint4 h4 = *reinterpret_cast<const uint4 *>(gen + S_AHI + off), l4 = *reinterpret_cast<const uint4 *>(gen + S_ALO + off);
					const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
					const float4 ma = *reinterpret_cast<const float4 *>(m2 + 32 * half + 8 * c8), mb = *reinterpret_cast<const float4 *>(m2 + 32 * half + 8 * c8 + 4);
					const float mq[2][4] = {{ma.x, ma.y, ma.z, ma.w}, {mb.x, mb.y, mb.z, mb.w}};
#pragma unroll
					for (int e = 0; e < 4; ++e) {
						const float x0 = __half2float(__ushort_as_half((unsigned short)(hw[e] & 0xffffu))) + __half2float(__ushort_as_half((unsigned short)(lw[e] & 0xffffu)));
						const float x1 = __half2float(__ushort_as_half((unsigned short)(hw[e] >> 16))) + __half2float(__ushort_as_half((unsigned short)(lw[e] >> 16)));
						const int c = 8 * c8 + 2 * e;
						part = fmaf(v[c], x0 + mq[e >> 1][2 * (e & 1)], part);
						part = fmaf(v[c + 1], x1 + mq[e >> 1][2 * (e & 1) + 1], part);
					}
				}
			}
			pend_t = fmaxf(fmaf(descale, part, mk), 0.0f);
			pend_k = k;
		}
		// nobody builds into a buffer (next round) whose MMAs may still read it: everybody sees the round's MMAs complete
		if (k < 0) g_mbar_wait(g_smem_u32(&m->bar), phase);
		phase ^= 1u;
		g_tc_fence_before();
	}
	if (pend_k >= 0 && j >= j_lo && j < T) ktab[j * 33 + pend_k] = tc_key(a, m, pend_t, j, pend_k, s0 + (uint32_t)j, ka, kb);
	__syncthreads(); // the keys are complete; the accumulators are free
}

__global__ void __launch_bounds__(256, 2) k_a2_tc(const TcArgs g) {
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
	for (int k = warp; k < 32; k += 8) {
		const float *Pk = Pc + (size_t)k * TD * TD;
		const float d = redux_max_f32(fmaxf(Pk[lane * (TD + 1)], Pk[(lane + 32) * (TD + 1)]));
		if (lane == 0) m->pmax[k] = d;
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
			// ---- A image (FP16 hi / lo, K-major, swizzled) of the centred items: thread = (step, 8 coordinates) ----
			for (int e = tid; e < TM * 8; e += 256) {
				const int j = e >> 3, q = e & 7;
				uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
				if (j < T) {
					const float4 *src = reinterpret_cast<const float4 *>(a.X + (size_t)m->items[j] * TD + 8 * q);
					const float4 p0 = __ldg(src), p1 = __ldg(src + 1);
					const float *xb = m->xbar + 8 * q;
					tc_split2((p0.x - xb[0]) * sx, (p0.y - xb[1]) * sx, hi.x, lo.x);
					tc_split2((p0.z - xb[2]) * sx, (p0.w - xb[3]) * sx, hi.y, lo.y);
					tc_split2((p1.x - xb[4]) * sx, (p1.y - xb[5]) * sx, hi.z, lo.z);
					tc_split2((p1.z - xb[6]) * sx, (p1.w - xb[7]) * sx, hi.w, lo.w);
				}
				*reinterpret_cast<uint4 *>(gen + S_AHI + g_sw128(j, 8 * q)) = hi;
				*reinterpret_cast<uint4 *>(gen + S_ALO + g_sw128(j, 8 * q)) = lo;
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
				if (warp < 4) {
#pragma unroll
					for (int o = 16; o > 0; o >>= 1) prod += __shfl_xor_sync(0xffffffffu, prod, o);
					if (lane == 0) m->red2[warp] = prod;
				}
				__syncthreads();
				const float t_s = m->red2[0] + m->red2[1], t_d = m->red2[2] + m->red2[3];
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
				if (warp < 2) { // the two changed diagonals
					const float *Pk = Pc + (size_t)(warp ? dst : src) * TD * TD;
					const float d = redux_max_f32(fmaxf(Pk[lane * (TD + 1)], Pk[(lane + 32) * (TD + 1)]));
					if (lane == 0) m->pmax[warp ? dst : src] = d;
				}
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

__global__ void k_a2_colmean(const double *X, int64_t N, int D, double *out) { // as k_colmean of npb_alg8_gemm.cu: mean and largest centred magnitude of column blockIdx.x
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
__global__ void k_a2_xscale(double *out, int D) {
	double mx = 0.0;
	for (int c = 0; c < D; ++c) mx = fmax(mx, out[D + 1 + c]);
	out[D] = (double)g_scale_exp((float)mx);
}

} // namespace

npb_status npb_launch_a2_tc(npb_chains *ch, const A2Args &a) {
	npb_ctx *ctx = ch->ctx;
	npb_dataset *ds = ch->ds;
	if (ch->D != TD) return npb_fail(ctx, NPB_E_UNSUPPORTED, "the tensor-core conjugate Algorithm 2 kernel covers D = 64");
	if (!ds->Xbar) NPB_CUDA_OK(cudaMalloc((void **)&ds->Xbar, sizeof(double) * (2 * TD + 1)));
	if (!ds->xbar_valid) { // the same contents the D = 64 Algorithm 8 path keeps there
		ds->xbar_valid = true;
		k_a2_colmean<<<TD, 256, 0, ctx->stream>>>(ds->X64, ds->N, TD, ds->Xbar);
		NPB_CUDA_OK(cudaGetLastError());
		k_a2_xscale<<<1, 1, 0, ctx->stream>>>(ds->Xbar, TD);
		NPB_CUDA_OK(cudaGetLastError());
	}
	if (!ctx->a2_tc_attr_set) {
		NPB_CUDA_OK(cudaFuncSetAttribute(k_a2_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
		ctx->a2_tc_attr_set = true;
	}
	TcArgs g;
	g.a = a;
	g.xbar = ds->Xbar;
	k_a2_tc<<<(unsigned)ch->C, 256, TC_SMEM, ctx->stream>>>(g);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}
