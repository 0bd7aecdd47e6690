// npb_alg8_gemm.cu -- the tensor path: Algorithm 8 / Algorithm 2 sweeps at D = 64 and D = 16 (Kmax = 32) with the whitened
// quadratic forms of a block of steps as a tcgen05 GEMM (BASELINE configs[3]: "tensor-core whitened quadratic forms") and the
// sequential race as a warp-per-chain consumer.  D = 64 first (fused kernel); the D = 16 variant (the headline shape: table
// kernel + race kernel) is the second half of the file.
//
// Same algorithm and the same reference lines as the D <= 16 kernels (NealAlgorithm8::update,
// src/np_neal_algorithm8.cpp:49-167; density: src/statistics/multivariatenormal.cpp:106-136), a different mapping: at
// D = 64 a cluster slot is 2145 floats, so neither the registers of a lane (npb_alg8_tile4.cuh) nor shared memory hold a
// chain's slot table, and a density is 2080 FMAs -- GEMM-shaped work.  A sweep is cut into blocks of 4096 steps; per
// block, in stream order:
//   k_pre_aimg     gathers the block's item rows (scan order), centres them on the dataset mean, scales them by a power
//                  of two (largest |x - xbar| just below 2^14), splits every coordinate into two FP16 terms (hi + lo)
//                  and writes them as the shared-memory IMAGE of the A operand (K-major, 128-byte swizzle: one row =
//                  the 64 coordinates), one 32 KB image (hi, lo) per 128-step tile;
//   k_pre_bimg     does the same for the slots whose parameters changed (births; everything at the first block of a
//                  launch): the upper-triangular factor T2 (scaled per slot) as the B operand image, nb = -T2 (mu - xbar),
//                  c2 and the descaling factor aside;
//   k_density_tc   persistent, one CTA per SM, unit of work = (chain, 4 slots): the four slots' B images stay resident
//                  in shared memory (64 KB) while the block's A images stream through a four-stage ring of
//                  cp.async.bulk copies (no register staging); one thread issues tcgen05.mma kind::f16 128 x 256 x 16
//                  (three FP16 products per FP32 product: hi*hi + hi*lo + lo*hi, FP32 accumulation in TMEM, two
//                  accumulator buffers of 256 columns); four epilogue warps read the accumulators back (tcgen05.ld
//                  32x32b), form c2 - sum_j (s y_j + nb_j)^2 per (step, slot) and write the block's log2-density table
//                  L[chain][step][slot].  The split is exact to 2^-25 of the operand's scale: FP16 carries 11 significant
//                  bits like TF32 and, with the operands scaled to ~2^14, the low terms stay above the subnormal
//                  quantum 2^-24, so the error is that of the dropped lo*lo product and of FP32 accumulation (3xTF32
//                  accuracy at twice its MMA rate; the parity test holds it to 1e-5 of the oracle).  T2 is upper
//                  triangular, so the first half of K only feeds rows j < 32: the rows of B are ordered
//                  (j / 32, slot, j % 32) and the first two K-steps are issued with N = 128, a quarter of the MMA work
//                  saved;
//   g_consume_chain  (two more warps of the same kernel, working on the PREVIOUS block's table) one warp per chain: the
//                  exponential race of every step of the block from L, the auxiliary keys of k_aux_keys and the member
//                  counts -- the consumer of npb_alg8_tile4.cuh with the producer warp replaced by a table in L2.  A birth
//                  writes theta' to the slot table, re-evaluates the slot's column of L for the rest of the block on the
//                  CUDA cores and marks the slot for k_pre_bimg.
// NPB_D64_DENSITY=fp32 replaces k_density_tc by a plain FP32 kernel (A/B measurements, cross-check in the tests).
#include "npb_tc_common.cuh"


// ---------------------------------------------------------------------------------------------------------
// column means of the dataset (double), once per dataset
// ---------------------------------------------------------------------------------------------------------
__global__ void k_colmean(const double *X, int64_t N, int D, double *out) {
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
	// largest centred magnitude of the column (out[D + 1 + c]); k_xscale reduces over the columns
	double mx = 0.0;
	for (int64_t i = threadIdx.x; i < N; i += 256) mx = fmax(mx, fabs(X[i * D + c] - mean));
	red[threadIdx.x] = mx;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		out[c] = mean;
		out[D + 1 + c] = red[0];
	}
}
// out[D] = exponent of the power-of-two scale of the A operand
__global__ void k_xscale(double *out, int D) {
	double mx = 0.0;
	for (int c = 0; c < D; ++c) mx = fmax(mx, out[D + 1 + c]);
	out[D] = (double)g_scale_exp((float)mx);
}

// ---------------------------------------------------------------------------------------------------------
// A images of one block of steps: thread = (step, 16-byte piece)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pre_aimg(const double *X64, const double *xbar, const int32_t *order, int nsteps, int ntiles,
		uint8_t *Aimg) {
	const int idx = blockIdx.x * 256 + threadIdx.x;
	const int s = idx >> 3, q = idx & 7; // step, piece of 8 coordinates
	if (s >= ntiles * G_M) return;
	const float sx = ldexpf(1.0f, (int)xbar[GD]);
	__align__(16) __half hi[8], lo[8];
	if (s < nsteps) {
		const double *x = X64 + (size_t)order[s] * GD + q * 8;
#pragma unroll
		for (int e = 0; e < 8; ++e) g_split((float)(x[e] - xbar[q * 8 + e]) * sx, hi[e], lo[e]);
	} else {
#pragma unroll
		for (int e = 0; e < 8; ++e) hi[e] = lo[e] = __float2half_rn(0.0f);
	}
	const int t = s / G_M, r = s % G_M;
	uint8_t *dst = Aimg + (size_t)t * G_ASTAGE + g_sw128(r, q * 8);
	*reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(hi);
	*reinterpret_cast<uint4 *>(dst + 16384) = *reinterpret_cast<const uint4 *>(lo);
}

// ---------------------------------------------------------------------------------------------------------
// B images and epilogue constants of the slots marked dirty (or born two blocks ago): CTA = (chain, slot)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pre_bimg(const float *theta, const double *xbar, uint8_t *dirty, const uint32_t *born, uint8_t *Bimg,
		float *Bconst) {
	__shared__ float th[GPS + 3];
	__shared__ float red[256];
	const int cs = blockIdx.x; // chain * 32 + slot
	if (!dirty[cs] && !((born[cs >> 5] >> (cs & 31)) & 1u)) return;
	const float *src = theta + (size_t)cs * GPS;
	float mx = 0.0f;
	for (int i = threadIdx.x; i < GPS; i += 256) {
		const float v = src[i];
		th[i] = v;
		if (i >= GD && i < GD + GTRI) mx = fmaxf(mx, fabsf(v));
	}
	red[threadIdx.x] = mx;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
		__syncthreads();
	}
	const int et = g_scale_exp(red[0]), ex = (int)xbar[GD];
	const float st = ldexpf(1.0f, et);
	if (threadIdx.x < GD) {
		const int j = threadIdx.x;
		float s = 0.0f;
		for (int c = j; c < GD; ++c) s = fmaf(th[GD + npb_tri_off(GD, j, c)], (float)((double)th[c] - xbar[c]), s);
		Bconst[(size_t)cs * G_CONST + j] = -s;
	}
	if (threadIdx.x == GD) Bconst[(size_t)cs * G_CONST + GD] = th[GD + GTRI];
	if (threadIdx.x == GD + 1) Bconst[(size_t)cs * G_CONST + GD + 1] = ldexpf(1.0f, -(ex + et));
	uint8_t *img = Bimg + (size_t)cs * G_SLOT_IMG;
	// chunk: 0 (hi, rows j < 32)  1 (hi, rows j >= 32)  2 (lo, j < 32)  3 (lo, j >= 32); a row is the 64 columns of T2 row j
	for (int pc = threadIdx.x; pc < 2 * 256; pc += 256) {
		const int jh = pc >> 8, jl = (pc >> 3) & 31, p = pc & 7;
		const int j = jh * 32 + jl;
		__align__(16) __half hi[8], lo[8];
#pragma unroll
		for (int e = 0; e < 8; ++e) {
			const int c = p * 8 + e;
			g_split(c >= j ? th[GD + npb_tri_off(GD, j, c)] * st : 0.0f, hi[e], lo[e]);
		}
		*reinterpret_cast<uint4 *>(img + jh * G_CHUNK + g_sw128(jl, p * 8)) = *reinterpret_cast<const uint4 *>(hi);
		*reinterpret_cast<uint4 *>(img + (2 + jh) * G_CHUNK + g_sw128(jl, p * 8)) = *reinterpret_cast<const uint4 *>(lo);
	}
	__syncthreads();
	if (threadIdx.x == 0) dirty[cs] = 0;
}



// One warp per chain.  A tile of 32 steps is first decided SPECULATIVELY with lane = step: every lane walks the 32 slots
// of its own step with the member counts as they stand at the start of the tile -- 32 independent instruction streams
// instead of one dependent chain of warp reductions.  If no step of the tile moves its item (the steady state of a
// converged chain: > 99.9 % of the tiles) the counts never changed and the tile is done.  Otherwise the steps before the
// first move are final, and the sequential pass (lane = slot, the consumer of npb_alg8_tile4.cuh) takes over from that
// step; both passes evaluate the same keys (same noise, same operation order), so the result does not depend on which
// pass decided a step.  A chain that moves a lot (burn-in) skips the speculative pass.
// Upper bound of the race key of a step's best auxiliary draw, from the FIRST Box-Muller pair of every draw only (|v| and
// z_par: key_m <= c0_2 - D log2|v| - (|xw|/|v| - s z_par)^2 + log2(alpha/m) + noise, the chi-square term dropped, noise <= 23 by
// neg_lg2_exp1's clamp; + 1 for rounding and the two index bits packed into the key).  Walks the same stream as aux_race
// (npb_alg8_tile4.cuh), skipping the words of a draw it does not need.
template <int CD, int M>
__device__ __forceinline__ float g_aux_bound(const Philox &ph, const PriorDev &pr, float rn, uint32_t sj, uint32_t sweep, float ik2) {
	return aux_race_bound<CD, M>(ph, pr, rn, sj, sweep, ik2);
}
// the exact key, packed like k_aux_keys packs it (draw index in the two low mantissa bits)
template <int CD, int M>
__device__ __forceinline__ uint32_t g_aux_exact(const Philox &ph, const PriorDev &pr, float rn, uint32_t sj, uint32_t sweep, float ik2) {
	float ak;
	int am;
	aux_race<CD, M>(ph, pr, rn, sj, sweep, ik2, ak, am);
	return (__float_as_uint(ak) & ~3u) | (uint32_t)am;
}


// tile: [32 * 33] floats, [slot * 33 + step]; lg_s, lg1_s: [32] floats each, log2 n_k and log2 (n_k - 1) of every slot
// (-inf without members) -- shared memory private to the calling warp
// PF: the next tile's rows prefetched into registers while the current one is worked on (32 registers; pays when few warps are
// resident: the fused D = 64 kernel, k_race at four CTAs per SM) or loaded when needed (k_race at eight CTAs per SM)
// AM: where the auxiliary keys come from -- 0 the k_aux_keys pre-pass (a.aux_keys), 1 bounds per step computed here, 2 group
// bounds from k_aux_bound (a.aux_max); exact keys on demand in modes 1 and 2
template <int CD, int M, bool PF = true, int AM = 0>
__device__ __forceinline__ void g_consume_chain(const PreArgs &p, const int chain, const int lane, float *tile, float *lg_s, float *lg1_s) {
	constexpr int CPS = npb_ps(CD);
	const SweepArgs &a = p.a;
	const int N = a.N, C = a.C;
	const Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
	float *thc = a.theta + (size_t)chain * 32 * CPS;
	float *Lc = p.L + (size_t)chain * p.BS * 32;
	const uint32_t sweep = a.sweep0 + (uint32_t)p.sw;
	const int32_t *order = a.scan_order + (size_t)p.sw * N;
	float n = (float)a.counts[(size_t)chain * 32 + lane];
	float lgn = n > 0.0f ? fast_lg2(n) : -INFINITY, lgn1 = n > 1.0f ? fast_lg2(n - 1.0f) : -INFINITY;
	lg_s[lane] = lgn;
	lg1_s[lane] = lgn1;
	int kocc = __popc(__ballot_sync(0xffffffffu, n > 0.0f));
	unsigned long long st_cand = 0ull, st_moved = 0ull, st_births = 0ull;
	int overflow = 0;
	const int ntile = (p.nsteps + 31) / 32;
	uint32_t born_mask = 0u;

	// The table of this block was computed while the previous block was still being consumed (the fused schedule; with
	// separate launches the bookkeeping is the same): the columns of the slots born there are re-evaluated here, for the whole block (lane = step).
	if (p.born_prev) {
		uint32_t fix = p.born_prev[chain];
		while (fix) {
			const int k = __ffs(fix) - 1;
			fix &= fix - 1;
			for (int s = lane; s < p.nsteps; s += 32)
				Lc[(size_t)s * 32 + k] = g_stream_density<CD>(thc + (size_t)k * CPS, a.X + (size_t)order[p.s0 + s] * CD);
		}
		__threadfence();
		__syncwarp();
	}

	float nxt[PF ? 32 : 1];
	if (PF) {
#pragma unroll
		for (int jj = 0; jj < 32; ++jj) nxt[PF ? jj : 0] = __ldcg(Lc + (size_t)jj * 32 + lane); // rows past nsteps exist (BS is padded)
	}
	bool reload = !PF;

	for (int ti = 0; ti < ntile; ++ti) {
		const int b0 = ti * 32;          // first step of the tile within the block
		const int sj = p.s0 + b0 + lane; // lane = step (prologue)
		const bool valid = b0 + lane < p.nsteps;
		const int item = valid ? order[sj] : 0;
		const int zold = valid ? (int)a.z[(size_t)item * C + chain] : 0;
		int znew = zold;
		const uint32_t T = npb_mix32(npb_mix32(ph.k0 ^ ((uint32_t)(p.s0 + b0) * 0x9E3779B1u)) ^ ph.k1 ^ (sweep * 0x85EBCA77u) ^ 0x5bd1e995u);
		const int cnt = min(32, p.nsteps - b0);
		// The auxiliary draws' race key of step sj: read from the pre-pass (k_aux_keys) if the launch ran one, else LAZY: an
		// upper bound from the draws' first normals now, the exact key (same function, same bits) only for the steps whose
		// own key does not clear the bound, and for the whole tile if it enters the sequential pass.
		// GROUP BOUNDS (a.aux_max from k_aux_bound, no aux_keys): the bound is the largest over the tile's 32 steps, read once;
		// the sequential pass needs the exact keys only if some step's own slot does not clear the bound (see there).
		constexpr bool lazy = AM != 0, grp = AM == 2;
		const float ik2 = a.prior.inv_sqrt_kappa * (float)NPB_HALF_LOG2E_SQRT;
		const float rn_j = (lazy && valid) ? __ldg(a.Xwn + item) : 0.0f;
		uint32_t auxp = 0xff800000u;
		bool aux_exact = !lazy || !valid;
		if (!lazy && valid) auxp = __ldg(a.aux_keys + ((size_t)p.sw * C + chain) * N + sj);
		float auxkey_j = __uint_as_float(auxp);
		if (grp && valid) auxkey_j = __ldg(a.aux_max + ((size_t)p.sw * C + chain) * a.aux_groups + ((p.s0 + b0) >> 5)); // a bound until aux_exact
		else if (lazy && valid) auxkey_j = g_aux_bound<CD, M>(ph, a.prior, rn_j, (uint32_t)sj, sweep, ik2);
		__syncwarp();
		if (reload) {
#pragma unroll
			for (int jj = 0; jj < 32; ++jj) tile[lane * 33 + jj] = __ldcg(Lc + (size_t)(b0 + jj) * 32 + lane);
			reload = !PF;
		} else {
#pragma unroll
			for (int jj = 0; jj < 32; ++jj) tile[lane * 33 + jj] = nxt[PF ? jj : 0];
		}
		if (PF && ti + 1 < ntile) {
#pragma unroll
			for (int jj = 0; jj < 32; ++jj) nxt[PF ? jj : 0] = __ldcg(Lc + (size_t)(b0 + 32 + jj) * 32 + lane);
		}
		__syncwarp();
		if (!PF && b0 + 32 + lane < p.nsteps) // the next tile's row of this lane on its way into L2 while this tile is decided
			asm volatile("prefetch.global.L2 [%0];" ::"l"(Lc + (size_t)(b0 + 32 + lane) * 32));
		unsigned cand_tile = 0u;

		int j0 = 0;
		if (p.spec) {
			// ---- step-parallel pass: lane = step, repeated after every move ----
			// Every lane finds its step's winner with the member counts as they stand: its own slot's key first; the race noise is
			// capped at G_NOISE_CAP (g_noise), so a slot whose noiseless key lies more than the cap (+1 for rounding) below the own
			// key cannot win -- two shared loads, two adds and a compare per (step, slot), no logarithm -- and only the slots a lane
			// cannot exclude are evaluated exactly (same keys, same operation order as the sequential pass).  The lane keeps its
			// winner (slot w, count-independent part basew, key top) and an upper bound ubmax of every other slot's key.
			// The first lane whose item does not stay is the chain's next move: the steps before it are final.  A move changes
			// two member counts and nothing else (parameters are frozen between births), so the later lanes only re-evaluate the
			// keys of those two slots: the gaining slot's exactly (it may take over), the losing slot's if it is the lane's
			// winner (it must still clear ubmax).  A lane that cannot be decided this way -- a tie with the bound, an auxiliary
			// draw that may win, a slot left without members -- ends the pass: the sequential pass takes over at its step.
			int w = zold;
			float basew = tile[zold * 33 + lane] + g_noise(T, (uint32_t)lane, (uint32_t)zold);
			const float lg_own = lg1_s[zold];
			float top = lg_own > -INFINITY ? basew + lg_own : -INFINITY;
			float ubmax = -INFINITY;
			unsigned need = 0u;
			const float thr = top - (G_NOISE_CAP + 1.0f); // (the own slot itself never falls below it: its noise is at most the cap)
#pragma unroll 8
			for (int k = 0; k < 32; ++k) {
				const float v = tile[k * 33 + lane] + lg_s[k];
				const bool nd = v >= thr;
				if (nd) need |= 1u << k;
				ubmax = fmaxf(ubmax, nd ? -INFINITY : v);
			}
			ubmax += G_NOISE_CAP + 1.0f;
			need &= ~(1u << zold);
			if (!valid) need = 0u;
			while (need) {
				const int k = __ffs(need) - 1;
				need &= need - 1;
				const float lg = lg_s[k];
				const float b = tile[k * 33 + lane] + g_noise(T, (uint32_t)lane, (uint32_t)k);
				const float key = lg > -INFINITY ? b + lg : -INFINITY;
				if (key > top || (key == top && k < w)) { ubmax = fmaxf(ubmax, top); w = k; basew = b; top = key; }
				else ubmax = fmaxf(ubmax, key);
			}
			if (!aux_exact && !(top >= auxkey_j)) { // the bound does not settle it: the exact key
				auxp = g_aux_exact<CD, M>(ph, a.prior, rn_j, (uint32_t)sj, sweep, ik2);
				auxkey_j = __uint_as_float(auxp);
				aux_exact = true;
			}
			for (;;) {
				const bool safe = top > ubmax && top >= auxkey_j;
				const unsigned evm = __ballot_sync(0xffffffffu, valid && lane >= j0 && (!safe || w != zold));
				const int je = evm ? __ffs(evm) - 1 : cnt;
				cand_tile += (unsigned)((je - j0) * (kocc + M));
				j0 = je;
				if (je >= cnt) break;
				if (!__shfl_sync(0xffffffffu, (int)safe, je)) break; // the sequential pass decides step je and what follows
				const int src = __shfl_sync(0xffffffffu, zold, je), dst = __shfl_sync(0xffffffffu, w, je);
				cand_tile += (unsigned)(kocc + M);
				if (lane == src) n -= 1.0f;
				if (lane == dst) n += 1.0f;
				if (__any_sync(0xffffffffu, lane == src && n <= 0.0f)) { kocc--; cand_tile--; }
				if (lane == src || lane == dst) {
					lgn = n > 0.0f ? fast_lg2(n) : -INFINITY;
					lgn1 = n > 1.0f ? fast_lg2(n - 1.0f) : -INFINITY;
					lg_s[lane] = lgn;
					lg1_s[lane] = lgn1;
				}
				__syncwarp();
				st_moved++;
				if (lane == je) znew = dst;
				j0 = je + 1;
				if (lane > je) {
					const float lgd = dst == zold ? lg1_s[dst] : lg_s[dst]; // the gaining slot's key went up
					if (w == dst) top = basew + lgd;
					else {
						const float b = tile[dst * 33 + lane] + g_noise(T, (uint32_t)lane, (uint32_t)dst);
						const float key = lgd > -INFINITY ? b + lgd : -INFINITY;
						if (key > top || (key == top && dst < w)) { ubmax = fmaxf(ubmax, top); w = dst; basew = b; top = key; }
						else ubmax = fmaxf(ubmax, key);
					}
					if (w == src) { // the losing slot's went down
						const float lgs = src == zold ? lg1_s[src] : lg_s[src];
						top = lgs > -INFINITY ? basew + lgs : -INFINITY;
					}
				}
			}
		}
		if (j0 < cnt) {
			// ---- sequential pass from step j0: lane = slot ----
			if (grp) {
				// Group bound: no auxiliary draw of this tile can win if every step's OWN slot clears the bound even with its
				// member count run down to one other member by the tile's up to 31 earlier moves (log2 n >= 0): then the best
				// slot key of every step exceeds every auxiliary key whatever the sequential pass does, and the keys are not needed.
				const bool safe = !valid || aux_exact ||
						(lg1_s[zold] >= 5.1f && tile[zold * 33 + lane] + g_noise(T, (uint32_t)lane, (uint32_t)zold) > auxkey_j);
				if (__all_sync(0xffffffffu, safe)) {
					if (!aux_exact) {
						auxp = 0xff800000u;
						auxkey_j = -INFINITY;
						aux_exact = true;
					}
				}
			}
			if (!aux_exact) {
				auxp = g_aux_exact<CD, M>(ph, a.prior, rn_j, (uint32_t)sj, sweep, ik2);
				auxkey_j = __uint_as_float(auxp);
			}
			const int zold_aux_j = zold | ((int)(auxp & 3u) << 16);
			float base_next = tile[lane * 33 + j0] + g_noise(T, (uint32_t)j0, (uint32_t)lane);
			int zo_aux_next = __shfl_sync(0xffffffffu, zold_aux_j, j0);
			float ak_next = __shfl_sync(0xffffffffu, auxkey_j, j0);
			for (int j = j0; j < cnt; ++j) {
				const int zo_aux = zo_aux_next;
				const float ak = ak_next;
				const float base = base_next;
				const int zo = zo_aux & 0xffff;
				float noise_next;
				{
					const int jn = min(j + 1, 31);
					noise_next = g_noise(T, (uint32_t)jn, (uint32_t)lane);
					base_next = tile[lane * 33 + jn] + noise_next;
					zo_aux_next = __shfl_sync(0xffffffffu, zold_aux_j, jn);
					ak_next = __shfl_sync(0xffffffffu, auxkey_j, jn);
				}
				const float lg = (zo == lane) ? lgn1 : lgn;
				const float key = lg > -INFINITY ? base + lg : -INFINITY; // a slot without (other) members never wins
				const float top = fmaxf(redux_max_f32(key), ak);
				const unsigned bal = __ballot_sync(0xffffffffu, key == top && key > -INFINITY);
				cand_tile += (unsigned)(kocc + M);
				int new_slot;
				bool born = false;
				if (bal != 0u) {
					new_slot = __ffs(bal) - 1;
				} else {
					born = true;
					new_slot = zo;
				}
				if (born || new_slot != zo) {
					// retract (membertrix.cpp:175-233)
					bool dead = false;
					if (zo == lane) {
						n -= 1.0f;
						dead = n <= 0.0f;
					}
					const bool died = __any_sync(0xffffffffu, dead);
					if (died) {
						kocc--;
						cand_tile--;
					}
					if (born) {
						// np_neal_algorithm8.cpp:136-145: the lowest free slot takes theta' of the winning auxiliary draw
						const unsigned fb = __ballot_sync(0xffffffffu, n <= 0.0f);
						const int fs = fb ? __ffs(fb) - 1 : -1;
						if (fs < 0) {
							overflow = 1; // no room: the item goes back where it was
							if (died) kocc++;
						} else {
							new_slot = fs;
							const int m = (zo_aux >> 16) & 0xff;
							const uint32_t step = (uint32_t)(p.s0 + b0 + j);
							const int bitem = order[step];
							g_birth_theta<CD>(ph, a.prior, a.Xw + (size_t)bitem * CD, __ldg(a.Xwn + bitem), step, sweep, m, lane, thc + (size_t)fs * CPS);
							__threadfence();
							__syncwarp();
							born_mask |= 1u << fs;
							kocc++;
							st_births++;
							// the newborn slot's column of L for the rest of the block (lane = step)
							for (int s = b0 + j + 1 + lane; s < p.nsteps; s += 32)
								Lc[(size_t)s * 32 + fs] = g_stream_density<CD>(thc + (size_t)fs * CPS, a.X + (size_t)order[p.s0 + s] * CD);
							__threadfence();
							__syncwarp();
							{
								const int s = b0 + lane;
								if (lane > j && s < p.nsteps) tile[fs * 33 + lane] = __ldcg(Lc + (size_t)s * 32 + fs);
							}
							__syncwarp();
							reload = true; // the prefetched rows of the next tile predate the column
							if (lane == fs) base_next = tile[fs * 33 + min(j + 1, 31)] + noise_next;
						}
					}
					if (new_slot == lane) n += 1.0f;
					lgn = n > 0.0f ? fast_lg2(n) : -INFINITY;
					lgn1 = n > 1.0f ? fast_lg2(n - 1.0f) : -INFINITY;
					st_moved++;
					if (lane == j) znew = new_slot;
				}
			}
			lg_s[lane] = lgn;
			lg1_s[lane] = lgn1;
		}
		if (valid && znew != zold) a.z[(size_t)item * C + chain] = (npb_z_t)znew;
		st_cand += cand_tile;
	}
	a.counts[(size_t)chain * 32 + lane] = (int)n;
	if (lane == 0) {
		p.born_out[chain] = born_mask;
		a.kocc[chain] = kocc;
		if (overflow) a.overflow[chain] = 1;
		a.st[(size_t)chain * 4 + 0] += st_cand;
		a.st[(size_t)chain * 4 + 1] += st_moved;
		a.st[(size_t)chain * 4 + 2] += st_births;
	}
}

// ---------------------------------------------------------------------------------------------------------
// k_density_tc: the density table of block k + 1 AND the race of block k in one persistent kernel.
// Warps 0 .. EW-1: epilogue (TMEM lanes 32 (w % 4) .. + 31 = steps of the tile), warp EW: MMA issue + TMEM allocation,
// warp EW + 1: bulk-copy producer, warps EW + 2, EW + 3: the race (g_consume_chain) of chains blockIdx.x and
// blockIdx.x + gridDim.x, ... of the PREVIOUS block, whose table sits in the other buffer.  (Two kernels on two streams did
// the same, but whether they overlapped was left to the block scheduler: sweeps came out at 66 or at 94 ms.)
// ---------------------------------------------------------------------------------------------------------
template <int EW, int M> // EW epilogue warps: 4 (each takes the unit's four slots) or 8 (two slots each); M auxiliary draws
__global__ void __launch_bounds__(EW * 32 + 128, 1) k_density_tc(const GemmArgs g, const PreArgs p, const int do_density, const int do_consume) {
	constexpr int SPW = 16 / EW; // slots per epilogue warp
	extern __shared__ uint8_t g_smem_raw[];
	const uint32_t raw = g_smem_u32(g_smem_raw);
	const uint32_t base = (raw + 1023u) & ~1023u;
	uint8_t *gen = g_smem_raw + (base - raw);
	// misc area: barriers, TMEM address, epilogue constants
	const uint32_t misc = base + G_BBYTES + G_STAGES * G_ASTAGE;
	const uint32_t bar_b_full = misc, bar_b_empty = misc + 8;
	const uint32_t bar_a_full = misc + 16, bar_a_empty = misc + 16 + 8 * G_STAGES;
	const uint32_t bar_t_full = misc + 16 + 16 * G_STAGES, bar_t_empty = bar_t_full + 16;
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + G_BBYTES + G_STAGES * G_ASTAGE + 256);
	float *econst = reinterpret_cast<float *>(gen + G_BBYTES + G_STAGES * G_ASTAGE + 512); // [4][G_CONST]
	float *cons_smem = reinterpret_cast<float *>(gen + G_BBYTES + G_STAGES * G_ASTAGE + G_SMEM_MISC); // [2][G_CONS_FLOATS]
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int n_units = do_density ? g.C * (32 / G_NS) : 0;

	if (warp == EW + 1 && lane == 0) {
		g_mbar_init(bar_b_full, 1);
		g_mbar_init(bar_b_empty, 1);
		for (int s = 0; s < G_STAGES; ++s) {
			g_mbar_init(bar_a_full + 8 * s, 1);
			g_mbar_init(bar_a_empty + 8 * s, 1);
		}
		for (int b = 0; b < 2; ++b) {
			g_mbar_init(bar_t_full + 8 * b, 1);
			g_mbar_init(bar_t_empty + 8 * b, EW * 32);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async;" ::: "memory");
	}
	if (warp == EW) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(g_smem_u32(tmem_slot)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	g_tc_fence_before();
	__syncthreads();
	g_tc_fence_after();
	const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

	if (warp == EW + 1) {
		// ===================== bulk-copy producer =====================
		if (lane == 0) {
			uint32_t a_it = 0, unit_it = 0;
			for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++unit_it) {
				g_mbar_wait(bar_b_empty, (unit_it & 1u) ^ 1u); // the MMAs of the previous unit have read B
				g_mbar_expect_tx(bar_b_full, G_BBYTES);
				const int c = u / (32 / G_NS), gq = u % (32 / G_NS);
				for (int sl = 0; sl < G_NS; ++sl) {
					const uint8_t *src = g.Bimg + ((size_t)c * 32 + gq * G_NS + sl) * G_SLOT_IMG;
					g_bulk_g2s(base + G_BHI + sl * G_CHUNK, src + 0 * G_CHUNK, G_CHUNK, bar_b_full);
					g_bulk_g2s(base + G_BHI + 16384 + sl * G_CHUNK, src + 1 * G_CHUNK, G_CHUNK, bar_b_full);
					g_bulk_g2s(base + G_BLO + sl * G_CHUNK, src + 2 * G_CHUNK, G_CHUNK, bar_b_full);
					g_bulk_g2s(base + G_BLO + 16384 + sl * G_CHUNK, src + 3 * G_CHUNK, G_CHUNK, bar_b_full);
				}
				for (int t = 0; t < g.ntiles; ++t, ++a_it) {
					const uint32_t s = a_it % G_STAGES, ph = (a_it / G_STAGES) & 1u;
					g_mbar_wait(bar_a_empty + 8 * s, ph ^ 1u);
					g_mbar_expect_tx(bar_a_full + 8 * s, G_ASTAGE);
					g_bulk_g2s(base + G_A0 + s * G_ASTAGE, g.Aimg + (size_t)t * G_ASTAGE, G_ASTAGE, bar_a_full + 8 * s);
				}
			}
		}
		__syncwarp();
	} else if (warp == EW) {
		// ===================== MMA issue (one thread) =====================
		if (lane == 0) {
			constexpr uint32_t ID128 = g_idesc(G_M, 128), ID256 = g_idesc(G_M, 256);
			uint32_t a_it = 0, tile_it = 0, unit_it = 0;
			for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++unit_it) {
				g_mbar_wait(bar_b_full, unit_it & 1u);
				g_tc_fence_after();
				for (int t = 0; t < g.ntiles; ++t, ++tile_it, ++a_it) {
					const uint32_t buf = tile_it & 1u;
					g_mbar_wait(bar_t_empty + 8 * buf, ((tile_it >> 1) & 1u) ^ 1u); // the epilogue has drained this accumulator
					const uint32_t s = a_it % G_STAGES;
					g_mbar_wait(bar_a_full + 8 * s, (a_it / G_STAGES) & 1u);
					g_tc_fence_after();
					const uint32_t dcol = tmem + buf * 256u;
					const uint32_t Ahi = base + G_A0 + s * G_ASTAGE, Alo = Ahi + 16384;
#pragma unroll
					for (int prod = 0; prod < 3; ++prod) {
						const uint32_t A = prod == 2 ? Alo : Ahi, B = base + (prod == 1 ? G_BLO : G_BHI);
#pragma unroll
						for (int k = 0; k < 4; ++k) { // K-step = 16 columns = 32 bytes of the swizzled row
							const uint64_t ad = g_desc(A + k * 32), bd = g_desc(B + k * 32);
							if (k < 2) {
								// columns c < 32 only meet rows j < 32 of the triangular factor: N = 128 (accumulator columns 0-127)
								g_mma_f16(dcol, ad, bd, ID128, (prod | k) != 0);
							} else if (prod == 0 && k == 2) {
								// columns 0-127 continue, columns 128-255 (rows j >= 32) start here
								g_mma_f16(dcol, ad, bd, ID128, 1u);
								g_mma_f16(dcol + 128u, ad, g_desc(B + 16384 + k * 32), ID128, 0u);
							} else {
								g_mma_f16(dcol, ad, bd, ID256, 1u);
							}
						}
					}
					g_tc_commit(bar_a_empty + 8 * s); // frees the stage once these MMAs have read it
					g_tc_commit(bar_t_full + 8 * buf);
				}
				g_tc_commit(bar_b_empty);
			}
		}
		__syncwarp();
	} else if (warp >= EW + 2) {
		// ===================== the race of the previous block: one warp per chain =====================
		if (do_consume) {
			float *sm = cons_smem + (warp - (EW + 2)) * G_CONS_FLOATS;
			for (int chain = blockIdx.x + (warp - (EW + 2)) * gridDim.x; chain < p.a.C; chain += 2 * gridDim.x)
				g_consume_chain<GD, M>(p, chain, lane, sm, sm + 32 * 33, sm + 32 * 33 + 32);
		}
	} else {
		// ===================== epilogue: thread = (step of the tile, pair of slots) =====================
		const int wq = warp & 3, eg = warp >> 2;
		const int row = wq * 32 + lane;
		uint32_t tile_it = 0;
		for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
			const int c = u / (32 / G_NS), gq = u % (32 / G_NS);
			asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
			{
				const float *src = g.Bconst + ((size_t)c * 32 + gq * G_NS) * G_CONST;
				for (int i = threadIdx.x; i < G_NS * G_CONST; i += EW * 32) econst[i] = __ldg(src + i);
			}
			asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
			float *Lc = g.L + ((size_t)c * g.BS) * 32 + gq * G_NS + eg * SPW;
			for (int t = 0; t < g.ntiles; ++t, ++tile_it) {
				const uint32_t buf = tile_it & 1u;
				g_mbar_wait(bar_t_full + 8 * buf, (tile_it >> 1) & 1u);
				g_tc_fence_after();
				const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + buf * 256u;
				float out[SPW];
#pragma unroll
				for (int h = 0; h < SPW; ++h) {
					const int sl = eg * SPW + h;
					const float *ec = econst + sl * G_CONST;
					const float dsc = ec[GD + 1];
					float v0[32], v1[32];
					g_tmem_ld32_nowait(taddr + sl * 32u, v0);
					g_tmem_ld32_nowait(taddr + 128u + sl * 32u, v1);
					g_tmem_wait_ld(v0, v1);
					if (h == SPW - 1) { // the last of the accumulator is in registers: hand the buffer back before the arithmetic
						g_tc_fence_before();
						g_mbar_arrive(bar_t_empty + 8 * buf);
					}
					float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
#pragma unroll
					for (int i = 0; i < 32; i += 4) {
						const float4 na = *reinterpret_cast<const float4 *>(ec + i);
						const float4 nb = *reinterpret_cast<const float4 *>(ec + 32 + i);
						const float a0 = fmaf(v0[i], dsc, na.x), a1 = fmaf(v0[i + 1], dsc, na.y), a2 = fmaf(v0[i + 2], dsc, na.z),
									    a3 = fmaf(v0[i + 3], dsc, na.w);
						const float b0 = fmaf(v1[i], dsc, nb.x), b1 = fmaf(v1[i + 1], dsc, nb.y), b2 = fmaf(v1[i + 2], dsc, nb.z),
									    b3 = fmaf(v1[i + 3], dsc, nb.w);
						q0 = fmaf(a0, a0, q0); q1 = fmaf(a1, a1, q1); q2 = fmaf(a2, a2, q2); q3 = fmaf(a3, a3, q3);
						q0 = fmaf(b0, b0, q0); q1 = fmaf(b1, b1, q1); q2 = fmaf(b2, b2, q2); q3 = fmaf(b3, b3, q3);
					}
					out[h] = ec[GD] - ((q0 + q1) + (q2 + q3));
				}
				if constexpr (SPW == 4) *reinterpret_cast<float4 *>(Lc + (size_t)(t * G_M + row) * 32) = make_float4(out[0], out[1], out[2], out[3]);
				else *reinterpret_cast<float2 *>(Lc + (size_t)(t * G_M + row) * 32) = make_float2(out[0], out[1]);
			}
		}
	}
	g_tc_fence_before();
	__syncthreads();
	if (warp == EW) {
		g_tc_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
	}
}

// ---------------------------------------------------------------------------------------------------------
// the same table on the FP32 pipe (NPB_D64_DENSITY=fp32): thread = step, loop over the 32 slots
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_density_fp32(const double *X64, const double *xbar, const int32_t *order, int nsteps,
		const float *theta, float *L, int BS) {
	const int s = blockIdx.x * 128 + threadIdx.x, c = blockIdx.y;
	if (s >= nsteps) return;
	float x[GD];
	{
		const double *xr = X64 + (size_t)order[s] * GD;
#pragma unroll
		for (int i = 0; i < GD; ++i) x[i] = (float)(xr[i] - xbar[i]);
	}
	for (int k = 0; k < 32; ++k) {
		const float *th = theta + ((size_t)c * 32 + k) * GPS;
		float q = 0.0f;
#pragma unroll 4
		for (int r = 0; r < GD; ++r) {
			float y = 0.0f;
			for (int cc = r; cc < GD; ++cc) y = fmaf(__ldg(th + GD + npb_tri_off(GD, r, cc)), x[cc] - (float)((double)__ldg(th + cc) - xbar[cc]), y);
			q = fmaf(y, y, q);
		}
		L[((size_t)c * BS + s) * 32 + k] = __ldg(th + GD + GTRI) - q;
	}
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
template npb_status npb_launch_aux_keys<64>(npb_chains *, const SweepArgs &);

static int g_block_steps(const npb_chains *ch) {
	int v = ch->sw.d64_block;
	if (v < 128) v = 128;
	if (v > (1 << 20)) v = 1 << 20;
	return (v + 127) & ~127;
}
static bool g_use_fp32(const npb_chains *ch) { return ch->sw.d64_fp32; }

static npb_status g_ensure(npb_chains *ch) {
	npb_ctx *ctx = ch->ctx;
	npb_dataset *ds = ch->ds;
	int BS = g_block_steps(ch);
	if ((int64_t)BS > ((ds->N + 127) & ~(int64_t)127)) BS = (int)((ds->N + 127) & ~(int64_t)127); // no larger than the sweep
	if (!ds->Xbar) NPB_CUDA_OK(cudaMalloc((void **)&ds->Xbar, sizeof(double) * (2 * GD + 1))); // means, scale exponent, column maxima
	if (!ds->xbar_valid) {
		ds->xbar_valid = true;
		k_colmean<<<GD, 256, 0, ctx->stream>>>(ds->X64, ds->N, GD, ds->Xbar);
		NPB_CUDA_OK(cudaGetLastError());
		k_xscale<<<1, 1, 0, ctx->stream>>>(ds->Xbar, GD);
		NPB_CUDA_OK(cudaGetLastError());
	}
	if (!ch->g_L) {
		const size_t C = (size_t)ch->C;
		NPB_CUDA_OK(cudaMalloc((void **)&ch->g_aimg, (size_t)(BS / G_M) * G_ASTAGE));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->g_bimg, C * 32 * G_SLOT_IMG));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->g_bconst, C * 32 * G_CONST * sizeof(float)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->g_L, 2 * C * (size_t)(BS + 32) * 32 * sizeof(float)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->g_dirty, C * 32));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->g_born, 2 * C * sizeof(uint32_t)));
		NPB_CUDA_OK(cudaMemsetAsync(ch->g_L, 0, 2 * C * (size_t)(BS + 32) * 32 * sizeof(float), ctx->stream));
		NPB_CUDA_OK(cudaMemsetAsync(ch->g_born, 0, 2 * C * sizeof(uint32_t), ctx->stream));
		ch->g_bs = BS;
		NPB_CUDA_OK(cudaFuncSetAttribute(k_density_tc<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
		NPB_CUDA_OK(cudaFuncSetAttribute(k_density_tc<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
	}
	return NPB_OK;
}

// One launch of the fused kernel: the log2-density table of `nsteps` steps (scan order `d_order`) of every chain into
// table buffer `buf` (d_order != NULL), and / or the race of the block described by *cons (cons != NULL).
static npb_status g_launch_block(npb_chains *ch, const int32_t *d_order, int nsteps, int buf, const PreArgs *cons) {
	npb_ctx *ctx = ch->ctx;
	const int C = (int)ch->C, BSP = ch->g_bs + 32;
	float *L = ch->g_L + (size_t)buf * C * BSP * 32;
	GemmArgs g;
	memset(&g, 0, sizeof(g));
	int do_density = 0;
	if (d_order && g_use_fp32(ch)) {
		dim3 grid((nsteps + 127) / 128, C);
		k_density_fp32<<<grid, 128, 0, ctx->stream>>>(ch->ds->X64, ch->ds->Xbar, d_order, nsteps, ch->theta, L, BSP);
		NPB_CUDA_OK(cudaGetLastError());
	} else if (d_order) {
		const int ntiles = (nsteps + G_M - 1) / G_M;
		k_pre_aimg<<<(ntiles * G_M * 8 + 255) / 256, 256, 0, ctx->stream>>>(ch->ds->X64, ch->ds->Xbar, d_order, nsteps, ntiles, ch->g_aimg);
		NPB_CUDA_OK(cudaGetLastError());
		k_pre_bimg<<<C * 32, 256, 0, ctx->stream>>>(ch->theta, ch->ds->Xbar, ch->g_dirty, ch->g_born + (size_t)buf * C, ch->g_bimg, ch->g_bconst);
		NPB_CUDA_OK(cudaGetLastError());
		g.Aimg = ch->g_aimg;
		g.Bimg = ch->g_bimg;
		g.Bconst = ch->g_bconst;
		g.L = L;
		g.ntiles = ntiles;
		g.BS = BSP;
		do_density = 1;
	}
	g.C = C;
	if (!do_density && !cons) return NPB_OK;
	const int n_sm = ctx->n_sm;
	const int n_units = C * (32 / G_NS);
	const int grid = n_units < n_sm ? n_units : n_sm;
	PreArgs p;
	if (cons) p = *cons;
	else memset(&p, 0, sizeof(p));
	if (ch->m_aux == 3) k_density_tc<4, 3><<<grid, 4 * 32 + 128, G_SMEM, ctx->stream>>>(g, p, do_density, cons ? 1 : 0);
	else k_density_tc<4, 1><<<grid, 4 * 32 + 128, G_SMEM, ctx->stream>>>(g, p, do_density, cons ? 1 : 0);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

npb_status npb_launch_alg8_gemm64(npb_chains *ch, const SweepArgs &a) {
	npb_ctx *ctx = ch->ctx;
	if (ch->m_aux != 3 && ch->m_aux != 1) return npb_fail(ctx, NPB_E_UNSUPPORTED, "m_aux must be 1 or 3 for the D = 64 sweep kernels");
	npb_status s = g_ensure(ch);
	if (s != NPB_OK) return s;
	s = npb_launch_aux_keys<64>(ch, a);
	if (s != NPB_OK) return s;
	const size_t C = (size_t)ch->C;
	// parameters may have changed since the last launch (init_from_params, update_params): every slot's image is rebuilt
	NPB_CUDA_OK(cudaMemsetAsync(ch->g_dirty, 1, C * 32, ctx->stream));
	// Software pipeline over the blocks of steps: launch k computes the table of block k into buffer k & 1 while its race
	// warps consume block k - 1 from the other buffer.  A slot born during block k - 1 is therefore missing from table k: its
	// column is re-evaluated by the race of block k (born_prev), and it enters the operand images with block k + 1
	// (k_pre_bimg reads the births of block k - 1 from g_born[(k + 1) & 1]).  The block counter runs on across launches, so
	// that batching the sweeps differently does not change which code evaluated which column: results are bit-identical
	// however a run is split, with the overlap (NPB_D64_OVERLAP=0: table and race in separate launches) or without.
	const int BS = ch->g_bs, N = a.N, BSP = BS + 32;
	const bool overlap = ch->sw.d64_overlap;
	PreArgs p, pending;
	bool have_pending = false;
	p.a = a;
	p.BS = BSP;
	p.spec = ch->sw.spec != 0;
	for (int sw = 0; sw < a.n_sweeps; ++sw) {
		for (int s0 = 0; s0 < N; s0 += BS, ++ch->g_k) {
			const int nsteps = N - s0 < BS ? N - s0 : BS;
			const int buf = (int)(ch->g_k & 1u);
			s = g_launch_block(ch, a.scan_order + (size_t)sw * N + s0, nsteps, buf, have_pending ? &pending : nullptr);
			if (s != NPB_OK) return s;
			p.L = ch->g_L + (size_t)buf * C * BSP * 32;
			p.born_prev = ch->g_born + (size_t)(buf ^ 1) * C;
			p.born_out = ch->g_born + (size_t)buf * C;
			p.sw = sw;
			p.s0 = s0;
			p.nsteps = nsteps;
			if (overlap) {
				pending = p;
				have_pending = true;
			} else {
				s = g_launch_block(ch, nullptr, 0, buf, &p);
				if (s != NPB_OK) return s;
			}
		}
	}
	if (have_pending) {
		s = g_launch_block(ch, nullptr, 0, 0, &pending);
		if (s != NPB_OK) return s;
	}
	return NPB_OK;
}

// parity probe (npb_chains_probe_tile_logdensity at D = 64): the [32 slots x 32 items] table exactly as the sweep reads it
__global__ void k_gemm64_probe_out(const float *L, const int *counts, int chain, int BSP, float *out) {
	const int k = threadIdx.x, j = blockIdx.x;
	const bool occupied = counts[(size_t)chain * 32 + k] > 0;
	out[k * 32 + j] = occupied ? L[((size_t)chain * BSP + j) * 32 + k] * NPB_LN2 : NAN;
}

npb_status npb_launch_gemm64_probe(npb_chains *ch, int chain, const int32_t *d_items, float *d_out) {
	npb_ctx *ctx = ch->ctx;
	npb_status s = g_ensure(ch);
	if (s != NPB_OK) return s;
	NPB_CUDA_OK(cudaMemsetAsync(ch->g_dirty, 1, (size_t)ch->C * 32, ctx->stream));
	s = g_launch_block(ch, d_items, 32, 0, nullptr);
	if (s != NPB_OK) return s;
	k_gemm64_probe_out<<<32, 32, 0, ctx->stream>>>(ch->g_L, ch->counts, chain, ch->g_bs + 32, d_out);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

// =========================================================================================================
// D = 16 (Kmax = 32) on the same tensor path: the default there (NPB_D16_PATH=fp32 selects k_alg8_sweep_tile4, the FP32-pipe
// kernel it measured 2x faster than at the headline shape).  A 16-D density is one K = 16 MMA step; the three FP16 products hi*hi + hi*lo + lo*hi are laid out ALONG K so that
// the 128-byte swizzled row of the D = 64 images carries them: A row = [x_hi | x_hi | x_lo | 0], B row = [T_hi | T_lo | T_hi | 0]
// (16 FP16 each) -- the fourth quarter carries the per-row offset nb = -T2 (mu - xbar) against a constant column of A, so the
// accumulator is y itself and the epilogue a plain sum of squares: four K-steps of one accumulation.  Unit of work = (chain, 16 slots): N = 256 = 16 slots x 16 rows, the unit's B
// image is one contiguous 32 KB copy.  With thousands of chains the race needs no fusion: k_race16 (one warp per chain)
// fills the GPU on its own, so a block is table kernel, then race kernel, on one stream (the schedule NPB_D64_OVERLAP=0
// tests at D = 64).
// =========================================================================================================

__global__ void __launch_bounds__(256) k_pre_aimg16(const double *X64, const double *xbar, const int32_t *order, int nsteps, int ntiles,
		uint8_t *Aimg) {
	const int idx = blockIdx.x * 256 + threadIdx.x;
	const int s = idx >> 1, q = idx & 1; // step, half of the 16 coordinates
	if (s >= ntiles * G_M) return;
	const float sx = ldexpf(1.0f, (int)xbar[HD]);
	__align__(16) __half hi[8], lo[8];
	if (s < nsteps) {
		const double *x = X64 + (size_t)order[s] * HD + q * 8;
#pragma unroll
		for (int e = 0; e < 8; ++e) g_split((float)(x[e] - xbar[q * 8 + e]) * sx, hi[e], lo[e]);
	} else {
#pragma unroll
		for (int e = 0; e < 8; ++e) hi[e] = lo[e] = __float2half_rn(0.0f);
	}
	const int t = s / G_M, r = s % G_M;
	uint8_t *row = Aimg + (size_t)t * H_ASTAGE;
	*reinterpret_cast<uint4 *>(row + g_sw128(r, q * 8)) = *reinterpret_cast<const uint4 *>(hi);      // K  0-15: x_hi
	*reinterpret_cast<uint4 *>(row + g_sw128(r, 16 + q * 8)) = *reinterpret_cast<const uint4 *>(hi); // K 16-31: x_hi
	*reinterpret_cast<uint4 *>(row + g_sw128(r, 32 + q * 8)) = *reinterpret_cast<const uint4 *>(lo); // K 32-47: x_lo
	// K 48-59: the constant 2^15 that multiplies the folded offset columns of B (k_pre_bimg16), K 60-63: zero
	__align__(16) __half one[8];
#pragma unroll
	for (int e = 0; e < 8; ++e) one[e] = __float2half_rn((q == 0 || e < 4) ? 32768.0f : 0.0f);
	*reinterpret_cast<uint4 *>(row + g_sw128(r, 48 + q * 8)) = *reinterpret_cast<const uint4 *>(one);
}

// CTA = 8 slots, warp = slot, lane = (row j, half of the columns)
__global__ void __launch_bounds__(256) k_pre_bimg16(const float *theta, const double *xbar, uint8_t *dirty, const uint32_t *born, uint8_t *Bimg,
		float *Bconst, int n_slots) {
	__shared__ float ths[8][HPS + 3];
	const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int cs = blockIdx.x * 8 + w; // chain * 32 + slot
	if (cs >= n_slots) return;
	if (!dirty[cs] && !((born[cs >> 5] >> (cs & 31)) & 1u)) return;
	float *th = ths[w];
	const float *src = theta + (size_t)cs * HPS;
	float mx = 0.0f;
	for (int i = lane; i < HPS; i += 32) {
		const float v = src[i];
		th[i] = v;
		if (i >= HD && i < HD + HTRI) mx = fmaxf(mx, fabsf(v));
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	__syncwarp();
	const int et = g_scale_exp(mx), ex = (int)xbar[HD];
	const float st = ldexpf(1.0f, et);
	// nb_j = -(T2 (mu - xbar))_j enters the GEMM as well: in accumulator units it is nbs_j = nb_j 2^(ex + et), written as
	// 2^15 * 4 * (h1 + h2 + h3) with three FP16 terms of nbs_j / 2^17, each repeated over four K columns against the constant
	// 2^15 of the A image -- if it fits (|nbs| < 2^31); a slot whose mean lies far outside the data keeps nb for the epilogue
	// (flag in Bconst[18]; the epilogue takes its short form only for units whose 16 slots are all folded).
	float nbj = 0.0f;
	if (lane < HD) {
		float s = 0.0f;
		for (int c = lane; c < HD; ++c) s = fmaf(th[HD + npb_tri_off(HD, lane, c)], (float)((double)th[c] - xbar[c]), s);
		nbj = -s;
	}
	const float up = ldexpf(1.0f, ex + et - 17);
	float nmax = fabsf(nbj) * up;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
	const bool fold = nmax < 16384.0f; // |nbs / 2^17| < 2^14
	if (lane < HD) Bconst[(size_t)cs * H_CONST + lane] = fold ? 0.0f : nbj;
	const float dsc = ldexpf(1.0f, -(ex + et));
	if (lane == HD) Bconst[(size_t)cs * H_CONST + HD] = th[HD + HTRI];
	if (lane == HD + 1) Bconst[(size_t)cs * H_CONST + HD + 1] = dsc;
	if (lane == HD + 2) Bconst[(size_t)cs * H_CONST + HD + 2] = fold ? 1.0f : 0.0f;
	if (lane == HD + 3) Bconst[(size_t)cs * H_CONST + HD + 3] = dsc * dsc;
	const int j = lane >> 1, q = lane & 1;
	__align__(16) __half hi[8], lo[8];
#pragma unroll
	for (int e = 0; e < 8; ++e) {
		const int c = q * 8 + e;
		g_split(c >= j ? th[HD + npb_tri_off(HD, j, c)] * st : 0.0f, hi[e], lo[e]);
	}
	uint8_t *img = Bimg + (size_t)cs * H_SLOT_IMG;
	*reinterpret_cast<uint4 *>(img + g_sw128(j, q * 8)) = *reinterpret_cast<const uint4 *>(hi);      // K  0-15: T_hi
	*reinterpret_cast<uint4 *>(img + g_sw128(j, 16 + q * 8)) = *reinterpret_cast<const uint4 *>(lo); // K 16-31: T_lo
	*reinterpret_cast<uint4 *>(img + g_sw128(j, 32 + q * 8)) = *reinterpret_cast<const uint4 *>(hi); // K 32-47: T_hi
	{
		const float w = fold ? __shfl_sync(0xffffffffu, nbj, j) * up : 0.0f; // nbs_j / 2^17
		const __half h1 = __float2half_rn(w);
		const float r1 = w - __half2float(h1);
		const __half h2 = __float2half_rn(r1);
		const __half h3 = __float2half_rn(r1 - __half2float(h2));
		const __half z = __float2half_rn(0.0f);
		__align__(16) __half off[8];
#pragma unroll
		for (int e = 0; e < 8; ++e) off[e] = q == 0 ? (e < 4 ? h1 : h2) : (e < 4 ? h3 : z); // K 48-51 h1, 52-55 h2, 56-59 h3, 60-63 0
		*reinterpret_cast<uint4 *>(img + g_sw128(j, 48 + q * 8)) = *reinterpret_cast<const uint4 *>(off);
	}
	__syncwarp();
	if (lane == 0) dirty[cs] = 0;
}

// warps 0 .. EW-1 epilogue, warp EW MMA issue + TMEM allocation, warp EW + 1 bulk-copy producer.
// Unit of work = NH half-chains (16 slots each, NH / 2 consecutive chains): their B images stay resident and every A tile that
// streams in is used NH times -- the A stream through L2 was this kernel's largest traffic (ncu: 9.1 GB per block with
// NH = 1, lts throughput 56 %).
template <int EW, int NH>
__global__ void __launch_bounds__(EW * 32 + 64, 1) k_density_tc16(const GemmArgs g) {
	constexpr int PPW = 32 / EW; // pairs of slots per epilogue warp: 8 (EW = 4), 4 (EW = 8) or 2 (EW = 16)
	constexpr uint32_t BB = NH * H_BBYTES;
	extern __shared__ uint8_t g_smem_raw[];
	const uint32_t raw = g_smem_u32(g_smem_raw);
	const uint32_t base = (raw + 1023u) & ~1023u;
	uint8_t *gen = g_smem_raw + (base - raw);
	constexpr uint32_t A0 = BB, MISC = BB + H_STAGES * H_ASTAGE;
	const uint32_t misc = base + MISC;
	const uint32_t bar_b_full = misc, bar_b_empty = misc + 8;
	const uint32_t bar_a_full = misc + 16, bar_a_empty = misc + 16 + 8 * H_STAGES;
	const uint32_t bar_t_full = misc + 16 + 16 * H_STAGES, bar_t_empty = bar_t_full + 16;
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + MISC + 256);
	float *econst = reinterpret_cast<float *>(gen + MISC + 512); // [NH * 16][H_CONST]
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int n_half = g.C * 2;
	const int n_units = (n_half + NH - 1) / NH;

	if (warp == EW + 1 && lane == 0) {
		g_mbar_init(bar_b_full, 1);
		g_mbar_init(bar_b_empty, 1);
		for (int s = 0; s < H_STAGES; ++s) {
			g_mbar_init(bar_a_full + 8 * s, 1);
			g_mbar_init(bar_a_empty + 8 * s, 1);
		}
		for (int b = 0; b < 2; ++b) {
			g_mbar_init(bar_t_full + 8 * b, 1);
			g_mbar_init(bar_t_empty + 8 * b, EW * 32);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async;" ::: "memory");
	}
	if (warp == EW) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(g_smem_u32(tmem_slot)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	g_tc_fence_before();
	__syncthreads();
	g_tc_fence_after();
	const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

	if (warp == EW + 1) {
		if (lane == 0) {
			uint32_t a_it = 0, unit_it = 0;
			for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++unit_it) {
				const int nh = min(NH, n_half - u * NH);
				g_mbar_wait(bar_b_empty, (unit_it & 1u) ^ 1u);
				g_mbar_expect_tx(bar_b_full, (uint32_t)nh * H_BBYTES);
				g_bulk_g2s(base, g.Bimg + (size_t)u * BB, (uint32_t)nh * H_BBYTES, bar_b_full);
				for (int t = 0; t < g.ntiles; ++t, ++a_it) {
					const uint32_t s = a_it % H_STAGES, ph = (a_it / H_STAGES) & 1u;
					g_mbar_wait(bar_a_empty + 8 * s, ph ^ 1u);
					g_mbar_expect_tx(bar_a_full + 8 * s, H_ASTAGE);
					g_bulk_g2s(base + A0 + s * H_ASTAGE, g.Aimg + (size_t)t * H_ASTAGE, H_ASTAGE, bar_a_full + 8 * s);
				}
			}
		}
		__syncwarp();
	} else if (warp == EW) {
		if (lane == 0) {
			constexpr uint32_t ID256 = g_idesc(G_M, 256);
			uint32_t a_it = 0, acc_it = 0, unit_it = 0;
			for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++unit_it) {
				const int nh = min(NH, n_half - u * NH);
				g_mbar_wait(bar_b_full, unit_it & 1u);
				g_tc_fence_after();
				for (int t = 0; t < g.ntiles; ++t, ++a_it) {
					const uint32_t s = a_it % H_STAGES;
					g_mbar_wait(bar_a_full + 8 * s, (a_it / H_STAGES) & 1u);
					g_tc_fence_after();
					const uint32_t A = base + A0 + s * H_ASTAGE;
					for (int hf = 0; hf < nh; ++hf, ++acc_it) {
						const uint32_t buf = acc_it & 1u;
						g_mbar_wait(bar_t_empty + 8 * buf, ((acc_it >> 1) & 1u) ^ 1u);
						g_tc_fence_after();
						const uint32_t dcol = tmem + buf * 256u, B = base + hf * H_BBYTES;
#pragma unroll
						for (int k = 0; k < 4; ++k) // hi*hi, hi*lo, lo*hi, the folded offsets: four K-steps of the same rows
							g_mma_f16(dcol, g_desc(A + k * 32), g_desc(B + k * 32), ID256, k != 0);
						g_tc_commit(bar_t_full + 8 * buf);
					}
					g_tc_commit(bar_a_empty + 8 * s);
				}
				g_tc_commit(bar_b_empty);
			}
		}
		__syncwarp();
	} else {
		// epilogue: thread = step of the tile; a warp takes PPW pairs of slots
		const int wq = warp & 3, eg = warp >> 2;
		const int row = wq * 32 + lane;
		uint32_t acc_it = 0;
		for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
			const int nh = min(NH, n_half - u * NH);
			asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
			{
				const float *src = g.Bconst + (size_t)u * NH * H_NS * H_CONST;
				for (int i = threadIdx.x; i < nh * H_NS * H_CONST; i += EW * 32) econst[i] = __ldg(src + i);
			}
			asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
			unsigned folded_mask = 0u; // bit hf: every slot of that half has its offsets inside the GEMM (short epilogue)
			for (int hf = 0; hf < nh; ++hf) {
				bool f = true;
#pragma unroll
				for (int sl = 0; sl < H_NS; ++sl) f = f && econst[(hf * H_NS + sl) * H_CONST + HD + 2] != 0.0f;
				folded_mask |= f ? (1u << hf) : 0u;
			}
			for (int t = 0; t < g.ntiles; ++t) {
				for (int hf = 0; hf < nh; ++hf, ++acc_it) {
					const int half = u * NH + hf;
					float *Lc = g.L + ((size_t)(half >> 1) * g.BS) * 32 + (half & 1) * H_NS + eg * PPW * 2;
					const bool folded = (folded_mask >> hf) & 1u;
					const uint32_t buf = acc_it & 1u;
					g_mbar_wait(bar_t_full + 8 * buf, (acc_it >> 1) & 1u);
					g_tc_fence_after();
					const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + buf * 256u + eg * PPW * 32;
					float out[PPW * 2];
#pragma unroll
					for (int pp = 0; pp < PPW; pp += 2) {
						float v0[32], v1[32];
						g_tmem_ld32_nowait(taddr + pp * 32u, v0);
						g_tmem_ld32_nowait(taddr + (pp + 1) * 32u, v1);
						g_tmem_wait_ld(v0, v1);
						if (pp + 2 >= PPW) { // the accumulator is in registers: hand the buffer back before the arithmetic
							g_tc_fence_before();
							g_mbar_arrive(bar_t_empty + 8 * buf);
						}
#pragma unroll
						for (int h = 0; h < 4; ++h) { // slots 2 pp .. 2 pp + 3 of this warp's range
							const float(&v)[32] = h < 2 ? v0 : v1;
							const int o = (h & 1) * 16;
							const float *ec = econst + (hf * H_NS + eg * PPW * 2 + pp * 2 + h) * H_CONST;
							float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
							if (folded) {
#pragma unroll
								for (int i = 0; i < 16; i += 4) {
									q0 = fmaf(v[o + i], v[o + i], q0); q1 = fmaf(v[o + i + 1], v[o + i + 1], q1);
									q2 = fmaf(v[o + i + 2], v[o + i + 2], q2); q3 = fmaf(v[o + i + 3], v[o + i + 3], q3);
								}
								out[pp * 2 + h] = fmaf(-ec[HD + 3], (q0 + q1) + (q2 + q3), ec[HD]);
							} else {
								const float dsc = ec[HD + 1];
#pragma unroll
								for (int i = 0; i < 16; i += 4) {
									const float4 nb = *reinterpret_cast<const float4 *>(ec + i);
									const float a0 = fmaf(v[o + i], dsc, nb.x), a1 = fmaf(v[o + i + 1], dsc, nb.y), a2 = fmaf(v[o + i + 2], dsc, nb.z),
											    a3 = fmaf(v[o + i + 3], dsc, nb.w);
									q0 = fmaf(a0, a0, q0); q1 = fmaf(a1, a1, q1); q2 = fmaf(a2, a2, q2); q3 = fmaf(a3, a3, q3);
								}
								out[pp * 2 + h] = ec[HD] - ((q0 + q1) + (q2 + q3));
							}
						}
					}
					float4 *dst = reinterpret_cast<float4 *>(Lc + (size_t)(t * G_M + row) * 32);
#pragma unroll
					for (int i = 0; i < PPW * 2; i += 4) dst[i / 4] = make_float4(out[i], out[i + 1], out[i + 2], out[i + 3]);
				}
			}
		}
	}
	g_tc_fence_before();
	__syncthreads();
	if (warp == EW) {
		g_tc_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
	}
}

// the race as a kernel of its own: four chains per CTA, 64 registers, eight CTAs per SM, the tile loaded when needed instead of
// prefetched into registers.  (Round 1: 128 registers with the prefetch, four CTAs per SM.  In a chain that keeps moving items
// the race is a sequential pass of ~52 dependent instructions per step and was issuing 49 % of the time at 4 warps per scheduler;
// twice the warps: 165 -> 144 ms per sweep at 15 % moved, 97 -> 95 ms with the chains stationary; 10 or 12 CTAs per SM spill.)
#ifndef NPB_RACE_CTAS
#define NPB_RACE_CTAS 8
#endif
template <int CD, int M, int AM>
__global__ void __launch_bounds__(128, NPB_RACE_CTAS) k_race(const PreArgs p) {
	__shared__ float sm[4][G_CONS_FLOATS];
	const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int chain = blockIdx.x * 4 + w;
	if (chain >= p.a.C) return;
	g_consume_chain<CD, M, false, AM>(p, chain, lane, sm[w], sm[w] + 32 * 33, sm[w] + 32 * 33 + 32);
}

extern template npb_status npb_launch_aux_keys<16>(npb_chains *, const SweepArgs &);

static int h_block_steps(const npb_chains *ch) {
	int v = ch->sw.d16_block; // default 8192 (4096 in round 1: the slots' operand images, 537 MB at 8192 chains, are re-read from HBM once per block)
	if (v < 128) v = 128;
	if (v > (1 << 16)) v = 1 << 16;
	return (v + 127) & ~127;
}

// buffers of the D = 16 tensor paths; the [chain][step][32] table only for the two-kernel path (NPB_D16_PATH=tc2)
npb_status npb_tc16_ensure(npb_chains *ch, bool need_table) {
	npb_ctx *ctx = ch->ctx;
	npb_dataset *ds = ch->ds;
	if (!ds->Xbar) NPB_CUDA_OK(cudaMalloc((void **)&ds->Xbar, sizeof(double) * (2 * HD + 1)));
	if (!ds->xbar_valid) {
		ds->xbar_valid = true;
		k_colmean<<<HD, 256, 0, ctx->stream>>>(ds->X64, ds->N, HD, ds->Xbar);
		NPB_CUDA_OK(cudaGetLastError());
		k_xscale<<<1, 1, 0, ctx->stream>>>(ds->Xbar, HD);
		NPB_CUDA_OK(cudaGetLastError());
	}
	const size_t C = (size_t)ch->C;
	if (!ch->g_aimg) {
		int BS = h_block_steps(ch);
		if ((int64_t)BS > ((ds->N + 127) & ~(int64_t)127)) BS = (int)((ds->N + 127) & ~(int64_t)127); // no larger than the sweep
		NPB_CUDA_OK(cudaMalloc((void **)&ch->g_aimg, (size_t)(BS / G_M) * H_ASTAGE));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->g_bimg, (C + 1) * 32 * H_SLOT_IMG));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->g_bconst, (C + 1) * 32 * H_CONST * sizeof(float)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->g_dirty, C * 32));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->g_born, 2 * C * sizeof(uint32_t)));
		NPB_CUDA_OK(cudaMemsetAsync(ch->g_born, 0, 2 * C * sizeof(uint32_t), ctx->stream));
		ch->g_bs = BS;
		NPB_CUDA_OK(cudaFuncSetAttribute(k_density_tc16<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM));
		NPB_CUDA_OK(cudaFuncSetAttribute(k_density_tc16<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM));
		NPB_CUDA_OK(cudaFuncSetAttribute(k_density_tc16<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM));
		NPB_CUDA_OK(cudaFuncSetAttribute(k_density_tc16<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM));
	}
	if (need_table && !ch->g_L) {
		const int BS = ch->g_bs;
		NPB_CUDA_OK(cudaMalloc((void **)&ch->g_L, C * (size_t)(BS + 32) * 32 * sizeof(float)));
		NPB_CUDA_OK(cudaMemsetAsync(ch->g_L, 0, C * (size_t)(BS + 32) * 32 * sizeof(float), ctx->stream));
	}
	return NPB_OK;
}
static npb_status h_ensure(npb_chains *ch) { return npb_tc16_ensure(ch, true); }

// operand images of one block of steps: A tiles of the items in scan order, B images of the slots marked dirty
npb_status npb_tc16_pre_block(npb_chains *ch, const int32_t *d_order, int nsteps, int born_buf) {
	npb_ctx *ctx = ch->ctx;
	const int C = (int)ch->C;
	const int ntiles = (nsteps + G_M - 1) / G_M;
	k_pre_aimg16<<<(ntiles * G_M * 2 + 255) / 256, 256, 0, ctx->stream>>>(ch->ds->X64, ch->ds->Xbar, d_order, nsteps, ntiles, ch->g_aimg);
	NPB_CUDA_OK(cudaGetLastError());
	k_pre_bimg16<<<(C * 32 + 7) / 8, 256, 0, ctx->stream>>>(ch->theta, ch->ds->Xbar, ch->g_dirty, ch->g_born + (size_t)born_buf * C, ch->g_bimg,
			ch->g_bconst, C * 32);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

static npb_status h_density_block(npb_chains *ch, const int32_t *d_order, int nsteps, int born_buf) {
	npb_ctx *ctx = ch->ctx;
	const int C = (int)ch->C;
	const int ntiles = (nsteps + G_M - 1) / G_M;
	npb_status ps = npb_tc16_pre_block(ch, d_order, nsteps, born_buf);
	if (ps != NPB_OK) return ps;
	GemmArgs g;
	g.Aimg = ch->g_aimg;
	g.Bimg = ch->g_bimg;
	g.Bconst = ch->g_bconst;
	g.L = ch->g_L;
	g.C = C;
	g.ntiles = ntiles;
	g.BS = ch->g_bs + 32;
	const int n_sm = ctx->n_sm;
	const int nh = ch->sw.d16_nh;
	const int n_units = (C * 2 + nh - 1) / nh;
	const int grid = n_units < n_sm ? n_units : n_sm;
	if (nh == 1) k_density_tc16<8, 1><<<grid, 8 * 32 + 64, H_SMEM, ctx->stream>>>(g);
	else if (nh == 2) k_density_tc16<8, 2><<<grid, 8 * 32 + 64, H_SMEM, ctx->stream>>>(g);
	else if (ch->sw.d16_epi == 8) k_density_tc16<8, 4><<<grid, 8 * 32 + 64, H_SMEM, ctx->stream>>>(g);
	else k_density_tc16<16, 4><<<grid, 16 * 32 + 64, H_SMEM, ctx->stream>>>(g);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

npb_status npb_launch_alg8_tc16(npb_chains *ch, const SweepArgs &a) {
	npb_ctx *ctx = ch->ctx;
	if (ch->m_aux != 3 && ch->m_aux != 1) return npb_fail(ctx, NPB_E_UNSUPPORTED, "m_aux must be 1 or 3 for the D = 16 tensor-core sweep");
	npb_status s = h_ensure(ch);
	if (s != NPB_OK) return s;
	// NPB_D16_AUX=lazy drops the k_aux_keys pre-pass (16 % of the step): the race then bounds the auxiliary keys from the
	// draws' first normals and evaluates them on demand (g_aux_bound / g_aux_exact; same keys, same results, 3.2 GB less
	// memory).  Measured equal (101 against 102 ms per sweep): the bound leaves out the chi-square term, which is what makes
	// an auxiliary draw hopeless at D = 16, so a third of the steps still need the exact key and every tile pays for it.
	const bool prepass = ch->sw.d16_aux_pre;
	// (same keys, same results in every mode; which one is a matter of time only)
	const bool grp = ch->sw.d16_aux_grp && !(ch->sw.d16_aux_auto && ch->moved_frac_last > 0.01);
	if (prepass) {
		s = npb_launch_aux_keys<16>(ch, a);
		if (s != NPB_OK) return s;
	} else if (grp) { // group maxima of a bound (k_aux_bound), exact keys on demand: the default while few items move
		s = npb_launch_aux_bound<16>(ch, a);
		if (s != NPB_OK) return s;
	}
	const size_t C = (size_t)ch->C;
	NPB_CUDA_OK(cudaMemsetAsync(ch->g_dirty, 1, C * 32, ctx->stream));
	const int BS = ch->g_bs, N = a.N;
	PreArgs p;
	p.a = a;
	if (!prepass) p.a.aux_keys = nullptr;
	if (prepass || !grp) p.a.aux_max = nullptr;
	p.BS = BS + 32;
	p.L = ch->g_L;
	p.spec = ch->sw.spec != 0;
	for (int sw = 0; sw < a.n_sweeps; ++sw) {
		for (int s0 = 0; s0 < N; s0 += BS, ++ch->g_k) {
			const int nsteps = N - s0 < BS ? N - s0 : BS;
			const int buf = (int)(ch->g_k & 1u);
			s = h_density_block(ch, a.scan_order + (size_t)sw * N + s0, nsteps, buf);
			if (s != NPB_OK) return s;
			p.born_prev = ch->g_born + (size_t)(buf ^ 1) * C;
			p.born_out = ch->g_born + (size_t)buf * C;
			p.sw = sw;
			p.s0 = s0;
			p.nsteps = nsteps;
			const unsigned blocks = (unsigned)((ch->C + 3) / 4);
			const int am = prepass ? 0 : (grp ? 2 : 1);
			if (ch->m_aux == 3 && am == 0) k_race<HD, 3, 0><<<blocks, 128, 0, ctx->stream>>>(p);
			else if (ch->m_aux == 3 && am == 1) k_race<HD, 3, 1><<<blocks, 128, 0, ctx->stream>>>(p);
			else if (ch->m_aux == 3) k_race<HD, 3, 2><<<blocks, 128, 0, ctx->stream>>>(p);
			else if (am == 0) k_race<HD, 1, 0><<<blocks, 128, 0, ctx->stream>>>(p);
			else if (am == 1) k_race<HD, 1, 1><<<blocks, 128, 0, ctx->stream>>>(p);
			else k_race<HD, 1, 2><<<blocks, 128, 0, ctx->stream>>>(p);
			NPB_CUDA_OK(cudaGetLastError());
		}
	}
	return NPB_OK;
}

npb_status npb_launch_tc16_probe(npb_chains *ch, int chain, const int32_t *d_items, float *d_out) {
	npb_ctx *ctx = ch->ctx;
	npb_status s = h_ensure(ch);
	if (s != NPB_OK) return s;
	NPB_CUDA_OK(cudaMemsetAsync(ch->g_dirty, 1, (size_t)ch->C * 32, ctx->stream));
	s = h_density_block(ch, d_items, 32, 0);
	if (s != NPB_OK) return s;
	k_gemm64_probe_out<<<32, 32, 0, ctx->stream>>>(ch->g_L, ch->counts, chain, ch->g_bs + 32, d_out);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}
