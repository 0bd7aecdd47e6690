// npb_alg8_tile4.cuh -- Algorithm 8 sweeps for D = 4, 8, 16 and Kmax = 32: four chains per CTA, register budget moved
// from the consumer warps to the producer warps with setmaxnreg, packed FP32 in the producer, state-independent work
// hoisted into a pre-pass kernel.
//
// Same algorithm as k_alg8_sweep_tile (npb_alg8_tile.cuh), a different mapping.  ncu on the two-warp CTA showed the FP32
// pipe 30 % busy with the schedulers issuing on 44 % of the cycles: every thread of the kernel was allocated the ~200
// registers the PRODUCER needs to keep one slot's parameters resident, so only 8 warps fitted an SM (two per scheduler)
// and the dependent chain of the consumer's sequential race could not be hidden.  Here
//   * warps 0-3 (one warpgroup) are the CONSUMERS of chains 4b .. 4b+3, warps 4-7 their PRODUCERS;
//   * the kernel is launched at 128 registers per thread, two CTAs per SM; the consumer warpgroup shrinks to 64
//     registers (setmaxnreg.dec) and the producer warpgroup grows to 192 (setmaxnreg.inc): 16 resident warps per SM
//     instead of 8 out of the same register file;
//   * the producer evaluates two items per instruction stream with fma.rn.f32x2 (FFMA2), the slot parameter being a
//     broadcast operand: 88 instructions per (item, 32 slots) instead of 181;
//   * the auxiliary draws of a step depend on (chain, step, sweep, item) only, so k_aux_keys computes the race key of
//     their best for every (chain, step) before the sweep; the sweep kernel reads one packed word per step;
//   * the master copy of the slot table stays in global memory (L2): producers read their slot once per launch and
//     after a birth, so shared memory only holds the two [slot x step] tiles, the staged item rows and the versions
//     (11 KB per chain);
//   * one rendezvous barrier per (chain, buffer): the producer arrives after filling buffer b, the consumer before
//     draining it; by the time the producer comes back to buffer b it has passed the other buffer's rendezvous, which
//     the consumer reaches only after draining b.
// Measured history and what bounds it now: profiles/README.md.
#pragma once
#include "npb_alg8_tile.cuh"

#define NPB_T4_CHAINS 4
#define NPB_T4_CONSUMER_REGS 64
#define NPB_T4_PRODUCER_REGS 192

template <int D>
struct Tile4Smem {
	float tile[2][32 * 33];   // [buffer][slot * 33 + step]
	float xs[NPB_TILE * D];   // item rows of the tile the producer is working on, pair-interleaved:
	                          // xs[(p * D + c) * 2 + h] = coordinate c of item 2p + h
	int ver_tile[2][32];      // version of slot k the buffer's column was computed from, -1 = not computed
	int ver_cur[32];          // current slot versions (bumped by a birth)
	unsigned occ;             // occupancy bit mask, maintained by the consumer
	unsigned pad[3];
};

// -log2 E' as neg_lg2_exp1 (npb_alg8_kernel.cuh) without the clamp: E' = 0 (probability 2^-25) gives +inf, a certain
// win, which is what an exponential that small means; 1 - v = 0 (probability 2^-32) gives -inf, a certain loss.
__device__ __forceinline__ float neg_lg2_exp1_open(uint32_t r) {
	const float omv = fmaf(__uint2float_rn(r), -2.3283064365386963e-10f, 1.0f - 2.3283064365386963e-10f);
	return -fast_lg2(-fast_lg2(omv));
}

__device__ __forceinline__ float redux_max_f32(float v) {
	float r;
	asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
	return r;
}

// Packed FP32 (sm_100a FFMA2): one instruction does two FMAs, on a 64-bit register pair; an operand written as
// {t, t} is encoded as a broadcast of the 32-bit register t (SASS "R.F32"), so the slot's parameters are not duplicated.
// The producer evaluates TWO items per instruction stream: pair = (item 2p, item 2p+1).
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_bcast(float t) {
	f32x2_t r;
	// volatile: keeps the pack next to its use, where ptxas folds it into the FFMA2 operand (hoisted out of the item
	// loop it would be materialised as a register PAIR per parameter, i.e. twice the registers)
	asm volatile("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(t));
	return r;
}
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
	f32x2_t r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
	f32x2_t r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
	return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float &lo, float &hi) {
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// log2 N(x | theta) of the two items of a pair.  Q[]: nb[D] = -(T2 mu), T2 packed upper row-wise, c2 -- so that
// y = T2 x - T2 mu needs no subtraction per item: row r starts from the broadcast -b[r] at its diagonal column.
// Column-major accumulation: D independent chains; x2 = D packed coordinates {x_2p[c], x_2p+1[c]} in shared memory.
template <int D>
__device__ __forceinline__ void log2density_pair(const float (&Q)[npb_psp(D)], const f32x2_t *x2, float &l0, float &l1) {
	constexpr int TRI = npb_tri(D);
	f32x2_t y[D];
#pragma unroll
	for (int c = 0; c < D; c += 2) {
		const ulonglong2 xx = *reinterpret_cast<const ulonglong2 *>(x2 + c); // coordinates c and c+1 of both items
#pragma unroll
		for (int h = 0; h < 2; ++h) {
			const int cc = c + h;
			const f32x2_t xc = h ? xx.y : xx.x;
#pragma unroll
			for (int r = 0; r < cc; ++r) y[r] = f2_fma(f2_bcast(Q[D + npb_tri_off(D, r, cc)]), xc, y[r]);
			y[cc] = f2_fma(f2_bcast(Q[D + npb_tri_off(D, cc, cc)]), xc, f2_bcast(Q[cc]));
		}
	}
	f32x2_t q0 = f2_bcast(0.0f), q1 = f2_bcast(0.0f);
#pragma unroll
	for (int r = 0; r < D; r += 2) {
		q0 = f2_fma(y[r], y[r], q0);
		q1 = f2_fma(y[r + 1], y[r + 1], q1);
	}
	float a0, a1, b0, b1;
	f2_unpack(q0, a0, a1);
	f2_unpack(q1, b0, b1);
	l0 = Q[D + TRI] - (a0 + b0);
	l1 = Q[D + TRI] - (a1 + b1);
}

// slot parameters in the producer's form: mu replaced by nb = -(T2 mu)
template <int D>
__device__ __forceinline__ void load_theta_nb(const float *thg, float (&Q)[npb_psp(D)]) {
	constexpr int PS = npb_ps(D);
#pragma unroll
	for (int e = 0; e < npb_psp(D); ++e) Q[e] = e < PS ? __ldcg(thg + e) : 0.0f;
	float nb[D];
#pragma unroll
	for (int r = 0; r < D; ++r) {
		float s = 0.0f;
#pragma unroll
		for (int c = r; c < D; ++c) s = fmaf(Q[D + npb_tri_off(D, r, c)], Q[c], s);
		nb[r] = -s;
	}
#pragma unroll
	for (int r = 0; r < D; ++r) Q[r] = nb[r];
}

// log2 N(x | theta) with theta streamed from global memory (consumer side, rare: a newborn slot's column)
template <int D>
__device__ __noinline__ float log2density_stream(const float *th, const float *xrow) {
	constexpr int TRI = npb_tri(D);
	float q = 0.0f;
	for (int r = 0; r < D; ++r) {
		float y = 0.0f;
		for (int c = r; c < D; ++c) y = fmaf(__ldcg(th + D + npb_tri_off(D, r, c)), __ldg(xrow + c) - __ldcg(th + c), y);
		q = fmaf(y, y, q);
	}
	return __ldcg(th + D + TRI) - q;
}

// One auxiliary draw theta' ~ G0 in the form the race needs (np_neal_algorithm8.cpp:79-84; dirichlet.h:91-93 ->
// normalinvwishart.h:44-64).  In prior-whitened coordinates theta' = (|v|, z), v ~ N(D, nu^2), z ~ N(0, I_D), and
//     log2 N(x | theta') = c0_2 - D log2|v| - |xw/|v| - s z|^2,      s = ik2            (npb_common.cuh)
// The density depends on z only through |a - s z|^2 with a = xw/|v|.  Split z along a: z = zpar a^ + zperp, then
//     |a - s z|^2 = (|a| - s zpar)^2 + s^2 |zperp|^2,   zpar ~ N(0,1),  |zperp|^2 ~ chi^2_{D-1}, direction uniform,
// all three independent.  So a draw's key needs THREE normals and (D-1)/2 uniforms instead of D+1 normals:
// chi^2_{2k+1} = -2 ln(U_1 ... U_k) + z2^2.  Only when a draw wins the race (a birth, ~1e-5 of the steps) is the full
// vector z materialised: zpar a^ + sqrt(R2) u^, u^ uniform on the unit sphere orthogonal to a^ (aux_birth_z below) --
// the joint law of (key, theta') is the reference's.  Stream per draw: 4 words (two Box-Muller pairs: v, zpar, z2,
// ) + ceil((D-1)/4) words of 16-bit uniforms.
template <int D>
__device__ __forceinline__ void aux_draw_chi(uint32_t (&as)[4], const PriorDev &pr, float &av, float &zpar, float &R2) {
	// word order of a draw's stream: the Box-Muller pair of (v, z_par), the packed uniforms of the chi-square, the pair of
	// z2^2 -- so that a bound of the draw's key (aux_draw_bound) needs the first four words only
	float g0, g1;
	{
		const uint32_t r0 = xoshiro_next(as), r1 = xoshiro_next(as);
		npb_normal2(r0, r1, g0, g1);
	}
	av = fmaxf(fabsf(pr.v_mean + pr.nu * g0), 1e-20f);
	zpar = g1;
	constexpr int KU = (D - 1) / 2; // uniforms; D <= 16 keeps their product far above FLT_MIN
	float prod = 1.0f, lsum = 0.0f;
#pragma unroll
	for (int i = 0; i < KU; i += 2) {
		const uint32_t w = xoshiro_next(as);
		prod *= __uint2float_rn((w & 0xffffu) + 1u) * (1.0f / 65536.0f);
		if (i + 1 < KU) prod *= __uint2float_rn((w >> 16) + 1u) * (1.0f / 65536.0f);
		if (KU > 7 && (i % 6) == 4) { // D > 16: one logarithm per six uniforms (their product stays above 2^-96)
			lsum += fast_lg2(prod);
			prod = 1.0f;
		}
	}
	R2 = -2.0f * NPB_LN2 * (lsum + fast_lg2(prod));
	if ((D - 1) & 1) {
		// z2^2 alone: the square of one Box-Muller output, (-2 ln u) cos^2(2 pi v), needs no square root and no sine
		const uint32_t r0 = xoshiro_next(as), r1 = xoshiro_next(as);
		const float c = __cosf(__uint2float_rn(r1) * (6.283185307179586f * 2.3283064365386963e-10f));
		R2 += -2.0f * NPB_LN2 * fast_lg2(npb_u01(r0)) * c * c;
	}
}

// Upper bound of the race key of one draw for one item from the first four words of its stream: v, z_par and the first
// four uniforms of the chi-square (-2 ln(U1 U2 U3 U4) <= R2), the race noise at its cap (neg_lg2_exp1 clamps at 23) and one
// more for rounding and the index bits packed into a key.  At the reference prior the dropped part of the chi-square is
// what makes the bound loose (by ~361 chi^2_7 log2 units at D = 16) and it does not matter: what is kept already puts a
// draw thousands of log2 units below any item's own cluster.
template <int D>
__device__ __forceinline__ float aux_draw_bound(uint32_t (&as)[4], const PriorDev &pr, float rn, float ik2) {
	// (single-instruction approximations of the square root and the reciprocal: their 2^-22 relative error moves the bound by
	// ~0.02 log2 units at most, inside the slack below; aux_draw_chi's IEEE versions cost two slow-path branches per draw)
	float g0, g1;
	{
		const uint32_t r0 = xoshiro_next(as), r1 = xoshiro_next(as);
		float rad, s, c;
		asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-2.0f * NPB_LN2 * __log2f(npb_u01(r0))));
		__sincosf(__uint2float_rn(r1) * (6.283185307179586f * 2.3283064365386963e-10f), &s, &c);
		g0 = rad * c;
		g1 = rad * s;
	}
	const float av = fmaxf(fabsf(pr.v_mean + pr.nu * g0), 1e-20f);
	constexpr int KU = (D - 1) / 2;
	float prod = 1.0f;
#pragma unroll
	for (int i = 0; i < (KU < 4 ? KU : 4); i += 2) {
		const uint32_t w = xoshiro_next(as);
		prod *= __uint2float_rn((w & 0xffffu) + 1u) * (1.0f / 65536.0f);
		if (i + 1 < KU) prod *= __uint2float_rn((w >> 16) + 1u) * (1.0f / 65536.0f);
	}
	const float lb = -2.0f * NPB_LN2 * fast_lg2(prod);
	float rav;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rav) : "f"(av));
	const float along = rn * rav - ik2 * g1;
	return pr.c0_2 - (float)D * fast_lg2(av) - fmaf(along, along, ik2 * ik2 * lb) + pr.log2_alpha_m + 24.0f;
}

// xoshiro state of the auxiliary stream of (chain, step, sweep): four words of a multiply-xorshift hash (the murmur3
// finaliser, full avalanche) of the chain's Philox key and the counters -- 36 instructions against Philox4x32-10's 70,
// in a stream that only has to be distinct per (chain, step, sweep) and reproducible at a birth.
__device__ __forceinline__ void aux_seed(const Philox &ph, uint32_t sj, uint32_t sweep, uint32_t (&as)[4]) {
	uint32_t h = npb_mix32(ph.k0 ^ (sj * 0x9E3779B1u));
	h = npb_mix32(h ^ ph.k1 ^ (sweep * 0x85EBCA77u));
	as[0] = h;
	as[1] = npb_mix32(h + 0x9E3779B9u);
	as[2] = npb_mix32(h + 0x3C6EF372u);
	as[3] = npb_mix32(h + 0xDAA66D2Bu) | 1u; // never the all-zero state
}

// stream of draw m: the step's state with the draw index hashed into two of its words
__device__ __forceinline__ void aux_stream(const uint32_t (&base)[4], int m, uint32_t (&as)[4]) {
	as[0] = npb_mix32(base[0] + (uint32_t)(m + 1) * 0x9E3779B9u);
	as[1] = npb_mix32(base[1] ^ ((uint32_t)(m + 1) * 0x7F4A7C15u));
	as[2] = base[2];
	as[3] = base[3];
}

// The M auxiliary draws of step sj for one item and the race among them: best key and which draw it was.
// rn = |xw| (k_whiten).
template <int D, int M>
__device__ __forceinline__ void aux_race(const Philox &ph, const PriorDev &pr, float rn, uint32_t sj, uint32_t sweep, float ik2,
		float &auxkey_j, int &auxm) {
	auxkey_j = -INFINITY;
	auxm = 0;
	uint32_t base[4];
	aux_seed(ph, sj, sweep, base);
#pragma unroll
	for (int m = 0; m < M; ++m) {
		uint32_t as[4];
		aux_stream(base, m, as);
		float av, zpar, R2;
		aux_draw_chi<D>(as, pr, av, zpar, R2);
		const float along = rn * __frcp_rn(av) - ik2 * zpar;
		const float q = fmaf(along, along, ik2 * ik2 * R2);
		const float key = (pr.c0_2 - (float)D * fast_lg2(av) - q + pr.log2_alpha_m) + neg_lg2_exp1(xoshiro_next(as));
		if (key > auxkey_j) { auxkey_j = key; auxm = m; }
	}
}

// upper bound of the packed race key of the best of step sj's M draws (aux_draw_bound): what k_aux_bound reduces per group
template <int D, int M>
__device__ __forceinline__ float aux_race_bound(const Philox &ph, const PriorDev &pr, float rn, uint32_t sj, uint32_t sweep, float ik2) {
	uint32_t base[4];
	aux_seed(ph, sj, sweep, base);
	float ub = -INFINITY;
#pragma unroll
	for (int m = 0; m < M; ++m) {
		uint32_t as[4];
		aux_stream(base, m, as);
		ub = fmaxf(ub, aux_draw_bound<D>(as, pr, rn, ik2));
	}
	return ub;
}

// Birth: coordinate `lane` (< D) of the full z of draw m of step `step`, consistent with the key that won the race.
// All lanes of the warp call it; xw is the item's whitened row, rn its norm.
template <int D>
__device__ __forceinline__ float aux_birth_z(const Philox &ph, const PriorDev &pr, const float *xw, float rn, uint32_t step,
		uint32_t sweep, int m, int lane, float &av_out) {
	uint32_t base[4], as[4];
	aux_seed(ph, step, sweep, base);
	aux_stream(base, m, as);
	float av = 1.0f, zpar = 0.0f, R2 = 0.0f;
	aux_draw_chi<D>(as, pr, av, zpar, R2);
	av_out = av;
	// unit vector along the item (any fixed direction if the item sits exactly on mu0)
	const float ahat = lane < D ? (rn > 0.0f ? __ldg(xw + lane) / rn : (lane == 0 ? 1.0f : 0.0f)) : 0.0f;
	// a standard normal vector from its own stream, made orthogonal to a^ and normalised
	uint32_t w[4];
	ph(step, 2u + (uint32_t)(lane >> 1), sweep, NPB_RNG_AUX, w);
	float n0, n1;
	npb_normal2(w[0], w[1], n0, n1);
	float g = lane < D ? ((lane & 1) ? n1 : n0) : 0.0f;
	float dot = g * ahat;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
	g -= dot * ahat;
	float nn = g * g;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
	const float uhat = nn > 0.0f ? g * rsqrtf(nn) : 0.0f;
	return zpar * ahat + sqrtf(R2) * uhat;
}

// Pre-pass: the auxiliary draws of a step do not depend on the chain's state, only on (chain, step, sweep, item), so their
// race is run for every (chain, step) of a launch by a plain data-parallel kernel before the sweep: one thread per
// (chain, step), one packed 32-bit word out ([sweep][chain][step], coalesced along steps for writer and reader).  This
// takes the 3 x (3 normals + 7 uniforms) per step off the sweep kernel's producer warp, whose FFMA2 stream then owns the
// FP32 pipe; the extra HBM traffic (8 bytes per reassignment) is ~1 ms of a 200 ms sweep.
template <int D, int M>
__global__ void __launch_bounds__(256) k_aux_keys(const SweepArgs a, uint32_t *out) {
	const int sj = blockIdx.x * 256 + threadIdx.x;
	const int sw = blockIdx.z;
	const bool in = sj < a.N;
	if (!in) return;
	const int item = a.scan_order[(size_t)sw * a.N + sj];
	const float rn = __ldg(a.Xwn + item);
	for (int chain = blockIdx.y; chain < a.C; chain += gridDim.y) { // grid.y is capped at 65535
		const Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
		float ak;
		int am;
		aux_race<D, M>(ph, a.prior, rn, (uint32_t)sj, a.sweep0 + (uint32_t)sw, a.prior.inv_sqrt_kappa * (float)NPB_HALF_LOG2E_SQRT, ak, am);
		if (in) out[((size_t)sw * a.C + chain) * a.N + sj] = (__float_as_uint(ak) & ~3u) | (uint32_t)am;
	}
}

// The fused D = 16 kernel (npb_alg8_fused16.cu) compares a step's pick with the largest auxiliary key of its 32-step group and
// evaluates a step's own key only when that does not settle it (a handful of steps per million at the reference prior), so its
// pre-pass needs no exact key at all: a BOUND per step (four stream words per draw instead of nine, no race noise), reduced to
// one number per (chain, group).  40 % of k_aux_keys' instructions and none of its 3.3 GB per sweep.
template <int D, int M>
__global__ void __launch_bounds__(256) k_aux_bound(const SweepArgs a) {
	const int sj = blockIdx.x * 256 + threadIdx.x;
	const int sw = blockIdx.z;
	const bool in = sj < a.N;
	const int item = in ? a.scan_order[(size_t)sw * a.N + sj] : 0;
	const float rn = __ldg(a.Xwn + item);
	for (int chain = blockIdx.y; chain < a.C; chain += gridDim.y) {
		const Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
		const float ub = aux_race_bound<D, M>(ph, a.prior, rn, (uint32_t)sj, a.sweep0 + (uint32_t)sw, a.prior.inv_sqrt_kappa * (float)NPB_HALF_LOG2E_SQRT);
		const float gm = redux_max_f32(in ? ub : -INFINITY);
		if ((threadIdx.x & 31) == 0 && in) a.aux_max[((size_t)sw * a.C + chain) * a.aux_groups + (sj >> 5)] = gm;
	}
}

template <int D>
npb_status npb_launch_aux_bound(npb_chains *ch, const SweepArgs &a) {
	npb_ctx *ctx = ch->ctx;
	if (ch->m_aux != 3 && ch->m_aux != 1) return npb_fail(ctx, NPB_E_UNSUPPORTED, "m_aux must be 1 or 3 for the D >= 4 sweep kernel");
	dim3 grid((unsigned)((a.N + 255) / 256), (unsigned)(ch->C < 65535 ? ch->C : 65535), (unsigned)a.n_sweeps);
	if (ch->m_aux == 3) k_aux_bound<D, 3><<<grid, 256, 0, ctx->stream>>>(a);
	else k_aux_bound<D, 1><<<grid, 256, 0, ctx->stream>>>(a);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

template <int D>
npb_status npb_launch_aux_keys(npb_chains *ch, const SweepArgs &a) {
	npb_ctx *ctx = ch->ctx;
	if (ch->m_aux != 3 && ch->m_aux != 1) return npb_fail(ctx, NPB_E_UNSUPPORTED, "m_aux must be 1 or 3 for the D >= 4 sweep kernel");
	dim3 grid((unsigned)((a.N + 255) / 256), (unsigned)(ch->C < 65535 ? ch->C : 65535), (unsigned)a.n_sweeps);
	if (ch->m_aux == 3) k_aux_keys<D, 3><<<grid, 256, 0, ctx->stream>>>(a, ch->aux_keys);
	else k_aux_keys<D, 1><<<grid, 256, 0, ctx->stream>>>(a, ch->aux_keys);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

template <int D, int M>
__global__ void __launch_bounds__(256, 2) k_alg8_sweep_tile4(const SweepArgs a) {
	constexpr int TRI = npb_tri(D), PS = npb_ps(D), PSP = npb_psp(D), KMAX = 32;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31;
	const int wid = threadIdx.x >> 5;
	const bool producer = wid >= NPB_T4_CHAINS;
	const int cl = wid & (NPB_T4_CHAINS - 1);
	Tile4Smem<D> &sm = reinterpret_cast<Tile4Smem<D> *>(smem_raw)[cl];
	const int chain = blockIdx.x * NPB_T4_CHAINS + cl;
	const int N = a.N, C = a.C;
	const int bar0 = 1 + cl * 2; // named barriers 1..8: rendezvous of (chain, buffer)
	const int tiles_per_sweep = (N + NPB_TILE - 1) / NPB_TILE;

	if (producer) {
		asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(NPB_T4_PRODUCER_REGS));
		if (chain >= C) return;
		// =========================== PRODUCER: lane = slot ===========================
		const int k = lane;
		const float *thg = a.theta + ((size_t)chain * KMAX + k) * PS;
		float P[PSP];
		load_theta_nb<D>(thg, P);
		int myver = 0;
		float *xs = sm.xs;
		int t = 0;
		// wait for the consumer's initial occupancy mask / versions
		named_bar_sync(bar0, 64);
		for (int sw = 0; sw < a.n_sweeps; ++sw) {
			const int32_t *order = a.scan_order + (size_t)sw * N;
			for (int ti = 0; ti < tiles_per_sweep; ++ti, ++t) {
				const int b = t & 1;
				const int s0 = ti * NPB_TILE;
				const int cnt = min(NPB_TILE, N - s0);
				{
					const int item = (lane < cnt) ? order[s0 + lane] : 0;
					const float4 *src = reinterpret_cast<const float4 *>(a.X + (size_t)item * D);
					float *dst = xs + (lane >> 1) * (2 * D) + (lane & 1);
#pragma unroll
					for (int c = 0; c < D / 4; ++c) {
						const float4 v = __ldg(src + c);
						dst[(4 * c + 0) * 2] = v.x; dst[(4 * c + 1) * 2] = v.y;
						dst[(4 * c + 2) * 2] = v.z; dst[(4 * c + 3) * 2] = v.w;
					}
				}
				__syncwarp();
				const bool occupied = (*((volatile unsigned *)&sm.occ) >> lane) & 1u;
				const int ver = ((volatile int *)sm.ver_cur)[k];
				__threadfence_block(); // the version is read before the parameters
				if (ver != myver) { // a birth re-used this slot: fetch the new parameters from the master copy
					load_theta_nb<D>(thg, P);
					myver = ver;
				}
				if (occupied) {
					// item pairs (a ragged last pair evaluates one row of padding; its tile entry is never read)
					for (int j = 0; j < cnt; j += 2) {
						float l0, l1;
						log2density_pair<D>(P, reinterpret_cast<const f32x2_t *>(xs + j * D), l0, l1);
						sm.tile[b][k * 33 + j] = l0;
						sm.tile[b][k * 33 + j + 1] = l1;
					}
				}
				sm.ver_tile[b][k] = occupied ? myver : -1;
				__syncwarp();
				__threadfence_block();
				named_bar_sync(bar0 + b, 64); // hand buffer b over
			}
		}
		return;
	}

	asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(NPB_T4_CONSUMER_REGS));
	if (chain >= C) return;
	// =========================== CONSUMER: lane = step (prologue) / lane = slot (steps) ===========================
	const Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
	float *thc = a.theta + (size_t)chain * KMAX * PS;
	float n = (float)a.counts[(size_t)chain * KMAX + lane];
	float lgn = n > 0.0f ? fast_lg2(n) : -INFINITY, lgn1 = n > 1.0f ? fast_lg2(n - 1.0f) : -INFINITY;
	int kocc = __popc(__ballot_sync(0xffffffffu, n > 0.0f));
	sm.ver_cur[lane] = 0;
	sm.ver_tile[0][lane] = -1;
	sm.ver_tile[1][lane] = -1;
	{
		const unsigned b = __ballot_sync(0xffffffffu, n > 0.0f);
		__syncwarp();
		if (lane == 0) sm.occ = b;
	}
	__threadfence_block();
	named_bar_sync(bar0, 64); // initial state published
	unsigned long long st_cand = 0ull, st_moved = 0ull, st_births = 0ull;
	int overflow = 0;

	// re-evaluates the tile column of `slot` for steps >= j_from of buffer b (lane = step) from the master copy
	auto fix_column = [&](int slot, int b, int j_from, int item, bool valid) {
		const float l = log2density_stream<D>(thc + (size_t)slot * PS, a.X + (size_t)item * D);
		if (valid && lane >= j_from) sm.tile[b][slot * 33 + lane] = l;
		__syncwarp();
	};

	int t = 0;
	for (int sw = 0; sw < a.n_sweeps; ++sw) {
		const uint32_t sweep = a.sweep0 + (uint32_t)sw;
		const int32_t *order = a.scan_order + (size_t)sw * N;
		for (int ti = 0; ti < tiles_per_sweep; ++ti, ++t) {
			const int b = t & 1;
			const int s0 = ti * NPB_TILE;
			const int sj = s0 + lane;
			const bool valid = sj < N;
			const int item = valid ? order[sj] : 0;
			const int zold = valid ? (int)a.z[(size_t)item * C + chain] : 0;
			int znew = zold;
			// race noise of the tile: one 32-bit LCG per lane (Knuth's multiplier; the top 24 bits of the state are what the
			// float conversion keeps), re-seeded every tile from a hash of the chain's key and the counters, so a stream is 33
			// draws long.  The
			// consumer's instruction count is what bounds a chain once the producers run packed FP32: xoshiro128++ was 9
			// of its ~64 instructions per step, this is 1.
			uint32_t rs = npb_mix32(npb_mix32(ph.k0 ^ ((uint32_t)sj * 0x9E3779B1u)) ^ ph.k1 ^ (sweep * 0x85EBCA77u) ^ 0x5bd1e995u);
			const int cnt = min(NPB_TILE, N - s0);

			// race key of the best of this step's M auxiliary draws and which draw it was (k_aux_keys, the state-independent
			// pre-pass): the two low mantissa bits carry the draw index and stay in the key, 3 ulp of a noise-dominated number
			const uint32_t auxp = valid ? __ldg(a.aux_keys + ((size_t)sw * C + chain) * N + sj) : 0xff800000u;
			const float auxkey_j = __uint_as_float(auxp);
			const int zold_aux_j = zold | ((int)(auxp & 3u) << 16);

			named_bar_sync(bar0 + b, 64); // the producer has filled buffer b
			{
				unsigned stale = __ballot_sync(0xffffffffu, n > 0.0f && sm.ver_tile[b][lane] != sm.ver_cur[lane]);
				while (stale) {
					const int k = __ffs(stale) - 1;
					stale &= stale - 1;
					fix_column(k, b, 0, item, valid);
				}
			}
			unsigned cand_tile = 0u;

			// The step loop is the critical path of a chain, so everything that does not depend on the previous step is
			// taken off its dependent chain: the tile entry, the race noise and the broadcasts of step j+1 are fetched
			// while step j is decided, and log2 of the member counts (n and n - 1) is kept in registers and refreshed
			// only when a count changes.
			float noise_next = neg_lg2_exp1_open(rs = rs * 1664525u + 1013904223u);
			float base_next = sm.tile[b][lane * 33] + noise_next;
			int zo_aux_next = __shfl_sync(0xffffffffu, zold_aux_j, 0);
			float ak_next = __shfl_sync(0xffffffffu, auxkey_j, 0);
			for (int j = 0; j < cnt; ++j) {
				const int zo_aux = zo_aux_next;
				const float ak = ak_next;
				const float base = base_next;
				const int zo = zo_aux & 0xffff;
				{
					const int jn = min(j + 1, NPB_TILE - 1);
					noise_next = neg_lg2_exp1_open(rs = rs * 1664525u + 1013904223u);
					base_next = sm.tile[b][lane * 33 + jn] + noise_next;
					zo_aux_next = __shfl_sync(0xffffffffu, zold_aux_j, jn);
					ak_next = __shfl_sync(0xffffffffu, auxkey_j, jn);
				}
				const float lg = (zo == lane) ? lgn1 : lgn;
				const float key = lg > -INFINITY ? base + lg : -INFINITY; // a slot without (other) members never wins
				// warp arg-max with the floating-point warp reduction of sm_100a (redux.sync.max.f32 -> CREDUX.MAX.F32)
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
						if (lane == 0) sm.occ &= ~(1u << zo);
					}
					if (born) {
						// np_neal_algorithm8.cpp:136-145: the lowest free slot takes theta' of the winning auxiliary draw
						const unsigned fb = __ballot_sync(0xffffffffu, n <= 0.0f);
						const int fs = fb ? __ffs(fb) - 1 : -1;
						if (fs < 0) {
							overflow = 1; // no room: the item goes back where it was
							if (died) { kocc++; if (lane == 0) sm.occ |= 1u << zo; }
						} else {
							new_slot = fs;
							const int m = (zo_aux >> 16) & 0xff;
							const uint32_t step = (uint32_t)(s0 + j);
							// materialise theta' of draw m of this step, consistent with the key that won (aux_birth_z)
							const int bitem = order[step];
							float av;
							const float zc = aux_birth_z<D>(ph, a.prior, a.Xw + (size_t)bitem * D, __ldg(a.Xwn + bitem), step, sweep, m, lane, av);
							const float g = zc * (av * a.prior.inv_sqrt_kappa);
							float mu_r = lane < D ? a.prior.mu0[lane] : 0.0f;
							for (int c = 0; c < D; ++c) {
								const float gc = __shfl_sync(0xffffffffu, g, c);
								if (lane <= c && lane < D) mu_r = fmaf(a.prior.S[npb_tri_off(D, lane, c)], gc, mu_r);
							}
							float *th = thc + (size_t)fs * PS;
							if (lane < D) th[lane] = mu_r;
							const float inv = 1.0f / av;
							for (int q = lane; q < TRI; q += 32) th[D + q] = a.prior.CT2[q] * inv;
							if (lane == 0) th[D + TRI] = a.prior.c0_2 - (float)D * log2f(av);
							__threadfence();
							__syncwarp();
							if (lane == 0) { // publish only once theta is complete
								sm.ver_cur[fs] += 1;
								__threadfence_block();
								sm.occ |= 1u << fs;
							}
							__syncwarp();
							kocc++;
							st_births++;
							fix_column(fs, b, j + 1, item, valid);
							// the prefetched entry of step j+1 was read before the newborn slot's column existed
							if (lane == fs) base_next = sm.tile[b][fs * 33 + min(j + 1, NPB_TILE - 1)] + noise_next;
						}
					}
					if (new_slot == lane) n += 1.0f;
					lgn = n > 0.0f ? fast_lg2(n) : -INFINITY;
					lgn1 = n > 1.0f ? fast_lg2(n - 1.0f) : -INFINITY;
					st_moved++;
					if (lane == j) znew = new_slot;
				}
			}
			if (valid && znew != zold) a.z[(size_t)item * C + chain] = (npb_z_t)znew;
			__syncwarp();
			st_cand += cand_tile;
			__threadfence_block();
		}
	}

	// ---- chain state back to memory (theta already lives there) ----
	a.counts[(size_t)chain * KMAX + lane] = (int)n;
	if (lane == 0) {
		a.kocc[chain] = kocc;
		if (overflow) a.overflow[chain] = 1;
		a.st[(size_t)chain * 4 + 0] += st_cand;
		a.st[(size_t)chain * 4 + 1] += st_moved;
		a.st[(size_t)chain * 4 + 2] += st_births;
	}
}

template <int D>
npb_status npb_launch_alg8_tile4(npb_chains *ch, const SweepArgs &a) {
	npb_ctx *ctx = ch->ctx;
	if (ch->m_aux != 3 && ch->m_aux != 1) return npb_fail(ctx, NPB_E_UNSUPPORTED, "m_aux must be 1 or 3 for the D >= 4 sweep kernel");
	const size_t shmem = sizeof(Tile4Smem<D>) * NPB_T4_CHAINS;
	const unsigned blocks = (unsigned)((ch->C + NPB_T4_CHAINS - 1) / NPB_T4_CHAINS);
	if (ch->m_aux == 3) {
		NPB_CUDA_OK(cudaFuncSetAttribute(k_alg8_sweep_tile4<D, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
		k_alg8_sweep_tile4<D, 3><<<blocks, 256, shmem, ctx->stream>>>(a);
	} else {
		NPB_CUDA_OK(cudaFuncSetAttribute(k_alg8_sweep_tile4<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
		k_alg8_sweep_tile4<D, 1><<<blocks, 256, shmem, ctx->stream>>>(a);
	}
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

// Parity probe: the [32 slots x 32 items] log-density tile exactly as the producer warp of k_alg8_sweep_tile4 computes
// it (same parameter form, same packed-FP32 instruction sequence), in natural-log units, for one chain and 32 given
// items.  out[slot * 32 + j]; slots without members are written as NaN.
template <int D>
__global__ void __launch_bounds__(32) k_tile4_density_probe(const SweepArgs a, int chain, const int32_t *items, float *out) {
	constexpr int PS = npb_ps(D), PSP = npb_psp(D);
	__shared__ __align__(16) float xs[NPB_TILE * D];
	const int lane = threadIdx.x;
	float P[PSP];
	load_theta_nb<D>(a.theta + ((size_t)chain * 32 + lane) * PS, P);
	{
		const float4 *src = reinterpret_cast<const float4 *>(a.X + (size_t)items[lane] * D);
		float *dst = xs + (lane >> 1) * (2 * D) + (lane & 1);
#pragma unroll
		for (int c = 0; c < D / 4; ++c) {
			const float4 v = __ldg(src + c);
			dst[(4 * c + 0) * 2] = v.x; dst[(4 * c + 1) * 2] = v.y;
			dst[(4 * c + 2) * 2] = v.z; dst[(4 * c + 3) * 2] = v.w;
		}
	}
	__syncwarp();
	const bool occupied = a.counts[(size_t)chain * 32 + lane] > 0;
	for (int j = 0; j < NPB_TILE; j += 2) {
		float l0, l1;
		log2density_pair<D>(P, reinterpret_cast<const f32x2_t *>(xs + j * D), l0, l1);
		out[lane * 32 + j] = occupied ? l0 * NPB_LN2 : NAN;
		out[lane * 32 + j + 1] = occupied ? l1 * NPB_LN2 : NAN;
	}
}

template <int D>
npb_status npb_launch_tile4_probe(npb_chains *ch, const SweepArgs &a, int chain, const int32_t *d_items, float *d_out) {
	npb_ctx *ctx = ch->ctx;
	k_tile4_density_probe<D><<<1, 32, 0, ctx->stream>>>(a, chain, d_items, d_out);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}
