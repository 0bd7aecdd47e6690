// npb_tc_common.cuh -- pieces shared by the tensor-path kernels (npb_alg8_gemm.cu: D = 64 fused kernel and the D = 16
// table + race pair; npb_alg8_fused16.cu: the D = 16 fused sweep kernel): operand-image geometry, mbarrier / bulk-copy /
// tcgen05 PTX wrappers, the counter-hash race noise, and the rare-path helpers (density of one (item, slot) from the slot
// table in global memory, theta' of a birth).
#pragma once
#include "npb_alg8_tile4.cuh"
#include <cstdlib>
#include <cuda_fp16.h>

namespace {
constexpr int GD = 64;
constexpr int GPS = npb_ps(GD);     // 2145
constexpr int GTRI = npb_tri(GD);   // 2080
constexpr int G_M = 128;            // steps per A tile (UMMA M)
constexpr int G_NS = 4;             // slots per unit of work (UMMA N = 64 * 4)
constexpr int G_STAGES = 4;         // A ring (one stage = one tile)
constexpr int G_CHUNK = 4096;       // 32 rows x 128 bytes (64 FP16), swizzled
constexpr int G_SLOT_IMG = 4 * G_CHUNK;   // hi: rows j < 32, rows j >= 32; lo: the same
constexpr int G_ASTAGE = 32768;     // 128 rows x 128 bytes, hi then lo
constexpr int G_BBYTES = 65536;     // Bhi 32K (256 rows), Blo 32K
constexpr int G_CONST = 68;         // nb[64], c2, descale, pad
constexpr int G_XEXP = 14;          // operands are scaled so that their largest magnitude is just below 2^14
constexpr int G_SMEM_MISC = 2048;
constexpr int G_CONS_FLOATS = 32 * 33 + 64; // per race warp: tile, lg_s, lg1_s
constexpr int G_SMEM = 1024 + G_BBYTES + G_STAGES * G_ASTAGE + G_SMEM_MISC + 2 * G_CONS_FLOATS * 4;
constexpr uint32_t G_BHI = 0, G_BLO = 32768, G_A0 = G_BBYTES;
}

struct GemmArgs {
	const uint8_t *Aimg;  // [ntiles][32 KB]
	const uint8_t *Bimg;  // [C * 32][16 KB]
	const float *Bconst;  // [C * 32][G_CONST]
	float *L;             // [C][BS][32]
	int C, ntiles, BS;
};

struct PreArgs {
	SweepArgs a;
	float *L;
	const uint32_t *born_prev; // [C] slots born during the previous block (their columns of L predate them), or NULL
	uint32_t *born_out;        // [C] slots born during this block
	int BS, sw, s0, nsteps;
	int spec;                  // 0: sequential pass only (NPB_D64_SPEC=0; the result must not depend on it)
	int flags;                 // k_sweep_tc16: measurement switches
	const npb_z_t *zblk;       // [C][zstride] k_sweep_tc16: old assignments of the block's steps, in step order (k_gather_z)
	int zstride;
	uint8_t *dirty;            // [C * 32] k_sweep_tc16: set for a slot born during the block (its operand image is stale)
};

// byte offset of FP16 element k (0..63) of row `row` in a K-major, 128-byte-swizzled region (rows of 128 bytes, base
// 1024-aligned)
__host__ __device__ __forceinline__ uint32_t g_sw128(uint32_t row, uint32_t k) {
	return row * 128u + ((((k >> 3) ^ (row & 7u)) << 4) | ((k & 7u) << 1));
}
// v = hi + lo, both FP16 (lo may be subnormal: quantum 2^-24)
__device__ __forceinline__ void g_split(float v, __half &hi, __half &lo) {
	hi = __float2half_rn(v);
	lo = __float2half_rn(v - __half2float(hi));
}
// the power of two that brings a magnitude just below 2^G_XEXP
__device__ __forceinline__ int g_scale_exp(float maxabs) {
	if (!(maxabs > 0.0f) || !isfinite(maxabs)) return 0;
	int m;
	frexpf(maxabs, &m); // maxabs = f 2^m, f in [0.5, 1)
	return G_XEXP - m;
}

// ---------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier, bulk copy, tcgen05
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t g_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_init(uint32_t bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void g_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_arrive(uint32_t bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(uint32_t bar, uint32_t parity) {
	asm volatile(
			"{\n\t"
			".reg .pred p;\n\t"
			"WAIT_LOOP:\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
			"@p bra WAIT_DONE;\n\t"
			"bra WAIT_LOOP;\n\t"
			"WAIT_DONE:\n\t"
			"}\n" ::"r"(bar), "r"(parity)
			: "memory");
}
// the same wait with a back-off between polls: for warps whose wait is long and whose polling would take issue slots from the
// warps they wait for
__device__ __forceinline__ void g_mbar_wait_sleep(uint32_t bar, uint32_t parity, unsigned ns) {
	uint32_t done = 0;
	while (true) {
		asm volatile(
				"{\n\t"
				".reg .pred p;\n\t"
				"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
				"selp.u32 %0, 1, 0, p;\n\t"
				"}\n"
				: "=r"(done)
				: "r"(bar), "r"(parity)
				: "memory");
		if (done) break;
		__nanosleep(ns);
	}
}
__device__ __forceinline__ void g_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
			"r"(bar)
			: "memory");
}
__device__ __forceinline__ void g_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void g_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void g_tc_commit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor:
// start address >> 4 at [0,14), leading byte offset >> 4 at [16,30) (= 1, unused with a swizzle), stride byte offset >> 4 at
// [32,46), version 1 at [46,48), layout type SWIZZLE_128B = 2 at [61,64))
__device__ __forceinline__ uint64_t g_desc(uint32_t saddr) {
	return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor, kind::f16: D = F32 (1 at [4,6)), A = B = F16 (0 at [7,10) and [10,13)), both K-major, N >> 3 at
// [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t g_idesc(int M, int N) {
	return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void g_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
			"{\n\t"
			".reg .pred p;\n\t"
			"setp.ne.b32 p, %4, 0;\n\t"
			"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
			"}\n" ::"r"(tmem_d),
			"l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
			: "memory");
}
// the load alone: the registers are valid after g_tmem_wait_ld, which takes the arrays as in/out operands so that no use
// of them is scheduled above the wait
__device__ __forceinline__ void g_tmem_ld32_nowait(uint32_t taddr, float (&v)[32]) {
	asm volatile(
			"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
			"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
			"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
			: "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
			  "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]),
			  "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]),
			  "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
			: "r"(taddr)
			: "memory");
}
__device__ __forceinline__ void g_tmem_wait_ld(float (&a)[32], float (&b)[32]) {
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
	for (int i = 0; i < 32; ++i) asm volatile("" : "+f"(a[i]), "+f"(b[i]));
}
__device__ __forceinline__ void g_tmem_ld32(uint32_t taddr, float (&v)[32]) {
	uint32_t r[32];
	asm volatile(
			"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
			"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
			"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
			: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
			  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
			  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
			  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
			: "r"(taddr)
			: "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
	for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------------------
// consumer: one warp per chain, lane = slot
// ---------------------------------------------------------------------------------------------------------
static __device__ __noinline__ float g_log2density_stream64(const float *th, const float *xrow) {
	float q = 0.0f;
	for (int r = 0; r < GD; ++r) {
		float y = 0.0f;
		for (int c = r; c < GD; ++c) y = fmaf(__ldcg(th + GD + npb_tri_off(GD, r, c)), __ldg(xrow + c) - __ldcg(th + c), y);
		q = fmaf(y, y, q);
	}
	return __ldcg(th + GD + GTRI) - q;
}

// Birth at D = 64: theta' of draw m of step `step`, consistent with the key that won the race (aux_birth_z of
// npb_alg8_tile4.cuh with two coordinates per lane: lane and lane + 32), written to the slot table.
static __device__ __noinline__ void g_birth_theta64(const Philox &ph, const PriorDev &pr, const float *xw, float rn, uint32_t step, uint32_t sweep,
		int m, int lane, float *th) {
	uint32_t base[4], as[4];
	aux_seed(ph, step, sweep, base);
	aux_stream(base, m, as);
	float av = 1.0f, zpar = 0.0f, R2 = 0.0f;
	aux_draw_chi<GD>(as, pr, av, zpar, R2);
	const float a0 = rn > 0.0f ? __ldg(xw + lane) / rn : (lane == 0 ? 1.0f : 0.0f);
	const float a1 = rn > 0.0f ? __ldg(xw + lane + 32) / rn : 0.0f;
	uint32_t w[4];
	ph(step, 2u + (uint32_t)lane, sweep, NPB_RNG_AUX, w);
	float g0, g1;
	npb_normal2(w[0], w[1], g0, g1);
	float dot = g0 * a0 + g1 * a1;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
	g0 -= dot * a0;
	g1 -= dot * a1;
	float nn = g0 * g0 + g1 * g1;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
	const float rs = nn > 0.0f ? rsqrtf(nn) : 0.0f, sr = sqrtf(R2), sc = av * pr.inv_sqrt_kappa;
	const float z0 = (zpar * a0 + sr * g0 * rs) * sc, z1 = (zpar * a1 + sr * g1 * rs) * sc;
	float mu_lo = pr.mu0[lane], mu_hi = pr.mu0[lane + 32];
	for (int c = 0; c < GD; ++c) {
		const float gc = __shfl_sync(0xffffffffu, c < 32 ? z0 : z1, c & 31);
		if (lane <= c) mu_lo = fmaf(__ldg(pr.S + npb_tri_off(GD, lane, c)), gc, mu_lo);
		if (lane + 32 <= c) mu_hi = fmaf(__ldg(pr.S + npb_tri_off(GD, lane + 32, c)), gc, mu_hi);
	}
	th[lane] = mu_lo;
	th[lane + 32] = mu_hi;
	const float inv = 1.0f / av;
	for (int q = lane; q < GTRI; q += 32) th[GD + q] = __ldg(pr.CT2 + q) * inv;
	if (lane == 0) th[GD + GTRI] = pr.c0_2 - (float)GD * log2f(av);
}

#define G_NOISE_CAP 20.0f
#define G_NOISE_FLOOR -6.0f
// race noise of (tile, step j, slot k): a counter hash instead of a per-lane stream, so that the step-parallel pass
// (lane = step) and the sequential pass (lane = slot) of g_consume_chain draw the same number for the same candidate
__device__ __forceinline__ float g_noise(uint32_t T, uint32_t j, uint32_t k) {
	uint32_t x = T + (j * 32u + k) * 0x9E3779B9u;
	x ^= x >> 16;
	x *= 0x85EBCA6Bu;
	x ^= x >> 13;
	x *= 0xC2B2AE35u;
	// capped above: -log2 E > 20 has probability 1.4e-6 per candidate and would shift a pick probability by less than that -- below
	// what FP32 log-densities resolve (1e-6 relative of ~1e2) -- and the cap is what lets the speculative passes exclude slots
	// without drawing their noise.  Bounded below: the 32-bit uniform gives E' <= 32, i.e. noise >= -5, except for the single
	// word that rounds 1 - v to 0 (probability 2^-32 per draw), whose -inf becomes G_NOISE_FLOOR: a slot's key is then never
	// below its noiseless key + G_NOISE_FLOOR, which lets a sole contender be decided without drawing its noise at all.
	return fmaxf(fminf(neg_lg2_exp1_open(x), G_NOISE_CAP), G_NOISE_FLOOR);
}

// D-generic front ends of the rare paths of the race: density of one (item, slot) from the slot table in global memory, and
// theta' of a birth (D = 64: two coordinates per lane; D <= 32: aux_birth_z of npb_alg8_tile4.cuh, one coordinate per lane)
template <int CD>
__device__ __forceinline__ float g_stream_density(const float *th, const float *xrow) {
	if constexpr (CD == 64) return g_log2density_stream64(th, xrow);
	else return log2density_stream<CD>(th, xrow);
}
template <int CD>
__device__ __forceinline__ void g_birth_theta(const Philox &ph, const PriorDev &pr, const float *xw, float rn, uint32_t step, uint32_t sweep,
		int m, int lane, float *th) {
	if constexpr (CD == 64) {
		g_birth_theta64(ph, pr, xw, rn, step, sweep, m, lane, th);
	} else {
		constexpr int TRI = npb_tri(CD);
		float av;
		const float zc = aux_birth_z<CD>(ph, pr, xw, rn, step, sweep, m, lane, av);
		const float g = zc * (av * pr.inv_sqrt_kappa);
		float mu_r = lane < CD ? pr.mu0[lane] : 0.0f;
		for (int c = 0; c < CD; ++c) {
			const float gc = __shfl_sync(0xffffffffu, g, c);
			if (lane <= c && lane < CD) mu_r = fmaf(pr.S[npb_tri_off(CD, lane, c)], gc, mu_r);
		}
		if (lane < CD) th[lane] = mu_r;
		const float inv = 1.0f / av;
		for (int q = lane; q < TRI; q += 32) th[CD + q] = pr.CT2[q] * inv;
		if (lane == 0) th[CD + TRI] = pr.c0_2 - (float)CD * log2f(av);
	}
}

// D = 16 operand-image geometry (k_pre_aimg16 / k_pre_bimg16 write it, k_density_tc16 and k_sweep_tc16 read it)
namespace {
constexpr int HD = 16;
constexpr int HPS = npb_ps(HD);       // 153
constexpr int HTRI = npb_tri(HD);     // 136
constexpr int H_NS = 16;              // slots per unit
constexpr int H_STAGES = 4;
constexpr int H_ASTAGE = 16384;       // 128 rows x 128 bytes
constexpr int H_SLOT_IMG = 2048;      // 16 rows x 128 bytes
constexpr int H_BBYTES = H_NS * H_SLOT_IMG; // 32 KB
constexpr int H_CONST = 20;           // nb[16] (zero if folded into the GEMM), c2, descale, folded flag, descale^2
constexpr int H_NH_MAX = 4;
constexpr int H_SMEM = 1024 + H_NH_MAX * H_BBYTES + H_STAGES * H_ASTAGE + 512 + H_NH_MAX * 16 * 20 * 4 + 256;
}
