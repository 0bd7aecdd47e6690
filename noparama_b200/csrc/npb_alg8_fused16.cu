// npb_alg8_fused16.cu -- D = 16, Kmax = 32: the Algorithm 8 / Algorithm 2 sweep as ONE kernel per block of steps: tcgen05
// quadratic forms, the density epilogue and the race of NealAlgorithm8::update (src/np_neal_algorithm8.cpp:49-167; density:
// src/statistics/multivariatenormal.cpp:106-136; pick: include/helper/dim1algebra.hpp:2078-2104) in the same CTA, so that the
// [step x slot] log2-density table never leaves the SM.  (Round 1 ran k_density_tc16 + k_race of npb_alg8_gemm.cu: the table,
// 128 bytes per reassignment, was written to HBM by one kernel and read back by the other -- 232 GB of DRAM traffic per sweep
// at the headline shape against 5 GB algorithmic, and the race's 30 ms per sweep were a table reader's.)
//
// CTA = one SM (576 threads, 96 registers, 224 KB of shared memory, all 512 TMEM columns), persistent over units of TWO chains
// (their B operand images, 128 KB, stay in shared memory for the block; an A tile of 128 steps streamed in by cp.async.bulk is
// used for both).  Warp roles:
//   warp 16       issues tcgen05.mma kind::f16, M = 128 steps x N = 256 (16 slots x 16 rows) x K = 64 (FP16x3 split + folded
//                 offsets, see k_pre_aimg16 / k_pre_bimg16), two TMEM accumulators of 256 columns;
//   warp 17       bulk-copy producer (B images of the unit, A tiles through a two-stage ring);
//   warps 0-7     EPILOGUE, two groups of four, one per accumulator: tcgen05.ld (thread = step, 16 slots, two 32-column loads per wait), packed
//                 FP32 (FFMA2) sum of the 16 squares, c2 - dsc^2 |y|^2 per (step, slot), one 16-byte store per 4 slots into a
//                 [step][36] row of a density tile in SHARED memory (18 KB per (chain, 128 steps), three tiles in rotation);
//   warps 8-15    DECISION, four per chain of the unit, lane = step of a 32-step sub-tile:
//                 (1) count-independent pass, all four sub-tiles in parallel: the lane reads its row (32 densities, kept in
//                     registers), finds the best density and the CONTENDERS -- slots within F_WINDOW = 32 log2 units of it:
//                     the race noise lies in [-6, 20] (g_noise) and log2 n_k in [0, 17], so no other slot can win whatever the
//                     counts become -- and caches up to three of them as base = d_k + noise_k (the noise of a sole contender is
//                     drawn lazily), plus a bound for the rest; the item's old slot, the largest auxiliary bound of the 32-step
//                     group (k_aux_bound) come in through a three-stage cp.async ring;
//                 (2) decision, in step order (a token goes round the chain's four warps): with the counts as they stand a step
//                     compares base_k + log2 n_k of its cached contenders (own slot first: it wins next to always) against the
//                     bound of everything else and the group's auxiliary bound; only when that does not settle it is the row
//                     evaluated in full (all 32 keys, the exact auxiliary key by f_aux_exact).  A step that stays changes
//                     nothing, so the token is handed on before any bookkeeping; a move is applied (retract, assign or birth:
//                     membertrix.cpp:147-233, np_neal_algorithm8.cpp:136-157), bumps the chain's version and later steps simply
//                     read the new counts -- their caches do not depend on counts.
//                 Every key is evaluated with the same operations in the same order wherever it is evaluated, so the result
//                 is the sequential sampler's, bit for bit, with or without the short cuts (p.spec = 0 evaluates every step
//                 in full, in order; tests/test_gpu_fused16.py compares the two and the round-1 kernel pair).
// A birth writes theta' to the slot table, marks the slot's operand image dirty (rebuilt by k_pre_bimg16 before the next
// block) and from then on the decision warps replace that slot's column of every tile of the block by CUDA-core densities.
#include "npb_tc_common.cuh"

namespace {
constexpr int F_EW = 8;                         // epilogue warps
constexpr int F_DW = 8;                         // decision warps, four per chain
constexpr int F_THREADS = (F_EW + 2 + F_DW) * 32;
constexpr int F_W_MMA = F_EW + F_DW, F_W_PROD = F_EW + F_DW + 1; // warps 0-7 epilogue, 8-15 decision, 16 MMA, 17 copies
constexpr int F_ASTAGES = 2;
constexpr int F_DTBUFS = 3;
constexpr float F_WINDOW = 32.0f;               // log2 units below a step's best density inside which a slot is one of its contenders
constexpr int F_DTS = 36;                       // row stride of a density tile in floats: [step][36], 32 slots + 4 of padding, so that
                                                // a lane's 128-bit accesses to its own row are conflict-free (rows 144 bytes apart)
constexpr int F_DT_FLOATS = G_M * F_DTS;
constexpr uint32_t F_B = 0;
constexpr uint32_t F_A = 4 * H_BBYTES;          // 131072
constexpr uint32_t F_DT = F_A + F_ASTAGES * H_ASTAGE;
constexpr uint32_t F_MISC = F_DT + F_DTBUFS * F_DT_FLOATS * 4;
constexpr uint32_t F_ECONST = F_MISC + 1024;    // [64 slots][H_CONST]
constexpr uint32_t F_CD = F_ECONST + 64 * H_CONST * 4; // [64 slots] (c2, descale^2) of the short epilogue
constexpr uint32_t F_CHAIN = F_CD + 64 * 8;

struct alignas(16) FChain {
	float lg[32], lg1[32]; // log2 n_k, log2 (n_k - 1); -inf without (other) members
	float occf[32];        // 0 with members, -inf without
	int n[32];
	float lgmax;           // upper bound of log2 n_k over the slots
	unsigned version;      // bumped by every change of the counts
	unsigned born_mask;    // slots born during this block: their column of every tile is re-evaluated on the CUDA cores
	unsigned born_seq;
	int kocc, overflow;
	int pad[2];
	unsigned long long pad2[4];
};
constexpr uint32_t F_RING = F_CHAIN + 2 * sizeof(FChain);  // [8 decision warps][3 stages][80 bytes]: 32 old assignments + the group's auxiliary maximum
constexpr int F_RING_STAGE = 80;
constexpr int F_SMEM = 1024 + F_RING + F_DW * 3 * F_RING_STAGE;
static_assert(F_SMEM <= 232448, "shared memory of k_sweep_tc16");

// barrier slots (8 bytes each) in the misc area
enum { FB_B_FULL = 0, FB_B_EMPTY = 1, FB_A_FULL = 2, FB_A_EMPTY = 4, FB_T_FULL = 6, FB_T_EMPTY = 8, FB_DT_FULL = 10, FB_DT_FREE = 13, FB_TOK = 16 };
}


__device__ __forceinline__ void f_tmem_wait_ld1(float (&a)[32]) {
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
	for (int i = 0; i < 32; ++i) asm volatile("" : "+f"(a[i]));
}

// log2-densities of two slots from 32 accumulator columns (16 per slot): c2 - |y|^2, the offsets folded into the GEMM
// (FOLDED) or added here.  q0 .. q3 accumulate columns i = 0, 1, 2, 3 mod 4 -- as two packed accumulators (FFMA2) in the
// folded form -- and are summed (q0 + q1) + (q2 + q3): the same operations in the same order as k_density_tc16.
template <bool FOLDED>
__device__ __forceinline__ void f_density2(const float (&v)[32], const float *ec, const float2 *cd, float &o0, float &o1) {
#pragma unroll
	for (int hh = 0; hh < 2; ++hh) {
		const int o = hh * 16;
		float r;
		if (FOLDED) {
			const float2 c = cd[hh];
			f32x2_t qa = f2_pack(0.0f, 0.0f), qb = qa;
#pragma unroll
			for (int i = 0; i < 16; i += 4) {
				const f32x2_t va = f2_pack(v[o + i], v[o + i + 1]), vb = f2_pack(v[o + i + 2], v[o + i + 3]);
				qa = f2_fma(va, va, qa);
				qb = f2_fma(vb, vb, qb);
			}
			float q0, q1, q2, q3;
			f2_unpack(qa, q0, q1);
			f2_unpack(qb, q2, q3);
			r = fmaf(-c.y, (q0 + q1) + (q2 + q3), c.x);
		} else {
			const float *e = ec + hh * H_CONST;
			const float dsc = e[HD + 1];
			float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
#pragma unroll
			for (int i = 0; i < 16; i += 4) {
				const float4 nb = *reinterpret_cast<const float4 *>(e + i);
				const float a0 = fmaf(v[o + i], dsc, nb.x), a1 = fmaf(v[o + i + 1], dsc, nb.y), a2 = fmaf(v[o + i + 2], dsc, nb.z),
						    a3 = fmaf(v[o + i + 3], dsc, nb.w);
				q0 = fmaf(a0, a0, q0); q1 = fmaf(a1, a1, q1); q2 = fmaf(a2, a2, q2); q3 = fmaf(a3, a3, q3);
			}
			r = e[HD] - ((q0 + q1) + (q2 + q3));
		}
		if (hh == 0) o0 = r; else o1 = r;
	}
}

// One epilogue thread's share of an accumulator: 128 columns = 8 slots of its step, two tcgen05.ld of 32 columns per wait (a
// staggered order -- one load in flight while the previous one is reduced -- measured 9 % slower); the accumulator goes back to
// the MMA warp as soon as the last columns have landed.
template <bool PROBE, bool ARRIVE = true>
__device__ __forceinline__ void f_epilogue_half(uint32_t taddr, const float2 *cd, float *drow, float *Lrow, uint32_t bar_t_empty) {
	float o[8];
#pragma unroll
	for (int pp = 0; pp < 2; ++pp) {
		float v0[32], v1[32];
		g_tmem_ld32_nowait(taddr + pp * 64u, v0);
		g_tmem_ld32_nowait(taddr + pp * 64u + 32u, v1);
		g_tmem_wait_ld(v0, v1);
		if (ARRIVE && pp == 1) { // the accumulator is in registers: hand the buffer back before the arithmetic
			g_tc_fence_before();
			g_mbar_arrive(bar_t_empty);
		}
		f_density2<true>(v0, nullptr, cd + 4 * pp, o[4 * pp + 0], o[4 * pp + 1]);
		f_density2<true>(v1, nullptr, cd + 4 * pp + 2, o[4 * pp + 2], o[4 * pp + 3]);
		*reinterpret_cast<float4 *>(drow + 4 * pp) = make_float4(o[4 * pp], o[4 * pp + 1], o[4 * pp + 2], o[4 * pp + 3]);
	}
	if (PROBE) { // parity probe: the table as the decision warps see it
		*reinterpret_cast<float4 *>(Lrow) = make_float4(o[0], o[1], o[2], o[3]);
		*reinterpret_cast<float4 *>(Lrow + 4) = make_float4(o[4], o[5], o[6], o[7]);
	}
}
// the same for a half-chain with a slot whose offsets stayed out of the GEMM (a mean far outside the data): rare, so a rolled
// loop over pairs of slots -- compact code rather than fast code
template <bool PROBE>
__device__ __noinline__ void f_epilogue_half_unfolded(uint32_t taddr, const float *ec, float *drow, float *Lrow, uint32_t bar_t_empty) {
#pragma unroll 1
	for (int pr = 0; pr < 4; ++pr) {
		float v[32], o0, o1;
		g_tmem_ld32(taddr + 32u * pr, v);
		if (pr == 3 && bar_t_empty) { // (0: the caller's second call hands the accumulator back)
			g_tc_fence_before();
			g_mbar_arrive(bar_t_empty);
		}
		f_density2<false>(v, ec + 2 * pr * H_CONST, nullptr, o0, o1);
		drow[2 * pr] = o0;
		drow[2 * pr + 1] = o1;
		if (PROBE) {
			Lrow[2 * pr] = o0;
			Lrow[2 * pr + 1] = o1;
		}
	}
}

// the packed race key of the best of a step's M auxiliary draws, exactly as k_aux_keys would have written it
template <int M>
__device__ __noinline__ uint32_t f_aux_exact(const Philox &ph, const PriorDev &pr, float rn, uint32_t sj, uint32_t sweep) {
	float ak;
	int am;
	aux_race<HD, M>(ph, pr, rn, sj, sweep, pr.inv_sqrt_kappa * (float)NPB_HALF_LOG2E_SQRT, ak, am);
	return (__float_as_uint(ak) & ~3u) | (uint32_t)am;
}

template <int M, bool PROBE>
__global__ void __launch_bounds__(F_THREADS, 1) k_sweep_tc16(const GemmArgs g, const PreArgs p) {
	const bool race = !PROBE && p.spec != 2; // (p.spec == 2: measurement switch, table pipeline alone, no decisions)
	extern __shared__ uint8_t g_smem_raw[];
	const uint32_t raw = g_smem_u32(g_smem_raw);
	const uint32_t base = (raw + 1023u) & ~1023u;
	uint8_t *gen = g_smem_raw + (base - raw);
	const uint32_t bars = base + F_MISC;
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + F_MISC + 512);
	float *econst = reinterpret_cast<float *>(gen + F_ECONST);
	float2 *cd2 = reinterpret_cast<float2 *>(gen + F_CD);
	float *Dt = reinterpret_cast<float *>(gen + F_DT);
	FChain *fcs = reinterpret_cast<FChain *>(gen + F_CHAIN);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int C = g.C;
	const int n_units = (C + 1) / 2;

	if (warp == F_W_PROD && lane == 0) {
		g_mbar_init(bars + 8 * FB_B_FULL, 1);
		g_mbar_init(bars + 8 * FB_B_EMPTY, 1);
		for (int s = 0; s < F_ASTAGES; ++s) {
			g_mbar_init(bars + 8 * (FB_A_FULL + s), 1);
			g_mbar_init(bars + 8 * (FB_A_EMPTY + s), 1);
		}
		for (int b = 0; b < 2; ++b) {
			g_mbar_init(bars + 8 * (FB_T_FULL + b), 1);
			g_mbar_init(bars + 8 * (FB_T_EMPTY + b), F_EW * 16);
		}
		for (int b = 0; b < F_DTBUFS; ++b) {
			g_mbar_init(bars + 8 * (FB_DT_FULL + b), F_EW);
			g_mbar_init(bars + 8 * (FB_DT_FREE + b), 4);
		}
		for (int i = 0; i < 8; ++i) g_mbar_init(bars + 8 * (FB_TOK + i), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async;" ::: "memory");
		// the token of either chain starts at its first decision warp
		g_mbar_arrive(bars + 8 * (FB_TOK + 0));
		g_mbar_arrive(bars + 8 * (FB_TOK + 4));
	}
	if (warp == F_W_MMA) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(g_smem_u32(tmem_slot)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	g_tc_fence_before();
	__syncthreads();
	g_tc_fence_after();
	const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

	if (warp == F_W_PROD) {
		// ===================== bulk-copy producer =====================
		if (lane == 0) {
			uint32_t a_it = 0, unit_it = 0;
			for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++unit_it) {
				const int ncc = min(2, C - 2 * u);
				g_mbar_wait(bars + 8 * FB_B_EMPTY, (unit_it & 1u) ^ 1u);
				g_mbar_expect_tx(bars + 8 * FB_B_FULL, (uint32_t)ncc * 2u * H_BBYTES);
				g_bulk_g2s(base + F_B, g.Bimg + (size_t)u * 4 * H_BBYTES, (uint32_t)ncc * 2u * H_BBYTES, bars + 8 * FB_B_FULL);
				for (int t = 0; t < g.ntiles; ++t, ++a_it) {
					const uint32_t s = a_it % F_ASTAGES, ph = (a_it / F_ASTAGES) & 1u;
					if (p.flags & 256) g_mbar_wait_sleep(bars + 8 * (FB_A_EMPTY + s), ph ^ 1u, 200);
					else g_mbar_wait(bars + 8 * (FB_A_EMPTY + s), ph ^ 1u);
					g_mbar_expect_tx(bars + 8 * (FB_A_FULL + s), H_ASTAGE);
					g_bulk_g2s(base + F_A + s * H_ASTAGE, g.Aimg + (size_t)t * H_ASTAGE, H_ASTAGE, bars + 8 * (FB_A_FULL + s));
				}
			}
		}
		__syncwarp();
	} else if (warp == F_W_MMA) {
		// ===================== MMA issue (one thread) =====================
		if (lane == 0) {
			constexpr uint32_t ID256 = g_idesc(G_M, 256);
			uint32_t a_it = 0, acc_it = 0, unit_it = 0;
			for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++unit_it) {
				const int nh = 2 * min(2, C - 2 * u);
				g_mbar_wait(bars + 8 * FB_B_FULL, unit_it & 1u);
				g_tc_fence_after();
				for (int t = 0; t < g.ntiles; ++t, ++a_it) {
					const uint32_t s = a_it % F_ASTAGES;
					g_mbar_wait(bars + 8 * (FB_A_FULL + s), (a_it / F_ASTAGES) & 1u);
					g_tc_fence_after();
					const uint32_t A = base + F_A + s * H_ASTAGE;
					for (int hf = 0; hf < nh; ++hf, ++acc_it) {
						const uint32_t buf = acc_it & 1u;
						g_mbar_wait(bars + 8 * (FB_T_EMPTY + buf), ((acc_it >> 1) & 1u) ^ 1u);
						g_tc_fence_after();
						const uint32_t dcol = tmem + buf * 256u, B = base + F_B + hf * H_BBYTES;
#pragma unroll
						for (int k = 0; k < 4; ++k) // hi*hi, hi*lo, lo*hi, the folded offsets: four K-steps of the same rows
							g_mma_f16(dcol, g_desc(A + k * 32), g_desc(B + k * 32), ID256, k != 0);
						g_tc_commit(bars + 8 * (FB_T_FULL + buf));
					}
					g_tc_commit(bars + 8 * (FB_A_EMPTY + s));
				}
				g_tc_commit(bars + 8 * FB_B_EMPTY);
			}
		}
		__syncwarp();
	} else if (warp < F_EW) {
		// ===================== epilogue: thread = step of the tile =====================
		// Two groups of four warps, group eg bound to TMEM accumulator eg (= half-chain eg of every chain tile, 16 slots): the two
		// epilogue warps of a scheduler work on different accumulators, out of phase, instead of waiting for the same loads and
		// then competing for the same issue slots (all eight on one accumulator, 8 slots each: 64.4 against 62.4 ms per sweep)
		const int wq = warp & 3, eg = warp >> 2;
		const int row = wq * 32 + lane;
		uint32_t acc_it = 0, ct = 0;
		for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
			const int ncc = min(2, C - 2 * u);
			asm volatile("bar.sync 1, %0;" ::"n"(F_EW * 32) : "memory");
			{
				const float *src = g.Bconst + (size_t)u * 64 * H_CONST;
				for (int i = threadIdx.x; i < ncc * 32 * H_CONST; i += F_EW * 32) econst[i] = __ldg(src + i);
				for (int i = threadIdx.x; i < ncc * 32; i += F_EW * 32) cd2[i] = make_float2(__ldg(src + i * H_CONST + HD), __ldg(src + i * H_CONST + HD + 3));
			}
			asm volatile("bar.sync 1, %0;" ::"n"(F_EW * 32) : "memory");
			unsigned folded_mask = 0u; // bit hf: every slot of that half-chain has its offsets inside the GEMM (short epilogue)
			for (int hf = 0; hf < 2 * ncc; ++hf) {
				bool f = true;
#pragma unroll
				for (int sl = 0; sl < H_NS; ++sl) f = f && econst[(hf * H_NS + sl) * H_CONST + HD + 2] != 0.0f;
				folded_mask |= f ? (1u << hf) : 0u;
			}
			for (int t = 0; t < g.ntiles; ++t) {
				for (int cc = 0; cc < ncc; ++cc, ++ct) {
					const uint32_t dbuf = ct % F_DTBUFS, dk = ct / F_DTBUFS;
					if (race) { // its previous tenant has been read by the decision warps
						if (p.flags & 512) g_mbar_wait_sleep(bars + 8 * (FB_DT_FREE + dbuf), (dk & 1u) ^ 1u, 100);
						else g_mbar_wait(bars + 8 * (FB_DT_FREE + dbuf), (dk & 1u) ^ 1u);
					}
					{
						// group eg takes accumulator eg of the chain tile: all 16 slots of half-chain eg for this warp's 32 steps
						const int h = eg, hf = cc * 2 + h;
						const uint32_t ai = acc_it + (uint32_t)h, buf = ai & 1u;
						acc_it += 2;
						g_mbar_wait(bars + 8 * (FB_T_FULL + buf), (ai >> 1) & 1u);
						g_tc_fence_after();
						const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + buf * 256u;
						float *dr = Dt + dbuf * F_DT_FLOATS + row * F_DTS + h * H_NS;
						float *Lrow = PROBE ? g.L + ((size_t)(2 * u + cc) * g.BS + (size_t)t * G_M + row) * 32 + h * H_NS : nullptr;
						if ((folded_mask >> hf) & 1u) {
							f_epilogue_half<PROBE, false>(taddr, cd2 + hf * H_NS, dr, Lrow, 0u);
							f_epilogue_half<PROBE, true>(taddr + 128u, cd2 + hf * H_NS + 8, dr + 8, PROBE ? Lrow + 8 : nullptr, bars + 8 * (FB_T_EMPTY + buf));
						} else {
							f_epilogue_half_unfolded<PROBE>(taddr, econst + (hf * H_NS) * H_CONST, dr, Lrow, 0u);
							f_epilogue_half_unfolded<PROBE>(taddr + 128u, econst + (hf * H_NS + 8) * H_CONST, dr + 8, PROBE ? Lrow + 8 : nullptr, bars + 8 * (FB_T_EMPTY + buf));
						}
					}
					if (race) {
						__syncwarp();
						if (lane == 0) g_mbar_arrive(bars + 8 * (FB_DT_FULL + dbuf));
					}
				}
			}
		}
	} else if (race) {
		// ===================== decision: four warps per chain, lane = step of a 32-step sub-tile =====================
		const int dw = warp - F_EW, cc = dw >> 2, wq = dw & 3;
		FChain &fcn = fcs[cc];
		volatile FChain &fc = fcs[cc];
		const float *lg_t = fcn.lg, *lg1_t = fcn.lg1, *occ_t = fcn.occf; // re-read after every barrier (memory clobbers)
		const SweepArgs &a = p.a;
		const int N = a.N;
		const uint32_t sweep = a.sweep0 + (uint32_t)p.sw;
		const int32_t *order = a.scan_order + (size_t)p.sw * N;
		const uint32_t tok_mine = bars + 8 * (FB_TOK + cc * 4 + wq), tok_next = bars + 8 * (FB_TOK + cc * 4 + ((wq + 1) & 3));
		uint32_t kt = 0, ct_base = 0;
		for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
			const int ncc = min(2, C - 2 * u);
			if (cc < ncc) {
				const int chain = 2 * u + cc;
				const Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
				float *thc = a.theta + (size_t)chain * 32 * HPS;
				// The step's inputs, two tiles ahead, by cp.async into a three-stage ring in shared memory (as register prefetches
				// ptxas spilled them straight after the load, i.e. waited for them): the old assignments, gathered per block into
				// step order by k_gather_z so that they stream instead of chasing scan order -> z through DRAM latency, and an
				// upper bound of the auxiliary keys of the sub-tile's 32 steps (k_aux_bound): it settles nearly every step, a
				// step's own key is drawn (f_aux_exact) only when the bound does not.
				const float *auxg = a.aux_max + ((size_t)p.sw * C + chain) * a.aux_groups + (p.s0 >> 5);
				const npb_z_t *zrow = p.zblk + (size_t)chain * p.zstride;
				uint8_t *ring = gen + F_RING + dw * (3 * F_RING_STAGE);
				auto prefetch = [&](int t) {
					const int sl = t * G_M + wq * 32;
					if (t < g.ntiles) {
						const uint32_t dst = g_smem_u32(ring + (t % 3) * F_RING_STAGE);
						if (lane < 16 && sl + 2 * lane < p.nsteps)
							asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * lane), "l"(zrow + sl + 2 * lane) : "memory");
						if (lane == 16 && sl < p.nsteps)
							asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 64u), "l"(auxg + (sl >> 5)) : "memory");
					}
					asm volatile("cp.async.commit_group;" ::: "memory");
				};
				prefetch(0);
				prefetch(1);
				unsigned acc_cand = 0u, acc_moved = 0u, acc_births = 0u, acc_redo = 0u; // per block: at most 65536 steps x 35
				for (int t = 0; t < g.ntiles; ++t, ++kt) {
					const uint32_t ct = ct_base + (uint32_t)(t * ncc + cc);
					const uint32_t dbuf = ct % F_DTBUFS, dk = ct / F_DTBUFS;
					float *drow = Dt + dbuf * F_DT_FLOATS + (wq * 32 + lane) * F_DTS; // d_k of my step = drow[k]
					const int sl0 = t * G_M + wq * 32;  // first step of the sub-tile within the block
					const bool valid = sl0 + lane < p.nsteps;
					prefetch(t + 2);
					asm volatile("cp.async.wait_group 2;" ::: "memory");
					__syncwarp();
					const int zold = valid ? (int)reinterpret_cast<const npb_z_t *>(ring + (t % 3) * F_RING_STAGE)[lane] : 0;
					const float akmax = sl0 < p.nsteps ? *reinterpret_cast<const float *>(ring + (t % 3) * F_RING_STAGE + 64) : -INFINITY;
					auto item_of = [&]() -> int { return valid ? __ldg(order + p.s0 + sl0 + lane) : 0; }; // (rare paths only)
					auto aux_of = [&]() -> uint32_t { // my step's packed auxiliary key, drawn on demand (rare paths only)
						return valid ? f_aux_exact<M>(ph, a.prior, __ldg(a.Xwn + item_of()), (uint32_t)(p.s0 + sl0 + lane), sweep) : 0xff800000u;
					};
					const uint32_t T = npb_mix32(npb_mix32(ph.k0 ^ ((uint32_t)(p.s0 + sl0) * 0x9E3779B1u)) ^ ph.k1 ^ (sweep * 0x85EBCA77u) ^ 0x5bd1e995u);
					int znew = zold;
					// ---- the step's CONTENDERS, independent of the member counts: the occupied slots whose density lies within
					// F_WINDOW of the best, with their race noise already added (base = d + noise); every other slot's key is
					// below rbound + log2 n_max whatever its noise.  At most three are kept; a step with more is evaluated in full.
					unsigned cpack = 0u;  // slots c0 < c1 < c2, five bits each
					int ncache = 4;       // 0..3, or 4 = not representable: full evaluation
					float base0 = -INFINITY, base1 = -INFINITY, base2 = -INFINITY, rbound = -INFINITY;
					bool noise0 = false;  // base0 already carries its slot's noise
					float d[32];          // my step's row of the density tile: the tile buffer goes back to the epilogue at once
					int w = zold;
					bool unsafe = true;
					unsigned v_pick = 0xffffffffu;
					// the pick from the contenders with the member counts as they stand; unsafe: the bound on the other slots does
					// not clear the winner's key (or there were too many contenders): full evaluation at the step's turn
					auto pick = [&]() {
						asm volatile("" ::: "memory");
						v_pick = fc.version;
						__threadfence_block();
						if (ncache == 1 && !noise0) {
							// one contender: it wins whatever its noise if even the lowest noise clears the other slots' bound and the
							// auxiliary keys; only otherwise is the noise drawn (once)
							const int c = (int)(cpack & 31u);
							const float lgx = (c == zold) ? lg1_t[c] : lg_t[c];
							const float lb = (base0 + G_NOISE_FLOOR) + lgx;
							if (lb > rbound + fc.lgmax + 1.0f && lb > akmax + 1.0f) {
								w = c;
								unsafe = false;
								return;
							}
							base0 = base0 + g_noise(T, (uint32_t)lane, (uint32_t)c);
							noise0 = true;
						}
						float best = -INFINITY;
						int ws = -1;
#pragma unroll
						for (int i = 0; i < 3; ++i) {
							const int c = (int)((cpack >> (5 * i)) & 31u);
							const float bs = i == 0 ? base0 : (i == 1 ? base1 : base2);
							const float lgx = (c == zold) ? lg1_t[c] : lg_t[c];
							const float key = (i < ncache && lgx > -INFINITY) ? bs + lgx : -INFINITY;
							if (key > best) { best = key; ws = c; } // slots in ascending order: the lower one keeps a tie
						}
						if (!(best > akmax)) { // slots win ties against the auxiliary draws
							const uint32_t auxp = aux_of();
							const float ak = __uint_as_float(auxp);
							if (ak > best) {
								best = ak;
								ws = 32 + (int)(auxp & 3u);
							}
						}
						w = ws;
						unsafe = ncache > 3 || !(best > rbound + fc.lgmax);
					};

					// (polled with a back-off: on power-capped boards eight spinning warps cost the epilogue clocks -- 96 -> 84 ms per
					// sweep on one box, no difference on another)
					if (p.flags & 64) g_mbar_wait(bars + 8 * (FB_DT_FULL + dbuf), dk & 1u);
					else g_mbar_wait_sleep(bars + 8 * (FB_DT_FULL + dbuf), dk & 1u, 200);
					bool redo = false;
					// two passes of one body: the row and its contenders before the token (after it, with the chain's state
					// just loaded, on the first tile of a unit), the token in between
#pragma unroll 1
					for (int phase = 0; phase < 2; ++phase) {
						if (phase == (t == 0 ? 1 : 0)) {
							asm volatile("" ::: "memory");
#pragma unroll
							for (int q = 0; q < 8; ++q) {
								const float4 x = *reinterpret_cast<const float4 *>(drow + 4 * q);
								d[4 * q + 0] = x.x; d[4 * q + 1] = x.y; d[4 * q + 2] = x.z; d[4 * q + 3] = x.w;
							}
							// (a slot born during this block: its column of the tile predates it; the steps are evaluated in full,
							// where a born slot's density comes from the CUDA cores)
							if (p.spec && fc.born_mask == 0u) {
								unsigned near = 0u;
								float lim = -INFINITY;
								bool scanned = false;
								if (valid) {
									// the usual case first: the item's own slot is the only contender -- one pass with its density as pivot
									const float vown = drow[zold] + occ_t[zold];
									const float plim = vown - F_WINDOW;
									unsigned pn = 0u;
#pragma unroll
									for (int q = 0; q < 8; ++q) {
										const float4 o = *reinterpret_cast<const float4 *>(occ_t + 4 * q); // 0 with members, -inf without
										pn |= (d[4 * q + 0] + o.x >= plim) ? (1u << (4 * q + 0)) : 0u;
										pn |= (d[4 * q + 1] + o.y >= plim) ? (1u << (4 * q + 1)) : 0u;
										pn |= (d[4 * q + 2] + o.z >= plim) ? (1u << (4 * q + 2)) : 0u;
										pn |= (d[4 * q + 3] + o.w >= plim) ? (1u << (4 * q + 3)) : 0u;
									}
									if (vown > -INFINITY && pn == (1u << zold)) { // nothing else within the window of it: the best, alone
										near = pn;
										lim = plim;
										scanned = true;
									}
								}
								if (!scanned) { // two passes over the row in the tile (rolled: code size matters more than this path's speed)
									float m1 = -INFINITY;
#pragma unroll 1
									for (int k = 0; k < 32; ++k) m1 = fmaxf(m1, drow[k] + occ_t[k]);
									lim = m1 - F_WINDOW;
#pragma unroll 1
									for (int k = 0; k < 32; ++k) {
										const float vk = drow[k] + occ_t[k];
										near |= (vk >= lim && vk > -INFINITY) ? (1u << k) : 0u;
									}
								}
								rbound = lim + (G_NOISE_CAP + 1.0f);
								ncache = min(__popc(near), 4);
								if (ncache <= 3) {
									int c0 = 0, c1 = 0, c2 = 0;
									if (near) { c0 = __ffs(near) - 1; near &= near - 1; }
									if (near) { c1 = __ffs(near) - 1; near &= near - 1; }
									if (near) { c2 = __ffs(near) - 1; }
									cpack = (unsigned)c0 | ((unsigned)c1 << 5) | ((unsigned)c2 << 10);
									// the contenders' race noise; a sole contender's is drawn only if its pick ever needs it (pick())
									base0 = drow[c0];
									noise0 = ncache >= 2 || ((p.flags & 8) && ncache == 1);
									if (noise0) base0 += g_noise(T, (uint32_t)lane, (uint32_t)c0);
									if (ncache >= 2) base1 = drow[c1] + g_noise(T, (uint32_t)lane, (uint32_t)c1);
									if (ncache == 3) base2 = drow[c2] + g_noise(T, (uint32_t)lane, (uint32_t)c2);
								}
							}
							__syncwarp(); // the tile buffer back to the epilogue warps
							if (lane == 0) g_mbar_arrive(bars + 8 * (FB_DT_FREE + dbuf));
							if (p.spec && fc.born_mask == 0u) pick();
						}
						if (phase == 0) {
							if (!(p.flags & 16)) g_mbar_wait(tok_mine, kt & 1u); // (flag 16: measurement only, wrong once an item moves)
							// ---- the chain's state is mine from here to the hand-over ----
							if (t == 0 && wq == 0) { // a new unit: this chain's counts
								const int n = a.counts[(size_t)chain * 32 + lane];
								fc.n[lane] = n;
								const float l0 = n > 0 ? fast_lg2((float)n) : -INFINITY;
								fc.lg[lane] = l0;
								fc.lg1[lane] = n > 1 ? fast_lg2((float)(n - 1)) : -INFINITY;
								fc.occf[lane] = n > 0 ? 0.0f : -INFINITY;
								const int ko = __popc(__ballot_sync(0xffffffffu, n > 0));
								const float lm = redux_max_f32(l0);
								if (lane == 0) {
									fc.kocc = ko;
									fc.lgmax = lm;
									fc.overflow = 0;
									fc.born_mask = 0u;
									fc.version = fc.version + 1u;
								}
							}
							__syncwarp();
							__threadfence_block();
						}
					}
					// The hand-over is the chain's serial path, four hops per tile: a sub-tile whose early pick still stands (no slot
					// born, no count changed since) and keeps every item where it is passes the token on before any bookkeeping.
					if (p.spec && t > 0 && !(t == g.ntiles - 1 && wq == 3) && fc.born_mask == 0u && fc.version == v_pick &&
							__ballot_sync(0xffffffffu, valid && (unsafe || w != zold)) == 0u) {
						const int kocc = fc.kocc;
						if (lane == 0) g_mbar_arrive(tok_next);
						acc_cand += (unsigned)(__popc(__ballot_sync(0xffffffffu, valid)) * (kocc + M));
						continue;
					}
					const unsigned born_now = fc.born_mask;
					if (!p.spec || born_now != 0u) {
						unsafe = true; // every step evaluated in full, in order
					} else if (v_pick == 0xffffffffu || fc.version != v_pick) { // counts changed since the early pick: the contenders' keys again
						redo = t > 0;
						pick();
					}
					// ---- validation in step order ----
					unsigned live = __ballot_sync(0xffffffffu, valid);
					int kocc = fc.kocc;
					unsigned cand = 0u, n_moved = 0u, n_births = 0u;
					while (true) {
						const unsigned pend = __ballot_sync(0xffffffffu, valid && (unsafe || w != zold)) & live;
						const int jn = pend ? __ffs(pend) - 1 : 32;
						{
							const unsigned below = jn >= 32 ? live : (live & ((1u << jn) - 1u));
							cand += (unsigned)(__popc(below) * (kocc + M));
							live &= ~below;
						}
						if (!pend) break;
						const int zo = __shfl_sync(0xffffffffu, zold, jn);
						if (__shfl_sync(0xffffffffu, (int)unsafe, jn)) {
							// full evaluation of step jn: lane = slot (the sequential sampler's step)
							const float lgx = (lane == zo) ? lg1_t[lane] : lg_t[lane];
							float dj = 0.0f; // d[lane] of step jn: its row sits in lane jn's registers
#pragma unroll
							for (int kk = 0; kk < 32; ++kk) {
								const float x = __shfl_sync(0xffffffffu, d[kk], jn);
								dj = lane == kk ? x : dj;
							}
							const unsigned bmk = fc.born_mask;
							if (bmk) { // slots born during this block: the tile's column predates them
								const int itj = __shfl_sync(0xffffffffu, item_of(), jn);
								if ((bmk >> lane) & 1u) dj = g_stream_density<HD>(thc + (size_t)lane * HPS, a.X + (size_t)itj * HD);
							}
							const float key = lgx > -INFINITY ? (dj + g_noise(T, (uint32_t)jn, (uint32_t)lane)) + lgx : -INFINITY;
							// the step's own auxiliary key only if the group's bound does not already lose to the best slot (akmax is the
							// same for the 32 steps of the sub-tile, so the branch is uniform); slots win ties
							const float kmax = redux_max_f32(key);
							uint32_t auxj = 0xff800000u;
							if (!(kmax > akmax)) auxj = __shfl_sync(0xffffffffu, aux_of(), jn);
							const float akj = __uint_as_float(auxj);
							const float top = fmaxf(kmax, akj);
							const unsigned bal = __ballot_sync(0xffffffffu, key == top && key > -INFINITY);
							const int ws = bal ? __ffs(bal) - 1 : 32 + (int)(auxj & 3u);
							if (lane == jn) {
								w = ws;
								unsafe = false;
							}
							if (ws == zo) continue; // it stays
						}
						// ---- step jn moves its item: retract, then assign or birth ----
						const int wj = __shfl_sync(0xffffffffu, w, jn);
						const int na = fc.n[zo] - 1;
						const bool died = na <= 0;
						const bool born = wj >= 32;
						int b = wj;
						bool ovf = false;
						cand += (unsigned)(kocc + M - (died ? 1 : 0));
						if (born) {
							// np_neal_algorithm8.cpp:136-145: the lowest free slot takes theta' of the winning auxiliary draw
							const unsigned fb = __ballot_sync(0xffffffffu, fc.n[lane] - (lane == zo ? 1 : 0) <= 0);
							if (fb) b = __ffs(fb) - 1;
							else { ovf = true; b = zo; } // no room: the item goes back where it was
						}
						n_moved++;
						if (!ovf) {
							if (born) {
								const int bitem = __shfl_sync(0xffffffffu, item_of(), jn);
								g_birth_theta<HD>(ph, a.prior, a.Xw + (size_t)bitem * HD, __ldg(a.Xwn + bitem), (uint32_t)(p.s0 + sl0 + jn), sweep, wj - 32, lane,
										thc + (size_t)b * HPS);
								__threadfence();
								__syncwarp();
								n_births++;
							}
							if (lane == 0) {
								int nb;
								if (b == zo) {
									nb = na + 1;
								} else {
									fc.n[zo] = na;
									fc.lg[zo] = na > 0 ? fast_lg2((float)na) : -INFINITY;
									fc.lg1[zo] = na > 1 ? fast_lg2((float)(na - 1)) : -INFINITY;
									if (na <= 0) fc.occf[zo] = -INFINITY;
									nb = fc.n[b] + 1;
								}
								fc.n[b] = nb;
								const float lb = fast_lg2((float)nb);
								fc.lg[b] = lb;
								fc.lg1[b] = nb > 1 ? fast_lg2((float)(nb - 1)) : -INFINITY;
								fc.occf[b] = 0.0f;
								fc.lgmax = fmaxf(fc.lgmax, lb); // never lowered within a unit: an upper bound is all it has to be
								if (born) {
									p.dirty[(size_t)chain * 32 + b] = 1;
									fc.born_mask = fc.born_mask | (1u << b);
								}
								__threadfence_block();
								fc.version = fc.version + 1u;
							}
							kocc += (born ? 1 : 0) - (died ? 1 : 0);
							__syncwarp();
							__threadfence_block();
							if (lane == jn) znew = b;
							if (born) {
								unsafe = true; // the newborn slot may be anybody's contender: the rest of the sub-tile in full
							} else if (lane > jn && valid && !unsafe) {
								pick(); // the counts of two slots changed: the contenders' keys again
							}
						} else if (lane == 0) {
							fc.overflow = 1;
						}
						if (lane == jn) { // final
							w = zold;
							unsafe = false;
						}
						live &= ~(1u << jn);
					}
					if (valid && znew != zold) a.z[(size_t)item_of() * C + chain] = (npb_z_t)znew;
					// statistics: per warp, flushed once per unit
					acc_cand += cand;
					acc_moved += n_moved;
					acc_births += n_births;
					acc_redo += (redo || n_moved) ? 1u : 0u;
					if (n_moved) {
						if (lane == 0) fc.kocc = kocc;
						__syncwarp();
					}
					if (t == g.ntiles - 1 && wq == 3) { // the unit's last sub-tile: the chain's state back to memory
						a.counts[(size_t)chain * 32 + lane] = fc.n[lane];
						if (lane == 0) {
							a.kocc[chain] = fc.kocc;
							if (fc.overflow) a.overflow[chain] = 1;
						}
					}
					__threadfence_block();
					if (lane == 0) g_mbar_arrive(tok_next);
				}
				if (lane == 0) { // (four warps per chain: atomics)
					atomicAdd(a.st + (size_t)chain * 4 + 0, (unsigned long long)acc_cand);
					atomicAdd(a.st + (size_t)chain * 4 + 1, (unsigned long long)acc_moved);
					atomicAdd(a.st + (size_t)chain * 4 + 2, (unsigned long long)acc_births);
					atomicAdd(a.st + (size_t)chain * 4 + 3, (unsigned long long)acc_redo);
				}
			}
			ct_base += (uint32_t)(g.ntiles * ncc);
		}
	}
	g_tc_fence_before();
	__syncthreads();
	if (warp == F_W_MMA) {
		g_tc_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
	}
}

// ---------------------------------------------------------------------------------------------------------
// Assignments of one block of steps, gathered from the item-major z[N][C] into step order, chain-major: zblk[c][s] =
// z[order[s]][c].  A tile of 32 steps x 64 chains goes through shared memory: reads are 128-byte runs of one item's row,
// writes 64-byte runs of one chain's steps.  The decision warps then stream their inputs (coalesced, L2-resident) instead
// of chasing scan order -> z through DRAM latency once per tile.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_z(const int32_t *order, int nsteps, const npb_z_t *z, int C, npb_z_t *zblk, int zstride) {
	__shared__ npb_z_t tile[32][64 + 2];
	const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
	const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6; // 64 x 4
	for (int r = ty; r < 32; r += 4) {
		const int s = s0 + r;
		if (s < nsteps && c0 + tx < C) tile[r][tx] = z[(size_t)order[s] * C + c0 + tx];
	}
	__syncthreads();
	const int sx = threadIdx.x & 31, cy = threadIdx.x >> 5; // 32 x 8
	for (int c = cy; c < 64; c += 8)
		if (s0 + sx < nsteps && c0 + c < C) zblk[(size_t)(c0 + c) * zstride + s0 + sx] = tile[sx][c];
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
npb_status npb_tc16_ensure(npb_chains *ch, bool need_table);
npb_status npb_tc16_pre_block(npb_chains *ch, const int32_t *d_order, int nsteps, int born_buf);

static npb_status f_launch(npb_chains *ch, const GemmArgs &g, const PreArgs &p, int race) {
	npb_ctx *ctx = ch->ctx;
	if (!ctx->fused_attr_set) { // (a function attribute is per device: one process may drive several, noparama_b200 --gpus G)
		NPB_CUDA_OK(cudaFuncSetAttribute(k_sweep_tc16<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
		NPB_CUDA_OK(cudaFuncSetAttribute(k_sweep_tc16<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
		NPB_CUDA_OK(cudaFuncSetAttribute(k_sweep_tc16<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
		ctx->fused_attr_set = true;
	}
	const int n_sm = ctx->n_sm;
	const int n_units = (g.C + 1) / 2;
	const int grid = n_units < n_sm ? n_units : n_sm;
	if (!race) k_sweep_tc16<1, true><<<grid, F_THREADS, F_SMEM, ctx->stream>>>(g, p);
	else if (ch->m_aux == 3) k_sweep_tc16<3, false><<<grid, F_THREADS, F_SMEM, ctx->stream>>>(g, p);
	else k_sweep_tc16<1, false><<<grid, F_THREADS, F_SMEM, ctx->stream>>>(g, p);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

npb_status npb_launch_alg8_fused16(npb_chains *ch, const SweepArgs &a) {
	npb_ctx *ctx = ch->ctx;
	if (ch->m_aux != 3 && ch->m_aux != 1) return npb_fail(ctx, NPB_E_UNSUPPORTED, "m_aux must be 1 or 3 for the D = 16 tensor-core sweep");
	npb_status s = npb_tc16_ensure(ch, false);
	if (s != NPB_OK) return s;
	// group maxima of a bound of the auxiliary keys (exact keys on demand in the kernel).  (Measured and dropped: the same bounds
	// block by block on a second, low-priority stream in 32-register CTAs meant to sit beside the sweep kernel's CTA -- the sweep
	// kernel's launches grew by what the bounds took, 51.4 -> 58.5 ms per sweep, and the step stayed at 65.7 ms.)
	s = npb_launch_aux_bound<16>(ch, a);
	if (s != NPB_OK) return s;
	const size_t C = (size_t)ch->C;
	// parameters may have changed since the last launch (init_from_params, update_params): every slot's image is rebuilt
	NPB_CUDA_OK(cudaMemsetAsync(ch->g_dirty, 1, C * 32, ctx->stream));
	const int BS = ch->g_bs, N = a.N;
	GemmArgs g;
	memset(&g, 0, sizeof(g));
	g.Aimg = ch->g_aimg;
	g.Bimg = ch->g_bimg;
	g.Bconst = ch->g_bconst;
	g.L = nullptr;
	g.C = (int)C;
	g.BS = BS + 32;
	if (!ch->g_zblk) NPB_CUDA_OK(cudaMalloc((void **)&ch->g_zblk, C * (size_t)BS * sizeof(npb_z_t)));
	PreArgs p;
	memset(&p, 0, sizeof(p));
	p.a = a;
	p.BS = BS + 32;
	p.dirty = ch->g_dirty;
	p.zblk = ch->g_zblk;
	p.zstride = BS;
	p.spec = ch->sw.spec;
	p.flags = ch->sw.f16_flags; // A/B measurement switches (results unchanged)
	for (int sw = 0; sw < a.n_sweeps; ++sw) {
		for (int s0 = 0; s0 < N; s0 += BS, ++ch->g_k) {
			const int nsteps = N - s0 < BS ? N - s0 : BS;
			s = npb_tc16_pre_block(ch, a.scan_order + (size_t)sw * N + s0, nsteps, 0);
			if (s != NPB_OK) return s;
			{
				dim3 gg((unsigned)((nsteps + 31) / 32), (unsigned)((C + 63) / 64));
				k_gather_z<<<gg, 256, 0, ctx->stream>>>(a.scan_order + (size_t)sw * N + s0, nsteps, a.z, (int)C, ch->g_zblk, BS);
				NPB_CUDA_OK(cudaGetLastError());
			}
			g.ntiles = (nsteps + G_M - 1) / G_M;
			p.sw = sw;
			p.s0 = s0;
			p.nsteps = nsteps;
			cudaEvent_t e0 = nullptr, e1 = nullptr;
			if (ch->time_kernels) { // (bench.py: the dominant kernel's own duration, per launch, on the launching stream)
				NPB_CUDA_OK(cudaEventCreate(&e0));
				NPB_CUDA_OK(cudaEventCreate(&e1));
				NPB_CUDA_OK(cudaEventRecord(e0, ctx->stream));
			}
			s = f_launch(ch, g, p, 1);
			if (s != NPB_OK) return s;
			if (ch->time_kernels) {
				NPB_CUDA_OK(cudaEventRecord(e1, ctx->stream));
				ch->kt_ev.push_back(e0);
				ch->kt_ev.push_back(e1);
			}
		}
	}
	return NPB_OK;
}

// parity probe: the [32 slots x 32 items] table exactly as the decision warps of k_sweep_tc16 read it
__global__ void k_fused16_probe_out(const float *L, const int *counts, int chain, int BSP, float *out) {
	const int k = threadIdx.x, j = blockIdx.x;
	const bool occupied = counts[(size_t)chain * 32 + k] > 0;
	out[k * 32 + j] = occupied ? L[((size_t)chain * BSP + j) * 32 + k] * NPB_LN2 : NAN;
}

npb_status npb_launch_fused16_probe(npb_chains *ch, int chain, const int32_t *d_items, float *d_out) {
	npb_ctx *ctx = ch->ctx;
	npb_status s = npb_tc16_ensure(ch, false);
	if (s != NPB_OK) return s;
	NPB_CUDA_OK(cudaMemsetAsync(ch->g_dirty, 1, (size_t)ch->C * 32, ctx->stream));
	s = npb_tc16_pre_block(ch, d_items, 32, 0);
	if (s != NPB_OK) return s;
	const int BSP = G_M;
	float *L = nullptr;
	NPB_CUDA_OK(cudaMalloc((void **)&L, (size_t)ch->C * BSP * 32 * sizeof(float)));
	GemmArgs g;
	memset(&g, 0, sizeof(g));
	g.Aimg = ch->g_aimg;
	g.Bimg = ch->g_bimg;
	g.Bconst = ch->g_bconst;
	g.L = L;
	g.C = (int)ch->C;
	g.ntiles = 1;
	g.BS = BSP;
	PreArgs p;
	memset(&p, 0, sizeof(p));
	s = f_launch(ch, g, p, 0);
	if (s == NPB_OK) {
		k_fused16_probe_out<<<32, 32, 0, ctx->stream>>>(L, ch->counts, chain, BSP, d_out);
		cudaError_t e = cudaGetLastError();
		if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
		if (e != cudaSuccess) s = npb_fail_cuda(ctx, e, "k_fused16_probe_out", __FILE__, __LINE__);
	}
	cudaFree(L);
	return s;
}
