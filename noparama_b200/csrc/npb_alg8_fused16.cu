// npb_alg8_fused16.cu -- D = 16, Kmax = 32: the Algorithm 8 / Algorithm 2 sweep as ONE kernel per block of steps: tcgen05
// quadratic forms, the density epilogue and the race of NealAlgorithm8::update (src/np_neal_algorithm8.cpp:49-167; density:
// src/statistics/multivariatenormal.cpp:106-136; pick: include/helper/dim1algebra.hpp:2078-2104) in the same CTA, so that the
// [step x slot] log2-density table never leaves the SM.  (Round 1 ran k_density_tc16 + k_race of npb_alg8_gemm.cu: the table,
// 128 bytes per reassignment, was written to HBM by one kernel and read back by the other -- 232 GB of DRAM traffic per sweep
// at the headline shape against 5 GB algorithmic, and the race's 30 ms per sweep were a table reader's.)
//
// CTA = one SM, persistent over units of TWO chains (their B operand images, 128 KB, stay in shared memory for the block;
// an A tile of 128 steps streamed in by cp.async.bulk is used for both).  Warp roles:
//   warp 8        issues tcgen05.mma kind::f16, M = 128 steps x N = 256 (16 slots x 16 rows) x K = 64 (FP16x3 split + folded
//                 offsets, see k_pre_aimg16 / k_pre_bimg16), two TMEM accumulators of 256 columns;
//   warp 9        bulk-copy producer (B images of the unit, A tiles through a two-stage ring);
//   warps 0-7     EPILOGUE: tcgen05.ld of an accumulator (thread = step, 8 slots each), c2 - |y|^2 per (step, slot), written
//                 to a [slot][step] tile in SHARED memory (16 KB per (chain, 128 steps), three tiles in rotation);
//   warps 10-17   DECISION, four per chain of the unit, lane = step of a 32-step sub-tile:
//                 (1) speculative pass, all four sub-tiles in parallel: with the member counts as they stand, the noiseless key
//                     d_k + log2 n_k of every slot, the exact key (counter-hash race noise, g_noise) of the best, and of every
//                     slot the capped noise cannot rule out; the winner w, its key and an upper bound r of every other key;
//                 (2) validation, in step order (a token goes round the chain's four warps): steps that stay are final as long
//                     as no earlier step moved; the first step that moves is applied (retract, assign or birth:
//                     membertrix.cpp:147-233, np_neal_algorithm8.cpp:136-157), and the LATER steps of the sub-tile re-evaluate
//                     only the two slots whose counts changed against their (w, key, r) -- the argmax of independent keys can
//                     only change through those two -- falling back to a full evaluation of a step when that does not settle
//                     it (key of the winner dropped to the bound).  A sub-tile whose speculation predates a change of its
//                     chain's counts (another sub-tile moved an item meanwhile) repeats pass (1) when it gets the token.
//                 Every key is evaluated with the same operations in the same order wherever it is evaluated, so the result
//                 is the sequential sampler's, bit for bit, with or without the speculation (p.spec = 0 evaluates every step
//                 in full, in order; tests/test_gpu_fused16.py compares the two and the round-1 kernel pair).
// A birth writes theta' to the slot table, marks the slot's operand image dirty (rebuilt by k_pre_bimg16 before the next
// block) and from then on the decision warps replace that slot's column of every tile of the block by CUDA-core densities.
#include "npb_tc_common.cuh"

namespace {
constexpr int F_EW = 8;                         // epilogue warps
constexpr int F_DW = 8;                         // decision warps, four per chain
constexpr int F_THREADS = (F_EW + 2 + F_DW) * 32;
constexpr int F_ASTAGES = 2;
constexpr int F_DTBUFS = 3;
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
	int n[32];
	unsigned version;      // bumped by every change of the tables above
	unsigned born_mask;    // slots born during this block: their column of every tile is re-evaluated on the CUDA cores
	unsigned born_seq;
	int kocc, overflow;
	int pad;
	unsigned long long st_cand, st_moved, st_births, st_redo;
};
constexpr int F_SMEM = 1024 + F_CHAIN + 2 * sizeof(FChain);
static_assert(F_SMEM <= 232448, "shared memory of k_sweep_tc16");

// barrier slots (8 bytes each) in the misc area
enum { FB_B_FULL = 0, FB_B_EMPTY = 1, FB_A_FULL = 2, FB_A_EMPTY = 4, FB_T_FULL = 6, FB_T_EMPTY = 8, FB_DT_FULL = 10, FB_DT_FREE = 13, FB_TOK = 16 };
}

template <int M, bool PROBE>
__global__ void __launch_bounds__(F_THREADS, 1) k_sweep_tc16(const GemmArgs g, const PreArgs p) {
	constexpr bool race = !PROBE;
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

	if (warp == F_EW + 1 && lane == 0) {
		g_mbar_init(bars + 8 * FB_B_FULL, 1);
		g_mbar_init(bars + 8 * FB_B_EMPTY, 1);
		for (int s = 0; s < F_ASTAGES; ++s) {
			g_mbar_init(bars + 8 * (FB_A_FULL + s), 1);
			g_mbar_init(bars + 8 * (FB_A_EMPTY + s), 1);
		}
		for (int b = 0; b < 2; ++b) {
			g_mbar_init(bars + 8 * (FB_T_FULL + b), 1);
			g_mbar_init(bars + 8 * (FB_T_EMPTY + b), F_EW * 32);
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
		fcs[0].version = 0u;
		fcs[1].version = 0u;
	}
	if (warp == F_EW) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(g_smem_u32(tmem_slot)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	g_tc_fence_before();
	__syncthreads();
	g_tc_fence_after();
	const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

	if (warp == F_EW + 1) {
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
					g_mbar_wait(bars + 8 * (FB_A_EMPTY + s), ph ^ 1u);
					g_mbar_expect_tx(bars + 8 * (FB_A_FULL + s), H_ASTAGE);
					g_bulk_g2s(base + F_A + s * H_ASTAGE, g.Aimg + (size_t)t * H_ASTAGE, H_ASTAGE, bars + 8 * (FB_A_FULL + s));
				}
			}
		}
		__syncwarp();
	} else if (warp == F_EW) {
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
		// ===================== epilogue: thread = step of the tile, 8 slots of the accumulator =====================
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
					float *drow = Dt + dbuf * F_DT_FLOATS + row * F_DTS + eg * 8;
					if (race) g_mbar_wait(bars + 8 * (FB_DT_FREE + dbuf), (dk & 1u) ^ 1u); // its previous tenant has been decided
#pragma unroll 1
					for (int h = 0; h < 2; ++h, ++acc_it) {
						const int hf = cc * 2 + h;
						const bool folded = (folded_mask >> hf) & 1u;
						const uint32_t buf = acc_it & 1u;
						g_mbar_wait(bars + 8 * (FB_T_FULL + buf), (acc_it >> 1) & 1u);
						g_tc_fence_after();
						const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + buf * 256u + eg * 128;
						const float *ecb = econst + (hf * H_NS + eg * 8) * H_CONST;
						const float2 *cdb = cd2 + hf * H_NS + eg * 8;
#pragma unroll
						for (int pp = 0; pp < 4; pp += 2) {
							float v0[32], v1[32];
							g_tmem_ld32_nowait(taddr + pp * 32u, v0);
							g_tmem_ld32_nowait(taddr + (pp + 1) * 32u, v1);
							g_tmem_wait_ld(v0, v1);
							if (pp == 2) { // the accumulator is in registers: hand the buffer back before the arithmetic
								g_tc_fence_before();
								g_mbar_arrive(bars + 8 * (FB_T_EMPTY + buf));
							}
							float out[4];
							if (folded) {
#pragma unroll
								for (int hh = 0; hh < 4; ++hh) {
									const float(&v)[32] = hh < 2 ? v0 : v1;
									const int o = (hh & 1) * 16;
									const float2 cd = cdb[pp * 2 + hh];
									// q0 .. q3 as two packed accumulators (FFMA2): (q0, q1) and (q2, q3), same sums as the scalar form
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
									out[hh] = fmaf(-cd.y, (q0 + q1) + (q2 + q3), cd.x);
								}
							} else {
#pragma unroll
								for (int hh = 0; hh < 4; ++hh) {
									const float(&v)[32] = hh < 2 ? v0 : v1;
									const int o = (hh & 1) * 16;
									const float *ec = ecb + (pp * 2 + hh) * H_CONST;
									const float dsc = ec[HD + 1];
									float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
#pragma unroll
									for (int i = 0; i < 16; i += 4) {
										const float4 nb = *reinterpret_cast<const float4 *>(ec + i);
										const float a0 = fmaf(v[o + i], dsc, nb.x), a1 = fmaf(v[o + i + 1], dsc, nb.y), a2 = fmaf(v[o + i + 2], dsc, nb.z),
												    a3 = fmaf(v[o + i + 3], dsc, nb.w);
										q0 = fmaf(a0, a0, q0); q1 = fmaf(a1, a1, q1); q2 = fmaf(a2, a2, q2); q3 = fmaf(a3, a3, q3);
									}
									out[hh] = ec[HD] - ((q0 + q1) + (q2 + q3));
								}
							}
							*reinterpret_cast<float4 *>(drow + h * H_NS + pp * 2) = make_float4(out[0], out[1], out[2], out[3]);
							if (PROBE) { // parity probe: the table as the decision warps see it
								float *Lrow = g.L + ((size_t)(2 * u + cc) * g.BS + (size_t)t * G_M + row) * 32 + h * H_NS + eg * 8 + pp * 2;
								*reinterpret_cast<float4 *>(Lrow) = make_float4(out[0], out[1], out[2], out[3]);
							}
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
		const int dw = warp - (F_EW + 2), cc = dw >> 2, wq = dw & 3;
		FChain &fcn = fcs[cc];
		volatile FChain &fc = fcs[cc];
		const float *lg_t = fcn.lg, *lg1_t = fcn.lg1; // re-read after every barrier (the accesses sit behind memory clobbers)
		const SweepArgs &a = p.a;
		const int N = a.N;
		const uint32_t sweep = a.sweep0 + (uint32_t)p.sw;
		const int32_t *order = a.scan_order + (size_t)p.sw * N;
		const uint32_t tok_mine = bars + 8 * (FB_TOK + cc * 4 + wq), tok_next = bars + 8 * (FB_TOK + cc * 4 + ((wq + 1) & 3));
		uint32_t kt = 0, ct_base = 0;
		bool hot = false; // my previous sub-tile had to be re-speculated or moved items: the chain is mixing, speculate late
		for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
			const int ncc = min(2, C - 2 * u);
			if (cc < ncc) {
				const int chain = 2 * u + cc;
				const Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
				float *thc = a.theta + (size_t)chain * 32 * HPS;
				const uint32_t *auxc = a.aux_keys + ((size_t)p.sw * C + chain) * N + p.s0;
				// software pipeline of the step's inputs: item two tiles ahead, old assignment and auxiliary key one tile ahead
				auto ld_item = [&](int t) -> int { const int sl = t * G_M + wq * 32 + lane; return sl < p.nsteps ? __ldg(order + p.s0 + sl) : 0; };
				auto ld_z = [&](int t, int it) -> int { const int sl = t * G_M + wq * 32 + lane; return sl < p.nsteps ? (int)__ldcg(a.z + (size_t)it * C + chain) : 0; };
				auto ld_aux = [&](int t) -> uint32_t { const int sl = t * G_M + wq * 32 + lane; return sl < p.nsteps ? __ldg(auxc + sl) : 0xff800000u; };
				int item_n = ld_item(0), item_nn = ld_item(1);
				int zold_n = ld_z(0, item_n);
				uint32_t auxp_n = ld_aux(0);
				for (int t = 0; t < g.ntiles; ++t, ++kt) {
					const uint32_t ct = ct_base + (uint32_t)(t * ncc + cc);
					const uint32_t dbuf = ct % F_DTBUFS, dk = ct / F_DTBUFS;
					float *drow = Dt + dbuf * F_DT_FLOATS + (wq * 32 + lane) * F_DTS; // d_k of my step = drow[k]
					const int sl0 = t * G_M + wq * 32;  // first step of the sub-tile within the block
					const bool valid = sl0 + lane < p.nsteps;
					const int item = item_n, zold = zold_n;
					const uint32_t auxp = auxp_n;
					item_n = item_nn;
					item_nn = ld_item(t + 2);
					zold_n = ld_z(t + 1, item_n);
					auxp_n = ld_aux(t + 1);
					const float ak = __uint_as_float(auxp);
					const uint32_t T = npb_mix32(npb_mix32(ph.k0 ^ ((uint32_t)(p.s0 + sl0) * 0x9E3779B1u)) ^ ph.k1 ^ (sweep * 0x85EBCA77u) ^ 0x5bd1e995u);
					int znew = zold;
					// speculation state of my step
					int w = -1;
					float keyw = -INFINITY, keyr = -INFINITY;
					bool uncertain = false;
					unsigned seq_seen = 0u;
					bool patched = false;

					auto key_exact = [&](int k) -> float {
						const float lgx = (k == zold) ? lg1_t[k] : lg_t[k];
						return lgx > -INFINITY ? (drow[k] + g_noise(T, (uint32_t)lane, (uint32_t)k)) + lgx : -INFINITY;
					};
					// columns of the slots born during this block: the operand images predate them
					auto patch = [&]() {
						unsigned bm = fc.born_mask;
						if (bm == 0u) return;
						const unsigned seq = fc.born_seq;
						if (patched && seq == seq_seen) return;
						while (bm) {
							const int k = __ffs(bm) - 1;
							bm &= bm - 1;
							const float l = g_stream_density<HD>(thc + (size_t)k * HPS, a.X + (size_t)item * HD);
							if (valid) drow[k] = l;
						}
						patched = true;
						seq_seen = seq;
					};
					// pass (1): winner, its key, and a bound of every other key, with the counts as they stand
					auto speculate = [&]() {
						asm volatile("" ::: "memory");
						float nk[32];
						float pm = -INFINITY; // running maximum of the keys with the slot index in the five low mantissa bits
#pragma unroll
						for (int q = 0; q < 8; ++q) {
							const float4 d = *reinterpret_cast<const float4 *>(drow + 4 * q);
							const float4 l = *reinterpret_cast<const float4 *>(lg_t + 4 * q);
							nk[4 * q + 0] = d.x + l.x; nk[4 * q + 1] = d.y + l.y; nk[4 * q + 2] = d.z + l.z; nk[4 * q + 3] = d.w + l.w;
						}
#pragma unroll
						for (int k = 0; k < 32; ++k) pm = fmaxf(pm, __uint_as_float((__float_as_uint(nk[k]) & ~31u) | (uint32_t)k));
						// the pivot: (about) the best noiseless key -- any slot would do, the best one prunes the most
						const int k1 = (int)(__float_as_uint(pm) & 31u);
						float best = key_exact(k1);
						w = best > -INFINITY ? k1 : -1;
						const float thr = best - (G_NOISE_CAP + 1.0f);
						unsigned need = 0u;
						float m2p = -INFINITY;
#pragma unroll
						for (int k = 0; k < 32; ++k) {
							const bool nd = nk[k] >= thr && nk[k] > -INFINITY;
							need |= nd ? (1u << k) : 0u;
							m2p = nd ? m2p : fmaxf(m2p, nk[k]);
						}
						need &= ~(1u << k1);
						keyr = m2p + (G_NOISE_CAP + 1.0f);
						while (need) {
							const int k = __ffs(need) - 1;
							need &= need - 1;
							const float key = key_exact(k);
							if (key > best || (key == best && key > -INFINITY && k < w)) {
								keyr = fmaxf(keyr, best);
								best = key;
								w = k;
							} else {
								keyr = fmaxf(keyr, key);
							}
						}
						if (ak > best) { // slots win ties against the auxiliary draws
							keyr = fmaxf(keyr, best);
							w = 32 + (int)(auxp & 3u);
							keyw = ak;
						} else {
							keyr = fmaxf(keyr, ak);
							keyw = best;
						}
						uncertain = false;
					};

					g_mbar_wait(bars + 8 * (FB_DT_FULL + dbuf), dk & 1u);
					unsigned v_spec = 0xffffffffu;
					if (p.spec && t > 0 && !hot) {
						v_spec = fc.version;
						__threadfence_block();
						patch();
						speculate();
					}
					g_mbar_wait(tok_mine, kt & 1u);
					// ---- the chain's state is mine from here to the hand-over ----
					if (t == 0 && wq == 0) { // a new unit: this chain's counts
						const int n = a.counts[(size_t)chain * 32 + lane];
						fc.n[lane] = n;
						fc.lg[lane] = n > 0 ? fast_lg2((float)n) : -INFINITY;
						fc.lg1[lane] = n > 1 ? fast_lg2((float)(n - 1)) : -INFINITY;
						const int ko = __popc(__ballot_sync(0xffffffffu, n > 0));
						if (lane == 0) {
							fc.kocc = ko;
							fc.overflow = 0;
							fc.born_mask = 0u;
							fc.born_seq = 0u;
							fc.st_cand = fc.st_moved = fc.st_births = fc.st_redo = 0ull;
							fc.version = fc.version + 1u;
						}
					}
					__syncwarp();
					__threadfence_block();
					bool redo = false;
					if (!p.spec) {
						patch();
						uncertain = valid; // every step evaluated in full, in order
						w = zold;
					} else if (fc.version != v_spec) {
						patch();
						speculate();
						redo = v_spec != 0xffffffffu;
					}
					// ---- pass (2): validation in step order ----
					unsigned live = __ballot_sync(0xffffffffu, valid);
					int kocc = fc.kocc;
					unsigned long long cand = 0ull;
					unsigned n_moved = 0u, n_births = 0u;
					while (true) {
						const unsigned pend = __ballot_sync(0xffffffffu, valid && (uncertain || w != zold)) & live;
						const int jn = pend ? __ffs(pend) - 1 : 32;
						{
							const unsigned below = jn >= 32 ? live : (live & ((1u << jn) - 1u));
							cand += (unsigned long long)(__popc(below) * (kocc + M));
							live &= ~below;
						}
						if (!pend) break;
						const int zo = __shfl_sync(0xffffffffu, zold, jn);
						if (__shfl_sync(0xffffffffu, (int)uncertain, jn)) {
							// full evaluation of step jn: lane = slot (the sequential sampler's step)
							const float akj = __shfl_sync(0xffffffffu, ak, jn);
							const uint32_t auxj = __shfl_sync(0xffffffffu, auxp, jn);
							const float lgx = (lane == zo) ? lg1_t[lane] : lg_t[lane];
							const float key = lgx > -INFINITY ? (Dt[dbuf * F_DT_FLOATS + (wq * 32 + jn) * F_DTS + lane] + g_noise(T, (uint32_t)jn, (uint32_t)lane)) + lgx : -INFINITY;
							const float top = fmaxf(redux_max_f32(key), akj);
							const unsigned bal = __ballot_sync(0xffffffffu, key == top && key > -INFINITY);
							const int ws = bal ? __ffs(bal) - 1 : 32 + (int)(auxj & 3u);
							const float rest = redux_max_f32(lane == ws ? -INFINITY : key);
							if (lane == jn) {
								w = ws;
								keyw = top;
								keyr = bal ? fmaxf(rest, akj) : rest;
								uncertain = false;
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
						cand += (unsigned long long)(kocc + M - (died ? 1 : 0));
						if (born) {
							// np_neal_algorithm8.cpp:136-145: the lowest free slot takes theta' of the winning auxiliary draw
							const unsigned fb = __ballot_sync(0xffffffffu, fc.n[lane] - (lane == zo ? 1 : 0) <= 0);
							if (fb) b = __ffs(fb) - 1;
							else { ovf = true; b = zo; } // no room: the item goes back where it was
						}
						n_moved++;
						if (!ovf) {
							if (born) {
								const int bitem = __shfl_sync(0xffffffffu, item, jn);
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
									nb = fc.n[b] + 1;
								}
								fc.n[b] = nb;
								fc.lg[b] = fast_lg2((float)nb);
								fc.lg1[b] = nb > 1 ? fast_lg2((float)(nb - 1)) : -INFINITY;
								if (born) {
									p.dirty[(size_t)chain * 32 + b] = 1;
									fc.born_mask = fc.born_mask | (1u << b);
									fc.born_seq = fc.born_seq + 1u;
								}
								__threadfence_block();
								fc.version = fc.version + 1u;
							}
							kocc += (born ? 1 : 0) - (died ? 1 : 0);
							__syncwarp();
							if (born) { // the newborn slot's column for my step
								const float l = g_stream_density<HD>(thc + (size_t)b * HPS, a.X + (size_t)item * HD);
								if (valid) drow[b] = l;
								seq_seen = fc.born_seq;
							}
							if (lane == jn) znew = b;
							// later steps: only the keys of the two slots whose counts changed can alter their pick
							if (lane > jn && valid && !uncertain) {
#pragma unroll 1
								for (int e = 0; e < 2; ++e) {
									const int k = e == 0 ? zo : b;
									if (e == 1 && b == zo) break;
									const float kn = key_exact(k);
									if (w == k) {
										keyw = kn;
										if (!(kn > keyr)) uncertain = true;
									} else if (kn > keyw || (kn == keyw && kn > -INFINITY && k < w)) {
										keyr = keyw;
										w = k;
										keyw = kn;
									} else {
										keyr = fmaxf(keyr, kn);
									}
								}
							}
						} else if (lane == 0) {
							fc.overflow = 1;
						}
						if (lane == jn) { // final
							w = zold;
							uncertain = false;
						}
						live &= ~(1u << jn);
					}
					if (valid && znew != zold) a.z[(size_t)item * C + chain] = (npb_z_t)znew;
					hot = redo || n_moved != 0u;
					if (lane == 0) {
						fc.kocc = kocc;
						fc.st_cand = fc.st_cand + cand;
						fc.st_moved = fc.st_moved + n_moved;
						fc.st_births = fc.st_births + n_births;
						if (hot) fc.st_redo = fc.st_redo + 1ull;
					}
					__syncwarp();
					if (t == g.ntiles - 1 && wq == 3) { // the unit's last sub-tile: the chain's state back to memory
						a.counts[(size_t)chain * 32 + lane] = fc.n[lane];
						if (lane == 0) {
							a.kocc[chain] = fc.kocc;
							if (fc.overflow) a.overflow[chain] = 1;
							a.st[(size_t)chain * 4 + 0] += fc.st_cand;
							a.st[(size_t)chain * 4 + 1] += fc.st_moved;
							a.st[(size_t)chain * 4 + 2] += fc.st_births;
							a.st[(size_t)chain * 4 + 3] += fc.st_redo;
						}
					}
					__threadfence_block();
					__syncwarp();
					if (lane == 0) {
						g_mbar_arrive(tok_next);
						g_mbar_arrive(bars + 8 * (FB_DT_FREE + dbuf));
					}
				}
			}
			ct_base += (uint32_t)(g.ntiles * ncc);
		}
	}
	g_tc_fence_before();
	__syncthreads();
	if (warp == F_EW) {
		g_tc_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
	}
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
extern template npb_status npb_launch_aux_keys<16>(npb_chains *, const SweepArgs &);
npb_status npb_tc16_ensure(npb_chains *ch, bool need_table);
npb_status npb_tc16_pre_block(npb_chains *ch, const int32_t *d_order, int nsteps, int born_buf);

static npb_status f_launch(npb_chains *ch, const GemmArgs &g, const PreArgs &p, int race) {
	npb_ctx *ctx = ch->ctx;
	static bool attr_set = false;
	if (!attr_set) {
		NPB_CUDA_OK(cudaFuncSetAttribute(k_sweep_tc16<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
		NPB_CUDA_OK(cudaFuncSetAttribute(k_sweep_tc16<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
		NPB_CUDA_OK(cudaFuncSetAttribute(k_sweep_tc16<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
		attr_set = true;
	}
	int n_sm = 0;
	NPB_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device));
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
	s = npb_launch_aux_keys<16>(ch, a);
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
	PreArgs p;
	memset(&p, 0, sizeof(p));
	p.a = a;
	p.BS = BS + 32;
	p.dirty = ch->g_dirty;
	p.spec = [] { const char *e = getenv("NPB_D64_SPEC"); return !(e && e[0] == '0'); }();
	for (int sw = 0; sw < a.n_sweeps; ++sw) {
		for (int s0 = 0; s0 < N; s0 += BS, ++ch->g_k) {
			const int nsteps = N - s0 < BS ? N - s0 : BS;
			s = npb_tc16_pre_block(ch, a.scan_order + (size_t)sw * N + s0, nsteps, 0);
			if (s != NPB_OK) return s;
			g.ntiles = (nsteps + G_M - 1) / G_M;
			p.sw = sw;
			p.s0 = s0;
			p.nsteps = nsteps;
			s = f_launch(ch, g, p, 1);
			if (s != NPB_OK) return s;
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
