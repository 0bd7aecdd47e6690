// npb_alg8_tile.cuh -- Algorithm 8 sweeps for D = 4, 8, 16: warp-specialised CTA, one chain per CTA.
//
// At D = 16 a cluster slot is 153 floats, so a chain's slot table no longer fits the registers of ONE warp (the
// D <= 3 kernel, npb_alg8_kernel.cuh).  Here a CTA of 1 + KMAX/32 warps owns one chain:
//   * each PRODUCER warp keeps 32 slots in registers, one slot per lane (mean, packed triangular precision factor,
//     log2 normaliser: npb_psp(D) registers).  For a tile of 32 consecutive steps it stages the 32 item rows in
//     shared memory and walks them: the item is broadcast (D/4 128-bit shared loads), every lane evaluates its own
//     slot -- D(D+1)/2 + 2D FMAs per (item, slot), the FP32 work of the path -- and writes one entry of the
//     [slot x step] tile.  Parameters never travel: only 4 D bytes of shared-memory traffic per item per warp.
//     (A first version kept the table in shared memory and broadcast theta instead; ncu showed it bound by the
//     shared-memory pipe -- one wavefront per float -- at 24 % FMA utilisation.)
//   * warp 0, the CONSUMER, draws the three auxiliary parameters of every step of the tile (lane = step), then runs
//     the 32 sequential steps from the finished tile (lane = slot): key = density + log2(n - own) - log2 E
//     (exponential race, see npb_alg8_kernel.cuh), warp arg-max, count update.
// Producers and consumer hand the two tile buffers over with named barriers (bar.sync / bar.arrive), so the FMA-bound
// producers and the latency-bound consumer overlap on the same schedulers.  Cluster parameters are frozen between
// births (np_cluster.h:49-51 slices every update away), which is what makes the tile precomputable; a birth
// writes the new theta to the shared-memory master copy and bumps the slot's version: the producer lane reloads its
// registers, and the consumer re-evaluates that one column of tiles computed from an older version.
#pragma once
#include "npb_alg8_kernel.cuh"

#define NPB_TILE 32
#define NPB_BAR_FULL 1   // ids 1,2: producers -> consumer, buffer 0/1
#define NPB_BAR_EMPTY 3  // ids 3,4: consumer -> producers

__host__ __device__ constexpr int npb_psp(int D) { return (npb_ps(D) + 3) & ~3; } // slot stride of the master copy

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// log2 N(x | theta), theta = P[npb_psp(D)]: mu[D], T2 packed upper row-wise, c2
template <int D>
__device__ __forceinline__ float log2density_regs(const float (&P)[npb_psp(D)], const float (&x)[D]) {
	constexpr int TRI = npb_tri(D);
	float d[D];
#pragma unroll
	for (int c = 0; c < D; ++c) d[c] = x[c] - P[c];
	// column-major accumulation: D independent chains, and consecutive FMAs share the operand d[c] (register reuse
	// cache), which keeps three-register FFMAs off the register-bank conflict path
	float y[D];
#pragma unroll
	for (int c = 0; c < D; ++c) {
		y[c] = P[D + npb_tri_off(D, c, c)] * d[c];
#pragma unroll
		for (int r = 0; r < c; ++r) y[r] = fmaf(P[D + npb_tri_off(D, r, c)], d[c], y[r]);
	}
	float q0 = 0.0f, q1 = 0.0f;
#pragma unroll
	for (int r = 0; r < D; r += 2) {
		q0 = fmaf(y[r], y[r], q0);
		q1 = fmaf(y[r + 1], y[r + 1], q1);
	}
	return P[D + TRI] - (q0 + q1);
}

template <int D>
__device__ __forceinline__ void load_theta(const float *th, float (&P)[npb_psp(D)]) {
#pragma unroll
	for (int i = 0; i < npb_psp(D) / 4; ++i) {
		const float4 v = reinterpret_cast<const float4 *>(th)[i];
		P[4 * i + 0] = v.x; P[4 * i + 1] = v.y; P[4 * i + 2] = v.z; P[4 * i + 3] = v.w;
	}
}

template <int D>
__device__ __forceinline__ void load_row(const float *row, float (&x)[D]) {
#pragma unroll
	for (int c = 0; c < D / 4; ++c) {
		const float4 v = reinterpret_cast<const float4 *>(row)[c];
		x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
	}
}

// The M auxiliary draws of a step come from one xoshiro128++ stream seeded by a Philox block of (chain, step, sweep):
// draw m uses (D+2)/2 Box-Muller pairs: normal 0 is v, normals 1..D are z; after the M draws come M words of race noise.
template <int D>
__device__ __forceinline__ void aux_normals(uint32_t (&rs)[4], float (&g)[2 * ((D + 2) / 2)]) {
#pragma unroll
	for (int p = 0; p < (D + 2) / 2; ++p) {
		const uint32_t r0 = xoshiro_next(rs), r1 = xoshiro_next(rs);
		npb_normal2(r0, r1, g[2 * p], g[2 * p + 1]);
	}
}

template <int D, int KMAX>
struct TileSmem {
	static constexpr int PSP = npb_psp(D);
	float theta[KMAX * PSP];            // master copy of the slot table (written by the consumer on a birth)
	float tile[2][KMAX * 33];           // [buffer][slot * 33 + step]: log2-density of the step's item under the slot
	float xs[KMAX / 32][NPB_TILE * D];  // per producer warp: the item rows of the tile it is working on
	int ver_tile[2][KMAX];              // version of slot k the buffer's column was computed from, -1 = not computed
	int ver_cur[KMAX];                  // current slot versions (bumped by a birth)
	unsigned occ[2];                    // occupancy bit mask (slots 0-31, 32-63), maintained by the consumer
	int role_swap;                      // see the role assignment in the kernel
};

template <int D, int KMAX, int M>
__global__ void __launch_bounds__(32 + KMAX) k_alg8_sweep_tile(const SweepArgs a) {
	constexpr int TRI = npb_tri(D), PS = npb_ps(D), PSP = npb_psp(D), NL = KMAX / 32, NTHREADS = 32 + KMAX;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	TileSmem<D, KMAX> &sm = *reinterpret_cast<TileSmem<D, KMAX> *>(smem_raw);
	const int lane = threadIdx.x & 31;
	const int chain = blockIdx.x;
	const int N = a.N, C = a.C;
	const Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
	// Role assignment.  The hardware hands a CTA's warps consecutive warp slots and slot % 4 is the scheduler, so with
	// two-warp CTAs "warp 0 consumes, warp 1 produces" would put every FMA-heavy producer of an SM on schedulers 1 and 3
	// (ncu: issue-active 0.88 max / 0.25 min across schedulers).  Bit 2 of the first warp's slot flips the roles of
	// every other CTA that shares a scheduler pair.
	if (threadIdx.x == 0) {
		unsigned wslot;
		asm volatile("mov.u32 %0, %%warpid;" : "=r"(wslot));
		sm.role_swap = (KMAX == 32) ? (int)((wslot >> 2) & 1u) : 0;
	}
	__syncthreads();
	const int warp = (int)(threadIdx.x >> 5) ^ sm.role_swap; // virtual warp id: 0 = consumer, 1.. = producers

	// ---- master copy of the slot table into shared memory (all warps), occupancy from the counts ----
	{
		const float *th = a.theta + (size_t)chain * KMAX * PS;
		for (int i = threadIdx.x; i < KMAX * PSP; i += NTHREADS) {
			const int k = i / PSP, e = i - k * PSP;
			sm.theta[i] = e < PS ? th[(size_t)k * PS + e] : 0.0f;
		}
		for (int k = threadIdx.x; k < KMAX; k += NTHREADS) sm.ver_cur[k] = 0;
		if (warp == 0) {
#pragma unroll
			for (int s = 0; s < NL; ++s) {
				const unsigned b = __ballot_sync(0xffffffffu, a.counts[(size_t)chain * KMAX + s * 32 + lane] > 0);
				if (lane == 0) sm.occ[s] = b;
			}
			if (NL == 1 && lane == 0) sm.occ[1] = 0u;
		}
	}
	__syncthreads();
	const int tiles_per_sweep = (N + NPB_TILE - 1) / NPB_TILE;

	if (warp > 0) {
		// =========================== PRODUCER: lane = slot ===========================
		const int p = warp - 1;
		const int k = p * 32 + lane;
		float P[PSP];
		load_theta<D>(sm.theta + k * PSP, P);
		int myver = 0;
		float *xs = sm.xs[p];
		int t = 0;
		for (int sw = 0; sw < a.n_sweeps; ++sw) {
			const int32_t *order = a.scan_order + (size_t)sw * N;
			for (int ti = 0; ti < tiles_per_sweep; ++ti, ++t) {
				const int b = t & 1;
				const int s0 = ti * NPB_TILE;
				const int cnt = min(NPB_TILE, N - s0);
				// stage the tile's item rows (lane = step)
				{
					const int item = (lane < cnt) ? order[s0 + lane] : 0;
					const float4 *src = reinterpret_cast<const float4 *>(a.X + (size_t)item * D);
					float4 *dst = reinterpret_cast<float4 *>(xs + lane * D);
#pragma unroll
					for (int c = 0; c < D / 4; ++c) dst[c] = src[c];
				}
				__syncwarp();
				named_bar_sync(NPB_BAR_EMPTY + b, NTHREADS); // the consumer has released buffer b
				const bool occupied = (((volatile unsigned *)sm.occ)[p] >> lane) & 1u;
				const int ver = ((volatile int *)sm.ver_cur)[k];
				__threadfence_block(); // the version is read before the parameters
				if (ver != myver) { // a birth re-used this slot: fetch the new parameters from the master copy
					load_theta<D>(sm.theta + k * PSP, P);
					myver = ver;
				}
				if (occupied) {
					// two items per iteration: twice the independent FMA chains for the same parameter registers
					int j = 0;
					for (; j + 1 < cnt; j += 2) {
						float x0[D], x1[D];
						load_row<D>(xs + j * D, x0);
						load_row<D>(xs + (j + 1) * D, x1);
						const float l0 = log2density_regs<D>(P, x0);
						const float l1 = log2density_regs<D>(P, x1);
						sm.tile[b][k * 33 + j] = l0;
						sm.tile[b][k * 33 + j + 1] = l1;
					}
					if (j < cnt) {
						float x0[D];
						load_row<D>(xs + j * D, x0);
						sm.tile[b][k * 33 + j] = log2density_regs<D>(P, x0);
					}
				}
				sm.ver_tile[b][k] = occupied ? myver : -1;
				__syncwarp(); // lanes of empty slots skipped the walk: reconverge before the (aligned) barrier
				__threadfence_block();
				named_bar_arrive(NPB_BAR_FULL + b, NTHREADS);
			}
		}
		return;
	}

	// =========================== CONSUMER: lane = step (prologue) / lane = slot (steps) ===========================
	float n[NL];
	int myslot[NL];
#pragma unroll
	for (int s = 0; s < NL; ++s) {
		myslot[s] = s * 32 + lane;
		n[s] = (float)a.counts[(size_t)chain * KMAX + myslot[s]];
	}
	int kocc = 0;
#pragma unroll
	for (int s = 0; s < NL; ++s) kocc += __popc(__ballot_sync(0xffffffffu, n[s] > 0.0f));
	unsigned long long st_cand = 0ull, st_moved = 0ull, st_births = 0ull;
	int overflow = 0;
	const float ik2 = a.prior.inv_sqrt_kappa * (float)NPB_HALF_LOG2E_SQRT;
	named_bar_arrive(NPB_BAR_EMPTY + 0, NTHREADS);
	named_bar_arrive(NPB_BAR_EMPTY + 1, NTHREADS);

	// re-evaluates the tile column of `slot` for steps >= j_from of buffer b (lane = step) from the master copy
	auto fix_column = [&](int slot, int b, int j_from, int item, bool valid) {
		float x[D], P[PSP];
		load_row<D>(a.X + (size_t)item * D, x);
		load_theta<D>(sm.theta + slot * PSP, P);
		const float l = log2density_regs<D>(P, x);
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
			// ---- auxiliary draws of step sj (np_neal_algorithm8.cpp:79-84,119-126) and their race keys ----
			float auxkey_j = -INFINITY;
			int auxm = 0;
			{
				float xw[D];
				load_row<D>(a.Xw + (size_t)item * D, xw);
				uint32_t as[4];
				ph((uint32_t)sj, 1u, sweep, NPB_RNG_AUX, as);
				float lkey[M];
#pragma unroll
				for (int m = 0; m < M; ++m) {
					float g[2 * ((D + 2) / 2)];
					aux_normals<D>(as, g);
					const float v = a.prior.v_mean + a.prior.nu * g[0];
					const float av = fmaxf(fabsf(v), 1e-20f);
					const float inv = __frcp_rn(av);
					float q = 0.0f;
#pragma unroll
					for (int c = 0; c < D; ++c) {
						const float y = xw[c] * inv - g[1 + c] * ik2;
						q = fmaf(y, y, q);
					}
					lkey[m] = a.prior.c0_2 - (float)D * fast_lg2(av) - q + a.prior.log2_alpha_m;
				}
#pragma unroll
				for (int m = 0; m < M; ++m) {
					const float key = lkey[m] + neg_lg2_exp1(xoshiro_next(as));
					if (key > auxkey_j) { auxkey_j = key; auxm = m; }
				}
			}
			const int zold_aux_j = zold | (auxm << 16);
			uint32_t rs[4];
			ph((uint32_t)sj, 0u, sweep, NPB_RNG_PICK, rs);
			const int cnt = min(NPB_TILE, N - s0);

			named_bar_sync(NPB_BAR_FULL + b, NTHREADS); // the producers have filled buffer b
			// columns computed from an older version of the slot (born or re-born since), or not computed at all
#pragma unroll
			for (int s = 0; s < NL; ++s) {
				unsigned stale = __ballot_sync(0xffffffffu, n[s] > 0.0f && sm.ver_tile[b][myslot[s]] != sm.ver_cur[myslot[s]]);
				while (stale) {
					const int k = s * 32 + __ffs(stale) - 1;
					stale &= stale - 1;
					fix_column(k, b, 0, item, valid);
				}
			}
			unsigned cand_tile = 0u;

			for (int j = 0; j < cnt; ++j) {
				const int zo_aux = __shfl_sync(0xffffffffu, zold_aux_j, j);
				const float ak = __shfl_sync(0xffffffffu, auxkey_j, j);
				const int zo = zo_aux & 0xffff;
				float key[NL];
				float kmax = -INFINITY;
#pragma unroll
				for (int s = 0; s < NL; ++s) {
					const float base = sm.tile[b][myslot[s] * 33 + j] + neg_lg2_exp1(xoshiro_next(rs));
					const float ne = (zo == myslot[s]) ? n[s] - 1.0f : n[s];
					key[s] = ne > 0.0f ? base + fast_lg2(ne) : -INFINITY;
					kmax = fmaxf(kmax, key[s]);
				}
				const int my_enc = float_order_key(kmax);
				const int top = max(__reduce_max_sync(0xffffffffu, my_enc), float_order_key(ak));
				const unsigned bal = __ballot_sync(0xffffffffu, my_enc == top && kmax > -INFINITY);
				cand_tile += (unsigned)(kocc + M);
				int new_slot;
				bool born = false;
				if (bal != 0u) {
					int code = myslot[NL - 1];
#pragma unroll
					for (int s = NL - 2; s >= 0; --s)
						if (key[s] == kmax) code = myslot[s];
					new_slot = __shfl_sync(0xffffffffu, code, __ffs(bal) - 1);
				} else {
					born = true;
					new_slot = zo;
				}
				if (born || new_slot != zo) {
					// retract (membertrix.cpp:175-233)
					bool dead = false;
#pragma unroll
					for (int s = 0; s < NL; ++s)
						if (zo == myslot[s]) {
							n[s] -= 1.0f;
							dead = n[s] <= 0.0f;
						}
					const bool died = __any_sync(0xffffffffu, dead);
					if (died) {
						kocc--;
						cand_tile--;
						if (lane == 0) sm.occ[zo >> 5] &= ~(1u << (zo & 31));
					}
					if (born) {
						// np_neal_algorithm8.cpp:136-145: the lowest free slot takes theta' of the winning auxiliary draw
						int fs = -1;
#pragma unroll
						for (int s = 0; s < NL; ++s) {
							const unsigned fb = __ballot_sync(0xffffffffu, n[s] <= 0.0f);
							if (fs < 0 && fb) fs = s * 32 + __ffs(fb) - 1;
						}
						if (fs < 0) {
							overflow = 1; // no room: the item goes back where it was
							if (died) { kocc++; if (lane == 0) sm.occ[zo >> 5] |= 1u << (zo & 31); }
						} else {
							new_slot = fs;
							const int m = (zo_aux >> 16) & 0xff;
							const uint32_t step = (uint32_t)(s0 + j);
							// re-derive theta' of draw m of this step from the step's stream: lane d takes normal 1+d
							uint32_t as[4];
							ph(step, 1u, sweep, NPB_RNG_AUX, as);
							float gg[2 * ((D + 2) / 2)];
							for (int mm = 0; mm <= m; ++mm) aux_normals<D>(as, gg);
							const float v = a.prior.v_mean + a.prior.nu * gg[0];
							const float av = fmaxf(fabsf(v), 1e-20f);
							float g = 0.0f;
#pragma unroll
							for (int c = 0; c < D; ++c)
								if (lane == c) g = gg[1 + c] * (av * a.prior.inv_sqrt_kappa);
							float mu_r = lane < D ? a.prior.mu0[lane] : 0.0f;
							for (int c = 0; c < D; ++c) {
								const float gc = __shfl_sync(0xffffffffu, g, c);
								if (lane <= c && lane < D) mu_r = fmaf(a.prior.S[npb_tri_off(D, lane, c)], gc, mu_r);
							}
							float *th = sm.theta + fs * PSP;
							if (lane < D) th[lane] = mu_r;
							const float inv = 1.0f / av;
							for (int q = lane; q < TRI; q += 32) th[D + q] = a.prior.CT2[q] * inv;
							if (lane == 0) th[D + TRI] = a.prior.c0_2 - (float)D * log2f(av);
							__threadfence_block();
							__syncwarp();
							if (lane == 0) { // publish only once theta is complete
								sm.ver_cur[fs] += 1;
								__threadfence_block();
								sm.occ[fs >> 5] |= 1u << (fs & 31);
							}
							__syncwarp();
							kocc++;
							st_births++;
							fix_column(fs, b, j + 1, item, valid);
						}
					}
#pragma unroll
					for (int s = 0; s < NL; ++s)
						if (new_slot == myslot[s]) n[s] += 1.0f;
					st_moved++;
					if (lane == j) znew = new_slot;
				}
			}
			if (valid && znew != zold) a.z[(size_t)item * C + chain] = (npb_z_t)znew;
			__syncwarp();
			st_cand += cand_tile;
			__threadfence_block();
			named_bar_arrive(NPB_BAR_EMPTY + b, NTHREADS);
		}
	}

	// ---- chain state back to memory ----
	{
		float *th = a.theta + (size_t)chain * KMAX * PS;
		int *cn = a.counts + (size_t)chain * KMAX;
#pragma unroll
		for (int s = 0; s < NL; ++s) cn[myslot[s]] = (int)n[s];
		for (int i = lane; i < KMAX * PS; i += 32) {
			const int k = i / PS, e = i - k * PS;
			th[i] = sm.theta[k * PSP + e];
		}
		if (lane == 0) {
			a.kocc[chain] = kocc;
			if (overflow) a.overflow[chain] = 1;
			a.st[(size_t)chain * 4 + 0] += st_cand;
			a.st[(size_t)chain * 4 + 1] += st_moved;
			a.st[(size_t)chain * 4 + 2] += st_births;
		}
	}
}

template <int D, int KMAX>
npb_status npb_launch_alg8_tile(npb_chains *ch, const SweepArgs &a) {
	npb_ctx *ctx = ch->ctx;
	if (ch->m_aux != 3) return npb_fail(ctx, NPB_E_UNSUPPORTED, "m_aux must be 3 for the D >= 4 sweep kernel");
	const size_t shmem = sizeof(TileSmem<D, KMAX>);
	NPB_CUDA_OK(cudaFuncSetAttribute(k_alg8_sweep_tile<D, KMAX, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
	// all of the unified L1/shared array as shared memory: the CTA count per SM is what hides the consumer's latency
	NPB_CUDA_OK(cudaFuncSetAttribute(k_alg8_sweep_tile<D, KMAX, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	k_alg8_sweep_tile<D, KMAX, 3><<<(unsigned)ch->C, 32 + KMAX, shmem, ctx->stream>>>(a);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}
