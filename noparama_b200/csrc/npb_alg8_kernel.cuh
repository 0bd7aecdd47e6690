// npb_alg8_kernel.cuh -- templates of the register-resident Algorithm 8 sweep kernel (included by npb_alg8.cu for
// the shared helpers and by npb_alg8_inst.cu, which is compiled once per (D, SPL) pair so the build parallelises).
#pragma once
#include "npb_internal.h"

#define NPB_SWEEP_WARPS 2

__device__ __forceinline__ float fast_lg2(float x) {
	float y;
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}

// normal number `f` (flat index) of the Philox stream (c0, *, c2, c3): call f>>2, Box-Muller pair (f>>1)&1
__device__ inline float npb_normal_at(const Philox &ph, uint32_t c0, uint32_t c2, uint32_t c3, int f) {
	uint32_t w[4];
	ph(c0, (uint32_t)(f >> 2), c2, c3, w);
	float a, b;
	if (f & 2) npb_normal2(w[2], w[3], a, b); else npb_normal2(w[0], w[1], a, b);
	return (f & 1) ? b : a;
}

// One draw from the base measure in device form (generic D, used by init and by the generic-D kernels):
// v = D + nu*g0 ; mu = mu0 + (|v|/sqrt(kappa)) S g[1..D] ; T2 = CT2/|v| ; c2 = c0_2 - D log2|v|
// normals are numbers f0 .. f0+D of the stream (c0,*,c2,c3).
__device__ inline void npb_draw_theta(const PriorDev &pr, const Philox &ph, uint32_t c0, uint32_t c2c, uint32_t c3, int f0,
		float *out /* [PS]: mu, T2, c2 */) {
	const int D = pr.D, TRI = npb_tri(D);
	float v = pr.v_mean + pr.nu * npb_normal_at(ph, c0, c2c, c3, f0);
	float av = fmaxf(fabsf(v), 1e-20f);
	float sc = av * pr.inv_sqrt_kappa;
	for (int i = 0; i < D; ++i) out[i] = pr.mu0[i];
	for (int j = 0; j < D; ++j) {
		float g = npb_normal_at(ph, c0, c2c, c3, f0 + 1 + j) * sc;
		for (int i = 0; i <= j; ++i) out[i] += pr.S[npb_tri_off(D, i, j)] * g;
	}
	float inv = 1.0f / av;
	for (int t = 0; t < TRI; ++t) out[D + t] = pr.CT2[t] * inv;
	out[D + TRI] = pr.c0_2 - (float)D * log2f(av);
}

// ---------------------------------------------------------------------------------------------------------
// The register-resident sweep kernel: D <= 3, Kmax = 32*SPL slots, M auxiliary draws.
//
// The categorical draw of a step (np_neal_algorithm8.cpp:130, weights p(x|theta_k) n_k and p(x|theta'_m) alpha/M)
// is taken as an exponential race -- candidate k wins iff  log2 w_k - log2 E_k  is the largest, E_k ~ Exp(1)
// i.i.d. -- which has exactly the probabilities w_k / sum(w) of the reference's cumulative-sum rule but needs one
// warp max-reduction instead of a five-stage prefix scan, no normalisation and no underflow handling (weights
// never leave the log domain).  (The double-precision replay path, npb_replay.cu, keeps the literal
// cumulative-sum rule for bit-exact parity with recorded draws.)
//
// A step is split in two stages so that consecutive steps overlap inside one warp:
//   stage A(j)  -- independent of the chain state step j-1 changes: broadcast item j, log2-densities of every
//                  slot, Exp(1) noise, base_s = l_s - log2 E_s;
//   stage B(j)  -- the sequential part: key_s = base_s + log2(n_s - own)  (-inf for an empty slot), warp arg-max,
//                  count update.
// The step loop issues A(j+1) together with B(j) and only ever writes the counts: a step won by an auxiliary
// draw (a birth, ~3e-5 of the steps) leaves the loop and is finished by generic code at tile level, which then
// re-enters the loop.  The number of register levels in use (NL = ceil(highest occupied slot / 32)) is a
// compile-time parameter of the step loop, so the per-level code is straight-line.
// ---------------------------------------------------------------------------------------------------------
template <int D, int SPL, int M>
struct SweepState {
	static constexpr int TRI = npb_tri(D);
	float n[SPL], mu[SPL][D], T[SPL][TRI], c2[SPL];
	int kocc, nlev, overflow;
	unsigned cand_tile;
	unsigned long long st_cand, st_moved, st_births;
};

template <int D, int M, int NG>
struct TileRegs {          // what lane j holds for step s0 + j of the current tile
	int zold_aux;          // old slot | (winning auxiliary index << 16)
	int znew;
	float x[D];
	float auxkey;          // max_m [ log2 p(x|theta'_m) alpha/M - log2 E_m ]
	float av[M];
	float g[NG];
	uint32_t rs[4];        // rs[0]: LCG state of this lane for the tile's race noise (seeded from a Philox block per tile)
};

template <int NL>
struct StageA {
	float base[NL];
	float auxkey;
	int zo_aux;
};

__device__ __forceinline__ int float_order_key(float v) {
	const int i = __float_as_int(v);
	return i ^ ((i >> 31) & 0x7fffffff); // order-preserving map float -> signed int
}

// xoshiro128++ (Blackman & Vigna): the per-lane generator of the race noise.  Its 128-bit state is one Philox4x32-10
// block drawn per (chain, lane, tile), so a stream is only a few hundred draws long and restartable per tile.
__device__ __forceinline__ uint32_t xoshiro_next(uint32_t (&s)[4]) {
	const uint32_t r = __funnelshift_l(s[0] + s[3], s[0] + s[3], 7) + s[0];
	const uint32_t t = s[1] << 9;
	s[2] ^= s[0];
	s[3] ^= s[1];
	s[1] ^= s[2];
	s[0] ^= s[3];
	s[2] ^= t;
	s[3] = __funnelshift_l(s[3], s[3], 11);
	return r;
}

// -log2 E' for E' = -log2(1 - v) = E / ln 2, E ~ Exp(1), v = (r + 1) 2^-32: the common factor ln 2 shifts every
// key alike and drops out of the arg-max.  E' is clamped at 2^-23 (an event of probability 1e-7) so that the key
// stays finite where 1 - v rounds to 1.
__device__ __forceinline__ float neg_lg2_exp1(uint32_t r) {
	const float omv = fmaf(__uint2float_rn(r), -2.3283064365386963e-10f, 1.0f - 2.3283064365386963e-10f);
	const float e = fmaxf(-fast_lg2(omv), 1.1920929e-7f);
	return -fast_lg2(e);
}

template <int D, int SPL, int M, int NG, int NL>
__device__ __forceinline__ void stage_a(const SweepState<D, SPL, M> &S, TileRegs<D, M, NG> &t, int j, StageA<NL> &o) {
	float xs[D];
#pragma unroll
	for (int d = 0; d < D; ++d) xs[d] = __shfl_sync(0xffffffffu, t.x[d], j);
	o.auxkey = __shfl_sync(0xffffffffu, t.auxkey, j);
	o.zo_aux = __shfl_sync(0xffffffffu, t.zold_aux, j);
#pragma unroll
	for (int s = 0; s < NL; ++s) {
		float dd[D];
#pragma unroll
		for (int d = 0; d < D; ++d) dd[d] = xs[d] - S.mu[s][d];
		float q = 0.0f;
#pragma unroll
		for (int r = 0; r < D; ++r) {
			float y = S.T[s][npb_tri_off(D, r, r)] * dd[r];
#pragma unroll
			for (int c = r + 1; c < D; ++c) y = fmaf(S.T[s][npb_tri_off(D, r, c)], dd[c], y);
			q = fmaf(y, y, q);
		}
		// race noise: one 32-bit LCG per lane (Numerical Recipes multiplier; the float conversion keeps the top 24 bits of
		// the state), re-seeded from a Philox block every tile -- a stream is at most 32 * levels draws long.  xoshiro128++
		// was 9 of the ~29 instructions a candidate level costs at D = 2; the clamp of neg_lg2_exp1 is dropped too
		// (E' = 0, probability 2^-25, is a certain win, which is what an exponential that small means).
		t.rs[0] = t.rs[0] * 1664525u + 1013904223u;
		{
			const float omv = fmaf(__uint2float_rn(t.rs[0]), -2.3283064365386963e-10f, 1.0f - 2.3283064365386963e-10f);
			o.base[s] = (S.c2[s] - q) - fast_lg2(-fast_lg2(omv));
		}
	}
}

#define NPB_STEP_BIRTH 0x100

// Runs steps j0 .. cnt-1 of the current tile with NL register levels.  Returns cnt when the tile is done, or
// (j | NPB_STEP_BIRTH) when step j was won by an auxiliary draw: nothing of step j has been applied yet and the caller
// finishes it.  Only S.n and the counters are written here.
template <int D, int SPL, int M, int NG, int NL>
__device__ __forceinline__ int run_steps(SweepState<D, SPL, M> &S, TileRegs<D, M, NG> &t, int lane, int j0, int cnt) {
	StageA<NL> bufA, bufB;
	int myslot[NL];
#pragma unroll
	for (int s = 0; s < NL; ++s) myslot[s] = s * 32 + lane;
	int j = j0;
	int ret = cnt;
	stage_a<D, SPL, M, NG, NL>(S, t, j0, bufA);

	// one step: stage A of step j+1 into `nxt`, stage B of step j from `cur`; returns false to leave the loop
	auto step = [&](const StageA<NL> &cur, StageA<NL> &nxt) -> bool {
		stage_a<D, SPL, M, NG, NL>(S, t, (j + 1) & 31, nxt);
		const int zo = cur.zo_aux & 0xffff;
		float key[NL];
		float kmax = -INFINITY;
#pragma unroll
		for (int s = 0; s < NL; ++s) {
			const float ne = (zo == myslot[s]) ? S.n[s] - 1.0f : S.n[s];
			key[s] = cur.base[s] + fast_lg2(ne);
			kmax = fmaxf(kmax, key[s]);
		}
		const int my_enc = float_order_key(kmax);
		const int top = max(__reduce_max_sync(0xffffffffu, my_enc), float_order_key(cur.auxkey));
		const unsigned b = __ballot_sync(0xffffffffu, my_enc == top && kmax > -INFINITY);
		if (b == 0u) { // an auxiliary draw won (or nothing had weight at all): a cluster is born
			ret = j | NPB_STEP_BIRTH;
			return false;
		}
		int code = myslot[NL - 1];
#pragma unroll
		for (int s = NL - 2; s >= 0; --s)
			if (key[s] == kmax) code = myslot[s];
		const int new_slot = __shfl_sync(0xffffffffu, code, __ffs(b) - 1);
		S.cand_tile += (unsigned)(S.kocc + M);
		if (new_slot != zo) {
			// retract + assign (membertrix.cpp:175-233, 147-164); an emptied cluster disappears (Q6)
			bool dead = false;
#pragma unroll
			for (int s = 0; s < NL; ++s) {
				if (zo == myslot[s]) {
					S.n[s] -= 1.0f;
					dead = S.n[s] <= 0.0f;
				}
				if (new_slot == myslot[s]) S.n[s] += 1.0f;
			}
			if (__any_sync(0xffffffffu, dead)) { S.kocc--; S.cand_tile--; }
			S.st_moved++;
			if (lane == j) t.znew = new_slot;
		}
		++j;
		return j < cnt;
	};
	for (;;) {
		if (!step(bufA, bufB)) break;
		if (!step(bufB, bufA)) break;
	}
	return ret;
}

template <int D, int SPL, int M, int NG, int NL>
struct Dispatch {
	static __device__ __forceinline__ int run(SweepState<D, SPL, M> &S, TileRegs<D, M, NG> &t, int lane, int j, int cnt) {
		if (S.nlev <= NL) return run_steps<D, SPL, M, NG, NL>(S, t, lane, j, cnt);
		return Dispatch<D, SPL, M, NG, (NL < SPL ? NL + 1 : SPL)>::run(S, t, lane, j, cnt);
	}
};
template <int D, int SPL, int M, int NG>
struct Dispatch<D, SPL, M, NG, SPL> {
	static __device__ __forceinline__ int run(SweepState<D, SPL, M> &S, TileRegs<D, M, NG> &t, int lane, int j, int cnt) {
		return run_steps<D, SPL, M, NG, SPL>(S, t, lane, j, cnt);
	}
};

// Finishes step j of the tile, which an auxiliary draw won (np_neal_algorithm8.cpp:136-145): retract the item, give
// the lowest free slot theta' of the winning draw, assign.  Generic over the levels (rare path, kept out of the
// NL-specialised loops).
template <int D, int SPL, int M, int NG>
__device__ __forceinline__ void finish_birth(SweepState<D, SPL, M> &S, TileRegs<D, M, NG> &t, const SweepArgs &a, int lane, int j) {
	constexpr int TRI = npb_tri(D);
	const int zo_aux = __shfl_sync(0xffffffffu, t.zold_aux, j);
	const int zo = zo_aux & 0xffff;
	const int aux_pick = (zo_aux >> 16) & 0xff;
	S.cand_tile += (unsigned)(S.kocc + M);
	bool dead = false;
#pragma unroll
	for (int s = 0; s < SPL; ++s)
		if (zo == s * 32 + lane) {
			S.n[s] -= 1.0f;
			dead = S.n[s] <= 0.0f;
		}
	const bool died = __any_sync(0xffffffffu, dead);
	if (died) { S.kocc--; S.cand_tile--; }
	int fsub = -1, fl = 0;
#pragma unroll
	for (int s = 0; s < SPL; ++s) {
		const unsigned fb = __ballot_sync(0xffffffffu, S.n[s] <= 0.0f);
		if (fsub < 0 && fb) { fsub = s; fl = __ffs(fb) - 1; }
	}
	int new_slot;
	if (fsub < 0) {
		// no room: flag the chain and put the item back where it was
		S.overflow = 1;
		new_slot = zo;
		if (died) S.kocc++;
	} else {
		new_slot = fsub * 32 + fl;
		// the lane that owns this step holds the auxiliary draws: re-derive theta' of the winner
		float avp = 1.0f, gz[D];
#pragma unroll
		for (int m = 0; m < M; ++m)
			if (m == aux_pick) {
				avp = t.av[m];
#pragma unroll
				for (int d = 0; d < D; ++d) gz[d] = t.g[m * (D + 1) + 1 + d];
			}
		avp = __shfl_sync(0xffffffffu, avp, j);
		float munew[D];
#pragma unroll
		for (int d = 0; d < D; ++d) munew[d] = a.prior.mu0[d];
		const float sc = avp * a.prior.inv_sqrt_kappa;
#pragma unroll
		for (int c = 0; c < D; ++c) {
			const float gc = __shfl_sync(0xffffffffu, gz[c], j) * sc;
#pragma unroll
			for (int r = 0; r <= c; ++r) munew[r] = fmaf(a.prior.S[npb_tri_off(D, r, c)], gc, munew[r]);
		}
		const float inv = 1.0f / avp;
		const float c2new = a.prior.c0_2 - (float)D * log2f(avp);
#pragma unroll
		for (int s = 0; s < SPL; ++s)
			if (s == fsub && lane == fl) {
#pragma unroll
				for (int d = 0; d < D; ++d) S.mu[s][d] = munew[d];
#pragma unroll
				for (int q = 0; q < TRI; ++q) S.T[s][q] = a.prior.CT2[q] * inv;
				S.c2[s] = c2new;
			}
		S.kocc++;
		S.st_births++;
		if (fsub + 1 > S.nlev) S.nlev = fsub + 1;
	}
#pragma unroll
	for (int s = 0; s < SPL; ++s)
		if (new_slot == s * 32 + lane) S.n[s] += 1.0f;
	S.st_moved++;
	if (lane == j) t.znew = new_slot;
}

template <int D, int SPL, int M>
__global__ void __launch_bounds__(NPB_SWEEP_WARPS * 32) k_alg8_sweep_reg(const SweepArgs a) {
	constexpr int TRI = npb_tri(D), PS = npb_ps(D);
	constexpr int NPAIR = (M * (D + 1) + 1) / 2;  // Box-Muller pairs for the M (D+1) normals of a step
	constexpr int NC = (2 * NPAIR + M + 3) / 4;   // Philox calls per step: the normals + M words of race noise
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int chain = blockIdx.x * NPB_SWEEP_WARPS + warp;
	if (chain >= a.C) return;
	const int N = a.N, C = a.C;

	// ---- chain state into registers ----
	SweepState<D, SPL, M> S;
	{
		const float *th = a.theta + (size_t)chain * a.Kmax * PS;
		const int *cn = a.counts + (size_t)chain * a.Kmax;
#pragma unroll
		for (int s = 0; s < SPL; ++s) {
			const int slot = s * 32 + lane;
			S.n[s] = (float)cn[slot];
#pragma unroll
			for (int d = 0; d < D; ++d) S.mu[s][d] = th[(size_t)slot * PS + d];
#pragma unroll
			for (int q = 0; q < TRI; ++q) S.T[s][q] = th[(size_t)slot * PS + D + q];
			S.c2[s] = th[(size_t)slot * PS + D + TRI];
		}
	}
	S.kocc = 0;
	S.nlev = 1;
	S.overflow = 0;
	S.cand_tile = 0u;
	S.st_cand = S.st_moved = S.st_births = 0ull;
#pragma unroll
	for (int s = 0; s < SPL; ++s) S.kocc += __popc(__ballot_sync(0xffffffffu, S.n[s] > 0.0f));

	const Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
	const float ik2 = a.prior.inv_sqrt_kappa * (float)NPB_HALF_LOG2E_SQRT;
	const float fD = (float)D;

	for (int sw = 0; sw < a.n_sweeps; ++sw) {
		const uint32_t sweep = a.sweep0 + (uint32_t)sw;
		const int32_t *order = a.scan_order + (size_t)sw * N;
		// levels only grow inside a sweep; shrink them here when the top ones have emptied
		{
			int nl = 1;
#pragma unroll
			for (int s = 0; s < SPL; ++s)
				if (__ballot_sync(0xffffffffu, S.n[s] > 0.0f)) nl = s + 1;
			S.nlev = nl;
		}
		for (int s0 = 0; s0 < N; s0 += 32) {
			// ---------------- tile prologue: lane j owns step s0 + j ----------------
			TileRegs<D, M, 2 * NPAIR> t;
			const int sj = s0 + lane;
			const bool valid = sj < N;
			const int item = valid ? order[sj] : 0;
			const int zold = valid ? (int)a.z[(size_t)item * C + chain] : 0;
			t.znew = zold;
			float xw[D];
#pragma unroll
			for (int d = 0; d < D; ++d) {
				t.x[d] = a.X[(size_t)item * D + d];
				xw[d] = a.Xw[(size_t)item * D + d];
			}
			uint32_t rw[NC * 4];
#pragma unroll
			for (int c = 0; c < NC; ++c) ph((uint32_t)sj, (uint32_t)c, sweep, NPB_RNG_AUX, rw + 4 * c);
			// words 0 .. 2 NPAIR - 1 make the normals (Box-Muller pairs), the next M words the auxiliaries' race noise
#pragma unroll
			for (int p = 0; p < NPAIR; ++p) npb_normal2(rw[2 * p], rw[2 * p + 1], t.g[2 * p], t.g[2 * p + 1]);
			t.auxkey = -INFINITY;
			int auxm = 0;
#pragma unroll
			for (int m = 0; m < M; ++m) {
				const float v = a.prior.v_mean + a.prior.nu * t.g[m * (D + 1)];
				const float av = fmaxf(fabsf(v), 1e-20f);
				const float inv = __frcp_rn(av);
				float q = 0.0f;
#pragma unroll
				for (int d = 0; d < D; ++d) {
					const float y = xw[d] * inv - t.g[m * (D + 1) + 1 + d] * ik2;
					q = fmaf(y, y, q);
				}
				t.av[m] = av;
				const float key = (a.prior.c0_2 - fD * fast_lg2(av) - q + a.prior.log2_alpha_m) + neg_lg2_exp1(rw[2 * NPAIR + m]);
				if (key > t.auxkey) { t.auxkey = key; auxm = m; }
			}
			t.zold_aux = zold | (auxm << 16);
			ph((uint32_t)sj, 0u, sweep, NPB_RNG_PICK, t.rs);
			const int cnt = min(32, N - s0);

			// ---------------- the steps of the tile ----------------
			int j = 0;
			while (j < cnt) {
				const int r = Dispatch<D, SPL, M, 2 * NPAIR, 1>::run(S, t, lane, j, cnt);
				j = r & 0xff;
				if (r & NPB_STEP_BIRTH) {
					finish_birth<D, SPL, M, 2 * NPAIR>(S, t, a, lane, j);
					++j;
				}
			}

			// ---------------- tile epilogue ----------------
			if (valid && t.znew != zold) a.z[(size_t)item * C + chain] = (npb_z_t)t.znew;
			S.st_cand += S.cand_tile;
			S.cand_tile = 0u;
		}
		__syncwarp();
	}

	// ---- chain state back to memory ----
	{
		float *th = a.theta + (size_t)chain * a.Kmax * PS;
		int *cn = a.counts + (size_t)chain * a.Kmax;
#pragma unroll
		for (int s = 0; s < SPL; ++s) {
			const int slot = s * 32 + lane;
			cn[slot] = (int)S.n[s];
			if (S.n[s] > 0.0f) {
#pragma unroll
				for (int d = 0; d < D; ++d) th[(size_t)slot * PS + d] = S.mu[s][d];
#pragma unroll
				for (int q = 0; q < TRI; ++q) th[(size_t)slot * PS + D + q] = S.T[s][q];
				th[(size_t)slot * PS + D + TRI] = S.c2[s];
			}
		}
		if (lane == 0) {
			a.kocc[chain] = S.kocc;
			if (S.overflow) a.overflow[chain] = 1;
			a.st[(size_t)chain * 4 + 0] += S.st_cand;
			a.st[(size_t)chain * 4 + 1] += S.st_moved;
			a.st[(size_t)chain * 4 + 2] += S.st_births;
		}
	}
}

template <int D, int SPL>
npb_status npb_launch_alg8_reg(npb_chains *ch, const SweepArgs &a) {
	npb_ctx *ctx = ch->ctx;
	int64_t blocks = (ch->C + NPB_SWEEP_WARPS - 1) / NPB_SWEEP_WARPS;
	if (ch->m_aux == 3) k_alg8_sweep_reg<D, SPL, 3><<<(unsigned)blocks, NPB_SWEEP_WARPS * 32, 0, ctx->stream>>>(a);
	else if (ch->m_aux == 1) k_alg8_sweep_reg<D, SPL, 1><<<(unsigned)blocks, NPB_SWEEP_WARPS * 32, 0, ctx->stream>>>(a);
	else return npb_fail(ctx, NPB_E_UNSUPPORTED, "m_aux must be 1 or 3 for the register-resident sweep kernel");
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}
