// npb_alg8.cu -- Neal Algorithm 8 sweeps for thousands of lockstep chains (sm_100a).
//
// Replaces, per chain and per item, NealAlgorithm8::update (src/np_neal_algorithm8.cpp:49-167):
//   retract (membertrix.cpp:175-233), M draws from the base measure (np_neal_algorithm8.cpp:79-84 ->
//   normalinvwishart.h:44-64), K weighted densities p(x|theta_k) n_k (:93-109) and M densities
//   p(x|theta'_m) alpha/M (:119-126), random_weighted_pick (dim1algebra.hpp:2078-2104), assign or addCluster
//   (:136-157) -- inside the sweep loop of MCMC::run (np_mcmc.cpp:109-163).
//
// Layout: one warp per chain.  Cluster slot `s` of a chain lives on lane (s & 31), register level (s >> 5):
// its member count, mean, triangular precision factor and log-normaliser stay in registers for the whole
// launch (D <= 3) so a step touches no memory but the item's coordinates and its assignment.  The scan order
// of a sweep is a keyed permutation shared by all chains (npb_common.cuh), evaluated 32 steps at a time:
// lane j of the warp prefetches item, old assignment and coordinates of step s0+j and draws that step's
// auxiliary parameters from Philox, so the sequential part of a step is: keys -> warp arg-max -> count update
// (an exponential race, exact in distribution; see npb_alg8_kernel.cuh).  Weights stay in the log2 domain.
#include "npb_alg8_tile4.cuh"
#include <cstdlib>

// ---------------------------------------------------------------------------------------------------------
// init: np_mcmc.cpp:49-91 -- K0 clusters from the base measure (np_init_clusters.cpp:24-41), every item to a
// uniformly chosen cluster, clusters that got nothing dropped (cleanup, membertrix.cpp:343-364).
// One warp per chain.
// ---------------------------------------------------------------------------------------------------------
__global__ void k_chains_init(SweepArgs a, int K0, const float *theta_given /* [K0, PS] or NULL */, uint32_t epoch) {
	extern __shared__ int sh_counts[]; // [warps][Kmax]
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int chain = blockIdx.x * (blockDim.x >> 5) + warp;
	if (chain >= a.C) return;
	const int D = a.prior.D, PS = npb_ps(D);
	int *cnt = sh_counts + warp * a.Kmax;
	for (int k = lane; k < a.Kmax; k += 32) cnt[k] = 0;
	__syncwarp();
	Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
	float *th = a.theta + (size_t)chain * a.Kmax * PS;
	for (int k = lane; k < a.Kmax; k += 32) {
		float *o = th + (size_t)k * PS;
		if (k < K0 && theta_given) for (int t = 0; t < PS; ++t) o[t] = theta_given[(size_t)k * PS + t];
		else if (k < K0) npb_draw_theta(a.prior, ph, (uint32_t)k, epoch, NPB_RNG_INIT_THETA, 0, o);
		else for (int t = 0; t < PS; ++t) o[t] = 0.0f;
	}
	for (int i = lane; i < a.N; i += 32) {
		uint32_t w[4];
		ph((uint32_t)i, 0u, epoch, NPB_RNG_INIT_Z, w);
		int k = (int)__umulhi(w[0], (uint32_t)K0);
		a.z[(size_t)i * a.C + chain] = (npb_z_t)k;
		atomicAdd(&cnt[k], 1);
	}
	__syncwarp();
	int occ = 0;
	for (int k = lane; k < a.Kmax; k += 32) {
		a.counts[(size_t)chain * a.Kmax + k] = cnt[k];
		occ += cnt[k] > 0;
	}
	occ = __reduce_add_sync(0xffffffffu, occ);
	if (lane == 0) {
		a.kocc[chain] = occ;
		a.overflow[chain] = 0;
		a.st[(size_t)chain * 4 + 0] = a.st[(size_t)chain * 4 + 1] = a.st[(size_t)chain * 4 + 2] = a.st[(size_t)chain * 4 + 3] = 0ull;
	}
}

// ---------------------------------------------------------------------------------------------------------
// whitening of the dataset against the prior: Xw = CT2 (x - mu0)  (CT2 upper triangular packed)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_whiten(const float *X, float *Xw, float *Xwn, int64_t N, PriorDev pr) {
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	const int D = pr.D;
	const float *x = X + i * D;
	float *o = Xw + i * D;
	float q = 0.0f;
	for (int r = 0; r < D; ++r) {
		float s = 0.0f;
		for (int c = r; c < D; ++c) s += pr.CT2[npb_tri_off(D, r, c)] * (x[c] - pr.mu0[c]);
		o[r] = s;
		q += s * s;
	}
	Xwn[i] = sqrtf(q);
}
// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
PriorDev npb_prior_dev(const npb_ctx *ctx, int m_aux) {
	const PriorHost &p = ctx->prior;
	PriorDev d;
	memset(&d, 0, sizeof(d));
	d.D = p.D;
	d.m_aux = m_aux;
	d.flags = p.flags;
	d.inv_sqrt_kappa = (float)(1.0 / sqrt(p.kappa));
	d.nu = (float)p.nu;
	d.v_mean = (float)p.D;
	d.c0_2 = (float)(-0.5 * (p.D * log2(2.0 * M_PI) + p.logdetA / log(2.0)));
	d.log2_alpha_m = (float)log2(p.alpha / (double)(m_aux > 0 ? m_aux : 1));
	for (int i = 0; i < p.D; ++i) d.mu0[i] = (float)p.mu0[i];
	d.CT2 = ctx->d_CT2;
	d.S = ctx->d_S;
	return d;
}

static SweepArgs make_args(npb_chains *ch, int n_sweeps) {
	SweepArgs a;
	a.X = ch->ds->X32;
	a.Xw = ch->ds->Xw;
	a.Xwn = ch->ds->Xwn;
	a.z = ch->z;
	a.theta = ch->theta;
	a.counts = ch->counts;
	a.st = ch->st;
	a.kocc = ch->kocc;
	a.overflow = ch->overflow;
	a.N = (int)ch->ds->N;
	a.C = (int)ch->C;
	a.Kmax = ch->Kmax;
	a.sweep0 = ch->sweep;
	a.n_sweeps = n_sweeps;
	a.seed = ch->seed;
	a.scan_order = ch->scan_order;
	a.aux_keys = ch->aux_keys;
	a.aux_max = ch->aux_max;
	a.aux_groups = (int)((ch->ds->N + 31) / 32);
	a.prior = npb_prior_dev(ch->ctx, ch->m_aux);
	return a;
}

npb_status npb_launch_whiten(npb_dataset *ds) {
	npb_ctx *ctx = ds->ctx;
	PriorDev pd = npb_prior_dev(ctx, 1);
	int threads = 256;
	int64_t blocks = (ds->N + threads - 1) / threads;
	k_whiten<<<(unsigned)blocks, threads, 0, ctx->stream>>>(ds->X32, ds->Xw, ds->Xwn, ds->N, pd);
	NPB_CUDA_OK(cudaGetLastError());
	ds->whitened_epoch = ctx->prior_epoch;
	return NPB_OK;
}

npb_status npb_launch_chains_init(npb_chains *ch, int K0, const float *d_theta_given) {
	npb_ctx *ctx = ch->ctx;
	if (ctx->prior.family) {
		if (d_theta_given) return npb_fail(ctx, NPB_E_UNSUPPORTED, "scalar-noise families: chains start from prior draws only");
		return npb_launch_sn_init(ch, K0);
	}
	SweepArgs a = make_args(ch, 0);
	const int warps = 4;
	int64_t blocks = (ch->C + warps - 1) / warps;
	size_t shmem = (size_t)warps * ch->Kmax * sizeof(int);
	k_chains_init<<<(unsigned)blocks, warps * 32, shmem, ctx->stream>>>(a, K0, d_theta_given, d_theta_given ? ++ch->init_epoch : 0u); // distinct initial assignments for repeated initialisations of a handle
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

// explicit instantiations live in npb_alg8_inst.cu (one object per (D, SPL))
#define NPB_DECL(D, SPL) extern template npb_status npb_launch_alg8_reg<D, SPL>(npb_chains *, const SweepArgs &);
NPB_DECL(2, 1) NPB_DECL(2, 2) NPB_DECL(2, 4) NPB_DECL(2, 8) NPB_DECL(2, 16)
NPB_DECL(3, 1) NPB_DECL(3, 2) NPB_DECL(3, 4) NPB_DECL(3, 8)
#undef NPB_DECL
#define NPB_TDECL(D, K) extern template npb_status npb_launch_alg8_tile<D, K>(npb_chains *, const SweepArgs &);
NPB_TDECL(4, 32) NPB_TDECL(4, 64) NPB_TDECL(8, 32) NPB_TDECL(8, 64) NPB_TDECL(16, 32) NPB_TDECL(16, 64)
#undef NPB_TDECL
#define NPB_T4DECL(D) extern template npb_status npb_launch_alg8_tile4<D>(npb_chains *, const SweepArgs &); \
	extern template npb_status npb_launch_aux_keys<D>(npb_chains *, const SweepArgs &); \
	extern template npb_status npb_launch_tile4_probe<D>(npb_chains *, const SweepArgs &, int, const int32_t *, float *);
NPB_T4DECL(4) NPB_T4DECL(8) NPB_T4DECL(16)
#undef NPB_T4DECL

// scan order of sweeps sweep0 .. sweep0+n_sweeps-1 into order[n_sweeps][N] (npb_common.cuh: keyed permutation with
// cycle walking; evaluated once per sweep here instead of per chain inside the sweep kernel)
__global__ void k_scan_order(int32_t *order, int N, uint64_t seed, uint32_t sweep0) {
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= N) return;
	const ScanOrder so = npb_scan_order(seed, sweep0 + blockIdx.y, (uint32_t)N);
	order[(size_t)blockIdx.y * N + s] = (int32_t)npb_scan_item(so, (uint32_t)s);
}

npb_status npb_launch_scan_order(npb_chains *ch, int n_sweeps) {
	npb_ctx *ctx = ch->ctx;
	const int N = (int)ch->ds->N;
	dim3 grid((N + 255) / 256, n_sweeps);
	k_scan_order<<<grid, 256, 0, ctx->stream>>>(ch->scan_order, N, ch->seed, ch->sweep);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

static npb_status launch_chunk(npb_chains *ch, int n_sweeps) {
	npb_ctx *ctx = ch->ctx;
	npb_status so = npb_launch_scan_order(ch, n_sweeps);
	if (so != NPB_OK) return so;
	SweepArgs a = make_args(ch, n_sweeps);
	npb_status s = NPB_E_UNSUPPORTED;
	int key = ch->D * 1000 + ch->Kmax / 32;
	// Kmax = 32, D >= 4: four chains per CTA with setmaxnreg (npb_alg8_tile4.cuh); NPB_TILE_KERNEL=2warp selects the
	// earlier one-chain-per-CTA kernel for A/B measurements
	const bool two_warp = ch->sw.two_warp;
	if (!two_warp && ch->Kmax == 32 && (ch->D == 4 || ch->D == 8 || ch->D == 16)) key = -ch->D;
	if (ch->Kmax == 32 && ch->D == 64) key = -64;
	// D = 16, Kmax = 32: the tensor path of npb_alg8_gemm.cu (2x the FP32-pipe kernel at the headline shape);
	// NPB_D16_PATH=fp32 selects k_alg8_sweep_tile4 (read at every launch: A/B measurements, tests of both)
	// D = 16, Kmax = 32: the tensor paths.  k_sweep_tc16 (npb_alg8_fused16.cu, one fused kernel, the density table never leaves
	// the SM) and round 1's table kernel + race kernel (npb_alg8_gemm.cu) evaluate the same keys and give the same assignments
	// bit for bit (tests/test_gpu_fused16.py), so the choice is a matter of speed only: the fused kernel runs two chains per SM
	// and wins while few items move (92 against 101 ms per sweep at the headline shape); the kernel pair runs 55 chains per SM
	// through its sequential pass and wins in a mixing chain (165 against 370 ms at 15 % moved).  NPB_D16_PATH: unset / auto =
	// by the moved fraction of the handle's last sweep with statistics (unknown: fused); tc = fused; tc2 = pair; fp32 =
	// k_alg8_sweep_tile4.
	if (ch->Kmax == 32 && ch->D == 16) {
		const char *e = ch->opt_d16_path;
		if (e[0] == 'f') key = -16;
		else if (e[0] == 't' && e[1] == 'c' && e[2] == '2') key = -1602;
		else if (e[0] == 't') key = -1600;
		else key = ch->moved_frac_last > 0.01 ? -1602 : -1600;
	}
	switch (key) {
	// pre-pass (state independent, fully parallel): the race key of every (chain, step)'s auxiliary draws; then the sweep
	case -4: s = npb_launch_aux_keys<4>(ch, a); if (s == NPB_OK) s = npb_launch_alg8_tile4<4>(ch, a); break;
	case -8: s = npb_launch_aux_keys<8>(ch, a); if (s == NPB_OK) s = npb_launch_alg8_tile4<8>(ch, a); break;
	case -16: s = npb_launch_aux_keys<16>(ch, a); if (s == NPB_OK) s = npb_launch_alg8_tile4<16>(ch, a); break;
	// D = 64: block-wise tcgen05 density table + warp-per-chain race (npb_alg8_gemm.cu)
	case -64: s = npb_launch_alg8_gemm64(ch, a); break;
	case -1600: s = npb_launch_alg8_fused16(ch, a); break;
	case -1602: s = npb_launch_alg8_tc16(ch, a); break;
	case 2001: s = npb_launch_alg8_reg<2, 1>(ch, a); break;
	case 2002: s = npb_launch_alg8_reg<2, 2>(ch, a); break;
	case 2004: s = npb_launch_alg8_reg<2, 4>(ch, a); break;
	case 2008: s = npb_launch_alg8_reg<2, 8>(ch, a); break;
	case 2016: s = npb_launch_alg8_reg<2, 16>(ch, a); break;
	case 3001: s = npb_launch_alg8_reg<3, 1>(ch, a); break;
	case 3002: s = npb_launch_alg8_reg<3, 2>(ch, a); break;
	case 3004: s = npb_launch_alg8_reg<3, 4>(ch, a); break;
	case 3008: s = npb_launch_alg8_reg<3, 8>(ch, a); break;
	case 4001: s = npb_launch_alg8_tile<4, 32>(ch, a); break;
	case 4002: s = npb_launch_alg8_tile<4, 64>(ch, a); break;
	case 8001: s = npb_launch_alg8_tile<8, 32>(ch, a); break;
	case 8002: s = npb_launch_alg8_tile<8, 64>(ch, a); break;
	case 16001: s = npb_launch_alg8_tile<16, 32>(ch, a); break;
	case 16002: s = npb_launch_alg8_tile<16, 64>(ch, a); break;
	default:
		return npb_fail(ctx, NPB_E_UNSUPPORTED,
				"Alg. 8 sweep kernels cover D = 2 (Kmax 32..512), D = 3 (Kmax 32..256), D = 4, 8, 16 (Kmax 32/64) and D = 64 (Kmax 32)");
	}
	if (s == NPB_OK) ch->sweep += (uint32_t)n_sweeps;
	return s;
}

npb_status npb_launch_alg8_sweep(npb_chains *ch, int n_sweeps) {
	npb_ctx *ctx = ch->ctx;
	if (!ch->scan_order) {
		// as many sweeps per launch as fit 64 MB of scan orders
		const size_t per_sweep = (size_t)ch->ds->N * sizeof(int32_t);
		size_t cap = (64u << 20) / per_sweep;
		ch->scan_cap = (int)(cap < 1 ? 1 : (cap > 1024 ? 1024 : cap));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->scan_order, per_sweep * ch->scan_cap));
	}
	if (ctx->prior.family) { // `-c regression` / `-c angular`: npb_scalarnoise.cu
		for (int done = 0; done < n_sweeps;) {
			const int n = (n_sweeps - done < ch->scan_cap) ? n_sweeps - done : ch->scan_cap;
			npb_status s = npb_launch_sn_sweep(ch, n);
			if (s != NPB_OK) return s;
			done += n;
		}
		return NPB_OK;
	}
	const bool two_warp = ch->sw.two_warp;
	if ((!two_warp || ch->D == 64) && ch->Kmax == 32 && ch->D >= 4 && !ch->aux_keys) {
		// as many sweeps per launch as fit 4 GB of auxiliary keys (one 32-bit word per chain and step)
		const size_t per_sweep = (size_t)ch->ds->N * ch->C * sizeof(uint32_t);
		size_t cap = ((size_t)4 << 30) / per_sweep;
		ch->aux_cap = (int)(cap < 1 ? 1 : (cap > (size_t)ch->scan_cap ? (size_t)ch->scan_cap : cap));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->aux_keys, per_sweep * ch->aux_cap));
		if (ch->D == 16) NPB_CUDA_OK(cudaMalloc((void **)&ch->aux_max, (size_t)ch->aux_cap * ch->C * ((ch->ds->N + 31) / 32) * sizeof(float)));
	}
	const int per_launch = ch->aux_keys ? ch->aux_cap : ch->scan_cap;
	for (int done = 0; done < n_sweeps;) {
		const int n = (n_sweeps - done < per_launch) ? n_sweeps - done : per_launch;
		npb_status s = launch_chunk(ch, n);
		if (s != NPB_OK) return s;
		done += n;
	}
	return NPB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// The single-item seam: one NealAlgorithm8::update(membertrix&, {item}) (np_neal_algorithm8.cpp:49-167) on chains
// chain0 .. chain0+n-1, one warp per chain, any D <= NPB_MAX_D and any Kmax.  Not a throughput path: parameters are read
// from the global slot table, the auxiliary draws are materialised in full (npb_draw_theta) into shared memory.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_update_item_alg8(SweepArgs a, int chain0, int item, uint32_t call) {
	extern __shared__ float s_aux[]; // [M][PS]
	const int lane = threadIdx.x, chain = chain0 + blockIdx.x;
	const int D = a.prior.D, PS = npb_ps(D), TRI = npb_tri(D), M = a.prior.m_aux, Kmax = a.Kmax;
	const Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
	float *theta = a.theta + (size_t)chain * Kmax * PS;
	int *counts = a.counts + (size_t)chain * Kmax;
	const float *x = a.X + (size_t)item * D;
	const int zo = (int)a.z[(size_t)item * a.C + chain];
	for (int m = lane; m < M; m += 32) npb_draw_theta(a.prior, ph, (uint32_t)item, 0x51A61E00u + call, NPB_RNG_AUX, m * (D + 1), s_aux + m * PS);
	__syncwarp();
	auto log2dens = [&](const float *p) {
		float q = 0.0f;
		for (int r = 0; r < D; ++r) {
			float y = 0.0f;
			for (int c = r; c < D; ++c) y = fmaf(p[D + npb_tri_off(D, r, c)], x[c] - p[c], y);
			q = fmaf(y, y, q);
		}
		return p[D + TRI] - q;
	};
	// exponential race over the occupied slots (weights p n_k, the item itself retracted) and the M auxiliary draws
	float best = -INFINITY;
	int best_c = -1, cand = 0;
	for (int c0 = 0; c0 < Kmax + M; c0 += 32) {
		const int c = c0 + lane;
		float key = -INFINITY;
		if (c < Kmax + M) {
			uint32_t w[4];
			ph((uint32_t)item, (uint32_t)c, 0x51A61E00u + call, NPB_RNG_PICK, w);
			const float noise = neg_lg2_exp1(w[0]);
			if (c < Kmax) {
				const int n = counts[c] - (c == zo ? 1 : 0);
				if (n > 0) { key = log2dens(theta + (size_t)c * PS) + fast_lg2((float)n) + noise; cand++; }
			} else {
				key = log2dens(s_aux + (c - Kmax) * PS) + a.prior.log2_alpha_m + noise;
				cand++;
			}
		}
		if (key > best) { best = key; best_c = c; }
	}
	// warp arg-max (ties: lowest candidate index)
	for (int o = 16; o > 0; o >>= 1) {
		const float ob = __shfl_xor_sync(0xffffffffu, best, o);
		const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
		if (ob > best || (ob == best && oc >= 0 && (best_c < 0 || oc < best_c))) { best = ob; best_c = oc; }
	}
	cand = __reduce_add_sync(0xffffffffu, cand);
	if (best_c < 0) return; // nothing had weight: the item stays (the reference would pick index 0, Q6)
	int new_slot = best_c;
	bool born = false;
	if (best_c >= Kmax) { // np_neal_algorithm8.cpp:136-145: lowest free slot once the item is retracted
		born = true;
		int fs = -1;
		for (int k0 = 0; k0 < Kmax && fs < 0; k0 += 32) {
			const int k = k0 + lane;
			const bool free_k = k < Kmax && (counts[k] - (k == zo ? 1 : 0)) <= 0;
			const unsigned fb = __ballot_sync(0xffffffffu, free_k);
			if (fb) fs = k0 + __ffs(fb) - 1;
		}
		if (fs < 0) { if (lane == 0) a.overflow[chain] = 1; return; }
		new_slot = fs;
		const float *src = s_aux + (best_c - Kmax) * PS;
		for (int t = lane; t < PS; t += 32) theta[(size_t)fs * PS + t] = src[t];
	}
	__syncwarp();
	if (lane == 0) {
		const int before = counts[zo];
		if (new_slot != zo || born) {
			counts[zo] = before - 1;
			counts[new_slot] += 1;
			a.z[(size_t)item * a.C + chain] = (npb_z_t)new_slot;
			int occ = a.kocc[chain];
			if (before - 1 == 0 && new_slot != zo) occ--;
			if (born && !(new_slot == zo)) occ++;
			else if (born && new_slot == zo) occ += 0; // the emptied singleton's slot is re-used by the newborn
			a.kocc[chain] = occ;
			a.st[(size_t)chain * 4 + 1] += 1ull;
			if (born) a.st[(size_t)chain * 4 + 2] += 1ull;
		}
		a.st[(size_t)chain * 4 + 0] += (unsigned long long)cand;
	}
}

npb_status npb_launch_update_item(npb_chains *ch, int64_t chain0, int64_t n, int64_t item) {
	npb_ctx *ctx = ch->ctx;
	SweepArgs a = make_args(ch, 0);
	const size_t shmem = (size_t)ch->m_aux * npb_ps(ch->D) * sizeof(float);
	k_update_item_alg8<<<(unsigned)n, 32, shmem, ctx->stream>>>(a, (int)chain0, (int)item, ch->item_calls++);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}

npb_status npb_launch_tile_probe(npb_chains *ch, int chain, const int32_t *d_items, float *d_out) {
	SweepArgs a = make_args(ch, 0);
	if (ch->Kmax != 32) return npb_fail(ch->ctx, NPB_E_UNSUPPORTED, "the tile probe covers the Kmax = 32 kernels");
	if (ch->D == 16) {
		const char *e = ch->opt_d16_path;
		if (e[0] == 't' && e[1] == 'c' && e[2] == '2') return npb_launch_tc16_probe(ch, chain, d_items, d_out);
		if (e[0] != 'f') return npb_launch_fused16_probe(ch, chain, d_items, d_out);
	}
	switch (ch->D) {
	case 64: return npb_launch_gemm64_probe(ch, chain, d_items, d_out);
	case 4: return npb_launch_tile4_probe<4>(ch, a, chain, d_items, d_out);
	case 8: return npb_launch_tile4_probe<8>(ch, a, chain, d_items, d_out);
	case 16: return npb_launch_tile4_probe<16>(ch, a, chain, d_items, d_out);
	default: return npb_fail(ch->ctx, NPB_E_UNSUPPORTED, "the tile probe covers D = 4, 8, 16, 64");
	}
}
