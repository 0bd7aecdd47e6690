// npb_splitmerge.cu -- Jain-Neal and triadic split-merge proposals for thousands of lockstep chains (sm_100a).
//
// Replaces, per chain and per subset of 2 / 3 items, JainNealAlgorithm::update
// (src/np_jain_neal_algorithm.cpp:424-502; split :215-314, merge :316-409, propose_split :91-185) and
// TriadicAlgorithm::update (src/np_triadic_algorithm.cpp:633-795; split :298-464, merge :470-618, propose_split
// :208-291, propose_merge :135-205, ratioStateProb :78-94, ratioR :116-129) inside the sweep loop of MCMC::run
// (np_mcmc.cpp:109-163, subsets with a repeated item skipped, :155-158).
//
// One CTA per chain.  A proposal is
//   plan     thread 0 reads the slots of the picks and decides split / merge exactly like the reference's update();
//   pool     the member lists of the source clusters (membertrix::getAssignments, membertrix.cpp:315-322): the chain keeps
//            every slot's items contiguous in a permutation array, rebuilt from its row of the (chain-major) assignment
//            copy only after an accepted move; the pool is visited in a keyed pseudo-random order (the reference shuffles
//            the list, dim1algebra.hpp:2066-2073);
//   phase A  all threads: log2-densities of a chunk of pool members under the <= 3 parameter sets involved
//            (sources + the fresh prior draw) -- the FP32 work, D(D+1)/2 + 2D FMAs per (member, theta) -- and the
//            per-cluster sums  sum_{x in c} log p(x|theta_c)  of the acceptance ratio (fp64 accumulators);
//   phase B  warp 0: the restricted sequential scan (SAMS): member t goes to part q with probability proportional to
//            p(x_t|theta_q) |P_q|, where |P_q| depends on the members before it -- inherently sequential, but all it
//            needs per member is Q precomputed numbers and one uniform, prepared 32 members at a time (lane = member);
//   accept   thread 0: log acceptance = prior ratio (lgamma terms) + rR + likelihood ratio; on accept all threads
//            rewrite the assignments of the pool, and thread 0 updates counts / theta / occupied count.
// Differences from the reference, on purpose: SAMS weights are normalised in the log2 domain (the reference multiplies
// linear-domain doubles, which underflow to 0 for far clusters); random streams are Philox keyed by (seed, chain)
// with counters (proposal, member, sweep, purpose).  Bug-compatible on purpose: the Jain-Neal SAMS rule
// "log-density PLUS member count fed to a cumulative-sum pick" including its lower_bound behaviour on a decreasing
// cumulative sum, and the acceptance ratios without the SAMS proposal probability.
#include "npb_internal.h"
#include "npb_alg8_kernel.cuh"

#define SM_THREADS 64
#define SM_MB 2                      // members per thread per chunk (register blocking of the theta reads)
#define SM_CHUNK (SM_THREADS * SM_MB) // pool members per chunk

enum { SM_JN_SPLIT = 0, SM_JN_MERGE = 1, SM_TRI_SPLIT = 2, SM_TRI_MERGE = 3 };

struct SMArgs {
	const float *X;          // [N, D]
	npb_z_t *zt;             // [C, N] chain-major working copy of the assignments
	float *theta;            // [C, Kmax, PS]
	int *counts;             // [C, Kmax]
	int *kocc, *overflow;    // [C]
	unsigned long long *smst; // [C, 12]: attempts[4], accepts[4], sams allocations, proposals done, 2 reserved
	const int32_t *order;    // [3, N] scan orders of the sweep (np_mcmc.cpp:120-125: one permutation per subset position)
	int32_t *pool;           // [C, N] scratch: pool members, ascending item id
	uint8_t *dec;            // [C, N] scratch: part chosen for the member visited at position t
	float *detail;           // optional [C, 16] per-proposal detail of the LAST proposal (tests), may be NULL
	int N, C, Kmax, sampler, s0, s1, zstride;
	uint32_t sweep;
	uint64_t seed;
	double alpha;            // Dirichlet-process concentration (np_main.cpp:164)
	int jn_bugcompat;        // Q8 rule on (reference behaviour) / off (linear-domain SAMS like the triadic sampler)
	int seq_scan;            // 1: the restricted scan one member at a time (option spec = 0); 0: 32 members at a time, to a fixed point (same result)
	PriorDev prior;
};

struct SMPlan {
	int type, nth, nsrc, Q, nskip;
	int picks[3];      // picks[q], q < nskip, seeds part q and is skipped in the pool
	int th_slot[3];    // slot behind theta k (-1: the fresh prior draw)
	int tgt_slot[3];   // slot part q is written to on accept (-1: the new cluster)
	int dying;         // slot removed by an accepted merge, -1 otherwise
	int stat;          // 0 merge 2->1, 1 split 1->2, 2 merge 3->2, 3 split 2->3
	int valid;
};

template <int D>
__device__ __forceinline__ void sm_load_row(const float *row, float (&x)[D]) {
	if constexpr (D % 4 == 0) {
#pragma unroll
		for (int c = 0; c < D / 4; ++c) {
			const float4 v = __ldg(reinterpret_cast<const float4 *>(row) + c);
			x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
		}
	} else if constexpr (D == 2) {
		const float2 v = __ldg(reinterpret_cast<const float2 *>(row));
		x[0] = v.x; x[1] = v.y;
	} else {
#pragma unroll
		for (int c = 0; c < D; ++c) x[c] = __ldg(row + c);
	}
}

template <int D>
__device__ __forceinline__ void sm_log2density4(const float *th /* shared: mu, T2, c2 */, const float (&x)[SM_MB][D], float (&out)[SM_MB]) {
	constexpr int TRI = npb_tri(D);
	float d[SM_MB][D], q[SM_MB];
#pragma unroll
	for (int m = 0; m < SM_MB; ++m) {
		q[m] = 0.0f;
#pragma unroll
		for (int c = 0; c < D; ++c) d[m][c] = x[m][c] - th[c];
	}
#pragma unroll
	for (int r = 0; r < D; ++r) {
		float y[SM_MB];
#pragma unroll
		for (int m = 0; m < SM_MB; ++m) y[m] = 0.0f;
#pragma unroll
		for (int c = r; c < D; ++c) {
			const float t = th[D + npb_tri_off(D, r, c)];
#pragma unroll
			for (int m = 0; m < SM_MB; ++m) y[m] = fmaf(t, d[m][c], y[m]);
		}
#pragma unroll
		for (int m = 0; m < SM_MB; ++m) q[m] = fmaf(y[m], y[m], q[m]);
	}
#pragma unroll
	for (int m = 0; m < SM_MB; ++m) out[m] = th[D + TRI] - q[m];
}

__device__ __forceinline__ double sm_block_sum(double v, double *red /* [SM_THREADS/32] shared */) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	double s = 0.0;
	for (int w = 0; w < SM_THREADS / 32; ++w) s += red[w];
	return s;
}

template <int D>
__global__ void __launch_bounds__(SM_THREADS, 16) k_split_merge(SMArgs a) {
	constexpr int PS = npb_ps(D);
	__shared__ float s_th[3][PS + 3];
	__shared__ float s_ld[3][SM_CHUNK];
	__shared__ int s_id[SM_CHUNK];
	__shared__ SMPlan s_plan;
	__shared__ ScanOrder s_perm;
	__shared__ int s_npool, s_accept, s_newslot;
	__shared__ int s_npart[3];
	__shared__ double s_S[3][3];
	__shared__ double s_red[SM_THREADS / 32];
	extern __shared__ int s_off[]; // [Kmax + 1] start of every slot's member list in perm (rebuilt after an accepted move)

	const int chain = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int N = a.N;
	npb_z_t *zc = a.zt + (size_t)chain * a.zstride;
	int32_t *pool = a.pool + (size_t)chain * N;
	uint8_t *dec = a.dec + (size_t)chain * N;
	float *theta = a.theta + (size_t)chain * a.Kmax * PS;
	int *counts = a.counts + (size_t)chain * a.Kmax;
	unsigned long long *st = a.smst + (size_t)chain * 12;
	const Philox ph((uint32_t)a.seed, (uint32_t)(a.seed >> 32) + (uint32_t)chain);
	const uint32_t c3 = NPB_RNG_SM | (a.sweep << 8);

	// member lists of the chain: offsets from the counts, items in ascending order from the chain's row of z
	// (deterministic: a batch of 32 items is grouped by slot with match.any, the group's first lane advances the cursor)
	auto rebuild_lists = [&]() {
		if (warp == 0) {
			int run = 0;
			for (int k0 = 0; k0 < a.Kmax; k0 += 32) {
				const int c = counts[k0 + lane];
				int incl = c;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const int t = __shfl_up_sync(0xffffffffu, incl, o);
					if (lane >= o) incl += t;
				}
				s_off[k0 + lane] = run + incl - c;
				run += __shfl_sync(0xffffffffu, incl, 31);
			}
			__syncwarp();
			for (int i0 = 0; i0 < N; i0 += 32) {
				const int i = i0 + lane;
				const unsigned active = __ballot_sync(0xffffffffu, i < N);
				if (i < N) {
					const int zz = zc[i];
					const unsigned grp = __match_any_sync(active, zz);
					const int rank = __popc(grp & ((1u << lane) - 1u));
					const int b = s_off[zz];
					pool[b + rank] = i;
					__syncwarp(active);
					if (rank == 0) s_off[zz] = b + __popc(grp);
				}
				__syncwarp();
			}
			// the cursors ran to the end of their segments: back to the starts
			for (int k0 = 0; k0 < a.Kmax; k0 += 32) s_off[k0 + lane] -= counts[k0 + lane];
		}
		__syncthreads();
	};
	rebuild_lists();

	for (int s = a.s0; s < a.s1; ++s) {
		// ---------------- plan (np_jain_neal_algorithm.cpp:424-502 / np_triadic_algorithm.cpp:633-795) ----------------
		if (tid == 0) {
			SMPlan p;
			p.valid = 1; p.dying = -1; p.nskip = 0;
			const int nsub = a.sampler == NPB_JAIN_NEAL ? 2 : 3;
			int pk[3], cl[3];
			for (int j = 0; j < nsub; ++j) pk[j] = a.order[(size_t)j * N + s];
			if (pk[0] == pk[1] || (nsub == 3 && (pk[0] == pk[2] || pk[1] == pk[2]))) p.valid = 0; // np_mcmc.cpp:155-158
			if (p.valid) {
				for (int j = 0; j < nsub; ++j) cl[j] = zc[pk[j]];
				if (a.sampler == NPB_JAIN_NEAL) {
					if (cl[0] == cl[1]) { // split: move starts with data_i -> new, remain with data_j -> current
						p.type = SM_JN_SPLIT; p.stat = 1; p.nth = 2; p.nsrc = 1; p.Q = 2; p.nskip = 2;
						p.th_slot[0] = cl[0]; p.th_slot[1] = -1;
						p.tgt_slot[0] = cl[0]; p.tgt_slot[1] = -1;
						p.picks[0] = pk[1]; p.picks[1] = pk[0];
					} else { // merge: every member of cluster(data_i) goes to cluster(data_j)
						p.type = SM_JN_MERGE; p.stat = 0; p.nth = 2; p.nsrc = 1; p.Q = 1; p.nskip = 0;
						p.th_slot[0] = cl[0]; p.th_slot[1] = cl[1];
						p.tgt_slot[0] = cl[1];
						p.dying = cl[0];
					}
				} else {
					const int uniq = 1 + (cl[1] != cl[0]) + (cl[2] != cl[0] && cl[2] != cl[1]);
					uint32_t w[4];
					ph((uint32_t)s, 0xFFFFFFF0u, 0u, c3, w);
					const float u0 = npb_u01(w[0]); // np_triadic_algorithm.cpp:675, always drawn
					// duplicate_pick (dim1algebra.hpp:2115-2137): first index equal to an earlier one, else the last
					const int dup = (cl[1] == cl[0]) ? 1 : 2;
					if (uniq == 1) { // :676-700 split 1 -> 2 on picks {0, 2}
						p.type = SM_TRI_SPLIT; p.stat = 1; p.nth = 2; p.nsrc = 1; p.Q = 2; p.nskip = 2;
						p.th_slot[0] = cl[0]; p.th_slot[1] = -1;
						p.tgt_slot[0] = cl[0]; p.tgt_slot[1] = -1;
						p.picks[0] = pk[0]; p.picks[1] = pk[2];
					} else if (u0 < 0.5f) { // :701-727 merge 2 -> 1 after erasing the duplicate pick
						int a0 = 0, a1 = 1;
						if (dup == 1) a1 = 2;
						p.type = SM_TRI_MERGE; p.stat = 0; p.nth = 2; p.nsrc = 2; p.Q = 1; p.nskip = 1;
						p.th_slot[0] = cl[a0]; p.th_slot[1] = cl[a1];
						p.tgt_slot[0] = cl[a0];
						p.picks[0] = pk[a0];
						p.dying = cl[a1];
					} else if (uniq == 2) { // :738-765 split 2 -> 3, duplicate pick swapped to the end
						int o0 = 0, o1 = 1, o2 = 2;
						if (dup == 1) { o1 = 2; o2 = 1; }
						p.type = SM_TRI_SPLIT; p.stat = 3; p.nth = 3; p.nsrc = 2; p.Q = 3; p.nskip = 3;
						p.th_slot[0] = cl[o0]; p.th_slot[1] = cl[o1]; p.th_slot[2] = -1;
						p.tgt_slot[0] = cl[o0]; p.tgt_slot[1] = cl[o1]; p.tgt_slot[2] = -1;
						p.picks[0] = pk[o0]; p.picks[1] = pk[o1]; p.picks[2] = pk[o2];
					} else { // :766-786 merge 3 -> 2: the third cluster is dissolved, its pick re-allocated
						p.type = SM_TRI_MERGE; p.stat = 2; p.nth = 3; p.nsrc = 3; p.Q = 2; p.nskip = 2;
						p.th_slot[0] = cl[0]; p.th_slot[1] = cl[1]; p.th_slot[2] = cl[2];
						p.tgt_slot[0] = cl[0]; p.tgt_slot[1] = cl[1];
						p.picks[0] = pk[0]; p.picks[1] = pk[1];
						p.dying = cl[2];
					}
				}
			}
			s_plan = p;
			s_npool = 0;
			for (int q = 0; q < 3; ++q) {
				s_npart[q] = 0;
				for (int k = 0; k < 3; ++k) s_S[q][k] = 0.0;
			}
		}
		__syncthreads();
		const SMPlan p = s_plan;
		if (!p.valid) { __syncthreads(); continue; }

		// ---------------- thetas: sources from the chain's table, the new one from the base measure ----------------
		for (int k = 0; k < p.nth; ++k) {
			if (p.th_slot[k] >= 0) {
				for (int t = tid; t < PS; t += SM_THREADS) s_th[k][t] = theta[(size_t)p.th_slot[k] * PS + t];
			} else if (tid == 0) {
				npb_draw_theta(a.prior, ph, (uint32_t)s, 0xFFFFFFF1u, c3, 0, s_th[k]); // sample_base, :233 / :345
			}
		}
		if (tid == 32) {
			// visiting order of the pool: keyed permutation, re-keyed per (chain, proposal)
			uint32_t w[4];
			ph((uint32_t)s, 0xFFFFFFF2u, 0u, c3, w);
			s_perm = npb_scan_order(((uint64_t)w[1] << 32) | w[0], w[2], 2u);
		}

		// ---------------- pool: the member lists of the source clusters, one after the other (membertrix::getAssignments
		// per cluster, np_triadic_algorithm.cpp:230-238); perm[s_off[k] .. s_off[k] + counts[k]) holds slot k's items ----------------
		int seg_start[3] = {0, 0, 0}, seg_end[3] = {0, 0, 0};
		int base = 0;
		for (int k = 0; k < p.nsrc; ++k) {
			seg_start[k] = s_off[p.th_slot[k]] - base; // pool position t of source k maps to perm[t + seg_start[k]]
			base += counts[p.th_slot[k]];
			seg_end[k] = base;
		}
		auto pool_item = [&](int t) -> int {
			const int k = t < seg_end[0] ? 0 : (t < seg_end[1] ? 1 : 2);
			return pool[t + seg_start[k]];
		};
		const int npool = base;
		__syncthreads(); // the keys of s_perm (thread 32) are in place before its size is patched (thread 0)
		if (tid == 0) {
			s_perm.N = (uint32_t)npool;
			uint32_t bits = 2;
			while ((1u << bits) < (uint32_t)npool) bits++;
			s_perm.half_bits = (bits + 1) / 2;
			s_perm.half_mask = (1u << s_perm.half_bits) - 1u;
		}
		__syncthreads();

		// ---------------- chunks: phase A (densities, all threads) + phase B (sequential scan, warp 0) ----------------
		double own[3] = {0.0, 0.0, 0.0};   // sum over members of source c of log2 p(x|theta_c)
		double q1[3] = {0.0, 0.0, 0.0};    // Q == 1: sum over the pool of log2 p(x|theta_k)
		float npart[3] = {0.0f, 0.0f, 0.0f};
		double S[3][3];
#pragma unroll
		for (int q = 0; q < 3; ++q)
#pragma unroll
			for (int k = 0; k < 3; ++k) S[q][k] = 0.0;
		if (p.type != SM_JN_MERGE)
			for (int q = 0; q < p.nskip; ++q) npart[q] = 1.0f;
		unsigned long long sams = 0;
		const ScanOrder perm = s_perm;

		for (int t0 = 0; t0 < npool; t0 += SM_CHUNK) {
			const int cnt = min(SM_CHUNK, npool - t0);
			// phase A
			{
				float x[SM_MB][D];
				int id[SM_MB], ownk[SM_MB];
#pragma unroll
				for (int m = 0; m < SM_MB; ++m) {
					const int j = tid + m * SM_THREADS;
					id[m] = -1; ownk[m] = -1;
					if (j < cnt) {
						id[m] = pool_item((int)npb_scan_item(perm, (uint32_t)(t0 + j)));
						const int zz = zc[id[m]];
						ownk[m] = zz == p.th_slot[0] ? 0 : (zz == p.th_slot[1] ? 1 : 2);
					}
					sm_load_row<D>(a.X + (size_t)(id[m] < 0 ? 0 : id[m]) * D, x[m]);
				}
				for (int k = 0; k < p.nth; ++k) {
					float l[SM_MB];
					sm_log2density4<D>(s_th[k], x, l);
#pragma unroll
					for (int m = 0; m < SM_MB; ++m) {
						const int j = tid + m * SM_THREADS;
						if (j < cnt) {
							s_ld[k][j] = l[m];
							if (ownk[m] == k) own[k] += (double)l[m];
							if (p.Q == 1) q1[k] += (double)l[m];
						}
					}
				}
#pragma unroll
				for (int m = 0; m < SM_MB; ++m) {
					const int j = tid + m * SM_THREADS;
					if (j < cnt) s_id[j] = id[m];
				}
			}
			__syncthreads();
			// phase B
			if (p.Q == 1) {
				for (int j = tid; j < cnt; j += SM_THREADS) dec[t0 + j] = 0;
			} else if (warp == 0) {
				for (int j0 = 0; j0 < cnt; j0 += 32) {
					const int my = j0 + lane;
					const bool ok = my < cnt;
					const int id = ok ? s_id[my] : -1;
					float l0 = ok ? s_ld[0][my] : 0.0f, l1 = ok ? s_ld[1][my] : 0.0f;
					float l2 = (ok && p.nth > 2) ? s_ld[2][my] : -INFINITY;
					uint32_t w[4];
					ph((uint32_t)s, (uint32_t)(t0 + my), 1u, c3, w);
					const float u = npb_u01(w[0]);
					float e0, e1, e2;
					if (p.type == SM_JN_SPLIT && a.jn_bugcompat) {
						e0 = l0 * NPB_LN2; e1 = l1 * NPB_LN2; e2 = 0.0f; // natural-log densities (:158,165)
					} else {
						const float lq2 = p.Q > 2 ? l2 : -INFINITY;
						const float mx = fmaxf(fmaxf(l0, l1), lq2);
						e0 = exp2f(l0 - mx); e1 = exp2f(l1 - mx); e2 = p.Q > 2 ? exp2f(lq2 - mx) : 0.0f;
					}
					int skipq = -1;
					for (int q = 0; q < p.nskip; ++q)
						if (id == p.picks[q]) skipq = q;
					int mydec = 0;
					const int lim = min(32, cnt - j0);
					if (!a.seq_scan) {
						// 32 members at a time.  A member's part depends on the earlier members only through the part sizes, so every
						// lane decides with the sizes implied by the earlier lanes' CURRENT decisions (ballot + popc), and the warp
						// repeats until no decision changes: at that fixed point every lane's decision is the one the sequential scan
						// makes (induction over the lanes; lane j is final after at most j + 1 rounds, in practice two or three:
						// one more member rarely tips a draw).  Same arithmetic per decision as the loop below.
						const bool live = my < cnt && skipq < 0;
						const unsigned lt = (1u << lane) - 1u;
						int d = skipq >= 0 ? skipq : 0;
						for (int round = 0; round < 34; ++round) {
							const unsigned m0 = __ballot_sync(0xffffffffu, live && d == 0), m1 = __ballot_sync(0xffffffffu, live && d == 1);
							const unsigned m2 = __ballot_sync(0xffffffffu, live && d == 2);
							const float n0 = npart[0] + (float)__popc(m0 & lt), n1 = npart[1] + (float)__popc(m1 & lt), n2 = npart[2] + (float)__popc(m2 & lt);
							int dn;
							if (p.type == SM_JN_SPLIT && a.jn_bugcompat) {
								const float c0 = e0 + n0, c1 = c0 + (e1 + n1);
								const float thr = u * c1;
								dn = (c1 < thr) ? 1 : ((c0 < thr) ? 1 : 0);
							} else {
								const float w0 = e0 * n0, w1 = e1 * n1, w2 = e2 * n2;
								const float c0 = w0, c1 = w0 + w1, c2 = c1 + w2;
								const float thr = u * (p.Q > 2 ? c2 : c1);
								dn = (c0 >= thr) ? 0 : ((c1 >= thr || p.Q == 2) ? 1 : 2);
							}
							const bool changed = live && dn != d;
							if (live) d = dn;
							if (!__any_sync(0xffffffffu, changed)) break;
						}
						{
							const unsigned m0 = __ballot_sync(0xffffffffu, live && d == 0), m1 = __ballot_sync(0xffffffffu, live && d == 1);
							const unsigned m2 = __ballot_sync(0xffffffffu, live && d == 2);
							npart[0] += (float)__popc(m0); npart[1] += (float)__popc(m1); npart[2] += (float)__popc(m2);
							sams += (unsigned long long)__popc(m0 | m1 | m2);
						}
						mydec = d;
					} else
					for (int j = 0; j < lim; ++j) {
						const int bskip = __shfl_sync(0xffffffffu, skipq, j);
						const float b0 = __shfl_sync(0xffffffffu, e0, j), b1 = __shfl_sync(0xffffffffu, e1, j);
						const float b2 = __shfl_sync(0xffffffffu, e2, j), bu = __shfl_sync(0xffffffffu, u, j);
						int d;
						if (bskip >= 0) {
							d = bskip; // a pick: it seeded its part already
						} else {
							if (p.type == SM_JN_SPLIT && a.jn_bugcompat) {
								// weights [logp(x|cur) + |remain|, logp(x|new) + |move|], cumulative sum, u * total,
								// std::lower_bound on two entries (decreasing sums fall through to "move")
								const float c0 = b0 + npart[0], c1 = c0 + (b1 + npart[1]);
								const float thr = bu * c1;
								d = (c1 < thr) ? 1 : ((c0 < thr) ? 1 : 0);
							} else {
								const float w0 = b0 * npart[0], w1 = b1 * npart[1], w2 = b2 * npart[2];
								const float c0 = w0, c1 = w0 + w1, c2 = c1 + w2;
								const float thr = bu * (p.Q > 2 ? c2 : c1);
								d = (c0 >= thr) ? 0 : ((c1 >= thr || p.Q == 2) ? 1 : 2);
							}
							npart[0] += d == 0; npart[1] += d == 1; npart[2] += d == 2;
							sams++;
						}
						if (lane == j) mydec = d;
					}
					// the likelihood sums of the acceptance ratio: every lane adds ITS member's log-densities to the part
					// that member went to (lane-private partial sums, combined after the last chunk) -- one step per 32
					// members instead of one per member on the sequential path
					if (ok) {
#pragma unroll
						for (int q = 0; q < 3; ++q)
							if (mydec == q) {
								S[q][0] += (double)l0; S[q][1] += (double)l1;
								if (p.nth > 2) S[q][2] += (double)l2;
							}
						dec[t0 + my] = (uint8_t)mydec;
					}
				}
			}
			__syncthreads();
		}
		// ---------------- acceptance ----------------
		if (warp == 0) { // lane-private partial sums of phase B -> every lane of warp 0 holds the totals
#pragma unroll
			for (int q = 0; q < 3; ++q)
#pragma unroll
				for (int k = 0; k < 3; ++k)
#pragma unroll
					for (int o = 16; o > 0; o >>= 1) S[q][k] += __shfl_xor_sync(0xffffffffu, S[q][k], o);
		}
		double ownsum[3], q1sum[3];
		for (int k = 0; k < 3; ++k) {
			ownsum[k] = sm_block_sum(own[k], s_red);
			q1sum[k] = p.Q == 1 ? sm_block_sum(q1[k], s_red) : 0.0;
		}
		if (tid == 0) {
			const double LN2 = 0.6931471805599453;
			double logA = 0.0;
			int np[3] = {(int)npart[0], (int)npart[1], (int)npart[2]};
			int nsrc_cnt[3];
			for (int k = 0; k < p.nsrc; ++k) nsrc_cnt[k] = counts[p.th_slot[k]];
			const double la = log(a.alpha);
			if (p.type == SM_JN_SPLIT) {
				// move = part 1 (new), remain = part 0; lsrc / ldest over the moved items only (:256-272)
				logA = la + lgamma((double)np[1]) + lgamma((double)np[0]) - lgamma((double)(np[0] + np[1])) + (S[1][1] - S[1][0]) * LN2;
			} else if (p.type == SM_JN_MERGE) {
				const int n0 = nsrc_cnt[0], n1 = counts[p.th_slot[1]];
				np[0] = n0 + n1;
				logA = -(la + lgamma((double)n0) + lgamma((double)n1) - lgamma((double)(n0 + n1))) + (q1sum[1] - q1sum[0]) * LN2;
			} else {
				if (p.Q == 1) { np[0] = npool; S[0][0] = q1sum[0]; }
				double lg = 0.0, rL = 0.0;
				for (int q = 0; q < p.Q; ++q) { lg += lgamma((double)np[q]); rL += S[q][q]; }
				for (int k = 0; k < p.nsrc; ++k) { lg -= lgamma((double)nsrc_cnt[k]); rL -= ownsum[k]; }
				const double rP = la + lg; // split: + ; merge: the same expression with the roles swapped => sign below
				double rR;
				if (p.type == SM_TRI_SPLIT) rR = p.nsrc == 1 ? log(0.5) : -log(0.5);
				else rR = p.nsrc == 2 ? -log(0.5) : log(0.5);
				// merge: rP = -(la + sum_c lgamma n_c - sum_q lgamma |P_q|) = -la + lg   (:78-94 with split == false)
				logA = (p.type == SM_TRI_SPLIT ? rP : (lg - la)) + rR + rL * LN2;
			}
			uint32_t w[4];
			ph((uint32_t)s, 0xFFFFFFF3u, 0u, c3, w);
			const double u = (double)npb_u01(w[0]);
			int accept = !(exp(logA) < u);
			int newslot = -1;
			if (accept && (p.type == SM_JN_SPLIT || p.type == SM_TRI_SPLIT)) {
				for (int k = 0; k < a.Kmax; ++k)
					if (counts[k] == 0) { newslot = k; break; }
				if (newslot < 0) { accept = 0; a.overflow[chain] = 1; }
			}
			st[p.stat] += 1ull;
			st[4 + p.stat] += (unsigned long long)accept;
			// a 2 -> 1 merge still walks the pool through the (single-weight) pick in the reference (:159-199)
			st[8] += (p.type == SM_TRI_MERGE && p.Q == 1) ? (unsigned long long)(npool - p.nskip) : sams;
			st[9] += 1ull;
			if (accept) {
				for (int q = 0; q < p.Q; ++q) {
					const int slot = p.tgt_slot[q] >= 0 ? p.tgt_slot[q] : newslot;
					counts[slot] = np[q];
				}
				if (p.dying >= 0) { counts[p.dying] = 0; a.kocc[chain] -= 1; }
				if (newslot >= 0) a.kocc[chain] += 1;
			}
			s_accept = accept;
			s_newslot = newslot;
			if (a.detail) {
				float *dt = a.detail + (size_t)chain * 16;
				dt[0] = (float)p.type; dt[1] = (float)p.stat; dt[2] = (float)logA; dt[3] = (float)accept;
				dt[4] = (float)np[0]; dt[5] = (float)np[1]; dt[6] = (float)np[2]; dt[7] = (float)npool;
				dt[8] = (float)newslot; dt[9] = (float)p.dying; dt[10] = (float)u; dt[11] = (float)p.Q;
			}
		}
		// S lives in warp 0's registers (identical on its lanes); thread 0 used it above.
		__syncthreads();
		if (s_accept) {
			const int newslot = s_newslot;
			int tg[3];
			for (int q = 0; q < 3; ++q) tg[q] = (q < p.Q) ? (p.tgt_slot[q] >= 0 ? p.tgt_slot[q] : newslot) : 0;
			for (int t = tid; t < npool; t += SM_THREADS) {
				const int id = pool_item((int)npb_scan_item(perm, (uint32_t)t));
				zc[id] = (npb_z_t)tg[dec[t]];
			}
			if (newslot >= 0)
				for (int t = tid; t < PS; t += SM_THREADS) theta[(size_t)newslot * PS + t] = s_th[p.nth - 1][t];
			__syncthreads();
			rebuild_lists(); // the partition changed
		}
		__syncthreads();
	}
}

// [N, C] item-major <-> [C, zstride] chain-major
__global__ void k_z_transpose(const npb_z_t *in, npb_z_t *out, int rows, int cols, int in_stride, int out_stride) {
	__shared__ npb_z_t tile[32][33];
	const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
	for (int r = threadIdx.y; r < 32; r += blockDim.y)
		if (r0 + r < rows && c0 + threadIdx.x < cols) tile[r][threadIdx.x] = in[(size_t)(r0 + r) * in_stride + c0 + threadIdx.x];
	__syncthreads();
	for (int c = threadIdx.y; c < 32; c += blockDim.y)
		if (c0 + c < cols && r0 + threadIdx.x < rows) out[(size_t)(c0 + c) * out_stride + r0 + threadIdx.x] = tile[threadIdx.x][c];
}

__global__ void k_scan_order3(int32_t *order, int N, uint64_t seed, uint32_t sweep) {
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= N) return;
	// one independent permutation per subset position (np_mcmc.cpp:120-125)
	const ScanOrder so = npb_scan_order(seed ^ (0x9E3779B97F4A7C15ull * (blockIdx.y + 1)), sweep, (uint32_t)N);
	order[(size_t)blockIdx.y * N + s] = (int32_t)npb_scan_item(so, (uint32_t)s);
}

npb_status npb_launch_split_merge(npb_chains *ch, int sampler, int64_t n_proposals, int whole_sweeps, float *d_detail) {
	npb_ctx *ctx = ch->ctx;
	const int N = (int)ch->ds->N, C = (int)ch->C;
	const int zstride = (N + 7) & ~7; // rows of the chain-major copy are read with 16-byte loads
	if (!ch->sm_zt) {
		NPB_CUDA_OK(cudaMalloc((void **)&ch->sm_zt, (size_t)zstride * C * sizeof(npb_z_t)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->sm_pool, (size_t)N * C * sizeof(int32_t)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->sm_dec, (size_t)N * C));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->sm_order, (size_t)3 * N * sizeof(int32_t)));
	}
	SMArgs a;
	memset(&a, 0, sizeof(a));
	a.X = ch->ds->X32;
	a.zt = ch->sm_zt;
	a.theta = ch->theta;
	a.counts = ch->counts;
	a.kocc = ch->kocc;
	a.overflow = ch->overflow;
	a.smst = ch->smst;
	a.order = ch->sm_order;
	a.pool = ch->sm_pool;
	a.dec = ch->sm_dec;
	a.detail = d_detail;
	a.N = N; a.C = C; a.Kmax = ch->Kmax; a.sampler = sampler; a.zstride = zstride;
	a.seed = ch->seed;
	a.jn_bugcompat = 1;
	a.seq_scan = ch->sw.spec == 0;
	a.alpha = ctx->prior.alpha;
	a.prior = npb_prior_dev(ctx, 1);
	dim3 tb(32, 8);
	dim3 g1((C + 31) / 32, (N + 31) / 32), g2((N + 31) / 32, (C + 31) / 32);
	k_z_transpose<<<g1, tb, 0, ctx->stream>>>(ch->z, ch->sm_zt, N, C, C, zstride);
	NPB_CUDA_OK(cudaGetLastError());
	const size_t smem_off = (size_t)(ch->Kmax + 1) * sizeof(int);
	int64_t left = whole_sweeps ? (int64_t)whole_sweeps * N : n_proposals;
	while (left > 0) {
		const int n = (int)(left < N ? left : N);
		dim3 go((N + 255) / 256, 3);
		k_scan_order3<<<go, 256, 0, ctx->stream>>>(ch->sm_order, N, ch->seed, ch->sweep);
		NPB_CUDA_OK(cudaGetLastError());
		a.s0 = 0; a.s1 = n; a.sweep = ch->sweep;
		switch (ch->D) {
		case 2: k_split_merge<2><<<C, SM_THREADS, smem_off, ctx->stream>>>(a); break;
		case 3: k_split_merge<3><<<C, SM_THREADS, smem_off, ctx->stream>>>(a); break;
		case 4: k_split_merge<4><<<C, SM_THREADS, smem_off, ctx->stream>>>(a); break;
		case 8: k_split_merge<8><<<C, SM_THREADS, smem_off, ctx->stream>>>(a); break;
		case 16: k_split_merge<16><<<C, SM_THREADS, smem_off, ctx->stream>>>(a); break;
		default: return npb_fail(ctx, NPB_E_UNSUPPORTED, "split-merge kernels cover D = 2, 3, 4, 8, 16");
		}
		NPB_CUDA_OK(cudaGetLastError());
		ch->sweep += 1;
		left -= n;
	}
	k_z_transpose<<<g2, tb, 0, ctx->stream>>>(ch->sm_zt, ch->z, C, N, zstride, C);
	NPB_CUDA_OK(cudaGetLastError());
	return NPB_OK;
}
