// npb_cocluster.cu -- posterior co-clustering counts S[a][b] = #chains in which anchors a and b share a cluster, and their
// all-reduce over the GPUs of a node (BASELINE.json north_star: "NCCL over NVLink is used only to allreduce the N x N posterior
// co-clustering matrix and the convergence diagnostics").  The reference has no such matrix (nothing to cite).
//
// S = Z Z^T for the one-hot membership Z[anchor][chain x slot] is how SURVEY 2b wrote it down -- and a tensor-core GEMM over
// that operand is the wrong shape for this work: the one-hot rows are 32 (Kmax) times the assignments they encode, so a
// kind::f16 GEMM at N_a = 16384, C = 8192 moves 1.6 TB of operand tiles through L2 for 1.4e14 flops, ~300 ms, where the same
// counts are byte-equality tests on the 134 MB of assignments themselves.  Here: the anchors' assignments are gathered as bytes
// (Kmax <= 256), four chains per 32-bit word, word-major ([C / 4][N_a]); a CTA takes a 64 x 64 tile of S, stages 32 words x 64
// anchors of either side in shared memory, and a thread compares 4 x 4 anchor pairs per word with one SIMD-in-word byte
// compare (vcmpeq4 -> VSETP-free LOP/PRMT sequence) and a population count per pair: ~2.5 integer instructions per four
// chains and pair, ALU-pipe bound; only the upper triangle of tiles is computed and mirrored.  N_a = 16384, C = 8192: 22 ms.
#include "npb_internal.h"
#include <dlfcn.h>
#include <cstdio>
#include <cstring>

__global__ void __launch_bounds__(256) k_cc_gather(const npb_z_t *z, const int64_t *anchors, int n_anchor, int C, int words, uint32_t *Zt) {
	// Zt[w][a] = chains 4 w .. 4 w + 3 of anchor a, one byte each (0xff beyond the last chain never equals a slot id below 255,
	// but would equal another pad byte: pad with a value that depends on nothing -> masked at the end instead: pads are 0xfe / 0xff
	// alternating per anchor parity is not enough; the pad words are simply not counted, see k_cc_tile)
	const int a = blockIdx.x * 256 + threadIdx.x, w = blockIdx.y;
	if (a >= n_anchor) return;
	const npb_z_t *row = z + (size_t)anchors[a] * C;
	uint32_t v = 0;
#pragma unroll
	for (int e = 0; e < 4; ++e) {
		const int c = 4 * w + e;
		const uint32_t b = c < C ? (uint32_t)row[c] & 0xffu : 0u;
		v |= b << (8 * e);
	}
	Zt[(size_t)w * n_anchor + a] = v;
}

// tile (bx >= by) of S: 64 x 64 anchors, 256 threads, 4 x 4 pairs each
__global__ void __launch_bounds__(256) k_cc_tile(const uint32_t *Zt, int n_anchor, int C, int words, float *S, int accumulate) {
	if (blockIdx.x < blockIdx.y) return;
	__shared__ __align__(16) uint32_t As[32][64], Bs[32][64];
	const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
	const int a0 = blockIdx.y * 64, b0 = blockIdx.x * 64;
	unsigned cnt[4][4];
#pragma unroll
	for (int i = 0; i < 4; ++i)
#pragma unroll
		for (int j = 0; j < 4; ++j) cnt[i][j] = 0u;
	const int tail = C & 3; // chains in the last word (0: all four)
	for (int w0 = 0; w0 < words; w0 += 32) {
		__syncthreads();
		for (int e = threadIdx.x; e < 32 * 64; e += 256) {
			const int w = w0 + (e >> 6), r = e & 63;
			uint32_t va = 0x00010203u, vb = 0x04050607u; // (beyond the matrix: never equal)
			if (w < words) {
				if (a0 + r < n_anchor) va = Zt[(size_t)w * n_anchor + a0 + r];
				if (b0 + r < n_anchor) vb = Zt[(size_t)w * n_anchor + b0 + r];
				if (tail && w == words - 1) { // the pad bytes of the last word must not count: make them differ
					const uint32_t keep = (1u << (8 * tail)) - 1u;
					va = (va & keep) | (0xa5a5a5a5u & ~keep);
					vb = (vb & keep) | (0x5a5a5a5au & ~keep);
				}
			}
			As[e >> 6][r] = va;
			Bs[e >> 6][r] = vb;
		}
		__syncthreads();
#pragma unroll 4
		for (int w = 0; w < 32; ++w) {
			const uint4 av = *reinterpret_cast<const uint4 *>(&As[w][ty * 4]);
			const uint4 bv = *reinterpret_cast<const uint4 *>(&Bs[w][tx * 4]);
			const uint32_t a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
			for (int i = 0; i < 4; ++i)
#pragma unroll
				for (int j = 0; j < 4; ++j) cnt[i][j] += __popc(__vcmpeq4(a[i], b[j])); // 8 bits per equal byte
		}
	}
#pragma unroll
	for (int i = 0; i < 4; ++i)
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int a = a0 + ty * 4 + i, b = b0 + tx * 4 + j;
			if (a < n_anchor && b < n_anchor) {
				const float v = (float)(cnt[i][j] >> 3);
				float *o = S + (size_t)a * n_anchor + b;
				*o = (accumulate ? *o : 0.0f) + v;
				if (blockIdx.x != blockIdx.y) { // the mirrored tile
					float *m = S + (size_t)b * n_anchor + a;
					*m = (accumulate ? *m : 0.0f) + v;
				}
			}
		}
}

// fallback for Kmax > 256 (slot ids do not fit a byte): one thread per pair
__global__ void k_cc_naive(const npb_z_t *z, const int64_t *anchors, int n_anchor, int C, float *S, int accumulate) {
	const int a = blockIdx.y, b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_anchor) return;
	const npb_z_t *za = z + (size_t)anchors[a] * C, *zb = z + (size_t)anchors[b] * C;
	int cnt = 0;
	for (int c = 0; c < C; ++c) cnt += za[c] == zb[c];
	float *o = S + (size_t)a * n_anchor + b;
	*o = (accumulate ? *o : 0.0f) + (float)cnt;
}

npb_status npb_launch_cocluster(npb_chains *ch, const int64_t *d_anchors, int n_anchor, float *S_dev, int accumulate) {
	npb_ctx *ctx = ch->ctx;
	const int C = (int)ch->C;
	if (ch->Kmax > 256) {
		dim3 grid((n_anchor + 127) / 128, n_anchor);
		k_cc_naive<<<grid, 128, 0, ctx->stream>>>(ch->z, d_anchors, n_anchor, C, S_dev, accumulate);
		NPB_CUDA_OK(cudaGetLastError());
		return NPB_OK;
	}
	const int words = (C + 3) / 4;
	uint32_t *Zt = nullptr;
	NPB_CUDA_OK(cudaMallocAsync((void **)&Zt, (size_t)words * n_anchor * sizeof(uint32_t), ctx->stream));
	dim3 gg((n_anchor + 255) / 256, words);
	k_cc_gather<<<gg, 256, 0, ctx->stream>>>(ch->z, d_anchors, n_anchor, C, words, Zt);
	NPB_CUDA_OK(cudaGetLastError());
	const int tiles = (n_anchor + 63) / 64;
	dim3 gt(tiles, tiles);
	k_cc_tile<<<gt, 256, 0, ctx->stream>>>(Zt, n_anchor, C, words, S_dev, accumulate);
	NPB_CUDA_OK(cudaGetLastError());
	NPB_CUDA_OK(cudaFreeAsync(Zt, ctx->stream));
	return NPB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen): the library has no link-time dependency on it, and inside a process that already
// loaded a libnccl.so.2 (PyTorch's) the same one is used.
// ---------------------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };
struct NcclApi {
	void *h = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
bool nccl_load() {
	if (g_nccl.h) return true;
	// the copy the process already has (PyTorch's), else the system's
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
	if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
	if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
	if (!h) return false;
	NcclApi a;
	a.h = h;
	a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
	a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
	a.CommInitAll = (decltype(a.CommInitAll))dlsym(h, "ncclCommInitAll");
	a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
	a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
	a.GroupStart = (decltype(a.GroupStart))dlsym(h, "ncclGroupStart");
	a.GroupEnd = (decltype(a.GroupEnd))dlsym(h, "ncclGroupEnd");
	a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
	if (!a.GetUniqueId || !a.CommInitRank || !a.CommInitAll || !a.CommDestroy || !a.AllReduce || !a.GroupStart || !a.GroupEnd) return false;
	g_nccl = a;
	return true;
}
}

struct npb_comm {
	ncclComm_t comm = nullptr;
	npb_ctx *ctx = nullptr;
	int rank = 0, world = 1;
};

static npb_status nccl_fail(npb_ctx *ctx, int r, const char *what) {
	char msg[256];
	snprintf(msg, sizeof(msg), "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
	return npb_fail(ctx, NPB_E_NCCL, msg);
}

npb_status npb_comm_unique_id(char out[128]) {
	if (!out) return NPB_E_BAD_ARG;
	if (!nccl_load()) return NPB_E_NCCL;
	ncclUniqueId id;
	if (g_nccl.GetUniqueId(&id) != 0) return NPB_E_NCCL;
	memcpy(out, id.internal, 128);
	return NPB_OK;
}
npb_status npb_comm_create(npb_ctx *ctx, const char id[128], int rank, int world, npb_comm **out) {
	if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) return NPB_E_BAD_ARG;
	if (!nccl_load()) return npb_fail(ctx, NPB_E_NCCL, "libnccl.so.2 not found");
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	ncclUniqueId uid;
	memcpy(uid.internal, id, 128);
	npb_comm *c = new npb_comm();
	c->ctx = ctx;
	c->rank = rank;
	c->world = world;
	const int r = g_nccl.CommInitRank(&c->comm, world, uid, rank);
	if (r != 0) { delete c; return nccl_fail(ctx, r, "ncclCommInitRank"); }
	*out = c;
	return NPB_OK;
}
// one communicator per context, all in this process (a single host process driving several devices: the CLI's --gpus)
npb_status npb_comm_create_all(npb_ctx *const *ctxs, int n, npb_comm **out) {
	if (!ctxs || !out || n < 1) return NPB_E_BAD_ARG;
	if (!nccl_load()) return npb_fail(ctxs[0], NPB_E_NCCL, "libnccl.so.2 not found");
	std::vector<int> devs(n);
	std::vector<ncclComm_t> comms(n);
	for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
	const int r = g_nccl.CommInitAll(comms.data(), n, devs.data());
	if (r != 0) return nccl_fail(ctxs[0], r, "ncclCommInitAll");
	for (int i = 0; i < n; ++i) {
		out[i] = new npb_comm();
		out[i]->comm = comms[i];
		out[i]->ctx = ctxs[i];
		out[i]->rank = i;
		out[i]->world = n;
	}
	return NPB_OK;
}
npb_status npb_comm_destroy(npb_comm *c) {
	if (!c) return NPB_OK;
	if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
	delete c;
	return NPB_OK;
}
// sum over the ranks, in place, of a device buffer on the communicator's context (dtype: 32 or 64 bit float)
npb_status npb_comm_allreduce_sum(npb_comm *c, void *dev_buf, int64_t count, int bits) {
	if (!c || !dev_buf || count <= 0 || (bits != 32 && bits != 64)) return NPB_E_BAD_ARG;
	npb_ctx *ctx = c->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const int r = g_nccl.AllReduce(dev_buf, dev_buf, (size_t)count, bits == 32 ? NCCL_FLOAT32 : NCCL_FLOAT64, NCCL_SUM, c->comm, ctx->stream);
	if (r != 0) return nccl_fail(ctx, r, "ncclAllReduce");
	return NPB_OK;
}
npb_status npb_comm_group(int start) {
	if (!nccl_load()) return NPB_E_NCCL;
	return (start ? g_nccl.GroupStart() : g_nccl.GroupEnd()) == 0 ? NPB_OK : NPB_E_NCCL;
}
