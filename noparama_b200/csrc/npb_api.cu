// npb_api.cu -- the C ABI of include/npb200.h: handle management, host<->device staging, launch order.
#include "npb_internal.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

npb_status npb_launch_logdensity(npb_ctx *, npb_dataset *, const int64_t *, int64_t, int, const double *, const double *,
		const double *, const float *, const float *, const float *, int, double *);
npb_status npb_launch_logdensity_sum(npb_ctx *, npb_dataset *, const int64_t *, int64_t, int, const double *, const double *,
		const double *, double *);
npb_status npb_launch_metrics(npb_chains *, const int32_t *, int, double *, double *, double *, double *, int32_t *);
npb_status npb_launch_cocluster(npb_chains *, const int64_t *, int, float *, int);
npb_status npb_launch_fma_peak(npb_ctx *ctx, double *tflops);
npb_status npb_launch_replay(npb_ctx *ctx, const double *X, int N, int D, int M, double alpha, int nslots, double *theta,
		int *counts, int32_t *z, int64_t n_steps, const int32_t *item, const int64_t *order_off, const int32_t *order,
		const double *aux_theta, const double *u, const int32_t *new_slot, int32_t *picked, int64_t z_every,
		int32_t *z_after, int *status);

npb_status npb_fail_cuda(npb_ctx *ctx, cudaError_t e, const char *expr, const char *file, int line) {
	if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s: %s (%s:%d)", cudaGetErrorString(e), expr, file, line);
	return e == cudaErrorMemoryAllocation ? NPB_E_NOMEM : NPB_E_CUDA;
}
npb_status npb_fail(npb_ctx *ctx, npb_status s, const char *msg) {
	if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s", msg);
	return s;
}

// small RAII device buffer for call-scoped staging
template <typename T>
struct DevBuf {
	T *p = nullptr;
	cudaError_t alloc(size_t n) { return cudaMalloc((void **)&p, (n ? n : 1) * sizeof(T)); }
	~DevBuf() { if (p) cudaFree(p); }
};

extern "C" {

const char *npb_status_str(npb_status s) {
	switch (s) {
	case NPB_OK: return "ok";
	case NPB_E_BAD_ARG: return "bad argument";
	case NPB_E_CUDA: return "CUDA error";
	case NPB_E_KMAX_OVERFLOW: return "a chain needed more than Kmax clusters";
	case NPB_E_NOT_POSITIVE: return "covariance not invertible / precision not positive definite";
	case NPB_E_UNSUPPORTED: return "unsupported configuration";
	case NPB_E_REPLAY_MISMATCH: return "replay diverged from the recorded trace";
	case NPB_E_NOMEM: return "out of device memory";
	case NPB_E_NCCL: return "NCCL error";
	case NPB_E_ALREADY_ASSIGNED: return "already assigned";
	case NPB_E_ASSIGNMENT_REMAINING: return "assignment remaining";
	case NPB_E_ASSIGNMENT_ABSENT: return "assignment absent";
	default: return "unknown status";
	}
}

npb_status npb_ctx_create(int device, npb_ctx **out) {
	if (!out) return NPB_E_BAD_ARG;
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return NPB_E_CUDA; // no CPU fallback
	if (device < 0 || device >= ndev) return NPB_E_BAD_ARG;
	npb_ctx *ctx = new (std::nothrow) npb_ctx();
	if (!ctx) return NPB_E_NOMEM;
	ctx->device = device;
	if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
			cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
		delete ctx;
		return NPB_E_CUDA;
	}
	if (cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || ctx->n_sm <= 0) {
		npb_ctx_destroy(ctx);
		return NPB_E_CUDA;
	}
	*out = ctx;
	return NPB_OK;
}

npb_status npb_ctx_destroy(npb_ctx *ctx) {
	if (!ctx) return NPB_OK;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->d_CT2) cudaFree(ctx->d_CT2);
	if (ctx->d_S) cudaFree(ctx->d_S);
	cudaEventDestroy(ctx->ev0);
	cudaEventDestroy(ctx->ev1);
	cudaStreamDestroy(ctx->stream);
	delete ctx;
	return NPB_OK;
}

void *npb_ctx_stream(npb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
const char *npb_ctx_last_error(npb_ctx *ctx) { return ctx ? ctx->err : ""; }

npb_status npb_ctx_synchronize(npb_ctx *ctx) {
	if (!ctx) return NPB_E_BAD_ARG;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}

// ---- dataset ------------------------------------------------------------------------------------------------
__global__ void k_to_float(const double *in, float *out, int64_t n) {
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = (float)in[i];
}

static npb_status dataset_push(npb_dataset *ds, const double *X) {
	npb_ctx *ctx = ds->ctx;
	const int64_t n = ds->N * ds->D;
	memcpy(ds->h_stage, X, sizeof(double) * n);
	NPB_CUDA_OK(cudaMemcpyAsync(ds->X64, ds->h_stage, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
	k_to_float<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ds->X64, ds->X32, n);
	NPB_CUDA_OK(cudaGetLastError());
	ds->whitened_epoch = 0;
	ds->xbar_valid = false; // recomputed on demand into the same buffer (npb_alg8_gemm.cu)
	return NPB_OK;
}

npb_status npb_dataset_upload(npb_ctx *ctx, const double *X, int64_t N, int D, npb_dataset **out) {
	NpbRange nvtx_range("npb:dataset_upload");
	if (!ctx || !X || !out || N <= 0 || D <= 0 || D > NPB_MAX_D || N > 0x7fffffff) return NPB_E_BAD_ARG;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	npb_dataset *ds = new (std::nothrow) npb_dataset();
	if (!ds) return NPB_E_NOMEM;
	ds->ctx = ctx;
	ds->N = N;
	ds->D = D;
	const size_t n = (size_t)N * D;
	cudaError_t e;
	if ((e = cudaMalloc((void **)&ds->X64, n * sizeof(double))) != cudaSuccess ||
			(e = cudaMalloc((void **)&ds->X32, n * sizeof(float))) != cudaSuccess ||
			(e = cudaMalloc((void **)&ds->Xw, n * sizeof(float))) != cudaSuccess ||
			(e = cudaMalloc((void **)&ds->Xwn, (size_t)N * sizeof(float))) != cudaSuccess ||
			(e = cudaMallocHost((void **)&ds->h_stage, n * sizeof(double))) != cudaSuccess) {
		npb_dataset_destroy(ds);
		return npb_fail_cuda(ctx, e, "dataset allocation", __FILE__, __LINE__);
	}
	npb_status s = dataset_push(ds, X);
	if (s != NPB_OK) { npb_dataset_destroy(ds); return s; }
	if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
		npb_dataset_destroy(ds);
		return npb_fail_cuda(ctx, e, "dataset upload", __FILE__, __LINE__);
	}
	*out = ds;
	return NPB_OK;
}

npb_status npb_dataset_update(npb_dataset *ds, const double *X) {
	NpbRange nvtx_range("npb:dataset_update");
	if (!ds || !X) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ds->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream)); // the staging buffer may still be in flight
	return dataset_push(ds, X);
}

npb_status npb_dataset_destroy(npb_dataset *ds) {
	if (!ds) return NPB_OK;
	if (ds->n_chains_alive > 0) return npb_fail(ds->ctx, NPB_E_BAD_ARG, "chain handles still reference this dataset: destroy them first");
	cudaSetDevice(ds->ctx->device);
	cudaStreamSynchronize(ds->ctx->stream);
	if (ds->X64) cudaFree(ds->X64);
	if (ds->X32) cudaFree(ds->X32);
	if (ds->Xw) cudaFree(ds->Xw);
	if (ds->Xwn) cudaFree(ds->Xwn);
	if (ds->h_stage) cudaFreeHost(ds->h_stage);
	if (ds->Xbar) cudaFree(ds->Xbar);
	delete ds;
	return NPB_OK;
}

// ---- prior --------------------------------------------------------------------------------------------------
npb_status npb_prior_set_niw(npb_ctx *ctx, int D, const double *mu0, double kappa, double nu, const double *Lambda,
		double alpha, int flags) {
	if (!ctx || !mu0 || !Lambda || D <= 0 || D > NPB_MAX_D || !(kappa > 0) || !(alpha > 0)) return NPB_E_BAD_ARG;
	if (!(flags & NPB_BUGCOMPAT_DEGENERATE_IW))
		return npb_fail(ctx, NPB_E_UNSUPPORTED, "only the reference's degenerate inverse-Wishart draw (Q2) is implemented");
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	PriorHost p;
	p.D = D;
	p.flags = flags;
	p.kappa = kappa;
	p.nu = nu;
	p.alpha = alpha;
	p.mu0.assign(mu0, mu0 + D);
	p.Lambda.assign(Lambda, Lambda + (size_t)D * D);
	if (!npb_prepare_prior(p)) return npb_fail(ctx, NPB_E_NOT_POSITIVE, "Lambda is not positive definite");
	p.set = true;
	const int TRI = npb_tri(D);
	std::vector<float> ct2(TRI), s(TRI);
	for (int t = 0; t < TRI; ++t) {
		ct2[t] = (float)(p.CT[t] * NPB_HALF_LOG2E_SQRT);
		s[t] = (float)p.S[t];
	}
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	if (ctx->d_CT2) cudaFree(ctx->d_CT2);
	if (ctx->d_S) cudaFree(ctx->d_S);
	ctx->d_CT2 = ctx->d_S = nullptr;
	NPB_CUDA_OK(cudaMalloc((void **)&ctx->d_CT2, TRI * sizeof(float)));
	NPB_CUDA_OK(cudaMalloc((void **)&ctx->d_S, TRI * sizeof(float)));
	NPB_CUDA_OK(cudaMemcpy(ctx->d_CT2, ct2.data(), TRI * sizeof(float), cudaMemcpyHostToDevice));
	NPB_CUDA_OK(cudaMemcpy(ctx->d_S, s.data(), TRI * sizeof(float), cudaMemcpyHostToDevice));
	ctx->prior = p;
	ctx->prior_epoch++;
	return NPB_OK;
}

// `-c regression` / `-c angular` (np_main.cpp:322-328, :357-364): scalar-noise likelihood, normal-inverse-gamma base measure
npb_status npb_prior_set_nig(npb_ctx *ctx, int family, const double *mu0, const double *Lambda, double ig_alpha, double ig_beta, double alpha) {
	if (!ctx || !mu0 || !Lambda || !(ig_alpha > 0) || !(ig_beta > 0) || !(alpha > 0)) return NPB_E_BAD_ARG;
	if (family != NPB_FAMILY_REGRESSION && family != NPB_FAMILY_ANGULAR) return npb_fail(ctx, NPB_E_BAD_ARG, "family: NPB_FAMILY_REGRESSION | NPB_FAMILY_ANGULAR");
	const double det = Lambda[0] * Lambda[3] - Lambda[1] * Lambda[2];
	if (!(Lambda[0] > 0) || !(det > 0) || Lambda[1] != Lambda[2]) return npb_fail(ctx, NPB_E_NOT_POSITIVE, "Lambda is not symmetric positive definite");
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	PriorHost p;
	p.family = family;
	p.D = family == NPB_FAMILY_REGRESSION ? 3 : 2; // width of a data row: (1, a, b) | (a, b) (np_main.cpp:83-101)
	p.kappa = p.nu = 1.0;
	p.alpha = alpha;
	p.ig_alpha = ig_alpha;
	p.ig_beta = ig_beta;
	p.mu0.assign(mu0, mu0 + 2);
	p.Lambda.assign(Lambda, Lambda + 4);
	p.set = true;
	ctx->prior = p;
	ctx->prior_epoch++;
	return NPB_OK;
}

// [n_rows, K] log-densities (natural log) of dataset rows under K scalar-noise parameter sets (mu [K, 2], sigma [K]): the raw
// parameters are brought into the slot layout on the host, the evaluation is k_logdensity's (scalarnoise_multivariatenormal.cpp:182-250)
npb_status npb_scalarnoise_logdensity_batch(npb_ctx *ctx, npb_dataset *ds, int family, const int64_t *rows, int64_t n_rows, const double *mu,
		const double *sigma, int K, double *out) {
	if (!ctx || !ds || !mu || !sigma || !out || K <= 0 || n_rows <= 0) return NPB_E_BAD_ARG;
	if (family != NPB_FAMILY_REGRESSION && family != NPB_FAMILY_ANGULAR) return NPB_E_BAD_ARG;
	const int D = ds->D, TRI = npb_tri(D);
	if (D != (family == NPB_FAMILY_REGRESSION ? 3 : 2)) return npb_fail(ctx, NPB_E_BAD_ARG, "rows are (1, a, b) for regression and (a, b) for angular");
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	std::vector<double> h((size_t)K * (D + TRI + 1));
	double *hm = h.data(), *hT = hm + (size_t)K * D, *hc = hT + (size_t)K * TRI;
	for (int k = 0; k < K; ++k) {
		if (!(sigma[k] > 0)) return npb_fail(ctx, NPB_E_NOT_POSITIVE, "sigma must be positive");
		npb_sn_slot_from_raw(family, D, mu + 2 * (size_t)k, sigma[k], hm + (size_t)k * D, hT + (size_t)k * TRI, hc + k);
	}
	double *d = nullptr, *d_out = nullptr;
	int64_t *d_rows = nullptr;
	cudaError_t e = cudaMalloc((void **)&d, h.size() * sizeof(double));
	if (e == cudaSuccess) e = cudaMalloc((void **)&d_out, (size_t)n_rows * K * sizeof(double));
	if (e == cudaSuccess && rows) e = cudaMalloc((void **)&d_rows, (size_t)n_rows * sizeof(int64_t));
	if (e == cudaSuccess) e = cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
	if (e == cudaSuccess && rows) e = cudaMemcpyAsync(d_rows, rows, (size_t)n_rows * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
	npb_status s = NPB_OK;
	if (e == cudaSuccess) s = npb_launch_logdensity(ctx, ds, d_rows, n_rows, K, d, d + (size_t)K * D, d + (size_t)K * (D + TRI), nullptr, nullptr, nullptr, 64, d_out);
	if (e == cudaSuccess && s == NPB_OK) e = cudaMemcpyAsync(out, d_out, (size_t)n_rows * K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	cudaFree(d);
	cudaFree(d_out);
	cudaFree(d_rows);
	if (e != cudaSuccess) return npb_fail_cuda(ctx, e, "npb_scalarnoise_logdensity_batch", __FILE__, __LINE__);
	return s;
}

// count raw draws (mu_0, mu_1, sigma) from the normal-inverse-gamma base measure as the sweep kernel makes them (parity probe)
npb_status npb_chains_sample_base_nig(npb_chains *ch, int64_t chain, int count, float *out) {
	if (!ch || !out || count <= 0 || chain < 0 || chain >= ch->C) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	if (!ctx->prior.family) return npb_fail(ctx, NPB_E_BAD_ARG, "the context's prior is not a normal-inverse-gamma one");
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	float *d = nullptr;
	NPB_CUDA_OK(cudaMalloc((void **)&d, (size_t)count * 3 * sizeof(float)));
	npb_status s = npb_launch_sn_sample_base(ch, (int)chain, count, d);
	cudaError_t e = cudaSuccess;
	if (s == NPB_OK) e = cudaMemcpyAsync(out, d, (size_t)count * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	cudaFree(d);
	if (e != cudaSuccess) return npb_fail_cuda(ctx, e, "npb_chains_sample_base_nig", __FILE__, __LINE__);
	return s;
}

// ---- density ------------------------------------------------------------------------------------------------
struct ThetaStaging {
	std::vector<double> T, c;
};
static npb_status stage_thetas(npb_ctx *ctx, int D, int K, const double *mu, const double *Sigma, ThetaStaging &st) {
	const int TRI = npb_tri(D);
	st.T.resize((size_t)K * TRI);
	st.c.resize(K);
	for (int k = 0; k < K; ++k) {
		double logdet;
		if (!npb_prepare_theta(D, mu + (size_t)k * D, Sigma + (size_t)k * D * D, st.T.data() + (size_t)k * TRI, &logdet))
			return npb_fail(ctx, NPB_E_NOT_POSITIVE, "Sigma is not invertible with a positive definite symmetric precision");
		// -log sqrt((2 pi)^D det Sigma), multivariatenormal.cpp:131-133
		st.c[k] = -0.5 * (D * std::log(2.0 * M_PI) + logdet);
	}
	return NPB_OK;
}

static npb_status logdensity_common(npb_ctx *ctx, npb_dataset *ds, const int64_t *rows, int64_t n_rows, const double *mu,
		const double *Sigma, int K, int precision, bool sum, double *out) {
	if (!ctx || !ds || !mu || !Sigma || !out || K <= 0 || n_rows < 0 || (precision != 32 && precision != 64)) return NPB_E_BAD_ARG;
	if (ds->ctx != ctx) return NPB_E_BAD_ARG;
	if (n_rows == 0) {
		if (sum) for (int k = 0; k < K; ++k) out[k] = 0.0;
		return NPB_OK;
	}
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const int D = ds->D, TRI = npb_tri(D);
	if (rows)
		for (int64_t r = 0; r < n_rows; ++r)
			if (rows[r] < 0 || rows[r] >= ds->N) return NPB_E_BAD_ARG;
	ThetaStaging st;
	npb_status s = stage_thetas(ctx, D, K, mu, Sigma, st);
	if (s != NPB_OK) return s;
	DevBuf<double> d_mu, d_T, d_c, d_out;
	DevBuf<float> f_mu, f_T, f_c;
	DevBuf<int64_t> d_rows;
	const size_t n_out = sum ? (size_t)K : (size_t)n_rows * K;
	NPB_CUDA_OK(d_mu.alloc((size_t)K * D));
	NPB_CUDA_OK(d_T.alloc((size_t)K * TRI));
	NPB_CUDA_OK(d_c.alloc(K));
	NPB_CUDA_OK(d_out.alloc(n_out));
	NPB_CUDA_OK(cudaMemcpyAsync(d_mu.p, mu, sizeof(double) * K * D, cudaMemcpyHostToDevice, ctx->stream));
	NPB_CUDA_OK(cudaMemcpyAsync(d_T.p, st.T.data(), sizeof(double) * K * TRI, cudaMemcpyHostToDevice, ctx->stream));
	NPB_CUDA_OK(cudaMemcpyAsync(d_c.p, st.c.data(), sizeof(double) * K, cudaMemcpyHostToDevice, ctx->stream));
	std::vector<float> hmu, hT, hc;
	if (precision == 32) {
		hmu.resize((size_t)K * D); hT.resize((size_t)K * TRI); hc.resize(K);
		for (size_t i = 0; i < hmu.size(); ++i) hmu[i] = (float)mu[i];
		for (size_t i = 0; i < hT.size(); ++i) hT[i] = (float)st.T[i];
		for (int k = 0; k < K; ++k) hc[k] = (float)st.c[k];
		NPB_CUDA_OK(f_mu.alloc(hmu.size()));
		NPB_CUDA_OK(f_T.alloc(hT.size()));
		NPB_CUDA_OK(f_c.alloc(K));
		NPB_CUDA_OK(cudaMemcpyAsync(f_mu.p, hmu.data(), sizeof(float) * hmu.size(), cudaMemcpyHostToDevice, ctx->stream));
		NPB_CUDA_OK(cudaMemcpyAsync(f_T.p, hT.data(), sizeof(float) * hT.size(), cudaMemcpyHostToDevice, ctx->stream));
		NPB_CUDA_OK(cudaMemcpyAsync(f_c.p, hc.data(), sizeof(float) * K, cudaMemcpyHostToDevice, ctx->stream));
	}
	if (rows) {
		NPB_CUDA_OK(d_rows.alloc(n_rows));
		NPB_CUDA_OK(cudaMemcpyAsync(d_rows.p, rows, sizeof(int64_t) * n_rows, cudaMemcpyHostToDevice, ctx->stream));
	}
	if (sum) s = npb_launch_logdensity_sum(ctx, ds, rows ? d_rows.p : nullptr, n_rows, K, d_mu.p, d_T.p, d_c.p, d_out.p);
	else s = npb_launch_logdensity(ctx, ds, rows ? d_rows.p : nullptr, n_rows, K, d_mu.p, d_T.p, d_c.p, f_mu.p, f_T.p, f_c.p, precision, d_out.p);
	if (s != NPB_OK) return s;
	NPB_CUDA_OK(cudaMemcpyAsync(out, d_out.p, sizeof(double) * n_out, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}

npb_status npb_logdensity_batch(npb_ctx *ctx, npb_dataset *ds, const int64_t *rows, int64_t n_rows, const double *mu,
		const double *Sigma, int K, int precision, double *out) {
	return logdensity_common(ctx, ds, rows, n_rows, mu, Sigma, K, precision, false, out);
}
npb_status npb_logdensity_sum(npb_ctx *ctx, npb_dataset *ds, const int64_t *rows, int64_t n_rows, const double *mu,
		const double *Sigma, int K, double *out) {
	return logdensity_common(ctx, ds, rows, n_rows, mu, Sigma, K, 64, true, out);
}

// ---- chains -------------------------------------------------------------------------------------------------
__global__ void k_best_init(double *best, int C);

static npb_status ensure_whitened(npb_dataset *ds) {
	// (npb_prior_set_niw may have been called with another dimension since the dataset was uploaded)
	if (!ds->ctx->prior.set || ds->ctx->prior.D != ds->D) return npb_fail(ds->ctx, NPB_E_BAD_ARG, "the context's prior is not of the dataset's dimension");
	if (ds->ctx->prior.family) return NPB_OK; // the scalar-noise kernels read the rows as they are
	if (ds->whitened_epoch == ds->ctx->prior_epoch) return NPB_OK;
	return npb_launch_whiten(ds);
}

static npb_status set_switch(npb_chains *ch, const char *name, const char *value);
npb_status npb_chains_create(npb_ctx *ctx, npb_dataset *ds, int64_t n_chains, int Kmax, int m_aux, int K0, uint64_t seed,
		npb_chains **out) {
	if (!ctx || !ds || !out || ds->ctx != ctx || n_chains <= 0 || n_chains > 0x7fffffff || Kmax <= 0 || Kmax > 65535 ||
			m_aux <= 0 || m_aux > NPB_MAX_AUX || K0 <= 0 || K0 > Kmax)
		return NPB_E_BAD_ARG;
	if (!ctx->prior.set || ctx->prior.D != ds->D) return npb_fail(ctx, NPB_E_BAD_ARG, "set a prior of the dataset's dimension first");
	if (Kmax % 32) return npb_fail(ctx, NPB_E_BAD_ARG, "Kmax must be a multiple of 32");
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	npb_chains *ch = new (std::nothrow) npb_chains();
	if (!ch) return NPB_E_NOMEM;
	ch->ctx = ctx;
	ch->ds = ds;
	ch->C = n_chains;
	ch->Kmax = Kmax;
	ch->m_aux = m_aux;
	ch->K0 = K0;
	ch->D = ds->D;
	ch->seed = seed;
	{ // behaviour switches are read once, here (npb_chains_set_option changes them afterwards)
		const char *e = getenv("NPB_D16_PATH");
		strncpy(ch->opt_d16_path, e && e[0] ? e : "auto", sizeof(ch->opt_d16_path) - 1);
		static const char *const names[] = {"NPB_D64_SPEC", "NPB_F16_FLAGS", "NPB_D16_BLOCK", "NPB_D64_BLOCK", "NPB_D16_EPI", "NPB_D16_NH",
				"NPB_D16_AUX", "NPB_D64_OVERLAP", "NPB_D64_DENSITY", "NPB_TILE_KERNEL", "NPB_A2_TILE", "NPB_A2_TC", "NPB_A2_TC16"};
		static const char *const opts[] = {"spec", "f16_flags", "d16_block", "d64_block", "d16_epi", "d16_nh", "d16_aux", "d64_overlap",
				"d64_density", "tile_kernel", "a2_tile", "a2_tc", "a2_tc16"};
		for (int i = 0; i < 13; ++i)
			if ((e = getenv(names[i])) && e[0] && set_switch(ch, opts[i], e) != NPB_OK) {
				npb_chains_destroy(ch);
				return npb_fail(ctx, NPB_E_BAD_ARG, "bad value in an NPB_* environment switch");
			}
	}
	const size_t PS = npb_ps(ds->D);
	cudaError_t e;
	if ((e = cudaMalloc((void **)&ch->z, (size_t)ds->N * n_chains * sizeof(npb_z_t))) != cudaSuccess ||
			(e = cudaMalloc((void **)&ch->theta, (size_t)n_chains * Kmax * PS * sizeof(float))) != cudaSuccess ||
			(e = cudaMalloc((void **)&ch->counts, (size_t)n_chains * Kmax * sizeof(int))) != cudaSuccess ||
			(e = cudaMalloc((void **)&ch->st, (size_t)n_chains * 4 * sizeof(unsigned long long))) != cudaSuccess ||
			(e = cudaMalloc((void **)&ch->kocc, (size_t)n_chains * sizeof(int))) != cudaSuccess ||
			(e = cudaMalloc((void **)&ch->overflow, (size_t)n_chains * sizeof(int))) != cudaSuccess ||
			(e = cudaMalloc((void **)&ch->smst, (size_t)n_chains * 12 * sizeof(unsigned long long))) != cudaSuccess ||
			(e = cudaMemsetAsync(ch->smst, 0, (size_t)n_chains * 12 * sizeof(unsigned long long), ctx->stream)) != cudaSuccess) {
		npb_chains_destroy(ch);
		return npb_fail_cuda(ctx, e, "chain state allocation", __FILE__, __LINE__);
	}
	npb_status s = ensure_whitened(ds);
	if (s == NPB_OK) s = npb_launch_chains_init(ch, ch->K0, nullptr);
	if (s != NPB_OK) { npb_chains_destroy(ch); return s; }
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	ds->n_chains_alive++;
	ch->counted = true;
	*out = ch;
	return NPB_OK;
}

// the measurement / behaviour switches of npb_chains::Switches by option name; NPB_E_BAD_ARG for an unknown name or value
static npb_status set_switch(npb_chains *ch, const char *name, const char *value) {
	npb_chains::Switches &w = ch->sw;
	const int v = atoi(value);
	if (!strcmp(name, "spec")) { if (v < 0 || v > 2) return NPB_E_BAD_ARG; w.spec = v; }
	else if (!strcmp(name, "f16_flags")) w.f16_flags = v;
	else if (!strcmp(name, "d16_block")) { if (v < 1) return NPB_E_BAD_ARG; w.d16_block = v; }
	else if (!strcmp(name, "d64_block")) { if (v < 1) return NPB_E_BAD_ARG; w.d64_block = v; }
	else if (!strcmp(name, "d16_epi")) { if (v != 8 && v != 16) return NPB_E_BAD_ARG; w.d16_epi = v; }
	else if (!strcmp(name, "d16_nh")) { if (v != 1 && v != 2 && v != 4) return NPB_E_BAD_ARG; w.d16_nh = v; }
	else if (!strcmp(name, "d16_aux")) { w.d16_aux_pre = value[0] == 'p'; w.d16_aux_grp = value[0] != 'l' && value[0] != 'p'; w.d16_aux_auto = false; } // pre | lazy | bound
	else if (!strcmp(name, "d64_overlap")) w.d64_overlap = value[0] != '0';
	else if (!strcmp(name, "d64_density")) w.d64_fp32 = value[0] == 'f';
	else if (!strcmp(name, "tile_kernel")) w.two_warp = value[0] == '2';
	else if (!strcmp(name, "a2_tile")) { if (v < 0 || v > 128) return NPB_E_BAD_ARG; w.a2_tile = v; }
	else if (!strcmp(name, "a2_tc")) w.a2_tc = value[0] != '0';
	else if (!strcmp(name, "a2_tc16")) w.a2_tc16 = value[0] != '0';
	else return NPB_E_BAD_ARG;
	return NPB_OK;
}

npb_status npb_chains_set_option(npb_chains *ch, const char *name, const char *value) {
	if (!ch || !name || !value) return NPB_E_BAD_ARG;
	if (set_switch(ch, name, value) == NPB_OK) return NPB_OK;
	if (!strcmp(name, "d16_path")) {
		if (strcmp(value, "auto") && strcmp(value, "tc") && strcmp(value, "tc2") && strcmp(value, "fp32"))
			return npb_fail(ch->ctx, NPB_E_BAD_ARG, "d16_path: auto | tc | tc2 | fp32");
		strncpy(ch->opt_d16_path, value, sizeof(ch->opt_d16_path) - 1);
		return NPB_OK;
	}
	if (!strcmp(name, "time_kernels")) {
		ch->time_kernels = value[0] == '1';
		return NPB_OK;
	}
	return npb_fail(ch->ctx, NPB_E_BAD_ARG, "unknown option");
}

// accumulated duration and number of launches of the dominant sweep kernel since the last call (option "time_kernels")
npb_status npb_chains_kernel_time(npb_chains *ch, double *ms, int64_t *launches) {
	if (!ch || !ms || !launches) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	for (size_t i = 0; i + 1 < ch->kt_ev.size(); i += 2) {
		float t = 0.0f;
		NPB_CUDA_OK(cudaEventElapsedTime(&t, ch->kt_ev[i], ch->kt_ev[i + 1]));
		ch->kt_ms += t;
		ch->kt_launches++;
		cudaEventDestroy(ch->kt_ev[i]);
		cudaEventDestroy(ch->kt_ev[i + 1]);
	}
	ch->kt_ev.clear();
	*ms = ch->kt_ms;
	*launches = ch->kt_launches;
	ch->kt_ms = 0.0;
	ch->kt_launches = 0;
	return NPB_OK;
}

npb_status npb_chains_destroy(npb_chains *ch) {
	if (!ch) return NPB_OK;
	cudaSetDevice(ch->ctx->device);
	cudaStreamSynchronize(ch->ctx->stream);
	if (ch->counted && ch->ds) ch->ds->n_chains_alive--;
	for (cudaEvent_t e : ch->kt_ev) cudaEventDestroy(e);
	for (void *q : {(void *)ch->a2_sx, (void *)ch->a2_sxx, (void *)ch->a2_work, (void *)ch->a2_prior, (void *)ch->a2_mu, (void *)ch->a2_P, (void *)ch->a2_ld,
			(void *)ch->a2_G, (void *)ch->a2_lp0})
		if (q) cudaFree(q);
	if (ch->z_prev) cudaFree(ch->z_prev);
	if (ch->dl_idx) cudaFree(ch->dl_idx);
	if (ch->dl_val) cudaFree(ch->dl_val);
	if (ch->dl_count) cudaFree(ch->dl_count);
	if (ch->h_dl_idx) cudaFreeHost(ch->h_dl_idx);
	if (ch->h_dl_val) cudaFreeHost(ch->h_dl_val);
	if (ch->z) cudaFree(ch->z);
	if (ch->theta) cudaFree(ch->theta);
	if (ch->counts) cudaFree(ch->counts);
	if (ch->st) cudaFree(ch->st);
	if (ch->kocc) cudaFree(ch->kocc);
	if (ch->overflow) cudaFree(ch->overflow);
	if (ch->h_z) cudaFreeHost(ch->h_z);
	if (ch->scan_order) cudaFree(ch->scan_order);
	if (ch->aux_keys) cudaFree(ch->aux_keys);
	if (ch->smst) cudaFree(ch->smst);
	if (ch->sm_zt) cudaFree(ch->sm_zt);
	if (ch->sm_pool) cudaFree(ch->sm_pool);
	if (ch->sm_dec) cudaFree(ch->sm_dec);
	if (ch->sm_order) cudaFree(ch->sm_order);
	if (ch->sm_detail) cudaFree(ch->sm_detail);
	if (ch->pstats) cudaFree(ch->pstats);
	if (ch->best_z) cudaFree(ch->best_z);
	if (ch->best_theta) cudaFree(ch->best_theta);
	if (ch->best_counts) cudaFree(ch->best_counts);
	if (ch->best_jll) cudaFree(ch->best_jll);
	if (ch->cur_jll) cudaFree(ch->cur_jll);
	if (ch->pLambda0) cudaFree(ch->pLambda0);
	if (ch->pfail) cudaFree(ch->pfail);
	if (ch->g_aimg) cudaFree(ch->g_aimg);
	if (ch->g_bimg) cudaFree(ch->g_bimg);
	if (ch->g_bconst) cudaFree(ch->g_bconst);
	if (ch->g_L) cudaFree(ch->g_L);
	if (ch->g_dirty) cudaFree(ch->g_dirty);
	if (ch->g_born) cudaFree(ch->g_born);
	if (ch->g_zblk) cudaFree(ch->g_zblk);
	if (ch->aux_max) cudaFree(ch->aux_max);
	delete ch;
	return NPB_OK;
}

npb_status npb_chains_init_from_params(npb_chains *ch, int K, const double *mu, const double *Sigma) {
	if (!ch || !mu || !Sigma || K <= 0 || K > ch->Kmax) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const int D = ch->D, TRI = npb_tri(D), PS = npb_ps(D);
	std::vector<float> th((size_t)K * PS);
	std::vector<double> T(TRI);
	for (int k = 0; k < K; ++k) {
		double logdet;
		if (!npb_prepare_theta(D, mu + (size_t)k * D, Sigma + (size_t)k * D * D, T.data(), &logdet))
			return npb_fail(ctx, NPB_E_NOT_POSITIVE, "Sigma is not invertible with a positive definite symmetric precision");
		float *o = th.data() + (size_t)k * PS;
		for (int d = 0; d < D; ++d) o[d] = (float)mu[(size_t)k * D + d];
		for (int t = 0; t < TRI; ++t) o[D + t] = (float)(T[t] * NPB_HALF_LOG2E_SQRT);
		o[D + TRI] = (float)(-0.5 * (D * std::log2(2.0 * M_PI) + logdet / std::log(2.0)));
	}
	DevBuf<float> d_th;
	NPB_CUDA_OK(d_th.alloc(th.size()));
	NPB_CUDA_OK(cudaMemcpyAsync(d_th.p, th.data(), sizeof(float) * th.size(), cudaMemcpyHostToDevice, ctx->stream));
	npb_status s = npb_launch_chains_init(ch, K, d_th.p);
	if (s != NPB_OK) return s;
	if (ch->best_jll) { // a new run: nothing kept yet (MCMC::run starts from -inf, np_mcmc.cpp:187-203)
		k_best_init<<<(unsigned)((ch->C + 255) / 256), 256, 0, ctx->stream>>>(ch->best_jll, (int)ch->C);
		NPB_CUDA_OK(cudaGetLastError());
	}
	ch->moved_frac_last = -1.0;
	ch->z_gen++;
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}

int64_t npb_chains_count(npb_chains *ch) { return ch ? ch->C : 0; }
int npb_chains_kmax(npb_chains *ch) { return ch ? ch->Kmax : 0; }

__global__ void k_scatter_chain_z(npb_z_t *z, const int32_t *zin, int N, int C, int chain) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < N) z[(size_t)i * C + chain] = (npb_z_t)zin[i];
}
__global__ void k_gather_chain_z(const npb_z_t *z, int32_t *zout, int N, int C, int chain0, int n) {
	int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (int64_t)N * n) return;
	int c = (int)(idx % n), i = (int)(idx / n); // consecutive threads read consecutive chains (coalesced)
	zout[(size_t)c * N + i] = (int32_t)z[(size_t)i * C + chain0 + c];
}

npb_status npb_chains_set_state(npb_chains *ch, int64_t chain, const int32_t *z, int K, const int32_t *slots,
		const double *mu, const double *Sigma) {
	if (!ch || !z || !slots || !mu || !Sigma || chain < 0 || chain >= ch->C || K <= 0 || K > ch->Kmax) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const int D = ch->D, TRI = npb_tri(D), PS = npb_ps(D), N = (int)ch->ds->N;
	std::vector<float> th((size_t)ch->Kmax * PS, 0.0f);
	std::vector<int> cnt(ch->Kmax, 0);
	std::vector<char> used(ch->Kmax, 0);
	std::vector<double> T(TRI);
	for (int k = 0; k < K; ++k) {
		const int s = slots[k];
		if (s < 0 || s >= ch->Kmax || used[s]) return NPB_E_BAD_ARG;
		used[s] = 1;
		double logdet;
		if (!npb_prepare_theta(D, mu + (size_t)k * D, Sigma + (size_t)k * D * D, T.data(), &logdet))
			return npb_fail(ctx, NPB_E_NOT_POSITIVE, "Sigma is not invertible with a positive definite symmetric precision");
		float *o = th.data() + (size_t)s * PS;
		for (int d = 0; d < D; ++d) o[d] = (float)mu[(size_t)k * D + d];
		for (int t = 0; t < TRI; ++t) o[D + t] = (float)(T[t] * NPB_HALF_LOG2E_SQRT);
		o[D + TRI] = (float)(-0.5 * (D * std::log2(2.0 * M_PI) + logdet / std::log(2.0)));
	}
	for (int i = 0; i < N; ++i) {
		if (z[i] < 0 || z[i] >= ch->Kmax || !used[z[i]]) return NPB_E_BAD_ARG;
		cnt[z[i]]++;
	}
	int occ = 0;
	for (int s = 0; s < ch->Kmax; ++s) occ += cnt[s] > 0;
	DevBuf<int32_t> d_z;
	NPB_CUDA_OK(d_z.alloc(N));
	NPB_CUDA_OK(cudaMemcpyAsync(d_z.p, z, sizeof(int32_t) * N, cudaMemcpyHostToDevice, ctx->stream));
	k_scatter_chain_z<<<(N + 255) / 256, 256, 0, ctx->stream>>>(ch->z, d_z.p, N, (int)ch->C, (int)chain);
	NPB_CUDA_OK(cudaGetLastError());
	NPB_CUDA_OK(cudaMemcpyAsync(ch->theta + (size_t)chain * ch->Kmax * PS, th.data(), sizeof(float) * th.size(), cudaMemcpyHostToDevice, ctx->stream));
	NPB_CUDA_OK(cudaMemcpyAsync(ch->counts + (size_t)chain * ch->Kmax, cnt.data(), sizeof(int) * cnt.size(), cudaMemcpyHostToDevice, ctx->stream));
	NPB_CUDA_OK(cudaMemcpyAsync(ch->kocc + chain, &occ, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	ch->z_gen++;
	return NPB_OK;
}

static npb_status collect_stats(npb_chains *ch, const std::vector<unsigned long long> &before,
		const std::vector<unsigned long long> &sm_before, int sampler, int n_sweeps, float ms, npb_sweep_stats *stats) {
	npb_ctx *ctx = ch->ctx;
	const size_t C = (size_t)ch->C;
	std::vector<unsigned long long> st(C * 4), sm(C * 12);
	NPB_CUDA_OK(cudaMemcpyAsync(sm.data(), ch->smst, sizeof(unsigned long long) * C * 12, cudaMemcpyDeviceToHost, ctx->stream));
	std::vector<int> kocc(C), ovf(C);
	NPB_CUDA_OK(cudaMemcpyAsync(st.data(), ch->st, sizeof(unsigned long long) * C * 4, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaMemcpyAsync(kocc.data(), ch->kocc, sizeof(int) * C, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaMemcpyAsync(ovf.data(), ch->overflow, sizeof(int) * C, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	memset(stats, 0, sizeof(*stats));
	stats->reassignments = sampler == NPB_ALG8 ? (int64_t)C * ch->ds->N * n_sweeps : 0;
	double sumK = 0;
	for (size_t c = 0; c < C; ++c) {
		for (int j = 0; j < 4; ++j) {
			stats->sm_attempts[j] += (int64_t)(sm[c * 12 + j] - sm_before[c * 12 + j]);
			stats->sm_accepts[j] += (int64_t)(sm[c * 12 + 4 + j] - sm_before[c * 12 + 4 + j]);
		}
		stats->sams_allocations += (int64_t)(sm[c * 12 + 8] - sm_before[c * 12 + 8]);
		if (sampler != NPB_ALG8) stats->reassignments += (int64_t)(sm[c * 12 + 9] - sm_before[c * 12 + 9]); // proposals made
		stats->candidates += (int64_t)(st[c * 4 + 0] - before[c * 4 + 0]);
		stats->moved += (int64_t)(st[c * 4 + 1] - before[c * 4 + 1]);
		stats->new_clusters += (int64_t)(st[c * 4 + 2] - before[c * 4 + 2]);
		sumK += kocc[c];
		if (kocc[c] > stats->max_K) stats->max_K = kocc[c];
		stats->overflow_chains += ovf[c] != 0;
	}
	stats->mean_K = sumK / (double)C;
	stats->kernel_ms = ms;
	if (sampler == NPB_ALG8 && stats->reassignments > 0) ch->moved_frac_last = (double)stats->moved / (double)stats->reassignments;
	return NPB_OK;
}

static npb_status sweep_common(npb_chains *ch, int sampler, int n_sweeps, int64_t n_proposals, npb_sweep_stats *stats,
		const double *X, uint16_t *z_out) {
	NpbRange nvtx_range(sampler == NPB_ALG8 ? "npb:sweep alg8" : sampler == NPB_ALG2 ? "npb:sweep alg2" : sampler == NPB_ALG2_CONJUGATE ? "npb:sweep alg2 conjugate" : sampler == NPB_JAIN_NEAL ? "npb:split-merge jain-neal" : "npb:split-merge triadic");
	if (!ch || n_sweeps < 0 || n_proposals < 0) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	if (sampler == NPB_ALG2) {
		// np_neal_algorithm2.cpp:32-120 (not compiled by the reference): K weights p(x|theta_k) n_k and ONE prior draw
		// weighted p(x|theta') alpha, a Bernoulli for "new" and then a pick among the existing -- one categorical draw over
		// K + 1 candidates, i.e. the Algorithm 8 kernels with a single auxiliary draw
		if (ch->m_aux != 1) return npb_fail(ctx, NPB_E_BAD_ARG, "Algorithm 2 runs on chains created with m_aux = 1");
		sampler = NPB_ALG8;
	}
	const bool conjugate = sampler == NPB_ALG2_CONJUGATE;
	if (conjugate) sampler = NPB_ALG8; // (statistics are the Algorithm 8 ones: candidates, moved, births)
	if (sampler != NPB_ALG8 && sampler != NPB_JAIN_NEAL && sampler != NPB_TRIADIC) return NPB_E_BAD_ARG;
	if (sampler == NPB_ALG8 && n_proposals) return NPB_E_BAD_ARG;
	std::vector<unsigned long long> before, sm_before;
	if (stats) {
		before.resize((size_t)ch->C * 4);
		sm_before.resize((size_t)ch->C * 12);
		NPB_CUDA_OK(cudaMemcpyAsync(before.data(), ch->st, sizeof(unsigned long long) * before.size(), cudaMemcpyDeviceToHost, ctx->stream));
		NPB_CUDA_OK(cudaMemcpyAsync(sm_before.data(), ch->smst, sizeof(unsigned long long) * sm_before.size(), cudaMemcpyDeviceToHost, ctx->stream));
		NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	}
	npb_status s;
	if (X) {
		s = npb_dataset_update(ch->ds, X);
		if (s != NPB_OK) return s;
	}
	s = ensure_whitened(ch->ds);
	if (s != NPB_OK) return s;
	if (ctx->prior.family && sampler != NPB_ALG8)
		return npb_fail(ctx, NPB_E_UNSUPPORTED, "scalar-noise families (regression / angular): Algorithm 8 only on the device");
	NPB_CUDA_OK(cudaEventRecord(ctx->ev0, ctx->stream));
	if (conjugate && n_sweeps > 0) {
		s = npb_launch_alg2_conjugate(ch, n_sweeps); // keeps its own statistics in step with the assignments
		if (s != NPB_OK) return s;
	} else if (sampler == NPB_ALG8 && n_sweeps > 0) {
		s = npb_launch_alg8_sweep(ch, n_sweeps);
		if (s != NPB_OK) return s;
		ch->z_gen++;
	} else if (sampler != NPB_ALG8 && (n_sweeps > 0 || n_proposals > 0)) {
		if (!ch->sm_detail) {
			NPB_CUDA_OK(cudaMalloc((void **)&ch->sm_detail, (size_t)ch->C * 16 * sizeof(float)));
			NPB_CUDA_OK(cudaMemsetAsync(ch->sm_detail, 0, (size_t)ch->C * 16 * sizeof(float), ctx->stream));
		}
		s = npb_launch_split_merge(ch, sampler, n_proposals, n_sweeps, ch->sm_detail);
		if (s != NPB_OK) return s;
		ch->z_gen++;
	}
	NPB_CUDA_OK(cudaEventRecord(ctx->ev1, ctx->stream));
	bool z_direct = false;
	if (z_out) {
		const size_t bytes = (size_t)ch->ds->N * ch->C * sizeof(npb_z_t);
		// a caller buffer that is already page-locked takes the DMA directly; pageable memory goes through the
		// handle's own pinned staging buffer
		cudaPointerAttributes attr;
		if (cudaPointerGetAttributes(&attr, z_out) == cudaSuccess && attr.type == cudaMemoryTypeHost) z_direct = true;
		else cudaGetLastError();
		if (!z_direct && !ch->h_z) NPB_CUDA_OK(cudaMallocHost((void **)&ch->h_z, bytes));
		NPB_CUDA_OK(cudaMemcpyAsync(z_direct ? (void *)z_out : (void *)ch->h_z, ch->z, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	}
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	if (z_out && !z_direct) memcpy(z_out, ch->h_z, (size_t)ch->ds->N * ch->C * sizeof(npb_z_t));
	float ms = 0.0f;
	NPB_CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	if (stats) {
		s = collect_stats(ch, before, sm_before, sampler, n_sweeps, ms, stats);
		if (s != NPB_OK) return s;
		if (stats->overflow_chains) return npb_fail(ctx, NPB_E_KMAX_OVERFLOW, "at least one chain ran out of cluster slots");
	}
	return NPB_OK;
}

npb_status npb_chains_sweep(npb_chains *ch, int sampler, int n_sweeps, npb_sweep_stats *stats) {
	return sweep_common(ch, sampler, n_sweeps, 0, stats, nullptr, nullptr);
}
npb_status npb_chains_split_merge(npb_chains *ch, int sampler, int64_t n_proposals, npb_sweep_stats *stats) {
	if (sampler != NPB_JAIN_NEAL && sampler != NPB_TRIADIC) return NPB_E_BAD_ARG;
	return sweep_common(ch, sampler, 0, n_proposals, stats, nullptr, nullptr);
}
npb_status npb_chains_update_params(npb_chains *ch, int mode, const double *mu0, double kappa0, double nu0, const double *Lambda0) {
	NpbRange nvtx_range("npb:update_params");
	if (!ch || (mode != NPB_UPDATE_POSTERIOR_DRAW && mode != NPB_UPDATE_POSTERIOR_MEAN)) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const PriorHost &p = ctx->prior;
	const bool own = mu0 && Lambda0;
	if (!own && (mu0 || Lambda0)) return NPB_E_BAD_ARG;
	if (own && !(kappa0 > 0.0 && nu0 > ch->D - 1.0)) return NPB_E_BAD_ARG;
	return npb_launch_update_params(ch, mode, own ? mu0 : p.mu0.data(), own ? kappa0 : p.kappa, own ? nu0 : p.nu,
			own ? Lambda0 : p.Lambda.data());
}

npb_status npb_chains_last_proposal(npb_chains *ch, float *detail_out) {
	if (!ch || !detail_out) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	if (!ch->sm_detail) return npb_fail(ctx, NPB_E_BAD_ARG, "no split-merge proposal has been made on this handle");
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	NPB_CUDA_OK(cudaMemcpyAsync(detail_out, ch->sm_detail, (size_t)ch->C * 16 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}
npb_status npb_chains_sweep_host(npb_chains *ch, const double *X, int sampler, int n_sweeps, uint16_t *z_out,
		npb_sweep_stats *stats) {
	NpbRange nvtx_range("npb:sweep_host (h2d + sweep + d2h)");
	if (!X) return NPB_E_BAD_ARG;
	return sweep_common(ch, sampler, n_sweeps, 0, stats, X, z_out);
}

npb_status npb_chains_get_assignments(npb_chains *ch, int64_t chain0, int64_t n, int32_t *z_out) {
	if (!ch || !z_out || chain0 < 0 || n <= 0 || chain0 + n > ch->C) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const int N = (int)ch->ds->N;
	DevBuf<int32_t> d;
	NPB_CUDA_OK(d.alloc((size_t)N * n));
	int64_t total = (int64_t)N * n;
	k_gather_chain_z<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(ch->z, d.p, N, (int)ch->C, (int)chain0, (int)n);
	NPB_CUDA_OK(cudaGetLastError());
	NPB_CUDA_OK(cudaMemcpyAsync(z_out, d.p, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}

static npb_status get_params_from(npb_chains *ch, const float *theta, const int *counts_dev, int64_t chain, int cap, int *K, int32_t *slots,
		int64_t *counts, double *mu, double *Sigma) {
	if (!ch || !K || chain < 0 || chain >= ch->C) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const int D = ch->D, TRI = npb_tri(D), PS = npb_ps(D);
	std::vector<float> th((size_t)ch->Kmax * PS);
	std::vector<int> cnt(ch->Kmax);
	NPB_CUDA_OK(cudaMemcpyAsync(th.data(), theta + (size_t)chain * ch->Kmax * PS, sizeof(float) * th.size(), cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaMemcpyAsync(cnt.data(), counts_dev + (size_t)chain * ch->Kmax, sizeof(int) * cnt.size(), cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	int k = 0;
	std::vector<double> T(TRI);
	for (int s = 0; s < ch->Kmax; ++s) {
		if (cnt[s] <= 0) continue;
		if (k < cap) {
			if (slots) slots[k] = s;
			if (counts) counts[k] = cnt[s];
			const float *o = th.data() + (size_t)s * PS;
			if (ctx->prior.family) {
				// scalar-noise slot -> (mu_0, mu_1) in mu[k, 0..1] and sigma^2 in Sigma[k, 0, 0], everything else 0; the angular
				// family reports the canonical (d, theta) its density uses (prepare(), scalarnoise_multivariatenormal.cpp:31-47)
				const double g = ctx->prior.family == NPB_FAMILY_REGRESSION ? (double)o[D + 2] : hypot((double)o[D], (double)o[D + 1]);
				if (mu) {
					for (int d = 0; d < D; ++d) mu[(size_t)k * D + d] = 0.0;
					if (ctx->prior.family == NPB_FAMILY_REGRESSION) {
						mu[(size_t)k * D] = -(double)o[D] / g;
						mu[(size_t)k * D + 1] = -(double)o[D + 1] / g;
					} else {
						mu[(size_t)k * D] = hypot((double)o[0], (double)o[1]);
						double th_ = atan2((double)o[D] / g, -(double)o[D + 1] / g);
						mu[(size_t)k * D + 1] = th_ < 0 ? th_ + 2.0 * M_PI : th_;
					}
				}
				if (Sigma) {
					for (int t = 0; t < D * D; ++t) Sigma[(size_t)k * D * D + t] = 0.0;
					const double sg = NPB_HALF_LOG2E_SQRT / g;
					Sigma[(size_t)k * D * D] = sg * sg;
				}
				k++;
				continue;
			}
			if (mu) for (int d = 0; d < D; ++d) mu[(size_t)k * D + d] = o[d];
			if (Sigma) {
				for (int t = 0; t < TRI; ++t) T[t] = (double)o[D + t] / NPB_HALF_LOG2E_SQRT;
				npb_theta_to_sigma(D, T.data(), Sigma + (size_t)k * D * D);
			}
		}
		k++;
	}
	*K = k;
	return k <= cap ? NPB_OK : NPB_E_BAD_ARG;
}

npb_status npb_chains_get_params(npb_chains *ch, int64_t chain, int cap, int *K, int32_t *slots, int64_t *counts, double *mu,
		double *Sigma) {
	if (!ch) return NPB_E_BAD_ARG;
	return get_params_from(ch, ch->theta, ch->counts, chain, cap, K, slots, counts, mu, Sigma);
}

// the clusters of the KEPT state (npb_chains_consider_max_likelihood), not of the current one: slots are re-used after a
// death, so the current slot table does not describe a snapshot taken earlier
npb_status npb_chains_get_best_params(npb_chains *ch, int64_t chain, int cap, int *K, int32_t *slots, int64_t *counts, double *mu,
		double *Sigma) {
	if (!ch) return NPB_E_BAD_ARG;
	if (!ch->best_theta) return npb_fail(ch->ctx, NPB_E_BAD_ARG, "no state has been kept on this handle yet");
	return get_params_from(ch, ch->best_theta, ch->best_counts, chain, cap, K, slots, counts, mu, Sigma);
}

// ---- the single-item seam of membertrix: retract + assign of one item of one chain (membertrix.cpp:147-233) ----------
// status: 0 ok, NPB_E_ALREADY_ASSIGNED (the item already sits in that cluster), NPB_E_ASSIGNMENT_ABSENT (no such cluster: the
// slot has no members and no parameters were given), NPB_E_KMAX_OVERFLOW (a new cluster found no free slot)
__global__ void k_move_item(npb_z_t *z, int *counts, int *kocc, float *theta, int C, int Kmax, int PS, int chain, int item, int slot,
		const float *theta_new, int *status) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	int *cnt = counts + (size_t)chain * Kmax;
	const int zo = (int)z[(size_t)item * C + chain];
	int to = slot;
	if (theta_new) { // addCluster + assign: the lowest free slot once the item is retracted (np_neal_algorithm8.cpp:136-145)
		to = -1;
		for (int k = 0; k < Kmax && to < 0; ++k)
			if (cnt[k] - (k == zo ? 1 : 0) <= 0) to = k;
		if (to < 0) { status[0] = NPB_E_KMAX_OVERFLOW; return; }
		for (int t = 0; t < PS; ++t) theta[((size_t)chain * Kmax + to) * PS + t] = theta_new[t];
	} else {
		if (to < 0 || to >= Kmax || cnt[to] <= 0) { status[0] = NPB_E_ASSIGNMENT_ABSENT; return; }
		if (to == zo) { status[0] = NPB_E_ALREADY_ASSIGNED; return; }
	}
	int occ = kocc[chain];
	if (to != zo) {
		cnt[zo] -= 1;
		if (cnt[zo] == 0) occ--;
		if (cnt[to] == 0) occ++;
		cnt[to] += 1;
	}
	kocc[chain] = occ;
	z[(size_t)item * C + chain] = (npb_z_t)to;
	status[0] = 0;
	status[1] = to;
}

static npb_status move_item(npb_chains *ch, int64_t chain, int64_t item, int slot, const double *mu, const double *Sigma, int *slot_out) {
	if (!ch || chain < 0 || chain >= ch->C || item < 0 || item >= ch->ds->N) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const int D = ch->D, TRI = npb_tri(D), PS = npb_ps(D);
	DevBuf<float> d_th;
	DevBuf<int> d_status;
	NPB_CUDA_OK(d_status.alloc(2));
	if (mu) {
		std::vector<float> th(PS);
		std::vector<double> T(TRI);
		double logdet;
		if (!npb_prepare_theta(D, mu, Sigma, T.data(), &logdet))
			return npb_fail(ctx, NPB_E_NOT_POSITIVE, "Sigma is not invertible with a positive definite symmetric precision");
		for (int d = 0; d < D; ++d) th[d] = (float)mu[d];
		for (int t = 0; t < TRI; ++t) th[D + t] = (float)(T[t] * NPB_HALF_LOG2E_SQRT);
		th[D + TRI] = (float)(-0.5 * (D * std::log2(2.0 * M_PI) + logdet / std::log(2.0)));
		NPB_CUDA_OK(d_th.alloc(PS));
		NPB_CUDA_OK(cudaMemcpyAsync(d_th.p, th.data(), sizeof(float) * PS, cudaMemcpyHostToDevice, ctx->stream));
	}
	k_move_item<<<1, 32, 0, ctx->stream>>>(ch->z, ch->counts, ch->kocc, ch->theta, (int)ch->C, ch->Kmax, PS, (int)chain, (int)item, slot,
			mu ? d_th.p : nullptr, d_status.p);
	NPB_CUDA_OK(cudaGetLastError());
	int st[2] = {0, 0};
	NPB_CUDA_OK(cudaMemcpyAsync(st, d_status.p, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	if (st[0] != 0) return npb_fail(ctx, (npb_status)st[0], st[0] == NPB_E_ALREADY_ASSIGNED ? "the item is already assigned to that cluster"
			: st[0] == NPB_E_ASSIGNMENT_ABSENT ? "no such cluster on this chain" : "no free cluster slot");
	if (slot_out) *slot_out = st[1];
	ch->z_gen++;
	return NPB_OK;
}
npb_status npb_chain_move_item(npb_chains *ch, int64_t chain, int64_t item, int slot) {
	return move_item(ch, chain, item, slot, nullptr, nullptr, nullptr);
}
npb_status npb_chain_move_item_new(npb_chains *ch, int64_t chain, int64_t item, const double *mu, const double *Sigma, int *slot_out) {
	if (!mu || !Sigma) return NPB_E_BAD_ARG;
	return move_item(ch, chain, item, -1, mu, Sigma, slot_out);
}

// every chain of the handle takes the state of chain `src` (assignments, slot table, counts): the starting point of the
// categorical tests, where 2^17 chains make the same first reassignment independently
__global__ void k_broadcast_z(npb_z_t *z, int N, int C, int src) {
	const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= (size_t)N * C) return;
	const size_t item = i / C;
	z[i] = z[item * C + src];
}
__global__ void k_broadcast_slots(float *theta, int *counts, int *kocc, int *overflow, int C, int Kmax, int PS, int src) {
	const int c = blockIdx.x;
	if (c == src) return;
	for (int t = threadIdx.x; t < Kmax * PS; t += blockDim.x) theta[(size_t)c * Kmax * PS + t] = theta[(size_t)src * Kmax * PS + t];
	for (int t = threadIdx.x; t < Kmax; t += blockDim.x) counts[(size_t)c * Kmax + t] = counts[(size_t)src * Kmax + t];
	if (threadIdx.x == 0) { kocc[c] = kocc[src]; overflow[c] = 0; }
}
npb_status npb_chains_broadcast_state(npb_chains *ch, int64_t src) {
	if (!ch || src < 0 || src >= ch->C) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const size_t n = (size_t)ch->ds->N * ch->C;
	// (z of chain src is read while other chains' entries of the same rows are written: entries of chain src itself are rewritten
	// with their own value, so the race is benign)
	k_broadcast_z<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ch->z, (int)ch->ds->N, (int)ch->C, (int)src);
	NPB_CUDA_OK(cudaGetLastError());
	k_broadcast_slots<<<(unsigned)ch->C, 128, 0, ctx->stream>>>(ch->theta, ch->counts, ch->kocc, ch->overflow, (int)ch->C, ch->Kmax, npb_ps(ch->D), (int)src);
	NPB_CUDA_OK(cudaGetLastError());
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	ch->moved_frac_last = -1.0;
	ch->z_gen++;
	return NPB_OK;
}

// membertrix::remove (membertrix.cpp:213-228): a cluster can only be removed once it has no members -- at which point the
// device has dropped it already (a slot without members is free)
npb_status npb_chain_remove_cluster(npb_chains *ch, int64_t chain, int slot) {
	if (!ch || chain < 0 || chain >= ch->C || slot < 0 || slot >= ch->Kmax) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	int n = 0;
	NPB_CUDA_OK(cudaMemcpyAsync(&n, ch->counts + (size_t)chain * ch->Kmax + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	if (n > 0) return npb_fail(ctx, NPB_E_ASSIGNMENT_REMAINING, "the cluster still has members");
	return NPB_OK;
}

// ---- conjugate Algorithm 2: parity probes ------------------------------------------------------------------------
// out [n_items, 33]: NIW posterior-predictive log-density of the items under every cluster of `chain` (NaN without members),
// as the sweep kernel evaluates it for a cluster the item is not a member of; column 32 = the prior predictive
npb_status npb_chains_alg2_logpred(npb_chains *ch, int64_t chain, const int32_t *items, int n_items, float *out) {
	if (!ch || !items || !out || n_items <= 0 || chain < 0 || chain >= ch->C) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	for (int j = 0; j < n_items; ++j)
		if (items[j] < 0 || items[j] >= ch->ds->N) return NPB_E_BAD_ARG;
	DevBuf<int32_t> d_items;
	DevBuf<float> d_out;
	NPB_CUDA_OK(d_items.alloc(n_items));
	NPB_CUDA_OK(d_out.alloc((size_t)n_items * 33));
	NPB_CUDA_OK(cudaMemcpyAsync(d_items.p, items, sizeof(int32_t) * n_items, cudaMemcpyHostToDevice, ctx->stream));
	npb_status s = npb_launch_alg2_probe(ch, (int)chain, d_items.p, n_items, d_out.p);
	if (s != NPB_OK) return s;
	NPB_CUDA_OK(cudaMemcpyAsync(out, d_out.p, sizeof(float) * n_items * 33, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}
// the sufficient statistics the conjugate path keeps for `chain`: counts [32], sum x [32, D], sum x x^T [32, D, D] (fp64)
npb_status npb_chains_alg2_suffstats(npb_chains *ch, int64_t chain, int32_t *counts, double *sx, double *sxx) {
	if (!ch || !counts || !sx || !sxx || chain < 0 || chain >= ch->C) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	if (!ch->a2_sx) return npb_fail(ctx, NPB_E_BAD_ARG, "the conjugate Algorithm 2 path has not run on this handle");
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const int D = ch->D;
	NPB_CUDA_OK(cudaMemcpyAsync(counts, ch->counts + (size_t)chain * 32, sizeof(int) * 32, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaMemcpyAsync(sx, ch->a2_sx + (size_t)chain * 32 * D, sizeof(double) * 32 * D, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaMemcpyAsync(sxx, ch->a2_sxx + (size_t)chain * 32 * D * D, sizeof(double) * 32 * D * D, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}

// ---- incremental result copy ------------------------------------------------------------------------------------
// entries of z that differ from the snapshot: (index, value) appended to a list (a warp reserves its run with one atomic),
// the snapshot brought up to date.  Eight entries per thread (one 16-byte load of each array).
__global__ void __launch_bounds__(256) k_z_delta(const npb_z_t *z, npb_z_t *prev, size_t n8, uint32_t *idx, npb_z_t *val, size_t cap,
		unsigned long long *count) {
	const size_t i8 = (size_t)blockIdx.x * 256 + threadIdx.x;
	uint4 a = make_uint4(0, 0, 0, 0), b = a;
	if (i8 < n8) {
		a = reinterpret_cast<const uint4 *>(z)[i8];
		b = reinterpret_cast<const uint4 *>(prev)[i8];
	}
	const unsigned short *pa = reinterpret_cast<const unsigned short *>(&a), *pb = reinterpret_cast<const unsigned short *>(&b);
	unsigned diff = 0u;
#pragma unroll
	for (int e = 0; e < 8; ++e) diff |= (pa[e] != pb[e]) ? (1u << e) : 0u;
	const int mine = __popc(diff);
	// exclusive scan of the counts over the warp, one atomic per warp
	int incl = mine;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const int t = __shfl_up_sync(0xffffffffu, incl, o);
		if ((int)(threadIdx.x & 31) >= o) incl += t;
	}
	const int total = __shfl_sync(0xffffffffu, incl, 31);
	if (total == 0) return;
	unsigned long long base = 0ull;
	if ((threadIdx.x & 31) == 31) base = atomicAdd(count, (unsigned long long)total);
	base = __shfl_sync(0xffffffffu, base, 31);
	if (mine) {
		size_t o = (size_t)base + (size_t)(incl - mine);
#pragma unroll
		for (int e = 0; e < 8; ++e) {
			if ((diff >> e) & 1u) {
				if (o < cap) {
					idx[o] = (uint32_t)(i8 * 8 + e);
					val[o] = pa[e];
				}
				++o;
			}
		}
		reinterpret_cast<uint4 *>(prev)[i8] = a;
	}
}

npb_status npb_chains_sweep_host_delta(npb_chains *ch, const double *X, int sampler, int n_sweeps, uint16_t *z_mirror,
		npb_sweep_stats *stats, int64_t *n_changed) {
	NpbRange nvtx_range("npb:sweep_host_delta (h2d + sweep + changed entries d2h)");
	if (!ch || !z_mirror) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const size_t n = (size_t)ch->ds->N * ch->C;
	if (n >= ((size_t)1 << 32) || (n & 7)) return npb_fail(ctx, NPB_E_UNSUPPORTED, "the incremental copy needs N * chains below 2^32 and a multiple of 8");
	npb_status s = sweep_common(ch, sampler, n_sweeps, 0, stats, X, nullptr);
	if (s != NPB_OK) return s;
	const bool first = ch->z_prev == nullptr;
	if (first) {
		ch->dl_cap = n / 4;
		NPB_CUDA_OK(cudaMalloc((void **)&ch->z_prev, n * sizeof(npb_z_t)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->dl_idx, ch->dl_cap * sizeof(uint32_t)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->dl_val, ch->dl_cap * sizeof(npb_z_t)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->dl_count, sizeof(unsigned long long)));
	}
	unsigned long long cnt = 0ull;
	if (!first) {
		NPB_CUDA_OK(cudaMemsetAsync(ch->dl_count, 0, sizeof(unsigned long long), ctx->stream));
		k_z_delta<<<(unsigned)((n / 8 + 255) / 256), 256, 0, ctx->stream>>>(ch->z, ch->z_prev, n / 8, ch->dl_idx, ch->dl_val, ch->dl_cap, ch->dl_count);
		NPB_CUDA_OK(cudaGetLastError());
		NPB_CUDA_OK(cudaMemcpyAsync(&cnt, ch->dl_count, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
		NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	}
	if (first || cnt > ch->dl_cap) {
		// everything travels: the first call, or more than a quarter of the entries changed (the list would be larger than the array)
		cudaPointerAttributes attr;
		const bool direct = cudaPointerGetAttributes(&attr, z_mirror) == cudaSuccess && attr.type == cudaMemoryTypeHost;
		if (!direct) cudaGetLastError();
		if (!direct && !ch->h_z) NPB_CUDA_OK(cudaMallocHost((void **)&ch->h_z, n * sizeof(npb_z_t)));
		NPB_CUDA_OK(cudaMemcpyAsync(direct ? (void *)z_mirror : (void *)ch->h_z, ch->z, n * sizeof(npb_z_t), cudaMemcpyDeviceToHost, ctx->stream));
		if (first) NPB_CUDA_OK(cudaMemcpyAsync(ch->z_prev, ch->z, n * sizeof(npb_z_t), cudaMemcpyDeviceToDevice, ctx->stream));
		NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
		if (!direct) memcpy(z_mirror, ch->h_z, n * sizeof(npb_z_t));
		if (n_changed) *n_changed = first ? (int64_t)n : (int64_t)cnt;
		return NPB_OK;
	}
	if (cnt > 0) {
		if (cnt > ch->h_dl_cap) {
			if (ch->h_dl_idx) cudaFreeHost(ch->h_dl_idx);
			if (ch->h_dl_val) cudaFreeHost(ch->h_dl_val);
			ch->h_dl_idx = nullptr;
			ch->h_dl_val = nullptr;
			size_t cap = ch->h_dl_cap ? ch->h_dl_cap : 4096;
			while (cap < cnt) cap *= 2;
			NPB_CUDA_OK(cudaMallocHost((void **)&ch->h_dl_idx, cap * sizeof(uint32_t)));
			NPB_CUDA_OK(cudaMallocHost((void **)&ch->h_dl_val, cap * sizeof(npb_z_t)));
			ch->h_dl_cap = cap;
		}
		NPB_CUDA_OK(cudaMemcpyAsync(ch->h_dl_idx, ch->dl_idx, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		NPB_CUDA_OK(cudaMemcpyAsync(ch->h_dl_val, ch->dl_val, cnt * sizeof(npb_z_t), cudaMemcpyDeviceToHost, ctx->stream));
		NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
		for (size_t i = 0; i < cnt; ++i) z_mirror[ch->h_dl_idx[i]] = ch->h_dl_val[i];
	}
	if (n_changed) *n_changed = (int64_t)cnt;
	return NPB_OK;
}

npb_status npb_chains_metrics(npb_chains *ch, const int32_t *truth, double *purity, double *rand_index, double *adjusted_rand,
		double *joint_loglik, int32_t *K) {
	NpbRange nvtx_range("npb:metrics");
	if (!ch) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const int N = (int)ch->ds->N;
	const size_t C = (size_t)ch->C;
	int Ktrue = 1;
	DevBuf<int32_t> d_truth, d_K;
	DevBuf<double> d_out; // purity, ri, ari, jll
	if (truth) {
		for (int i = 0; i < N; ++i) {
			if (truth[i] < 0) return NPB_E_BAD_ARG;
			if (truth[i] + 1 > Ktrue) Ktrue = truth[i] + 1;
		}
		NPB_CUDA_OK(d_truth.alloc(N));
		NPB_CUDA_OK(cudaMemcpyAsync(d_truth.p, truth, sizeof(int32_t) * N, cudaMemcpyHostToDevice, ctx->stream));
	}
	NPB_CUDA_OK(d_out.alloc(C * 4));
	NPB_CUDA_OK(d_K.alloc(C));
	npb_status s = npb_launch_metrics(ch, truth ? d_truth.p : nullptr, Ktrue, d_out.p, d_out.p + C, d_out.p + 2 * C,
			joint_loglik ? d_out.p + 3 * C : nullptr, d_K.p);
	if (s != NPB_OK) return s;
	std::vector<double> h(C * 4);
	std::vector<int32_t> hK(C);
	NPB_CUDA_OK(cudaMemcpyAsync(h.data(), d_out.p, sizeof(double) * C * 4, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaMemcpyAsync(hK.data(), d_K.p, sizeof(int32_t) * C, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	if (truth) {
		if (purity) memcpy(purity, h.data(), sizeof(double) * C);
		if (rand_index) memcpy(rand_index, h.data() + C, sizeof(double) * C);
		if (adjusted_rand) memcpy(adjusted_rand, h.data() + 2 * C, sizeof(double) * C);
	}
	if (joint_loglik) memcpy(joint_loglik, h.data() + 3 * C, sizeof(double) * C);
	if (K) memcpy(K, hK.data(), sizeof(int32_t) * C);
	return NPB_OK;
}

npb_status npb_cocluster(npb_chains *ch, const int64_t *anchors, int64_t n_anchor, float *S_dev, int accumulate) {
	NpbRange nvtx_range("npb:cocluster");
	if (!ch || !anchors || !S_dev || n_anchor <= 0 || n_anchor > 65535) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	for (int64_t a = 0; a < n_anchor; ++a)
		if (anchors[a] < 0 || anchors[a] >= ch->ds->N) return NPB_E_BAD_ARG;
	DevBuf<int64_t> d_a;
	NPB_CUDA_OK(d_a.alloc(n_anchor));
	NPB_CUDA_OK(cudaMemcpyAsync(d_a.p, anchors, sizeof(int64_t) * n_anchor, cudaMemcpyHostToDevice, ctx->stream));
	npb_status s = npb_launch_cocluster(ch, d_a.p, (int)n_anchor, S_dev, accumulate);
	if (s != NPB_OK) return s;
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}

npb_status npb_cocluster_allreduce(npb_chains *ch, const int64_t *anchors, int64_t n_anchor, npb_comm *comm, float *S_dev) {
	NpbRange nvtx_range("npb:cocluster_allreduce");
	npb_status s = npb_cocluster(ch, anchors, n_anchor, S_dev, 0);
	if (s != NPB_OK || !comm) return s;
	s = npb_comm_allreduce_sum(comm, S_dev, n_anchor * n_anchor, 32);
	if (s != NPB_OK) return s;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}

// the same into a HOST matrix (a host above the ABI needs no CUDA of its own)
npb_status npb_cocluster_host(npb_chains *ch, const int64_t *anchors, int64_t n_anchor, npb_comm *comm, float *S_host) {
	if (!ch || !S_host || n_anchor <= 0) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	DevBuf<float> d_S;
	NPB_CUDA_OK(d_S.alloc((size_t)n_anchor * n_anchor));
	npb_status s = npb_cocluster_allreduce(ch, anchors, n_anchor, comm, d_S.p);
	if (s != NPB_OK) return s;
	NPB_CUDA_OK(cudaMemcpyAsync(S_host, d_S.p, sizeof(float) * n_anchor * n_anchor, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}

npb_status npb_scan_order_host(uint64_t seed, uint32_t sweep, int64_t N, int32_t *order_out) {
	if (!order_out || N <= 0 || N > 0x7fffffff) return NPB_E_BAD_ARG;
	const ScanOrder so = npb_scan_order(seed, sweep, (uint32_t)N);
	for (int64_t s = 0; s < N; ++s) order_out[s] = (int32_t)npb_scan_item(so, (uint32_t)s);
	return NPB_OK;
}

npb_status npb_fp32_peak(npb_ctx *ctx, double *tflops) {
	if (!ctx || !tflops) return NPB_E_BAD_ARG;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	return npb_launch_fma_peak(ctx, tflops);
}

npb_status npb_chain_update_alg8(npb_chains *ch, int64_t chain, int64_t item) {
	if (ch) ch->z_gen++;
	if (!ch) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	// chain < 0: the same item on every chain (what a lockstep driver of the reference's per-item loop does)
	if (chain >= ch->C || item < 0 || item >= ch->ds->N) return NPB_E_BAD_ARG;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	npb_status s = ensure_whitened(ch->ds);
	if (s != NPB_OK) return s;
	s = chain < 0 ? npb_launch_update_item(ch, 0, ch->C, item) : npb_launch_update_item(ch, chain, 1, item);
	if (s != NPB_OK) return s;
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}

// theta in the replay kernel's double layout: mu[D], T upper packed, c = -0.5 (D log 2pi + log det Sigma)
static bool pack_theta64(int D, const double *mu, const double *Sigma, double *out) {
	const int TRI = npb_tri(D);
	double logdet;
	if (!npb_prepare_theta(D, mu, Sigma, out + D, &logdet)) return false;
	for (int d = 0; d < D; ++d) out[d] = mu[d];
	out[D + TRI] = -0.5 * (D * std::log(2.0 * M_PI) + logdet);
	return true;
}

npb_status npb_replay_alg8(npb_ctx *ctx, npb_dataset *ds, int m_aux, int nslots, const int32_t *z0, int K0,
		const int32_t *slots0, const double *mu0, const double *Sigma0, int64_t n_steps, const int32_t *item,
		const int64_t *order_off, const int32_t *order, const double *aux_mu, const double *aux_Sigma,
		const double *u, const int32_t *new_slot, int32_t *picked_out, int64_t z_every, int32_t *z_after_out) {
	if (!ctx || !ds || ds->ctx != ctx || !z0 || !slots0 || !mu0 || !Sigma0 || !item || !order_off || !order || !aux_mu ||
			!aux_Sigma || !u || !new_slot || !picked_out || m_aux <= 0 || nslots <= 0 || K0 <= 0 || K0 > nslots || n_steps <= 0)
		return NPB_E_BAD_ARG;
	if (!ctx->prior.set || ctx->prior.D != ds->D) return npb_fail(ctx, NPB_E_BAD_ARG, "set a prior of the dataset's dimension first");
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const int D = ds->D, PS = npb_ps(D), N = (int)ds->N;
	std::vector<double> theta((size_t)nslots * PS, 0.0);
	std::vector<int> counts(nslots, 0);
	for (int k = 0; k < K0; ++k) {
		if (slots0[k] < 0 || slots0[k] >= nslots) return NPB_E_BAD_ARG;
		if (!pack_theta64(D, mu0 + (size_t)k * D, Sigma0 + (size_t)k * D * D, theta.data() + (size_t)slots0[k] * PS))
			return npb_fail(ctx, NPB_E_NOT_POSITIVE, "initial Sigma not invertible");
	}
	for (int i = 0; i < N; ++i) {
		if (z0[i] < 0 || z0[i] >= nslots) return NPB_E_BAD_ARG;
		counts[z0[i]]++;
	}
	for (int64_t s = 0; s < n_steps; ++s)
		if (item[s] < 0 || item[s] >= N || order_off[s + 1] < order_off[s]) return NPB_E_BAD_ARG;
	std::vector<double> aux((size_t)n_steps * m_aux * PS);
	for (int64_t t = 0; t < n_steps * m_aux; ++t)
		if (!pack_theta64(D, aux_mu + (size_t)t * D, aux_Sigma + (size_t)t * D * D, aux.data() + (size_t)t * PS))
			return npb_fail(ctx, NPB_E_NOT_POSITIVE, "auxiliary Sigma not invertible");
	const int64_t n_order = order_off[n_steps];
	const int64_t n_snap = (z_after_out && z_every > 0) ? n_steps / z_every : 0;
	DevBuf<double> d_theta, d_aux, d_u;
	DevBuf<int> d_counts, d_status;
	DevBuf<int32_t> d_z, d_item, d_order, d_new, d_picked, d_zafter;
	DevBuf<int64_t> d_off;
	NPB_CUDA_OK(d_theta.alloc(theta.size()));
	NPB_CUDA_OK(d_aux.alloc(aux.size()));
	NPB_CUDA_OK(d_u.alloc(n_steps));
	NPB_CUDA_OK(d_counts.alloc(nslots));
	NPB_CUDA_OK(d_status.alloc(1));
	NPB_CUDA_OK(d_z.alloc(N));
	NPB_CUDA_OK(d_item.alloc(n_steps));
	NPB_CUDA_OK(d_order.alloc(n_order));
	NPB_CUDA_OK(d_new.alloc(n_steps));
	NPB_CUDA_OK(d_picked.alloc(n_steps));
	NPB_CUDA_OK(d_zafter.alloc((size_t)n_snap * N));
	NPB_CUDA_OK(d_off.alloc(n_steps + 1));
	cudaStream_t st = ctx->stream;
	int zero = 0;
	NPB_CUDA_OK(cudaMemcpyAsync(d_theta.p, theta.data(), sizeof(double) * theta.size(), cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_aux.p, aux.data(), sizeof(double) * aux.size(), cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_u.p, u, sizeof(double) * n_steps, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_counts.p, counts.data(), sizeof(int) * nslots, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_status.p, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_z.p, z0, sizeof(int32_t) * N, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_item.p, item, sizeof(int32_t) * n_steps, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_order.p, order, sizeof(int32_t) * n_order, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_new.p, new_slot, sizeof(int32_t) * n_steps, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_off.p, order_off, sizeof(int64_t) * (n_steps + 1), cudaMemcpyHostToDevice, st));
	npb_status s = npb_launch_replay(ctx, ds->X64, N, D, m_aux, ctx->prior.alpha, nslots, d_theta.p, d_counts.p, d_z.p, n_steps,
			d_item.p, d_off.p, d_order.p, d_aux.p, d_u.p, d_new.p, d_picked.p, z_every, n_snap ? d_zafter.p : nullptr, d_status.p);
	if (s != NPB_OK) return s;
	int status = 0;
	NPB_CUDA_OK(cudaMemcpyAsync(picked_out, d_picked.p, sizeof(int32_t) * n_steps, cudaMemcpyDeviceToHost, st));
	if (n_snap) NPB_CUDA_OK(cudaMemcpyAsync(z_after_out, d_zafter.p, sizeof(int32_t) * n_snap * N, cudaMemcpyDeviceToHost, st));
	NPB_CUDA_OK(cudaMemcpyAsync(&status, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
	NPB_CUDA_OK(cudaStreamSynchronize(st));
	if (status != 0) return npb_fail(ctx, status, "replay stopped: candidate index past the end or slot out of range");
	return NPB_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// split-merge replay (npb_replay_sm.cu)
// ---------------------------------------------------------------------------------------------------------------------
npb_status npb_launch_replay_sm(npb_ctx *ctx, const RsmArgs &a);

// theta in the split-merge replay layout: mu[D], T upper packed, c, cst = sqrt((2 pi)^D det Sigma)
static bool pack_theta64r(int D, const double *mu, const double *Sigma, double *out) {
	const int TRI = npb_tri(D);
	double logdet;
	if (!npb_prepare_theta(D, mu, Sigma, out + D, &logdet)) return false;
	for (int d = 0; d < D; ++d) out[d] = mu[d];
	out[D + TRI] = -0.5 * (D * std::log(2.0 * M_PI) + logdet);
	out[D + TRI + 1] = std::exp(0.5 * (D * std::log(2.0 * M_PI) + logdet));
	return true;
}

npb_status npb_replay_split_merge(npb_ctx *ctx, npb_dataset *ds, int sampler, int nslots, const int32_t *z0, int K0,
		const int32_t *slots0, const double *mu0, const double *Sigma0, int64_t n_prop, const int32_t *picks, const double *u0,
		const double *new_mu, const double *new_Sigma, const int64_t *pool_off, const int32_t *pool, const double *us,
		const double *uacc, const int32_t *new_slot, int32_t *type_out, int32_t *dec_out, int32_t *accept_out, double *logA_out,
		int32_t *z_final_out) {
	if (!ctx || !ds || ds->ctx != ctx || !z0 || !slots0 || !mu0 || !Sigma0 || !picks || !u0 || !new_mu || !new_Sigma || !pool_off ||
			!pool || !us || !uacc || !new_slot || !type_out || !dec_out || !accept_out || !logA_out || nslots <= 0 || K0 <= 0 ||
			K0 > nslots || n_prop <= 0 || (sampler != NPB_JAIN_NEAL && sampler != NPB_TRIADIC))
		return NPB_E_BAD_ARG;
	if (!ctx->prior.set || ctx->prior.D != ds->D) return npb_fail(ctx, NPB_E_BAD_ARG, "set a prior of the dataset's dimension first");
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const int D = ds->D, PSR = npb_ps(D) + 1, N = (int)ds->N;
	const int nsub = sampler == NPB_JAIN_NEAL ? 2 : 3;
	std::vector<double> theta((size_t)nslots * PSR, 0.0);
	std::vector<int> counts(nslots, 0);
	for (int k = 0; k < K0; ++k) {
		if (slots0[k] < 0 || slots0[k] >= nslots) return NPB_E_BAD_ARG;
		if (!pack_theta64r(D, mu0 + (size_t)k * D, Sigma0 + (size_t)k * D * D, theta.data() + (size_t)slots0[k] * PSR))
			return npb_fail(ctx, NPB_E_NOT_POSITIVE, "initial Sigma not invertible");
	}
	for (int i = 0; i < N; ++i) {
		if (z0[i] < 0 || z0[i] >= nslots) return NPB_E_BAD_ARG;
		counts[z0[i]]++;
	}
	const int64_t L = pool_off[n_prop];
	for (int64_t p = 0; p < n_prop; ++p) {
		if (pool_off[p + 1] < pool_off[p]) return NPB_E_BAD_ARG;
		for (int j = 0; j < nsub; ++j)
			if (picks[p * 3 + j] < 0 || picks[p * 3 + j] >= N) return NPB_E_BAD_ARG;
	}
	for (int64_t t = 0; t < L; ++t)
		if (pool[t] < 0 || pool[t] >= N) return NPB_E_BAD_ARG;
	std::vector<double> thn((size_t)n_prop * PSR);
	for (int64_t p = 0; p < n_prop; ++p)
		if (!pack_theta64r(D, new_mu + (size_t)p * D, new_Sigma + (size_t)p * D * D, thn.data() + (size_t)p * PSR))
			return npb_fail(ctx, NPB_E_NOT_POSITIVE, "a recorded prior draw has a singular Sigma");
	DevBuf<double> d_theta, d_thn, d_u0, d_us, d_uacc, d_logA;
	DevBuf<int> d_counts, d_status;
	DevBuf<int32_t> d_z, d_picks, d_pool, d_new, d_type, d_dec, d_acc;
	DevBuf<int64_t> d_off;
	NPB_CUDA_OK(d_theta.alloc(theta.size()));
	NPB_CUDA_OK(d_thn.alloc(thn.size()));
	NPB_CUDA_OK(d_u0.alloc(n_prop));
	NPB_CUDA_OK(d_us.alloc(L));
	NPB_CUDA_OK(d_uacc.alloc(n_prop));
	NPB_CUDA_OK(d_logA.alloc(n_prop));
	NPB_CUDA_OK(d_counts.alloc(nslots));
	NPB_CUDA_OK(d_status.alloc(1));
	NPB_CUDA_OK(d_z.alloc(N));
	NPB_CUDA_OK(d_picks.alloc(n_prop * 3));
	NPB_CUDA_OK(d_pool.alloc(L));
	NPB_CUDA_OK(d_new.alloc(n_prop));
	NPB_CUDA_OK(d_type.alloc(n_prop));
	NPB_CUDA_OK(d_dec.alloc(L));
	NPB_CUDA_OK(d_acc.alloc(n_prop));
	NPB_CUDA_OK(d_off.alloc(n_prop + 1));
	cudaStream_t st = ctx->stream;
	int zero = 0;
	NPB_CUDA_OK(cudaMemcpyAsync(d_theta.p, theta.data(), sizeof(double) * theta.size(), cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_thn.p, thn.data(), sizeof(double) * thn.size(), cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_u0.p, u0, sizeof(double) * n_prop, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_us.p, us, sizeof(double) * L, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_uacc.p, uacc, sizeof(double) * n_prop, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_counts.p, counts.data(), sizeof(int) * nslots, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_status.p, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_z.p, z0, sizeof(int32_t) * N, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_picks.p, picks, sizeof(int32_t) * n_prop * 3, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_pool.p, pool, sizeof(int32_t) * L, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_new.p, new_slot, sizeof(int32_t) * n_prop, cudaMemcpyHostToDevice, st));
	NPB_CUDA_OK(cudaMemcpyAsync(d_off.p, pool_off, sizeof(int64_t) * (n_prop + 1), cudaMemcpyHostToDevice, st));
	RsmArgs a;
	a.X = ds->X64; a.N = N; a.D = D; a.sampler = sampler; a.nslots = nslots; a.alpha = ctx->prior.alpha;
	a.theta = d_theta.p; a.counts = d_counts.p; a.z = d_z.p; a.n_prop = n_prop; a.picks = d_picks.p; a.u0 = d_u0.p;
	a.th_new = d_thn.p; a.pool_off = d_off.p; a.pool = d_pool.p; a.us = d_us.p; a.uacc = d_uacc.p; a.new_slot = d_new.p;
	a.type_out = d_type.p; a.dec_out = d_dec.p; a.accept_out = d_acc.p; a.logA_out = d_logA.p; a.status = d_status.p;
	npb_status s = npb_launch_replay_sm(ctx, a);
	if (s != NPB_OK) return s;
	int status = 0;
	NPB_CUDA_OK(cudaMemcpyAsync(type_out, d_type.p, sizeof(int32_t) * n_prop, cudaMemcpyDeviceToHost, st));
	NPB_CUDA_OK(cudaMemcpyAsync(dec_out, d_dec.p, sizeof(int32_t) * L, cudaMemcpyDeviceToHost, st));
	NPB_CUDA_OK(cudaMemcpyAsync(accept_out, d_acc.p, sizeof(int32_t) * n_prop, cudaMemcpyDeviceToHost, st));
	NPB_CUDA_OK(cudaMemcpyAsync(logA_out, d_logA.p, sizeof(double) * n_prop, cudaMemcpyDeviceToHost, st));
	if (z_final_out) NPB_CUDA_OK(cudaMemcpyAsync(z_final_out, d_z.p, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
	NPB_CUDA_OK(cudaMemcpyAsync(&status, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
	NPB_CUDA_OK(cudaStreamSynchronize(st));
	if (status != 0) return npb_fail(ctx, (npb_status)status, "split-merge replay left the recorded trajectory");
	return NPB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// max-likelihood snapshot: MCMC::considerMaxLikelihood (np_mcmc.cpp:187-203) for every chain at once
// ---------------------------------------------------------------------------------------------------------------------
__global__ void k_best_init(double *best, int C) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c < C) best[c] = -INFINITY;
}
// chains whose current joint log-likelihood beats the kept one copy their column of z (item-major: a warp covers 32
// neighbouring chains of one item, so the copy is coalesced where it happens)
__global__ void k_best_copy_z(const npb_z_t *z, npb_z_t *best_z, const double *cur, const double *best, int N, int C) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= C || !(cur[c] > best[c])) return;
	for (int i = blockIdx.y; i < N; i += gridDim.y) best_z[(size_t)i * C + c] = z[(size_t)i * C + c];
}
__global__ void k_best_copy_params(const float *theta, const int *counts, float *best_theta, int *best_counts, const double *cur,
		double *best, int C, int Kmax, int PS) {
	const int c = blockIdx.x;
	if (!(cur[c] > best[c])) return;
	for (int t = threadIdx.x; t < Kmax * PS; t += blockDim.x) best_theta[(size_t)c * Kmax * PS + t] = theta[(size_t)c * Kmax * PS + t];
	for (int t = threadIdx.x; t < Kmax; t += blockDim.x) best_counts[(size_t)c * Kmax + t] = counts[(size_t)c * Kmax + t];
	__syncthreads();
	if (threadIdx.x == 0) best[c] = cur[c]; // last: the z copy (earlier launch) and this block read the old value
}

npb_status npb_chains_consider_max_likelihood(npb_chains *ch, double *joint_loglik_out, double *best_out) {
	if (!ch) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const int N = (int)ch->ds->N, C = (int)ch->C, PS = npb_ps(ch->D);
	if (!ch->best_z) {
		NPB_CUDA_OK(cudaMalloc((void **)&ch->best_z, (size_t)N * C * sizeof(npb_z_t)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->best_theta, (size_t)C * ch->Kmax * PS * sizeof(float)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->best_counts, (size_t)C * ch->Kmax * sizeof(int)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->best_jll, (size_t)C * sizeof(double)));
		NPB_CUDA_OK(cudaMalloc((void **)&ch->cur_jll, (size_t)C * 4 * sizeof(double)));
		k_best_init<<<(C + 255) / 256, 256, 0, ctx->stream>>>(ch->best_jll, C);
		NPB_CUDA_OK(cudaGetLastError());
	}
	DevBuf<int32_t> d_K;
	NPB_CUDA_OK(d_K.alloc(C));
	// joint log-likelihood of the current state: sum_k sum_{i in k} log p(x_i | theta_k)
	npb_status s = npb_launch_metrics(ch, nullptr, 1, ch->cur_jll + C, ch->cur_jll + 2 * (size_t)C, ch->cur_jll + 3 * (size_t)C, ch->cur_jll, d_K.p);
	if (s != NPB_OK) return s;
	dim3 gz((C + 127) / 128, 64);
	k_best_copy_z<<<gz, 128, 0, ctx->stream>>>(ch->z, ch->best_z, ch->cur_jll, ch->best_jll, N, C);
	NPB_CUDA_OK(cudaGetLastError());
	k_best_copy_params<<<C, 128, 0, ctx->stream>>>(ch->theta, ch->counts, ch->best_theta, ch->best_counts, ch->cur_jll, ch->best_jll, C,
			ch->Kmax, PS);
	NPB_CUDA_OK(cudaGetLastError());
	if (joint_loglik_out) NPB_CUDA_OK(cudaMemcpyAsync(joint_loglik_out, ch->cur_jll, sizeof(double) * C, cudaMemcpyDeviceToHost, ctx->stream));
	if (best_out) NPB_CUDA_OK(cudaMemcpyAsync(best_out, ch->best_jll, sizeof(double) * C, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}

npb_status npb_chains_get_best_assignments(npb_chains *ch, int64_t chain0, int64_t n, int32_t *z_out) {
	if (!ch || !z_out || chain0 < 0 || n <= 0 || chain0 + n > ch->C) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	if (!ch->best_z) return npb_fail(ctx, NPB_E_BAD_ARG, "npb_chains_consider_max_likelihood has not been called on this handle");
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	const int N = (int)ch->ds->N;
	DevBuf<int32_t> d;
	NPB_CUDA_OK(d.alloc((size_t)n * N));
	k_gather_chain_z<<<(unsigned)(((size_t)n * N + 255) / 256), 256, 0, ctx->stream>>>(ch->best_z, d.p, N, (int)ch->C, (int)chain0, (int)n);
	NPB_CUDA_OK(cudaGetLastError());
	NPB_CUDA_OK(cudaMemcpyAsync(z_out, d.p, sizeof(int32_t) * n * N, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}

npb_status npb_chains_probe_tile_logdensity(npb_chains *ch, int64_t chain, const int32_t *items32, float *out) {
	if (!ch || !items32 || !out || chain < 0 || chain >= ch->C) return NPB_E_BAD_ARG;
	npb_ctx *ctx = ch->ctx;
	NPB_CUDA_OK(cudaSetDevice(ctx->device));
	for (int j = 0; j < 32; ++j)
		if (items32[j] < 0 || items32[j] >= ch->ds->N) return NPB_E_BAD_ARG;
	DevBuf<int32_t> d_items;
	DevBuf<float> d_out;
	NPB_CUDA_OK(d_items.alloc(32));
	NPB_CUDA_OK(d_out.alloc(32 * 32));
	NPB_CUDA_OK(cudaMemcpyAsync(d_items.p, items32, sizeof(int32_t) * 32, cudaMemcpyHostToDevice, ctx->stream));
	npb_status s = npb_launch_tile_probe(ch, (int)chain, d_items.p, d_out.p);
	if (s != NPB_OK) return s;
	NPB_CUDA_OK(cudaMemcpyAsync(out, d_out.p, sizeof(float) * 32 * 32, cudaMemcpyDeviceToHost, ctx->stream));
	NPB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
	return NPB_OK;
}
