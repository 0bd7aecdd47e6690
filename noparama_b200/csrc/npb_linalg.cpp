// npb_linalg.cpp -- host-side (double) preparation of cluster parameters for the device kernels.
//
// The reference evaluates  q = d^T inverse(Sigma) d  and  det(Sigma)  with Eigen's general (LU) routines on
// every call (src/statistics/multivariatenormal.cpp:87-90) and accepts any invertible Sigma (its known-answer
// test uses a non-symmetric one, test/test_mvn_likelihood.cpp:22).  Here each theta is prepared ONCE:
//     P  = inverse(Sigma)            (Gauss-Jordan, partial pivoting; det from the pivots)
//     Ps = (P + P^T)/2               (q only sees the symmetric part)
//     Ps = C C^T,  T = C^T           (upper triangular)  =>  q = |T d|^2
// and the kernels work with (mu, T, log det Sigma).
#include "npb_internal.h"
#include <cmath>
#include <vector>

static bool gauss_jordan_inverse(int n, const double *A, double *inv, double *det_out) {
	std::vector<double> a(A, A + (size_t)n * n);
	for (int i = 0; i < n; ++i)
		for (int j = 0; j < n; ++j) inv[i * n + j] = (i == j) ? 1.0 : 0.0;
	double det = 1.0;
	for (int c = 0; c < n; ++c) {
		int p = c;
		for (int r = c + 1; r < n; ++r)
			if (std::fabs(a[r * n + c]) > std::fabs(a[p * n + c])) p = r;
		if (a[p * n + c] == 0.0 || !std::isfinite(a[p * n + c])) return false;
		if (p != c) {
			for (int j = 0; j < n; ++j) {
				std::swap(a[p * n + j], a[c * n + j]);
				std::swap(inv[p * n + j], inv[c * n + j]);
			}
			det = -det;
		}
		double piv = a[c * n + c];
		det *= piv;
		double ip = 1.0 / piv;
		for (int j = 0; j < n; ++j) {
			a[c * n + j] *= ip;
			inv[c * n + j] *= ip;
		}
		for (int r = 0; r < n; ++r) {
			if (r == c) continue;
			double f = a[r * n + c];
			if (f == 0.0) continue;
			for (int j = 0; j < n; ++j) {
				a[r * n + j] -= f * a[c * n + j];
				inv[r * n + j] -= f * inv[c * n + j];
			}
		}
	}
	*det_out = det;
	return true;
}

static bool cholesky_lower(int n, const double *A, double *C) {
	for (int i = 0; i < n * n; ++i) C[i] = 0.0;
	for (int j = 0; j < n; ++j) {
		double s = A[j * n + j];
		for (int k = 0; k < j; ++k) s -= C[j * n + k] * C[j * n + k];
		if (!(s > 0.0) || !std::isfinite(s)) return false;
		double d = std::sqrt(s);
		C[j * n + j] = d;
		for (int i = j + 1; i < n; ++i) {
			double t = A[i * n + j];
			for (int k = 0; k < j; ++k) t -= C[i * n + k] * C[j * n + k];
			C[i * n + j] = t / d;
		}
	}
	return true;
}

bool npb_prepare_theta(int D, const double * /*mu*/, const double *Sigma, double *T, double *logdet) {
	std::vector<double> P((size_t)D * D), Ps((size_t)D * D), C((size_t)D * D);
	double det;
	if (!gauss_jordan_inverse(D, Sigma, P.data(), &det)) return false;
	if (!(det > 0.0)) return false;
	for (int i = 0; i < D; ++i)
		for (int j = 0; j < D; ++j) Ps[i * D + j] = 0.5 * (P[i * D + j] + P[j * D + i]);
	if (!cholesky_lower(D, Ps.data(), C.data())) return false;
	for (int i = 0; i < D; ++i)
		for (int j = i; j < D; ++j) T[npb_tri_off(D, i, j)] = C[j * D + i];
	*logdet = std::log(det);
	return true;
}

// Sigma' = v^2 A with A = L^T L, L = lower Cholesky of Lambda (include/statistics/invwishart.h:38-43)
bool npb_prepare_prior(PriorHost &p) {
	const int D = p.D;
	std::vector<double> L((size_t)D * D), A((size_t)D * D), Ainv((size_t)D * D), C((size_t)D * D);
	if (!cholesky_lower(D, p.Lambda.data(), L.data())) return false;
	for (int i = 0; i < D; ++i)
		for (int j = 0; j < D; ++j) {
			double s = 0.0;
			for (int k = 0; k < D; ++k) s += L[k * D + i] * L[k * D + j];
			A[i * D + j] = s;
		}
	double det;
	if (!gauss_jordan_inverse(D, A.data(), Ainv.data(), &det) || !(det > 0.0)) return false;
	for (int i = 0; i < D; ++i)
		for (int j = i + 1; j < D; ++j) Ainv[i * D + j] = Ainv[j * D + i] = 0.5 * (Ainv[i * D + j] + Ainv[j * D + i]);
	if (!cholesky_lower(D, Ainv.data(), C.data())) return false;
	p.CT.assign(npb_tri(D), 0.0);
	for (int i = 0; i < D; ++i)
		for (int j = i; j < D; ++j) p.CT[npb_tri_off(D, i, j)] = C[j * D + i];
	// S = (C^T)^-1, upper triangular: back substitution column by column
	std::vector<double> U((size_t)D * D, 0.0), Sd((size_t)D * D, 0.0);
	for (int i = 0; i < D; ++i)
		for (int j = i; j < D; ++j) U[i * D + j] = C[j * D + i];
	for (int c = 0; c < D; ++c)
		for (int i = c; i >= 0; --i) {
			double s = (i == c) ? 1.0 : 0.0;
			for (int k = i + 1; k <= c; ++k) s -= U[i * D + k] * Sd[k * D + c];
			Sd[i * D + c] = s / U[i * D + i];
		}
	p.S.assign(npb_tri(D), 0.0);
	for (int i = 0; i < D; ++i)
		for (int j = i; j < D; ++j) p.S[npb_tri_off(D, i, j)] = Sd[i * D + j];
	p.logdetA = std::log(det);
	return true;
}

void npb_theta_to_sigma(int D, const double *T, double *Sigma) {
	std::vector<double> P((size_t)D * D, 0.0);
	for (int i = 0; i < D; ++i)
		for (int j = 0; j < D; ++j) {
			double s = 0.0;
			for (int k = 0; k <= (i < j ? i : j); ++k) s += T[npb_tri_off(D, k, i)] * T[npb_tri_off(D, k, j)];
			P[i * D + j] = s;
		}
	double det;
	if (!gauss_jordan_inverse(D, P.data(), Sigma, &det))
		for (int i = 0; i < D * D; ++i) Sigma[i] = NAN;
}
