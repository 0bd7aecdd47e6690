/*
 * np_oracle.cpp -- CPU oracle for the Gibbs-reassignment hot path of mrquincle/noparama.
 *
 * TEST INFRASTRUCTURE ONLY (see np_oracle.h).  Nothing under noparama_b200/ may
 * link, import or call this file.
 *
 * This is a from-scratch, Eigen-free restatement of the reference algorithm.  Every
 * function cites the reference file:line it follows (paths relative to /root/reference).
 * The random number plumbing deliberately uses the very same libstdc++ <random>
 * objects the reference uses (minstd_rand0 via default_random_engine,
 * std::normal_distribution with process-wide saved state, std::discrete_distribution,
 * std::shuffle on a process-wide mt19937, std::unordered_map iteration order), so that
 * with identical seeds the oracle walks the same trajectory the reference would.
 *
 * The three Eigen routines on the path (PartialPivLU inverse/determinant, LLT and
 * SelfAdjointEigenSolver) are restated from Eigen 3's published algorithms: Eigen is
 * an un-vendored, un-pinned dependency that is absent from this image, so agreement
 * at the last ulp is unpinned; the density is pinned by the reference's own
 * known-answer test (test/test_mvn_likelihood.cpp:33,44).
 */
#include "np_oracle.h"

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <random>
#include <set>
#include <unordered_map>
#include <vector>

namespace {

typedef std::default_random_engine engine_t; /* np_main.cpp:180, == minstd_rand0 in libstdc++ */

/* ------------------------------------------------------------------------------------------------
 * Linear algebra restated from Eigen 3 (dynamic-size MatrixXd code paths).  Row-major storage here;
 * only the order of floating point operations per entry follows Eigen, not its memory layout.
 * ---------------------------------------------------------------------------------------------- */

/* Eigen::PartialPivLU (unblocked path, used for sizes < 16; for larger sizes Eigen blocks the
 * trailing update, which changes only the rounding).  Returns the permutation sign. */
static int lu_decompose(int n, std::vector<double> &a, std::vector<int> &piv) {
	int sign = 1;
	piv.resize(n);
	for (int k = 0; k < n; ++k) {
		int p = k;
		double best = std::fabs(a[k * n + k]);
		for (int i = k + 1; i < n; ++i) {
			double v = std::fabs(a[i * n + k]);
			if (v > best) { best = v; p = i; }
		}
		piv[k] = p;
		if (best != 0.0) {
			if (p != k) {
				for (int j = 0; j < n; ++j) std::swap(a[k * n + j], a[p * n + j]);
				sign = -sign;
			}
			double pivot = a[k * n + k];
			for (int i = k + 1; i < n; ++i) a[i * n + k] /= pivot;
		}
		for (int i = k + 1; i < n; ++i) {
			double l = a[i * n + k];
			for (int j = k + 1; j < n; ++j) a[i * n + j] -= l * a[k * n + j];
		}
	}
	return sign;
}

/* MatrixXd::determinant() -> PartialPivLU::determinant(): sign * prod(diag(U)) */
static double lu_determinant(int n, const double *A) {
	if (n == 0) return 1.0;
	std::vector<double> a(A, A + n * n);
	std::vector<int> piv;
	int sign = lu_decompose(n, a, piv);
	double det = sign;
	for (int i = 0; i < n; ++i) det *= a[i * n + i];
	return det;
}

/* MatrixXd::inverse() -> PartialPivLU::solve(Identity): P, then unit-lower forward, then upper backward */
static void lu_inverse(int n, const double *A, double *Ainv) {
	std::vector<double> a(A, A + n * n);
	std::vector<int> piv;
	lu_decompose(n, a, piv);
	std::vector<double> b(n * n, 0.0);
	for (int i = 0; i < n; ++i) b[i * n + i] = 1.0;
	for (int k = 0; k < n; ++k)
		if (piv[k] != k)
			for (int j = 0; j < n; ++j) std::swap(b[k * n + j], b[piv[k] * n + j]);
	for (int c = 0; c < n; ++c) {
		for (int i = 0; i < n; ++i) {
			double s = b[i * n + c];
			for (int j = 0; j < i; ++j) s -= a[i * n + j] * b[j * n + c];
			b[i * n + c] = s;
		}
		for (int i = n - 1; i >= 0; --i) {
			double s = b[i * n + c];
			for (int j = i + 1; j < n; ++j) s -= a[i * n + j] * b[j * n + c];
			b[i * n + c] = s / a[i * n + i];
		}
	}
	std::memcpy(Ainv, b.data(), sizeof(double) * n * n);
}

/* Lambda.llt().matrixL()  (invwishart.h:40): lower Cholesky factor, row-major */
static void llt_lower(int n, const double *A, double *L) {
	std::fill(L, L + n * n, 0.0);
	for (int j = 0; j < n; ++j) {
		double s = A[j * n + j];
		for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
		double d = std::sqrt(s);
		L[j * n + j] = d;
		for (int i = j + 1; i < n; ++i) {
			double t = A[i * n + j];
			for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
			L[i * n + j] = t / d;
		}
	}
}

/* SelfAdjointEigenSolver (multivariatenormal.cpp:42-44): eigenvalues ascending, eigenvectors in columns.
 * A matrix that is already diagonal needs no rotation in Eigen's tridiagonal QL iteration, so it yields
 * unit eigenvectors sorted by eigenvalue; that case (every prior the reference ships: Lambda = c I) is
 * reproduced exactly.  Non-diagonal input uses cyclic Jacobi: same eigen-pairs, eigenvector signs unpinned. */
static void sym_eigen(int n, const double *A, std::vector<double> &evals, std::vector<double> &evecs) {
	std::vector<double> a(A, A + n * n);
	evecs.assign(n * n, 0.0);
	for (int i = 0; i < n; ++i) evecs[i * n + i] = 1.0;
	bool diagonal = true;
	for (int i = 0; i < n && diagonal; ++i)
		for (int j = 0; j < n; ++j)
			if (i != j && a[i * n + j] != 0.0) { diagonal = false; break; }
	if (!diagonal) {
		for (int sweep = 0; sweep < 64; ++sweep) {
			double off = 0.0;
			for (int i = 0; i < n; ++i)
				for (int j = i + 1; j < n; ++j) off += a[i * n + j] * a[i * n + j];
			if (off < 1e-300) break;
			for (int p = 0; p < n; ++p)
				for (int q = p + 1; q < n; ++q) {
					if (a[p * n + q] == 0.0) continue;
					double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * a[p * n + q]);
					double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
					double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
					for (int k = 0; k < n; ++k) {
						double akp = a[k * n + p], akq = a[k * n + q];
						a[k * n + p] = c * akp - s * akq;
						a[k * n + q] = s * akp + c * akq;
					}
					for (int k = 0; k < n; ++k) {
						double apk = a[p * n + k], aqk = a[q * n + k];
						a[p * n + k] = c * apk - s * aqk;
						a[q * n + k] = s * apk + c * aqk;
					}
					for (int k = 0; k < n; ++k) {
						double vkp = evecs[k * n + p], vkq = evecs[k * n + q];
						evecs[k * n + p] = c * vkp - s * vkq;
						evecs[k * n + q] = s * vkp + c * vkq;
					}
				}
		}
	}
	evals.resize(n);
	for (int i = 0; i < n; ++i) evals[i] = a[i * n + i];
	/* ascending order, stable (Eigen sorts with a selection sort on the eigenvalues) */
	for (int i = 0; i < n - 1; ++i) {
		int k = i;
		for (int j = i + 1; j < n; ++j)
			if (evals[j] < evals[k]) k = j;
		if (k != i) {
			std::swap(evals[i], evals[k]);
			for (int r = 0; r < n; ++r) std::swap(evecs[r * n + i], evecs[r * n + k]);
		}
	}
}

/* ------------------------------------------------------------------------------------------------
 * Parameters ("suffies", np_suffies.h:187-200) and the density (multivariatenormal.cpp)
 * ---------------------------------------------------------------------------------------------- */
struct Theta {
	std::vector<double> mu;    /* [D]   */
	std::vector<double> sigma; /* [D,D] row-major */
	/* cache used only when NPO_PER_CALL_LU is off; computed with the very same routines, so the
	 * numbers are bit-identical to the per-call path */
	mutable bool cached = false;
	mutable std::vector<double> inv;
	mutable double constant = 0.0;
};

struct Counters {
	int64_t density_evals = 0;
};

/* multivariate_normal_distribution::probability / logprobability, clustering branch
 * (multivariatenormal.cpp:83-92 and :125-134).  exponent = ((-0.5 * d^T) * inverse) * d in Eigen's
 * left-to-right evaluation order; constant = sqrt(pow(2 pi, D) * det). */
struct Density {
	bool per_call_lu = true;
	int family = 0; /* 0: multivariate normal; NPO_FAMILY_REGRESSION / NPO_FAMILY_ANGULAR: scalarnoise_multivariatenormal.cpp */
	Counters *cnt = nullptr;
	std::vector<double> inv_tmp, row_tmp, diff_tmp;

	/* scalarnoise_multivariate_normal_distribution::probability / logprobability (scalarnoise_multivariatenormal.cpp:77-150,
	 * :182-250): a one-dimensional normal of standard deviation sigma on a residual -- regression_mode (:86-101, :190-207): the
	 * row is (x_0 .. x_{D-1}, y) and the residual y - x . mu; angular_mode (:120-140, :225-249): the row is a point of the plane,
	 * mu = (d, theta) is a line in normal form, made canonical by prepare() (:31-47) every time the object is bound to a theta,
	 * and the residual |d - q_1| with q = R(theta) x.  The probability divides by sqrt(2 pi sigma^2), the log-probability
	 * subtracts log(2 pi sigma^2) / 2. */
	double scalarnoise_exponent(const Theta &th, const double *x) const {
		const double var = th.sigma[0];
		double diff;
		if (family == NPO_FAMILY_REGRESSION) {
			const int D = (int)th.mu.size();
			double y_proj = 0.0;
			for (int i = 0; i < D; ++i) y_proj += x[i] * th.mu[i];
			diff = x[D] - y_proj;
		} else {
			/* prepare() calls the unqualified abs(): in this translation unit (as in a build against the real Eigen, which
			 * pulls in <cmath> and <cstdlib> but not <math.h>) overload resolution finds only ::abs(int), so both parameters
			 * are TRUNCATED to integers before the absolute value (quirk Q12; verified against oracle/_ref) */
			const double d = (double)std::abs((int)th.mu[0]);
			const double theta = std::fmod((double)std::abs((int)th.mu[1]), 2 * M_PI);
			const double q1 = -std::sin(theta) * x[0] + std::cos(theta) * x[1];
			diff = std::abs(d - q1);
		}
		const double inverse = 1 / (var * var);
		return -0.5 * (diff * diff) * inverse;
	}

	void terms(const Theta &th, const double *x, double &exponent, double &constant) {
		if (family) { /* (callers below never get here for the scalar-noise families) */
			exponent = scalarnoise_exponent(th, x);
			constant = std::sqrt(2 * M_PI * th.sigma[0] * th.sigma[0]);
			if (cnt) cnt->density_evals++;
			return;
		}
		const int D = (int)th.mu.size();
		const double *inv;
		if (per_call_lu) {
			inv_tmp.resize(D * D);
			lu_inverse(D, th.sigma.data(), inv_tmp.data());
			inv = inv_tmp.data();
		} else {
			if (!th.cached) {
				th.inv.resize(D * D);
				lu_inverse(D, th.sigma.data(), th.inv.data());
				th.constant = std::sqrt(std::pow(2 * M_PI, D) * lu_determinant(D, th.sigma.data()));
				th.cached = true;
			}
			inv = th.inv.data();
		}
		diff_tmp.resize(D);
		row_tmp.resize(D);
		for (int i = 0; i < D; ++i) diff_tmp[i] = x[i] - th.mu[i];
		for (int j = 0; j < D; ++j) {
			double s = 0.0;
			for (int i = 0; i < D; ++i) s += (-0.5 * diff_tmp[i]) * inv[i * D + j];
			row_tmp[j] = s;
		}
		double e = 0.0;
		for (int j = 0; j < D; ++j) e += row_tmp[j] * diff_tmp[j];
		exponent = e;
		if (per_call_lu)
			constant = std::sqrt(std::pow(2 * M_PI, D) * lu_determinant(D, th.sigma.data()));
		else
			constant = th.constant;
		if (cnt) cnt->density_evals++;
	}
	double probability(const Theta &th, const double *x) {
		double e, c;
		terms(th, x, e, c);
		return std::exp(e) / c;
	}
	double logprobability(const Theta &th, const double *x) {
		if (family) {
			if (cnt) cnt->density_evals++;
			const double constant = 2 * M_PI * th.sigma[0] * th.sigma[0];
			return scalarnoise_exponent(th, x) - std::log(constant) / 2;
		}
		double e, c;
		terms(th, x, e, c);
		return e - std::log(c);
	}
};

/* ------------------------------------------------------------------------------------------------
 * Process-wide random state of one reference run (Q3, Q4, Q5 of SURVEY 7.4)
 * ---------------------------------------------------------------------------------------------- */
struct Prior {
	int D;
	std::vector<double> mu0, Lambda;
	double kappa, nu, alpha;
	/* scalar-noise families: normal-inverse-gamma base measure (normalinvgamma.h), D = 2 parameters */
	int family = 0;
	double ig_alpha = 1.0, ig_beta = 1.0;
};

struct Process {
	/* engines: main's generator and the COPIES each component keeps (np_init_clusters.h:19,
	 * np_update_clusters.h:17, np_neal_algorithm8.h:20 ...), all taken before the first draw */
	engine_t gen_main, gen_init, gen_upd, gen_pop;
	/* function-static distributions (normal.h:61 and multivariatenormal.cpp:41): one per process,
	 * parameters frozen at first use, saved spare value shared by all callers */
	std::normal_distribution<> static_scalar;
	std::normal_distribution<> static_std;
	/* gamma.h:41: function-static std::gamma_distribution, parameters frozen at the first draw */
	std::gamma_distribution<> static_gamma;
	/* static mt19937 of the 1-argument random_order (dim1algebra.hpp:2066-2073) */
	std::mt19937 mt;

	Process(const Prior &prior, uint32_t seed_main, uint32_t seed_shuffle)
		: gen_main(seed_main), gen_init(gen_main), gen_upd(gen_main), gen_pop(gen_main),
		  static_scalar((double)prior.D, prior.nu) /* mean=D, "variance" passed as stddev: invwishart.h:30-31 */,
		  static_std(0.0, 1.0), static_gamma(prior.ig_alpha, prior.ig_beta), mt(seed_shuffle) {}
};

/* dirichlet_process::sample_base (dirichlet.h:91-93) -> normal_inverse_wishart_distribution::operator()
 * (normalinvwishart.h:44-64) -> inverse_wishart_distribution::operator() (invwishart.h:34-46) ->
 * normal_distribution::operator() (normal.h:57-66) -> multivariate_normal_distribution::operator()
 * (multivariatenormal.cpp:37-50). */
static Theta *sample_base(const Prior &prior, Process &proc, engine_t &gen) {
	const int D = prior.D;
	Theta *th = new Theta();
	th->mu.resize(D);
	if (prior.family) {
		/* normal_inverse_gamma_distribution::operator() (normalinvgamma.h:58-84): sigma^2 = 1 / Gamma(alpha, beta) (gamma.h:37-46,
		 * std::gamma_distribution(shape, scale)), then mu ~ N(mu0, sigma^2 Lambda^-1) through
		 * multivariate_normal_distribution::operator() (multivariatenormal.cpp:37-50) */
		const double val = proc.static_gamma(gen);
		const double sigma_pow2 = 1.0 / val;
		th->sigma.assign(1, std::sqrt(sigma_pow2));
		std::vector<double> inv(D * D), cov(D * D), evals, evecs;
		lu_inverse(D, prior.Lambda.data(), inv.data());
		for (int i = 0; i < D * D; ++i) cov[i] = sigma_pow2 * inv[i];
		sym_eigen(D, cov.data(), evals, evecs);
		std::vector<double> z(D);
		for (int d = 0; d < D; ++d) z[d] = proc.static_std(gen);
		for (int i = 0; i < D; ++i) {
			double acc = 0.0;
			for (int k = 0; k < D; ++k) acc += (evecs[i * D + k] * std::sqrt(evals[k])) * z[k];
			th->mu[i] = prior.mu0[i] + acc;
		}
		return th;
	}
	th->sigma.resize(D * D);
	/* invwishart.h:38-43: v scalar; x = L^T * v ; sigma = x * x^T = v^2 L^T L */
	double v = proc.static_scalar(gen);
	std::vector<double> L(D * D), X(D * D);
	llt_lower(D, prior.Lambda.data(), L.data());
	for (int i = 0; i < D; ++i)
		for (int j = 0; j < D; ++j) X[i * D + j] = L[j * D + i] * v; /* (L^T)(i,j) * v */
	for (int i = 0; i < D; ++i)
		for (int j = 0; j < D; ++j) {
			double s = 0.0;
			for (int k = 0; k < D; ++k) s += X[i * D + k] * X[j * D + k];
			th->sigma[i * D + j] = s;
		}
	/* normalinvwishart.h:56-61: mu ~ N(mu0, sigma / kappa) through the eigen-decomposition */
	std::vector<double> cov(D * D), evals, evecs;
	for (int i = 0; i < D * D; ++i) cov[i] = th->sigma[i] / prior.kappa;
	sym_eigen(D, cov.data(), evals, evecs);
	std::vector<double> z(D);
	for (int d = 0; d < D; ++d) z[d] = proc.static_std(gen); /* unaryExpr, coefficient order */
	for (int i = 0; i < D; ++i) {
		double s = 0.0;
		for (int k = 0; k < D; ++k) s += (evecs[i * D + k] * std::sqrt(evals[k])) * z[k];
		th->mu[i] = prior.mu0[i] + s;
	}
	return th;
}

/* algebra::random_weighted_pick (dim1algebra.hpp:2078-2104) with the uniform made explicit */
static int weighted_pick_u(const double *w, int n, double u01, double *u_scaled_out = nullptr) {
	if (n == 0) return 0;
	std::vector<double> cumsum(n);
	std::partial_sum(w, w + n, cumsum.begin());
	double u = u01 * cumsum.back();
	if (u_scaled_out) *u_scaled_out = u;
	return (int)(std::lower_bound(cumsum.begin(), cumsum.end(), u) - cumsum.begin());
}
static int weighted_pick(const std::vector<double> &w, engine_t &gen, double *u01_out = nullptr) {
	if (w.empty()) return 0;
	std::uniform_real_distribution<> dis;
	double u = dis(gen);
	if (u01_out) *u01_out = u;
	return weighted_pick_u(w.data(), (int)w.size(), u);
}

/* ------------------------------------------------------------------------------------------------
 * membertrix (include/membertrix.h, src/membertrix.cpp)
 * ---------------------------------------------------------------------------------------------- */
struct Cluster {
	Theta *theta; /* shared, never deep-copied (np_cluster.h:27-33; clone shares it, membertrix.cpp:68) */
	int slot;     /* oracle-only: stable storage index used by the GPU replay */
};

enum np_error_t { error_none, error_already_assigned, error_assignment_remaining, error_assignment_absent };

struct Membertrix {
	bool dense; /* keep the N x cols bool matrix (membertrix.h:30): the reference's cost profile */
	int N = 0, cols = 0;
	std::vector<uint8_t> matrix;                          /* column-major N x cols when dense */
	std::vector<int> z;                                   /* cluster id per item, -1 if unassigned */
	std::unordered_map<int, Cluster *> cluster_objects;   /* membertrix.h:58 */
	std::unordered_map<int, std::vector<int> *> clusters_dataset; /* membertrix.h:64, item ids instead of pointers */

	explicit Membertrix(bool dense_) : dense(dense_) {}

	/* membertrix.cpp:124-138 */
	int addData() {
		int idx = N++;
		z.push_back(-1);
		if (dense) resize(N, cols);
		return idx;
	}
	/* membertrix.cpp:87-118 */
	int addCluster(Cluster *c) {
		int id = cols;
		cluster_objects.insert({id, c});
		clusters_dataset.insert({id, new std::vector<int>()});
		cols = id + 1;
		if (dense) resize(N, cols);
		return id;
	}
	int mat_rows = 0, mat_cols = 0;
	void resize(int n, int c) { /* Eigen conservativeResize: allocate + copy the overlap */
		std::vector<uint8_t> m((size_t)n * c, 0);
		int rn = std::min(mat_rows, n), rc = std::min(mat_cols, c);
		for (int j = 0; j < rc; ++j)
			if (rn) std::memcpy(&m[(size_t)j * n], &matrix[(size_t)j * mat_rows], rn);
		matrix.swap(m);
		mat_rows = n;
		mat_cols = c;
	}
	bool row_any(int i) const {
		if (!dense) return z[i] >= 0;
		for (int j = 0; j < cols; ++j)
			if (matrix[(size_t)j * N + i]) return true;
		return false;
	}
	/* membertrix.cpp:147-164 */
	np_error_t assign(int cluster_id, int i) {
		if (row_any(i)) return error_already_assigned;
		if (dense) matrix[(size_t)cluster_id * N + i] = 1;
		z[i] = cluster_id;
		clusters_dataset.at(cluster_id)->push_back(i);
		return error_none;
	}
	bool assigned(int i) const { return row_any(i); }
	/* membertrix.cpp:235-244 */
	int getClusterId(int i) const {
		if (!dense) return z[i];
		for (int j = 0; j < cols; ++j)
			if (matrix[(size_t)j * N + i]) return j;
		return -1;
	}
	/* membertrix.cpp:315-322: column scan, ascending item id, APPENDS to the list it is given */
	void getAssignments(int cluster_id, std::vector<int> &ids) const {
		if (dense) {
			for (int i = 0; i < N; ++i)
				if (matrix[(size_t)cluster_id * N + i]) ids.push_back(i);
		} else {
			for (int i = 0; i < N; ++i)
				if (z[i] == cluster_id) ids.push_back(i);
		}
	}
	bool empty(int id) { return clusters_dataset.at(id)->size() == 0; }
	size_t count(int id) const { return clusters_dataset.at(id)->size(); }
	/* membertrix.cpp:213-228 */
	np_error_t remove(int id) {
		if (!empty(id)) return error_assignment_remaining;
		delete cluster_objects.at(id);
		cluster_objects.erase(id);
		delete clusters_dataset.at(id);
		clusters_dataset.erase(id);
		return error_none;
	}
	/* membertrix.cpp:175-211 */
	np_error_t retract(int cluster_id, int i, bool auto_remove = true) {
		if (!row_any(i)) return error_assignment_absent;
		if (dense) matrix[(size_t)cluster_id * N + i] = 0;
		z[i] = -1;
		std::vector<int> *cl = clusters_dataset.at(cluster_id);
		auto elem = std::find(cl->begin(), cl->end(), i);
		cl->erase(elem);
		if (auto_remove && empty(cluster_id)) remove(cluster_id);
		if (row_any(i)) return error_assignment_remaining;
		return error_none;
	}
	/* membertrix.cpp:230-233 */
	np_error_t retract(int i, bool auto_remove = true) { return retract(getClusterId(i), i, auto_remove); }

	/* membertrix.cpp:343-364: erases from the cluster map only */
	int cleanup() {
		int removed = 0;
		for (auto it = cluster_objects.begin(); it != cluster_objects.end();) {
			if (empty(it->first)) {
				it = cluster_objects.erase(it);
				removed++;
			} else
				++it;
		}
		return removed;
	}
	/* relabel (membertrix.cpp:259-262) == operator=(by value) == compacting copy-constructor
	 * (membertrix.cpp:34-55): clusters re-added in hash-map iteration order, members re-assigned in
	 * ascending item order; the cluster objects themselves are shared. */
	void relabel() {
		Membertrix other(dense);
		for (int i = 0; i < N; ++i) other.addData();
		for (auto pair : cluster_objects) {
			int old_id = pair.first;
			int new_id = other.addCluster(pair.second);
			if (dense) {
				for (int i = 0; i < N; ++i)
					if (matrix[(size_t)old_id * N + i]) other.assign(new_id, i);
			} else {
				for (int i = 0; i < N; ++i)
					if (z[i] == old_id) other.assign(new_id, i);
			}
		}
		for (auto &p : clusters_dataset) delete p.second;
		matrix.swap(other.matrix);
		z.swap(other.z);
		cluster_objects.swap(other.cluster_objects);
		clusters_dataset.swap(other.clusters_dataset);
		cols = other.cols;
		mat_rows = other.mat_rows;
		mat_cols = other.mat_cols;
		other.clusters_dataset.clear();
	}
	~Membertrix() {
		for (auto &p : clusters_dataset) delete p.second;
	}
};

/* ------------------------------------------------------------------------------------------------
 * clustering_performance (clustering_performance.cpp:14-82); int64/double instead of int (Q12)
 * ---------------------------------------------------------------------------------------------- */
static void metrics(const int *A, const int *B, int n, double out3[3]) {
	out3[0] = out3[1] = out3[2] = 0.0;
	if (n <= 0) return;
	int sizeA = *std::max_element(A, A + n) + 1;
	int sizeB = *std::max_element(B, B + n) + 1;
	std::vector<int64_t> f((size_t)sizeA * sizeB, 0), R(sizeA, 0), C(sizeB, 0);
	for (int i = 0; i < n; ++i) f[(size_t)A[i] * sizeB + B[i]]++;
	int64_t N = 0;
	for (int a = 0; a < sizeA; ++a)
		for (int b = 0; b < sizeB; ++b) {
			int64_t v = f[(size_t)a * sizeB + b];
			R[a] += v;
			C[b] += v;
			N += v;
		}
	if (N == 0) return;
	int64_t colmax_sum = 0;
	for (int b = 0; b < sizeB; ++b) {
		int64_t m = 0;
		for (int a = 0; a < sizeA; ++a) m = std::max(m, f[(size_t)a * sizeB + b]);
		colmax_sum += m;
	}
	out3[0] = colmax_sum / (double)N;
	int64_t a = 0, b = 0, c = 0;
	for (auto v : f) a += (v * v - v) / 2;
	for (auto v : R) b += (v * v - v) / 2;
	for (auto v : C) c += (v * v - v) / 2;
	double S = ((double)N * (double)N - (double)N) / 2.0;
	if (S == 0) return;
	out3[1] = (2.0 * a - b - c) / S + 1;
	double bc = (double)b * (double)c / S;
	double bpc = ((double)b + (double)c) / 2.0;
	if (bc == bpc) return;
	out3[2] = ((double)a - bc) / (bpc - bc);
}

} // namespace

/* ================================================================================================
 * The run object
 * ============================================================================================== */
struct npo_run {
	Prior prior;
	npo_options opt;
	int N = 0, D = 0;
	std::vector<double> X;
	npo_stats stats;
	/* final + max-likelihood states */
	std::vector<int> z_final, z_maxlik;
	std::vector<std::vector<double>> final_mu, final_sigma;
	std::vector<int64_t> final_counts;
	/* initial state in slot numbering */
	std::vector<int> init_z, init_slots;
	std::vector<double> init_mu, init_sigma;
	/* trace */
	std::vector<int> tr_item, tr_K, tr_order, tr_picked, tr_new_slot, tr_z_after;
	std::vector<int64_t> tr_order_off;
	std::vector<double> tr_aux_mu, tr_aux_sigma, tr_u;
	std::vector<double> sweep_reassign_seconds, sweep_total_seconds;
	std::vector<int> K_after_call; /* cluster count after every sampler.update() (RECORD_TRACE) */
	/* split-merge replay trace (np_oracle_sm.inc: sm_record), one entry per proposal */
	std::vector<int> sm_picks, sm_type, sm_pool, sm_dec, sm_accept, sm_new_slot;
	std::vector<double> sm_u0, sm_th_mu, sm_th_sigma, sm_us, sm_logA, sm_uacc;
	std::vector<int64_t> sm_pool_off{0};
	/* NOT the reference: initial clusters given by the caller instead of drawn from the prior (bench regime) */
	std::vector<double> given_mu, given_sigma;
	/* slot allocator */
	std::vector<char> slot_used;
	int alloc_slot() {
		for (size_t s = 0; s < slot_used.size(); ++s)
			if (!slot_used[s]) { slot_used[s] = 1; return (int)s; }
		slot_used.push_back(1);
		return (int)slot_used.size() - 1;
	}
};

namespace {

struct Sampler {
	npo_run &run;
	const Prior &prior;
	Process proc;
	Membertrix trix;
	Density density;
	Counters counters;
	const int N, D;
	const double *X;
	double sumK = 0.0;
	bool record;
	bool log_domain;

	Sampler(npo_run &r)
		: run(r), prior(r.prior), proc(r.prior, r.opt.seed_main, r.opt.seed_shuffle),
		  trix((r.opt.flags & NPO_DENSE_MATRIX) != 0), N(r.N), D(r.D), X(r.X.data()),
		  record((r.opt.flags & NPO_RECORD_TRACE) != 0), log_domain((r.opt.flags & NPO_LOG_DOMAIN) != 0) {
		density.per_call_lu = (r.opt.flags & NPO_PER_CALL_LU) != 0;
		density.family = r.prior.family;
		density.cnt = &counters;
	}
	const double *x(int i) const { return X + (size_t)i * D; }

	void free_slot_if_dead(int cluster_id, int slot) {
		if (trix.cluster_objects.find(cluster_id) == trix.cluster_objects.end()) run.slot_used[slot] = 0;
	}

	/* NealAlgorithm8::update (np_neal_algorithm8.cpp:49-167) */
	void alg8_update(int i) {
		const int M = run.opt.M_aux;
		npo_stats &st = run.stats;
		int old_id = trix.getClusterId(i);
		int old_slot = trix.cluster_objects.at(old_id)->slot;
		trix.retract(i); /* :63 */
		free_slot_if_dead(old_id, old_slot);

		auto clusters = trix.cluster_objects; /* :68 -- a COPY of the unordered_map */
		size_t K = clusters.size();
		std::vector<int> cluster_ids(K);

		std::vector<Theta *> aux(M); /* :79-84 */
		for (int m = 0; m < M; ++m) aux[m] = sample_base(prior, proc, proc.gen_pop);

		const double *obs = x(i);
		std::vector<double> w(K + M);
		int k = 0;
		for (auto pair : clusters) { /* :93-109, hash-map iteration order */
			cluster_ids[k] = pair.first;
			const Theta &th = *pair.second->theta;
			if (log_domain) {
				w[k] = density.logprobability(th, obs) + std::log((double)trix.count(pair.first));
			} else {
				if (density.per_call_lu) (void)density.probability(th, obs); /* :105 evaluated twice */
				w[k] = density.probability(th, obs) * trix.count(pair.first); /* :107 */
			}
			k++;
		}
		for (int m = 0; m < M; ++m) { /* :119-126 */
			if (log_domain)
				w[K + m] = density.logprobability(*aux[m], obs) + std::log(prior.alpha / (double)M);
			else
				w[K + m] = density.probability(*aux[m], obs) * prior.alpha / (double)M;
		}
		if (log_domain) { /* NOT the reference: same categorical without underflow */
			double mx = *std::max_element(w.begin(), w.end());
			for (auto &v : w) v = std::exp(v - mx);
		}
		double u01 = 0.0;
		size_t index = weighted_pick(w, proc.gen_pop, &u01); /* :130 */

		if (record) {
			run.tr_item.push_back(i);
			run.tr_K.push_back((int)K);
			for (size_t j = 0; j < K; ++j) run.tr_order.push_back(clusters.at(cluster_ids[j])->slot);
			run.tr_order_off.push_back((int64_t)run.tr_order.size());
			for (int m = 0; m < M; ++m) {
				run.tr_aux_mu.insert(run.tr_aux_mu.end(), aux[m]->mu.begin(), aux[m]->mu.end());
				run.tr_aux_sigma.insert(run.tr_aux_sigma.end(), aux[m]->sigma.begin(), aux[m]->sigma.end());
			}
			run.tr_u.push_back(u01);
			run.tr_picked.push_back((int)index);
		}
		st.candidates += (int64_t)(K + M);
		sumK += (double)K;
		int new_slot = -1;
		if (index >= K) { /* :136-145 */
			Theta *keep = aux[index - K];
			aux[index - K] = nullptr;
			Cluster *c = new Cluster{keep, run.alloc_slot()};
			new_slot = c->slot;
			int id = trix.addCluster(c);
			trix.assign(id, i);
			st.new_cluster_events++;
			st.moved++;
		} else { /* :146-157 */
			int id = cluster_ids[index];
			assert(trix.count(id) != 0);
			trix.assign(id, i);
			if (id != old_id) st.moved++;
		}
		if (record) run.tr_new_slot.push_back(new_slot);
		for (int m = 0; m < M; ++m) delete aux[m]; /* :161-163 (the reference leaks the Suffies) */
		assert(trix.assigned(i));
	}

	/* UpdateClusters::update (np_update_clusters.cpp:71-141).  setSuffies slices (np_cluster.h:49-51, Q1):
	 * theta never changes; the step still draws from gen_upd and the shared static normals. */
	void update_clusters(int mh_steps) {
		const bool eval = density.per_call_lu; /* faithful cost profile: evaluate both dataset sums */
		std::uniform_real_distribution<double> dist(0.0, 1.0);
		for (int t = 0; t < mh_steps; ++t) {
			for (auto pair : trix.cluster_objects) {
				int key = pair.first;
				if (trix.empty(key)) continue;
				double lik = 1.0, plik = 1.0;
				std::vector<int> *ds = trix.clusters_dataset.at(key);
				if (eval) {
					lik = 0.0;
					for (int i : *ds) lik += density.logprobability(*pair.second->theta, x(i));
				}
				Theta *prop = sample_base(prior, proc, proc.gen_upd);
				if (eval) {
					plik = 0.0;
					for (int i : *ds) plik += density.logprobability(*prop, x(i));
				}
				if (!lik) { delete prop; continue; }
				double alpha = std::exp(plik - lik);
				double reject = dist(proc.gen_upd);
				(void)(reject < alpha); /* accept/reject both leave theta unchanged (Q1) */
				delete prop;
			}
		}
	}

	/* MCMC::considerMaxLikelihood (np_mcmc.cpp:187-203) */
	void consider_max_likelihood() {
		double cur = 0.0;
		for (auto pair : trix.cluster_objects) {
			std::vector<int> *ds = trix.clusters_dataset.at(pair.first);
			double s = 0.0;
			for (int i : *ds) s += density.logprobability(*pair.second->theta, x(i));
			cur += s;
		}
		if (cur > run.stats.max_loglik) {
			run.stats.max_loglik = cur;
			snapshot(run.z_maxlik);
		}
	}
	/* compact labels in hash-map iteration order, like Results::calculateContingencyMatrix
	 * (np_results.cpp:17-37: copy-assign => relabel, then getClusterId per item) */
	void snapshot(std::vector<int> &zout) {
		zout.assign(N, -1);
		int label = 0;
		for (auto pair : trix.cluster_objects) {
			for (int i : *trix.clusters_dataset.at(pair.first)) zout[i] = label;
			label++;
		}
	}

	void record_z_after() {
		for (int i = 0; i < N; ++i) run.tr_z_after.push_back(trix.cluster_objects.at(trix.z[i])->slot);
	}

	/* split/merge samplers, np_oracle_sm.inc */
	void sm_update(const std::vector<int> &subset);
	double logprob_items(const Theta &th, const std::vector<int> &items);
	bool jn_split(int data_i, int data_j, int cur_id);
	bool jn_merge(int id0, int id1);
	void jn_update(const std::vector<int> &ids);
	void tri_allocate(std::vector<std::vector<int>> &pdata, const std::vector<int> &picks, int Q,
			const std::vector<int> &source_ids, const std::vector<const Theta *> &target);
	bool tri_split(const std::vector<int> &picks, std::vector<int> &cluster_ids);
	bool tri_merge(const std::vector<int> &picks, std::vector<int> &cluster_ids);
	void tri_update(const std::vector<int> &picks);
	void sm_record(int type, const Theta *th_new, const std::vector<int> &pool, const std::vector<double> &us,
			const std::vector<int> &dec, double logA, double u, bool accept, int new_slot);
	std::vector<int> tri_pool, tri_dec;
	std::vector<double> tri_us;

	/* MCMC::run (np_mcmc.cpp:48-175) */
	void mcmc_run() {
		npo_stats &st = run.stats;
		const int K0 = run.opt.K0, T = run.opt.T;
		auto t_start = std::chrono::steady_clock::now();
		for (int i = 0; i < N; ++i) trix.addData(); /* :58-63 */
		for (int k = 0; k < K0; ++k) {              /* :66, np_init_clusters.cpp:24-41 */
			Theta *th;
			if (!run.given_mu.empty()) { /* caller-supplied parameters (npo_mcmc_run_given) */
				th = new Theta();
				th->mu.assign(run.given_mu.begin() + (size_t)k * D, run.given_mu.begin() + (size_t)(k + 1) * D);
				th->sigma.assign(run.given_sigma.begin() + (size_t)k * D * D, run.given_sigma.begin() + (size_t)(k + 1) * D * D);
			} else {
				th = sample_base(prior, proc, proc.gen_init);
			}
			trix.addCluster(new Cluster{th, -1});
		}
		std::vector<double> weights(K0, 1 / (double)K0); /* :69-73 */
		std::discrete_distribution<int> distribution(weights.begin(), weights.end());
		for (int i = 0; i < N; ++i) trix.assign(distribution(proc.gen_main), i); /* :76-85 */
		trix.cleanup(); /* :89-91 */
		trix.cleanup();
		/* oracle-only: stable slots for the survivors, ascending cluster id */
		{
			std::vector<int> ids;
			for (auto &p : trix.cluster_objects) ids.push_back(p.first);
			std::sort(ids.begin(), ids.end());
			for (int id : ids) trix.cluster_objects.at(id)->slot = run.alloc_slot();
			run.init_z.resize(N);
			for (int i = 0; i < N; ++i) run.init_z[i] = trix.cluster_objects.at(trix.z[i])->slot;
			for (int id : ids) {
				Cluster *c = trix.cluster_objects.at(id);
				run.init_slots.push_back(c->slot);
				run.init_mu.insert(run.init_mu.end(), c->theta->mu.begin(), c->theta->mu.end());
				run.init_sigma.insert(run.init_sigma.end(), c->theta->sigma.begin(), c->theta->sigma.end());
			}
		}
		const int subset_count = run.opt.algorithm == NPO_ALG8 ? 1 : (run.opt.algorithm == NPO_JAIN_NEAL ? 2 : 3);
		const int Msteps = N; /* :94-99 */
		double reassign_seconds = 0.0;
		if (record) run.tr_order_off.push_back(0);
		for (int t = 0; t < T; ++t) { /* :109 */
			if (t % 10 == 0) trix.relabel(); /* :111-114 */
			std::vector<std::vector<int>> indices(subset_count); /* :120-125 */
			for (int j = 0; j < subset_count; ++j) {
				indices[j].resize(N);
				std::iota(indices[j].begin(), indices[j].end(), 0);
				std::shuffle(indices[j].begin(), indices[j].end(), proc.mt);
			}
			auto t0 = std::chrono::steady_clock::now();
			for (int i = 0; i < Msteps; ++i) { /* :146-163 */
				std::vector<int> subset(subset_count);
				std::set<int> uniq;
				for (int j = 0; j < subset_count; ++j) {
					subset[j] = indices[j][i];
					uniq.insert(subset[j]);
				}
				if ((int)uniq.size() != subset_count) continue; /* Q11 */
				if (run.opt.algorithm == NPO_ALG8)
					alg8_update(subset[0]);
				else
					sm_update(subset);
				st.updates++;
				if (record) run.K_after_call.push_back((int)trix.cluster_objects.size());
			}
			{
				double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
				reassign_seconds += dt;
				run.sweep_reassign_seconds.push_back(dt);
			}
			if (record) record_z_after();
			if (run.opt.flags & NPO_UPDATE_CLUSTERS) update_clusters(run.opt.mh_steps); /* :170 */
			if ((run.opt.flags & NPO_MAX_LIKELIHOOD) && t % 5 == 0) consider_max_likelihood(); /* :172-174 */
			run.sweep_total_seconds.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
		}
		st.seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
		st.seconds_reassign = reassign_seconds;
		st.density_evals = counters.density_evals;
		st.mean_K = st.updates ? sumK / (double)st.updates : 0.0;
		st.K_final = (int)trix.cluster_objects.size();
		snapshot(run.z_final);
		if (run.z_maxlik.empty()) run.z_maxlik = run.z_final;
		for (auto pair : trix.cluster_objects) {
			run.final_mu.push_back(pair.second->theta->mu);
			run.final_sigma.push_back(pair.second->theta->sigma);
			run.final_counts.push_back((int64_t)trix.count(pair.first));
		}
	}
};

#include "np_oracle_sm.inc"

} // namespace

/* ================================================================================================
 * C interface
 * ============================================================================================== */
extern "C" {

static Theta make_theta(int D, const double *mu, const double *Sigma) {
	Theta th;
	th.mu.assign(mu, mu + D);
	th.sigma.assign(Sigma, Sigma + D * D);
	return th;
}

double npo_mvn_pdf(int D, const double *mu, const double *Sigma, const double *x) {
	Density d;
	Theta th = make_theta(D, mu, Sigma);
	return d.probability(th, x);
}
double npo_mvn_logpdf(int D, const double *mu, const double *Sigma, const double *x) {
	Density d;
	Theta th = make_theta(D, mu, Sigma);
	return d.logprobability(th, x);
}
/* multivariatenormal.cpp:96-104 */
double npo_mvn_pdf_dataset(int D, const double *mu, const double *Sigma, const double *X, int n) {
	Density d;
	Theta th = make_theta(D, mu, Sigma);
	double r = 1.0;
	for (int i = 0; i < n; ++i) r *= d.probability(th, X + (size_t)i * D);
	return r;
}
/* multivariatenormal.cpp:138-146 */
double npo_mvn_logpdf_dataset(int D, const double *mu, const double *Sigma, const double *X, int n) {
	Density d;
	Theta th = make_theta(D, mu, Sigma);
	double r = 0.0;
	for (int i = 0; i < n; ++i) r += d.logprobability(th, X + (size_t)i * D);
	return r;
}
void npo_mvn_logpdf_batch(int D, const double *mu, const double *Sigma, int K, const double *X, int n, double *out) {
	Density d;
	d.per_call_lu = false;
	for (int k = 0; k < K; ++k) {
		Theta th = make_theta(D, mu + (size_t)k * D, Sigma + (size_t)k * D * D);
		for (int i = 0; i < n; ++i) out[(size_t)i * K + k] = d.logprobability(th, X + (size_t)i * D);
	}
}

int npo_weighted_pick_u(const double *w, int n, double u) { return weighted_pick_u(w, n, u); }

/* test/test_weighted_vector.cpp:10-40 */
void npo_weighted_pick_freq(const double *w, int n, int draws, uint32_t seed, int64_t *freq) {
	engine_t gen(seed);
	std::vector<double> wv(w, w + n);
	for (int i = 0; i < n; ++i) freq[i] = 0;
	for (int t = 0; t < draws; ++t) {
		int j = weighted_pick(wv, gen);
		if (j >= 0 && j < n) freq[j]++;
	}
}

static Prior make_prior(const npo_prior *p) {
	Prior pr;
	pr.D = p->D;
	pr.mu0.assign(p->mu0, p->mu0 + p->D);
	pr.Lambda.assign(p->Lambda, p->Lambda + p->D * p->D);
	pr.kappa = p->kappa;
	pr.nu = p->nu;
	pr.alpha = p->alpha;
	return pr;
}

void npo_sample_base(const npo_prior *prior, uint32_t seed, int count, double *mu_out, double *Sigma_out) {
	Prior pr = make_prior(prior);
	Process proc(pr, seed, 0);
	const int D = pr.D;
	for (int c = 0; c < count; ++c) {
		Theta *th = sample_base(pr, proc, proc.gen_main);
		std::memcpy(mu_out + (size_t)c * D, th->mu.data(), sizeof(double) * D);
		std::memcpy(Sigma_out + (size_t)c * D * D, th->sigma.data(), sizeof(double) * D * D);
		delete th;
	}
}

double npo_lu_determinant(int n, const double *A) { return lu_determinant(n, A); }
void npo_lu_inverse(int n, const double *A, double *Ainv) { lu_inverse(n, A, Ainv); }

void npo_metrics(const int *truth, const int *result, int N, double out3[3]) { metrics(truth, result, N, out3); }

/* test/test_membertrix.cpp:20-96: 5 points, 4 clusters, random assignment; after retracting every member of
 * a random cluster l the cluster count must drop by exactly (l had members ? 1 : 0).  Returns 0 on success. */
int npo_membertrix_selftest(uint32_t seed, int dense) {
	engine_t generator(seed);
	Membertrix trix(dense != 0);
	const int N = 5, K = 4;
	for (int i = 0; i < N; ++i)
		if (trix.addData() != i) return 1;
	std::vector<Cluster *> owned;
	for (int k = 0; k < K; ++k) {
		Cluster *c = new Cluster{new Theta(), k};
		if (trix.addCluster(c) != k) return 2;
	}
	std::uniform_int_distribution<int> distribution(0, K - 1);
	std::vector<int> assigned_to(N);
	for (int i = 0; i < N; ++i) {
		int k = distribution(generator);
		assigned_to[i] = k;
		if (trix.assign(k, i) != error_none) return 3;
		if (trix.assign(k, i) != error_already_assigned) return 4;
	}
	int l = distribution(generator);
	int members = (int)trix.count(l);
	int before = (int)trix.cluster_objects.size();
	for (int i = 0; i < N; ++i)
		if (assigned_to[i] == l)
			if (trix.retract(i) != error_none) return 5;
	int after = (int)trix.cluster_objects.size();
	if (before - after != (members > 0 ? 1 : 0)) return 6;
	for (int i = 0; i < N; ++i)
		if (assigned_to[i] == l && trix.retract(i) != error_assignment_absent) return 7;
	trix.relabel();
	if ((int)trix.cluster_objects.size() != after) return 8;
	for (int i = 0; i < N; ++i)
		if (assigned_to[i] != l && !trix.assigned(i)) return 9;
	return 0;
}

npo_run *npo_mcmc_run(const npo_prior *prior, const npo_options *opt, const double *X, int N) {
	npo_run *r = new npo_run();
	r->prior = make_prior(prior);
	r->opt = *opt;
	r->N = N;
	r->D = prior->D;
	r->X.assign(X, X + (size_t)N * prior->D);
	std::memset(&r->stats, 0, sizeof(r->stats));
	r->stats.max_loglik = -std::numeric_limits<double>::infinity();
	Sampler s(*r);
	s.mcmc_run();
	return r;
}
/* NOT the reference: the same run started from K0 = K caller-supplied clusters (mu [K,D], Sigma [K,D,D]) instead of K0
 * prior draws -- the regime bench.py measures at D = 16, where the reference's own prior collapses to one cluster */
npo_run *npo_mcmc_run_given(const npo_prior *prior, const npo_options *opt, const double *X, int N, int K, const double *mu,
		const double *Sigma) {
	npo_run *r = new npo_run();
	r->prior = make_prior(prior);
	r->opt = *opt;
	r->opt.K0 = K;
	r->N = N;
	r->D = prior->D;
	r->X.assign(X, X + (size_t)N * prior->D);
	r->given_mu.assign(mu, mu + (size_t)K * prior->D);
	r->given_sigma.assign(Sigma, Sigma + (size_t)K * prior->D * prior->D);
	memset(&r->stats, 0, sizeof(r->stats));
	r->stats.max_loglik = -std::numeric_limits<double>::infinity();
	Sampler s(*r);
	s.mcmc_run();
	return r;
}
/* the same run with the scalar-noise likelihood and the normal-inverse-gamma base measure of `-c regression` / `-c angular`
 * (np_main.cpp:322-328, :357-364): X [N, row] with row = 3 (1, a, b) for regression, 2 for angular (np_main.cpp:83-101) */
static Prior make_nig_prior(int family, const double *mu0, const double *Lambda, double ig_alpha, double ig_beta, double dp_alpha) {
	Prior pr;
	pr.D = 2;
	pr.mu0.assign(mu0, mu0 + 2);
	pr.Lambda.assign(Lambda, Lambda + 4);
	pr.kappa = pr.nu = 1.0;
	pr.alpha = dp_alpha;
	pr.family = family;
	pr.ig_alpha = ig_alpha;
	pr.ig_beta = ig_beta;
	return pr;
}
npo_run *npo_mcmc_run_scalarnoise(int family, const double *mu0, const double *Lambda, double ig_alpha, double ig_beta, double dp_alpha,
		const npo_options *opt, const double *X, int N) {
	if (family != NPO_FAMILY_REGRESSION && family != NPO_FAMILY_ANGULAR) return nullptr;
	npo_run *r = new npo_run();
	r->prior = make_nig_prior(family, mu0, Lambda, ig_alpha, ig_beta, dp_alpha);
	r->opt = *opt;
	r->opt.flags &= ~NPO_RECORD_TRACE; /* the replay traces are laid out for (mu [D], Sigma [D, D]) */
	r->N = N;
	r->D = family == NPO_FAMILY_REGRESSION ? 3 : 2; /* row width of X */
	r->X.assign(X, X + (size_t)N * r->D);
	std::memset(&r->stats, 0, sizeof(r->stats));
	r->stats.max_loglik = -std::numeric_limits<double>::infinity();
	Sampler s(*r);
	s.mcmc_run();
	return r;
}
double npo_scalarnoise_logpdf(int family, const double *mu, double sigma, const double *x) {
	Density d;
	d.family = family;
	Theta th;
	th.mu.assign(mu, mu + 2);
	th.sigma.assign(1, sigma);
	return d.logprobability(th, x);
}
double npo_scalarnoise_pdf(int family, const double *mu, double sigma, const double *x) {
	Density d;
	d.family = family;
	Theta th;
	th.mu.assign(mu, mu + 2);
	th.sigma.assign(1, sigma);
	return d.probability(th, x);
}
void npo_sample_base_nig(const double *mu0, const double *Lambda, double ig_alpha, double ig_beta, uint32_t seed, int count, double *mu_out,
		double *sigma_out) {
	Prior pr = make_nig_prior(NPO_FAMILY_REGRESSION, mu0, Lambda, ig_alpha, ig_beta, 1.0);
	Process proc(pr, seed, 0);
	for (int c = 0; c < count; ++c) {
		Theta *th = sample_base(pr, proc, proc.gen_main);
		mu_out[2 * c] = th->mu[0];
		mu_out[2 * c + 1] = th->mu[1];
		sigma_out[c] = th->sigma[0];
		delete th;
	}
}
/* final clusters of a scalar-noise run: mu [K, 2], sigma [K] */
int npo_run_params_scalarnoise(const npo_run *r, int *K, double *mu, double *sigma, int64_t *counts, int cap) {
	int k = (int)r->final_mu.size();
	*K = k;
	if (k > cap || !r->prior.family) return -1;
	for (int j = 0; j < k; ++j) {
		mu[2 * j] = r->final_mu[j][0];
		mu[2 * j + 1] = r->final_mu[j][1];
		sigma[j] = r->final_sigma[j][0];
		counts[j] = r->final_counts[j];
	}
	return 0;
}
void npo_run_free(npo_run *r) { delete r; }
void npo_run_stats(const npo_run *r, npo_stats *out) { *out = r->stats; }
void npo_run_assignments(const npo_run *r, int which, int *z_out) {
	const std::vector<int> &z = which ? r->z_maxlik : r->z_final;
	std::memcpy(z_out, z.data(), sizeof(int) * z.size());
}
int npo_run_params(const npo_run *r, int *K, double *mu, double *Sigma, int64_t *counts, int cap) {
	int k = (int)r->final_mu.size();
	*K = k;
	if (k > cap) return -1;
	const int D = r->D;
	for (int j = 0; j < k; ++j) {
		std::memcpy(mu + (size_t)j * D, r->final_mu[j].data(), sizeof(double) * D);
		std::memcpy(Sigma + (size_t)j * D * D, r->final_sigma[j].data(), sizeof(double) * D * D);
		counts[j] = r->final_counts[j];
	}
	return 0;
}
int npo_run_init_K(const npo_run *r) { return (int)r->init_slots.size(); }
void npo_run_init_state(const npo_run *r, int *z0, int *slots, double *mu, double *Sigma) {
	std::memcpy(z0, r->init_z.data(), sizeof(int) * r->init_z.size());
	std::memcpy(slots, r->init_slots.data(), sizeof(int) * r->init_slots.size());
	std::memcpy(mu, r->init_mu.data(), sizeof(double) * r->init_mu.size());
	std::memcpy(Sigma, r->init_sigma.data(), sizeof(double) * r->init_sigma.size());
}
void npo_run_sweep_seconds(const npo_run *r, double *reassign, double *total) {
	std::memcpy(reassign, r->sweep_reassign_seconds.data(), sizeof(double) * r->sweep_reassign_seconds.size());
	std::memcpy(total, r->sweep_total_seconds.data(), sizeof(double) * r->sweep_total_seconds.size());
}
int64_t npo_run_K_after_len(const npo_run *r) { return (int64_t)r->K_after_call.size(); }
void npo_run_K_after(const npo_run *r, int *out) { std::copy(r->K_after_call.begin(), r->K_after_call.end(), out); }
int64_t npo_sm_trace_proposals(const npo_run *r) { return (int64_t)r->sm_type.size(); }
int64_t npo_sm_trace_pool_len(const npo_run *r) { return (int64_t)r->sm_pool.size(); }
void npo_sm_trace_copy(const npo_run *r, int *picks, double *u0, int *type, double *th_mu, double *th_sigma, int64_t *pool_off,
		int *pool, double *us, int *dec, double *logA, double *uacc, int *accept, int *new_slot) {
	std::copy(r->sm_picks.begin(), r->sm_picks.end(), picks);
	std::copy(r->sm_u0.begin(), r->sm_u0.end(), u0);
	std::copy(r->sm_type.begin(), r->sm_type.end(), type);
	std::copy(r->sm_th_mu.begin(), r->sm_th_mu.end(), th_mu);
	std::copy(r->sm_th_sigma.begin(), r->sm_th_sigma.end(), th_sigma);
	std::copy(r->sm_pool_off.begin(), r->sm_pool_off.end(), pool_off);
	std::copy(r->sm_pool.begin(), r->sm_pool.end(), pool);
	std::copy(r->sm_us.begin(), r->sm_us.end(), us);
	std::copy(r->sm_dec.begin(), r->sm_dec.end(), dec);
	std::copy(r->sm_logA.begin(), r->sm_logA.end(), logA);
	std::copy(r->sm_uacc.begin(), r->sm_uacc.end(), uacc);
	std::copy(r->sm_accept.begin(), r->sm_accept.end(), accept);
	std::copy(r->sm_new_slot.begin(), r->sm_new_slot.end(), new_slot);
}
int64_t npo_trace_steps(const npo_run *r) { return (int64_t)r->tr_item.size(); }
int64_t npo_trace_order_len(const npo_run *r) { return (int64_t)r->tr_order.size(); }
int npo_trace_max_slot(const npo_run *r) { return (int)r->slot_used.size(); }
void npo_trace_copy(const npo_run *r, int *item, int *K, int64_t *order_off, int *order, double *aux_mu,
		double *aux_Sigma, double *u, int *picked, int *new_slot, int *z_after) {
#define CP(dst, src) if (dst) std::memcpy(dst, (src).data(), sizeof((src)[0]) * (src).size())
	CP(item, r->tr_item);
	CP(K, r->tr_K);
	CP(order_off, r->tr_order_off);
	CP(order, r->tr_order);
	CP(aux_mu, r->tr_aux_mu);
	CP(aux_Sigma, r->tr_aux_sigma);
	CP(u, r->tr_u);
	CP(picked, r->tr_picked);
	CP(new_slot, r->tr_new_slot);
	CP(z_after, r->tr_z_after);
#undef CP
}

} /* extern "C" */

#include "np_oracle_alg2.inc"
