/*
 * oracle/ref_driver.cpp -- runs the UNMODIFIED reference sampler (mrquincle/noparama) from a binary request file.
 *
 * TEST INFRASTRUCTURE ONLY.  Linked (oracle/Makefile, target _ref) with the reference's own translation units,
 * compiled where they lie under /root/reference against oracle/eigen_shim (Eigen itself is absent from the image).
 * It does what np_main.cpp:320-470 does for `-c clustering` -- build likelihood, NIW prior, Dirichlet process,
 * InitClusters, UpdateClusters, the chosen sampler and MCMC, then MCMC::run -- with three differences that do not touch
 * the path: (1) data, prior and D come from the request instead of the 2-D file reader; (2) the two random_device
 * seeds (np_main.cpp:180, dim1algebra.hpp:2070) are fixed by the request, through oracle/ref_seed_hook.h;
 * (3) the sampler is wrapped in a forwarding decorator that clocks the update() calls (np_mcmc.cpp:146-163) and
 * records the cluster count after each call plus the assignments after each sweep.
 *
 * One process = one run: the reference keeps process-wide static distributions (normal.h:61,
 * multivariatenormal.cpp:41) whose cached state is part of the random stream.
 *
 * usage: np_ref_run ALG T SEED_MAIN SEED_SHUFFLE request.bin result.bin [record_every_sweep [FAMILY]]
 *   ALG: algorithm8 | jain_neal_split | triadic           (np_main.cpp:228-238)
 *   FAMILY: clustering (default) | regression | angular   (np_main.cpp:196-205, `-c`)
 *   request.bin : int32 N, int32 D, double kappa, nu, alpha, mu0[P], Lambda[P*P] (row-major), X[N*D] (row-major); P = D for
 *                 clustering; for regression / angular P = 2 (np_main.cpp:324,358), the rows are what read_data builds
 *                 (np_main.cpp:83-101: (1, a, b) resp. (a, b)) and kappa, nu carry the inverse-gamma's alpha, beta (np_main.cpp:360-361)
 *   result.bin  : int32 N, int32 T, int64 calls, double seconds_run, double seconds_update,
 *                 int32 K_final, int32 n_snap, int32 z_final[N], int32 z_maxlik[N],
 *                 int64 n_K, int32 K_after_call[n_K], int32 z_snap[n_snap][N],
 *                 int64 n_s, double cumulative_update_seconds_at_sweep_end[n_s]  (Alg. 8 only)
 */
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include <np_data.h>
#include <np_suffies.h>
#include <membertrix.h>
#include <np_mcmc.h>
#include <np_init_clusters.h>
#include <np_update_clusters.h>
#include <np_neal_algorithm8.h>
#include <np_jain_neal_algorithm.h>
#include <np_triadic_algorithm.h>
#include <statistics/dirichlet.h>
#include <statistics/multivariatenormal.h>
#include <statistics/normalinvwishart.h>
#include <statistics/normalinvgamma.h>
#include <statistics/scalarnoise_multivariatenormal.h>

unsigned npo_ref_seed_value = 0; /* read by oracle/ref_seed_hook.h */

namespace {

/* labels in the order Results::calculateContingencyMatrix uses (np_results.cpp:17-37): copy => relabel, getClusterId */
void labels_of(const membertrix &m, int N, std::vector<int32_t> &z) {
	membertrix copy(m);
	z.resize(N);
	for (int i = 0; i < N; ++i) z[i] = (int32_t)copy.getClusterId(i);
}

class TimedSampler : public UpdateClusterPopulation {
public:
	UpdateClusterPopulation &inner;
	int N, subset, record;
	double seconds = 0.0;
	int64_t calls = 0;
	std::vector<int32_t> K_after;
	std::vector<double> sweep_seconds; /* Alg. 8: cumulative update() seconds at the end of every sweep */
	std::vector<std::vector<int32_t>> snaps;
	TimedSampler(UpdateClusterPopulation &s, int n, int subset_count, int rec) : inner(s), N(n), subset(subset_count), record(rec) {}
	void update(membertrix &trix, const data_ids_t &ids) override {
		auto t0 = std::chrono::steady_clock::now();
		inner.update(trix, ids);
		seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		++calls;
		if (subset == 1 && calls % N == 0) sweep_seconds.push_back(seconds);
		if (record) {
			K_after.push_back((int32_t)trix.getClusterCount());
			/* Alg. 8 makes exactly N calls per sweep (np_mcmc.cpp:146-163); split-merge sweeps skip collisions (Q11) */
			if (subset == 1 && calls % N == 0) {
				std::vector<int32_t> z(N);
				for (int i = 0; i < N; ++i) z[i] = (int32_t)trix.getClusterId(i);
				snaps.push_back(z);
			}
		}
	}
	void printStatistics() override { inner.printStatistics(); }
};

template <typename T> void rd(FILE *f, T *p, size_t n) {
	if (fread(p, sizeof(T), n, f) != n) { std::fprintf(stderr, "np_ref_run: short request file\n"); std::exit(2); }
}
template <typename T> void wr(FILE *f, const T *p, size_t n) {
	if (n && fwrite(p, sizeof(T), n, f) != n) { std::fprintf(stderr, "np_ref_run: write failed\n"); std::exit(2); }
}

} // namespace

int main(int argc, char **argv) {
	if (argc < 7) {
		std::fprintf(stderr, "usage: %s ALG T SEED_MAIN SEED_SHUFFLE request.bin result.bin [record]\n", argv[0]);
		return 1;
	}
	const std::string alg = argv[1];
	const int T = std::atoi(argv[2]);
	const unsigned seed_main = (unsigned)std::strtoul(argv[3], nullptr, 10);
	npo_ref_seed_value = (unsigned)std::strtoul(argv[4], nullptr, 10);
	const int record = argc > 7 ? std::atoi(argv[7]) : 0;
	const std::string family = argc > 8 ? argv[8] : "clustering";
	if (family != "clustering" && family != "regression" && family != "angular") {
		std::fprintf(stderr, "Unknown family: %s\n", family.c_str());
		return 1;
	}

	FILE *fin = std::fopen(argv[5], "rb");
	if (!fin) { std::perror(argv[5]); return 2; }
	int32_t N, D;
	double kappa, nu, alpha;
	rd(fin, &N, 1); rd(fin, &D, 1); rd(fin, &kappa, 1); rd(fin, &nu, 1); rd(fin, &alpha, 1);
	const int P = family == "clustering" ? D : 2;
	std::vector<double> mu0(P), Lambda((size_t)P * P), X((size_t)N * D);
	rd(fin, mu0.data(), mu0.size()); rd(fin, Lambda.data(), Lambda.size()); rd(fin, X.data(), X.size());
	std::fclose(fin);

	/* the reference logs every step to std::cout (pretty_print.hpp:41-48); keep the run quiet */
	std::ostringstream sink;
	std::streambuf *old_cout = std::cout.rdbuf(sink.rdbuf());

	std::default_random_engine generator(seed_main); /* np_main.cpp:180 with the seed fixed */

	dataset_t dataset;
	for (int i = 0; i < N; ++i) dataset.push_back(new data_t(X.begin() + (size_t)i * D, X.begin() + (size_t)(i + 1) * D));

	Suffies_Dirichlet suffies_dirichlet;
	suffies_dirichlet.alpha = alpha;
	distribution_t *likelihood, *prior;
	if (family == "clustering") {
		/* np_main.cpp:330-334 */
		Suffies_MultivariateNormal *suffies_mvn = new Suffies_MultivariateNormal(D);
		suffies_mvn->mu.setZero();
		suffies_mvn->sigma = Eigen::MatrixXd::Identity(D, D);
		likelihood = new multivariate_normal_distribution(*suffies_mvn);
		/* np_main.cpp:353-372 */
		Suffies_NormalInvWishart *niw = new Suffies_NormalInvWishart(D);
		for (int d = 0; d < D; ++d) niw->mu(d) = mu0[d];
		niw->kappa = kappa;
		niw->nu = nu;
		for (int r = 0; r < D; ++r)
			for (int c = 0; c < D; ++c) niw->Lambda(r, c) = Lambda[(size_t)r * D + c];
		prior = new normal_inverse_wishart_distribution(*niw);
	} else {
		/* np_main.cpp:322-328 */
		Suffies_ScalarNoise_MultivariateNormal *suffies_sn = new Suffies_ScalarNoise_MultivariateNormal(2);
		suffies_sn->mu.setZero();
		suffies_sn->sigma = 1;
		likelihood = new scalarnoise_multivariate_normal_distribution(*suffies_sn, family == "regression" ? regression_mode : angular_mode);
		/* np_main.cpp:357-364 */
		Suffies_NormalInvGamma *nig = new Suffies_NormalInvGamma(2);
		for (int d = 0; d < 2; ++d) nig->mu(d) = mu0[d];
		nig->alpha = kappa;
		nig->beta = nu;
		for (int r = 0; r < 2; ++r)
			for (int c = 0; c < 2; ++c) nig->Lambda(r, c) = Lambda[(size_t)r * 2 + c];
		prior = new normal_inverse_gamma_distribution(*nig);
	}
	dirichlet_process hyper(suffies_dirichlet, *prior);

	InitClusters init_clusters(generator, hyper);           /* np_main.cpp:388 */
	UpdateClusters update_clusters(generator, *likelihood, hyper); /* np_main.cpp:415 */

	int subset_count = 1;
	UpdateClusterPopulation *sampler = nullptr; /* np_main.cpp:424-459 */
	if (alg == "algorithm8") {
		sampler = new NealAlgorithm8(generator, *likelihood, hyper);
		subset_count = 1;
	} else if (alg == "jain_neal_split") {
		sampler = new JainNealAlgorithm(generator, *likelihood, hyper);
		subset_count = 2;
	} else if (alg == "triadic") {
		sampler = new TriadicAlgorithm(generator, *likelihood, hyper);
		subset_count = 3;
	} else {
		std::fprintf(stderr, "Unknown algorithm: %s\n", alg.c_str());
		return 1;
	}
	TimedSampler timed(*sampler, N, subset_count, record);

	MCMC &mcmc = *new MCMC(generator, init_clusters, update_clusters, timed, subset_count, *likelihood); /* np_main.cpp:468 */
	auto t0 = std::chrono::steady_clock::now();
	mcmc.run(dataset, T); /* np_main.cpp:471 */
	double seconds_run = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

	std::vector<int32_t> z_final, z_maxlik;
	labels_of(mcmc.getMembershipMatrix(), N, z_final);
	labels_of(mcmc.getMaxLikelihoodMatrix(), N, z_maxlik);
	int32_t K_final = (int32_t)mcmc.getMembershipMatrix().getClusterCount();
	std::cout.rdbuf(old_cout);

	FILE *fo = std::fopen(argv[6], "wb");
	if (!fo) { std::perror(argv[6]); return 2; }
	int32_t T32 = T, n_snap = (int32_t)timed.snaps.size();
	wr(fo, &N, 1); wr(fo, &T32, 1); wr(fo, &timed.calls, 1); wr(fo, &seconds_run, 1); wr(fo, &timed.seconds, 1);
	wr(fo, &K_final, 1); wr(fo, &n_snap, 1);
	wr(fo, z_final.data(), z_final.size()); wr(fo, z_maxlik.data(), z_maxlik.size());
	int64_t nk = (int64_t)timed.K_after.size();
	wr(fo, &nk, 1);
	wr(fo, timed.K_after.data(), timed.K_after.size());
	for (auto &z : timed.snaps) wr(fo, z.data(), z.size());
	int64_t ns = (int64_t)timed.sweep_seconds.size();
	wr(fo, &ns, 1);
	wr(fo, timed.sweep_seconds.data(), timed.sweep_seconds.size());
	std::fclose(fo);
	std::printf("np_ref_run %s N=%d D=%d T=%d calls=%lld K_final=%d run %.3fs update %.3fs\n", alg.c_str(), N, D, T,
			(long long)timed.calls, K_final, seconds_run, timed.seconds);
	return 0;
}
