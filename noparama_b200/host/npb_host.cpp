// npb_host.cpp -- see npb_host.h.  Everything here is plumbing above the C ABI; no arithmetic of the path lives on the host.
#include "npb_host.h"

#include <filesystem>
#include <fstream>
#include <map>

#include <algorithm>
#include <cassert>
#include <iostream>

namespace npb {

device::device(int index) {
	npb_status s = npb_ctx_create(index, &ctx_);
	if (s != NPB_OK) throw npb_error(s, std::string("npb200: no usable CUDA device (") + npb_status_str(s) + "); there is no CPU path");
}
device::~device() { npb_ctx_destroy(ctx_); }
void device::check(npb_status s) const {
	if (s == NPB_OK) return;
	std::string msg = npb_status_str(s);
	const char *detail = npb_ctx_last_error(ctx_);
	if (detail && *detail) msg += std::string(": ") + detail;
	throw npb_error(s, msg);
}

membertrix::membertrix(device &dev, dataset_t &dataset, int D) : dev_(dev), dataset_(&dataset), N_((int)dataset.size()), D_(D) {
	// addData for every item (membertrix.cpp:124-138): the rows go to the device once
	std::vector<double> X((size_t)N_ * D_);
	for (int i = 0; i < N_; ++i) {
		assert((int)dataset[i]->size() >= D_);
		std::copy(dataset[i]->begin(), dataset[i]->begin() + D_, X.begin() + (size_t)i * D_);
	}
	dev_.check(npb_dataset_upload(dev_.ctx(), X.data(), N_, D_, &ds_));
}
membertrix::~membertrix() { if (!detached_) npb_dataset_destroy(ds_); }

void membertrix::refresh() {
	if (!dirty_ || !chains || detached_) return;
	const int cap = npb_chains_kmax(chains);
	z_.resize(N_);
	slots_.resize(cap);
	counts_.resize(cap);
	mu_.resize((size_t)cap * D_);
	sigma_.resize((size_t)cap * D_ * D_);
	dev_.check(npb_chains_get_assignments(chains, chain_, 1, z_.data()));
	int K = 0;
	dev_.check(npb_chains_get_params(chains, chain_, cap, &K, slots_.data(), counts_.data(), mu_.data(), sigma_.data()));
	slots_.resize(K);
	counts_.resize(K);
	dirty_ = false;
}
cluster_id_t membertrix::getClusterId(data_id_t i) {
	refresh();
	if (i < 0 || i >= N_ || pending_.count(i)) return -1;
	return z_[i];
}

// ---- mutators: the reference's single-item interface (membertrix.cpp:87-233) over the device state ----
cluster_id_t membertrix::addCluster(const Suffies_MultivariateNormal &suffies) {
	// the reference hands out the next column index (membertrix.cpp:108-117); here a host-side id beyond the device's slot range
	// until the cluster gets its first member, at which point it becomes the lowest free device slot
	const cluster_id_t id = npb_chains_kmax(chains) + (cluster_id_t)added_.size() + (added_.empty() ? 0 : added_.rbegin()->first - npb_chains_kmax(chains));
	added_[id] = suffies;
	return id;
}
np_error_t membertrix::retract(data_id_t i, bool auto_remove) {
	refresh();
	if (i < 0 || i >= N_ || pending_.count(i)) return error_assignment_absent; // membertrix.cpp:177-180
	(void)auto_remove; // the device drops a cluster with its last member (membertrix.cpp:200-203); the host copy follows on refresh
	pending_[i] = z_[i];
	return error_none;
}
np_error_t membertrix::assign(cluster_id_t k, data_id_t i) {
	refresh();
	if (i < 0 || i >= N_) return error_assignment_absent;
	auto pend = pending_.find(i);
	if (pend == pending_.end()) return error_already_assigned; // membertrix.cpp:148-151: retract first
	npb_status s;
	auto add = added_.find(k);
	if (add != added_.end()) {
		int slot = -1;
		s = npb_chain_move_item_new(chains, chain_, i, add->second.mu.data(), add->second.sigma.data(), &slot);
		if (s == NPB_OK) added_.erase(add);
	} else if (k == pend->second) {
		s = NPB_OK; // back where it was: nothing moved on the device
	} else {
		s = npb_chain_move_item(chains, chain_, i, k);
	}
	if (s == NPB_E_ASSIGNMENT_ABSENT) return error_assignment_absent;
	if (s == NPB_E_ALREADY_ASSIGNED) return error_already_assigned;
	dev_.check(s);
	pending_.erase(pend);
	dirty_ = true;
	return error_none;
}
np_error_t membertrix::remove(cluster_id_t k) {
	if (added_.erase(k)) return error_none;
	const npb_status s = npb_chain_remove_cluster(chains, chain_, k);
	if (s == NPB_E_ASSIGNMENT_REMAINING) return error_assignment_remaining; // membertrix.cpp:214-216
	dev_.check(s);
	return error_none;
}
np_error_t membertrix::cleanup() {
	added_.clear(); // clusters without members; on the device a slot without members is free already
	return error_none;
}
membertrix *membertrix::clone() {
	refresh();
	membertrix *c = new membertrix(*this);
	c->detached_ = true; // a copy of the host view: it owns no device handle and never refreshes
	c->chains = nullptr;
	c->ds_ = nullptr;
	return c;
}
size_t membertrix::getClusterCount() {
	refresh();
	return slots_.size();
}
std::map<cluster_id_t, Suffies_MultivariateNormal> membertrix::getClusters() {
	refresh();
	std::map<cluster_id_t, Suffies_MultivariateNormal> out;
	for (size_t k = 0; k < slots_.size(); ++k) {
		Suffies_MultivariateNormal s(D_);
		std::copy(mu_.begin() + k * D_, mu_.begin() + (k + 1) * D_, s.mu.begin());
		std::copy(sigma_.begin() + k * D_ * D_, sigma_.begin() + (k + 1) * D_ * D_, s.sigma.begin());
		out[slots_[k]] = s;
	}
	return out;
}
size_t membertrix::count(cluster_id_t k) {
	refresh();
	for (size_t j = 0; j < slots_.size(); ++j)
		if (slots_[j] == k) return (size_t)counts_[j];
	return 0;
}
void membertrix::getAssignments(cluster_id_t k, data_ids_t &ids) {
	refresh();
	for (int i = 0; i < N_; ++i)
		if (z_[i] == k) ids.push_back(i);
}

void NealAlgorithm8::sweep(membertrix &cluster_matrix, int n_sweeps) {
	npb_sweep_stats st{};
	dev_.check(npb_chains_sweep(cluster_matrix.chains, NPB_ALG8, n_sweeps, &st));
	record(st);
	cluster_matrix.invalidate();
}
void NealAlgorithm2::sweep(membertrix &cluster_matrix, int n_sweeps) {
	npb_sweep_stats st{};
	dev().check(npb_chains_sweep(cluster_matrix.chains, NPB_ALG2, n_sweeps, &st));
	record(st);
	cluster_matrix.invalidate();
}
void NealAlgorithm2Conjugate::sweep(membertrix &cluster_matrix, int n_sweeps) {
	npb_sweep_stats st{};
	dev().check(npb_chains_sweep(cluster_matrix.chains, NPB_ALG2_CONJUGATE, n_sweeps, &st));
	record(st);
	cluster_matrix.invalidate();
}
void NealAlgorithm8::update(membertrix &cluster_matrix, const data_ids_t &data_ids) {
	assert(data_ids.size() == 1); // np_neal_algorithm8.cpp:54
	if (calls_++ % cluster_matrix.size() == 0) sweep(cluster_matrix, 1);
}
void NealAlgorithm8::printStatistics() { // np_neal_algorithm8.cpp:169-176
	std::cout << "Statistics:" << std::endl;
	std::cout << " # of new cluster events accepted: " << accepted_ << std::endl;
	std::cout << " # of new cluster events rejected: " << rejected_ << std::endl;
}

void SplitMergeAlgorithm::sweep(membertrix &cluster_matrix, int n_sweeps) {
	dev_.check(npb_chains_sweep(cluster_matrix.chains, sampler_, n_sweeps, &last_));
	for (int i = 0; i < 4; ++i) {
		attempts_[i] += last_.sm_attempts[i];
		accepts_[i] += last_.sm_accepts[i];
	}
	cluster_matrix.invalidate();
}
void SplitMergeAlgorithm::update(membertrix &cluster_matrix, const data_ids_t &data_ids) {
	assert((int)data_ids.size() == subsets_); // np_jain_neal_algorithm.cpp:429, np_triadic_algorithm.cpp:647
	if (calls_++ % cluster_matrix.size() == 0) sweep(cluster_matrix, 1);
}
void SplitMergeAlgorithm::printStatistics() {
	static const char *type[4] = {"merge from 2 to 1", "split from 1 to 2", "merge from 3 to 2", "split from 2 to 3"};
	std::cout << "Statistics:" << std::endl;
	for (int i = 0; i < (sampler_ == NPB_JAIN_NEAL ? 2 : 4); ++i) {
		std::cout << " # of " << type[i] << " attempts: " << attempts_[i] << std::endl;
		std::cout << "   o of accepted " << type[i] << " cluster events: " << accepts_[i] << std::endl;
		std::cout << "   o of rejected " << type[i] << " cluster events: " << attempts_[i] - accepts_[i] << std::endl;
	}
}

void UpdateClusters::update(membertrix &cluster_matrix, int) {
	if (!fix_) return; // the reference's update changes no parameter
	dev_.check(npb_chains_update_params(cluster_matrix.chains, NPB_UPDATE_POSTERIOR_DRAW, nullptr, 0.0, 0.0, nullptr));
	cluster_matrix.invalidate();
}

MCMC::MCMC(device &dev, dirichlet_process &hyper, UpdateClusterPopulation &ucp, int64_t chains, int Kmax, int K0, int m_aux, uint64_t seed)
	: dev_(dev), hyper_(hyper), ucp_(ucp), chains_(chains), Kmax_(Kmax), K0_(K0), m_aux_(m_aux), seed_(seed) {}
MCMC::~MCMC() {
	if (trix_) {
		npb_chains_destroy(trix_->chains);
		delete trix_;
	}
}

void MCMC::run(dataset_t &dataset, int T, bool per_item_seam, UpdateClusters *update_clusters) {
	const Suffies_NormalInvWishart &niw = hyper_.getSuffies();
	if (!trix_) {
		trix_ = new membertrix(dev_, dataset, niw.D);
		if (hyper_.mode() == clustering_mode) {
			dev_.check(npb_prior_set_niw(dev_.ctx(), niw.D, niw.mu.data(), niw.kappa, niw.nu, niw.Lambda.data(), hyper_.alpha(), NPB_BUGCOMPAT_DEFAULT));
		} else { // np_main.cpp:322-328, 357-364
			const Suffies_NormalInvGamma &nig = hyper_.getSuffiesNIG();
			dev_.check(npb_prior_set_nig(dev_.ctx(), hyper_.mode() == regression_mode ? NPB_FAMILY_REGRESSION : NPB_FAMILY_ANGULAR, nig.mu.data(),
					nig.Lambda.data(), nig.alpha, nig.beta, hyper_.alpha()));
		}
		// np_mcmc.cpp:49-91: K0 prior clusters, uniform assignment, empty clusters dropped
		dev_.check(npb_chains_create(dev_.ctx(), trix_->dataset_handle(), chains_, Kmax_, m_aux_, K0_, seed_, &trix_->chains));
	}
	const int N = trix_->size();
	const bool upd = update_clusters && update_clusters->fixes_q1();
	// np_mcmc.cpp:109-175: per sweep t the population update (:146-163), UpdateClusters::update (:170) and, when
	// t % 5 == 0, considerMaxLikelihood (:172-174)
	if (!per_item_seam && !upd) {
		// nothing happens between the sweeps t = 5j+1 .. 5j+4, so they share a launch with sweep 5(j+1)
		for (int t = 0; t < T;) {
			const int n = (t == 0) ? 1 : std::min(5, T - t);
			ucp_.sweep(*trix_, n);
			t += n;
			if ((t - 1) % 5 == 0) considerMaxLikelihood();
		}
		return;
	}
	const int sc = ucp_.subset_count();
	for (int t = 0; t < T; ++t) {
		if (per_item_seam) {
			// the reference's own loop shape: one update() per subset of subset_count items
			for (int i = 0; i < N; ++i) {
				data_ids_t subset(sc);
				for (int j = 0; j < sc; ++j) subset[j] = (i + j * 7919) % N; // placeholders: the device draws the real subsets
				ucp_.update(*trix_, subset);
			}
		} else {
			ucp_.sweep(*trix_, 1);
		}
		if (upd) update_clusters->update(*trix_, 20);
		if (t % 5 == 0) considerMaxLikelihood();
	}
}

void MCMC::considerMaxLikelihood() {
	dev_.check(npb_chains_consider_max_likelihood(trix_->chains, nullptr, nullptr));
}
double MCMC::coclusterDiagonalCheck(const std::vector<int64_t> &anchors, npb_comm *comm) {
	const size_t n = anchors.size();
	std::vector<float> h(n * n);
	dev_.check(npb_cocluster_host(trix_->chains, anchors.data(), (int64_t)n, comm, h.data()));
	double d = 0.0;
	for (size_t i = 0; i < n; ++i) d += h[i * n + i];
	return d / (double)n;
}
std::map<cluster_id_t, Suffies_MultivariateNormal> MCMC::getMaxLikelihoodClusters(int64_t chain) {
	// the clusters as they were when the state was kept: slot ids are re-used after a death, so the CURRENT slot table may
	// describe other clusters under the same ids
	const int cap = npb_chains_kmax(trix_->chains), D = trix_->dim();
	std::vector<int32_t> slots(cap);
	std::vector<int64_t> counts(cap);
	std::vector<double> mu((size_t)cap * D), sigma((size_t)cap * D * D);
	int K = 0;
	dev_.check(npb_chains_get_best_params(trix_->chains, chain, cap, &K, slots.data(), counts.data(), mu.data(), sigma.data()));
	std::map<cluster_id_t, Suffies_MultivariateNormal> out;
	for (int k = 0; k < K; ++k) {
		Suffies_MultivariateNormal s(D);
		std::copy(mu.begin() + (size_t)k * D, mu.begin() + (size_t)(k + 1) * D, s.mu.begin());
		std::copy(sigma.begin() + (size_t)k * D * D, sigma.begin() + (size_t)(k + 1) * D * D, s.sigma.begin());
		out[slots[k]] = s;
	}
	return out;
}
std::vector<int32_t> MCMC::getMaxLikelihoodAssignments(int64_t chain) {
	std::vector<int32_t> z(trix_->size());
	dev_.check(npb_chains_get_best_assignments(trix_->chains, chain, 1, z.data()));
	return z;
}

// clustering_performance::{calculateContingencyMatrix, calculateSimilarity} (clustering_performance.cpp:14-82) on the
// host for ONE labelling, in 64-bit integers (the reference's int overflows for N > 46340)
void Results::scores() {
	const size_t N = z_.size();
	if (truth_.size() != N || N < 2) return;
	std::map<std::pair<int, int>, long long> nab;
	std::map<int, long long> ra, cb;
	for (size_t i = 0; i < N; ++i) { nab[{truth_[i], z_[i]}]++; ra[truth_[i]]++; cb[z_[i]]++; }
	auto c2 = [](long long n) { return (double)n * (double)(n - 1) / 2.0; };
	std::map<int, long long> colmax;
	double a = 0, b = 0, c = 0;
	for (auto &kv : nab) { a += c2(kv.second); colmax[kv.first.second] = std::max(colmax[kv.first.second], kv.second); }
	for (auto &kv : ra) b += c2(kv.second);
	for (auto &kv : cb) c += c2(kv.second);
	double pm = 0;
	for (auto &kv : colmax) pm += (double)kv.second;
	const double S = c2((long long)N);
	purity = pm / (double)N;
	rand_index = (2 * a - b - c) / S + 1;
	adjusted_rand = (a - b * c / S) / ((b + c) / 2 - b * c / S);
}

void Results::write(const std::string &workspace, const std::string &path, const std::string &basename) {
	namespace fs = std::filesystem;
	scores();
	const std::string ws_path = workspace + path;
	fs::create_directories(ws_path);
	const std::string ws_latest = workspace + "LATEST";
	std::error_code ec;
	fs::remove(ws_latest, ec);
	fs::create_symlink(path, ws_latest, ec);
	int k = 0;
	for (auto &kv : clusters_) { // np_results.cpp:71-97: one file per cluster with its items
		std::ofstream of(ws_path + "/" + basename + std::to_string(k) + ".txt");
		for (size_t i = 0; i < z_.size(); ++i)
			if (z_[i] == kv.first) {
				for (double d : *dataset_[i]) of << d << " ";
				of << std::endl;
			}
		k++;
	}
	{ // writeOctave, np_results.cpp:112-196
		std::ofstream of(ws_path + "/" + basename + ".txt");
		const int K = (int)clusters_.size();
		if (K > 0) {
			const int D = (int)clusters_.begin()->second.mu.size();
			of << "# name: mu" << std::endl << "# type: matrix" << std::endl << "# rows: " << K << std::endl << "# columns: " << D << std::endl;
			for (auto &kv : clusters_) {
				for (int d = 0; d < D; ++d) of << (d ? " " : "") << kv.second.mu[d];
				of << std::endl;
			}
			of << std::endl << std::endl;
			of << "# name: sigma" << std::endl << "# type: matrix" << std::endl << "# ndims: 3" << std::endl << " " << D << " " << D << " " << K << std::endl;
			for (auto &kv : clusters_) { // column-major like Eigen's default storage printed with " \n" separators
				for (int c = 0; c < D; ++c)
					for (int r = 0; r < D; ++r) of << kv.second.sigma[(size_t)r * D + c] << " " << std::endl;
				of << std::endl;
			}
		}
	}
	{ // clustering_performance::write, clustering_performance.cpp:84-93
		std::ofstream of(ws_path + "/" + basename + ".score.txt");
		of << "Purity: " << purity << std::endl;
		of << "Rand Index: " << rand_index << std::endl;
		of << "Adjusted Rand Index: " << adjusted_rand << std::endl;
	}
}

clustering_scores MCMC::scores(const std::vector<int> &ground_truth) {
	clustering_scores s;
	const size_t C = (size_t)chains_;
	s.purity.resize(C); s.rand_index.resize(C); s.adjusted_rand.resize(C); s.joint_loglik.resize(C); s.K.resize(C);
	std::vector<int32_t> gt(ground_truth.begin(), ground_truth.end());
	dev_.check(npb_chains_metrics(trix_->chains, gt.empty() ? nullptr : gt.data(), s.purity.data(), s.rand_index.data(),
			s.adjusted_rand.data(), s.joint_loglik.data(), s.K.data()));
	return s;
}

} // namespace npb
