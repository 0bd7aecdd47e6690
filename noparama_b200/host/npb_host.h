// npb_host.h -- C++ host side above the C ABI (include/npb200.h): the reference's sampler seam, kept by name.
//
// The reference wires  MCMC(generator, InitClusters, UpdateClusters, UpdateClusterPopulation, subset_count, likelihood)
// by hand in main (src/np_main.cpp:391-468) and drives  UpdateClusterPopulation::update(membertrix&, data_ids)  once
// per item per sweep (src/np_mcmc.cpp:146-163).  These classes keep those names, argument meanings and error behaviour
// (np_error_t), without Eigen, and delegate every computation to libnpb200.so.  One object graph drives `chains`
// lockstep chains; chain 0 is what the single-chain accessors of the reference interface see.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/npb200.h"

namespace npb {

// ---- include/np_data.h:9-25, include/np_cluster.h:12-18 ----
typedef std::vector<double> data_t;
typedef std::vector<data_t *> dataset_t;
typedef int data_id_t;
typedef std::vector<data_id_t> data_ids_t;
typedef int cluster_id_t;
typedef std::vector<cluster_id_t> cluster_ids_t;

// ---- include/membertrix.h:16-23 ----
enum np_error_t { error_none, error_already_assigned, error_assignment_remaining, error_assignment_absent };

// ---- include/np_suffies.h:80-95,187-200 and Suffies_Dirichlet ----
struct Suffies_NormalInvWishart {
	int D;
	std::vector<double> mu;     // [D]
	double kappa, nu;
	std::vector<double> Lambda; // [D,D] row-major
	explicit Suffies_NormalInvWishart(int d) : D(d), mu(d, 0.0), kappa(1.0), nu(d + 2.0), Lambda((size_t)d * d, 0.0) {}
};
struct Suffies_MultivariateNormal {
	int D;
	std::vector<double> mu, sigma;
	explicit Suffies_MultivariateNormal(int d = 0) : D(d), mu(d, 0.0), sigma((size_t)d * d, 0.0) {}
};
struct Suffies_Dirichlet {
	double alpha = 1.0;
};
// include/np_suffies.h:112-128: the base measure of `-c regression` / `-c angular` (constants np_main.cpp:357-364)
struct Suffies_NormalInvGamma {
	int D;
	std::vector<double> mu;     // [2]
	double alpha, beta;
	std::vector<double> Lambda; // [2,2] row-major
	explicit Suffies_NormalInvGamma(int d = 2) : D(d), mu(d, 0.0), alpha(10.0), beta(0.1), Lambda((size_t)d * d, 0.0) {}
};
// include/statistics/scalarnoise_multivariatenormal.h:14
enum representation_mode_t { clustering_mode, regression_mode, angular_mode, points3d_mode };

struct npb_error : std::runtime_error {
	npb_status status;
	npb_error(npb_status s, const std::string &what) : std::runtime_error(what), status(s) {}
};

// One CUDA device + stream.  There is no CPU path: construction throws without a GPU.
class device {
public:
	explicit device(int index = 0);
	~device();
	npb_ctx *ctx() const { return ctx_; }
	void check(npb_status s) const;
private:
	npb_ctx *ctx_ = nullptr;
	device(const device &) = delete;
};

// dirichlet_process(alpha, normal_inverse_wishart_distribution)  (include/statistics/dirichlet.h:20-93)
class dirichlet_process {
public:
	dirichlet_process(const Suffies_Dirichlet &d, const Suffies_NormalInvWishart &niw) : alpha_(d.alpha), niw_(niw) {}
	// scalar-noise likelihood (regression_mode: rows (1, a, b); angular_mode: rows (a, b)) with a normal-inverse-gamma base measure
	dirichlet_process(const Suffies_Dirichlet &d, const Suffies_NormalInvGamma &nig, representation_mode_t mode)
		: alpha_(d.alpha), niw_(mode == regression_mode ? 3 : 2), nig_(nig), mode_(mode) {}
	const Suffies_NormalInvWishart &getSuffies() const { return niw_; } // (niw.D = width of a data row in either case)
	const Suffies_NormalInvGamma &getSuffiesNIG() const { return nig_; }
	representation_mode_t mode() const { return mode_; }
	double alpha() const { return alpha_; }
private:
	double alpha_;
	Suffies_NormalInvWishart niw_;
	Suffies_NormalInvGamma nig_;
	representation_mode_t mode_ = clustering_mode;
};

// The membership state (include/membertrix.h:52-314) of the device chains.  Cluster ids are device slot ids: stable,
// so relabel() has nothing to do.  Reads refresh a host copy of one chain on demand.
class membertrix {
public:
	membertrix(device &dev, dataset_t &dataset, int D);
	~membertrix();
	int size() const { return N_; }
	int dim() const { return D_; }
	data_t *getDatum(data_id_t i) { return (*dataset_)[i]; }                            // membertrix.cpp:140-144
	cluster_id_t getClusterId(data_id_t i);                                              // :235-244
	size_t getClusterCount();                                                            // :246-248
	std::map<cluster_id_t, Suffies_MultivariateNormal> getClusters();                    // :250-257 (parameters by value)
	size_t count(cluster_id_t k);                                                        // :328-330
	void getAssignments(cluster_id_t k, data_ids_t &ids);                                // :315-322
	bool assigned(data_id_t i) { return getClusterId(i) >= 0; }                          // :166-168
	void relabel() {}                                                                    // :259-262
	// mutators (membertrix.h:106-152, membertrix.cpp:87-233, 343-364).  The device holds complete states only (every item
	// in a cluster), so a retracted item is pending on the host until it is assigned again; assign() then moves it on the
	// device in one call (npb_chain_move_item / npb_chain_move_item_new).  A cluster added with addCluster lives on the
	// host until its first member arrives.  Error codes are the reference's np_error_t.
	cluster_id_t addCluster(const Suffies_MultivariateNormal &suffies);                  // :87-118
	np_error_t assign(cluster_id_t k, data_id_t i);                                      // :147-164
	np_error_t retract(data_id_t i, bool auto_remove = true);                            // :175-233 (both overloads)
	np_error_t remove(cluster_id_t k);                                                   // :213-228
	np_error_t cleanup();                                                                // :343-364: drop clusters without members
	bool empty(cluster_id_t k) { return count(k) == 0; }                                 // :324-326
	int pending() const { return (int)pending_.size(); }                                 // items retracted and not yet assigned
	membertrix *clone();                                                                 // :57-78: a detached, read-only copy
	void select_chain(int64_t chain) { chain_ = chain; dirty_ = true; }
	// device side
	npb_dataset *dataset_handle() const { return ds_; }
	npb_chains *chains = nullptr; // owned by MCMC
	void invalidate() { dirty_ = true; }
private:
	void refresh();
	device &dev_;
	dataset_t *dataset_;
	npb_dataset *ds_ = nullptr;
	int N_, D_;
	int64_t chain_ = 0;
	bool dirty_ = true;
	std::vector<int32_t> z_;
	std::vector<int32_t> slots_;
	std::vector<int64_t> counts_;
	std::vector<double> mu_, sigma_;
	std::map<data_id_t, cluster_id_t> pending_;                       // retracted item -> the cluster it still occupies on the device
	std::map<cluster_id_t, Suffies_MultivariateNormal> added_;        // clusters without members yet: host only, ids >= Kmax
	bool detached_ = false;                                            // clone(): no device behind it
};

// include/np_update_cluster_population.h:13-44
class UpdateClusterPopulation {
public:
	virtual ~UpdateClusterPopulation() {}
	virtual void update(membertrix &cluster_matrix, const data_ids_t &data_ids) = 0;
	virtual void printStatistics() = 0;
	virtual void sweep(membertrix &cluster_matrix, int n_sweeps) = 0; // batched form: n_sweeps sweeps of every chain
	virtual int sampler() const = 0;
	virtual int subset_count() const = 0;
};

// NealAlgorithm8 (include/np_neal_algorithm8.h:52-56, src/np_neal_algorithm8.cpp:49-176).  update() keeps the
// per-item calling convention of MCMC::run: the first call of a sweep performs the whole sweep of every chain on the
// device, the remaining N-1 calls of that sweep are absorbed.
class NealAlgorithm8 : public UpdateClusterPopulation {
public:
	NealAlgorithm8(device &dev, dirichlet_process &nonparametrics) : dev_(dev), hyper_(nonparametrics) {}
	void update(membertrix &cluster_matrix, const data_ids_t &data_ids) override;
	void sweep(membertrix &cluster_matrix, int n_sweeps) override;
	void printStatistics() override;
	int sampler() const override { return NPB_ALG8; }
	int subset_count() const override { return 1; }
	const npb_sweep_stats &last() const { return last_; }
protected:
	device &dev() { return dev_; }
	void record(const npb_sweep_stats &st) {
		last_ = st;
		accepted_ += st.new_clusters;
		rejected_ += st.reassignments - st.new_clusters;
	}
private:
	device &dev_;
	dirichlet_process &hyper_;
	int64_t calls_ = 0;
	int64_t accepted_ = 0, rejected_ = 0;
	npb_sweep_stats last_{};
};

// NealAlgorithm2 (include/np_neal_algorithm2.h, src/np_neal_algorithm2.cpp:32-120; not compiled by the reference): one prior
// draw weighted alpha beside the K occupied clusters -- the same device sweep with m_aux = 1 (MCMC must be built with
// m_aux = 1 for it).
class NealAlgorithm2 : public NealAlgorithm8 {
public:
	NealAlgorithm2(device &dev, dirichlet_process &nonparametrics) : NealAlgorithm8(dev, nonparametrics) {}
	int sampler() const override { return NPB_ALG2; }
	void sweep(membertrix &cluster_matrix, int n_sweeps) override;
};

// The CONJUGATE Algorithm 2 the reference meant (np_neal_algorithm2.cpp:32-120 with updateSuffies; never completed there): collapsed
// Gibbs with the NIW posterior predictive, sufficient statistics up- and down-dated per move (NPB_ALG2_CONJUGATE; Kmax = 32).
class NealAlgorithm2Conjugate : public NealAlgorithm8 {
public:
	NealAlgorithm2Conjugate(device &dev, dirichlet_process &nonparametrics) : NealAlgorithm8(dev, nonparametrics) {}
	int sampler() const override { return NPB_ALG2_CONJUGATE; }
	void sweep(membertrix &cluster_matrix, int n_sweeps) override;
};

// The split-merge samplers (include/np_jain_neal_algorithm.h:75-79, include/np_triadic_algorithm.h:73-77).  update()
// keeps MCMC::run's per-subset calling convention (np_mcmc.cpp:146-163): subsets of 2 / 3 items, asserted like the
// reference (np_jain_neal_algorithm.cpp:429, np_triadic_algorithm.cpp:647); the first call of a sweep runs the whole
// sweep of every chain on the device (the device draws its own lockstep subsets), the others are absorbed.
class SplitMergeAlgorithm : public UpdateClusterPopulation {
public:
	SplitMergeAlgorithm(device &dev, dirichlet_process &nonparametrics, int sampler, int subsets)
		: dev_(dev), hyper_(nonparametrics), sampler_(sampler), subsets_(subsets) {}
	void update(membertrix &cluster_matrix, const data_ids_t &data_ids) override;
	void sweep(membertrix &cluster_matrix, int n_sweeps) override;
	void printStatistics() override; // np_jain_neal_algorithm.cpp:504-531, np_triadic_algorithm.cpp:797-826
	int sampler() const override { return sampler_; }
	int subset_count() const override { return subsets_; }
	const npb_sweep_stats &last() const { return last_; }
	const int64_t *attempts() const { return attempts_; }
	const int64_t *accepts() const { return accepts_; }
private:
	device &dev_;
	dirichlet_process &hyper_;
	int sampler_, subsets_;
	int64_t calls_ = 0;
	int64_t attempts_[4] = {0, 0, 0, 0}, accepts_[4] = {0, 0, 0, 0};
	npb_sweep_stats last_{};
};
class JainNealAlgorithm : public SplitMergeAlgorithm {
public:
	JainNealAlgorithm(device &dev, dirichlet_process &nonparametrics) : SplitMergeAlgorithm(dev, nonparametrics, NPB_JAIN_NEAL, 2) {}
};
class TriadicAlgorithm : public SplitMergeAlgorithm {
public:
	TriadicAlgorithm(device &dev, dirichlet_process &nonparametrics) : SplitMergeAlgorithm(dev, nonparametrics, NPB_TRIADIC, 3) {}
};

// UpdateClusters (include/np_update_clusters.h, src/np_update_clusters.cpp:71-141), called after every sweep
// (np_mcmc.cpp:170).  In the reference its result is sliced away (np_cluster.h:49-51): parameters never change,
// so the default here is the same no-op; with fix_q1 the parameters of every cluster are redrawn from the conjugate
// normal-inverse-Wishart posterior of its members on the device (npb_chains_update_params).
class UpdateClusters {
public:
	UpdateClusters(device &dev, dirichlet_process &nonparametrics, bool fix_q1 = false) : dev_(dev), hyper_(nonparametrics), fix_(fix_q1) {}
	void update(membertrix &cluster_matrix, int number_mh_steps);
	bool fixes_q1() const { return fix_; }
private:
	device &dev_;
	dirichlet_process &hyper_;
	bool fix_;
};

struct clustering_scores { // src/clustering_performance.cpp:38-82
	std::vector<double> purity, rand_index, adjusted_rand, joint_loglik;
	std::vector<int32_t> K;
};

// MCMC (include/np_mcmc.h:71-95, src/np_mcmc.cpp:48-175) over `chains` lockstep chains.
class MCMC {
public:
	MCMC(device &dev, dirichlet_process &hyper, UpdateClusterPopulation &update_cluster_population, int64_t chains = 1,
			int Kmax = 256, int K0 = 20, int m_aux = 3, uint64_t seed = 20261018);
	~MCMC();
	void run(dataset_t &dataset, int T, bool per_item_seam = false, UpdateClusters *update_clusters = nullptr);
	membertrix &getMembershipMatrix() { return *trix_; }                  // np_mcmc.h:88
	std::map<cluster_id_t, Suffies_MultivariateNormal> getMaxLikelihoodClusters(int64_t chain = 0); // the kept state's clusters
	std::vector<int32_t> getMaxLikelihoodAssignments(int64_t chain = 0);  // np_mcmc.h:90: the state kept by considerMaxLikelihood
	void considerMaxLikelihood();                                         // np_mcmc.cpp:187-203, every chain at once
	// co-clustering counts of `anchors` over this object's chains (npb_cocluster_allreduce: summed over the communicator's ranks when
	// one is given); returns the mean diagonal entry = the total number of chains that contributed
	double coclusterDiagonalCheck(const std::vector<int64_t> &anchors, npb_comm *comm);
	clustering_scores scores(const std::vector<int> &ground_truth);        // np_results.cpp:17-37 + clustering_performance
	int64_t chains() const { return chains_; }
private:
	device &dev_;
	dirichlet_process &hyper_;
	UpdateClusterPopulation &ucp_;
	int64_t chains_;
	int Kmax_, K0_, m_aux_;
	uint64_t seed_;
	membertrix *trix_ = nullptr;
};

// Results (include/np_results.h, src/np_results.cpp:39-196): writes one clustering of one chain the way the reference does --
// <workspace>/<path>/<basename><k>.txt with the items of cluster k, <basename>.txt with the parameters as Octave
// matrices "mu" and "sigma", <basename>.score.txt with purity / Rand / adjusted Rand (clustering_performance.cpp:84-93),
// and the symlink <workspace>/LATEST -> <path>.
class Results {
public:
	Results(dataset_t &dataset, const std::vector<int32_t> &assignments, const std::map<cluster_id_t, Suffies_MultivariateNormal> &clusters,
			const std::vector<int> &ground_truth)
		: dataset_(dataset), z_(assignments), clusters_(clusters), truth_(ground_truth) {}
	void write(const std::string &workspace, const std::string &path, const std::string &basename);
	double purity = 0, rand_index = 0, adjusted_rand = 0;
private:
	void scores();
	dataset_t &dataset_;
	std::vector<int32_t> z_;
	std::map<cluster_id_t, Suffies_MultivariateNormal> clusters_;
	std::vector<int> truth_;
};

} // namespace npb
