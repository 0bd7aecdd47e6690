// npb_main.cpp -- the reference's command line (src/np_main.cpp:187-267) on top of the device path:
//   noparama_b200 -d <datafile> -a <algorithm8|jain_neal_split|triadic> -T <sweeps> -c clustering|regression|angular [--chains C] [--gpus G] [--seed S] [--kmax K] [--seam] [--fix-q1]
// --gpus G  the chains are split over G devices of this node, one host thread and one object graph per device (chains are
//     independent: no exchange during the run); the scores of all chains are gathered at the end
// -d  text file, one item per line: D coordinates then the ground-truth label (the reference reads exactly 2 + 1
//     columns, np_main.cpp:93-101; here D = columns - 1 <= 3 for the register kernel, 4/8/16 for the tile kernel)
// -a  algorithm8 | jain_neal_split | triadic (np_main.cpp:228-238) | algorithm2 (commented out there, :222-227)
// -T  sweeps (default 2000, np_main.cpp:242)    -c  clustering | regression | angular (points3d is out of scope)
// Prior and constants as hard-wired in the reference: alpha = 1, NIW{mu = 6, kappa = 1/500, nu = D + 2, Lambda = 0.01 I}
// (np_main.cpp:164,367-371), K0 = 20, M = 3.  No 200-row subsampling (np_main.cpp:166-167): every row is used.
#include "npb_host.h"

#include <cstdio>
#include <ctime>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <numeric>
#include <sstream>
#include <memory>
#include <mutex>
#include <random>
#include <thread>

using namespace npb;

static void usage() {
	std::cout << "usage: noparama_b200 -d datafile -a algorithm8|algorithm2|algorithm2_conjugate|jain_neal_split|triadic -T sweeps -c clustering|regression|angular [--chains C] [--gpus G] [--seed S] [--kmax K] [--seam] [--fix-q1] [--output DIR] [--selftest-membertrix]\n";
}

// test/test_membertrix.cpp (the reference's unit test of the state class) through the host mutators over the DEVICE state:
// 5 items, 4 clusters, a random assignment; after retracting every member of a random cluster the cluster count must drop
// by exactly one if it had members, by none otherwise (test_membertrix.cpp:75-91).  Then the error codes.
static int membertrix_selftest(device &dev, dirichlet_process &hyper, dataset_t &dataset, unsigned long long seed) {
	const int D = (int)dataset[0]->size();
	dataset_t five(dataset.begin(), dataset.begin() + 5);
	NealAlgorithm8 alg8(dev, hyper);
	MCMC mcmc(dev, hyper, alg8, 1, 32, 4, 3, seed);
	mcmc.run(five, 0, false, nullptr); // init only: K0 = 4 prior clusters, uniform assignment, empty ones dropped
	membertrix &trix = mcmc.getMembershipMatrix();
	std::mt19937 gen((unsigned)seed);
	int fails = 0;
	auto expect = [&](bool ok, const char *what) { if (!ok) { std::cout << "FAIL: " << what << std::endl; fails++; } };
	for (int round = 0; round < 8; ++round) {
		auto clusters = trix.getClusters();
		const size_t before = trix.getClusterCount();
		std::vector<cluster_id_t> ids;
		for (auto &kv : clusters) ids.push_back(kv.first);
		const cluster_id_t l = ids[gen() % ids.size()];
		data_ids_t members;
		trix.getAssignments(l, members);
		// move every member of l to a new cluster with the same parameters (retract, addCluster, assign)
		const cluster_id_t fresh = trix.addCluster(clusters[l]);
		cluster_id_t target = fresh;
		for (data_id_t i : members) {
			expect(trix.retract(i) == error_none, "retract of an assigned item");
			expect(trix.retract(i) == error_assignment_absent, "second retract must report error_assignment_absent");
			expect(trix.assign(target, i) == error_none, "assign after retract");
			expect(trix.assign(target, i) == error_already_assigned, "second assign must report error_already_assigned");
			target = trix.getClusterId(i); // the host id has become a device slot
		}
		const size_t after = trix.getClusterCount();
		expect(after == before, "moving all members of a cluster to a new one keeps the cluster count");
		expect(trix.count(l) == 0 || target == l, "the emptied cluster is gone (or its slot was re-used by the new one)");
		if (!members.empty() && ids.size() > 1) {
			// now merge the new cluster into another one: the count must drop by exactly one
			cluster_id_t other = -1;
			for (auto &kv : trix.getClusters()) if (kv.first != target) { other = kv.first; break; }
			expect(trix.remove(target) == error_assignment_remaining, "remove of a cluster with members must report error_assignment_remaining");
			data_ids_t m2;
			trix.getAssignments(target, m2);
			for (data_id_t i : m2) { trix.retract(i); expect(trix.assign(other, i) == error_none, "assign to an existing cluster"); }
			expect(trix.getClusterCount() == after - 1, "retracting every member of a cluster drops exactly one cluster");
			expect(trix.assign(31, 0) == error_already_assigned, "assign without retract");
		}
		int total = 0;
		for (auto &kv : trix.getClusters()) total += (int)trix.count(kv.first);
		expect(total == 5, "every item is in exactly one cluster");
	}
	membertrix *copy = trix.clone();
	expect(copy->getClusterCount() == trix.getClusterCount() && copy->getClusterId(3) == trix.getClusterId(3), "clone() carries the state");
	delete copy;
	std::cout << (fails ? "membertrix selftest FAILED" : "membertrix selftest passed") << std::endl;
	return fails ? 3 : 0;
}

int main(int argc, char **argv) {
	std::string datafile, algorithm = "algorithm8", config = "clustering";
	int T = 2000, kmax = 256;
	long long chains = 1;
	unsigned long long seed = 20261018ull;
	bool seam = false, fix_q1 = false, selftest = false;
	int gpus = 1;
	std::string output; // workspace root for the results files; empty = do not write (the reference always writes to output/)
	for (int i = 1; i < argc; ++i) {
		std::string a = argv[i];
		auto next = [&](const char *what) -> const char * {
			if (i + 1 >= argc) { std::cerr << "missing value for " << what << std::endl; exit(1); }
			return argv[++i];
		};
		if (a == "-d") datafile = next("-d");
		else if (a == "-a") algorithm = next("-a");
		else if (a == "-T") T = atoi(next("-T"));
		else if (a == "-c") config = next("-c");
		else if (a == "--chains") chains = atoll(next("--chains"));
		else if (a == "--seed") seed = strtoull(next("--seed"), nullptr, 10);
		else if (a == "--kmax") kmax = atoi(next("--kmax"));
		else if (a == "--gpus") gpus = atoi(next("--gpus"));
		else if (a == "--selftest-membertrix") selftest = true;
		else if (a == "--seam") seam = true;
		else if (a == "--fix-q1") fix_q1 = true;
		else if (a == "--output") output = next("--output");
		else if (a == "-h" || a == "-?") { usage(); return 0; }
		else { std::cerr << "unknown option " << a << std::endl; usage(); return 1; }
	}
	if (datafile.empty()) { usage(); return 1; }
	if (config != "clustering" && config != "regression" && config != "angular") { std::cerr << "Unknown likelihood" << std::endl; return 107; } // np_main.cpp:196-205, exit code :349
	const representation_mode_t mode = config == "regression" ? regression_mode : (config == "angular" ? angular_mode : clustering_mode);
	if (algorithm == "algorithm2_conjugate" && kmax != 32) {
		std::cout << "algorithm2_conjugate keeps 32 cluster slots per chain: --kmax 32" << std::endl;
		kmax = 32;
	}
	if (mode != clustering_mode && algorithm != "algorithm8") {
		std::cerr << "-c " << config << " runs Algorithm 8 on the device (the split-merge samplers are multivariate-normal only)" << std::endl;
		return 1;
	}
	if (algorithm != "algorithm8" && algorithm != "algorithm2" && algorithm != "algorithm2_conjugate" && algorithm != "jain_neal_split" &&
			algorithm != "triadic") { // np_main.cpp:222-238
		std::cerr << "Unknown algorithm: " << algorithm << std::endl;
		return 1;
	}

	// read_data (np_main.cpp:57-148), generalised to D columns + label
	std::ifstream in(datafile);
	if (!in) { std::cerr << "cannot open " << datafile << std::endl; return 7; }
	dataset_t dataset;
	std::vector<int> ground_truth;
	std::string line;
	int D = -1;
	while (std::getline(in, line)) {
		std::istringstream ss(line);
		std::vector<double> row;
		double v;
		while (ss >> v) row.push_back(v);
		if (row.empty()) continue;
		if (D < 0) D = (int)row.size() - 1;
		if ((int)row.size() != D + 1 || D < 1) { std::cerr << "ragged line in " << datafile << std::endl; return 7; }
		ground_truth.push_back((int)row.back());
		row.pop_back();
		if (mode == regression_mode) row.insert(row.begin(), 1.0); // np_main.cpp:83-92: prepend the constant
		dataset.push_back(new data_t(row));
	}
	if (mode != clustering_mode && D != 2) { std::cerr << "-c " << config << " expects lines \"a b label\" (np_main.cpp:76-101)" << std::endl; return 7; }
	if (mode == regression_mode) D = 3;
	std::cout << "Read " << dataset.size() << " items of dimension " << D << std::endl;

	// one device: its own context, object graph and share of the chains
	auto run_on_device = [&](device &dev, npb_comm *comm, int dev_index, long long chains, unsigned long long seed, bool lead, clustering_scores &sc_out, std::mutex &io) -> int {
	try {
			Suffies_Dirichlet sd;
			sd.alpha = 1.0; // np_main.cpp:164
			Suffies_NormalInvWishart niw(D);
			for (int d = 0; d < D; ++d) { niw.mu[d] = 6.0; niw.Lambda[(size_t)d * D + d] = 0.01; }
			niw.kappa = 1.0 / 500;
			niw.nu = D + 2.0;
			Suffies_NormalInvGamma nig(2); // np_main.cpp:357-364
			nig.Lambda[0] = nig.Lambda[3] = 0.01;
			dirichlet_process hyper = mode == clustering_mode ? dirichlet_process(sd, niw) : dirichlet_process(sd, nig, mode);
			// np_main.cpp:424-459
			NealAlgorithm8 alg8(dev, hyper);
			JainNealAlgorithm jain_neal(dev, hyper);
			TriadicAlgorithm triadic(dev, hyper);
			NealAlgorithm2 alg2(dev, hyper); // np_main.cpp:222-227,425-431 (commented out in the reference)
			NealAlgorithm2Conjugate alg2c(dev, hyper); // the collapsed form np_neal_algorithm2.cpp:32-120 describes
			UpdateClusterPopulation &sampler = algorithm == "algorithm8" ? (UpdateClusterPopulation &)alg8
					: algorithm == "algorithm2_conjugate" ? (UpdateClusterPopulation &)alg2c
					: (algorithm == "algorithm2" ? (UpdateClusterPopulation &)alg2
					: (algorithm == "jain_neal_split" ? (UpdateClusterPopulation &)jain_neal : (UpdateClusterPopulation &)triadic));
			MCMC mcmc(dev, hyper, sampler, chains, kmax, 20, algorithm == "algorithm2" ? 1 : 3, seed);
			std::cout << "Run MCMC for " << T << " steps, " << chains << " chain(s)" << std::endl;
			UpdateClusters update_clusters(dev, hyper, fix_q1); // np_main.cpp:415
			if (selftest) return membertrix_selftest(dev, hyper, dataset, seed);
			mcmc.run(dataset, T, seam, &update_clusters);
			sampler.printStatistics();
			clustering_scores sc = mcmc.scores(ground_truth);
			sc_out = sc;
			// posterior co-clustering counts of (up to) 1024 evenly spaced items over this device's chains, summed over the devices
			// (ncclAllReduce inside the library; the only exchange of a run)
			std::vector<int64_t> anchors;
			const int64_t Nn = (int64_t)dataset.size(), step = std::max<int64_t>(1, Nn / 1024);
			for (int64_t i = 0; i < Nn && anchors.size() < 1024; i += step) anchors.push_back(i);
			const double cc = mcmc.coclusterDiagonalCheck(anchors, comm);
			std::lock_guard<std::mutex> lock(io);
			if (lead) std::cout << "Co-clustering matrix of " << anchors.size() << " anchors over all chains: mean self-count " << cc
					      << (comm ? " (all-reduced over NCCL)" : "") << std::endl;
			if (!lead) {
				std::cout << "device " << dev_index << ": " << chains << " chain(s) done" << std::endl;
				return 0;
			}
			auto mean = [](const std::vector<double> &v) { return std::accumulate(v.begin(), v.end(), 0.0) / v.size(); };
			// clustering_performance.cpp:77-79 prints the three scores; chain 0 first, then the mean over chains
			std::cout << "Purity: " << sc.purity[0] << std::endl;
			std::cout << "Rand Index: " << sc.rand_index[0] << std::endl;
			std::cout << "Adjusted Rand Index: " << sc.adjusted_rand[0] << std::endl;
			std::cout << "Clusters: " << sc.K[0] << "  Loglikelihood now: " << sc.joint_loglik[0] << std::endl;
			if (chains > 1)
				std::cout << "Mean over " << chains << " chains: purity " << mean(sc.purity) << " rand " << mean(sc.rand_index)
					  << " adjusted rand " << mean(sc.adjusted_rand) << std::endl;
			membertrix &trix = mcmc.getMembershipMatrix();
			if (!output.empty()) {
				// np_main.cpp:476-497: the last state ("snapshot") and the max-likelihood state ("results") of chain 0, written
				// to <output>/<algorithm>/<datafile name>/<timestamp>/ with the reference's file layout (np_results.cpp:39-196)
				std::string base = datafile.substr(datafile.find_last_of('/') == std::string::npos ? 0 : datafile.find_last_of('/') + 1);
				const std::string workspace = output + "/" + algorithm + "/" + base + "/";
				char stamp[32];
				std::time_t now = std::time(nullptr);
				std::strftime(stamp, sizeof(stamp), "%Y%m%d_%H:%M", std::localtime(&now));
				std::vector<int32_t> z_last(trix.size());
				for (int i = 0; i < trix.size(); ++i) z_last[i] = trix.getClusterId(i);
				Results snapshot(dataset, z_last, trix.getClusters(), ground_truth);
				snapshot.write(workspace, stamp, "snapshot");
				// the kept state with ITS clusters (a slot that died since the snapshot may have been re-used by another cluster)
				Results results(dataset, mcmc.getMaxLikelihoodAssignments(0), mcmc.getMaxLikelihoodClusters(0), ground_truth);
				results.write(workspace, stamp, "results");
				std::cout << "Wrote snapshot* and results* to " << workspace << stamp << std::endl;
				std::cout << "Max-likelihood state: purity " << results.purity << " rand " << results.rand_index << " adjusted rand "
					  << results.adjusted_rand << std::endl;
			}
			for (auto &kv : trix.getClusters())
				std::cout << " cluster " << kv.first << " [#" << trix.count(kv.first) << "] mu " << kv.second.mu[0] << (D > 1 ? " " : "")
					  << (D > 1 ? std::to_string(kv.second.mu[1]) : "") << std::endl;
		} catch (const npb_error &e) {
			std::cerr << "npb200: " << e.what() << std::endl;
			return e.status == NPB_E_CUDA ? 2 : 1;
		}
		return 0;
	};
	if (gpus < 1 || chains < gpus) { std::cerr << "--gpus must be between 1 and --chains" << std::endl; return 1; }
	int rc = 0;
	std::vector<clustering_scores> all(gpus);
	std::mutex io;
	// one context per device, created up front so that the communicators can be (ncclCommInitAll through npb_comm_create_all)
	std::vector<std::unique_ptr<device>> devs;
	std::vector<npb_comm *> comms(gpus, nullptr);
	try {
		for (int g = 0; g < gpus; ++g) devs.emplace_back(new device(g));
		if (gpus > 1 && !selftest) {
			std::vector<npb_ctx *> ctxs;
			for (auto &d : devs) ctxs.push_back(d->ctx());
			devs[0]->check(npb_comm_create_all(ctxs.data(), gpus, comms.data()));
		}
	} catch (const npb_error &e) {
		std::cerr << "npb200: " << e.what() << std::endl;
		return e.status == NPB_E_CUDA ? 2 : 1;
	}
	auto run_device = [&](int g) {
		// chains g * chains / gpus .. (g + 1) * chains / gpus - 1 on device g (contiguous blocks, dataset replicated)
		const long long c0 = chains * g / gpus, c1 = chains * (g + 1) / gpus;
		int r = run_on_device(*devs[g], comms[g], g, c1 - c0, seed + 0x9E3779B97F4A7C15ull * (unsigned long long)g, g == 0, all[g], io);
		if (r) rc = r;
	};
	if (gpus == 1) {
		run_device(0);
	} else {
		std::vector<std::thread> th;
		for (int g = 0; g < gpus; ++g) th.emplace_back(run_device, g);
		for (auto &t : th) t.join();
		if (!rc) {
			double pu = 0, ri = 0, ar = 0, k = 0;
			size_t n = 0;
			for (auto &sc : all)
				for (size_t c = 0; c < sc.purity.size(); ++c, ++n) { pu += sc.purity[c]; ri += sc.rand_index[c]; ar += sc.adjusted_rand[c]; k += sc.K[c]; }
			std::cout << "All " << n << " chains on " << gpus << " devices: purity " << pu / n << " rand " << ri / n << " adjusted rand " << ar / n
				  << " clusters " << k / n << std::endl;
		}
	}
	for (auto c : comms) npb_comm_destroy(c);
	for (auto p : dataset) delete p;
	return rc;
}
