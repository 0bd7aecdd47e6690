"""The C++ host side (noparama_b200/host): the reference's CLI flags on the device path."""
import os
import re
import subprocess

import numpy as np
import pytest

from noparama_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "noparama_b200", "host", "noparama_b200")


def ensure_built():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.dirname(CLI)])


def write_data(path):
    X, y = syn.config(1)
    with open(path, "w") as f:
        for row, lab in zip(X, y):
            f.write("%.9f %.9f %d\n" % (row[0], row[1], lab))


def test_cli_usage_and_refusals(tmp_path):
    ensure_built()
    r = subprocess.run([CLI, "-h"], capture_output=True, text=True)
    assert r.returncode == 0 and "-d datafile" in r.stdout
    data = tmp_path / "twogaussians.data"
    write_data(str(data))
    # the reference exits 107 on an unknown likelihood (np_main.cpp:346,385) and 1 on an unknown algorithm (:236)
    assert subprocess.run([CLI, "-d", str(data), "-c", "points3d"], capture_output=True).returncode == 107
    # the scalar-noise families run Algorithm 8 only on the device
    r = subprocess.run([CLI, "-d", str(data), "-c", "regression", "-a", "triadic"], capture_output=True, text=True)
    assert r.returncode == 1 and "Algorithm 8" in r.stderr
    r = subprocess.run([CLI, "-d", str(data), "-a", "algorithm9"], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown algorithm" in r.stderr
    assert subprocess.run([CLI, "-d", str(tmp_path / "missing")], capture_output=True).returncode == 7


def test_cli_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ensure_built()
    data = tmp_path / "d.data"
    write_data(str(data))
    r = subprocess.run([CLI, "-d", str(data), "-T", "5"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_cli_runs_config1(tmp_path):
    ensure_built()
    data = tmp_path / "twogaussians.data"
    write_data(str(data))
    outs = []
    for extra in ([], ["--seam"]):
        r = subprocess.run([CLI, "-d", str(data), "-a", "algorithm8", "-T", "300", "-c", "clustering", "--chains", "64",
                            "--kmax", "64", "--seed", "7"] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout)
        pur = float(re.search(r"^Purity: ([0-9.]+)", r.stdout, re.M).group(1))
        ri = float(re.search(r"^Rand Index: ([0-9.]+)", r.stdout, re.M).group(1))
        mean_pur = float(re.search(r"chains: purity ([0-9.]+)", r.stdout).group(1))
        assert pur > 0.95 and mean_pur > 0.98 and ri < pur  # README.rst:55
        assert "new cluster events accepted" in r.stdout
    # the per-item seam (one update() per item, np_mcmc.cpp:162) and the batched sweep walk the same trajectory
    assert outs[0] == outs[1]


@pytest.mark.gpu
def test_cli_runs_algorithm2(tmp_path):
    ensure_built()
    data = tmp_path / "twogaussians.data"
    write_data(str(data))
    r = subprocess.run([CLI, "-d", str(data), "-a", "algorithm2", "-T", "200", "-c", "clustering", "--chains", "32", "--kmax", "64",
                        "--seed", "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert float(re.search(r"chains: purity ([0-9.]+)", r.stdout).group(1)) > 0.97


@pytest.mark.gpu
def test_cli_runs_conjugate_algorithm2(tmp_path):
    """-a algorithm2_conjugate: the collapsed sampler np_neal_algorithm2.cpp:32-120 describes, through the C++ host classes"""
    ensure_built()
    data = tmp_path / "twogaussians.data"
    write_data(str(data))
    r = subprocess.run([CLI, "-d", str(data), "-a", "algorithm2_conjugate", "-T", "100", "-c", "clustering", "--chains", "32", "--kmax", "64",
                        "--seed", "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "--kmax 32" in r.stdout
    assert float(re.search(r"chains: purity ([0-9.]+)", r.stdout).group(1)) > 0.9


@pytest.mark.gpu
@pytest.mark.parametrize("algorithm", ["jain_neal_split", "triadic"])
def test_cli_runs_split_merge(tmp_path, algorithm):
    ensure_built()
    data = tmp_path / "twogaussians.data"
    write_data(str(data))
    outs = []
    for extra in ([], ["--seam"]):
        r = subprocess.run([CLI, "-d", str(data), "-a", algorithm, "-T", "40", "-c", "clustering", "--chains", "32", "--kmax",
                            "64", "--seed", "9"] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout)
        assert "merge from 2 to 1 attempts" in r.stdout and "split from 1 to 2 attempts" in r.stdout
        assert ("merge from 3 to 2 attempts" in r.stdout) == (algorithm == "triadic")
        assert re.search(r"^Purity: ([0-9.]+)", r.stdout, re.M)
    assert outs[0] == outs[1]  # per-subset seam (np_mcmc.cpp:162) and batched sweeps: same trajectory


@pytest.mark.gpu
def test_cli_fix_q1_refits_parameters(tmp_path):
    """--fix-q1: UpdateClusters::update (np_mcmc.cpp:170) with a working parameter update (conjugate NIW posterior draw)."""
    ensure_built()
    data = tmp_path / "twogaussians.data"
    write_data(str(data))
    r = subprocess.run([CLI, "-d", str(data), "-a", "algorithm8", "-T", "60", "-c", "clustering", "--chains", "32", "--kmax", "64",
                        "--seed", "3", "--fix-q1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pur = float(re.search(r"chains: purity ([0-9.]+)", r.stdout).group(1))
    assert pur > 0.95


@pytest.mark.gpu
def test_cli_writes_results_like_the_reference(tmp_path):
    """--output: snapshot* / results* files, Octave parameter file, score file and LATEST symlink (np_results.cpp:39-196,
    np_main.cpp:476-497); the results are the max-likelihood state kept every 5 sweeps (np_mcmc.cpp:172-174,187-203)."""
    ensure_built()
    data = tmp_path / "twogaussians.data"
    write_data(str(data))
    out = tmp_path / "out"
    r = subprocess.run([CLI, "-d", str(data), "-a", "algorithm8", "-T", "100", "-c", "clustering", "--chains", "8", "--kmax", "64",
                        "--seed", "5", "--output", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ws = out / "algorithm8" / "twogaussians.data"
    latest = ws / "LATEST"
    assert latest.is_symlink()
    d = ws / os.readlink(str(latest))
    for base in ("snapshot", "results"):
        score = (d / (base + ".score.txt")).read_text()
        assert score.startswith("Purity: ") and "Adjusted Rand Index: " in score
        octave = (d / (base + ".txt")).read_text()
        assert "# name: mu" in octave and "# name: sigma" in octave and "# ndims: 3" in octave
        K = int(re.search(r"# rows: (\d+)", octave).group(1))
        rows = sum(len((d / ("%s%d.txt" % (base, k))).read_text().splitlines()) for k in range(K))
        assert rows <= 200 and rows > 150  # every item of a cluster that still has parameters is dumped once
    pur = float(re.search(r"Purity: ([0-9.]+)", (d / "results.score.txt").read_text()).group(1))
    assert pur > 0.95


@pytest.mark.gpu
def test_cli_membertrix_mutators_selftest(tmp_path):
    """membertrix::{addCluster, assign, retract, remove, clone} of the host mirror over the device state, with the reference's
    np_error_t codes: the reference's own unit test of the class (test/test_membertrix.cpp:75-91) through the single-item ABI
    calls npb_chain_move_item / npb_chain_move_item_new / npb_chain_remove_cluster."""
    ensure_built()
    data = tmp_path / "twogaussians.data"
    write_data(str(data))
    for seed in ("3", "11"):
        r = subprocess.run([CLI, "-d", str(data), "-a", "algorithm8", "-c", "clustering", "--seed", seed, "--selftest-membertrix"],
                           capture_output=True, text=True)
        assert r.returncode == 0 and "membertrix selftest passed" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cli_gpus_flag_shards_chains(tmp_path):
    """--gpus G: one host thread and object graph per device, chains split in contiguous blocks, scores gathered (here G = 1 and,
    where a second device exists, G = 2)."""
    ensure_built()
    data = tmp_path / "twogaussians.data"
    write_data(str(data))
    n_dev = len([ln for ln in subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.splitlines() if ln.startswith("GPU ")])
    for g in (1, 2):
        if g > n_dev:
            continue
        r = subprocess.run([CLI, "-d", str(data), "-a", "algorithm8", "-T", "60", "-c", "clustering", "--chains", "16", "--kmax", "64",
                            "--gpus", str(g)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        if g > 1:
            m = re.search(r"All 16 chains on 2 devices: purity ([0-9.]+)", r.stdout)
            assert m and float(m.group(1)) > 0.9, r.stdout
    r = subprocess.run([CLI, "-d", str(data), "-a", "algorithm8", "-T", "5", "-c", "clustering", "--chains", "2", "--gpus", "4"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "--gpus" in r.stderr
    # 16-D, Kmax = 32: the tcgen05 kernels (their shared-memory opt-in is per device) on every device of the process
    X, y = syn.gmm(1500, 16, 4, 12)
    d16 = tmp_path / "d16.data"
    with open(str(d16), "w") as f:
        for row, lab in zip(X, y):
            f.write(" ".join("%.9f" % v for v in row) + " %d\n" % lab)
    for g in (1, 2):
        if g > n_dev:
            continue
        r = subprocess.run([CLI, "-d", str(d16), "-a", "algorithm8", "-T", "6", "-c", "clustering", "--chains", "12", "--kmax", "32",
                            "--gpus", str(g)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("config", ["regression", "angular"])
def test_cli_runs_the_scalar_noise_families(tmp_path, config):
    """`-c regression` / `-c angular` (np_main.cpp:196-205): lines "a b label", two lines to find"""
    ensure_built()
    X, y = (syn.regression_lines(300, 2, 4) if config == "regression" else syn.angular_lines(300, 2, 4))
    data = tmp_path / "lines.data"
    with open(str(data), "w") as f:
        for row, lab in zip(X, y):
            f.write("%.9f %.9f %d\n" % (row[-2], row[-1], lab))
    r = subprocess.run([CLI, "-d", str(data), "-a", "algorithm8", "-T", "200", "-c", config, "--chains", "64", "--kmax", "64", "--seed", "3"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mean_pur = float(re.search(r"chains: purity ([0-9.]+)", r.stdout).group(1))
    assert mean_pur > 0.8, r.stdout
