"""GPU (pytest -m gpu): the ABI calls added in round 2 -- the incremental result copy (both of its modes), the kept state's own
clusters, the single-item membertrix calls with the np_error_t mirror codes, and the lifetime / dimension guards."""
import numpy as np
import pytest
import torch  # noqa: F401  (first: the library binds NCCL at run time and must find PyTorch's copy, not load the system's before it)

from noparama_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def test_sweep_host_delta_both_modes(npb, ctx):
    """npb_chains_sweep_host_delta keeps the caller's mirror equal to the device state: the first call copies everything, a
    mixing chain (more than a quarter of the entries change: 2-D, reference initialisation, first sweeps) takes the full copy,
    a settled chain the compacted list; sweep_host (full copy every time) is the cross-check."""
    X, _ = syn.config(1)
    ds = npb.Dataset(ctx, X)
    prior = npb.NormalInverseWishart(**syn.reference_prior(2))
    a = npb.MCMC(ctx, ds, prior, chains=16, Kmax=64, seed=4)
    b = npb.MCMC(ctx, ds, prior, chains=16, Kmax=64, seed=4)
    mirror = np.zeros((ds.N, 16), dtype=np.uint16)
    full = np.zeros((ds.N, 16), dtype=np.uint16)
    modes = set()
    for it in range(40):
        _, n = a.chains.sweep_host_delta(X, npb.ALG8, 1, z_mirror=mirror)
        b.chains.sweep_host(X, npb.ALG8, 1, z_out=full)
        assert np.array_equal(mirror, full), it
        assert np.array_equal(mirror.T.astype(np.int32), a.getMembershipMatrix())
        modes.add("all" if n == mirror.size else ("full" if n > mirror.size // 4 else "list"))
    assert "all" in modes and "list" in modes, modes
    # a converged high-D chain: next to nothing travels
    X16, y16 = syn.gmm(4000, 16, 8, 5)
    ds16 = npb.Dataset(ctx, X16)
    mc = npb.MCMC(ctx, ds16, npb.NormalInverseWishart(**syn.reference_prior(16)), chains=32, Kmax=32, seed=3)
    given = (np.stack([X16[y16 == k].mean(0) for k in range(8)]), np.tile(np.eye(16), (8, 1, 1)))
    mc.chains.init_from_params(*given)
    m16 = np.zeros((ds16.N, 32), dtype=np.uint16)
    counts = [mc.chains.sweep_host_delta(X16, npb.ALG8, 1, z_mirror=m16)[1]]  # first call: everything
    mc.chains.init_from_params(given[0][::-1].copy(), given[1])  # the same clusters in reverse slot order: every entry changes
    counts += [mc.chains.sweep_host_delta(X16, npb.ALG8, 1, z_mirror=m16)[1] for _ in range(3)]
    assert counts[0] == m16.size and counts[1] > m16.size // 4 and counts[3] < m16.size // 100, counts
    assert np.array_equal(m16.T.astype(np.int32), mc.getMembershipMatrix())
    ds.close()
    ds16.close()


def test_best_params_survive_slot_reuse(npb, ctx):
    """The kept (max-likelihood) state comes with ITS clusters: empty a slot after the snapshot and let another cluster take
    it -- npb_chains_get_best_params still reports the snapshot's parameters, npb_chains_get_params the new ones; a
    re-initialisation of the handle forgets the kept state."""
    X, _ = syn.config(1)
    ds = npb.Dataset(ctx, X)
    prior = npb.NormalInverseWishart(**syn.reference_prior(2))
    prior.bind(ctx)
    ch = npb.Chains(ctx, ds, 2, Kmax=32, K0=4, seed=9)
    mu = np.array([[0.0, 0.0], [5.0, 5.0]])
    Sig = np.tile(np.eye(2), (2, 1, 1))
    z = (np.arange(ds.N) % 2).astype(np.int32)
    ch.set_state(0, z, [0, 1], mu, Sig)
    ch.set_state(1, z, [0, 1], mu, Sig)
    cur, best = ch.consider_max_likelihood()
    slots0, counts0, mu0, _ = ch.best_params(0)
    assert list(slots0) == [0, 1] and np.allclose(mu0, mu, atol=1e-6)
    # every member of slot 1 moves to slot 0, then a new cluster (other parameters) takes the free slot 1
    for i in np.nonzero(z == 1)[0]:
        ch.move_item(0, int(i), 0)
    new_slot = ch.move_item_new(0, 0, np.array([9.0, -3.0]), np.eye(2) * 2.0)
    assert new_slot == 1
    slots_now, counts_now, mu_now, _ = ch.params(0)
    assert list(slots_now) == [0, 1] and list(counts_now) == [ds.N - 1, 1] and np.allclose(mu_now[1], [9.0, -3.0], atol=1e-6)
    slots_b, counts_b, mu_b, _ = ch.best_params(0)
    assert np.array_equal(counts_b, counts0) and np.allclose(mu_b, mu, atol=1e-6)  # the snapshot's clusters, not the current table's
    assert np.array_equal(ch.best_assignments(0, 1)[0], z)
    ch.init_from_params(mu, Sig)
    cur2, best2 = ch.consider_max_likelihood()
    assert np.array_equal(cur2, best2)  # the kept state was reset: the first state after a re-initialisation is kept
    ch.close()
    ds.close()


def test_move_item_status_codes_and_guards(npb, ctx):
    X, _ = syn.config(1)
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(2)).bind(ctx)
    ch = npb.Chains(ctx, ds, 1, Kmax=32, K0=4, seed=2)
    z = (np.arange(ds.N) % 2).astype(np.int32)
    ch.set_state(0, z, [0, 1], np.array([[0.0, 0.0], [5.0, 5.0]]), np.tile(np.eye(2), (2, 1, 1)))
    with pytest.raises(npb.NpbError) as e:
        ch.move_item(0, 0, 0)  # already there
    assert e.value.status == -16
    with pytest.raises(npb.NpbError) as e:
        ch.move_item(0, 0, 7)  # no such cluster
    assert e.value.status == -18
    assert ctx._lib.npb_chain_remove_cluster(ch._h, 0, 1) == -17  # members remaining
    assert ctx._lib.npb_chain_remove_cluster(ch._h, 0, 9) == 0
    # a dataset cannot be destroyed under its chains, and a prior of another dimension is refused at the next sweep
    assert ctx._lib.npb_dataset_destroy(ds._h) == -1
    npb.NormalInverseWishart(**syn.reference_prior(3)).bind(ctx)
    with pytest.raises(npb.NpbError):
        ch.sweep(npb.ALG8, 1)
    npb.NormalInverseWishart(**syn.reference_prior(2)).bind(ctx)
    ch.sweep(npb.ALG8, 1)
    ch.close()
    ds.close()


@pytest.mark.parametrize("chains,n_anchor", [(37, 150), (256, 1024), (1001, 333)])
def test_cocluster_counts_exact(npb, ctx, chains, n_anchor):
    """k_cc_gather + k_cc_tile (byte-compare of four chains per word, 64 x 64 tiles, upper triangle mirrored) against numpy:
    exact counts for chain counts that are not multiples of four and anchor counts that are not multiples of the tile; and the
    same through a one-rank NCCL communicator (npb_comm_create / ncclAllReduce bound at run time)."""
    X, _ = syn.config(1)
    X = np.tile(X, (8, 1))[:n_anchor + 57]
    ds = npb.Dataset(ctx, X)
    npb.NormalInverseWishart(**syn.reference_prior(2)).bind(ctx)
    ch = npb.Chains(ctx, ds, chains, Kmax=64, K0=20, seed=3)
    ch.sweep(npb.ALG8, 3)
    anchors = np.random.default_rng(1).permutation(len(X))[:n_anchor]
    z = ch.assignments()[:, anchors]  # [C, A]
    want = (z[:, :, None] == z[:, None, :]).sum(0).astype(np.float32)
    got = ch.cocluster(anchors)
    assert np.array_equal(got, want)
    assert np.all(np.diag(got) == chains)
    comm = npb.Comm(ctx, npb.Comm.unique_id(ctx), 0, 1)
    got2 = ch.cocluster(anchors, comm)
    comm.close()
    assert np.array_equal(got2, want)
    ch.close()
    ds.close()
