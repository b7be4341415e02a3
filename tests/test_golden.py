"""Golden vectors produced by executing the reference's own source (tests/golden/make_golden.py, run in the build
container where /root/reference exists) versus (a) the oracle, on CPU, and (b) the CUDA path, on the GPU box.

This is what pins the oracle: get_conn_padded content for Ising and LocalOperator (incl. the compaction/padding
semantics), the numba-packed lookup tables, log_cosh, local_value_kernel_jax, the block statistics, the exchange
clusters, the chain-length rounding.
"""

import os

import numpy as np
import pytest

import oracle
from oracle import graph as ograph
from oracle import hilbert as ohilbert
from oracle import operators as oops
from oracle import rbm as orbm

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))


# ============================================================================================ oracle (CPU)
def test_spin_index_map():
    assert np.array_equal(ohilbert.states_to_local_indices(G["spin_states"]), G["spin_indices"])


def test_log_cosh_golden():
    np.testing.assert_allclose(orbm.log_cosh(G["log_cosh_x"]), G["log_cosh_y"], rtol=1e-15, atol=1e-15)
    y32 = orbm.log_cosh(G["log_cosh_x"].astype(np.float32))
    assert y32.dtype == np.float32
    np.testing.assert_allclose(y32, G["log_cosh_y32"], rtol=3e-6, atol=1e-6)


@pytest.mark.parametrize("tag", ["ising1d", "ising2d", "ising_h0", "ising_negJ"])
def test_oracle_ising_conn_golden(tag):
    L, nd, h, J = G[f"{tag}_cfg"]
    edges, _ = ograph.hypercube_edges(int(L), int(nd))
    assert np.array_equal(edges, G[f"{tag}_edges"])
    xp, mels = oops.ising_conn_padded(G[f"{tag}_sigma"], edges, float(h), float(J))
    assert xp.dtype == np.int8 and np.array_equal(xp, G[f"{tag}_xp"])
    assert np.array_equal(mels, G[f"{tag}_mels"])
    assert np.array_equal(oops.ising_n_conn(G[f"{tag}_sigma"], edges, float(h), float(J)), G[f"{tag}_nconn"])


def _tables(tag):
    if tag == "heis1d":
        return oops.heisenberg_tables(G["heis1d_edges"], None, 1.0, True)
    if tag == "j1j2":
        return oops.heisenberg_tables(G["j1j2_edges"], G["j1j2_colors"], [1.0, 0.5], [False, False])
    ops = list(G["generic_mats1"]) + list(G["generic_mats2"])
    aon = [(i,) for i in range(5)] + [tuple(p) for p in G["generic_pairs"].tolist()]
    return oops.pack_internals(oops.canonical_operators_dict(ops, aon), 0.25)


def test_heisenberg_bond_matrices_golden():
    assert np.array_equal(oops.SZ_SZ, G["heis_sz_sz"]) and np.array_equal(oops.EXCHANGE, G["heis_exchange"])


@pytest.mark.parametrize("tag", ["heis1d", "j1j2", "generic"])
def test_oracle_packed_tables_golden(tag):
    """numba pack_internals(_jax) of the reference vs the oracle's restatement, entry by entry (NaN padding included)."""
    t = _tables(tag)
    assert t["max_conn_size"] == int(G[f"{tag}_K"]) and t["nonzero_diagonal"] == bool(G[f"{tag}_nonzero_diagonal"])
    for g in range(len(t["acting_on"])):
        for name in ("acting_on", "n_conns", "diag_mels", "x_prime", "mels", "basis"):
            np.testing.assert_array_equal(np.asarray(t[name][g], dtype=np.float64), np.asarray(G[f"{tag}_g{g}_{name}"], dtype=np.float64),
                                          err_msg=f"{tag} group {g} {name}")


@pytest.mark.parametrize("tag", ["heis1d", "j1j2", "generic"])
def test_oracle_localop_conn_golden(tag):
    xp, mels, nconn = oops.local_operator_conn_padded(G[f"{tag}_sigma"], _tables(tag))
    assert np.array_equal(xp, G[f"{tag}_xp"])
    assert np.array_equal(nconn, G[f"{tag}_nconn"])
    np.testing.assert_allclose(mels, G[f"{tag}_mels"], rtol=1e-15, atol=1e-15)


def test_oracle_local_value_kernel_golden():
    from oracle import estimators as oest

    W, b, a = G["eloc_W"], G["eloc_b"], G["eloc_a"]
    np.testing.assert_allclose(orbm.logpsi(G["eloc_sigma"], W, b, a), G["eloc_logpsi"], rtol=1e-13)
    e = oest.local_value_kernel(G["eloc_sigma"], lambda x: oops.ising_conn_padded(x, G["eloc_edges"], 3.0, 1.0), W, b, a)
    np.testing.assert_allclose(e, G["eloc_ising_h3"], rtol=1e-12)


@pytest.mark.parametrize("tag", ["stats_16x63", "stats_64x100", "stats_1x1000", "stats_33x65", "stats_40x1", "stats_5x7"])
def test_oracle_statistics_golden(tag):
    st = oracle.stats.statistics(G[f"{tag}_data"])
    got = np.array([st["mean"], st["error_of_mean"], st["variance"], st["tau_corr"], st["R_hat"]])
    np.testing.assert_allclose(got, G[f"{tag}_result"], rtol=1e-12, equal_nan=True)


def test_oracle_clusters_golden():
    e, _ = ograph.hypercube_edges(8, 1)
    assert np.array_equal(ograph.compute_clusters(8, e, 2), G["clusters_chain8_d2"])
    assert np.array_equal(ograph.distances(8, e), G["clusters_chain8_d2_dist"])
    e, _ = ograph.hypercube_edges(4, 2)
    assert np.array_equal(ograph.compute_clusters(16, e, 1), G["clusters_sq4_d1"])
    from oracle import sampler as osampler

    assert np.array_equal(osampler.hoppable_mask(G["clusters_mask_sigma"], G["clusters_chain8_d2"]), G["clusters_mask"])


def test_chain_length_golden():
    import warnings

    from netket_b200.vqs import compute_chain_length

    for nc, ns, cl in G["chain_length_cases"].tolist():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert compute_chain_length(nc, ns) == cl


# ============================================================================================ CUDA path (GPU box)
def _cuda(x):
    import torch

    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["ising1d", "ising2d", "ising_h0", "ising_negJ"])
def test_cuda_ising_conn_golden(cuda, tag):
    import netket_b200 as nk

    L, nd, h, J = G[f"{tag}_cfg"]
    g = nk.graph.Hypercube(int(L), int(nd))
    op = nk.operator.Ising(nk.hilbert.Spin(0.5, g.n_nodes), g, h=float(h), J=float(J))
    xp, mels = op.get_conn_padded(_cuda(G[f"{tag}_sigma"]))
    assert np.array_equal(xp.cpu().numpy(), G[f"{tag}_xp"])
    assert np.array_equal(mels.cpu().numpy(), G[f"{tag}_mels"])
    assert np.array_equal(op.n_conn(_cuda(G[f"{tag}_sigma"])).cpu().numpy(), G[f"{tag}_nconn"])


def _nk_localop(nk, tag):
    if tag == "heis1d":
        g = nk.graph.Chain(10)
        return nk.operator.Heisenberg(nk.hilbert.Spin(0.5, 10, total_sz=0), g)
    if tag == "j1j2":
        g = nk.graph.Hypercube(4, 2, max_neighbor_order=2)
        return nk.operator.Heisenberg(nk.hilbert.Spin(0.5, 16, total_sz=0), g, J=[1.0, 0.5])
    ops = list(G["generic_mats1"]) + list(G["generic_mats2"])
    aon = [[i] for i in range(5)] + G["generic_pairs"].tolist()
    return nk.operator.LocalOperator(nk.hilbert.Spin(0.5, 5), ops, aon, constant=0.25)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["heis1d", "j1j2", "generic"])
def test_cuda_localop_conn_golden(cuda, tag):
    import netket_b200 as nk

    op = _nk_localop(nk, tag)
    assert op.max_conn_size == int(G[f"{tag}_K"])
    xp, mels = op.get_conn_padded(_cuda(G[f"{tag}_sigma"]))
    assert np.array_equal(xp.cpu().numpy(), G[f"{tag}_xp"])
    assert np.array_equal(op.n_conn(_cuda(G[f"{tag}_sigma"])).cpu().numpy(), G[f"{tag}_nconn"])
    np.testing.assert_allclose(mels.cpu().numpy(), G[f"{tag}_mels"], rtol=1e-14, atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_cuda_logpsi_and_eloc_golden(cuda, dtype, tol):
    import torch

    import netket_b200 as nk

    g = nk.graph.Hypercube(4, 2)
    hi = nk.hilbert.Spin(0.5, 16)
    op = nk.operator.Ising(hi, g, h=3.0)
    var = {"params": {"Dense": {"kernel": _cuda(G["eloc_W"].astype(dtype)), "bias": _cuda(G["eloc_b"].astype(dtype))},
                      "visible_bias": _cuda(G["eloc_a"].astype(dtype))}}
    model = nk.models.RBM(alpha=2, param_dtype=dtype)
    lp = model.apply(var, _cuda(G["eloc_sigma"]))
    np.testing.assert_allclose(lp.cpu().numpy(), G["eloc_logpsi"], rtol=tol, atol=tol * np.abs(G["eloc_logpsi"]).max())
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=16), model, variables=var, n_samples=16, seed=0)
    for path in (1, 0):
        e = vs._eloc_on_samples(op, _cuda(G["eloc_sigma"]), path=path)
        np.testing.assert_allclose(e.cpu().numpy(), G["eloc_ising_h3"], rtol=tol, atol=tol * np.abs(G["eloc_ising_h3"]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["stats_16x63", "stats_64x100", "stats_1x1000", "stats_33x65", "stats_40x1", "stats_5x7"])
def test_cuda_statistics_golden(cuda, tag):
    import netket_b200 as nk

    st = nk.stats.statistics(_cuda(G[f"{tag}_data"]))
    got = np.array([st.mean, st.error_of_mean, st.variance, st.tau_corr, st.R_hat])
    np.testing.assert_allclose(got, G[f"{tag}_result"], rtol=1e-10, atol=1e-14, equal_nan=True)
