"""Host-side logic of the Python mirror that needs no GPU: graphs, Hilbert space, operator table packing
(against the oracle's restatement of compile_helpers.py), chain-length rounding, Stats."""

import math
import warnings

import numpy as np
import pytest

import netket_b200 as nk
from oracle import graph as ograph
from oracle import hilbert as ohilbert
from oracle import operators as oops


@pytest.mark.parametrize("L,n_dim,order", [(8, 1, 1), (10, 2, 1), (10, 2, 2), (4, 3, 1), (6, 1, 2), (5, 2, 2)])
def test_hypercube_edges(L, n_dim, order):
    g = nk.graph.Hypercube(L, n_dim, pbc=True, max_neighbor_order=order)
    e, c = ograph.hypercube_edges(L, n_dim, True, order)
    assert np.array_equal(np.asarray(g.edges()), e) and g.edge_colors == c.tolist()
    assert g.n_nodes == L ** n_dim and g.n_edges == len(e)
    assert np.array_equal(g.distances(), ograph.distances(g.n_nodes, e))
    assert g.is_bipartite() == ograph.is_bipartite(g.n_nodes, e)
    if order == 1:
        assert g.n_edges == n_dim * L ** n_dim
    assert [x[:2] for x in g.edges(return_color=True)] == g.edges()
    assert len(g.edges(filter_color=0)) == n_dim * L ** n_dim


def test_open_chain_and_generic_graph():
    g = nk.graph.Chain(5, pbc=False)
    assert g.edges() == [(0, 1), (1, 2), (2, 3), (3, 4)] and g.is_bipartite()
    tri = nk.graph.Graph([(0, 1), (1, 2), (2, 0)])
    assert not tri.is_bipartite() and tri.n_nodes == 3
    assert tri.distances()[0, 2] == 1


def test_spin_hilbert():
    hi = nk.hilbert.Spin(0.5, 6)
    assert hi.size == 6 and hi.n_states == 64 and not hi.constrained and hi.n_down == -1
    assert np.array_equal(hi.all_states(), ohilbert.all_states(6))
    assert np.array_equal(hi.states_to_local_indices(np.array([1, -1])), [0, 1])
    assert np.array_equal(hi.local_indices_to_states(np.array([0, 1])), [1, -1])
    hc = nk.hilbert.Spin(0.5, 6, total_sz=1)
    assert hc.constrained and hc.n_down == 2 and hc.n_states == math.comb(6, 2)
    assert np.array_equal(hc.all_states(), ohilbert.all_states(6, total_sz=1))
    with pytest.raises(ValueError):
        nk.hilbert.Spin(0.5, 5, total_sz=0)
    with pytest.raises(NotImplementedError):
        nk.hilbert.Spin(1.0, 4)
    assert nk.hilbert.Spin(0.5, 4) == nk.hilbert.Spin(0.5, 4) != nk.hilbert.Spin(0.5, 4, total_sz=0)


@pytest.mark.parametrize("J,sign_rule,order,L,n_dim", [(1.0, None, 1, 10, 1), ([1.0, 0.5], None, 2, 10, 2), ([1.0, 2.0], [True, False], 2, 6, 1),
                                                        (0.75, False, 1, 4, 2)])
def test_heisenberg_tables_match_oracle_packing(J, sign_rule, order, L, n_dim):
    g = nk.graph.Hypercube(L, n_dim, max_neighbor_order=order)
    hi = nk.hilbert.Spin(0.5, g.n_nodes)
    op = nk.operator.Heisenberg(hi, g, J=J, sign_rule=sign_rule)
    e, c = ograph.hypercube_edges(L, n_dim, True, order)
    sr = sign_rule if sign_rule is not None else ([False] * len(J) if isinstance(J, list) else ograph.is_bipartite(g.n_nodes, e))
    ref = oops.heisenberg_tables(e, c, J, sr)
    t = op._pack()
    assert t["max_conn_size"] == ref["max_conn_size"] and t["nonzero_diagonal"] == ref["nonzero_diagonal"]
    assert len(t["groups"]) == len(ref["acting_on"]) == 1
    G = t["groups"][0]
    assert np.array_equal(G["acting_on"], ref["acting_on"][0])
    assert np.array_equal(G["diag_mels"], ref["diag_mels"][0])
    assert np.array_equal(G["n_conns"], ref["n_conns"][0])
    nc = ref["n_conns"][0]
    for o in range(G["n_ops"]):
        for r in range(4):
            k = nc[o, r]
            assert np.array_equal(G["mels"][o, r, :k], ref["mels"][0][o, r, :k])
            assert np.array_equal(G["x_prime"][o, r, :k], ref["x_prime"][0][o, r, :k].astype(np.int8))
    assert op.is_hermitian and op.max_conn_size == 1 + g.n_edges


def test_local_operator_canonicalisation_and_errors():
    hi = nk.hilbert.Spin(0.5, 4)
    rs = np.random.default_rng(1)
    A = rs.normal(size=(4, 4))
    op = nk.operator.LocalOperator(hi, [A, A], [[3, 1], [1, 3]])
    ref = oops.canonical_operators_dict([A, A], [(3, 1), (1, 3)])
    assert op.acting_on == [(1, 3)]
    np.testing.assert_allclose(op.operators[0], ref[(1, 3)])
    with pytest.raises(ValueError, match="invalid set of sites"):
        nk.operator.LocalOperator(hi, [A], [[0, 4]])
    with pytest.raises(ValueError, match="duplicated"):
        nk.operator.LocalOperator(hi, [A], [[1, 1]])
    with pytest.raises(ValueError, match="must have shape"):
        nk.operator.LocalOperator(hi, [np.eye(2)], [[0, 1]])
    with pytest.raises(NotImplementedError):
        nk.operator.LocalOperator(hi, [np.eye(8)], [[0, 1, 2]])
    empty = nk.operator.LocalOperator(hi)
    assert empty.max_conn_size == 0 and empty.n_operators == 0
    ident = nk.operator.LocalOperator(hi, constant=2.0)
    assert ident.max_conn_size == 1
    with pytest.raises(ValueError, match="non-bipartite"):
        nk.operator.Heisenberg(nk.hilbert.Spin(0.5, 3), nk.graph.Graph([(0, 1), (1, 2), (2, 0)]), sign_rule=True)


def test_ising_host_properties():
    g = nk.graph.Hypercube(10, 2)
    hi = nk.hilbert.Spin(0.5, 100)
    op = nk.operator.Ising(hi, g, h=3.0)
    assert op.max_conn_size == 101 and op.dtype == np.float64 and op.is_hermitian
    assert nk.operator.Ising(hi, g, h=0.0).max_conn_size == 1
    assert nk.operator.Ising(hi, g, h=1.0, dtype=np.float32).dtype == np.float32
    with pytest.raises(ValueError):
        nk.operator.Ising(nk.hilbert.Spin(0.5, 50), g, h=1.0)


def test_compute_chain_length_rounds_up_with_warning():
    """state.py:60-79 and test/variational/test_variational.py:182-187: 16 chains, 1000 samples -> 1008, chain_length 63."""
    from netket_b200.vqs import compute_chain_length

    with pytest.warns(UserWarning, match="increased to 1008"):
        assert compute_chain_length(16, 1000) == 63
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert compute_chain_length(16, 1008) == 63
    with pytest.raises(ValueError):
        compute_chain_length(16, 0)


def test_sampler_construction_rules():
    hi = nk.hilbert.Spin(0.5, 8)
    sa = nk.sampler.MetropolisLocal(hi)
    assert sa.n_chains == sa.n_chains_per_rank == 16 and sa.sweep_size == 8 and sa.machine_pow == 2.0 and not sa.is_exact
    assert nk.sampler.MetropolisLocal(hi, n_chains=24, sweep_size=3).sweep_size == 3
    with pytest.raises(ValueError):
        nk.sampler.MetropolisLocal(hi, machine_pow=-1)
    with pytest.raises(ValueError):
        nk.sampler.MetropolisLocal(hi, n_chains=4, n_chains_per_rank=4)
    with pytest.raises(TypeError):
        nk.sampler.MetropolisSampler(hi, "local")
    with pytest.raises(TypeError):
        nk.sampler.MetropolisLocal(hi, reset_chains=1)
    g = nk.graph.Chain(8)
    ex = nk.sampler.MetropolisExchange(hi, graph=g, d_max=2)
    assert np.array_equal(ex.rule.clusters, ograph.compute_clusters(8, np.asarray(g.edges()), 2)) and len(ex.rule.clusters) == 16
    with pytest.raises(ValueError):
        nk.sampler.MetropolisExchange(hi)
    # ExchangeRule(probabilities=...): one weight per graph distance, or per cluster (rules/exchange.py:86-123,200-203)
    wr = nk.sampler.ExchangeRule(graph=g, d_max=2, probabilities=[0.7, 0.3])
    D = np.asarray(g.distances())
    assert np.array_equal(wr.probabilities, np.array([0.7, 0.3])[D[wr.clusters[:, 0], wr.clusters[:, 1]] - 1])
    assert nk.sampler.MetropolisExchange(hi, clusters=[(0, 1), (2, 5)], probabilities=[1.0, 2.0]).rule.probabilities.tolist() == [1.0, 2.0]
    with pytest.raises(ValueError, match="positive"):
        nk.sampler.ExchangeRule(graph=g, probabilities=[0.0])
    with pytest.raises(TypeError, match="don't match"):
        nk.sampler.ExchangeRule(clusters=[(0, 1), (2, 5)], probabilities=[1.0])
    assert "MetropolisSampler" in repr(sa) and "ExchangeRule(# of clusters: 16)" in repr(ex)


def test_stats_object():
    st = nk.stats.Stats(mean=-10.25, error_of_mean=0.013, variance=0.17, tau_corr=0.4, R_hat=1.003)
    assert st.Mean == st.mean and st.Sigma == st.error_of_mean and st.R == st.R_hat and st["variance"] == 0.17
    assert st.to_dict() == {"Mean": -10.25, "Variance": 0.17, "Sigma": 0.013, "R_hat": 1.003, "TauCorr": 0.4}
    assert repr(st) == "-10.250 ± 0.013 [σ²=0.17, R̂=1.003]"
    assert "nan" in repr(nk.stats.Stats(mean=1.0)).lower()
    with pytest.raises(AttributeError):
        st.nope


def test_no_cpu_fallback_for_numerics():
    """Without a GPU every numerical entry point raises instead of silently computing on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    hi = nk.hilbert.Spin(0.5, 4)
    with pytest.raises(nk.NkError):
        hi.random_state(1, 4)
    with pytest.raises(nk.NkError):
        nk.operator.Ising(hi, nk.graph.Chain(4), h=1.0).get_conn_padded(np.ones((2, 4)))
    with pytest.raises(nk.NkError):
        nk.stats.statistics(np.zeros((4, 4)))


def test_msgpack_wire_format_of_state_dicts():
    """The bytes of `flax.serialization.to_bytes` as restated in netket_b200/serialization.py: a hand-assembled byte string in
    that format (ExtType 1 = packb((shape, dtype name, C bytes)), ExtType 3 for NumPy scalars, ExtType 2 for complex) decodes to the
    expected tree, and a tree of arrays / scalars / None round-trips bit for bit."""
    import msgpack

    from netket_b200 import serialization as ser

    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    ext_a = msgpack.ExtType(1, msgpack.packb(((2, 3), "float32", a.tobytes()), use_bin_type=True))
    ext_s = msgpack.ExtType(3, msgpack.packb(((), "int64", np.int64(7).tobytes()), use_bin_type=True))
    ext_c = msgpack.ExtType(2, msgpack.packb((1.5, -2.0)))
    blob = msgpack.packb({"variables": {"params": {"Dense": {"kernel": ext_a}}}, "n_samples": 1008, "k": ext_s, "z": ext_c, "chunk_size": None})
    tree = ser.msgpack_restore(blob)
    assert np.array_equal(tree["variables"]["params"]["Dense"]["kernel"], a) and tree["variables"]["params"]["Dense"]["kernel"].dtype == np.float32
    assert tree["n_samples"] == 1008 and tree["k"] == 7 and isinstance(tree["k"], np.int64) and tree["z"] == complex(1.5, -2.0)
    assert tree["chunk_size"] is None
    # our writer produces exactly those bytes for the same tree
    same = ser.msgpack_serialize({"variables": {"params": {"Dense": {"kernel": a}}}, "n_samples": 1008, "k": np.int64(7), "z": complex(1.5, -2.0),
                                  "chunk_size": None})
    assert same == blob
    big = {"sampler_state": {"σ": np.random.default_rng(0).choice([-1, 1], size=(5, 7)).astype(np.int8), "rng": np.array([3, 9], dtype=np.uint64),
                             "n_accepted_proc": np.arange(5, dtype=np.int64)}, "w": np.random.default_rng(1).normal(size=(3, 4))}
    back = ser.msgpack_restore(ser.msgpack_serialize(big))
    for k, v in big["sampler_state"].items():
        assert np.array_equal(back["sampler_state"][k], v) and back["sampler_state"][k].dtype == v.dtype
    assert np.array_equal(back["w"], big["w"])
    with pytest.raises(ValueError, match="object arrays"):
        ser.msgpack_serialize({"x": np.array([object()])})
