"""Parity of the CUDA path (through the C ABI, via the Python mirror) against the CPU oracle.

Tolerances (BASELINE.json north_star): connected elements bit-exact; logpsi / E_loc 1e-12 relative in fp64 and
1e-5 in fp32; fixed-proposal-stream chains identical to the oracle's.
"""

import numpy as np
import pytest
import torch

import oracle
from oracle import estimators as oest
from oracle import graph as ograph
from oracle import hilbert as ohilbert
from oracle import operators as oops
from oracle import rbm as orbm
from oracle import rng as orng
from oracle import sampler as osampler

pytestmark = pytest.mark.gpu

RTOL = {np.float64: 1e-12, np.float32: 1e-5}


def _nk():
    import netket_b200 as nk

    return nk


def _params(N, alpha, dtype, std=0.01, seed=1234, hidden_bias=True, visible_bias=True):
    W, b, a = orbm.init_params(N, alpha, seed=seed, std=std, dtype=dtype, use_hidden_bias=hidden_bias,
                               use_visible_bias=visible_bias)
    dense = {"kernel": torch.from_numpy(W).cuda()}
    if b is not None:
        dense["bias"] = torch.from_numpy(b).cuda()
    p = {"Dense": dense}
    if a is not None:
        p["visible_bias"] = torch.from_numpy(a).cuda()
    return (W, b, a), {"params": p}


def _sigma(B, N, seed=0, total_sz=None):
    return ohilbert.random_state(seed, B, N, total_sz)


# ----------------------------------------------------------------------------------------- RBM.apply
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("N,alpha,std", [(20, 1, 0.01), (100, 4, 0.01), (100, 4, 0.1), (22, 2, 0.5), (7, 3, 1.0), (130, 1.5, 0.05)])
def test_logpsi(cuda, dtype, N, alpha, std):
    nk = _nk()
    (W, b, a), var = _params(N, alpha, dtype, std)
    sig = _sigma(257, N, seed=3)
    model = nk.models.RBM(alpha=alpha, param_dtype=dtype)
    out, theta = model.apply(var, torch.from_numpy(sig).cuda(), return_theta=True)
    ref = orbm.logpsi(sig, W.astype(np.float64), b.astype(np.float64), a.astype(np.float64))
    ref_theta = orbm.theta(sig, W.astype(np.float64), b.astype(np.float64))
    scale = np.abs(ref).max()
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=RTOL[dtype], atol=RTOL[dtype] * scale)
    np.testing.assert_allclose(theta.cpu().numpy(), ref_theta, rtol=RTOL[dtype], atol=RTOL[dtype] * np.abs(ref_theta).max())
    assert out.dtype == (torch.float64 if dtype == np.float64 else torch.float32)


@pytest.mark.parametrize("hb,vb", [(False, True), (True, False), (False, False)])
def test_logpsi_bias_flags(cuda, hb, vb):
    nk = _nk()
    (W, b, a), var = _params(12, 2, np.float64, 0.3, hidden_bias=hb, visible_bias=vb)
    sig = _sigma(33, 12)
    out = nk.models.RBM(alpha=2, use_hidden_bias=hb, use_visible_bias=vb).apply(var, torch.from_numpy(sig).cuda())
    np.testing.assert_allclose(out.cpu().numpy(), orbm.logpsi(sig, W, b, a), rtol=1e-12)


def test_logpsi_batch_shapes_and_empty(cuda):
    nk = _nk()
    (W, b, a), var = _params(10, 1, np.float64, 0.2)
    m = nk.models.RBM(alpha=1)
    sig = _sigma(24, 10).reshape(2, 3, 4, 10)
    out = m.apply(var, torch.from_numpy(sig).cuda())
    assert tuple(out.shape) == (2, 3, 4)
    np.testing.assert_allclose(out.cpu().numpy(), orbm.logpsi(sig, W, b, a), rtol=1e-12)
    assert tuple(m.apply(var, torch.zeros((0, 10), dtype=torch.int8, device="cuda")).shape) == (0,)
    one = m.apply(var, torch.from_numpy(sig[0, 0, 0]).cuda())
    assert one.ndim == 0


# ----------------------------------------------------------------------------------------- get_conn_padded
@pytest.mark.parametrize("L,n_dim,h,J", [(10, 1, 1.321, 1.0), (20, 1, 1.0, 1.0), (10, 2, 3.0, 1.0), (5, 1, 0.7, -0.5), (4, 2, 0.0, 1.0),
                                          (3, 3, 2.0, 0.25)])
def test_ising_conn_bit_exact(cuda, L, n_dim, h, J):
    nk = _nk()
    g = nk.graph.Hypercube(L, n_dim, pbc=True)
    hi = nk.hilbert.Spin(0.5, g.n_nodes)
    op = nk.operator.Ising(hi, g, h=h, J=J)
    e, _ = ograph.hypercube_edges(L, n_dim)
    assert np.array_equal(op.edges, e)
    sig = _sigma(301, g.n_nodes, seed=5)
    xp, mels = op.get_conn_padded(torch.from_numpy(sig).cuda())
    rxp, rmels = oops.ising_conn_padded(sig, e, h, J)
    assert xp.dtype == torch.int8 and mels.dtype == torch.float64
    assert tuple(xp.shape) == rxp.shape and op.max_conn_size == rxp.shape[1]
    assert np.array_equal(xp.cpu().numpy(), rxp)
    assert np.array_equal(mels.cpu().numpy(), rmels)
    nc = op.n_conn(torch.from_numpy(sig).cuda())
    assert np.array_equal(nc.cpu().numpy(), oops.ising_n_conn(sig, e, h, J))


def test_ising_conn_contract(cuda):
    """rank/dtype contract of test/operator/test_operator.py:345-359 and numpy round trip."""
    nk = _nk()
    g = nk.graph.Chain(6)
    hi = nk.hilbert.Spin(0.5, 6)
    op = nk.operator.Ising(hi, g, h=1.0, dtype=np.float32)
    for shape in [(6,), (1, 6), (2, 6), (2, 3, 6)]:
        v = np.ones(shape, dtype=np.float64)
        vp, mels = op.get_conn_padded(v)
        assert vp.ndim == v.ndim + 1 and mels.ndim == v.ndim
        assert vp.dtype == v.dtype and mels.dtype == op.dtype
    xp, mels = op.get_conn_padded(torch.zeros((0, 6), dtype=torch.int8, device="cuda"))
    assert tuple(xp.shape) == (0, 7, 6) and tuple(mels.shape) == (0, 7)


def _heis(nk, L, n_dim, total_sz, J, sign_rule, order=1):
    g = nk.graph.Hypercube(L, n_dim, pbc=True, max_neighbor_order=order)
    hi = nk.hilbert.Spin(0.5, g.n_nodes, total_sz=total_sz)
    op = nk.operator.Heisenberg(hi, g, J=J, sign_rule=sign_rule)
    e, c = ograph.hypercube_edges(L, n_dim, max_neighbor_order=order)
    sr = sign_rule
    if sr is None:
        sr = [False] * len(J) if isinstance(J, (list, tuple)) else ograph.is_bipartite(g.n_nodes, e)
    tables = oops.heisenberg_tables(e, c, J=J, sign_rule=sr)
    return g, hi, op, tables


@pytest.mark.parametrize("L,n_dim,total_sz,J,sign_rule,order", [
    (10, 1, 0, 1.0, None, 1), (22, 1, 0, 1.0, None, 1), (4, 2, None, 1.0, False, 1), (10, 2, 0, [1.0, 0.5], None, 2),
    (6, 1, None, [1.0, 2.0], [True, False], 2), (5, 1, None, 0.75, False, 1)])
def test_localop_conn_bit_exact(cuda, L, n_dim, total_sz, J, sign_rule, order):
    nk = _nk()
    g, hi, op, tables = _heis(nk, L, n_dim, total_sz, J, sign_rule, order)
    assert op.max_conn_size == tables["max_conn_size"]
    sig = _sigma(203, g.n_nodes, seed=11, total_sz=total_sz)
    sig[0] = 1  # all-up: zero off-diagonal elements, non-zero diagonal
    xp, mels = op.get_conn_padded(torch.from_numpy(sig).cuda())
    nconn = op.n_conn(torch.from_numpy(sig).cuda())
    rxp, rmels, rn = oops.local_operator_conn_padded(sig, tables)
    assert np.array_equal(xp.cpu().numpy(), rxp)
    assert np.array_equal(nconn.cpu().numpy(), rn)
    # matrix elements: entries are +-J, +-2J; the diagonal is a sum of +-J_b whose order differs from numpy's
    np.testing.assert_allclose(mels.cpu().numpy(), rmels, rtol=1e-14, atol=1e-13)
    m = mels.cpu().numpy()
    for r in range(m.shape[0]):  # padding must be trailing zeros (test/operator/test_operator.py:507-556)
        k = int(nconn[r])
        assert np.all(m[r, :k] != 0) and np.all(m[r, k:] == 0)
        assert np.array_equal(xp[r, k:].cpu().numpy(), np.broadcast_to(sig[r], (m.shape[1] - k, sig.shape[1])))


def test_localop_site_and_bond_terms(cuda):
    """GraphOperator with site_ops (1-site group) + bond_ops (2-site group), unsorted supports, summed duplicates."""
    nk = _nk()
    N = 6
    hi = nk.hilbert.Spin(0.5, N)
    rs = np.random.default_rng(0)
    ops, aon = [], []
    for i in range(N):
        m = rs.normal(size=(2, 2)); ops.append(m + m.T); aon.append([i])
    for (i, j) in [(0, 1), (2, 1), (5, 3), (1, 0), (4, 5)]:
        m = rs.normal(size=(4, 4)); m[np.abs(m) < 0.4] = 0.0; ops.append(m + m.T); aon.append([i, j])
    op = nk.operator.LocalOperator(hi, ops, aon, constant=0.3)
    tables = oops.pack_internals(oops.canonical_operators_dict(ops, aon), 0.3)
    sig = ohilbert.all_states(N)
    xp, mels = op.get_conn_padded(torch.from_numpy(sig).cuda())
    rxp, rmels, rn = oops.local_operator_conn_padded(sig, tables)
    assert np.array_equal(xp.cpu().numpy(), rxp)
    np.testing.assert_allclose(mels.cpu().numpy(), rmels, rtol=1e-14, atol=1e-14)
    assert np.array_equal(op.n_conn(torch.from_numpy(sig).cuda()).cpu().numpy(), rn)
    # dense matrix equals the Kronecker construction
    H = op.to_dense()
    Hk = oops.kron_dense(N, [(a[0], m) for m, a in zip(ops, aon) if len(a) == 1],
                         [(min(a), max(a), m if a[0] < a[1] else oops._reorder_kronecker_product(m, a)[0])
                          for m, a in zip(ops, aon) if len(a) == 2]) + 0.3 * np.eye(2 ** N)
    np.testing.assert_allclose(H, Hk, atol=1e-13)


# ----------------------------------------------------------------------------------------- E_loc
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("L,n_dim,alpha,std,h", [(20, 1, 1, 0.01, 1.0), (10, 2, 4, 0.01, 3.0), (10, 2, 4, 0.1, 3.0), (4, 2, 2, 0.5, 0.5),
                                                  (6, 1, 1, 0.3, 0.0)])
def test_eloc_ising(cuda, dtype, L, n_dim, alpha, std, h):
    nk = _nk()
    g = nk.graph.Hypercube(L, n_dim)
    N = g.n_nodes
    hi = nk.hilbert.Spin(0.5, N)
    op = nk.operator.Ising(hi, g, h=h)
    (W, b, a), var = _params(N, alpha, dtype, std)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=16), nk.models.RBM(alpha=alpha, param_dtype=dtype),
                        variables=var, n_samples=16, seed=1)
    sig = _sigma(128, N, seed=9)
    out = vs._eloc_on_samples(op, torch.from_numpy(sig).cuda(), path=1)
    W64, b64, a64 = W.astype(np.float64), b.astype(np.float64), a.astype(np.float64)
    e, _ = ograph.hypercube_edges(L, n_dim)
    ref = oest.local_value_kernel(sig, lambda x: oops.ising_conn_padded(x, e, h, 1.0), W64, b64, a64)
    assert out.dtype == torch.float64  # promote(operator float64, params)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=RTOL[dtype], atol=RTOL[dtype] * np.abs(ref).max())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("L,n_dim,total_sz,J,sign_rule,order,alpha,std", [
    (22, 1, 0, 1.0, None, 1, 2, 0.01), (10, 2, 0, [1.0, 0.5], None, 2, 4, 0.05), (4, 2, None, 1.0, True, 1, 1, 0.4)])
def test_eloc_heisenberg(cuda, dtype, L, n_dim, total_sz, J, sign_rule, order, alpha, std):
    nk = _nk()
    g, hi, op, tables = _heis(nk, L, n_dim, total_sz, J, sign_rule, order)
    N = g.n_nodes
    (W, b, a), var = _params(N, alpha, dtype, std)
    vs = nk.vqs.MCState(nk.sampler.MetropolisExchange(hi, graph=g, n_chains=16), nk.models.RBM(alpha=alpha, param_dtype=dtype),
                        variables=var, n_samples=16, seed=1)
    sig = _sigma(64, N, seed=9, total_sz=total_sz)
    out = vs._eloc_on_samples(op, torch.from_numpy(sig).cuda())
    ref = oest.local_value_kernel(sig, lambda x: oops.local_operator_conn_padded(x, tables), W.astype(np.float64),
                                  b.astype(np.float64), a.astype(np.float64))
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=RTOL[dtype], atol=RTOL[dtype] * np.abs(ref).max())


# ----------------------------------------------------------------------------------------- random_state
@pytest.mark.parametrize("N,total_sz", [(20, None), (100, None), (130, None), (22, 0), (100, 0), (9, 1.5), (8, -2)])
def test_random_state_matches_oracle(cuda, N, total_sz):
    nk = _nk()
    hi = nk.hilbert.Spin(0.5, N, total_sz=total_sz)
    out = hi.random_state(15324, 64, chain_offset=7).cpu().numpy()
    ref = ohilbert.random_state(15324, 64, N, total_sz, chain_offset=7)
    assert out.dtype == np.int8 and np.array_equal(out, ref)
    assert set(np.unique(out)) <= {-1, 1}
    if total_sz is not None:
        assert np.all(out.astype(int).sum(axis=1) == round(2 * total_sz))


# ----------------------------------------------------------------------------------------- sweeps (chain reproduction)
def _sampler_case(nk, rule, N, alpha, dtype, std, B, L=None, n_dim=1, total_sz=None, sweep_size=None):
    L = L or N
    g = nk.graph.Hypercube(L, n_dim)
    hi = nk.hilbert.Spin(0.5, g.n_nodes, total_sz=total_sz)
    (W, b, a), var = _params(g.n_nodes, alpha, dtype, std)
    model = nk.models.RBM(alpha=alpha, param_dtype=dtype)
    if rule == "local":
        sa = nk.sampler.MetropolisLocal(hi, n_chains=B, sweep_size=sweep_size)
        clusters = None
    else:
        sa = nk.sampler.MetropolisExchange(hi, graph=g, d_max=2, n_chains=B, sweep_size=sweep_size)
        e, _ = ograph.hypercube_edges(L, n_dim)
        clusters = ograph.compute_clusters(g.n_nodes, e, 2)
        assert np.array_equal(clusters, sa.rule.clusters)
    return g, hi, (W, b, a), var, model, sa, clusters


@pytest.mark.parametrize("rule,N,alpha,std,total_sz", [("local", 20, 1, 0.01, None), ("local", 20, 1, 0.3, None), ("local", 16, 2, 1.0, None),
                                                        ("exchange", 22, 2, 0.01, 0), ("exchange", 12, 2, 0.5, 0), ("exchange", 10, 1, 0.4, 1)])
def test_sweep_reproduces_oracle_chain_fp64(cuda, rule, N, alpha, std, total_sz):
    """In-kernel Philox stream and explicit-stream mode both reproduce the oracle's chains bit for bit (fp64)."""
    nk = _nk()
    B, CL = 24, 3
    g, hi, (W, b, a), var, model, sa, clusters = _sampler_case(nk, rule, N, alpha, np.float64, std, B, total_sz=total_sz)
    st = sa.init_state(model, var, seed=15324)
    sig0 = st.σ.cpu().numpy()
    seed, t0 = st.rng
    ref = osampler.sample_chain(rule, sig0, W, b, a, chain_length=CL, seed=seed, t0=t0, clusters=clusters)
    (samples, logp), st2 = sa.sample(model, var, state=st, chain_length=CL, return_log_probabilities=True, _path=1)
    assert np.array_equal(samples.cpu().numpy(), ref["samples"])
    np.testing.assert_allclose(logp.cpu().numpy(), ref["log_prob_samples"], rtol=1e-10, atol=1e-10)
    assert np.array_equal(st2.n_accepted_proc.cpu().numpy(), ref["n_accepted"])
    assert st2.n_steps_proc == ref["n_steps"] and st2.rng == (seed, ref["t"])
    assert np.array_equal(st.σ.cpu().numpy(), sig0), "input state must not be mutated"
    # continuing from the new state continues the same Philox stream
    ref2 = osampler.sample_chain(rule, ref["sigma"], W, b, a, chain_length=2, seed=seed, t0=ref["t"], clusters=clusters)
    s2, st3 = sa.sample(model, var, state=st2, chain_length=2, _path=1)
    assert np.array_equal(s2.cpu().numpy(), ref2["samples"])
    # explicit proposal stream
    T = CL * sa.sweep_size
    rs = np.random.default_rng(5)
    w0 = rs.integers(0, 2 ** 32, size=(T, B), dtype=np.uint64).astype(np.uint32)
    u = rs.random((T, B))
    ref3 = osampler.sample_chain(rule, sig0, W, b, a, chain_length=CL, stream=(w0, u), clusters=clusters)
    s3, _ = sa.sample(model, var, state=st, chain_length=CL, _stream=(w0, u), _path=1)
    assert np.array_equal(s3.cpu().numpy(), ref3["samples"])


@pytest.mark.parametrize("rule,N,alpha,std,total_sz", [("local", 20, 1, 0.3, None), ("exchange", 12, 2, 0.5, 0)])
def test_sweep_reproduces_oracle_chain_fp32(cuda, rule, N, alpha, std, total_sz):
    """fp32: identical up to accept-boundary ties (decisions with |u - exp(arg)| ~ 1e-7): >= 90% of chains identical."""
    nk = _nk()
    B, CL = 64, 2
    g, hi, (W, b, a), var, model, sa, clusters = _sampler_case(nk, rule, N, alpha, np.float32, std, B, total_sz=total_sz)
    st = sa.init_state(model, var, seed=99)
    seed, t0 = st.rng
    ref = osampler.sample_chain(rule, st.σ.cpu().numpy(), W, b, a, chain_length=CL, seed=seed, t0=t0, clusters=clusters)
    samples, st2 = sa.sample(model, var, state=st, chain_length=CL, _path=1)
    same = np.all(samples.cpu().numpy() == ref["samples"], axis=(1, 2))
    assert same.mean() >= 0.9, same.mean()


def test_sweep_size_and_discard(cuda):
    nk = _nk()
    g, hi, (W, b, a), var, model, sa, _ = _sampler_case(nk, "local", 10, 1, np.float64, 0.3, 8, sweep_size=7)
    st = sa.init_state(model, var, seed=1)
    seed, t0 = st.rng
    ref = osampler.sample_chain("local", st.σ.cpu().numpy(), W, b, a, chain_length=5, sweep_size=7, seed=seed, t0=t0)
    samples, _, _, st2 = sa._launch(model, var, st, 3, n_discard=2, path=1)
    assert np.array_equal(samples.cpu().numpy(), ref["samples"][:, 2:, :])
    assert np.array_equal(st2.n_accepted_proc.cpu().numpy(), ref["n_accepted"])
    assert st2.n_steps_proc == 8 * 5 * 7


def test_sharded_chains_equal_single_device_run(cuda):
    """chain_offset: chains [4, 12) of a 12-chain run are reproduced by a rank that owns only those."""
    nk = _nk()
    g, hi, (W, b, a), var, model, sa, _ = _sampler_case(nk, "local", 12, 1, np.float64, 0.3, 12)
    st = sa.init_state(model, var, seed=4)
    full, _ = sa.sample(model, var, state=st, chain_length=2, _path=1)
    sa8 = nk.sampler.MetropolisLocal(hi, n_chains=8)
    st8 = sa8.init_state(model, var, seed=4)
    st8 = st8.replace(σ=hi.random_state(nk.utils.mix_seed(4, 0), 8, chain_offset=4), chain_offset=4)
    part, _ = sa8.sample(model, var, state=st8, chain_length=2, _path=1)
    assert np.array_equal(part.cpu().numpy(), full[4:12].cpu().numpy())


# ----------------------------------------------------------------------------------------- fused sweep + E_loc
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_eloc_equals_standalone_and_oracle(cuda, dtype):
    nk = _nk()
    g = nk.graph.Hypercube(4, 2)
    hi = nk.hilbert.Spin(0.5, 16)
    op = nk.operator.Ising(hi, g, h=3.0)
    (W, b, a), var = _params(16, 4, dtype, 0.1)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=32), nk.models.RBM(alpha=4, param_dtype=dtype), variables=var,
                        n_samples=32 * 6, n_discard_per_chain=3, sampler_seed=3)
    eloc_fused = vs.local_estimators(op)          # no cached samples -> fused launch
    samples = vs.samples
    assert tuple(eloc_fused.shape) == (32, 6) and tuple(samples.shape) == (32, 6, 16)
    eloc_alone = vs._eloc_on_samples(op, samples, path=1)
    e, _ = ograph.hypercube_edges(4, 2)
    ref = oest.local_estimators(samples.cpu().numpy(), lambda x: oops.ising_conn_padded(x, e, 3.0, 1.0), W.astype(np.float64),
                                b.astype(np.float64), a.astype(np.float64))
    from tolerances import F32_TOL, F64_TOL, assert_rel

    tol = F64_TOL if dtype == np.float64 else F32_TOL
    assert_rel(eloc_alone.cpu().numpy(), ref, tol, "stand-alone E_loc (theta form)")
    assert_rel(eloc_fused.cpu().numpy(), ref, tol, "fused E_loc")
    st = vs.expect(op)
    ost = oracle.stats.statistics(eloc_fused.cpu().numpy())
    for k in ("mean", "variance", "error_of_mean", "tau_corr", "R_hat"):
        np.testing.assert_allclose(getattr(st, k), ost[k], rtol=1e-10, equal_nan=True)


def test_fused_eloc_heisenberg(cuda):
    nk = _nk()
    g = nk.graph.Chain(12)
    hi = nk.hilbert.Spin(0.5, 12, total_sz=0)
    op = nk.operator.Heisenberg(hi, g)
    (W, b, a), var = _params(12, 2, np.float64, 0.2)
    vs = nk.vqs.MCState(nk.sampler.MetropolisExchange(hi, graph=g, n_chains=16), nk.models.RBM(alpha=2), variables=var,
                        n_samples=64, sampler_seed=3)
    eloc = vs.local_estimators(op)
    e, c = ograph.hypercube_edges(12, 1)
    tables = oops.heisenberg_tables(e, c, 1.0, True)
    ref = oest.local_estimators(vs.samples.cpu().numpy(), lambda x: oops.local_operator_conn_padded(x, tables), W, b, a)
    np.testing.assert_allclose(eloc.cpu().numpy(), ref, rtol=1e-9, atol=1e-9)
    assert np.all(vs.samples.cpu().numpy().astype(int).sum(axis=-1) == 0)


# ----------------------------------------------------------------------------------------- statistics
@pytest.mark.parametrize("shape", [(16, 63), (1, 1000), (64, 1), (33, 100), (128, 64), (5, 7), (2, 2)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_statistics(cuda, shape, dtype):
    nk = _nk()
    rs = np.random.default_rng(1)
    x = (rs.normal(size=shape).cumsum(axis=1) * 0.1 + rs.normal(size=shape) - 40.0).astype(dtype)
    st = nk.stats.statistics(torch.from_numpy(x).cuda())
    ref = oracle.stats.statistics(x.astype(np.float64))
    tol = 1e-10 if dtype == np.float64 else 1e-10
    for k in ("mean", "variance", "error_of_mean", "tau_corr", "R_hat"):
        np.testing.assert_allclose(getattr(st, k), ref[k], rtol=tol, atol=1e-12, equal_nan=True, err_msg=k)
    assert set(st.to_dict()) == {"Mean", "Variance", "Sigma", "R_hat", "TauCorr"}


# ----------------------------------------------------------------------------------------- fast (product-form) path
def _fast_case(nk, L, n_dim, alpha, std, B, h=3.0, seed=15324):
    g = nk.graph.Hypercube(L, n_dim)
    N = g.n_nodes
    hi = nk.hilbert.Spin(0.5, N)
    (W, b, a), var = _params(N, alpha, np.float32, std)
    model = nk.models.RBM(alpha=alpha, param_dtype=np.float32)
    sa = nk.sampler.MetropolisLocal(hi, n_chains=B)
    op = nk.operator.Ising(hi, g, h=h)
    e, _ = ograph.hypercube_edges(L, n_dim)
    return g, hi, (W, b, a), var, model, sa, op, e


@pytest.mark.parametrize("L,n_dim,alpha,std", [(10, 2, 4, 0.01), (10, 2, 4, 0.1), (20, 1, 1, 0.3), (4, 2, 2, 0.05), (6, 2, 3, 0.02), (5, 2, 4, 0.2)])
def test_fast_sweep_follows_oracle_chain(cuda, L, n_dim, alpha, std):
    """Same Philox stream => same chains as the oracle except at fp32 accept-boundary ties."""
    nk = _nk()
    B, CL = 96, 2
    g, hi, (W, b, a), var, model, sa, op, e = _fast_case(nk, L, n_dim, alpha, std, B)
    st = sa.init_state(model, var, seed=7)
    seed, t0 = st.rng
    ref = osampler.sample_chain("local", st.σ.cpu().numpy(), W.astype(np.float64), b.astype(np.float64), a.astype(np.float64),
                                chain_length=CL, seed=seed, t0=t0,
                                stream=None if True else None)
    # the oracle above runs in fp64 but must see the fp32 uniforms: rebuild the stream explicitly
    words, u32 = orng.proposal_stream(seed, t0, CL * hi.size, np.arange(B), np.float32)
    ref = osampler.sample_chain("local", st.σ.cpu().numpy(), W.astype(np.float64), b.astype(np.float64), a.astype(np.float64),
                                chain_length=CL, stream=(words[..., 0], u32.astype(np.float64)))
    (samples, logp), st2 = sa.sample(model, var, state=st, chain_length=CL, return_log_probabilities=True, _path=2)
    same = np.all(samples.cpu().numpy() == ref["samples"], axis=(1, 2))
    assert same.mean() >= 0.9, same.mean()
    np.testing.assert_allclose(logp.cpu().numpy()[same], ref["log_prob_samples"][same], rtol=2e-5, atol=2e-4)
    assert np.array_equal(st2.n_accepted_proc.cpu().numpy()[same], ref["n_accepted"][same])
    np.testing.assert_allclose(st2.log_prob.cpu().numpy()[same], ref["log_prob"][same], rtol=2e-5, atol=2e-4)
    assert np.array_equal(st2.σ.cpu().numpy(), samples[:, -1].cpu().numpy())


@pytest.mark.parametrize("L,n_dim,alpha,std,CL", [(10, 2, 4, 0.01, 4), (10, 2, 4, 0.1, 4), (10, 2, 4, 0.05, 48), (4, 2, 2, 0.3, 8), (20, 1, 1, 0.2, 8)])
def test_fast_fused_eloc_matches_oracle(cuda, L, n_dim, alpha, std, CL):
    """E_loc produced inside the fast sweep kernel vs the oracle's materialised E_loc on the same samples (1e-5);
    CL=48 checks that the (C,S) recurrences do not drift over ~5000 proposals."""
    nk = _nk()
    B = 48
    g, hi, (W, b, a), var, model, sa, op, e = _fast_case(nk, L, n_dim, alpha, std, B)
    st = sa.init_state(model, var, seed=11)
    samples, _, eloc, st2 = sa._launch(model, var, st, CL, n_discard=1, operator=op, path=2)
    ref = oest.local_estimators(samples.cpu().numpy(), lambda x: oops.ising_conn_padded(x, e, op.h, op.J), W.astype(np.float64),
                                b.astype(np.float64), a.astype(np.float64))
    assert eloc.dtype == torch.float64
    np.testing.assert_allclose(eloc.cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())
    # and the generic kernel's stand-alone E_loc agrees as well
    vs = nk.vqs.MCState(sa, model, variables=var, n_samples=B, seed=1)
    alone = vs._eloc_on_samples(op, samples, path=1)
    np.testing.assert_allclose(eloc.cpu().numpy(), alone.cpu().numpy(), rtol=2e-5, atol=2e-5 * np.abs(ref).max())


def test_fast_path_hands_over_to_generic_for_large_weights(cuda):
    """max|tanh 2W| beyond the product form's range: NK_PATH_AUTO must give exactly the generic kernel's chains."""
    nk = _nk()
    g, hi, (W, b, a), var, model, sa, op, e = _fast_case(nk, 4, 2, 16, 1.5, 32)  # 8 hidden units per lane
    assert np.abs(W).max() > 4.0
    st = sa.init_state(model, var, seed=3)
    s_auto, _, e_auto, st_a = sa._launch(model, var, st, 3, operator=op, path=0)
    s_gen, _, e_gen, st_g = sa._launch(model, var, st, 3, operator=op, path=1)
    assert np.array_equal(s_auto.cpu().numpy(), s_gen.cpu().numpy())
    assert np.array_equal(e_auto.cpu().numpy(), e_gen.cpu().numpy())
    assert np.array_equal(st_a.n_accepted_proc.cpu().numpy(), st_g.n_accepted_proc.cpu().numpy())


@pytest.mark.parametrize("std", [0.01, 0.4])
def test_fast_sampler_chi_square(cuda, std):
    """test/sampler/test_sampler.py:399-457 on the fast path: histogram vs exact |psi|^2 (4 sites, M = 8)."""
    from scipy import stats as sstats

    nk = _nk()
    N = 4
    hi = nk.hilbert.Spin(0.5, N)
    (W, b, a), var = _params(N, 2, np.float32, std)
    model = nk.models.RBM(alpha=2, param_dtype=np.float32)
    sa = nk.sampler.MetropolisLocal(hi, n_chains=512, sweep_size=8)
    st = sa.init_state(model, var, seed=5)
    samples, _, _, st2 = sa._launch(model, var, st, 100, n_discard=20, path=2)
    p = osampler.exact_distribution(W, b, a, ohilbert.all_states(N))
    counts = np.bincount(ohilbert.states_to_numbers(samples.cpu().numpy().reshape(-1, N), N), minlength=16)
    # chains are autocorrelated: thin to every 4th sweep before the chi-square
    thin = samples[:, ::4].cpu().numpy().reshape(-1, N)
    counts = np.bincount(ohilbert.states_to_numbers(thin, N), minlength=16)
    pv = sstats.chisquare(counts, p * counts.sum()).pvalue
    assert pv > 1e-3, pv
    assert 0.0 < st2.acceptance <= 1.0


# ----------------------------------------------------------------------------------------- theta GEMM (tcgen05)
@pytest.mark.parametrize("N,M,B,bias", [(100, 400, 1000, True), (100, 400, 65, False), (20, 20, 300, True), (16, 64, 128, True),
                                         (128, 512, 257, True), (37, 112, 129, True), (100, 400, 5000, True)])
def test_theta_gemm_tensor_core(cuda, N, M, B, bias):
    """nk_theta_gemm (tcgen05, exact 3-way bf16 split of W) vs the oracle's fp64 theta: fp32 accuracy, any row/column tail."""
    import ctypes as C

    from netket_b200 import _lib

    rs = np.random.default_rng(7)
    W = (rs.normal(size=(N, M)) * 0.3).astype(np.float32)
    b = rs.normal(size=M).astype(np.float32) if bias else None
    sig = _sigma(B, N, seed=2)
    Wt, st = torch.from_numpy(W).cuda(), torch.from_numpy(sig).cuda()
    bt = torch.from_numpy(b).cuda() if bias else None
    rbm = _lib.nk_rbm_t(W=Wt.data_ptr(), b=bt.data_ptr() if bias else None, a=None, N=N, M=M, dtype=0, reserved=0)
    L = _lib.lib()
    ws = torch.empty(int(L.nk_theta_gemm_workspace_bytes(C.byref(rbm), B)), dtype=torch.uint8, device="cuda")
    theta = torch.full((B, M), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(L.nk_theta_gemm(_lib.stream_ptr(), C.byref(rbm), _lib.ptr(st), B, _lib.ptr(theta), _lib.ptr(ws)))
    ref = orbm.theta(sig, W.astype(np.float64), None if b is None else b.astype(np.float64))
    got = theta.cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=2e-6 * np.abs(ref).max())


@pytest.mark.parametrize("N,M,B,bias", [(100, 400, 1000, True), (100, 400, 65, False), (20, 20, 300, True), (22, 44, 128, True),
                                         (37, 113, 129, True), (7, 5, 3, True), (144, 576, 257, True), (250, 64, 200, True)])
def test_theta_gemm_dmma_fp64(cuda, N, M, B, bias):
    """nk_theta_gemm in fp64 (DMMA m8n8k4 on the FP64 tensor cores) vs the oracle's theta: 1e-13, any row / column / K tail."""
    import ctypes as C

    from netket_b200 import _lib

    rs = np.random.default_rng(7)
    W = rs.normal(size=(N, M)) * 0.3
    b = rs.normal(size=M) if bias else None
    sig = _sigma(B, N, seed=2)
    Wt, st = torch.from_numpy(W).cuda(), torch.from_numpy(sig).cuda()
    bt = torch.from_numpy(b).cuda() if bias else None
    rbm = _lib.nk_rbm_t(W=Wt.data_ptr(), b=bt.data_ptr() if bias else None, a=None, N=N, M=M, dtype=1, reserved=0)
    L = _lib.lib()
    ws = torch.empty(max(1, int(L.nk_theta_gemm_workspace_bytes(C.byref(rbm), B))), dtype=torch.uint8, device="cuda")
    theta = torch.full((B, M), float("nan"), dtype=torch.float64, device="cuda")
    _lib.check(L.nk_theta_gemm(_lib.stream_ptr(), C.byref(rbm), _lib.ptr(st), B, _lib.ptr(theta), _lib.ptr(ws)))
    ref = orbm.theta(sig, W, b)
    got = theta.cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())
