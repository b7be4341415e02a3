"""Parity of the general product-form sweep kernel (sweep_prod: fp32 / fp64, LocalRule / ExchangeRule, fused Ising /
LocalOperator local energy; NK_PATH_PROD = 3) against the CPU oracle.

fp64: chains identical to the oracle's (the kernel re-decides in full double precision every proposal whose fixed-point
test is inside the approximation's error band), log-probabilities and E_loc to 1e-10 / 1e-11 relative (the running
(A, B) pairs accumulate one rounding per accepted move).  fp32: chains identical up to accept-boundary ties, E_loc 1e-5.
"""

import numpy as np
import pytest
import torch

from oracle import estimators as oest
from oracle import graph as ograph
from oracle import hilbert as ohilbert
from oracle import operators as oops
from oracle import rbm as orbm
from oracle import rng as orng
from oracle import sampler as osampler
from tolerances import F32_TOL, F32_TOL_LARGE_W, F64_TOL, assert_rel, f32_tol, record, rel_err

pytestmark = pytest.mark.gpu

PROD = 3  # NK_PATH_PROD


def _nk():
    import netket_b200 as nk

    return nk


def _params(N, alpha, dtype, std=0.01, seed=1234):
    W, b, a = orbm.init_params(N, alpha, seed=seed, std=std, dtype=dtype)
    var = {"params": {"Dense": {"kernel": torch.from_numpy(W).cuda(), "bias": torch.from_numpy(b).cuda()},
                      "visible_bias": torch.from_numpy(a).cuda()}}
    return (W, b, a), var


def _f64(*xs):
    return tuple(x.astype(np.float64) for x in xs)


def _case(nk, rule, L, n_dim, alpha, dtype, std, B, total_sz=None, d_max=1, sweep_size=None):
    g = nk.graph.Hypercube(L, n_dim)
    N = g.n_nodes
    hi = nk.hilbert.Spin(0.5, N, total_sz=total_sz)
    (W, b, a), var = _params(N, alpha, dtype, std)
    model = nk.models.RBM(alpha=alpha, param_dtype=dtype)
    e, col = ograph.hypercube_edges(L, n_dim)
    if rule == "local":
        sa = nk.sampler.MetropolisLocal(hi, n_chains=B, sweep_size=sweep_size)
        clusters = None
    else:
        sa = nk.sampler.MetropolisExchange(hi, graph=g, d_max=d_max, n_chains=B, sweep_size=sweep_size)
        clusters = ograph.compute_clusters(N, e, d_max)
        assert np.array_equal(clusters, sa.rule.clusters)
    return g, hi, (W, b, a), var, model, sa, clusters, e, col


# ----------------------------------------------------------------------------------------- fp64 chains
@pytest.mark.parametrize("rule,L,n_dim,alpha,std,total_sz,d_max", [
    ("local", 20, 1, 1, 0.01, None, 1), ("local", 20, 1, 1, 0.3, None, 1), ("local", 16, 1, 2, 0.6, None, 1),
    ("local", 10, 2, 4, 0.01, None, 1), ("local", 10, 2, 4, 0.1, None, 1), ("local", 6, 2, 3, 0.05, None, 1),
    ("exchange", 22, 1, 2, 0.01, 0, 1), ("exchange", 12, 1, 2, 0.5, 0, 2), ("exchange", 10, 1, 1, 0.4, 1, 2),
    ("exchange", 4, 2, 4, 0.1, 0, 2), ("exchange", 10, 2, 4, 0.02, 0, 1), ("exchange", 6, 2, 2, 0.05, 0, 2)])
def test_prod_reproduces_oracle_chain_fp64(cuda, rule, L, n_dim, alpha, std, total_sz, d_max):
    nk = _nk()
    B, CL = 24, 3
    g, hi, (W, b, a), var, model, sa, clusters, e, col = _case(nk, rule, L, n_dim, alpha, np.float64, std, B, total_sz, d_max)
    st = sa.init_state(model, var, seed=15324)
    sig0 = st.σ.cpu().numpy()
    seed, t0 = st.rng
    ref = osampler.sample_chain(rule, sig0, W, b, a, chain_length=CL, seed=seed, t0=t0, clusters=clusters)
    (samples, logp), st2 = sa.sample(model, var, state=st, chain_length=CL, return_log_probabilities=True, _path=PROD)
    assert np.array_equal(samples.cpu().numpy(), ref["samples"])
    scale = max(1.0, np.abs(ref["log_prob_samples"]).max())
    np.testing.assert_allclose(logp.cpu().numpy(), ref["log_prob_samples"], rtol=1e-10, atol=1e-11 * scale)
    np.testing.assert_allclose(st2.log_prob.cpu().numpy(), ref["log_prob"], rtol=1e-10, atol=1e-11 * scale)
    assert np.array_equal(st2.n_accepted_proc.cpu().numpy(), ref["n_accepted"])
    assert st2.n_steps_proc == ref["n_steps"] and st2.rng == (seed, ref["t"])
    assert np.array_equal(st.σ.cpu().numpy(), sig0), "input state must not be mutated"
    assert np.array_equal(st2.σ.cpu().numpy(), ref["sigma"])
    # continuing from the new state continues the same Philox stream
    ref2 = osampler.sample_chain(rule, ref["sigma"], W, b, a, chain_length=2, seed=seed, t0=ref["t"], clusters=clusters)
    s2, _ = sa.sample(model, var, state=st2, chain_length=2, _path=PROD)
    assert np.array_equal(s2.cpu().numpy(), ref2["samples"])
    # explicit proposal stream
    T = CL * sa.sweep_size
    rs = np.random.default_rng(5)
    w0 = rs.integers(0, 2 ** 32, size=(T, B), dtype=np.uint64).astype(np.uint32)
    u = rs.random((T, B))
    ref3 = osampler.sample_chain(rule, sig0, W, b, a, chain_length=CL, stream=(w0, u), clusters=clusters)
    s3, _ = sa.sample(model, var, state=st, chain_length=CL, _stream=(w0, u), _path=PROD)
    assert np.array_equal(s3.cpu().numpy(), ref3["samples"])
    # AUTO takes the same kernel for fp64
    s4, _ = sa.sample(model, var, state=st, chain_length=CL, _path=0)
    assert np.array_equal(s4.cpu().numpy(), ref["samples"])


def test_prod_sweep_size_discard_and_machine_pow(cuda):
    nk = _nk()
    g, hi, (W, b, a), var, model, sa, _, e, col = _case(nk, "local", 10, 1, 1, np.float64, 0.3, 8, sweep_size=7)
    st = sa.init_state(model, var, seed=1)
    seed, t0 = st.rng
    ref = osampler.sample_chain("local", st.σ.cpu().numpy(), W, b, a, chain_length=5, sweep_size=7, seed=seed, t0=t0)
    samples, _, _, st2 = sa._launch(model, var, st, 3, n_discard=2, path=PROD)
    assert np.array_equal(samples.cpu().numpy(), ref["samples"][:, 2:, :])
    assert np.array_equal(st2.n_accepted_proc.cpu().numpy(), ref["n_accepted"])
    assert st2.n_steps_proc == 8 * 5 * 7
    sa1 = nk.sampler.MetropolisLocal(hi, n_chains=8, machine_pow=1.0)
    st = sa1.init_state(model, var, seed=2)
    seed, t0 = st.rng
    ref = osampler.sample_chain("local", st.σ.cpu().numpy(), W, b, a, chain_length=4, seed=seed, t0=t0, machine_pow=1.0)
    samples, _ = sa1.sample(model, var, state=st, chain_length=4, _path=PROD)
    assert np.array_equal(samples.cpu().numpy(), ref["samples"])


@pytest.mark.parametrize("L,n_dim,alpha,std,total_sz,d_max,probs", [
    (12, 1, 2, 0.3, 0, 2, [0.7, 0.3]), (6, 2, 2, 0.05, 0, 2, [1.0, 2.5]), (10, 2, 4, 0.02, 0, 1, None), (22, 1, 2, 0.1, 1, 3, [3.0, 1.0, 0.25])])
def test_prod_weighted_exchange_reproduces_oracle_chain(cuda, L, n_dim, alpha, std, total_sz, d_max, probs):
    """ExchangeRule(probabilities=) (rules/exchange.py:86-123,155-182) on the product-form kernel: inverse-CDF cluster choice over
    the hoppable clusters' weights and the weighted log_prob_corr.  probs=None: one random weight per cluster."""
    nk = _nk()
    B, CL = 24, 3
    g = nk.graph.Hypercube(L, n_dim)
    N = g.n_nodes
    hi = nk.hilbert.Spin(0.5, N, total_sz=total_sz)
    (W, b, a), var = _params(N, alpha, np.float64, std)
    model = nk.models.RBM(alpha=alpha, param_dtype=np.float64)
    e, col = ograph.hypercube_edges(L, n_dim)
    clusters = ograph.compute_clusters(N, e, d_max)
    if probs is None:
        rule = nk.sampler.ExchangeRule(clusters=clusters, probabilities=np.random.default_rng(5).uniform(0.1, 2.0, len(clusters)))
    else:
        rule = nk.sampler.ExchangeRule(graph=g, d_max=d_max, probabilities=probs)
    assert np.array_equal(clusters, rule.clusters)
    sa = nk.sampler.MetropolisSampler(hi, rule, n_chains=B)
    st = sa.init_state(model, var, seed=15324)
    seed, t0 = st.rng
    ref = osampler.sample_chain("exchange", st.σ.cpu().numpy(), W, b, a, chain_length=CL, seed=seed, t0=t0, clusters=clusters,
                                probabilities=rule.probabilities)
    for path in (PROD, 0):  # AUTO takes the same kernel
        (samples, logp), st2 = sa.sample(model, var, state=st, chain_length=CL, return_log_probabilities=True, _path=path)
        assert np.array_equal(samples.cpu().numpy(), ref["samples"])
        np.testing.assert_allclose(logp.cpu().numpy(), ref["log_prob_samples"], rtol=1e-10, atol=1e-10)
        assert np.array_equal(st2.n_accepted_proc.cpu().numpy(), ref["n_accepted"])
    # the weights matter: the uniform rule draws different chains from the same stream
    su, _ = nk.sampler.MetropolisExchange(hi, graph=g, d_max=d_max, n_chains=B).sample(model, var, state=st, chain_length=CL, _path=PROD)
    assert not np.array_equal(su.cpu().numpy(), ref["samples"])
    # fp32 follows the same chains except at accept-boundary ties; fused Heisenberg E_loc on the weighted chain
    (W32, b32, a32), var32 = _params(N, alpha, np.float32, std)
    model32 = nk.models.RBM(alpha=alpha, param_dtype=np.float32)
    st32 = sa.init_state(model32, var32, seed=15324)
    W64, b64, a64 = _f64(W32, b32, a32)
    ref32 = osampler.sample_chain("exchange", st32.σ.cpu().numpy(), W64, b64, a64, chain_length=CL, seed=st32.rng[0], t0=st32.rng[1],
                                  clusters=clusters, probabilities=rule.probabilities)
    op = nk.operator.Heisenberg(hi, g)
    samples32, _, eloc32, _ = sa._launch(model32, var32, st32, CL, operator=op, path=PROD)
    same = np.all(samples32.cpu().numpy() == ref32["samples"], axis=(1, 2))
    assert same.mean() >= 0.8, same.mean()
    tables = oops.heisenberg_tables(e, col, J=1.0, sign_rule=ograph.is_bipartite(N, e))
    refe = oest.local_estimators(samples32.cpu().numpy(), lambda x: oops.local_operator_conn_padded(x, tables), W64, b64, a64)
    assert_rel(eloc32.cpu().numpy(), refe, F32_TOL_LARGE_W if std > 0.2 else F32_TOL)


# ----------------------------------------------------------------------------------------- fp32 chains
@pytest.mark.parametrize("rule,L,n_dim,alpha,std,total_sz,d_max", [
    ("local", 20, 1, 1, 0.3, None, 1), ("local", 10, 2, 4, 0.05, None, 1), ("exchange", 12, 1, 2, 0.5, 0, 2),
    ("exchange", 10, 2, 4, 0.02, 0, 1), ("exchange", 4, 2, 3, 0.2, 0, 2)])
def test_prod_follows_oracle_chain_fp32(cuda, rule, L, n_dim, alpha, std, total_sz, d_max):
    """Same Philox stream => same chains as the fp64 oracle except at fp32 accept-boundary ties."""
    nk = _nk()
    B, CL = 96, 2
    g, hi, (W, b, a), var, model, sa, clusters, e, col = _case(nk, rule, L, n_dim, alpha, np.float32, std, B, total_sz, d_max)
    st = sa.init_state(model, var, seed=7)
    seed, t0 = st.rng
    words, u32 = orng.proposal_stream(seed, t0, CL * hi.size, np.arange(B), np.float32)
    W64, b64, a64 = _f64(W, b, a)
    ref = osampler.sample_chain(rule, st.σ.cpu().numpy(), W64, b64, a64, chain_length=CL, stream=(words[..., 0], u32.astype(np.float64)),
                                clusters=clusters)
    (samples, logp), st2 = sa.sample(model, var, state=st, chain_length=CL, return_log_probabilities=True, _path=PROD)
    same = np.all(samples.cpu().numpy() == ref["samples"], axis=(1, 2))
    assert same.mean() >= 0.9, same.mean()
    np.testing.assert_allclose(logp.cpu().numpy()[same], ref["log_prob_samples"][same], rtol=2e-5, atol=2e-4)
    assert np.array_equal(st2.n_accepted_proc.cpu().numpy()[same], ref["n_accepted"][same])
    assert np.array_equal(st2.σ.cpu().numpy(), samples[:, -1].cpu().numpy())
    if total_sz is not None:
        assert np.all(samples.cpu().numpy().astype(int).sum(axis=-1) == round(2 * total_sz))


# ----------------------------------------------------------------------------------------- fused E_loc
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("L,n_dim,alpha,std,h,CL", [(10, 2, 4, 0.01, 3.0, 3), (10, 2, 4, 0.1, 3.0, 3), (20, 1, 1, 0.2, 1.0, 6),
                                                     (4, 2, 2, 0.3, 0.5, 8), (6, 1, 1, 0.3, 0.0, 2), (10, 2, 4, 0.05, 3.0, 24)])
def test_prod_fused_eloc_ising(cuda, dtype, L, n_dim, alpha, std, h, CL):
    nk = _nk()
    B = 40
    g, hi, (W, b, a), var, model, sa, _, e, col = _case(nk, "local", L, n_dim, alpha, dtype, std, B)
    op = nk.operator.Ising(hi, g, h=h)
    st = sa.init_state(model, var, seed=11)
    samples, _, eloc, st2 = sa._launch(model, var, st, CL, n_discard=1, operator=op, path=PROD)
    W64, b64, a64 = _f64(W, b, a)
    ref = oest.local_estimators(samples.cpu().numpy(), lambda x: oops.ising_conn_padded(x, e, h, 1.0), W64, b64, a64)
    assert eloc.dtype == torch.float64
    tol = F64_TOL if dtype == np.float64 else F32_TOL
    assert_rel(eloc.cpu().numpy(), ref, tol)


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


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("rule,L,n_dim,total_sz,J,sign_rule,order,alpha,std", [
    ("exchange", 12, 1, 0, 1.0, None, 1, 2, 0.2), ("exchange", 22, 1, 0, 1.0, None, 1, 2, 0.01),
    ("exchange", 10, 2, 0, [1.0, 0.5], None, 2, 4, 0.03), ("local", 4, 2, None, 1.0, True, 1, 1, 0.4),
    ("exchange", 6, 1, 0, [1.0, 2.0], [True, False], 2, 3, 0.1)])
def test_prod_fused_eloc_heisenberg(cuda, dtype, rule, L, n_dim, total_sz, J, sign_rule, order, alpha, std):
    """cfg-2 / cfg-4 shapes: Heisenberg and J1-J2 bond operators (LocalOperator tables) fused into the sweep kernel."""
    nk = _nk()
    g, hi, op, tables = _heis(nk, L, n_dim, total_sz, J, sign_rule, order)
    N = g.n_nodes
    (W, b, a), var = _params(N, alpha, dtype, std)
    model = nk.models.RBM(alpha=alpha, param_dtype=dtype)
    sa = (nk.sampler.MetropolisExchange(hi, graph=g, n_chains=20) if rule == "exchange" else nk.sampler.MetropolisLocal(hi, n_chains=20))
    st = sa.init_state(model, var, seed=5)
    samples, _, eloc, _ = sa._launch(model, var, st, 4, n_discard=1, operator=op, path=PROD)
    W64, b64, a64 = _f64(W, b, a)
    ref = oest.local_estimators(samples.cpu().numpy(), lambda x: oops.local_operator_conn_padded(x, tables), W64, b64, a64)
    tol = F64_TOL if dtype == np.float64 else F32_TOL
    assert_rel(eloc.cpu().numpy(), ref, tol)
    if total_sz is not None:
        assert np.all(samples.cpu().numpy().astype(int).sum(axis=-1) == round(2 * total_sz))
    # the generic kernel on the same start state and stream gives the same chain in fp64, hence the same E_loc
    if dtype == np.float64:
        s_gen, _, e_gen, _ = sa._launch(model, var, st, 4, n_discard=1, operator=op, path=1)
        assert np.array_equal(s_gen.cpu().numpy(), samples.cpu().numpy())
        np.testing.assert_allclose(eloc.cpu().numpy(), e_gen.cpu().numpy(), rtol=1e-10, atol=1e-10 * np.abs(ref).max())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_prod_fused_eloc_general_local_operator(cuda, dtype):
    """1-site + 2-site terms with arbitrary (symmetric) matrices: single flips, exchange-like and same-sign double flips,
    entries that map sigma onto itself do not occur off the diagonal; several entries per row (ncmax = 3)."""
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
    (W, b, a), var = _params(N, 3, dtype, 0.3)
    model = nk.models.RBM(alpha=3, param_dtype=dtype)
    sa = nk.sampler.MetropolisLocal(hi, n_chains=64)
    st = sa.init_state(model, var, seed=9)
    samples, _, eloc, _ = sa._launch(model, var, st, 3, operator=op, path=PROD)
    W64, b64, a64 = _f64(W, b, a)
    ref = oest.local_estimators(samples.cpu().numpy(), lambda x: oops.local_operator_conn_padded(x, tables), W64, b64, a64)
    tol = F64_TOL if dtype == np.float64 else F32_TOL
    assert_rel(eloc.cpu().numpy(), ref, tol)


# ----------------------------------------------------------------------------------------- hand-over, statistics
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("rule", ["local", "exchange"])
def test_prod_hands_over_to_generic_for_large_weights(cuda, dtype, rule):
    nk = _nk()
    g, hi, (W, b, a), var, model, sa, clusters, e, col = _case(nk, rule, 4, 2, 16, dtype, 1.5, 32, 0 if rule == "exchange" else None, 1)
    assert np.abs(W).max() > 4.0
    op = nk.operator.Ising(hi, g, h=1.0)
    st = sa.init_state(model, var, seed=3)
    s_auto, _, e_auto, st_a = sa._launch(model, var, st, 3, operator=op, path=0)
    s_gen, _, e_gen, st_g = sa._launch(model, var, st, 3, operator=op, path=1)
    assert np.array_equal(s_auto.cpu().numpy(), s_gen.cpu().numpy())
    assert np.array_equal(e_auto.cpu().numpy(), e_gen.cpu().numpy())
    assert np.array_equal(st_a.n_accepted_proc.cpu().numpy(), st_g.n_accepted_proc.cpu().numpy())


@pytest.mark.parametrize("dtype,rule,std", [(np.float32, "exchange", 0.3), (np.float64, "exchange", 0.3), (np.float64, "local", 0.4),
                                            (np.float32, "local", 0.05)])
def test_prod_sampler_chi_square(cuda, dtype, rule, std):
    """test/sampler/test_sampler.py:399-457: histogram vs exact |psi|^2 (6 sites; the exchange rule stays in total_sz = 0)."""
    from scipy import stats as sstats

    nk = _nk()
    N = 6
    g = nk.graph.Chain(N)
    total_sz = 0 if rule == "exchange" else None
    hi = nk.hilbert.Spin(0.5, N, total_sz=total_sz)
    (W, b, a), var = _params(N, 2, dtype, std)
    model = nk.models.RBM(alpha=2, param_dtype=dtype)
    sa = (nk.sampler.MetropolisExchange(hi, graph=g, d_max=2, n_chains=512, sweep_size=8) if rule == "exchange"
          else nk.sampler.MetropolisLocal(hi, n_chains=512, sweep_size=8))
    st = sa.init_state(model, var, seed=5)
    samples, _, _, st2 = sa._launch(model, var, st, 100, n_discard=20, path=PROD)
    states = ohilbert.all_states(N)
    if total_sz is not None:
        states = states[states.astype(int).sum(axis=1) == 0]
    p = osampler.exact_distribution(*_f64(W, b, a), states)
    thin = samples[:, ::4].cpu().numpy().reshape(-1, N)
    idx = {tuple(s): k for k, s in enumerate(states)}
    counts = np.bincount([idx[tuple(s)] for s in thin], minlength=len(states))
    pv = sstats.chisquare(counts, p * counts.sum()).pvalue
    assert pv > 1e-3, pv
    assert 0.0 < st2.acceptance <= 1.0


# ----------------------------------------------------------------------------------------- stand-alone local estimator
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind,L,n_dim,alpha,std", [("ising", 10, 2, 4, 0.01), ("ising", 10, 2, 4, 0.1), ("ising", 20, 1, 1, 0.3),
                                                     ("heis", 22, 1, 2, 0.05), ("j1j2", 10, 2, 4, 0.03), ("heis", 4, 2, 2, 0.4)])
def test_prod_standalone_eloc(cuda, dtype, kind, L, n_dim, alpha, std):
    """nk_eloc_*_rbm on arbitrary configurations (not produced by a chain), product-form kernel vs oracle and vs the
    theta-form kernel; leading batch dimensions are kept."""
    nk = _nk()
    if kind == "ising":
        g = nk.graph.Hypercube(L, n_dim)
        hi = nk.hilbert.Spin(0.5, g.n_nodes)
        op = nk.operator.Ising(hi, g, h=3.0)
        e, _ = ograph.hypercube_edges(L, n_dim)
        conn = lambda x: oops.ising_conn_padded(x, e, 3.0, 1.0)  # noqa: E731
        total_sz = None
    else:
        total_sz = 0
        g, hi, op, tables = _heis(nk, L, n_dim, total_sz, [1.0, 0.5] if kind == "j1j2" else 1.0, None, 2 if kind == "j1j2" else 1)
        conn = lambda x: oops.local_operator_conn_padded(x, tables)  # noqa: E731
    N = g.n_nodes
    (W, b, a), var = _params(N, alpha, dtype, std)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(nk.hilbert.Spin(0.5, N), n_chains=16), nk.models.RBM(alpha=alpha, param_dtype=dtype),
                        variables=var, n_samples=16, seed=1)
    sig = ohilbert.random_state(9, 3 * 67, N, total_sz).reshape(3, 67, N)
    out = vs._eloc_on_samples(op, torch.from_numpy(sig).cuda(), path=PROD)
    assert tuple(out.shape) == (3, 67) and out.dtype == torch.float64
    ref = oest.local_value_kernel(sig.reshape(-1, N), conn, *_f64(W, b, a)).reshape(3, 67)
    tol = F64_TOL if dtype == np.float64 else F32_TOL
    assert_rel(out.cpu().numpy(), ref, tol)
    gen = vs._eloc_on_samples(op, torch.from_numpy(sig).cuda(), path=1)
    np.testing.assert_allclose(out.cpu().numpy(), gen.cpu().numpy(), rtol=2 * tol, atol=2 * tol * np.abs(ref).max())
    auto = vs._eloc_on_samples(op, torch.from_numpy(sig).cuda(), path=0)
    assert np.array_equal(auto.cpu().numpy(), out.cpu().numpy())


def test_prod_standalone_eloc_hands_over_for_large_weights(cuda):
    nk = _nk()
    g = nk.graph.Hypercube(4, 2)
    hi = nk.hilbert.Spin(0.5, 16)
    op = nk.operator.Ising(hi, g, h=1.0)
    (W, b, a), var = _params(16, 16, np.float64, 1.5)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=16), nk.models.RBM(alpha=16), variables=var, n_samples=16, seed=1)
    sig = torch.from_numpy(ohilbert.random_state(2, 50, 16)).cuda()
    assert np.array_equal(vs._eloc_on_samples(op, sig, path=0).cpu().numpy(), vs._eloc_on_samples(op, sig, path=1).cpu().numpy())
    with pytest.raises(nk.NkError):
        # beyond the product-form kernels' limits (N <= 1024): NK_PATH_PROD must refuse, not fall back
        nk.vqs.MCState(nk.sampler.MetropolisLocal(nk.hilbert.Spin(0.5, 1030), n_chains=16), nk.models.RBM(alpha=1),
                       n_samples=16, seed=1)._eloc_on_samples(nk.operator.Ising(nk.hilbert.Spin(0.5, 1030), nk.graph.Hypercube(1030, 1), h=1.0),
                                                              torch.ones((4, 1030), dtype=torch.int8, device="cuda"), path=PROD)


# ----------------------------------------------------------------------------------------- large systems
# M > 512: several warps per chain (partial sums combined through shared memory);  N > 128: sigma bits beyond four words;
# tables larger than shared memory: rows read through L2 (generic-address loads).
LARGE = [("local", 20, 1, 32, 0.01, None, 1),      # N=20, M=640: 2 warps per chain
         ("local", 16, 1, 100, 0.005, None, 1),    # N=16, M=1600: 4-5 warps per chain
         ("local", 12, 2, 4, 0.02, None, 1),       # N=144, M=576: 2 warps per chain, N > 128
         ("local", 10, 1, 110, 0.01, None, 1),     # N=10, M=1100: two 550-unit segments would need an uninstantiated shape -> 3 warps
         ("local", 14, 2, 2, 0.03, None, 1),       # N=196, M=392: one warp per chain, table (326 KB fp32) not resident
         ("exchange", 12, 2, 1, 0.05, 0, 1),       # N=144 exchange, 288 clusters
         ("exchange", 20, 1, 32, 0.01, 0, 2),      # N=20, M=640 exchange: 2 warps per chain, every warp keeps the hoppable-cluster words
         ("exchange", 6, 2, 32, 0.01, 0, 1),       # N=36, M=1152 exchange on a 2d lattice
         ("exchange", 14, 2, 2, 0.03, 0, 1)]       # N=196, M=392 exchange: the fp32 table is not resident either (rows through L2)


@pytest.mark.parametrize("rule,L,n_dim,alpha,std,total_sz,d_max", LARGE)
def test_prod_large_reproduces_oracle_chain_fp64(cuda, rule, L, n_dim, alpha, std, total_sz, d_max):
    nk = _nk()
    B, CL = 12, 2
    g, hi, (W, b, a), var, model, sa, clusters, e, col = _case(nk, rule, L, n_dim, alpha, np.float64, std, B, total_sz, d_max)
    st = sa.init_state(model, var, seed=15324)
    seed, t0 = st.rng
    ref = osampler.sample_chain(rule, st.σ.cpu().numpy(), W, b, a, chain_length=CL, seed=seed, t0=t0, clusters=clusters)
    (samples, logp), st2 = sa.sample(model, var, state=st, chain_length=CL, return_log_probabilities=True, _path=PROD)
    assert np.array_equal(samples.cpu().numpy(), ref["samples"])
    scale = max(1.0, np.abs(ref["log_prob_samples"]).max())
    np.testing.assert_allclose(logp.cpu().numpy(), ref["log_prob_samples"], rtol=1e-10, atol=1e-11 * scale)
    np.testing.assert_allclose(st2.log_prob.cpu().numpy(), ref["log_prob"], rtol=1e-10, atol=1e-11 * scale)
    assert np.array_equal(st2.n_accepted_proc.cpu().numpy(), ref["n_accepted"])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("rule,L,n_dim,alpha,std,total_sz,d_max", LARGE)
def test_prod_large_fused_eloc(cuda, dtype, rule, L, n_dim, alpha, std, total_sz, d_max):
    nk = _nk()
    B = 30
    g, hi, (W, b, a), var, model, sa, clusters, e, col = _case(nk, rule, L, n_dim, alpha, dtype, std, B, total_sz, d_max)
    if rule == "local":
        op = nk.operator.Ising(hi, g, h=2.0)
        conn = lambda x: oops.ising_conn_padded(x, e, 2.0, 1.0)  # noqa: E731
    else:
        op = nk.operator.Heisenberg(hi, g)
        tables = oops.heisenberg_tables(e, col, 1.0, ograph.is_bipartite(g.n_nodes, e))
        conn = lambda x: oops.local_operator_conn_padded(x, tables)  # noqa: E731
    st = sa.init_state(model, var, seed=11)
    samples, _, eloc, st2 = sa._launch(model, var, st, 2, n_discard=1, operator=op, path=PROD)
    ref = oest.local_estimators(samples.cpu().numpy(), conn, *_f64(W, b, a))
    tol = F64_TOL if dtype == np.float64 else f32_tol(W.shape[1])
    assert_rel(eloc.cpu().numpy(), ref, tol)
    if dtype == np.float32 and W.shape[1] > 512:  # the reference algorithm itself in float32 (NumPy): its own rounding error
        ref32 = oest.local_estimators(samples.cpu().numpy()[:6], conn, W, b, a)
        record("reference algorithm in float32 vs float64 oracle, M = %d" % W.shape[1], tol, rel_err(ref32, ref[:6], 0.1))
    # fp32: the chain itself follows the oracle's up to accept-boundary ties
    if dtype == np.float32:
        seed, t0 = st.rng
        words, u32 = orng.proposal_stream(seed, t0, 3 * sa.sweep_size, np.arange(B), np.float32)
        r = osampler.sample_chain(rule, st.σ.cpu().numpy(), *_f64(W, b, a), chain_length=3, stream=(words[..., 0], u32.astype(np.float64)),
                                  clusters=clusters)
        same = np.all(samples.cpu().numpy() == r["samples"][:, 1:], axis=(1, 2))
        assert same.mean() >= 0.8, same.mean()
    # stand-alone estimator on the same samples
    vs = nk.vqs.MCState(sa, model, variables=var, n_samples=B, seed=1)
    alone = vs._eloc_on_samples(op, samples, path=PROD)
    assert_rel(alone.cpu().numpy(), ref, tol)


# ----------------------------------------------------------------------------------------- edge shapes
@pytest.mark.parametrize("N,M,B,CL,sweep_size", [(2, 1, 1, 3, None), (3, 5, 7, 2, 1), (128, 128, 5, 1, None), (129, 64, 3, 1, 40),
                                                  (16, 512, 9, 2, None), (16, 513, 9, 2, None), (31, 33, 33, 2, 100), (64, 65, 4, 1, None)])
def test_prod_edge_shapes_fp64(cuda, N, M, B, CL, sweep_size):
    """Smallest / boundary shapes of the product-form kernels (one chain, one hidden unit, N = 128 / 129, M = 512 / 513,
    sweep_size of 1 and larger than N): fp64 chains, log-probabilities and fused E_loc against the oracle."""
    nk = _nk()
    rs = np.random.default_rng(N * 1000 + M)
    std = 0.2 if M < 512 else 0.04  # (16 hidden units per lane: larger weights leave the product form's range)
    W = rs.normal(size=(N, M)) * std
    b = rs.normal(size=M) * 0.2
    a = rs.normal(size=N) * 0.2
    var = {"params": {"Dense": {"kernel": torch.from_numpy(W).cuda(), "bias": torch.from_numpy(b).cuda()},
                      "visible_bias": torch.from_numpy(a).cuda()}}
    g = nk.graph.Hypercube(N, 1, pbc=N > 2)
    hi = nk.hilbert.Spin(0.5, N)
    op = nk.operator.Ising(hi, g, h=0.7)
    model = nk.models.RBM(alpha=M / N)
    assert model.n_hidden(N) == M
    sa = nk.sampler.MetropolisLocal(hi, n_chains=B, sweep_size=sweep_size)
    st = sa.init_state(model, var, seed=5)
    seed, t0 = st.rng
    ref = osampler.sample_chain("local", st.σ.cpu().numpy(), W, b, a, chain_length=CL, sweep_size=sweep_size, seed=seed, t0=t0)
    samples, logp, eloc, st2 = sa._launch(model, var, st, CL, operator=op, return_log_probabilities=True, path=PROD)
    assert np.array_equal(samples.cpu().numpy(), ref["samples"])
    np.testing.assert_allclose(logp.cpu().numpy(), ref["log_prob_samples"], rtol=1e-10, atol=1e-10)
    assert np.array_equal(st2.n_accepted_proc.cpu().numpy(), ref["n_accepted"])
    e = np.asarray(g.edges(), dtype=np.int64).reshape(-1, 2)
    eref = oest.local_estimators(ref["samples"], lambda x: oops.ising_conn_padded(x, e, 0.7, 1.0), W, b, a)
    assert_rel(eloc.cpu().numpy(), eref, F64_TOL)


def test_prod_burn_in_only_and_empty_batch(cuda):
    """chain_length = 0 (pure burn-in) advances the chains exactly like the first sweeps of a longer run; B = 0 is a no-op."""
    nk = _nk()
    g, hi, (W, b, a), var, model, sa, _, e, col = _case(nk, "local", 12, 1, 2, np.float64, 0.3, 6)
    st = sa.init_state(model, var, seed=8)
    seed, t0 = st.rng
    ref = osampler.sample_chain("local", st.σ.cpu().numpy(), W, b, a, chain_length=3, seed=seed, t0=t0)
    samples, _, _, st2 = sa._launch(model, var, st, 0, n_discard=3, path=PROD)
    assert tuple(samples.shape) == (6, 0, 12)
    assert np.array_equal(st2.σ.cpu().numpy(), ref["sigma"])
    assert st2.rng == (seed, ref["t"])
    import ctypes as C

    from netket_b200 import _lib

    rbm = nk.models.RBM.c_struct(var)
    ch = _lib.nk_chains_t(sigma=None, log_prob=None, n_accepted=None, workspace=None, B=0, seed=0, t=0, chain_offset=0)
    args = _lib.nk_sweep_t()
    args.rule, args.chain_length, args.n_discard, args.sweep_size, args.machine_pow, args.path = 0, 1, 0, 12, 2.0, 0
    assert _lib.lib().nk_sweep(_lib.stream_ptr(), C.byref(rbm), C.byref(ch), C.byref(args)) == 0


def test_forced_product_path_still_produces_outputs_for_large_weights(cuda):
    """NK_PATH_PROD with weights outside the product form's numerical range: the in-stream hand-over runs all the same."""
    nk = _nk()
    g, hi, (W, b, a), var, model, sa, _, e, col = _case(nk, "local", 4, 2, 16, np.float64, 1.5, 16)
    op = nk.operator.Ising(hi, g, h=1.0)
    st = sa.init_state(model, var, seed=3)
    s_prod, _, e_prod, _ = sa._launch(model, var, st, 3, operator=op, path=PROD)
    s_gen, _, e_gen, _ = sa._launch(model, var, st, 3, operator=op, path=1)
    assert np.array_equal(s_prod.cpu().numpy(), s_gen.cpu().numpy())
    assert np.array_equal(e_prod.cpu().numpy(), e_gen.cpu().numpy())
    assert set(np.unique(s_prod.cpu().numpy())) <= {-1, 1}


# ----------------------------------------------------------------------------------------- exchange-rule edge cases
@pytest.mark.parametrize("L,n_dim,d_max,total_sz,expect_clusters", [(10, 2, 3, 0, 1200),     # > 1024 clusters: second hop word per lane
                                                                     (6, 2, 6, 0, 630),       # all pairs: 35 clusters per site > 32 -> hand-over
                                                                     (8, 1, 1, 4, 8),         # all spins up: no hoppable cluster, chains frozen
                                                                     (4, 1, 1, 0, 4)])
def test_prod_exchange_edge_cases_fp64(cuda, L, n_dim, d_max, total_sz, expect_clusters):
    nk = _nk()
    B, CL = 6, 2
    g, hi, (W, b, a), var, model, sa, clusters, e, col = _case(nk, "exchange", L, n_dim, 2, np.float64, 0.1, B, total_sz, d_max)
    assert clusters.shape[0] == expect_clusters
    st = sa.init_state(model, var, seed=21)
    seed, t0 = st.rng
    ref = osampler.sample_chain("exchange", st.σ.cpu().numpy(), W, b, a, chain_length=CL, seed=seed, t0=t0, clusters=clusters)
    for path in (PROD, 1):  # product-form and theta-form kernels
        (samples, logp), st2 = sa.sample(model, var, state=st, chain_length=CL, return_log_probabilities=True, _path=path)
        assert np.array_equal(samples.cpu().numpy(), ref["samples"])
        np.testing.assert_allclose(logp.cpu().numpy(), ref["log_prob_samples"], rtol=1e-10, atol=1e-10)
        assert np.array_equal(st2.n_accepted_proc.cpu().numpy(), ref["n_accepted"])
        if total_sz == L ** n_dim / 2:  # no hoppable cluster: the reference's nan correction rejects every (identity) proposal
            assert int(st2.n_accepted_proc.sum()) == 0 and np.array_equal(st2.σ.cpu().numpy(), st.σ.cpu().numpy())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("hb,vb", [(False, True), (True, False), (False, False)])
def test_prod_without_biases(cuda, dtype, hb, vb):
    nk = _nk()
    N, alpha, B = 12, 3, 10
    W, b, a = orbm.init_params(N, alpha, seed=3, std=0.2, dtype=dtype, use_hidden_bias=hb, use_visible_bias=vb)
    dense = {"kernel": torch.from_numpy(W).cuda()}
    if hb:
        dense["bias"] = torch.from_numpy(b).cuda()
    p = {"Dense": dense}
    if vb:
        p["visible_bias"] = torch.from_numpy(a).cuda()
    var = {"params": p}
    g = nk.graph.Hypercube(N, 1)
    hi = nk.hilbert.Spin(0.5, N)
    op = nk.operator.Ising(hi, g, h=1.1)
    model = nk.models.RBM(alpha=alpha, param_dtype=dtype, use_hidden_bias=hb, use_visible_bias=vb)
    sa = nk.sampler.MetropolisLocal(hi, n_chains=B)
    st = sa.init_state(model, var, seed=2)
    samples, _, eloc, _ = sa._launch(model, var, st, 3, operator=op, path=PROD)
    e, _ = ograph.hypercube_edges(N, 1)
    W64 = W.astype(np.float64)
    b64 = None if b is None else b.astype(np.float64)
    a64 = None if a is None else a.astype(np.float64)
    ref = oest.local_estimators(samples.cpu().numpy(), lambda x: oops.ising_conn_padded(x, e, 1.1, 1.0), W64, b64, a64)
    tol = F64_TOL if dtype == np.float64 else F32_TOL
    assert_rel(eloc.cpu().numpy(), ref, tol)
    if dtype == np.float64:
        seed, t0 = st.rng
        r = osampler.sample_chain("local", st.σ.cpu().numpy(), W64, b64, a64, chain_length=3, seed=seed, t0=t0)
        assert np.array_equal(samples.cpu().numpy(), r["samples"])


# ----------------------------------------------------------------------------------------- more sampler / operator edge cases
@pytest.mark.parametrize("rule,machine_pow", [("local", 1.0), ("local", 3.0), ("local", 0.5), ("exchange", 1.0), ("exchange", 2.5),
                                              ("local", 0.0)])
def test_prod_machine_pow_fp64(cuda, rule, machine_pow):
    """sampler/base.py:139-150: machine_pow is any non-negative real; 0 samples the uniform distribution (always accept)."""
    nk = _nk()
    N, B = 12, 10
    g = nk.graph.Hypercube(N, 1)
    total_sz = 0 if rule == "exchange" else None
    hi = nk.hilbert.Spin(0.5, N, total_sz=total_sz)
    (W, b, a), var = _params(N, 2, np.float64, 0.4)
    model = nk.models.RBM(alpha=2)
    e, _ = ograph.hypercube_edges(N, 1)
    if rule == "local":
        sa, clusters = nk.sampler.MetropolisLocal(hi, n_chains=B, machine_pow=machine_pow), None
    else:
        sa = nk.sampler.MetropolisExchange(hi, graph=g, d_max=2, n_chains=B, machine_pow=machine_pow)
        clusters = ograph.compute_clusters(N, e, 2)
    st = sa.init_state(model, var, seed=17)
    seed, t0 = st.rng
    ref = osampler.sample_chain(rule, st.σ.cpu().numpy(), W, b, a, chain_length=3, seed=seed, t0=t0, clusters=clusters,
                                machine_pow=machine_pow)
    for path in (PROD, 1):
        (samples, logp), st2 = sa.sample(model, var, state=st, chain_length=3, return_log_probabilities=True, _path=path)
        assert np.array_equal(samples.cpu().numpy(), ref["samples"]), path
        np.testing.assert_allclose(logp.cpu().numpy(), ref["log_prob_samples"], rtol=1e-10, atol=1e-10)
        assert np.array_equal(st2.n_accepted_proc.cpu().numpy(), ref["n_accepted"])
    if machine_pow == 0.0:
        assert int(st2.n_accepted_proc.sum()) == B * 3 * N


def test_prod_sharded_chains_and_large_counters(cuda):
    """chain_offset (multi-GPU sharding) and Philox counters / seeds beyond 32 bits on the product-form kernel."""
    nk = _nk()
    g, hi, (W, b, a), var, model, sa, _, e, col = _case(nk, "local", 12, 1, 1, np.float64, 0.3, 12)
    seed, t0 = (1 << 40) + 12345, (1 << 33) + 7
    st = sa.init_state(model, var, seed=4).replace(rng=(seed, t0))
    full, stf = sa.sample(model, var, state=st, chain_length=2, _path=PROD)
    ref = osampler.sample_chain("local", st.σ.cpu().numpy(), W, b, a, chain_length=2, seed=seed, t0=t0)
    assert np.array_equal(full.cpu().numpy(), ref["samples"])
    assert stf.rng == (seed, t0 + 2 * 12)
    sa8 = nk.sampler.MetropolisLocal(hi, n_chains=8)
    st8 = sa8.init_state(model, var, seed=4).replace(σ=st.σ[4:12].clone(), chain_offset=4, rng=(seed, t0))
    part, _ = sa8.sample(model, var, state=st8, chain_length=2, _path=PROD)
    assert np.array_equal(part.cpu().numpy(), full[4:12].cpu().numpy())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_prod_local_operator_edge_cases(cuda, dtype):
    """1-site-only operators (a transverse field written as a LocalOperator equals Ising with J = 0), a pure constant, entries
    below mel_cutoff, and a term given twice (summed by the reference's `_append_matrix`)."""
    nk = _nk()
    N, B = 10, 12
    g = nk.graph.Hypercube(N, 1)
    hi = nk.hilbert.Spin(0.5, N)
    (W, b, a), var = _params(N, 2, dtype, 0.3)
    model = nk.models.RBM(alpha=2, param_dtype=dtype)
    sa = nk.sampler.MetropolisLocal(hi, n_chains=B)
    st = sa.init_state(model, var, seed=6)
    sx = np.array([[0.0, 1.0], [1.0, 0.0]])
    sz = np.array([[1.0, 0.0], [0.0, -1.0]])
    tol = F64_TOL if dtype == np.float64 else F32_TOL
    W64, b64, a64 = _f64(W, b, a)
    # transverse field as LocalOperator == Ising(h, J=0)
    op_x = nk.operator.LocalOperator(hi, [-1.3 * sx] * N, [[i] for i in range(N)])
    samples, _, e_x, _ = sa._launch(model, var, st, 2, operator=op_x, path=PROD)
    e, _ = ograph.hypercube_edges(N, 1)
    ref = oest.local_estimators(samples.cpu().numpy(), lambda x: oops.ising_conn_padded(x, e, 1.3, 0.0), W64, b64, a64)
    assert_rel(e_x.cpu().numpy(), ref, tol)
    # constant + diagonal terms only, with an off-diagonal entry below the cutoff and a duplicated term
    tiny = np.array([[0.0, 1e-12], [1e-12, 0.0]])
    ops = [0.5 * sz, tiny, 0.5 * sz, np.kron(sz, sz)]
    aon = [[0], [1], [0], [2, 3]]
    op_d = nk.operator.LocalOperator(hi, ops, aon, constant=0.7)
    tables = oops.pack_internals(oops.canonical_operators_dict(ops, aon), 0.7)
    samples, _, e_d, _ = sa._launch(model, var, st, 2, operator=op_d, path=PROD)
    ref = oest.local_estimators(samples.cpu().numpy(), lambda x: oops.local_operator_conn_padded(x, tables), W64, b64, a64)
    assert_rel(e_d.cpu().numpy(), ref, tol)
    s = samples.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(e_d.cpu().numpy(), 0.7 + s[..., 0] + s[..., 2] * s[..., 3], rtol=tol, atol=tol)


@pytest.mark.parametrize("N,M", [(128, 256), (100, 512), (128, 512), (16, 4), (64, 500), (127, 132)])
def test_auto_path_boundaries_fp32(cuda, N, M):
    """NK_PATH_AUTO in fp32 at the limits of the kernels' coverage (tuned kernel: table resident, M % 4 == 0; general kernel:
    one warp, resident; general kernel: table read through L2): fused E_loc vs oracle and the chains vs the fp64 oracle."""
    nk = _nk()
    B, CL = 24, 2
    rs = np.random.default_rng(N + M)
    W = (rs.normal(size=(N, M)) * 0.03).astype(np.float32)
    b = (rs.normal(size=M) * 0.1).astype(np.float32)
    a = (rs.normal(size=N) * 0.1).astype(np.float32)
    var = {"params": {"Dense": {"kernel": torch.from_numpy(W).cuda(), "bias": torch.from_numpy(b).cuda()},
                      "visible_bias": torch.from_numpy(a).cuda()}}
    g = nk.graph.Hypercube(N, 1)
    hi = nk.hilbert.Spin(0.5, N)
    op = nk.operator.Ising(hi, g, h=1.0)
    model = nk.models.RBM(alpha=M / N, param_dtype=np.float32)
    assert model.n_hidden(N) == M
    sa = nk.sampler.MetropolisLocal(hi, n_chains=B)
    st = sa.init_state(model, var, seed=9)
    samples, _, eloc, st2 = sa._launch(model, var, st, CL, operator=op, path=0)
    W64, b64, a64 = _f64(W, b, a)
    e = np.asarray(g.edges(), dtype=np.int64).reshape(-1, 2)
    ref = oest.local_estimators(samples.cpu().numpy(), lambda x: oops.ising_conn_padded(x, e, 1.0, 1.0), W64, b64, a64)
    assert_rel(eloc.cpu().numpy(), ref, F32_TOL)
    seed, t0 = st.rng
    words, u32 = orng.proposal_stream(seed, t0, CL * N, np.arange(B), np.float32)
    r = osampler.sample_chain("local", st.σ.cpu().numpy(), W64, b64, a64, chain_length=CL, stream=(words[..., 0], u32.astype(np.float64)))
    same = np.all(samples.cpu().numpy() == r["samples"], axis=(1, 2))
    assert same.mean() >= 0.85, same.mean()


# ----------------------------------------------------------------------------------------- large weights (trained networks)
def _prod_flags(sa, B, M, esz):
    """The hand-over / range flags the prep kernels left in the sampler's workspace (netket_b200/csrc/api.cu layout)."""
    ws = next(iter(sa.__dict__["_ws_cache"].values()))
    off = (B * M * esz + 255) & ~255
    return ws[off:off + 32].view(torch.int32).cpu().numpy()


@pytest.mark.parametrize("dtype,std,expect_split,expect_wide_e", [(np.float32, 0.2, 2, 0), (np.float32, 0.35, 2, 0), (np.float32, 0.05, 1, 0),
                                                                   (np.float64, 0.2, 1, 0), (np.float64, 0.35, 1, 1), (np.float64, 0.5, 1, 1)])
def test_prod_wide_range_weights(cuda, dtype, std, expect_split, expect_wide_e):
    """max|W| up to ~1.5 at N=100, M=400 stays on the product-form kernel (fp32: two logarithms per lane product and a
    short renormalisation period; fp64: local energy reduced as (mantissa, exponent) pairs once the row sums of |W| could take a
    product of M factors out of the double range) instead of handing over to the theta-form kernel; chains and E_loc still
    match the oracle."""
    nk = _nk()
    B, CL = 48, 3
    g, hi, (W, b, a), var, model, sa, _, e, col = _case(nk, "local", 10, 2, 4, dtype, std, B)
    op = nk.operator.Ising(hi, g, h=3.0)
    st = sa.init_state(model, var, seed=13)
    samples, _, eloc, st2 = sa._launch(model, var, st, CL, n_discard=1, operator=op, path=PROD)
    flags = _prod_flags(sa, B, 400, np.dtype(dtype).itemsize)
    assert flags[0] == 0, "the product-form kernel handed over to the theta-form kernel"
    assert flags[1] >= 1 and flags[6] == expect_split and flags[7] == expect_wide_e, flags[:8]
    W64, b64, a64 = _f64(W, b, a)
    ref = oest.local_estimators(samples.cpu().numpy(), lambda x: oops.ising_conn_padded(x, e, 3.0, 1.0), W64, b64, a64)
    tol = F64_TOL if dtype == np.float64 else F32_TOL_LARGE_W
    assert_rel(eloc.cpu().numpy(), ref, tol)
    seed, t0 = st.rng
    if dtype == np.float64:
        r = osampler.sample_chain("local", st.σ.cpu().numpy(), W64, b64, a64, chain_length=CL + 1, seed=seed, t0=t0)
        assert np.array_equal(samples.cpu().numpy(), r["samples"][:, 1:])
    else:
        words, u32 = orng.proposal_stream(seed, t0, (CL + 1) * 100, np.arange(B), np.float32)
        r = osampler.sample_chain("local", st.σ.cpu().numpy(), W64, b64, a64, chain_length=CL + 1, stream=(words[..., 0], u32.astype(np.float64)))
        same = np.all(samples.cpu().numpy() == r["samples"][:, 1:], axis=(1, 2))
        assert same.mean() >= 0.85, same.mean()


@pytest.mark.parametrize("std,fast_gives_up,prod_gives_up", [(0.05, 0, 0), (0.2, 0, 0), (0.35, 0, 0), (0.5, 1, 0), (0.9, 1, 1)])
def test_fp32_auto_chain_large_weights(cuda, std, fast_gives_up, prod_gives_up):
    """NK_PATH_AUTO, fp32 TFIM: tuned kernel -> (weights beyond its range) general kernel in wide mode -> (beyond that) theta-form
    kernel, all enqueued in-stream; whichever kernel does the work, E_loc and the chains match the oracle."""
    nk = _nk()
    B, CL = 48, 2
    g, hi, (W, b, a), var, model, sa, _, e, col = _case(nk, "local", 10, 2, 4, np.float32, std, B)
    op = nk.operator.Ising(hi, g, h=3.0)
    st = sa.init_state(model, var, seed=13)
    samples, _, eloc, st2 = sa._launch(model, var, st, CL, n_discard=1, operator=op, path=0)
    flags = _prod_flags(sa, B, 400, 4)
    assert (int(flags[0] != 0), int(flags[5] != 0)) == (fast_gives_up, prod_gives_up), flags[:8]
    W64, b64, a64 = _f64(W, b, a)
    ref = oest.local_estimators(samples.cpu().numpy(), lambda x: oops.ising_conn_padded(x, e, 3.0, 1.0), W64, b64, a64)
    assert_rel(eloc.cpu().numpy(), ref, F32_TOL_LARGE_W)
    seed, t0 = st.rng
    words, u32 = orng.proposal_stream(seed, t0, (CL + 1) * 100, np.arange(B), np.float32)
    r = osampler.sample_chain("local", st.σ.cpu().numpy(), W64, b64, a64, chain_length=CL + 1, stream=(words[..., 0], u32.astype(np.float64)))
    same = np.all(samples.cpu().numpy() == r["samples"][:, 1:], axis=(1, 2))
    assert same.mean() >= 0.85, same.mean()
    assert set(np.unique(samples.cpu().numpy())) <= {-1, 1}
