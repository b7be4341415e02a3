"""The oracle against the reference's own known answers and invariants.  CPU only."""

import numpy as np
import pytest
from scipy import stats as sstats

import oracle
from oracle import ed, estimators, graph, hilbert, operators as ops, rbm, rng, sampler
from oracle import stats as ostats


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    h = lambda x: [int(v) for v in x]  # noqa: E731
    assert h(rng.philox4x32_10([0, 0, 0, 0], [0, 0])) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert h(rng.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert h(rng.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_uniform_ranges():
    w, u64 = rng.proposal_stream(3, 0, 50, np.arange(40), np.float64)
    _, u32 = rng.proposal_stream(3, 0, 50, np.arange(40), np.float32)
    assert u64.dtype == np.float64 and u32.dtype == np.float32
    assert (u64 >= 0).all() and (u64 < 1).all() and (u32 >= 0).all() and (u32 < 1).all()
    idx = rng.index_from_word(w[..., 0], 20)
    assert idx.min() >= 0 and idx.max() < 20
    assert abs(u64.mean() - 0.5) < 0.05


def test_spin_index_convention():
    """spin.py:165-171: index 0 <-> +1, index 1 <-> -1; state 0 of all_states is all-up."""
    assert np.array_equal(hilbert.states_to_local_indices(np.array([1, -1])), [0, 1])
    assert np.array_equal(hilbert.local_indices_to_states(np.array([0, 1])), [1, -1])
    st = hilbert.all_states(3)
    assert np.array_equal(st[0], [1, 1, 1]) and np.array_equal(st[1], [1, 1, -1]) and np.array_equal(st[-1], [-1, -1, -1])
    assert np.array_equal(hilbert.states_to_numbers(st, 3), np.arange(8))
    assert len(hilbert.all_states(6, total_sz=0)) == 20


def test_random_state_in_space():
    """test/hilbert/test_hilbert.py:180: random states live in the (constrained) space."""
    s = hilbert.random_state(1, 500, 10)
    assert s.dtype == np.int8 and set(np.unique(s)) == {-1, 1}
    assert abs(s.mean()) < 0.1
    for tsz in (0, 1, -2):
        s = hilbert.random_state(1, 200, 10, total_sz=tsz)
        assert np.all(s.sum(axis=1) == 2 * tsz)
    # uniform over the C(4,2)=6 constrained states
    s = hilbert.random_state(7, 6000, 4, total_sz=0)
    _, counts = np.unique(hilbert.states_to_numbers(s, 4), return_counts=True)
    assert len(counts) == 6 and sstats.chisquare(counts).pvalue > 1e-3


def test_lanczos_doctest_eigenvalues():
    """netket/exact.py:61-63: lanczos_ed(Ising(h=1, Chain(8)), k=3)."""
    e, _ = graph.hypercube_edges(8, 1)
    w = ed.full_ed(lambda x: ops.ising_conn_padded(x, e, 1.0, 1.0), 8, k=3)
    np.testing.assert_allclose(w, [-10.25166179, -10.05467898, -8.69093921], atol=1e-8)


def test_heisenberg_constrained_equals_full_ground_energy():
    """test/exact/test_groundstate.py:73-97."""
    e, c = graph.hypercube_edges(8, 1)
    t = ops.heisenberg_tables(e, c, 1.0, True)
    f = lambda x: ops.local_operator_conn_padded(x, t)  # noqa: E731
    w_full = ed.full_ed(f, 8, k=1)
    w_con = ed.full_ed(f, 8, total_sz=0, k=1)
    np.testing.assert_allclose(w_full, w_con, atol=1e-10)
    np.testing.assert_allclose(w_con[0], -14.60437363, atol=1e-6)  # 4 * (-3.651093408) Bethe-ansatz N=8 value


@pytest.mark.parametrize("L,h,J", [(10, 1.321, 1.0), (6, 0.0, 2.0), (7, 0.3, -1.5)])
def test_ising_dense_and_invariants(L, h, J):
    """dense equality with an independent Kronecker construction (test_operator.py:393-399), hermiticity (:235-271),
    n_conn == number of nonzero mels (:742-757), rank/dtype contract (:345-359)."""
    e, _ = graph.hypercube_edges(L, 1)
    f = lambda x: ops.ising_conn_padded(x, e, h, J)  # noqa: E731
    H = ops.to_dense(f, L)
    sx = np.array([[0.0, 1.0], [1.0, 0.0]])
    Hk = ops.kron_dense(L, [(i, -h * sx) for i in range(L)], [(a, b, J * ops.SZ_SZ) for a, b in e.tolist()])
    np.testing.assert_allclose(H, Hk, atol=1e-13)
    np.testing.assert_allclose(H, H.T, atol=1e-13)
    st = hilbert.all_states(L)
    xp, mels = f(st)
    assert xp.dtype == st.dtype and mels.dtype == np.float64
    assert xp.shape == (2 ** L, 1 if h == 0 else L + 1, L)
    assert np.array_equal((mels != 0).sum(axis=1), ops.ising_n_conn(st, e, h, J))
    v = np.ones((2, 3, L))
    vp, m = f(v)
    assert vp.shape == (2, 3, xp.shape[1], L) and m.shape == (2, 3, xp.shape[1]) and vp.dtype == v.dtype


@pytest.mark.parametrize("J,sign_rule,order,n_dim,L", [(1.0, True, 1, 1, 8), (1.0, False, 1, 1, 8), ([1.0, 0.5], [False, False], 2, 2, 3),
                                                        ([1.0, 2.0], [True, False], 2, 1, 6)])
def test_heisenberg_dense_and_padding(J, sign_rule, order, n_dim, L):
    pbc = not (n_dim == 2 and L == 3)
    e, c = graph.hypercube_edges(L, n_dim, pbc=True, max_neighbor_order=order)
    N = L ** n_dim
    t = ops.heisenberg_tables(e, c, J, sign_rule)
    f = lambda x: ops.local_operator_conn_padded(x, t)  # noqa: E731
    H = ops.to_dense(f, N)
    Js = J if isinstance(J, list) else [J]
    srs = sign_rule if isinstance(sign_rule, list) else [sign_rule]
    Hk = ops.kron_dense(N, [], [(a, b, Js[col] * (ops.SZ_SZ - ops.EXCHANGE if srs[col] else ops.SZ_SZ + ops.EXCHANGE))
                                for (a, b), col in zip(e.tolist(), c.tolist())])
    np.testing.assert_allclose(H, Hk, atol=1e-13)
    st = hilbert.all_states(N)
    xp, mels, nconn = f(st)
    assert mels.shape[1] == t["max_conn_size"]
    for r in range(len(st)):  # padding = trailing zeros with x' = x (test_operator.py:507-556)
        k = nconn[r]
        assert np.all(np.abs(mels[r, :k]) > 1e-10) and np.all(mels[r, k:] == 0)
        assert np.all(xp[r, k:] == st[r])
    assert pbc or True


def test_local_operator_canonicalisation():
    """Unsorted supports are permuted (helpers.py:151-213); equal supports are summed (base.py:136-147)."""
    rs = np.random.default_rng(0)
    A = rs.normal(size=(4, 4))
    d = ops.canonical_operators_dict([A], [(3, 1)])
    assert list(d.keys()) == [(1, 3)]
    SW = np.eye(4)[[0, 2, 1, 3]]
    np.testing.assert_allclose(d[(1, 3)], SW @ A @ SW)
    d = ops.canonical_operators_dict([A, A], [(0, 1), (1, 0)])
    np.testing.assert_allclose(d[(0, 1)], A + SW @ A @ SW)
    H = ops.to_dense(lambda x: ops.local_operator_conn_padded(x, ops.pack_internals(ops.canonical_operators_dict([A + A.T], [(2, 0)]))), 3)
    Hk = ops.kron_dense(3, [], [(0, 2, SW @ (A + A.T) @ SW)])
    np.testing.assert_allclose(H, Hk, atol=1e-13)


def test_log_cosh_and_logpsi():
    x = np.linspace(-30, 30, 601)
    np.testing.assert_allclose(rbm.log_cosh(x), np.log(np.cosh(x)), rtol=1e-12, atol=1e-13)
    W, b, a = rbm.init_params(6, 2, std=0.4)
    st = hilbert.all_states(6)
    ref = np.log(np.cosh(st @ W + b)).sum(axis=1) + st @ a
    np.testing.assert_allclose(rbm.logpsi(st, W, b, a), ref, rtol=1e-12)


def _chi2_p(samples, p_exact, N):
    nums = hilbert.states_to_numbers(samples.reshape(-1, N), N)
    counts = np.bincount(nums, minlength=len(p_exact))
    keep = p_exact * counts.sum() > 5
    return sstats.chisquare(counts[keep], p_exact[keep] / p_exact[keep].sum() * counts[keep].sum()).pvalue


@pytest.mark.parametrize("rule", ["local", "exchange"])
def test_sampler_chi_square(rule):
    """test/sampler/test_sampler.py:399-457: histogram of samples vs exact |psi|^2 on a 4-site chain."""
    N = 4
    W, b, a = rbm.init_params(N, 2, std=0.5, seed=1234)
    total_sz = 0 if rule == "exchange" else None
    e, _ = graph.hypercube_edges(N, 1)
    clusters = graph.compute_clusters(N, e, 1) if rule == "exchange" else None
    states = hilbert.all_states(N)
    p = sampler.exact_distribution(W, b, a, states)
    if total_sz is not None:
        keep = states.sum(axis=1) == 0
        p = np.where(keep, p, 0.0)
        p /= p.sum()
    sig = hilbert.random_state(15324, 64, N, total_sz)
    out = sampler.sample_chain(rule, sig, W, b, a, chain_length=240, seed=15324, clusters=clusters, sweep_size=4 * N)
    pv = _chi2_p(out["samples"][:, 40:], p, N)
    assert pv > 0.005, pv
    # returned log-probs equal recomputed machine_pow * logpsi (test_sampler.py:342-361)
    np.testing.assert_allclose(out["log_prob_samples"], 2 * rbm.logpsi(out["samples"], W, b, a), rtol=1e-12)
    assert 0 < out["n_accepted"].sum() <= out["n_steps"]


def test_expect_within_5_sigma_of_exact():
    """test/variational/test_variational.py:362-408 with Metropolis samples: <H> from E_loc vs psi^T H psi."""
    N = 6
    e, _ = graph.hypercube_edges(N, 1)
    W, b, a = rbm.init_params(N, 2, std=0.3)
    f = lambda x: ops.ising_conn_padded(x, e, 1.0, 1.0)  # noqa: E731
    psi = rbm.to_array(W, b, a, hilbert.all_states(N))
    exact = ed.expectation(f, N, psi)
    sig = hilbert.random_state(3, 128, N)
    out = sampler.sample_chain("local", sig, W, b, a, chain_length=80, seed=3)
    eloc = estimators.local_estimators(out["samples"][:, 16:], f, W, b, a)
    st = ostats.statistics(eloc)
    assert abs(st["mean"] - exact) < 5 * st["error_of_mean"], (st, exact)
    # zero-variance property: an exact eigenstate has constant local energy
    w, v = ed.full_ed(f, N, compute_eigenvectors=True)
    H = ops.to_dense(f, N)
    np.testing.assert_allclose((H @ v[:, 0]) / v[:, 0], w[0], rtol=1e-9)


def test_statistics_identities():
    """test/stats/test_stats.py:60-84: mean/var identity; shapes; NaN rules."""
    rs = np.random.default_rng(0)
    x = rs.normal(size=(16, 63))
    st = ostats.statistics(x)
    np.testing.assert_allclose(st["mean"], x.mean())
    np.testing.assert_allclose(st["variance"], x.var())
    # 16 chains < 32 batches -> not batch_good; l_block = 1 -> n_blocks = 1008 >= 32 -> block_good
    np.testing.assert_allclose(st["error_of_mean"], np.sqrt(x.var() / x.size))
    assert np.isnan(ostats.statistics(x[:4, :5])["error_of_mean"])  # neither estimate is good
    st1 = ostats.statistics(x[0])
    assert np.isnan(st1["R_hat"])
    big = rs.normal(size=(64, 100))
    stb = ostats.statistics(big)
    np.testing.assert_allclose(stb["error_of_mean"], np.sqrt(big.mean(axis=1).var() / 64))
    assert 0.9 < stb["R_hat"] < 1.1


def test_rbm_log_derivatives_match_finite_differences():
    """Pins oracle/forces.py: the closed-form dlogpsi/dp equals the numerical derivative of oracle.rbm.logpsi, and the
    forces follow expect_forces.py:81-104 (centred local energies, division by n_samples)."""
    from oracle import forces as oforces

    N, alpha = 5, 2
    W, b, a = rbm.init_params(N, alpha, std=0.4)
    sig = hilbert.random_state(1, 7, N)
    OW, Ob, Oa = oforces.log_derivatives(sig, W, b, a)
    h = 1e-6
    for (i, j) in [(0, 0), (3, 7), (4, 9)]:
        Wp, Wm = W.copy(), W.copy()
        Wp[i, j] += h
        Wm[i, j] -= h
        fd = (rbm.logpsi(sig, Wp, b, a) - rbm.logpsi(sig, Wm, b, a)) / (2 * h)
        np.testing.assert_allclose(OW[:, i, j], fd, rtol=1e-7, atol=1e-8)
    for j in (0, 5):
        bp, bm = b.copy(), b.copy()
        bp[j] += h
        bm[j] -= h
        np.testing.assert_allclose(Ob[:, j], (rbm.logpsi(sig, W, bp, a) - rbm.logpsi(sig, W, bm, a)) / (2 * h), rtol=1e-7, atol=1e-8)
    ap, am = a.copy(), a.copy()
    ap[2] += h
    am[2] -= h
    np.testing.assert_allclose(Oa[:, 2], (rbm.logpsi(sig, W, b, ap) - rbm.logpsi(sig, W, b, am)) / (2 * h), rtol=1e-7, atol=1e-8)
    e = np.random.default_rng(0).normal(size=7)
    f = oforces.forces(sig, e, W, b, a)
    w = (e - e.mean()) / 7
    np.testing.assert_allclose(f["W"], np.einsum("s,sij->ij", w, OW), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(f["b"], w @ Ob, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(f["a"], w @ Oa, rtol=1e-12, atol=1e-14)
    # exact gradient of <H>: for exact sampling (all states weighted by |psi|^2) 2 F equals d<H>/dp by finite differences
    e_g, _ = graph.hypercube_edges(N, 1)
    conn = lambda x: ops.ising_conn_padded(x, e_g, 1.0, 1.0)  # noqa: E731
    states = hilbert.all_states(N)

    def energy(Wx):
        psi = rbm.to_array(Wx, b, a, states)
        return ed.expectation(conn, N, psi)

    p = sampler.exact_distribution(W, b, a, states)
    eloc = estimators.local_value_kernel(states, conn, W, b, a)
    OWs, _, _ = oforces.log_derivatives(states, W, b, a)
    mean = float(p @ eloc)
    g_exact = 2.0 * np.einsum("s,sij->ij", p * (eloc - mean), OWs)
    Wp, Wm = W.copy(), W.copy()
    Wp[1, 3] += h
    Wm[1, 3] -= h
    np.testing.assert_allclose(g_exact[1, 3], (energy(Wp) - energy(Wm)) / (2 * h), rtol=1e-6, atol=1e-8)


def test_qgt_oracle_is_consistent():
    """oracle/qgt.py: the Jacobian equals finite differences of logpsi, the dense S equals the reference's matrix-free recipe
    (qgt_onthefly_logic.py:33-43), S is symmetric positive semi-definite and annihilates nothing it should not."""
    from oracle import qgt as oqgt, rbm as orbm

    rs = np.random.default_rng(5)
    N, M = 6, 9
    W, b, a = rs.normal(size=(N, M)) * 0.3, rs.normal(size=M) * 0.2, rs.normal(size=N) * 0.2
    sig = (1 - 2 * rs.integers(0, 2, size=(64, N))).astype(np.int8)
    O = oqgt.jacobian(sig, W, b, a)
    p = np.concatenate([W.ravel(), b, a])

    def lp(q):
        return orbm.logpsi(sig, q[:N * M].reshape(N, M), q[N * M:N * M + M], q[N * M + M:])

    for k in (0, 7, N * M + 3, N * M + M + 2):
        e = np.zeros_like(p)
        e[k] = 1e-6
        np.testing.assert_allclose((lp(p + e) - lp(p - e)) / 2e-6, O[:, k], atol=1e-8)
    S = oqgt.qgt_dense(sig, W, b, a, 0.0)
    np.testing.assert_allclose(S, S.T, atol=1e-14)
    assert np.linalg.eigvalsh(S).min() > -1e-12
    v = rs.normal(size=p.size)
    np.testing.assert_allclose(S @ v + 0.01 * v, oqgt.mat_vec(sig, W, b, a, v, 0.01), atol=1e-13)
    g = rs.normal(size=p.size)
    x = oqgt.sr_solve(sig, W, b, a, g, 0.01)
    np.testing.assert_allclose(oqgt.mat_vec(sig, W, b, a, x, 0.01), g, atol=1e-10)
    # without biases the Jacobian keeps the kernel block only
    assert oqgt.jacobian(sig, W, None, None).shape == (64, N * M)
