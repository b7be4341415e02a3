"""Stochastic Reconfiguration on the GPU (SR / QGT item): the matrix-free quantum geometric tensor
(nk_rbm_jvp + nk_forces_rbm) against the dense definition of oracle/qgt.py, the SR solve against numpy's, and the VMC driver
(netket/driver/vmc.py) with SGD and with SR.  Tolerances: 1e-10 relative to the largest entry in fp64, 2e-5 in fp32."""

import math

import numpy as np
import pytest
import torch

from oracle import qgt as oqgt

pytestmark = pytest.mark.gpu


def make_state(nk, L=6, alpha=2, dtype=np.float64, n_chains=64, n_samples=512, use_hidden_bias=True, use_visible_bias=True, seed=3):
    g = nk.graph.Chain(L)
    hi = nk.hilbert.Spin(0.5, L)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=n_chains),
                        nk.models.RBM(alpha=alpha, param_dtype=dtype, use_hidden_bias=use_hidden_bias, use_visible_bias=use_visible_bias),
                        n_samples=n_samples, seed=seed)
    return vs, nk.operator.Ising(hi, g, h=1.0)


def host_params(nk, vs):
    W, b, a = nk.models.RBM.unpack(vs.variables)
    f = lambda t: None if t is None else t.cpu().numpy().astype(np.float64)  # noqa: E731
    return f(W), f(b), f(a)


@pytest.mark.parametrize("dtype,biases,with_tanh", [(np.float64, (True, True), True), (np.float64, (True, True), False),
                                                    (np.float64, (False, False), True), (np.float64, (True, False), True),
                                                    (np.float32, (True, True), True), (np.float32, (False, True), False)])
def test_qgt_matvec_and_dense_match_oracle(cuda, dtype, biases, with_tanh):
    import netket_b200 as nk
    from netket_b200.optimizer import QGTOnTheFly, tree_to_flat

    vs, H = make_state(nk, dtype=dtype, use_hidden_bias=biases[0], use_visible_bias=biases[1])
    if with_tanh:
        vs.expect_and_grad(H)     # samples + tanh(theta) written by the sweep kernel
    else:
        vs.sample()               # samples only: the QGT recomputes tanh(theta) with one theta GEMM
    S = QGTOnTheFly(vs, diag_shift=0.02)
    W, b, a = host_params(nk, vs)
    ref = oqgt.qgt_dense(vs.samples.cpu().numpy(), W, b, a, 0.02)
    n = ref.shape[0]
    assert S.shape == (n, n) and n == vs.n_parameters
    tol = (1e-10 if dtype == np.float64 else 2e-5) * np.abs(ref).max()
    rs = np.random.default_rng(0)
    for _ in range(3):
        v = rs.normal(size=n)
        got = (S @ torch.from_numpy(v).to(cuda)).cpu().numpy()
        np.testing.assert_allclose(got, ref @ v, rtol=0, atol=tol * np.abs(v).max() * 4)
    np.testing.assert_allclose(S.to_dense().cpu().numpy(), ref, rtol=0, atol=tol)
    # pytree in, pytree out
    tree = {k: ({kk: torch.ones_like(vv) for kk, vv in v.items()} if isinstance(v, dict) else torch.ones_like(v)) for k, v in vs.parameters.items()}
    out = S @ tree
    assert set(out) == set(vs.parameters) and out["Dense"]["kernel"].shape == vs.parameters["Dense"]["kernel"].shape
    np.testing.assert_allclose(tree_to_flat(out).cpu().numpy(), ref @ np.ones(n), rtol=0, atol=tol * 4)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sr_solve_matches_dense_solve(cuda, dtype):
    import netket_b200 as nk
    from netket_b200.optimizer import SR, cg, tree_to_flat
    from functools import partial

    vs, H = make_state(nk, dtype=dtype, n_samples=2048)
    E, grad = vs.expect_and_grad(H)
    sr = SR(diag_shift=0.01, solver=partial(cg, tol=1e-10 if dtype == np.float64 else 1e-6))
    dp = sr(vs, grad)
    W, b, a = host_params(nk, vs)
    ref = oqgt.sr_solve(vs.samples.cpu().numpy(), W, b, a, tree_to_flat(grad).cpu().numpy(), 0.01)
    got = tree_to_flat(dp).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=0, atol=(1e-7 if dtype == np.float64 else 2e-3) * np.abs(ref).max())
    assert sr.info["n_iter"] > 0 and sr._lhs.n_matvec >= sr.info["n_iter"]
    # the next call starts from the previous solution (solver_restart=False): same system -> no further iterations needed
    it0 = sr.info["n_iter"]
    sr(vs, grad)
    assert sr.info["n_iter"] <= max(2, it0 // 4)
    with pytest.raises(TypeError, match="scheduled `diag_shift`"):
        SR(diag_shift=lambda step: 0.1)(vs, grad)
    assert SR(diag_shift=lambda step: 0.1 / (1 + step))(vs, grad, 3) is not None
    with pytest.raises(NotImplementedError, match="diag_scale"):
        SR(diag_scale=0.1)(vs, grad)


def exact_ground_energy(nk, L, h):
    from oracle import ed as oed, operators as oops

    edges = np.asarray(nk.graph.Chain(L).edges(), dtype=np.int32)
    return float(oed.full_ed(lambda x: oops.ising_conn_padded(x, edges, h, 1.0), L, k=1)[0])


def test_vmc_driver_sgd_and_sr_lower_the_energy(cuda):
    """Examples/Ising1d: VMC with SGD, then with the SR preconditioner; SR gets closer to the exact ground energy in the
    same number of steps (the reference's test/groundstate/test_vmc.py checks convergence to ED the same way)."""
    import netket_b200 as nk

    L = 8
    e0 = exact_ground_energy(nk, L, 1.0)
    results = {}
    for name, pre in (("sgd", None), ("sr", nk.optimizer.SR(diag_shift=0.1))):
        vs, H = make_state(nk, L=L, alpha=1, n_chains=256, n_samples=4096, seed=7)
        drv = nk.driver.VMC(H, nk.optimizer.Sgd(learning_rate=0.1), variational_state=vs, preconditioner=pre)
        log = nk.driver.RuntimeLog()
        e_start = vs.expect(H).mean
        drv.run(100, out=log, show_progress=False)
        assert drv.step_count == 100 and len(log["Energy"]["Mean"]) == 100
        assert log["Energy"]["Mean"].iters[:3] == [0, 1, 2]
        e_end = float(np.mean(log["Energy"]["Mean"].values[-5:]))
        assert e_end < e_start - 0.5
        results[name] = e_end
        assert math.isfinite(drv.energy.mean) and drv.energy.error_of_mean > 0
    assert results["sr"] < results["sgd"] + 0.05
    assert abs(results["sr"] - e0) / abs(e0) < 0.03, (results, e0)
    assert all(r > e0 - 0.2 for r in results.values())     # variational within the MC error


def test_driver_api(cuda):
    import netket_b200 as nk

    vs, H = make_state(nk)
    drv = nk.driver.VMC(H, nk.optimizer.Momentum(0.02, beta=0.8), variational_state=vs)
    p0 = vs.parameters["Dense"]["kernel"].clone()
    steps = list(drv.iter(6, 2))
    assert steps == [0, 2, 4] and drv.step_count == 6
    assert not torch.equal(p0, vs.parameters["Dense"]["kernel"])
    drv.advance(2)
    assert drv.step_count == 8
    stopped = []
    drv.run(10, show_progress=False, callback=lambda step, log, d: (stopped.append(step) or len(stopped) < 3))
    assert len(stopped) == 3
    out = drv.run(2, out=[nk.driver.RuntimeLog(), nk.driver.RuntimeLog()], obs={"H2": H}, show_progress=False)
    assert len(out) == 2 and "H2" in out[0].data and "Energy" in out[1].data
    other = nk.hilbert.Spin(0.5, 4)
    with pytest.raises(TypeError, match="should match"):
        nk.driver.VMC(nk.operator.Ising(other, nk.graph.Chain(4), h=1.0), nk.optimizer.Sgd(0.1), variational_state=vs)
    with pytest.raises(ValueError, match="must be a number"):
        drv.run("ten")
    with pytest.warns(UserWarning, match="rank deficient"):
        small, Hs = make_state(nk, n_chains=16, n_samples=16)
        nk.driver.VMC(Hs, nk.optimizer.Sgd(0.1), variational_state=small, preconditioner=nk.optimizer.SR())


@pytest.mark.parametrize("N,M,B,dtype,biases", [
    (10, 20, 300, np.float32, (True, True)),      # tcgen05 GEMM, one column tile
    (20, 320, 1000, np.float32, (True, True)),    # two column tiles
    (100, 400, 4096, np.float32, (True, True)),   # cfg-3 shape
    (33, 52, 257, np.float32, (False, True)),     # odd N, no hidden bias
    (16, 18, 130, np.float32, (True, False)),     # M % 4 != 0: scalar stores; no visible bias
    (6, 12, 200, np.float32, (True, True)),       # M < 16: GEMM on CUDA cores + separate row dot
    (20, 40, 500, np.float64, (True, True)),      # fp64: DMMA GEMM + separate row dot
])
def test_rbm_jvp_matches_oracle_jacobian(cuda, N, M, B, dtype, biases):
    """nk_rbm_jvp through the C ABI: y = O v and its sum, against the oracle's Jacobian (every kernel route of the product)."""
    import ctypes as C
    from netket_b200 import _lib

    rs = np.random.default_rng(N * 1000 + M)
    W, b, a = rs.normal(size=(N, M)) * 0.2, rs.normal(size=M) * 0.2, rs.normal(size=N) * 0.2
    V, vb, va = rs.normal(size=(N, M)), rs.normal(size=M) if biases[0] else None, rs.normal(size=N) if biases[1] else None
    sig = (1 - 2 * rs.integers(0, 2, size=(B, N))).astype(np.int8)
    O = oqgt.jacobian(sig, W, b, a)                       # columns [W | b | a]
    v = np.concatenate([V.ravel(), vb if vb is not None else np.zeros(M), va if va is not None else np.zeros(N)])
    want = O @ v
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    dev = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x)).to(device=cuda, dtype=tdt)  # noqa: E731
    tanh = dev(np.tanh(sig.astype(np.float64) @ W + b))
    Vd, vbd, vad, s8 = dev(V), dev(vb), dev(va), torch.from_numpy(sig).to(cuda)
    vr = _lib.nk_rbm_t(W=_lib.ptr(Vd), b=_lib.ptr(vbd) if vbd is not None else None, a=_lib.ptr(vad) if vad is not None else None, N=N, M=M,
                       dtype=_lib.dtype_code(tdt), reserved=0)
    L = _lib.lib()
    ws = torch.empty(max(int(L.nk_theta_gemm_workspace_bytes(C.byref(vr), B)), 1), dtype=torch.uint8, device=cuda)
    scratch = torch.empty((B, M), dtype=tdt, device=cuda)
    y = torch.full((B,), 7.0, dtype=torch.float64, device=cuda)
    ysum = torch.full((1,), 7.0, dtype=torch.float64, device=cuda)
    with torch.cuda.device(cuda):
        _lib.check(L.nk_rbm_jvp(_lib.stream_ptr(cuda), C.byref(vr), _lib.ptr(s8), B, _lib.ptr(tanh), _lib.ptr(y), _lib.ptr(ysum),
                                _lib.ptr(scratch), _lib.ptr(ws)))
    tol = (1e-11 if dtype == np.float64 else 3e-6) * np.abs(want).max() * np.sqrt(M)
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=0, atol=tol)
    np.testing.assert_allclose(ysum.item(), want.sum(), rtol=0, atol=tol * np.sqrt(B) * 4)
