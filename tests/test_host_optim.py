"""Host logic of the optimiser / driver / streaming callers on CPU tensors (no kernels involved): the conjugate-gradient
solver with jax.scipy's stopping rule, parameter-tree <-> flat-vector maps, optimiser updates, loggers, stopping tests."""

import math

import numpy as np
import pytest
import torch

from netket_b200 import convergence as conv
from netket_b200 import driver as drv
from netket_b200 import optimizer as opt


def spd(n, seed=0, cond=50.0):
    rs = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rs.normal(size=(n, n)))
    return (q * np.geomspace(1.0, cond, n)) @ q.T


def test_cg_matches_dense_solve_and_stopping_rule():
    A = torch.from_numpy(spd(40))
    b = torch.from_numpy(np.random.default_rng(1).normal(size=40))
    x, info = opt.cg(lambda v: A @ v, b, tol=1e-12)
    np.testing.assert_allclose(x.numpy(), np.linalg.solve(A.numpy(), b.numpy()), rtol=1e-9)
    assert 0 < info["n_iter"] <= 40 * 10 and info["residual"] <= 1e-12 * float(b.norm()) * 1.0001
    # loose tolerance: fewer iterations, residual below tol * |b| (jax.scipy.sparse.linalg.cg: |r| <= max(tol |b|, atol))
    x2, info2 = opt.cg(lambda v: A @ v, b, tol=1e-2)
    assert info2["n_iter"] < info["n_iter"] and float((A @ x2 - b).norm()) <= 1e-2 * float(b.norm()) * 1.01
    # atol dominates; maxiter caps; a converged x0 needs no iteration
    _, info3 = opt.cg(lambda v: A @ v, b, tol=0.0, atol=0.5 * float(b.norm()))
    assert info3["residual"] <= 0.5 * float(b.norm())
    _, info4 = opt.cg(lambda v: A @ v, b, tol=1e-14, maxiter=3)
    assert info4["n_iter"] == 3
    _, info5 = opt.cg(lambda v: A @ v, b, x0=x, tol=1e-8)
    assert info5["n_iter"] == 0
    xz, infoz = opt.cg(lambda v: A @ v, torch.zeros(40, dtype=torch.float64))
    assert infoz["n_iter"] == 0 and float(xz.abs().max()) == 0.0


@pytest.mark.parametrize("hidden_bias,visible_bias", [(True, True), (False, True), (True, False), (False, False)])
def test_parameter_tree_round_trip(hidden_bias, visible_bias):
    N, M = 5, 7
    tree = {"Dense": {"kernel": torch.arange(N * M, dtype=torch.float32).reshape(N, M)}}
    if hidden_bias:
        tree["Dense"]["bias"] = torch.arange(M, dtype=torch.float32) + 100
    if visible_bias:
        tree["visible_bias"] = torch.arange(N, dtype=torch.float32) + 200
    flat = opt.tree_to_flat(tree)
    assert flat.dtype == torch.float64 and flat.numel() == N * M + (M if hidden_bias else 0) + (N if visible_bias else 0)
    assert flat[N * M - 1] == N * M - 1 and (not hidden_bias or flat[N * M] == 100) and (not visible_bias or flat[-1] == 200 + N - 1)
    back = opt.flat_to_tree(flat, tree)
    assert set(back) == set(tree) and set(back["Dense"]) == set(tree["Dense"])
    assert back["Dense"]["kernel"].dtype == torch.float32 and torch.equal(back["Dense"]["kernel"], tree["Dense"]["kernel"])
    if visible_bias:
        assert torch.equal(back["visible_bias"], tree["visible_bias"])


def test_sgd_and_momentum_updates():
    p = {"Dense": {"kernel": torch.ones(2, 3)}, "visible_bias": torch.zeros(2)}
    g = {"Dense": {"kernel": torch.full((2, 3), 2.0)}, "visible_bias": torch.ones(2)}
    sgd = opt.Sgd(0.1)
    st = sgd.init(p)
    upd, st = sgd.update(g, st, p)
    new = opt.apply_updates(p, upd)
    assert torch.allclose(new["Dense"]["kernel"], torch.full((2, 3), 0.8)) and torch.allclose(new["visible_bias"], torch.full((2,), -0.1))
    assert st["count"] == 1
    sched = opt.Sgd(lambda step: 1.0 / (1 + step))
    s = sched.init(p)
    u0, s = sched.update(g, s, p)
    u1, s = sched.update(g, s, p)
    assert torch.allclose(u0["visible_bias"], torch.full((2,), -1.0)) and torch.allclose(u1["visible_bias"], torch.full((2,), -0.5))
    mom = opt.Momentum(0.1, beta=0.5)
    ms = mom.init(p)
    m0, ms = mom.update(g, ms, p)
    m1, ms = mom.update(g, ms, p)
    assert torch.allclose(m0["visible_bias"], torch.full((2,), -0.1)) and torch.allclose(m1["visible_bias"], torch.full((2,), -0.15))
    assert opt.identity_preconditioner(None, g) is g


def test_history_and_runtime_log():
    h = conv.HistoryDict()
    h.push({"a": 1.0, "b": 2.0}, step=0).push({"a": 3.0}, step=5)
    assert h["a"].iters == [0, 5] and h["a"].values == [1.0, 3.0] and len(h["b"]) == 1

    class FakeStats:
        def to_dict(self):
            return {"Mean": -1.5, "Sigma": 0.1}

    log = drv.RuntimeLog()
    log(0, {"Energy": FakeStats(), "acc": 0.5})
    log(1, {"Energy": FakeStats(), "acc": 0.6})
    assert log["Energy"]["Mean"].values == [-1.5, -1.5] and log["Energy"]["Sigma"].iters == [0, 1] and log.data["acc"].values == [0.5, 0.6]
    log.flush()


def test_stopping_tests_of_expect_to_precision():
    class Acc:
        def __init__(self, mean, err):
            self._s = type("S", (), {"mean": mean, "error_of_mean": err})()

        def get_stats(self):
            return self._s

    assert conv._rel_err(0.0, 0.0) == 0.0 and conv._rel_err(0.1, 0.0) == math.inf and conv._rel_err(0.1, 2.0) == 0.05
    assert conv._not_converged(Acc(-10.0, 0.2), atol=0.1, rtol=None)
    assert not conv._not_converged(Acc(-10.0, 0.05), atol=0.1, rtol=None)
    assert conv._not_converged(Acc(-10.0, 0.05), atol=0.1, rtol=1e-3)          # both tolerances must hold
    assert not conv._not_converged(Acc(-10.0, 0.005), atol=0.1, rtol=1e-3)
    assert not conv._not_converged(Acc(-10.0, math.nan), atol=0.1, rtol=None)   # NaN > atol is False, as in the reference
    assert conv._postfix(Acc(-10.0, 0.05), 0.1, 1e-2) == {"err": "0.05", "atol": "0.1", "rel_err": "0.005", "rtol": "0.01"}

    class NotMetropolis:
        sampler = object()

    for fn in (conv.expect_to_precision, conv.check_mc_convergence, conv.thermalise_mcmc):
        with pytest.raises(ValueError, match="MetropolisSampler"):
            fn(NotMetropolis(), None, **({"atol": 0.1} if fn is conv.expect_to_precision else {}))
    with pytest.raises(ValueError, match="atol.*rtol"):
        conv.expect_to_precision(NotMetropolis(), None)
    with pytest.raises(ValueError, match="rtol must be > 0"):
        conv.expect_to_precision(NotMetropolis(), None, rtol=0.0)


class QuadraticState:
    """A stand-in variational state whose 'energy' is |p - target|^2 / 2 (gradient p - target): the driver's loop logic
    (netket/driver/vmc.py:141-161, abstract_variational_driver.py:349-528) runs on CPU tensors."""

    hilbert = "hi"
    n_samples, n_parameters = 100, 8

    def __init__(self):
        self.parameters = {"Dense": {"kernel": torch.zeros(2, 3, dtype=torch.float64)}, "visible_bias": torch.zeros(2, dtype=torch.float64)}
        self.target = {"Dense": {"kernel": torch.full((2, 3), 2.0, dtype=torch.float64)}, "visible_bias": torch.full((2,), -1.0, dtype=torch.float64)}
        self.resets = 0

    def reset(self):
        self.resets += 1

    def _loss(self):
        d = opt.tree_to_flat(self.parameters) - opt.tree_to_flat(self.target)
        return 0.5 * float(d @ d)

    def expect_and_grad(self, op):
        from netket_b200.stats import Stats

        g = opt._tree_map2(lambda p, t: p - t, self.parameters, self.target)
        return Stats(mean=self._loss(), error_of_mean=0.0, variance=0.0), g

    def expect(self, op):
        from netket_b200.stats import Stats

        return Stats(mean=self._loss())


class Op:
    hilbert = "hi"


def test_vmc_driver_loop_on_cpu():
    vs = QuadraticState()
    d = drv.VMC(Op(), opt.Sgd(0.5), variational_state=vs)
    log = drv.RuntimeLog()
    d.run(20, out=log, obs={"again": Op()}, show_progress=False)
    assert d.step_count == 20 and vs.resets >= 20
    e = log["Energy"]["Mean"]
    assert e.iters == list(range(20)) and all(b < a for a, b in zip(e.values, e.values[1:])) and e.values[-1] < 1e-9
    assert len(log["again"]["Mean"]) == 20
    assert torch.allclose(vs.parameters["visible_bias"], torch.full((2,), -1.0, dtype=torch.float64), atol=1e-5)
    # iter(): yields the step count every `step` steps; advance(); callbacks stop the run; step_size thins the log
    assert list(d.iter(6, 3)) == [20, 23] and d.step_count == 26
    d.advance(4)
    assert d.step_count == 30
    seen = []
    d.run(10, show_progress=False, callback=[lambda s, l, drv_: True, lambda s, l, drv_: (seen.append(s) or len(seen) < 2)])
    assert seen == [30, 31]
    log2 = drv.RuntimeLog()
    d.run(6, out=log2, step_size=2, show_progress=False)
    assert len(log2["Energy"]["Mean"]) == 3
    d.reset()
    assert d.step_count == 0
    # a preconditioner is called as (state, grad, step); None means identity
    calls = []

    def halve(state, grad, step=None):
        calls.append(step)
        return opt._tree_map(lambda g: 0.5 * g, grad)

    d2 = drv.VMC(Op(), opt.Sgd(1.0), variational_state=QuadraticState(), preconditioner=halve)
    d2.advance(3)
    assert calls == [0, 1, 2] and abs(d2.state.parameters["visible_bias"][0].item() + 1.0 * (1 - 0.5 ** 3)) < 1e-12
    d2.preconditioner = None
    assert d2.preconditioner is opt.identity_preconditioner

    class Other:
        hilbert = "other"

    with pytest.raises(TypeError, match="should match"):
        drv.VMC(Other(), opt.Sgd(0.1), variational_state=vs)
    with pytest.raises(ValueError, match="must be a number"):
        d.run("ten")
