"""``nk.optimizer``: plain SGD and the Stochastic-Reconfiguration preconditioner with a matrix-free quantum geometric
tensor (the SR / QGT item; netket/optimizer/sr.py:56-215, qgt/qgt_onthefly.py,
qgt/qgt_onthefly_logic.py:33-43, preconditioner.py).

``S v = O^H ((O v - mean(O v)) / n) + diag_shift v`` with ``O[s, :] = d log psi(sigma_s) / d p``.  For the RBM both halves
are kernels of libnkb200 on the samples and ``tanh(theta)`` that the sweep kernel already wrote:

* ``O v``   = ``nk_rbm_jvp``  (theta GEMM with the tangent ``(V, v_b)`` on the tensor cores + one HBM-bound row dot),
* ``O^H w`` = ``nk_forces_rbm`` (the tcgen05 / DMMA contraction of the forces) with ``w`` in place of ``E_loc - mean``.

Between GPUs one matrix-vector product all-reduces ``1 + n_parameters`` doubles.  The linear solve (conjugate gradients on
``n_parameters``-long float64 vectors) is host-driven vector algebra on the device, as in the reference (jax.scipy's cg).
"""

import ctypes as C

import torch

from . import _lib
from .models import RBM
from .stats import _allreduce
from .utils import world


# ------------------------------------------------------------------------------------------------ parameter trees
def tree_to_flat(tree, like=None):
    """``{"Dense": {"kernel", "bias"}, "visible_bias"}`` -> flat float64 vector ``[W | b | a]`` (the layout of nk_forces_rbm)."""
    parts = [tree["Dense"]["kernel"].reshape(-1)]
    if "bias" in tree["Dense"]:
        parts.append(tree["Dense"]["bias"].reshape(-1))
    if "visible_bias" in tree:
        parts.append(tree["visible_bias"].reshape(-1))
    return torch.cat([p.to(torch.float64) for p in parts])


def flat_to_tree(flat, like):
    """Inverse of :func:`tree_to_flat`; leaves take the dtype and shape of ``like``'s."""
    W = like["Dense"]["kernel"]
    N, M = W.shape
    pos = N * M
    dense = {"kernel": flat[:pos].reshape(N, M).to(W.dtype)}
    if "bias" in like["Dense"]:
        dense["bias"] = flat[pos:pos + M].to(W.dtype)
        pos += M
    out = {"Dense": dense}
    if "visible_bias" in like:
        out["visible_bias"] = flat[pos:pos + N].to(W.dtype)
    return out


# ------------------------------------------------------------------------------------------------ optimisers
class Sgd:
    """``nk.optimizer.Sgd(learning_rate)`` (netket/optimizer/__init__.py: optax.sgd): ``p <- p - learning_rate * dp``."""

    def __init__(self, learning_rate):
        self.learning_rate = learning_rate

    def init(self, params):
        return {"count": 0}

    def update(self, grads, state, params=None):
        lr = self.learning_rate(state["count"]) if callable(self.learning_rate) else self.learning_rate
        upd = _tree_map(lambda g: -lr * g, grads)
        return upd, {"count": state["count"] + 1}

    def __repr__(self):
        return f"Sgd(learning_rate={self.learning_rate})"


class Momentum(Sgd):
    """``nk.optimizer.Momentum(learning_rate, beta)``: ``m <- beta m + dp; p <- p - learning_rate m``."""

    def __init__(self, learning_rate, beta=0.9, nesterov=False):
        super().__init__(learning_rate)
        self.beta, self.nesterov = beta, nesterov

    def init(self, params):
        return {"count": 0, "trace": _tree_map(torch.zeros_like, params)}

    def update(self, grads, state, params=None):
        lr = self.learning_rate(state["count"]) if callable(self.learning_rate) else self.learning_rate
        trace = _tree_map2(lambda g, m: g + self.beta * m, grads, state["trace"])
        step = _tree_map2(lambda g, m: g + self.beta * m, grads, trace) if self.nesterov else trace
        return _tree_map(lambda m: -lr * m, step), {"count": state["count"] + 1, "trace": trace}


def _tree_map(f, t):
    return {k: (_tree_map(f, v) if isinstance(v, dict) else f(v)) for k, v in t.items()}


def _tree_map2(f, a, b):
    return {k: (_tree_map2(f, a[k], b[k]) if isinstance(a[k], dict) else f(a[k], b[k])) for k in a}


def apply_updates(params, updates):
    """optax.apply_updates."""
    return _tree_map2(lambda p, u: p + u.to(p.dtype), params, updates)


# ------------------------------------------------------------------------------------------------ linear solvers
def cg(A, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None):
    """Conjugate gradients with jax.scipy.sparse.linalg.cg's stopping rule (``|r| <= max(tol |b|, atol)``, ``maxiter``
    defaulting to 10 x size).  ``A`` is a callable on flat float64 device vectors.  Returns ``(x, info)`` with
    ``info = {"n_iter", "residual"}``."""
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    default_cap = maxiter is None
    if maxiter is None:
        maxiter = 10 * b.numel()
    bs = float(torch.dot(b, b))
    stop2 = max(tol * tol * bs, atol * atol)
    r = b - A(x) if x0 is not None else b.clone()
    p = r.clone()
    gamma = float(torch.dot(r, r))
    best = gamma
    k = 0
    while gamma > stop2 and k < maxiter:
        Ap = A(p)
        alpha = gamma / float(torch.dot(p, Ap))
        x.add_(p, alpha=alpha)
        r.add_(Ap, alpha=-alpha)
        gamma_new = float(torch.dot(r, r))
        p.mul_(gamma_new / gamma).add_(r)
        gamma = gamma_new
        k += 1
        if default_cap and k >= 2000 and gamma > stop2 and gamma >= 0.25 * best:
            # no factor-2 progress of |r| over the last 1000 iterations: the operator is only consistent to its working
            # precision (fp32 matvec) and the recurrence has stagnated above `tol`; stop instead of running to 10 x size
            import warnings

            warnings.warn(f"cg stagnated at |r| = {gamma ** 0.5:.3e} (target {stop2 ** 0.5:.3e}) after {k} iterations", RuntimeWarning,
                          stacklevel=2)
            break
        if k % 1000 == 0:
            best = gamma
    return x, {"n_iter": k, "residual": gamma ** 0.5, "converged": gamma <= stop2}


# ------------------------------------------------------------------------------------------------ QGT
class QGTOnTheFly:
    """Matrix-free ``S + diag_shift``: ``qgt @ v``, ``qgt.solve(solver, y)``, ``qgt.to_dense()`` (qgt/qgt_onthefly.py)."""

    def __init__(self, vstate, *, diag_shift=0.0, diag_scale=None, holomorphic=None, chunk_size=None, **kwargs):
        if diag_scale not in (None, 0, 0.0):
            raise NotImplementedError("QGTOnTheFly: diag_scale needs the diagonal of S, which a matrix-free QGT does not form "
                                      "(the reference's QGTOnTheFly raises for it as well)")
        if not isinstance(vstate.model, RBM):
            raise NotImplementedError("QGTOnTheFly: closed-form log-derivatives are implemented for netket_b200.models.RBM")
        self.diag_shift = float(diag_shift)
        self._vstate = vstate
        self._params = vstate.parameters
        variables = vstate.variables
        self._rbm = RBM.c_struct(variables)
        W, _, _ = RBM.unpack(variables)
        self._W = W
        dev = W.device
        N, M = self._rbm.N, self._rbm.M
        samples = vstate.samples
        self._s8 = samples.reshape(-1, N).contiguous()
        Ns = self._Ns = self._s8.shape[0]
        self._n_total = Ns * world()[1]
        L = _lib.lib()
        ws_bytes = max(int(L.nk_theta_gemm_workspace_bytes(C.byref(self._rbm), Ns)), int(L.nk_forces_workspace_bytes(C.byref(self._rbm), Ns)), 1)
        self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        tanh = vstate._tanh
        if tanh is None or tanh.numel() != Ns * M:  # samples drawn without tanh(theta): one theta GEMM for the batch
            tanh = torch.empty((Ns, M), dtype=W.dtype, device=dev)
            with torch.cuda.device(dev):
                _lib.check(L.nk_rbm_tanh_theta(_lib.stream_ptr(dev), C.byref(self._rbm), _lib.ptr(self._s8), Ns, _lib.ptr(tanh),
                                               _lib.ptr(self._ws)))
        self._tanh = tanh
        self._scratch = torch.empty((Ns, M), dtype=W.dtype, device=dev)
        self._y = torch.empty(Ns, dtype=torch.float64, device=dev)
        self._n = N * M + (M if "bias" in self._params["Dense"] else 0) + (N if "visible_bias" in self._params else 0)
        self.n_matvec = 0

    @property
    def shape(self):
        return (self._n, self._n)

    def _matvec_flat(self, v):
        """(S + diag_shift) v on a flat float64 vector."""
        L = _lib.lib()
        dev = self._W.device
        N, M = self._rbm.N, self._rbm.M
        tree = flat_to_tree(v, self._params)
        V = tree["Dense"]["kernel"].contiguous()
        vb = tree["Dense"].get("bias")
        va = tree.get("visible_bias")
        vr = _lib.nk_rbm_t(W=_lib.ptr(V), b=_lib.ptr(vb.contiguous()) if vb is not None else None,
                           a=_lib.ptr(va.contiguous()) if va is not None else None, N=N, M=M, dtype=_lib.dtype_code(V.dtype), reserved=0)
        sums = torch.empty(N * M + M + N, dtype=torch.float64, device=dev)
        head = torch.empty(1, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.check(L.nk_rbm_jvp(st, C.byref(vr), _lib.ptr(self._s8), self._Ns, _lib.ptr(self._tanh), _lib.ptr(self._y),
                                    _lib.ptr(head), _lib.ptr(self._scratch), _lib.ptr(self._ws)))
            _allreduce(head)
            mean = float(head.item()) / self._n_total
            _lib.check(L.nk_forces_rbm(st, C.byref(self._rbm), _lib.ptr(self._s8), self._Ns, _lib.ptr(self._y), _lib.NK_F64, mean,
                                       _lib.ptr(sums), None, _lib.ptr(self._tanh)))
            _allreduce(sums)
        self.n_matvec += 1
        pos, parts = N * M, [sums[:N * M]]
        if vb is not None:
            parts.append(sums[pos:pos + M])
        if va is not None:
            parts.append(sums[pos + M:])
        res = torch.cat(parts) if len(parts) > 1 else parts[0]
        return res / self._n_total + self.diag_shift * v

    def __matmul__(self, v):
        if isinstance(v, dict):
            return flat_to_tree(self._matvec_flat(tree_to_flat(v)), v)
        return self._matvec_flat(v.to(torch.float64))

    def solve(self, solve_fun, y, *, x0=None):
        """``x`` with ``(S + diag_shift) x = y`` (pytree in, pytree out) and the solver's info."""
        flat = isinstance(y, torch.Tensor)
        b = y.to(torch.float64) if flat else tree_to_flat(y)
        if x0 is not None and not isinstance(x0, torch.Tensor):
            x0 = tree_to_flat(x0)
        x, info = solve_fun(self._matvec_flat, b, x0=x0)
        return (x if flat else flat_to_tree(x, y)), info

    def to_dense(self):
        """The dense ``(n_parameters, n_parameters)`` matrix, column by column (small problems / tests)."""
        eye = torch.zeros(self._n, dtype=torch.float64, device=self._W.device)
        cols = []
        for k in range(self._n):
            eye.zero_()
            eye[k] = 1.0
            cols.append(self._matvec_flat(eye))
        return torch.stack(cols, dim=1)

    def __repr__(self):
        return f"QGTOnTheFly(diag_shift={self.diag_shift}, n_parameters={self._n}, n_samples={self._n_total})"


class SR:
    """Stochastic Reconfiguration / natural gradient: ``dp`` solves ``(S + diag_shift) dp = grad`` (sr.py:56-215,
    preconditioner.py:128-175: ``sr(vstate, grad, step)``; the previous solution seeds the next solve unless
    ``solver_restart``)."""

    def __init__(self, qgt=None, solver=cg, *, diag_shift=0.01, diag_scale=None, solver_restart=False, **kwargs):
        self.qgt_constructor = QGTOnTheFly if qgt is None else qgt
        self.solver = solver
        self.diag_shift, self.diag_scale = diag_shift, diag_scale
        self.solver_restart = solver_restart
        self.qgt_kwargs = kwargs
        self.x0 = None
        self.info = None
        self._lhs = None

    def lhs_constructor(self, vstate, step=None):
        shift, scale = self.diag_shift, self.diag_scale
        if callable(shift):
            if step is None:
                raise TypeError("If you use a scheduled `diag_shift`, you must call the preconditioner with an extra argument `step`.")
            shift = shift(step)
        if callable(scale):
            if step is None:
                raise TypeError("If you use a scheduled `diag_scale`, you must call the preconditioner with an extra argument `step`.")
            scale = scale(step)
        return self.qgt_constructor(vstate, diag_shift=shift, diag_scale=scale, **self.qgt_kwargs)

    def __call__(self, vstate, gradient, step=None):
        self._lhs = self.lhs_constructor(vstate, step)
        x0 = None if self.solver_restart else self.x0
        self.x0, self.info = self._lhs.solve(self.solver, gradient, x0=x0)
        return self.x0

    def __repr__(self):
        return f"SR(qgt={getattr(self.qgt_constructor, '__name__', self.qgt_constructor)}, diag_shift={self.diag_shift}, solver_restart={self.solver_restart})"


def identity_preconditioner(vstate, gradient, step=None):
    return gradient
