"""Quantum geometric tensor of the RBM and the SR solve (TEST INFRASTRUCTURE; parity unpinned against a golden vector: the
reference builds S v from jax.linearize / linear_transpose, which cannot run here - the log-derivatives it rests on are
checked against finite differences in tests/test_oracle.py, and S below is the definition the reference's own tests compare
the matrix-free product with, test/optimizer/test_qgt_itersolve.py).

``S = <dO^H dO>`` with ``dO = O - <O>`` over the samples (netket/optimizer/qgt/qgt_onthefly_logic.py:33-43:
``S v = O^H ((O v - mean(O v)) / n) + diag_shift v``); ``SR``: ``(S + diag_shift) dp = grad`` (netket/optimizer/sr.py:56-215).
Parameter order: ``[W.ravel() | b | a]``.
"""

import numpy as np

from .forces import log_derivatives


def jacobian(samples, W, b, a):
    """O[s, k] = d log psi(sigma_s) / d p_k, shape (n_samples, n_parameters)."""
    sig = np.asarray(samples).reshape(-1, W.shape[0])
    OW, Ob, Oa = log_derivatives(sig, W.astype(np.float64), None if b is None else b.astype(np.float64), a)
    parts = [OW.reshape(sig.shape[0], -1)]
    if b is not None:
        parts.append(Ob)
    if a is not None:
        parts.append(Oa.astype(np.float64))
    return np.concatenate(parts, axis=1)


def qgt_dense(samples, W, b, a, diag_shift=0.0):
    O = jacobian(samples, W, b, a)
    dO = O - O.mean(axis=0, keepdims=True)
    return dO.T @ dO / O.shape[0] + diag_shift * np.eye(O.shape[1])


def mat_vec(samples, W, b, a, v, diag_shift=0.0):
    """The matrix-free form, step by step as the reference does it."""
    O = jacobian(samples, W, b, a)
    w = O @ v / O.shape[0]
    w = w - w.mean()
    return O.T @ w + diag_shift * v


def sr_solve(samples, W, b, a, grad, diag_shift=0.01):
    return np.linalg.solve(qgt_dense(samples, W, b, a, diag_shift), grad)
