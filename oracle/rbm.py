"""RBM log-amplitude (TEST INFRASTRUCTURE).

Follows netket/models/rbm.py:57-81 (Dense -> log_cosh -> sum, + visible bias) and
netket/nn/activation.py:78-84 (log_cosh(x) = |x| + log1p(exp(-2|x|)) - log 2).
Parameter layout is Flax's: kernel W (N, M) "in x out", hidden bias b (M,), visible
bias a (N,); default init normal(0.01) (rbm.py:29).
"""

import numpy as np


def log_cosh(x):
    """activation.py:78-84, real input."""
    x = np.asarray(x)
    sgn = -2 * np.signbit(x).astype(x.dtype) + 1
    x = x * sgn
    return x + np.log1p(np.exp(-2.0 * x)) - np.log(np.asarray(2.0, dtype=x.dtype))


def init_params(N, alpha, seed=1234, std=0.01, dtype=np.float64, use_hidden_bias=True, use_visible_bias=True):
    """Synthetic random-init parameters (BASELINE.md 'Configs and synthetic inputs').

    Drawn in fp64 from numpy.random.default_rng(seed) in the order W, b, a and then cast.
    """
    M = int(alpha * N)
    g = np.random.default_rng(seed)
    W = g.normal(0.0, std, size=(N, M))
    b = g.normal(0.0, std, size=(M,))
    a = g.normal(0.0, std, size=(N,))
    W = W.astype(dtype)
    b = b.astype(dtype) if use_hidden_bias else None
    a = a.astype(dtype) if use_visible_bias else None
    return W, b, a


def theta(sigma, W, b=None):
    """Hidden pre-activations theta = sigma W + b, computed in W.dtype (flax Dense promotes int8)."""
    th = np.asarray(sigma).astype(W.dtype) @ W
    if b is not None:
        th = th + b
    return th


def logpsi(sigma, W, b=None, a=None):
    """log psi(sigma) for a batch sigma[..., N] -> [...]."""
    sigma = np.asarray(sigma)
    x = log_cosh(theta(sigma, W, b)).sum(axis=-1)
    if a is not None:
        x = x + sigma.astype(W.dtype) @ a
    return x


def to_array(W, b, a, states, normalize=True):
    """Full wave-function on a list of basis states (netket/nn/utils.py:42-98)."""
    lp = logpsi(states, W.astype(np.float64), None if b is None else b.astype(np.float64),
                None if a is None else a.astype(np.float64))
    psi = np.exp(lp - lp.max())
    if normalize:
        psi = psi / np.linalg.norm(psi)
    return psi
