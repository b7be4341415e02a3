"""Forces / gradient of <O> for the RBM (TEST INFRASTRUCTURE).  Pinned by tests/golden/sampler_vectors.npz:
``forces_expect_hermitian`` executed from the reference's source (tests/golden/make_golden_sampler.py; its reverse-mode vjp
replaced by Richardson-extrapolated central differences of the reference's forward pass), tests/test_golden_sampler.py;
the closed-form log-derivatives are also checked against finite differences of ``oracle.rbm.logpsi`` in tests/test_oracle.py.

``forces_expect_hermitian`` (netket/vqs/mc/mc_state/expect_forces.py:69-112):

    O_loc -= mean(O_loc)
    forces = vjp(w -> logpsi(w, sigma))(conj(O_loc) / n_samples)         # sum_s dlogpsi(sigma_s)/dw * dE_s / n_samples

``force_to_grad`` (netket/vqs/mc/common.py:103-118): gradient = 2 * Re(forces) for real parameters.
RBM (netket/models/rbm.py:57-81): dlogpsi/dW_ij = sigma_i tanh(theta_j), /db_j = tanh(theta_j), /da_i = sigma_i.
"""

import numpy as np

from .rbm import theta


def log_derivatives(sigma, W, b, a):
    """Per-sample log-derivatives: (O_W[B,N,M], O_b[B,M], O_a[B,N])."""
    sigma = np.asarray(sigma).astype(W.dtype)
    t = np.tanh(theta(sigma, W, b))
    return sigma[:, :, None] * t[:, None, :], t, sigma


def forces(samples, eloc, W, b, a, mean=None, n_total=None):
    """samples[..., N], eloc[...] -> dict(W, b, a) of forces (b / a entries are None when the RBM has no such bias)."""
    sig = np.asarray(samples).reshape(-1, W.shape[0]).astype(W.dtype)
    e = np.asarray(eloc, dtype=np.float64).reshape(-1)
    mean = e.mean() if mean is None else mean
    n = e.size if n_total is None else n_total
    w = ((e - mean) / n).astype(W.dtype)
    t = np.tanh(theta(sig, W, b))
    tw = t * w[:, None]
    return dict(W=sig.T @ tw, b=None if b is None else tw.sum(axis=0), a=None if a is None else sig.T @ w)


def grad(samples, eloc, W, b, a, **kw):
    f = forces(samples, eloc, W, b, a, **kw)
    return {k: (None if v is None else 2.0 * v) for k, v in f.items()}


def expect_and_forces(samples, conn_fn, W, b, a):
    """``expect_and_forces`` (expect_forces.py:39-112) on samples[n_chains, chain_length, N]: (mean of E_loc, forces in the
    parameter layout of the reference: kernel / bias / visible_bias)."""
    from .estimators import local_estimators

    eloc = local_estimators(samples, conn_fn, W, b, a)
    f = forces(samples, eloc, W, b, a)
    return float(eloc.mean()), {"kernel": f["W"], "bias": f["b"], "visible_bias": f["a"]}
