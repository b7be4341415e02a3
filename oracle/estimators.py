"""Local estimators (TEST INFRASTRUCTURE).

``local_value_kernel_jax`` (netket/vqs/mc/kernels.py:62-71):

    sigma', mel = O.get_conn_padded(sigma)
    E_loc = sum_k mel_k * exp(logpsi(sigma'_k) - logpsi(sigma))

evaluated exactly as the reference does: connected configurations are *materialised* and
run through the full forward pass.  ``local_estimators`` collapses (n_chains, chain_len, N)
-> (B, N) and reshapes the result back (netket/vqs/mc/mc_state/local_estimators.py:34-38).
"""

import numpy as np

from .rbm import logpsi


def local_value_kernel(sigma, conn_fn, W, b, a):
    """sigma[B,N] -> E_loc[B] in promote(mel dtype, W dtype)."""
    sigma = np.asarray(sigma)
    xp, mels = conn_fn(sigma)[:2]
    lp = logpsi(sigma, W, b, a)
    lpp = logpsi(xp.reshape(-1, xp.shape[-1]), W, b, a).reshape(xp.shape[:-1])
    return np.sum(mels * np.exp(lpp - lp[..., None]), axis=-1)


def local_estimators(samples, conn_fn, W, b, a, chunk=4096):
    """samples[n_chains, chain_len, N] -> E_loc[n_chains, chain_len]."""
    samples = np.asarray(samples)
    shp = samples.shape[:-1]
    flat = samples.reshape(-1, samples.shape[-1])
    out = []
    for s in range(0, flat.shape[0], chunk):
        out.append(local_value_kernel(flat[s:s + chunk], conn_fn, W, b, a))
    return np.concatenate(out).reshape(shp)
