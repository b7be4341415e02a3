"""MC statistics, block variant (TEST INFRASTRUCTURE).

Follows netket/stats/mc_stats_old.py:28-196 (the default ``statistics``; the FFT variant
of netket/stats/mc_stats.py:303-331 is opt-in behind a flag and out of scope):
variances are ddof=0; ``l_block = max(1, L // 32)``; blocks are consecutive per-chain means
over the first floor(L/l)*l samples; tau = ((ts/n) * var_of_means / variance - 1)/2;
error_of_mean from batches if ``tau_batch < 6 L and n_batches >= 32`` else from blocks if
``tau_block < 6 l_block and n_blocks >= 32`` else NaN; split-R_hat with the total variance
as W (:165-190); NaN R_hat for a single chain.
"""

import numpy as np


def statistics(data, batch_size=32):
    data = np.atleast_1d(np.asarray(data))
    if data.ndim == 1:
        data = data.reshape(1, -1)
    if data.ndim > 2:
        raise NotImplementedError("Statistics are implemented only for ndim<=2")
    n_chains, L = data.shape
    mean = data.mean()
    variance = data.var()
    ts = data.size
    b_means = data.mean(axis=1)
    batch_var, n_batches = b_means.var(), b_means.size
    l_block = max(1, L // batch_size)
    n_b = int(np.floor(L / float(l_block)))
    blocks = data[:, : n_b * l_block].reshape(-1, l_block).mean(axis=1)
    n_blocks = blocks.size
    block_var = blocks.var() if n_blocks > 0 else np.nan
    with np.errstate(divide="ignore", invalid="ignore"):
        tau_batch = ((ts / n_batches) * batch_var / variance - 1) * 0.5
        tau_block = ((ts / n_blocks) * block_var / variance - 1) * 0.5 if n_blocks > 0 else np.nan
    batch_good = bool(tau_batch < 6 * L) and n_batches >= batch_size
    block_good = bool(tau_block < 6 * l_block) and n_blocks >= batch_size
    if batch_good:
        error_of_mean = np.sqrt(batch_var / n_batches)
        tau_corr = max(tau_batch, 0.0)
    elif block_good:
        error_of_mean = np.sqrt(block_var / n_blocks)
        tau_corr = max(tau_block, 0.0)
    else:
        error_of_mean = np.nan
        tau_corr = np.nan
    if n_batches > 1:
        N = L
        half = data if N % 2 == 0 else data[:, :-1]
        hv = half.reshape(2 * n_chains, N // 2).mean(axis=1).var()
        with np.errstate(divide="ignore", invalid="ignore"):
            R_hat = np.sqrt((N - 1) / N + hv / variance)
    else:
        R_hat = np.nan
    return dict(mean=mean, error_of_mean=error_of_mean, variance=variance, tau_corr=tau_corr, R_hat=R_hat)


# ------------------------------------------------------------------------------------------------------------------
# FFT variant (opt-in in the reference: flag NETKET_EXPERIMENTAL_FFT_AUTOCORRELATION; netket/stats/mc_stats.py:303-331,
# netket/stats/_autocorr.py:32-86).  Pinned by tests/golden/online_stats_vectors.npz (keys fft_*).  No product code yet.
# ------------------------------------------------------------------------------------------------------------------
def autocorr_1d(x):
    """Normalised autocorrelation function of one series; the direct O(L^2) sum the zero-padded FFT evaluates."""
    x = np.asarray(x, dtype=np.float64)
    d = x - x.mean()
    L = d.size
    acf = np.array([np.dot(d[: L - k], d[k:]) for k in range(L)])
    return acf / acf[0]


def integrated_time(x, c=5):
    """Sokal's automatic window: tau(M) = 2 sum_{k<=M} rho_k - 1 at the first M with M >= c tau(M).  As the reference codes it
    (`auto_window`, _autocorr.py:62-64): `argmin(M < c tau)`, which is 0 when no M qualifies, and the last M when every M does."""
    taus = 2.0 * np.cumsum(autocorr_1d(x)) - 1.0
    m = np.arange(taus.size) < c * taus
    window = int(np.argmin(m)) if m.any() else taus.size - 1
    return taus[window]


def statistics_fft(data):
    data = np.atleast_1d(np.asarray(data, dtype=np.float64))
    if data.ndim == 1:
        data = data.reshape(1, -1)
    n_chains, L = data.shape
    mean, variance = data.mean(), data.var()
    taus = np.array([integrated_time(row) for row in data])
    if n_chains > 1:
        error_of_mean = np.sqrt(data.mean(axis=1).var() / n_chains)
        half = data if L % 2 == 0 else data[:, :-1]
        with np.errstate(divide="ignore", invalid="ignore"):
            R_hat = np.sqrt((L - 1) / L + half.reshape(2 * n_chains, L // 2).mean(axis=1).var() / variance)
    else:
        l_block = max(1, L // 32)
        n_b = L // l_block
        blocks = data[:, : n_b * l_block].reshape(-1, l_block).mean(axis=1)
        error_of_mean = np.sqrt(blocks.var() / blocks.size)
        R_hat = np.nan
    return dict(mean=mean, error_of_mean=error_of_mean, variance=variance, tau_corr=taus.mean(), R_hat=R_hat, tau_corr_max=taus.max())
