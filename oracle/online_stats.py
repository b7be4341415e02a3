"""Streaming MC statistics (TEST INFRASTRUCTURE): per-chain Welford state + online autocovariance at lags 0..max_lag.

CPU restatement of the reference's accumulator, pinned by tests/golden/online_stats_vectors.npz (the reference's own
source executed by tests/golden/make_golden_online.py):

  * update                      netket/_src/stats/online_stats/kernels.py:25-190 (`_acf_core`, `_update_arrays`)
  * derived quantities          netket/_src/stats/online_stats/accumulator.py:226-447
  * expand_max_lag / thin_by_2  netket/_src/stats/online_stats/operations.py:132-261
  * window_saturated / reliable netket/_src/vqs/check_mc_convergence.py:243-272

The autocovariance update is written on the concatenated series z = [buffer | batch] instead of the reference's two
masked windows: the pair (t, t-k) counts when t lies in the batch and t-k is a stored sample; that is the union of the
reference's "within-batch" and "cross-batch" pairs.
"""

import math

import numpy as np


class OnlineStats:
    def __init__(self, n_chains, dtype=np.float64, decay=None, max_lag=64):
        L = int(max_lag)
        n_acf = L + 1 if L > 0 else 0
        self.max_lag, self.decay = L, decay
        self.count = np.zeros(n_chains)
        self.mean_c = np.zeros(n_chains, dtype=dtype)
        self.M2 = np.zeros(n_chains)
        self.cross = np.zeros((n_chains, n_acf))
        self.m1 = np.zeros((n_chains, n_acf))
        self.m2 = np.zeros((n_chains, n_acf))
        self.pairs = np.zeros((n_chains, n_acf))
        self.buf = np.zeros((n_chains, L))
        self.buf_len = 0
        self.n_samples = 0

    @property
    def n_chains(self):
        return self.count.shape[0]

    def copy(self):
        new = OnlineStats.__new__(OnlineStats)
        for k, v in self.__dict__.items():
            setattr(new, k, v.copy() if isinstance(v, np.ndarray) else v)
        return new

    # ------------------------------------------------------------------ update (kernels.py:116-190)
    def update(self, data):
        data = np.asarray(data)
        if data.ndim == 1:
            data = data[None, :]
        if data.ndim != 2:
            raise ValueError(f"data must be 1D or 2D, got {data.ndim}D")
        C, n = data.shape
        if C != self.n_chains:
            raise ValueError(f"Number of chains changed: expected {self.n_chains}, got {C}")
        new = self.copy()
        L = self.max_lag
        for c in range(C):
            x = data[c]
            bm = x.mean()  # in the data's dtype, like the reference
            bM2 = float(np.sum(np.abs(x - bm) ** 2))
            cnt, M2 = float(self.count[c]), float(self.M2[c])
            if self.decay is not None:
                cnt *= self.decay
                M2 *= self.decay
            tot = cnt + n
            safe = tot if tot > 0 else 1.0
            delta = bm - self.mean_c[c]
            new.mean_c[c] = self.mean_c[c] + delta * (n / safe)
            new.M2[c] = M2 + bM2 + abs(delta) ** 2 * (cnt * n / safe)
            new.count[c] = tot
            if L == 0:
                continue
            z = np.concatenate([self.buf[c], x.astype(np.float64)])
            first = L - self.buf_len  # first stored sample of z
            for k in range(L + 1):
                s_cross = s_lag = s_cur = 0.0
                npair = 0
                for t in range(L, L + n):
                    if t - k >= first:
                        s_cross += z[t] * z[t - k]
                        s_lag += z[t - k]
                        s_cur += z[t]
                        npair += 1
                d = 1.0 if self.decay is None else self.decay
                new.cross[c, k] = self.cross[c, k] * d + s_cross
                new.m1[c, k] = self.m1[c, k] * d + s_lag
                new.m2[c, k] = self.m2[c, k] * d + s_cur
                new.pairs[c, k] = self.pairs[c, k] * d + npair
            new.buf[c] = z[n:]  # the last max_lag samples, right-aligned (kernels.py:103-113)
        new.buf_len = min(self.buf_len + n, L)
        new.n_samples = self.n_samples + C * n
        return new

    # ------------------------------------------------------------------ derived quantities (accumulator.py:240-447)
    @property
    def mean(self):
        total = self.count.sum()
        if total == 0:
            return math.nan
        return float((self.count * self.mean_c).sum() / total)

    @property
    def variance(self):
        total = self.count.sum()
        if total == 0:
            return math.nan
        g = (self.count * self.mean_c).sum() / total
        return float((self.M2.sum() + (self.count * np.abs(self.mean_c - g) ** 2).sum()) / total)

    @property
    def acf(self):
        if self.max_lag == 0:
            return None
        n = np.maximum(self.pairs, 1.0)
        cov = (self.cross / n - (self.m1 / n) * (self.m2 / n)).mean(axis=0)
        if cov[0] <= 0:
            return None
        return cov / cov[0]

    @property
    def tau_corr_acf(self):
        """Geyer initial positive + initial monotone sequence on the pair sums rho[2t] + rho[2t+1]."""
        rho = self.acf
        if rho is None:
            return math.nan
        m = len(rho) // 2
        if m == 0:
            return math.nan
        total, running_min = 0.0, math.inf
        for t in range(m):
            p = rho[2 * t] + rho[2 * t + 1]
            if p <= 0:
                if t == 0:
                    return 1.0
                break
            running_min = min(running_min, p)
            total += running_min
        return max(2.0 * total - 1.0, 1.0)

    @property
    def tau_corr_batch(self):
        if self.n_chains < 2:
            return math.nan
        v = self.variance
        if math.isnan(v) or v <= 0:
            return math.nan
        n_eff = self.n_samples / self.n_chains
        return max((n_eff * np.var(self.mean_c.astype(np.float64)) / v - 1) * 0.5, 0.0)

    @property
    def tau_corr(self):
        tau = self.tau_corr_acf
        return self.tau_corr_batch if math.isnan(tau) else tau

    @property
    def R_hat(self):
        if self.n_chains < 2:
            return math.nan
        W = (self.M2 / np.maximum(self.count, 1.0)).mean()
        if W <= 0:
            return math.nan
        N = self.count.mean()
        return math.sqrt((N - 1) / N + np.var(self.mean_c.astype(np.float64)) / W)

    @property
    def error_of_mean(self):
        if self.count.sum() == 0:
            return math.nan
        if self.n_chains > 1:
            return math.sqrt(np.var(self.mean_c.astype(np.float64)) / self.n_chains)
        tau = self.tau_corr_acf
        if not math.isnan(tau):
            return math.sqrt(self.variance * tau / self.n_samples)
        return math.nan  # the batch estimate needs two chains

    def get_stats(self):
        if self.count.sum() == 0:
            return dict(mean=math.nan, error_of_mean=math.nan, variance=math.nan, tau_corr=math.nan, R_hat=math.nan)
        return dict(mean=self.mean, error_of_mean=self.error_of_mean, variance=self.variance, tau_corr=self.tau_corr, R_hat=self.R_hat)


def online_statistics(data, old_estimator=None, *, decay=None, max_lag=64):
    """operations.py:55-129."""
    data = np.asarray(data)
    if old_estimator is None:
        old_estimator = OnlineStats(1 if data.ndim == 1 else data.shape[0], dtype=data.dtype, decay=decay, max_lag=max_lag)
    return old_estimator.update(data)


def expand_max_lag(est, new_max_lag):
    """operations.py:132-192: new lags start empty; the buffer grows on the left (stays right-aligned)."""
    new_max_lag = int(new_max_lag)
    if new_max_lag <= est.max_lag:
        raise ValueError(f"new_max_lag={new_max_lag} must be > current max_lag={est.max_lag}")
    new = est.copy()
    old_len = est.max_lag + 1 if est.max_lag > 0 else 0
    extra = new_max_lag + 1 - old_len
    for f in ("cross", "m1", "m2", "pairs"):
        setattr(new, f, np.pad(getattr(est, f), ((0, 0), (0, extra))))
    new.buf = np.pad(est.buf, ((0, 0), (new_max_lag - est.max_lag, 0)))
    new.max_lag = new_max_lag
    return new


def thin_acf_by_2(est):
    """operations.py:195-261: even lags become lags 0..max_lag//2, every other buffered sample is kept."""
    if est.max_lag < 2:
        raise ValueError(f"max_lag={est.max_lag} must be >= 2 to thin by 2")
    new = est.copy()
    L2 = est.max_lag // 2
    for f in ("cross", "m1", "m2", "pairs"):
        setattr(new, f, getattr(est, f)[:, 0:2 * L2 + 1:2].copy())
    start = est.max_lag - 2 * L2
    new.buf = est.buf[:, start::2].copy()
    new.buf_len = est.buf_len // 2
    new.max_lag = L2
    return new


def acf_window_saturated(est):
    """check_mc_convergence.py:243-255: every pair sum inside the window is positive."""
    rho = est.acf
    if rho is None or len(rho) // 2 == 0:
        return False
    m = len(rho) // 2
    return all(rho[2 * t] + rho[2 * t + 1] > 0 for t in range(m))


def tau_corr_reliable(est):
    """check_mc_convergence.py:258-272."""
    if acf_window_saturated(est):
        return False
    tau = est.tau_corr_acf
    if math.isnan(tau) or tau <= 0:
        return False
    return (est.n_samples / est.n_chains) / tau >= 50
