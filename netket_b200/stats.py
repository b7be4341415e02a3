"""``nk.stats.statistics`` / ``Stats`` (netket/stats/mc_stats_old.py:52-196, netket/stats/mc_stats.py:84-179).

The per-device work is ``nk_stats_partial`` (8 doubles); across GPUs only those scalars travel
(``torch.distributed.all_reduce`` over NCCL/NVLink), then ``nk_stats_finalize`` does the scalar arithmetic.
"""

import ctypes as C
import math

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .utils import world


def _fmt(mean, err, var):
    if not (err == err) or err <= 0 or math.isinf(err):
        return f"{mean:.3e} ± {err:.1e} [σ²={var:.1e}"
    digits = max(0, -int(math.floor(math.log10(err))) + 1)
    return f"{mean:.{digits}f} ± {err:.{digits}f} [σ²={var:.{max(digits - 1, 1)}f}"


class Stats:
    """Dict-compatible result of ``statistics`` (mc_stats.py:84-179)."""

    __slots__ = ("mean", "error_of_mean", "variance", "tau_corr", "R_hat", "tau_corr_max")

    def __init__(self, mean=math.nan, error_of_mean=math.nan, variance=math.nan, tau_corr=math.nan, R_hat=math.nan,
                 tau_corr_max=math.nan):
        self.mean, self.error_of_mean, self.variance = mean, error_of_mean, variance
        self.tau_corr, self.R_hat, self.tau_corr_max = tau_corr, R_hat, tau_corr_max

    def to_dict(self):
        return {"Mean": self.mean, "Variance": self.variance, "Sigma": self.error_of_mean, "R_hat": self.R_hat,
                "TauCorr": self.tau_corr}

    def to_compound(self):
        return "Mean", self.to_dict()

    _ALIASES = {"Mean": "mean", "Variance": "variance", "Sigma": "error_of_mean", "R": "R_hat", "TauCorr": "tau_corr",
                "TauCorrMax": "tau_corr_max"}

    def __getattr__(self, name):
        alias = Stats._ALIASES.get(name)
        if alias is None:
            raise AttributeError(f"'Stats' object object has no attribute '{name}'")
        return getattr(self, alias)

    def __getitem__(self, name):
        return getattr(self, name)

    @property
    def shape(self):
        return ()

    def __repr__(self):
        ext = f", R̂={self.R_hat:.3f}" if not math.isnan(self.R_hat) else ""
        return _fmt(self.mean, self.error_of_mean, self.variance) + f"{ext}]"


def _allreduce(t):
    """Sum a small tensor over all ranks (NCCL on GPU; gloo in the CPU tests)."""
    _, ws = world()
    if ws > 1:
        if dist.get_backend() == "gloo" and t.is_cuda:
            c = t.cpu()
            dist.all_reduce(c)
            t.copy_(c)
        else:
            dist.all_reduce(t)
    return t


def finalize(sums, mean, n_chains_total, L):
    """Host arithmetic of _statistics (mc_stats_old.py:87-196) on globally reduced partial sums."""
    p = (C.c_double * _lib.NK_STATS_NPARTIAL)(*[float(v) for v in sums])
    out = (C.c_double * 5)()
    _lib.check(_lib.lib().nk_stats_finalize(p, float(mean), int(n_chains_total), int(L), out))
    return Stats(out[0], out[1], out[2], out[3], out[4])


def _allreduce_max_bits(t):
    """MAX over ranks of order-preserving integer encodings (nk_stats_tau's out[1]) viewed as int64."""
    _, ws = world()
    if ws > 1:
        v = t.view(torch.int64)
        # the encoding is an unsigned order; flipping the top bit makes it the signed order int64 MAX uses
        v ^= -0x8000000000000000
        if dist.get_backend() == "gloo" and t.is_cuda:
            c = v.cpu()
            dist.all_reduce(c, op=dist.ReduceOp.MAX)
            v.copy_(c)
        else:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        v ^= -0x8000000000000000
    return t


def statistics_fft(data):
    """The opt-in FFT variant (netket/stats/mc_stats.py:303-331): mean / variance / R_hat as in the block version, error of the
    mean from the chain means (from blocks for a single chain) whatever the quality flags, ``tau_corr`` = chain average of the
    integrated autocorrelation time with Sokal's window (netket/stats/_autocorr.py:40-86), plus ``tau_corr_max``."""
    data = _as_stats_input(data)
    n_chains, L = data.shape
    dev = data.device
    code = _lib.dtype_code(data.dtype)
    part = torch.zeros(_lib.NK_STATS_NPARTIAL, dtype=torch.float64, device=dev)
    tau = torch.zeros(4, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        _lib.check(_lib.lib().nk_stats_partial(st, _lib.ptr(data), code, n_chains, L, 0, 0.0, _lib.ptr(part)))
        head = torch.cat([part[:1], torch.tensor([float(n_chains)], dtype=torch.float64, device=dev)])
        _allreduce(head)
        total, n_chains_total = head.tolist()
        n_chains_total = int(round(n_chains_total))
        mean = total / (n_chains_total * L)
        _lib.check(_lib.lib().nk_stats_partial(st, _lib.ptr(data), code, n_chains, L, 1, mean, _lib.ptr(part)))
        _lib.check(_lib.lib().nk_stats_tau(st, _lib.ptr(data), code, n_chains, L, 5.0, _lib.ptr(tau)))
        _allreduce(part)
        sums = torch.stack([tau[0], tau[2]])
        _allreduce(sums)
        _allreduce_max_bits(tau[1:2])
    p = part.tolist()
    tau_sum, n_nan = sums.tolist()
    tau_max = float(_lib.lib().nk_stats_tau_max_decode(float(tau[1].item())))
    ts = float(n_chains_total * L)
    dm = p[7] / ts
    variance = p[0] / ts - dm * dm
    nan = math.nan
    if n_chains_total > 1:
        nb = float(n_chains_total)
        batch_var = p[2] / nb - (p[1] / nb) ** 2
        err = math.sqrt(max(batch_var, 0.0) / nb)
        half = L // 2
        hv = (p[6] / (2 * nb) - (p[5] / (2 * nb)) ** 2) if half > 0 else nan
        rhat = math.sqrt((L - 1.0) / L + hv / variance) if variance > 0 else nan
    else:
        l_block = max(1, L // 32)
        n_blocks = n_chains_total * (L // l_block)
        block_var = p[4] / n_blocks - (p[3] / n_blocks) ** 2
        err = math.sqrt(max(block_var, 0.0) / n_blocks)
        rhat = nan
    if n_nan > 0:
        tau_avg, tau_max = nan, nan
    else:
        tau_avg = tau_sum / n_chains_total
    return Stats(mean + dm, err, variance, tau_avg, rhat, tau_max)


def _as_stats_input(data):
    if not isinstance(data, torch.Tensor):
        from .utils import default_device

        data = torch.from_numpy(np.ascontiguousarray(np.asarray(data))).to(default_device())
    _lib.require_cuda(data, "data")
    if data.ndim == 0:
        data = data.reshape(1, 1)
    elif data.ndim == 1:
        data = data.reshape(1, -1)
    elif data.ndim > 2:
        raise NotImplementedError("Statistics are implemented only for ndim<=2")
    if data.dtype not in (torch.float32, torch.float64):
        raise TypeError("statistics: float32 / float64 data only (real-parameter RBM => real local energies)")
    return data.contiguous()


def statistics(data):
    """Statistics of ``data[n_chains, L]`` (or a 1-D time series) held on this rank's GPU.  Under
    torch.distributed the chains of all ranks are combined (chain axis sharded, as the reference does with its
    mesh axis "S": netket/stats/mc_stats_old.py:96-107).  With the flag ``netket_experimental_fft_autocorrelation``
    (netket_b200.config) the FFT variant runs instead, as in netket/stats/mc_stats.py:296-301."""
    from .config import config

    if config.netket_experimental_fft_autocorrelation:
        return statistics_fft(data)
    if not isinstance(data, torch.Tensor):
        from .utils import default_device

        data = torch.from_numpy(np.ascontiguousarray(np.asarray(data))).to(default_device())
    _lib.require_cuda(data, "data")
    if data.ndim == 0:
        data = data.reshape(1, 1)
    elif data.ndim == 1:
        data = data.reshape(1, -1)
    elif data.ndim > 2:
        raise NotImplementedError("Statistics are implemented only for ndim<=2")
    if data.dtype not in (torch.float32, torch.float64):
        raise TypeError("statistics: float32 / float64 data only (real-parameter RBM => real local energies)")
    data = data.contiguous()
    n_chains, L = data.shape
    dev = data.device
    code = _lib.dtype_code(data.dtype)
    part = torch.zeros(_lib.NK_STATS_NPARTIAL, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        _lib.check(_lib.lib().nk_stats_partial(st, _lib.ptr(data), code, n_chains, L, 0, 0.0, _lib.ptr(part)))
        cnt = torch.tensor([float(n_chains)], dtype=torch.float64, device=dev)
        head = torch.cat([part[:1], cnt])
        _allreduce(head)
        total, n_chains_total = head.tolist()
        n_chains_total = int(round(n_chains_total))
        mean = total / (n_chains_total * L)
        _lib.check(_lib.lib().nk_stats_partial(st, _lib.ptr(data), code, n_chains, L, 1, mean, _lib.ptr(part)))
        _allreduce(part)
    return finalize(part.tolist(), mean, n_chains_total, L)


# ------------------------------------------------------------------------------------------------------------------
# Streaming statistics (netket/_src/stats/online_stats/*): the accumulator behind thermalise_mcmc,
# check_mc_convergence and expect_to_precision.
# ------------------------------------------------------------------------------------------------------------------
_ONLINE_FIELDS = ("_chain_count", "_chain_mean", "_chain_M2", "_cross_sum", "_m1_sum", "_m2_sum", "_pair_count", "_chain_buf")


class OnlineStats:
    """``OnlineStats`` (accumulator.py:31-447): per-chain Welford state and the autocovariance at lags ``0..max_lag``,
    updated batch by batch on the GPU (nk_online_stats_update, one warp per chain).  The fields carry the reference's
    names and shapes; they are float64 device tensors (the reference keeps ``_chain_mean`` in the data's dtype).  Under
    torch.distributed every rank holds its own chains; the derived quantities all-reduce ``4 + max_lag + 4`` doubles.
    """

    def __init__(self, n_chains, dtype=None, *, decay=None, max_lag=64, device=None):
        from .utils import default_device

        max_lag = int(max_lag)
        if not 0 <= max_lag <= _lib.NK_ONLINE_MAX_LAG:
            raise ValueError(f"max_lag must lie in [0, {_lib.NK_ONLINE_MAX_LAG}]")
        if decay is not None and not 0.0 < float(decay) <= 1.0:
            raise ValueError("decay must lie in (0, 1]")
        dev = torch.device(device) if device is not None else default_device()
        acf_len = max_lag + 1 if max_lag > 0 else 0
        self.max_lag = max_lag
        self._decay = decay
        z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)  # noqa: E731
        self._chain_count, self._chain_mean, self._chain_M2 = z(n_chains), z(n_chains), z(n_chains)
        self._cross_sum, self._m1_sum, self._m2_sum, self._pair_count = (z(n_chains, acf_len) for _ in range(4))
        self._chain_buf = z(n_chains, max_lag)
        self._buf_len = 0
        self._n_samples_total = 0  # of this rank
        self._summary = None

    # ------------------------------------------------------------------ construction / update
    @classmethod
    def from_data(cls, data, *, decay=None, max_lag=64):
        data = _as_device_2d(data)
        return cls(data.shape[0], data.dtype, decay=decay, max_lag=max_lag, device=data.device).update(data)

    def replace(self, **kw):
        new = OnlineStats.__new__(OnlineStats)
        new.__dict__.update(self.__dict__)
        new.__dict__.update(kw)
        new._summary = None
        return new

    def _c_struct(self):
        p = lambda t: _lib.ptr(t) if t.numel() else None  # noqa: E731
        return _lib.nk_online_stats_t(p(self._chain_count), p(self._chain_mean), p(self._chain_M2), p(self._cross_sum), p(self._m1_sum),
                                      p(self._m2_sum), p(self._pair_count), p(self._chain_buf), self.n_chains, self.max_lag, self._buf_len)

    def update(self, data, *, inplace=False):
        """Merge a batch ``(n_chains, n)`` (or ``(n,)`` for one chain) and return the updated accumulator
        (accumulator.py:166-221).  ``inplace=True`` reuses this object's buffers (the streaming loops do)."""
        data = _as_device_2d(data)
        if data.ndim != 2:
            raise ValueError(f"data must be 1D or 2D, got {data.ndim}D")
        n_chains_new, n = data.shape
        if n_chains_new != self.n_chains:
            raise ValueError(f"Number of chains changed: expected {self.n_chains}, got {n_chains_new}")
        dev = self._chain_count.device
        if data.device != dev:
            raise ValueError("data lives on another device than the accumulator")
        out = self if inplace else self.replace(**{f: torch.empty_like(getattr(self, f)) for f in _ONLINE_FIELDS})
        with torch.cuda.device(dev):
            a, b = self._c_struct(), out._c_struct()
            _lib.check(_lib.lib().nk_online_stats_update(_lib.stream_ptr(dev), C.byref(a), C.byref(b), _lib.ptr(data),
                                                         _lib.dtype_code(data.dtype), n, 1.0 if self._decay is None else float(self._decay)))
        out._buf_len = min(self._buf_len + n, self.max_lag)
        out._n_samples_total = self._n_samples_total + n_chains_new * n
        out._summary = None
        return out

    # ------------------------------------------------------------------ derived quantities
    @property
    def n_chains(self):
        return self._chain_mean.shape[0]

    @property
    def chain_means(self):
        return self._chain_mean

    @property
    def decay(self):
        return 1.0 if self._decay is None else self._decay

    def _summarise(self):
        if self._summary is not None:
            return self._summary
        dev = self._chain_count.device
        n_lag = self.max_lag + 1 if self.max_lag > 0 else 0
        p0 = torch.zeros(5, dtype=torch.float64, device=dev)
        p1 = torch.zeros(_lib.NK_ONLINE_NSUM + n_lag, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            st, s = _lib.stream_ptr(dev), self._c_struct()
            _lib.check(_lib.lib().nk_online_stats_summary(st, C.byref(s), 0, 0.0, 0.0, _lib.ptr(p0)))
            p0[3] = float(self.n_chains)
            p0[4] = float(self._n_samples_total)
            _allreduce(p0)
            h0 = p0.tolist()
            n_chains, n_samples = int(round(h0[3])), int(round(h0[4]))
            gmean = h0[1] / h0[0] if h0[0] > 0 else 0.0
            mbar = h0[2] / n_chains if n_chains > 0 else 0.0
            _lib.check(_lib.lib().nk_online_stats_summary(st, C.byref(s), 1, gmean, mbar, _lib.ptr(p1)))
            _allreduce(p1)
            h1 = p1.tolist()
        self._summary = online_finalize(h0[:3], h1, n_chains, n_samples, self.max_lag)
        return self._summary

    mean = property(lambda self: self._summarise()["out"][0], doc="count-weighted mean of the chain means (accumulator.py:240-250)")
    variance = property(lambda self: self._summarise()["out"][2], doc="accumulator.py:252-263")
    tau_corr = property(lambda self: self._summarise()["out"][3], doc="ACF-based if available, else batch (accumulator.py:265-271)")
    R_hat = property(lambda self: self._summarise()["out"][4], doc="accumulator.py:379-395")
    tau_corr_batch = property(lambda self: self._summarise()["out"][5], doc="accumulator.py:273-309")
    tau_corr_acf = property(lambda self: self._summarise()["out"][6], doc="Geyer IPS + IMS (accumulator.py:311-351)")
    error_of_mean = property(lambda self: self._summarise()["out"][1], doc="accumulator.py:430-447")

    @property
    def acf(self):
        """Normalised autocorrelation function averaged over chains, ``(max_lag + 1,)`` NumPy array, or ``None``."""
        return self._summarise()["acf"]

    @property
    def n_samples(self):
        """Samples accumulated over all ranks (never decayed)."""
        return self._summarise()["n_samples"]

    def get_stats(self):
        s = self._summarise()
        if s["empty"]:
            return Stats()
        o = s["out"]
        return Stats(mean=o[0], error_of_mean=o[1], variance=o[2], tau_corr=o[3], R_hat=o[4])

    def to_dict(self):
        return self.get_stats().to_dict()

    def to_compound(self):
        return self.get_stats().to_compound()

    def __repr__(self):
        return repr(self.get_stats())


def online_finalize(p0, p1, n_chains, n_samples, max_lag):
    """Host arithmetic of the derived quantities (nk_online_stats_finalize) on the globally reduced sums of
    nk_online_stats_summary: ``p0`` = 3 doubles of phase 0, ``p1`` = ``4 + max_lag + 1`` doubles of phase 1."""
    n_lag = max_lag + 1 if max_lag > 0 else 0
    a0 = (C.c_double * 3)(*[float(v) for v in p0])
    a1 = (C.c_double * (_lib.NK_ONLINE_NSUM + n_lag))(*[float(v) for v in p1])
    out = (C.c_double * _lib.NK_ONLINE_NOUT)()
    acf = (C.c_double * max(n_lag, 1))()
    _lib.check(_lib.lib().nk_online_stats_finalize(a0, a1, max(int(n_chains), 1), int(n_samples), int(max_lag), out, acf))
    acf = np.array(acf[:n_lag])
    empty = float(p0[0]) == 0.0
    return dict(out=list(out), acf=None if n_lag == 0 or empty or np.isnan(acf[0]) else acf, n_chains=int(n_chains), n_samples=int(n_samples),
                empty=empty)


def _as_device_2d(data):
    if isinstance(data, OnlineStats):
        raise TypeError("expected samples, got an accumulator")
    data = getattr(data, "data", data)  # LocalEstimators wrapper
    if not isinstance(data, torch.Tensor):
        from .utils import default_device

        data = torch.from_numpy(np.ascontiguousarray(np.asarray(data))).to(default_device())
    _lib.require_cuda(data, "data")
    if data.dtype not in (torch.float32, torch.float64):
        raise TypeError("online statistics: float32 / float64 data only (real-parameter RBM => real local energies)")
    if data.ndim == 1:
        data = data[None, :]
    return data.contiguous()


def online_statistics(data, old_estimator=None, *, decay=None, max_lag=64, inplace=False):
    """Functional API (operations.py:55-129): ``est = online_statistics(batch, est)``."""
    data = _as_device_2d(data)
    if old_estimator is None:
        old_estimator = OnlineStats(data.shape[0], data.dtype, decay=decay, max_lag=max_lag, device=data.device)
        inplace = True
    return old_estimator.update(data, inplace=inplace)


def expand_max_lag(estimator, new_max_lag):
    """operations.py:132-192: lags ``0..old`` are kept, new lags start empty, the buffer grows on the left."""
    old, new_max_lag = estimator.max_lag, int(new_max_lag)
    if new_max_lag <= old:
        raise ValueError(f"new_max_lag={new_max_lag} must be > current max_lag={old}")
    if new_max_lag > _lib.NK_ONLINE_MAX_LAG:
        raise ValueError(f"max_lag must lie in [0, {_lib.NK_ONLINE_MAX_LAG}]")
    extra = new_max_lag + 1 - (old + 1 if old > 0 else 0)
    pad = torch.nn.functional.pad
    return estimator.replace(max_lag=new_max_lag, _chain_buf=pad(estimator._chain_buf, (new_max_lag - old, 0)).contiguous(),
                             **{f: pad(getattr(estimator, f), (0, extra)).contiguous()
                                for f in ("_cross_sum", "_m1_sum", "_m2_sum", "_pair_count")})


def thin_acf_by_2(estimator):
    """operations.py:195-261: re-index the lag accumulators for a twice coarser sampling cadence."""
    old = estimator.max_lag
    if old < 2:
        raise ValueError(f"max_lag={old} must be >= 2 to thin by 2")
    new = old // 2
    return estimator.replace(max_lag=new, _chain_buf=estimator._chain_buf[:, old - 2 * new::2].contiguous(), _buf_len=estimator._buf_len // 2,
                             **{f: getattr(estimator, f)[:, 0:2 * new + 1:2].contiguous()
                                for f in ("_cross_sum", "_m1_sum", "_m2_sum", "_pair_count")})


def acf_window_saturated(estimator):
    """True if the Geyer sequence found no non-positive pair inside the lag window (check_mc_convergence.py:243-255)."""
    return bool(estimator._summarise()["out"][7])


def tau_corr_reliable(estimator):
    """check_mc_convergence.py:258-272: window not saturated and at least 50 effective samples per chain."""
    return bool(estimator._summarise()["out"][8])


# ------------------------------------------------------------------------------------------------------------------
# K-channel streaming statistics with a combinator (netket/_src/stats/online_stats/accumulator_batch.py:71-257)
# ------------------------------------------------------------------------------------------------------------------
class StatsBatch:
    """Statistics of an array-valued estimator (netket/stats/mc_stats.py:183-207): ``mean`` and ``error_of_mean`` have the
    shape of the combinator's output; scalar diagnostics are not defined."""

    __slots__ = ("mean", "error_of_mean")

    def __init__(self, mean, error_of_mean):
        self.mean, self.error_of_mean = mean, error_of_mean

    @property
    def shape(self):
        return tuple(self.mean.shape)

    def __repr__(self):
        return f"StatsBatch(shape={self.shape}, max_err={float(self.error_of_mean.abs().max()):.4g})"


def _delta_method_stats(f, X, Cov):
    """accumulator_batch.py:31-68: Var[f_i(X)] ~ J_i^T Cov J_i with J = d f / d X (forward-mode Jacobian of the torch
    combinator); a scalar combinator gives a ``Stats``, an array-valued one a ``StatsBatch``; ``Cov = None`` -> NaN errors."""
    mean = f(X)
    scalar = mean.ndim == 0
    if Cov is None:
        if scalar:
            return Stats(mean=float(mean), error_of_mean=math.nan)
        return StatsBatch(mean, torch.full_like(mean, math.nan))
    J = torch.func.jacfwd(f)(X)  # (*out_shape, K)
    err = torch.sqrt(torch.clamp(torch.einsum("...k,kl,...l->...", J, Cov, J), min=0.0))
    if scalar:
        return Stats(mean=float(mean), error_of_mean=float(err))
    return StatsBatch(mean, err)


class OnlineStatsBatch:
    """K ``OnlineStats`` accumulators (one per channel of an ``(n_chains, chain_length, K)`` estimator, each updated by the
    streaming kernel) and a combinator ``f: (K,) -> scalar | array`` of their means; errors by the delta method on the
    covariance of the chain means (accumulator_batch.py:71-225).  ``combinator`` is a function of a float64 torch vector
    (differentiated with ``torch.func.jacfwd``, the stand-in for ``jax.jacfwd``)."""

    def __init__(self, estimators, combinator):
        self.estimators = tuple(estimators)
        self.combinator = combinator

    @classmethod
    def from_data(cls, data, combinator, *, max_lag=64):
        data = data if isinstance(data, torch.Tensor) else torch.as_tensor(np.asarray(data))
        if data.ndim != 3:
            raise ValueError(f"data must be 3D, got {data.ndim}D")
        from .utils import default_device

        dev = data.device if data.is_cuda else default_device()
        ests = tuple(OnlineStats(data.shape[0], max_lag=max_lag, device=dev) for _ in range(data.shape[-1]))
        return cls(ests, combinator).update(data)

    def update(self, data):
        data = data if isinstance(data, torch.Tensor) else torch.as_tensor(np.asarray(data))
        if data.ndim != 3 or data.shape[-1] != len(self.estimators):
            raise ValueError(f"data must have shape (n_chains, chain_length, {len(self.estimators)})")
        dev = self.estimators[0]._chain_count.device
        data = data.to(dev)
        return OnlineStatsBatch(tuple(e.update(data[..., k].contiguous()) for k, e in enumerate(self.estimators)), self.combinator)

    @property
    def n_samples(self):
        return self.estimators[0].n_samples

    @property
    def n_chains(self):
        """Chains of all ranks."""
        t = torch.tensor([float(self.estimators[0].n_chains)], dtype=torch.float64, device=self.estimators[0]._chain_count.device)
        return int(round(float(_allreduce(t).item())))

    def _means(self):
        dev = self.estimators[0]._chain_count.device
        return torch.tensor([e.get_stats().mean for e in self.estimators], dtype=torch.float64, device=dev)

    @property
    def mean(self):
        return self.combinator(self._means())

    @property
    def error_of_mean(self):
        return self.get_stats().error_of_mean

    def get_stats(self):
        X = self._means()
        n_chains = self.n_chains
        if n_chains < 2 or bool(torch.isnan(X).any()):
            return _delta_method_stats(self.combinator, X, None)
        dev = torch.stack([e.chain_means for e in self.estimators]) - X[:, None]  # this rank's chains
        Cov = _allreduce(dev @ dev.T) / n_chains ** 2                             # sum over the chains of all ranks
        return _delta_method_stats(self.combinator, X, Cov)

    def to_dict(self):
        return self.get_stats().to_dict()

    def to_compound(self):
        return self.get_stats().to_compound()

    def __repr__(self):
        return repr(self.get_stats())


def online_statistics_batch(data, combinator, old_estimator=None, *, max_lag=64):
    """accumulator_batch.py:228-248."""
    if old_estimator is None:
        return OnlineStatsBatch.from_data(data, combinator, max_lag=max_lag)
    return old_estimator.update(data)
