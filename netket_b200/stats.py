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


def statistics(data):
    """Statistics of ``data[n_chains, L]`` (or a 1-D time series) held on this rank's GPU.  Under
    torch.distributed the chains of all ranks are combined (chain axis sharded, as the reference does with its
    mesh axis "S": netket/stats/mc_stats_old.py:96-107)."""
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
