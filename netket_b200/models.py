"""``nk.models.RBM`` with real parameters (netket/models/rbm.py:32-81).

logpsi(sigma) = sum_j log cosh(sum_i sigma_i W_ij + b_j) + sum_i a_i sigma_i, evaluated by ``nk_rbm_logpsi``.
Parameters use Flax's pytree layout ``{"params": {"Dense": {"kernel": (N, M), "bias": (M,)}, "visible_bias": (N,)}}``
so that state-dicts interchange with the reference.  Other ansaetze are out of scope: the fused
kernels recognise this model only and everything else raises.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib
from .utils import default_device


class RBM:
    def __init__(self, alpha=1, param_dtype=np.float64, use_hidden_bias=True, use_visible_bias=True, *,
                 kernel_init_std=0.01, hidden_bias_init_std=0.01, visible_bias_init_std=0.01, activation=None, precision=None):
        if activation is not None:
            raise NotImplementedError("netket_b200 RBM implements the log_cosh activation only")
        self.alpha = alpha
        self.param_dtype = np.dtype(param_dtype)
        _lib.dtype_code(self.param_dtype)  # validates float32 / float64
        self.use_hidden_bias = bool(use_hidden_bias)
        self.use_visible_bias = bool(use_visible_bias)
        self._std = (kernel_init_std, hidden_bias_init_std, visible_bias_init_std)

    def n_hidden(self, N):
        return int(self.alpha * N)

    def init(self, seed, sigma_or_N, *, device=None):
        """Random-init parameters, normal(stddev=0.01) by default (rbm.py:29,50-55), drawn in fp64 from
        numpy.random.default_rng(seed) in the order W, b, a and then cast (BASELINE.md synthetic inputs)."""
        N = int(sigma_or_N) if np.isscalar(sigma_or_N) else int(sigma_or_N.shape[-1])
        M = self.n_hidden(N)
        device = default_device(device)
        g = np.random.default_rng(seed)
        W = g.normal(0.0, self._std[0], size=(N, M))
        b = g.normal(0.0, self._std[1], size=(M,))
        a = g.normal(0.0, self._std[2], size=(N,))
        td = _lib.torch_dtype(self.param_dtype)
        dense = {"kernel": torch.from_numpy(W).to(device=device, dtype=td)}
        if self.use_hidden_bias:
            dense["bias"] = torch.from_numpy(b).to(device=device, dtype=td)
        params = {"Dense": dense}
        if self.use_visible_bias:
            params["visible_bias"] = torch.from_numpy(a).to(device=device, dtype=td)
        return {"params": params}

    @staticmethod
    def unpack(variables):
        """-> (W, b|None, a|None) tensors from a Flax-layout pytree (or its ``params`` sub-dict)."""
        p = variables.get("params", variables)
        W = p["Dense"]["kernel"]
        return W, p["Dense"].get("bias"), p.get("visible_bias")

    @staticmethod
    def c_struct(variables):
        W, b, a = RBM.unpack(variables)
        _lib.require_cuda(W, "Dense.kernel")
        for t, n in ((b, "Dense.bias"), (a, "visible_bias")):
            if t is not None:
                _lib.require_cuda(t, n)
                if t.dtype != W.dtype:
                    raise TypeError("all RBM parameters must share one dtype")
        N, M = W.shape
        if b is not None and b.shape != (M,):
            raise ValueError("Dense.bias has the wrong shape")
        if a is not None and a.shape != (N,):
            raise ValueError("visible_bias has the wrong shape")
        return _lib.nk_rbm_t(W=W.data_ptr(), b=b.data_ptr() if b is not None else None,
                             a=a.data_ptr() if a is not None else None, N=N, M=M, dtype=_lib.dtype_code(W.dtype), reserved=0)

    def apply(self, variables, sigma, *, return_theta=False):
        """logpsi for sigma[..., N] (int8 CUDA tensor, or numpy which is uploaded)."""
        W, _, _ = self.unpack(variables)
        rbm = self.c_struct(variables)
        is_np = not isinstance(sigma, torch.Tensor)
        if is_np:
            sigma = torch.from_numpy(np.ascontiguousarray(np.asarray(sigma).astype(np.int8))).to(W.device)
        if sigma.shape[-1] != rbm.N:
            raise ValueError(f"input has {sigma.shape[-1]} sites, the model has {rbm.N}")
        batch = tuple(sigma.shape[:-1])
        s8 = sigma.reshape(-1, rbm.N).to(torch.int8).contiguous()
        B = s8.shape[0]
        out = torch.empty((B,), dtype=W.dtype, device=W.device)
        theta = torch.empty((B, rbm.M), dtype=W.dtype, device=W.device) if return_theta else None
        with torch.cuda.device(W.device):
            _lib.check(_lib.lib().nk_rbm_logpsi(_lib.stream_ptr(W.device), C.byref(rbm), _lib.ptr(s8), B, _lib.ptr(out),
                                                _lib.ptr(theta)))
        out = out.reshape(batch)
        if is_np:
            out = out.cpu().numpy()
        if return_theta:
            return out, theta.reshape(*batch, rbm.M)
        return out

    __call__ = apply

    def __repr__(self):
        return (f"RBM(alpha={self.alpha}, param_dtype={self.param_dtype}, use_hidden_bias={self.use_hidden_bias}, "
                f"use_visible_bias={self.use_visible_bias})")
