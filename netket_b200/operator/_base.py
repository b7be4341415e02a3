"""The ``DiscreteJaxOperator`` contract (netket/operator/_discrete_operator_jax.py:29-139,199-230):
``max_conn_size``, ``get_conn_padded(x) -> (x', mels)``, ``n_conn(x)``; arbitrary leading batch
dimensions are flattened and restored; ``x'.dtype == x.dtype``; ``mels.dtype == op.dtype``
(test/operator/test_operator.py:345-359)."""

import numpy as np
import torch

from .. import _lib
from ..utils import default_device


class _Batch:
    """Normalises the input of get_conn_padded to an int8 CUDA tensor [B, N] and restores type/shape after."""

    def __init__(self, x, N):
        self.is_numpy = not isinstance(x, torch.Tensor)
        if self.is_numpy:
            x = np.asarray(x)
            self.np_dtype = x.dtype
            t = torch.from_numpy(np.ascontiguousarray(x.astype(np.int8))).to(default_device())
        else:
            if not x.is_cuda:
                raise _lib.NkError("operator inputs must be CUDA tensors or numpy arrays")
            t = x
        if t.shape[-1] != N:
            raise ValueError(f"the last dimension of x must be the Hilbert size {N}, got {tuple(t.shape)}")
        self.in_dtype = t.dtype
        self.batch_shape = tuple(t.shape[:-1])
        self.device = t.device
        self.x8 = t.reshape(-1, N).to(torch.int8).contiguous()
        self.B = self.x8.shape[0]

    def restore_states(self, xp, K):
        xp = xp.reshape(*self.batch_shape, K, xp.shape[-1])
        if self.is_numpy:
            return xp.cpu().numpy().astype(self.np_dtype)
        return xp if self.in_dtype == torch.int8 else xp.to(self.in_dtype)

    def restore(self, t, *tail):
        t = t.reshape(*self.batch_shape, *tail)
        return t.cpu().numpy() if self.is_numpy else t


class DiscreteJaxOperator:
    """Abstract base: subclasses implement ``_conn(x8[B,N]) -> (xp[B,K,N] int8, mels[B,K], n_conn[B] or None)``."""

    def __init__(self, hilbert, dtype):
        self._hilbert = hilbert
        self._dtype = np.dtype(dtype)

    @property
    def hilbert(self):
        return self._hilbert

    @property
    def dtype(self):
        return self._dtype

    @property
    def size(self):
        return self._hilbert.size

    @property
    def max_conn_size(self):
        raise NotImplementedError

    def get_conn_padded(self, x):
        b = _Batch(x, self.hilbert.size)
        xp, mels, _ = self._conn(b.x8, want_nconn=False)
        K = self.max_conn_size
        return b.restore_states(xp, K), b.restore(mels, K)

    def n_conn(self, x, out=None):
        if out is not None:
            raise NotImplementedError("operators do not support passing the `out` argument to operator.n_conn().")
        b = _Batch(x, self.hilbert.size)
        return b.restore(self._n_conn(b.x8))

    def to_dense(self):
        """Dense matrix in the reference's basis ordering (small Hilbert spaces only; host numpy)."""
        hi = self.hilbert
        states = hi.all_states()
        N = hi.size
        w = 1 << np.arange(N - 1, -1, -1, dtype=np.int64)
        nums = (((1 - states.astype(np.int64)) // 2) * w).sum(axis=1)
        lut = -np.ones(1 << N, dtype=np.int64)
        lut[nums] = np.arange(len(states))
        xp, mels = self.get_conn_padded(states)
        D = len(states)
        cols = lut[(((1 - xp.astype(np.int64)) // 2) * w).sum(axis=-1)]
        H = np.zeros((D, D), dtype=self.dtype)
        for r in range(D):
            nz = mels[r] != 0
            if np.any(cols[r][nz] < 0):
                raise ValueError("operator connects outside the constrained Hilbert space")
            np.add.at(H[r], cols[r][nz], mels[r][nz])
        return H
