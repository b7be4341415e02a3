"""``nk.operator.Ising`` = ``IsingJax`` (netket/operator/_ising/jax.py:35-175, _ising/base.py:32-224).

H = -h sum_i sx_i + J sum_<ij> sz_i sz_j.  ``get_conn_padded`` runs ``nk_ising_conn``; inside
``MCState.expect`` the operator is never materialised (``nk_eloc_ising_rbm`` / fused sweep).
"""

import ctypes as C

import numpy as np
import torch

from .. import _lib
from ._base import DiscreteJaxOperator


class IsingJax(DiscreteJaxOperator):
    def __init__(self, hilbert, graph, h, J=1.0, dtype=None):
        if len(hilbert.local_states) != 2:
            raise ValueError("IsingJax only supports Hamiltonians with two local states")
        if dtype is None:
            dtype = np.result_type(float, np.asarray(h if h is not None else 0.0).dtype, np.asarray(J).dtype)
        super().__init__(hilbert, dtype)
        if np.dtype(dtype) not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise NotImplementedError("netket_b200 Ising supports real float32/float64 matrix elements")
        self._h = 0.0 if h is None else float(h)
        self._J = float(J)
        edges = graph.edges() if hasattr(graph, "edges") else graph
        edges = np.asarray([tuple(e[:2]) for e in edges], dtype=np.int32).reshape(-1, 2)
        if edges.size and (edges.min() < 0 or edges.max() >= hilbert.size):
            raise ValueError("graph edges refer to sites outside the Hilbert space")
        self._edges_np = edges
        self._edges_dev = {}

    @property
    def h(self):
        return self._h

    @property
    def J(self):
        return self._J

    @property
    def edges(self):
        return self._edges_np

    @property
    def is_hermitian(self):
        return True

    @property
    def max_conn_size(self):
        """N+1 (_ising/base.py:158-161); 1 when h == 0 (StaticZero, jax.py:60-61,127-131)."""
        return 1 if self._h == 0.0 else self.hilbert.size + 1

    def _edges_on(self, device):
        key = str(device)
        if key not in self._edges_dev:
            self._edges_dev[key] = torch.from_numpy(self._edges_np.copy()).to(device)
        return self._edges_dev[key]

    def _c_struct(self, device):
        e = self._edges_on(device)
        return _lib.nk_ising_t(edges=e.data_ptr(), n_edges=int(self._edges_np.shape[0]), reserved=0, h=self._h, J=self._J)

    def _conn(self, x8, want_nconn=False):
        B, N = x8.shape
        K = self.max_conn_size
        dev = x8.device
        xp = torch.empty((B, K, N), dtype=torch.int8, device=dev)
        mels = torch.empty((B, K), dtype=_lib.torch_dtype(self.dtype), device=dev)
        op = self._c_struct(dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nk_ising_conn(_lib.stream_ptr(dev), C.byref(op), _lib.ptr(x8), B, N, _lib.ptr(xp),
                                                _lib.ptr(mels), _lib.dtype_code(self.dtype)))
        return xp, mels, None

    def _n_conn(self, x8):
        B, N = x8.shape
        out = torch.empty((B,), dtype=torch.int32, device=x8.device)
        op = self._c_struct(x8.device)
        with torch.cuda.device(x8.device):
            _lib.check(_lib.lib().nk_ising_n_conn(_lib.stream_ptr(x8.device), C.byref(op), _lib.ptr(x8), B, N, _lib.ptr(out)))
        return out

    def __repr__(self):
        return f"IsingJax(J={self._J}, h={self._h}; dim={self.hilbert.size})"


Ising = IsingJax
