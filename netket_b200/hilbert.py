"""Spin-1/2 Hilbert space: the part of netket.hilbert the hot path needs.

Mirrors ``nk.hilbert.Spin(s, N, total_sz)`` (netket/hilbert/spin.py:56-241) for s = 1/2:
local states ``[+1, -1]`` (StaticRange start=1, step=-2 => local index 0 <-> +1, 1 <-> -1,
spin.py:165-171), optional total magnetisation constraint (``SumConstraint(2*total_sz)``, :181-187),
``random_state`` (netket/hilbert/random/homogeneous.py:35-72, random/fock.py:77-97) which runs on the
GPU through ``nk_random_state``.  Everything else in netket.hilbert is out of scope.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib
from .utils import default_device, split_seed


class Spin:
    def __init__(self, s=0.5, N=1, *, total_sz=None):
        if round(2 * s + 1) != 2:
            raise NotImplementedError("netket_b200 implements Spin(s=1/2) only (SURVEY.md §2: other local dimensions are out of scope)")
        if N < 1:
            raise ValueError("N must be positive")
        self._s = 0.5
        self._N = int(N)
        if total_sz is not None:
            m = round(2 * total_sz)
            if abs(m) > N or (N + m) % 2 != 0:
                raise ValueError(f"Cannot fix the total magnetization: 2|M| = {abs(m)} incompatible with N = {N}")
        self._total_sz = total_sz

    # -- structure ------------------------------------------------------------------------
    @property
    def size(self):
        return self._N

    @property
    def shape(self):
        return (2,) * self._N

    @property
    def local_states(self):
        return np.array([1, -1], dtype=np.int8)

    @property
    def local_size(self):
        return 2

    def size_at_index(self, i):
        return 2

    def states_at_index(self, i):
        return [1, -1]

    @property
    def constrained(self):
        return self._total_sz is not None

    @property
    def total_sz(self):
        return self._total_sz

    @property
    def is_finite(self):
        return True

    @property
    def n_down(self):
        """Number of -1 sites fixed by the constraint (n_excitations of homogeneous.py:64-67), or -1."""
        if self._total_sz is None:
            return -1
        return (self._N - round(2 * self._total_sz)) // 2

    @property
    def n_states(self):
        if self._total_sz is None:
            return 2 ** self._N
        from math import comb

        return comb(self._N, self.n_down)

    # -- conversions ----------------------------------------------------------------------
    def states_to_local_indices(self, x):
        """(x - start)/step with start=1, step=-2 (netket/utils/static_range.py:148-170)."""
        if isinstance(x, torch.Tensor):
            return ((1 - x.to(torch.int64)) // 2).to(torch.uint8)
        return ((1 - np.asarray(x).astype(np.int64)) // 2).astype(np.uint8)

    def local_indices_to_states(self, idx, dtype=None):
        if isinstance(idx, torch.Tensor):
            return (1 - 2 * idx.to(torch.int64)).to(dtype or torch.int8)
        return (1 - 2 * np.asarray(idx).astype(np.int64)).astype(dtype or np.int8)

    def all_states(self):
        """All basis states (host numpy, small N only), index 0 = all up; constrained spaces keep the order."""
        if self._N > 26:
            raise ValueError("Hilbert space too large to enumerate")
        nums = np.arange(1 << self._N, dtype=np.int64)
        bits = (nums[:, None] >> np.arange(self._N - 1, -1, -1)[None, :]) & 1
        st = (1 - 2 * bits).astype(np.int8)
        if self._total_sz is not None:
            st = st[st.astype(np.int64).sum(axis=1) == round(2 * self._total_sz)]
        return st

    # -- random states ----------------------------------------------------------------------
    def random_state(self, key=None, size=None, dtype=None, *, chain_offset=0, device=None):
        """``hilbert.random_state(key, n)`` -> sigma[n, N] int8 on the GPU (Philox STREAM_INIT stream)."""
        seed = split_seed(key)
        scalar = size is None
        n = 1 if scalar else int(np.prod(size))
        device = default_device(device)
        out = torch.empty((n, self._N), dtype=torch.int8, device=device)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().nk_random_state(_lib.stream_ptr(device), _lib.ptr(out), n, self._N, self.n_down,
                                                  C.c_uint64(seed), C.c_uint64(chain_offset)))
        if dtype is not None and dtype not in (torch.int8, np.int8):
            out = out.to(_lib.torch_dtype(dtype))
        if scalar:
            return out[0]
        return out.reshape(*(size if isinstance(size, (tuple, list)) else (size,)), self._N)

    def __eq__(self, o):
        return isinstance(o, Spin) and o._N == self._N and o._total_sz == self._total_sz

    def __hash__(self):
        return hash(("Spin", self._N, self._total_sz))

    def __repr__(self):
        c = f", total_sz={self._total_sz}" if self._total_sz is not None else ""
        return f"Spin(s=1/2, N={self._N}{c})"
