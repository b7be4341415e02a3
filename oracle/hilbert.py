"""Spin-1/2 Hilbert space conventions and random states (TEST INFRASTRUCTURE).

Follows
  * netket/hilbert/spin.py:158-193       local states = StaticRange(start=1, step=-2, length=2)
                                          => local index 0 <-> sigma=+1, index 1 <-> sigma=-1;
                                          total_sz -> SumConstraint(round(2*total_sz))
  * netket/utils/static_range.py:148-192  states_to_numbers = (x-start)/step ; numbers_to_states
  * netket/hilbert/random/homogeneous.py:35-47   unconstrained random_state (uniform local index)
  * netket/hilbert/random/homogeneous.py:50-72 + random/fock.py:77-97  constrained: n_excitations
        ones followed by a uniform random permutation
  * netket/hilbert/random/homogeneous.py:146-170 flip_state_scalar: 2 local states => idx <-> 1-idx
"""

import numpy as np

from . import rng

START = 1
STEP = -2


def states_to_local_indices(x):
    """sigma (+1/-1) -> local index (0/1).  static_range.py:148-170."""
    return ((np.asarray(x).astype(np.int64) - START) // STEP).astype(np.int64)


def local_indices_to_states(idx, dtype=np.int8):
    """local index -> sigma.  static_range.py:172-192."""
    return np.ascontiguousarray((START + STEP * np.asarray(idx).astype(np.int64)).astype(dtype))


def n_excitations(N, total_sz):
    """Number of local-index-1 (spin down) sites for a given total_sz.

    homogeneous.py:64-67: n_excitations = (sum_value - start*size)//step with
    sum_value = round(2*total_sz)  (spin.py:181-187).
    """
    sum_value = round(2 * total_sz)
    return (sum_value - START * N) // STEP


def all_states(N, total_sz=None):
    """All basis states in the reference's ordering (index 0 = all up = +1).

    The ordering is the lexicographic order of local-index strings, first site most
    significant (netket/hilbert/index/unconstrained.py; constrained spaces keep the same
    relative order, netket/hilbert/index/constraints.py).
    """
    n = 1 << N
    nums = np.arange(n, dtype=np.int64)
    bits = (nums[:, None] >> np.arange(N - 1, -1, -1)[None, :]) & 1
    states = local_indices_to_states(bits)
    if total_sz is not None:
        keep = states.astype(np.int64).sum(axis=1) == round(2 * total_sz)
        states = states[keep]
    return states


def states_to_numbers(states, N):
    """Index of a state in the unconstrained ordering above."""
    idx = states_to_local_indices(states)
    w = 1 << np.arange(N - 1, -1, -1, dtype=np.int64)
    return (idx * w).sum(axis=-1)


def random_state(seed, n_chains, N, total_sz=None, chain_offset=0):
    """Initial configurations, drawn from the Philox STREAM_INIT stream.

    Distribution follows the reference (uniform over the product space, or a uniform
    random permutation of a fixed multiset, fock.py:91-97); the bit stream is this
    repository's definition (oracle/rng.py) and is reproduced by ``nk_random_state``:

      unconstrained: bit i of the 128-bit block k=i//128 (words w0..w3 little-endian) is
                     the local index of site i  (0 -> +1, 1 -> -1);
      constrained:   start from [1]*n_exc + [0]*(N-n_exc) (local indices), then
                     Fisher-Yates for i = N-1 .. 1:  j = (word_i * (i+1)) >> 32, swap(i, j),
                     word_i = word (i%4) of Philox block i//4.
    """
    chains = np.arange(n_chains, dtype=np.uint64) + np.uint64(chain_offset)
    if total_sz is None:
        n_blocks = (N + 127) // 128
        words = rng.step_words(seed, np.arange(n_blocks, dtype=np.uint64)[None, :], chains[:, None], rng.STREAM_INIT)
        words = words.reshape(n_chains, n_blocks * 4)
        site = np.arange(N)
        bits = (words[:, site // 32] >> (site % 32).astype(np.uint32)) & np.uint32(1)
        return np.ascontiguousarray(local_indices_to_states(bits))
    n_exc = n_excitations(N, total_sz)
    if not (0 <= n_exc <= N):
        raise ValueError("total_sz incompatible with N")
    idx = np.zeros((n_chains, N), dtype=np.int64)
    idx[:, :n_exc] = 1
    n_blocks = (N + 3) // 4
    words = rng.step_words(seed, np.arange(n_blocks, dtype=np.uint64)[None, :], chains[:, None], rng.STREAM_INIT)
    words = words.reshape(n_chains, n_blocks * 4)
    rows = np.arange(n_chains)
    for i in range(N - 1, 0, -1):
        j = rng.index_from_word(words[:, i], i + 1)
        tmp = idx[rows, i].copy()
        idx[rows, i] = idx[rows, j]
        idx[rows, j] = tmp
    return np.ascontiguousarray(local_indices_to_states(idx))
