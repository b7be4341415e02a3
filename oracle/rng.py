"""Counter-based proposal stream shared by the oracle and the CUDA kernels.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws, per Metropolis step and per chain, one proposal index and one
uniform (netket/sampler/rules/local.py:41-47, netket/sampler/metropolis.py:429,444,
netket/sampler/rules/exchange.py:149,162-167) from JAX's threefry generator.  That
bit-stream is unpinned (no golden vectors, JAX absent), so the stream is *defined* here:

    Philox4x32-10 (Salmon et al., SC'11; Random123), key = (seed_lo, seed_hi),
    counter = (t_lo, t_hi, chain_lo, (chain_hi & 0x00ffffff) | (stream << 24))

with ``t`` the index of the Metropolis step of that chain since ``init_state`` and
``chain`` the *global* chain index (so a run sharded over G GPUs reproduces the 1-GPU
run).  The four output words of step ``t`` are used as

    w0 -> proposal index      idx = (w0 * n) >> 32           (Lemire multiply-shift)
    w1 -> uniform, fp32 mode  u   = (w1 >> 8) * 2^-24        in [0, 1)
    w1,w2 -> uniform, fp64    u   = ((w1 >> 5) * 2^26 + (w2 >> 6)) * 2^-53
    w3 -> spare

STREAM_STEP (0) feeds the Metropolis steps, STREAM_INIT (1) feeds ``random_state``.
"""

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)

STREAM_STEP = 0
STREAM_INIT = 1

_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """Vectorised Philox4x32-10.

    ctr: uint32 array (..., 4); key: uint32 array (..., 2) (broadcastable).
    Returns uint32 array (..., 4).
    """
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    c0 = ctr[..., 0].astype(np.uint64)
    c1 = ctr[..., 1].astype(np.uint64)
    c2 = ctr[..., 2].astype(np.uint64)
    c3 = ctr[..., 3].astype(np.uint64)
    k0 = np.broadcast_to(key[..., 0], c0.shape).astype(np.uint64)
    k1 = np.broadcast_to(key[..., 1], c0.shape).astype(np.uint64)
    for r in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0), lo1, (hi0 ^ c3 ^ k1), lo0
        if r != 9:
            k0 = (k0 + np.uint64(PHILOX_W0)) & _MASK32
            k1 = (k1 + np.uint64(PHILOX_W1)) & _MASK32
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _counter(t, chain, stream):
    t = np.asarray(t, dtype=np.uint64)
    chain = np.asarray(chain, dtype=np.uint64)
    t, chain = np.broadcast_arrays(t, chain)
    c = np.empty(t.shape + (4,), dtype=np.uint32)
    c[..., 0] = (t & _MASK32).astype(np.uint32)
    c[..., 1] = (t >> np.uint64(32)).astype(np.uint32)
    c[..., 2] = (chain & _MASK32).astype(np.uint32)
    c[..., 3] = (((chain >> np.uint64(32)) & np.uint64(0x00FFFFFF)) | (np.uint64(stream) << np.uint64(24))).astype(
        np.uint32
    )
    return c


def _key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32)


def step_words(seed, t, chain, stream=STREAM_STEP):
    """The four Philox words of (chain, step t).  t, chain broadcast."""
    return philox4x32_10(_counter(t, chain, stream), _key(seed))


def index_from_word(w0, n):
    """idx = floor(w0 * n / 2^32)  (what the kernel computes with __umulhi)."""
    return ((np.asarray(w0, dtype=np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def uniform_from_words(words, dtype):
    """u in [0,1) in the working precision (see module docstring)."""
    w1 = words[..., 1]
    w2 = words[..., 2]
    if np.dtype(dtype) == np.float32:
        return ((w1 >> np.uint32(8)).astype(np.float32) * np.float32(2.0**-24)).astype(np.float32)
    hi = (w1 >> np.uint32(5)).astype(np.float64)
    lo = (w2 >> np.uint32(6)).astype(np.float64)
    return (hi * 67108864.0 + lo) * (2.0**-53)


def proposal_stream(seed, t0, n_steps, chains, dtype):
    """Raw words for steps t0..t0+n_steps-1 of the given global chain indices.

    Returns words[n_steps, n_chains, 4] uint32 and u[n_steps, n_chains].
    """
    t = (np.uint64(t0) + np.arange(n_steps, dtype=np.uint64))[:, None]
    ch = np.asarray(chains, dtype=np.uint64)[None, :]
    words = step_words(seed, t, ch, STREAM_STEP)
    return words, uniform_from_words(words, dtype)
