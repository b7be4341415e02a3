"""CPU baseline: the reference's *algorithm* for one VMC inner-loop step, multithreaded (TEST/BENCH INFRASTRUCTURE).

The north star names NetKet's jax[cpu] path as the CPU baseline.  jax/flax are not installed in this image (container
and GPU box share it) and there is no network, so this file is a port that keeps the reference's algorithm and
cost structure, running on all host cores through torch's CPU kernels (MKL/oneDNN GEMM + OpenMP elementwise):

  * sampling: per Metropolis step, a FULL forward pass (B x N) @ (N x M) + log_cosh + row sums on the proposed
    configurations of all chains (netket/sampler/metropolis.py:427-460, rules/local.py:40-49);
  * local energy: connected configurations are MATERIALISED, (B, K, N), and run through a full forward pass
    (B*K x N) @ (N x M) (netket/vqs/mc/kernels.py:62-71, netket/operator/_ising/jax.py:147-165), chunked over
    samples only to bound memory (the reference's chunk_size, netket/vqs/mc/kernels.py:198-223).

It is labelled "port" (never "NetKet") wherever a number from it is reported.  tests/test_reference_algorithm.py
checks it against the NumPy oracle on the same proposal stream.
"""

import time

import numpy as np
import torch

from . import rng


def _log_cosh(x):
    ax = x.abs()
    return ax + torch.log1p(torch.exp(-2.0 * ax)) - 0.6931471805599453


def logpsi(sigma, W, b, a):
    """sigma (B, N) float tensor of +-1."""
    th = sigma @ W
    if b is not None:
        th = th + b
    out = _log_cosh(th).sum(dim=-1)
    if a is not None:
        out = out + sigma @ a
    return out


def sweep_local(sigma, W, b, a, w0, u, machine_pow=2.0):
    """len(w0) Metropolis steps with LocalRule on all chains; sigma (B, N) float (+-1), modified in place.
    w0[T, B] uint32 proposal words, u[T, B] uniforms.  Returns n_accepted (B,)."""
    B, N = sigma.shape
    rows = torch.arange(B)
    logp = machine_pow * logpsi(sigma, W, b, a)
    nacc = torch.zeros(B, dtype=torch.int64)
    idx_all = torch.from_numpy(rng.index_from_word(w0, N))
    u = torch.from_numpy(np.ascontiguousarray(u)).to(W.dtype)
    for t in range(idx_all.shape[0]):
        idx = idx_all[t]
        sp = sigma.clone()
        sp[rows, idx] = -sp[rows, idx]
        logp_p = machine_pow * logpsi(sp, W, b, a)
        acc = u[t] < torch.exp(logp_p - logp)
        sigma[acc] = sp[acc]
        logp = torch.where(acc, logp_p, logp)
        nacc += acc
    return nacc


def eloc_ising(sigma, W, b, a, edges, h, J, chunk=2048):
    """E_loc of the TFIM with materialised connected configurations.  sigma (B, N) float."""
    B, N = sigma.shape
    e0 = torch.from_numpy(np.asarray(edges[:, 0], dtype=np.int64))
    e1 = torch.from_numpy(np.asarray(edges[:, 1], dtype=np.int64))
    out = torch.empty(B, dtype=torch.float64)
    flip = 1.0 - 2.0 * torch.eye(N + 1, N, dtype=W.dtype).roll(1, 0)
    flip[0] = 1.0  # slot 0 = sigma itself; slot k = site k-1 flipped
    for s in range(0, B, chunk):
        x = sigma[s:s + chunk]
        xp = x[:, None, :] * flip[None]                                    # (b, K, N) materialised
        mel0 = J * (x[:, e0] * x[:, e1]).sum(dim=-1)                       # diagonal slot
        lp = logpsi(x, W, b, a)
        lpp = logpsi(xp.reshape(-1, N), W, b, a).reshape(x.shape[0], N + 1)
        ratio = torch.exp(lpp - lp[:, None]).to(torch.float64)
        out[s:s + chunk] = mel0.to(torch.float64) * ratio[:, 0] - h * ratio[:, 1:].sum(dim=-1)
    return out


def vmc_step(sigma, W, b, a, edges, h, J, chain_length, seed, t0, machine_pow=2.0):
    """chain_length x (one sweep of N proposals + E_loc of the resulting sample) for all chains.
    Returns (eloc[B, chain_length], n_accepted[B], t_next)."""
    B, N = sigma.shape
    dtype = np.float32 if W.dtype == torch.float32 else np.float64
    eloc = torch.empty((B, chain_length), dtype=torch.float64)
    nacc = torch.zeros(B, dtype=torch.int64)
    for s in range(chain_length):
        words, u = rng.proposal_stream(seed, t0, N, np.arange(B), dtype)
        nacc += sweep_local(sigma, W, b, a, words[..., 0], u, machine_pow)
        t0 += N
        eloc[:, s] = eloc_ising(sigma, W, b, a, edges, h, J)
    return eloc, nacc, t0


def time_vmc_steps(N, M, B, chain_length, edges, h, J, dtype, steps, warmup, seed=15324, threads=None):
    """Times `steps` calls of vmc_step on random-init parameters; returns (samples_per_s, ms_per_step, threads)."""
    from . import hilbert as ohilbert
    from . import rbm as orbm

    if threads:
        torch.set_num_threads(threads)
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    Wn, bn, an = orbm.init_params(N, M // N, seed=1234, std=0.01, dtype=dtype)
    W, b, a = (torch.from_numpy(x).to(tdt) for x in (Wn, bn, an))
    sigma = torch.from_numpy(ohilbert.random_state(seed, B, N)).to(tdt)
    t0 = 0
    for _ in range(warmup):
        _, _, t0 = vmc_step(sigma, W, b, a, edges, h, J, chain_length, seed, t0)
    tic = time.perf_counter()
    for _ in range(steps):
        _, _, t0 = vmc_step(sigma, W, b, a, edges, h, J, chain_length, seed, t0)
    dt = time.perf_counter() - tic
    return B * chain_length * steps / dt, dt / steps * 1e3, torch.get_num_threads()
