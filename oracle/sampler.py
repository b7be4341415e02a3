"""Metropolis-Hastings sampler, reference algorithm (TEST INFRASTRUCTURE).

Follows netket/sampler/metropolis.py:382-505:
  * ``_reset``  (:382-414): log_prob = machine_pow * Re logpsi(sigma); counters zeroed.
  * ``_sample_next`` loop body (:427-460): propose, **full forward pass** on the proposal,
    ``accept = u < exp(logp' - logp [+ corr])``, select; ``n_accepted_proc += accept``;
    ``n_steps_proc += n_chains``.
  * ``_sample_chain`` (:466-505): ``chain_length`` x ``sweep_size`` steps, one recorded
    sample per sweep, output (n_chains, chain_length, N).
Rules:
  * LocalRule (netket/sampler/rules/local.py:40-49): uniform site per chain, deterministic
    flip for 2 local states (netket/hilbert/random/homogeneous.py:157-159), no correction.
  * ExchangeRule (netket/sampler/rules/exchange.py:143-184): one *hoppable* cluster
    (sigma_i != sigma_j) chosen uniformly, swap, correction
    log n_hop(sigma) - log n_hop(sigma').  With ``probabilities`` (:86-123,155-160,177-182)
    the cluster is drawn with weight mask * p by inverse CDF -- JAX's ``random.choice``
    algorithm, ``searchsorted(cumsum(w), total * r)`` -- at r = (w0 + 1/2) / 2^32, and the
    correction is log sum_c w_c(sigma) - log sum_c w_c(sigma').
The per-step randomness comes from the explicit proposal stream of oracle/rng.py (or from
arrays passed in ``stream=``), never from a hidden generator.
"""

import numpy as np

from . import rng
from .rbm import logpsi


def hoppable_mask(sigma, clusters):
    """_compute_different_clusters_mask (exchange.py:208-218), integer sigma."""
    return sigma[..., clusters[:, 0]] != sigma[..., clusters[:, 1]]


def _propose_local(sigma, w0):
    N = sigma.shape[1]
    idx = rng.index_from_word(w0, N)
    sp = sigma.copy()
    rows = np.arange(sigma.shape[0])
    sp[rows, idx] = -sp[rows, idx]
    return sp, None, idx


def _propose_exchange_weighted(sigma, w0, clusters, prob):
    B = sigma.shape[0]
    mask = hoppable_mask(sigma, clusters)
    w = mask * prob[None, :]                      # _weighted_mask (exchange.py:155-160)
    cs = np.cumsum(w, axis=1)
    tot = cs[:, -1]
    r = tot * ((np.asarray(w0, dtype=np.float64) + 0.5) * 2.0 ** -32)
    sel = np.minimum((cs < r[:, None]).sum(axis=1), clusters.shape[0] - 1)   # searchsorted(cumsum, r), side="left"
    rows = np.arange(B)
    si, sj = clusters[sel, 0], clusters[sel, 1]
    sp = sigma.copy()
    ok = tot > 0
    sp[rows[ok], si[ok]] = sigma[rows[ok], sj[ok]]
    sp[rows[ok], sj[ok]] = sigma[rows[ok], si[ok]]
    tot_p = (hoppable_mask(sp, clusters) * prob[None, :]).sum(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        corr = np.log(tot) - np.log(tot_p)
    corr = np.where(ok, corr, np.nan)
    return sp, corr, sel


def _propose_exchange(sigma, w0, clusters):
    B = sigma.shape[0]
    mask = hoppable_mask(sigma, clusters)
    n_hop = mask.sum(axis=1)
    k = rng.index_from_word(w0, np.maximum(n_hop, 1).astype(np.uint64))
    # k-th hoppable cluster in cluster order
    cs = np.cumsum(mask, axis=1)
    sel = np.argmax((cs == (k + 1)[:, None]) & mask, axis=1)
    rows = np.arange(B)
    si = clusters[sel, 0]
    sj = clusters[sel, 1]
    sp = sigma.copy()
    ok = n_hop > 0
    sp[rows[ok], si[ok]] = sigma[rows[ok], sj[ok]]
    sp[rows[ok], sj[ok]] = sigma[rows[ok], si[ok]]
    n_hop_p = hoppable_mask(sp, clusters).sum(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        corr = np.log(n_hop.astype(np.float64)) - np.log(n_hop_p.astype(np.float64))
    # no hoppable cluster: the reference's correction is log(0) - log(0) = nan (rules/exchange.py:177-182), so that
    # `u < exp(nan)` is False: the (identity) proposal is rejected and not counted as accepted
    corr = np.where(ok, corr, np.nan)
    return sp, corr, sel


def reset_log_prob(sigma, W, b, a, machine_pow=2.0):
    """metropolis.py:399-403."""
    return (W.dtype.type(machine_pow) * logpsi(sigma, W, b, a)).astype(W.dtype)


def sample_chain(
    rule,
    sigma,
    W,
    b,
    a,
    *,
    chain_length,
    sweep_size=None,
    machine_pow=2.0,
    seed=0,
    t0=0,
    chain_offset=0,
    clusters=None,
    stream=None,
    return_trace=False,
    probabilities=None,
):
    """Run ``chain_length`` sweeps.  sigma[B,N] int8 (+/-1).  rule in {"local","exchange"}.

    stream: optional (w0[T,B] uint32, u[T,B]) explicit proposal stream, T = chain_length*sweep_size.
    Returns dict with samples[B,chain_length,N], log_prob_samples[B,chain_length], sigma, log_prob,
    n_accepted[B] (int64), n_steps (= T*B, metropolis.py:459), t (= t0 + T); with ``return_trace`` also
    ``trace``: per step (selected site / cluster [B], accepted [B], margin [B] = log u - (logp' - logp + corr)).
    """
    sigma = np.array(sigma, dtype=np.int8, copy=True)
    B, N = sigma.shape
    dtype = W.dtype
    sweep_size = N if sweep_size is None else sweep_size
    T = chain_length * sweep_size
    if stream is None:
        chains = np.arange(B, dtype=np.uint64) + np.uint64(chain_offset)
        words, u = rng.proposal_stream(seed, t0, T, chains, dtype)
        w0 = words[..., 0]
    else:
        w0, u = stream
        w0 = np.asarray(w0, dtype=np.uint32)
        u = np.asarray(u, dtype=dtype)
    pw = dtype.type(machine_pow)
    logp = reset_log_prob(sigma, W, b, a, machine_pow)
    n_acc = np.zeros(B, dtype=np.int64)
    samples = np.empty((B, chain_length, N), dtype=np.int8)
    lps = np.empty((B, chain_length), dtype=dtype)
    trace = [] if return_trace else None
    if rule == "exchange":
        clusters = np.asarray(clusters)
    t = 0
    for s in range(chain_length):
        for _ in range(sweep_size):
            if rule == "local":
                sp, corr, sel = _propose_local(sigma, w0[t])
            elif rule == "exchange" and probabilities is not None:
                sp, corr, sel = _propose_exchange_weighted(sigma, w0[t], clusters, np.asarray(probabilities, dtype=np.float64))
            elif rule == "exchange":
                sp, corr, sel = _propose_exchange(sigma, w0[t], clusters)
            else:
                raise NotImplementedError(rule)
            logp_p = (pw * logpsi(sp, W, b, a)).astype(dtype)
            arg = logp_p - logp
            if corr is not None:
                arg = arg + corr.astype(dtype)
            with np.errstate(over="ignore", invalid="ignore"):
                acc = u[t] < np.exp(arg)
            sigma = np.where(acc[:, None], sp, sigma)
            logp = np.where(acc, logp_p, logp)
            n_acc += acc
            if return_trace:
                with np.errstate(divide="ignore", invalid="ignore"):
                    margin = np.log(u[t].astype(np.float64)) - arg.astype(np.float64)  # accept <=> margin < 0
                trace.append((sel.copy(), acc.copy(), margin))
            t += 1
        samples[:, s, :] = sigma
        lps[:, s] = logp
    out = dict(samples=samples, log_prob_samples=lps, sigma=sigma, log_prob=logp, n_accepted=n_acc,
               n_steps=T * B, t=t0 + T)
    if return_trace:
        out["trace"] = trace
    return out


def exact_distribution(W, b, a, states, machine_pow=2.0):
    """|psi|^machine_pow normalised over ``states`` (test/sampler/test_sampler.py:399-457)."""
    lp = machine_pow * logpsi(states, W.astype(np.float64), None if b is None else b.astype(np.float64),
                              None if a is None else a.astype(np.float64))
    p = np.exp(lp - lp.max())
    return p / p.sum()
