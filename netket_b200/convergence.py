"""Streaming callers of the sample + local-estimator loop:

* ``expect_to_precision``   netket/_src/vqs/expect_to_precision.py:69-164
* ``check_mc_convergence``  netket/_src/vqs/check_mc_convergence.py:31-240
* ``thermalise_mcmc``       netket/_src/vqs/check_mc_convergence.py:275-457

Every iteration is ``state.sample(n_discard_per_chain=0); state.local_estimators(op)`` followed by an update of the
:class:`~netket_b200.stats.OnlineStats` accumulator.  Here the first two are ONE launch (sweeps fused with the local
energy, ``MCState._sample_and_estimate``) and the accumulator update is one more (``nk_online_stats_update``); per
iteration only the ``4 + max_lag + 4`` summary doubles come back to the host for the stopping test.
"""

import copy
import math
import warnings

from tqdm.auto import tqdm

from .sampler import MetropolisSampler
from .stats import acf_window_saturated, expand_max_lag, online_statistics, tau_corr_reliable, thin_acf_by_2
from .utils import world


class History:
    """One logged quantity: parallel lists ``iters`` / ``values`` (netket/utils/history/history.py)."""

    def __init__(self):
        self.iters, self.values = [], []

    def __len__(self):
        return len(self.iters)

    def __repr__(self):
        return f"History(n={len(self)})"


class HistoryDict(dict):
    """``{name: History}`` with the reference's ``push(values, step)`` (netket/utils/history/history_dict.py:147-180)."""

    def push(self, value, step):
        for k, v in value.items():
            h = self.setdefault(k, History())
            h.iters.append(int(step))
            h.values.append(v)
        return self


def _rel_err(err, scale):
    if scale == 0.0:
        return 0.0 if err == 0.0 else math.inf
    return err / scale


def _not_converged(stats, atol, rtol):
    s = stats.get_stats()
    err, scale = abs(float(s.error_of_mean)), abs(float(s.mean))
    if atol is not None and err > atol:
        return True
    if rtol is not None and _rel_err(err, scale) > rtol:
        return True
    return False


def _postfix(stats, atol, rtol):
    s = stats.get_stats()
    err, scale = abs(float(s.error_of_mean)), abs(float(s.mean))
    d = {"err": f"{err:.4g}"}
    if atol is not None:
        d["atol"] = f"{atol:.4g}"
    if rtol is not None:
        d["rel_err"] = f"{_rel_err(err, scale):.4g}"
        d["rtol"] = f"{rtol:.4g}"
    return d


def _require_metropolis(state, who):
    if not isinstance(state.sampler, MetropolisSampler):
        raise ValueError(f"{who} only works for MetropolisSampler.")


def expect_to_precision(state, op, *, atol=None, rtol=None, max_iter=10_000, max_lag=64, verbose=True):
    """Sample until the standard error of the mean of ``op`` meets ``atol`` and / or ``rtol``; returns the final
    :class:`OnlineStats`.  ``op`` may be a list / tuple / dict of operators: every entry is iterated until it has
    converged (the reference flattens an operator pytree the same way)."""
    if atol is None and rtol is None:
        raise ValueError("At least one of 'atol' or 'rtol' must be specified.")
    if atol is not None and atol <= 0:
        raise ValueError("atol must be > 0.")
    if rtol is not None and rtol <= 0:
        raise ValueError("rtol must be > 0.")
    _require_metropolis(state, "expect_to_precision")
    rank0 = world()[0] == 0

    if isinstance(op, dict):
        keys, leaves = list(op.keys()), list(op.values())
        rebuild = lambda out: dict(zip(keys, out))  # noqa: E731
    elif isinstance(op, (list, tuple)):
        leaves, rebuild = list(op), type(op)
    else:
        leaves, rebuild = [op], (lambda out: out[0])

    def accumulate(active, old, n_discard):
        fused = len(leaves) == 1
        if not fused:
            state.sample(n_discard_per_chain=n_discard)
        out = list(old)
        for i in active:
            le = state._sample_and_estimate(leaves[i], n_discard) if fused else state.local_estimators(leaves[i])
            out[i] = online_statistics(le, old[i], max_lag=max_lag, inplace=True)
        return out

    stats = accumulate(range(len(leaves)), [None] * len(leaves), None)
    active = [i for i in range(len(leaves)) if _not_converged(stats[i], atol, rtol)]
    it = 0
    with tqdm(total=max_iter, desc="Sampling", unit="iter", disable=not verbose or not rank0) as pbar:
        pbar.set_postfix(_postfix(stats[active[0] if active else 0], atol, rtol))
        try:
            while active and it < max_iter:
                stats = accumulate(active, stats, 0)
                active = [i for i in active if _not_converged(stats[i], atol, rtol)]
                pbar.set_postfix(_postfix(stats[active[0] if active else 0], atol, rtol))
                pbar.update(1)
                it += 1
        except KeyboardInterrupt:
            if rank0:
                pbar.write("  Early termination requested by user.")
        if verbose and rank0:
            if it >= max_iter:
                pbar.write("  Reached max_iter before target precision.")
            pbar.write(f"  [done] error = {abs(float(stats[0].get_stats().error_of_mean)):g}")
    return rebuild(stats)


def check_mc_convergence(state_, op, min_chain_length=50, plot=False, max_chain_length=500):
    """Diagnose whether ``sweep_size`` decorrelates successive samples of ``op``'s local estimator.  Works on a copy of
    the state with ``sweep_size = 1``; whenever the Geyer sequence runs out of lag window the sweep size is doubled and
    the accumulator re-indexed (``thin_acf_by_2`` + ``expand_max_lag``).  Returns ``(OnlineStats, HistoryDict)``."""
    _require_metropolis(state_, "check_mc_convergence")
    if plot:
        raise NotImplementedError("check_mc_convergence(plot=True): plotting is outside this package's scope")
    sampler = state_.sampler
    state = copy.copy(state_)
    del state_
    orig_sweep_size = sampler.sweep_size
    max_lag = 32
    sampler_state = state.sampler_state
    state._set_sampler_keep_state(sampler.replace(sweep_size=1), sampler_state)
    rank0 = world()[0] == 0

    chain_length = state.chain_length
    min_iters = math.ceil(min_chain_length / chain_length)
    max_iters = max_chain_length // chain_length

    stats = online_statistics(state._sample_and_estimate(op, None), max_lag=max_lag)
    it = 0
    hist = HistoryDict()
    with tqdm(desc="MC convergence", total=max_chain_length, initial=chain_length, unit=" spl/chain", leave=True,
              disable=not rank0) as pbar:
        while it < min_iters or acf_window_saturated(stats) or not tau_corr_reliable(stats):
            stats = online_statistics(state._sample_and_estimate(op, 0), stats, inplace=True)
            s = stats.get_stats()
            hist.push({"mean": stats.mean, "error_of_mean": s.error_of_mean, "variance": stats.variance, "R_hat": stats.R_hat,
                       "tau_corr_acf": stats.tau_corr_acf, "tau_corr_batch": stats.tau_corr_batch,
                       "sweep_size": state.sampler.sweep_size}, step=stats._n_samples_total // stats.n_chains)
            saturated = acf_window_saturated(stats)
            pbar.set_postfix({"sweep": state.sampler.sweep_size, "mean": f"{stats.mean:.4g}", "tau": f"{stats.tau_corr_acf:.3g}",
                              "R_hat": f"{stats.R_hat:.4f}", "sat": saturated})
            pbar.update(chain_length)
            if saturated:
                new_sweep = state.sampler.sweep_size * 2
                if rank0:
                    pbar.write(f"  [iter {it}] ACF window saturated — doubling sweep size to {new_sweep}")
                old_max_lag = stats.max_lag
                stats = expand_max_lag(thin_acf_by_2(stats), old_max_lag)
                # the reference re-installs the sampler state it captured before the loop (check_mc_convergence.py:206-207)
                state._set_sampler_keep_state(state.sampler.replace(sweep_size=new_sweep), sampler_state)
            it += 1
            if it >= max_iters:
                if rank0:
                    pbar.write(f"  Reached maximum chain length ({max_chain_length} samples/chain). Stopping.")
                break

    final_sweep = state.sampler.sweep_size
    tau_acf = stats.tau_corr_acf
    tau_mc_steps = tau_acf * final_sweep
    tau_sweeps = tau_mc_steps / orig_sweep_size
    good = tau_sweeps < 1.0
    if rank0:
        detail = f"  [tau_acf={tau_acf:.3g} x internal sweep_size={final_sweep}]" if final_sweep > 1 else ""
        print("\n---- MC Convergence Results ----\n"
              f"  Final statistics         : {stats}\n\n"
              f"  tau_corr (MC steps)      : {tau_mc_steps:.3g}{detail}\n\n"
              f"  MCState.sweep_size       : {orig_sweep_size}\n"
              f"  tau_corr (MC sweeps)     : {tau_sweeps:.3g}  "
              f"{'(< 1 sweep: good)' if good else '(>= 1 sweep: consider increasing sweep_size)'}\n\n"
              f"  Minimum sweep_size       : ~{2.0 * tau_mc_steps:.1f}  (= 2 x tau_corr in MC steps)\n"
              "--------------------------------")
    return stats, hist


def thermalise_mcmc(state, op, *, min_chain_length=10, max_chain_length=100, rhat_tol=1.05, decay=0.9, patience=1, verbose=True,
                    raise_on_failure=False):
    """Advance the chains of ``state`` (in place) until the exponentially windowed R-hat of ``op``'s local estimator has
    stayed below ``rhat_tol`` for ``patience`` consecutive batches.  Returns ``(OnlineStats, HistoryDict)``."""
    _require_metropolis(state, "thermalise_mcmc")
    if state.sampler.n_chains < 2:
        raise ValueError(f"thermalise_mcmc requires at least 2 chains to compute R̂. Current n_chains={state.sampler.n_chains}.")
    rank0 = world()[0] == 0
    chain_length = state.chain_length
    min_iters = max(0, math.ceil(min_chain_length / chain_length) - 1)
    max_iters = max(0, max_chain_length // chain_length - 1)

    stats = online_statistics(state._sample_and_estimate(op, 0), max_lag=0, decay=decay)
    hist = HistoryDict()

    def log():
        s = stats.get_stats()
        hist.push({"mean": stats.mean, "error_of_mean": s.error_of_mean, "variance": stats.variance, "R_hat": float(stats.R_hat)},
                  step=stats._n_samples_total // stats.n_chains)
        r = float(stats.R_hat)
        return r, (not math.isnan(r) and r < rhat_tol)

    rhat, good = log()
    consecutive_good = 1 if good else 0
    it = 0
    with tqdm(desc="MC thermalisation", total=max_chain_length, initial=chain_length, unit=" spl/chain", leave=True,
              disable=not (rank0 and verbose)) as pbar:
        while it < min_iters or consecutive_good < patience:
            if it >= max_iters:
                msg = (f"thermalise_mcmc reached the maximum chain length ({max_chain_length} samples/chain) without converging "
                       f"(R̂={rhat:.4f} >= {rhat_tol}). Consider increasing max_chain_length or sweep_size.")
                if rank0:
                    pbar.write(msg)
                if raise_on_failure:
                    raise RuntimeError(msg)
                warnings.warn(msg, UserWarning, stacklevel=2)
                break
            stats = online_statistics(state._sample_and_estimate(op, 0), stats, inplace=True)
            rhat, good = log()
            consecutive_good = consecutive_good + 1 if good else 0
            pbar.set_postfix({"mean": f"{stats.mean:.4g}", "R_hat": f"{rhat:.4f}", "patience": f"{consecutive_good}/{patience}"})
            pbar.update(chain_length)
            it += 1
    return stats, hist
