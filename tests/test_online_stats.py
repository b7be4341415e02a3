"""Streaming statistics on CPU: the oracle against the golden vectors (the reference's own source executed,
tests/golden/make_golden_online.py), and the host arithmetic of libnkb200 (nk_online_stats_finalize, no GPU needed)
against the oracle.  Tolerances: 1e-11 relative for float64 data; float32 data 1e-5 relative / 1e-6 absolute (the
reference forms the lag products in float32, the oracle and the kernels in float64)."""

import math
import os

import numpy as np
import pytest

from oracle import online_stats as oos

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "online_stats_vectors.npz"))
FIELDS = {"_chain_count": "count", "_chain_mean": "mean_c", "_chain_M2": "M2", "_cross_sum": "cross", "_m1_sum": "m1", "_m2_sum": "m2",
          "_pair_count": "pairs", "_chain_buf": "buf"}
SUMMARY = ("mean", "variance", "tau_corr", "tau_corr_batch", "tau_corr_acf", "R_hat", "error_of_mean", "n_samples")


def tolerances(dtype):
    return (1e-5, 1e-6) if dtype == np.float32 else (1e-11, 1e-11)


def check_against_golden(tag, e, rtol, atol):
    s = GOLD[f"{tag}_summary"]
    mine = [getattr(e, k) for k in SUMMARY] + [float(oos.acf_window_saturated(e)), float(oos.tau_corr_reliable(e))]
    np.testing.assert_allclose(mine, s, rtol=rtol, atol=atol, equal_nan=True, err_msg=tag)
    if e.acf is None:
        assert GOLD[f"{tag}_acf"].size == 0
    else:
        np.testing.assert_allclose(e.acf, GOLD[f"{tag}_acf"], rtol=rtol, atol=atol)
    assert int(GOLD[f"{tag}_buf_len"]) == e.buf_len
    if f"{tag}_chain_count" in GOLD:
        for k, v in FIELDS.items():
            np.testing.assert_allclose(getattr(e, v), GOLD[f"{tag}{k}"], rtol=rtol, atol=atol, err_msg=tag + k)


def golden_run(tag, update, on_step=None):
    """Feed the batches of golden case `tag` through `update(batch, est, decay, max_lag)`."""
    _, L, decay = GOLD[f"{tag}_cfg"]
    data, lens = GOLD[f"{tag}_data"], GOLD[f"{tag}_lens"]
    e, pos = None, 0
    for i, n in enumerate(lens):
        e = update(data[:, pos:pos + n], e, None if np.isnan(decay) else float(decay), int(L))
        pos += n
        if on_step:
            on_step(f"{tag}_s{i}", e)
    return e


@pytest.mark.parametrize("tag", [str(t) for t in GOLD["cases"]])
def test_oracle_matches_reference_vectors(tag):
    rtol, atol = tolerances(GOLD[f"{tag}_data"].dtype)
    e = golden_run(tag, lambda x, est, d, L: oos.online_statistics(x, est, decay=d, max_lag=L),
                   lambda t, est: check_against_golden(t, est, rtol, atol))
    if f"{tag}_more" in GOLD:  # the coarsening step of check_mc_convergence
        thin = oos.thin_acf_by_2(e)
        check_against_golden(f"{tag}_thin", thin, rtol, atol)
        wide = oos.expand_max_lag(thin, e.max_lag)
        check_against_golden(f"{tag}_expand", wide, rtol, atol)
        check_against_golden(f"{tag}_after", wide.update(GOLD[f"{tag}_more"]), rtol, atol)


def numpy_sums(e, gmean=None, mbar=None):
    """What nk_online_stats_summary produces for the chains of `e` (the kernel itself is tested in -m gpu)."""
    p0 = np.array([e.count.sum(), (e.count * e.mean_c).sum(), e.mean_c.sum()], dtype=np.float64)
    if gmean is None:
        return p0
    mu = e.mean_c.astype(np.float64)
    head = [e.M2.sum(), (e.count * (mu - gmean) ** 2).sum(), ((mu - mbar) ** 2).sum(), (e.M2 / np.maximum(e.count, 1.0)).sum()]
    n = np.maximum(e.pairs, 1.0)
    cov = (e.cross / n - (e.m1 / n) * (e.m2 / n)).sum(axis=0)
    return np.concatenate([head, cov])


def finalize_from_oracle_state(e):
    from netket_b200 import stats as nkstats

    p0 = numpy_sums(e)
    p1 = numpy_sums(e, p0[1] / p0[0], p0[2] / e.n_chains)
    return nkstats.online_finalize(p0, p1, e.n_chains, e.n_samples, e.max_lag)


@pytest.mark.parametrize("tag", [str(t) for t in GOLD["cases"]])
def test_host_finalize_matches_oracle(lib_built, tag):
    def on_step(t, e):
        s = finalize_from_oracle_state(e)
        want = [e.mean, e.error_of_mean, e.variance, e.tau_corr, e.R_hat, e.tau_corr_batch, e.tau_corr_acf,
                float(oos.acf_window_saturated(e)), float(oos.tau_corr_reliable(e))]
        np.testing.assert_allclose(s["out"], want, rtol=1e-10, atol=1e-12, equal_nan=True, err_msg=t)
        if e.acf is None:
            assert s["acf"] is None
        else:
            np.testing.assert_allclose(s["acf"], e.acf, rtol=1e-10, atol=1e-12)

    golden_run(tag, lambda x, est, d, L: oos.online_statistics(x, est, decay=d, max_lag=L), on_step)


def ar1(phi, n_chains=8, n_samples=500, seed=42):
    rng = np.random.default_rng(seed)
    data = np.zeros((n_chains, n_samples))
    data[:, 0] = rng.standard_normal(n_chains)
    for t in range(1, n_samples):
        data[:, t] = phi * data[:, t - 1] + np.sqrt(1 - phi ** 2) * rng.standard_normal(n_chains)
    return data


def test_oracle_window_diagnostics():
    """The cases of the reference's test/variational/test_check_mc_convergence.py:55-93."""
    assert oos.acf_window_saturated(oos.online_statistics(ar1(0.9), max_lag=8))
    assert not oos.acf_window_saturated(oos.online_statistics(ar1(0.9), max_lag=64))
    assert not oos.acf_window_saturated(oos.online_statistics(ar1(0.0), max_lag=32))
    assert not oos.tau_corr_reliable(oos.online_statistics(ar1(0.9), max_lag=8))
    assert oos.tau_corr_reliable(oos.online_statistics(ar1(0.0, n_samples=500), max_lag=32))
    assert not oos.tau_corr_reliable(oos.online_statistics(ar1(0.0, n_samples=5), max_lag=4))


def test_oracle_chunked_equals_oneshot():
    """test/stats/test_online_stats.py:169-204,374-392: feeding a series in pieces visits the same lag pairs."""
    x = ar1(0.7, n_chains=4, n_samples=240, seed=3) + 2.0
    one = oos.online_statistics(x, max_lag=16)
    e = None
    for lo, hi in [(0, 5), (5, 6), (6, 40), (40, 41), (41, 200), (200, 240)]:
        e = oos.online_statistics(x[:, lo:hi], e, max_lag=16)
    for f in FIELDS.values():
        np.testing.assert_allclose(getattr(e, f), getattr(one, f), rtol=1e-11, atol=1e-11, err_msg=f)
    assert math.isclose(e.tau_corr_acf, one.tau_corr_acf, rel_tol=1e-10)
    np.testing.assert_allclose(one.mean, x.mean(), rtol=1e-12)
    np.testing.assert_allclose(one.variance, x.var(), rtol=1e-12)


def test_oracle_errors_and_empty():
    e = oos.OnlineStats(3, max_lag=4)
    assert all(math.isnan(v) for v in e.get_stats().values())
    with pytest.raises(ValueError, match="Number of chains changed"):
        e.update(np.zeros((2, 5)))
    with pytest.raises(ValueError, match="must be >"):
        oos.expand_max_lag(e, 4)
    with pytest.raises(ValueError, match="thin"):
        oos.thin_acf_by_2(oos.OnlineStats(3, max_lag=1))


def test_host_finalize_edge_cases(lib_built):
    """Empty accumulator (all NaN, as OnlineStats().get_stats() -> Stats()), a constant series (no ACF: cov[0] = 0), one sample
    per chain, and the argument checks of the host entry point."""
    from netket_b200 import stats as nkstats

    s = nkstats.online_finalize([0.0, 0.0, 0.0], [0.0] * (4 + 9), 4, 0, 8)
    assert s["empty"] and all(math.isnan(v) for v in s["out"][:7]) and s["out"][7:] == [0.0, 0.0] and s["acf"] is None
    const = oos.online_statistics(np.full((3, 20), 2.5), max_lag=4)
    got = finalize_from_oracle_state(const)
    assert got["acf"] is None and const.acf is None and math.isnan(got["out"][6]) and got["out"][0] == 2.5 and got["out"][2] == 0.0
    assert math.isnan(got["out"][4]) and math.isnan(const.R_hat)              # W = 0: no R_hat
    assert math.isnan(got["out"][5]) and math.isnan(const.tau_corr_batch)     # variance = 0: no batch tau
    one = oos.online_statistics(np.random.default_rng(0).normal(size=(6, 1)), max_lag=4)
    got = finalize_from_oracle_state(one)
    want = [one.mean, one.error_of_mean, one.variance, one.tau_corr, one.R_hat, one.tau_corr_batch, one.tau_corr_acf]
    np.testing.assert_allclose(got["out"][:7], want, rtol=1e-12, equal_nan=True)
    with pytest.raises(Exception, match="bad arguments"):
        nkstats.online_finalize([1.0, 1.0, 1.0], [0.0] * 4, 4, 4, 5000)


@pytest.mark.parametrize("tag", [str(t) for t in GOLD["fft_cases"]])
def test_fft_statistics_oracle_matches_reference_vectors(tag):
    """The opt-in FFT variant of `statistics` (mc_stats.py:303-331, _autocorr.py:32-86): oracle only so far."""
    from oracle import stats as ostats

    data = GOLD[f"{tag}_data"]
    np.testing.assert_allclose(ostats.autocorr_1d(data[0]), GOLD[f"{tag}_acf0"], rtol=1e-9, atol=1e-12)
    r = ostats.statistics_fft(data)
    got = [r["mean"], r["error_of_mean"], r["variance"], r["tau_corr"], r["R_hat"], r["tau_corr_max"]]
    np.testing.assert_allclose(got, GOLD[f"{tag}_result"], rtol=1e-9, atol=1e-12, equal_nan=True)
