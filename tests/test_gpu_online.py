"""Streaming statistics and their callers on the GPU.

The CUDA accumulator (nk_online_stats_update / _summary, through netket_b200.stats.OnlineStats) is compared with the
oracle (oracle/online_stats.py, itself pinned to the reference's source by tests/test_online_stats.py) and with the golden
vectors directly; then the reference's own test ideas (test/stats/test_online_stats.py,
test/variational/test_check_mc_convergence.py) run against it.  Tolerance: 1e-10 relative for float64 data (sums in another
order), 1e-5 / 1e-6 for float32 data (the reference forms lag products in float32)."""

import math
import warnings

import numpy as np
import pytest
import torch

from oracle import online_stats as oos
from test_online_stats import FIELDS, GOLD, SUMMARY, ar1, golden_run

pytestmark = pytest.mark.gpu


def dev_fields(e):
    return {k: getattr(e, k).cpu().numpy() for k in FIELDS}


def assert_matches_oracle(e, ref, rtol=1e-10, atol=1e-11, msg=""):
    for k, v in FIELDS.items():
        np.testing.assert_allclose(getattr(e, k).cpu().numpy(), np.asarray(getattr(ref, v), dtype=np.float64), rtol=rtol, atol=atol,
                                   err_msg=f"{msg}{k}")
    assert e._buf_len == ref.buf_len and e._n_samples_total == ref.n_samples
    got = [e.mean, e.variance, e.tau_corr, e.tau_corr_batch, e.tau_corr_acf, e.R_hat, e.error_of_mean, e.n_samples]
    want = [getattr(ref, k) for k in SUMMARY]
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, equal_nan=True, err_msg=msg)
    if ref.acf is None:
        assert e.acf is None
    else:
        np.testing.assert_allclose(e.acf, ref.acf, rtol=rtol, atol=max(atol, 1e-10))


@pytest.mark.parametrize("tag", [str(t) for t in GOLD["cases"]])
def test_cuda_accumulator_matches_reference_vectors(cuda, tag):
    from netket_b200 import stats as nkstats

    f32 = GOLD[f"{tag}_data"].dtype == np.float32
    rtol, atol = (1e-5, 1e-6) if f32 else (1e-10, 1e-11)

    def check(t, e):
        s = GOLD[f"{t}_summary"]
        got = [e.mean, e.variance, e.tau_corr, e.tau_corr_batch, e.tau_corr_acf, e.R_hat, e.error_of_mean, e.n_samples,
               float(nkstats.acf_window_saturated(e)), float(nkstats.tau_corr_reliable(e))]
        np.testing.assert_allclose(got, s, rtol=rtol, atol=atol, equal_nan=True, err_msg=t)
        if e.acf is None:
            assert GOLD[f"{t}_acf"].size == 0
        else:
            np.testing.assert_allclose(e.acf, GOLD[f"{t}_acf"], rtol=rtol, atol=max(atol, 1e-10))
        if f"{t}_chain_count" in GOLD:
            for k in FIELDS:
                np.testing.assert_allclose(getattr(e, k).cpu().numpy(), GOLD[f"{t}{k}"], rtol=rtol, atol=atol, err_msg=t + k)
        st = e.get_stats()
        assert st.mean == e.mean and st.variance == e.variance

    def update(x, est, decay, L):
        return nkstats.online_statistics(torch.from_numpy(np.ascontiguousarray(x)).to(cuda), est, decay=decay, max_lag=L)

    e = golden_run(tag, update, check)
    if f"{tag}_more" in GOLD:
        thin = nkstats.thin_acf_by_2(e)
        check(f"{tag}_thin", thin)
        wide = nkstats.expand_max_lag(thin, e.max_lag)
        check(f"{tag}_expand", wide)
        check(f"{tag}_after", wide.update(torch.from_numpy(GOLD[f"{tag}_more"]).to(cuda)))


@pytest.mark.parametrize("n_chains,max_lag,lens,dtype,decay", [
    (300, 64, [16, 16, 1, 200, 64, 65, 63], np.float64, None),   # batches below / at / above the 64-sample chunk of the kernel
    (37, 200, [50, 500, 3, 130], np.float64, None),               # several lag rounds per lane, buffer longer than a chunk
    (1, 31, [400], np.float64, None),                             # one chain: error from the ACF estimate
    (64, 33, [10, 10, 10], np.float32, 0.85),                     # float32 data with decay
    (5000, 8, [4, 4], np.float64, None),                          # more chains than resident warps
    (9, 1000, [700, 700], np.float64, None),                      # fewer warps per block (shared memory per warp)
])
def test_cuda_accumulator_matches_oracle(cuda, n_chains, max_lag, lens, dtype, decay):
    from netket_b200 import stats as nkstats

    x = (ar1(0.6, n_chains=n_chains, n_samples=sum(lens), seed=n_chains) * 1.7 - 4.0).astype(dtype)
    e, ref, pos = None, None, 0
    rtol, atol = (2e-5, 2e-5) if dtype == np.float32 else (1e-10, 1e-10)
    for n in lens:
        batch = np.ascontiguousarray(x[:, pos:pos + n])
        pos += n
        e = nkstats.online_statistics(torch.from_numpy(batch).to(cuda), e, decay=decay, max_lag=max_lag)
        ref = oos.online_statistics(batch.astype(np.float64), ref, decay=decay, max_lag=max_lag)
        assert_matches_oracle(e, ref, rtol, atol, msg=f"after {pos} ")


def test_update_is_functional_unless_inplace(cuda):
    """test/stats/test_online_stats.py:524-537: update returns a new accumulator and leaves the old one untouched."""
    from netket_b200 import stats as nkstats

    x = torch.from_numpy(ar1(0.5, n_chains=6, n_samples=40)).to(cuda)
    a = nkstats.online_statistics(x[:, :20].contiguous(), max_lag=8)
    before, mean_before = dev_fields(a), a.mean
    b = a.update(x[:, 20:].contiguous())
    after = dev_fields(a)
    for k in FIELDS:
        np.testing.assert_array_equal(before[k], after[k])
    assert a.mean == mean_before and a.n_samples == 120 and b.n_samples == 240 and b is not a
    c = a.update(x[:, 20:].contiguous(), inplace=True)
    assert c is a and a.n_samples == 240
    for k in FIELDS:
        np.testing.assert_array_equal(dev_fields(a)[k], dev_fields(b)[k])
    one = nkstats.OnlineStats.from_data(x, max_lag=8)  # chunked == one shot (:169-204, 374-392)
    for k in FIELDS:
        np.testing.assert_allclose(dev_fields(one)[k], dev_fields(b)[k], rtol=1e-11, atol=1e-11)


def test_reference_behaviour_cases(cuda):
    """Ideas of test/stats/test_online_stats.py: batch equivalence, diverging chains, decay, max_lag = 0, 1-D input,
    logging protocol, error handling."""
    from netket_b200 import stats as nkstats

    rs = np.random.default_rng(0)
    x = rs.normal(size=(16, 200)) * 2.0 + 1.5
    e = nkstats.online_statistics(x)
    np.testing.assert_allclose(e.mean, x.mean(), rtol=1e-12)              # :45-55
    np.testing.assert_allclose(e.variance, x.var(), rtol=1e-12)
    batch = nkstats.statistics(x)                                         # :150-167 (iid: both estimates agree roughly)
    assert 0.5 < e.error_of_mean / batch.error_of_mean < 2.0
    assert e.n_chains == 16 and e.n_samples == 3200 and e.acf.shape == (65,) and e.acf[0] == 1.0   # :307-314, 394-407
    assert 0.98 < e.R_hat < 1.02                                          # :234-241
    assert e.tau_corr == e.tau_corr_acf                                   # :456-468
    assert set(e.to_dict()) == {"Mean", "Variance", "Sigma", "R_hat", "TauCorr"} and e.to_compound()[0] == "Mean"   # :470-497
    shifted = x + np.arange(16)[:, None] * 5.0                            # :217-232
    assert nkstats.online_statistics(shifted).R_hat > 1.5
    tau = nkstats.online_statistics(ar1(0.9, n_chains=32, n_samples=4000), max_lag=128).tau_corr_acf   # :351-372
    assert 12.0 < tau < 28.0                                              # (1 + phi) / (1 - phi) = 19
    old, new = rs.normal(size=(8, 50)) + 10.0, rs.normal(size=(8, 50)) - 10.0   # :243-263
    d = nkstats.online_statistics(old, decay=0.1, max_lag=0)
    for _ in range(5):
        d = nkstats.online_statistics(new, d)
    assert abs(d.mean + 10.0) < 0.5 and d.decay == 0.1
    z = nkstats.online_statistics(x, max_lag=0)                           # :735-760
    assert z.acf is None and math.isnan(z.tau_corr_acf) and z.tau_corr == z.tau_corr_batch and z._cross_sum.shape == (16, 0)
    z = nkstats.expand_max_lag(z, 4)                                      # :686-701
    assert z._cross_sum.shape == (16, 5) and z._chain_buf.shape == (16, 4) and float(z._pair_count.sum()) == 0.0
    z = z.update(torch.from_numpy(x).to(cuda))
    assert z.acf is not None
    one = nkstats.online_statistics(x[0])                                 # :286-305
    assert one.n_chains == 1 and math.isnan(one.R_hat) and math.isnan(one.tau_corr_batch) and one.error_of_mean > 0
    empty = nkstats.OnlineStats(4, max_lag=8)                             # :510-522
    assert math.isnan(empty.get_stats().mean) and math.isnan(empty.mean)
    with pytest.raises(ValueError, match="Number of chains changed"):     # :499-508
        e.update(torch.zeros((3, 5), dtype=torch.float64, device=cuda))
    with pytest.raises(ValueError, match="must be >"):                    # :720-733
        nkstats.expand_max_lag(e, 64)
    with pytest.raises(ValueError, match="thin"):
        nkstats.thin_acf_by_2(nkstats.online_statistics(x, max_lag=1))
    with pytest.raises(TypeError, match="float32 / float64"):
        nkstats.online_statistics(torch.zeros((2, 4), dtype=torch.int32, device=cuda))
    with pytest.raises(ValueError, match="max_lag"):
        nkstats.OnlineStats(4, max_lag=5000)


def test_window_diagnostics(cuda):
    """test/variational/test_check_mc_convergence.py:55-93."""
    from netket_b200.stats import acf_window_saturated, online_statistics, tau_corr_reliable

    assert acf_window_saturated(online_statistics(ar1(0.9), max_lag=8))
    assert not acf_window_saturated(online_statistics(ar1(0.9), max_lag=64))
    assert not acf_window_saturated(online_statistics(ar1(0.0), max_lag=32))
    assert not tau_corr_reliable(online_statistics(ar1(0.9), max_lag=8))
    assert tau_corr_reliable(online_statistics(ar1(0.0, n_samples=500), max_lag=32))
    assert not tau_corr_reliable(online_statistics(ar1(0.0, n_samples=5), max_lag=4))


def small_state(nk, n_chains=16, n_samples=16, dtype=np.float64, seed=0, **kw):
    hi = nk.hilbert.Spin(0.5, 4)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=n_chains, **kw), nk.models.RBM(alpha=1, param_dtype=dtype),
                        n_samples=n_samples, seed=seed)
    return vs, nk.operator.Ising(hi, nk.graph.Chain(4), h=1.0)


def test_expect_to_precision(cuda):
    """test/variational/test_check_mc_convergence.py:119-183."""
    import netket_b200 as nk

    vs, H = small_state(nk)
    with pytest.raises(ValueError, match="atol.*rtol"):
        vs.expect_to_precision(H, verbose=False)
    with pytest.raises(ValueError, match="atol must be > 0"):
        vs.expect_to_precision(H, atol=-1.0, verbose=False)
    s = vs.expect_to_precision(H, atol=0.5, max_iter=500, verbose=False).get_stats()
    assert math.isfinite(s.mean) and s.error_of_mean <= 0.5
    s = vs.expect_to_precision(H, rtol=0.1, max_iter=500, verbose=False).get_stats()
    assert math.isfinite(s.mean) and s.error_of_mean / abs(s.mean) <= 0.1
    acc = vs.expect_to_precision(H, atol=1e-10, max_iter=3, verbose=False)
    assert math.isfinite(acc.mean) and acc.n_samples == 16 * 4          # the first batch + max_iter more
    both = vs.expect_to_precision({"H": H, "H2": nk.operator.Ising(vs.hilbert, nk.graph.Chain(4), h=0.5)}, atol=0.3, max_iter=500,
                                  verbose=False)
    assert set(both) == {"H", "H2"} and all(v.error_of_mean <= 0.3 for v in both.values())


def test_expect_to_precision_agrees_with_exact_energy(cuda):
    """The streamed estimate converges to <H> of the RBM state (enumeration over the 2^8 configurations)."""
    import netket_b200 as nk
    from oracle import estimators as oest, hilbert as ohilbert, operators as oops, sampler as osampler

    hi = nk.hilbert.Spin(0.5, 8)
    g = nk.graph.Chain(8)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=1024), nk.models.RBM(alpha=2, param_dtype=np.float64), n_samples=1024 * 8,
                        seed=5)
    H = nk.operator.Ising(hi, g, h=1.0)
    acc = vs.expect_to_precision(H, atol=2e-3, max_iter=2000, verbose=False)
    W, b, a = (t.cpu().numpy() for t in nk.models.RBM.unpack(vs.variables))
    states = ohilbert.all_states(8)
    p = osampler.exact_distribution(W, b, a, states)
    edges = np.asarray(g.edges(), dtype=np.int32)
    eloc = oest.local_value_kernel(states, lambda x: oops.ising_conn_padded(x, edges, 1.0, 1.0), W, b, a)
    exact = float((p * eloc).sum())
    assert acc.error_of_mean <= 2e-3
    assert abs(acc.mean - exact) < 6 * acc.error_of_mean + 1e-9, (acc.mean, exact, acc.error_of_mean)


def test_streaming_loop_equals_manual_loop(cuda):
    """expect_to_precision is `sample(n_discard_per_chain=0); local_estimators(op)` + accumulator updates: a manual loop over
    the public API on an identically seeded state gives the same numbers (fused and unfused launches draw the same chains)."""
    import netket_b200 as nk
    from netket_b200.stats import online_statistics

    vs1, H = small_state(nk, n_chains=32, n_samples=128, seed=11)
    vs2, _ = small_state(nk, n_chains=32, n_samples=128, seed=11)
    acc = vs1.expect_to_precision(H, atol=1e-12, max_iter=4, max_lag=16, verbose=False)
    vs2.sample()
    man = online_statistics(vs2.local_estimators(H), max_lag=16)
    for _ in range(4):
        vs2.sample(n_discard_per_chain=0)
        man = online_statistics(vs2.local_estimators(H), man)
    for k in FIELDS:
        np.testing.assert_allclose(getattr(acc, k).cpu().numpy(), getattr(man, k).cpu().numpy(), rtol=1e-12, atol=1e-12, err_msg=k)
    assert torch.equal(vs1.sampler_state.σ, vs2.sampler_state.σ)


def test_check_mc_convergence(cuda):
    """test/variational/test_check_mc_convergence.py:97-104 + the state is left untouched."""
    import netket_b200 as nk

    vs, H = small_state(nk)
    sigma0, sweep0 = vs.sampler_state.σ.clone(), vs.sampler.sweep_size
    stats, hist = vs.check_mc_convergence(H, max_chain_length=50)
    assert math.isfinite(stats.mean) and math.isfinite(stats.variance)
    assert "mean" in hist and "tau_corr_acf" in hist and len(hist["mean"]) >= 1
    assert hist["sweep_size"].values[0] == 1
    assert torch.equal(vs.sampler_state.σ, sigma0) and vs.sampler.sweep_size == sweep0
    # a long run on many chains resolves tau of the single-spin-flip chain: reliable and not saturated at the end
    vs, H = small_state(nk, n_chains=256, n_samples=256 * 8)
    stats, hist = vs.check_mc_convergence(H, min_chain_length=50, max_chain_length=4000)
    from netket_b200.stats import acf_window_saturated, tau_corr_reliable

    assert not acf_window_saturated(stats) and tau_corr_reliable(stats) and stats.tau_corr_acf >= 1.0
    with pytest.raises(NotImplementedError):
        vs.check_mc_convergence(H, plot=True)


def test_thermalise(cuda):
    import netket_b200 as nk

    vs, H = small_state(nk, n_chains=64, n_samples=64 * 4)
    sigma0 = vs.sampler_state.σ.clone()
    stats, hist = vs.thermalise(H, min_chain_length=20, max_chain_length=400, verbose=False)
    assert stats.max_lag == 0 and stats.decay == 0.9 and stats.R_hat < 1.05
    assert hist["R_hat"].iters[0] == 4 and hist["R_hat"].iters[-1] == stats._n_samples_total // stats.n_chains
    assert len(hist["R_hat"]) >= 5                        # min_chain_length = 20 at 4 samples per batch
    assert not torch.equal(vs.sampler_state.σ, sigma0)    # the chains were advanced in place
    # an unreachable tolerance: warning, or an error on request
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        vs.thermalise(H, rhat_tol=0.5, max_chain_length=20, verbose=False)
    assert any("without converging" in str(x.message) for x in w)
    with pytest.raises(RuntimeError, match="without converging"):
        vs.thermalise(H, rhat_tol=0.5, max_chain_length=20, verbose=False, raise_on_failure=True)
    one, H1 = small_state(nk, n_chains=1, n_samples=4)
    with pytest.raises(ValueError, match="at least 2 chains"):
        one.thermalise(H1, verbose=False)


def test_online_stats_batch_delta_method(cuda):
    """OnlineStatsBatch (accumulator_batch.py:71-225): K channels accumulated by the streaming kernel, a scalar and an
    array-valued combinator, errors by the delta method on the covariance of the chain means - against the same formulas in
    NumPy (Cov = D D^T / n_chains^2, Var f = J Cov J^T) with the analytic Jacobians."""
    import netket_b200 as nk

    rs = np.random.default_rng(11)
    n_chains, K = 24, 3
    batches = [rs.normal(size=(n_chains, L, K)) * [1.0, 0.5, 2.0] + [1.0, -2.0, 0.5] for L in (7, 16, 5)]
    f = lambda X: X[2] - X[0] * X[1]  # noqa: E731  (a connected correlator)
    acc = None
    for b in batches:
        acc = nk.stats.online_statistics_batch(torch.from_numpy(b).cuda(), f, acc, max_lag=8)
    allx = np.concatenate(batches, axis=1)
    X = allx.mean(axis=(0, 1))
    cm = allx.mean(axis=1)  # chain means (n_chains, K)
    D = (cm - X).T
    Cov = D @ D.T / n_chains ** 2
    J = np.array([-X[1], -X[0], 1.0])
    st = acc.get_stats()
    np.testing.assert_allclose(st.mean, X[2] - X[0] * X[1], rtol=1e-12)
    np.testing.assert_allclose(st.error_of_mean, np.sqrt(J @ Cov @ J), rtol=1e-10)
    assert acc.n_samples == n_chains * 28 and acc.n_chains == n_chains
    g = lambda X: torch.stack([X[0] + X[1], X[0] * X[2]])  # noqa: E731
    accg = nk.stats.OnlineStatsBatch.from_data(torch.from_numpy(allx).cuda(), g, max_lag=8)
    sb = accg.get_stats()
    Jg = np.array([[1.0, 1.0, 0.0], [X[2], 0.0, X[0]]])
    assert sb.shape == (2,)
    np.testing.assert_allclose(sb.mean.cpu().numpy(), [X[0] + X[1], X[0] * X[2]], rtol=1e-12)
    np.testing.assert_allclose(sb.error_of_mean.cpu().numpy(), np.sqrt(np.einsum("ik,kl,il->i", Jg, Cov, Jg)), rtol=1e-10)
    one = nk.stats.OnlineStatsBatch.from_data(torch.from_numpy(allx[:1]).cuda(), f, max_lag=4)
    assert np.isnan(one.get_stats().error_of_mean)  # fewer than two chains: no covariance (accumulator_batch.py:199-200)
    with pytest.raises(ValueError, match="3D"):
        nk.stats.OnlineStatsBatch.from_data(torch.zeros((4, 4)).cuda(), f)


def test_local_estimator_containers(cuda):
    """LocalEstimators / LocalEstimatorsBatch (netket/_src/stats/local_estimators.py:47-246): `to_stats`, `to_online_stats`,
    `accumulate`, attribute forwarding; the batch container's one-shot delta-method statistics against NumPy."""
    import netket_b200 as nk

    vs, H = small_state(nk, n_samples=64)
    le = vs.local_estimators(H)
    assert isinstance(le, nk.stats.LocalEstimators) and tuple(le.shape) == tuple(le.data.shape) and le.dtype == le.data.dtype
    with pytest.raises(AttributeError, match="no attribute 'foo'"):
        le.foo
    st = le.to_stats()
    np.testing.assert_allclose(st.mean, vs.expect(H).mean, rtol=1e-12)
    acc = le.accumulate(max_lag=4)
    assert isinstance(acc, nk.stats.OnlineStats)
    np.testing.assert_allclose(acc.mean, st.mean, rtol=1e-10)
    vs.sample(n_discard_per_chain=0)
    le2 = vs.local_estimators(H)
    acc2 = le2.accumulate(acc)
    both = torch.cat([le.data, le2.data], dim=1)
    np.testing.assert_allclose(acc2.mean, both.mean().item(), rtol=1e-10)
    assert acc2.n_samples == both.numel()
    # K channels: the variance observable  <E^2> - <E>^2  and an array-valued combinator
    e = le.data.to(torch.float64)
    comb = lambda mu: mu[1] - mu[0] ** 2  # noqa: E731
    leb = nk.stats.LocalEstimatorsBatch(torch.stack([e, e * e], dim=-1), comb)
    assert leb.n_channels == 2
    with pytest.raises(AttributeError, match="use le.data.mean"):
        leb.mean
    x = leb.data.cpu().numpy()
    cm = x.mean(axis=1)
    X = cm.mean(axis=0)
    D = (cm - X).T
    Cov = D @ D.T / cm.shape[0] ** 2
    J = np.array([-2.0 * X[0], 1.0])
    sb = leb.to_stats()
    np.testing.assert_allclose(sb.mean, X[1] - X[0] ** 2, rtol=1e-11)
    np.testing.assert_allclose(sb.error_of_mean, np.sqrt(J @ Cov @ J), rtol=1e-8)
    accb = leb.accumulate(max_lag=4)
    assert isinstance(accb, nk.stats.OnlineStatsBatch)
    np.testing.assert_allclose(accb.get_stats().mean, sb.mean, rtol=1e-10)
    np.testing.assert_allclose(accb.get_stats().error_of_mean, sb.error_of_mean, rtol=1e-8)
    e2 = le2.data.to(torch.float64)
    accb2 = nk.stats.LocalEstimatorsBatch(torch.stack([e2, e2 * e2], dim=-1), comb).accumulate(accb)
    assert accb2.n_samples == 2 * e.numel()
    arr = nk.stats.LocalEstimatorsBatch(leb.data, lambda mu: torch.stack([mu[0], mu[1] - mu[0] ** 2])).to_stats()
    assert arr.shape == (2,)
    np.testing.assert_allclose(arr.mean.cpu().numpy(), [X[0], X[1] - X[0] ** 2], rtol=1e-11)
    one = nk.stats.LocalEstimatorsBatch(leb.data[:1], comb).to_stats()  # a single chain: covariance of the samples
    flat = x[:1].reshape(-1, 2)
    X1 = flat.mean(axis=0)
    C1 = (flat - X1).T @ (flat - X1) / flat.shape[0] ** 2
    J1 = np.array([-2.0 * X1[0], 1.0])
    np.testing.assert_allclose(one.error_of_mean, np.sqrt(J1 @ C1 @ J1), rtol=1e-8)
