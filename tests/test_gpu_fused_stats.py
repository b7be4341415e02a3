"""In-kernel statistics (nk_sweep_t.stats_out, north_star item 4): `MCState.expect` without cached samples is ONE launch that
reduces the partial sums of netket/stats/mc_stats_old.py:87-196 in the sweep kernel's epilogue; the result must equal
`statistics(local_estimators)` (test/variational/test_variational.py:455-476) and the oracle's `statistics`."""

import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from oracle import rbm as orbm

pytestmark = pytest.mark.gpu


def _state(nk, L, dims, alpha, dtype, n_chains, chain_length, std=0.05, rule="local"):
    g = nk.graph.Hypercube(L, dims)
    N = g.n_nodes
    if rule == "local":
        hi = nk.hilbert.Spin(0.5, N)
        op = nk.operator.Ising(hi, g, h=3.0 if dims == 2 else 1.0)
        sa = nk.sampler.MetropolisLocal(hi, n_chains=n_chains)
    else:
        hi = nk.hilbert.Spin(0.5, N, total_sz=0)
        op = nk.operator.Heisenberg(hi, g)
        sa = nk.sampler.MetropolisExchange(hi, graph=g, n_chains=n_chains)
    W, b, a = orbm.init_params(N, alpha, seed=1234, std=std, dtype=dtype)
    var = {"params": {"Dense": {"kernel": torch.from_numpy(W).cuda(), "bias": torch.from_numpy(b).cuda()},
                      "visible_bias": torch.from_numpy(a).cuda()}}
    vs = nk.vqs.MCState(sa, nk.models.RBM(alpha=alpha, param_dtype=dtype), variables=var, n_samples=n_chains * chain_length,
                        n_discard_per_chain=3, sampler_seed=15324)
    return vs, op


@pytest.mark.parametrize("cfg", [
    ("local", 10, 2, 4, np.float32, 256, 16),   # tuned fp32 kernel: reduction inside sweep_fast_kernel
    ("local", 10, 2, 4, np.float32, 96, 63),    # odd chain length: half-chain and block sums drop the last sample
    ("local", 10, 2, 4, np.float32, 64, 130),   # l_block = 4
    ("local", 10, 2, 4, np.float64, 128, 16),   # general kernel + K6 pass over eloc_out
    ("local", 20, 1, 1, np.float64, 16, 63),    # cfg-1
    ("exchange", 12, 1, 2, np.float64, 64, 32),
    ("local", 10, 2, 4, np.float32, 1, 40),     # one chain: R_hat is NaN, block statistics decide
])
def test_fused_expect_equals_statistics_of_local_estimators(cuda, cfg):
    import netket_b200 as nk

    rule, L, dims, alpha, dtype, n_chains, cl = cfg
    vs, op = _state(nk, L, dims, alpha, dtype, n_chains, cl, rule=rule)
    for step in range(3):  # step 0: shift 0 (far from the mean), later steps: shift = the previous mean
        vs.reset()
        st = vs.expect(op)
        eloc = vs.local_estimators(op)
        two_pass = nk.stats.statistics(eloc)
        ref = oracle.stats.statistics(eloc.cpu().numpy().astype(np.float64))
        for k in ("mean", "variance", "error_of_mean", "tau_corr", "R_hat"):
            np.testing.assert_allclose(getattr(st, k), getattr(two_pass, k), rtol=2e-10, atol=1e-12, equal_nan=True, err_msg=f"{k} step {step}")
            np.testing.assert_allclose(getattr(st, k), ref[k], rtol=1e-9, atol=1e-11, equal_nan=True, err_msg=f"{k} step {step} (oracle)")


def test_stats_out_through_the_c_abi_any_shift(cuda):
    """nk_sweep with stats_out: the phase-1 sums of nk_stats_partial over eloc_out, for several shifts, on the tuned kernel and
    on the general kernel (forced path)."""
    import netket_b200 as nk
    from netket_b200 import _lib

    vs, op = _state(nk, 10, 2, 4, np.float32, 192, 20)
    sa = vs.sampler
    for path in (_lib.NK_PATH_AUTO, _lib.NK_PATH_PROD, _lib.NK_PATH_GENERIC):
        for shift in (0.0, -223.0, 1000.0):
            st0 = sa.init_state(vs.model, vs.variables, seed=5)
            _, _, eloc, _, part = sa._launch(vs.model, vs.variables, st0, 20, n_discard=2, operator=op, path=path, stats_shift=shift)
            x = eloc.cpu().numpy().astype(np.float64)
            d = x - shift
            m = d.mean(axis=1)
            halves = d.reshape(192, 2, 10).mean(axis=2)
            want = np.array([np.sum(d * d), m.sum(), (m * m).sum(), d.sum(), np.sum(d * d), halves.sum(), (halves ** 2).sum(), d.sum()])
            got = part.cpu().numpy()
            assert got[8] == 192
            np.testing.assert_allclose(got[:8], want, rtol=1e-12, err_msg=f"path {path} shift {shift}")


def test_two_temporaries_do_not_share_the_cache(cuda):
    """the E_loc cache was keyed by id(op); two temporaries with different h must give different energies."""
    import netket_b200 as nk

    vs, _ = _state(nk, 10, 2, 4, np.float32, 64, 8)
    g = nk.graph.Hypercube(10, 2)
    vs.sample()
    means = [vs.expect(nk.operator.Ising(vs.hilbert, g, h=h)).mean for h in (1.0, 2.0, 3.0)]
    assert abs(means[0] - means[1]) > 1.0 and abs(means[1] - means[2]) > 1.0, means


def test_user_registered_operator_goes_through_the_multimethods(cuda):
    """docs/advanced/custom-operators/local-estimators.ipynb cells 11-12: a user type registered on `local_estimators` (or on
    `get_local_kernel_arguments` + `get_local_kernel`) is what `vs.expect` runs; the built-in kernel is reachable as a function."""
    import netket_b200 as nk

    vs, op = _state(nk, 10, 2, 4, np.float64, 64, 8)

    class Shifted:  # H + c, evaluated through the generic (arguments, kernel) route
        def __init__(self, parent, c):
            self.parent, self.c, self.hilbert, self.dtype = parent, c, parent.hilbert, parent.dtype

    @nk.vqs.get_local_kernel_arguments.dispatch
    def _(vstate: nk.vqs.MCState, o: Shifted):
        return vstate.samples, o

    @nk.vqs.get_local_kernel.dispatch
    def _(vstate: nk.vqs.MCState, o: Shifted):
        return lambda logpsi, pars, sigma, a: nk.vqs.local_value_kernel_rbm(logpsi, pars, sigma, a.parent) + a.c

    vs.sample()
    base = vs.expect(op)
    sh = vs.expect(Shifted(op, 2.5))
    np.testing.assert_allclose(sh.mean, base.mean + 2.5, rtol=1e-12)
    np.testing.assert_allclose(sh.variance, base.variance, rtol=1e-9)

    class Twice:  # registered directly on local_estimators, one overload per chunk_size kind as the reference's message asks
        def __init__(self, parent):
            self.parent, self.hilbert = parent, parent.hilbert

    @nk.vqs.local_estimators.dispatch
    def _(vstate: nk.vqs.MCState, o: Twice, chunk_size: None):
        return nk.vqs.LocalEstimators(2.0 * vstate.local_estimators(o.parent))

    np.testing.assert_allclose(vs.expect(Twice(op)).mean, 2.0 * base.mean, rtol=1e-12)
    with pytest.raises(NotImplementedError):
        vs.expect(object())


# ---------------------------------------------------------------------------------------------- FFT variant of `statistics`
import os  # noqa: E402

_GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "online_stats_vectors.npz"))


@pytest.mark.parametrize("tag", [str(t) for t in _GOLD["fft_cases"]])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fft_statistics_match_reference_vectors(cuda, tag, dtype):
    """netket/stats/mc_stats.py:303-331 + _autocorr.py:40-86 executed from the reference's source (make_golden_online.py):
    mean, error_of_mean, variance, tau_corr (Sokal window, chain average), R_hat, tau_corr_max."""
    import netket_b200 as nk

    data = _GOLD[f"{tag}_data"].astype(dtype)
    want = _GOLD[f"{tag}_result"]
    st = nk.stats.statistics_fft(torch.from_numpy(data).cuda())
    got = np.array([st.mean, st.error_of_mean, st.variance, st.tau_corr, st.R_hat, st.tau_corr_max])
    if dtype == np.float64:
        np.testing.assert_allclose(got, want, rtol=1e-10, equal_nan=True)
    else:
        ref = oracle.stats.statistics_fft(data.astype(np.float64))  # the same rounded inputs, in double
        want32 = np.array([ref[k] for k in ("mean", "error_of_mean", "variance", "tau_corr", "R_hat", "tau_corr_max")])
        np.testing.assert_allclose(got, want32, rtol=1e-9, equal_nan=True)


def test_statistics_flag_switches_to_the_fft_variant(cuda):
    import netket_b200 as nk

    rs = np.random.default_rng(3)
    x = torch.from_numpy(rs.normal(size=(8, 64)).cumsum(axis=1) * 0.3).cuda()
    plain = nk.stats.statistics(x)
    assert np.isnan(plain.tau_corr_max)
    nk.config.update("netket_experimental_fft_autocorrelation", True)
    try:
        fft = nk.stats.statistics(x)
    finally:
        nk.config.update("netket_experimental_fft_autocorrelation", False)
    ref = oracle.stats.statistics_fft(x.cpu().numpy())
    for k in ("mean", "error_of_mean", "variance", "tau_corr", "R_hat", "tau_corr_max"):
        np.testing.assert_allclose(getattr(fft, k), ref[k], rtol=1e-10, err_msg=k)
    # a constant chain has no autocorrelation function: tau is NaN, as jnp's 0 / 0
    c = torch.ones((4, 16), dtype=torch.float64, device="cuda")
    assert np.isnan(nk.stats.statistics_fft(c).tau_corr)


@pytest.mark.parametrize("std", [0.5, 0.9])
def test_optimistic_launch_repeats_with_the_full_chain_of_kernels_for_large_weights(cuda, std):
    """NK_SWEEP_NO_HANDOVER: `expect` enqueues only the tuned fp32 kernel; weights beyond its range (std 0.5: max|W| ~ 2,
    std 0.9: ~ 3.8) make it raise NaN in the partial sums, and MCState repeats the launch with the hand-over kernels."""
    import netket_b200 as nk
    from netket_b200 import _lib

    vs, op = _state(nk, 10, 2, 4, np.float32, 128, 8, std=std)
    st0 = vs.sampler_state
    n0 = _lib.lib().nk_launch_count()
    stats = vs.expect(op)
    assert np.isfinite(stats.mean) and np.isfinite(stats.variance)
    eloc = vs.local_estimators(op)
    ref = oracle.stats.statistics(eloc.cpu().numpy().astype(np.float64))
    np.testing.assert_allclose(stats.mean, ref["mean"], rtol=1e-10)
    np.testing.assert_allclose(stats.variance, ref["variance"], rtol=1e-9)
    # the same chains as a plain (non-optimistic) launch from the same state
    sa = vs.sampler
    s2, _, e2, _ = sa._launch(vs.model, vs.variables, st0.replace(n_steps_proc=0, n_accepted_proc=torch.zeros_like(st0.n_accepted_proc)),
                              8, n_discard=3, operator=op)
    assert torch.equal(s2, vs.samples) and torch.equal(e2, eloc)
    assert _lib.lib().nk_launch_count() - n0 > 8  # two rounds of launches happened


def test_expect_is_at_most_four_of_our_kernels_per_step(cuda):
    """launches per `vs.reset(); vs.expect(H)` on the headline path: theta prep + theta GEMM + fused sweep."""
    import netket_b200 as nk
    from netket_b200 import _lib

    vs, op = _state(nk, 10, 2, 4, np.float32, 256, 16, std=0.01)
    vs.expect(op)
    vs.reset()
    n0 = _lib.lib().nk_launch_count()
    vs.expect(op)
    assert _lib.lib().nk_launch_count() - n0 <= 4
