"""BASELINE.json configurations at (or near) full size: size-independent properties of the CUDA path, plus exact
expectation values on the small systems where the Hilbert space can be enumerated.

cfg-1 Ising1d L=20 fp64 local (16 chains x 63);  cfg-2 Heisenberg1d L=22 total_sz=0 exchange;  cfg-3 TFIM 10x10
2^16 chains fp32/fp64;  cfg-4 J1-J2 10x10 fp64 exchange 2^18 samples.
"""

import numpy as np
import pytest
import torch

import oracle
from oracle import estimators as oest
from oracle import graph as ograph
from oracle import hilbert as ohilbert
from oracle import operators as oops
from oracle import rbm as orbm
from oracle import sampler as osampler

pytestmark = pytest.mark.gpu


def _nk():
    import netket_b200 as nk

    return nk


def _var(N, alpha, dtype, std=0.01, seed=1234):
    W, b, a = orbm.init_params(N, alpha, seed=seed, std=std, dtype=dtype)
    var = {"params": {"Dense": {"kernel": torch.from_numpy(W).cuda(), "bias": torch.from_numpy(b).cuda()},
                      "visible_bias": torch.from_numpy(a).cuda()}}
    return (W.astype(np.float64), b.astype(np.float64), a.astype(np.float64)), var


def test_cfg1_ising1d_example_shape(cuda):
    """Examples/Ising1d: 16 chains, n_samples 1008 -> chain_length 63 (test_variational.py:182-187), fp64, E_loc vs oracle."""
    nk = _nk()
    g = nk.graph.Hypercube(20, 1)
    hi = nk.hilbert.Spin(0.5, 20)
    op = nk.operator.Ising(hi, g, h=1.0)
    (W, b, a), var = _var(20, 1, np.float64)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=16), nk.models.RBM(alpha=1), variables=var, n_samples=1000,
                        n_discard_per_chain=10, sampler_seed=15324)
    assert vs.n_samples == 1008 and vs.chain_length == 63
    st = vs.expect(op)
    samples = vs.samples.cpu().numpy()
    assert samples.shape == (16, 63, 20)
    e, _ = ograph.hypercube_edges(20, 1)
    ref = oest.local_estimators(samples, lambda x: oops.ising_conn_padded(x, e, 1.0, 1.0), W, b, a)
    np.testing.assert_allclose(vs.local_estimators(op).cpu().numpy(), ref, rtol=1e-11, atol=1e-11 * np.abs(ref).max())
    ost = oracle.stats.statistics(ref)
    np.testing.assert_allclose(st.mean, ost["mean"], rtol=1e-10)
    np.testing.assert_allclose(st.error_of_mean, ost["error_of_mean"], rtol=1e-7)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["ising", "heisenberg"])
def test_energy_within_sigma_of_exact_expectation(cuda, dtype, kind):
    """test/variational/test_variational.py:362-408 on the product-form kernels: MC <H> vs sum_sigma |psi|^2 E_loc (enumerated)."""
    nk = _nk()
    N = 12
    g = nk.graph.Hypercube(N, 1)
    e, c = ograph.hypercube_edges(N, 1)
    if kind == "ising":
        hi = nk.hilbert.Spin(0.5, N)
        op = nk.operator.Ising(hi, g, h=1.0)
        conn = lambda x: oops.ising_conn_padded(x, e, 1.0, 1.0)  # noqa: E731
        sa = nk.sampler.MetropolisLocal(hi, n_chains=2048)
        states = ohilbert.all_states(N)
    else:
        hi = nk.hilbert.Spin(0.5, N, total_sz=0)
        op = nk.operator.Heisenberg(hi, g)
        tables = oops.heisenberg_tables(e, c, 1.0, True)
        conn = lambda x: oops.local_operator_conn_padded(x, tables)  # noqa: E731
        sa = nk.sampler.MetropolisExchange(hi, graph=g, d_max=2, n_chains=2048)
        states = ohilbert.all_states(N, 0)
    (W, b, a), var = _var(N, 2, dtype, std=0.25)
    p = osampler.exact_distribution(W, b, a, states)
    exact = float(np.sum(p * oest.local_value_kernel(states, conn, W, b, a)))
    vs = nk.vqs.MCState(sa, nk.models.RBM(alpha=2, param_dtype=dtype), variables=var, n_samples=2048 * 64, n_discard_per_chain=64,
                        sampler_seed=7)
    st = vs.expect(op)
    assert st.error_of_mean < 0.1
    assert abs(st.mean - exact) < 5 * st.error_of_mean + (0 if dtype == np.float64 else 1e-4), (st, exact)
    assert st.R_hat < 1.1


def test_cfg3_full_size_fp32_vs_fp64(cuda):
    """2^16 chains: fp32 (tuned kernel) and fp64 (general kernel) agree on acceptance within 1% and on the energy within
    their error bars; fused E_loc equals the stand-alone estimator on the samples the launch produced."""
    nk = _nk()
    g = nk.graph.Hypercube(10, 2)
    hi = nk.hilbert.Spin(0.5, 100)
    op = nk.operator.Ising(hi, g, h=3.0)
    res = {}
    for dtype in (np.float32, np.float64):
        _, var = _var(100, 4, dtype)
        vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=1 << 16), nk.models.RBM(alpha=4, param_dtype=dtype), variables=var,
                            n_samples=(1 << 16) * 4, n_discard_per_chain=5, sampler_seed=15324)
        st = vs.expect(op)
        fused = vs.local_estimators(op)
        alone = vs._eloc_on_samples(op, vs.samples)
        tol = 2e-5 if dtype == np.float32 else 1e-11
        scale = float(fused.abs().max())
        assert float((fused - alone).abs().max()) <= tol * scale
        s = vs.samples
        assert s.dtype == torch.int8 and tuple(s.shape) == (1 << 16, 4, 100) and bool(((s == 1) | (s == -1)).all())
        res[dtype] = (st, vs.sampler_state.acceptance)
    (s32, a32), (s64, a64) = res[np.float32], res[np.float64]
    assert abs(a32 - a64) < 0.01 * a64
    assert abs(s32.mean - s64.mean) < 5 * np.hypot(s32.error_of_mean, s64.error_of_mean)


def test_cfg4_j1j2_full_size(cuda):
    """J1-J2 10x10, RBM alpha=4 fp64, MetropolisExchange, 2^18 samples: magnetisation conserved, fused == stand-alone,
    and the theta-form kernel run on a slice of the chains gives bit-identical samples."""
    nk = _nk()
    g = nk.graph.Hypercube(10, 2, max_neighbor_order=2)
    hi = nk.hilbert.Spin(0.5, 100, total_sz=0)
    op = nk.operator.Heisenberg(hi, g, J=[1.0, 0.5], sign_rule=[False, False])
    assert op.max_conn_size == 401
    _, var = _var(100, 4, np.float64)
    B, CL = 1 << 14, 16
    sa = nk.sampler.MetropolisExchange(hi, graph=g, d_max=1, n_chains=B)
    assert sa.rule.clusters.shape == (400, 2)
    model = nk.models.RBM(alpha=4)
    st0 = sa.init_state(model, var, seed=15324)
    samples, _, eloc, st = sa._launch(model, var, st0, CL, n_discard=2, operator=op)
    assert tuple(samples.shape) == (B, CL, 100) and tuple(eloc.shape) == (B, CL)
    assert bool((samples.to(torch.int32).sum(dim=-1) == 0).all())
    vs = nk.vqs.MCState(sa, model, variables=var, n_samples=B, seed=1)
    alone = vs._eloc_on_samples(op, samples)
    assert float((alone - eloc).abs().max()) <= 1e-10 * float(eloc.abs().max())
    assert 0.5 < st.acceptance <= 1.0
    # a 256-chain slice with the theta-form kernel: same Philox stream (global chain index) -> identical chains
    sa_small = nk.sampler.MetropolisExchange(hi, graph=g, d_max=1, n_chains=256)
    sts = sa_small.init_state(model, var, seed=15324).replace(σ=st0.σ[512:768].clone(), chain_offset=512)
    s_gen, _, e_gen, _ = sa_small._launch(model, var, sts, 4, n_discard=2, operator=op, path=1)
    assert torch.equal(s_gen, samples[512:768, :4])
    assert float((e_gen - eloc[512:768, :4]).abs().max()) <= 1e-10 * float(eloc.abs().max())


def test_cfg2_heisenberg1d(cuda):
    """Heisenberg1d L=22 total_sz=0 (sign rule on: bipartite), RBM alpha=2 fp64, MetropolisExchange 16 chains x 256 = 4096 samples."""
    nk = _nk()
    g = nk.graph.Hypercube(22, 1)
    hi = nk.hilbert.Spin(0.5, 22, total_sz=0)
    op = nk.operator.Heisenberg(hi, g)
    (W, b, a), var = _var(22, 2, np.float64)
    vs = nk.vqs.MCState(nk.sampler.MetropolisExchange(hi, graph=g, n_chains=16), nk.models.RBM(alpha=2), variables=var,
                        n_samples=4096, sampler_seed=15324)
    assert vs.chain_length == 256
    st = vs.expect(op)
    samples = vs.samples.cpu().numpy()
    assert np.all(samples.astype(int).sum(axis=-1) == 0)
    e, c = ograph.hypercube_edges(22, 1)
    tables = oops.heisenberg_tables(e, c, 1.0, True)
    ref = oest.local_estimators(samples, lambda x: oops.local_operator_conn_padded(x, tables), W, b, a)
    np.testing.assert_allclose(vs.local_estimators(op).cpu().numpy(), ref, rtol=1e-10, atol=1e-10 * np.abs(ref).max())
    np.testing.assert_allclose(st.mean, ref.mean(), rtol=1e-10)


def test_cfg3_streaming_statistics_full_size(cuda):
    """2^16 chains x 16 local energies per batch (cfg-3's E_loc shape): size-independent properties of the accumulator.
    Feeding a batch in four pieces equals feeding it at once; mean / variance equal the float64 moments of all samples
    (torch reference, 1e-12); split chains give the same pooled moments; the ACF starts at 1."""
    from netket_b200 import stats as nkstats

    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((2 ** 16, 16), dtype=torch.float64, device=cuda, generator=g).cumsum(dim=1) * 0.3 - 300.0
    y = torch.randn((2 ** 16, 16), dtype=torch.float64, device=cuda, generator=g) * 0.5 - 300.0
    one = nkstats.online_statistics(y, nkstats.online_statistics(x, max_lag=64))
    pieces = None
    for part in (x[:, :3], x[:, 3:4], x[:, 4:16], y[:, :9], y[:, 9:]):
        pieces = nkstats.online_statistics(part.contiguous(), pieces, max_lag=64)
    for f in ("_chain_count", "_chain_mean", "_chain_M2", "_cross_sum", "_m1_sum", "_m2_sum", "_pair_count", "_chain_buf"):
        torch.testing.assert_close(getattr(pieces, f), getattr(one, f), rtol=1e-11, atol=1e-9, msg=f)
    allx = torch.cat([x, y], dim=1)
    np.testing.assert_allclose(one.mean, allx.mean().item(), rtol=1e-13)
    np.testing.assert_allclose(one.variance, allx.var(unbiased=False).item(), rtol=1e-10)
    assert one.n_samples == 2 ** 21 and one.acf[0] == 1.0 and one.acf.shape == (65,)
    assert float(one._pair_count[0, 0]) == 32.0 and float(one._pair_count[7, 31]) == 1.0 and float(one._pair_count[7, 32]) == 0.0
    np.testing.assert_allclose(one.error_of_mean, (allx.mean(dim=1).var(unbiased=False) / 2 ** 16).sqrt().item(), rtol=1e-10)
    # float32 input: same state up to the rounding of the data
    one32 = nkstats.online_statistics(y.float(), nkstats.online_statistics(x.float(), max_lag=64))
    np.testing.assert_allclose(one32.mean, one.mean, rtol=1e-7)
    np.testing.assert_allclose(one32.variance, one.variance, rtol=1e-3)


def test_cfg3_qgt_full_size_properties(cuda):
    """The matrix-free S on cfg-3 (2^18 samples here, 40 500 parameters): symmetric (u.Sv == v.Su), positive (v.Sv >= shift |v|^2),
    linear, and constant shifts of log psi do not matter (S annihilates nothing but its shift on the all-equal direction of O)."""
    nk = _nk()
    from netket_b200.optimizer import QGTOnTheFly

    g = nk.graph.Hypercube(10, 2)
    hi = nk.hilbert.Spin(0.5, 100)
    H = nk.operator.Ising(hi, g, h=3.0)
    for dtype, tol in ((np.float32, 2e-4), (np.float64, 1e-10)):
        vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=2 ** 14), nk.models.RBM(alpha=4, param_dtype=dtype), n_samples=2 ** 18,
                            seed=3)
        vs.expect_and_grad(H)
        S = QGTOnTheFly(vs, diag_shift=0.01)
        n = S.shape[0]
        gen = torch.Generator(device="cuda").manual_seed(1)
        u = torch.randn(n, dtype=torch.float64, device=cuda, generator=gen)
        v = torch.randn(n, dtype=torch.float64, device=cuda, generator=gen)
        Su, Sv = S @ u, S @ v
        a, b = torch.dot(u, Sv).item(), torch.dot(v, Su).item()
        assert abs(a - b) <= tol * max(abs(a), abs(b), torch.dot(v, Sv).item())
        assert torch.dot(v, Sv).item() >= 0.01 * torch.dot(v, v).item() * (1 - 1e-6)
        lin = S @ (2.0 * u - 3.0 * v)
        ref = 2.0 * Su - 3.0 * Sv
        assert (lin - ref).abs().max().item() <= tol * ref.abs().max().item() * 10
