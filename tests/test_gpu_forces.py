"""expect_and_forces / expect_and_grad against the oracle: the RBM's closed-form log-derivatives
contracted with the centred local energies (netket/vqs/mc/mc_state/expect_forces.py:69-112, vqs/mc/common.py:103-118).
Tolerances: fp64 1e-10 relative to max|F| (double atomics: summation order differs), fp32 2e-5."""

import numpy as np
import pytest
import torch

from oracle import forces as oforces
from oracle import graph as ograph
from oracle import rbm as orbm

pytestmark = pytest.mark.gpu


def _nk():
    import netket_b200 as nk

    return nk


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind,L,n_dim,alpha,std,hb,vb,B,CL", [
    ("ising", 10, 2, 4, 0.05, True, True, 256, 8), ("ising", 20, 1, 1, 0.3, True, True, 16, 63), ("ising", 6, 1, 3, 0.4, False, True, 33, 5),
    ("ising", 7, 1, 2, 0.2, True, False, 20, 3), ("heis", 12, 1, 2, 0.2, True, True, 64, 4), ("ising", 12, 2, 4, 0.02, True, True, 40, 2),
    ("ising", 20, 1, 32, 0.01, True, True, 32, 2)])
def test_forces_and_grad_match_oracle(cuda, dtype, kind, L, n_dim, alpha, std, hb, vb, B, CL):
    nk = _nk()
    g = nk.graph.Hypercube(L, n_dim)
    N = g.n_nodes
    W, b, a = orbm.init_params(N, alpha, seed=1234, std=std, dtype=dtype, use_hidden_bias=hb, use_visible_bias=vb)
    dense = {"kernel": torch.from_numpy(W).cuda()}
    if hb:
        dense["bias"] = torch.from_numpy(b).cuda()
    p = {"Dense": dense}
    if vb:
        p["visible_bias"] = torch.from_numpy(a).cuda()
    var = {"params": p}
    if kind == "ising":
        hi = nk.hilbert.Spin(0.5, N)
        op = nk.operator.Ising(hi, g, h=1.5)
        sa = nk.sampler.MetropolisLocal(hi, n_chains=B)
    else:
        hi = nk.hilbert.Spin(0.5, N, total_sz=0)
        op = nk.operator.Heisenberg(hi, g)
        sa = nk.sampler.MetropolisExchange(hi, graph=g, n_chains=B)
    vs = nk.vqs.MCState(sa, nk.models.RBM(alpha=alpha, param_dtype=dtype, use_hidden_bias=hb, use_visible_bias=vb), variables=var,
                        n_samples=B * CL, n_discard_per_chain=2, sampler_seed=3)
    st, F = vs.expect_and_forces(op)
    st2, G = vs.expect_and_grad(op)
    eloc = vs.local_estimators(op).cpu().numpy()
    samples = vs.samples.cpu().numpy()
    W64 = W.astype(np.float64)
    b64 = None if b is None else b.astype(np.float64)
    a64 = None if a is None else a.astype(np.float64)
    ref = oforces.forces(samples, eloc, W64, b64, a64)
    tol = 1e-10 if dtype == np.float64 else 2e-5
    np.testing.assert_allclose(st.mean, eloc.mean(), rtol=1e-10)
    np.testing.assert_allclose(st2.mean, st.mean, rtol=1e-13)  # statistics reduce with atomics: order varies
    got = {"W": F["Dense"]["kernel"], "b": F["Dense"].get("bias"), "a": F.get("visible_bias")}
    got2 = {"W": G["Dense"]["kernel"], "b": G["Dense"].get("bias"), "a": G.get("visible_bias")}
    for k in ("W", "b", "a"):
        if ref[k] is None:
            assert got[k] is None
            continue
        assert got[k].dtype == (torch.float64 if dtype == np.float64 else torch.float32)
        scale = max(np.abs(ref[k]).max(), 1e-30)
        np.testing.assert_allclose(got[k].cpu().numpy(), ref[k], rtol=0, atol=tol * scale, err_msg=k)
        np.testing.assert_allclose(got2[k].cpu().numpy(), 2.0 * ref[k], rtol=0, atol=2 * tol * scale, err_msg=k)
    assert tuple(got["W"].shape) == (N, alpha * N)


def test_gradient_descends_the_energy(cuda):
    """A few plain SGD steps with expect_and_grad lower <H> of a 1-d Ising chain (the caller of this path, driver/vmc.py:141-161)."""
    nk = _nk()
    N = 10
    g = nk.graph.Hypercube(N, 1)
    hi = nk.hilbert.Spin(0.5, N)
    op = nk.operator.Ising(hi, g, h=1.0)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=512), nk.models.RBM(alpha=2), n_samples=512 * 16, n_discard_per_chain=8,
                        seed=1, sampler_seed=2)
    e0 = vs.expect(op).mean
    for _ in range(30):
        _, G = vs.expect_and_grad(op)
        p = vs.parameters
        new = {"Dense": {"kernel": p["Dense"]["kernel"] - 0.05 * G["Dense"]["kernel"], "bias": p["Dense"]["bias"] - 0.05 * G["Dense"]["bias"]},
               "visible_bias": p["visible_bias"] - 0.05 * G["visible_bias"]}
        vs.parameters = new
    e1 = vs.expect(op).mean
    assert e1 < e0 - 1.0, (e0, e1)
    assert e1 > -1.2738 * N - 0.5  # exact ground-state energy per site of the critical chain is -4/pi


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_forces_with_and_without_fused_tanh(cuda, dtype):
    """The sweep kernel writes tanh(theta) of every recorded sample for the forces; when the samples were drawn earlier
    (`vs.sample()`), theta is recomputed for the batch instead.  Both routes give the oracle's forces."""
    nk = _nk()
    g = nk.graph.Hypercube(10, 2)
    hi = nk.hilbert.Spin(0.5, 100)
    op = nk.operator.Ising(hi, g, h=3.0)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=64), nk.models.RBM(alpha=4, param_dtype=dtype), n_samples=64 * 4,
                        n_discard_per_chain=2, seed=5, sampler_seed=6)
    W, b, a = (t.cpu().numpy().astype(np.float64) for t in nk.models.RBM.unpack(vs.variables))
    tol = 1e-10 if dtype == np.float64 else 2e-5
    for fused in (True, False):
        vs.reset()
        if not fused:
            vs.sample()
            assert vs._tanh is None
        _, F = vs.expect_and_forces(op)
        assert (vs._tanh is not None) == fused
        if fused:  # the kernel's tanh(theta) itself
            th = np.tanh(orbm.theta(vs.samples.cpu().numpy().reshape(-1, 100), W, b))
            np.testing.assert_allclose(vs._tanh.cpu().numpy().reshape(-1, 400), th, rtol=0, atol=1e-12 if dtype == np.float64 else 2e-6)
        ref = oforces.forces(vs.samples.cpu().numpy(), vs.local_estimators(op).cpu().numpy(), W, b, a)
        for k, got in (("W", F["Dense"]["kernel"]), ("b", F["Dense"]["bias"]), ("a", F["visible_bias"])):
            np.testing.assert_allclose(got.cpu().numpy(), ref[k], rtol=0, atol=tol * np.abs(ref[k]).max(), err_msg=f"{k} fused={fused}")
