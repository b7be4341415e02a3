"""The CUDA sweep kernels in fixed-proposal-stream mode against chains produced by EXECUTING THE REFERENCE'S OWN SOURCE
(tests/golden/sampler_vectors.npz, see tests/golden/make_golden_sampler.py): north_star's "a fixed-proposal-stream mode
reproduces the reference's chains", on every kernel path; and `expect_and_forces` against `forces_expect_hermitian` run
from the reference's source."""

import numpy as np
import pytest
import torch

from test_golden_sampler import CASES, G, load_case

pytestmark = pytest.mark.gpu

PATHS = {"auto": 0, "theta-form": 1, "product-form": 3}


def _sampler(nk, c):
    N = c["L"] ** c["n_dim"]
    hi = nk.hilbert.Spin(0.5, N, total_sz=c["total_sz"])
    B = c["sigma0"].shape[0]
    if c["rule"] == "local":
        sa = nk.sampler.MetropolisLocal(hi, n_chains=B, sweep_size=c["sweep_size"], machine_pow=c["machine_pow"])
    else:
        sa = nk.sampler.MetropolisSampler(hi, nk.sampler.ExchangeRule(clusters=c["clusters"], probabilities=c["probs"]), n_chains=B,
                                          sweep_size=c["sweep_size"], machine_pow=c["machine_pow"])
    return hi, sa


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("tag", CASES)
def test_cuda_chain_equals_reference_source_chain_fp64(cuda, tag, path):
    import netket_b200 as nk

    c = load_case(tag)
    hi, sa = _sampler(nk, c)
    var = {"params": {"Dense": {"kernel": torch.from_numpy(c["W"]).cuda(), "bias": torch.from_numpy(c["b"]).cuda()},
                      "visible_bias": torch.from_numpy(c["a"]).cuda()}}
    model = nk.models.RBM(alpha=c["alpha"])
    st = sa.init_state(model, var, seed=1).replace(σ=torch.from_numpy(c["sigma0"]).cuda())
    (samples, logp), st2 = sa.sample(model, var, state=st, chain_length=c["n_sweeps"], return_log_probabilities=True,
                                     _stream=(c["w0"], c["u"]), _path=PATHS[path])
    assert np.array_equal(samples.cpu().numpy(), c["samples"])
    np.testing.assert_allclose(logp.cpu().numpy(), c["logp"], rtol=1e-10, atol=1e-10)
    assert np.array_equal(st2.n_accepted_proc.cpu().numpy(), c["nacc"])
    assert st2.n_steps_proc == c["nsteps"]


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-7), (np.float32, 2e-5)])
@pytest.mark.parametrize("tag", ["forces_1d", "forces_2d"])
def test_cuda_forces_equal_reference_source_forces(cuda, tag, dtype, tol):
    import netket_b200 as nk

    L, nd, alpha, nch, cl = (int(v) for v in G[f"{tag}_cfg"])
    W, b, a, sig = (G[f"{tag}_{k}"] for k in ("W", "b", "a", "sigma"))
    g = nk.graph.Hypercube(L, nd)
    hi = nk.hilbert.Spin(0.5, g.n_nodes)
    op = nk.operator.Ising(hi, g, h=float(G[f"{tag}_h"]))
    assert np.array_equal(np.asarray(op.edges), G[f"{tag}_edges"])
    var = {"params": {"Dense": {"kernel": torch.from_numpy(W.astype(dtype)).cuda(), "bias": torch.from_numpy(b.astype(dtype)).cuda()},
                      "visible_bias": torch.from_numpy(a.astype(dtype)).cuda()}}
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=nch), nk.models.RBM(alpha=alpha, param_dtype=dtype), variables=var,
                        n_samples=nch * cl, seed=1)
    vs._samples = torch.from_numpy(sig).cuda()  # the golden batch instead of drawn samples
    stats, F = vs.expect_and_forces(op)
    np.testing.assert_allclose(stats.mean, float(G[f"{tag}_mean"]), rtol=1e-12 if dtype == np.float64 else 1e-5)
    scale = np.abs(G[f"{tag}_F_kernel"]).max()
    np.testing.assert_allclose(F["Dense"]["kernel"].cpu().numpy(), G[f"{tag}_F_kernel"], rtol=tol, atol=tol * scale)
    np.testing.assert_allclose(F["Dense"]["bias"].cpu().numpy(), G[f"{tag}_F_bias"], rtol=tol, atol=tol * scale)
    np.testing.assert_allclose(F["visible_bias"].cpu().numpy(), G[f"{tag}_F_visible"], rtol=tol, atol=tol * scale)
