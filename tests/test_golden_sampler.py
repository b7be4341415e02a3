"""The oracle's Metropolis loop and force estimator against vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE
(tests/golden/make_golden_sampler.py: `MetropolisSampler._sample_next`, `LocalRule.transition`, `ExchangeRule.transition`
with and without `probabilities`, `forces_expect_hermitian`) with injected draws: same proposal stream => same chains."""

import os

import numpy as np
import pytest

from oracle import forces as oforces
from oracle import graph as ograph
from oracle import operators as oops
from oracle import sampler as osampler

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sampler_vectors.npz"))

CASES = ["mh_local_1d", "mh_local_2d", "mh_local_sweep5", "mh_exch_1d", "mh_exch_2d_d2", "mh_exch_sz1", "mh_exch_weighted",
         "mh_exch_weighted3"]


def load_case(tag):
    L, n_dim, alpha, B, n_sweeps, sweep_size, d_max, total_sz = (int(v) for v in G[f"{tag}_cfg"])
    kw = dict(W=G[f"{tag}_W"], b=G[f"{tag}_b"], a=G[f"{tag}_a"], sigma0=G[f"{tag}_sigma0"], w0=G[f"{tag}_w0"], u=G[f"{tag}_u"],
              samples=G[f"{tag}_samples"], logp=G[f"{tag}_logp"], nacc=G[f"{tag}_nacc"], nsteps=int(G[f"{tag}_nsteps"]),
              machine_pow=float(G[f"{tag}_pow"]), n_sweeps=n_sweeps, sweep_size=sweep_size, L=L, n_dim=n_dim, alpha=alpha, d_max=d_max,
              total_sz=None if total_sz == -99 else total_sz)
    kw["rule"] = "exchange" if f"{tag}_clusters" in G.files else "local"
    kw["clusters"] = G[f"{tag}_clusters"] if kw["rule"] == "exchange" else None
    kw["probs"] = G[f"{tag}_probs"] if f"{tag}_probs" in G.files else None
    return kw


@pytest.mark.parametrize("tag", CASES)
def test_oracle_chain_equals_reference_source_chain(tag):
    c = load_case(tag)
    r = osampler.sample_chain(c["rule"], c["sigma0"], c["W"], c["b"], c["a"], chain_length=c["n_sweeps"], sweep_size=c["sweep_size"],
                              machine_pow=c["machine_pow"], stream=(c["w0"], c["u"]), clusters=c["clusters"], probabilities=c["probs"])
    assert np.array_equal(r["samples"], c["samples"])
    np.testing.assert_allclose(r["log_prob_samples"], c["logp"], rtol=1e-13, atol=1e-13)
    assert np.array_equal(r["n_accepted"], c["nacc"])
    assert r["n_steps"] == c["nsteps"]
    assert 0 < c["nacc"].sum() < c["nsteps"], "the vectors must exercise both branches of the accept"


def test_oracle_clusters_equal_the_generator_s(tmp_path):
    for tag in CASES:
        c = load_case(tag)
        if c["rule"] == "exchange":
            e, _ = ograph.hypercube_edges(c["L"], c["n_dim"])
            assert np.array_equal(ograph.compute_clusters(c["L"] ** c["n_dim"], e, c["d_max"]), c["clusters"])


@pytest.mark.parametrize("tag", ["forces_1d", "forces_2d"])
def test_oracle_forces_equal_reference_source_forces(tag):
    """expect_forces.py:67-112 executed from source (vjp by Richardson-extrapolated central differences of the reference's forward
    pass, so the vectors carry ~1e-9 of differentiation error)."""
    W, b, a, sig, edges = (G[f"{tag}_{k}"] for k in ("W", "b", "a", "sigma", "edges"))
    h = float(G[f"{tag}_h"])
    mean, F = oforces.expect_and_forces(sig, lambda x: oops.ising_conn_padded(x, edges, h, 1.0), W, b, a)
    np.testing.assert_allclose(mean, float(G[f"{tag}_mean"]), rtol=1e-12)
    np.testing.assert_allclose(F["kernel"], G[f"{tag}_F_kernel"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(F["bias"], G[f"{tag}_F_bias"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(F["visible_bias"], G[f"{tag}_F_visible"], rtol=1e-7, atol=1e-8)
