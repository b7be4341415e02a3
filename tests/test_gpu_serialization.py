"""State-dict round trips (netket/vqs/mc/mc_state/state.py:962-1016,
test/sampler/test_metropolis_serialization.py:38-111): a restored MCState re-draws the same samples and continues the same
chains; a file written with a different number of chains restores the parameters and the RNG but keeps fresh chains."""

import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _state(nk, n_chains, dtype=np.float64, seed=3):
    g = nk.graph.Hypercube(10, 1)
    hi = nk.hilbert.Spin(0.5, 10)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=n_chains), nk.models.RBM(alpha=2, param_dtype=dtype), n_samples=n_chains * 4,
                        n_discard_per_chain=2, seed=seed, sampler_seed=seed + 1)
    return g, hi, vs


def test_mcstate_round_trip_resumes_the_same_chains(cuda):
    import netket_b200 as nk

    g, hi, vs = _state(nk, 16)
    op = nk.operator.Ising(hi, g, h=1.0)
    vs.sample()                       # advance once so that the stored sampler state is not the initial one
    vs.reset()
    sd = pickle.loads(pickle.dumps(vs.to_state_dict()))  # plain dicts of NumPy arrays / Python scalars
    assert set(sd) == {"variables", "sampler_state", "n_samples", "n_discard_per_chain", "chunk_size"}
    # the reference's keys (metropolis.py:50-74): log_prob is not serialised, rule_state is (None for these rules)
    assert set(sd["sampler_state"]) == {"σ", "rng", "rule_state", "n_steps_proc", "n_accepted_proc"}
    assert sd["sampler_state"]["σ"].dtype == np.int8 and sd["sampler_state"]["rng"].dtype == np.uint64
    e1 = vs.expect(op)
    s1 = vs.samples.clone()
    # a differently initialised state of the same structure, restored from the dict
    _, _, other = _state(nk, 16, seed=99)
    assert not torch.equal(other.parameters["Dense"]["kernel"], vs.parameters["Dense"]["kernel"])
    restored = other.from_state_dict(sd)
    assert torch.equal(restored.parameters["Dense"]["kernel"], vs.parameters["Dense"]["kernel"])
    assert restored.n_samples == vs.n_samples and restored.n_discard_per_chain == 2
    e2 = restored.expect(op)
    assert torch.equal(restored.samples, s1)
    assert e2.mean == e1.mean or abs(e2.mean - e1.mean) < 1e-12 * abs(e1.mean)
    # serialising while samples are cached stores the state they were drawn from: the copy re-draws the same samples
    sd2 = vs.to_state_dict()
    again = other.from_state_dict(sd2)
    assert torch.equal(again.samples, s1)
    # the msgpack bytes of flax.serialization.to_bytes / from_bytes (wire format restated in serialization.py)
    blob = vs.to_bytes()
    assert isinstance(blob, bytes)
    from_blob = other.from_bytes(blob)
    assert torch.equal(from_blob.samples, s1) and from_blob.n_samples == vs.n_samples
    assert torch.equal(from_blob.parameters["Dense"]["kernel"], vs.parameters["Dense"]["kernel"])
    # and both continue identically
    vs.reset()
    restored.reset()
    assert torch.equal(vs.sample(), restored.sample())


def test_restore_with_a_different_number_of_chains(cuda):
    import netket_b200 as nk

    _, _, vs = _state(nk, 16)
    vs.sample()
    sd = vs.to_state_dict()
    _, _, small = _state(nk, 8, seed=99)
    fresh_sigma = small.sampler_state.σ.clone()
    restored = small.from_state_dict(sd)
    assert torch.equal(restored.parameters["visible_bias"], vs.parameters["visible_bias"])
    assert torch.equal(restored.sampler_state.σ, fresh_sigma)            # shapes differ: the fresh chains are kept
    assert restored.sampler_state.rng == vs.sampler_state.rng or restored.sampler_state.rng[0] == sd["sampler_state"]["rng"][0]
    assert tuple(restored.sample().shape)[0] == 8
    with pytest.raises(ValueError):
        _, _, wrong = _state(nk, 8, seed=1)
        wrong._variables["params"]["Dense"]["kernel"] = torch.zeros((10, 30), dtype=torch.float64, device="cuda")
        wrong.from_state_dict(sd)
