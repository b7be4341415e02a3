"""State dicts of the sampler state and of MCState.

Mirrors the reference's Flax state-dict serialisation (netket/vqs/mc/mc_state/state.py:962-1016 `serialize_MCState` /
`deserialize_MCState`; netket/sampler/metropolis.py:50-74,131-137): nested dicts of NumPy arrays with the reference's
keys.  As there, ``log_prob`` is derived state and is not written, and the sampler-state fields are restored in the
"relaxed, ignore errors" mode: a field whose shape does not fit the target (e.g. a file written with a different number
of chains, test/sampler/test_metropolis_serialization.py:38-111) keeps the target's freshly initialised value.
The JAX PRNG key of the reference is replaced by the (seed, counter) pair of this library's Philox stream, stored as a
``uint64[2]`` array under the same key ``rng``.
"""

import numpy as np
import torch


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def sampler_state_to_state_dict(state):
    return {"σ": _np(state.σ), "rng": np.asarray(state.rng, dtype=np.uint64), "n_steps_proc": np.asarray(state.n_steps_proc, dtype=np.int64),
            "n_accepted_proc": _np(state.n_accepted_proc)}


def sampler_state_from_state_dict(target, d):
    """Restore into ``target`` (a freshly initialised state of the sampler that will be used): relaxed, ignore errors."""
    upd = {}
    dev = target.σ.device
    if "σ" in d and tuple(np.shape(d["σ"])) == tuple(target.σ.shape):
        upd["σ"] = torch.as_tensor(np.asarray(d["σ"], dtype=np.int8), device=dev)
        if "n_accepted_proc" in d and tuple(np.shape(d["n_accepted_proc"])) == tuple(target.n_accepted_proc.shape):
            upd["n_accepted_proc"] = torch.as_tensor(np.asarray(d["n_accepted_proc"], dtype=np.int64), device=dev)
            upd["n_steps_proc"] = int(np.asarray(d.get("n_steps_proc", 0)))
    if "rng" in d and np.size(d["rng"]) == 2:
        r = np.asarray(d["rng"], dtype=np.uint64).reshape(2)
        upd["rng"] = (int(r[0]), int(r[1]))
    return target.replace(**upd)


def _tree_to_numpy(tree):
    return {k: (_tree_to_numpy(v) if isinstance(v, dict) else _np(v)) for k, v in tree.items()}


def _tree_from_numpy(target, tree):
    out = {}
    for k, v in target.items():
        if isinstance(v, dict):
            out[k] = _tree_from_numpy(v, tree[k])
        else:
            arr = np.asarray(tree[k])
            if tuple(arr.shape) != tuple(v.shape):
                raise ValueError(f"parameter {k!r}: stored shape {arr.shape} does not match {tuple(v.shape)}")
            out[k] = torch.as_tensor(arr, device=v.device).to(v.dtype)
    return out


def serialize_MCState(vstate):
    """state.py:963-981.  Samples are not written; the sampler state is the one the cached samples were drawn *from*, so that
    a restored state re-draws the same samples."""
    st = vstate._sampler_state_previous if vstate._samples is not None and vstate._sampler_state_previous is not None else vstate.sampler_state
    return {"variables": _tree_to_numpy(vstate.variables), "sampler_state": sampler_state_to_state_dict(st),
            "n_samples": vstate.n_samples, "n_discard_per_chain": vstate.n_discard_per_chain, "chunk_size": vstate.chunk_size}


def deserialize_MCState(vstate, state_dict):
    """state.py:984-1009: returns a copy of ``vstate`` with the stored variables, sampler state and sampling settings."""
    import copy

    new = copy.copy(vstate)
    new.reset()
    new.variables = _tree_from_numpy(vstate.variables, state_dict["variables"])
    new.sampler_state = sampler_state_from_state_dict(vstate.sampler_state, state_dict["sampler_state"])
    new.n_samples = state_dict["n_samples"]
    new.n_discard_per_chain = state_dict["n_discard_per_chain"]
    new.chunk_size = state_dict["chunk_size"]
    return new
