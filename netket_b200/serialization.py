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


def _gather_chains(t):
    """This rank's chain shard -> the global array (chains of rank 0, rank 1, ...): the reference serialises the global
    ``(n_chains, ...)`` arrays (its fields are sharded jax arrays, netket/sampler/metropolis.py:50-74).  Collective."""
    from .utils import world

    _, ws = world()
    if ws == 1:
        return _np(t)
    import torch.distributed as dist

    src = t.detach()
    if dist.get_backend() == "gloo":
        src = src.cpu()
    parts = [torch.empty_like(src) for _ in range(ws)]
    dist.all_gather(parts, src.contiguous())
    return torch.cat(parts, dim=0).cpu().numpy()


def sampler_state_to_state_dict(state):
    """Keys and shapes of the reference's state dict: global ``σ (n_chains, N)`` and ``n_accepted_proc (n_chains,)``, the
    per-process step counter, ``rule_state`` (None for LocalRule / ExchangeRule), no ``log_prob`` (derived)."""
    return {"σ": _gather_chains(state.σ), "rng": np.asarray(state.rng, dtype=np.uint64), "rule_state": None,
            "n_steps_proc": np.asarray(state.n_steps_proc, dtype=np.int64), "n_accepted_proc": _gather_chains(state.n_accepted_proc)}


def sampler_state_from_state_dict(target, d):
    """Restore into ``target`` (a freshly initialised state of the sampler that will be used): relaxed, ignore errors.
    A global array is sliced to this rank's chains ``[rank * B, (rank + 1) * B)``; a field whose (global) shape does not
    fit keeps the target's value and a warning says so."""
    import warnings

    from .utils import world

    rank, ws = world()
    upd = {}
    dev = target.σ.device
    B, N = target.σ.shape

    def shard(name, arr, trailing):
        arr = np.asarray(arr)
        if tuple(arr.shape) == (B * ws,) + trailing:
            return arr[rank * B:(rank + 1) * B]
        warnings.warn(f"sampler state: stored {name} of shape {tuple(arr.shape)} does not fit {(B * ws,) + trailing}; "
                      "keeping the freshly initialised value", UserWarning, stacklevel=3)
        return None

    if "σ" in d:
        sg = shard("σ", d["σ"], (N,))
        if sg is not None:
            upd["σ"] = torch.as_tensor(np.ascontiguousarray(sg.astype(np.int8)), device=dev)
            if "n_accepted_proc" in d:
                na = shard("n_accepted_proc", d["n_accepted_proc"], ())
                if na is not None:
                    upd["n_accepted_proc"] = torch.as_tensor(np.ascontiguousarray(na.astype(np.int64)), device=dev)
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


# ------------------------------------------------------------------------------------------------------------------
# Bytes: the msgpack wire format of `flax.serialization.to_bytes / from_bytes` (what NetKet users write to `.mpack` files:
# docs `flax.serialization.to_bytes(vstate)`).  flax (>= 0.10.6, pyproject.toml:44) is a third-party dependency that is neither in
# /root/reference nor in this image, so its published format is RESTATED here and is unverified against flax itself ("parity
# unpinned" for this format; the round trip and a hand-assembled byte string are what tests/test_host_logic.py and tests/test_gpu_serialization.py check):
#   * the state dict is packed with msgpack (`strict_types=True`), dict keys are strings;
#   * an ndarray is ExtType(1, packb((shape, dtype.name, C-order bytes), use_bin_type=True));
#   * a Python complex is ExtType(2, packb((re, im))); a NumPy scalar is ExtType(3, <the ndarray encoding of asarray(x)>).
# Arrays above 2^30 bytes are chunked by flax; this path never writes one (parameters and one rank-gathered sigma).
# ------------------------------------------------------------------------------------------------------------------
_EXT_NDARRAY, _EXT_COMPLEX, _EXT_NPSCALAR = 1, 2, 3
_MAX_CHUNK = 2 ** 30


def _ndarray_to_bytes(arr):
    import msgpack

    arr = np.asarray(arr)  # (tobytes("C") below linearises any layout; ascontiguousarray would turn a 0-d scalar into shape (1,))
    if arr.dtype.hasobject:
        raise ValueError("object arrays cannot be serialised")
    if arr.nbytes > _MAX_CHUNK:
        raise ValueError("arrays above 2^30 bytes are not written by this path (flax would chunk them)")
    return msgpack.packb((arr.shape, arr.dtype.name, arr.tobytes("C")), use_bin_type=True)


def _ndarray_from_bytes(data):
    import msgpack

    shape, dtype_name, buf = msgpack.unpackb(data, raw=True)
    return np.frombuffer(buf, dtype=np.dtype(dtype_name.decode()), count=-1, offset=0).reshape(shape, order="C").copy()


def _ext_pack(x):
    import msgpack

    if isinstance(x, torch.Tensor):
        x = _np(x)
    if isinstance(x, np.ndarray):
        return msgpack.ExtType(_EXT_NDARRAY, _ndarray_to_bytes(x))
    if isinstance(x, np.generic):
        return msgpack.ExtType(_EXT_NPSCALAR, _ndarray_to_bytes(np.asarray(x)))
    if isinstance(x, complex):
        return msgpack.ExtType(_EXT_COMPLEX, msgpack.packb((x.real, x.imag)))
    return x


def _ext_unpack(code, data):
    import msgpack

    if code == _EXT_NDARRAY:
        return _ndarray_from_bytes(data)
    if code == _EXT_NPSCALAR:
        return _ndarray_from_bytes(data)[()]
    if code == _EXT_COMPLEX:
        re, im = msgpack.unpackb(data)
        return complex(re, im)
    return msgpack.ExtType(code, data)


def msgpack_serialize(state_dict):
    """``flax.serialization.msgpack_serialize``: nested dict of arrays / scalars -> bytes."""
    import msgpack

    return msgpack.packb(state_dict, default=_ext_pack, strict_types=True)


def msgpack_restore(data):
    """``flax.serialization.msgpack_restore``: bytes -> nested dict of NumPy arrays / scalars."""
    import msgpack

    return msgpack.unpackb(data, ext_hook=_ext_unpack, raw=False)


def to_bytes(vstate):
    """``flax.serialization.to_bytes(vstate)``: the msgpack of ``serialize_MCState(vstate)``."""
    return msgpack_serialize(serialize_MCState(vstate))


def from_bytes(vstate, data):
    """``flax.serialization.from_bytes(vstate, data)``: a copy of ``vstate`` restored from the bytes."""
    return deserialize_MCState(vstate, msgpack_restore(data))
