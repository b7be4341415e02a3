"""Small host-side helpers (device selection, seeds, torch.distributed plumbing)."""

import os

import numpy as np
import torch
import torch.distributed as dist

_MASK64 = (1 << 64) - 1


def default_device(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        from ._lib import NkError

        raise NkError("no CUDA device: netket_b200 runs on B200 GPUs only (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def world():
    """(rank, world_size) of the one-process-per-GPU job (1 process if torch.distributed is not initialised)."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def split_seed(key):
    """Accepts None / int / numpy integer and returns a 64-bit seed (the analogue of nk.jax.PRNGKey,
    netket/jax/_utils_random.py:25-77: None draws a fresh seed, which is broadcast from rank 0)."""
    if key is None:
        s = int.from_bytes(os.urandom(8), "little")
        rank, ws = world()
        if ws > 1:
            t = torch.tensor([s & ((1 << 62) - 1)], dtype=torch.int64)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, src=0)
            s = int(t.item())
        return s & _MASK64
    return int(key) & _MASK64


def mix_seed(seed, salt):
    """SplitMix64 step: derive independent sub-seeds (the analogue of jax.random.split)."""
    z = (int(seed) + 0x9E3779B97F4A7C15 * (int(salt) + 1)) & _MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
    return z ^ (z >> 31)


def as_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)
