"""A NumPy stand-in for the few pieces of jax / jax.numpy that the reference's hot-path *functions* use, so that their
own source can be executed in this container (jax is not installed) to generate golden vectors.

Used only by tests/golden/make_golden.py.  Nothing here re-states reference logic: it re-states JAX semantics
(``.at[].set`` / ``.add``, ``vmap``, ``lax.select``, ``lax.cond``, ``lax.scan``, ``lax.dynamic_slice_in_dim``,
``jnp.where(size=, fill_value=)``) on top of NumPy.
"""

import functools
import types

import numpy as np


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, val):
        out = np.array(self.arr, copy=True).view(Arr)
        out[self.idx] = val
        return out

    def add(self, val):
        out = np.array(self.arr, copy=True).view(Arr)
        np.add.at(out, self.idx, val)
        return out


class Arr(np.ndarray):
    """ndarray with the functional-update property ``.at``."""

    @property
    def at(self):
        return _At(self)


def _wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, Arr):
        return x.view(Arr)
    if isinstance(x, (np.generic,)):
        return np.asarray(x).view(Arr)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    return x


def _lift(fn):
    @functools.wraps(fn)
    def g(*a, **k):
        return _wrap(fn(*a, **k))

    return g


def _where(cond, x=None, y=None, *, size=None, fill_value=None):
    if x is not None or y is not None:
        return _wrap(np.where(cond, x, y))
    (idx,) = np.nonzero(np.asarray(cond))
    if size is not None:
        out = np.full(size, fill_value if fill_value is not None else 0, dtype=idx.dtype)
        n = min(size, idx.size)
        out[:n] = idx[:n]
        idx = out
    return (_wrap(idx),)


def _make_jnp():
    jnp = types.ModuleType("jax.numpy")
    for name in ("zeros", "ones", "full", "arange", "eye", "asarray", "array", "abs", "exp", "log", "log1p", "sqrt", "sum", "mean",
                 "var", "hstack", "broadcast_to", "expand_dims", "take", "signbit", "atleast_1d", "swapaxes", "dot",
                 "isclose", "concatenate", "stack", "max", "min", "floor", "zeros_like", "pad", "real", "minimum", "maximum", "all", "cumsum", "argmin", "any", "conjugate"):
        setattr(jnp, name, _lift(getattr(np, name)))
    jnp.where = _where
    jnp.clip = lambda a, min=None, max=None: _wrap(np.clip(a, min, max))  # noqa: A002  (jnp.clip(x, 0) == lower bound only)
    jnp.nan = np.nan
    fft = types.ModuleType("jax.numpy.fft")
    fft.fft = _lift(np.fft.fft)
    fft.ifft = _lift(np.fft.ifft)
    jnp.fft = fft
    jnp.inf = np.inf
    jnp.int32, jnp.int64, jnp.float32, jnp.float64, jnp.bool_ = np.int32, np.int64, np.float32, np.float64, np.bool_
    jnp.bool = np.bool_
    jnp.integer = np.integer
    jnp.issubdtype = np.issubdtype
    jnp.ndarray = np.ndarray
    return jnp


def _index(arg, ax, i):
    if ax is None:
        return arg
    if isinstance(arg, (tuple, list)):
        axes = ax if isinstance(ax, (tuple, list)) else [ax] * len(arg)
        return type(arg)(_index(a, x, i) for a, x in zip(arg, axes))
    return _wrap(np.take(np.asarray(arg), i, axis=ax))


def _size(arg, ax):
    if ax is None:
        return None
    if isinstance(arg, (tuple, list)):
        axes = ax if isinstance(ax, (tuple, list)) else [ax] * len(arg)
        for a, x in zip(arg, axes):
            n = _size(a, x)
            if n is not None:
                return n
        return None
    return np.asarray(arg).shape[ax]


def _stack(outs):
    first = outs[0]
    if isinstance(first, tuple):
        return tuple(_stack([o[k] for o in outs]) for k in range(len(first)))
    return _wrap(np.stack([np.asarray(o) for o in outs], axis=0))


def vmap(fun=None, in_axes=0, out_axes=0):
    if fun is None:
        return functools.partial(vmap, in_axes=in_axes, out_axes=out_axes)
    assert out_axes == 0

    @functools.wraps(fun)
    def wrapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            n = _size(a, ax)
            if n is not None:
                break
        outs = [fun(*[_index(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)]
        return _stack(outs)

    return wrapped


def jit(fun=None, **kw):
    if fun is None:
        return functools.partial(jit, **kw)
    return fun


def _cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def _dynamic_slice_in_dim(operand, start_index, slice_size, axis=0):
    """jax.lax.dynamic_slice_in_dim: the start index is clamped so that the slice stays inside the operand."""
    n = np.asarray(operand).shape[axis]
    start = int(np.clip(int(start_index), 0, n - slice_size))
    idx = [slice(None)] * np.asarray(operand).ndim
    idx[axis] = slice(start, start + slice_size)
    return _wrap(np.asarray(operand)[tuple(idx)])


def _scan(f, init, xs):
    """jax.lax.scan for bodies that return (carry, None)."""
    carry = init
    for x in xs:
        carry, y = f(carry, x)
        assert y is None
    return carry, None


def make_jax():
    jax = types.ModuleType("jax")
    jnp = _make_jnp()
    jax.numpy = jnp
    jax.jit = jit
    jax.vmap = vmap
    lax = types.ModuleType("jax.lax")
    lax.select = lambda m, a, b: _wrap(np.where(m, a, b))
    lax.cond = _cond
    lax.scan = _scan
    lax.dynamic_slice_in_dim = _dynamic_slice_in_dim
    jax.lax = lax
    jax.Array = np.ndarray
    tree_util = types.ModuleType("jax.tree_util")
    tree_util.register_pytree_node_class = lambda c: c
    jax.tree_util = tree_util
    return jax, jnp
