"""``nk.vqs.MCState``: sample() / samples / log_value / local_estimators / expect.

Mirrors netket/vqs/mc/mc_state/state.py:123-959 for the (Metropolis sampler, RBM) pair:

* ``n_samples`` is rounded up to a multiple of ``n_chains`` (``compute_chain_length``, state.py:60-79), default
  = first multiple of ``n_chains`` >= 1000 (:306-309); ``n_discard_per_chain`` default 5 (:466-468).
* ``sample()`` always resets the sampler first, runs ``n_discard_per_chain`` burn-in sweeps and then
  ``chain_length`` recorded sweeps (:521-576).  ``samples`` is cached until ``reset()``; assigning ``parameters``
  / ``variables`` calls ``reset()`` (netket/vqs/base.py:88-116).
* ``expect(O)`` = ``local_estimators(O).to_stats()`` (mc_state/expect.py:124-128, local_estimators.py:34-82).
  If no samples are cached the sweep kernel is launched *fused* with the local energy (theta stays on chip and
  sigma' is never materialised); with cached samples the stand-alone E_loc kernel runs on them.  Both give the
  same numbers (tests/test_mcstate_gpu.py).
* ``chunk_size`` is accepted and ignored: results are chunk-invariant by the reference's own tests
  (test/variational/test_variational.py:497-515) and the fused kernels have nothing to chunk.
"""

import ctypes as C
import warnings

import numpy as np
import torch

from . import _lib
from ._dispatch import Dispatcher
from .models import RBM
from .operator import IsingJax, LocalOperatorJax
from .stats import Stats, statistics
from .utils import mix_seed, split_seed, world


def compute_chain_length(n_chains, n_samples):
    """state.py:60-79."""
    if n_samples <= 0:
        raise ValueError(f"Invalid number of samples: n_samples={n_samples}")
    chain_length = int(np.ceil(n_samples / n_chains))
    n_new = chain_length * n_chains
    if n_new != n_samples:
        _, ws = world()
        warnings.warn(f"n_samples={n_samples} ({n_samples // ws} per device) does not divide n_chains={n_chains}, "
                      f"increased to {n_new} ({n_new // ws} per device)", UserWarning, stacklevel=3)
    return chain_length


# array attributes the reference forwards to ``.data`` for backward compatibility (local_estimators.py:37-44, 84-94); the
# torch.Tensor spelling of the same names
_ARRAY_LIKE_ATTRS = frozenset({"reshape", "mean", "std", "sum", "min", "max", "flatten", "ravel", "T", "dtype", "real", "imag",
                               "conj", "tolist", "item", "size", "ndim", "shape", "cpu", "numpy", "device"})


class LocalEstimators:
    """Per-sample scalar estimators, ``data`` of shape (n_chains, chain_length) (netket/_src/stats/local_estimators.py:47-129):
    ``to_stats()`` one-shot ``Stats``; ``to_online_stats()`` / ``accumulate(old)`` start or update the streaming accumulator
    (``OnlineStats``, K7).  Array attributes are forwarded to ``.data`` as the reference does."""

    def __init__(self, data):
        self.data = data

    def __getattr__(self, name):
        # the reference forwards a list of array attributes (and everything else works there through __jax_array__); a torch
        # tensor's methods (.abs(), .to(), ...) are the counterpart of jnp functions applied to the container
        if name != "data" and not name.startswith("_") and (name in _ARRAY_LIKE_ATTRS or hasattr(self.data, name)):
            return getattr(self.data, name)
        raise AttributeError(f"LocalEstimators has no attribute {name!r}.")

    # the reference's container converts implicitly to an array (`__jax_array__`, local_estimators.py:75-83): the torch spelling
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        from torch.utils._pytree import tree_map

        unwrap = lambda x: x.data if isinstance(x, LocalEstimators) else x  # noqa: E731
        return func(*tree_map(unwrap, args), **tree_map(unwrap, kwargs or {}))

    def __array__(self, dtype=None):
        a = self.data.detach().cpu().numpy()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return self.data[idx]

    def __neg__(self):
        return -self.data

    def __add__(self, o):
        return self.data + getattr(o, "data", o)

    __radd__ = __add__

    def __sub__(self, o):
        return self.data - getattr(o, "data", o)

    def __rsub__(self, o):
        return getattr(o, "data", o) - self.data

    def __mul__(self, o):
        return self.data * getattr(o, "data", o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self.data / getattr(o, "data", o)

    def __pow__(self, o):
        return self.data ** o

    def to_stats(self):
        return statistics(self.data)

    def to_online_stats(self, *, max_lag=64):
        from .stats import online_statistics

        return online_statistics(self.data, None, max_lag=max_lag)

    def accumulate(self, old=None, *, max_lag=64):
        """Fold this batch into an online accumulator (``old`` from a previous call, or None to start one)."""
        if old is None:
            return self.to_online_stats(max_lag=max_lag)
        return old.update(self.data)


class LocalEstimatorsBatch:
    """K-channel estimators for nonlinear observables, ``data`` of shape (n_chains, chain_length, K), with a ``combinator``
    ``(K,) -> scalar | array`` of the channel means whose error is propagated by the delta method
    (netket/_src/stats/local_estimators.py:132-246).  ``combinator`` is a function of a float64 torch vector."""

    def __init__(self, data, combinator):
        self.data = data
        self.combinator = combinator

    @property
    def n_channels(self):
        return int(self.data.shape[-1])

    def __getattr__(self, name):
        if name in _ARRAY_LIKE_ATTRS:
            raise AttributeError(f"LocalEstimatorsBatch has no attribute {name!r}. The underlying array is at .data; "
                                 f"use le.data.{name} instead.")
        raise AttributeError(f"LocalEstimatorsBatch has no attribute {name!r}.")

    def to_stats(self):
        """Delta-method statistics of this batch alone: covariance of the chain means (of the samples when there is one chain);
        over all ranks.  ``Stats`` for a scalar combinator, ``StatsBatch`` for an array-valued one."""
        import torch

        from .stats import _allreduce, _delta_method_stats

        d = self.data.to(torch.float64)
        K = d.shape[-1]
        rows = d.mean(dim=1) if d.shape[0] >= 1 else d.reshape(0, K)  # chain means of this rank
        # first and second moments over ALL chains (one all-reduce of K + K^2 + 1 doubles), then the centred covariance
        pack = torch.cat([rows.sum(dim=0), (rows.T @ rows).reshape(-1), torch.tensor([float(rows.shape[0])], dtype=torch.float64,
                                                                                     device=d.device)])
        pack = _allreduce(pack)
        n = float(pack[-1].item())
        X = pack[:K] / n
        if n < 2:  # one chain in total: the samples themselves (local_estimators.py:206-210)
            flat = d.reshape(-1, K)
            dev = flat - X[None, :]
            Cov = (dev.T @ dev) / float(flat.shape[0]) ** 2
        else:
            Cov = (pack[K:K + K * K].reshape(K, K) / n - X[:, None] * X[None, :]) / n
        return _delta_method_stats(self.combinator, X, Cov)

    def to_online_stats(self, *, max_lag=64):
        from .stats import OnlineStatsBatch

        return OnlineStatsBatch.from_data(self.data, self.combinator, max_lag=max_lag)

    def accumulate(self, old=None, *, max_lag=64):
        if old is None:
            return self.to_online_stats(max_lag=max_lag)
        return old.update(self.data)


# ------------------------------------------------------------------------------------------------------------------
# Seam S4: the local-estimator multimethods (netket/vqs/mc/common.py:31-99; registrations of mc_state/expect.py:54-78,
# 124-128, local_estimators.py:41-82).  `MCState.local_estimators / expect` go through them, so a user overload
#     @nk.vqs.local_estimators.dispatch
#     def _(vstate: nk.vqs.MCState, op: MyOperator, chunk_size: None): ...
# takes effect exactly as in the reference (docs/advanced/custom-operators/local-estimators.ipynb).
# ------------------------------------------------------------------------------------------------------------------
get_local_kernel_arguments = Dispatcher(
    "get_local_kernel_arguments", "(vstate, O) -> (sigma, args): the samples and whatever the local kernel needs (common.py:31-46)")
get_local_kernel = Dispatcher(
    "get_local_kernel", "(vstate, O[, chunk_size]) -> kernel(logpsi, pars, sigma, args[, chunk_size=]) -> O_loc[B] (common.py:49-67)")
local_estimators = Dispatcher(
    "local_estimators", "(vstate, O, chunk_size) -> LocalEstimators with data (n_chains, chain_length) (common.py:70-99)")
expect = Dispatcher("expect", "(vstate, O, chunk_size) -> Stats (mc_state/expect.py:124-128)")


def check_hilbert(A, B):
    """common.py:24-28."""
    if not A == B:
        raise NotImplementedError(f"Non matching hilbert spaces {A} and {B}")


def local_value_kernel_rbm(logpsi, pars, sigma, op, *, chunk_size=None):
    """``local_value_kernel_jax`` (netket/vqs/mc/kernels.py:62-71) for (RBM, Ising | LocalOperator): O_loc[...] of sigma[..., N]
    from nk_eloc_ising_rbm / nk_eloc_localop_rbm - sigma' is never materialised.  ``logpsi``: the RBM module or its bound
    ``apply``; any other ansatz raises (there is no generic / CPU path).  ``chunk_size`` is accepted and ignored."""
    model = getattr(logpsi, "__self__", logpsi)
    if not isinstance(model, RBM):
        raise NotImplementedError(f"{type(model).__name__}: the fused local-value kernel recognises netket_b200.models.RBM only")
    return _eloc_on_samples(pars, op, sigma)


def _eloc_on_samples(variables, op, sigma, path=_lib.NK_PATH_AUTO, ws_holder=None):
    """Stand-alone E_loc on sigma[..., N] (any leading dimensions)."""
    rbm = RBM.c_struct(variables)
    W, _, _ = RBM.unpack(variables)
    dev = W.device
    shape = tuple(sigma.shape[:-1])
    s8 = sigma.reshape(-1, rbm.N).to(torch.int8).contiguous()
    B = s8.shape[0]
    out_dtype = torch.promote_types(_lib.torch_dtype(op.dtype), W.dtype)
    out = torch.empty((B,), dtype=out_dtype, device=dev)
    st = op._c_struct(dev)
    ws = None
    if path != _lib.NK_PATH_GENERIC:  # scratch of the product-form kernel (theta + tables), cached per size
        nbytes = int(_lib.lib().nk_sweep_workspace_bytes(C.byref(rbm), B))
        if nbytes > 0:
            holder = ws_holder if ws_holder is not None else _eloc_on_samples.__dict__
            cur = holder.get("_eloc_ws")
            if cur is None or cur.numel() < nbytes or cur.device != dev:
                cur = holder["_eloc_ws"] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            ws = cur
    wsp = _lib.ptr(ws) if ws is not None else None
    with torch.cuda.device(dev):
        if isinstance(op, IsingJax):
            _lib.check(_lib.lib().nk_eloc_ising_rbm(_lib.stream_ptr(dev), C.byref(rbm), C.byref(st), _lib.ptr(s8), B,
                                                    _lib.ptr(out), _lib.dtype_code(out_dtype), path, wsp))
        elif isinstance(op, LocalOperatorJax):
            _lib.check(_lib.lib().nk_eloc_localop_rbm(_lib.stream_ptr(dev), C.byref(rbm), C.byref(st), _lib.ptr(s8), B,
                                                      _lib.ptr(out), _lib.dtype_code(out_dtype), path, wsp))
        else:
            raise NotImplementedError(f"no local-value kernel for {type(op).__name__}")
    return out.reshape(shape)


class MCState:
    def __init__(self, sampler, model=None, *, n_samples=None, n_samples_per_rank=None, n_discard_per_chain=None,
                 chunk_size=None, variables=None, seed=None, sampler_seed=None):
        if model is None:
            raise ValueError("MCState needs a model (netket_b200.models.RBM)")
        if not isinstance(model, RBM):
            raise NotImplementedError(f"{type(model).__name__}: MCState's fused kernels recognise netket_b200.models.RBM only")
        self._model = model
        self._sampler = sampler
        seed = split_seed(seed)
        self._sampler_seed = split_seed(sampler_seed) if sampler_seed is not None else mix_seed(seed, 7)
        if variables is None:
            variables = model.init(seed & 0xFFFFFFFF, sampler.hilbert.size)
        self._variables = variables
        self.sampler_state = sampler.init_state(model, variables, seed=self._sampler_seed)
        self._samples = None
        self._eloc_cache = None   # (operator, E_loc): the operator object itself is held, an id() can be recycled
        self._stats_cache = None  # (operator, Stats) of the fused launch
        self._shift_hint = None   # (operator, last mean): shift of the in-kernel statistics (any estimate is exact)
        self._eloc_ws = None
        self._forces_ws = None
        self._sampler_state_previous = None
        self._tanh = None
        self._chain_length = None
        _, ws = world()
        if n_samples is not None and n_samples_per_rank is not None:
            raise ValueError("Only one argument between `n_samples` and `n_samples_per_rank`can be specified at the same time.")
        if n_samples_per_rank is not None:
            n_samples = n_samples_per_rank * ws
        if n_samples is None:
            n_samples = sampler.n_chains * int(np.ceil(1000 / sampler.n_chains))  # state.py:306-309
        self.n_samples = n_samples
        self.n_discard_per_chain = n_discard_per_chain
        self.chunk_size = chunk_size

    # ------------------------------------------------------------------ properties
    @property
    def model(self):
        return self._model

    @property
    def hilbert(self):
        return self._sampler.hilbert

    @property
    def sampler(self):
        return self._sampler

    @sampler.setter
    def sampler(self, s):
        self._sampler = s
        self.sampler_state = s.init_state(self._model, self._variables, seed=self._sampler_seed)
        self.reset()

    @property
    def variables(self):
        return self._variables

    @variables.setter
    def variables(self, v):
        self._variables = v
        self.reset()

    @property
    def parameters(self):
        return self._variables["params"]

    @parameters.setter
    def parameters(self, p):
        self._variables = {**self._variables, "params": p}
        self.reset()

    @property
    def model_state(self):
        return {}

    @property
    def n_parameters(self):
        W, b, a = RBM.unpack(self._variables)
        return W.numel() + (b.numel() if b is not None else 0) + (a.numel() if a is not None else 0)

    @property
    def n_samples(self):
        return self._chain_length * self._sampler.n_chains

    @n_samples.setter
    def n_samples(self, n):
        self.chain_length = compute_chain_length(self._sampler.n_chains, n)

    @property
    def n_samples_per_rank(self):
        return self._chain_length * self._sampler.n_chains_per_rank

    @property
    def chain_length(self):
        return self._chain_length

    @chain_length.setter
    def chain_length(self, L):
        if L <= 0:
            raise ValueError(f"Invalid chain length: chain_length={L}")
        self._chain_length = int(L)
        self.reset()

    @property
    def n_discard_per_chain(self):
        return self._n_discard

    @n_discard_per_chain.setter
    def n_discard_per_chain(self, n):
        if n is not None and n < 0:
            raise ValueError(f"Invalid number of discarded samples: n_discard_per_chain={n}")
        self._n_discard = 5 if n is None else int(n)  # state.py:466-468

    @property
    def chunk_size(self):
        return self._chunk_size

    @chunk_size.setter
    def chunk_size(self, c):
        if c is not None and (not isinstance(c, int) or c <= 0):
            raise ValueError("Chunk size must be a positive integer or None.")
        self._chunk_size = c

    # ------------------------------------------------------------------ sampling
    def reset(self):
        """Drop the cached samples so that the next access re-samples (state.py:514-519)."""
        self._samples = None
        self._eloc_cache = None
        self._stats_cache = None
        self._tanh = None

    def _run(self, chain_length, n_discard, operator=None, path=_lib.NK_PATH_AUTO, want_tanh=False, fused_stats=False):
        sa = self._sampler
        # sampler.reset: counters zeroed (and chains re-randomised if reset_chains); log_prob is rebuilt in-kernel
        st = self.sampler_state.replace(n_steps_proc=0, n_accepted_proc=torch.zeros_like(self.sampler_state.n_accepted_proc))
        if sa.reset_chains:
            st = sa.reset(self._model, self._variables, st)
        self._sampler_state_previous = self.sampler_state  # what the samples are drawn from (serialisation, state.py:555)
        self._tanh = None
        if want_tanh:  # tanh(theta) of every sample, written by the sweep kernel for the forces (1.7 GB at 2^20 x 400 fp32)
            W, _, _ = RBM.unpack(self._variables)
            self._tanh = torch.empty((sa.n_chains_per_rank, chain_length, W.shape[1]), dtype=W.dtype, device=W.device)
        shift = None
        if fused_stats and operator is not None:
            hint = self._shift_hint
            shift = hint[1] if (hint is not None and hint[0] is operator) else 0.0
        out = sa._launch(self._model, self._variables, st, chain_length, n_discard=n_discard, operator=operator, path=path,
                         tanh_out=self._tanh, stats_shift=shift, no_handover=shift is not None and path == _lib.NK_PATH_AUTO)
        stats = None
        if shift is not None:
            stats = self._finish_stats(out[2], out[4], shift, chain_length)
            if stats is None:  # the optimistic single-kernel launch met weights outside its range: the full chain of kernels
                out = sa._launch(self._model, self._variables, st, chain_length, n_discard=n_discard, operator=operator, path=path,
                                 tanh_out=self._tanh, stats_shift=shift)
                stats = self._finish_stats(out[2], out[4], shift, chain_length)
        samples, _, eloc, st = out[:4]
        self.sampler_state = st
        if shift is not None:
            self._stats_cache = (operator, stats)
            self._shift_hint = (operator, stats.mean)
        return samples, eloc

    def _finish_stats(self, eloc, part, shift, L):
        """ONE all-reduce (NK_STATS_NPARTIAL sums + the chain count) and ONE host read per fused expect.  The sums are shifted
        by `shift`; if that estimate turns out to be so far from the mean that the one-pass variance formulas would lose
        more than ~1e-11 to cancellation (first call on a low-variance state), the two-pass kernel runs instead."""
        from .stats import _allreduce, finalize

        _allreduce(part)
        p = part.tolist()
        if p[0] != p[0]:
            return None  # NK_SWEEP_NO_HANDOVER: the tuned kernel gave up on every rank (the weights are the same everywhere)
        n_chains_total = int(round(p[_lib.NK_STATS_NPARTIAL]))
        sums = p[:_lib.NK_STATS_NPARTIAL]
        ts = float(n_chains_total * L)
        dm = sums[7] / ts
        var = sums[0] / ts - dm * dm
        if not (dm * dm * max(L, 1) <= 1.0e5 * var):
            return statistics(eloc)
        return finalize(sums, shift, n_chains_total, L)

    def sample(self, *, chain_length=None, n_samples=None, n_discard_per_chain=None):
        """state.py:521-576."""
        if n_samples is not None and chain_length is not None:
            raise ValueError("Cannot specify both `chain_length` and `n_samples`.")  # state.py:540-546
        if n_samples is None and chain_length is None:
            chain_length = self._chain_length
        elif chain_length is None:
            chain_length = compute_chain_length(self._sampler.n_chains, n_samples)
        if n_discard_per_chain is None:
            n_discard_per_chain = self._n_discard
        self.reset()
        self._samples, _ = self._run(chain_length, n_discard_per_chain)
        return self._samples

    @property
    def samples(self):
        if self._samples is None:
            self.sample()
        return self._samples

    def log_value(self, sigma):
        """state.py:594-610."""
        return self._model.apply(self._variables, sigma)

    # ------------------------------------------------------------------ estimators
    def _check_operator(self, op):
        if not isinstance(op, (IsingJax, LocalOperatorJax)):
            raise NotImplementedError(
                f"no local-estimator kernel registered for (MCState, {type(op).__name__}): supported operators are Ising, "
                "Heisenberg / GraphOperator / LocalOperator with 1- and 2-site terms")
        if op.hilbert != self.hilbert:
            raise TypeError("Hilbert spaces of the state and of the operator do not match")

    def _eloc_on_samples(self, op, sigma, path=_lib.NK_PATH_AUTO):
        """Stand-alone E_loc on sigma[..., N]."""
        return _eloc_on_samples(self._variables, op, sigma, path, self.__dict__)

    def local_estimators(self, op, *, chunk_size=None):
        """O_loc for every sample, shape (n_chains_per_rank, chain_length) (state.py:612-692): the `local_estimators`
        multimethod on (type(self), type(op), chunk_size).  Returns the reference's container (`LocalEstimators`: `.data`,
        `to_stats()`, `accumulate()`; usable wherever a tensor is, like the reference's implicit array conversion)."""
        return local_estimators(self, op, self.chunk_size if chunk_size is None else chunk_size)

    def expect(self, op):
        """<O> with MC statistics (state.py:695-712): the `expect` multimethod.  For the built-in operators and no cached
        samples this is ONE launch: sweeps, local energies and the statistics' partial sums come out of the same kernel,
        followed by one all-reduce and one host read."""
        return expect(self, op, self.chunk_size)

    def _local_estimators_fused(self, op):
        """Built-in operators: cached E_loc, else the fused launch (no samples yet), else the stand-alone kernel."""
        self._check_operator(op)
        if self._eloc_cache is not None and self._eloc_cache[0] is op:
            return self._eloc_cache[1]
        if self._samples is None:
            self._samples, eloc = self._run(self._chain_length, self._n_discard, operator=op, fused_stats=True)
        else:
            eloc = self._eloc_on_samples(op, self._samples)
        self._eloc_cache = (op, eloc)
        return eloc

    # ------------------------------------------------------------------ streaming callers
    def _sample_and_estimate(self, op, n_discard_per_chain=None):
        """``self.sample(n_discard_per_chain=...); return self.local_estimators(op)`` as ONE fused launch."""
        self._check_operator(op)
        self.reset()
        n_discard = self._n_discard if n_discard_per_chain is None else n_discard_per_chain
        self._samples, eloc = self._run(self._chain_length, n_discard, operator=op)
        self._eloc_cache = (op, eloc)
        return eloc

    def _set_sampler_keep_state(self, sampler, sampler_state):
        """``self.sampler = sampler; self.sampler_state = sampler_state`` without drawing a fresh sampler state in between
        (check_mc_convergence.py:132-134: same chains, another sweep size)."""
        self._sampler = sampler
        self.sampler_state = sampler_state
        self.reset()

    def check_mc_convergence(self, op, *, min_chain_length=50, max_chain_length=500, plot=False):
        """state.py:829-844."""
        from .convergence import check_mc_convergence

        return check_mc_convergence(self, op, min_chain_length=min_chain_length, max_chain_length=max_chain_length, plot=plot)

    def thermalise(self, op, *, min_chain_length=10, max_chain_length=100, rhat_tol=1.05, decay=0.9, patience=1, verbose=True,
                   raise_on_failure=False):
        """state.py:846-869."""
        from .convergence import thermalise_mcmc

        return thermalise_mcmc(self, op, min_chain_length=min_chain_length, max_chain_length=max_chain_length, rhat_tol=rhat_tol,
                               decay=decay, patience=patience, verbose=verbose, raise_on_failure=raise_on_failure)

    def expect_to_precision(self, op, *, atol=None, rtol=None, max_iter=10_000, max_lag=64, verbose=True):
        """state.py:871-942."""
        from .convergence import expect_to_precision

        return expect_to_precision(self, op, atol=atol, rtol=rtol, max_iter=max_iter, max_lag=max_lag, verbose=verbose)

    def expect_and_forces(self, op, *, mutable=False):
        """``(Stats, forces)`` with ``forces[k] = < d log psi / d p_k * (E_loc - <E_loc>) >`` in the layout of
        ``self.parameters`` (netket/vqs/mc/mc_state/expect_forces.py:39-112).  The RBM's log-derivatives are closed
        forms, so the vjp is one contraction over the samples (nk_forces_rbm); between GPUs only the
        ``N*M + M + N`` double sums are all-reduced."""
        return self._forces(op, 1.0)

    def expect_and_grad(self, op, *, use_covariance=None, mutable=False):
        """``(Stats, grad)``: for the real-parameter RBM the gradient of <O> is ``2 Re F``
        (``force_to_grad``, netket/vqs/mc/common.py:103-118; hermitian operators only, as the reference's default)."""
        if use_covariance is False:
            raise NotImplementedError("expect_and_grad(use_covariance=False): the non-hermitian estimator is not implemented")
        return self._forces(op, 2.0)

    def _forces(self, op, factor):
        from .stats import _allreduce

        self._check_operator(op)
        if self._samples is None:  # one fused launch: sweeps + E_loc + tanh(theta) of every sample
            self._samples, eloc = self._run(self._chain_length, self._n_discard, operator=op, want_tanh=True, fused_stats=True)
            self._eloc_cache = (op, eloc)
        stats = self.expect(op)
        eloc = self.local_estimators(op).data
        samples = self.samples
        rbm = RBM.c_struct(self._variables)
        W, b, a = RBM.unpack(self._variables)
        dev = W.device
        N, M = rbm.N, rbm.M
        s8 = samples.reshape(-1, N).contiguous()
        e = eloc.reshape(-1).contiguous()
        Ns = s8.shape[0]
        L = _lib.lib()
        tanh = self._tanh if (self._tanh is not None and self._tanh.numel() == Ns * M) else None
        if tanh is None:  # samples were drawn earlier without tanh(theta): recompute theta for the batch
            nbytes = int(L.nk_forces_workspace_bytes(C.byref(rbm), Ns))
            if self._forces_ws is None or self._forces_ws.numel() < nbytes or self._forces_ws.device != dev:
                self._forces_ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        n = N * M + M + N
        sums = torch.empty(n, dtype=torch.float64, device=dev)
        _, ws = world()
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.check(L.nk_forces_rbm(st, C.byref(rbm), _lib.ptr(s8), Ns, _lib.ptr(e), _lib.dtype_code(e.dtype), float(stats.mean),
                                       _lib.ptr(sums), _lib.ptr(self._forces_ws) if tanh is None else None,
                                       _lib.ptr(tanh) if tanh is not None else None))
            _allreduce(sums)  # the only cross-device traffic of the gradient: n_parameters doubles
            out = torch.empty(n, dtype=W.dtype, device=dev)
            _lib.check(L.nk_forces_finalize(st, _lib.ptr(sums), factor / float(Ns * ws), n, _lib.ptr(out), _lib.dtype_code(W.dtype)))
        dense = {"kernel": out[: N * M].reshape(N, M)}
        if b is not None:
            dense["bias"] = out[N * M: N * M + M]
        forces = {"Dense": dense}
        if a is not None:
            forces["visible_bias"] = out[N * M + M:]
        return stats, forces

    # ------------------------------------------------------------------ serialisation (state.py:962-1016)
    def to_state_dict(self):
        from .serialization import serialize_MCState

        return serialize_MCState(self)

    def from_state_dict(self, state_dict):
        """Returns a copy of this state restored from ``state_dict`` (the reference's ``flax.serialization.from_state_dict``)."""
        from .serialization import deserialize_MCState

        return deserialize_MCState(self, state_dict)

    def to_bytes(self):
        """``flax.serialization.to_bytes(vstate)``: msgpack bytes in flax's wire format (serialization.py)."""
        from .serialization import to_bytes

        return to_bytes(self)

    def from_bytes(self, data):
        """``flax.serialization.from_bytes(vstate, data)``: a copy of this state restored from the bytes."""
        from .serialization import from_bytes

        return from_bytes(self, data)

    def __repr__(self):
        return (f"MCState(\n  hilbert = {self.hilbert},\n  sampler = {self._sampler},\n  n_samples = {self.n_samples},\n"
                f"  n_discard_per_chain = {self._n_discard},\n  sampler_state = {self.sampler_state},\n"
                f"  n_parameters = {self.n_parameters})")


# ------------------------------------------------------------------------------------------------------------------
# registrations for (MCState, Ising | LocalOperator) - mc_state/expect.py:54-78,124-128; local_estimators.py:41-82
# ------------------------------------------------------------------------------------------------------------------
_BUILTIN_OPS = (IsingJax, LocalOperatorJax)


@get_local_kernel_arguments.dispatch
def _(vstate: MCState, op: _BUILTIN_OPS):
    check_hilbert(vstate.hilbert, op.hilbert)
    return vstate.samples, op  # a DiscreteJaxOperator is its own kernel argument (expect.py:69-73)


@get_local_kernel.dispatch
def _(vstate: MCState, op: _BUILTIN_OPS):
    return local_value_kernel_rbm


@get_local_kernel.dispatch
def _(vstate: MCState, op: _BUILTIN_OPS, chunk_size: (int, type(None))):
    return local_value_kernel_rbm  # nothing is materialised, so there is nothing to chunk


@get_local_kernel.dispatch(precedence=-10)
def _(vstate, op, chunk_size: None):  # common.py:63-67
    return get_local_kernel(vstate, op)


@local_estimators.dispatch
def _(vstate: MCState, op: _BUILTIN_OPS, chunk_size: (int, type(None))):
    return LocalEstimators(vstate._local_estimators_fused(op))


@local_estimators.dispatch(precedence=-50)
def _(vstate: MCState, op, chunk_size: (int, type(None))):
    """Generic route of local_estimators.py:41-82 for operators registered through get_local_kernel(_arguments)."""
    sigma, args = get_local_kernel_arguments(vstate, op)
    kernel = get_local_kernel(vstate, op, chunk_size)
    data = kernel(vstate.model, vstate.variables, sigma.reshape(-1, sigma.shape[-1]), args)
    return LocalEstimators(data.reshape(sigma.shape[:-1]))


@local_estimators.dispatch(precedence=-100)
def _(vstate, op, chunk_size):  # common.py:84-99
    raise NotImplementedError(
        f"local_estimators is not implemented for the combination of vstate type {type(vstate).__name__} and operator type "
        f"{type(op).__name__}.\nTo add support, register a dispatch. For MCState + custom AbstractOperator, define separate "
        f"chunk_size=None and chunk_size=int overloads to avoid ambiguity:\n    @nk.vqs.local_estimators.dispatch\n"
        f"    def _(vstate: YourState, op: YourOp, chunk_size: None):\n        ...\n    @nk.vqs.local_estimators.dispatch\n"
        f"    def _(vstate: YourState, op: YourOp, chunk_size: int):\n        ...")


@expect.dispatch
def _(vstate: MCState, op, chunk_size: (int, type(None))):  # mc_state/expect.py:124-128
    le = local_estimators(vstate, op, chunk_size)
    sc = vstate._stats_cache
    if sc is not None and sc[0] is op and vstate._eloc_cache is not None and vstate._eloc_cache[1] is le.data:
        return sc[1]  # reduced inside the sweep kernel
    return le.to_stats()
