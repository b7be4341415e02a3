"""``nk.sampler.MetropolisLocal`` / ``MetropolisExchange`` on top of ``nk_sweep``.

Mirrors the ``Sampler`` seam S1 (netket/sampler/base.py:254-463, netket/sampler/metropolis.py:206-683):
``init_state / reset / sample / samples / sample_next``, functional state updates (inputs are never mutated),
``n_chains`` rounded up to a multiple of the number of ranks with a warning (metropolis.py:179-203,296-302),
``sweep_size`` default = hilbert.size (:283-284), 16 chains per rank by default (:287-289), ``machine_pow`` real
and >= 0 (base.py:139-150).  Only ``LocalRule`` and ``ExchangeRule`` (uniform or ``probabilities=``) exist: any other
rule raises — there is no generic/CPU fallback.

Multi-GPU: one process per GPU (torch.distributed); this rank owns chains
``[rank * n_chains_per_rank, (rank+1) * n_chains_per_rank)`` and no collective runs during sampling
(the property test/sampler/test_sampler.py:514-548 pins for the reference).
"""

import ctypes as C
import dataclasses
import warnings

import numpy as np
import torch

from . import _lib
from .models import RBM
from .utils import default_device, mix_seed, split_seed, world


# --------------------------------------------------------------------------------------- rules
class MetropolisRule:
    """Base class of transition rules (netket/sampler/rules/base.py:37-144)."""


class LocalRule(MetropolisRule):
    """One uniformly random site per chain is flipped (netket/sampler/rules/local.py:23-52)."""

    code = _lib.NK_RULE_LOCAL

    def __repr__(self):
        return "LocalRule()"

    def __eq__(self, o):
        return isinstance(o, LocalRule)

    def __hash__(self):
        return hash("LocalRule")


class ExchangeRule(MetropolisRule):
    """Exchange of two sites of a randomly chosen *hoppable* cluster, with the log-ratio correction
    (netket/sampler/rules/exchange.py:25-187).  ``clusters``: explicit list, or all pairs of ``graph`` within
    distance ``d_max`` in np.argwhere order (``compute_clusters``, :190-205)."""

    code = _lib.NK_RULE_EXCHANGE

    def __init__(self, *, clusters=None, graph=None, d_max=1, probabilities=None):
        if probabilities is not None:  # exchange.py:112-115
            probabilities = np.atleast_1d(np.asarray(probabilities, dtype=np.float64))
            if not np.all(probabilities > 0):
                raise ValueError("Probabilities must be positive")
        if clusters is None and graph is not None:
            D = np.asarray(graph.distances())
            cl = np.argwhere(D <= d_max)
            clusters = cl[cl[:, 0] < cl[:, 1]]
            if probabilities is not None:  # one weight per graph distance (compute_clusters, :200-203)
                assert probabilities.shape == (d_max,), f"Expected {d_max = } probabilities, got {probabilities}"
                probabilities = probabilities[D[clusters[:, 0], clusters[:, 1]] - 1]
        elif not (clusters is not None and graph is None):
            raise ValueError("You must either provide the list of exchange-clusters or a netket graph, from which "
                             "clusters will be computed using the maximum distance d_max. ")
        self.clusters = np.ascontiguousarray(np.asarray(clusters, dtype=np.int32).reshape(-1, 2))
        if probabilities is not None and len(probabilities) != len(self.clusters):
            raise TypeError(f"Number of clusters and probabilities don't match: {len(self.clusters)} != {len(probabilities)}")
        self.probabilities = None if probabilities is None else np.ascontiguousarray(probabilities, dtype=np.float64)
        self._dev = {}

    def probabilities_on(self, device):
        if self.probabilities is None:
            return None
        k = ("p", str(device))
        if k not in self._dev:
            self._dev[k] = torch.from_numpy(self.probabilities).to(device)
        return self._dev[k]

    def clusters_on(self, device):
        k = str(device)
        if k not in self._dev:
            self._dev[k] = torch.from_numpy(self.clusters).to(device)
        return self._dev[k]

    def __repr__(self):
        return f"ExchangeRule(# of clusters: {len(self.clusters)})"


# --------------------------------------------------------------------------------------- state
@dataclasses.dataclass(frozen=True)
class MetropolisSamplerState:
    """netket/sampler/metropolis.py:42-137.  ``rng`` is (seed, t): Philox key and the number of Metropolis
    steps each chain has performed (the counter).  ``log_prob`` is derived state (not serialised in the reference)."""

    σ: torch.Tensor               # (n_chains_per_rank, N) int8
    rng: tuple                    # (seed, t)
    log_prob: torch.Tensor        # (n_chains_per_rank,)
    n_steps_proc: int = 0
    n_accepted_proc: torch.Tensor = None  # (n_chains_per_rank,) int64
    rule_state: object = None
    chain_offset: int = 0

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)

    @property
    def n_steps(self):
        """Total number of moves performed across all ranks since the last reset (:121-124); every rank runs the same
        launches, so this is the per-process counter times the number of ranks."""
        _, ws = world()
        return self.n_steps_proc * ws

    @property
    def n_accepted(self):
        """Accepted moves summed over all ranks.  COLLECTIVE under torch.distributed (one all-reduce): call it on every
        rank; ``repr`` and logging on a single rank should use ``acceptance_proc``."""
        from .stats import _allreduce

        s = self.n_accepted_proc.sum().to(torch.float64).reshape(1)
        return int(_allreduce(s).item())

    @property
    def acceptance(self):
        """Fraction of accepted moves since the last reset; None before any sampling (:97-108).  Collective, like
        ``n_accepted``."""
        if self.n_steps == 0:
            return None
        return self.n_accepted / self.n_steps

    @property
    def acceptance_proc(self):
        """Acceptance over this rank's chains only (no communication)."""
        if self.n_steps_proc == 0:
            return None
        return int(self.n_accepted_proc.sum().item()) / self.n_steps_proc

    def __repr__(self):  # per-rank counters: printing a state on one rank must not start a collective
        if self.n_steps_proc > 0:
            na = int(self.n_accepted_proc.sum().item())
            return (f"MetropolisSamplerState(# accepted = {na}/{self.n_steps_proc} "
                    f"({na / self.n_steps_proc * 100}%), rng state={self.rng})")
        return f"MetropolisSamplerState(rng state={self.rng})"


# --------------------------------------------------------------------------------------- sampler
class MetropolisSampler:
    def __init__(self, hilbert, rule, *, sweep_size=None, reset_chains=False, n_chains=None, n_chains_per_rank=None,
                 chunk_size=None, machine_pow=2, dtype=None):
        if not isinstance(rule, MetropolisRule):
            raise TypeError(f"The second positional argument, rule, must be a MetropolisRule but `type(rule)={type(rule)}`.")
        if not isinstance(rule, (LocalRule, ExchangeRule)):
            raise NotImplementedError(f"{type(rule).__name__}: only LocalRule and ExchangeRule are implemented")
        if not isinstance(reset_chains, bool):
            raise TypeError("reset_chains must be a boolean.")
        if not (np.isscalar(machine_pow) and np.isreal(machine_pow) and machine_pow >= 0):
            raise ValueError(f"machine_pow ({machine_pow}) must be a non-negative real number.")
        _, ws = world()
        if n_chains is not None and n_chains_per_rank is not None:
            raise ValueError("Cannot specify both `n_chains` and `n_chains_per_rank`")
        if n_chains is None and n_chains_per_rank is None:
            n_chains_per_rank = 16  # default_n_chains_per_rank, metropolis.py:287-289
        if n_chains is not None:
            n_chains_per_rank = max(int(np.ceil(n_chains / ws)), 1)
            if n_chains_per_rank * ws != n_chains:
                warnings.warn(f"Using {n_chains_per_rank} chains per rank among {ws} ranks (total="
                              f"{n_chains_per_rank * ws} instead of n_chains={n_chains}). To directly control the number "
                              f"of chains on every rank, specify `n_chains_per_rank` when constructing the sampler.",
                              category=UserWarning, stacklevel=2)
        if n_chains_per_rank <= 0:
            raise ValueError("n_chains must be positive")
        if sweep_size is None:
            sweep_size = hilbert.size
        if sweep_size < 1:
            raise ValueError("sweep_size must be >= 1")
        if chunk_size is not None and (not isinstance(chunk_size, int) or chunk_size <= 0):
            raise ValueError("chunk_size must be a positive integer or None")
        if dtype is not None and np.dtype(dtype) != np.int8:
            raise NotImplementedError("netket_b200 samplers store configurations as int8 (the reference's default)")
        self.hilbert = hilbert
        self.rule = rule
        self.sweep_size = int(sweep_size)
        self.reset_chains = reset_chains
        self.n_chains_per_rank = int(n_chains_per_rank)
        self.n_chains = int(n_chains_per_rank) * ws
        self.chunk_size = chunk_size  # accepted and ignored: the fused kernel never materialises a batch
        self.machine_pow = float(machine_pow)
        self.dtype = np.dtype(np.int8)

    is_exact = False

    @property
    def n_batches(self):
        return self.n_chains_per_rank

    def replace(self, **kw):
        args = dict(sweep_size=self.sweep_size, reset_chains=self.reset_chains, n_chains_per_rank=self.n_chains_per_rank,
                    chunk_size=self.chunk_size, machine_pow=self.machine_pow)
        hilbert = kw.pop("hilbert", self.hilbert)
        rule = kw.pop("rule", self.rule)
        if "n_chains" in kw:
            args.pop("n_chains_per_rank")
        args.update(kw)
        return MetropolisSampler(hilbert, rule, **args)

    # ------------------------------------------------------------------ internals
    @staticmethod
    def _check_machine(machine):
        if not isinstance(machine, RBM):
            raise NotImplementedError(
                f"{type(machine).__name__}: the fused sampler recognises netket_b200.models.RBM only "
                "(no generic apply-function path, no CPU fallback)")

    def _workspace(self, rbm, B, device):
        """Scratch for the fast path (theta from the GEMM + hand-over flag), cached per (shape, device)."""
        nbytes = int(_lib.lib().nk_sweep_workspace_bytes(C.byref(rbm), B))
        if nbytes <= 0:
            return None
        key = (str(device), nbytes)
        cache = self.__dict__.setdefault("_ws_cache", {})
        if key not in cache:
            cache.clear()
            cache[key] = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return cache[key]

    def _random_state(self, seed, chain_offset, device):
        return self.hilbert.random_state(seed, self.n_chains_per_rank, chain_offset=chain_offset, device=device)

    # ------------------------------------------------------------------ Sampler API
    def init_state(self, machine, parameters, seed=None):
        """``Sampler.init_state`` (base.py:254-287) -> ``_init_state`` (metropolis.py:353-380)."""
        self._check_machine(machine)
        W, _, _ = RBM.unpack(parameters)
        device = W.device if W.is_cuda else default_device()
        seed = split_seed(seed)
        rank, _ = world()
        off = rank * self.n_chains_per_rank
        B = self.n_chains_per_rank
        if self.reset_chains:
            sigma = torch.zeros((B, self.hilbert.size), dtype=torch.int8, device=device)
        else:
            sigma = self._random_state(mix_seed(seed, 0), off, device)
        log_prob = torch.full((B,), -float("inf"), dtype=W.dtype, device=device)
        nacc = torch.zeros((B,), dtype=torch.int64, device=device)
        return MetropolisSamplerState(σ=sigma, rng=(seed, 0), log_prob=log_prob, n_steps_proc=0, n_accepted_proc=nacc,
                                      chain_offset=off)

    def reset(self, machine, parameters, state=None):
        """``_reset`` (metropolis.py:382-414): optionally re-randomise, recompute log_prob, zero the counters."""
        self._check_machine(machine)
        if state is None:
            state = self.init_state(machine, parameters)
        sigma = state.σ
        seed, t = state.rng
        if self.reset_chains:
            sigma = self._random_state(mix_seed(seed, 1 + t), state.chain_offset, sigma.device)
        log_prob = (self.machine_pow * machine.apply(parameters, sigma)).to(state.log_prob.dtype)
        return state.replace(σ=sigma, log_prob=log_prob, n_steps_proc=0, n_accepted_proc=torch.zeros_like(state.n_accepted_proc))

    def _launch(self, machine, parameters, state, chain_length, *, n_discard=0, return_log_probabilities=False,
                operator=None, stream=None, path=_lib.NK_PATH_AUTO, want_samples=True, tanh_out=None, stats_shift=None,
                no_handover=False):
        """One ``nk_sweep`` call.  Returns (samples, logp|None, eloc|None, new_state).  ``tanh_out``: optional tensor
        ``(B, chain_length, M)`` that receives tanh(theta) of every recorded sample (input of ``nk_forces_rbm``).
        ``stats_shift``: with an operator, also reduce the statistics' partial sums inside the launch (shifted by this
        estimate of the mean); the return value then has a fifth element, a float64 device tensor
        ``[NK_STATS_NPARTIAL sums | n_chains of this rank]`` ready for ONE all-reduce."""
        self._check_machine(machine)
        rbm = RBM.c_struct(parameters)
        N = self.hilbert.size
        if rbm.N != N:
            raise ValueError(f"the model has {rbm.N} visible units, the Hilbert space {N} sites")
        W, _, _ = RBM.unpack(parameters)
        dev = W.device
        B = self.n_chains_per_rank
        sigma = state.σ.clone()  # functional semantics: the input state is not mutated
        nacc = state.n_accepted_proc.clone()
        log_prob = torch.empty((B,), dtype=W.dtype, device=dev)
        seed, t = state.rng
        ws = self._workspace(rbm, B, dev) if path != _lib.NK_PATH_GENERIC else None
        chains = _lib.nk_chains_t(sigma=sigma.data_ptr(), log_prob=log_prob.data_ptr(), n_accepted=nacc.data_ptr(),
                                  workspace=ws.data_ptr() if ws is not None else None, B=B, seed=seed, t=t,
                                  chain_offset=state.chain_offset)
        samples = torch.empty((B, chain_length, N), dtype=torch.int8, device=dev) if want_samples else None
        logp = torch.empty((B, chain_length), dtype=W.dtype, device=dev) if return_log_probabilities else None
        a = _lib.nk_sweep_t()
        a.rule = self.rule.code
        a.chain_length = int(chain_length)
        a.n_discard = int(n_discard)
        a.sweep_size = self.sweep_size
        a.machine_pow = self.machine_pow
        a.samples_out = samples.data_ptr() if samples is not None else None
        a.logp_out = logp.data_ptr() if logp is not None else None
        keep = []
        if stream is not None:
            w0, u = stream
            T = (n_discard + chain_length) * self.sweep_size
            w0 = torch.as_tensor(np.asarray(w0, dtype=np.uint32).view(np.int32)).to(dev).contiguous()
            u = torch.as_tensor(np.asarray(u)).to(device=dev, dtype=W.dtype).contiguous()
            if tuple(w0.shape) != (T, B) or tuple(u.shape) != (T, B):
                raise ValueError(f"explicit proposal stream must have shape {(T, B)}")
            keep += [w0, u]
            a.stream_w0, a.stream_u = w0.data_ptr(), u.data_ptr()
        if isinstance(self.rule, ExchangeRule):
            cl = self.rule.clusters_on(dev)
            a.clusters, a.n_clusters = cl.data_ptr(), int(cl.shape[0])
            pr = self.rule.probabilities_on(dev)
            if pr is not None:
                a.cluster_probs = pr.data_ptr()
        a.path = path
        if tanh_out is not None:
            if tuple(tanh_out.shape) != (B, chain_length, rbm.M) or tanh_out.dtype != W.dtype or not tanh_out.is_contiguous():
                raise ValueError("tanh_out must be a contiguous (n_chains, chain_length, n_hidden) tensor of the parameter dtype")
            a.tanh_out = tanh_out.data_ptr()
        eloc = None
        if operator is not None:
            from .operator import IsingJax, LocalOperatorJax

            out_dtype = torch.promote_types(_lib.torch_dtype(operator.dtype), W.dtype)
            eloc = torch.empty((B, chain_length), dtype=out_dtype, device=dev)
            a.eloc_out, a.eloc_dtype = eloc.data_ptr(), _lib.dtype_code(out_dtype)
            if isinstance(operator, IsingJax):
                st = operator._c_struct(dev)
                keep.append(st)
                a.ising = C.pointer(st)
            elif isinstance(operator, LocalOperatorJax):
                st = operator._c_struct(dev)
                keep.append(st)
                a.localop = C.pointer(st)
            else:
                raise NotImplementedError(f"no fused local-energy kernel for {type(operator).__name__}")
        part = None
        if stats_shift is not None:
            if operator is None:
                raise ValueError("stats_shift needs an operator")
            part = torch.empty(_lib.NK_STATS_NPARTIAL + 1, dtype=torch.float64, device=dev)
            part[_lib.NK_STATS_NPARTIAL] = float(B)
            a.stats_out, a.stats_shift = part.data_ptr(), float(stats_shift)
            if no_handover:  # only the kernel NK_PATH_AUTO tries first; part[0] is NaN if it had to give up (caller repeats)
                a.flags = _lib.NK_SWEEP_NO_HANDOVER
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nk_sweep(_lib.stream_ptr(dev), C.byref(rbm), C.byref(chains), C.byref(a)))
        n_steps = (n_discard + chain_length) * self.sweep_size
        new_state = state.replace(σ=sigma, rng=(seed, int(chains.t)), log_prob=log_prob, n_accepted_proc=nacc,
                                  n_steps_proc=state.n_steps_proc + n_steps * B)
        if part is not None:
            return samples, logp, eloc, new_state, part
        return samples, logp, eloc, new_state

    def sample(self, machine, parameters, *, state=None, chain_length=1, return_log_probabilities=False,
               _stream=None, _path=_lib.NK_PATH_AUTO):
        """``Sampler.sample`` (base.py:344-383) -> ``_sample_chain`` (metropolis.py:466-505).
        Returns ``(samples[n_chains_per_rank, chain_length, N], state)`` or ``((samples, log_prob), state)``."""
        if state is None:
            state = self.reset(machine, parameters)
        samples, logp, _, state = self._launch(machine, parameters, state, chain_length,
                                               return_log_probabilities=return_log_probabilities, stream=_stream, path=_path)
        if return_log_probabilities:
            return (samples, logp), state
        return samples, state

    def sample_next(self, machine, parameters, state=None):
        """One sweep; returns ``(state, sigma)`` (metropolis.py:324-351, inverted order on purpose)."""
        samples, state = self.sample(machine, parameters, state=state, chain_length=1)
        return state, samples[:, 0, :]

    def samples(self, machine, parameters, *, state=None, chain_length=1):
        """Generator over ``chain_length`` successive batches (base.py:385-410)."""
        if state is None:
            state = self.reset(machine, parameters)
        for _ in range(chain_length):
            s, state = self.sample(machine, parameters, state=state, chain_length=1)
            yield s[:, 0, :]

    def __repr__(self):
        return (f"{type(self).__name__}(\n  hilbert = {self.hilbert},\n  rule = {self.rule},\n  n_chains = {self.n_chains},"
                f"\n  sweep_size = {self.sweep_size},\n  reset_chains = {self.reset_chains},\n  machine_power = "
                f"{self.machine_pow},\n  dtype = int8)")


def MetropolisLocal(hilbert, **kwargs):
    """``nk.sampler.MetropolisLocal`` (metropolis.py:532-570)."""
    return MetropolisSampler(hilbert, LocalRule(), **kwargs)


def MetropolisExchange(hilbert, *, clusters=None, graph=None, d_max=1, probabilities=None, **kwargs):
    """``nk.sampler.MetropolisExchange`` (metropolis.py:573-683)."""
    return MetropolisSampler(hilbert, ExchangeRule(clusters=clusters, graph=graph, d_max=d_max, probabilities=probabilities), **kwargs)
