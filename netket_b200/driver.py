"""``nk.driver.VMC``: energy minimisation (netket/driver/vmc.py:32-176; loop of
netket/_src/driver/abstract_variational_driver.py:349-528, abstract_optimization_driver.py:74-95).

One iteration: ``state.reset(); E, grad = state.expect_and_grad(H)`` (one fused sweep + E_loc + tanh(theta) launch, the
tensor-core force contraction, one all-reduce) -> ``dp = preconditioner(state, grad, step)`` -> optimiser update.
"""

import numbers

from tqdm.auto import tqdm

from .convergence import HistoryDict
from .optimizer import SR, apply_updates, identity_preconditioner
from .utils import world


class RuntimeLog:
    """In-memory logger (netket/logging/runtime_log.py): ``log.data["Energy"]["Mean"]`` -> ``History``."""

    def __init__(self):
        self.data = HistoryDict()

    def __call__(self, step, item, variational_state=None):
        flat = {}
        for name, val in item.items():
            d = val.to_dict() if hasattr(val, "to_dict") else val
            if isinstance(d, dict):
                for k, v in d.items():
                    flat[(name, k)] = v
            else:
                flat[(name, None)] = d
        for (name, k), v in flat.items():
            if k is None:
                self.data.push({name: v}, step)
            else:
                self.data.setdefault(name, HistoryDict()).push({k: v}, step)

    def __getitem__(self, key):
        return self.data[key]

    def flush(self, variational_state=None):
        pass


class VMC:
    def __init__(self, hamiltonian, optimizer, *, variational_state, preconditioner=identity_preconditioner):
        if variational_state.hilbert != hamiltonian.hilbert:
            raise TypeError(f"the variational_state has hilbert space {variational_state.hilbert} (this is normally defined by the "
                            f"hilbert space in the sampler), but the hamiltonian has hilbert space {hamiltonian.hilbert}. "
                            "The two should match.")
        self._ham = hamiltonian
        self._variational_state = variational_state
        self.optimizer = optimizer
        self._optimizer_state = optimizer.init(variational_state.parameters)
        self.preconditioner = preconditioner
        self._step_count = 0
        self._loss_stats = None
        self._loss_grad = None
        self._dp = None
        self._loss_name = "Energy"
        if isinstance(preconditioner, SR) and variational_state.n_samples <= variational_state.n_parameters:
            import warnings

            warnings.warn(f"n_samples={variational_state.n_samples} <= n_parameters={variational_state.n_parameters}: the S matrix "
                          "is rank deficient; consider more samples or a larger diag_shift", UserWarning, stacklevel=2)

    # ------------------------------------------------------------------ properties
    @property
    def preconditioner(self):
        return self._preconditioner

    @preconditioner.setter
    def preconditioner(self, val):
        self._preconditioner = identity_preconditioner if val is None else val

    @property
    def state(self):
        return self._variational_state

    @property
    def step_count(self):
        return self._step_count

    @property
    def energy(self):
        return self._loss_stats

    # ------------------------------------------------------------------ one step
    def compute_loss_and_update(self):
        """vmc.py:141-161."""
        self.state.reset()
        self._loss_stats, self._loss_grad = self.state.expect_and_grad(self._ham)
        self._dp = self.preconditioner(self.state, self._loss_grad, self.step_count)
        return self._loss_stats, self._dp

    def update_parameters(self, dp):
        """abstract_optimization_driver.py:74-95."""
        updates, self._optimizer_state = self.optimizer.update(dp, self._optimizer_state, self.state.parameters)
        self.state.parameters = apply_updates(self.state.parameters, updates)

    def reset(self):
        self.state.reset()
        self._step_count = 0

    def iter(self, n_steps, step=1):
        """Yield the step count every ``step`` optimisation steps (abstract_variational_driver.py:453-485)."""
        for _ in range(0, n_steps, step):
            for i in range(step):
                self._loss_stats, dp = self.compute_loss_and_update()
                if i == 0:
                    yield self.step_count
                self._step_count += 1
                self.update_parameters(dp)

    def advance(self, steps=1):
        for _ in self.iter(steps):
            pass

    def estimate(self, observables):
        if observables is None:
            return {}
        return {k: self.state.expect(o) for k, o in observables.items()}

    def run(self, n_iter, out=(), obs=None, step_size=1, show_progress=True, callback=None):
        """abstract_variational_driver.py:349-451.  ``out``: a logger or an iterable of loggers called as
        ``logger(step, log_data, state)``; ``callback(step, log_data, driver) -> bool`` stops the run when it returns False."""
        if not isinstance(n_iter, numbers.Number):
            raise ValueError("n_iter, the first positional argument to `run`, must be a number!")
        if out is None:
            out = ()
        loggers = tuple(out) if isinstance(out, (list, tuple)) else (out,)
        callbacks = () if callback is None else (tuple(callback) if isinstance(callback, (list, tuple)) else (callback,))
        rank0 = world()[0] == 0
        with tqdm(total=n_iter, disable=not show_progress or not rank0, dynamic_ncols=True) as pbar:
            for step in self.iter(n_iter, step_size):
                log_data = self.estimate(obs)
                log_data[self._loss_name] = self._loss_stats
                pbar.set_postfix_str(f"{self._loss_name}={self._loss_stats}")
                if not all(cb(step, log_data, self) for cb in callbacks):
                    break
                if rank0:
                    for logger in loggers:
                        logger(self.step_count, log_data, self.state)
                pbar.update(step_size)
        for logger in loggers:
            logger.flush(self.state)
        return loggers

    def __repr__(self):
        return f"Vmc(\n  step_count = {self.step_count},\n  state = {self.state})"
