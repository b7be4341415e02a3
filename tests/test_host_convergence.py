"""Control flow of the streaming callers (netket_b200/convergence.py) on CPU: the accumulator is replaced by an adapter around
the oracle's (tests may use the oracle; the product's accumulator is CUDA-only and is tested in -m gpu), the state by a fake
that emits AR(1) series.  Checks the loop logic the reference prescribes: iteration counts, stopping rules, histories,
sweep-size doubling, warnings (netket/_src/vqs/expect_to_precision.py:69-164, check_mc_convergence.py:31-457)."""

import math
import warnings

import numpy as np
import pytest

from netket_b200 import convergence as conv
from netket_b200.sampler import MetropolisSampler
from netket_b200.stats import Stats
from oracle import online_stats as oos


class Acc:
    """The product accumulator's interface on top of the oracle's."""

    def __init__(self, o):
        self.o = o

    max_lag = property(lambda s: s.o.max_lag)
    n_chains = property(lambda s: s.o.n_chains)
    _n_samples_total = property(lambda s: s.o.n_samples)
    n_samples = property(lambda s: s.o.n_samples)
    mean = property(lambda s: s.o.mean)
    variance = property(lambda s: s.o.variance)
    R_hat = property(lambda s: s.o.R_hat)
    tau_corr_acf = property(lambda s: s.o.tau_corr_acf)
    tau_corr_batch = property(lambda s: s.o.tau_corr_batch)
    error_of_mean = property(lambda s: s.o.error_of_mean)
    decay = property(lambda s: 1.0 if s.o.decay is None else s.o.decay)

    def get_stats(self):
        return Stats(**self.o.get_stats())

    def __repr__(self):
        return repr(self.get_stats())


def fake_online_statistics(data, old=None, *, decay=None, max_lag=64, inplace=False):
    return Acc(oos.online_statistics(np.asarray(data), None if old is None else old.o, decay=decay, max_lag=max_lag))


class FakeSampler(MetropolisSampler):
    def __init__(self, n_chains, sweep_size=4):  # no Hilbert space needed for the loop logic
        self.n_chains, self.sweep_size = n_chains, sweep_size

    def replace(self, **kw):
        new = FakeSampler(self.n_chains, self.sweep_size)
        new.__dict__.update(kw)
        return new


class FakeState:
    """Emits (n_chains, chain_length) blocks of an AR(1) series per chain; phi shrinks with the sweep size like a real chain's."""

    def __init__(self, n_chains=8, chain_length=5, phi=0.5, offset=-3.0, spread=0.0, seed=0):
        self.sampler = FakeSampler(n_chains)
        self.sampler_state = "s0"
        self.chain_length = chain_length
        self.phi1, self.offset = phi ** (1.0 / 4), offset   # per elementary step; phi at sweep_size 4
        self.rs = np.random.default_rng(seed)
        self.x = self.rs.normal(size=n_chains)
        self.shift = spread * np.arange(n_chains)            # unthermalised chains: constant per-chain offsets that decay
        self.calls = []

    def _set_sampler_keep_state(self, sampler, sampler_state):
        self.sampler, self.sampler_state = sampler, sampler_state

    def _sample_and_estimate(self, op, n_discard=None):
        self.calls.append((op, n_discard, self.sampler.sweep_size))
        phi = self.phi1 ** self.sampler.sweep_size
        out = np.empty((self.sampler.n_chains, self.chain_length))
        for t in range(self.chain_length):
            self.x = phi * self.x + math.sqrt(1 - phi * phi) * self.rs.normal(size=self.x.size)
            self.shift = self.shift * 0.7
            out[:, t] = self.offset + self.x + self.shift
        return out

    def sample(self, n_discard_per_chain=None):
        self._pending = n_discard_per_chain

    def local_estimators(self, op):
        return self._sample_and_estimate(op, self._pending) if not hasattr(self, "_cached") else self._cached


@pytest.fixture(autouse=True)
def oracle_accumulator(monkeypatch):
    monkeypatch.setattr(conv, "online_statistics", fake_online_statistics)
    monkeypatch.setattr(conv, "acf_window_saturated", lambda a: oos.acf_window_saturated(a.o))
    monkeypatch.setattr(conv, "tau_corr_reliable", lambda a: oos.tau_corr_reliable(a.o))
    monkeypatch.setattr(conv, "thin_acf_by_2", lambda a: Acc(oos.thin_acf_by_2(a.o)))
    monkeypatch.setattr(conv, "expand_max_lag", lambda a, n: Acc(oos.expand_max_lag(a.o, n)))


def test_expect_to_precision_loop():
    st = FakeState(n_chains=16, chain_length=8)
    acc = conv.expect_to_precision(st, "H", atol=0.05, max_iter=1000, verbose=False)
    assert acc.error_of_mean <= 0.05 and abs(acc.mean + 3.0) < 0.3
    # first batch with the default discards (None), every further one with n_discard = 0; stops as soon as the error is met
    assert st.calls[0] == ("H", None, 4) and all(c[1] == 0 for c in st.calls[1:])
    assert acc.n_samples == len(st.calls) * 16 * 8
    again = FakeState(n_chains=16, chain_length=8)
    short = conv.expect_to_precision(again, "H", atol=1e-9, max_iter=3, verbose=False)
    assert len(again.calls) == 4 and short.n_samples == 4 * 16 * 8              # max_iter more after the first batch
    both = conv.expect_to_precision(FakeState(n_chains=16, chain_length=8), "H", atol=10.0, rtol=0.004, max_iter=1000, verbose=False)
    assert both.error_of_mean / abs(both.mean) <= 0.004                          # both tolerances must hold
    # several operators: every one is iterated until it has converged, results keep the container
    multi = FakeState(n_chains=16, chain_length=8)
    out = conv.expect_to_precision(multi, {"a": "A", "b": "B"}, atol=0.05, max_iter=1000, verbose=False)
    assert set(out) == {"a", "b"} and all(v.error_of_mean <= 0.05 for v in out.values())
    lst = conv.expect_to_precision(FakeState(n_chains=16, chain_length=8), ["A", "B"], atol=0.2, max_iter=50, verbose=False)
    assert isinstance(lst, list) and len(lst) == 2


def test_expect_to_precision_verbose_reports(capsys):
    conv.expect_to_precision(FakeState(n_chains=16, chain_length=8), "H", atol=1e-9, max_iter=2, verbose=True)
    text = capsys.readouterr()
    assert "Reached max_iter before target precision." in text.out + text.err and "[done] error" in text.out + text.err


def test_thermalise_loop_and_failure_modes():
    st = FakeState(n_chains=16, chain_length=4, spread=3.0)        # chains start far apart and relax
    stats, hist = conv.thermalise_mcmc(st, "O", min_chain_length=12, max_chain_length=400, rhat_tol=1.05, patience=2, verbose=False)
    assert stats.max_lag == 0 and stats.decay == 0.9 and stats.R_hat < 1.05
    r = hist["R_hat"]
    assert r.iters[0] == 4 and r.iters == [4 * (i + 1) for i in range(len(r))] and len(r) == len(st.calls)
    assert r.values[0] > 1.05 and r.values[-1] < 1.05 and r.values[-2] < 1.05  # patience = 2 consecutive good batches
    assert all(c[1] == 0 for c in st.calls)                                     # never discards: the chains are being advanced
    assert len(r) >= math.ceil(12 / 4)
    bad = FakeState(n_chains=16, chain_length=4, spread=3.0)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        conv.thermalise_mcmc(bad, "O", max_chain_length=12, rhat_tol=0.5, verbose=False)
    assert any("without converging" in str(x.message) for x in w) and len(bad.calls) == 12 // 4
    with pytest.raises(RuntimeError, match="without converging"):
        conv.thermalise_mcmc(FakeState(n_chains=16, chain_length=4), "O", max_chain_length=12, rhat_tol=0.5, verbose=False,
                             raise_on_failure=True)
    with pytest.raises(ValueError, match="at least 2 chains"):
        conv.thermalise_mcmc(FakeState(n_chains=1), "O", verbose=False)


def test_check_mc_convergence_doubles_the_sweep_size(capsys):
    # strongly correlated at sweep_size 1 (phi = 0.97 per step): the 32-lag window saturates, the sweep size doubles until the
    # Geyer sequence terminates inside the window; the caller's state is left alone
    st = FakeState(n_chains=32, chain_length=10, phi=0.97 ** 4, seed=1)
    sampler0, state0 = st.sampler, st.sampler_state
    stats, hist = conv.check_mc_convergence(st, "H", min_chain_length=20, max_chain_length=20000)
    assert st.sampler is sampler0 and st.sampler_state == state0 and sampler0.sweep_size == 4
    sweeps = hist["sweep_size"].values
    assert sweeps[0] == 1 and sweeps[-1] > 1 and all(b in (a, 2 * a) for a, b in zip(sweeps, sweeps[1:]))
    assert stats.max_lag == 32                                                   # thinned to 16, expanded back
    assert not oos.acf_window_saturated(stats.o) and oos.tau_corr_reliable(stats.o)
    captured = capsys.readouterr()
    assert "MC Convergence Results" in captured.out and "doubling sweep size" in captured.out + captured.err
    tau_steps = stats.tau_corr_acf * sweeps[-1]
    assert 25 < tau_steps < 130                                                  # (1 + phi) / (1 - phi) = 66 elementary steps
    # an easy chain stops after min_chain_length without any doubling
    easy = FakeState(n_chains=32, chain_length=10, phi=0.05, seed=2)
    stats2, hist2 = conv.check_mc_convergence(easy, "H", min_chain_length=50, max_chain_length=5000)
    assert set(hist2["sweep_size"].values) == {1} and 1.5 < stats2.tau_corr_acf < 4.5   # phi = 0.05 ** (1 / 4) per step: tau = 2.8
    with pytest.raises(NotImplementedError):
        conv.check_mc_convergence(easy, "H", plot=True)
