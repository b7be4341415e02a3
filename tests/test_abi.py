"""The C-ABI library: loads without a GPU, exports every symbol include/nkb200.h declares, validates arguments
before touching CUDA, and its host-only entry points work (nk_stats_finalize vs the oracle).  CPU only."""

import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "nkb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(nk_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if not n.endswith("_t")))


def test_header_declares_what_python_binds(lib_built):
    from netket_b200 import _lib

    declared = header_functions()
    assert set(declared) == set(_lib.SYMBOLS), (set(declared) ^ set(_lib.SYMBOLS))


def test_library_exports_every_declared_symbol(lib_built):
    L = C.CDLL(lib_built)
    for name in header_functions():
        assert hasattr(L, name), f"{name} declared in include/nkb200.h but not exported"


def test_no_torch_or_python_in_the_abi(lib_built):
    """plain C boundary: the shared object must not link against torch / python."""
    import subprocess

    out = subprocess.run(["ldd", lib_built], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out, out


def test_version_and_error_string(lib_built):
    from netket_b200 import _lib

    L = _lib.lib()
    assert L.nk_version() >= 100
    assert isinstance(L.nk_last_error(), bytes)


def test_argument_validation_happens_before_cuda(lib_built):
    """Bad arguments return NK_EINVAL with a message; no CUDA call is needed for that (runs on the CPU-only box)."""
    from netket_b200 import _lib

    L = _lib.lib()
    rbm = _lib.nk_rbm_t(W=None, b=None, a=None, N=4, M=4, dtype=0, reserved=0)
    assert L.nk_rbm_logpsi(None, C.byref(rbm), None, 1, None, None) == -1
    assert b"rbm.W is NULL" in L.nk_last_error()
    rbm = _lib.nk_rbm_t(W=1, b=None, a=None, N=4, M=4, dtype=7, reserved=0)
    assert L.nk_rbm_logpsi(None, C.byref(rbm), None, 1, None, None) == -1
    assert b"dtype" in L.nk_last_error()
    rbm = _lib.nk_rbm_t(W=1, b=None, a=None, N=4, M=4, dtype=0, reserved=0)
    ch = _lib.nk_chains_t(sigma=1, log_prob=1, n_accepted=1, workspace=None, B=2, seed=0, t=0, chain_offset=0)
    a = _lib.nk_sweep_t()
    a.rule, a.chain_length, a.n_discard, a.sweep_size, a.machine_pow = 5, 1, 0, 4, 2.0
    assert L.nk_sweep(None, C.byref(rbm), C.byref(ch), C.byref(a)) == -1
    assert b"only LocalRule and ExchangeRule" in L.nk_last_error()
    a.rule, a.machine_pow = 0, -1.0
    assert L.nk_sweep(None, C.byref(rbm), C.byref(ch), C.byref(a)) == -1
    assert b"machine_pow" in L.nk_last_error()
    a.machine_pow, a.rule = 2.0, 1
    assert L.nk_sweep(None, C.byref(rbm), C.byref(ch), C.byref(a)) == -1
    assert b"ExchangeRule needs clusters" in L.nk_last_error()
    assert L.nk_random_state(None, None, 4, 8, 9, 0, 0) == -1
    assert L.nk_stats_partial(None, None, 0, 4, 4, 3, 0.0, None) == -1
    op = _lib.nk_localop_t()
    op.n_groups = 1
    op.groups[0].n_sites = 3
    assert L.nk_localop_conn(None, C.byref(op), None, 1, 4, None, None, 1, None) == -1
    assert b"supported: 1, 2" in L.nk_last_error()
    assert L.nk_sweep_workspace_bytes(C.byref(rbm), 10) > 0
    # entry points added for the stand-alone estimator and the forces: argument validation happens before any CUDA call
    a.rule, a.path = 0, 7
    assert L.nk_sweep(None, C.byref(rbm), C.byref(ch), C.byref(a)) == -1
    assert b"bad path" in L.nk_last_error()
    ising = _lib.nk_ising_t(edges=None, n_edges=0, reserved=0, h=1.0, J=1.0)
    assert L.nk_eloc_ising_rbm(None, C.byref(rbm), C.byref(ising), 1, 2, 1, 5, 0, None) == -1
    assert b"eloc_dtype" in L.nk_last_error()
    assert L.nk_eloc_ising_rbm(None, C.byref(rbm), C.byref(ising), 1, 2, 1, 1, 9, None) == -1
    assert b"bad path" in L.nk_last_error()
    assert L.nk_forces_rbm(None, C.byref(rbm), 1, 4, 1, 1, 0.0, None, 1, None) == -1
    assert L.nk_forces_rbm(None, C.byref(rbm), None, 4, 1, 1, 0.0, 1, 1, None) == -1
    assert L.nk_forces_rbm(None, C.byref(rbm), 1, 4, 1, 1, 0.0, 1, None, None) == -1  # neither workspace nor tanh(theta)
    assert b"NULL buffer" in L.nk_last_error()
    assert L.nk_forces_finalize(None, None, 1.0, 4, None, 0) == -1
    assert L.nk_forces_workspace_bytes(C.byref(rbm), 1000) >= 1000 * 4 * 4
    big = _lib.nk_rbm_t(W=1, b=None, a=None, N=400, M=3200, dtype=0, reserved=0)
    assert L.nk_sweep_workspace_bytes(C.byref(big), 8) >= 8 * 3200 * 4 + 400 * 3200 * 4  # theta + the G table (several warps per chain)
    # streaming statistics
    st = _lib.nk_online_stats_t(1, 1, 1, None, None, None, None, None, 4, 8, 0)
    assert L.nk_online_stats_update(None, C.byref(st), C.byref(st), 1, 1, 4, 1.0) == -1
    assert b"NULL lag array" in L.nk_last_error()
    st = _lib.nk_online_stats_t(1, 1, 1, 1, 1, 1, 1, 1, 4, 8, 9)
    assert L.nk_online_stats_update(None, C.byref(st), C.byref(st), 1, 1, 4, 1.0) == -1
    assert b"buf_len" in L.nk_last_error()
    st = _lib.nk_online_stats_t(1, 1, 1, 1, 1, 1, 1, 1, 4, 8, 0)
    assert L.nk_online_stats_update(None, C.byref(st), C.byref(st), 1, 1, 0, 1.0) == -1
    assert b"at least one sample" in L.nk_last_error()
    assert L.nk_online_stats_update(None, C.byref(st), C.byref(st), 1, 1, 4, 1.5) == -1
    assert b"decay" in L.nk_last_error()
    assert L.nk_online_stats_update(None, C.byref(st), C.byref(st), 1, 3, 4, 1.0) == -1
    other = _lib.nk_online_stats_t(1, 1, 1, 1, 1, 1, 1, 1, 4, 16, 0)
    assert L.nk_online_stats_update(None, C.byref(st), C.byref(other), 1, 1, 4, 1.0) == -1
    assert b"shapes differ" in L.nk_last_error()
    wide = _lib.nk_online_stats_t(1, 1, 1, 1, 1, 1, 1, 1, 4, 5000, 0)
    assert L.nk_online_stats_summary(None, C.byref(wide), 0, 0.0, 0.0, 1) == -1
    assert b"max_lag" in L.nk_last_error()
    assert L.nk_online_stats_summary(None, C.byref(st), 2, 0.0, 0.0, 1) == -1
    assert L.nk_online_stats_finalize(None, None, 1, 1, 0, None, None) == -1


def _partials(x, mu):
    """What nk_stats_partial(phase=1) produces, in NumPy (the GPU kernel is tested against the oracle in -m gpu)."""
    n_chains, L = x.shape
    d = x - mu
    lb = max(1, L // 32)
    nb = L // lb
    blocks = d[:, : nb * lb].reshape(n_chains, nb, lb).mean(axis=2)
    half = L // 2
    halves = d[:, : 2 * half].reshape(n_chains, 2, half).mean(axis=2) if half > 0 else np.zeros((n_chains, 0))
    m = d.mean(axis=1)
    return np.array([np.sum(d * d), m.sum(), (m * m).sum(), blocks.sum(), (blocks ** 2).sum(), halves.sum(), (halves ** 2).sum(),
                     d.sum()])


@pytest.mark.parametrize("shape", [(16, 63), (64, 100), (1, 1000), (40, 1), (5, 7), (128, 64), (33, 65)])
def test_stats_finalize_matches_oracle(lib_built, shape):
    import oracle
    from netket_b200 import stats as nkstats

    rs = np.random.default_rng(0)
    x = rs.normal(size=shape).cumsum(axis=1) * 0.2 + rs.normal(size=shape) - 7.0
    mean = x.mean()
    for shift in (mean, mean + 0.37, 0.0):  # any shift gives the same statistics (the one-pass, in-kernel reduction uses an estimate)
        st = nkstats.finalize(_partials(x, shift), shift, shape[0], shape[1])
        ref = oracle.stats.statistics(x)
        for k in ("mean", "variance", "error_of_mean", "tau_corr", "R_hat"):
            np.testing.assert_allclose(getattr(st, k), ref[k], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=k)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from netket_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.NkError, match="no CPU or eager fallback"):
        _lib.lib()
