"""The host-buffer C ABI (nk_ctx_create / nk_ctx_step_host / nk_ctx_get_sigma_host / nk_ctx_destroy: what bench.py's `e2e`
times; mirrors `vs.parameters = ...; vs.reset(); vs.expect(H)` of netket/vqs/mc/mc_state/state.py:514-576,695-712) against
the oracle: host parameters in, E_loc + statistics + acceptance out."""

import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import estimators as oest
from oracle import graph as ograph
from oracle import hilbert as ohilbert
from oracle import operators as oops
from oracle import rbm as orbm
from oracle import sampler as osampler

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_ctx_step_host_matches_oracle(cuda, dtype):
    from netket_b200 import _lib

    L = _lib.lib()
    Ls, N, alpha, B, CL, n_discard, h, J, seed, off = 4, 16, 2, 24, 5, 2, 1.7, 1.0, 15324, 3
    M = alpha * N
    e, _ = ograph.hypercube_edges(Ls, 2)
    e32 = np.ascontiguousarray(e, dtype=np.int32)
    W, b, a = orbm.init_params(N, alpha, seed=1234, std=0.2, dtype=dtype)
    ctx = C.c_void_p()
    _lib.check(L.nk_ctx_create(C.byref(ctx), 0, N, M, _lib.dtype_code(np.dtype(dtype)), B, CL, e32.ctypes.data_as(C.c_void_p), e32.shape[0],
                               h, J, seed, off))
    try:
        eloc = np.empty((B, CL), dtype=dtype)
        stats = (C.c_double * 6)()
        sig_host = np.empty((B, N), dtype=np.int8)
        sig0 = ohilbert.random_state(seed, B, N, None, chain_offset=off)
        _lib.check(L.nk_ctx_get_sigma_host(ctx, sig_host.ctypes.data_as(C.c_void_p)))
        assert np.array_equal(sig_host, sig0)
        W64, b64, a64 = W.astype(np.float64), b.astype(np.float64), a.astype(np.float64)
        t0, sigma = 0, sig0
        for step, nd in enumerate((n_discard, 0)):  # two steps: the second continues the chains and the Philox counter
            _lib.check(L.nk_ctx_step_host(ctx, W.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
                                          nd, eloc.ctypes.data_as(C.c_void_p), stats))
            _lib.check(L.nk_ctx_get_sigma_host(ctx, sig_host.ctypes.data_as(C.c_void_p)))
            if dtype == np.float64:
                ref = osampler.sample_chain("local", sigma, W64, b64, a64, chain_length=nd + CL, seed=seed, t0=t0, chain_offset=off)
                samples = ref["samples"][:, nd:]
                assert np.array_equal(sig_host, ref["sigma"])
                eref = oest.local_estimators(samples, lambda x: oops.ising_conn_padded(x, e, h, J), W64, b64, a64)
                np.testing.assert_allclose(eloc, eref, rtol=1e-11, atol=1e-11 * np.abs(eref).max())
                np.testing.assert_allclose(stats[5], ref["n_accepted"].sum() / ref["n_steps"], rtol=1e-12)
                t0, sigma = ref["t"], ref["sigma"]
            ost = oracle.stats.statistics(eloc.astype(np.float64))
            for k, name in enumerate(("mean", "error_of_mean", "variance", "tau_corr", "R_hat")):
                np.testing.assert_allclose(stats[k], ost[name], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=name)
            assert 0.0 < stats[5] <= 1.0 and set(np.unique(sig_host)) <= {-1, 1}
    finally:
        L.nk_ctx_destroy(ctx)


def test_ctx_rejects_bad_arguments(cuda):
    from netket_b200 import _lib

    L = _lib.lib()
    ctx = C.c_void_p()
    assert L.nk_ctx_create(C.byref(ctx), 0, 0, 4, 0, 8, 2, None, 0, 1.0, 1.0, 0, 0) == -1
    assert L.nk_ctx_create(None, 0, 4, 4, 0, 8, 2, None, 0, 1.0, 1.0, 0, 0) == -1
    assert L.nk_ctx_create(C.byref(ctx), 0, 4, 4, 7, 8, 2, None, 0, 1.0, 1.0, 0, 0) == -1
    assert b"dtype" in L.nk_last_error()
