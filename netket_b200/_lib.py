"""ctypes binding of libnkb200.so (include/nkb200.h).

There is deliberately NO fallback: if the CUDA library is missing or a call fails, an exception is
raised.  torch is used only for device memory / streams (``tensor.data_ptr()``,
``torch.cuda.current_stream()``).
"""

import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NKB200_LIB") or os.path.join(_HERE, "lib", "libnkb200.so")  # env override: developer A/B builds

NK_F32, NK_F64 = 0, 1
NK_RULE_LOCAL, NK_RULE_EXCHANGE = 0, 1
NK_PATH_AUTO, NK_PATH_GENERIC, NK_PATH_FAST, NK_PATH_PROD = 0, 1, 2, 3
NK_STATS_NPARTIAL = 8
NK_ONLINE_NSUM, NK_ONLINE_NOUT, NK_ONLINE_MAX_LAG = 4, 9, 4096


class NkError(RuntimeError):
    pass


class nk_rbm_t(C.Structure):
    _fields_ = [("W", C.c_void_p), ("b", C.c_void_p), ("a", C.c_void_p), ("N", C.c_int32), ("M", C.c_int32),
                ("dtype", C.c_int32), ("reserved", C.c_int32)]


class nk_ising_t(C.Structure):
    _fields_ = [("edges", C.c_void_p), ("n_edges", C.c_int32), ("reserved", C.c_int32), ("h", C.c_double), ("J", C.c_double)]


class nk_localop_group_t(C.Structure):
    _fields_ = [("n_ops", C.c_int32), ("n_sites", C.c_int32), ("ncmax", C.c_int32), ("reserved", C.c_int32),
                ("acting_on", C.c_void_p), ("diag_mels", C.c_void_p), ("n_conns", C.c_void_p), ("mels", C.c_void_p),
                ("x_prime", C.c_void_p)]


class nk_localop_t(C.Structure):
    _fields_ = [("groups", nk_localop_group_t * 2), ("n_groups", C.c_int32), ("nonzero_diagonal", C.c_int32),
                ("max_conn_size", C.c_int32), ("reserved", C.c_int32), ("constant", C.c_double), ("mel_cutoff", C.c_double)]


class nk_chains_t(C.Structure):
    _fields_ = [("sigma", C.c_void_p), ("log_prob", C.c_void_p), ("n_accepted", C.c_void_p), ("workspace", C.c_void_p),
                ("B", C.c_int64), ("seed", C.c_uint64), ("t", C.c_uint64), ("chain_offset", C.c_uint64)]


class nk_online_stats_t(C.Structure):
    _fields_ = [("chain_count", C.c_void_p), ("chain_mean", C.c_void_p), ("chain_M2", C.c_void_p), ("cross_sum", C.c_void_p),
                ("m1_sum", C.c_void_p), ("m2_sum", C.c_void_p), ("pair_count", C.c_void_p), ("chain_buf", C.c_void_p),
                ("n_chains", C.c_int64), ("max_lag", C.c_int32), ("buf_len", C.c_int32)]


class nk_sweep_t(C.Structure):
    _fields_ = [("rule", C.c_int32), ("chain_length", C.c_int32), ("n_discard", C.c_int32), ("sweep_size", C.c_int32),
                ("machine_pow", C.c_double), ("samples_out", C.c_void_p), ("logp_out", C.c_void_p),
                ("stream_w0", C.c_void_p), ("stream_u", C.c_void_p), ("clusters", C.c_void_p), ("n_clusters", C.c_int32),
                ("path", C.c_int32), ("ising", C.POINTER(nk_ising_t)), ("localop", C.POINTER(nk_localop_t)),
                ("eloc_out", C.c_void_p), ("eloc_dtype", C.c_int32), ("flags", C.c_int32), ("tanh_out", C.c_void_p),
                ("stats_out", C.c_void_p), ("stats_shift", C.c_double), ("cluster_probs", C.c_void_p)]


class nk_ctx_desc_t(C.Structure):
    _fields_ = [("device", C.c_int32), ("N", C.c_int32), ("M", C.c_int32), ("dtype", C.c_int32), ("n_chains", C.c_int64),
                ("chain_length", C.c_int32), ("sweep_size", C.c_int32), ("rule", C.c_int32), ("n_clusters", C.c_int32),
                ("clusters_host", C.c_void_p), ("cluster_probs_host", C.c_void_p), ("machine_pow", C.c_double),
                ("n_down", C.c_int32), ("return_samples", C.c_int32), ("ising_host", C.POINTER(nk_ising_t)),
                ("localop_host", C.POINTER(nk_localop_t)), ("seed", C.c_uint64), ("chain_offset", C.c_uint64), ("stream", C.c_void_p),
                ("eloc_in_param_dtype", C.c_int32), ("reserved", C.c_int32)]


NK_SWEEP_NO_HANDOVER = 1
NK_CTX_NPARTIAL = NK_STATS_NPARTIAL + 2
NK_RESHIFT = 1

# every symbol include/nkb200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "nk_last_error": (C.c_char_p, []),
    "nk_version": (C.c_int, []),
    "nk_launch_count": (C.c_longlong, []),
    "nk_rbm_logpsi": (C.c_int, [C.c_void_p, C.POINTER(nk_rbm_t), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nk_theta_gemm_workspace_bytes": (C.c_int64, [C.POINTER(nk_rbm_t), C.c_int64]),
    "nk_theta_gemm": (C.c_int, [C.c_void_p, C.POINTER(nk_rbm_t), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nk_random_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64]),
    "nk_sweep_workspace_bytes": (C.c_int64, [C.POINTER(nk_rbm_t), C.c_int64]),
    "nk_sweep": (C.c_int, [C.c_void_p, C.POINTER(nk_rbm_t), C.POINTER(nk_chains_t), C.POINTER(nk_sweep_t)]),
    "nk_ising_conn": (C.c_int, [C.c_void_p, C.POINTER(nk_ising_t), C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_int32]),
    "nk_ising_n_conn": (C.c_int, [C.c_void_p, C.POINTER(nk_ising_t), C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "nk_localop_conn": (C.c_int, [C.c_void_p, C.POINTER(nk_localop_t), C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_int32, C.c_void_p]),
    "nk_eloc_ising_rbm": (C.c_int, [C.c_void_p, C.POINTER(nk_rbm_t), C.POINTER(nk_ising_t), C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_int32, C.c_int32, C.c_void_p]),
    "nk_eloc_localop_rbm": (C.c_int, [C.c_void_p, C.POINTER(nk_rbm_t), C.POINTER(nk_localop_t), C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "nk_forces_workspace_bytes": (C.c_int64, [C.POINTER(nk_rbm_t), C.c_int64]),
    "nk_forces_rbm": (C.c_int, [C.c_void_p, C.POINTER(nk_rbm_t), C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_double,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "nk_forces_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_int64, C.c_void_p, C.c_int32]),
    "nk_stats_partial": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_double, C.c_void_p]),
    "nk_stats_finalize": (C.c_int, [C.POINTER(C.c_double), C.c_double, C.c_int64, C.c_int64, C.POINTER(C.c_double)]),
    "nk_stats_tau": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_double, C.c_void_p]),
    "nk_stats_tau_max_decode": (C.c_double, [C.c_double]),
    "nk_rbm_tanh_theta": (C.c_int, [C.c_void_p, C.POINTER(nk_rbm_t), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nk_rbm_jvp": (C.c_int, [C.c_void_p, C.POINTER(nk_rbm_t), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_void_p]),
    "nk_online_stats_update": (C.c_int, [C.c_void_p, C.POINTER(nk_online_stats_t), C.POINTER(nk_online_stats_t), C.c_void_p, C.c_int32,
                                         C.c_int64, C.c_double]),
    "nk_online_stats_summary": (C.c_int, [C.c_void_p, C.POINTER(nk_online_stats_t), C.c_int32, C.c_double, C.c_double, C.c_void_p]),
    "nk_online_stats_finalize": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64, C.c_int64, C.c_int32,
                                           C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "nk_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_uint64, C.c_uint64]),
    "nk_ctx_create2": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(nk_ctx_desc_t)]),
    "nk_ctx_destroy": (None, [C.c_void_p]),
    "nk_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "nk_ctx_partials_device": (C.c_void_p, [C.c_void_p]),
    "nk_ctx_step_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "nk_ctx_step_end": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]),
    "nk_ctx_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                   C.POINTER(C.c_double)]),
    "nk_ctx_get_sigma_host": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nk_microbench": (C.c_int, [C.c_int32, C.POINTER(C.c_double)]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built (python -m netket_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NkError(
                f"{LIB_PATH} is missing: build it with `python -m netket_b200.build` (nvcc, sm_100a). "
                "netket_b200 has no CPU or eager fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise NkError(f"libnkb200 error {rc}: {lib().nk_last_error().decode()}")


def dtype_code(dtype):
    if isinstance(dtype, torch.dtype):
        if dtype == torch.float32:
            return NK_F32
        if dtype == torch.float64:
            return NK_F64
    else:
        d = np.dtype(dtype)
        if d == np.float32:
            return NK_F32
        if d == np.float64:
            return NK_F64
    raise TypeError(f"unsupported dtype {dtype} (float32 / float64 only)")


def torch_dtype(dtype):
    if isinstance(dtype, torch.dtype):
        return dtype
    return {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.int8): torch.int8,
            np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}[np.dtype(dtype)]


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise NkError(f"{name} must be a CUDA tensor (netket_b200 has no CPU path)")
    if not t.is_contiguous():
        raise NkError(f"{name} must be contiguous")
    return t
