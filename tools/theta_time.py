"""Developer probe: time nk_theta_gemm and print on-chip peak rates (nk_microbench)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_b200 as nk
from netket_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda", 0)
for which, name in ((0, "smem GB/s"), (1, "L2 GB/s"), (2, "fp32 GFLOP/s"), (3, "MUFU Gop/s"), (4, "fp64 GFLOP/s")):
    r = C.c_double(); _lib.check(L.nk_microbench(which, C.byref(r))); print(name, r.value)
for dtype in (np.float32, np.float64):
    B, N, alpha = 65536, 100, 4
    model = nk.models.RBM(alpha=alpha, param_dtype=dtype)
    var = model.init(1234, N, device=dev)
    rbm = nk.models.RBM.c_struct(var)
    sig = torch.randint(0, 2, (B, N), device=dev, dtype=torch.int8) * 2 - 1
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    theta = torch.empty((B, N * alpha), dtype=tdt, device=dev)
    ws = torch.empty(max(1, int(L.nk_theta_gemm_workspace_bytes(C.byref(rbm), B))), dtype=torch.uint8, device=dev)
    def f():
        _lib.check(L.nk_theta_gemm(_lib.stream_ptr(dev), C.byref(rbm), _lib.ptr(sig), B, _lib.ptr(theta), _lib.ptr(ws)))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(dtype.__name__, "theta_gemm ms", ms, "GB/s written", B * N * alpha * theta.element_size() / ms / 1e6)
