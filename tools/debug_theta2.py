import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from netket_b200 import _lib
L = _lib.lib()
def run(W, sig, b=None):
    N, M = W.shape; B = sig.shape[0]
    Wt, st = torch.from_numpy(W.astype(np.float32)).cuda(), torch.from_numpy(sig.astype(np.int8)).cuda()
    bt = torch.from_numpy(b.astype(np.float32)).cuda() if b is not None else None
    rbm = _lib.nk_rbm_t(W=Wt.data_ptr(), b=bt.data_ptr() if b is not None else None, a=None, N=N, M=M, dtype=0, reserved=0)
    ws = torch.empty(int(L.nk_theta_gemm_workspace_bytes(C.byref(rbm), B)), dtype=torch.uint8, device="cuda")
    th = torch.full((B, M), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(L.nk_theta_gemm(_lib.stream_ptr(), C.byref(rbm), _lib.ptr(st), B, _lib.ptr(th), _lib.ptr(ws)))
    torch.cuda.synchronize()
    return th.cpu().numpy()
rs = np.random.default_rng(0)
for (N, M, B, bias) in [(16, 64, 128, True), (16, 64, 128, False), (100, 400, 128, False), (100, 208, 128, False), (100, 64, 128, False), (48, 64, 128, False), (100, 400, 1000, True), (16, 64, 300, False)]:
    W = (rs.normal(size=(N, M)) * 0.3).astype(np.float32)
    b = rs.normal(size=M).astype(np.float32) if bias else None
    sg = (1 - 2 * rs.integers(0, 2, size=(B, N))).astype(np.int8)
    t = run(W, sg, b); ref = sg.astype(np.float64) @ W.astype(np.float64) + (b.astype(np.float64) if bias else 0)
    err = np.abs(t - ref)
    print((N, M, B, bias), "max err", err.max(), "bad rows", np.unique(np.argwhere(err > 1e-4)[:, 0])[:10], "bad cols", np.unique(np.argwhere(err > 1e-4)[:, 1])[:10])
