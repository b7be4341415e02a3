import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from netket_b200 import _lib
from oracle import hilbert as ohilbert, rbm as orbm
N, M, B, bias = 16, 64, 128, True
rs = np.random.default_rng(7)
W = (rs.normal(size=(N, M)) * 0.3).astype(np.float32)
b = rs.normal(size=M).astype(np.float32) if bias else None
sig = ohilbert.random_state(2, B, N, None)
print(sig.dtype, sig.shape, sig.flags['C_CONTIGUOUS'], sig.strides)
Wt, st = torch.from_numpy(W).cuda(), torch.from_numpy(sig).cuda()
print(st.dtype, st.is_contiguous(), st.stride())
bt = torch.from_numpy(b).cuda() if bias else None
rbm = _lib.nk_rbm_t(W=Wt.data_ptr(), b=bt.data_ptr() if bias else None, a=None, N=N, M=M, dtype=0, reserved=0)
L = _lib.lib()
ws = torch.empty(int(L.nk_theta_gemm_workspace_bytes(C.byref(rbm), B)), dtype=torch.uint8, device="cuda")
theta = torch.full((B, M), float("nan"), dtype=torch.float32, device="cuda")
_lib.check(L.nk_theta_gemm(_lib.stream_ptr(), C.byref(rbm), _lib.ptr(st), B, _lib.ptr(theta), _lib.ptr(ws)))
ref = orbm.theta(sig, W.astype(np.float64), None if b is None else b.astype(np.float64))
got = theta.cpu().numpy()
print(np.abs(got-ref).max())
ref2 = sig.astype(np.float64) @ W.astype(np.float64) + b
print(np.abs(got-ref2).max())
