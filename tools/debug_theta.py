import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from netket_b200 import _lib
L = _lib.lib()
def run(W, sig, b=None):
    N, M = W.shape; B = sig.shape[0]
    Wt, st = torch.from_numpy(W.astype(np.float32)).cuda(), torch.from_numpy(sig.astype(np.int8)).cuda()
    rbm = _lib.nk_rbm_t(W=Wt.data_ptr(), b=None, a=None, N=N, M=M, dtype=0, reserved=0)
    ws = torch.empty(int(L.nk_theta_gemm_workspace_bytes(C.byref(rbm), B)), dtype=torch.uint8, device="cuda")
    th = torch.full((B, M), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(L.nk_theta_gemm(_lib.stream_ptr(), C.byref(rbm), _lib.ptr(st), B, _lib.ptr(th), _lib.ptr(ws)))
    torch.cuda.synchronize()
    return th.cpu().numpy()
np.set_printoptions(linewidth=250, precision=2, suppress=True)
N, M, B = 16, 64, 128
sig = np.ones((B, N), dtype=np.int8)
# test 1: W[k][j] = 1 for k == 0 only; sigma[r][0] = -1 for odd r -> theta[r][j] = +-1
W = np.zeros((N, M)); W[0, :] = 1
s1 = sig.copy(); s1[1::2, 0] = -1
t = run(W, s1); print("test1 rows 0..9, cols 0..15 (expect alternating +1/-1 rows):"); print(t[:10, :16])
# test 2: W[k][j] = j (k == 0), sigma all +1 -> theta[r][j] = j
W = np.zeros((N, M)); W[0, :] = np.arange(M)
t = run(W, sig); print("test2 row 0 (expect 0..63):"); print(t[0]); print("row 9:", t[9][:16])
# test 3: W[k][0] = k+1 (only column 0), sigma[r][k] = -1 iff k == r % 16  -> theta[r][0] = sum(k+1) - 2 (r%16 + 1) = 136 - 2(r%16+1)
W = np.zeros((N, M)); W[:, 0] = np.arange(N) + 1
s3 = sig.copy(); s3[np.arange(B), np.arange(B) % 16] = -1
t = run(W, s3); print("test3 col 0 rows 0..17 (expect 134,132,...):"); print(t[:18, 0]); print("col1 (expect 0):", t[:4, 1])
print("---- random W, N=16, no bias")
rs = np.random.default_rng(0)
W = rs.normal(size=(16, 64)) * 0.3
sg = (1 - 2 * rs.integers(0, 2, size=(128, 16))).astype(np.int8)
t = run(W, sg); ref = sg.astype(np.float64) @ W.astype(np.float32).astype(np.float64)
print("max err", np.abs(t - ref).max()); print(t[0, :8]); print(ref[0, :8])
print("---- structured N=32 (2 K-steps): W[k][0]=k+1")
N = 32
W = np.zeros((N, 64)); W[:, 0] = np.arange(N) + 1
s3 = np.ones((128, N), dtype=np.int8); s3[np.arange(128), np.arange(128) % N] = -1
t = run(W, s3); print(t[:34, 0]); print("expect", (528 - 2 * (np.arange(34) % 32 + 1)))
print("---- random W, N=32")
W = rs.normal(size=(32, 64)) * 0.3
sg = (1 - 2 * rs.integers(0, 2, size=(128, 32))).astype(np.int8)
t = run(W, sg); ref = sg.astype(np.float64) @ W.astype(np.float32).astype(np.float64)
print("max err", np.abs(t - ref).max())
print("---- W = 1/3 everywhere N=16 (needs parts 2,3)")
W = np.full((16, 64), 1.0 / 3.0); sg = np.ones((128, 16), dtype=np.int8)
t = run(W, sg); print(t[0, :4], "expect", 16 * np.float32(1 / 3))
