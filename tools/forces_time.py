"""Developer probe: time of one VMC step with gradient (vs.reset(); vs.expect_and_grad(H)) and of nk_forces_rbm alone, cfg-3."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_b200 as nk
for dtype in (np.float32, np.float64):
    g = nk.graph.Hypercube(10, 2); hi = nk.hilbert.Spin(0.5, 100); op = nk.operator.Ising(hi, g, h=3.0)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=1 << 16), nk.models.RBM(alpha=4, param_dtype=dtype),
                        n_samples=(1 << 16) * 16, n_discard_per_chain=0, seed=1234, sampler_seed=15324)
    vs.sample(n_discard_per_chain=5)
    for fn, name in ((lambda: (vs.reset(), vs.expect(op)), "expect"), (lambda: (vs.reset(), vs.expect_and_grad(op)), "expect_and_grad")):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        print(np.dtype(dtype).name, name, e0.elapsed_time(e1) / 5, "ms/step")

# ---- the forces call alone (theta GEMM over all samples + contraction), tensor-core vs CUDA-core contraction
import ctypes as C, subprocess
from netket_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda", 0)
for dtype in (np.float32, np.float64):
    Ns, N, M = 1 << 20, 100, 400
    model = nk.models.RBM(alpha=4, param_dtype=dtype)
    var = model.init(1234, N, device=dev)
    rbm = nk.models.RBM.c_struct(var)
    sig = (torch.randint(0, 2, (Ns, N), device=dev, dtype=torch.int8) * 2 - 1)
    eloc = torch.randn(Ns, dtype=torch.float64, device=dev)
    ws = torch.empty(int(L.nk_forces_workspace_bytes(C.byref(rbm), Ns)), dtype=torch.uint8, device=dev)
    sums = torch.empty(N * M + M + N, dtype=torch.float64, device=dev)
    def f():
        _lib.check(L.nk_forces_rbm(_lib.stream_ptr(dev), C.byref(rbm), _lib.ptr(sig), Ns, _lib.ptr(eloc), 1, 0.0, _lib.ptr(sums), _lib.ptr(ws), None))
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    print(np.dtype(dtype).name, "nk_forces_rbm (theta GEMM + contraction), 2^20 samples:", e0.elapsed_time(e1) / 5, "ms",
          "[NKB200_FORCES_CUDA_CORE=" + str(os.environ.get("NKB200_FORCES_CUDA_CORE")) + "]")
