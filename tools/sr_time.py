"""Times the matrix-free QGT product and one SR-preconditioned VMC step on cfg-3 (2^16 chains x 16, N=100, M=400)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_b200 as nk
from netket_b200.optimizer import QGTOnTheFly, tree_to_flat

g = nk.graph.Hypercube(10, 2); hi = nk.hilbert.Spin(0.5, 100); H = nk.operator.Ising(hi, g, h=3.0)
for dtype in (np.float32, np.float64):
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=2 ** 16), nk.models.RBM(alpha=4, param_dtype=dtype), n_samples=2 ** 20, seed=1)
    E, G = vs.expect_and_grad(H)
    S = QGTOnTheFly(vs, diag_shift=0.01)
    v = tree_to_flat(G)
    for _ in range(3):
        S @ v
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        S @ v
    b.record(); torch.cuda.synchronize()
    t_mv = a.elapsed_time(b) / 10
    drv = nk.driver.VMC(H, nk.optimizer.Sgd(0.05), variational_state=vs, preconditioner=nk.optimizer.SR(diag_shift=0.01))
    drv.advance(1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    drv.advance(3)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{np.dtype(dtype).name}: QGT mat-vec (jvp GEMM + row dot + force contraction, 2^20 samples, 40500 parameters) {t_mv:.2f} ms; "
          f"VMC step with SR (cg tol 1e-5): {(t1 - t0) / 3 * 1e3:.1f} ms, cg iterations of the last solve {drv.preconditioner.info['n_iter']}, E = {drv.energy}")
    del S, drv, vs
    torch.cuda.empty_cache()
