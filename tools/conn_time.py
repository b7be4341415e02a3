"""Developer probe: get_conn_padded kernels (HBM-write-bound) and the statistics kernel against the HBM copy bandwidth."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_b200 as nk

def timed(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

peak = None
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
a = torch.empty(1 << 28, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
ms = timed(lambda: b.copy_(a)); copy_gbs = 2 * a.numel() * 4 / ms / 1e6
print(f"device copy bandwidth here: {copy_gbs:.0f} GB/s (MEASURED_PEAKS.json hbm_gbs = {peak})")
del a, b
g = nk.graph.Hypercube(10, 2); hi = nk.hilbert.Spin(0.5, 100)
B = 1 << 16
x = hi.random_state(1, B)
op = nk.operator.Ising(hi, g, h=3.0)
ms = timed(lambda: op.get_conn_padded(x)); by = B * 101 * 100 + B * 101 * 8 + B * 100
print(f"Ising.get_conn_padded  B={B} N=100 K=101: {ms:.3f} ms, {by / ms / 1e6:.0f} GB/s of output ({by / ms / 1e6 / copy_gbs * 2:.2f} of the write half of the copy bandwidth)")
g2 = nk.graph.Hypercube(10, 2, max_neighbor_order=2); hi0 = nk.hilbert.Spin(0.5, 100, total_sz=0)
op2 = nk.operator.Heisenberg(hi0, g2, J=[1.0, 0.5], sign_rule=[False, False])
B2 = 1 << 14
x2 = hi0.random_state(1, B2)
ms = timed(lambda: op2.get_conn_padded(x2)); by = B2 * 401 * 100 + B2 * 401 * 8 + B2 * 100
print(f"J1-J2 LocalOperator.get_conn_padded  B={B2} N=100 K=401: {ms:.3f} ms, {by / ms / 1e6:.0f} GB/s of output")
e = torch.randn(B, 16, dtype=torch.float64, device="cuda")
ms = timed(lambda: nk.stats.statistics(e)); print(f"statistics  ({B} x 16 fp64, two phases + host finalize): {ms:.3f} ms")
m = nk.models.RBM(alpha=4, param_dtype=np.float32); var = m.init(1, 100)
ms = timed(lambda: m.apply(var, x)); print(f"RBM.apply fp32  B={B}: {ms:.3f} ms")
