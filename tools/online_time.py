"""Times the streaming-statistics kernels at the headline shape (2^16 chains x 16 samples per batch, max_lag 64) and one
iteration of expect_to_precision on cfg-3.  CUDA events on torch's current stream (the library launches there)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_b200 as nk
from netket_b200 import stats as nkstats


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for dtype in (torch.float32, torch.float64):
    for C_, n, L in [(2 ** 16, 16, 64), (2 ** 16, 16, 32), (2 ** 16, 16, 0), (2 ** 14, 64, 64), (2 ** 10, 1024, 64)]:
        x = torch.randn((C_, n), dtype=dtype, device="cuda")
        e = nkstats.online_statistics(x, max_lag=L)
        t_up = timed(lambda: e.update(x, inplace=True))
        def summ():
            e._summary = None
            e._summarise()
        t_su = timed(summ)
        words = 3 + (4 * (L + 1) + L if L else 0)
        byt = C_ * (2 * words * 8 + n * x.element_size())
        print(f"{str(dtype):14s} chains={C_:6d} n={n:5d} max_lag={L:3d}: update {t_up * 1e3:8.1f} us ({byt / t_up / 1e6:7.1f} GB/s of state + batch), "
              f"summary (2 kernels + 2 host reads) {t_su * 1e3:8.1f} us")

g = nk.graph.Hypercube(10, 2)
hi = nk.hilbert.Spin(0.5, 100)
H = nk.operator.Ising(hi, g, h=3.0)
vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=2 ** 16), nk.models.RBM(alpha=4, param_dtype=np.float32), n_samples=2 ** 20, seed=1)
t0 = time.perf_counter()
acc = vs.expect_to_precision(H, atol=1e-9, max_iter=10, verbose=False)
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"expect_to_precision cfg-3 fp32: 11 batches of 2^20 samples in {(t1 - t0) * 1e3:.1f} ms = {(t1 - t0) / 11 * 1e3:.2f} ms per batch "
      f"(sweep + E_loc + accumulator + stopping test); {acc}  tau_acf={acc.tau_corr_acf:.2f}")
st, hist = vs.thermalise(H, verbose=False)
print("thermalise:", st, "R_hat", st.R_hat, "batches", len(hist["R_hat"]))
