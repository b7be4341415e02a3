"""Small launches of the streaming-statistics kernels, for compute-sanitizer:

    compute-sanitizer --tool racecheck python tools/sanitize_online.py
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from netket_b200 import stats as nkstats

rs = np.random.default_rng(0)
for n_chains, max_lag, lens, dtype in [(40, 64, [16, 100, 3], np.float64), (9, 200, [50, 130], np.float32), (33, 0, [8, 8], np.float64),
                                       (5, 1000, [70], np.float64)]:
    e = None
    for n in lens:
        e = nkstats.online_statistics(torch.from_numpy(rs.normal(size=(n_chains, n)).astype(dtype)).cuda(), e, max_lag=max_lag)
    if max_lag >= 2:
        e = nkstats.expand_max_lag(nkstats.thin_acf_by_2(e), max_lag).update(torch.from_numpy(rs.normal(size=(n_chains, 7)).astype(dtype)).cuda())
    torch.cuda.synchronize()
    print(f"ok chains={n_chains} max_lag={max_lag} {np.dtype(dtype).name}: {e.get_stats()} tau_acf={e.tau_corr_acf:.3f}")

# matrix-free QGT product and one SR step (nk_rbm_jvp, nk_rbm_tanh_theta + the force contraction)
import netket_b200 as nk
for dtype in (np.float32, np.float64):
    hi = nk.hilbert.Spin(0.5, 12)
    H = nk.operator.Ising(hi, nk.graph.Chain(12), h=1.0)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=40), nk.models.RBM(alpha=2, param_dtype=dtype), n_samples=160, seed=1)
    vs.sample()
    S = nk.optimizer.QGTOnTheFly(vs, diag_shift=0.01)      # tanh(theta) recomputed
    drv = nk.driver.VMC(H, nk.optimizer.Sgd(0.05), variational_state=vs, preconditioner=nk.optimizer.SR(diag_shift=0.05))
    drv.advance(2)
    torch.cuda.synchronize()
    print(f"ok SR {np.dtype(dtype).name}: E = {drv.energy}, cg iterations {drv.preconditioner.info['n_iter']}")
