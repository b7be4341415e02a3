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
