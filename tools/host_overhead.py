"""Developer probe: host-side cost of one `vs.reset(); vs.expect(H)` step (what strong scaling at 8 GPUs is left with once the
kernels take < 2 ms): wall time per step at a small number of chains and a cProfile of the Python path.

    python tools/host_overhead.py [--chains 8192] [--steps 200]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_b200 as nk  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=8192)
ap.add_argument("--steps", type=int, default=200)
a = ap.parse_args()
g = nk.graph.Hypercube(10, 2, pbc=True)
hi = nk.hilbert.Spin(0.5, g.n_nodes)
ha = nk.operator.Ising(hi, g, h=3.0)
model = nk.models.RBM(alpha=4, param_dtype=np.float32)
vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=a.chains), model, n_samples=a.chains * 16, n_discard_per_chain=5, seed=1234,
                    sampler_seed=4321)
for _ in range(5):
    vs.reset()
    vs.expect(ha)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tic = time.perf_counter()
e0.record()
for _ in range(a.steps):
    vs.reset()
    vs.expect(ha)
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - tic) / a.steps * 1e3
print(f"chains={a.chains}: {wall:.3f} ms per step wall, {e0.elapsed_time(e1) / a.steps:.3f} ms between CUDA events")
# kernel time alone: the same launches without the host read in between
sa, st = vs.sampler, vs.sampler_state
torch.cuda.synchronize()
e0.record()
for _ in range(a.steps):
    sa._launch(vs.model, vs.variables, st, 16, operator=ha, want_samples=True, stats_shift=0.0, no_handover=True)
e1.record()
torch.cuda.synchronize()
print(f"  back-to-back launches without a host read: {e0.elapsed_time(e1) / a.steps:.3f} ms per step")
pr = cProfile.Profile()
pr.enable()
for _ in range(a.steps):
    vs.reset()
    vs.expect(ha)
pr.disable()
ps = pstats.Stats(pr).sort_stats("cumulative")
ps.print_stats(28)
