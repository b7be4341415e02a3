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
