"""Two QGT products on cfg-3 fp32 (for an ncu launch list: which kernels make up a product)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_b200 as nk
from netket_b200.optimizer import QGTOnTheFly, tree_to_flat
dtype = np.float32 if len(sys.argv) < 2 or sys.argv[1] == "f32" else np.float64
g = nk.graph.Hypercube(10, 2); hi = nk.hilbert.Spin(0.5, 100); H = nk.operator.Ising(hi, g, h=3.0)
vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=2 ** 16), nk.models.RBM(alpha=4, param_dtype=dtype), n_samples=2 ** 20, seed=1)
E, G = vs.expect_and_grad(H)
S = QGTOnTheFly(vs, diag_shift=0.01)
v = tree_to_flat(G)
torch.cuda.synchronize()
print("MARK")
for _ in range(2):
    S @ v
torch.cuda.synchronize()
