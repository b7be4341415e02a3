"""One update + one summary of the streaming statistics at the headline shape (2^16 chains x 16, max_lag 64) and one QGT
product on cfg-3 fp32 - the launches an `ncu --set full -k regex:online_|jvp_dot` capture looks at."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_b200 as nk
from netket_b200 import stats as nkstats
from netket_b200.optimizer import QGTOnTheFly, tree_to_flat

x = torch.randn((2 ** 16, 16), dtype=torch.float32, device="cuda")
e = nkstats.online_statistics(x, max_lag=64)
for _ in range(3):
    e.update(x, inplace=True)
e._summarise()
g = nk.graph.Hypercube(10, 2); hi = nk.hilbert.Spin(0.5, 100); H = nk.operator.Ising(hi, g, h=3.0)
vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=2 ** 16), nk.models.RBM(alpha=4, param_dtype=np.float32), n_samples=2 ** 20, seed=1)
E, G = vs.expect_and_grad(H)
S = QGTOnTheFly(vs, diag_shift=0.01)
S @ tree_to_flat(G)
torch.cuda.synchronize()
