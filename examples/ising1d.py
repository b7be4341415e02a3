"""Examples/Ising1d/ising1d.py of the reference (BASELINE.json cfg-1) on the B200 path: same objects, keyword arguments and
driver calls as the reference script (SGD with the SR preconditioner, netket/driver/vmc.py, netket/optimizer/sr.py).

    python examples/ising1d.py [n_iter]
"""

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_b200 as nk  # noqa: E402

L = 20
g = nk.graph.Hypercube(length=L, n_dim=1, pbc=True)
hi = nk.hilbert.Spin(s=1 / 2, N=g.n_nodes)
ha = nk.operator.Ising(hilbert=hi, graph=g, h=1.0)
ma = nk.models.RBM(alpha=1, use_visible_bias=True, param_dtype=float)
sa = nk.sampler.MetropolisLocal(hi, n_chains=16)
op = nk.optimizer.Sgd(learning_rate=0.1)
sr = nk.optimizer.SR(diag_shift=0.1)
vs = nk.vqs.MCState(sa, ma, n_samples=1008, n_discard_per_chain=10, seed=0, sampler_seed=1)

gs = nk.driver.VMC(ha, op, variational_state=vs, preconditioner=sr)
log = nk.driver.RuntimeLog()
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 300
gs.run(n_iter=n_iter, out=log, show_progress=False)
torch.cuda.synchronize()

energies = log["Energy"]["Mean"]
for it in list(range(0, n_iter, max(1, n_iter // 10))) + [n_iter - 1]:
    print(f"iter {energies.iters[it]:4d}  E = {energies.values[it]:.4f} ± {log['Energy']['Sigma'].values[it]:.4f}")
print("final:", gs.energy, " acceptance =", f"{vs.sampler_state.acceptance:.3f}")
print("precise estimate of the final state:", vs.expect_to_precision(ha, atol=5e-3, verbose=False))
print("exact ground-state energy of the L=20 critical chain: -25.4910 (netket.exact.lanczos_ed)")
