"""Examples/Ising1d/ising1d.py of the reference (BASELINE.json cfg-1) on the B200 path.

Same objects and keyword arguments as the reference script; the optimisation loop is written out because drivers and
optimisers are outside this repository's scope: plain SGD, `p <- p - lr * grad`, which is what
`nk.driver.VMC(ha, nk.optimizer.Sgd(0.02), variational_state=vs)` does without a preconditioner
(netket/driver/vmc.py:141-161).

    python examples/ising1d.py [n_iter]
"""

import sys

import torch

import netket_b200 as nk

L = 20
g = nk.graph.Hypercube(length=L, n_dim=1, pbc=True)
hi = nk.hilbert.Spin(s=1 / 2, N=g.n_nodes)
ha = nk.operator.Ising(hilbert=hi, graph=g, h=1.0)
ma = nk.models.RBM(alpha=1, use_visible_bias=True, param_dtype=float)
sa = nk.sampler.MetropolisLocal(hi, n_chains=16)
vs = nk.vqs.MCState(sa, ma, n_samples=1008, n_discard_per_chain=10, seed=0, sampler_seed=1)

lr = 0.02
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for it in range(n_iter):
    energy, grad = vs.expect_and_grad(ha)
    p = vs.parameters
    vs.parameters = {"Dense": {"kernel": p["Dense"]["kernel"] - lr * grad["Dense"]["kernel"],
                               "bias": p["Dense"]["bias"] - lr * grad["Dense"]["bias"]},
                     "visible_bias": p["visible_bias"] - lr * grad["visible_bias"]}
    if it % 25 == 0 or it == n_iter - 1:
        print(f"iter {it:4d}  E = {energy}  acceptance = {vs.sampler_state.acceptance:.3f}")
torch.cuda.synchronize()
print("exact ground-state energy of the L=20 critical chain: -25.4910 (netket.exact.lanczos_ed)")
