import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_b200 as nk
g = nk.graph.Hypercube(6, 2); hi = nk.hilbert.Spin(0.5, 36); op = nk.operator.Ising(hi, g, h=3.0)
for dtype in (np.float32, np.float64):
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=1024), nk.models.RBM(alpha=4, param_dtype=dtype), n_samples=1024 * 8,
                        n_discard_per_chain=4, seed=1, sampler_seed=2)
    for it in range(200):
        e, G = vs.expect_and_grad(op)
        p = vs.parameters
        vs.parameters = {"Dense": {"kernel": p["Dense"]["kernel"] - 0.02 * G["Dense"]["kernel"], "bias": p["Dense"]["bias"] - 0.02 * G["Dense"]["bias"]},
                         "visible_bias": p["visible_bias"] - 0.02 * G["visible_bias"]}
    W = vs.parameters["Dense"]["kernel"]
    e_prod = vs.expect(op)
    alone = vs._eloc_on_samples(op, vs.samples, path=1)   # theta-form kernel on the same samples
    fused = vs.local_estimators(op)
    rel = float((alone - fused).abs().max() / fused.abs().max())
    print(np.dtype(dtype).name, "after 200 SGD steps: E =", e_prod, " max|W| =", float(W.abs().max()), " rowabs =", float(W.abs().sum(1).max()),
          " E_loc product-form vs theta-form max rel diff =", rel, " acceptance =", vs.sampler_state.acceptance)
print("exact TFIM 6x6 h=3 ground energy per site ~ -3.17 -> E ~ -114")
