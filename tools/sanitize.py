"""Small launches of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool racecheck python tools/sanitize.py [case indices]
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_b200 as nk


def run(dtype, rule, L, n_dim, alpha, op_kind, path=0, B=40, CL=2, grad=False, probs=None):
    g = nk.graph.Hypercube(L, n_dim)
    N = g.n_nodes
    hi = nk.hilbert.Spin(0.5, N, total_sz=0 if rule == "exchange" else None)
    op = nk.operator.Ising(hi, g, h=2.0) if op_kind == "ising" else nk.operator.Heisenberg(hi, g)
    sa = (nk.sampler.MetropolisLocal(hi, n_chains=B) if rule == "local"
          else nk.sampler.MetropolisExchange(hi, graph=g, d_max=2, n_chains=B, probabilities=probs))
    model = nk.models.RBM(alpha=alpha, param_dtype=dtype)
    var = model.init(1234, N)
    st = sa.init_state(model, var, seed=3)
    samples, _, eloc, st = sa._launch(model, var, st, CL, n_discard=1, operator=op, path=path)
    vs = nk.vqs.MCState(sa, model, variables=var, n_samples=B * CL, sampler_seed=3)
    e2 = vs._eloc_on_samples(op, samples)
    if grad:
        vs.expect_and_grad(op)
    torch.cuda.synchronize()
    print(f"ok {np.dtype(dtype).name} {rule} N={N} M={int(alpha * N)} {op_kind} path={path}: E = {eloc.mean().item():.4f} / {e2.mean().item():.4f}")


CASES = [
    lambda: run(np.float32, "local", 6, 2, 4, "ising", grad=True),  # sweep_fast + theta tcgen05 + forces tcgen05
    lambda: run(np.float32, "local", 6, 2, 4, "ising", path=3),  # sweep_prod f32 local
    lambda: run(np.float64, "local", 6, 2, 4, "ising", grad=True),  # sweep_prod f64 local + theta DMMA + forces DMMA
    lambda: run(np.float64, "exchange", 6, 2, 2, "heis"),  # sweep_prod f64 exchange + LocalOperator E_loc
    lambda: run(np.float32, "exchange", 12, 1, 2, "heis"),  # sweep_prod f32 exchange
    lambda: run(np.float32, "local", 20, 1, 32, "ising"),  # MULTI (2 warps per chain)
    lambda: run(np.float64, "local", 16, 1, 40, "ising"),  # MULTI fp64
    lambda: run(np.float64, "local", 14, 2, 2, "ising"),  # N = 196: rows through L2
    lambda: run(np.float64, "local", 6, 2, 4, "ising", path=1),  # theta-form kernels
    lambda: run(np.float64, "local", 10, 2, 4, "ising", B=24),  # cfg-3 shape: fp64 shadow kernel (TMA ring, producer warp, parked state)
    lambda: run(np.float32, "local", 10, 2, 4, "ising", B=24),  # cfg-3 shape: sweep_fast <3, 1>
    lambda: run(np.float64, "exchange", 6, 2, 2, "heis", probs=[0.7, 0.3]),  # weighted cluster choice on the product-form kernel
    lambda: run(np.float32, "exchange", 20, 1, 32, "heis", B=12),  # multi-warp exchange (M = 640)
    lambda: run(np.float64, "exchange", 6, 2, 32, "heis", B=12),  # multi-warp exchange fp64 (M = 1152)
]
only = [int(x) for x in sys.argv[1:]]  # e.g. `python tools/sanitize.py 9 10`: only those cases
for idx, case in enumerate(CASES):
    if not only or idx in only:
        case()
