"""Developer probe: throughput of the fused sweep + E_loc kernels on the BASELINE.json configurations other than cfg-3.

    python tools/prod_probe.py --cfg 1|2|4|5 [--chains N] [--cl K] [--dtype float64] [--path 0|1|3] [--reps 3]

cfg 1: Ising1d L=20, RBM alpha=1, MetropolisLocal;  cfg 2: Heisenberg1d L=22 total_sz=0, RBM alpha=2, MetropolisExchange;
cfg 4: J1-J2 10x10 (J2=0.5), RBM alpha=4, MetropolisExchange (d_max given by --dmax);  cfg 5: TFIM 20x20, RBM alpha=8 (N=400, M=3200).
Prints samples/s (CUDA events on the launching stream) and the acceptance.  Not part of the product or of bench.py.
"""

import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import netket_b200 as nk  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=4)
    ap.add_argument("--chains", type=int, default=16384)
    ap.add_argument("--cl", type=int, default=16)
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--path", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--dmax", type=int, default=1)
    ap.add_argument("--std", type=float, default=0.01)
    ap.add_argument("--no-eloc", action="store_true")
    a = ap.parse_args()
    dtype = np.dtype(a.dtype).type
    if a.cfg == 1:
        g = nk.graph.Hypercube(20, 1)
        hi = nk.hilbert.Spin(0.5, g.n_nodes)
        op = nk.operator.Ising(hi, g, h=1.0)
        alpha = 1
        sa = nk.sampler.MetropolisLocal(hi, n_chains=a.chains)
    elif a.cfg == 5:
        g = nk.graph.Hypercube(20, 2)
        hi = nk.hilbert.Spin(0.5, g.n_nodes)
        op = nk.operator.Ising(hi, g, h=3.0)
        alpha = 8
        sa = nk.sampler.MetropolisLocal(hi, n_chains=a.chains)
    elif a.cfg == 2:
        g = nk.graph.Hypercube(22, 1)
        hi = nk.hilbert.Spin(0.5, g.n_nodes, total_sz=0)
        op = nk.operator.Heisenberg(hi, g)
        alpha = 2
        sa = nk.sampler.MetropolisExchange(hi, graph=g, d_max=a.dmax, n_chains=a.chains)
    else:
        g = nk.graph.Hypercube(10, 2, max_neighbor_order=2)
        hi = nk.hilbert.Spin(0.5, g.n_nodes, total_sz=0)
        op = nk.operator.Heisenberg(hi, g, J=[1.0, 0.5], sign_rule=[False, False])
        alpha = 4
        sa = nk.sampler.MetropolisExchange(hi, graph=g, d_max=a.dmax, n_chains=a.chains)
    model = nk.models.RBM(alpha=alpha, param_dtype=dtype)
    rs = np.random.default_rng(1234)
    N, M = g.n_nodes, alpha * g.n_nodes
    var = {"params": {"Dense": {"kernel": torch.from_numpy((rs.normal(size=(N, M)) * a.std).astype(dtype)).cuda(),
                                "bias": torch.from_numpy((rs.normal(size=M) * a.std).astype(dtype)).cuda()},
                      "visible_bias": torch.from_numpy((rs.normal(size=N) * a.std).astype(dtype)).cuda()}}
    st = sa.init_state(model, var, seed=15324)
    oper = None if a.no_eloc else op
    _, _, _, st = sa._launch(model, var, st, 2, operator=oper, path=a.path, want_samples=False)
    torch.cuda.synchronize()
    times = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st0 = st.replace(n_steps_proc=0, n_accepted_proc=torch.zeros_like(st.n_accepted_proc))
        e0.record()
        samples, _, eloc, st = sa._launch(model, var, st0, a.cl, operator=oper, path=a.path, want_samples=True)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    n = a.chains * a.cl
    extra = f", clusters={sa.rule.clusters.shape[0]}" if hasattr(sa.rule, "clusters") else ""
    print(f"cfg={a.cfg} path={a.path} dtype={a.dtype} N={N} M={M} chains={a.chains} cl={a.cl}{extra}: {ms:.3f} ms/launch "
          f"(min {min(times):.3f}), {n / ms * 1e3:.4g} samples/s, acceptance={st.acceptance:.4f}"
          + (f", mean E = {eloc.mean().item():.5f}" if eloc is not None else ""))


if __name__ == "__main__":
    main()
