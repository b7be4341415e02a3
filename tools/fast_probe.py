"""Developer probe: throughput and accuracy of the sweep(+E_loc) kernels on one GPU.

    python tools/fast_probe.py [--chains 65536] [--cl 16] [--std 0.01] [--path 0|1|2] [--dtype float32] [--reps 5]

Prints samples/s (CUDA events on the launching stream), acceptance, and the deviation of the kernel's running
log_prob / E_loc from the oracle on a few chains.  Not part of the product or of bench.py.
"""

import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import netket_b200 as nk  # noqa: E402
from oracle import estimators as oest  # noqa: E402
from oracle import graph as ograph  # noqa: E402
from oracle import operators as oops  # noqa: E402
from oracle import rbm as orbm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=65536)
    ap.add_argument("--cl", type=int, default=16)
    ap.add_argument("--std", type=float, default=0.01)
    ap.add_argument("--path", type=int, default=0)
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--L", type=int, default=10)
    ap.add_argument("--alpha", type=int, default=4)
    ap.add_argument("--no-eloc", action="store_true")
    ap.add_argument("--check", type=int, default=64, help="chains checked against the oracle")
    a = ap.parse_args()
    dtype = np.dtype(a.dtype).type
    g = nk.graph.Hypercube(a.L, 2)
    N = g.n_nodes
    hi = nk.hilbert.Spin(0.5, N)
    op = nk.operator.Ising(hi, g, h=3.0)
    W, b, av = orbm.init_params(N, a.alpha, seed=1234, std=a.std, dtype=dtype)
    var = {"params": {"Dense": {"kernel": torch.from_numpy(W).cuda(), "bias": torch.from_numpy(b).cuda()},
                      "visible_bias": torch.from_numpy(av).cuda()}}
    model = nk.models.RBM(alpha=a.alpha, param_dtype=dtype)
    sa = nk.sampler.MetropolisLocal(hi, n_chains=a.chains)
    st = sa.init_state(model, var, seed=15324)
    oper = None if a.no_eloc else op
    # burn-in + warm-up
    _, _, _, st = sa._launch(model, var, st, 5, operator=oper, path=a.path, want_samples=False)
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
    print(f"path={a.path} dtype={a.dtype} chains={a.chains} cl={a.cl} std={a.std}: {ms:.3f} ms/launch (min {min(times):.3f}), "
          f"{n / ms * 1e3:.4g} samples/s, acceptance={st.acceptance:.4f}")
    try:  # -DNK_SH_PROFILE builds: cycles of one warp per phase of the fp64 shadow kernel
        import ctypes
        from netket_b200 import _lib as _L
        fn = _L.lib().nk_debug_shadow_profile
        buf = (ctypes.c_ulonglong * 16)()
        if fn(buf) == 0 and buf[7]:
            v = [int(x) for x in buf]
            ns = v[7]
            print(f"  shadow profile (cycles per sweep of one warp, {ns} sweeps): S={v[0] / ns:.0f} exact={v[1] / ns:.0f} "
                  f"({v[6] / ns:.3f}/sweep) U={v[2] / ns:.0f} (ring wait {v[8] / ns:.0f}) E+out={v[4] / ns:.0f} (ring wait {v[9] / ns:.0f})")
    except AttributeError:
        pass
    # accuracy on the first chains
    k = min(a.check, a.chains)
    s_np = samples[:k].cpu().numpy()
    W64, b64, a64 = W.astype(np.float64), b.astype(np.float64), av.astype(np.float64)
    lp_ref = 2.0 * orbm.logpsi(s_np[:, -1], W64, b64, a64)
    lp = st.log_prob[:k].cpu().numpy()
    print(f"  log_prob drift after {a.cl} sweeps: max |d| = {np.abs(lp - lp_ref).max():.3e}, mean d = {(lp - lp_ref).mean():.3e}")
    if eloc is not None:
        e, _ = ograph.hypercube_edges(a.L, 2)
        ref = oest.local_estimators(s_np, lambda x: oops.ising_conn_padded(x, e, 3.0, 1.0), W64, b64, a64)
        err = np.abs(eloc[:k].cpu().numpy() - ref) / np.abs(ref).max()
        print(f"  E_loc max rel err = {err.max():.3e} (last sweep {err[:, -1].max():.3e}), mean E = {eloc.mean().item():.4f}")


if __name__ == "__main__":
    main()
