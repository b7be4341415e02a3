"""Developer probe: strong scaling of BASELINE cfg-4 (J1-J2 10x10, fp64, MetropolisExchange, 2^18 samples) and cfg-5
(TFIM 20x20, alpha=8, fp32) over the ranks of a torchrun launch (or one process).  Chains are sharded, no collective
inside the timed region; time = max over ranks (CUDA events), throughput = total samples / time.

    python -m torch.distributed.run --nproc-per-node N tools/scale_probe.py
"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, ws, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if ws > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import netket_b200 as nk


def run(name, g, hi, op, sa_fn, alpha, dtype, total_chains, cl):
    B = total_chains // ws
    sa = sa_fn(B)
    model = nk.models.RBM(alpha=alpha, param_dtype=dtype)
    var = model.init(1234, g.n_nodes)
    st = sa.init_state(model, var, seed=15324)
    _, _, _, st = sa._launch(model, var, st, 1, operator=op, want_samples=False)
    torch.cuda.synchronize()
    if ws > 1:
        dist.barrier()
    times = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, _, eloc, st = sa._launch(model, var, st, cl, operator=op, want_samples=True)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.median(times))], dtype=torch.float64, device="cuda")
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name}: ws={ws} chains/GPU={B} cl={cl}: {t.item():.2f} ms, {total_chains * cl / t.item() * 1e3:.4g} samples/s total")


g4 = nk.graph.Hypercube(10, 2, max_neighbor_order=2)
hi4 = nk.hilbert.Spin(0.5, 100, total_sz=0)
op4 = nk.operator.Heisenberg(hi4, g4, J=[1.0, 0.5], sign_rule=[False, False])
run("cfg-4 J1-J2 fp64 exchange, 2^18 samples", g4, hi4, op4, lambda B: nk.sampler.MetropolisExchange(hi4, graph=g4, d_max=1, n_chains_per_rank=B),
    4, np.float64, 1 << 14, 16)
g5 = nk.graph.Hypercube(20, 2)
hi5 = nk.hilbert.Spin(0.5, 400)
op5 = nk.operator.Ising(hi5, g5, h=3.0)
run("cfg-5 TFIM 20x20 alpha=8 fp32, 2^16 chains x 2", g5, hi5, op5, lambda B: nk.sampler.MetropolisLocal(hi5, n_chains_per_rank=B), 8, np.float32, 1 << 16, 2)
if ws > 1:
    dist.barrier(); dist.destroy_process_group()
