"""Developer probe (torchrun): the host-buffer e2e loop of bench.py repeated, with and without the nvidia-smi clock sampler
running beside it - to tell a per-process effect from a transient one when e2e and the device-resident value disagree."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ws, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
device = torch.device("cuda", lr)
if ws > 1:
    dist.init_process_group("nccl", device_id=device)
import netket_b200 as nk  # noqa: E402

rs = np.random.default_rng(bench.WEIGHT_SEED)
W = (rs.normal(size=(bench.N_SITES, bench.N_HIDDEN)) * 0.01).astype(np.float32)
b = (rs.normal(size=bench.N_HIDDEN) * 0.01).astype(np.float32)
a = (rs.normal(size=bench.N_SITES) * 0.01).astype(np.float32)
for sampler_on in (False, True, False, True):
    clk = bench.ClockSampler(lr)
    if sampler_on and rank == 0:
        clk.start()
    vals = []
    for rep in range(4):
        r = bench.measure_e2e(nk, torch, dist, rank, ws, device, "float32", 20, 3, (W, b, a))
        vals.append(r["value"])
    if sampler_on and rank == 0:
        clk.stop()
    if rank == 0:
        print(f"clock sampler {'on ' if sampler_on else 'off'}: e2e samples/s over 4 x 20 steps: " + " ".join(f"{v:.4g}" for v in vals), flush=True)
