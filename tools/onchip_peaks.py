"""Writes the on-chip roofline denominators of the current GPU, with the method of each measurement, as JSON
(committed as profiles/r02_onchip_peaks.json; bench.py re-measures the same numbers in every run through nk_microbench).

    python tools/onchip_peaks.py > gpurun_out/onchip_peaks.json
"""
import ctypes as C
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from netket_b200 import _lib  # noqa: E402

METHODS = {
    0: ("smem_read_GBps", "mb_smem_kernel (netket_b200/csrc/microbench.cu): one CTA of 1024 threads per SM, 128 KB of shared memory, "
        "every thread issues conflict-free LDS.128 (8 per iteration, 4000 iterations, addresses advanced by 37 * 32 words), accumulating "
        "into registers; bytes = SMs * 1024 threads * iterations * 8 * 16; time = best of 3 CUDA-event timings after one warm-up"),
    1: ("l2_read_GBps", "mb_l2_kernel: 4 CTAs of 512 threads per SM stream a 32 MB buffer (L2-resident: 126 MB L2) 20 times with 128-bit "
        "ld.global.cg loads (L1 bypassed); bytes = 32 MB * 20"),
    2: ("fp32_fma_GFLOPs", "mb_fma_kernel<float>: 2 CTAs of 1024 threads per SM, 8 independent FFMA chains per thread, 4000 x 8 iterations; 2 flop per FMA"),
    3: ("mufu_Gops", "mb_mufu_kernel: 4 independent lg2.approx chains per thread, 2 CTAs of 1024 threads per SM"),
    4: ("fp64_fma_GFLOPs", "mb_fma_kernel<double>: as fp32, DFMA; 2 flop per FMA"),
}

torch.cuda.init()
L = _lib.lib()
out = {"gpu": torch.cuda.get_device_name(0), "sms": torch.cuda.get_device_properties(0).multi_processor_count, "entries": {}}
try:
    q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
    out["clocks_sm_mhz_idle_query"] = q.stdout.strip()
except OSError:
    pass
for which, (name, how) in METHODS.items():
    vals = []
    for _ in range(3):
        r = C.c_double()
        _lib.check(L.nk_microbench(which, C.byref(r)))
        vals.append(float(r.value))
    out["entries"][name] = {"value": max(vals), "runs": vals, "method": how}
sm = out["sms"]
out["theory"] = {"smem_read_GBps_at_1965MHz": 128 * sm * 1.965, "note": "128 B/clk/SM of shared-memory bandwidth (B300_MICROARCH.md:235-240) x SMs x 1.965 GHz"}
print(json.dumps(out, indent=1))
