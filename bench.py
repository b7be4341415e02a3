#!/usr/bin/env python
"""bench.py — E_loc samples/sec (1 sweep/sample), RBM alpha=4, 10x10 TFIM (BASELINE.json metric), on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype float32|float64]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one batch: `vs.reset(); vs.expect(H)`, i.e. for every chain
`chain_length` x (one sweep of N Metropolis proposals + the local energy of the resulting sample) + the MC
statistics, reduced inside the sweep kernel (only 10 scalars cross GPUs, one NCCL all-reduce, one host read).
Workload: N=100, M=400, h=3, J=1, W,b,a ~ N(0, 0.01^2) from default_rng(1234), 2^16 chains per GPU
(weak scaling: chains are independent, no data-path collective), chain_length 16, 5 untimed burn-in sweeps.

Printed JSON (one line, rank 0):
  value      whole-job samples/s with inputs resident in HBM, CUDA events around the K steps, max over ranks
  e2e        the same metric through the host-buffer C ABI (nk_ctx_step_begin / all-reduce / nk_ctx_step_end): parameters
             H2D from pinned memory, E_loc + statistics D2H, the NCCL all-reduce of the partial sums, every step
  roofline   dominant kernel (fused sweep + E_loc + statistics): algorithmic shared-memory operand bytes / kernel time vs
             the shared-memory read bandwidth measured in this run (nk_microbench)
  strong     the split BASELINE.json names: cfg-3 with 2^16 chains IN TOTAL and cfg-5 with 2^20 chains IN TOTAL over the N GPUs
  other_configs  cfg-4 (J1-J2, fp64, MetropolisExchange) and cfg-5 (TFIM 20x20, alpha=8, fp32) with their L2 rooflines
  cpu_baseline  the reference *algorithm* (oracle/reference_algorithm.py, kind "port") on this box's host cores

`--impl reference` prints the same metric / config for that CPU port, K timed steps after W warm-up steps as asked, every
step a bounded sample (1024 chains x 1 sweep + E_loc) of the workload.
"""

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L_SIDE, ALPHA, H_FIELD, J_COUP = 10, 4, 3.0, 1.0
N_SITES, N_HIDDEN = L_SIDE * L_SIDE, ALPHA * L_SIDE * L_SIDE
CHAINS_PER_GPU = 1 << 16
CHAIN_LENGTH = 16
WEIGHT_SEED, SAMPLER_SEED = 1234, 15324
METRIC = "E_loc samples/sec (1 sweep/sample), RBM a=4 10x10 TFIM"
CPU_SAMPLE_CHAINS = 1024


def workload_name(dtype):
    return (f"TFIM 10x10 pbc h=3 J=1, RBM alpha=4 (N=100, M=400) {dtype}, MetropolisLocal, {CHAINS_PER_GPU} chains/GPU x "
            f"chain_length {CHAIN_LENGTH} (sweep_size=N), fused E_loc + statistics per step")


def config_dict(dtype, ws):
    """The `config` of the JSON line: the same for both arms (the reference arm times a bounded sample of this workload)."""
    return {"workload": workload_name(dtype),
            "parallelism": f"chains sharded over {ws} GPU(s), one process per GPU; only the statistics scalars are all-reduced (NCCL)",
            "l2": "no explicit L2 flush: the path is bound by shared-memory operand reads (W is staged once per CTA); per step it "
                  "streams 105 MB of theta scratch and 105 MB of samples through HBM, more than the 126 MB L2",
            "weights": "W,b,a ~ N(0, 0.01^2), numpy default_rng(1234); sampler seed 15324; 5 burn-in sweeps untimed"}


def bytes_per_sample(esz):
    """Algorithmic on-chip operand bytes per (sweep + E_loc) sample: N proposals + N flips, one W row of M elements each
    (SURVEY.md §8d): 2 * N * M * sizeof(T)."""
    return 2 * N_SITES * N_HIDDEN * esz


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i",
                                          str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------- reference arm
def edges_np():
    from oracle import graph as ograph

    return ograph.hypercube_edges(L_SIDE, 2)[0]


def run_netket_cpu(dtype, steps, warmup):
    """NetKet itself on jax[cpu] (tried first).  Raises ImportError where jax / netket cannot be imported - which is
    every box of this image: jax is not installed and `pip install --no-index` of the reference fails on it - and run_cpu then
    times the port of the reference algorithm instead."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    os.environ.setdefault("JAX_PLATFORMS", "cpu")
    import jax  # noqa: F401  (ImportError here is the expected outcome)
    import netket as rnk

    g = rnk.graph.Hypercube(length=L_SIDE, n_dim=2, pbc=True)
    hi = rnk.hilbert.Spin(s=0.5, N=g.n_nodes)
    ha = rnk.operator.Ising(hi, g, h=H_FIELD, J=J_COUP)
    vs = rnk.vqs.MCState(rnk.sampler.MetropolisLocal(hi, n_chains=CPU_SAMPLE_CHAINS), rnk.models.RBM(alpha=ALPHA, param_dtype=np.dtype(dtype).type),
                         n_samples=CPU_SAMPLE_CHAINS, n_discard_per_chain=0, seed=WEIGHT_SEED, sampler_seed=SAMPLER_SEED)

    def step():
        vs.reset()
        le = vs.local_estimators(ha)
        jax.block_until_ready(getattr(le, "data", le))

    for _ in range(warmup + 1):  # + 1: compilation
        step()
    tic = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - tic
    val = CPU_SAMPLE_CHAINS * steps / dt
    return val, dt / steps * 1e3, {"value": val, "unit": "samples/s", "cores": os.cpu_count(), "kind": "reference",
                                   "sample": f"{CPU_SAMPLE_CHAINS} chains x 1 (sweep + E_loc) per step, {steps} timed steps",
                                   "what": "NetKet on jax[cpu]: vs.reset(); vs.local_estimators(H)"}


def run_cpu(dtype, steps, warmup):
    """The reference on the host cores, on a bounded sample of the workload (see cpu_baseline.sample): NetKet itself where it can be
    imported, else the port of its algorithm."""
    why = ""
    try:
        return run_netket_cpu(dtype, steps, warmup)
    except Exception as exc:  # ImportError (no jax) in this image; anything else: the port is always there
        why = f"{type(exc).__name__}: {str(exc).splitlines()[0] if str(exc) else ''}"[:160]
    from oracle import reference_algorithm as ra

    # all the host threads the box has (torchrun exports OMP_NUM_THREADS=1, which would otherwise cap the baseline at one core)
    val, ms, threads = ra.time_vmc_steps(N_SITES, N_HIDDEN, CPU_SAMPLE_CHAINS, 1, edges_np(), H_FIELD, J_COUP, np.dtype(dtype).type,
                                         steps=steps, warmup=warmup, seed=SAMPLER_SEED, threads=os.cpu_count())
    sample = (f"{CPU_SAMPLE_CHAINS} chains x 1 (sweep + E_loc) per step, {steps} timed steps after {warmup} warm-up; same "
              "N, M, h, J, parameters and algorithmic structure as the GPU workload; throughput is linear in the number "
              "of chains once the GEMMs saturate the cores")
    return val, ms, {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample,
                     "what": "reference algorithm (full forward pass per proposal, materialised connected states), torch CPU "
                             "kernels on all host threads; NetKet's jax[cpu] path itself was tried first and cannot run here ("
                             + why + ")"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)  # exactly what was asked for
    val, ms, cpu = run_cpu(args.dtype, steps, warmup)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps,
           "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic", "config": config_dict(args.dtype, ws),
           "cpu_baseline": cpu, "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------- our arm
def build_state(nk, torch, dtype, device, chains, chain_length=CHAIN_LENGTH):
    g = nk.graph.Hypercube(L_SIDE, 2, pbc=True)
    hi = nk.hilbert.Spin(0.5, g.n_nodes)
    ha = nk.operator.Ising(hi, g, h=H_FIELD, J=J_COUP)
    model = nk.models.RBM(alpha=ALPHA, param_dtype=dtype)
    var = model.init(WEIGHT_SEED, N_SITES, device=device)  # W, b, a ~ N(0, 0.01^2) from numpy default_rng(1234)
    Wt, bt, at = nk.models.RBM.unpack(var)
    sa = nk.sampler.MetropolisLocal(hi, n_chains_per_rank=chains)
    vs = nk.vqs.MCState(sa, model, variables=var, n_samples_per_rank=chains * chain_length, n_discard_per_chain=0,
                        sampler_seed=SAMPLER_SEED)
    return g, hi, ha, vs, (Wt.cpu().numpy(), bt.cpu().numpy(), at.cpu().numpy())


def timed_steps(torch, dist, ws, fn, steps, warmup):
    """W warm-up + K timed steps, CUDA events on the current stream, barrier + synchronize on both sides, max over ranks."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if ws > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if ws > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if ws > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def microbench(which):
    from netket_b200 import _lib

    res = C.c_double()
    _lib.check(_lib.lib().nk_microbench(which, C.byref(res)))
    return float(res.value)


def measure(nk, torch, dist, rank, ws, device, dtype, steps, warmup):
    from netket_b200 import _lib

    L = _lib.lib()
    esz = 4 if dtype == "float32" else 8
    g, hi, ha, vs, (W, b, a) = build_state(nk, torch, dtype, device, CHAINS_PER_GPU)
    # burn-in: 5 sweeps, untimed (netket/vqs/mc/mc_state/state.py:466-468)
    vs.sample(chain_length=CHAIN_LENGTH, n_discard_per_chain=5)
    result = {}

    def step():
        vs.reset()
        result["stats"] = vs.expect(ha)

    step()  # centres the shift of the in-kernel statistics on the energy (the first call may take the two-pass route)
    torch.cuda.synchronize()
    launches0 = L.nk_launch_count()
    ms = timed_steps(torch, dist, ws, step, steps, warmup)
    launches = (L.nk_launch_count() - launches0) // (steps + warmup)  # our kernels per step
    samples_per_step = CHAINS_PER_GPU * CHAIN_LENGTH * ws
    value = samples_per_step * steps / (ms * 1e-3)

    # ---- dominant kernel alone: the fused sweep + E_loc (+ statistics) launch = nk_sweep minus the theta kernel before it
    sa = vs.sampler
    st0 = vs.sampler_state
    shift = float(result["stats"].mean)

    def sweep_only():
        sa._launch(vs.model, vs.variables, st0, CHAIN_LENGTH, operator=ha, want_samples=True, stats_shift=shift)

    ms_sweep_call = timed_steps(torch, dist, 1, sweep_only, max(3, steps // 2), 2) / max(3, steps // 2)
    rbm = nk.models.RBM.c_struct(vs.variables)
    theta = torch.empty((CHAINS_PER_GPU, N_HIDDEN), dtype=vs.variables["params"]["Dense"]["kernel"].dtype, device=device)
    scratch = torch.empty(int(L.nk_theta_gemm_workspace_bytes(C.byref(rbm), CHAINS_PER_GPU)), dtype=torch.uint8, device=device)

    def theta_only():
        _lib.check(L.nk_theta_gemm(_lib.stream_ptr(device), C.byref(rbm), _lib.ptr(st0.σ), CHAINS_PER_GPU, _lib.ptr(theta),
                                   _lib.ptr(scratch)))

    ms_theta = timed_steps(torch, dist, 1, theta_only, 5, 2) / 5
    ms_kernel = ms_sweep_call - ms_theta
    alg_bytes = CHAINS_PER_GPU * CHAIN_LENGTH * bytes_per_sample(esz)
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    peak = microbench(0)
    kernel = ("sweep_fast_kernel<3,1> (fused sweep + E_loc + statistics, fp32 LocalRule specialisation)" if dtype == "float32"
              else "sweep_shadow_kernel<3,1> (fused sweep + E_loc + statistics: accept / reject on an fp32 shadow of the state against "
                   "a resident fp32 table, re-decided in fp64 inside the shadow's error band; the fp64 state follows the net flips of "
                   "each sweep and the fp64 local energy is formed from double rows streamed through a TMA ring)")
    roofline = {"bound": "smem", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None,
                "traffic_note": "DRAM bytes are not measurable inside bench.py (no profiler in the timed run); the ncu captures "
                                "under profiles/ give dram__bytes_read + write per launch (theta in, samples + E_loc out: "
                                "~0.06 % of the on-chip operand bytes this roofline counts)",
                "kernel": kernel, "kernel_ms": ms_kernel, "theta_kernel_ms": ms_theta,
                "algorithmic_bytes_per_sample": bytes_per_sample(esz),
                "peak_source": "measured in this run by nk_microbench(0) (LDS.128 shared-memory read bandwidth, all SMs; method "
                               "and a tracked copy in profiles/r02_onchip_peaks.json); MEASURED_PEAKS.json holds no on-chip figure "
                               "(its HBM copy number does not bound this path: HBM traffic is ~120 B/sample)",
                "method": "CUDA events on the launching stream around nk_sweep, minus the separately timed theta kernel"}
    if dtype == "float64":
        # competing bounds of the fp64 kernel, measured in the same run: the FP64 pipe (E: one DFMA + one DMUL per table element
        # of every site; U: one DMUL per element of every net-flipped site, a fraction (1 - exp(-2 acc)) / 2 of the sites) and the
        # L2 -> shared stream of the double table (two passes per sample and CTA, shared by the 11 chains a CTA sweeps at a time)
        import math
        dp_peak_inst = microbench(4) / 2.0  # G lane-instructions/s (the microbenchmark counts 2 flop per DFMA)
        acc = float(vs.sampler_state.acceptance)
        net_flip = 0.5 * (1.0 - math.exp(-2.0 * acc))
        dp_per_sample = N_SITES * N_HIDDEN * (2.0 + net_flip)
        dp_ach = CHAINS_PER_GPU * CHAIN_LENGTH * dp_per_sample / (ms_kernel * 1e-3) / 1e9
        l2_peak = microbench(1)
        l2_bytes = 2.0 * N_SITES * N_HIDDEN * 8.0 / 11.0
        roofline["competing"] = {
            "fp64_pipe": {"achieved": dp_ach, "peak": dp_peak_inst, "unit": "G lane-instructions/s", "frac": dp_ach / dp_peak_inst},
            "l2_stream": {"achieved": CHAINS_PER_GPU * CHAIN_LENGTH * l2_bytes / (ms_kernel * 1e-3) / 1e9, "peak": l2_peak, "unit": "GB/s",
                          "frac": CHAINS_PER_GPU * CHAIN_LENGTH * l2_bytes / (ms_kernel * 1e-3) / 1e9 / l2_peak},
            "note": "neither bound binds: at 168 registers per thread (26 doubles of state + two rows in flight) the kernel runs 3 warps "
                    "per scheduler and is bound by dependent-issue latency (ncu: issue slots 45 % busy, stall 'wait' 2.2 cycles per "
                    "instruction; profiles/r02_ncu_summary.txt)"}
    return {"value": value, "ms_per_step": ms / steps, "launches": launches, "roofline": roofline, "stats": result["stats"],
            "acceptance": vs.sampler_state.acceptance, "params": (W, b, a)}


class _DevPtr:
    """A raw device pointer as a zero-copy torch tensor (the C ABI's partial sums, all-reduced in place by torch.distributed)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def measure_e2e(nk, torch, dist, rank, ws, device, dtype, steps, warmup, params):
    """Host-buffer C ABI: parameters from pinned host memory every step, E_loc + statistics back to the host; at N > 1 the
    statistics' partial sums are all-reduced over NCCL between nk_ctx_step_begin and nk_ctx_step_end."""
    from netket_b200 import _lib

    L = _lib.lib()
    W, b, a = params
    esz = W.dtype.itemsize
    ctx = C.c_void_p()
    e = np.ascontiguousarray(np.asarray(nk.graph.Hypercube(L_SIDE, 2, pbc=True).edges()), dtype=np.int32)
    op = _lib.nk_ising_t(edges=e.ctypes.data_as(C.c_void_p), n_edges=e.shape[0], reserved=0, h=H_FIELD, J=J_COUP)
    d = _lib.nk_ctx_desc_t()
    d.device, d.N, d.M, d.dtype = device.index, N_SITES, N_HIDDEN, _lib.dtype_code(W.dtype)
    d.n_chains, d.chain_length, d.sweep_size, d.rule, d.machine_pow, d.n_down = CHAINS_PER_GPU, CHAIN_LENGTH, 0, _lib.NK_RULE_LOCAL, 2.0, -1
    d.return_samples = 1  # the samples are written to HBM as in the `value` path (and stay there, as MCState.samples do)
    d.ising_host = C.pointer(op)
    d.seed, d.chain_offset = SAMPLER_SEED, rank * CHAINS_PER_GPU
    d.stream = None  # a stream of the context's own; the all-reduce below is issued on it (nk_ctx_stream)
    d.eloc_in_param_dtype = 1
    _lib.check(L.nk_ctx_create2(C.byref(ctx), C.byref(d)))
    part = torch.as_tensor(_DevPtr(L.nk_ctx_partials_device(ctx), _lib.NK_CTX_NPARTIAL), device=device)
    # the context's stream as a torch stream: torch.distributed orders a collective after the work already queued on the CURRENT
    # stream and makes that stream wait for it - so the all-reduce must be issued with the context's stream current
    ctx_stream = torch.cuda.ExternalStream(int(L.nk_ctx_stream(ctx)), device=device)
    pin = lambda x: torch.from_numpy(x).pin_memory()  # noqa: E731
    Wp, bp, ap = pin(W), pin(b), pin(a)
    eloc = torch.empty((CHAINS_PER_GPU, CHAIN_LENGTH), dtype=Wp.dtype).pin_memory()
    stats = (C.c_double * 6)()

    def step(n_discard=0):
        _lib.check(L.nk_ctx_step_begin(ctx, Wp.data_ptr(), bp.data_ptr(), ap.data_ptr(), n_discard))
        if ws > 1:
            with torch.cuda.stream(ctx_stream):
                dist.all_reduce(part)  # NK_CTX_NPARTIAL doubles: the only cross-device traffic of a step
        rc = L.nk_ctx_step_end(ctx, eloc.data_ptr(), None, stats)
        if rc not in (0, _lib.NK_RESHIFT):
            _lib.check(rc)

    step(5)
    for _ in range(warmup):
        step()
    if ws > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tic = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - tic
    if ws > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    del part
    L.nk_ctx_destroy(ctx)
    return {"value": CHAINS_PER_GPU * CHAIN_LENGTH * ws * steps / dt, "unit": "samples/s",
            "h2d_bytes_per_step": int((W.size + b.size + a.size) * esz),
            "d2h_bytes_per_step": int(CHAINS_PER_GPU * CHAIN_LENGTH * esz + 8 * _lib.NK_CTX_NPARTIAL),
            "api": "nk_ctx_step_begin / nk_ctx_step_end (include/nkb200.h): W,b,a from pinned host memory; sweeps + E_loc + "
                   "statistics sums in one launch, samples written to HBM; the partial sums all-reduced over NCCL (N > 1); "
                   "E_loc[chains, chain_length], the 5 statistics and the acceptance back to host memory; one synchronisation",
            "mean_energy": stats[0], "acceptance": stats[5], "collective": "nccl all_reduce of 10 doubles" if ws > 1 else None}


def measure_config(nk, torch, dist, ws, device, *, kind, chains_per_rank, chain_length, steps, warmup, dtype):
    """Device-resident steps of another BASELINE configuration (cfg-3 at another split, cfg-4, cfg-5)."""
    if kind == "cfg3":
        g = nk.graph.Hypercube(10, 2)
        hi = nk.hilbert.Spin(0.5, 100)
        op = nk.operator.Ising(hi, g, h=3.0)
        sa = nk.sampler.MetropolisLocal(hi, n_chains_per_rank=chains_per_rank)
        alpha = 4
    elif kind == "cfg5":
        g = nk.graph.Hypercube(20, 2)
        hi = nk.hilbert.Spin(0.5, 400)
        op = nk.operator.Ising(hi, g, h=3.0)
        sa = nk.sampler.MetropolisLocal(hi, n_chains_per_rank=chains_per_rank)
        alpha = 8
    else:  # cfg4: J1-J2 10x10, Examples/HeisenbergJ1J2
        g = nk.graph.Hypercube(10, 2, max_neighbor_order=2)
        hi = nk.hilbert.Spin(0.5, 100, total_sz=0)
        op = nk.operator.Heisenberg(hi, g, J=[1.0, 0.5], sign_rule=[False, False])
        sa = nk.sampler.MetropolisExchange(hi, graph=g, d_max=1, n_chains_per_rank=chains_per_rank)
        alpha = 4
    model = nk.models.RBM(alpha=alpha, param_dtype=dtype)
    var = model.init(WEIGHT_SEED, hi.size, device=device)
    vs = nk.vqs.MCState(sa, model, variables=var, n_samples_per_rank=chains_per_rank * chain_length, n_discard_per_chain=0,
                        sampler_seed=SAMPLER_SEED)
    vs.sample(chain_length=1, n_discard_per_chain=5)
    out = {}

    def step():
        vs.reset()
        out["stats"] = vs.expect(op)

    step()
    ms = timed_steps(torch, dist, ws, step, steps, warmup)
    n = chains_per_rank * chain_length * ws
    res = {"value": n * steps / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms / steps, "steps": steps, "chains_total": chains_per_rank * ws,
           "chain_length": chain_length, "dtype": dtype, "mean_energy": out["stats"].mean, "acceptance": vs.sampler_state.acceptance}
    return res, vs, op


def measure_strong_and_others(nk, torch, dist, rank, ws, device, steps):
    """The split BASELINE.json's configs name (strong scaling: the TOTAL number of chains is fixed), and the rooflines of the
    L2-resident configurations.  Bounded: a few steps each."""
    strong, others = {}, {}
    k = max(3, min(steps, 10))
    r, _, _ = measure_config(nk, torch, dist, ws, device, kind="cfg3", chains_per_rank=(1 << 16) // ws, chain_length=CHAIN_LENGTH, steps=k,
                             warmup=3, dtype="float32")
    strong["cfg3_f32_65536_chains_total"] = r
    r, _, _ = measure_config(nk, torch, dist, ws, device, kind="cfg5", chains_per_rank=(1 << 20) // ws, chain_length=1, steps=3, warmup=3,
                             dtype="float32")
    strong["cfg5_f32_1048576_chains_total"] = r
    l2_peak = microbench(1)
    # cfg-5 roofline: every proposal and every flip of E_loc reads one table row of M floats through L2
    b5 = 2 * 400 * 3200 * 4
    r5 = dict(r)
    r5["roofline"] = {"bound": "l2", "achieved": r["value"] / ws * b5 / 1e9, "peak": l2_peak, "unit": "GB/s",
                      "frac": r["value"] / ws * b5 / 1e9 / l2_peak, "algorithmic_bytes_per_sample": b5,
                      "kernel": "sweep_prod_kernel<float,2,2,LocalRule,MULTI> (10 warps per chain, rows through L2)",
                      "note": "step time (sweep kernel + theta GEMM + statistics), not the kernel alone"}
    others["cfg5_tfim20x20_alpha8_f32"] = r5
    # cfg-4: 2^18 samples over the N GPUs = 2^14 / N chains per GPU x 16
    r4, vs4, _ = measure_config(nk, torch, dist, ws, device, kind="cfg4", chains_per_rank=max(1, (1 << 14) // ws), chain_length=16, steps=k,
                                warmup=3, dtype="float64")
    s = vs4.samples
    g4 = nk.graph.Hypercube(10, 2, max_neighbor_order=2)
    e4 = torch.as_tensor(np.asarray(g4.edges()), device=device).long()
    anti = float((s[..., e4[:, 0]] != s[..., e4[:, 1]]).float().mean().item())  # fraction of bonds with an exchange term
    b4 = (2 * 100 * 400 + 2 * anti * 400 * 400) * 8  # sweep: 2 rows per proposal; E_loc: 2 rows per antiparallel bond
    r4["roofline"] = {"bound": "smem+l2", "achieved": r4["value"] / ws * b4 / 1e9, "peak": microbench(0), "unit": "GB/s",
                      "frac": r4["value"] / ws * b4 / 1e9 / microbench(0), "algorithmic_bytes_per_sample": b4,
                      "antiparallel_bond_fraction": anti, "l2_peak": l2_peak,
                      "kernel": "sweep_prod_kernel<double,6,1,ExchangeRule> (66 of 100 rows resident, the others through L2)",
                      "note": "step time, against the shared-memory read bandwidth (the resident rows' bound)"}
    others["cfg4_j1j2_f64_exchange"] = r4
    strong["cfg4_f64_262144_samples_total"] = {k2: r4[k2] for k2 in ("value", "unit", "ms_per_step", "steps", "chains_total", "chain_length")}
    # K4: Ising get_conn_padded on the cfg-3 batch - the path's HBM-write-bound kernel - and `statistics` on
    # given data, against the HBM copy bandwidth the driver measured (MEASURED_PEAKS.json) or the profiling guide's fallback
    try:
        hbm_peak, hbm_src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        hbm_peak, hbm_src = 7700.0, "fallback: B200_PROFILING.md nominal 7.7 TB/s (MEASURED_PEAKS.json absent)"
    g3 = nk.graph.Hypercube(L_SIDE, 2, pbc=True)
    hi3 = nk.hilbert.Spin(0.5, N_SITES)
    op3 = nk.operator.Ising(hi3, g3, h=H_FIELD)
    x3 = hi3.random_state(1, CHAINS_PER_GPU, device=device)
    ms_conn = timed_steps(torch, dist, 1, lambda: op3.get_conn_padded(x3), 5, 3) / 5
    by = CHAINS_PER_GPU * (N_SITES + 1) * N_SITES + CHAINS_PER_GPU * (N_SITES + 1) * 8 + CHAINS_PER_GPU * N_SITES
    others["k4_ising_get_conn_padded_cfg3"] = {
        "ms": ms_conn, "batch": CHAINS_PER_GPU, "n_conn": N_SITES + 1,
        "roofline": {"bound": "hbm", "achieved": by / (ms_conn * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": by / (ms_conn * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": by, "peak_source": hbm_src,
                     "kernel": "ising_conn_kernel (sigma' [B, N+1, N] int8 + mels [B, N+1] fp64 written, sigma read)"}}
    e3 = torch.randn(CHAINS_PER_GPU, CHAIN_LENGTH, dtype=torch.float64, device=device)
    ms_stat = timed_steps(torch, dist, 1, lambda: nk.stats.statistics(e3), 5, 3) / 5
    others["k6_statistics_on_given_data"] = {"ms": ms_stat, "shape": [CHAINS_PER_GPU, CHAIN_LENGTH],
                                            "note": "two-pass stats_partial_kernel + all-reduce + one host read; inside a step the sums "
                                                    "come from the sweep kernel's epilogue instead"}
    return strong, others


def main_ours(args):
    import torch
    import torch.distributed as dist

    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: netket_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if ws > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    import __graft_entry__

    if rank == 0:
        __graft_entry__.build()
    if ws > 1:
        dist.barrier()
    import netket_b200 as nk

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    main = measure(nk, torch, dist, rank, ws, device, args.dtype, args.steps, args.warmup)
    e2e = measure_e2e(nk, torch, dist, rank, ws, device, args.dtype, args.steps, args.warmup, main["params"])
    clk = clocks.stop() if rank == 0 else None
    other = "float64" if args.dtype == "float32" else "float32"
    extra = None
    if not args.no_second_dtype:
        k2 = max(3, args.steps // 2) if other == "float64" else args.steps
        m2 = measure(nk, torch, dist, rank, ws, device, other, k2, 3)
        extra = {"dtype": other, "value": m2["value"], "unit": "samples/s", "steps": k2, "ms_per_step": m2["ms_per_step"],
                 "roofline": m2["roofline"], "mean_energy": m2["stats"].mean}
    strong = others = None
    if not args.no_extra:
        strong, others = measure_strong_and_others(nk, torch, dist, rank, ws, device, args.steps)
    cpu = None
    if rank == 0 and ws == 1 and not args.no_cpu:
        _, _, cpu = run_cpu(args.dtype, 12, 1)
    if rank == 0:
        st = main["stats"]
        out = {"metric": METRIC, "value": main["value"], "unit": "samples/s", "n_gpus": ws, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
               "config": config_dict(args.dtype, ws),
               "e2e": e2e, "gpu_launches": int(main["launches"] * args.steps), "gpu_launches_per_step": int(main["launches"]),
               "roofline": main["roofline"], "clocks": clk,
               "result": {"energy_mean": st.mean, "energy_sigma": st.error_of_mean, "variance": st.variance, "R_hat": st.R_hat,
                          "tau_corr": st.tau_corr, "acceptance": main["acceptance"]}}
        if extra is not None:
            out["second_dtype"] = extra
        if strong is not None:
            out["strong"] = strong
            out["other_configs"] = others
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-second-dtype", action="store_true", help="skip the shorter run in the other precision")
    ap.add_argument("--no-extra", action="store_true", help="skip the strong-scaling split and the cfg-4 / cfg-5 legs")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "ours":
        a.warmup = 3
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
