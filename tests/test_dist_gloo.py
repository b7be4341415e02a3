"""The N>1 host logic on CPU: two gloo processes.  Chains are sharded with no data-path collective; only the
statistics scalars are all-reduced.  The per-rank partial sums that the GPU kernel would
produce are computed with NumPy here, so the test exercises exactly the combine path of netket_b200.stats."""

import os
import socket
import sys
import warnings

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partials(x, mu):
    n_chains, L = x.shape
    d = x - mu
    lb = max(1, L // 32)
    nb = L // lb
    blocks = d[:, : nb * lb].reshape(n_chains, nb, lb).mean(axis=2)
    half = L // 2
    halves = d[:, : 2 * half].reshape(n_chains, 2, half).mean(axis=2)
    m = d.mean(axis=1)
    return np.array([np.sum(d * d), m.sum(), (m * m).sum(), blocks.sum(), (blocks ** 2).sum(), halves.sum(), (halves ** 2).sum(),
                     d.sum()])


def _worker(rank, ws, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    import netket_b200 as nk
    from netket_b200 import stats as nkstats
    from netket_b200.utils import split_seed, world
    import oracle

    res = {}
    assert world() == (rank, ws)
    # --- sampler sharding arithmetic (metropolis.py:179-203,296-302)
    hi = nk.hilbert.Spin(0.5, 8)
    sa = nk.sampler.MetropolisLocal(hi, n_chains=32)
    res["chains"] = (sa.n_chains, sa.n_chains_per_rank)
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        sb = nk.sampler.MetropolisLocal(hi, n_chains=33)
    res["rounded"] = (sb.n_chains, sb.n_chains_per_rank, any("chains per rank" in str(w.message) for w in wlist))
    res["default"] = nk.sampler.MetropolisLocal(hi).n_chains
    # --- seeds: None draws on rank 0 and broadcasts
    res["seed"] = split_seed(None)
    # --- statistics: every rank holds its shard of chains; combine == statistics of the full array
    rs = np.random.default_rng(123)
    full = rs.normal(size=(24, 50)).cumsum(axis=1) * 0.3 - 5.0
    shard = full[rank * 12:(rank + 1) * 12]
    head = torch.tensor([shard.sum(), float(shard.shape[0])], dtype=torch.float64)
    nkstats._allreduce(head)
    total, n_chains_total = head.tolist()
    mean = total / (n_chains_total * shard.shape[1])
    part = torch.from_numpy(_partials(shard, mean))
    nkstats._allreduce(part)
    st = nkstats.finalize(part.tolist(), mean, int(n_chains_total), shard.shape[1])
    ref = oracle.stats.statistics(full)
    res["stats_ok"] = all(np.isclose(getattr(st, k), ref[k], rtol=1e-10, equal_nan=True)
                          for k in ("mean", "variance", "error_of_mean", "tau_corr", "R_hat"))
    # --- chain sharding of the proposal stream: rank r's chains are chains [12 r, 12 r + 12) of the global run
    from oracle import hilbert as ohilbert, rbm as orbm, sampler as osampler

    W, b, a = orbm.init_params(8, 2, std=0.3)
    sig_full = ohilbert.random_state(9, 24, 8)
    sig_mine = ohilbert.random_state(9, 12, 8, chain_offset=12 * rank)
    res["init_ok"] = bool(np.array_equal(sig_mine, sig_full[12 * rank:12 * rank + 12]))
    run_full = osampler.sample_chain("local", sig_full, W, b, a, chain_length=2, seed=5)
    run_mine = osampler.sample_chain("local", sig_mine, W, b, a, chain_length=2, seed=5, chain_offset=12 * rank)
    res["chains_ok"] = bool(np.array_equal(run_mine["samples"], run_full["samples"][12 * rank:12 * rank + 12]))
    # --- forces: every rank contributes the sums over its shard of samples (what nk_forces_rbm produces); the all-reduced
    # sums scaled by 1 / n_samples_total equal the forces of the full batch (expect_forces.py:81-104)
    from oracle import forces as oforces

    samples = run_full["samples"]
    eloc = np.random.default_rng(4).normal(size=samples.shape[:2])
    mine = slice(12 * rank, 12 * rank + 12)
    f_mine = oforces.forces(samples[mine], eloc[mine], W, b, a, mean=eloc.mean(), n_total=1)  # raw sums of this shard
    sums = torch.from_numpy(np.concatenate([f_mine["W"].ravel(), f_mine["b"], f_mine["a"]]))
    nkstats._allreduce(sums)
    f_full = oforces.forces(samples, eloc, W, b, a)
    ref_vec = np.concatenate([f_full["W"].ravel(), f_full["b"], f_full["a"]])
    res["forces_ok"] = bool(np.allclose(sums.numpy() / eloc.size, ref_vec, rtol=1e-12, atol=1e-14))
    # --- streaming statistics: every rank accumulates its own chains; the derived quantities come from all-reduced sums
    # (3 doubles, then 4 + max_lag + 1), which equal those of one accumulator over all chains
    from oracle import online_stats as oos
    from test_online_stats import numpy_sums

    e_full = e_mine = None
    for lo, hi in [(0, 7), (7, 30), (30, 50)]:
        e_full = oos.online_statistics(full[:, lo:hi], e_full, max_lag=12)
        e_mine = oos.online_statistics(shard[:, lo:hi], e_mine, max_lag=12)
    p0 = torch.from_numpy(np.concatenate([numpy_sums(e_mine), [float(e_mine.n_chains), float(e_mine.n_samples)]]))
    nkstats._allreduce(p0)
    h0 = p0.tolist()
    p1 = torch.from_numpy(numpy_sums(e_mine, h0[1] / h0[0], h0[2] / h0[3]))
    nkstats._allreduce(p1)
    s = nkstats.online_finalize(h0[:3], p1.tolist(), int(h0[3]), int(h0[4]), 12)
    want = [e_full.mean, e_full.error_of_mean, e_full.variance, e_full.tau_corr, e_full.R_hat, e_full.tau_corr_batch, e_full.tau_corr_acf]
    res["online_ok"] = bool(np.allclose(s["out"][:7], want, rtol=1e-10, equal_nan=True) and np.allclose(s["acf"], e_full.acf, rtol=1e-10)
                            and s["n_chains"] == 24 and s["n_samples"] == 24 * 50)
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_process_gloo():
    ws = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(ws, port, out), nprocs=ws, join=True)
    r0, r1 = out[0], out[1]
    for r in (r0, r1):
        assert r["chains"] == (32, 16)
        assert r["rounded"] == (34, 17, True)
        assert r["default"] == 32  # 16 chains per rank
        assert r["stats_ok"] and r["init_ok"] and r["chains_ok"] and r["forces_ok"] and r["online_ok"]
    assert r0["seed"] == r1["seed"]
