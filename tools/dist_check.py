"""Developer check on N GPUs (torchrun): sharded sampling reproduces the single-device chains, statistics and forces
all-reduce over NCCL agree with a single-device run over the same global chains."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import netket_b200 as nk
g = nk.graph.Hypercube(6, 2); hi = nk.hilbert.Spin(0.5, 36); op = nk.operator.Ising(hi, g, h=2.0)
for dtype in (np.float64, np.float32):
    model = nk.models.RBM(alpha=2, param_dtype=dtype)
    var = model.init(1234, 36, device=torch.device("cuda", lr))
    B = 64
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains_per_rank=B), model, variables=var, n_samples_per_rank=B * 8,
                        n_discard_per_chain=2, sampler_seed=77)
    st, G = vs.expect_and_grad(op)
    if rank == 0:
        # single-process reference over all ws*B chains (torch.distributed makes world() = ws, so emulate by slices)
        pass
    # gather samples / eloc to rank 0 and recompute there with the oracle
    samples = [torch.empty_like(vs.samples) for _ in range(ws)]; dist.all_gather(samples, vs.samples)
    eloc = [torch.empty_like(vs.local_estimators(op).data) for _ in range(ws)]; dist.all_gather(eloc, vs.local_estimators(op).data)
    if rank == 0:
        from oracle import forces as oforces, stats as ostats
        S = torch.cat(samples).cpu().numpy(); E = torch.cat(eloc).cpu().numpy()
        W = var["params"]["Dense"]["kernel"].cpu().numpy().astype(np.float64); b = var["params"]["Dense"]["bias"].cpu().numpy().astype(np.float64)
        a = var["params"]["visible_bias"].cpu().numpy().astype(np.float64)
        ref = oforces.grad(S, E, W, b, a); rst = ostats.statistics(E)
        tol = 1e-10 if dtype == np.float64 else 2e-5
        ok = abs(st.mean - rst["mean"]) < 1e-9 * abs(rst["mean"]) and abs(st.error_of_mean - rst["error_of_mean"]) < 1e-6 * rst["error_of_mean"]
        ok = ok and np.allclose(G["Dense"]["kernel"].cpu().numpy(), ref["W"], rtol=0, atol=tol * np.abs(ref["W"]).max())
        ok = ok and np.allclose(G["visible_bias"].cpu().numpy(), ref["a"], rtol=0, atol=tol * np.abs(ref["a"]).max())
        # chains are distinct across ranks (global chain index keys the stream)
        distinct = len({S[i * B].tobytes() for i in range(ws)}) == ws
        print(f"dist_check {np.dtype(dtype).name} ws={ws}: stats+grad {'OK' if ok else 'MISMATCH'}, distinct shards {distinct}, E = {st}")
    # matrix-free QGT: every rank contracts its own samples, 1 + n_parameters doubles are all-reduced per product
    from netket_b200.optimizer import QGTOnTheFly, SR, tree_to_flat
    vs.reset(); st, G = vs.expect_and_grad(op)
    Sq = QGTOnTheFly(vs, diag_shift=0.01)
    vvec = torch.from_numpy(np.random.default_rng(11).normal(size=Sq.shape[0])).cuda()
    got_sv = (Sq @ vvec).cpu().numpy()
    dp = tree_to_flat(SR(diag_shift=0.01)(vs, G)).cpu().numpy()
    samples = [torch.empty_like(vs.samples) for _ in range(ws)]; dist.all_gather(samples, vs.samples)
    if rank == 0:
        from oracle import qgt as oqgt
        S_all = torch.cat(samples).cpu().numpy()
        W = var["params"]["Dense"]["kernel"].cpu().numpy().astype(np.float64); b = var["params"]["Dense"]["bias"].cpu().numpy().astype(np.float64)
        a = var["params"]["visible_bias"].cpu().numpy().astype(np.float64)
        want_sv = oqgt.mat_vec(S_all, W, b, a, vvec.cpu().numpy(), 0.01)
        want_dp = oqgt.sr_solve(S_all, W, b, a, tree_to_flat(G).cpu().numpy(), 0.01)
        tol = 1e-10 if dtype == np.float64 else 3e-5
        ok = np.abs(got_sv - want_sv).max() <= tol * np.abs(want_sv).max() and np.abs(dp - want_dp).max() <= 2e-3 * np.abs(want_dp).max()
        print(f"dist_check {np.dtype(dtype).name} ws={ws}: QGT product + SR solve over {S_all.shape[0] * S_all.shape[1]} samples {'OK' if ok else 'MISMATCH'} "
              f"(product err {np.abs(got_sv - want_sv).max() / np.abs(want_sv).max():.1e}, solve err {np.abs(dp - want_dp).max() / np.abs(want_dp).max():.1e})")
    # streaming statistics: every rank accumulates its own chains, the derived quantities all-reduce the summary sums
    from netket_b200.stats import online_statistics
    acc = None
    batches = []
    for it in range(3):
        e = vs._sample_and_estimate(op, None if it == 0 else 0)
        acc = online_statistics(e, acc, max_lag=16, inplace=True)
        allb = [torch.empty_like(e) for _ in range(ws)]; dist.all_gather(allb, e); batches.append(torch.cat(allb).cpu().numpy())
    got = [acc.mean, acc.error_of_mean, acc.variance, acc.tau_corr, acc.R_hat, acc.tau_corr_batch, acc.tau_corr_acf, acc.n_samples]
    if rank == 0:
        from oracle import online_stats as oos
        ref = None
        for bt in batches:
            ref = oos.online_statistics(bt.astype(np.float64), ref, max_lag=16)
        want = [ref.mean, ref.error_of_mean, ref.variance, ref.tau_corr, ref.R_hat, ref.tau_corr_batch, ref.tau_corr_acf, ref.n_samples]
        ok = np.allclose(got, want, rtol=1e-9 if dtype == np.float64 else 1e-5) and np.allclose(acc.acf, ref.acf, rtol=1e-5, atol=1e-6)
        print(f"dist_check {np.dtype(dtype).name} ws={ws}: online statistics over {ref.n_chains} chains {'OK' if ok else 'MISMATCH'}: {acc}")
    st2 = vs.expect_to_precision(op, rtol=2e-3, max_iter=50, verbose=False)
    if rank == 0:
        print(f"dist_check expect_to_precision ws={ws}: {st2}, n_samples={st2.n_samples}")
dist.barrier(); dist.destroy_process_group()
