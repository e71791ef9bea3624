#!/usr/bin/env python
"""Multi-GPU check of the in-library collective (run under torchrun, one rank per GPU; also runs on one GPU):

  1. rbv_lnprob_batch_allgather: every rank ends with the single-GPU lnprob, bit for bit
  2. DistributedDeviceSampler (rbv_stretch_run_dist: the step incl. the NCCL all-gather is one CUDA graph):
     chain bit-identical to the single-GPU DeviceEnsembleSampler chain with the same seed, on every rank,
     also with seed=None / a rank-dependent initial state (rank 0's are broadcast)
  3. DistributedDeviceSliceSampler (rbv_slice_run with the rows of every iteration split over the ranks):
     chain bit-identical to the single-GPU DeviceEnsembleSliceSampler chain
  4. MCMC steps/s at C5a scale (stretch) and C2 (slice), JSON line on rank 0

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_dist_sampler.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from rbvfit_b200 import dist as rdist
    from rbvfit_b200.sampler import DeviceEnsembleSampler, DistributedDeviceSampler
    from rbvfit_b200.slice_sampler import DeviceEnsembleSliceSampler, DistributedDeviceSliceSampler
    rank, world, local = rdist.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    torch.cuda.set_device(local)
    part = rdist.WalkerPartition(rank, world)
    out = {"world": world}

    def say(msg):
        print(f"[rank {rank}/{world}] {msg}", flush=True)

    # ---- C2: exactness (78 walkers, odd half sizes exercised with 77)
    w, models, like, thetas, spectra = bench.build_problem("C2", local)
    ref_like = like
    w2, models2, dlike, _t, _s = bench.build_problem("C2", local)       # its own context: gets the communicator
    if world > 1:
        assert dlike.engine.comm_init()
    ok = thetas[np.all((thetas >= w["lb"]) & (thetas <= w["ub"]), axis=1)]
    th = torch.as_tensor(thetas, device=f"cuda:{local}")
    a = ref_like.lnprob_device(th).cpu().numpy()
    b = rdist.DistributedLikelihood(dlike, part).lnprob_device(th).cpu().numpy()
    assert np.array_equal(a, b, equal_nan=True), "all-gathered lnprob differs from the single-GPU launch"
    say(f"C2 lnprob all-gather == single-GPU launch on {len(a)} rows")
    for W in (78, 77):
        p0 = ok[:W]
        ref = DeviceEnsembleSampler(W, like.ndim, ref_like, seed=5)
        ref.run_mcmc(p0, 25, skip_initial_state_check=True)
        dsm = DistributedDeviceSampler(W, like.ndim, dlike, part, seed=5)
        dsm.run_mcmc(p0 + (1e-3 * rank if W == 77 else 0.0), 15, skip_initial_state_check=True)   # rank 0's state wins
        dsm.run_mcmc(None, 10)
        same = np.array_equal(dsm.get_chain(), ref.get_chain()) and np.array_equal(dsm.get_log_prob(), ref.get_log_prob())
        same_acc = np.array_equal(dsm.acceptance_fraction, ref.acceptance_fraction)
        say(f"C2 stretch W={W}: chain bit-identical to the single-GPU chain: {same}, acceptance equal: {same_acc}")
        assert same and same_acc
    # seed=None: rank 0's OS-entropy seed is broadcast -> all ranks hold the same chain
    dsm = DistributedDeviceSampler(78, like.ndim, dlike, part, seed=None)
    dsm.run_mcmc(ok[:78], 8, skip_initial_state_check=True)
    if world > 1:
        c = torch.as_tensor(dsm.get_chain(), device=f"cuda:{local}")
        c0 = c.clone()
        torch.distributed.broadcast(c0, src=0)
        assert torch.equal(c, c0), "ranks diverged with seed=None"
    say("C2 stretch seed=None: chains equal on every rank")
    # slice sampler
    Ws = 2 * like.ndim + 8
    rng = np.random.default_rng(5)
    p0 = np.clip(w["theta_true"] + 1e-3 * rng.standard_normal((Ws, like.ndim)), w["lb"] + 1e-10, w["ub"] - 1e-10)
    ref = DeviceEnsembleSliceSampler(Ws, like.ndim, ref_like, seed=3)
    ref.run_mcmc(p0, 30)
    dss = DistributedDeviceSliceSampler(Ws, like.ndim, dlike, part, seed=3)
    dss.run_mcmc(p0, 30)
    same = np.array_equal(dss.get_chain(), ref.get_chain()) and dss.mu == ref.mu
    say(f"C2 slice W={Ws}: chain / mu / call count identical to the single-GPU run: {same}")
    assert same
    if world > 1:
        torch.distributed.barrier()
    rates = []
    for _ in range(3):
        t0 = time.perf_counter()
        dss.run_mcmc(None, 100)
        rates.append(100 / (time.perf_counter() - t0))
    out["c2_slice_steps_per_sec"] = sorted(rates)[1]
    ref_like.close()
    dlike.close()

    # ---- C5a scale
    w, models, like, thetas, spectra = bench.build_problem("C5a", local)
    if world > 1:
        assert like.engine.comm_init()
    ok = thetas[np.all((thetas >= w["lb"]) & (thetas <= w["ub"]), axis=1)]
    W = len(ok) - (len(ok) % 2)
    dsm = DistributedDeviceSampler(W, like.ndim, like, part, seed=7)
    dsm.run_mcmc(ok[:W], 4, skip_initial_state_check=True)
    if world > 1:
        torch.distributed.barrier()
    rates, nsteps = [], 24
    for _ in range(3):
        t0 = time.perf_counter()
        dsm.run_mcmc(None, nsteps)
        rates.append(nsteps / (time.perf_counter() - t0))
    sps = sorted(rates)[1]
    # digest of the chain: identical for every number of ranks (compare the lines of the 1/2/4/8-GPU runs)
    digest = float(np.sum(dsm.get_chain()[:4].astype(np.float64) * np.arange(1, 5)[:, None, None]))
    out.update(c5a_walkers=int(W), c5a_pixels=int(like.total_pixels), c5a_stretch_steps_per_sec=sps,
               c5a_stretch_runs=rates, c5a_walker_pixel_per_sec=sps * W * like.total_pixels,
               c5a_acceptance=float(dsm.acceptance_fraction.mean()), c5a_chain_digest=repr(digest),
               kernel=like.engine.last_kernel, peer_memory_allgather=bool(like.engine.peer_attached),
               peer_error=like.engine.peer_error() if world > 1 else 0)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
