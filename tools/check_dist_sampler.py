#!/usr/bin/env python
"""Multi-GPU check of DistributedDeviceSampler (run under torchrun, one rank per GPU): the chain must equal the
single-GPU DeviceEnsembleSampler chain for the same seed (to 1e-9: lnprob differs in the last bit between tile
geometries), on every rank.  Also times a C5a-scale run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_dist_sampler.py
"""
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
    from rbvfit_b200 import workloads as wl
    from rbvfit_b200.sampler import DeviceEnsembleSampler, DistributedDeviceSampler
    rank, world, local = rdist.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    torch.cuda.set_device(local)
    part = rdist.WalkerPartition(rank, world)
    # 1. exactness on C2 (78 walkers, odd half sizes exercised with 77)
    w, models, like, thetas, spectra = bench.build_problem("C2", local)
    ok = thetas[np.all((thetas >= w["lb"]) & (thetas <= w["ub"]), axis=1)]
    for W in (78, 77):
        p0 = ok[:W]
        ref = DeviceEnsembleSampler(W, like.ndim, like, seed=5)
        ref.run_mcmc(p0, 25, skip_initial_state_check=True)
        dsm = DistributedDeviceSampler(W, like.ndim, like, part, seed=5)
        dsm.run_mcmc(p0, 15, skip_initial_state_check=True)
        dsm.run_mcmc(None, 10)
        d = np.max(np.abs(dsm.get_chain() - ref.get_chain()))
        same_acc = np.array_equal(dsm.acceptance_fraction, ref.acceptance_fraction)
        print(f"[rank {rank}/{world}] C2 W={W}: max |chain - single-GPU chain| = {d:.3e}, acceptance equal: {same_acc}",
              flush=True)
        assert d <= 1e-9 and same_acc
    like.close()
    # 2. C5a scale
    w, models, like, thetas, spectra = bench.build_problem("C5a", local)
    ok = thetas[np.all((thetas >= w["lb"]) & (thetas <= w["ub"]), axis=1)]
    W = len(ok) - (len(ok) % 2)
    ref = DeviceEnsembleSampler(W, like.ndim, like, seed=7)          # fused path (finalize_kernel at this size)
    ref.run_mcmc(ok[:W], 2, skip_initial_state_check=True)
    dsm = DistributedDeviceSampler(W, like.ndim, like, part, seed=7)
    dsm.run_mcmc(ok[:W], 2, skip_initial_state_check=True)
    d = np.max(np.abs(dsm.get_chain() - ref.get_chain()))
    print(f"[rank {rank}/{world}] C5a W={W}: max |chain - single-GPU chain| after 2 steps = {d:.3e}", flush=True)
    assert d <= 1e-9 and np.array_equal(dsm.acceptance_fraction, ref.acceptance_fraction)
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    nsteps = 40
    dsm.run_mcmc(None, nsteps)
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"C5a W={W} x {like.total_pixels} px on {world} GPU(s): {nsteps / dt:.1f} MCMC steps/s "
              f"({nsteps * W * like.total_pixels / dt:.3e} walker*px/s through the sampler), acceptance "
              f"{dsm.acceptance_fraction.mean():.3f}", flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
