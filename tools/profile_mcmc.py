#!/usr/bin/env python
"""Short driver for ncu / timing: device-resident sampling on a small workload (C1 or C2), stretch move
(rbv_stretch_run) or ensemble slice move (--slice, rbv_slice_run)."""
import argparse
import os
import sys
import time

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C1")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--slice", action="store_true", help="zeus-style ensemble slice move instead of the stretch move")
    args = ap.parse_args()
    import numpy as np
    import bench
    from rbvfit_b200.sampler import DeviceEnsembleSampler
    w, models, like, thetas, spectra = bench.build_problem(args.workload, 0)
    ok = thetas[np.all((thetas >= w["lb"]) & (thetas <= w["ub"]), axis=1)]
    W = len(ok) - (len(ok) % 2)
    p0 = ok[:W]
    if args.slice:
        from rbvfit_b200.slice_sampler import DeviceEnsembleSliceSampler
        W = max(2 * like.ndim + 8, 8)
        rng = np.random.default_rng(5)
        p0 = np.clip(w["theta_true"] + 1e-3 * rng.standard_normal((W, like.ndim)), w["lb"] + 1e-10, w["ub"] - 1e-10)
        s = DeviceEnsembleSliceSampler(W, like.ndim, like, seed=3)
        s.run_mcmc(p0, 40)
        c0, b0 = s.ncall, s.nbatches
        t0 = time.perf_counter()
        s.run_mcmc(None, args.steps)
        dt = time.perf_counter() - t0
        print(f"{args.workload} W={W} px={like.total_pixels} slice move: {args.steps / dt:.1f} steps/s "
              f"({dt / args.steps * 1e6:.1f} us/step), {(s.nbatches - b0) / args.steps:.1f} batches and "
              f"{(s.ncall - c0) / args.steps:.0f} lnprob rows per step, mu = {s.mu:.3f}")
        return
    s = DeviceEnsembleSampler(W, like.ndim, like, seed=1, use_graph=not args.no_graph)
    s.run_mcmc(p0, 20, skip_initial_state_check=True)
    t0 = time.perf_counter()
    s.run_mcmc(None, args.steps)
    dt = time.perf_counter() - t0
    print(f"{args.workload} W={W} px={like.total_pixels} graph={not args.no_graph}: {args.steps / dt:.1f} steps/s "
          f"({dt / args.steps * 1e6:.1f} us/step), acceptance {s.acceptance_fraction.mean():.3f}")


if __name__ == "__main__":
    main()
