#!/usr/bin/env python
"""Short driver for ncu / timing: device-resident stretch-move sampling on a small workload (C1 or C2)."""
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
    args = ap.parse_args()
    import numpy as np
    import bench
    from rbvfit_b200.sampler import DeviceEnsembleSampler
    w, models, like, thetas, spectra = bench.build_problem(args.workload, 0)
    ok = thetas[np.all((thetas >= w["lb"]) & (thetas <= w["ub"]), axis=1)]
    W = len(ok) - (len(ok) % 2)
    p0 = ok[:W]
    s = DeviceEnsembleSampler(W, like.ndim, like, seed=1, use_graph=not args.no_graph)
    s.run_mcmc(p0, 20, skip_initial_state_check=True)
    t0 = time.perf_counter()
    s.run_mcmc(None, args.steps)
    dt = time.perf_counter() - t0
    print(f"{args.workload} W={W} px={like.total_pixels} graph={not args.no_graph}: {args.steps / dt:.1f} steps/s "
          f"({dt / args.steps * 1e6:.1f} us/step), acceptance {s.acceptance_fraction.mean():.3f}")


if __name__ == "__main__":
    main()
