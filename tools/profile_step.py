#!/usr/bin/env python
"""Short driver for ncu: builds a workload and runs a few lnprob batches (device-resident theta)."""
import argparse
import os
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C5a")
    ap.add_argument("--walkers", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--precision", default="fp64")
    args = ap.parse_args()
    import torch
    import bench
    from rbvfit_b200 import workloads as wl
    w, models, like, thetas, spectra = bench.build_problem(args.workload, 0)
    if args.walkers and args.walkers != len(thetas):
        thetas = wl.make_ensemble(w, args.walkers)
    like.set_precision(args.precision)
    th = torch.as_tensor(thetas, device="cuda:0")
    evs = []
    for _ in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = like.lnprob_device(th)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    npx = like.total_pixels * len(thetas)
    print(f"[{args.precision}] {args.workload} W={len(thetas)} px={like.total_pixels} ms={['%.3f' % m for m in ms]} "
          f"-> {npx / (min(ms) * 1e-3):.4e} walker*px/s; finite={int(torch.isfinite(out).sum())}")


if __name__ == "__main__":
    main()
