#!/usr/bin/env python
"""Per-warp timeline of one voigt_stream_kernel launch (experiments only).

Needs a library built with -DRBV_STREAM_TIMELINE (RBVFIT_B200_LIB=build/variants/lib_timeline.so): every warp records
when it drew its first ticket, when it finished its last item, and how many items / segments it ran.  Prints how the
launch ends: the spread of the warps' finish times against the launch length, per SM and overall.
"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C5a")
    ap.add_argument("--walkers", type=int, nargs="+", default=[1024, 8192])
    args = ap.parse_args()
    import torch
    import bench
    from rbvfit_b200 import _lib
    from rbvfit_b200 import workloads as wl
    lib = _lib.load()
    w, models, like, thetas, spectra = bench.build_problem(args.workload, 0)
    n_warps = 148 * 2 * 8
    buf = np.zeros(4 * n_warps, dtype=np.uint64)
    for W in args.walkers:
        th = torch.as_tensor(wl.make_ensemble(w, W), device="cuda:0")
        for _ in range(3):
            like.lnprob_device(th)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        like.lnprob_device(th)
        b.record()
        torch.cuda.synchronize()
        rc = lib.rbv_debug_stream_timeline(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(buf.size))
        assert rc == 0
        t = buf.reshape(n_warps, 4).astype(np.float64)
        t0 = t[:, 0].min()
        start = (t[:, 0] - t0) * 1e-3
        end = (t[:, 1] - t0) * 1e-3
        length = end.max()
        print(f"W={W}: event time {a.elapsed_time(b) * 1e3:.1f} us, kernel span {length:.1f} us; warps start "
              f"{start.min():.1f}..{start.max():.1f} us")
        q = np.percentile(end, [0, 1, 10, 25, 50, 75, 90, 99, 100])
        print("   finish-time percentiles (us): " + " ".join(f"{x:.1f}" for x in q))
        print(f"   mean finish {end.mean():.1f} us = {end.mean() / length:.3f} of the span "
              f"(idle tail {1 - end.mean() / length:.3f}); items/warp {t[:, 2].mean():.2f} "
              f"(min {t[:, 2].min():.0f} max {t[:, 2].max():.0f}); segments/warp mean {t[:, 3].mean():.1f} "
              f"min {t[:, 3].min():.0f} max {t[:, 3].max():.0f}")
        # per SM (CTA b runs on some SM; two CTAs per SM -- group CTAs by pairs is not exact, report per CTA)
        per_cta = end.reshape(-1, 8).max(axis=1)
        print(f"   per-CTA last finish: min {per_cta.min():.1f} median {np.median(per_cta):.1f} max {per_cta.max():.1f} us")
        seg_rate = t[:, 3] / np.maximum(end - start, 1e-9)
        print(f"   segments per us per warp: min {seg_rate.min():.4f} median {np.median(seg_rate):.4f} max {seg_rate.max():.4f}")


if __name__ == "__main__":
    main()
