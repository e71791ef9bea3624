#!/usr/bin/env python
"""GPU check of the far-field (Chebyshev) path against the direct evaluation: lnprob and model flux under both
modes on every workload, plus timings.  Development tool (run under gpurun); the parity tests proper live in
tests/test_gpu_parity.py / test_gpu_properties.py."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import build_problem  # noqa: E402


def timed(like, th_dev, n=5):
    for _ in range(2):
        like.lnprob_device(th_dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        like.lnprob_device(th_dev)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    out = {}
    for name, nw in (("C1", 50), ("C2", 80), ("C3", 50), ("C4", 32), ("C4w", 32), ("C5a", int(os.environ.get("C5A_W", "1024")))):
        w, models, like, thetas, spectra = build_problem(name, 0)
        from rbvfit_b200 import workloads as wl
        thetas = wl.make_ensemble(w, nw)
        th_dev = torch.as_tensor(thetas, device="cuda")
        res = {}
        for mode in ("direct", "chebyshev"):
            like.engine.set_farfield(mode)
            res[mode] = like.lnprob(thetas)
            res[mode + "_ms"] = timed(like, th_dev)
            first = list(models)[0]
            eng_flux = like.engine.model_flux(0, thetas[:4])
            res[mode + "_flux"] = eng_flux
        a, b = res["direct"], res["chebyshev"]
        fin = np.isfinite(a)
        rel = np.max(np.abs(a[fin] - b[fin]) / np.abs(a[fin]))
        dflux = np.nanmax(np.abs(res["direct_flux"] - res["chebyshev_flux"]))
        same = np.array_equal(np.isfinite(a), np.isfinite(b))
        px = like.total_pixels * nw
        out[name] = dict(lnprob_rel=float(rel), flux_abs=float(dflux), same_pattern=bool(same),
                         direct_ms=res["direct_ms"], cheb_ms=res["chebyshev_ms"],
                         direct_wpx_s=px / res["direct_ms"] * 1e3, cheb_wpx_s=px / res["chebyshev_ms"] * 1e3)
        print(name, json.dumps(out[name]), flush=True)
        like.close()


if __name__ == "__main__":
    main()
