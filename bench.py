#!/usr/bin/env python
"""Benchmark of the rbvfit likelihood hot path on B200 (BASELINE.json metric: walker.pixel lnprob evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C5a] [--impl reference]

One "step" = one lnprob evaluation of the whole walker ensemble of the workload (default C5a: 8192 walkers x
100 000 px, L = 33 lines, K = 23 LSF taps -- the configuration the metric is quoted on; it fits one GPU).
With N > 1 (launched by torchrun, one rank per GPU) the ensemble is split by walkers (strong scaling: the
total work per step is fixed) and every step ends with the all-gather of lnprob issued by the library: one kernel
over NVLink peer memory (`config.collective`; NCCL when the peer attach is unavailable or RBVFIT_B200_PEER=0).

`value`  device-resident throughput: theta already in HBM, CUDA events around each step on the launch
         stream, L2 flushed (untimed) between steps, max over ranks.
`e2e`    the same metric through the public API `lnprob(theta_host)` -> numpy: per step H2D of the theta rows
         from pinned memory, the kernel, the gather and the D2H of lnprob, timed by wall clock around
         synchronous calls.
`mcmc`   (N = 1) the second half of the metric, MCMC steps/s vs the CPU reference path: C1 stretch move
         (device-resident, host-driven, CPU serial / fork pool), the same move at C5a's own scale, and C2 with the
         zeus-style ensemble slice move (device-resident CUDA-graph WHILE loop, host-driven, CPU fork pool).
Other workloads: --workload C1 | C2 | C3 | C4 | C4w | C5a_L4, and C5b [--sightlines S] (survey mode: sightlines
sharded over the ranks with no collective; its `mcmc` key is the per-sightline ensembles sampled in lockstep).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "walker_pixel_lnprob_evals_per_sec"
UNIT = "walker*pixel/s"
# BASELINE.json configs 1-4 (+ SURVEY 8d's L = 4 companion of C5a): reported next to the headline in `configs`
CONFIG_WORKLOADS = ["C1", "C2", "C3", "C4", "C4w", "C5a_L4"]


# --------------------------------------------------------------------------------------------- helpers
def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with NVML during the timed region."""

    def __init__(self, index: int, period: float = 0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def build_problem(workload: str, device: int):
    """Synthetic spectra (model(theta_true) + noise, generated with the GPU model itself), the likelihood
    object and the walker ensemble."""
    from rbvfit_b200 import FitConfiguration, workloads as wl
    from rbvfit_b200 import lsf
    from rbvfit_b200.likelihood import GpuLikelihood
    from rbvfit_b200.model import GpuVoigtModel
    w = wl.get_workload(workload)
    cfg = FitConfiguration()
    for (z, ion, trans, comps) in w["systems"]:
        cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
    models = {}
    for name, inst in w["instruments"].items():
        taps = lsf.cos_like_taps(321) if inst.get("lsf") == "cos_like" else None
        models[name] = GpuVoigtModel(cfg, FWHM=inst["FWHM"], device=device, lsf_taps=taps)
    compiled = {n: m.compile() for n, m in models.items()}
    spectra = wl.make_spectra(w, lambda n, th, wave: compiled[n].model_flux(th, wave))
    inst_data = {n: dict(model=models[n], **spectra[n]) for n in models}
    like = GpuLikelihood(inst_data, w["lb"], w["ub"], device=device)
    thetas = wl.make_ensemble(w)
    return w, models, like, thetas, spectra


def oracle_problem(workload: str):
    """The same workload for the CPU arm (oracle port of the reference, numpy + scipy.special.wofz)."""
    from oracle import voigt_oracle as vo
    from rbvfit_b200 import workloads as wl
    w = wl.get_workload(workload)
    cfg = vo.OracleConfig()
    for (z, ion, trans, comps) in w["systems"]:
        cfg.add_system(z, ion, trans, comps)
    models = {}
    for name, inst in w["instruments"].items():
        taps = vo.cos_like_lsf(321) if inst.get("lsf") == "cos_like" else None
        models[name] = vo.lower(cfg, FWHM=inst["FWHM"], custom_taps=taps)
    spectra = wl.make_spectra(w, lambda n, th, wave: vo.model_flux(models[n], th, wave))
    comp = vo.compile_instruments({n: dict(model=models[n], **spectra[n]) for n in models})
    return w, comp, wl.make_ensemble(w)


def reference_problem(workload: str):
    """The workload built with the reference's OWN classes (FitConfiguration -> VoigtModel -> vfit), imported
    unmodified from the staged copy (oracle/_ref, see oracle/build_ref.py) behind oracle/refshim.py."""
    import contextlib
    import io
    from oracle import refshim, voigt_oracle as vo
    from rbvfit_b200 import workloads as wl
    FitConfiguration, VoigtModel, mc, _vm = refshim.import_reference()
    from astropy.convolution import CustomKernel
    w = wl.get_workload(workload)
    config = FitConfiguration()
    for (z, ion, trans, comps) in w["systems"]:
        config.add_system(z=z, ion=ion, transitions=list(trans), components=comps)
    models = {}
    for name, inst in w["instruments"].items():
        m = VoigtModel(config, FWHM=inst["FWHM"])
        if inst.get("lsf") == "cos_like":            # what _setup_kernel does with a linetools table (:458-460)
            m.kernel = CustomKernel(vo.cos_like_lsf(321))
        models[name] = m
    compiled = {n: m.compile() for n, m in models.items()}
    spectra = wl.make_spectra(w, lambda n, th, wave: compiled[n].model_flux(th, wave))
    with contextlib.redirect_stdout(io.StringIO()):
        fitter = mc.vfit({n: dict(model=models[n], **spectra[n]) for n in models}, w["theta_true"], w["lb"], w["ub"])
    return w, fitter, wl.make_ensemble(w), mc


def cpu_arm(workload: str, steps: int, warmup: int, sample_walkers=None):
    """Times the reference's CPU implementation of the path on a bounded walker sample of the workload, with every
    host core.  With the staged reference present (`kind` "reference"): the reference's own `vfit.lnprob` mapped over
    the rows by `mp.get_context('fork').Pool()` -- its `use_pool=True` path, vfit_mcmc.py:41-45, 408-414, which is
    what emcee does with `pool=pool` once per half-step (the bound method is pickled with every chunk).  Otherwise
    (`kind` "port") the oracle port under a fork pool that inherits the problem."""
    from oracle import refshim, voigt_oracle as vo
    cores = len(os.sched_getaffinity(0))
    live = refshim.available()
    if live:
        w, fitter, thetas, mc = reference_problem(workload)
        total_px = sum(len(d["wave"]) for d in fitter.instrument_data.values())
        work = sum(len(d["wave"]) * d["model"].__self__.data.n_lines for d in fitter.instrument_data.values())
    else:
        w, comp, thetas = oracle_problem(workload)
        total_px = sum(len(d["wave"]) for d in comp.values())
        work = sum(len(d["wave"]) * d["model"].n_lines for d in comp.values())
    if sample_walkers is None:
        # ~140 ns per (line, pixel) wofz on one core -> aim at ~2 s of wall clock per pool call
        per_walker = 1.4e-7 * work
        sample_walkers = int(min(len(thetas), max(cores, round(cores * 2.0 / max(per_walker, 1e-4)))))
    sample = thetas[:sample_walkers]
    times = []
    if live:
        pool = mc.OptimizedPool(processes=cores)                          # lives across steps, like emcee's
        rows = [row for row in sample]
        step = lambda: list(pool.map(fitter.lnprob, rows))               # noqa: E731
    else:
        pool = vo.make_pool(comp, w["lb"], w["ub"], processes=cores)
        step = lambda: vo.lnprob_pool(comp, sample, w["lb"], w["ub"], pool=pool)   # noqa: E731
    try:
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            step()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    finally:
        pool.close()
        pool.join()
    best = min(times)
    mean = sum(times) / len(times)
    how = ("the reference's own vfit.lnprob under mp.get_context('fork').Pool().map (unmodified rbvfit 2.4.0, "
           "numpy + scipy.special.wofz)") if live else "oracle port: numpy + scipy.special.wofz"
    return {"value": sample_walkers * total_px / mean, "best": sample_walkers * total_px / best,
            "ms_per_step": mean * 1e3, "cores": cores, "sample_walkers": sample_walkers,
            "total_px": total_px, "kind": "reference" if live else "port",
            "sample": f"{sample_walkers} of {len(thetas)} walkers x {total_px} px of {workload} per step, "
                      f"fork pool over {cores} cores ({how})"}


# --------------------------------------------------------------------------------------------- arms
def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, args.steps)          # exactly K timed steps after W warm-up steps; a step is ~2 s of pool work
    r = cpu_arm(args.workload, steps, max(0, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": max(0, args.warmup), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "sample": r["sample"]},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def measure_workload(workload, rank, world, local, steps, warmup, far_field="chebyshev", fp64_peak=None, flush=None):
    """One bench line's worth of numbers for `workload` on this rank's GPU (every rank calls it; the walker batch is
    partitioned over the ranks and the step ends with the all-gather of lnprob):
      value  device-resident throughput (theta in HBM, CUDA events per step, L2 flushed untimed between steps,
             max over ranks); counts all W rows, the ~1 % out-of-bounds rows included (they return -inf unevaluated)
      e2e    host theta -> host lnprob through the public API, H2D and D2H inside the timed region
      roofline  FP64-pipe bound, F = algorithmic flops per walker.pixel of the algorithm that runs."""
    import torch
    from rbvfit_b200 import dist as rdist
    from rbvfit_b200 import roofline as rf
    dev = torch.device("cuda", local)
    w, models, like, thetas, spectra = build_problem(workload, local)
    part = rdist.WalkerPartition(rank, world)
    if world > 1:
        like.engine.comm_init()                 # the lnprob all-gather is issued by the library, in-stream
    dlike = rdist.DistributedLikelihood(like, part)
    W, ndim = thetas.shape
    total_px = like.total_pixels
    lo, hi = part.rows(W)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            torch.distributed.barrier()

    if fp64_peak is None:
        fp64_peak = like.engine.measure_fp64_peak(300.0)          # TFLOP/s, DFMA chain, measured on this box
    name0 = like.names[0]
    data0 = models[name0].compile().data
    n_taps = 1 if data0.kernel is None else len(data0.kernel)
    F, F_direct, tiers, tiers_direct = 0.0, 0.0, {}, {}
    for n in like.names:                                        # pixel-weighted over instruments
        d = models[n].compile().data
        k = 1 if d.kernel is None else len(d.kernel)
        share = len(spectra[n]["wave"]) / total_px
        Fd, td = rf.flops_per_walker_pixel(d, w["theta_true"], spectra[n]["wave"], k)
        Ff, tf = rf.flops_farfield(d, w["theta_true"], spectra[n]["wave"], k)
        F_direct += Fd * share
        F += Ff * share
        tiers[n], tiers_direct[n] = tf, td
    if far_field == "direct":
        F, tiers = F_direct, tiers_direct
    like.engine.set_farfield(far_field)

    theta_dev = torch.as_tensor(thetas, device=dev)
    if flush is None:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2
    for _ in range(warmup):
        out = dlike.lnprob_device(theta_dev)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    launches0 = like.engine.launch_count
    barrier()
    for k in range(steps):
        flush.zero_()                                     # untimed L2 flush
        ev[k][0].record()
        out = dlike.lnprob_device(theta_dev)              # prep + lnprob kernel + finalize (+ all-gather, N > 1)
        ev[k][1].record()
    barrier()
    launches = like.engine.launch_count - launches0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(total_ms, op=torch.distributed.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / steps
    value = W * total_px / (ms_per_step * 1e-3)
    lnp = out.cpu().numpy()
    kernel = like.engine.last_kernel
    # the kernels alone on this rank (no collective): the roofline's denominator
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if hi > lo:
        rows = theta_dev[lo:hi].contiguous()
        for k in range(steps):
            flush.zero_()
            kev[k][0].record()
            like.lnprob_device(rows)
            kev[k][1].record()
        torch.cuda.synchronize(dev)
        k_ms = sum(a.elapsed_time(b) for a, b in kev) / steps
    else:
        k_ms = float("nan")

    # e2e through the public API (host buffers): theta sits in page-locked host memory, as the contract asks, so the
    # H2D copy reads it in place (a pageable array costs a staging copy on top -- reported as e2e.pageable)
    thetas_pin = like.pinned_theta(W)
    thetas_pin[:] = thetas
    for _ in range(min(warmup, 3)):
        dlike.lnprob(thetas_pin)
        dlike.lnprob(thetas)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res_pageable = dlike.lnprob(thetas)
    torch.cuda.synchronize(dev)
    e2e_pageable_ms = (time.perf_counter() - t0) * 1e3 / steps
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = dlike.lnprob(thetas_pin)
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(e2e_s, op=torch.distributed.ReduceOp.MAX)
    e2e_ms = float(e2e_s.item()) * 1e3 / steps
    e2e_value = W * total_px / (e2e_ms * 1e-3)
    assert np.array_equal(np.asarray(res), lnp, equal_nan=True), "e2e and device-resident results differ"

    n_local = hi - lo
    n_inb = int(np.count_nonzero(np.isfinite(lnp[lo:hi]) | np.isnan(lnp[lo:hi])))   # rows actually evaluated
    achieved = F * n_inb * total_px / (k_ms * 1e-3) / 1e12
    res = {
        "workload": workload, "value": value, "ms_per_step": ms_per_step, "kernel": kernel, "clocks": clocks,
        "launches": int(launches), "fp64_peak": fp64_peak,
        "config": {"workload": workload, "walkers": W, "pixels": total_px, "ndim": ndim,
                   "lines": int(sum(models[n].compile().data.n_lines for n in like.names)), "lsf_taps": n_taps,
                   "partition": f"walkers/{world}",
                   "collective": ("none (one rank)" if world == 1 else
                                  "all-gather of lnprob: one kernel over NVLink peer memory (rbv_peer_attach)"
                                  if like.engine.peer_attached else "all-gather of lnprob: ncclAllGather in-stream"),
                   "l2": "flushed between timed steps (256 MiB memset, untimed)",
                   "walkers_out_of_bounds": int(np.count_nonzero(np.isneginf(lnp))),
                   "value_counts": "all walkers, the out-of-bounds rows included (they return -inf unevaluated); "
                                   "roofline.achieved counts the evaluated rows only"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(n_local * ndim * 8), "d2h_bytes_per_step": int(W * 8),
                "input": "page-locked host array (GpuLikelihood.pinned_theta), read in place by the H2D copy",
                "pageable": {"value": W * total_px / (e2e_pageable_ms * 1e-3), "ms_per_step": e2e_pageable_ms,
                             "note": "this rank's wall clock with a pageable numpy theta (staging copy included)"}},
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak, "kernel": f"voigt_{kernel}_kernel", "kernel_ms": k_ms,
                     "far_field": far_field, "algorithmic_flops_per_walker_pixel": F, "tiers": tiers,
                     "direct_equivalent": {"algorithmic_flops_per_walker_pixel": F_direct,
                                           "tflops": F_direct * n_inb * total_px / (k_ms * 1e-3) / 1e12,
                                           "frac": F_direct * n_inb * total_px / (k_ms * 1e-3) / 1e12 / fp64_peak,
                                           "tiers": tiers_direct},
                     "hbm_algorithmic_gbs": (n_local * ndim * 8 + n_local * 8 + 4 * 8 * total_px) / (k_ms * 1e-3) / 1e9},
    }
    return res, (w, models, like, thetas, spectra, theta_dev, lnp, flush, F_direct, n_inb)


def run_gpu(args, rank, world, local):
    import torch
    torch.cuda.set_device(local)
    r, ctx = measure_workload(args.workload, rank, world, local, args.steps, args.warmup, args.far_field)
    w, models, like, thetas, spectra, theta_dev, lnp, flush, F_direct, n_inb = ctx
    W, total_px, fp64_peak = thetas.shape[0], like.total_pixels, r["fp64_peak"]
    peaks, peaks_kind = _peaks()
    prof = {}
    ppath = os.path.join(ROOT, "profiles", "latest_ncu_summary.json")
    if os.path.exists(ppath):
        with open(ppath) as fh:
            prof = json.load(fh)
    roof = dict(r["roofline"])
    roof.update({
        "traffic": prof.get("dram_bytes_per_launch"),
        "traffic_capture": None if not prof else f"{prof.get('report')}: {prof.get('note')}",
        "peak_source": "DFMA dependent-chain probe measured in this run (rbv_measure_fp64_peak); "
                       "MEASURED_PEAKS.json has no FP64 entry (spec: 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2)",
        "algorithm_note": "far wings of each 1024-px super-chunk are summed at 8 Chebyshev nodes and interpolated "
                          "(a-priori gated, |dtau| <= 1e-12): F counts THAT algorithm "
                          "(rbvfit_b200/roofline.py:flops_farfield); direct_equivalent applies SURVEY 8(d)'s "
                          "per-(line,pixel) rule to the same throughput and may exceed the peak",
        "hbm_sanity": {"algorithmic_gbs": roof.pop("hbm_algorithmic_gbs"), "peak_gbs": peaks.get("hbm_gbs"),
                       "peaks": peaks_kind}})
    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": r["config"], "clocks": r["clocks"],
        "e2e": r["e2e"],
        "gpu_launches": r["launches"],      # prep_kernel + lnprob kernel + finalize kernel (+ all-gather kernel) per step
        "roofline": roof,
    }
    extras = not args.no_extras
    # ---- MCMC steps/s at the bench workload's own scale, on all N GPUs (every rank takes part)
    if extras:
        large = mcmc_large_leg(like, w, thetas, total_px, rank, world)
    # ---- the other BASELINE configurations (parity-test sizes, milliseconds each), same measurement
    configs = {}
    if extras and args.workload == "C5a":
        for name in CONFIG_WORKLOADS:
            c, cctx = measure_workload(name, rank, world, local, max(5, args.steps // 2), 3, args.far_field,
                                       fp64_peak=fp64_peak, flush=flush)
            cctx[2].close()
            configs[name] = {"value": c["value"], "unit": UNIT, "ms_per_step": c["ms_per_step"],
                             "e2e": c["e2e"]["value"], "kernel": c["kernel"], "config": c["config"],
                             "roofline": {k: c["roofline"][k] for k in
                                          ("achieved", "peak", "frac", "kernel_ms",
                                           "algorithmic_flops_per_walker_pixel")},
                             "roofline_direct_rule": {
                                 "algorithmic_flops_per_walker_pixel":
                                     c["roofline"]["direct_equivalent"]["algorithmic_flops_per_walker_pixel"],
                                 "frac": c["roofline"]["direct_equivalent"]["frac"]}}
        configs["C5b"] = sightline_leg(args, rank, world, local, fp64_peak, flush,
                                       steps=max(5, args.steps // 2), with_mcmc=False)
    if rank != 0:
        return
    if configs:
        line["configs"] = configs
    if world == 1 and extras:
        if args.far_field == "chebyshev":
            line["direct_far_wings"] = direct_leg(like, theta_dev, lnp, W, total_px, args.steps, flush, F_direct,
                                                  n_inb, fp64_peak)
        line["fp32_gated"] = fp32_gated_leg(like, theta_dev, thetas, W, total_px, args.steps, flush)
        line["mcmc"] = mcmc_leg(local, with_cpu=not args.no_cpu)
        line["mcmc"]["zeus"] = mcmc_zeus_leg(local, with_cpu=not args.no_cpu)
    if extras:
        line.setdefault("mcmc", {})["large_ensemble"] = large
    if not args.no_cpu and world == 1:      # the CPU baseline is reported at N = 1 only
        r2 = cpu_arm(args.workload, steps=2, warmup=1)
        line["cpu_baseline"] = {"value": r2["value"], "unit": UNIT, "cores": r2["cores"], "kind": r2["kind"],
                                "sample": r2["sample"]}
    print(json.dumps(line))


def build_sightlines(first, count, device, walkers=64):
    """C5b: `count` independent sightlines starting at index `first` (C1's structure at z ~ U(0.3, 0.4), own noise
    realisation, 2048 px, `walkers` walkers each)."""
    from rbvfit_b200 import FitConfiguration, workloads as wl
    from rbvfit_b200.likelihood import SightlineBatch
    from rbvfit_b200.model import GpuVoigtModel
    sight, thetas = [], []
    w0 = wl.c5b_sightline(0)
    for sidx in range(first, first + count):
        w = wl.c5b_sightline(sidx)
        cfg = FitConfiguration()
        for (z, ion, trans, comps) in w["systems"]:
            cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
        m = GpuVoigtModel(cfg, FWHM="6.5", device=device)
        # truth spectrum without going through a per-sightline flux engine: unit continuum + noise is enough
        # for a throughput workload whose cost does not depend on the data values
        rng = np.random.default_rng(w["seed"])
        wave = w["instruments"]["COS"]["wave"]
        sight.append(dict(model=m, wave=wave, flux=1.0 + wl.SIGMA * rng.standard_normal(wave.size),
                          error=np.full(wave.size, wl.SIGMA)))
        thetas.append(wl.make_ensemble(w, walkers))
    return SightlineBatch(sight, w0["lb"], w0["ub"], device=device), np.array(thetas)


def sightline_leg(args, rank, world, local, fp64_peak=None, flush=None, steps=None, with_mcmc=True,
                  n_sightlines=None, walkers=64):
    """Workload C5b (survey mode): S independent sightlines (C1's structure at its own redshift, 2048 px, 64 walkers
    each) sharded across the ranks -- rank r owns a contiguous block, no collective on the data path, only the final
    max over ranks of the elapsed time.  Returns the result dict on every rank."""
    import torch
    from rbvfit_b200 import roofline as rf
    from rbvfit_b200.dist import SightlinePartition
    dev = torch.device("cuda", local)
    n_sightlines = n_sightlines or args.sightlines
    steps = steps or args.steps
    first, count = SightlinePartition(rank, world).owned(n_sightlines)
    batch, thetas = build_sightlines(first, count, local, walkers)
    S, Ws, ndim = thetas.shape
    P = batch.pixels
    th_dev = torch.as_tensor(thetas.reshape(S * Ws, ndim), device=dev)
    if flush is None:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    if fp64_peak is None:
        fp64_peak = batch.engine.measure_fp64_peak(300.0)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            torch.distributed.barrier()
    for _ in range(max(args.warmup, 3)):
        out = batch.lnprob_device(th_dev, Ws)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    l0 = batch.engine.launch_count
    barrier()
    for k in range(steps):
        flush.zero_()
        ev[k][0].record()
        out = batch.lnprob_device(th_dev, Ws)
        ev[k][1].record()
    barrier()
    launches = batch.engine.launch_count - l0
    clocks = sampler.stop()
    local_ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    total_ms = torch.tensor([local_ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(total_ms, op=torch.distributed.ReduceOp.MAX)
    ms = float(total_ms.item())
    value = n_sightlines * Ws * P / (ms * 1e-3)
    # e2e: host theta (page-locked, read in place by the H2D copy) -> host lnprob through SightlineBatch.lnprob
    thetas_pin = batch.pinned_theta(Ws)
    thetas_pin[:] = thetas
    for _ in range(2):
        batch.lnprob(thetas_pin)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = batch.lnprob(thetas_pin)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(e2e_s, op=torch.distributed.ReduceOp.MAX)
    e2e_ms = float(e2e_s.item()) * 1e3 / steps
    # roofline of this rank's launch: F of the sightline structure (C1's lines on a 2048-px grid), evaluated rows only
    from rbvfit_b200 import FitConfiguration, workloads as wl
    from rbvfit_b200.model import GpuVoigtModel
    w0 = wl.c5b_sightline(first)
    cfg = FitConfiguration()
    for (z, ion, trans, comps) in w0["systems"]:
        cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
    d0 = GpuVoigtModel(cfg, FWHM="6.5", device=local).compile().data
    wave0 = w0["instruments"]["COS"]["wave"]
    F, _t = rf.flops_farfield(d0, w0["theta_true"], wave0, 23)
    F_direct, _t = rf.flops_per_walker_pixel(d0, w0["theta_true"], wave0, 23)
    n_inb = int(np.count_nonzero(np.isfinite(res)))
    achieved = F * n_inb * P / (local_ms * 1e-3) / 1e12
    result = {
        "value": value, "unit": UNIT, "ms_per_step": ms, "kernel": batch.engine.last_kernel, "clocks": clocks,
        "launches": int(launches),
        "e2e": {"value": n_sightlines * Ws * P / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(S * Ws * ndim * 8), "d2h_bytes_per_step": int(S * Ws * 8)},
        "config": {"workload": "C5b", "sightlines": n_sightlines, "walkers_per_sightline": Ws, "pixels": P,
                   "lines": 4, "lsf_taps": 23, "partition": f"sightlines/{world}",
                   "l2": "flushed between timed steps (256 MiB memset, untimed)"},
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak, "kernel": f"voigt_{batch.engine.last_kernel}_kernel",
                     "kernel_ms": local_ms, "algorithmic_flops_per_walker_pixel": F,
                     "direct_equivalent": {"algorithmic_flops_per_walker_pixel": F_direct,
                                           "frac": F_direct * n_inb * P / (local_ms * 1e-3) / 1e12 / fp64_peak},
                     "note": "rank 0's launch over its own sightlines; F of sightline 0 (every pixel of this workload "
                             "lies within 45 Doppler widths of all four lines: no far field to exploit)"},
        "finite_fraction": float(np.isfinite(res).mean())}
    if with_mcmc:
        # MCMC in survey mode: one stretch-move ensemble per sightline, all in lockstep on the device
        # (rbv_stretch_run_sightlines); every rank samples its own sightlines, no collective
        from rbvfit_b200.sampler import SightlineEnsembleSampler
        p0 = thetas.copy()
        bad = ~np.all((p0 >= batch.lb) & (p0 <= batch.ub), axis=2)
        p0[bad] = np.clip(p0[bad], batch.lb + 1e-9, batch.ub - 1e-9)
        smp = SightlineEnsembleSampler(Ws, ndim, batch, seed=6)
        smp.run_mcmc(p0, 3, skip_initial_state_check=True)
        mc_steps = 20
        barrier()
        t0 = time.perf_counter()
        smp.run_mcmc(None, mc_steps)
        mc_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(mc_s, op=torch.distributed.ReduceOp.MAX)
        mc_sps = mc_steps / float(mc_s.item())
        result["mcmc"] = {"sampler": "device-resident stretch move, one ensemble per sightline in lockstep "
                                     "(rbv_stretch_run_sightlines), chain D2H included",
                          "steps_per_sec": mc_sps, "sightline_steps_per_sec": mc_sps * n_sightlines,
                          "walker_pixel_per_sec": mc_sps * n_sightlines * Ws * P,
                          "acceptance": float(smp.acceptance_fraction.mean())}
    batch.close()
    return result


def run_gpu_sightlines(args, rank, world, local):
    """`--workload C5b`: the survey-mode line on its own."""
    import torch
    torch.cuda.set_device(local)
    r = sightline_leg(args, rank, world, local)
    if rank != 0:
        return
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": r["config"], "clocks": r["clocks"],
            "e2e": r["e2e"], "gpu_launches": r["launches"], "roofline": r["roofline"],
            "finite_fraction": r["finite_fraction"]}
    if "mcmc" in r:
        line["mcmc"] = r["mcmc"]
    print(json.dumps(line))


def direct_leg(like, theta_dev, lnp_default, W, total_px, steps, flush, F_direct, n_inb, fp64_peak):
    """The same batch with the far field switched off (every (line, pixel) pair evaluated on its own): the kernel
    SURVEY.md 8(d)'s flop rule describes, timed exactly like `value`, plus its agreement with the default path."""
    import torch
    like.engine.set_farfield("direct")
    try:
        for _ in range(2):
            out = like.lnprob_device(theta_dev)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            flush.zero_()
            ev[k][0].record()
            out = like.lnprob_device(theta_dev)
            ev[k][1].record()
        torch.cuda.synchronize()
    finally:
        like.engine.set_farfield("chebyshev")
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    got = out.cpu().numpy()
    fin = np.isfinite(lnp_default)
    dev = float(np.max(np.abs(got[fin] - lnp_default[fin]) / np.abs(lnp_default[fin]))) if fin.any() else 0.0
    tf = F_direct * n_inb * total_px / (ms * 1e-3) / 1e12
    return {"value": W * total_px / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
            "algorithmic_flops_per_walker_pixel": F_direct, "achieved_tflops": tf, "frac": tf / fp64_peak,
            "max_rel_dev_vs_default": dev}


def fp32_gated_leg(like, theta_dev, thetas, W, total_px, steps, flush):
    """The FP32-gated far-wing variant (north_star): kept only if it passes the tolerance check against the
    FP64 kernel on a walker sample; timed exactly like `value`."""
    import torch
    ok = like.set_precision("fp32-gated", check_thetas=thetas[:256])
    out = {"accepted": bool(ok), "max_rel_dev_vs_fp64": like.last_precision_check, "tolerance": 1e-9,
           "gate": "per pixel the FP32-evaluated far-wing contributions sum to <= 4e-6 (|dtau| <= 1e-11)"}
    if ok:
        for _ in range(3):
            like.lnprob_device(theta_dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        for k in range(steps):
            flush.zero_()
            ev[k][0].record()
            like.lnprob_device(theta_dev)
            ev[k][1].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / steps
        out.update(value=W * total_px / (ms * 1e-3), unit=UNIT, ms_per_step=ms)
    like.set_precision("fp64")
    return out


def mcmc_large_leg(like, w, thetas, total_px, rank=0, world=1, nsteps=48):
    """MCMC steps/s at the bench workload's own scale (C5a: ~8000 walkers x 100 000 px) on all N GPUs: the
    device-resident stretch move on the in-bounds rows of the bench ensemble; one step = every walker updated once =
    one full ensemble evaluation.  N = 1: rbv_stretch_run (one CUDA graph per step).  N > 1: rbv_stretch_run_dist --
    the rows of every half-step are split over the ranks and the NCCL all-gather of their lnprob sits inside the
    captured step; the chain is the single-GPU chain bit for bit (its digest is printed for comparison across N)."""
    import torch
    from rbvfit_b200 import dist as rdist
    from rbvfit_b200.sampler import DeviceEnsembleSampler, DistributedDeviceSampler
    ok = thetas[np.all((thetas >= w["lb"]) & (thetas <= w["ub"]), axis=1)]
    W = len(ok) - (len(ok) % 2)
    if world > 1:
        smp = DistributedDeviceSampler(W, like.ndim, like, rdist.WalkerPartition(rank, world), seed=4)
    else:
        smp = DeviceEnsembleSampler(W, like.ndim, like, seed=4)
    smp.run_mcmc(ok[:W], 4, skip_initial_state_check=True)
    torch.cuda.synchronize()
    rates = []
    for _ in range(3):                 # median of three runs: one run is 0.1 s of wall clock, a host stall doubles it
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        smp.run_mcmc(None, nsteps)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=like.engine.tdev)
        if world > 1:
            torch.distributed.all_reduce(dt, op=torch.distributed.ReduceOp.MAX)
        rates.append(nsteps / float(dt.item()))
    sps = sorted(rates)[1]
    digest = float(np.sum(smp.get_chain()[:4] * np.arange(1, 5)[:, None, None]))
    return {"walkers": int(W), "pixels": int(total_px), "steps": nsteps, "n_gpus": world, "steps_per_sec": sps,
            "steps_per_sec_runs": rates, "walker_pixel_per_sec": sps * W * total_px,
            "acceptance": float(smp.acceptance_fraction.mean()),
            "chain_digest_first_4_steps": repr(digest),
            "sampler": "rbv_stretch_run (CUDA graph)" if world == 1 else
                       "rbv_stretch_run_dist (CUDA graph incl. the in-place all-gather of lnprob per half-step)",
            "note": "median of 3 runs, max over ranks; includes the D2H copy of the chain (W x ndim x 8 B per step)"}


def mcmc_zeus_leg(device, with_cpu=True, nsteps=200, cpu_steps=2):
    """MCMC steps/s of the configuration BASELINE.json pairs with the zeus sampler: C2 (3 redshifts x CIV + SiIV + HI
    Lyman series, 33 lines, 36 parameters, 20 000 px, 80 walkers), ensemble slice sampling.  Device-resident loop
    (rbv_slice_run, what vfit.runmcmc(sampler='zeus') uses) vs the host-driven sampler on GPU batches vs the same
    host-driven sampler on the CPU oracle under a fork pool.  One step = every walker moved once (~6 likelihood
    evaluations per walker)."""
    from rbvfit_b200.slice_sampler import DeviceEnsembleSliceSampler, EnsembleSliceSampler
    w, _models, like, thetas, spectra = build_problem("C2", device)
    ok = thetas[np.all((thetas >= w["lb"]) & (thetas <= w["ub"]), axis=1)]
    W = 2 * like.ndim + 8
    rng = np.random.default_rng(5)
    p0 = np.clip(w["theta_true"] + 1e-3 * rng.standard_normal((W, like.ndim)), w["lb"] + 1e-10, w["ub"] - 1e-10)
    p0[:min(W, len(ok)) // 2] = ok[:min(W, len(ok)) // 2]          # half of the bench ensemble's own rows
    dev = DeviceEnsembleSliceSampler(W, like.ndim, like, seed=3)
    dev.run_mcmc(p0, 40)                                           # mu adapts here
    calls0, batches0, rates = dev.ncall, dev.nbatches, []
    for _ in range(3):
        t0 = time.perf_counter()
        dev.run_mcmc(None, nsteps)
        rates.append(nsteps / (time.perf_counter() - t0))
    total_px = sum(len(s["wave"]) for s in spectra.values())
    evals = (dev.ncall - calls0) / (3.0 * nsteps)
    host = EnsembleSliceSampler(W, like.ndim, like.lnprob, seed=3)
    host.run_mcmc(p0, 40)
    hb0 = host.nbatches
    t0 = time.perf_counter()
    host.run_mcmc(None, nsteps // 2)
    host_sps = (nsteps // 2) / (time.perf_counter() - t0)
    out = {"workload": f"C2 ({W} walkers x {total_px} px, L=33, ndim={like.ndim}, ensemble slice sampling)",
           "steps_per_sec": sorted(rates)[1], "steps_per_sec_runs": rates,
           "lnprob_rows_per_step": evals, "walker_pixel_per_sec": sorted(rates)[1] * evals * total_px,
           "batches_per_step": (dev.nbatches - batches0) / (3.0 * nsteps), "mu": dev.mu,
           "sampler": "device-resident (rbv_slice_run), chain D2H included; median of 3 runs",
           "host_driven_steps_per_sec": host_sps,
           "host_driven_batches_per_step": (host.nbatches - hb0) / float(nsteps // 2)}
    if with_cpu:
        from oracle import voigt_oracle as vo
        _w, ocomp, _t = oracle_problem("C2")
        cores = len(os.sched_getaffinity(0))
        pool = vo.make_pool(ocomp, w["lb"], w["ub"], processes=cores)
        try:
            cpu = EnsembleSliceSampler(W, like.ndim, lambda th: vo.lnprob_pool(ocomp, th, w["lb"], w["ub"], pool=pool),
                                       seed=3, mu=host.mu, tune=False)
            cpu.run_mcmc(p0, 1)
            t0 = time.perf_counter()
            cpu.run_mcmc(None, cpu_steps)
            out["cpu_pool_steps_per_sec"] = cpu_steps / (time.perf_counter() - t0)
            out["cpu_cores"] = cores
        finally:
            pool.close()
            pool.join()
    return out


def mcmc_leg(device, with_cpu=True, nsteps=300, cpu_steps=12):
    """Second half of the metric: MCMC steps/s (one step = every walker updated once) on config C1
    (MgII doublet, 2 components, 2048 px, 50 walkers, stretch move) -- GPU-batched sampler through the public
    ``vfit`` API vs the same sampler driven by the CPU oracle (serial and fork pool)."""
    import contextlib
    import io
    from rbvfit_b200 import FitConfiguration, workloads as wl
    from rbvfit_b200.model import GpuVoigtModel
    from rbvfit_b200.sampler import EnsembleSampler
    from rbvfit_b200.vfit_mcmc import vfit
    w = wl.get_workload("C1")
    cfg = FitConfiguration()
    for (z, ion, trans, comps) in w["systems"]:
        cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
    model = GpuVoigtModel(cfg, FWHM="6.5", device=device)
    comp = model.compile()
    spectra = wl.make_spectra(w, lambda n, th, wave: comp.model_flux(th, wave))
    s = spectra["COS"]
    fitter = vfit({"COS": dict(model=model, **s)}, w["theta_true"], w["lb"], w["ub"], no_of_Chain=50,
                  no_of_steps=nsteps, seed=1, device=device)
    p0 = fitter._initialize_walkers(w["theta_true"])
    smp = EnsembleSampler(50, 6, fitter.lnprob, seed=2)
    smp.run_mcmc(p0, 20)
    t0 = time.perf_counter()
    smp.run_mcmc(None, nsteps)
    gpu_sps = nsteps / (time.perf_counter() - t0)
    # the same move with the whole loop on the device (rbv_stretch_run, CUDA graph replay): what vfit.runmcmc uses
    from rbvfit_b200.sampler import DeviceEnsembleSampler
    dsm = DeviceEnsembleSampler(50, 6, fitter._like, seed=3)
    dsm.run_mcmc(p0, 50)
    dev_steps, rates = 1000, []
    for _ in range(3):                                   # median of three runs (launch jitter on a shared host)
        t0 = time.perf_counter()
        dsm.run_mcmc(None, dev_steps)
        rates.append(dev_steps / (time.perf_counter() - t0))
    dev_sps = sorted(rates)[1]
    out = {"workload": "C1 (50 walkers x 2048 px, L=4, stretch move)", "steps_per_sec": dev_sps,
           "walker_pixel_per_sec": dev_sps * 50 * 2048, "acceptance": float(dsm.acceptance_fraction.mean()),
           "sampler": "device-resident (rbv_stretch_run, CUDA graph), chain D2H included; median of 3 x 1000 steps",
           "steps_per_sec_runs": rates,
           "host_driven_steps_per_sec": gpu_sps, "host_driven_acceptance": float(smp.acceptance_fraction.mean())}
    if with_cpu:
        from oracle import voigt_oracle as vo
        ocfg = vo.OracleConfig()
        for (z, ion, trans, comps) in w["systems"]:
            ocfg.add_system(z, ion, trans, comps)
        om = vo.lower(ocfg, FWHM="6.5")
        ocomp = vo.compile_instruments({"COS": dict(model=om, **s)})
        cpu = EnsembleSampler(50, 6, lambda th: vo.lnprob_batch(ocomp, th, w["lb"], w["ub"]), seed=2)
        cpu.run_mcmc(p0, 2)
        t0 = time.perf_counter()
        cpu.run_mcmc(None, cpu_steps)
        out["cpu_serial_steps_per_sec"] = cpu_steps / (time.perf_counter() - t0)
        cores = len(os.sched_getaffinity(0))
        pool = vo.make_pool(ocomp, w["lb"], w["ub"], processes=cores)
        try:
            cpup = EnsembleSampler(50, 6, lambda th: vo.lnprob_pool(ocomp, th, w["lb"], w["ub"], pool=pool), seed=2)
            cpup.run_mcmc(p0, 2)
            t0 = time.perf_counter()
            cpup.run_mcmc(None, cpu_steps)
            out["cpu_pool_steps_per_sec"] = cpu_steps / (time.perf_counter() - t0)
            out["cpu_cores"] = cores
        finally:
            pool.close()
            pool.join()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--far-field", default="chebyshev", choices=["chebyshev", "direct"],
                    help="far-wing accumulation: interpolated far field (default) or every pair evaluated directly")
    ap.add_argument("--workload", default="C5a")
    ap.add_argument("--sightlines", type=int, default=1024, help="number of sightlines of workload C5b")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp32-gated and MCMC legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world_env == 1 and args.impl == "ours":
        # not under torchrun: relaunch ourselves with one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return
    from rbvfit_b200 import dist as rdist
    rank, world, local = rdist.init_from_env("nccl" if world_env > 1 else None)
    try:
        if args.workload == "C5b":
            run_gpu_sightlines(args, rank, world, local)
        else:
            run_gpu(args, rank, world, local)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
