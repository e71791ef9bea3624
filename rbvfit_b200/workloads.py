"""Synthetic workloads C1..C5 of SURVEY.md section 8(d) / BASELINE.json ``configs``.

Pure descriptions + seeded random inputs; the forward model used to make the "observed"
spectrum is passed in by the caller (tests pass the CPU oracle, ``bench.py`` passes the GPU
model), so this module depends on neither.

Every workload is a dict:
    systems      [(z, ion, [transitions], components), ...]   in add_system order
    instruments  {name: {"wave": f64[P], "FWHM": str|None, "lsf": None|"cos_like"}}
    theta_true, lb, ub   f64[ndim]        (bounds = reference set_bounds defaults,
                                            vfit_mcmc.py:761-766)
    nwalkers     ensemble size
"""
from __future__ import annotations

from typing import Callable, Dict

import numpy as np

SIGMA = 0.05


def default_bounds(nguess, bguess, vguess):
    """Traditional branch of set_bounds (vfit_mcmc.py:760-766)."""
    nguess, bguess, vguess = map(np.asarray, (nguess, bguess, vguess))
    lb = np.concatenate([nguess - 2.0, np.clip(bguess - 40.0, 2.0, None), vguess - 50.0])
    ub = np.concatenate([nguess + 2.0, np.clip(bguess + 40.0, None, 150.0), vguess + 50.0])
    return lb, ub


def _c1(P=2048):
    n, b, v = [14.2, 14.5], [40.0, 30.0], [-25.0, 35.0]
    lb, ub = default_bounds(n, b, v)
    return dict(
        name="C1", seed=20261,
        systems=[(0.348, "MgII", [2796.3, 2803.5], 2)],
        instruments={"COS": dict(wave=np.linspace(3755.0, 3795.0, P), FWHM="6.5", lsf=None)},
        theta_true=np.array(n + b + v), lb=lb, ub=ub, nwalkers=50)


def _c2_systems():
    out = []
    for z in (2.0, 2.3, 2.6):
        out.append((z, "CIV", [1548.2, 1550.77], 2))
        out.append((z, "SiIV", [1393.76, 1402.77], 1))
        out.append((z, "HI", [1215.67, 1025.72, 972.54, 949.74, 937.80], 1))
    return out


def _c2(P=20000, nwalkers=80, name="C2", seed=20262):
    rng = np.random.default_rng(seed)
    C = 12
    n = rng.uniform(13.0, 14.5, C)
    b = rng.uniform(10.0, 40.0, C)
    v = rng.uniform(-100.0, 100.0, C)
    lb, ub = default_bounds(n, b, v)
    return dict(
        name=name, seed=seed, systems=_c2_systems(),
        instruments={"SPEC": dict(wave=np.linspace(3300.0, 5700.0, P), FWHM="6.5", lsf=None)},
        theta_true=np.concatenate([n, b, v]), lb=lb, ub=ub, nwalkers=nwalkers)


def _c3():
    w = _c1()
    w.update(name="C3", seed=20263, instruments={
        "COS": dict(wave=np.linspace(3755.0, 3795.0, 4096), FWHM=None, lsf="cos_like"),
        "HIRES": dict(wave=np.linspace(3750.0, 3800.0, 16384), FWHM="3.0", lsf=None),
    })
    return w


def _c4(wide=False, P=20000):
    n, b, v = [21.0, 13.5, 13.2], [40.0, 12.0, 9.0], [0.0, -30.0, 25.0]
    lb, ub = default_bounds(n, b, v)
    inst = dict(wave=np.linspace(3500.0, 5500.0, P), FWHM=None if wide else "6.5",
                lsf="cos_like" if wide else None)
    return dict(
        name="C4w" if wide else "C4", seed=20264,
        systems=[(2.5, "HI", [1215.67, 1025.72], 1), (2.5, "SiII", [1526.71, 1304.37, 1260.42], 2)],
        instruments={"SPEC": inst},
        theta_true=np.array(n + b + v), lb=lb, ub=ub, nwalkers=32)


def get_workload(name: str) -> dict:
    if name == "C1":
        return _c1()
    if name == "C2":
        return _c2()
    if name == "C3":
        return _c3()
    if name == "C4":
        return _c4(False)
    if name == "C4w":
        return _c4(True)
    if name == "C5a":
        return _c2(P=100000, nwalkers=8192, name="C5a", seed=20265)
    if name == "C5a_L4":
        # "also report an L = 4 model at the same W, P" -- C1's MgII doublet on C5a's grid and ensemble
        w = _c1()
        w.update(name="C5a_L4", seed=20267, nwalkers=8192,
                 instruments={"SPEC": dict(wave=np.linspace(3300.0, 5700.0, 100000), FWHM="6.5", lsf=None)})
        return w
    raise KeyError(name)


def c5b_sightline(index: int) -> dict:
    """One of the 1024 independent sightlines of C5b: C1's structure at its own redshift."""
    rng = np.random.default_rng(20266 + 7919 * index)
    z = float(rng.uniform(0.3, 0.4))
    w = _c1()
    lo, hi = 2796.35 * (1 + z) - 20.0, 2803.53 * (1 + z) + 13.0
    w.update(name=f"C5b[{index}]", seed=20266 + 7919 * index,
             systems=[(z, "MgII", [2796.3, 2803.5], 2)],
             instruments={"COS": dict(wave=np.linspace(lo, hi, 2048), FWHM="6.5", lsf=None)},
             nwalkers=64)
    return w


def make_ensemble(w: dict, nwalkers=None, frac_out_of_bounds=0.01, seed_offset=100) -> np.ndarray:
    """theta_w = clip(theta_true + [0.05 dex, 1 km/s, 2 km/s] N(0,1)) with ~1 % of the rows
    pushed outside the bounds (exercises the -inf path)."""
    W = int(nwalkers or w["nwalkers"])
    rng = np.random.default_rng(w["seed"] + seed_offset)
    C = w["theta_true"].size // 3
    scale = np.concatenate([np.full(C, 0.05), np.full(C, 1.0), np.full(C, 2.0)])
    th = w["theta_true"][None, :] + scale[None, :] * rng.standard_normal((W, 3 * C))
    th = np.clip(th, w["lb"] + 1e-10, w["ub"] - 1e-10)
    n_out = int(round(frac_out_of_bounds * W)) if W >= 20 else 0
    if frac_out_of_bounds > 0 and W >= 4:
        n_out = max(n_out, 1)
    rows = rng.choice(W, size=n_out, replace=False) if n_out else []
    for r in rows:
        j = int(rng.integers(0, 3 * C))
        th[r, j] = w["ub"][j] + 1.0 if rng.random() < 0.5 else w["lb"][j] - 1.0
    return th


def make_spectra(w: dict, model_flux: Callable[[str, np.ndarray, np.ndarray], np.ndarray],
                 error_dtype=np.float64) -> Dict[str, dict]:
    """Observed spectra = model(theta_true) + N(0, sigma); ``model_flux(inst_name, theta, wave)``."""
    rng = np.random.default_rng(w["seed"])
    out = {}
    for name, inst in w["instruments"].items():
        wave = inst["wave"]
        truth = np.asarray(model_flux(name, w["theta_true"], wave), dtype=np.float64)
        flux = truth + SIGMA * rng.standard_normal(wave.size)
        error = np.full(wave.size, SIGMA, dtype=error_dtype)
        if error_dtype != np.float64:
            flux = flux.astype(error_dtype)
        out[name] = dict(wave=wave, flux=flux, error=error)
    return out
