"""TEST INFRASTRUCTURE ONLY -- numpy/scipy restatement of rbvfit's likelihood hot path.

Every function cites the reference lines it follows (paths relative to /root/reference).
The arithmetic (operation order, dtypes, third-party calls) is kept identical to the
reference so that results agree to rounding; ``tests/test_oracle_vs_golden.py`` pins this
module against fixtures produced by the real reference code (``oracle/make_golden.py``).

Third-party arithmetic on this path:
  * scipy.special.wofz           (present in this image; same call as the reference)
  * scipy.ndimage.convolve1d     (present; same call as the reference)
  * astropy.convolution          (ABSENT -> restated below from astropy >= 5.3 documented
                                  behaviour; parity on this input is unpinned)
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
from scipy import ndimage
from scipy.special import wofz

_LINES_JSON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rbvfit_b200", "data",
                           "atomic_lines.json")
_line_cache = None


# --------------------------------------------------------------------------- atomic data
def _load_lines():
    """Line list = src/rbvfit/lines/atom_full.dat (name, wrest, f, gamma), shipped as JSON by
    tools/make_line_table.py.  Follows rb_setline.py:38-52: wavelengths float64, f and gamma
    stored as **float32**."""
    global _line_cache
    if _line_cache is None:
        with open(_LINES_JSON) as fh:
            rows = json.load(fh)["lines"]
        wave = np.array([float(r[1]) for r in rows], dtype=np.float64)
        fval = np.zeros(len(rows), dtype=np.float32)
        gamma = np.zeros(len(rows), dtype=np.float32)
        name = np.empty(len(rows), dtype=object)
        for i, r in enumerate(rows):
            fval[i] = float(r[2])
            gamma[i] = float(r[3])
            name[i] = f"{r[0]} {int(float(r[1]))}"
        _line_cache = (wave, fval, gamma, name)
    return _line_cache


def rb_setline(lambda_rest: float, method: str = "closest") -> dict:
    """rb_setline.py:25-64 ('atom' list only)."""
    wave, fval, gamma, name = _load_lines()
    if method == "Exact":
        idx = np.where(np.abs(lambda_rest - wave) < 1e-3)
    elif method == "closest":
        idx = np.array([np.abs(lambda_rest - wave).argmin()])
    else:
        raise ValueError("Specify a valid matching method: 'closest' or 'Exact'")
    return {"wave": wave[idx], "fval": fval[idx], "name": name[idx], "gamma": gamma[idx]}


# --------------------------------------------------------------------------- configuration
@dataclass
class OracleIonGroup:
    ion_name: str
    transitions: List[float]
    components: int
    redshift: float


@dataclass
class OracleSystem:
    redshift: float
    ion_groups: List[OracleIonGroup] = field(default_factory=list)


class OracleConfig:
    """The part of FitConfiguration the hot path reads (core/fit_configuration.py:318-351,
    380-391; IonGroup.validate_transitions :79-133 snaps every transition to the database
    wavelength)."""

    def __init__(self):
        self.systems: List[OracleSystem] = []

    def add_system(self, z, ion, transitions, components=1):
        system = None
        for s in self.systems:
            if abs(s.redshift - z) < 1e-6:
                system = s
                break
        if system is None:
            system = OracleSystem(z)
            self.systems.append(system)
        snapped = [rb_setline(w, "closest")["wave"][0] for w in transitions]
        system.ion_groups.append(OracleIonGroup(ion, snapped, int(components), z))
        return self


@dataclass
class LoweredModel:
    """CompiledModelData (core/voigt_model.py:265-280)."""
    atomic_lambda0: np.ndarray
    atomic_gamma: np.ndarray      # float32, as in the reference
    atomic_f: np.ndarray          # float32, as in the reference
    z_factors: np.ndarray
    N_indices: np.ndarray
    b_indices: np.ndarray
    v_indices: np.ndarray
    kernel_taps: Optional[np.ndarray]   # None | array (already what ndimage/astropy would apply)
    kernel_kind: str                    # 'none' | 'gaussian' | 'custom'
    n_lines: int
    total_components: int
    voigt_method: str = "wofz"


def gaussian_kernel_taps(fwhm) -> np.ndarray:
    """astropy.convolution.Gaussian1DKernel(stddev=FWHM/2.355).array restated
    (reference call site core/voigt_model.py:462-464).

    astropy >= 5.3: size = ceil(8*stddev) bumped to the next odd integer; taps = Gaussian1D
    (amplitude 1/(sqrt(2 pi) stddev)) sampled at integer offsets (discretisation mode
    'center'); the array is then normalised to unit sum."""
    sigma = float(fwhm) / 2.355
    size = int(math.ceil(8 * sigma))
    if size % 2 == 0:
        size += 1
    half = size // 2
    x = np.arange(-half, half + 1, dtype=np.float64)
    arr = (1.0 / (np.sqrt(2 * np.pi) * sigma)) * np.exp(-0.5 * (x / sigma) ** 2)
    return arr / arr.sum()


def convolve_extend(flux: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """astropy.convolution.convolve(flux, CustomKernel(taps), boundary='extend') restated
    (call site core/voigt_model.py:225-230): odd-length kernel, normalised by its sum
    (normalize_kernel=True), true convolution (kernel flipped) with edge replication."""
    taps = np.asarray(taps, dtype=np.float64)
    if taps.size % 2 == 0:
        raise ValueError("Kernel size must be odd in all axes.")
    k = taps / taps.sum()
    half = taps.size // 2
    padded = np.pad(np.asarray(flux, dtype=np.float64), half, mode="edge")
    return np.convolve(padded, k, mode="valid")


def lower(config, FWHM="6.5", voigt_method="wofz", custom_taps=None) -> LoweredModel:
    """VoigtModel.__init__/_cache_atomic_parameters/_setup_fast_mapping/_setup_kernel
    (core/voigt_model.py:371-464).  ``custom_taps`` stands in for the linetools COS table."""
    lam, gam, fos, zf = [], [], [], []
    for system in config.systems:
        for group in system.ion_groups:
            for wavelength in group.transitions:
                for _ in range(group.components):
                    info = rb_setline(wavelength, "closest")
                    lam.append(info["wave"][0])
                    gam.append(info["gamma"][0])
                    fos.append(info["fval"][0])
                    zf.append(1.0 + system.redshift)
    total = sum(g.components for s in config.systems for g in s.ion_groups)
    idx, base = [], 0
    for system in config.systems:
        for group in system.ion_groups:
            for _w in group.transitions:
                for c in range(group.components):
                    idx.append(base + c)
            base += group.components
    idx = np.array(idx)
    if custom_taps is not None:
        taps, kind = np.asarray(custom_taps, dtype=np.float64), "custom"
    elif FWHM is None:
        taps, kind = None, "none"
    else:
        taps, kind = gaussian_kernel_taps(FWHM), "gaussian"
    return LoweredModel(
        atomic_lambda0=np.array(lam), atomic_gamma=np.array(gam), atomic_f=np.array(fos),
        z_factors=np.array(zf), N_indices=idx, b_indices=idx + total, v_indices=idx + 2 * total,
        kernel_taps=taps, kernel_kind=kind, n_lines=len(lam), total_components=total,
        voigt_method=voigt_method)


# --------------------------------------------------------------------------- forward model
def H_tepper_garcia(x, a):
    """core/voigt_approx.py:35-86 (the reference's own algebra, kept as is)."""
    x2 = x * x
    G = np.exp(-x2)
    sqrt_pi = np.sqrt(np.pi)
    eps = np.maximum(1e-2, 100.0 * np.abs(a) / sqrt_pi)
    safe = np.maximum(x2, eps)
    numer = G * (4.0 * safe ** 2 + 7.0 * safe + 4.0) - 1.5
    denom = safe * (safe + 1.0) ** 2
    H_tg = G - (a / sqrt_pi) * numer / denom
    H_core = G * (1.0 - 2.0 * a / sqrt_pi)
    return np.where(x2 < eps, H_core, H_tg)


def voigt_tau(lambda0, gamma, f, N_linear, b_values, wave_rest, voigt_method="wofz"):
    """_vectorized_voigt_tau, core/voigt_model.py:100-159."""
    c_freq = 2.99792458e18
    atomic_constant = 4.48898479507e3
    lambda0_bc = lambda0[:, np.newaxis]
    gamma_bc = gamma[:, np.newaxis]
    f_bc = f[:, np.newaxis]
    N_bc = N_linear[:, np.newaxis]
    b_bc = b_values[:, np.newaxis]
    b_f = b_bc / lambda0_bc * 1e13
    freq0 = c_freq / lambda0_bc
    freq = c_freq / wave_rest
    constant = atomic_constant / (freq0 * b_bc)
    a = gamma_bc / (4 * np.pi * b_f)
    x = (freq - freq0) / b_f
    if voigt_method == "fast":
        H = H_tepper_garcia(x, a)
    else:
        H = wofz(x + 1j * a).real
    return N_bc * f_bc * constant * H


def model_flux(m: LoweredModel, theta, wave, convolve=True, return_tau=False):
    """_evaluate_compiled_model, core/voigt_model.py:162-230."""
    theta = np.asarray(theta)
    wave = np.asarray(wave)
    N_linear = 10 ** theta[m.N_indices]
    b_values = theta[m.b_indices]
    v_values = theta[m.v_indices]
    c = 299792.458
    z_total = m.z_factors * (1 + v_values / c) - 1
    wave_rest = wave[np.newaxis, :] / (1 + z_total[:, np.newaxis])
    tau_all = voigt_tau(m.atomic_lambda0, m.atomic_gamma, m.atomic_f, N_linear, b_values, wave_rest,
                        voigt_method=m.voigt_method)
    tau_total = np.sum(tau_all, axis=0)
    if return_tau:
        return tau_total
    flux = np.exp(-tau_total)
    if convolve and m.kernel_taps is not None:
        if m.kernel_kind == "gaussian":
            flux = ndimage.convolve1d(flux, m.kernel_taps, mode="nearest")
        else:
            flux = convolve_extend(flux, m.kernel_taps)
    return flux


def model_flux_piecewise_lsf(models, starts, theta, wave):
    """Wavelength-dependent LSF (an EXTENSION -- the reference applies one kernel per instrument,
    core/voigt_model.py:444-464 -- SURVEY 8(f) rank 3), by its definition: output pixel p is the unconvolved model
    convolved with the kernel of the block p lies in (block b = pixels with starts[b] <= wave < starts[b + 1], the
    first block from the first pixel), edges of the spectrum replicated.  ``models``: one LoweredModel per block (same
    lines, its own kernel)."""
    wave = np.asarray(wave)
    raw = model_flux(models[0], theta, wave, convolve=False)
    first = [0] + [int(np.searchsorted(wave, w0, side="left")) for w0 in list(starts)[1:]]
    edges = first + [wave.size]
    out = np.empty_like(raw)
    for b, m in enumerate(models):
        if m.kernel_taps is None:
            conv = raw
        elif m.kernel_kind == "gaussian":
            conv = ndimage.convolve1d(raw, m.kernel_taps, mode="nearest")
        else:
            conv = convolve_extend(raw, m.kernel_taps)
        out[edges[b]:edges[b + 1]] = conv[edges[b]:edges[b + 1]]
    return out


# --------------------------------------------------------------------------- likelihood
def compile_instruments(instruments: Dict[str, dict]) -> Dict[str, dict]:
    """vfit._compile_models, vfit_mcmc.py:234-259 -- weights inherit the dtype of ``error``."""
    out = {}
    for name, d in instruments.items():
        err = np.asarray(d["error"])
        with np.errstate(divide="ignore", invalid="ignore"):
            out[name] = {
                "model": d["model"],
                "wave": np.asarray(d["wave"]),
                "flux": np.asarray(d["flux"]),
                "error": err,
                "inv_sigma2": 1.0 / (err ** 2),
                "log_inv_sigma2": np.log(1.0 / (err ** 2)),
            }
    return out


def lnprior(theta, lb, ub):
    """vfit.lnprior, vfit_mcmc.py:291-295."""
    if np.any(theta < lb) or np.any(theta > ub):
        return -np.inf
    return 0.0


def lnlike(compiled: Dict[str, dict], theta):
    """vfit.lnlike, vfit_mcmc.py:297-319."""
    try:
        total = 0.0
        for _name, d in compiled.items():
            model_dat = model_flux(d["model"], theta, d["wave"])
            total += -0.5 * np.sum((d["flux"] - model_dat) ** 2 * d["inv_sigma2"] - d["log_inv_sigma2"])
        return total
    except Exception:
        return -np.inf


def lnprob(compiled: Dict[str, dict], theta, lb, ub):
    """vfit.lnprob, vfit_mcmc.py:348-353."""
    theta = np.asarray(theta)
    lp = lnprior(theta, lb, ub)
    if not np.isfinite(lp):
        return -np.inf
    with np.errstate(invalid="ignore", over="ignore"):
        return lp + lnlike(compiled, theta)


def lnprob_batch(compiled, thetas, lb, ub):
    return np.array([lnprob(compiled, t, lb, ub) for t in np.atleast_2d(thetas)])


# --------------------------------------------------------------------------- CPU baseline
_POOL_STATE = {}


def _pool_eval(theta):
    s = _POOL_STATE
    return lnprob(s["compiled"], theta, s["lb"], s["ub"])


def make_pool(compiled, lb, ub, processes=None):
    """``use_pool=True`` equivalent (vfit_mcmc.py:41-45, 413): a fork pool that lives for the whole run.
    The likelihood state is inherited by fork instead of being pickled per task, which is *kinder* to the
    CPU arm than emcee's pickling of the bound method per chunk."""
    import multiprocessing as mp
    _POOL_STATE.update(compiled=compiled, lb=np.asarray(lb), ub=np.asarray(ub))
    return mp.get_context("fork").Pool(processes)


def lnprob_pool(compiled, thetas, lb, ub, processes=None, pool=None):
    """Map lnprob over the rows of ``thetas`` with a fork pool (what emcee does per half-step)."""
    own = pool is None
    if own:
        pool = make_pool(compiled, lb, ub, processes)
    try:
        out = pool.map(_pool_eval, list(np.atleast_2d(thetas)))
    finally:
        if own:
            pool.close()
            pool.join()
    return np.array(out)


# --------------------------------------------------------------------------- synthetic LSF
def cos_like_lsf(K=321, seed=7):
    """Synthetic COS-like line-spread function: narrow core + broad asymmetric wings.
    (The real linetools COS tables are unavailable offline)"""
    half = K // 2
    x = np.arange(-half, half + 1, dtype=np.float64)
    core = np.exp(-0.5 * (x / 2.8) ** 2)
    wing_l = 0.055 * np.exp(-np.abs(x) / 22.0) * (x < 0)
    wing_r = 0.035 * np.exp(-np.abs(x) / 31.0) * (x >= 0)
    k = core + wing_l + wing_r
    return k / k.sum()
