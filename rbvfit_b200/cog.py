"""Curve of growth through the GPU forward model (reference: src/rbvfit/compute_cog.py:23-184).

The reference evaluates ``model_compiled.model_flux([N, b, 0], wave)`` once per (N, b) grid point in a Python
double loop (compute_cog.py:84-90, 177-181) and integrates ``1 - flux`` with the trapezoidal rule on a 1000-point
grid of +-5 Angstrom around the line (:49-59).  Here the whole (N, b) grid is ONE ``model_flux`` batch; the grid,
the parameter layout, the absence of an LSF and the trapezoidal rule are the reference's.  Plotting is out of scope.
"""
from __future__ import annotations

import numpy as np

from .config import FitConfiguration
from .lines import rb_setline
from .model import GpuVoigtModel

_FALLBACK_IONS = [((1200, 1230), "HI"), ((2790, 2810), "MgII"), ((1190, 1200), "SiII"), ((1240, 1260), "NV"),
                  ((1520, 1530), "CIV"), ((1030, 1040), "OVI")]          # compute_cog.py:143-158


def _trapezoid(y, x):
    """np.trapz(y, x=x) along the last axis (np.trapz is gone in numpy 2.x; same arithmetic)."""
    d = np.diff(x)
    return np.sum(d * (y[..., 1:] + y[..., :-1]) / 2.0, axis=-1)


def equivalent_widths(model_compiled, Nlist, blist, lam_rest, n_wave: int = 1000, half_width: float = 5.0):
    """W[i, j] = integral of (1 - flux) for logN = Nlist[i], b = blist[j], v = 0 (set_one_absorber :23-64)."""
    Nlist = np.asarray(Nlist, dtype=np.float64)
    blist = np.asarray(blist, dtype=np.float64)
    wave = np.linspace(lam_rest - half_width, lam_rest + half_width, n_wave)
    NN, BB = np.meshgrid(Nlist, blist, indexing="ij")
    theta = np.stack([NN.ravel(), BB.ravel(), np.zeros(NN.size)], axis=1)
    flux = model_compiled.model_flux(theta, wave)                      # one device batch for the whole grid
    return _trapezoid(1.0 - flux, wave).reshape(len(Nlist), len(blist))


class compute_cog:
    """Same constructor and attributes (``st``, ``model_compiled``, ``Nlist``, ``blist``, ``Wlist``) as the
    reference class (compute_cog.py:92-184)."""

    def __init__(self, lam_guess, Nlist, blist, device=None, verbose: bool = False):
        self.st = rb_setline(lam_guess, "closest")
        wave_val = float(np.atleast_1d(self.st["wave"])[0])
        self.st["wave"] = wave_val
        self.st["fval"] = float(np.atleast_1d(self.st["fval"])[0])
        self.st["gamma"] = float(np.atleast_1d(self.st["gamma"])[0])
        config = FitConfiguration()
        try:
            config.add_system(z=0.0, ion="auto", transitions=[wave_val], components=1)
        except Exception:
            ion = next((name for (lo, hi), name in _FALLBACK_IONS if lo < wave_val < hi), "HI")
            config.add_system(z=0.0, ion=ion, transitions=[wave_val], components=1)
        model = GpuVoigtModel(config, FWHM=None, device=device)       # no LSF for COG calculations (:160-162)
        self.model_compiled = model.compile()
        self.Nlist = np.array(Nlist)
        self.blist = np.array(blist)
        if verbose:
            print(f"Computing COG for transition: {self.st['name']} at {wave_val:.2f} A "
                  f"({len(self.Nlist)} x {len(self.blist)} grid, one device batch)")
        self.Wlist = equivalent_widths(self.model_compiled, self.Nlist, self.blist, wave_val)

    def plot_cog(self, *args, **kwargs):
        raise NotImplementedError("plotting is outside the hot path (SURVEY.md section 8f); use Wlist / Nlist / blist")
