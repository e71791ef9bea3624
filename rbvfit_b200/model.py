"""GPU drop-in for rbvfit's ``VoigtModel`` / ``CompiledVoigtModel``
(reference: src/rbvfit/core/voigt_model.py:265-558).

    config = FitConfiguration(); config.add_system(z=0.348, ion='MgII', transitions=[2796.3, 2803.5], components=2)
    model = GpuVoigtModel(config, FWHM='6.5')           # same arguments as VoigtModel
    flux  = model.compile().model_flux(theta, wave)     # theta: (ndim,) -> (P,)   or   (W, ndim) -> (W, P)

Same names, argument meaning and error behaviour as the reference; the evaluation runs on the B200 through
the C ABI (include/rbvfit_b200.h).  There is no CPU path: without a CUDA device every evaluation raises.
"""
from __future__ import annotations

import copy
import hashlib
from dataclasses import dataclass
from typing import Any, Dict, Optional, Union

import numpy as np

from . import lines as _lines
from . import lsf as _lsf
from .engine import Engine


@dataclass
class CompiledModelData:
    """Same fields as the reference container (voigt_model.py:265-280); ``kernel`` holds the LSF taps
    (``None`` = no convolution) and ``kernel_normalize`` whether astropy's convolve would normalise them."""
    atomic_lambda0: np.ndarray
    atomic_gamma: np.ndarray
    atomic_f: np.ndarray
    z_factors: np.ndarray
    N_indices: np.ndarray
    b_indices: np.ndarray
    v_indices: np.ndarray
    kernel: Optional[np.ndarray]
    n_lines: int
    total_components: int
    voigt_method: str = "wofz"
    kernel_normalize: bool = False


class _FluxEvaluator:
    """Caches one flux-only engine per wavelength grid (model_flux takes ``wave`` on every call)."""

    def __init__(self, data: CompiledModelData, device=None, max_grids: int = 4):
        self.data = data
        self.device = device
        self.max_grids = max_grids
        self._engines: Dict[Any, Engine] = {}

    def _engine_for(self, wave: np.ndarray, data: CompiledModelData) -> Engine:
        wave = np.ascontiguousarray(wave, dtype=np.float64)
        key = (wave.size, hashlib.blake2b(wave.tobytes(), digest_size=16).digest(), id(data))
        eng = self._engines.get(key)
        if eng is None:
            if len(self._engines) >= self.max_grids:
                self._engines.pop(next(iter(self._engines))).close()
            eng = Engine(self.device)
            eng.add_instrument(data, wave, taps=data.kernel, normalize_taps=data.kernel_normalize)
            self._engines[key] = eng
        return eng

    def flux(self, theta, wave, convolve=True, data: Optional[CompiledModelData] = None):
        data = data or self.data
        theta = np.asarray(theta, dtype=np.float64)
        single = theta.ndim == 1
        th = np.atleast_2d(theta)
        if th.shape[1] < 3 * data.total_components:
            raise IndexError(f"theta has {th.shape[1]} parameters, the model needs {3 * data.total_components}")
        out = self._engine_for(wave, data).model_flux(0, th, convolve=convolve)   # convolve=0: LSF skipped in-kernel
        return out[0] if single else out

    def close(self):
        for e in self._engines.values():
            e.close()
        self._engines = {}


class GpuCompiledVoigtModel:
    """Picklable compiled model (``CompiledVoigtModel``, voigt_model.py:283-323): only the lowered arrays
    are pickled, device state is re-created lazily in the receiving process."""

    def __init__(self, data_container: CompiledModelData, device=None):
        self.data = data_container
        self._device = device
        self._eval: Optional[_FluxEvaluator] = None

    def _evaluator(self) -> _FluxEvaluator:
        if self._eval is None:
            self._eval = _FluxEvaluator(self.data, self._device)
        return self._eval

    def model_flux(self, theta: np.ndarray, wavelength: np.ndarray) -> np.ndarray:
        return self._evaluator().flux(theta, wavelength, convolve=True)

    def __call__(self, theta, wavelength):
        return self.model_flux(theta, wavelength)

    def __getstate__(self):
        return {"data": self.data, "_device": self._device}

    def __setstate__(self, state):
        self.data = state["data"]
        self._device = state.get("_device")
        self._eval = None


class GpuVoigtModel:
    """Same constructor and attributes as ``VoigtModel`` (voigt_model.py:334-384)."""

    def __init__(self, config, FWHM: Union[str, float, None] = "6.5", grating: str = "G130M",
                 life_position: str = "1", cen_wave: str = "1300A", voigt_method: str = "wofz",
                 device: Optional[int] = None, lsf_taps=None):
        _valid = ("wofz", "fast")
        if voigt_method not in _valid:
            raise ValueError(f"voigt_method must be one of {_valid}, got '{voigt_method}'")
        self.voigt_method = voigt_method
        self.config = config
        self.config.validate()
        self.device = device
        ip = getattr(config, "instrumental_params", {}) or {}
        self.FWHM = ip.get("FWHM", FWHM)                  # config wins over the argument (:371)
        self.grating = ip.get("grating", grating)
        self.life_position = ip.get("life_position", life_position)
        self.cen_wave = ip.get("cen_wave", cen_wave)
        self._lsf_taps = None if lsf_taps is None else np.asarray(lsf_taps, dtype=np.float64)
        self._setup_kernel()
        self._cache_atomic_parameters()
        self._setup_fast_mapping()
        self._compiled = False
        self._eval: Optional[_FluxEvaluator] = None

    # ------------------------------------------------------------------ set-up (host, not hot)
    def _setup_kernel(self):
        """voigt_model.py:444-464.  ``kernel`` = taps as applied; custom tables go through astropy's
        convolve, which normalises the kernel, Gaussian ones through ndimage.convolve1d verbatim."""
        self.kernel_normalize = False
        if self._lsf_taps is not None:                    # explicit LSF table (stands in for linetools)
            if self._lsf_taps.size % 2 == 0:
                raise ValueError("Kernel size must be odd in all axes.")
            self.kernel = self._lsf_taps.copy()
            self.kernel_normalize = True
        elif self.FWHM is None:
            self.kernel = None
        elif isinstance(self.FWHM, str) and self.FWHM == "COS":
            self.kernel = _lsf.cos_taps(self.grating, self.life_position, self.cen_wave)
            self.kernel_normalize = True
        else:
            self.kernel = _lsf.gaussian_taps(float(self.FWHM))

    def _cache_atomic_parameters(self):
        """voigt_model.py:386-412: system -> ion group -> transition -> component; f, gamma stay float32."""
        lam, gam, fos, zf = [], [], [], []
        for system in self.config.systems:
            for group in system.ion_groups:
                for wavelength in group.transitions:
                    for _ in range(group.components):
                        info = _lines.rb_setline(wavelength, "closest")
                        lam.append(info["wave"][0])
                        gam.append(info["gamma"][0])
                        fos.append(info["fval"][0])
                        zf.append(1.0 + system.redshift)
        self.atomic_lambda0 = np.array(lam)
        self.atomic_gamma = np.array(gam)
        self.atomic_f = np.array(fos)
        self.z_factors = np.array(zf)
        self.n_lines = len(lam)

    def _setup_fast_mapping(self):
        """voigt_model.py:414-442: every transition of an ion group points at that group's component slots."""
        self.total_components = sum(g.components for s in self.config.systems for g in s.ion_groups)
        idx, base = [], 0
        for system in self.config.systems:
            for group in system.ion_groups:
                for _w in group.transitions:
                    for c in range(group.components):
                        idx.append(base + c)
                base += group.components
        self.N_indices = np.array(idx)
        self.b_indices = self.N_indices + self.total_components
        self.v_indices = self.N_indices + 2 * self.total_components

    def _container(self, kernel, voigt_method) -> CompiledModelData:
        return CompiledModelData(
            atomic_lambda0=self.atomic_lambda0.copy(), atomic_gamma=self.atomic_gamma.copy(),
            atomic_f=self.atomic_f.copy(), z_factors=self.z_factors.copy(), N_indices=self.N_indices.copy(),
            b_indices=self.b_indices.copy(), v_indices=self.v_indices.copy(), kernel=copy.deepcopy(kernel),
            n_lines=self.n_lines, total_components=self.total_components, voigt_method=voigt_method,
            kernel_normalize=self.kernel_normalize)

    # ------------------------------------------------------------------ public API
    def compile(self, verbose: bool = False) -> GpuCompiledVoigtModel:
        """voigt_model.py:466-507."""
        if verbose:
            print(f"Compiling GpuVoigtModel: {3 * self.total_components} parameters, {self.n_lines} lines")
            print(f"FWHM: {self.FWHM}")
        compiled = GpuCompiledVoigtModel(self._container(self.kernel, self.voigt_method), self.device)
        self._compiled = True
        if verbose:
            print("GpuVoigtModel compiled successfully")
        return compiled

    def evaluate(self, theta, wavelength, return_components: bool = False, return_unconvolved: bool = False,
                 validate_theta: bool = False):
        """voigt_model.py:509-558.  As in the reference this path always uses the exact (wofz) profile
        and ``components`` are the UNCONVOLVED per-line fluxes exp(-tau_i) (:232-238)."""
        theta = np.asarray(theta, dtype=np.float64)
        if validate_theta and theta.shape[-1] != 3 * self.total_components:
            raise ValueError(f"theta must have {3 * self.total_components} parameters")
        if self._eval is None:
            self._eval = _FluxEvaluator(self._container(self.kernel, "wofz"), self.device,
                                        max_grids=2 * (self.n_lines + 2))
            self._line_data: Dict[int, CompiledModelData] = {}
        flux = self._eval.flux(theta, wavelength, convolve=not return_unconvolved and self.kernel is not None)
        if not return_components:
            return flux
        N_linear = 10 ** theta[self.N_indices]
        z_total = self.z_factors * (1 + theta[self.v_indices] / 299792.458) - 1
        comps, info = [], []
        for i in range(self.n_lines):
            d = self._line_data.get(i)
            if d is None:
                d = CompiledModelData(
                    atomic_lambda0=self.atomic_lambda0[i:i + 1].copy(), atomic_gamma=self.atomic_gamma[i:i + 1].copy(),
                    atomic_f=self.atomic_f[i:i + 1].copy(), z_factors=self.z_factors[i:i + 1].copy(),
                    N_indices=self.N_indices[i:i + 1].copy(), b_indices=self.b_indices[i:i + 1].copy(),
                    v_indices=self.v_indices[i:i + 1].copy(), kernel=None, n_lines=1,
                    total_components=self.total_components, voigt_method="wofz")
                self._line_data[i] = d
            comps.append(self._eval.flux(theta, wavelength, convolve=False, data=d))
            info.append({"line_index": i, "lambda0": float(self.atomic_lambda0[i]),
                         "gamma": float(self.atomic_gamma[i]), "f_value": float(self.atomic_f[i]),
                         "z_total": float(z_total[i]), "N_value": float(N_linear[i]),
                         "b_value": float(theta[self.b_indices][i]), "v_value": float(theta[self.v_indices][i])})
        return {"flux": flux, "components": comps, "component_info": info}

    @property
    def is_compiled(self) -> bool:
        return self._compiled

    def get_info(self) -> str:
        lsf = "none" if self.kernel is None else f"{len(self.kernel)} taps (FWHM={self.FWHM})"
        return (f"GpuVoigtModel: {self.n_lines} lines, {self.total_components} components, "
                f"{3 * self.total_components} parameters, LSF {lsf}, voigt_method={self.voigt_method}")

    def print_info(self) -> None:
        print(self.get_info())
