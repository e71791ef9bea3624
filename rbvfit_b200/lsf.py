"""Line-spread-function taps exactly as the reference would hand them to the convolution
(reference: VoigtModel._setup_kernel, src/rbvfit/core/voigt_model.py:444-464).

If astropy is importable its ``Gaussian1DKernel(...).array`` is used verbatim so the taps can never drift
from what the reference applies; otherwise the documented astropy >= 5.3 construction is restated here
(odd size >= ceil(8 sigma), centre-sampled Gaussian, normalised to unit sum).
"""
from __future__ import annotations

import math

import numpy as np


def gaussian_taps_restated(fwhm_pixels) -> np.ndarray:
    """astropy >= 5.3's ``Gaussian1DKernel(stddev)`` restated: odd size >= ceil(8 sigma), centre-sampled Gaussian,
    normalised to sum 1.  Parity against astropy itself is pinned wherever astropy is importable
    (tests/test_reference_contract.py); it is not in the build image nor on the GPU pool."""
    sigma = float(fwhm_pixels) / 2.355
    size = int(math.ceil(8 * sigma))
    if size % 2 == 0:
        size += 1
    half = size // 2
    x = np.arange(-half, half + 1, dtype=np.float64)
    arr = (1.0 / (np.sqrt(2 * np.pi) * sigma)) * np.exp(-0.5 * (x / sigma) ** 2)
    return arr / arr.sum()


def gaussian_taps(fwhm_pixels) -> np.ndarray:
    """Gaussian1DKernel(stddev=FWHM/2.355).array  (voigt_model.py:462-464): astropy's own array whenever astropy is
    importable (so the taps cannot drift from the reference's), else the restated construction."""
    try:  # pragma: no cover - astropy is absent in the build image
        from astropy.convolution import Gaussian1DKernel
        return np.asarray(Gaussian1DKernel(stddev=float(fwhm_pixels) / 2.355).array, dtype=np.float64)
    except ImportError:
        return gaussian_taps_restated(fwhm_pixels)


def cos_taps(grating, life_position, cen_wave) -> np.ndarray:
    """The linetools COS table column used by the reference (voigt_model.py:449-460)."""
    try:
        from linetools.spectra.lsf import LSF
    except ImportError as exc:
        raise ImportError("COS LSF requires linetools package") from exc
    lsf = LSF(dict(name="COS", grating=grating, life_position=life_position, cen_wave=cen_wave))
    _, data = lsf.load_COS_data()
    return np.asarray(data[cen_wave].data, dtype=np.float64)


def cos_like_taps(n_taps: int = 321) -> np.ndarray:
    """Synthetic COS-like LSF (narrow core + broad asymmetric wings) for benchmarks when the linetools
    tables are unavailable; same shape family as the oracle's ``cos_like_lsf``."""
    half = n_taps // 2
    x = np.arange(-half, half + 1, dtype=np.float64)
    k = (np.exp(-0.5 * (x / 2.8) ** 2) + 0.055 * np.exp(-np.abs(x) / 22.0) * (x < 0)
         + 0.035 * np.exp(-np.abs(x) / 31.0) * (x >= 0))
    return k / k.sum()
