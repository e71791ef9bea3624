"""TEST INFRASTRUCTURE ONLY -- run the UNMODIFIED reference in the build container.

``/root/reference/src/rbvfit`` imports astropy, emcee, corner and matplotlib at module top
(core/voigt_model.py:13, rb_setline.py:3, vfit_mcmc.py:21-25); none is installed here and
there is no network.  ``install()`` registers minimal stand-ins in ``sys.modules`` so that
``rbvfit.core.voigt_model`` and ``rbvfit.vfit_mcmc`` import and the reference's own numeric
code (``_vectorized_voigt_tau``, ``_evaluate_compiled_model``, ``vfit.lnprob``) executes on
numpy + scipy.  The stand-ins restate astropy >= 5.3 behaviour:

  * ``Gaussian1DKernel(stddev)``: odd size >= ceil(8 stddev), centre-sampled, normalised.
  * ``CustomKernel(array)``: array kept verbatim.
  * ``convolve(array, kernel, boundary='extend')``: kernel / kernel.sum(), true convolution,
    edge replication.
  * ``astropy.io.ascii.read(file)``: whitespace table with columns col1..colN.

The reference is taken from /root/reference/src (build container) or, where that does not exist (the GPU box),
from ``oracle/_ref`` -- the pip install of the unmodified reference staged by ``oracle/build_ref.py``.
Users: ``oracle/make_golden.py``, the ``reference``-marked tests, ``bench.py --impl reference``.
"""
from __future__ import annotations

import math
import os
import sys
import types

import numpy as np

STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_SRC = ("/root/reference/src" if os.path.isdir("/root/reference/src/rbvfit")
                 and os.environ.get("RBVFIT_B200_REF") != "staged" else STAGED)


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "rbvfit", "vfit_mcmc.py"))


def kind() -> str:
    """Where the reference comes from: 'source tree' (/root/reference) or 'staged' (oracle/_ref)."""
    return "staged" if REFERENCE_SRC == STAGED else "source tree"


class _Kernel1D:
    def __init__(self, array):
        self._array = np.asarray(array, dtype=np.float64)

    @property
    def array(self):
        return self._array

    @property
    def shape(self):
        return self._array.shape


class Gaussian1DKernel(_Kernel1D):
    def __init__(self, stddev, **kwargs):
        size = int(math.ceil(8 * stddev))
        if size % 2 == 0:
            size += 1
        half = size // 2
        x = np.arange(-half, half + 1, dtype=np.float64)
        arr = (1.0 / (np.sqrt(2 * np.pi) * stddev)) * np.exp(-0.5 * (x / stddev) ** 2)
        super().__init__(arr / arr.sum())


class CustomKernel(_Kernel1D):
    pass


def convolve(array, kernel, boundary="fill", **kwargs):
    if boundary != "extend":
        raise NotImplementedError("shim implements boundary='extend' only")
    k = kernel.array if hasattr(kernel, "array") else np.asarray(kernel, dtype=np.float64)
    if k.size % 2 == 0:
        raise ValueError("Kernel size must be odd in all axes.")
    k = k / k.sum()
    half = k.size // 2
    padded = np.pad(np.asarray(array, dtype=np.float64), half, mode="edge")
    return np.convolve(padded, k, mode="valid")


class _Row(dict):
    pass


class _Table(list):
    pass


def _ascii_read(filename, **kwargs):
    table = _Table()
    with open(filename) as fh:
        for line in fh:
            parts = line.split()
            if not parts or parts[0].startswith("#"):
                continue
            row = _Row()
            for i, p in enumerate(parts):
                try:
                    row[f"col{i + 1}"] = float(p)
                except ValueError:
                    row[f"col{i + 1}"] = p
            table.append(row)
    return table


def install():
    """Register the stand-ins and put the reference on sys.path.  Idempotent."""
    if not available():
        raise RuntimeError("neither /root/reference nor oracle/_ref (python -m oracle.build_ref) is present")
    if "astropy" not in sys.modules:
        astropy = types.ModuleType("astropy")
        conv = types.ModuleType("astropy.convolution")
        conv.convolve = convolve
        conv.Gaussian1DKernel = Gaussian1DKernel
        conv.CustomKernel = CustomKernel
        io = types.ModuleType("astropy.io")
        ascii_mod = types.ModuleType("astropy.io.ascii")
        ascii_mod.read = _ascii_read
        io.ascii = ascii_mod
        astropy.convolution = conv
        astropy.io = io
        sys.modules.update({"astropy": astropy, "astropy.convolution": conv, "astropy.io": io,
                            "astropy.io.ascii": ascii_mod})
    for name in ("emcee", "corner", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)


def import_reference():
    """Returns (FitConfiguration, VoigtModel, vfit_mcmc module, voigt_model module)."""
    install()
    from rbvfit.core.fit_configuration import FitConfiguration
    from rbvfit.core import voigt_model as vm
    import rbvfit.vfit_mcmc as mc
    return FitConfiguration, vm.VoigtModel, mc, vm
