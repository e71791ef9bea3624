"""rbvfit_b200 -- B200-native (sm_100a) implementation of rbvfit's likelihood hot path.

Drop-in behind the reference's  FitConfiguration -> VoigtModel(config, FWHM) -> vfit(...) -> runmcmc()
API: ``GpuVoigtModel`` duck-types ``VoigtModel``; ``vfit`` / ``GpuLikelihood`` evaluate whole walker
ensembles per call on the GPU through the C ABI declared in include/rbvfit_b200.h.
Importing this package does not need a GPU; evaluating anything does (no CPU fallback).
"""
from .config import FitConfiguration
from .lines import rb_setline

__all__ = ["FitConfiguration", "rb_setline", "GpuVoigtModel", "GpuCompiledVoigtModel", "GpuLikelihood",
           "vfit", "set_bounds", "EnsembleSampler"]
__version__ = "0.1.0"


def __getattr__(name):      # lazy: keeps `import rbvfit_b200` free of torch / CUDA
    if name in ("GpuVoigtModel", "GpuCompiledVoigtModel", "CompiledModelData"):
        from . import model
        return getattr(model, name)
    if name == "GpuLikelihood":
        from .likelihood import GpuLikelihood
        return GpuLikelihood
    if name in ("vfit", "set_bounds"):
        from . import vfit_mcmc
        return getattr(vfit_mcmc, name)
    if name == "EnsembleSampler":
        from .sampler import EnsembleSampler
        return EnsembleSampler
    raise AttributeError(name)
