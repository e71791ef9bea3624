"""The reference's own results class consumes what this package produces: the UNMODIFIED
`rbvfit.core.unified_results.UnifiedResults` (imported from /root/reference behind the oracle's shim, plus an empty
h5py stand-in -- only save/load touch HDF5) is constructed from a fitter that exposes this package's sampler objects
and FitConfiguration, exactly the attributes `rbvfit_b200.vfit` carries after `runmcmc()`.  Build container only."""
import copy
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.reference


def _fitter(sampler_name):
    from rbvfit_b200 import FitConfiguration
    from rbvfit_b200.sampler import EnsembleSampler
    from rbvfit_b200.slice_sampler import EnsembleSliceSampler
    mu, sig = np.array([14.0, 30.0, 0.0]), np.array([0.1, 2.0, 3.0])

    def lnp(x):
        return -0.5 * np.sum(((np.atleast_2d(x) - mu) / sig) ** 2, axis=1)

    cls = EnsembleSampler if sampler_name == "emcee" else EnsembleSliceSampler
    s = cls(16, 3, lnp, seed=1)
    s.run_mcmc(mu + 1e-2 * np.random.default_rng(0).standard_normal((16, 3)), 300)
    cfg = FitConfiguration()
    cfg.add_system(z=0.348, ion="MgII", transitions=[2796.3, 2803.5], components=1)
    cfg = copy.deepcopy(cfg)
    cfg.instrumental_params = {"FWHM": "6.5"}
    wave = np.linspace(3760.0, 3790.0, 50)

    class Fitter:      # the attribute surface of rbvfit_b200.vfit that UnifiedResults reads
        pass
    f = Fitter()
    f.sampler, f.sampler_name = s, sampler_name
    f.best_theta = np.median(s.get_chain(discard=100, flat=True), axis=0)
    f.theta, f.lb, f.ub = mu, mu - 5 * sig, mu + 5 * sig
    f.no_of_Chain, f.no_of_steps, f.multi_instrument = 16, 300, False
    f.instrument_data = {"COS": {"wave": wave, "flux": np.ones(50), "error": np.full(50, 0.05), "model": None}}
    f.instrument_configs = {"COS": cfg}
    return f, mu, sig


@pytest.mark.parametrize("sampler_name", ["emcee", "zeus"])
def test_reference_unified_results_accepts_our_fitter(sampler_name):
    from oracle import refshim
    refshim.install()
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    from rbvfit.core import unified_results as ur
    f, mu, sig = _fitter(sampler_name)
    r = ur.UnifiedResults(f)
    assert r.chain.shape == (300, 16, 3) and r.samples.ndim == 2 and r.samples.shape[1] == 3
    assert np.array_equal(r.best_fit, f.best_theta)
    assert np.array_equal(r.bounds_lb, f.lb) and np.array_equal(r.bounds_ub, f.ub)
    assert r.n_walkers == 16 and r.n_steps == 300 and r.sampler_name == sampler_name
    assert 0.2 < r.acceptance_fraction <= 1.0
    assert r.config_metadata is not None and set(r.instrument_data) == {"COS"}
    assert np.all(np.abs(r.samples.mean(axis=0) - mu) < 0.5 * sig)
    assert r.correlation_matrix().shape == (3, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("workload", ["C1", "C3"])
def test_reference_vfit_drives_gpu_model(workload):
    """SURVEY 8(b) model plug point: the UNMODIFIED reference `vfit` (staged by oracle/build_ref.py, imported through
    the shim) takes a `GpuVoigtModel` as `instrument_data[...]['model']` -- it sees `.config` + `.compile()` and calls
    `compile(verbose=False).model_flux(theta, wave)` per instrument (vfit_mcmc.py:238-248), then does its own
    chi^2 in numpy (:297-319).  Its lnprob must agree with this package's fused-device `vfit.lnprob` and with the
    oracle on the same rows."""
    import contextlib
    import io
    from oracle import refshim, voigt_oracle as vo
    from rbvfit_b200 import FitConfiguration, lsf, workloads as wl
    from rbvfit_b200.model import GpuVoigtModel
    from rbvfit_b200.vfit_mcmc import vfit as gpu_vfit
    refshim.install()
    import rbvfit.vfit_mcmc as mc
    w = wl.get_workload(workload)
    cfg, ocfg = FitConfiguration(), vo.OracleConfig()
    for (z, ion, trans, comps) in w["systems"]:
        cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
        ocfg.add_system(z, ion, trans, comps)
    models, omodels = {}, {}
    for name, inst in w["instruments"].items():
        cos = inst.get("lsf") == "cos_like"
        models[name] = GpuVoigtModel(cfg, FWHM=inst["FWHM"], lsf_taps=lsf.cos_like_taps(321) if cos else None)
        omodels[name] = vo.lower(ocfg, FWHM=inst["FWHM"], custom_taps=vo.cos_like_lsf(321) if cos else None)
    spectra = wl.make_spectra(w, lambda n, th, wave: vo.model_flux(omodels[n], th, wave))
    inst_data = {n: dict(model=models[n], **spectra[n]) for n in models}
    with contextlib.redirect_stdout(io.StringIO()):
        ref_fitter = mc.vfit(inst_data, w["theta_true"], w["lb"], w["ub"])       # the reference's own class
    ours = gpu_vfit(inst_data, w["theta_true"], w["lb"], w["ub"])
    thetas = wl.make_ensemble(w, 24)
    with np.errstate(all="ignore"):
        ref = np.array([ref_fitter.lnprob(t) for t in thetas])
    got = ours.lnprob(thetas)
    ora = vo.lnprob_batch(vo.compile_instruments({n: dict(model=omodels[n], **spectra[n]) for n in omodels}),
                          thetas, w["lb"], w["ub"])
    assert np.isneginf(ref).any() and np.array_equal(np.isneginf(ref), np.isneginf(got))
    fin = np.isfinite(ref)
    assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= 1e-9
    assert np.max(np.abs(ora[fin] - ref[fin]) / np.abs(ref[fin])) <= 1e-9
    # the reference's optimiser entry (vfit_mcmc.py:355-360) runs on the GPU-backed model as well
    assert np.isfinite(ref_fitter.lnprob(w["theta_true"]))
