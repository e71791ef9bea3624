"""GPU parity on randomly drawn problems (fixed seeds): random systems / ions / component counts, redshifts,
wavelength windows and pixel counts, LSFs, column densities up to the damped regime, Doppler widths from 3 to 80 km/s
-- lnprob and model flux against the CPU oracle at the north-star tolerances."""
import numpy as np
import pytest

from fuzz_util import draw_problem as _draw

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(24))
def test_random_problem_matches_oracle(seed):
    from oracle import voigt_oracle as vo
    from rbvfit_b200 import FitConfiguration
    from rbvfit_b200.likelihood import GpuLikelihood
    from rbvfit_b200.model import GpuVoigtModel
    systems, wave, fwhm, taps, theta, thetas, lb, ub, rng = _draw(1000 + seed)
    cfg, ocfg = FitConfiguration(), vo.OracleConfig()
    for (z, ion, trans, comps) in systems:
        cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
        ocfg.add_system(z, ion, trans, comps)
    model = GpuVoigtModel(cfg, FWHM=fwhm, lsf_taps=taps)
    om = vo.lower(ocfg, FWHM=fwhm, custom_taps=taps)
    truth = vo.model_flux(om, theta, wave)
    flux = truth + 0.03 * rng.standard_normal(wave.size)
    error = np.full(wave.size, 0.03)
    like = GpuLikelihood({"S": dict(model=model, wave=wave, flux=flux, error=error)}, lb, ub)
    comp = vo.compile_instruments({"S": dict(model=om, wave=wave, flux=flux, error=error)})
    got, ref = like.lnprob(thetas), vo.lnprob_batch(comp, thetas, lb, ub)
    assert np.array_equal(np.isneginf(got), np.isneginf(ref)) and np.isneginf(ref).sum() >= 1
    fin = np.isfinite(ref)
    assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= 1e-9
    gf = model.compile().model_flux(thetas[:3], wave)
    rf = np.array([vo.model_flux(om, t, wave) for t in thetas[:3]])
    assert np.max(np.abs(gf - rf)) <= 1e-10
    like.close()
