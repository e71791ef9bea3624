"""Pins the CPU oracle (oracle/voigt_oracle.py) against fixtures produced by the REAL reference
code (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from golden_util import CASES, Golden
from oracle import voigt_oracle as vo


@pytest.mark.parametrize("case", CASES)
def test_lowering_matches_reference(case):
    g = Golden(case)
    models = g.oracle_models()
    for n in g.instruments:
        m = models[n]
        assert np.array_equal(m.atomic_lambda0, g.inst(n, "lambda0"))
        assert m.atomic_gamma.dtype == np.float32 and m.atomic_f.dtype == np.float32
        assert np.array_equal(m.atomic_gamma, g.inst(n, "gamma"))
        assert np.array_equal(m.atomic_f, g.inst(n, "f"))
        assert np.array_equal(m.z_factors, g.inst(n, "zfac"))
        assert np.array_equal(m.N_indices, g.inst(n, "N_indices"))
        taps = g.inst(n, "taps")
        if taps.size:
            assert np.array_equal(m.kernel_taps, taps)
        else:
            assert m.kernel_taps is None


@pytest.mark.parametrize("case", CASES)
def test_flux_matches_reference(case):
    g = Golden(case)
    models = g.oracle_models()
    for n in g.instruments:
        ref = g.inst(n, "ref_flux")
        for k, row in enumerate(g.flux_rows):
            got = vo.model_flux(models[n], g.thetas[row], g.inst(n, "wave"))
            assert np.max(np.abs(got - ref[k])) <= 1e-15
        # VoigtModel.evaluate() drops voigt_method (core/voigt_model.py:537-548): always wofz
        import dataclasses
        m_eval = dataclasses.replace(models[n], voigt_method="wofz")
        unc = vo.model_flux(m_eval, g.thetas[g.flux_rows[0]], g.inst(n, "wave"), convolve=False)
        assert np.max(np.abs(unc - g.inst(n, "ref_flux_unconvolved")[0])) <= 1e-15


@pytest.mark.parametrize("case", CASES)
def test_lnprob_matches_reference(case):
    g = Golden(case)
    comp = g.oracle_compiled()
    got = vo.lnprob_batch(comp, g.thetas, g.lb, g.ub)
    ref = g.ref_lnprob
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(np.isneginf(got), np.isneginf(ref))
    fin = np.isfinite(ref)
    assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= 1e-14


def test_known_answer_test_script():
    """SURVEY.md section 8c known answers for examples/test_script.py + cos_data.npz."""
    g = Golden("test_script")
    assert g.ref_lnprob[0] == 205.55188977038017
    assert g.ref_lnprob[-1] == -np.inf
    assert g.inst("COS", "ref_flux")[0].min() == 0.007271589891281632
    assert g.inst("COS", "error").dtype == np.float32
    comp = g.oracle_compiled()
    assert comp["COS"]["inv_sigma2"].dtype == np.float32
    assert np.array_equal(comp["COS"]["inv_sigma2"], g.inst("COS", "inv_sigma2"))
    assert np.array_equal(comp["COS"]["log_inv_sigma2"], g.inst("COS", "log_inv_sigma2"))
    assert g.inst("COS", "taps").size == 9
    assert list(g.inst("COS", "N_indices")) == [0, 0, 1]


def test_c1_atomic_constants_are_float32():
    g = Golden("C1")
    f = g.inst("COS", "f")
    assert float(f[0]) == 0.6122999787330627 and float(f[2]) == 0.305400013923645
    assert list(g.inst("COS", "N_indices")) == [0, 1, 0, 1]
    assert g.inst("COS", "taps").size == 23


def test_wofz_lattice_scipy_vs_mpmath():
    import os
    from golden_util import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "wofz_lattice.npz"))
    from scipy.special import wofz
    now = wofz(z["x"] + 1j * z["a"]).real
    assert np.max(np.abs(now - z["mpmath"]) / z["mpmath"]) < 5e-14


def test_convolve_extend_equals_ndimage_nearest():
    from scipy import ndimage
    rng = np.random.default_rng(3)
    f = rng.random(300)
    k = vo.gaussian_kernel_taps("6.5")
    a = ndimage.convolve1d(f, k, mode="nearest")
    b = vo.convolve_extend(f, k)
    assert np.max(np.abs(a - b)) < 1e-15
    ka = vo.cos_like_lsf(321)         # asymmetric: true convolution (flipped kernel)
    a = ndimage.convolve1d(f, ka, mode="nearest")
    b = vo.convolve_extend(f, ka)
    assert np.max(np.abs(a - b)) < 5e-15     # 321-term sums, different association


@pytest.mark.reference
def test_refshim_runs_reference_live():
    """Build container only: the live reference agrees with the committed fixture."""
    from oracle import refshim
    FitConfiguration, VoigtModel, mc, vm = refshim.import_reference()
    g = Golden("C1")
    cfg = FitConfiguration()
    for (z, ion, trans, comps) in g.meta["systems"]:
        cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
    m = VoigtModel(cfg, FWHM="6.5").compile()
    row = g.flux_rows[0]
    assert np.array_equal(m.model_flux(g.thetas[row], g.inst("COS", "wave")), g.inst("COS", "ref_flux")[0])


@pytest.mark.reference
@pytest.mark.parametrize("seed", range(24))
def test_oracle_matches_live_reference_on_fuzz_problems(seed):
    """Build container only: on the SAME seeded random problems tests/test_gpu_fuzz.py runs against the oracle on the
    GPU box, the oracle agrees with the unmodified reference (FitConfiguration -> VoigtModel -> vfit.lnprob through
    oracle/refshim.py): lowered tables identical, model flux to 1e-15, lnprob to 1e-13, same -inf rows."""
    import contextlib
    import io
    from fuzz_util import draw_problem
    from oracle import refshim
    FitConfiguration, VoigtModel, mc, vm = refshim.import_reference()
    systems, wave, fwhm, taps, theta, thetas, lb, ub, rng = draw_problem(1000 + seed)
    cfg, ocfg = FitConfiguration(), vo.OracleConfig()
    for (z, ion, trans, comps) in systems:
        cfg.add_system(z=z, ion=ion, transitions=list(trans), components=comps)
        ocfg.add_system(z, ion, trans, comps)
    ref_model = VoigtModel(cfg, FWHM=fwhm)
    if taps is not None:
        from astropy.convolution import CustomKernel          # the shim's; injected as _setup_kernel would (:458-460)
        ref_model.kernel = CustomKernel(taps)
    om = vo.lower(ocfg, FWHM=fwhm, custom_taps=taps)
    rc = ref_model.compile()
    d = rc.data
    assert np.array_equal(d.atomic_lambda0, om.atomic_lambda0) and np.array_equal(d.z_factors, om.z_factors)
    assert np.array_equal(d.atomic_f, om.atomic_f) and np.array_equal(d.atomic_gamma, om.atomic_gamma)
    assert np.array_equal(d.N_indices, om.N_indices) and np.array_equal(d.v_indices, om.v_indices)
    truth = vo.model_flux(om, theta, wave)
    flux = truth + 0.03 * rng.standard_normal(wave.size)
    error = np.full(wave.size, 0.03)
    with contextlib.redirect_stdout(io.StringIO()):
        fitter = mc.vfit({"S": dict(model=ref_model, wave=wave, flux=flux, error=error)}, theta, lb, ub)
    comp = vo.compile_instruments({"S": dict(model=om, wave=wave, flux=flux, error=error)})
    with np.errstate(all="ignore"):
        ref = np.array([fitter.lnprob(t) for t in thetas])
    got = vo.lnprob_batch(comp, thetas, lb, ub)
    assert np.array_equal(np.isneginf(got), np.isneginf(ref)) and np.isneginf(ref).sum() >= 1
    fin = np.isfinite(ref)
    assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= 1e-13
    for t in thetas[:2]:
        assert np.max(np.abs(vo.model_flux(om, t, wave) - rc.model_flux(t, wave))) <= 1e-15
