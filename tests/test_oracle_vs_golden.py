"""Pins the CPU oracle (oracle/voigt_oracle.py) against fixtures produced by the REAL reference
code (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from golden_util import BIG_CASES, CASES, Golden, GoldenSightlines
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


@pytest.mark.parametrize("case", BIG_CASES)
def test_headline_geometry_matches_reference(case):
    """C5a / C5a_L4 (100 000 px): lowering, lnprob of the 10 fixture rows and the flux of two rows on the pixel
    subset, oracle vs the live reference's committed outputs."""
    g = Golden(case)
    n = g.instruments[0]
    m = g.oracle_models()[n]
    assert np.array_equal(m.atomic_lambda0, g.inst(n, "lambda0")) and np.array_equal(m.atomic_f, g.inst(n, "f"))
    assert np.array_equal(m.kernel_taps, g.inst(n, "taps"))
    wave, px = g.inst(n, "wave"), g.inst(n, "flux_px")
    assert wave.size == 100000 and px.size < wave.size // 4
    for k, row in enumerate(g.flux_rows):
        got = vo.model_flux(m, g.thetas[row], wave)
        assert np.max(np.abs(got[px] - g.inst(n, "ref_flux")[k])) <= 1e-15
    unc = vo.model_flux(m, g.thetas[0], wave, convolve=False)
    assert np.max(np.abs(unc[px] - g.inst(n, "ref_flux_unconvolved")[0])) <= 1e-15
    got = vo.lnprob_batch(g.oracle_compiled(), g.thetas, g.lb, g.ub)
    ref = g.ref_lnprob
    assert np.count_nonzero(np.isneginf(ref)) == 2 and np.array_equal(np.isneginf(got), np.isneginf(ref))
    fin = np.isfinite(ref)
    assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= 1e-14


def test_sightline_fixture_matches_reference():
    """C5b: every sightline's oracle likelihood against that sightline's own reference vfit."""
    g = GoldenSightlines()
    assert g.n == 8 and g.thetas.shape[1:] == (16, 6)
    for s in range(g.n):
        got = vo.lnprob_batch(g.oracle_compiled(s), g.thetas[s], g.lb, g.ub)
        ref = g.ref_lnprob[s]
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


def _ref_vs_oracle(instruments, thetas, lb, ub, flux_tol=1e-15, rtol=1e-13):
    """instruments: {name: (systems, FWHM, taps, wave, flux, error)} -> oracle vs LIVE reference (joint lnprob, flux)."""
    import contextlib
    import io
    from oracle import refshim
    FitConfiguration, VoigtModel, mc, vm = refshim.import_reference()
    ref_inst, ora_inst, pairs = {}, {}, []
    for name, (systems, fwhm, taps, wave, flux, error) in instruments.items():
        cfg, ocfg = FitConfiguration(), vo.OracleConfig()
        for (z, ion, trans, comps) in systems:
            cfg.add_system(z=z, ion=ion, transitions=list(trans), components=comps)
            ocfg.add_system(z, ion, trans, comps)
        rm = VoigtModel(cfg, FWHM=fwhm)
        if taps is not None:
            from astropy.convolution import CustomKernel
            rm.kernel = CustomKernel(taps)
        om = vo.lower(ocfg, FWHM=fwhm, custom_taps=taps)
        ref_inst[name] = dict(model=rm, wave=wave, flux=flux, error=error)
        ora_inst[name] = dict(model=om, wave=wave, flux=flux, error=error)
        pairs.append((rm.compile(), om, wave))
    thetas = np.atleast_2d(thetas)
    with contextlib.redirect_stdout(io.StringIO()):
        fitter = mc.vfit(ref_inst, np.clip(thetas[0], lb, ub), lb, ub)
    comp = vo.compile_instruments(ora_inst)
    with np.errstate(all="ignore"):
        ref = np.array([fitter.lnprob(t) for t in thetas])
    got = vo.lnprob_batch(comp, thetas, lb, ub)
    assert np.array_equal(np.isneginf(got), np.isneginf(ref)) and np.array_equal(np.isnan(got), np.isnan(ref))
    fin = np.isfinite(ref)
    if fin.any():
        assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= rtol
    for rc, om, wave in pairs:
        assert np.max(np.abs(vo.model_flux(om, thetas[0], wave) - rc.model_flux(thetas[0], wave))) <= flux_tol
    return ref


@pytest.mark.reference
def test_oracle_matches_live_reference_on_edge_cases():
    """Build container only: the edge cases tests/test_gpu_edges.py runs against the oracle on the GPU box -- tiny and
    ragged spectra (shorter than the LSF, one pixel), an LSF wider than the spectrum, descending and shuffled
    wavelength grids, 96 lines over 32 components, float32 errors, a joint fit with different line lists per
    instrument -- agree between the oracle and the unmodified reference."""
    mgii = [(0.348, "MgII", [2796.3, 2803.5], 2)]
    th = np.array([14.2, 14.5, 40.0, 30.0, -25.0, 35.0])
    lb, ub = th - np.array([2, 2, 38, 28, 50, 50.0]), th + np.array([2, 2, 40, 40, 50, 50.0])
    rng = np.random.default_rng(41)
    thetas = np.clip(th + rng.standard_normal((5, 6)) * [0.05, 0.05, 1, 1, 2, 2], lb, ub)
    thetas[3, 0] = ub[0] + 1.0

    def spec(wave, dtype=np.float64):
        wave = np.asarray(wave, dtype=np.float64)
        return wave, 1.0 + 0.05 * rng.standard_normal(wave.size), np.full(wave.size, 0.05, dtype=dtype)

    for P in (1, 2, 7, 22, 23, 257):
        wave = np.linspace(3762.0, 3786.0, P) if P > 1 else np.array([3769.5])
        ref = _ref_vs_oracle({"S": (mgii, "6.5", None) + spec(wave)}, thetas, lb, ub)
        assert np.isneginf(ref[3]) and np.isfinite(ref[0])
    x = np.arange(-160, 161)
    taps = np.exp(-0.5 * (x / 25.0) ** 2) * (1 + 0.3 * (x > 0))
    for P in (40, 321, 700):                                   # asymmetric 321-tap LSF, wider than the spectrum
        _ref_vs_oracle({"S": (mgii, None, taps / taps.sum()) + spec(np.linspace(3762.0, 3786.0, P))}, thetas, lb, ub,
                       flux_tol=5e-15)
    wave = np.linspace(3755.0, 3795.0, 3000)
    _ref_vs_oracle({"S": (mgii, "6.5", None) + spec(wave[::-1].copy())}, thetas, lb, ub)
    _ref_vs_oracle({"S": (mgii, "6.5", None) + spec(rng.permutation(wave))}, thetas, lb, ub)
    _ref_vs_oracle({"S": (mgii, "6.5", None) + spec(np.linspace(3755.0, 3795.0, 2048), np.float32)}, th, lb, ub)
    systems = []
    for z in np.linspace(1.9, 2.9, 8):
        systems.append((float(z), "CIV", [1548.2, 1550.77], 2))
        systems.append((float(z), "HI", [1215.67, 1025.72, 972.54, 949.74], 2))
    C = 32
    t96 = np.concatenate([rng.uniform(12.5, 14.5, C), rng.uniform(8, 45, C), rng.uniform(-120, 120, C)])
    _ref_vs_oracle({"S": (systems, "6.5", None) + spec(np.linspace(3400.0, 6100.0, 9000))},
                   t96 + 0.01 * rng.standard_normal((2, 3 * C)), t96 - 60.0, t96 + 60.0)
    _ref_vs_oracle({"A": ([(0.348, "MgII", [2796.3], 2)], "6.5", None) + spec(np.linspace(3760.0, 3776.0, 700)),
                    "B": (mgii, "3.0", None) + spec(np.linspace(3755.0, 3795.0, 5000))}, thetas, lb, ub)
