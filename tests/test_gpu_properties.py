"""GPU: size-independent properties at BASELINE.json's full sizes (C5a: 100 000 px, L = 33; C3: two instruments;
C4w: 321-tap LSF), where the CPU oracle would take minutes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _problem(workload, nwalkers, voigt_method="wofz"):
    from rbvfit_b200 import FitConfiguration, lsf, workloads as wl
    from rbvfit_b200.likelihood import GpuLikelihood
    from rbvfit_b200.model import GpuVoigtModel
    w = wl.get_workload(workload)
    cfg = FitConfiguration()
    for (z, ion, trans, comps) in w["systems"]:
        cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
    models = {}
    for name, inst in w["instruments"].items():
        taps = lsf.cos_like_taps(321) if inst.get("lsf") == "cos_like" else None
        models[name] = GpuVoigtModel(cfg, FWHM=inst["FWHM"], lsf_taps=taps, voigt_method=voigt_method)
    compiled = {n: m.compile() for n, m in models.items()}
    spectra = wl.make_spectra(w, lambda n, th, wave: compiled[n].model_flux(th, wave))
    like = GpuLikelihood({n: dict(model=models[n], **spectra[n]) for n in models}, w["lb"], w["ub"])
    return w, models, compiled, spectra, like, wl.make_ensemble(w, nwalkers)


def _host_lnprob_from_flux(compiled, spectra, like, thetas, lb, ub):
    """lnprob recomputed on the host with the reference's numpy formula (vfit_mcmc.py:309-311) from the GPU's
    model flux -- ties the flux kernel, the chi^2 kernel and the finalisation together."""
    out = np.zeros(len(thetas))
    for n, c in compiled.items():
        d = like.instrument_data[n]
        m = c.model_flux(thetas, spectra[n]["wave"])
        out += -0.5 * np.sum((d["flux"][None, :] - m) ** 2 * d["inv_sigma2"][None, :] - d["log_inv_sigma2"][None, :],
                             axis=1)
    bad = np.any((thetas < lb) | (thetas > ub), axis=1)
    out[bad] = -np.inf
    return out


@pytest.mark.parametrize("workload,nw", [("C5a", 48), ("C3", 64), ("C4w", 32), ("C2", 80)])
def test_lnprob_equals_host_formula_on_gpu_flux(workload, nw):
    w, models, compiled, spectra, like, thetas = _problem(workload, nw)
    got = like.lnprob(thetas)
    ref = _host_lnprob_from_flux(compiled, spectra, like, thetas, w["lb"], w["ub"])
    assert np.array_equal(np.isneginf(got), np.isneginf(ref))
    fin = np.isfinite(ref)
    assert fin.sum() >= nw - 3
    assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= 1e-11


def test_c5a_batching_permutation_and_tile_geometry_invariance():
    w, models, compiled, spectra, like, thetas = _problem("C5a", 700)
    full = like.lnprob(thetas)                       # 700 walkers -> big tiles (scale 4)
    assert np.isneginf(full).sum() >= 1
    # sub-batches of 5 walkers use the smallest tiles: different tile sizes, same answer
    idx = np.arange(0, 700, 70)
    small = np.concatenate([like.lnprob(thetas[i:i + 5]) for i in idx])
    ref = np.concatenate([full[i:i + 5] for i in idx])
    fin = np.isfinite(ref)
    assert np.array_equal(np.isneginf(small), np.isneginf(ref))
    assert np.max(np.abs(small[fin] - ref[fin]) / np.abs(ref[fin])) <= 1e-12
    perm = np.random.default_rng(1).permutation(700)
    assert np.array_equal(like.lnprob(thetas[perm]), full[perm], equal_nan=True)     # bit-exact
    assert np.array_equal(like.lnprob(thetas), full, equal_nan=True)                 # run-to-run
    # a page-locked theta is copied to the device in place (no staging copy): same bits, and the query tells them apart
    pin = like.pinned_theta(700)
    pin[:] = thetas
    assert like.engine.lib.rbv_host_pinned(pin.ctypes.data) == 1
    assert like.engine.lib.rbv_host_pinned(thetas.ctypes.data) == 0
    assert np.array_equal(like.lnprob(pin), full, equal_nan=True)
    pin[3] = thetas[5]                                                               # refilled in place between calls
    assert like.lnprob(pin)[3] == full[5]


def test_additivity_over_pixel_ranges_without_lsf():
    """Without an LSF the likelihood is a plain sum over pixels: lnL(spectrum) = lnL(left) + lnL(right)."""
    from rbvfit_b200 import FitConfiguration, workloads as wl
    from rbvfit_b200.likelihood import GpuLikelihood
    from rbvfit_b200.model import GpuVoigtModel
    w = wl.get_workload("C2")
    cfg = FitConfiguration()
    for (z, ion, trans, comps) in w["systems"]:
        cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
    model = GpuVoigtModel(cfg, FWHM=None)
    wave = w["instruments"]["SPEC"]["wave"]
    rng = np.random.default_rng(3)
    flux = 1.0 + 0.05 * rng.standard_normal(wave.size)
    err = np.full(wave.size, 0.05)
    thetas = wl.make_ensemble(w, 16, frac_out_of_bounds=0.0)
    cut = 7777

    def like_for(sl):
        return GpuLikelihood({"S": dict(model=model, wave=wave[sl], flux=flux[sl], error=err[sl])}, w["lb"], w["ub"])

    whole = like_for(slice(None)).lnprob(thetas)
    parts = like_for(slice(0, cut)).lnprob(thetas) + like_for(slice(cut, None)).lnprob(thetas)
    assert np.max(np.abs(whole - parts) / np.abs(whole)) <= 1e-12


def test_flux_bounds_and_monotonicity_in_column_density():
    """0 <= flux <= 1 everywhere and more column density never increases the (unconvolved) flux."""
    w, models, compiled, spectra, like, thetas = _problem("C4w", 4)
    m = models["SPEC"]
    wave = spectra["SPEC"]["wave"]
    th = w["theta_true"].copy()
    f0 = m.evaluate(th, wave, return_unconvolved=True)
    th2 = th.copy()
    th2[:3] += 0.3
    f1 = m.evaluate(th2, wave, return_unconvolved=True)
    assert f0.min() >= 0.0 and f0.max() <= 1.0
    assert np.all(f1 <= f0 + 1e-15)
    assert f0.min() == 0.0          # the DLA core is black: exp(-tau) underflows cleanly


def test_error_zero_gives_nan_and_empty_batch():
    w, models, compiled, spectra, like, thetas = _problem("C1", 8)
    assert like.lnprob(np.zeros((0, like.ndim))).shape == (0,)
    from rbvfit_b200.likelihood import GpuLikelihood
    s = spectra["COS"]
    err = s["error"].copy()
    err[100] = 0.0                  # examples/example_voigt_fitter.py:42-44 masks pixels this way
    bad = GpuLikelihood({"COS": dict(model=models["COS"], wave=s["wave"], flux=s["flux"], error=err)},
                        w["lb"], w["ub"])
    with np.errstate(all="ignore"):
        out = bad.lnprob(thetas)
    inb = ~np.any((thetas < w["lb"]) | (thetas > w["ub"]), axis=1)
    assert np.all(np.isnan(out[inb]))            # inf - inf, as numpy would give


def test_components_are_unconvolved_per_line_fluxes():
    w, models, compiled, spectra, like, thetas = _problem("C1", 4)
    m = models["COS"]
    wave = spectra["COS"]["wave"]
    res = m.evaluate(w["theta_true"], wave, return_components=True)
    assert set(res) == {"flux", "components", "component_info"} and len(res["components"]) == 4
    prod = np.prod(np.array(res["components"]), axis=0)          # exp(-sum tau_i) = prod exp(-tau_i)
    unc = m.evaluate(w["theta_true"], wave, return_unconvolved=True)
    assert np.max(np.abs(prod - unc)) <= 1e-13
    assert res["component_info"][0]["lambda0"] == float(m.atomic_lambda0[0])


def test_sightline_batch_equals_individual_fits():
    """C5b-style survey batch: one launch over S sightlines == S separate single-sightline likelihoods."""
    from rbvfit_b200 import FitConfiguration, workloads as wl
    from rbvfit_b200.likelihood import GpuLikelihood, SightlineBatch
    from rbvfit_b200.model import GpuVoigtModel
    S, Ws = 24, 16
    sight, thetas, singles = [], [], []
    w0 = wl.c5b_sightline(0)
    for s in range(S):
        w = wl.c5b_sightline(s)
        cfg = FitConfiguration()
        for (z, ion, trans, comps) in w["systems"]:
            cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
        m = GpuVoigtModel(cfg, FWHM="6.5")
        c = m.compile()
        sp = wl.make_spectra(w, lambda n, th, wave: c.model_flux(th, wave))["COS"]
        sight.append(dict(model=m, **sp))
        thetas.append(wl.make_ensemble(w, Ws, frac_out_of_bounds=0.1))
        singles.append(GpuLikelihood({"COS": dict(model=m, **sp)}, w["lb"], w["ub"]))
    thetas = np.array(thetas)
    batch = SightlineBatch(sight, w0["lb"], w0["ub"])
    got = batch.lnprob(thetas)
    ref = np.array([singles[s].lnprob(thetas[s]) for s in range(S)])
    assert got.shape == (S, Ws)
    assert np.isneginf(ref).sum() >= S
    # 16-walker calls use smaller tiles than the 384-walker batch (geometry follows the batch size): same values up
    # to the order of the tile sums ...
    assert np.array_equal(np.isneginf(got), np.isneginf(ref))
    fin = np.isfinite(ref)
    assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= 1e-12
    # ... and bit-identical when the single-sightline call has the batch's size (same kernel, same tiles)
    same = np.array([singles[s].lnprob(np.tile(thetas[s], (S, 1)))[:Ws] for s in range(0, S, 5)])
    assert np.array_equal(got[0:S:5], same, equal_nan=True)
    assert len(set(np.round(ref[np.isfinite(ref)], 3))) > S     # sightlines really differ
    pin = batch.pinned_theta(Ws)                     # page-locked input: same values
    pin[:] = thetas
    assert np.array_equal(batch.lnprob(pin), got, equal_nan=True)
    with pytest.raises(ValueError):
        batch.lnprob(thetas[:5])


@pytest.mark.parametrize("workload,nw", [("C5a", 96), ("C2", 80), ("C4", 32), ("C4w", 32), ("C3", 50)])
def test_far_field_interpolant_matches_direct_evaluation(workload, nw):
    """Default path (Chebyshev far field per 1024-px super-chunk) against the direct evaluation of every
    (line, pixel) pair: the a-priori gate promises |dtau| <= 1e-13 per pixel."""
    w, models, compiled, spectra, like, thetas = _problem(workload, nw)
    name0 = like.names[0]
    ok = thetas[np.all((thetas >= w["lb"]) & (thetas <= w["ub"]), axis=1)][:6]
    res = {}
    for mode in ("direct", "chebyshev"):
        like.set_far_field(mode)
        res[mode] = (like.lnprob(thetas), like.engine.model_flux(0, ok))
    (la, fa), (lb_, fb) = res["direct"], res["chebyshev"]
    assert np.array_equal(np.isneginf(la), np.isneginf(lb_))
    fin = np.isfinite(la)
    assert np.max(np.abs(la[fin] - lb_[fin]) / np.abs(la[fin])) <= 1e-12
    assert np.max(np.abs(fa - fb)) <= 2e-13
    with pytest.raises(ValueError):
        like.set_far_field("fmm")


def test_far_field_stress_strong_lines_and_extreme_widths():
    """Far-field gate under stress: damped (logN 22) and very narrow / very broad components next to weak ones, on
    a coarse and on a fine wavelength grid; direct and interpolated paths must agree to the flux tolerance."""
    from rbvfit_b200 import FitConfiguration
    from rbvfit_b200.likelihood import GpuLikelihood
    from rbvfit_b200.model import GpuVoigtModel
    rng = np.random.default_rng(5)
    cfg = FitConfiguration()
    cfg.add_system(z=2.5, ion="HI", transitions=[1215.67, 1025.72, 972.54], components=2)
    cfg.add_system(z=2.1, ion="CIV", transitions=[1548.2, 1550.77], components=2)
    cfg.add_system(z=1.9, ion="SiIV", transitions=[1393.76, 1402.77], components=1)
    model = GpuVoigtModel(cfg, FWHM="6.5")
    C = 5
    for P in (3000, 60000):
        wave = np.linspace(3200.0, 5600.0, P)
        rows = []
        for _ in range(24):
            logN = np.array([rng.uniform(19.0, 22.0), rng.uniform(12.0, 15.0), rng.uniform(12.0, 16.0),
                             rng.uniform(12.0, 14.0), rng.uniform(12.0, 17.0)])
            b = np.array([rng.uniform(15, 80), rng.uniform(1.0, 5.0), rng.uniform(2, 150), rng.uniform(5, 30),
                          rng.uniform(1.0, 200.0)])
            v = rng.uniform(-300, 300, C)
            rows.append(np.concatenate([logN, b, v]))
        thetas = np.array(rows)
        lb, ub = np.full(3 * C, -1e9), np.full(3 * C, 1e9)
        flux = 1.0 + 0.05 * rng.standard_normal(P)
        like = GpuLikelihood({"S": dict(model=model, wave=wave, flux=flux, error=np.full(P, 0.05))}, lb, ub)
        out = {}
        for mode in ("direct", "chebyshev"):
            like.set_far_field(mode)
            out[mode] = (like.lnprob(thetas), like.engine.model_flux(0, thetas[:8]))
        (la, fa), (lb_, fb) = out["direct"], out["chebyshev"]
        assert np.all(np.isfinite(la))
        assert np.max(np.abs(fa - fb)) <= 1e-12, P
        assert np.max(np.abs(la - lb_) / np.abs(la)) <= 1e-11, P
        like.close()


def test_wavelength_dependent_lsf_and_weight_mask_vs_oracle():
    """Extension: a spectrum whose LSF changes along the wavelength axis, run as a joint fit of its
    blocks with weight-0 halos -- lnprob and the stitched model flux against the oracle's direct definition (every
    output pixel under the kernel of its own block); plus the weight mask on its own (masked pixels drop out of both
    likelihood terms exactly)."""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_host_logic import _piecewise_problem
    from oracle import voigt_oracle as vo
    from rbvfit_b200 import lsf, workloads as wl
    from rbvfit_b200.likelihood import GpuLikelihood
    w, cfg, wave, flux, error, starts, models, lowered = _piecewise_problem()
    entries = lsf.piecewise_lsf_instruments("COS", wave, flux, error, list(zip(starts, models)))
    like = GpuLikelihood(entries, w["lb"], w["ub"])
    thetas = wl.make_ensemble(w, 40)
    got = like.lnprob(thetas)
    ref = np.empty(len(thetas))
    for i, th in enumerate(thetas):
        if np.any(th < w["lb"]) or np.any(th > w["ub"]):
            ref[i] = -np.inf
            continue
        m = vo.model_flux_piecewise_lsf(lowered, starts, th, wave)
        ref[i] = -0.5 * np.sum((flux - m) ** 2 / error ** 2 - np.log(1.0 / error ** 2))
    assert np.array_equal(np.isneginf(got), np.isneginf(ref)) and np.isneginf(ref).sum() >= 1
    fin = np.isfinite(ref)
    assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= 1e-9
    compiled = {n: d["model"].compile() for n, d in entries.items()}
    stitched = lsf.piecewise_lsf_flux(entries, lambda n, d: compiled[n].model_flux(thetas[0], d["wave"]))
    assert np.max(np.abs(stitched - vo.model_flux_piecewise_lsf(lowered, starts, thetas[0], wave))) <= 1e-10
    like.close()
    # the mask alone: a single-kernel instrument with every third pixel masked == the oracle's sum over the rest
    keep = (np.arange(wave.size) % 3) != 0
    like = GpuLikelihood({"COS": dict(model=models[0], wave=wave, flux=flux, error=error, weight_mask=keep)},
                         w["lb"], w["ub"])
    got = like.lnprob(thetas[:6])
    for i, th in enumerate(thetas[:6]):
        if not np.isfinite(got[i]):
            continue
        m = vo.model_flux(lowered[0], th, wave)
        r = -0.5 * np.sum(((flux - m) ** 2 / error ** 2 - np.log(1.0 / error ** 2))[keep])
        assert abs(got[i] - r) <= 1e-9 * abs(r)
    like.close()


def test_sightline_batch_of_long_many_line_spectra():
    """Survey batch whose spectra are long (several ranges of the stream kernel) and have 8 lines (boundary
    records): with the stream kernel forced (RBVFIT_B200_STREAM=1, the second run of the GPU suite) this is the
    sightline launch of voigt_stream_kernel<3, true> + finalize_stream_kernel; by default the tile kernel takes it.
    Either way: one launch over S sightlines == S separate likelihoods, and the oracle on a few rows."""
    from oracle import voigt_oracle as vo
    from rbvfit_b200 import FitConfiguration, workloads as wl
    from rbvfit_b200.likelihood import GpuLikelihood, SightlineBatch
    from rbvfit_b200.model import GpuVoigtModel
    S, Ws = 5, 12
    base = wl.get_workload("C4")
    sight, thetas, singles, lowered = [], [], [], []
    for s in range(S):
        w = dict(base)
        w["seed"] = base["seed"] + 31 * s
        z = 2.5 + 0.004 * s
        w["systems"] = [(z, ion, trans, comps) for (_z, ion, trans, comps) in base["systems"]]
        w["instruments"] = {"SPEC": dict(wave=np.linspace(3500.0, 5500.0, 6000) * (1 + 0.004 * s / 3.5), FWHM="6.5",
                                         lsf=None)}
        cfg = FitConfiguration()
        for (zz, ion, trans, comps) in w["systems"]:
            cfg.add_system(z=zz, ion=ion, transitions=trans, components=comps)
        m = GpuVoigtModel(cfg, FWHM="6.5")
        c = m.compile()
        sp = wl.make_spectra(w, lambda n, th, wave: c.model_flux(th, wave))["SPEC"]
        sight.append(dict(model=m, **sp))
        thetas.append(wl.make_ensemble(w, Ws, frac_out_of_bounds=0.1))
        singles.append(GpuLikelihood({"SPEC": dict(model=m, **sp)}, w["lb"], w["ub"]))
        lowered.append((vo.lower(cfg, FWHM="6.5"), sp))
    thetas = np.array(thetas)
    batch = SightlineBatch(sight, base["lb"], base["ub"])
    got = batch.lnprob(thetas)
    ref = np.array([singles[s].lnprob(thetas[s]) for s in range(S)])
    assert np.array_equal(np.isneginf(got), np.isneginf(ref)) and np.isneginf(ref).sum() >= S
    fin = np.isfinite(ref)
    assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= 1e-12
    for s in (0, S - 1):
        m, sp = lowered[s]
        comp = vo.compile_instruments({"SPEC": dict(model=m, **sp)})
        for i in range(3):
            o = vo.lnprob(comp, thetas[s, i], base["lb"], base["ub"])
            assert (np.isneginf(o) and np.isneginf(got[s, i])) or abs(got[s, i] - o) <= 1e-9 * abs(o)
    assert np.array_equal(batch.lnprob(thetas), got, equal_nan=True)
    batch.close()
    for one in singles:
        one.close()
