#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz with the REAL reference code.

Runs in the build container only (needs /root/reference).  For every case it builds the
reference's own FitConfiguration -> VoigtModel -> vfit objects (through oracle/refshim.py) and
records, for a seeded batch of theta:
    * the lowered model arrays (atomic_lambda0, atomic_gamma[f32], atomic_f[f32], z_factors,
      N_indices) and the LSF taps the reference applied,
    * the observed spectra handed to vfit,
    * reference model flux for a few walkers and reference lnprob for every walker.
The fixtures are committed; tests compare the oracle AND the CUDA path against them.

    python -m oracle.make_golden          # regenerate everything
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, ".."))
sys.path.insert(0, ROOT)

from oracle import refshim, voigt_oracle as vo          # noqa: E402
from rbvfit_b200 import workloads as wl                  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
N_FLUX_ROWS = 2


def _ref_models(w, voigt_method="wofz"):
    FitConfiguration, VoigtModel, mc, vm = refshim.import_reference()
    from astropy.convolution import CustomKernel
    config = FitConfiguration()
    for (z, ion, trans, comps) in w["systems"]:
        config.add_system(z=z, ion=ion, transitions=list(trans), components=comps)
    models = {}
    for name, inst in w["instruments"].items():
        m = VoigtModel(config, FWHM=inst["FWHM"], voigt_method=voigt_method)
        if inst.get("lsf") == "cos_like":
            # the reference's COS branch needs linetools; inject the table the same way
            # _setup_kernel would (core/voigt_model.py:458-460)
            m.kernel = CustomKernel(vo.cos_like_lsf(321))
        models[name] = m
    return config, models, mc


def build_case(case_name, w, voigt_method="wofz", error_dtype=np.float64, nwalkers=None,
               extra_thetas=None):
    config, models, mc = _ref_models(w, voigt_method)
    compiled = {n: m.compile() for n, m in models.items()}
    spectra = wl.make_spectra(w, lambda n, th, wave: compiled[n].model_flux(th, wave),
                              error_dtype=error_dtype)
    thetas = wl.make_ensemble(w, nwalkers)
    if extra_thetas is not None:
        thetas = np.vstack([thetas, extra_thetas])
    inst_data = {n: dict(model=models[n], wave=s["wave"], flux=s["flux"], error=s["error"])
                 for n, s in spectra.items()}
    with contextlib.redirect_stdout(io.StringIO()):
        fitter = mc.vfit(inst_data, w["theta_true"], w["lb"], w["ub"])
    with np.errstate(all="ignore"):
        ref_lnprob = np.array([fitter.lnprob(t) for t in thetas])
    finite = np.flatnonzero(np.isfinite(ref_lnprob))[:N_FLUX_ROWS]
    out = dict(
        meta=json.dumps(dict(case=case_name, workload=w["name"], voigt_method=voigt_method,
                             error_dtype=np.dtype(error_dtype).name,
                             instruments=list(w["instruments"].keys()),
                             systems=[(z, ion, list(t), c) for (z, ion, t, c) in w["systems"]])),
        thetas=thetas, lb=w["lb"], ub=w["ub"], ref_lnprob=ref_lnprob, flux_rows=finite)
    for n, m in models.items():
        d = compiled[n].data
        out[f"{n}__lambda0"] = d.atomic_lambda0
        out[f"{n}__gamma"] = d.atomic_gamma
        out[f"{n}__f"] = d.atomic_f
        out[f"{n}__zfac"] = d.z_factors
        out[f"{n}__N_indices"] = d.N_indices
        out[f"{n}__taps"] = (np.zeros(0) if d.kernel is None else np.asarray(d.kernel.array))
        out[f"{n}__kernel_kind"] = np.array(
            "none" if d.kernel is None else
            ("gaussian" if type(d.kernel).__name__ == "Gaussian1DKernel" else "custom"))
        out[f"{n}__wave"] = spectra[n]["wave"]
        out[f"{n}__flux"] = spectra[n]["flux"]
        out[f"{n}__error"] = spectra[n]["error"]
        out[f"{n}__ref_flux"] = np.array([compiled[n].model_flux(thetas[i], spectra[n]["wave"])
                                          for i in finite])
        out[f"{n}__ref_flux_unconvolved"] = np.array(
            [m.evaluate(thetas[i], spectra[n]["wave"], return_unconvolved=True) for i in finite[:1]])
    path = os.path.join(GOLDEN_DIR, f"{case_name}.npz")
    np.savez_compressed(path, **out)
    print(f"{case_name}: W={len(thetas)} finite={np.isfinite(ref_lnprob).sum()} "
          f"lnprob[0]={ref_lnprob[0]!r} -> {os.path.relpath(path, ROOT)} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


def build_test_script_case():
    """examples/test_script.py:38-63 + examples/cos_data.npz (float32 flux/error)."""
    FitConfiguration, VoigtModel, mc, vm = refshim.import_reference()
    d = np.load(os.path.join(refshim.REFERENCE_SRC, "rbvfit", "examples", "cos_data.npz"))
    systems = [(0.0, "SiII", [1190.416, 1193.290], 1), (0.162005, "HI", [1025.722], 1)]
    config = FitConfiguration()
    for (z, ion, trans, comps) in systems:
        config.add_system(z=z, ion=ion, transitions=trans, components=comps)
    fwhm = 2.394991274145626
    model = VoigtModel(config, FWHM=fwhm)
    theta0 = np.array([14.4228, 14.5105, 46.4066, 46.7789, -28.2628, -6.743])
    lb = np.array([10., 10., 1., 1., -500., -500.])
    ub = np.array([20., 20., 100., 100., 500., 500.])
    rng = np.random.default_rng(20260)
    thetas = np.vstack([theta0,
                        theta0 + np.array([0.1, 0.1, 3, 3, 4, 4]) * rng.standard_normal((14, 6)),
                        theta0 + np.array([100.0, 0, 0, 0, 0, 0])])
    inst = {"COS": dict(model=model, wave=d["wave"], flux=d["flux"], error=d["error"])}
    with contextlib.redirect_stdout(io.StringIO()):
        fitter = mc.vfit(inst, theta0, lb, ub)
    ref_lnprob = np.array([fitter.lnprob(t) for t in thetas])
    comp = model.compile()
    rows = np.array([0, 1])
    out = dict(
        meta=json.dumps(dict(case="test_script", workload="test_script", voigt_method="wofz",
                             error_dtype="float32", instruments=["COS"], systems=systems,
                             FWHM=fwhm)),
        thetas=thetas, lb=lb, ub=ub, ref_lnprob=ref_lnprob, flux_rows=rows)
    cd = comp.data
    out.update({"COS__lambda0": cd.atomic_lambda0, "COS__gamma": cd.atomic_gamma, "COS__f": cd.atomic_f,
                "COS__zfac": cd.z_factors, "COS__N_indices": cd.N_indices,
                "COS__taps": np.asarray(cd.kernel.array), "COS__kernel_kind": np.array("gaussian"),
                "COS__wave": d["wave"], "COS__flux": d["flux"], "COS__error": d["error"],
                "COS__ref_flux": np.array([comp.model_flux(thetas[i], d["wave"]) for i in rows]),
                "COS__ref_flux_unconvolved": np.array(
                    [model.evaluate(thetas[0], d["wave"], return_unconvolved=True)]),
                "COS__inv_sigma2": fitter.instrument_data["COS"]["inv_sigma2"],
                "COS__log_inv_sigma2": fitter.instrument_data["COS"]["log_inv_sigma2"]})
    path = os.path.join(GOLDEN_DIR, "test_script.npz")
    np.savez_compressed(path, **out)
    print(f"test_script: lnprob(theta0)={ref_lnprob[0]!r} last={ref_lnprob[-1]!r} "
          f"min flux={out['COS__ref_flux'][0].min()!r}")


def build_tutorial_case(case_name, files, fwhms, systems, nguess, bguess, vguess, wave_scale=1.0, seed=20270):
    """The reference's own multi-instrument tutorials on the REAL spectra that ship with it
    (examples/rbvfit2-multi-instrument-tutorial.py:82-226 and ...-tutorial2.py:76-220): the rb_spec JSON slices
    (`wave_slice`, `fnorm`, `enorm`; read directly, rbcodes is only a loader there), one VoigtModel per instrument
    with its own FWHM, bounds from the reference's `mc.set_bounds` defaults, joint lnprob."""
    FitConfiguration, VoigtModel, mc, vm = refshim.import_reference()
    config = FitConfiguration()
    for (z, ion, trans, comps) in systems:
        config.add_system(z=z, ion=ion, transitions=list(trans), components=comps)
    ex = os.path.join(refshim.REFERENCE_SRC, "rbvfit", "examples")
    inst, models = {}, {}
    for name, fname in files.items():
        with open(os.path.join(ex, fname)) as fh:
            d = json.load(fh)
        wave = np.asarray(d["wave_slice"], dtype=np.float64) * wave_scale
        models[name] = VoigtModel(config, FWHM=fwhms[name])
        inst[name] = dict(model=models[name], wave=wave, flux=np.asarray(d["fnorm"], dtype=np.float64),
                          error=np.asarray(d["enorm"], dtype=np.float64))
    theta0 = np.concatenate([nguess, bguess, vguess]).astype(np.float64)
    with contextlib.redirect_stdout(io.StringIO()):
        _bounds, lb, ub = mc.set_bounds(nguess, bguess, vguess)
        fitter = mc.vfit(inst, theta0, lb, ub)
    lb, ub = np.asarray(lb, dtype=np.float64), np.asarray(ub, dtype=np.float64)
    C = len(nguess)
    rng = np.random.default_rng(seed)
    scale = np.concatenate([np.full(C, 0.1), np.full(C, 3.0), np.full(C, 5.0)])
    outside = theta0.copy()
    outside[2 * C] = ub[2 * C] + 1.0
    thetas = np.vstack([theta0, np.clip(theta0 + scale * rng.standard_normal((18, 3 * C)), lb, ub), lb, ub, outside])
    with np.errstate(all="ignore"):
        ref_lnprob = np.array([fitter.lnprob(t) for t in thetas])
    finite = np.flatnonzero(np.isfinite(ref_lnprob))[:N_FLUX_ROWS]
    out = dict(
        meta=json.dumps(dict(case=case_name, workload=case_name, voigt_method="wofz", error_dtype="float64",
                             instruments=list(files.keys()), FWHM=fwhms,
                             systems=[(z, ion, list(t), c) for (z, ion, t, c) in systems])),
        thetas=thetas, lb=lb, ub=ub, ref_lnprob=ref_lnprob, flux_rows=finite)
    for n, m in models.items():
        comp = m.compile()
        d = comp.data
        out[f"{n}__lambda0"] = d.atomic_lambda0
        out[f"{n}__gamma"] = d.atomic_gamma
        out[f"{n}__f"] = d.atomic_f
        out[f"{n}__zfac"] = d.z_factors
        out[f"{n}__N_indices"] = d.N_indices
        out[f"{n}__taps"] = np.asarray(d.kernel.array)
        out[f"{n}__kernel_kind"] = np.array("gaussian")
        for k in ("wave", "flux", "error"):
            out[f"{n}__{k}"] = inst[n][k]
        out[f"{n}__ref_flux"] = np.array([comp.model_flux(thetas[i], inst[n]["wave"]) for i in finite])
        out[f"{n}__ref_flux_unconvolved"] = np.array(
            [m.evaluate(thetas[i], inst[n]["wave"], return_unconvolved=True) for i in finite[:1]])
    path = os.path.join(GOLDEN_DIR, f"{case_name}.npz")
    np.savez_compressed(path, **out)
    print(f"{case_name}: W={len(thetas)} finite={np.isfinite(ref_lnprob).sum()} lnprob[0]={ref_lnprob[0]!r} "
          f"taps={[len(out[n + '__taps']) for n in models]} -> {os.path.relpath(path, ROOT)} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")



def flux_pixel_subset(P, n_lines_hint=8, seed=7):
    """Pixels at which the big fixtures keep the reference flux: every 8th pixel plus dense 1024-pixel windows
    (seeded) -- a 100 000-pixel row costs 0.8 MB, the subset 0.2 MB."""
    rng = np.random.default_rng(seed)
    px = set(range(0, P, 8))
    for start in rng.integers(0, max(P - 1024, 1), size=n_lines_hint):
        px.update(range(int(start), min(int(start) + 1024, P)))
    px.update(range(0, min(64, P)))
    px.update(range(max(P - 64, 0), P))
    return np.array(sorted(px), dtype=np.int64)


def build_big_case(case_name, w, n_in=8, n_out=2):
    """The headline geometry (C5a: 100 000 px, 33 lines; C5a_L4: the 4-line companion): `n_in` in-bounds and `n_out`
    out-of-bounds rows of the workload's own 8192-walker ensemble through the REAL reference (~0.5 s per row).
    Compact storage: the wavelength grid as its linspace arguments, the constant error as a scalar, the reference
    flux of two rows on a pixel subset."""
    config, models, mc = _ref_models(w)
    assert len(models) == 1
    (name, model), = models.items()
    compiled = model.compile()
    spectra = wl.make_spectra(w, lambda n, th, wave: compiled.model_flux(th, wave))
    s = spectra[name]
    ens = wl.make_ensemble(w)                       # the ensemble bench.py evaluates
    inb = np.all((ens >= w["lb"]) & (ens <= w["ub"]), axis=1)
    rows = np.concatenate([np.flatnonzero(inb)[:n_in], np.flatnonzero(~inb)[:n_out]])
    thetas = ens[rows]
    with contextlib.redirect_stdout(io.StringIO()):
        fitter = mc.vfit({name: dict(model=model, **s)}, w["theta_true"], w["lb"], w["ub"])
    ref_lnprob = np.array([fitter.lnprob(t) for t in thetas])
    px = flux_pixel_subset(s["wave"].size)
    d = compiled.data
    wave = s["wave"]
    assert np.array_equal(wave, np.linspace(wave[0], wave[-1], wave.size))
    assert np.all(s["error"] == s["error"][0])
    out = dict(
        meta=json.dumps(dict(case=case_name, workload=w["name"], voigt_method="wofz", error_dtype="float64",
                             instruments=[name], ensemble_rows=[int(r) for r in rows],
                             systems=[(z, ion, list(t), c) for (z, ion, t, c) in w["systems"]])),
        thetas=thetas, lb=w["lb"], ub=w["ub"], ref_lnprob=ref_lnprob, flux_rows=np.array([0, 1]))
    out.update({f"{name}__lambda0": d.atomic_lambda0, f"{name}__gamma": d.atomic_gamma, f"{name}__f": d.atomic_f,
                f"{name}__zfac": d.z_factors, f"{name}__N_indices": d.N_indices,
                f"{name}__taps": np.asarray(d.kernel.array), f"{name}__kernel_kind": np.array("gaussian"),
                f"{name}__wave_linspace": np.array([wave[0], wave[-1], wave.size]),
                f"{name}__error_const": np.array(s["error"][0]), f"{name}__flux": s["flux"],
                f"{name}__flux_px": px,
                f"{name}__ref_flux": np.array([compiled.model_flux(thetas[i], wave)[px] for i in (0, 1)]),
                f"{name}__ref_flux_unconvolved": np.array(
                    [model.evaluate(thetas[0], wave, return_unconvolved=True)[px]])})
    path = os.path.join(GOLDEN_DIR, f"{case_name}.npz")
    np.savez_compressed(path, **out)
    print(f"{case_name}: rows={list(rows)} lnprob={ref_lnprob} -> {os.path.relpath(path, ROOT)} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


def build_c5b_case(n_sightlines=8, walkers=16):
    """Survey mode (C5b): `n_sightlines` independent sightlines (C1's structure at its own redshift, its own
    spectrum = reference model(theta_true) + noise) x `walkers` rows each, every sightline through its own
    reference vfit -- what S x vfit(...).lnprob does (vfit_mcmc.py:127-197, 348-353)."""
    zs, waves, fluxes, errs, thetas, lnps = [], [], [], [], [], []
    w0 = wl.c5b_sightline(0)
    for sidx in range(n_sightlines):
        w = wl.c5b_sightline(sidx)
        config, models, mc = _ref_models(w)
        model = models["COS"]
        compiled = model.compile()
        s = wl.make_spectra(w, lambda n, th, wave: compiled.model_flux(th, wave))["COS"]
        th = wl.make_ensemble(w, walkers)
        with contextlib.redirect_stdout(io.StringIO()):
            fitter = mc.vfit({"COS": dict(model=model, **s)}, w["theta_true"], w["lb"], w["ub"])
        zs.append(w["systems"][0][0])
        waves.append([s["wave"][0], s["wave"][-1], s["wave"].size])
        assert np.array_equal(s["wave"], np.linspace(s["wave"][0], s["wave"][-1], s["wave"].size))
        fluxes.append(s["flux"])
        errs.append(s["error"][0])
        thetas.append(th)
        lnps.append([fitter.lnprob(t) for t in th])
    path = os.path.join(GOLDEN_DIR, "C5b.npz")
    np.savez_compressed(path, meta=json.dumps(dict(case="C5b", n_sightlines=n_sightlines, walkers=walkers)),
                        z=np.array(zs), wave_linspace=np.array(waves), flux=np.array(fluxes),
                        error_const=np.array(errs), thetas=np.array(thetas), ref_lnprob=np.array(lnps),
                        lb=w0["lb"], ub=w0["ub"])
    print(f"C5b: {n_sightlines} sightlines x {walkers} walkers, finite={np.isfinite(np.array(lnps)).sum()} -> "
          f"{os.path.relpath(path, ROOT)} ({os.path.getsize(path) / 1024:.0f} KiB)")


def build_wofz_lattice():
    """Known-answer lattice for Re w(x + i a): scipy.special.wofz (the reference's call) and
    mpmath at 40 digits.  Covers core, mid, far wings and the whole a range."""
    import mpmath as mp
    from scipy.special import wofz
    mp.mp.dps = 40
    xs = np.concatenate([np.linspace(0, 9, 37), np.geomspace(9.5, 3e4, 28)])
    As = np.array([1e-8, 1e-6, 5e-6, 1.5e-4, 2.8e-3, 1e-2, 3.4e-2, 5e-2, 7e-2, 0.3, 1.0, 4.0, 20.0])
    X, A = np.meshgrid(xs, As)
    ref_scipy = wofz(X + 1j * A).real
    ref_mp = np.zeros_like(X)
    for i in range(X.shape[0]):
        for j in range(X.shape[1]):
            z = mp.mpc(X[i, j], A[i, j])
            ref_mp[i, j] = float((mp.exp(-z * z) * mp.erfc(-1j * z)).real)
    path = os.path.join(GOLDEN_DIR, "wofz_lattice.npz")
    np.savez_compressed(path, x=X, a=A, scipy=ref_scipy, mpmath=ref_mp)
    print("wofz lattice: max rel |scipy-mpmath| =", np.max(np.abs(ref_scipy - ref_mp) / ref_mp))


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    build_test_script_case()
    build_wofz_lattice()
    c1 = wl.get_workload("C1")
    # extra rows: exact-bound rows (inclusive bounds), a tiny-b row (a > A_FAST path), NaN row
    extra = np.vstack([c1["theta_true"], c1["theta_true"], c1["theta_true"], c1["theta_true"]])
    extra[0, 0] = c1["ub"][0]            # on the bound -> allowed
    extra[1, 2] = c1["lb"][2]            # b on its lower bound
    extra[2, 4] = np.nextafter(c1["ub"][4], np.inf)   # one ulp outside -> -inf
    extra[3, 1] = np.nan                 # NaN passes the prior and poisons the likelihood
    build_case("C1", c1, extra_thetas=extra)
    build_case("C1_f32err", c1, error_dtype=np.float32)
    build_case("C1_fast", c1, voigt_method="fast")
    c1n = wl.get_workload("C1")
    c1n["instruments"]["COS"]["FWHM"] = None
    build_case("C1_nolsf", c1n, nwalkers=8)
    c1c = wl.get_workload("C1")
    c1c["instruments"]["COS"].update(FWHM=None, lsf="cos_like")
    build_case("C1_coslsf", c1c, nwalkers=16)
    build_case("C2", wl.get_workload("C2"), nwalkers=24)
    build_case("C3", wl.get_workload("C3"), nwalkers=16)
    build_case("C4", wl.get_workload("C4"), nwalkers=16)
    build_case("C4w", wl.get_workload("C4w"), nwalkers=8)
    # small-b / large-a stress: bounds opened so that b down to 0.05 km/s is legal
    cs = wl.get_workload("C1")
    cs["lb"] = cs["lb"].copy(); cs["lb"][2:4] = 0.01
    th = np.tile(cs["theta_true"], (6, 1))
    th[:, 2] = [0.02, 0.1, 0.5, 1.0, 1.9, 3.0]
    th[:, 3] = [3.0, 1.9, 1.0, 0.5, 0.1, 0.05]
    build_case("C1_smallb", cs, nwalkers=4, extra_thetas=th)
    build_tutorial_cases()
    build_big_cases()


def build_big_cases():
    build_big_case("C5a", wl.get_workload("C5a"))
    build_big_case("C5a_L4", wl.get_workload("C5a_L4"))
    build_c5b_case()


def build_tutorial_cases():
    # examples/rbvfit2-multi-instrument-tutorial.py: OI 1302 of one absorber seen by XShooter and FIRE
    build_tutorial_case("tutorial_oi1302",
                        {"XShooter": "J159_7921_XShooter_OI1302.json", "FIRE": "J159_7921_FIRE_OI1302.json"},
                        {"XShooter": "2.2", "FIRE": "4.0"},
                        [(0.0, "OI", [1302.17], 1)], [14.4], [18.0], [200.0])
    # examples/rbvfit2-multi-instrument-tutorial2.py: the same slices moved to the observed frame (z = 6.074762),
    # fitted as a 4-component CIV doublet at z = 4.9484 by XShooter + FIRE + HIRES
    build_tutorial_case("tutorial_civ3",
                        {"XShooter": "J1030_9089_XShooter_OI1302.json", "FIRE": "J1030_9089_FIRE_OI1302.json",
                         "HIRES": "J1030_9089_HIRES_OI_air2vac_updated.json"},
                        {"XShooter": "2.2", "FIRE": "4.0", "HIRES": "4.285"},
                        [(4.9484, "CIV", [1548.2, 1550.3], 4)],
                        [13.25, 13.63, 13.12, 13.2], [23.0, 25.0, 50.0, 13.2], [-67.0, 0.0, -20.0, -20.0],
                        wave_scale=6.074762 + 1.0)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        build_big_cases()         # only the headline-geometry / survey-mode fixtures
    else:
        main()
