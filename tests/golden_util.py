"""Helpers shared by the parity tests: load a tests/golden fixture and rebuild the CPU oracle
(and, for GPU tests, the product objects) from its metadata."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = ["test_script", "C1", "C1_f32err", "C1_fast", "C1_nolsf", "C1_coslsf", "C1_smallb",
         "C2", "C3", "C4", "C4w", "tutorial_oi1302", "tutorial_civ3"]
# headline geometry (100 000 px): compact fixtures -- wavelength grid as linspace arguments, constant error as a
# scalar, reference flux on the pixel subset `flux_px`; 8 in-bounds + 2 out-of-bounds rows of the bench ensemble
BIG_CASES = ["C5a", "C5a_L4"]


class Golden:
    def __init__(self, case):
        self.case = case
        self.z = np.load(os.path.join(GOLDEN_DIR, f"{case}.npz"), allow_pickle=False)
        self.meta = json.loads(str(self.z["meta"]))
        self.instruments = self.meta["instruments"]
        self.thetas = self.z["thetas"]
        self.lb, self.ub = self.z["lb"], self.z["ub"]
        self.ref_lnprob = self.z["ref_lnprob"]
        self.flux_rows = self.z["flux_rows"]

    def inst(self, name, key):
        full = f"{name}__{key}"
        if full not in self.z.files:
            if key == "wave":
                lo, hi, n = self.z[f"{name}__wave_linspace"]
                return np.linspace(lo, hi, int(n))
            if key == "error":
                return np.full(int(self.z[f"{name}__wave_linspace"][2]), float(self.z[f"{name}__error_const"]))
            if key == "flux_px":                      # small fixtures keep every pixel
                return np.arange(self.z[f"{name}__wave"].size)
        return self.z[full]

    def kernel_kind(self, name):
        return str(self.z[f"{name}__kernel_kind"])

    def fwhm_for(self, name):
        """FWHM argument that reproduces the fixture's kernel through the public API."""
        kind = self.kernel_kind(name)
        if kind == "none":
            return None, None
        if kind == "custom":
            return None, self.inst(name, "taps")
        if "FWHM" in self.meta:
            f = self.meta["FWHM"]
            return (f[name] if isinstance(f, dict) else f), None
        from rbvfit_b200 import workloads as wl
        w = wl.get_workload(self.meta["workload"])
        return w["instruments"][name]["FWHM"], None

    def oracle_models(self):
        from oracle import voigt_oracle as vo
        cfg = vo.OracleConfig()
        for (z, ion, trans, comps) in self.meta["systems"]:
            cfg.add_system(z, ion, trans, comps)
        models = {}
        for name in self.instruments:
            fwhm, taps = self.fwhm_for(name)
            models[name] = vo.lower(cfg, FWHM=fwhm, voigt_method=self.meta["voigt_method"],
                                    custom_taps=taps)
        return models

    def oracle_compiled(self):
        from oracle import voigt_oracle as vo
        models = self.oracle_models()
        inst = {n: dict(model=models[n], wave=self.inst(n, "wave"), flux=self.inst(n, "flux"),
                        error=self.inst(n, "error")) for n in self.instruments}
        return vo.compile_instruments(inst)


class GoldenSightlines:
    """tests/golden/C5b.npz: S independent sightlines x W walkers, every sightline through its own reference vfit."""

    def __init__(self):
        z = np.load(os.path.join(GOLDEN_DIR, "C5b.npz"), allow_pickle=False)
        self.z_sys = z["z"]
        self.waves = [np.linspace(lo, hi, int(n)) for lo, hi, n in z["wave_linspace"]]
        self.flux = z["flux"]
        self.errors = [np.full(f.size, e) for f, e in zip(self.flux, z["error_const"])]
        self.thetas = z["thetas"]                 # [S, W, ndim]
        self.ref_lnprob = z["ref_lnprob"]         # [S, W]
        self.lb, self.ub = z["lb"], z["ub"]
        self.n = len(self.z_sys)

    def systems(self, s):
        return [(float(self.z_sys[s]), "MgII", [2796.3, 2803.5], 2)]

    def oracle_compiled(self, s):
        from oracle import voigt_oracle as vo
        cfg = vo.OracleConfig()
        for (z, ion, trans, comps) in self.systems(s):
            cfg.add_system(z, ion, trans, comps)
        m = vo.lower(cfg, FWHM="6.5")
        return vo.compile_instruments({"COS": dict(model=m, wave=self.waves[s], flux=self.flux[s],
                                                   error=self.errors[s])})
