"""Helpers shared by the parity tests: load a tests/golden fixture and rebuild the CPU oracle
(and, for GPU tests, the product objects) from its metadata."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = ["test_script", "C1", "C1_f32err", "C1_fast", "C1_nolsf", "C1_coslsf", "C1_smallb",
         "C2", "C3", "C4", "C4w", "tutorial_oi1302", "tutorial_civ3"]


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
        return self.z[f"{name}__{key}"]

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
