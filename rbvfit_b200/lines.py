"""Atomic line lookup: host-side mirror of rbvfit's ``rb_setline`` for the 'atom' list
(reference: src/rbvfit/rb_setline.py:25-64).

The parity contract includes the reference's dtype choices: rest wavelengths are float64 while the
oscillator strength f and the damping constant gamma are stored as **float32** (:42,44) -- e.g. MgII 2796
has f = 0.6122999787330627, not 0.6123.  The table itself (rbvfit_b200/data/atomic_lines.json) is the
reference's lines/atom_full.dat converted verbatim by tools/make_line_table.py.
"""
from __future__ import annotations

import json
import os
from functools import lru_cache

import numpy as np

_TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "atomic_lines.json")


@lru_cache(maxsize=1)
def _table():
    with open(_TABLE) as fh:
        rows = json.load(fh)["lines"]
    n = len(rows)
    wave = np.zeros(n, dtype=np.float64)
    fval = np.zeros(n, dtype=np.float32)
    gamma = np.zeros(n, dtype=np.float32)
    name = np.empty(n, dtype=object)
    for i, (ion, w, f, g) in enumerate(rows):
        wave[i] = float(w)
        fval[i] = float(f)
        gamma[i] = float(g)
        name[i] = f"{ion} {int(float(w))}"
    return wave, fval, gamma, name


def rb_setline(lambda_rest: float, method: str = "closest", linelist: str = "atom") -> dict:
    """Same contract as the reference: dict of length-1 arrays 'wave', 'fval', 'name', 'gamma'."""
    if linelist != "atom":
        raise ValueError("rbvfit_b200 ships the 'atom' line list only (the only one the hot path reads)")
    wave, fval, gamma, name = _table()
    if method == "Exact":
        idx = np.where(np.abs(lambda_rest - wave) < 1e-3)
    elif method == "closest":
        idx = np.array([np.abs(lambda_rest - wave).argmin()])
    else:
        raise ValueError("Specify a valid matching method: 'closest' or 'Exact'")
    return {"wave": wave[idx], "fval": fval[idx], "name": name[idx], "gamma": gamma[idx]}


def ion_of(line_name: str) -> str:
    """'MgII 2796' -> 'MgII' (IonGroup._extract_ion_name, core/fit_configuration.py:135-153)."""
    return line_name.split()[0] if line_name else "Unknown"
