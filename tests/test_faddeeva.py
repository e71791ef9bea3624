"""Device Voigt-Hjerting function H(a,x) = Re w(x+ia) against the known-answer lattice
(tests/golden/wofz_lattice.npz: scipy.special.wofz -- the reference's call -- and mpmath at 40 digits)."""
import os

import numpy as np
import pytest

from golden_util import GOLDEN_DIR


def _lattice():
    z = np.load(os.path.join(GOLDEN_DIR, "wofz_lattice.npz"))
    return z["x"], z["a"], z["scipy"], z["mpmath"]


@pytest.mark.gpu
def test_device_H_vs_lattice():
    from rbvfit_b200.engine import Engine
    x, a, ref_scipy, ref_mp = _lattice()
    eng = Engine()
    got = eng.voigt_h(x, a)
    rel_mp = np.abs(got - ref_mp) / ref_mp
    rel_sc = np.abs(got - ref_scipy) / ref_scipy
    print("max rel err vs mpmath", rel_mp.max(), "vs scipy", rel_sc.max())
    assert rel_mp.max() <= 5e-13      # stated accuracy of the device function
    assert rel_sc.max() <= 5e-13


@pytest.mark.gpu
def test_device_H_dense_sweep_vs_scipy():
    from scipy.special import wofz
    from rbvfit_b200.engine import Engine
    rng = np.random.default_rng(11)
    n = 200_000
    x = np.concatenate([rng.uniform(-9, 9, n // 2), rng.choice([-1, 1], n // 2) * np.exp(rng.uniform(np.log(7), np.log(5e4), n // 2))])
    a = np.exp(rng.uniform(np.log(1e-8), np.log(30.0), n))
    ref = wofz(x + 1j * a).real
    got = Engine().voigt_h(x, a)
    rel = np.abs(got - ref) / ref
    i = rel.argmax()
    print("dense sweep max rel", rel.max(), "at x,a =", x[i], a[i])
    assert rel.max() <= 1e-12


@pytest.mark.gpu
def test_device_H_tepper_garcia_matches_reference_formula():
    from oracle import voigt_oracle as vo
    from rbvfit_b200.engine import Engine
    rng = np.random.default_rng(5)
    x = rng.uniform(-40, 40, 50_000)
    a = np.exp(rng.uniform(np.log(1e-6), np.log(0.05), 50_000))
    ref = vo.H_tepper_garcia(x, a)
    got = Engine().voigt_h(x, a, method="fast")
    assert np.max(np.abs(got - ref)) <= 1e-14


@pytest.mark.gpu
def test_device_reciprocal_is_full_precision():
    from rbvfit_b200.engine import Engine
    err = Engine().selftest_rcp()
    print("rcp_pos max rel err", err)
    assert err <= 4e-16


def test_tables_are_up_to_date():
    """The committed header equals what tools/gen_faddeeva_tables.py generates (CPU, ~1 s)."""
    import importlib.util
    root = os.path.normpath(os.path.join(os.path.dirname(__file__), ".."))
    spec = importlib.util.spec_from_file_location("gen", os.path.join(root, "tools", "gen_faddeeva_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text = gen.render_header()
    assert text == open(os.path.join(root, "rbvfit_b200", "csrc", "faddeeva_tables.h")).read()


def test_numpy_emulation_of_device_algorithm():
    """CPU emulation of the device tiers (same tables, same formulas) against the lattice: guards the tables
    and thresholds without a GPU."""
    import importlib.util
    import math
    root = os.path.normpath(os.path.join(os.path.dirname(__file__), ".."))
    spec = importlib.util.spec_from_file_location("gen", os.path.join(root, "tools", "gen_faddeeva_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    tabs, ctab = gen.build_core_tables(), gen.build_ctab()
    x, a, ref_scipy, ref_mp = _lattice()
    x, a, ref = x.ravel(), a.ravel(), ref_mp.ravel()
    a2 = a * a
    d = x * x + a2
    out = np.zeros_like(x)
    # asymptotic tiers
    q = np.zeros((gen.ASYM_PMAX + 1, x.size))
    for p in range(1, gen.ASYM_PMAX + 1):
        for m in range(gen.ASYM_MMAX, -1, -1):
            q[p] = q[p] * a2 + ctab[p, m]
    rho = 1.0 / d
    for lo, hi, nq in ((64.0, 576.0, 13), (576.0, 4e4, 6), (4e4, np.inf, 3)):
        sel = (d >= lo) & (d < hi) & (a <= 1.0)
        s = np.zeros(x.size)
        for p in range(nq, 0, -1):
            s = (s + q[p]) * rho
        out[sel] = (a / math.sqrt(math.pi) * s)[sel]
    # core, table path
    sel = (d < 64.0) & (a <= 0.05)
    ax = np.abs(x[sel])
    j = np.minimum((ax * 4).astype(int), 31)
    t = ax * 8 - (2 * j + 1)
    G = np.zeros(ax.size)
    for k in range(3, -1, -1):
        tab = tabs[k]
        g = tab[-1, j]
        for dd in range(tab.shape[0] - 2, -1, -1):
            g = g * t + tab[dd, j]
        G = G * a2[sel] + g
    y = (a[sel] * x[sel]) ** 2
    c = np.zeros(ax.size)
    for k in range(5, -1, -1):
        c = c * y + (-4.0) ** k / math.factorial(2 * k)
    out[sel] = np.exp(a2[sel] - x[sel] ** 2) * c + a[sel] * G
    # core, Weideman path
    sel = (d < 64.0) & (a > 0.05)
    L, wc = gen.build_weideman()
    z = x[sel] + 1j * a[sel]
    Z = (L + 1j * z) / (L - 1j * z)
    out[sel] = (2 * np.polyval(wc, Z) / (L - 1j * z) ** 2 + (1 / math.sqrt(math.pi)) / (L - 1j * z)).real
    # a > 1: complex asymptotic series, 13 terms
    sel = (d >= 64.0) & (a > 1.0)
    zz = x[sel] + 1j * a[sel]
    r = 1.0 / zz
    s2 = r * r
    S = np.zeros_like(zz)
    ck = [1.0]
    for k in range(1, 13):
        ck.append(ck[-1] * (2 * k - 1) * 0.5)
    for k in range(12, -1, -1):
        S = S * s2 + ck[k]
    out[sel] = (1j / math.sqrt(math.pi) * r * S).real
    rel = np.abs(out - ref) / ref
    assert rel.max() <= 5e-13, rel.max()
