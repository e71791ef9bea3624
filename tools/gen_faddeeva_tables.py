#!/usr/bin/env python
"""Generate rbvfit_b200/csrc/faddeeva_tables.h -- the constant tables behind the
device Voigt-Hjerting function H(a,x) = Re w(x + i a).

Three ingredients (see DESIGN.md "Faddeeva"):

1. CORE (|z|^2 < 64, a <= A_FAST): the Taylor series of H in the damping parameter,
       H(a,x) = exp(a^2 - x^2) cos(2 a x) + a * sum_k a^(2k) g_k(x),
       g_k(x) = (-1)^(k+1) Im w^(2k+1)(x) / (2k+1)!,
   with g_0..g_3 tabulated as piecewise polynomials on 32 intervals of width 1/4 in |x|
   (Chebyshev interpolants converted to monomials in the local coordinate t in [-1,1]).
   The g_k are generated in 60-digit arithmetic from  V_0 = exp(-x^2) erfi(x),
   V_1 = -2 x V_0 + 2/sqrt(pi),  V_{n+1} = -2 x V_n - 2 n V_{n-1}.

2. ASYMPTOTIC (|z|^2 >= 64, any a): the truncated expansion
       w(z) ~ i/(sqrt(pi) z) sum_k (2k-1)!!/(2 z^2)^k
   rewritten with rho = 1/(x^2 + a^2) as  H = (a/sqrt(pi)) sum_p q_p(a^2) rho^p,
       q_p(a^2) = sum_m CTAB[p][m] a^(2m),
       CTAB[p][m] = c_k * [s^m] V_k(s),  k = p-1-m,  c_k = (2k-1)!!/2^k,
       V_k(s) = sum_j (-1)^j C(2k+1,2j+1) (1-s)^(k-j) s^j      (= sin((2k+1)phi)/sin(phi), s = sin^2 phi).
   Exact rational arithmetic; p = 1..13 (k <= 12).

3. GENERAL-a CORE (|z|^2 < 64, a > A_FAST): Weideman (1994) N = 40 rational approximation,
   coefficients from the FFT recipe of the paper.

4. FAR FIELD: m = 8 Chebyshev nodes t_k = cos(pi (2k+1)/(2m)) on [-1,1] and the m x m matrix that maps
   the values of a function at those nodes to the MONOMIAL coefficients of its interpolating polynomial
   (discrete Chebyshev transform followed by the T_j -> monomial change of basis), in 60-digit arithmetic.
   The tile kernel interpolates the summed far-wing optical depth of a 256-pixel chunk with it (DESIGN.md
   section 4c).

This is build tooling, not test infrastructure and not product code; the generated header is
committed so that building needs only nvcc.
"""
from __future__ import annotations

import math
import os
import sys
from fractions import Fraction
from math import comb

import numpy as np

CORE_XMAX = 8.0
CORE_H = 0.25
CORE_NINT = 32
CORE_DEGS = (10, 8, 6, 4)
ASYM_KMAX = 12            # k = 0..12  -> p = 1..13
ASYM_PMAX = ASYM_KMAX + 1
ASYM_MMAX = 6
WEID_N = 40
FF_M = 8                  # far-field interpolation nodes per chunk


# ----------------------------------------------------------------------------- core tables
def _V_seq(x, nmax, mp):
    sp = mp.sqrt(mp.pi)
    V0 = mp.exp(-x * x) * mp.erfi(x)
    V = [V0, -2 * x * V0 + 2 / sp]
    for n in range(1, nmax):
        V.append(-2 * x * V[n] - 2 * n * V[n - 1])
    return V


def g_funcs(x, kmax, mp):
    V = _V_seq(mp.mpf(x), 2 * kmax + 1, mp)
    return [(-1) ** (k + 1) * V[2 * k + 1] / mp.factorial(2 * k + 1) for k in range(kmax + 1)]


def _cheb_monomials(n, mp):
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for k in range(2, n):
        a = [mp.mpf(0)] + [2 * v for v in T[k - 1]]
        b = T[k - 2] + [mp.mpf(0)] * (len(a) - len(T[k - 2]))
        T.append([a[i] - b[i] for i in range(len(a))])
    return T


def _fit_interval(fvals_at, lo, hi, deg, mp):
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * i + 1) / (2 * n)) for i in range(n)]
    fv = [fvals_at((lo + hi) / 2 + (hi - lo) / 2 * t) for t in nodes]
    c = [sum(fv[i] * mp.cos(mp.pi * j * (2 * i + 1) / (2 * n)) for i in range(n)) * 2 / n for j in range(n)]
    c[0] /= 2
    T = _cheb_monomials(n, mp)
    mono = [mp.mpf(0)] * n
    for j in range(n):
        for i, v in enumerate(T[j]):
            mono[i] += c[j] * v
    return [float(v) for v in mono]


def build_core_tables():
    import mpmath as mp
    mp.mp.dps = 60
    kmax = len(CORE_DEGS) - 1
    tabs = []
    for k, deg in enumerate(CORE_DEGS):
        tab = np.zeros((deg + 1, CORE_NINT))
        for j in range(CORE_NINT):
            lo, hi = mp.mpf(j) * CORE_H, mp.mpf(j + 1) * CORE_H
            tab[:, j] = _fit_interval(lambda x, k=k: g_funcs(x, kmax, mp)[k], lo, hi, deg, mp)
        tabs.append(tab)
    return tabs


# ----------------------------------------------------------------------------- asymptotic table
def _Vcoef(k, mmax):
    out = [0] * (mmax + 1)
    for j in range(k + 1):
        cj = (-1) ** j * comb(2 * k + 1, 2 * j + 1)
        for i in range(k - j + 1):
            m = i + j
            if m <= mmax:
                out[m] += cj * comb(k - j, i) * (-1) ** i
    return out


def _dfact_odd(k):  # (2k-1)!!
    r = 1
    for i in range(1, 2 * k, 2):
        r *= i
    return r


def build_ctab():
    C = np.zeros((ASYM_PMAX + 1, ASYM_MMAX + 1))
    for p in range(1, ASYM_PMAX + 1):
        for m in range(ASYM_MMAX + 1):
            k = p - 1 - m
            if k < 0 or m > k or k > ASYM_KMAX:
                continue
            ck = Fraction(_dfact_odd(k), 2 ** k)
            C[p, m] = float(ck * _Vcoef(k, ASYM_MMAX)[m])
    return C


# ----------------------------------------------------------------------------- Weideman
def build_weideman(N=WEID_N):
    M = 2 * N
    M2 = 2 * M
    k = np.arange(-M + 1, M)
    L = math.sqrt(N / math.sqrt(2.0))
    theta = k * np.pi / M
    t = L * np.tan(theta / 2)
    f = np.exp(-t ** 2) * (L ** 2 + t ** 2)
    f = np.concatenate([[0.0], f])
    a = np.real(np.fft.fft(np.fft.fftshift(f))) / M2
    a = np.flipud(a[1:N + 1])          # highest power first (polyval order)
    return L, a


# ----------------------------------------------------------------------------- far field
def build_farfield(m=FF_M):
    """(nodes[m], MINV[m][m]) with  coef_j = sum_k MINV[j][k] f(t_k)  (monomial coefficients, j = power)."""
    import mpmath as mp
    mp.mp.dps = 60
    nodes = [mp.cos(mp.pi * (2 * k + 1) / (2 * m)) for k in range(m)]
    T = _cheb_monomials(m, mp)
    minv = [[mp.mpf(0)] * m for _ in range(m)]
    for j in range(m):                      # Chebyshev coefficient c_j = (2/m) sum_k f_k T_j(t_k), c_0 halved
        for k in range(m):
            cjk = mp.cos(mp.pi * j * (2 * k + 1) / (2 * m)) * 2 / m
            if j == 0:
                cjk /= 2
            for i, v in enumerate(T[j]):    # T_j = sum_i v t^i
                minv[i][k] += cjk * v
    return [float(t) for t in nodes], np.array([[float(v) for v in row] for row in minv])


# ----------------------------------------------------------------------------- emit
def _arr(name, values, per_line=4):
    vals = [f"{v:.17g}" for v in np.asarray(values, dtype=np.float64).ravel()]
    lines = []
    for i in range(0, len(vals), per_line):
        lines.append("    " + ", ".join(vals[i:i + per_line]) + ",")
    body = "\n".join(lines)
    return f"static const double {name}[{len(vals)}] = {{\n{body}\n}};\n"


def render_header():
    tabs = build_core_tables()
    ctab = build_ctab()
    L, wa = build_weideman()
    ff_nodes, ff_minv = build_farfield()
    out = []
    out.append("// GENERATED by tools/gen_faddeeva_tables.py -- do not edit by hand.\n")
    out.append("// Tables for the device Voigt-Hjerting function; see the generator's docstring.\n")
    out.append("#pragma once\n\n")
    out.append(f"#define RBV_CORE_NINT {CORE_NINT}\n")
    out.append(f"#define RBV_CORE_INV_H {1.0 / CORE_H:.17g}\n")
    out.append(f"#define RBV_CORE_H {CORE_H:.17g}\n")
    for k, deg in enumerate(CORE_DEGS):
        out.append(f"#define RBV_CORE_DEG{k} {deg}\n")
    offs = np.cumsum([0] + [(d + 1) * CORE_NINT for d in CORE_DEGS])
    for k in range(len(CORE_DEGS)):
        out.append(f"#define RBV_CORE_OFF{k} {int(offs[k])}\n")
    out.append(f"#define RBV_CORE_TABLE_LEN {int(offs[-1])}\n")
    out.append(f"#define RBV_ASYM_PMAX {ASYM_PMAX}\n")
    out.append(f"#define RBV_ASYM_MMAX {ASYM_MMAX}\n")
    out.append(f"#define RBV_WEID_N {WEID_N}\n")
    out.append(f"#define RBV_WEID_L {L:.17g}\n")
    out.append(f"#define RBV_FF_M {FF_M}\n\n")
    out.append("// g_k tables, layout [k][degree][interval] (interval fastest), monomials in t in [-1,1]\n")
    out.append(_arr("RBV_CORE_TABLE_HOST", np.concatenate([t.ravel() for t in tabs])))
    out.append("\n// CTAB[p][m], p = 0..PMAX (row 0 unused), m = 0..MMAX\n")
    out.append(_arr("RBV_ASYM_CTAB_HOST", ctab, per_line=ASYM_MMAX + 1))
    out.append("\n// Weideman N=40 polynomial coefficients, highest power first\n")
    out.append(_arr("RBV_WEID_COEF_HOST", wa))
    out.append("\n// far field: Chebyshev nodes on [-1,1] and the node-values -> monomial-coefficients matrix [power][node]\n")
    out.append(_arr("RBV_FF_NODES_HOST", ff_nodes))
    out.append(_arr("RBV_FF_MINV_HOST", ff_minv, per_line=FF_M))
    return "".join(out)


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    target = os.path.join(here, "..", "rbvfit_b200", "csrc", "faddeeva_tables.h")
    text = render_header()
    if "--check" in sys.argv:
        ok = os.path.exists(target) and open(target).read() == text
        print("up to date" if ok else "STALE")
        sys.exit(0 if ok else 1)
    with open(target, "w") as fh:
        fh.write(text)
    print(f"wrote {os.path.normpath(target)} ({len(text)} bytes)")


if __name__ == "__main__":
    main()
