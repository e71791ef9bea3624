"""Algorithmic work of the hot path, per SURVEY.md section 8(d) -- the figure ``roofline.achieved`` is
computed from (NOT the instructions the kernel happens to execute):

    F per walker.pixel = sum_lines c(x_lp) + 2 K + 32
    c = 80 (|x| < 12: exp(-x^2) + damping correction), 40 (12 <= |x| < 100: asymptotic series, 6-8 terms),
        21 (|x| >= 100: x by one FMA, x^2, reciprocal, 3-term series, scale, accumulate)
    2 K = one FMA per LSF tap; 32 = exp(-tau) (24) + chi^2 term (8).

Evaluated once per workload on the CPU from theta_true (host arithmetic only, no forward model).

With the far-field path (default) the kernel does LESS work than that rule assumes: the far
wings of a 1024-pixel super-chunk are summed at 8 Chebyshev nodes and interpolated.  ``flops_farfield`` counts the
algorithmic work of THAT algorithm (the honest numerator for its roofline); the rule above is kept as the
"direct-evaluation equivalent":

    per super-chunk : 8 nodes x n_ff lines x c (one series evaluation each: 21 at |x| >= 100, 40 below)
                      + 2 x 64 (node values -> coefficients)
    per pixel       : 16 if the super-chunk has a far field (t by one FMA + degree-7 Horner), else 0
                      + sum over the lines NOT in the far field of c(x_lp)           (same tiers as above)
                      + 2 K + 32
    line l is in the far field of a super-chunk when min |z|^2 >= 576 over it (6-term series valid) and
        72 |kappa_l| (hw / (2 xm))^8 / xm^2 <= 1e-12 / L,   kappa = N f K a / sqrt(pi),  xm = min |x|, hw = half width in x
"""
from __future__ import annotations

import numpy as np

C_CORE, C_MID, C_FAR = 80.0, 40.0, 21.0


def x_grid(data, theta, wave):
    """Normalised frequency offsets x[l, p] (voigt_model.py:142-150, 192-204)."""
    theta = np.asarray(theta, dtype=np.float64)
    b = theta[data.b_indices]
    v = theta[data.v_indices]
    lam0 = np.asarray(data.atomic_lambda0, dtype=np.float64)
    z_total = data.z_factors * (1 + v / 299792.458) - 1
    b_f = b / lam0 * 1e13
    nu0 = 2.99792458e18 / lam0
    nu = 2.99792458e18 * (1 + z_total)[:, None] / np.asarray(wave)[None, :]
    return (nu - nu0[:, None]) / b_f[:, None]


def flops_per_walker_pixel(data, theta, wave, n_taps):
    """Returns (F, tier_fractions)."""
    L, P = len(data.atomic_lambda0), len(wave)
    core = mid = 0
    for l0 in range(0, L, 8):          # chunked to bound memory at P = 100k
        sub = type("S", (), {})()
        sub.b_indices, sub.v_indices = data.b_indices[l0:l0 + 8], data.v_indices[l0:l0 + 8]
        sub.atomic_lambda0, sub.z_factors = data.atomic_lambda0[l0:l0 + 8], data.z_factors[l0:l0 + 8]
        ax = np.abs(x_grid(sub, theta, wave))
        core += int(np.count_nonzero(ax < 12.0))
        mid += int(np.count_nonzero((ax >= 12.0) & (ax < 100.0)))
    far = L * P - core - mid
    F = (C_CORE * core + C_MID * mid + C_FAR * far) / P + 2.0 * n_taps + 32.0
    return F, {"core": core / (L * P), "mid": mid / (L * P), "far": far / (L * P)}


SUPER_PIX = 1024
FF_NODES = 8
FF_EPS = 1e-12


def flops_farfield(data, theta, wave, n_taps):
    """Algorithmic flops per walker.pixel of the far-field algorithm (see module docstring); super-chunks are
    taken aligned to pixel 0 (the kernel's are offset by the tile origin and the LSF halo).
    Returns (F, {"farfield": fraction of (line, pixel) pairs served by the interpolant, "core"/"mid"/"far":
    fractions evaluated directly})."""
    theta = np.asarray(theta, dtype=np.float64)
    wave = np.asarray(wave, dtype=np.float64)
    L, P = len(data.atomic_lambda0), len(wave)
    lam0 = np.asarray(data.atomic_lambda0, dtype=np.float64)
    gam = np.asarray(data.atomic_gamma, dtype=np.float64)
    fos = np.asarray(data.atomic_f, dtype=np.float64)
    N = 10.0 ** theta[data.N_indices]
    b = theta[data.b_indices]
    v = theta[data.v_indices]
    b_f = b / lam0 * 1e13
    nu0 = 2.99792458e18 / lam0
    a = gam / (4 * np.pi * b_f)
    kappa = N * fos * (4.48898479507e3 / (nu0 * b)) * a / np.sqrt(np.pi)
    A = 2.99792458e18 * (data.z_factors * (1 + v / 299792.458)) / b_f
    B = nu0 / b_f
    u = 1.0 / wave
    flops = 0.0
    n_ffp = core = mid = far = 0
    for p0 in range(0, P, SUPER_PIX):
        uc = u[p0:p0 + SUPER_PIX]
        x1, x2 = A * uc.min() - B, A * uc.max() - B
        crosses = x1 * x2 < 0
        xm = np.where(crosses, 0.0, np.minimum(np.abs(x1), np.abs(x2)))
        hw = 0.5 * np.abs(x2 - x1)
        with np.errstate(divide="ignore", invalid="ignore"):
            bound = 72.0 * np.abs(kappa) * (hw / (2 * xm)) ** FF_NODES / xm ** 2
        ff = (~crosses) & (xm * xm + a * a >= 576.0) & (a < 1.0) & (bound <= FF_EPS / L)
        n_ff = int(ff.sum())
        npx = len(uc)
        if n_ff:
            node_cost = np.where(xm[ff] >= 100.0, C_FAR, C_MID).sum()
            flops += FF_NODES * node_cost + 2.0 * FF_NODES * FF_NODES + 16.0 * npx
            n_ffp += n_ff * npx
        d = np.flatnonzero(~ff)
        if len(d):
            ax = np.abs(A[d, None] * uc[None, :] - B[d, None])
            c = int(np.count_nonzero(ax < 12.0))
            m = int(np.count_nonzero((ax >= 12.0) & (ax < 100.0)))
            f = ax.size - c - m
            core, mid, far = core + c, mid + m, far + f
            flops += C_CORE * c + C_MID * m + C_FAR * f
    F = flops / P + 2.0 * n_taps + 32.0
    tot = float(L * P)
    return F, {"farfield": n_ffp / tot, "core": core / tot, "mid": mid / tot, "far": far / tot}
