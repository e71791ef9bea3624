"""Algorithmic work of the hot path, per SURVEY.md section 8(d) -- the figure ``roofline.achieved`` is
computed from (NOT the instructions the kernel happens to execute):

    F per walker.pixel = sum_lines c(x_lp) + 2 K + 32
    c = 80 (|x| < 12: exp(-x^2) + damping correction), 40 (12 <= |x| < 100: asymptotic series, 6-8 terms),
        21 (|x| >= 100: x by one FMA, x^2, reciprocal, 3-term series, scale, accumulate)
    2 K = one FMA per LSF tap; 32 = exp(-tau) (24) + chi^2 term (8).

Evaluated once per workload on the CPU from theta_true (host arithmetic only, no forward model).
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
