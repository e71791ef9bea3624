"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the device-resident stretch-move sampler
(rbvfit_b200/csrc/rbv_sampler.cuh, C ABI rbv_stretch_run), random streams included.

The algorithm is emcee's RedBlueMove + StretchMove (Goodman & Weare 2010; Foreman-Mackey et al. 2013), the sampler
the reference builds in vfit_mcmc.py:408-423; emcee itself is not vendored in the reference and not installed here.
What is specific to the device version and restated here bit for bit:

  * Philox4x32-10 keyed by the 64-bit seed, counter = (step lo, step hi, walker, purpose);
  * uniform (0,1) doubles from the top 53 bits of (x, y): (m + 0.5) * 2^-53;
  * the step's split: positions 0..W-1 -> walkers (a pos + b) mod W with (a, b) from purpose 0 / walker 0xffffffff,
    a advanced until gcd(a, W) = 1; the first ceil(W/2) positions are half 0;
  * half s: u and the partner index from purpose 1 + s (x, y -> u; z -> partner), the accept draw from purpose 3 + s;
  * survey mode (rbv_stretch_run_sightlines): ensemble e of W walkers uses the same streams with the walker counter
    offset by e W (``walker_offset``); the step's split is shared by all ensembles.

Driving this replica with the GPU's lnprob must reproduce the device chain exactly (tests/test_gpu_vfit.py); driving
it with an analytic Gaussian checks the algorithm itself on the CPU (tests/test_host_logic.py).
"""
from __future__ import annotations

from math import gcd

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def u01(hi, lo):
    m = ((hi << 32) | lo) >> 11
    return (float(m) + 0.5) * 1.1102230246251565e-16


def _rand(seed, step, walker, purpose):
    seed, step, walker, purpose = int(seed), int(step), int(walker), int(purpose)     # Python ints: no overflow
    return philox4x32_10((step & MASK, (step >> 32) & MASK, walker & MASK, purpose), (seed & MASK, (seed >> 32) & MASK))


def step_perm(seed, step, W):
    r = _rand(seed, step, 0xFFFFFFFF, 0)
    a, b = r[0] % W, r[1] % W
    while gcd(a, W) != 1 and W != 1:
        a = (a + 1) % W
    return a, b


def split_geometry(W, split):
    h = (W + 1) // 2
    return (0, h, h, W - h) if split == 0 else (h, W - h, 0, h)


def propose_half(coords, seed, step, split, a_scale=2.0, walker_offset=0):
    """Proposals of half `split` of step `step`: (walker index per row, proposals [nS, ndim], (ndim-1) ln z).
    ``walker_offset`` = e * W for ensemble e of a survey run (rbv_stretch_run_sightlines): the random streams are
    those of a single ensemble with the walker counter shifted."""
    W, ndim = coords.shape
    pa, pb = step_perm(seed, step, W)
    walker = [(pa * pos + pb) % W for pos in range(W)]
    offS, nS, offC, nC = split_geometry(W, split)
    idx = np.array([walker[offS + k] for k in range(nS)])
    q = np.empty((nS, ndim))
    fac = np.empty(nS)
    for k, i in enumerate(idx):
        r = _rand(seed, step, walker_offset + i, 1 + split)
        u = u01(r[0], r[1])
        j = walker[offC + r[2] % nC]
        t = (a_scale - 1.0) * u + 1.0
        zz = t * t / a_scale
        fac[k] = (ndim - 1.0) * np.log(zz)
        q[k] = coords[j] - (coords[j] - coords[i]) * zz
    return idx, q, fac


def accept_half(coords, lnp, nacc, seed, step, split, idx, q, fac, new, walker_offset=0):
    """Accept / reject in place; same draws as rbv_stretch_accept."""
    for k, i in enumerate(idx):
        r = _rand(seed, step, walker_offset + i, 3 + split)
        with np.errstate(invalid="ignore"):
            if np.log(u01(r[0], r[1])) < fac[k] + new[k] - lnp[i]:
                coords[i] = q[k]
                lnp[i] = new[k]
                nacc[i] += 1


def run(lnprob_fn, coords, lnp, nsteps, seed, a_scale=2.0, first_step=0, walker_offset=0):
    """Returns (chain [nsteps, W, ndim], lnp_chain [nsteps, W], n_accepted [W]); ``lnprob_fn`` maps
    (n, ndim) -> (n,)."""
    coords = np.array(coords, dtype=np.float64, copy=True)
    lnp = np.array(lnp, dtype=np.float64, copy=True)
    W, ndim = coords.shape
    chain = np.empty((nsteps, W, ndim))
    lps = np.empty((nsteps, W))
    nacc = np.zeros(W, dtype=np.int64)
    for s in range(nsteps):
        step = first_step + s
        for split in (0, 1):
            idx, q, fac = propose_half(coords, seed, step, split, a_scale, walker_offset)
            new = np.asarray(lnprob_fn(q), dtype=np.float64)
            accept_half(coords, lnp, nacc, seed, step, split, idx, q, fac, new, walker_offset)
        chain[s], lps[s] = coords, lnp
    return chain, lps, nacc
