"""Seeded random problems shared by the fuzz tests: random systems / ions / component counts, redshifts, wavelength
windows and pixel counts, LSFs, column densities up to the damped regime, Doppler widths from 3 to 80 km/s.
``tests/test_gpu_fuzz.py`` checks the CUDA path against the CPU oracle on them (GPU box);
``tests/test_oracle_vs_golden.py`` checks the oracle against the LIVE reference on the same problems (build
container), so the two together tie the CUDA path to the reference beyond the committed fixtures."""
import numpy as np

IONS = {
    "HI": [1215.67, 1025.72, 972.54], "CIV": [1548.2, 1550.77], "SiIV": [1393.76, 1402.77],
    "MgII": [2796.3, 2803.5], "FeII": [2600.17, 2586.65, 2382.77], "SiII": [1526.71, 1304.37, 1260.42],
    "OVI": [1031.93, 1037.62], "AlII": [1670.79],
}


def draw_problem(seed):
    rng = np.random.default_rng(seed)
    systems, rest = [], []
    for _ in range(int(rng.integers(1, 4))):
        z = float(rng.uniform(0.1, 3.0))
        for ion in rng.choice(list(IONS), size=int(rng.integers(1, 4)), replace=False):
            trans = IONS[ion][: int(rng.integers(1, len(IONS[ion]) + 1))]
            systems.append((z, str(ion), list(trans), int(rng.integers(1, 4))))
            rest += [t * (1 + z) for t in trans]
    # one system may not hold the same ion twice: merge duplicates away by construction (distinct z per system)
    centre = float(rng.choice(rest))
    half = float(rng.uniform(8.0, 400.0))
    P = int(rng.integers(300, 6000))
    wave = np.linspace(centre - half, centre + half, P)
    lsf = rng.choice(["none", "2.5", "4.0", "6.5", "custom"])
    taps, fwhm = None, None
    if lsf == "custom":
        x = np.arange(-30, 31)
        taps = np.exp(-0.5 * (x / rng.uniform(2, 6)) ** 2) * (1 + 0.3 * (x > 0))
        taps /= taps.sum()
    elif lsf != "none":
        fwhm = str(lsf)
    C = sum(s[3] for s in systems)
    n = rng.uniform(12.0, 16.0, C)
    for k, s in enumerate(np.repeat(np.arange(len(systems)), [s[3] for s in systems])):
        if systems[s][1] == "HI" and rng.random() < 0.3:
            n[k] = rng.uniform(18.5, 21.5)                     # Lyman-limit / damped systems
    b = rng.uniform(3.0, 80.0, C)
    v = rng.uniform(-200.0, 200.0, C)
    theta = np.concatenate([n, b, v])
    scale = np.concatenate([np.full(C, 0.05), np.full(C, 0.5), np.full(C, 2.0)])
    thetas = theta + scale * rng.standard_normal((6, 3 * C))
    lb = theta - np.concatenate([np.full(C, 2.0), np.minimum(b - 1.0, 40.0), np.full(C, 60.0)])
    ub = theta + np.concatenate([np.full(C, 2.0), np.full(C, 40.0), np.full(C, 60.0)])
    thetas = np.clip(thetas, lb, ub)
    j = int(rng.integers(0, 3 * C))
    thetas[4, j] = ub[j] + 1.0                                   # one row outside the box
    return systems, wave, fwhm, taps, theta, thetas, lb, ub, rng
