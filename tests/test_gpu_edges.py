"""GPU edge cases against the CPU oracle (same seeded inputs, sizes the oracle finishes in seconds): ragged and tiny
spectra, LSF wider than the spectrum, many lines, unsorted wavelength grids, single walkers, every tile geometry and
both chunk sizes.  Tolerances are the north-star ones (flux 1e-10, lnprob 1e-9 relative)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FLUX_TOL = 1e-10
LNPROB_RTOL = 1e-9


def _models(systems, FWHM="6.5", taps=None):
    from oracle import voigt_oracle as vo
    from rbvfit_b200 import FitConfiguration
    from rbvfit_b200.model import GpuVoigtModel
    cfg, ocfg = FitConfiguration(), vo.OracleConfig()
    for (z, ion, trans, comps) in systems:
        cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
        ocfg.add_system(z, ion, trans, comps)
    return GpuVoigtModel(cfg, FWHM=FWHM, lsf_taps=taps), vo.lower(ocfg, FWHM=FWHM, custom_taps=taps)


def _check(model, omodel, wave, thetas, lb, ub, rng, error_dtype=np.float64):
    from oracle import voigt_oracle as vo
    from rbvfit_b200.likelihood import GpuLikelihood
    wave = np.asarray(wave, dtype=np.float64)
    flux = 1.0 + 0.05 * rng.standard_normal(wave.size)
    error = np.full(wave.size, 0.05, dtype=error_dtype)
    like = GpuLikelihood({"S": dict(model=model, wave=wave, flux=flux, error=error)}, lb, ub)
    comp = vo.compile_instruments({"S": dict(model=omodel, wave=wave, flux=flux, error=error)})
    got = like.lnprob(thetas)
    ref = vo.lnprob_batch(comp, np.atleast_2d(thetas), lb, ub)
    got = np.atleast_1d(got)
    assert np.array_equal(np.isneginf(got), np.isneginf(ref))
    fin = np.isfinite(ref)
    if fin.any():
        assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= LNPROB_RTOL
    th0 = np.atleast_2d(thetas)[int(np.flatnonzero(fin)[0])] if fin.any() else np.atleast_2d(thetas)[0]
    gf = model.compile().model_flux(th0, wave)
    assert np.max(np.abs(gf - vo.model_flux(omodel, th0, wave))) <= FLUX_TOL
    like.close()
    return got


MGII = [(0.348, "MgII", [2796.3, 2803.5], 2)]
TH = np.array([14.2, 14.5, 40.0, 30.0, -25.0, 35.0])
LB, UB = TH - np.array([2, 2, 38, 28, 50, 50.0]), TH + np.array([2, 2, 40, 40, 50, 50.0])


@pytest.mark.parametrize("P", [1, 2, 7, 22, 23, 255, 256, 257, 1791, 1793, 2049, 4100])
def test_ragged_and_tiny_spectra(P):
    """Spectra shorter than the 23-tap LSF, one pixel, and sizes around the tile / block boundaries."""
    rng = np.random.default_rng(100 + P)
    model, om = _models(MGII)
    wave = np.linspace(3762.0, 3786.0, P) if P > 1 else np.array([3769.5])
    thetas = np.clip(TH + rng.standard_normal((9, 6)) * [0.05, 0.05, 1, 1, 2, 2], LB, UB)
    thetas[3, 0] = UB[0] + 1.0                                           # one row outside the prior box
    _check(model, om, wave, thetas, LB, UB, rng)


def test_lsf_wider_than_spectrum_and_widest_supported():
    from rbvfit_b200 import lsf
    rng = np.random.default_rng(7)
    taps = lsf.cos_like_taps(321)
    model, om = _models(MGII, FWHM=None, taps=taps)
    thetas = np.clip(TH + rng.standard_normal((5, 6)) * [0.05, 0.05, 1, 1, 2, 2], LB, UB)
    for P in (40, 320, 321, 700):
        _check(model, om, np.linspace(3762.0, 3786.0, P), thetas, LB, UB, rng)
    wide = np.exp(-0.5 * (np.arange(-1500, 1501) / 400.0) ** 2)
    model, om = _models(MGII, FWHM=None, taps=wide / wide.sum())
    _check(model, om, np.linspace(3740.0, 3800.0, 5000), thetas[:3], LB, UB, rng)
    widest = np.exp(-0.5 * (np.arange(-4096, 4097) / 900.0) ** 2)              # 8193 taps: the documented maximum
    model, om = _models(MGII, FWHM=None, taps=widest / widest.sum())
    _check(model, om, np.linspace(3740.0, 3800.0, 20000), thetas[:2], LB, UB, rng)
    from rbvfit_b200._lib import RbvError
    too_wide = np.ones(8195) / 8195.0
    model, om = _models(MGII, FWHM=None, taps=too_wide)
    with pytest.raises(RbvError):
        model.compile().model_flux(TH, np.linspace(3740.0, 3800.0, 9000))


def test_unsorted_and_descending_wavelength_grids():
    """The kernel takes the range of 1/lambda from block min/max tables: no monotonicity is assumed.  (The LSF acts
    on pixel order, as in the reference, so the oracle sees the same arrays.)"""
    rng = np.random.default_rng(11)
    model, om = _models(MGII)
    thetas = np.clip(TH + rng.standard_normal((6, 6)) * [0.05, 0.05, 1, 1, 2, 2], LB, UB)
    wave = np.linspace(3755.0, 3795.0, 3000)
    _check(model, om, wave[::-1].copy(), thetas, LB, UB, rng)
    _check(model, om, rng.permutation(wave), thetas, LB, UB, rng)


def test_many_lines_many_components():
    """L = 96 lines over 32 tied components (the per-warp line lists and classification run several rounds)."""
    rng = np.random.default_rng(13)
    systems = []
    for k, z in enumerate(np.linspace(1.9, 2.9, 8)):
        systems.append((float(z), "CIV", [1548.2, 1550.77], 2))
        systems.append((float(z), "HI", [1215.67, 1025.72, 972.54, 949.74], 2))
    model, om = _models(systems)
    C = om.total_components
    assert om.n_lines == 96 and C == 32
    n, b, v = rng.uniform(12.5, 14.5, C), rng.uniform(8, 45, C), rng.uniform(-120, 120, C)
    th = np.concatenate([n, b, v])
    lb, ub = th - np.concatenate([np.full(C, 2.0), np.full(C, 6.0), np.full(C, 50.0)]), th + 50.0
    thetas = th + rng.standard_normal((7, 3 * C)) * np.concatenate([np.full(C, 0.05), np.full(C, 1.0), np.full(C, 2.0)])
    _check(model, om, np.linspace(3400.0, 6100.0, 9000), thetas, lb, ub, rng)


def test_single_walker_scalar_and_float32_errors():
    rng = np.random.default_rng(17)
    model, om = _models(MGII)
    wave = np.linspace(3755.0, 3795.0, 2048)
    got = _check(model, om, wave, TH, LB, UB, rng, error_dtype=np.float32)       # (ndim,) -> scalar
    assert got.shape == (1,)


@pytest.mark.parametrize("level", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("ppt", [2, 8])
def test_every_tile_geometry_and_chunk_size(level, ppt):
    """RBVFIT_B200_GEOM / RBVFIT_B200_PPT pin the tile size and the phase-1 chunk size (read when a context is
    created): every combination must give the oracle's numbers on a spectrum with several tiles."""
    from oracle import voigt_oracle as vo
    from rbvfit_b200 import workloads as wl
    rng = np.random.default_rng(19)
    w = wl.get_workload("C2")
    model, om = _models(w["systems"])
    wave = np.linspace(3300.0, 5700.0, 6000)
    thetas = wl.make_ensemble(w, 12)
    os.environ["RBVFIT_B200_GEOM"], os.environ["RBVFIT_B200_PPT"] = str(level), str(ppt)
    try:
        _check(model, om, wave, thetas, w["lb"], w["ub"], rng)
    finally:
        del os.environ["RBVFIT_B200_GEOM"], os.environ["RBVFIT_B200_PPT"]
        from rbvfit_b200.engine import Engine
        Engine(0).close()                      # a context created without the variables resets the hooks


def test_joint_fit_with_different_line_lists_per_instrument():
    """Joint fit where the instruments see different transitions of the same ion (different L, same theta layout),
    different pixel counts and different LSFs: summed lnprob against the oracle."""
    from oracle import voigt_oracle as vo
    from rbvfit_b200.likelihood import GpuLikelihood
    rng = np.random.default_rng(23)
    m_a, o_a = _models([(0.348, "MgII", [2796.3], 2)], FWHM="6.5")
    m_b, o_b = _models([(0.348, "MgII", [2796.3, 2803.5], 2)], FWHM="3.0")
    assert o_a.n_lines == 2 and o_b.n_lines == 4
    wa, wb = np.linspace(3760.0, 3776.0, 700), np.linspace(3755.0, 3795.0, 5000)
    data = {}
    for name, (gm, om, wave) in {"A": (m_a, o_a, wa), "B": (m_b, o_b, wb)}.items():
        flux = vo.model_flux(om, TH, wave) + 0.05 * rng.standard_normal(wave.size)
        data[name] = dict(g=gm, o=om, wave=wave, flux=flux, error=np.full(wave.size, 0.05))
    like = GpuLikelihood({n: dict(model=d["g"], wave=d["wave"], flux=d["flux"], error=d["error"])
                          for n, d in data.items()}, LB, UB)
    comp = vo.compile_instruments({n: dict(model=d["o"], wave=d["wave"], flux=d["flux"], error=d["error"])
                                   for n, d in data.items()})
    thetas = np.clip(TH + rng.standard_normal((11, 6)) * [0.05, 0.05, 1, 1, 2, 2], LB, UB)
    thetas[5, 2] = LB[2] - 1.0
    got, ref = like.lnprob(thetas), vo.lnprob_batch(comp, thetas, LB, UB)
    assert np.array_equal(np.isneginf(got), np.isneginf(ref)) and np.isneginf(ref).sum() == 1
    fin = np.isfinite(ref)
    assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= LNPROB_RTOL
    like.close()


def test_no_writes_outside_caller_buffers():
    """Canary check through the raw C ABI: lnprob, workspace (exactly rbv_workspace_bytes), model flux, sampler state
    and chain buffers sit inside larger allocations filled with a sentinel; nothing outside the documented extents
    may change.  (compute-sanitizer is not available on the GPU pool.)"""
    import ctypes as C
    import torch
    from oracle import voigt_oracle as vo
    from rbvfit_b200 import workloads as wl
    from rbvfit_b200._lib import check
    from rbvfit_b200.likelihood import GpuLikelihood
    rng = np.random.default_rng(29)
    w = wl.get_workload("C2")
    model, om = _models(w["systems"])
    wave = np.linspace(3300.0, 5700.0, 5003)
    flux = 1.0 + 0.05 * rng.standard_normal(wave.size)
    like = GpuLikelihood({"S": dict(model=model, wave=wave, flux=flux, error=np.full(wave.size, 0.05))}, w["lb"], w["ub"])
    eng, lib = like.engine, like.engine.lib
    SENT, PAD = -12345.678, 4096
    for W in (1, 7, 130):
        thetas = wl.make_ensemble(w, max(W, 4))[:W]
        th = torch.as_tensor(thetas, device="cuda")
        nbytes = C.c_size_t(0)
        check(lib.rbv_workspace_bytes(eng._h, W, C.byref(nbytes)))
        nws = (int(nbytes.value) + 7) // 8
        big_ws = torch.full((nws + 2 * PAD,), SENT, dtype=torch.float64, device="cuda")
        big_out = torch.full((W + 2 * PAD,), SENT, dtype=torch.float64, device="cuda")
        check(lib.rbv_lnprob_batch(eng._h, th.data_ptr(), W, big_out[PAD:].data_ptr(), big_ws[PAD:].data_ptr(),
                                   int(nbytes.value), None), "rbv_lnprob_batch")
        torch.cuda.synchronize()
        assert torch.all(big_out[:PAD] == SENT) and torch.all(big_out[PAD + W:] == SENT)
        assert torch.all(big_ws[:PAD] == SENT) and torch.all(big_ws[PAD + nws:] == SENT)
        assert np.array_equal(big_out[PAD:PAD + W].cpu().numpy(), like.lnprob(thetas), equal_nan=True)
        P = wave.size
        big_flux = torch.full((W * P + 2 * PAD,), SENT, dtype=torch.float64, device="cuda")
        check(lib.rbv_model_flux_batch(eng._h, 0, th.data_ptr(), W, 1, big_flux[PAD:].data_ptr(), None, 0, None))
        torch.cuda.synchronize()
        assert torch.all(big_flux[:PAD] == SENT) and torch.all(big_flux[PAD + W * P:] == SENT)
        assert torch.all(torch.isfinite(big_flux[PAD:PAD + W * P]))
        # the same call with a workspace (line constants once per walker): same flux, workspace extent respected
        check(lib.rbv_flux_workspace_bytes(eng._h, 0, W, C.byref(nbytes)))
        nfw = (int(nbytes.value) + 7) // 8
        big_fw = torch.full((nfw + 2 * PAD,), SENT, dtype=torch.float64, device="cuda")
        big_flux2 = torch.full((W * P + 2 * PAD,), SENT, dtype=torch.float64, device="cuda")
        for conv in (1, 0):
            check(lib.rbv_model_flux_batch(eng._h, 0, th.data_ptr(), W, conv, big_flux2[PAD:].data_ptr(),
                                           big_fw[PAD:].data_ptr(), int(nbytes.value), None))
            torch.cuda.synchronize()
            assert torch.all(big_fw[:PAD] == SENT) and torch.all(big_fw[PAD + nfw:] == SENT)
            assert torch.all(big_flux2[:PAD] == SENT) and torch.all(big_flux2[PAD + W * P:] == SENT)
            if conv:
                assert torch.equal(big_flux2, big_flux)
        # convolve = 0 in-kernel == an instrument added without taps (what the reference's evaluate() computes)
        nrow = min(W, 2)
        unc = np.array([vo.model_flux(om, t, wave, convolve=False) for t in thetas[:nrow]])
        assert np.max(np.abs(big_flux2[PAD:PAD + nrow * P].cpu().numpy().reshape(nrow, P) - unc)) <= FLUX_TOL
    # sampler: coords / lnprob / chain / counters with canaries
    W, nd, nsteps = 22, like.ndim, 6
    ok = wl.make_ensemble(w, 60)
    ok = ok[np.all((ok >= w["lb"]) & (ok <= w["ub"]), axis=1)][:W]
    bufs = {}
    for name, n in (("coords", W * nd), ("lnp", W), ("chain", nsteps * W * nd), ("lps", nsteps * W)):
        bufs[name] = torch.full((n + 2 * PAD,), SENT, dtype=torch.float64, device="cuda")
    bufs["coords"][PAD:PAD + W * nd] = torch.as_tensor(ok.ravel(), device="cuda")
    bufs["lnp"][PAD:PAD + W] = torch.as_tensor(like.lnprob(ok), device="cuda")
    nacc = torch.full((W + 2 * PAD,), -7, dtype=torch.int32, device="cuda")
    nacc[PAD:PAD + W] = 0
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    check(lib.rbv_stretch_workspace_bytes(eng._h, W, C.byref(nbytes)))
    nws = (int(nbytes.value) + 7) // 8
    big_ws = torch.full((nws + 2 * PAD,), SENT, dtype=torch.float64, device="cuda")
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    check(lib.rbv_stretch_run(eng._h, bufs["coords"][PAD:].data_ptr(), bufs["lnp"][PAD:].data_ptr(), W, nsteps, 2.0,
                              99, 0, bufs["chain"][PAD:].data_ptr(), bufs["lps"][PAD:].data_ptr(),
                              nacc[PAD:].data_ptr(), flag.data_ptr(), big_ws[PAD:].data_ptr(), int(nbytes.value), 1,
                              st.cuda_stream), "rbv_stretch_run")
    torch.cuda.synchronize()
    for name, n in (("coords", W * nd), ("lnp", W), ("chain", nsteps * W * nd), ("lps", nsteps * W)):
        b = bufs[name]
        assert torch.all(b[:PAD] == SENT) and torch.all(b[PAD + n:] == SENT), name
        assert torch.all(torch.isfinite(b[PAD:PAD + n])) and not torch.any(b[PAD:PAD + n] == SENT), name
    assert torch.all(nacc[:PAD] == -7) and torch.all(nacc[PAD + W:] == -7) and int(nacc[PAD:PAD + W].sum()) > 0
    assert torch.all(big_ws[:PAD] == SENT) and torch.all(big_ws[PAD + nws:] == SENT)
    # slice sampler (odd ensemble: halves of 12 and 11 walkers, two candidate rows per walker)
    from rbvfit_b200._lib import RbvSliceTuning
    W = 23
    ok = wl.make_ensemble(w, 60)
    ok = ok[np.all((ok >= w["lb"]) & (ok <= w["ub"]), axis=1)][:W]
    for name, n in (("coords", W * nd), ("lnp", W), ("chain", nsteps * W * nd), ("lps", nsteps * W)):
        bufs[name] = torch.full((n + 2 * PAD,), SENT, dtype=torch.float64, device="cuda")
    bufs["coords"][PAD:PAD + W * nd] = torch.as_tensor(ok.ravel(), device="cuda")
    bufs["lnp"][PAD:PAD + W] = torch.as_tensor(like.lnprob(ok), device="cuda")
    check(lib.rbv_slice_workspace_bytes(eng._h, W, C.byref(nbytes)))
    nws = (int(nbytes.value) + 7) // 8
    big_ws = torch.full((nws + 2 * PAD,), SENT, dtype=torch.float64, device="cuda")
    tuning = RbvSliceTuning(mu=1.0, tolerance=0.05, tune=1, good=0, patience=5, maxsteps=10000, maxiter=10000)
    mus_t = torch.full((nsteps + 2,), SENT, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    check(lib.rbv_slice_run(eng._h, bufs["coords"][PAD:].data_ptr(), bufs["lnp"][PAD:].data_ptr(), W, nsteps,
                            C.byref(tuning), 99, 0, bufs["chain"][PAD:].data_ptr(), bufs["lps"][PAD:].data_ptr(),
                            mus_t[1:].data_ptr(), flag.data_ptr(), big_ws[PAD:].data_ptr(),
                            int(nbytes.value), 1, st.cuda_stream), "rbv_slice_run")
    torch.cuda.synchronize()
    mus = mus_t.cpu().numpy()
    for name, n in (("coords", W * nd), ("lnp", W), ("chain", nsteps * W * nd), ("lps", nsteps * W)):
        b = bufs[name]
        assert torch.all(b[:PAD] == SENT) and torch.all(b[PAD + n:] == SENT), name
        assert torch.all(torch.isfinite(b[PAD:PAD + n])) and not torch.any(b[PAD:PAD + n] == SENT), name
    assert torch.all(big_ws[:PAD] == SENT) and torch.all(big_ws[PAD + nws:] == SENT)
    assert mus[0] == SENT and mus[-1] == SENT and np.all(mus[1:-1] > 0) and int(flag.item()) == 0
    # (at least one widening and one shrinking launch per half-step: two logical iterations per launch by default)
    assert tuning.n_calls >= 3 * W * nsteps and tuning.n_batches >= 2 * 2 * nsteps and tuning.mu == mus[-2]
    # argument checks of the entry point
    assert lib.rbv_slice_run(eng._h, bufs["coords"][PAD:].data_ptr(), bufs["lnp"][PAD:].data_ptr(), 3, nsteps,
                             C.byref(tuning), 99, 0, None, None, None, flag.data_ptr(), big_ws[PAD:].data_ptr(),
                             int(nbytes.value), 1, st.cuda_stream) == 1                    # RBV_EINVAL: W < 4
    assert lib.rbv_slice_run(eng._h, bufs["coords"][PAD:].data_ptr(), bufs["lnp"][PAD:].data_ptr(), W, nsteps,
                             C.byref(tuning), 99, 0, None, None, None, flag.data_ptr(), big_ws[PAD:].data_ptr(),
                             int(nbytes.value) - 256, 1, st.cuda_stream) == 3              # RBV_ENOMEM
    # a half-step that needs more than maxiter iterations ends the run with RBV_ESTATE in both loop modes
    for use_graph in (1, 0):
        short = RbvSliceTuning(mu=1.0, tolerance=0.05, tune=1, good=0, patience=5, maxsteps=10000, maxiter=1)
        assert lib.rbv_slice_run(eng._h, bufs["coords"][PAD:].data_ptr(), bufs["lnp"][PAD:].data_ptr(), W, 2,
                                 C.byref(short), 99, 0, None, None, None, flag.data_ptr(), big_ws[PAD:].data_ptr(),
                                 int(nbytes.value), use_graph, st.cuda_stream) == 4        # RBV_ESTATE
        assert b"maxiter" in lib.rbv_last_error()
    like.close()
