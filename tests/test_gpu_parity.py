"""GPU parity tests: the CUDA path (through the C ABI) against the committed golden fixtures of the real
reference and against the CPU oracle.  Tolerances are the north-star ones: model flux |delta| <= 1e-10,
lnprob relative error <= 1e-9."""
import numpy as np
import pytest

from golden_util import BIG_CASES, CASES, Golden, GoldenSightlines

pytestmark = pytest.mark.gpu

FLUX_TOL = 1e-10
LNPROB_RTOL = 1e-9


def _gpu_models(g):
    from rbvfit_b200 import FitConfiguration
    from rbvfit_b200.model import GpuVoigtModel
    cfg = FitConfiguration()
    for (z, ion, trans, comps) in g.meta["systems"]:
        cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
    models = {}
    for n in g.instruments:
        fwhm, taps = g.fwhm_for(n)
        models[n] = GpuVoigtModel(cfg, FWHM=fwhm, voigt_method=g.meta["voigt_method"], lsf_taps=taps)
    return models


@pytest.mark.parametrize("case", CASES)
def test_lowering_matches_reference(case):
    g = Golden(case)
    for n, m in _gpu_models(g).items():
        assert np.array_equal(m.atomic_lambda0, g.inst(n, "lambda0"))
        assert m.atomic_gamma.dtype == np.float32 and np.array_equal(m.atomic_gamma, g.inst(n, "gamma"))
        assert m.atomic_f.dtype == np.float32 and np.array_equal(m.atomic_f, g.inst(n, "f"))
        assert np.array_equal(m.z_factors, g.inst(n, "zfac"))
        assert np.array_equal(m.N_indices, g.inst(n, "N_indices"))
        taps = g.inst(n, "taps")
        if taps.size:
            assert np.array_equal(m.kernel, taps)
        else:
            assert m.kernel is None


@pytest.mark.parametrize("case", CASES)
def test_model_flux_vs_golden(case):
    g = Golden(case)
    for n, m in _gpu_models(g).items():
        comp = m.compile()
        wave = g.inst(n, "wave")
        ref = g.inst(n, "ref_flux")
        got = comp.model_flux(g.thetas[g.flux_rows], wave)          # batched call
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) <= FLUX_TOL
        one = comp.model_flux(g.thetas[g.flux_rows[0]], wave)       # scalar call == batch row
        assert one.shape == (wave.size,)
        assert np.array_equal(one, got[0])
        unc = m.evaluate(g.thetas[g.flux_rows[0]], wave, return_unconvolved=True)
        assert np.max(np.abs(unc - g.inst(n, "ref_flux_unconvolved")[0])) <= FLUX_TOL


@pytest.mark.parametrize("case", CASES)
def test_lnprob_vs_golden(case):
    from rbvfit_b200.likelihood import GpuLikelihood
    g = Golden(case)
    models = _gpu_models(g)
    inst = {n: dict(model=models[n], wave=g.inst(n, "wave"), flux=g.inst(n, "flux"), error=g.inst(n, "error"))
            for n in g.instruments}
    like = GpuLikelihood(inst, g.lb, g.ub)
    got = like.lnprob(g.thetas)
    ref = g.ref_lnprob
    assert np.array_equal(np.isneginf(got), np.isneginf(ref))
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    fin = np.isfinite(ref)
    rel = np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])
    assert rel.max() <= LNPROB_RTOL, rel.max()
    # scalar call returns a float that agrees with the batch row to rounding (a one-row launch may pick another tile
    # size, i.e. another far-field / summation partition); a second call is bit-identical (fixed-order sums)
    k = int(np.flatnonzero(fin)[0])
    s = like.lnprob(g.thetas[k])
    assert isinstance(s, float) and abs(s - got[k]) <= 1e-13 * abs(got[k])
    assert np.array_equal(like.lnprob(g.thetas), got, equal_nan=True)
    # walker permutation invariance
    perm = np.random.default_rng(0).permutation(len(g.thetas))
    assert np.array_equal(like.lnprob(g.thetas[perm]), got[perm], equal_nan=True)
    like.close()


@pytest.mark.parametrize("case", [c for c in CASES if c != "C1_fast"])
def test_fp32_gated_variant_within_tolerance(case):
    """The FP32-gated far-wing variant must meet the SAME tolerances as the FP64 kernel (gate: a-priori bound
    |dtau| <= 1e-11, plus the a-posteriori check against the FP64 kernel)."""
    from rbvfit_b200.likelihood import GpuLikelihood
    g = Golden(case)
    models = _gpu_models(g)
    inst = {n: dict(model=models[n], wave=g.inst(n, "wave"), flux=g.inst(n, "flux"), error=g.inst(n, "error"))
            for n in g.instruments}
    like = GpuLikelihood(inst, g.lb, g.ub)
    ref64 = like.lnprob(g.thetas)
    assert like.set_precision("fp32-gated", check_thetas=g.thetas) is True
    assert like.last_precision_check <= LNPROB_RTOL
    got = like.lnprob(g.thetas)
    ref = g.ref_lnprob
    assert np.array_equal(np.isneginf(got), np.isneginf(ref)) and np.array_equal(np.isnan(got), np.isnan(ref))
    fin = np.isfinite(ref)
    assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= LNPROB_RTOL
    assert np.max(np.abs(got[fin] - ref64[fin]) / np.abs(ref64[fin])) <= 1e-10
    assert like.set_precision("fp64") and like.precision == "fp64"
    assert np.array_equal(like.lnprob(g.thetas), ref64, equal_nan=True)
    like.close()


# ---------------------------------------------------------------------------------------------------------------
# The geometry bench.py times (100 000 px) and the survey-mode launch, against the live reference's
# committed outputs.  Both kernels are pinned: "stream" = voigt_stream_kernel (what a big batch takes by default),
# "tile" = voigt_tile_kernel with its biggest tiles, 8 pixels per lane and the separate finalisation launch.
KERNELS = {"stream": "1", "tile": "0"}


@pytest.mark.parametrize("kernel", list(KERNELS))
@pytest.mark.parametrize("case", BIG_CASES)
def test_headline_geometry_vs_golden(case, kernel, monkeypatch):
    """The 10 fixture rows (8 in bounds, 2 outside) sit at random positions of a 700-walker batch, so the launch is
    the big-batch one (reference: voigt_model.py:162-230, vfit_mcmc.py:297-353)."""
    from rbvfit_b200 import workloads as wl
    from rbvfit_b200.likelihood import GpuLikelihood
    monkeypatch.setenv("RBVFIT_B200_STREAM", KERNELS[kernel])
    g = Golden(case)
    n = g.instruments[0]
    models = _gpu_models(g)
    wave, px = g.inst(n, "wave"), g.inst(n, "flux_px")
    like = GpuLikelihood({n: dict(model=models[n], wave=wave, flux=g.inst(n, "flux"), error=g.inst(n, "error"))},
                         g.lb, g.ub)
    w = wl.get_workload(case)
    batch = wl.make_ensemble(w, 700, seed_offset=300)
    rng = np.random.default_rng(11)
    pos = np.sort(rng.choice(700, size=len(g.thetas), replace=False))
    batch[pos] = g.thetas
    got_all = like.lnprob(batch)
    assert like.engine.last_kernel == kernel
    got, ref = got_all[pos], g.ref_lnprob
    assert np.array_equal(np.isneginf(got), np.isneginf(ref)) and not np.isnan(got_all).any()
    fin = np.isfinite(ref)
    rel = np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])
    assert rel.max() <= LNPROB_RTOL, rel.max()
    # bit-identical repeats; another batch size may pick another partition (tile size / range schedule, i.e. other
    # far-field intervals and another order of the chi^2 additions): same values to rounding
    assert np.array_equal(like.lnprob(batch), got_all)
    sub = like.lnprob(batch[: pos[3] + 1])[pos[:4]]
    assert np.max(np.abs(sub - got[:4]) / np.abs(got[:4])) <= 1e-13
    # model flux of two fixture rows on the fixture's pixel subset (flux mode of the tile kernel)
    comp = models[n].compile()
    flux = comp.model_flux(g.thetas[g.flux_rows], wave)
    assert np.max(np.abs(flux[:, px] - g.inst(n, "ref_flux"))) <= FLUX_TOL
    unc = models[n].evaluate(g.thetas[0], wave, return_unconvolved=True)
    assert np.max(np.abs(unc[px] - g.inst(n, "ref_flux_unconvolved")[0])) <= FLUX_TOL
    like.close()


def test_headline_geometry_kernels_agree(monkeypatch):
    """Same batch through both kernels: different summation trees, same numbers to a few ulp of the chi^2 sum."""
    from rbvfit_b200 import workloads as wl
    from rbvfit_b200.likelihood import GpuLikelihood
    g = Golden("C5a")
    n = g.instruments[0]
    out = {}
    for kernel, flag in KERNELS.items():
        monkeypatch.setenv("RBVFIT_B200_STREAM", flag)
        like = GpuLikelihood({n: dict(model=_gpu_models(g)[n], wave=g.inst(n, "wave"), flux=g.inst(n, "flux"),
                                      error=g.inst(n, "error"))}, g.lb, g.ub)
        out[kernel] = like.lnprob(wl.make_ensemble(wl.get_workload("C5a"), 600))
        assert like.engine.last_kernel == kernel
        like.close()
    fin = np.isfinite(out["tile"])
    assert np.array_equal(fin, np.isfinite(out["stream"]))
    assert np.max(np.abs(out["stream"][fin] - out["tile"][fin]) / np.abs(out["tile"][fin])) <= 1e-12


@pytest.mark.parametrize("kernel", list(KERNELS))
def test_sightline_batch_vs_golden_and_oracle(kernel, monkeypatch):
    """Survey mode (the wps > 0 launch): 8 sightlines x 16 walkers in ONE batch against each sightline's own
    reference vfit (fixture) and against the oracle per sightline (S x vfit(...).lnprob, vfit_mcmc.py:127-197,
    348-353)."""
    from oracle import voigt_oracle as vo
    from rbvfit_b200 import FitConfiguration
    from rbvfit_b200.likelihood import SightlineBatch
    from rbvfit_b200.model import GpuVoigtModel
    monkeypatch.setenv("RBVFIT_B200_STREAM", KERNELS[kernel])
    g = GoldenSightlines()
    sight = []
    for s in range(g.n):
        cfg = FitConfiguration()
        for (z, ion, trans, comps) in g.systems(s):
            cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
        sight.append(dict(model=GpuVoigtModel(cfg, FWHM="6.5"), wave=g.waves[s], flux=g.flux[s], error=g.errors[s]))
    batch = SightlineBatch(sight, g.lb, g.ub)
    got = batch.lnprob(g.thetas)
    assert batch.engine.last_kernel == kernel
    assert got.shape == g.ref_lnprob.shape
    assert np.array_equal(np.isneginf(got), np.isneginf(g.ref_lnprob)) and not np.isnan(got).any()
    fin = np.isfinite(g.ref_lnprob)
    assert np.max(np.abs(got[fin] - g.ref_lnprob[fin]) / np.abs(g.ref_lnprob[fin])) <= LNPROB_RTOL
    for s in range(g.n):
        ora = vo.lnprob_batch(g.oracle_compiled(s), g.thetas[s], g.lb, g.ub)
        f = np.isfinite(ora)
        assert np.array_equal(f, np.isfinite(got[s]))
        assert np.max(np.abs(got[s][f] - ora[f]) / np.abs(ora[f])) <= LNPROB_RTOL
    batch.close()
