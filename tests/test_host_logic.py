"""CPU tests of the host layer: configuration mirror, line table, LSF taps, workloads, sampler."""
import numpy as np
import pytest

from golden_util import Golden


def test_line_table_float32_rounding():
    from rbvfit_b200 import rb_setline
    r = rb_setline(2796.3, "closest")
    assert r["wave"][0] == 2796.3543 or abs(r["wave"][0] - 2796.35) < 0.01
    assert r["fval"].dtype == np.float32 and r["gamma"].dtype == np.float32
    assert float(r["fval"][0]) == 0.6122999787330627
    assert r["name"][0].startswith("MgII")
    with pytest.raises(ValueError):
        rb_setline(1000.0, "nearest")
    assert len(rb_setline(1215.6701, "Exact")["wave"]) == 1


def test_config_snaps_and_ties():
    from rbvfit_b200 import FitConfiguration
    g = Golden("C2")
    cfg = FitConfiguration()
    for (z, ion, trans, comps) in g.meta["systems"]:
        cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
    cfg.validate()
    assert len(cfg.systems) == 3 and [len(s.ion_groups) for s in cfg.systems] == [3, 3, 3]
    assert cfg.systems[0].ion_groups[0].transitions[0] == 1548.2049 or \
        abs(cfg.systems[0].ion_groups[0].transitions[0] - 1548.2) < 0.01
    with pytest.raises(ValueError):
        cfg.add_system(z=2.0, ion="CIV", transitions=[1548.2], components=1)     # duplicate ion
    with pytest.raises(ValueError):
        FitConfiguration().add_system(z=0.1, ion="MgII", transitions=[1548.2])    # wrong ion
    with pytest.raises(ValueError):
        FitConfiguration().validate()


def test_gaussian_taps_match_reference_fixture():
    from rbvfit_b200 import lsf
    assert np.array_equal(lsf.gaussian_taps("6.5"), Golden("C1").inst("COS", "taps"))
    assert np.array_equal(lsf.gaussian_taps(2.394991274145626), Golden("test_script").inst("COS", "taps"))
    assert lsf.gaussian_taps("3.0").size == 11 and lsf.gaussian_taps("4.0").size == 15
    with pytest.raises(ImportError):
        lsf.cos_taps("G130M", "1", "1300A")       # linetools absent -> same error type as the reference
    from oracle import voigt_oracle as vo
    assert np.array_equal(lsf.cos_like_taps(321), vo.cos_like_lsf(321))


def test_workloads_are_deterministic():
    from rbvfit_b200 import workloads as wl
    a, b = wl.get_workload("C2"), wl.get_workload("C2")
    assert np.array_equal(a["theta_true"], b["theta_true"]) and a["theta_true"].size == 36
    e1, e2 = wl.make_ensemble(a), wl.make_ensemble(b)
    assert np.array_equal(e1, e2)
    out = np.any((e1 < a["lb"]) | (e1 > a["ub"]), axis=1)
    assert out.sum() >= 1                       # the -inf path is exercised
    assert wl.get_workload("C5a")["nwalkers"] == 8192
    s = wl.c5b_sightline(3)
    assert 0.3 <= s["systems"][0][0] <= 0.4 and s["nwalkers"] == 64


def test_set_bounds_traditional():
    from rbvfit_b200.vfit_mcmc import set_bounds
    _, lb, ub = set_bounds([14.0], [20.0], [0.0])
    assert list(lb) == [12.0, 2.0, -50.0] and list(ub) == [16.0, 60.0, 50.0]
    _, lb, ub = set_bounds([14.0], [130.0], [0.0], vlow=[-10.0])
    assert ub[1] == 150.0 and lb[2] == -10.0


def test_stretch_sampler_recovers_gaussian():
    from rbvfit_b200.sampler import EnsembleSampler, integrated_time
    mu = np.array([1.0, -2.0, 0.5])
    sig = np.array([0.5, 2.0, 1.0])
    calls = []

    def lnp(x):
        calls.append(x.shape)
        return -0.5 * np.sum(((x - mu) / sig) ** 2, axis=1)

    rng = np.random.default_rng(1)
    s = EnsembleSampler(32, 3, lnp, seed=2)
    s.run_mcmc(mu + 1e-2 * rng.standard_normal((32, 3)), 1500)
    assert all(c[1] == 3 and c[0] in (16, 32) for c in calls)      # batched half-ensembles
    flat = s.get_chain(discard=300, flat=True)
    assert flat.shape == (1200 * 32, 3)
    assert np.all(np.abs(flat.mean(0) - mu) < 0.15 * sig)
    assert np.all(np.abs(flat.std(0) / sig - 1) < 0.1)
    af = s.acceptance_fraction
    assert af.shape == (32,) and 0.3 < af.mean() < 0.85
    tau = s.get_autocorr_time(quiet=True)
    assert tau.shape == (3,) and np.all(tau > 1) and np.all(tau < 200)
    assert s.get_chain().shape == (1500, 32, 3) and s.get_log_prob().shape == (1500, 32)
    with pytest.raises(ValueError):
        EnsembleSampler(8, 3, lnp).run_mcmc(np.zeros((8, 3)), 1)     # degenerate initial state


def test_slice_sampler_recovers_gaussian():
    """zeus-style ensemble slice sampler (restated, rbvfit_b200/slice_sampler.py): statistical parity on a
    correlated Gaussian, batched lnprob calls only, zeus's accessor contract."""
    from rbvfit_b200.slice_sampler import EnsembleSliceSampler
    mu = np.array([1.0, -2.0, 0.5])
    cov = np.array([[0.25, 0.3, 0.0], [0.3, 4.0, 0.5], [0.0, 0.5, 1.0]])
    icov = np.linalg.inv(cov)
    shapes = []

    def lnp(x):
        shapes.append(x.shape)
        d = x - mu
        return -0.5 * np.einsum("ni,ij,nj->n", d, icov, d)

    rng = np.random.default_rng(3)
    s = EnsembleSliceSampler(16, 3, lnp, seed=4)
    s.run_mcmc(mu + 1e-2 * rng.standard_normal((16, 3)), 700)
    assert all(len(sh) == 2 and sh[1] == 3 and 1 <= sh[0] <= 16 for sh in shapes)
    flat = s.get_chain(discard=100, flat=True)
    assert flat.shape == (600 * 16, 3)
    sig = np.sqrt(np.diag(cov))
    assert np.all(np.abs(flat.mean(0) - mu) < 0.15 * sig)
    assert np.all(np.abs(flat.std(0) / sig - 1) < 0.1)
    assert abs(np.corrcoef(flat.T)[0, 1] - 0.3) < 0.08
    assert s.acceptance_fraction.shape == (16,) and s.acceptance_fraction.min() > 0.95   # slice moves always move
    assert 0.05 < s.efficiency < 1.0 and not s.tune                                       # mu tuning converged
    tau = s.get_autocorr_time(quiet=True)
    assert tau.shape == (3,) and np.all(tau < 50)
    with pytest.raises(ValueError):
        EnsembleSliceSampler(5, 3, lnp)           # zeus: >= 2*ndim walkers, even


def test_roofline_flops_rule():
    from rbvfit_b200 import roofline as rf, workloads as wl
    from oracle import voigt_oracle as vo
    w = wl.get_workload("C1")
    cfg = vo.OracleConfig()
    for (z, ion, trans, comps) in w["systems"]:
        cfg.add_system(z, ion, trans, comps)
    m = vo.lower(cfg)
    F, tiers = rf.flops_per_walker_pixel(m, w["theta_true"], w["instruments"]["COS"]["wave"], 23)
    assert abs(sum(tiers.values()) - 1) < 1e-12
    assert 150 < F < 398            # ~280 for C1, upper bound 80 L + 2 K + 32 = 398


def test_roofline_farfield_rule():
    """Algorithmic flops of the far-field algorithm (rbvfit_b200/roofline.py): never more than the direct rule, equal
    to it when no line qualifies (C1: every pixel is within 200 Doppler widths of the doublet), ~5x less at C5a's
    structure (C2 grid here to keep the CPU suite fast)."""
    from rbvfit_b200 import roofline as rf, workloads as wl
    from oracle import voigt_oracle as vo
    out = {}
    for name in ("C1", "C2", "C4"):
        w = wl.get_workload(name)
        cfg = vo.OracleConfig()
        for (z, ion, trans, comps) in w["systems"]:
            cfg.add_system(z, ion, trans, comps)
        m = vo.lower(cfg)
        wave = list(w["instruments"].values())[0]["wave"]
        Fd, _ = rf.flops_per_walker_pixel(m, w["theta_true"], wave, 23)
        Ff, t = rf.flops_farfield(m, w["theta_true"], wave, 23)
        assert abs(sum(t.values()) - 1) < 1e-12
        out[name] = (Fd, Ff, t)
    assert out["C1"][0] == out["C1"][1] and out["C1"][2]["farfield"] == 0.0
    assert out["C2"][1] < 0.4 * out["C2"][0] and out["C2"][2]["farfield"] > 0.75
    assert out["C4"][1] < out["C4"][0] and 0.5 < out["C4"][2]["farfield"] < 0.9     # the DLA's wings stay direct
    # SURVEY 8(d)'s L = 4 companion of C5a (the MgII doublet on C5a's 100 000-pixel grid, 8192 walkers)
    w = wl.get_workload("C5a_L4")
    cfg = vo.OracleConfig()
    for (z, ion, trans, comps) in w["systems"]:
        cfg.add_system(z, ion, trans, comps)
    m = vo.lower(cfg)
    wave = w["instruments"]["SPEC"]["wave"]
    assert m.n_lines == 4 and wave.size == 100000 and wl.make_ensemble(w).shape == (8192, 6)
    Fd, td = rf.flops_per_walker_pixel(m, w["theta_true"], wave, 23)
    Ff, tf = rf.flops_farfield(m, w["theta_true"], wave, 23)
    assert 160 < Fd < 170 and 95 < Ff < 110 and tf["farfield"] > 0.9 and td["far"] > 0.95


def test_stretch_replay_philox_known_answers_and_gaussian_target():
    """oracle/stretch_replay.py (numpy restatement of the device sampler): Philox4x32-10 against the Random123
    known-answer vectors, the step permutation is a bijection with balanced halves, and the sampler -- random
    affine split included -- recovers a correlated Gaussian."""
    from oracle import stretch_replay as sr
    assert sr.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert sr.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert sr.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)
    for W in (2, 7, 50, 64):
        for step in range(5):
            a, b = sr.step_perm(12345, step, W)
            assert sorted((a * p + b) % W for p in range(W)) == list(range(W))
    us = [sr.u01(*sr._rand(9, s, 3, 1)[:2]) for s in range(2000)]
    assert 0.0 < min(us) and max(us) < 1.0 and abs(np.mean(us) - 0.5) < 0.02
    mu = np.array([1.0, -2.0, 0.5])
    cov = np.array([[0.25, 0.3, 0.0], [0.3, 4.0, 0.5], [0.0, 0.5, 1.0]])
    icov = np.linalg.inv(cov)

    def lnp(x):
        d = np.atleast_2d(x) - mu
        return -0.5 * np.einsum("ni,ij,nj->n", d, icov, d)

    rng = np.random.default_rng(0)
    p0 = mu + 0.1 * rng.standard_normal((24, 3))
    chain, lps, nacc = sr.run(lnp, p0, lnp(p0), 1500, seed=424242)
    assert np.allclose(lps[-1], lnp(chain[-1]))
    flat = chain[300:].reshape(-1, 3)
    sig = np.sqrt(np.diag(cov))
    assert np.all(np.abs(flat.mean(0) - mu) < 0.15 * sig)
    assert np.all(np.abs(flat.std(0) / sig - 1) < 0.12)
    assert abs(np.corrcoef(flat.T)[0, 1] - 0.3) < 0.1
    assert 0.3 < (nacc / 1500).mean() < 0.85


def test_slice_replay_recovers_gaussian_and_tunes_mu():
    """oracle/slice_replay.py (numpy restatement of rbv_slice_run: lockstep widen / shrink state machines, Philox
    streams): it recovers a correlated Gaussian, every stored lnprob belongs to its stored position, mu settles
    (expansions ~ contractions), a continued run equals one long run, and every half-step needs at least two
    batches (both bracket ends, one draw)."""
    from oracle import slice_replay as sl
    mu = np.array([1.0, -2.0, 0.5])
    cov = np.array([[0.25, 0.3, 0.0], [0.3, 4.0, 0.5], [0.0, 0.5, 1.0]])
    icov = np.linalg.inv(cov)

    def lnp(x):
        d = np.atleast_2d(x) - mu
        return -0.5 * np.einsum("ni,ij,nj->n", d, icov, d)

    rng = np.random.default_rng(1)
    p0 = mu + 0.1 * rng.standard_normal((12, 3))
    out = sl.run(lnp, p0, lnp(p0), 700, seed=20260, mu=1.0)
    chain, lps = out["chain"], out["lnp_chain"]
    assert np.allclose(lps, lnp(chain.reshape(-1, 3)).reshape(lps.shape))
    assert np.all(np.any(chain[1:] != chain[:-1], axis=2))            # a slice move always moves
    flat = chain[150:].reshape(-1, 3)
    sig = np.sqrt(np.diag(cov))
    assert np.all(np.abs(flat.mean(0) - mu) < 0.15 * sig)
    assert np.all(np.abs(flat.std(0) / sig - 1) < 0.12)
    assert abs(np.corrcoef(flat.T)[0, 1] - 0.3) < 0.1
    assert not out["tune"] and out["good"] == 6 and 0.2 < out["mu"] < 5.0
    assert out["nbatches"] >= 2 * 2 * 700 and out["ncall"] >= 3 * 12 * 700
    # continuation: 300 + 400 steps with the carried tuning state == 700 steps
    a = sl.run(lnp, p0, lnp(p0), 300, seed=20260, mu=1.0)
    b = sl.run(lnp, a["chain"][-1], a["lnp_chain"][-1], 400, seed=20260, mu=a["mu"], tune=a["tune"], good=a["good"],
               first_step=300)
    assert np.array_equal(np.concatenate([a["chain"], b["chain"]]), chain)
    # a tiny stepping-out budget still samples correctly (the bracket just cannot widen)
    c = sl.run(lnp, p0, lnp(p0), 50, seed=3, mu=0.01, maxsteps=1, tune=False)
    assert c["nexp"] == 0 and np.all(np.isfinite(c["lnp_chain"]))
    # two logical iterations per batch (the device default: the second one's candidates are speculative): the chain,
    # mu and every counter of the sequential algorithm, in between half and all of its batches -- also with a
    # stepping-out budget that runs out, and with an lnprob that is -inf outside a box (rejections on one side)
    spec = sl.run(lnp, p0, lnp(p0), 700, seed=20260, mu=1.0, depth=2)
    for key in ("chain", "lnp_chain", "mus"):
        assert np.array_equal(spec[key], out[key]), key
    assert [spec[k] for k in ("mu", "tune", "good", "nexp", "ncon", "ncall")] == \
           [out[k] for k in ("mu", "tune", "good", "nexp", "ncon", "ncall")]
    assert out["nbatches"] / 2 <= spec["nbatches"] < 0.75 * out["nbatches"]

    def boxed(x):
        x = np.atleast_2d(x)
        return np.where(np.all(np.abs(x - mu) < 3.0, axis=1), lnp(x), -np.inf)

    for kw in (dict(maxsteps=3, mu=0.05, tune=False), dict(mu=5.0)):
        one = sl.run(boxed, p0, boxed(p0), 120, seed=11, **kw)
        two = sl.run(boxed, p0, boxed(p0), 120, seed=11, depth=2, **kw)
        assert np.array_equal(one["chain"], two["chain"]) and np.array_equal(one["lnp_chain"], two["lnp_chain"])
        assert (one["mu"], one["nexp"], one["ncon"], one["ncall"]) == (two["mu"], two["nexp"], two["ncon"], two["ncall"])
        assert two["nbatches"] < one["nbatches"]


def test_sightline_sampler_accessors_without_device():
    """SightlineEnsembleSampler's bookkeeping (shapes of get_chain / get_log_prob with discard, thin, flat and per
    sightline; the emcee-shaped per-sightline view) on a hand-filled chain -- no device involved."""
    from rbvfit_b200.sampler import SightlineEnsembleSampler

    class _Batch:                 # what the sampler reads from a SightlineBatch before sampling
        engine, n_sightlines, ndim = object(), 3, 2

    S, W, nd, n = 3, 4, 2, 10
    smp = SightlineEnsembleSampler(W, nd, _Batch(), seed=1)
    assert smp.get_chain().shape == (0, S, W, nd) and smp.acceptance_fraction.shape == (S, W)
    chain = np.arange(n * S * W * nd, dtype=float).reshape(n, S, W, nd)
    smp._chain, smp._log_prob = chain, chain[..., 0] * 0.5
    smp._accepted[:] = np.arange(S * W).reshape(S, W)
    smp.iteration = n
    assert np.array_equal(smp.get_chain(discard=2, thin=2), chain[3::2])                  # emcee's thin convention
    assert smp.get_chain(flat=True).shape == (S, n * W, nd)
    assert np.array_equal(smp.get_chain(flat=True)[1], chain[:, 1].reshape(-1, nd))       # never mixes sightlines
    assert smp.get_log_prob(flat=True, discard=4).shape == (S, (n - 4) * W)
    assert np.array_equal(smp.get_chain(sightline=2, discard=1), chain[1:, 2])
    v = smp.sightline(1)
    assert (v.nwalkers, v.ndim, v.iteration) == (W, nd, n)
    assert np.array_equal(v.get_chain(flat=True, discard=3), chain[3:, 1].reshape(-1, nd))
    assert v.chain.shape == (W, n, nd) and v.flatchain.shape == (n * W, nd) and v.lnprobability.shape == (W, n)
    assert np.array_equal(v.acceptance_fraction, np.arange(4, 8) / n)
    assert v.get_autocorr_time(quiet=True).shape == (nd,)
    with pytest.raises(IndexError):
        smp.sightline(3)
    with pytest.raises(ValueError):
        SightlineEnsembleSampler(W, nd + 1, _Batch())
    with pytest.raises(ValueError):
        SightlineEnsembleSampler(1, nd, _Batch())
    with pytest.raises(TypeError):
        SightlineEnsembleSampler(W, nd, object())


@pytest.mark.reference
def test_product_lowering_matches_live_reference():
    """Build container only: the PRODUCT's host layer (rb_setline, FitConfiguration, GpuVoigtModel lowering, Gaussian
    LSF taps, set_bounds) against the unmodified reference on the 24 fuzz problems and a sweep of line look-ups --
    no device needed: lowering is host code."""
    import contextlib
    import io
    import pickle
    from fuzz_util import draw_problem
    from oracle import refshim
    from rbvfit_b200 import FitConfiguration, rb_setline
    from rbvfit_b200.model import GpuVoigtModel
    from rbvfit_b200.vfit_mcmc import set_bounds
    RefConfig, RefModel, mc, vm = refshim.import_reference()
    from rbvfit.rb_setline import rb_setline as ref_setline
    for lam in (1215.67, 1025.7, 1302.17, 1304.5, 1548.2, 1550.3, 2796.3, 2803.5, 1393.76, 977.02, 2600.17):
        with contextlib.redirect_stdout(io.StringIO()):
            a, b = rb_setline(lam, "closest"), ref_setline(lam, "closest")
        assert a["wave"] == b["wave"] and a["name"] == b["name"]
        assert np.asarray(a["fval"]).dtype == np.asarray(b["fval"]).dtype == np.float32
        assert a["fval"] == b["fval"] and a["gamma"] == b["gamma"]
    for seed in range(24):
        systems, wave, fwhm, taps, theta, thetas, lb, ub, rng = draw_problem(1000 + seed)
        cfg, rcfg = FitConfiguration(), RefConfig()
        for (z, ion, trans, comps) in systems:
            cfg.add_system(z=z, ion=ion, transitions=list(trans), components=comps)
            rcfg.add_system(z=z, ion=ion, transitions=list(trans), components=comps)
        ours = pickle.loads(pickle.dumps(GpuVoigtModel(cfg, FWHM=fwhm, lsf_taps=taps).compile())).data
        ref = RefModel(rcfg, FWHM=fwhm).compile().data
        for key in ("atomic_lambda0", "atomic_gamma", "atomic_f", "z_factors", "N_indices", "b_indices", "v_indices"):
            x, y = np.asarray(getattr(ours, key)), np.asarray(getattr(ref, key))
            assert x.dtype == y.dtype and np.array_equal(x, y), (seed, key)
        assert (ours.n_lines, ours.total_components) == (ref.n_lines, ref.total_components)
        if taps is None and fwhm is not None:
            assert np.array_equal(ours.kernel, np.asarray(ref.kernel.array))
        # the reference's own FitConfiguration object is accepted as is (duck-typed)
        again = GpuVoigtModel(rcfg, FWHM=fwhm, lsf_taps=taps).compile().data
        assert np.array_equal(again.atomic_lambda0, ref.atomic_lambda0) and np.array_equal(again.N_indices,
                                                                                           ref.N_indices)
        C = ref.total_components
        n, b, v = theta[:C], theta[C:2 * C], theta[2 * C:]
        with contextlib.redirect_stdout(io.StringIO()):
            _, rlb, rub = mc.set_bounds(list(n), list(b), list(v))
            _, olb, oub = set_bounds(list(n), list(b), list(v))
        assert np.array_equal(np.asarray(olb), np.asarray(rlb)) and np.array_equal(np.asarray(oub), np.asarray(rub))


def test_gelman_rubin_statistic():
    """vfit._print_diagnostics' R-hat (vfit_mcmc.py:631-647): ~1 for chains of one distribution, > 1.1 when the
    chains sit in different places; the closed form on a tiny case."""
    from rbvfit_b200.vfit_mcmc import gelman_rubin
    rng = np.random.default_rng(0)
    same = rng.standard_normal((16, 2000, 3))
    r = gelman_rubin(same)
    assert r.shape == (3,) and np.all(np.abs(r - 1.0) < 0.01)
    apart = same + np.arange(16)[:, None, None]
    assert np.all(gelman_rubin(apart) > 1.1)
    x = np.array([[[0.0], [2.0]], [[1.0], [5.0]]])           # m = 2 chains, n = 2 steps
    W, B_over_n = (2.0 + 8.0) / 2, 2.0                        # within-chain variances 2 and 8; means 1 and 3
    V = 0.5 * W + B_over_n + B_over_n / 2
    assert np.allclose(gelman_rubin(x), np.sqrt(V / W))
    with pytest.raises(ValueError):
        gelman_rubin(np.zeros((1, 10, 2)))


def _piecewise_problem():
    """C1's MgII doublet on a 2400-pixel grid with three LSF blocks: Gaussian FWHM 6.5, Gaussian FWHM 3.0, and an
    asymmetric 31-tap table (normalised by the convolution, like a CustomKernel)."""
    from oracle import voigt_oracle as vo
    from rbvfit_b200 import FitConfiguration, workloads as wl
    from rbvfit_b200.model import GpuVoigtModel
    w = wl.get_workload("C1")
    cfg = FitConfiguration()
    for (z, ion, trans, comps) in w["systems"]:
        cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
    wave = np.linspace(3755.0, 3795.0, 2400)
    x = np.arange(-15, 16)
    table = np.exp(-0.5 * (x / 3.0) ** 2) * (1 + 0.4 * (x > 0))
    specs = [("6.5", None), ("3.0", None), (None, table)]
    starts = [wave[0], 3768.3, 3781.7]
    models = [GpuVoigtModel(cfg, FWHM=f, lsf_taps=t) for f, t in specs]
    lowered = [vo.lower(cfg, FWHM=f if f is not None else "6.5", custom_taps=t) for f, t in specs]
    rng = np.random.default_rng(5)
    truth = vo.model_flux_piecewise_lsf(lowered, starts, w["theta_true"], wave)
    flux = truth + 0.05 * rng.standard_normal(wave.size)
    error = np.full(wave.size, 0.05)
    return w, cfg, wave, flux, error, starts, models, lowered


def test_piecewise_lsf_split_equals_its_definition():
    """Wavelength-dependent LSF (extension): the split into sub-instruments with weight-0 halos
    (rbvfit_b200.lsf.piecewise_lsf_instruments) reproduces the definition -- every output pixel convolved with the
    kernel of ITS block, spectrum edges replicated (oracle.model_flux_piecewise_lsf) -- on the CPU oracle: the split
    is host logic, no device involved."""
    from oracle import voigt_oracle as vo
    from rbvfit_b200 import lsf
    w, cfg, wave, flux, error, starts, models, lowered = _piecewise_problem()
    entries = lsf.piecewise_lsf_instruments("COS", wave, flux, error, list(zip(starts, models)))
    assert list(entries) == ["COS[0]", "COS[1]", "COS[2]"]
    # every pixel is counted exactly once, halos are K // 2 wide on interior sides only
    assert np.array_equal(np.concatenate([d["wave"][d["weight_mask"]] for d in entries.values()]), wave)
    halos = [(int(np.argmax(d["weight_mask"])), int(np.argmax(d["weight_mask"][::-1]))) for d in entries.values()]
    assert halos == [(0, 11), (5, 5), (15, 0)]
    theta = w["theta_true"] + np.array([0.05, -0.03, 2.0, -1.5, 3.0, -4.0])
    direct = vo.model_flux_piecewise_lsf(lowered, starts, theta, wave)
    by_name = dict(zip(entries, lowered))
    stitched = lsf.piecewise_lsf_flux(entries, lambda n, d: vo.model_flux(by_name[n], theta, d["wave"]))
    assert stitched.shape == wave.shape and np.max(np.abs(stitched - direct)) <= 2e-15
    # a single block is the single-kernel model
    one = lsf.piecewise_lsf_instruments("COS", wave, flux, error, [(wave[0], models[0])])
    assert np.all(one["COS[0]"]["weight_mask"]) and len(one["COS[0]"]["wave"]) == wave.size
    assert np.array_equal(vo.model_flux_piecewise_lsf(lowered[:1], starts[:1], theta, wave),
                          vo.model_flux(lowered[0], theta, wave))
    with pytest.raises(ValueError):
        lsf.piecewise_lsf_instruments("COS", wave, flux, error, [(3780.0, models[0]), (3770.0, models[1])])
    with pytest.raises(ValueError):
        lsf.piecewise_lsf_instruments("COS", wave, flux, error, [(wave[0], models[0]), (9999.0, models[1])])
