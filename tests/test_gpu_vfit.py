"""GPU: the ``vfit`` mirror end to end -- batched optimiser, walker initialisation, stretch-move sampling --
and its sampler-object contract; chain parity against the same sampler driven by the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _c1_fitter(nwalkers=32, nsteps=60, seed=5, sampler="emcee"):
    from oracle import voigt_oracle as vo
    from rbvfit_b200 import FitConfiguration, workloads as wl
    from rbvfit_b200.model import GpuVoigtModel
    from rbvfit_b200.vfit_mcmc import vfit
    w = wl.get_workload("C1")
    cfg, ocfg = FitConfiguration(), vo.OracleConfig()
    for (z, ion, trans, comps) in w["systems"]:
        cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
        ocfg.add_system(z, ion, trans, comps)
    omodel = vo.lower(ocfg, FWHM="6.5")
    spectra = wl.make_spectra(w, lambda n, th, wave: vo.model_flux(omodel, th, wave))
    s = spectra["COS"]
    model = GpuVoigtModel(cfg, FWHM="6.5")
    theta0 = np.clip(w["theta_true"] + np.array([0.05, -0.05, 2, -2, 3, -3.0]), w["lb"], w["ub"])
    fitter = vfit({"COS": dict(model=model, wave=s["wave"], flux=s["flux"], error=s["error"])}, theta0,
                  w["lb"], w["ub"], no_of_Chain=nwalkers, no_of_steps=nsteps, seed=seed, sampler=sampler)
    comp = vo.compile_instruments({"COS": dict(model=omodel, **s)})
    return w, fitter, comp, theta0


def test_vfit_validation_errors():
    from rbvfit_b200.vfit_mcmc import vfit
    with pytest.raises(TypeError):
        vfit([], [0.0], [0.0], [1.0])
    with pytest.raises(ValueError):
        vfit({}, [0.0], [0.0], [1.0])
    with pytest.raises(ValueError):
        vfit({"a": {"model": None, "wave": [1], "flux": [1]}}, [0.0], [0.0], [1.0])
    w, fitter, comp, theta0 = _c1_fitter()
    from rbvfit_b200.vfit_mcmc import vfit as V
    with pytest.raises(ValueError):
        V({"COS": dict(model=lambda t, x: x, wave=[1.0], flux=[1.0], error=[1.0])}, w["ub"] + 1, w["lb"], w["ub"])
    with pytest.raises(TypeError):       # arbitrary callables cannot run on the device; no CPU fallback
        V({"COS": dict(model=lambda t, x: x, wave=[1.0], flux=[1.0], error=[1.0])}, theta0, w["lb"], w["ub"])


def test_lnprob_lnlike_lnprior_match_oracle():
    from oracle import voigt_oracle as vo
    w, fitter, comp, theta0 = _c1_fitter()
    ref = vo.lnprob(comp, theta0, w["lb"], w["ub"])
    assert abs(fitter.lnprob(theta0) - ref) / abs(ref) <= 1e-9
    out = theta0.copy()
    out[0] = w["ub"][0] + 0.5
    assert fitter.lnprior(out) == -np.inf and fitter.lnprob(out) == -np.inf
    ll = fitter.lnlike(out)                      # lnlike ignores the prior, like the reference's
    assert abs(ll - vo.lnlike(comp, out)) / abs(ll) <= 1e-9
    assert fitter.lnprob(theta0) == fitter.lnprob(theta0[None, :])[0]


def test_optimize_and_walker_init():
    w, fitter, comp, theta0 = _c1_fitter(nwalkers=40)
    before = fitter.lnprob(theta0)
    popt = fitter.optimize_guess(theta0)
    after = fitter.lnprob(popt)
    assert after > before + 1.0
    assert np.all(popt >= w["lb"]) and np.all(popt <= w["ub"])
    # the optimum is at least as good as the truth the data were drawn from (up to the noise realisation);
    # the two blended components make (N1, N2) nearly degenerate, so compare likelihoods, not parameters
    assert after >= fitter.lnprob(w["theta_true"]) - 5.0
    assert np.all(np.abs(popt - w["theta_true"]) < np.array([0.3, 0.3, 10, 10, 8, 8]))
    g = fitter._initialize_walkers(popt)
    assert g.shape == (40, 6) and np.all(g > w["lb"]) and np.all(g < w["ub"])
    assert np.all(np.isfinite(fitter.lnprob(g)))


def test_runmcmc_contract_and_posterior():
    w, fitter, comp, theta0 = _c1_fitter(nwalkers=32, nsteps=400, seed=11)
    fitter.runmcmc(optimize=True, verbose=False, use_pool=True, progress=False)
    s = fitter.sampler
    assert fitter.mcmc_flag and s.get_chain().shape == (400, 32, 6)
    assert s.get_chain(discard=100, flat=True).shape == (300 * 32, 6)
    assert fitter.samples.shape == (300 * 32, 6) and fitter.best_theta.shape == (6,)
    af = s.acceptance_fraction
    assert af.shape == (32,) and 0.15 < af.mean() < 0.8
    tau = s.get_autocorr_time(quiet=True)
    assert tau.shape == (6,) and np.all(np.isfinite(tau))
    # the posterior brackets the truth the spectrum was generated from
    lo, hi = np.percentile(fitter.samples, [0.5, 99.5], axis=0)
    assert np.all(lo < w["theta_true"]) and np.all(w["theta_true"] < hi)
    assert np.all(fitter.low_theta <= fitter.best_theta) and np.all(fitter.best_theta <= fitter.high_theta)
    assert fitter.get_samples(burn_in=0.5).shape == (200 * 32, 6)
    assert s.n_logp_rows == 32 + 400 * 32        # one batched call per half-step, one row per proposal


def test_chain_matches_cpu_oracle_chain():
    """Same sampler, same seed: driven by the GPU lnprob and by the CPU oracle the chains coincide (an
    accept/reject flip needs |ln u - ln q| < 1e-9, i.e. essentially never in 1280 decisions)."""
    from oracle import voigt_oracle as vo
    from rbvfit_b200.sampler import EnsembleSampler
    w, fitter, comp, theta0 = _c1_fitter()
    rng = np.random.default_rng(2)
    p0 = np.clip(w["theta_true"] + 1e-3 * rng.standard_normal((16, 6)), w["lb"], w["ub"])
    gpu = EnsembleSampler(16, 6, fitter.lnprob, seed=9)
    cpu = EnsembleSampler(16, 6, lambda th: vo.lnprob_batch(comp, th, w["lb"], w["ub"]), seed=9)
    cg, lg = gpu.run_mcmc(p0, 80)
    cc, lc = cpu.run_mcmc(p0, 80)
    assert np.allclose(cg, cc, rtol=0, atol=1e-9)
    assert np.max(np.abs(lg - lc) / np.abs(lc)) <= 1e-9
    assert np.array_equal(gpu.acceptance_fraction, cpu.acceptance_fraction)


def test_device_sampler_bookkeeping_reproducibility_and_continuation():
    """rbv_stretch_run: the recorded lnprob of every stored position equals the likelihood recomputed for it, walkers
    never leave the prior box, a seed reproduces the chain bit for bit with and without the CUDA graph, and
    run(60) + run(None, 40) is the same chain as run(100)."""
    from rbvfit_b200.sampler import DeviceEnsembleSampler
    w, fitter, comp, theta0 = _c1_fitter()
    like = fitter._like
    rng = np.random.default_rng(3)
    p0 = np.clip(w["theta_true"] + 1e-3 * rng.standard_normal((20, 6)), w["lb"], w["ub"])
    a = DeviceEnsembleSampler(20, 6, like, seed=5, use_graph=True)
    a.run_mcmc(p0, 100)
    chain, lps = a.get_chain(), a.get_log_prob()
    assert chain.shape == (100, 20, 6) and lps.shape == (100, 20)
    for step in (0, 37, 99):
        assert np.array_equal(like.lnprob(chain[step]), lps[step])
    assert np.all(chain >= w["lb"]) and np.all(chain <= w["ub"]) and np.all(np.isfinite(lps))
    moved = np.any(chain[1:] != chain[:-1], axis=2)
    af = a.acceptance_fraction
    assert np.allclose(af, moved.sum(axis=0) / 100.0 + (np.any(chain[0] != p0, axis=1)) / 100.0)
    assert 0.2 < af.mean() < 0.9
    b = DeviceEnsembleSampler(20, 6, like, seed=5, use_graph=False)
    b.run_mcmc(p0, 60)
    b.run_mcmc(None, 40)
    assert np.array_equal(b.get_chain(), chain) and np.array_equal(b.get_log_prob(), lps)
    assert np.array_equal(b.acceptance_fraction, af)
    c = DeviceEnsembleSampler(20, 6, like, seed=6)
    c.run_mcmc(p0, 30)
    assert not np.array_equal(c.get_chain(), chain[:30])
    with pytest.raises(ValueError):
        DeviceEnsembleSampler(20, 6, like, seed=1).run_mcmc(np.zeros((20, 6)), 5)      # degenerate ensemble
    with pytest.raises(TypeError):
        DeviceEnsembleSampler(20, 6, fitter.lnprob)


def test_device_sampler_matches_numpy_replay():
    """The device chain against the numpy restatement of rbv_stretch_run (oracle/stretch_replay.py: same Philox
    streams, same split, same update rule) driven by the GPU lnprob: identical proposals, identical decisions."""
    from oracle import stretch_replay as sr
    from rbvfit_b200.sampler import DeviceEnsembleSampler
    w, fitter, comp, theta0 = _c1_fitter()
    like = fitter._like
    rng = np.random.default_rng(8)
    for W in (16, 21):                               # even and odd ensembles
        p0 = np.clip(w["theta_true"] + 1e-3 * rng.standard_normal((W, 6)), w["lb"], w["ub"])
        dev = DeviceEnsembleSampler(W, 6, like, seed=77)
        dev.run_mcmc(p0, 60)
        chain, lps, nacc = sr.run(like.lnprob, p0, like.lnprob(p0), 60, dev._seed)
        assert np.allclose(dev.get_chain(), chain, rtol=0, atol=1e-9)
        assert np.max(np.abs(dev.get_log_prob() - lps) / np.abs(lps)) <= 1e-9
        assert np.array_equal(np.rint(dev.acceptance_fraction * 60).astype(int), nacc)


def test_curve_of_growth_grid_and_chi_squared_match_oracle():
    """Other callers of model_flux: the COG grid as one batch (compute_cog.py:23-184) and the
    reduced chi-squared of UnifiedResults.chi_squared (unified_results.py:305-369), against the CPU oracle."""
    from oracle import voigt_oracle as vo
    from rbvfit_b200.cog import compute_cog
    Nlist, blist = np.linspace(12.0, 20.0, 9), [5.0, 20.0, 60.0]
    cog = compute_cog(1215.67, Nlist, blist)
    assert cog.Wlist.shape == (9, 3) and str(np.atleast_1d(cog.st["name"])[0]).startswith("HI")
    ocfg = vo.OracleConfig()
    ocfg.add_system(0.0, "HI", [cog.st["wave"]], 1)
    om = vo.lower(ocfg, FWHM=None)
    wave = np.linspace(cog.st["wave"] - 5.0, cog.st["wave"] + 5.0, 1000)
    for i in (0, 4, 8):
        for j in range(3):
            flx = vo.model_flux(om, np.array([Nlist[i], blist[j], 0.0]), wave)
            ref = np.sum(np.diff(wave) * ((1 - flx)[1:] + (1 - flx)[:-1]) / 2.0)
            assert abs(cog.Wlist[i, j] - ref) <= 1e-9 * max(1.0, ref)
    assert np.all(np.diff(cog.Wlist, axis=0) > 0)                       # W grows with N on every b track
    w, fitter, comp, theta0 = _c1_fitter()
    got = fitter.chi_squared(w["theta_true"])
    d = fitter.instrument_data["COS"]
    m = vo.model_flux(comp["COS"]["model"], w["theta_true"], d["wave"]) if "model" in comp.get("COS", {}) else None
    if m is not None:
        ref = float(np.sum(((d["flux"] - m) / d["error"]) ** 2)) / (len(d["wave"]) - 6)
        assert abs(got["COS"] - ref) <= 1e-9 * ref
    assert 0.8 < got["COS"] < 1.25                                       # noise realisation at the truth
    with pytest.raises(ValueError):
        fitter.chi_squared(instrument_name="HIRES")


def test_fit_quick_recovers_truth_with_sane_errors():
    """vfit.fit_quick (quick_fit_interface.py:10-128) through device batches: the chi^2 objective equals the
    oracle's, the optimum sits near the truth the spectrum was drawn from, errors are finite and bracket it."""
    from oracle import voigt_oracle as vo
    w, fitter, comp, theta0 = _c1_fitter()
    d = comp["COS"]
    probe = np.vstack([w["theta_true"], theta0])
    ref = np.array([np.sum(((d["flux"] - vo.model_flux(d["model"], th, d["wave"])) / d["error"]) ** 2) for th in probe])
    assert np.max(np.abs(fitter._chi2_batch(probe) - ref) / ref) <= 1e-9
    best, err = fitter.fit_quick(verbose=False)
    assert best.shape == (6,) and err.shape == (6,) and np.all(np.isfinite(err)) and np.all(err > 0)
    assert np.all(best >= w["lb"]) and np.all(best <= w["ub"])
    assert fitter._chi2_batch(best)[0] <= fitter._chi2_batch(w["theta_true"])[0] + 1e-6
    # (the diagonal-curvature errors ignore the N1-N2 / v1-v2 degeneracy of the blended doublet, as in the reference)
    assert np.all(np.abs(best - w["theta_true"]) < np.array([0.5, 0.5, 5.0, 5.0, 10.0, 10.0]))
    assert fitter.mcmc_flag is False and np.array_equal(fitter.theta_best, best)


def test_device_sampler_with_separate_finalisation_launch():
    """Big grids form lnprob (and the sampler's accept/reject) in finalize_kernel instead of the walker's last CTA;
    RBVFIT_B200_FINALIZE=1 forces that path on a small problem: same chain as the numpy replay, bit-identical
    lnprob to the in-kernel path."""
    import os
    from oracle import stretch_replay as sr
    from rbvfit_b200.sampler import DeviceEnsembleSampler
    w, fitter0, comp, theta0 = _c1_fitter()
    rng = np.random.default_rng(9)
    p0 = np.clip(w["theta_true"] + 1e-3 * rng.standard_normal((18, 6)), w["lb"], w["ub"])
    ref = fitter0._like.lnprob(p0)
    os.environ["RBVFIT_B200_FINALIZE"] = "1"
    try:
        w, fitter, comp, theta0 = _c1_fitter()            # new context: reads the hook
        like = fitter._like
        assert np.array_equal(like.lnprob(p0), ref)
        dev = DeviceEnsembleSampler(18, 6, like, seed=31)
        dev.run_mcmc(p0, 40)
        chain, lps, nacc = sr.run(like.lnprob, p0, like.lnprob(p0), 40, dev._seed)
        assert np.allclose(dev.get_chain(), chain, rtol=0, atol=1e-9)
        assert np.array_equal(np.rint(dev.acceptance_fraction * 40).astype(int), nacc)
    finally:
        del os.environ["RBVFIT_B200_FINALIZE"]
        from rbvfit_b200.engine import Engine
        Engine(0).close()


def test_distributed_device_sampler_single_rank_and_multi_gpu():
    """Multi-GPU form of the sampler (rbv_stretch_propose_eval / all-gather / rbv_stretch_accept): with one rank it
    reproduces the fused single-GPU chain bit for bit (same random streams, same kernel on the same rows); with
    >= 2 GPUs the torchrun check (tools/check_dist_sampler.py) must pass on every rank."""
    import subprocess
    import sys
    import torch
    from rbvfit_b200.dist import WalkerPartition
    from rbvfit_b200.sampler import DeviceEnsembleSampler, DistributedDeviceSampler
    w, fitter, comp, theta0 = _c1_fitter()
    like = fitter._like
    rng = np.random.default_rng(12)
    for W in (20, 23):
        p0 = np.clip(w["theta_true"] + 1e-3 * rng.standard_normal((W, 6)), w["lb"], w["ub"])
        ref = DeviceEnsembleSampler(W, 6, like, seed=41)
        ref.run_mcmc(p0, 50)
        dsm = DistributedDeviceSampler(W, 6, like, WalkerPartition(0, 1), seed=41)
        dsm.run_mcmc(p0, 30)
        dsm.run_mcmc(None, 20)
        assert np.array_equal(dsm.get_chain(), ref.get_chain())
        assert np.array_equal(dsm.get_log_prob(), ref.get_log_prob())
        assert np.array_equal(dsm.acceptance_fraction, ref.acceptance_fraction)
    if torch.cuda.device_count() >= 2:
        import os
        root = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
        res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                              "--master-addr", "127.0.0.1", "--master-port", "29541",
                              os.path.join(root, "tools", "check_dist_sampler.py")],
                             capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


def test_device_slice_sampler_matches_numpy_replay():
    """rbv_slice_run against its numpy restatement (oracle/slice_replay.py: same Philox streams, same lockstep
    state machines, same mu adaptation) driven by the GPU lnprob: identical candidates, identical decisions, the
    same counters, in graph mode (WHILE node, device-set condition) and in host-polled mode, which launches exactly
    one masked batch per half-step on top."""
    from oracle import slice_replay as sl
    from rbvfit_b200.slice_sampler import DeviceEnsembleSliceSampler
    w, fitter, comp, theta0 = _c1_fitter()
    like = fitter._like
    rng = np.random.default_rng(8)
    for W, nsteps in ((12, 40), (16, 25)):
        p0 = np.clip(w["theta_true"] + 1e-3 * rng.standard_normal((W, 6)), w["lb"], w["ub"])
        dev = DeviceEnsembleSliceSampler(W, 6, like, seed=77, depth=1)
        dev.run_mcmc(p0, nsteps)
        ref = sl.run(like.lnprob, p0, like.lnprob(p0), nsteps, dev._seed)
        assert np.allclose(dev.get_chain(), ref["chain"], rtol=0, atol=1e-9)
        assert np.max(np.abs(dev.get_log_prob() - ref["lnp_chain"]) / np.abs(ref["lnp_chain"])) <= 1e-9
        assert (dev.nexp, dev.ncon) == (ref["nexp"], ref["ncon"])
        assert dev.ncall == W + ref["ncall"] and dev.nbatches == 1 + ref["nbatches"]
        assert dev.mu == ref["mu"] and dev.tune == ref["tune"] and np.array_equal(dev.mus, ref["mus"][:len(dev.mus)])
        poll = DeviceEnsembleSliceSampler(W, 6, like, seed=77, use_graph=False, depth=1)   # host-polled loop: same chain
        poll.run_mcmc(p0, nsteps)
        assert np.array_equal(poll.get_chain(), dev.get_chain()) and np.array_equal(poll.get_log_prob(),
                                                                                    dev.get_log_prob())
        assert (poll.mu, poll.ncall, poll.nexp, poll.ncon) == (dev.mu, dev.ncall, dev.nexp, dev.ncon)
        assert poll.nbatches == dev.nbatches + 2 * nsteps        # one masked batch per half-step
        # two logical iterations per launch (the default: the second one's candidates are speculative): the chain, mu
        # and every counter of the sequential algorithm bit for bit, in fewer launches -- graph mode and host-polled
        for graph in (True, False):
            spec = DeviceEnsembleSliceSampler(W, 6, like, seed=77, use_graph=graph)
            assert spec.depth == 2
            spec.run_mcmc(p0, nsteps)
            assert np.array_equal(spec.get_chain(), dev.get_chain())
            assert np.array_equal(spec.get_log_prob(), dev.get_log_prob())
            assert (spec.mu, spec.ncall, spec.nexp, spec.ncon, spec.tune) == (dev.mu, dev.ncall, dev.nexp, dev.ncon,
                                                                              dev.tune)
            assert np.array_equal(spec.mus, dev.mus)
            ref_batches = dev.nbatches if graph else poll.nbatches
            assert spec.nbatches < ref_batches and 2 * (spec.nbatches - 1) >= dev.nbatches - 1 - 2 * nsteps
        ref2 = sl.run(like.lnprob, p0, like.lnprob(p0), nsteps, dev._seed, depth=2)     # the replay's own depth 2
        graph2 = DeviceEnsembleSliceSampler(W, 6, like, seed=77)
        graph2.run_mcmc(p0, nsteps)
        assert graph2.nbatches == 1 + ref2["nbatches"] and graph2.ncall == W + ref2["ncall"]
        assert np.allclose(graph2.get_chain(), ref2["chain"], rtol=0, atol=1e-9)


def test_device_slice_sampler_bookkeeping_continuation_and_posterior():
    """rbv_slice_run: the recorded lnprob of every stored position equals the likelihood recomputed for it, walkers
    stay inside the prior box (out-of-bounds candidates are -inf, i.e. outside every slice), every walker moves
    every step, run(30) + run(None, 20) equals run(50), and through vfit(sampler='zeus') the posterior brackets the
    truth with zeus's accessor contract."""
    from rbvfit_b200.slice_sampler import DeviceEnsembleSliceSampler, EnsembleSliceSampler
    w, fitter, comp, theta0 = _c1_fitter()
    like = fitter._like
    rng = np.random.default_rng(3)
    p0 = np.clip(w["theta_true"] + 1e-3 * rng.standard_normal((14, 6)), w["lb"], w["ub"])
    a = DeviceEnsembleSliceSampler(14, 6, like, seed=5)
    a.run_mcmc(p0, 50)
    chain, lps = a.get_chain(), a.get_log_prob()
    assert chain.shape == (50, 14, 6) and lps.shape == (50, 14)
    for step in (0, 17, 49):
        assert np.array_equal(like.lnprob(chain[step]), lps[step])
    assert np.all(chain >= w["lb"]) and np.all(chain <= w["ub"]) and np.all(np.isfinite(lps))
    assert np.all(np.any(chain[1:] != chain[:-1], axis=2)) and a.acceptance_fraction.min() > 0.95
    b = DeviceEnsembleSliceSampler(14, 6, like, seed=5)
    b.run_mcmc(p0, 30)
    b.run_mcmc(None, 20)
    assert np.array_equal(b.get_chain(), chain) and np.array_equal(b.get_log_prob(), lps)
    assert (b.mu, b.ncall, b.nexp, b.ncon) == (a.mu, a.ncall, a.nexp, a.ncon)
    c = DeviceEnsembleSliceSampler(14, 6, like, seed=6)
    c.run_mcmc(p0, 10)
    assert not np.array_equal(c.get_chain(), chain[:10])
    assert a.efficiency > 0.05 and a.get_last_sample()[0].shape == (14, 6)
    with pytest.raises(ValueError):
        DeviceEnsembleSliceSampler(5, 6, like)                    # zeus: >= 2 * ndim walkers, an even number
    with pytest.raises(TypeError):
        DeviceEnsembleSliceSampler(14, 6, fitter.lnprob)
    bad = p0.copy()
    bad[0, 0] = w["ub"][0] + 1.0
    with pytest.raises(ValueError):
        DeviceEnsembleSliceSampler(14, 6, like, seed=1).run_mcmc(bad, 3)
    # through the fitter
    w2, f2, _, _ = _c1_fitter(nwalkers=24, nsteps=250, seed=13, sampler="zeus")
    f2.runmcmc(optimize=True, verbose=False, progress=False)
    s = f2.sampler
    assert isinstance(s, DeviceEnsembleSliceSampler) and s.get_chain().shape == (250, 24, 6)
    assert f2.samples.shape == (150 * 24, 6)
    lo, hi = np.percentile(f2.samples, [0.5, 99.5], axis=0)
    assert np.all(lo < w2["theta_true"]) and np.all(w2["theta_true"] < hi)
    tau = s.get_autocorr_time(quiet=True)
    assert tau.shape == (6,) and np.all(np.isfinite(tau))
    w3, f3, _, _ = _c1_fitter(nwalkers=24, nsteps=110, seed=13, sampler="zeus")     # burn-in of 100 steps is fixed
    f3.device_sampler = False
    f3.runmcmc(optimize=False, verbose=False, progress=False)
    assert type(f3.sampler) is EnsembleSliceSampler and f3.sampler.get_chain().shape == (110, 24, 6)


def test_sightline_sampler_equals_independent_replays():
    """rbv_stretch_run_sightlines (survey mode: one ensemble per sightline, lockstep on the device): every
    sightline's chain equals the numpy replay of ITS ensemble driven by its own single-sightline GPU likelihood
    (ensembles never mix, random streams offset per sightline), sightline 0 equals the single-ensemble device sampler,
    bookkeeping and continuation hold, and the per-sightline view has emcee's accessors."""
    from oracle import stretch_replay as sr
    from rbvfit_b200 import FitConfiguration, workloads as wl
    from rbvfit_b200.likelihood import GpuLikelihood, SightlineBatch
    from rbvfit_b200.model import GpuVoigtModel
    from rbvfit_b200.sampler import DeviceEnsembleSampler, SightlineEnsembleSampler
    S, nsteps = 5, 40
    sight, singles = [], []
    w0 = wl.c5b_sightline(0)
    for s in range(S):
        w = wl.c5b_sightline(s)
        cfg = FitConfiguration()
        for (z, ion, trans, comps) in w["systems"]:
            cfg.add_system(z=z, ion=ion, transitions=trans, components=comps)
        m = GpuVoigtModel(cfg, FWHM="6.5")
        c = m.compile()
        sp = wl.make_spectra(w, lambda n, th, wave: c.model_flux(th, wave))["COS"]
        sight.append(dict(model=m, **sp))
        singles.append(GpuLikelihood({"COS": dict(model=m, **sp)}, w["lb"], w["ub"]))
    batch = SightlineBatch(sight, w0["lb"], w0["ub"])
    nd = batch.ndim
    rng = np.random.default_rng(12)
    for W in (14, 9):                                    # even and odd ensembles
        p0 = np.clip(w0["theta_true"] + 1e-3 * rng.standard_normal((S, W, nd)), w0["lb"], w0["ub"])
        smp = SightlineEnsembleSampler(W, nd, batch, seed=31)
        smp.run_mcmc(p0, nsteps)
        chain, lps = smp.get_chain(), smp.get_log_prob()
        assert chain.shape == (nsteps, S, W, nd) and lps.shape == (nsteps, S, W)
        assert smp.get_chain(flat=True, discard=10).shape == (S, (nsteps - 10) * W, nd)
        assert np.all(chain >= w0["lb"]) and np.all(chain <= w0["ub"]) and np.all(np.isfinite(lps))
        got = batch.lnprob(chain[-1])
        assert np.max(np.abs(got - lps[-1]) / np.abs(got)) <= 1e-12
        for e in range(S):
            like = singles[e]
            ref_chain, ref_lps, nacc = sr.run(like.lnprob, p0[e], like.lnprob(p0[e]), nsteps, smp._seed,
                                              walker_offset=e * W)
            assert np.allclose(chain[:, e], ref_chain, rtol=0, atol=1e-9), e
            assert np.max(np.abs(lps[:, e] - ref_lps) / np.abs(ref_lps)) <= 1e-9
            assert np.array_equal(np.rint(smp.acceptance_fraction[e] * nsteps).astype(int), nacc)
        assert 0.15 < smp.acceptance_fraction.mean() < 0.9
        one = DeviceEnsembleSampler(W, nd, singles[0], seed=31)
        one.run_mcmc(p0[0], nsteps)
        assert np.allclose(one.get_chain(), chain[:, 0], rtol=0, atol=1e-9)
        view = smp.sightline(2)
        assert view.get_chain(discard=5, flat=True).shape == ((nsteps - 5) * W, nd)
        assert np.array_equal(view.get_chain(), chain[:, 2]) and view.acceptance_fraction.shape == (W,)
        assert view.chain.shape == (W, nsteps, nd) and view.get_autocorr_time(quiet=True).shape == (nd,)
        cont = SightlineEnsembleSampler(W, nd, batch, seed=31)
        cont.run_mcmc(p0, 25)
        cont.run_mcmc(None, nsteps - 25)
        assert np.array_equal(cont.get_chain(), chain) and np.array_equal(cont.acceptance_fraction,
                                                                          smp.acceptance_fraction)
    with pytest.raises(ValueError):
        SightlineEnsembleSampler(14, nd, batch).run_mcmc(p0[:2], 3)
    with pytest.raises(TypeError):
        SightlineEnsembleSampler(14, nd, singles[0])
    with pytest.raises(IndexError):
        smp.sightline(S)
