"""Batched affine-invariant ensemble sampler with emcee's semantics and sampler-object contract.

The reference hands ``vfit.lnprob`` to ``emcee.EnsembleSampler(nwalkers, ndim, lnprob, pool=pool)``
(src/rbvfit/vfit_mcmc.py:408-423) and later reads ``get_chain(discard, flat)``, ``acceptance_fraction``
and ``get_autocorr_time()`` from it (:563-655; core/unified_results.py:163-299).  emcee is a third-party
dependency that is not vendored in the reference (requirements.txt:5, ``emcee>=3.0.0``) and is not
installed here, so the algorithm is restated from its published description (Goodman & Weare 2010;
Foreman-Mackey et al. 2013, emcee 3 ``RedBlueMove`` + ``StretchMove``):

  per step: shuffle the walkers into two halves; for each half S with complement C
      z_k   = ((a - 1) u_k + 1)^2 / a,  u_k ~ U(0,1),  a = 2
      Y_k   = C_{j(k)} - (C_{j(k)} - S_k) z_k,          j(k) uniform over the complement
      ln q  = (ndim - 1) ln z_k + lnp(Y_k) - lnp(S_k);  accept when ln U < ln q

The only structural change: ``lnp`` is evaluated for the whole half-ensemble in ONE call
(``vectorize=True`` contract) -- that call is the GPU batch.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from ._lib import RbvError


class AutocorrError(Exception):
    """Raised when the chain is too short for a reliable autocorrelation-time estimate (emcee's name)."""

    def __init__(self, tau, *args, **kwargs):
        self.tau = tau
        super().__init__(*args, **kwargs)


# ------------------------------------------------------------------------------------------ autocorrelation
def _next_pow_two(n):
    i = 1
    while i < n:
        i <<= 1
    return i


def function_1d(x):
    """Normalised autocorrelation function of a 1-D series (FFT based)."""
    x = np.atleast_1d(x)
    if x.ndim != 1:
        raise ValueError("invalid dimensions for 1D autocorrelation function")
    n = _next_pow_two(len(x))
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[: len(x)].real
    acf /= acf[0]
    return acf


def _auto_window(taus, c):
    m = np.arange(len(taus)) < c * taus
    if np.any(m):
        return int(np.argmin(m))
    return len(taus) - 1


def integrated_time(x, c=5, tol=50, quiet=False):
    """Integrated autocorrelation time per dimension; ``x`` is [nsteps, nwalkers, ndim] (Sokal windowing)."""
    x = np.atleast_1d(x)
    if x.ndim == 1:
        x = x[:, None, None]
    if x.ndim == 2:
        x = x[:, :, None]
    if x.ndim != 3:
        raise ValueError("invalid dimensions")
    n_t, n_w, n_d = x.shape
    tau_est = np.empty(n_d)
    windows = np.empty(n_d, dtype=int)
    for d in range(n_d):
        f = np.zeros(n_t)
        for k in range(n_w):
            f += function_1d(x[:, k, d])
        f /= n_w
        taus = 2.0 * np.cumsum(f) - 1.0
        windows[d] = _auto_window(taus, c)
        tau_est[d] = taus[windows[d]]
    flag = tol * tau_est > n_t
    if np.any(flag):
        msg = ("The chain is shorter than {0} times the integrated autocorrelation time for {1} parameter(s). "
               "Use this estimate with caution and run a longer chain!\n").format(tol, np.sum(flag))
        msg += "N/{0} = {1:.0f};\ntau: {2}".format(tol, n_t / tol, tau_est)
        if not quiet:
            raise AutocorrError(tau_est, msg)
    return tau_est


def walkers_independent(coords) -> bool:
    """emcee's initial-state check: the ensemble must span the parameter space."""
    if not np.all(np.isfinite(coords)):
        return False
    C = coords - np.mean(coords, axis=0)[None, :]
    C_colmax = np.amax(np.abs(C), axis=0)
    if np.any(C_colmax == 0):
        return False
    C /= C_colmax
    C_colsum = np.sqrt(np.sum(C ** 2, axis=0))
    C /= C_colsum
    return np.linalg.cond(C.astype(float)) <= 1e8


class _ChunkedHistory:
    """Chain / log-prob history as a list of per-run chunks, concatenated once when it is first read: appending the
    chain of a run must not re-copy everything sampled so far (a C5a-scale ensemble writes 2.3 MB per step)."""

    def _hist_get(self, name):
        chunks = self.__dict__["_hist_" + name]
        if len(chunks) > 1:
            chunks[:] = [np.concatenate(chunks, axis=0)]
        return chunks[0]

    def _hist_set(self, name, value):
        self.__dict__["_hist_" + name] = [value]

    def _hist_append(self, chain, lps):
        for name, arr in (("chain", chain), ("log_prob", lps)):
            chunks = self.__dict__["_hist_" + name]
            if len(chunks) == 1 and len(chunks[0]) == 0:
                chunks[:] = [arr]
            else:
                chunks.append(arr)

    _chain = property(lambda self: self._hist_get("chain"), lambda self, v: self._hist_set("chain", v))
    _log_prob = property(lambda self: self._hist_get("log_prob"), lambda self, v: self._hist_set("log_prob", v))


# ------------------------------------------------------------------------------------------ sampler
class EnsembleSampler(_ChunkedHistory):
    """Stretch-move ensemble sampler; ``log_prob_fn`` must map ``(n, ndim) -> (n,)``."""

    def __init__(self, nwalkers: int, ndim: int, log_prob_fn: Callable, a: float = 2.0, pool=None,
                 vectorize: bool = True, seed: Optional[int] = None):
        if nwalkers < 2 * ndim:
            # emcee refuses this only for the live check below; keep the same spirit
            pass
        if nwalkers % 2 != 0 and nwalkers < 2:
            raise ValueError("need at least two walkers")
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(ndim), float(a)
        self.log_prob_fn = log_prob_fn
        self.vectorize = vectorize
        self.pool = pool          # accepted for API compatibility; batches run on the device instead
        self._random = np.random.default_rng(seed)
        self.reset()

    def reset(self):
        self._chain = np.empty((0, self.nwalkers, self.ndim))
        self._log_prob = np.empty((0, self.nwalkers))
        self._accepted = np.zeros(self.nwalkers)
        self.iteration = 0
        self._last = None
        self.n_logp_calls = 0
        self.n_logp_rows = 0

    # -- likelihood plumbing
    def compute_log_prob(self, coords):
        if np.any(np.isinf(coords)):
            raise ValueError("At least one parameter value was infinite")
        if np.any(np.isnan(coords)):
            raise ValueError("At least one parameter value was NaN")
        if self.vectorize:
            lp = np.asarray(self.log_prob_fn(coords), dtype=np.float64)
        else:
            lp = np.array([float(self.log_prob_fn(c)) for c in coords])
        self.n_logp_calls += 1
        self.n_logp_rows += len(coords)
        if np.any(np.isnan(lp)):
            raise ValueError("Probability function returned NaN")
        return lp

    # -- one full step (two half-steps)
    def _step(self, coords, log_prob):
        nw, nd, a = self.nwalkers, self.ndim, self.a
        inds = np.arange(nw) % 2
        self._random.shuffle(inds)
        accepted = np.zeros(nw, dtype=bool)
        for split in range(2):
            S1 = inds == split
            s = coords[S1]
            c = coords[~S1]
            Ns, Nc = len(s), len(c)
            zz = ((a - 1.0) * self._random.random(Ns) + 1.0) ** 2.0 / a
            factors = (nd - 1.0) * np.log(zz)
            rint = self._random.integers(Nc, size=Ns)
            q = c[rint] - (c[rint] - s) * zz[:, None]
            new_lp = self.compute_log_prob(q)
            lnpdiff = factors + new_lp - log_prob[S1]
            acc = np.log(self._random.random(Ns)) < lnpdiff
            idx = np.flatnonzero(S1)[acc]
            coords[idx] = q[acc]
            log_prob[idx] = new_lp[acc]
            accepted[idx] = True
        return coords, log_prob, accepted

    def run_mcmc(self, initial_state, nsteps, progress=False, skip_initial_state_check=False, **_ignored):
        if initial_state is None:
            if self._last is None:
                raise ValueError("Cannot have `initial_state=None` if run_mcmc has never been called.")
            coords, log_prob = self._last
        else:
            coords = np.array(initial_state, dtype=np.float64, copy=True)
            if coords.shape != (self.nwalkers, self.ndim):
                raise ValueError("incompatible input dimensions {0}".format(coords.shape))
            if not skip_initial_state_check and not walkers_independent(coords):
                raise ValueError("Initial state has a large condition number. Make sure that your walkers are "
                                 "linearly independent for the best performance")
            log_prob = self.compute_log_prob(coords)
        chain = np.empty((nsteps, self.nwalkers, self.ndim))
        lps = np.empty((nsteps, self.nwalkers))
        it = range(nsteps)
        if progress:
            try:
                from tqdm import tqdm
                it = tqdm(it, total=nsteps)
            except ImportError:
                pass
        for i in it:
            coords, log_prob, acc = self._step(coords, log_prob)
            self._accepted += acc
            chain[i] = coords
            lps[i] = log_prob
        self._hist_append(chain, lps)
        self.iteration += nsteps
        self._last = (coords, log_prob)
        return coords, log_prob

    # -- emcee-shaped accessors (what UnifiedResults and vfit read)
    def _get(self, arr, discard=0, thin=1, flat=False):
        v = arr[discard + thin - 1:: thin]
        if flat:
            s = list(v.shape[1:])
            s[0] = np.prod(v.shape[:2])
            return v.reshape(s)
        return v

    def get_chain(self, discard=0, thin=1, flat=False):
        return self._get(self._chain, discard, thin, flat)

    def get_log_prob(self, discard=0, thin=1, flat=False):
        return self._get(self._log_prob, discard, thin, flat)

    @property
    def chain(self):
        return np.swapaxes(self._chain, 0, 1)

    @property
    def flatchain(self):
        return self.get_chain(flat=True)

    @property
    def lnprobability(self):
        return self._log_prob.T

    @property
    def acceptance_fraction(self):
        return self._accepted / max(self.iteration, 1)

    def get_autocorr_time(self, discard=0, thin=1, **kwargs):
        return thin * integrated_time(self.get_chain(discard=discard, thin=thin), **kwargs)

    def get_last_sample(self):
        return self._last


class DeviceEnsembleSampler(EnsembleSampler):
    """Same sampler, same accessor contract, but the whole loop runs on the GPU (``rbv_stretch_run``): proposals,
    the likelihood batch and accept/reject of every half-step are device kernels replayed from a CUDA graph, and
    the chain is copied to the host once per ``run_mcmc`` call.  ``likelihood`` is a ``GpuLikelihood``; random
    numbers come from a counter-based generator keyed by ``seed``, so a run is reproducible and can be continued
    (``run_mcmc(None, n)``) without changing the stream of a single long run."""

    def __init__(self, nwalkers: int, ndim: int, likelihood, a: float = 2.0, seed: Optional[int] = None,
                 use_graph: bool = True, **_ignored):
        if not hasattr(likelihood, "engine"):
            raise TypeError("DeviceEnsembleSampler needs a GpuLikelihood (the log-probability must run on the device)")
        if likelihood.ndim != ndim:
            raise ValueError(f"likelihood has ndim={likelihood.ndim}, sampler was given ndim={ndim}")
        self.likelihood = likelihood
        self.use_graph = bool(use_graph)
        self._seed = int(np.random.SeedSequence(seed).generate_state(1, dtype=np.uint64)[0])
        self._stream = None
        self._state = None
        super().__init__(nwalkers, ndim, likelihood.lnprob, a=a, seed=seed)

    def reset(self):
        super().reset()
        self._state = None

    def _append(self, chain_t, lps_t, nacc_t):
        """Device chain of one run -> host, appended without re-copying the first run.  (A pinned staging buffer was
        measured slower here: allocating 100 MB of page-locked memory per run costs more than the pageable copy.)"""
        chain = chain_t if isinstance(chain_t, np.ndarray) else chain_t.cpu().numpy()
        lps = lps_t if isinstance(lps_t, np.ndarray) else lps_t.cpu().numpy()
        nacc = nacc_t.cpu().numpy()
        self._accepted += nacc
        self._hist_append(chain, lps)
        n = chain.shape[0]
        self.iteration += n
        if n:
            self._last = (chain[-1].copy(), lps[-1].copy())
        return self._last

    def run_mcmc(self, initial_state, nsteps, progress=False, skip_initial_state_check=False, **_ignored):
        import torch
        eng = self.likelihood.engine
        dev = eng.tdev
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=dev)
        torch.cuda.current_stream(dev).synchronize()
        with torch.cuda.stream(self._stream):
            if initial_state is None:
                if self._state is None:
                    raise ValueError("Cannot have `initial_state=None` if run_mcmc has never been called.")
                coords_t, lnp_t = self._state
            else:
                coords = np.array(initial_state, dtype=np.float64, copy=True)
                if coords.shape != (self.nwalkers, self.ndim):
                    raise ValueError("incompatible input dimensions {0}".format(coords.shape))
                if np.any(np.isinf(coords)):
                    raise ValueError("At least one parameter value was infinite")
                if np.any(np.isnan(coords)):
                    raise ValueError("At least one parameter value was NaN")
                if not skip_initial_state_check and not walkers_independent(coords):
                    raise ValueError("Initial state has a large condition number. Make sure that your walkers are "
                                     "linearly independent for the best performance")
                coords_t = torch.as_tensor(coords, device=dev)
                lnp_t = eng.lnprob_device(coords_t)
                self.n_logp_calls += 1
                self.n_logp_rows += self.nwalkers
                if bool(torch.isnan(lnp_t).any()):
                    raise ValueError("Probability function returned NaN")
            nacc_t = torch.zeros(self.nwalkers, dtype=torch.int32, device=dev)
            flag_t = torch.zeros(1, dtype=torch.int32, device=dev)
            if self.use_graph and nsteps >= 4:
                # the chain streams to the host in blocks while the run goes on (rbv_stretch_run_sink)
                chain_t, lps_t = eng.stretch_run_to_host(coords_t, lnp_t, nsteps, self.a, self._seed, self.iteration,
                                                         nacc_t, flag_t)
            else:
                chain_t = torch.empty((nsteps, self.nwalkers, self.ndim), dtype=torch.float64, device=dev)
                lps_t = torch.empty((nsteps, self.nwalkers), dtype=torch.float64, device=dev)
                eng.stretch_run(coords_t, lnp_t, nsteps, self.a, self._seed, self.iteration, chain_t, lps_t, nacc_t,
                                flag_t, use_graph=self.use_graph)
            self._stream.synchronize()
            flag = int(flag_t.item())
            if flag & 4:
                raise RuntimeError("rbv_stretch_run: grid barrier timed out (cooperative launch not co-resident)")
            if flag & 1:
                raise ValueError("Probability function returned NaN")
            self._state = (coords_t, lnp_t)
            self.n_logp_calls += 2 * nsteps
            self.n_logp_rows += self.nwalkers * nsteps
            return self._append(chain_t, lps_t, nacc_t)


class DistributedDeviceSampler(DeviceEnsembleSampler):
    """The device-resident stretch move over several GPUs (one process per GPU): the ensemble state is replicated,
    every half-step each rank evaluates its rows of the proposals, the ranks all-gather the 8-byte lnprob values over
    NCCL / NVLink, and every rank applies the same accept/reject -- the random streams are counter based, so the
    chain is the one ``DeviceEnsembleSampler`` produces on a single GPU with the same seed, bit for bit (the launch
    geometry of a rank's share is chosen as for the whole half-step).  Nothing but lnprob values crosses the wire.

    With a communicator attached to the engine (``Engine.comm_init``, done here when ``torch.distributed`` runs on
    NCCL) the whole run is ONE C call: ``rbv_stretch_run_dist`` captures a step -- proposals, this rank's likelihood
    launch, the in-place NCCL all-gather, accept/reject -- in a CUDA graph and replays it; no Python and no host
    synchronisation inside the run.  Without one (gloo in the CPU tests, a stub engine) the half-step is driven from
    here: ``rbv_stretch_propose_eval`` -> ``torch.distributed.all_gather_into_tensor`` -> ``rbv_stretch_accept``.

    Every rank must run with the same seed and the same initial ensemble: ``seed`` (None = OS entropy) and
    ``initial_state`` are taken from rank 0 and broadcast."""

    def __init__(self, nwalkers: int, ndim: int, likelihood, partition, a: float = 2.0, seed: Optional[int] = None,
                 use_graph: bool = True, **_ignored):
        super().__init__(nwalkers, ndim, likelihood, a=a, seed=seed, use_graph=use_graph)
        self.partition = partition
        from .dist import replicate_seed
        self._seed = replicate_seed(seed, partition.rank, partition.world, partition.group)
        eng = likelihood.engine
        if partition.world > 1 and not getattr(eng, "has_comm", True):
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_backend(partition.group) == "nccl":
                eng.comm_init(partition.group)

    def run_mcmc(self, initial_state, nsteps, progress=False, skip_initial_state_check=False, **_ignored):
        import torch
        from .dist import replicate_array
        eng = self.likelihood.engine
        dev = eng.tdev
        part = self.partition
        W, nd = self.nwalkers, self.ndim
        in_library = bool(getattr(eng, "has_comm", False)) and eng.comm_world == part.world
        if in_library and self._stream is None:
            self._stream = torch.cuda.Stream(device=dev)
        if initial_state is None:
            if self._state is None:
                raise ValueError("Cannot have `initial_state=None` if run_mcmc has never been called.")
            coords_t, lnp_t = self._state
        else:
            coords = np.array(initial_state, dtype=np.float64, copy=True)
            if coords.shape != (W, nd):
                raise ValueError("incompatible input dimensions {0}".format(coords.shape))
            coords = replicate_array(coords, part.rank, part.world, part.group)      # rank 0's ensemble everywhere
            if not np.all(np.isfinite(coords)):
                raise ValueError("At least one parameter value was infinite or NaN")
            if not skip_initial_state_check and not walkers_independent(coords):
                raise ValueError("Initial state has a large condition number. Make sure that your walkers are "
                                 "linearly independent for the best performance")
            coords_t = torch.as_tensor(coords, device=dev)
            if in_library:
                lnp_t = eng.lnprob_allgather_device(coords_t).contiguous()
            else:
                lo, hi = part.rows(W)
                local = eng.lnprob_device(coords_t[lo:hi].contiguous()) if hi > lo else coords_t.new_empty(0)
                lnp_t = part.gather(local, W).contiguous()
            if bool(torch.isnan(lnp_t).any()):
                raise ValueError("Probability function returned NaN")
        if in_library:
            torch.cuda.current_stream(dev).synchronize()
            with torch.cuda.stream(self._stream):
                nacc_t = torch.zeros(W, dtype=torch.int32, device=dev)
                flag_t = torch.zeros(1, dtype=torch.int32, device=dev)
                if self.use_graph and nsteps >= 4:
                    chain_t, lps_t = eng.stretch_run_to_host(coords_t, lnp_t, nsteps, self.a, self._seed,
                                                             self.iteration, nacc_t, flag_t, distributed=True)
                else:
                    chain_t = torch.empty((nsteps, W, nd), dtype=torch.float64, device=dev)
                    lps_t = torch.empty((nsteps, W), dtype=torch.float64, device=dev)
                    eng.stretch_run_dist(coords_t, lnp_t, nsteps, self.a, self._seed, self.iteration, chain_t, lps_t,
                                         nacc_t, flag_t, use_graph=self.use_graph)
                self._stream.synchronize()
            if getattr(eng, "peer_attached", False) and eng.peer_error():
                raise RbvError("the peer-memory all-gather gave up waiting for a rank (30 s without its rows): a "
                               "rank died or ran a different sequence of calls; rows it did not receive are NaN")
            if int(flag_t.item()) & 1:
                raise ValueError("Probability function returned NaN")
            self._state = (coords_t, lnp_t)
            self.n_logp_calls += 2 * nsteps
            self.n_logp_rows += W * nsteps
            return self._append(chain_t, lps_t, nacc_t)
        chain_t = torch.empty((nsteps, W, nd), dtype=torch.float64, device=dev)
        lps_t = torch.empty((nsteps, W), dtype=torch.float64, device=dev)
        nacc_t = torch.zeros(W, dtype=torch.int32, device=dev)
        flag_t = torch.zeros(1, dtype=torch.int32, device=dev)
        h = (W + 1) // 2
        # one buffer for the proposals' lnprob: rank r owns rows [r c, (r + 1) c) of a half-step, which is also its
        # slot of the all-gather output, so the collective runs in place and nothing is allocated inside the loop
        c_max = part.chunk(h)
        rows_t = torch.empty(c_max * part.world, dtype=torch.float64, device=dev)
        for s in range(nsteps):
            step = self.iteration + s
            for split in (0, 1):
                nS = h if split == 0 else W - h
                lo, hi = part.rows(nS)
                c = part.chunk(nS)
                eng.stretch_propose_eval(coords_t, self.a, self._seed, step, split, lo, hi, rows_t)
                if part.world > 1:
                    import torch.distributed as dist
                    dist.all_gather_into_tensor(rows_t[: c * part.world], rows_t[part.rank * c:(part.rank + 1) * c],
                                                group=part.group)
                eng.stretch_accept(coords_t, lnp_t, self.a, self._seed, step, split, rows_t, chain_t[s], lps_t[s],
                                   nacc_t, flag_t)
        torch.cuda.synchronize(dev)
        if int(flag_t.item()) & 1:
            raise ValueError("Probability function returned NaN")
        self._state = (coords_t, lnp_t)
        self.n_logp_calls += 2 * nsteps
        self.n_logp_rows += W * nsteps
        return self._append(chain_t, lps_t, nacc_t)


class _SightlineView:
    """One sightline of a ``SightlineEnsembleSampler`` behind emcee's accessor contract (what ``vfit`` /
    ``UnifiedResults`` read from ``fitter.sampler``, core/unified_results.py:163-299)."""

    def __init__(self, parent, index: int):
        self._p, self._s = parent, int(index)
        self.nwalkers, self.ndim = parent.nwalkers, parent.ndim

    @property
    def iteration(self):
        return self._p.iteration

    def get_chain(self, discard=0, thin=1, flat=False):
        return self._p.get_chain(discard=discard, thin=thin, flat=flat, sightline=self._s)

    def get_log_prob(self, discard=0, thin=1, flat=False):
        return self._p.get_log_prob(discard=discard, thin=thin, flat=flat, sightline=self._s)

    @property
    def chain(self):
        return np.swapaxes(self.get_chain(), 0, 1)

    @property
    def flatchain(self):
        return self.get_chain(flat=True)

    @property
    def lnprobability(self):
        return self.get_log_prob().T

    @property
    def acceptance_fraction(self):
        return self._p.acceptance_fraction[self._s]

    def get_autocorr_time(self, discard=0, thin=1, **kwargs):
        return thin * integrated_time(self.get_chain(discard=discard, thin=thin), **kwargs)


class SightlineEnsembleSampler(_ChunkedHistory):
    """Survey mode (BASELINE.json config 5): one stretch-move ensemble PER SIGHTLINE of a ``SightlineBatch``, all S
    ensembles advancing in lockstep on the device (``rbv_stretch_run_sightlines``) -- the reference would run S
    separate ``vfit(...).runmcmc()`` calls one after the other (vfit_mcmc.py:492-561).  Per half-step: one proposal
    kernel over all S * W/2 rows, ONE sightline likelihood launch, one accept kernel; no host round trip, the chain
    is copied to the host once per ``run_mcmc`` call.  Ensembles never mix.  Sightline 0 reproduces the chain of a
    ``DeviceEnsembleSampler`` with the same seed on that sightline alone (same random streams).  Sightlines shard
    across GPUs with no collective: every rank runs this sampler over its own ``SightlineBatch``.

    ``sightline(s)`` returns a per-sightline object with emcee's accessors (``get_chain(discard, thin, flat)``,
    ``acceptance_fraction``, ``get_autocorr_time`` ...)."""

    def __init__(self, nwalkers: int, ndim: int, batch, a: float = 2.0, seed: Optional[int] = None):
        if not hasattr(batch, "engine") or not hasattr(batch, "n_sightlines"):
            raise TypeError("SightlineEnsembleSampler needs a SightlineBatch")
        if batch.ndim != ndim:
            raise ValueError(f"the sightline batch has ndim={batch.ndim}, sampler was given ndim={ndim}")
        if nwalkers < 2:
            raise ValueError("need at least two walkers per sightline")
        self.batch = batch
        self.n_sightlines = int(batch.n_sightlines)
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(ndim), float(a)
        self._seed = int(np.random.SeedSequence(seed).generate_state(1, dtype=np.uint64)[0])
        self._stream = None
        self.reset()

    def reset(self):
        S, W, nd = self.n_sightlines, self.nwalkers, self.ndim
        self._chain = np.empty((0, S, W, nd))
        self._log_prob = np.empty((0, S, W))
        self._accepted = np.zeros((S, W))
        self.iteration = 0
        self._state = None
        self._last = None

    def run_mcmc(self, initial_state, nsteps, progress=False, skip_initial_state_check=False, **_ignored):
        import torch
        eng = self.batch.engine
        dev = eng.tdev
        S, W, nd = self.n_sightlines, self.nwalkers, self.ndim
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=dev)
        torch.cuda.current_stream(dev).synchronize()
        with torch.cuda.stream(self._stream):
            if initial_state is None:
                if self._state is None:
                    raise ValueError("Cannot have `initial_state=None` if run_mcmc has never been called.")
                coords_t, lnp_t = self._state
            else:
                coords = np.array(initial_state, dtype=np.float64, copy=True)
                if coords.shape != (S, W, nd):
                    raise ValueError("incompatible input dimensions {0}".format(coords.shape))
                if not np.all(np.isfinite(coords)):
                    raise ValueError("At least one parameter value was infinite or NaN")
                if not skip_initial_state_check and not all(walkers_independent(c) for c in coords):
                    raise ValueError("Initial state has a large condition number. Make sure that your walkers are "
                                     "linearly independent for the best performance")
                coords_t = torch.as_tensor(coords, device=dev)
                lnp_t = eng.lnprob_sightlines_device(coords_t.view(S * W, nd), W).view(S, W)
                if bool(torch.isnan(lnp_t).any()):
                    raise ValueError("Probability function returned NaN")
            chain_t = torch.empty((nsteps, S, W, nd), dtype=torch.float64, device=dev)
            lps_t = torch.empty((nsteps, S, W), dtype=torch.float64, device=dev)
            nacc_t = torch.zeros((S, W), dtype=torch.int32, device=dev)
            flag_t = torch.zeros(1, dtype=torch.int32, device=dev)
            eng.stretch_run_sightlines(coords_t, lnp_t, nsteps, self.a, self._seed, self.iteration, chain_t, lps_t,
                                       nacc_t, flag_t)
            self._stream.synchronize()
            if int(flag_t.item()) & 1:
                raise ValueError("Probability function returned NaN")
            self._state = (coords_t, lnp_t)
            chain, lps = chain_t.cpu().numpy(), lps_t.cpu().numpy()
            self._accepted += nacc_t.cpu().numpy()
        self._hist_append(chain, lps)
        self.iteration += nsteps
        if nsteps:
            self._last = (chain[-1].copy(), lps[-1].copy())
        return self._last

    def _get(self, arr, discard, thin, flat, sightline):
        v = arr[discard + thin - 1:: thin]
        if sightline is not None:
            v = v[:, int(sightline)]
            if flat:
                return v.reshape((-1,) + v.shape[2:])
            return v
        if flat:                                  # [S, steps * W, ...]: flat per sightline, never across sightlines
            v = np.swapaxes(v, 0, 1)
            return v.reshape((v.shape[0], -1) + v.shape[3:])
        return v

    def get_chain(self, discard=0, thin=1, flat=False, sightline=None):
        """[steps, S, W, ndim]; ``sightline=s`` -> [steps, W, ndim]; ``flat`` merges the steps and walkers of each
        sightline (never across sightlines)."""
        return self._get(self._chain, discard, thin, flat, sightline)

    def get_log_prob(self, discard=0, thin=1, flat=False, sightline=None):
        return self._get(self._log_prob, discard, thin, flat, sightline)

    @property
    def acceptance_fraction(self):
        """[S, W]"""
        return self._accepted / max(self.iteration, 1)

    def sightline(self, index: int) -> _SightlineView:
        if not 0 <= int(index) < self.n_sightlines:
            raise IndexError(index)
        return _SightlineView(self, index)

    def get_last_sample(self):
        return self._last
