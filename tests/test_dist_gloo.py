"""CPU, world_size = 2, gloo: the walker partition + lnprob all-gather used by the multi-GPU path.
A stub evaluator stands in for the device call (this tests host logic, it is not a CPU fallback)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.normpath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from rbvfit_b200 import dist as rdist
    r, w, _ = rdist.init_from_env("gloo")
    part = rdist.WalkerPartition(r, w)
    theta = torch.arange(W * 3, dtype=torch.float64).reshape(W, 3)          # replicated on every rank
    seen = []

    def stub_eval(rows):
        seen.append(rows.shape[0])
        return (rows ** 2).sum(dim=1) + 0.5

    full = part.evaluate(theta, stub_eval)
    ret[rank] = (full.numpy().copy(), seen, part.rows(W))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("W", [8, 7, 1])
def test_walker_partition_gloo(W):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, W, ret), nprocs=world, join=True)
    theta = np.arange(W * 3, dtype=np.float64).reshape(W, 3)
    ref = (theta ** 2).sum(axis=1) + 0.5
    spans = []
    for r in range(world):
        full, seen, rows = ret[r]
        assert np.array_equal(full, ref)                  # every rank ends with the full vector
        assert seen == [rows[1] - rows[0]]                # ... having evaluated only its own rows
        spans.append(rows)
    assert spans[0][0] == 0 and spans[-1][1] == W and spans[0][1] == spans[1][0]


def test_partition_rows_cover_everything():
    from rbvfit_b200.dist import WalkerPartition
    for world in (1, 2, 3, 8):
        for W in (0, 1, 5, 8, 8192, 8191):
            got = []
            for r in range(world):
                lo, hi = WalkerPartition(r, world).rows(W)
                got.extend(range(lo, hi))
            assert got == list(range(W))


# ---------------------------------------------------------------------------------------------------------------
# DistributedDeviceSampler over gloo: the engine is replaced by a CPU stub that restates rbv_stretch_propose_eval /
# rbv_stretch_accept with oracle/stretch_replay.py (host logic under test: row partition of the half-steps, the
# in-place all-gather with padding, replicated state; NOT a CPU fallback of the product)
MU = np.array([1.0, -2.0, 0.5])
SIG = np.array([0.5, 2.0, 1.0])


def _gauss(x):
    return -0.5 * np.sum(((np.atleast_2d(x) - MU) / SIG) ** 2, axis=1)


class _StubEngine:
    def __init__(self):
        self.tdev = torch.device("cpu")
        self.rows_evaluated = 0

    def lnprob_device(self, theta_t):
        return torch.as_tensor(_gauss(theta_t.numpy()))

    def stretch_propose_eval(self, coords_t, a, seed, step, split, lo, hi, rows_t):
        from oracle import stretch_replay as sr
        self._half = sr.propose_half(coords_t.numpy(), seed, step, split, a)
        rows_t[lo:hi] = torch.as_tensor(_gauss(self._half[1][lo:hi]))
        self.rows_evaluated += hi - lo

    def stretch_accept(self, coords_t, lnp_t, a, seed, step, split, rows_t, chain_row_t, lps_row_t, nacc_t, flag_t):
        from oracle import stretch_replay as sr
        idx, q, fac = self._half
        nacc = nacc_t.numpy()
        sr.accept_half(coords_t.numpy(), lnp_t.numpy(), nacc, seed, step, split, idx, q, fac,
                       rows_t.numpy()[: len(idx)])
        for i in idx:
            chain_row_t[i] = coords_t[i]
            lps_row_t[i] = lnp_t[i]


class _StubLikelihood:
    ndim = 3

    def __init__(self):
        self.engine = _StubEngine()

    def lnprob(self, theta):
        return _gauss(theta)


def _sampler_worker(rank, world, port, W, nsteps, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.cuda
    torch.cuda.synchronize = lambda *a, **k: None            # no device in this test
    from rbvfit_b200 import dist as rdist
    from rbvfit_b200.sampler import DistributedDeviceSampler
    r, w, _ = rdist.init_from_env("gloo")
    like = _StubLikelihood()
    smp = DistributedDeviceSampler(W, 3, like, rdist.WalkerPartition(r, w), seed=77)
    p0 = MU + 0.1 * np.random.default_rng(5).standard_normal((W, 3))
    smp.run_mcmc(p0, nsteps - 4)
    smp.run_mcmc(None, 4)
    ret[rank] = (smp.get_chain().copy(), smp.get_log_prob().copy(), smp.acceptance_fraction.copy(), smp._seed,
                 like.engine.rows_evaluated)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("W", [12, 11])
def test_distributed_device_sampler_gloo(W):
    from oracle import stretch_replay as sr
    world, nsteps = 2, 14
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sampler_worker, args=(world, port, W, nsteps, ret), nprocs=world, join=True)
    p0 = MU + 0.1 * np.random.default_rng(5).standard_normal((W, 3))
    chain, lps, nacc = sr.run(_gauss, p0, _gauss(p0), nsteps, ret[0][3])
    for r in range(world):
        c, l, af, seed, rows = ret[r]
        assert np.array_equal(c, chain) and np.array_equal(l, lps)          # every rank: the single-process chain
        assert np.array_equal(np.rint(af * nsteps).astype(int), nacc)
    assert ret[0][4] + ret[1][4] == nsteps * W                              # each proposal evaluated exactly once
    assert ret[0][4] > 0 and ret[1][4] > 0


def _sightline_worker(rank, world, port, S, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from rbvfit_b200 import dist as rdist
    r, w, _ = rdist.init_from_env("gloo")
    part = rdist.SightlinePartition(r, w)
    first, count = part.owned(S)
    # stand-in for "sample my sightlines, summarise each": 3 percentiles x 2 parameters per sightline
    ids = torch.arange(first, first + count, dtype=torch.float64)
    local = ids[:, None, None] * 10.0 + torch.arange(6, dtype=torch.float64).reshape(3, 2)
    full = part.gather(local, S)
    ret[rank] = (full.numpy().copy(), (first, count))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("S", [8, 5, 1])
def test_sightline_partition_gloo(S):
    """Survey mode over 2 ranks: contiguous sightline blocks that cover 0..S-1 exactly once (the last rank may own
    fewer or none), no collective until the final gather of the per-sightline summaries, which arrive in sightline
    order on every rank."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sightline_worker, args=(world, port, S, ret), nprocs=world, join=True)
    ref = np.arange(S, dtype=np.float64)[:, None, None] * 10.0 + np.arange(6, dtype=np.float64).reshape(3, 2)
    covered = []
    for r in range(world):
        full, (first, count) = ret[r]
        assert full.shape == (S, 3, 2) and np.array_equal(full, ref)
        covered += list(range(first, first + count))
    assert covered == list(range(S))


# ---------------------------------------------------------------------------------------------------------------
# zeus-style slice sampling over several ranks: the host-driven EnsembleSliceSampler with the partitioned likelihood
# as its log-probability (every rank draws the same numpy random numbers from the shared seed, evaluates only its
# rows of each -- ragged, often tiny -- batch and gathers lnprob)
def _slice_worker(rank, world, port, W, nsteps, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from rbvfit_b200 import dist as rdist
    from rbvfit_b200.slice_sampler import EnsembleSliceSampler
    r, w, _ = rdist.init_from_env("gloo")
    part = rdist.WalkerPartition(r, w)
    rows = [0]

    def lnprob(theta):             # what DistributedLikelihood.lnprob does, with a stub in place of the device call
        th = torch.as_tensor(np.ascontiguousarray(theta))

        def local_eval(sub):
            rows[0] += sub.shape[0]
            return torch.as_tensor(_gauss(sub.numpy())) if sub.shape[0] else torch.empty(0, dtype=torch.float64)

        return part.evaluate(th, local_eval).numpy()

    smp = EnsembleSliceSampler(W, 3, lnprob, seed=11)
    p0 = MU + 0.1 * np.random.default_rng(6).standard_normal((W, 3))
    smp.run_mcmc(p0, nsteps)
    ret[rank] = (smp.get_chain().copy(), smp.get_log_prob().copy(), smp.mu, smp.ncall, rows[0])
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_slice_sampler_over_partitioned_likelihood_gloo():
    from rbvfit_b200.slice_sampler import EnsembleSliceSampler
    world, W, nsteps = 2, 8, 12
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_slice_worker, args=(world, port, W, nsteps, ret), nprocs=world, join=True)
    ref = EnsembleSliceSampler(W, 3, _gauss, seed=11)
    p0 = MU + 0.1 * np.random.default_rng(6).standard_normal((W, 3))
    ref.run_mcmc(p0, nsteps)
    for r in range(world):
        chain, lps, mu, ncall, rows = ret[r]
        assert np.array_equal(chain, ref.get_chain()) and np.array_equal(lps, ref.get_log_prob())
        assert mu == ref.mu and ncall == ref.ncall
    assert ret[0][4] + ret[1][4] == ref.ncall and ret[0][4] > 0 and ret[1][4] > 0     # every row evaluated once


def _replicate_worker(rank, world, port, W, nsteps, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.cuda
    torch.cuda.synchronize = lambda *a, **k: None            # no device in this test
    from rbvfit_b200 import dist as rdist
    from rbvfit_b200.sampler import DistributedDeviceSampler
    r, w, _ = rdist.init_from_env("gloo")
    s0 = rdist.replicate_seed(1000 + r, r, w)                # ranks ask for different seeds: rank 0's wins
    a0 = rdist.replicate_array(np.full((2, 3), float(r)), r, w)
    like = _StubLikelihood()
    smp = DistributedDeviceSampler(W, 3, like, rdist.WalkerPartition(r, w), seed=None)    # OS entropy on rank 0
    p0 = MU + 0.1 * np.random.default_rng(5 + r).standard_normal((W, 3))                  # rank-dependent start
    smp.run_mcmc(p0, nsteps)
    ret[rank] = (s0, a0, smp._seed, smp.get_chain().copy(), like.engine.rows_evaluated)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_seed_and_initial_state_are_replicated_from_rank0_gloo():
    """Device samplers run replicated from counter-based random streams: with seed=None (each rank would draw its
    own OS entropy) or rank-dependent initial states the ranks would apply all-gathered lnprob values to proposals
    they were not computed for.  Rank 0's seed and ensemble are broadcast instead; every rank ends with rank 0's
    chain, which is the single-process chain for that seed and start."""
    from oracle import stretch_replay as sr
    world, W, nsteps = 2, 10, 9
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_replicate_worker, args=(world, port, W, nsteps, ret), nprocs=world, join=True)
    assert ret[0][0] == ret[1][0] and ret[0][0] == int(np.random.SeedSequence(1000).generate_state(1, dtype=np.uint64)[0])
    assert np.array_equal(ret[0][1], np.zeros((2, 3))) and np.array_equal(ret[1][1], np.zeros((2, 3)))
    assert ret[0][2] == ret[1][2]
    assert np.array_equal(ret[0][3], ret[1][3])
    p0 = MU + 0.1 * np.random.default_rng(5).standard_normal((W, 3))         # rank 0's start
    chain, _lps, _nacc = sr.run(_gauss, p0, _gauss(p0), nsteps, ret[0][2])
    assert np.array_equal(ret[0][3], chain)
    assert ret[0][4] + ret[1][4] == nsteps * W
