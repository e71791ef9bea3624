"""Multi-GPU: one process per GPU, walkers partitioned across ranks, one tiny all-gather of lnprob per batch.

The path shards by independent units: lnprob of a walker depends only on its own
theta row, so rank r evaluates rows [r*ceil(W/G), (r+1)*ceil(W/G)) of every batch against its own replica
of the (MB-sized) spectra and line tables, and the only exchange is the all-gather of the W lnprob values
(8 B per walker) over NCCL / NVLink.  Every rank runs the same sampler from the same seed, so ensemble
state stays replicated and no theta ever crosses the wire.  Independent sightlines (survey mode) shard the
same way with no collective at all until the final gather (``SightlinePartition``).

``WalkerPartition`` holds the host-side logic (row ranges, padding, gather) and works with any
torch.distributed backend -- NCCL on the GPUs, gloo in the CPU tests, where a stub evaluator stands in for
the device call.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import numpy as np


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank).
    A single-process run (no WORLD_SIZE) returns (0, 1, 0) without initialising anything."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29531")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(local)
                dist.init_process_group(backend, rank=rank, world_size=world,
                                        device_id=torch.device("cuda", local))
            else:
                dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def replicate_seed(seed, rank: int, world: int, group=None) -> int:
    """The 64-bit sampler seed every rank must share: rank 0's (drawn from OS entropy when ``seed`` is None) is
    broadcast.  Device samplers run replicated from counter-based random streams -- ranks with different seeds would
    build different proposals and apply all-gathered lnprob values to rows they were not computed for."""
    s = int(np.random.SeedSequence(seed).generate_state(1, dtype=np.uint64)[0])
    if world == 1:
        return s
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([s >> 32, s & 0xFFFFFFFF], dtype=torch.int64, device=dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    hi, lo = (int(v) for v in t.cpu().tolist())
    return (hi << 32) | lo


def replicate_array(arr, rank: int, world: int, group=None) -> np.ndarray:
    """Rank 0's copy of a float64 array (initial ensemble) on every rank."""
    a = np.ascontiguousarray(arr, dtype=np.float64)
    if world == 1:
        return a
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.as_tensor(a).to(dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return t.cpu().numpy()


class WalkerPartition:
    """Row partition of a [W, ndim] batch over the ranks of a process group."""

    def __init__(self, rank: int, world: int, group=None):
        if not (0 <= rank < world):
            raise ValueError("bad rank/world")
        self.rank, self.world, self.group = rank, world, group

    def chunk(self, W: int) -> int:
        return (W + self.world - 1) // self.world

    def rows(self, W: int, rank: Optional[int] = None) -> Tuple[int, int]:
        r = self.rank if rank is None else rank
        c = self.chunk(W)
        lo = min(r * c, W)
        return lo, min(lo + c, W)

    def gather(self, local, W: int):
        """``local``: torch tensor with this rank's results (len = rows(W) span, any device).
        Returns the full length-W tensor on every rank."""
        import torch
        if self.world == 1:
            return local[:W]
        import torch.distributed as dist
        c = self.chunk(W)
        padded = local
        if local.numel() != c:
            padded = torch.full((c,), float("nan"), dtype=local.dtype, device=local.device)
            padded[: local.numel()] = local
        full = torch.empty(c * self.world, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, padded.contiguous(), group=self.group)
        return full[:W]

    def evaluate(self, theta, local_eval: Callable, W: Optional[int] = None):
        """``local_eval(theta_rows) -> tensor[len(rows)]``; returns the gathered length-W tensor."""
        W = int(theta.shape[0]) if W is None else W
        lo, hi = self.rows(W)
        return self.gather(local_eval(theta[lo:hi]), W)


class SightlinePartition:
    """Survey mode ("sightline-parallel"): rank r owns a contiguous block of the S independent
    sightlines, builds its own ``SightlineBatch`` / ``SightlineEnsembleSampler`` over them and never talks to the
    other ranks while sampling; ``gather`` is the one collective at the very end, for per-sightline summaries
    (posterior percentiles, acceptance fractions -- a few numbers per sightline, never the chains)."""

    def __init__(self, rank: int, world: int, group=None):
        if not (0 <= rank < world):
            raise ValueError("bad rank/world")
        self.rank, self.world, self.group = rank, world, group

    def block(self, n_sightlines: int) -> int:
        return (n_sightlines + self.world - 1) // self.world

    def owned(self, n_sightlines: int, rank: Optional[int] = None) -> Tuple[int, int]:
        """(first, count) of the sightlines of ``rank`` (default: this rank); trailing ranks may own none."""
        r = self.rank if rank is None else rank
        per = self.block(n_sightlines)
        first = min(r * per, n_sightlines)
        return first, min(per, n_sightlines - first)

    def gather(self, local, n_sightlines: int):
        """``local``: torch tensor [count, ...] of this rank's per-sightline summaries (any device); returns the
        [n_sightlines, ...] tensor on every rank, in sightline order."""
        import torch
        if self.world == 1:
            return local[:n_sightlines]
        import torch.distributed as dist
        per = self.block(n_sightlines)
        tail = tuple(local.shape[1:])
        padded = torch.full((per,) + tail, float("nan"), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
        full = torch.empty((per * self.world,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, padded.contiguous(), group=self.group)
        return full[:n_sightlines]


class DistributedLikelihood:
    """``lnprob(theta[W, ndim]) -> lnprob[W]`` over all ranks; same call signature as GpuLikelihood."""

    def __init__(self, likelihood, partition: WalkerPartition):
        self.like = likelihood
        self.part = partition
        self.ndim = likelihood.ndim
        self.total_pixels = likelihood.total_pixels
        self._theta_dev = None

    def lnprob_device(self, theta_t):
        """Device-resident: every rank holds the full theta tensor, evaluates its rows, gathers lnprob.  With a
        communicator attached to the engine (``Engine.comm_init``) the all-gather is issued by the library itself
        (``rbv_lnprob_batch_allgather``): one C call, no torch.distributed op on the hot path."""
        W = int(theta_t.shape[0])
        eng = getattr(self.like, "engine", None)
        if eng is not None and getattr(eng, "has_comm", False) and eng.comm_world == self.part.world:
            return eng.lnprob_allgather_device(theta_t)
        lo, hi = self.part.rows(W)
        local = self.like.lnprob_device(theta_t[lo:hi].contiguous()) if hi > lo else theta_t.new_empty(0)
        return self.part.gather(local, W)

    def lnprob(self, theta):
        """HOST theta (replicated on every rank) -> HOST lnprob; H2D of this rank's rows only."""
        import torch
        theta = np.asarray(theta, dtype=np.float64)
        single = theta.ndim == 1
        th = np.atleast_2d(theta)
        W = th.shape[0]
        if self.part.world == 1:
            out = self.like.lnprob(th)
            return float(out[0]) if single else out
        lo, hi = self.part.rows(W)
        eng = self.like.engine
        if getattr(eng, "has_comm", False) and eng.comm_world == self.part.world:
            # collective inside the library: only this rank's rows cross PCIe (into their place of the replicated
            # device buffer -- the library reads nothing else), then one C call evaluates and all-gathers
            eng._reserve(W, th.shape[1])
            if hi > lo:
                if th.flags.c_contiguous and eng.lib.rbv_host_pinned(th.ctypes.data):
                    eng._theta_dev[lo:hi].copy_(torch.from_numpy(th[lo:hi]), non_blocking=True)   # page-locked: in place
                else:
                    eng._theta_pin_np[lo:hi] = th[lo:hi]
                    eng._theta_dev[lo:hi].copy_(eng._theta_pin[lo:hi], non_blocking=True)
            out = eng.lnprob_allgather_device(eng._theta_dev[:W]).cpu().numpy()
            return float(out[0]) if single else out
        if hi > lo:
            eng._reserve(hi - lo, th.shape[1])
            eng._theta_pin_np[: hi - lo] = th[lo:hi]
            dev_rows = eng._theta_dev[: hi - lo]
            dev_rows.copy_(eng._theta_pin[: hi - lo], non_blocking=True)
            local = eng.lnprob_device(dev_rows)
        else:
            local = torch.empty(0, dtype=torch.float64, device=eng.tdev)
        full = self.part.gather(local, W)
        out = full.cpu().numpy()
        return float(out[0]) if single else out

    __call__ = lnprob
