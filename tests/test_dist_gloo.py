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
