#!/usr/bin/env python
"""All-gather inside the library: peer-memory kernel against NCCL, same process, same launches (run under torchrun).

Two contexts per rank on the same workload, one attached to peer memory (rbv_peer_attach), one left on NCCL
(RBVFIT_B200_PEER=0 while its communicator is created); per workload: the two must agree bit for bit, then the
device time of `rbv_lnprob_batch_allgather` is measured with CUDA events over many calls (max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/peer_latency.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as tdist
    import bench
    from rbvfit_b200 import dist as rdist
    rank, world, local = rdist.init_from_env("nccl")
    torch.cuda.set_device(local)
    part = rdist.WalkerPartition(rank, world)
    out = {"world": world}
    for name, calls in (("C1", 400), ("C2", 400), ("C5a", 40)):
        w, _m, like_p, thetas, _s = bench.build_problem(name, local)
        _w, _m2, like_n, _t, _s2 = bench.build_problem(name, local)
        assert like_p.engine.comm_init()
        os.environ["RBVFIT_B200_PEER"] = "0"
        assert like_n.engine.comm_init()
        del os.environ["RBVFIT_B200_PEER"]
        assert like_p.engine.peer_attached and not like_n.engine.peer_attached
        th = torch.as_tensor(thetas, device=f"cuda:{local}")
        res = {}
        vals = {}
        for tag, like in (("peer", like_p), ("nccl", like_n)):
            dl = rdist.DistributedLikelihood(like, part)
            for _ in range(10):
                v = dl.lnprob_device(th)
            vals[tag] = v.clone()
            torch.cuda.synchronize()
            tdist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(calls):
                dl.lnprob_device(th)
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / calls * 1e3], device=f"cuda:{local}", dtype=torch.float64)
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
            res[tag] = float(t.item())
        assert torch.equal(vals["peer"].view(torch.int64), vals["nccl"].view(torch.int64)), "peer != nccl"
        res["rows"] = int(len(thetas))
        res["peer_error"] = like_p.engine.peer_error()
        out[name] = res
        like_p.close()
        like_n.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    tdist.barrier()
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
