#!/usr/bin/env python
"""Short driver for ncu: C5b sightline batch (S sightlines x 64 walkers x 2048 px)."""
import os
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
batch, thetas = bench.build_sightlines(0, S, 0, 64)
Sx, Ws, nd = thetas.shape
th = torch.as_tensor(thetas.reshape(Sx * Ws, nd), device="cuda:0")
evs = []
for _ in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = batch.lnprob_device(th, Ws)
    b.record()
    evs.append((a, b))
torch.cuda.synchronize()
print("C5b", S, [round(a.elapsed_time(b), 3) for a, b in evs])
