#!/usr/bin/env python
"""Where a C5a-scale run_mcmc call spends its wall clock: the C call (graph capture + replay + sync) vs the chain
hand-off to the host.  Development tool."""
import os
import sys
import time

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from rbvfit_b200.sampler import DeviceEnsembleSampler  # noqa: E402

w, models, like, thetas, spectra = bench.build_problem("C5a", 0)
ok = thetas[np.all((thetas >= w["lb"]) & (thetas <= w["ub"]), axis=1)]
W = len(ok) - (len(ok) % 2)
smp = DeviceEnsembleSampler(W, like.ndim, like, seed=4)
smp.run_mcmc(ok[:W], 4, skip_initial_state_check=True)
eng = like.engine
for nsteps in (12, 48):
    for rep in range(2):
        coords_t, lnp_t = smp._state
        st = smp._stream
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(st):
            chain_t = torch.empty((nsteps, W, like.ndim), dtype=torch.float64, device="cuda")
            lps_t = torch.empty((nsteps, W), dtype=torch.float64, device="cuda")
            nacc_t = torch.zeros(W, dtype=torch.int32, device="cuda")
            flag_t = torch.zeros(1, dtype=torch.int32, device="cuda")
            t1 = time.perf_counter()
            eng.stretch_run(coords_t, lnp_t, nsteps, 2.0, 1234, 100, chain_t, lps_t, nacc_t, flag_t, use_graph=True)
            st.synchronize()
            t2 = time.perf_counter()
            chain = chain_t.cpu().numpy()
            lps = lps_t.cpu().numpy()
            t3 = time.perf_counter()
        print(f"nsteps={nsteps}: alloc {1e3 * (t1 - t0):.2f} ms, C call {1e3 * (t2 - t1):.2f} ms "
              f"({1e3 * (t2 - t1) / nsteps:.3f} ms/step), D2H {1e3 * (t3 - t2):.2f} ms "
              f"({chain.nbytes / 1e6:.0f} MB -> {chain.nbytes / 1e9 / (t3 - t2):.1f} GB/s)")
