#!/usr/bin/env python
"""Dev probe: where the end-to-end step time goes (H2D set_data, sweeps, D2H summary)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
S, C = 131072, 16
g = "tests/golden/"
pr = np.load(g + "prior_stats_nROI48.npz"); ds = np.load(g + "dataset_s0.1.npz")
idx = np.arange(S) % 4
y = torch.empty((S, 48, 54), dtype=torch.float32, pin_memory=True); y.numpy()[:] = (ds["tac_noisy_sampled"] / ds["dt"][None, None, :]).astype(np.float32)[idx]
c = torch.empty((S, 54), dtype=torch.float32, pin_memory=True); c.numpy()[:] = ds["vartacref"].astype(np.float32)[idx]
k = torch.full((S,), 0.0126, dtype=torch.float32).pin_memory()
out = torch.empty((S, 96, 8), dtype=torch.float32, pin_memory=True)
s = MHSampler(n_chains=C, max_tacs=S, seed=1)
s.set_frames(ds["time_vector"], ds["dt"]); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
s.set_data(y.numpy()[:4], c.numpy()[:4], k.numpy()[:4], ds["sigma_noise"].astype(np.float32))
s.set_data_ptr(S, y.data_ptr(), c.data_ptr(), k.data_ptr(), None)
s.reset(); s.plan(10**6, 0, 1); s.advance(2); s.summary_ptr(out.data_ptr())
for rep in range(3):
    t0 = time.perf_counter(); s.set_data_ptr(S, y.data_ptr(), c.data_ptr(), k.data_ptr(), None)
    t1 = time.perf_counter(); s.advance(5)
    t2 = time.perf_counter(); s.summary_ptr(out.data_ptr())
    t3 = time.perf_counter()
    print("set_data %.1f ms (%.1f GB/s)  advance(5) %.1f ms  summary %.1f ms (%.1f GB/s)" % (
        1e3 * (t1 - t0), S * 10588 / (t1 - t0) / 1e9, 1e3 * (t2 - t1), 1e3 * (t3 - t2), S * 3072 / (t3 - t2) / 1e9))
