#!/usr/bin/env python
"""Dev probe: chain-steps/s of the sweep kernel on a tiled golden dataset."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
C = int(sys.argv[2]) if len(sys.argv) > 2 else 16
SW = int(sys.argv[3]) if len(sys.argv) > 3 else 20
WARM = int(sys.argv[4]) if len(sys.argv) > 4 else 20
g = "tests/golden/"
pr = np.load(g + "prior_stats_nROI48.npz"); ds = np.load(g + "dataset_s0.1.npz")
n0 = ds["varDVR"].shape[0]
idx = np.arange(S) % n0
y = (ds["tac_noisy_sampled"] / ds["dt"][None, None, :]).astype(np.float32)[idx]
cr = ds["vartacref"].astype(np.float32)[idx]
k2p = ds["vark2p"].astype(np.float32)[idx]
s = MHSampler(n_chains=C, max_tacs=S, max_draws=0, seed=1)
s.set_frames(ds["time_vector"], ds["dt"]); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
s.set_data(y, cr, k2p, ds["sigma_noise"].astype(np.float32))
s.reset(); s.plan(10**6, WARM, 1)
s.advance(WARM)  # tuning sweeps (pymc table every 100), then steady-state draw sweeps are timed
for rep in range(3):
    s.advance(SW)
    ms, nl = s.last_kernel_ms()
    print("S=%d C=%d sweeps=%d: %.2f ms, %d launches -> %.3e chain-steps/s" % (S, C, SW, ms, nl, S * C * 96 * SW / (ms * 1e-3)), flush=True)
