#!/usr/bin/env python
"""Dev probe: seconds per 48-ROI posterior for small jobs (BASELINE configs[1])."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
g = "tests/golden/"
pr = np.load(g + "prior_stats_nROI48.npz"); ds = np.load(g + "dataset_s0.1.npz")
y = ds["tac_noisy_sampled"] / ds["dt"][None, None, :]
for C in (4, 64, 256):
    s = MHSampler(n_chains=C, max_tacs=1, max_draws=2000, seed=1)
    s.set_frames(ds["time_vector"], ds["dt"]); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
    s.set_data(y[:1], ds["vartacref"][:1], ds["vark2p"][:1], ds["sigma_noise"])
    s.run(draws=200, tune=200)
    t0 = time.perf_counter(); s.run(draws=2000, tune=4000); t1 = time.perf_counter(); sm = s.summary(); t2 = time.perf_counter()
    print("chains %3d: 6000 sweeps %.3f s (%.2e chain-steps/s), summary %.3f s, rhat max %.3f" % (C, t1 - t0, C * 96 * 6000 / (t1 - t0), t2 - t1, np.nanmax(sm[0, :, 5])))
