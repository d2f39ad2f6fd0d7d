#!/usr/bin/env python
"""The reference's shipped sampler length (40 000 tune + 20 000 draws, file name it2.0e+04_brn4.0e+04) on one
golden TAC: wall time, convergence diagnostics and agreement with the CPU oracle's posterior moments."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
pr = np.load(os.path.join(g, "prior_stats_nROI48.npz")); ds = np.load(os.path.join(g, "dataset_s0.1.npz"))
ref = np.load(os.path.join(g, "oracle_posterior_tac0.npz"))
y = ds["tac_noisy_sampled"] / ds["dt"][None, None, :]
for C in (4, 64):
    s = MHSampler(n_chains=C, max_tacs=1, max_draws=20000, seed=2025)
    s.set_frames(ds["time_vector"], ds["dt"]); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
    s.set_data(y[:1], ds["vartacref"][:1], ds["vark2p"][:1], ds["sigma_noise"])
    s.run(draws=100, tune=100)                      # warm the context
    t0 = time.perf_counter(); s.run(draws=20000, tune=40000); t1 = time.perf_counter(); sm = s.summary()[0]; t2 = time.perf_counter()
    z = (sm[:, 0] - ref["mean"]) / np.sqrt(sm[:, 2].astype(np.float64) ** 2 + ref["mcse_mean"] ** 2)
    print("chains %2d: sampling %.2f s + diagnostics %.2f s | r_hat max %.4f  ess_bulk min %.0f  accept %.2f..%.2f | "
          "means vs CPU oracle: max|z| %.2f rms %.2f | sd ratio %.3f..%.3f" % (
              C, t1 - t0, t2 - t1, sm[:, 5].max(), sm[:, 3].min(), sm[:, 6].min(), sm[:, 6].max(),
              np.abs(z).max(), np.sqrt((z ** 2).mean()), (sm[:, 1] / ref["sd"]).min(), (sm[:, 1] / ref["sd"]).max()))
