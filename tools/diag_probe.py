#!/usr/bin/env python
"""Dev probe: time of the stored-draw diagnostics (K3: rank-normalised R-hat, bulk / tail / mean / sd ESS, MCSE, hdi) for one
TAC.  Usage: python tools/diag_probe.py [CHAINS [DRAWS]]   (e.g. 256 20000 = the reference's length at configs[2]'s width)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
g = "tests/golden/"
pr = np.load(g + "prior_stats_nROI48.npz"); ds = np.load(g + "dataset_s0.1.npz")
y = ds["tac_noisy_sampled"] / ds["dt"][None, None, :]
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
D = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
s = MHSampler(n_chains=C, max_tacs=1, max_draws=D, seed=1)
s.set_frames(ds["time_vector"], ds["dt"]); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
s.set_data(y[:1], ds["vartacref"][:1], ds["vark2p"][:1], ds["sigma_noise"])
s.run(draws=D, tune=4000)
for rep in range(3):
    s.advance(0)
    t0 = time.perf_counter(); sm = s.summary(); t1 = time.perf_counter()
    print("chains %d x %d draws: summary %.3f s, ess_bulk min %.0f, r_hat max %.3f" % (C, D, t1 - t0, sm[0, :, 3].min(), sm[0, :, 5].max()))
