#!/usr/bin/env python
"""Dev probe: normal vs wide kernel over job sizes (chain pairs) -> where the automatic switch belongs."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
g = "tests/golden/"
pr = np.load(g + "prior_stats_nROI48.npz"); ds = np.load(g + "dataset_s0.1.npz")
y = ds["tac_noisy_sampled"] / ds["dt"][None, None, :]
for nt, C in [(1, 64), (1, 256), (4, 128), (4, 256), (5, 256), (8, 256), (12, 256), (16, 256), (32, 256), (1, 6), (100, 6)]:
    idx = [i % 4 for i in range(nt)]
    res = []
    for wide in ("0", "1", None):
        if wide is None: os.environ.pop("PETMH_WIDE", None)
        else: os.environ["PETMH_WIDE"] = wide
        s = MHSampler(n_chains=C, max_tacs=nt, max_draws=0, seed=1)
        s.set_frames(ds["time_vector"], ds["dt"]); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
        s.set_data(y[idx], ds["vartacref"][idx], ds["vark2p"][idx], ds["sigma_noise"])
        s.run(draws=100, tune=200)
        s.run(draws=400, tune=600)
        res.append(s.last_kernel_ms()[0])
        s.close()
    print("tacs %3d chains %4d pairs %5d: normal %.1f ms, wide %.1f ms, auto %.1f ms (1000 sweeps)" % (nt, C, nt * C // 2, res[0], res[1], res[2]))
