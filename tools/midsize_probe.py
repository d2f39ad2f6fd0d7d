import os, sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
from pet_posterior_distribution_b200 import MHSampler
g = "tests/golden/"
pr = np.load(g + "prior_stats_nROI48.npz"); ds = np.load(g + "dataset_s0.1.npz")
y = ds["tac_noisy_sampled"] / ds["dt"][None, None, :]
for S, C in ((13, 256), (100, 16), (25, 128), (3, 1024)):
    idx = np.arange(S) % 4
    s = MHSampler(n_chains=C, max_tacs=S, max_draws=0, seed=1)
    s.set_frames(ds["time_vector"], ds["dt"]); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
    s.set_data(y[idx], ds["vartacref"][idx], ds["vark2p"][idx], ds["sigma_noise"])
    s.reset(); s.plan(10**6, 400, 1); s.advance(400)
    for rep in range(2):
        s.advance(200); ms, nl = s.last_kernel_ms()
    print("S=%d C=%d: 200 sweeps %.2f ms -> %.3e chain-steps/s" % (S, C, ms, S * C * 96 * 200 / (ms * 1e-3)), flush=True)
    s.close()
