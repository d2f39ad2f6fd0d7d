#!/usr/bin/env python
"""Dev probe: a sub-wave job with the config-2 test set (tools/run_config2.py prepare DIR): samples 0..12 x 256 chains,
6000 sweeps -- the per-GPU share of BASELINE configs[2] on 8 GPUs.  In a sub-wave job every launch waits for its slowest
CTA, so one ROI on the exact-operator fallback slows the whole job."""
import glob, os, pickle, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
root = sys.argv[1] if len(sys.argv) > 1 else "/tmp/c2"
d = pickle.load(open(glob.glob(os.path.join(root, "sim_data/nROI48/*_test/data_*.pik"))[0], "rb"))
pr = pickle.load(open(os.path.join(root, "prior_stats_nROI48.pik"), "rb"))
dt = np.asarray(d["dt"])
idx = list(range(13))
y = np.asarray(d["tac_noisy_sampled"])[idx] / dt[None, None, :]
s = MHSampler(n_chains=256, max_tacs=13, max_draws=0, seed=0)
s.set_frames(d["time_vector"], dt); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
s.set_data(y, np.asarray(d["vartacref"])[idx], np.asarray(d["vark2p"], np.float64)[idx], d["sigma_noise"])
s.run(draws=100, tune=100)
t0 = time.perf_counter(); s.run(draws=2000, tune=4000); t1 = time.perf_counter()
ms, nl = s.last_kernel_ms()
print("13 TACs x 256 chains x 6000 sweeps: %.3f s (sweep kernels %.1f ms, %d launches) -> %.3e chain-steps/s" % (t1 - t0, ms, nl, 13 * 256 * 96 * 6000 / (ms * 1e-3)))
