#!/usr/bin/env python
"""Dev probe: chain-steps/s of the sweep kernel on GPU-generated (K4) training-style TACs -- the bench's data, where
some ROIs sit outside the Chebyshev range and take the exact-operator fallback.
Usage: python tools/perf_probe_synth.py [S C SWEEPS TUNE]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
from pet_posterior_distribution_b200 import sample_sim_data as gen

S = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
C = int(sys.argv[2]) if len(sys.argv) > 2 else 16
SW = int(sys.argv[3]) if len(sys.argv) > 3 else 100
TUNE = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
prior = gen.load_prior()
t, dtv = gen.frame_grid()
sig64 = gen.noise_table(np.random.default_rng(1234), 0.1, t, dtv)
s = MHSampler(n_chains=C, max_tacs=S, max_draws=0, seed=2026)
s.set_frames(t, dtv)
s.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
s.synth(S, 4321, prior["mu_tac_ref"], prior["Cov_tac_ref"], float(prior["mu_k2p"]), sig64)
g = s.synth_get()
k2a = float(prior["mu_k2p"]) * g["R1"] / g["DVR"]
_, (lo, hi) = s.cheb_operator(0)
oob = (k2a < lo) | (k2a > hi)
print("truth outside the Chebyshev range: %.2f %% of ROIs, %.1f %% of TACs; attempts max %d" % (100 * oob.mean(), 100 * oob.any(1).mean(), g["attempts"].max()))
s.reset(); s.plan(10 ** 6, TUNE, 1)
s.advance(TUNE)
ms, nl = s.last_kernel_ms()
print("tuning %d sweeps: %.1f ms -> %.3e chain-steps/s" % (TUNE, ms, S * C * 96 * TUNE / (ms * 1e-3)))
for rep in range(3):
    s.advance(SW)
    ms, nl = s.last_kernel_ms()
    print("S=%d C=%d sweeps=%d tune=%d: %.2f ms, %d launches -> %.3e chain-steps/s" % (S, C, SW, TUNE, ms, nl, S * C * 96 * SW / (ms * 1e-3)), flush=True)
q, sc = s.state()
k2a_q = float(prior["mu_k2p"]) * q[:, :, 48:] / q[:, :, :48]
oq = (k2a_q < lo) | (k2a_q > hi)
print("chain states outside the range: %.2f %% of (chain, ROI), %.1f %% of chains, %.1f %% of TACs" % (100 * oq.mean(), 100 * oq.any(2).mean(), 100 * oq.any(2).any(1).mean()))
