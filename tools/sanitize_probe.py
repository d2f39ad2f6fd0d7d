#!/usr/bin/env python
"""Dev probe for compute-sanitizer: one small call of every kernel family of libpetmh.so (throughput and wide sweep kernels,
taped mode, parity hooks, rank and moments diagnostics, cross-chain ESS, K4 generator with and without the test rule, the k2-free
SRTM sampler, the general-grid helpers, checkpoint round trip).  Usage (GPU box):
  compute-sanitizer --tool memcheck  python tools/sanitize_probe.py
  compute-sanitizer --tool racecheck python tools/sanitize_probe.py small"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
from pet_posterior_distribution_b200 import kinetic_model as km
small = len(sys.argv) > 1 and sys.argv[1] == "small"
g = "tests/golden/"
pr = np.load(g + "prior_stats_nROI48.npz"); ds = np.load(g + "dataset_s0.1.npz")
y = ds["tac_noisy_sampled"] / ds["dt"][None, None, :]
t, dt = ds["time_vector"], ds["dt"]


def sampler(C, nt, draws, wide):
    os.environ["PETMH_WIDE"] = wide
    s = MHSampler(n_chains=C, max_tacs=nt, max_draws=draws, seed=3)
    s.set_frames(t, dt); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
    s.set_data(y[:nt], ds["vartacref"][:nt], ds["vark2p"][:nt], ds["sigma_noise"])
    return s


n_sw = 6 if small else 24
for wide, C, nt in (("0", 5, 2), ("1", 4, 1), ("2", 2, 1)):
    s = sampler(C, nt, 16, wide)
    s.run(draws=16, tune=n_sw)
    s.chains(); s.summary(); s.summary_ext(); s.ess_cross_chain(); s.posterior_cov()
    blob = s.checkpoint(); s.restore(blob); s.advance(0)
    s.close()
    print("sweep kernel wide", wide, "ok", flush=True)
s = sampler(3, 2, 0, "0")
s.run(draws=12, tune=n_sw); s.summary()                      # moments mode
s.forward(0, ds["varDVR"][0], ds["varR1"][0]); s.loglik(1, ds["varDVR"][1], ds["varR1"][1]); s.operator(0); s.cheb_operator(1)
s.forward_srtm(0, ds["varDVR"][0], 0.0126 * ds["varR1"][0], ds["varR1"][0]); s.philox_raw(5, 1, 0)
rng = np.random.default_rng(0)
nt_ = 4
s.run_taped(0, rng.standard_normal((2, nt_, 2, 48)).astype(np.float32), np.log(1 - rng.random((2, nt_, 2, 48))).astype(np.float32),
            np.stack([np.stack([np.stack([rng.permutation(48) for _ in range(2)]) for _ in range(nt_)]) for _ in range(2)]).astype(np.uint8), tune=2)
s.sample_srtm(0.0126 * pr["mu_R1"], np.eye(48) * 1e-6, draws=3, tune=3)
print("hooks / taped / srtm ok", flush=True)
for rule in (None, 0.8):
    s.synth_test_rule(rule, pr["Cov_DVR"], pr["Cov_R1"], pr["Cov_tac_ref"])
    s.synth(2, 9, pr["mu_tac_ref"], pr["Cov_tac_ref"], float(pr["mu_k2p"]), ds["sigma_noise"]); s.synth_get()
s.close()
print("synth ok", flush=True)
E = np.exp(-np.linspace(0.004, 0.03, 5)[None, :] * t[:, None])
km.estimate_continuous_convolution(t, ds["vartacref"][0], E); km.estimate_continuous_convolution(t[:9], E[:9, 0], E[:9, 1], num_points_resample=15)
km.interp1d_linear_vec(np.array([t[0] - 1, t[0], 3.3, t[-1]]), t, E); km.SRTM.make_time_exponential(-np.linspace(0.004, 0.03, 5), t)
print("helpers ok", flush=True)
