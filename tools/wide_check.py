#!/usr/bin/env python
"""Dev probe: the wide (small-job) kernels (PETMH_WIDE = 1: three warps per chain pair, 2: nine) against the throughput
kernel (0) on the same seeds -- every array must be bit-identical -- and, per variant, chains of a 5-chain run against the
first 5 chains of a 6-chain run (streams are keyed by chain id, not by chain count)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
g = "tests/golden/"
pr = np.load(g + "prior_stats_nROI48.npz"); ds = np.load(g + "dataset_s0.1.npz")
y = ds["tac_noisy_sampled"] / ds["dt"][None, None, :]


def run(C, nt, wide, draws, tune):
    os.environ["PETMH_WIDE"] = wide
    s = MHSampler(n_chains=C, max_tacs=nt, max_draws=draws, seed=11)
    s.set_frames(ds["time_vector"], ds["dt"]); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
    s.set_data(y[:nt], ds["vartacref"][:nt], ds["vark2p"][:nt], ds["sigma_noise"])
    s.run(draws=draws, tune=tune)
    dvr, r1 = s.chains(); q, sc = s.state()
    out = dict(dvr=dvr.copy(), r1=r1.copy(), q=q.copy(), sc=sc.copy(), summ=s.summary().copy())
    s.close()
    return out


for C, nt, draws, tune in [(2, 1, 4, 0), (2, 1, 40, 0), (5, 2, 40, 230), (64, 1, 40, 230)]:
    ref = run(C, nt, "0", draws, tune)
    for wide in ("1", "2"):
        got = run(C, nt, wide, draws, tune)
        for k in ref:
            a, b = ref[k], got[k]
            neq = ~((a == b) | (np.isnan(a) & np.isnan(b)))
            print("chains %d tacs %d draws %d tune %d wide %s %-4s: %d of %d differ%s" % (
                C, nt, draws, tune, wide, k, int(neq.sum()), a.size, (" first " + str(np.argwhere(neq)[:3].tolist())) if neq.any() else ""))
for wide in ("0", "1", "2"):
    for tune in (0, 100, 230):
        a5, a6 = run(5, 1, wide, 40, tune)["dvr"], run(6, 1, wide, 40, tune)["dvr"]
        print("wide %s tune %d: 5-chain run == first 5 chains of the 6-chain run: %s" % (wide, tune, np.array_equal(a5[0], a6[0, :5])))
