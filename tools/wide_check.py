#!/usr/bin/env python
"""Dev probe: wide (small-job) kernel vs normal kernel, same seeds -> report differences."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
g = "tests/golden/"
pr = np.load(g + "prior_stats_nROI48.npz"); ds = np.load(g + "dataset_s0.1.npz")
y = ds["tac_noisy_sampled"] / ds["dt"][None, None, :]
for C, nt, draws, tune in [(2, 1, 4, 0), (2, 1, 40, 0), (5, 2, 40, 230), (64, 1, 40, 230)]:
    out = {}
    for wide in ("0", "1"):
        os.environ["PETMH_WIDE"] = wide
        s = MHSampler(n_chains=C, max_tacs=nt, max_draws=draws, seed=11)
        s.set_frames(ds["time_vector"], ds["dt"]); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
        s.set_data(y[:nt], ds["vartacref"][:nt], ds["vark2p"][:nt], ds["sigma_noise"])
        s.run(draws=draws, tune=tune)
        dvr, r1 = s.chains(); q, sc = s.state()
        out[wide] = dict(dvr=dvr.copy(), r1=r1.copy(), q=q.copy(), sc=sc.copy(), summ=s.summary().copy())
        s.close()
    for k in out["0"]:
        a, b = out["0"][k], out["1"][k]
        neq = ~((a == b) | (np.isnan(a) & np.isnan(b)))
        print(C, nt, draws, tune, k, "n_diff", int(neq.sum()), "of", a.size, "max|d|", float(np.nanmax(np.abs(a - b))) if neq.any() else 0.0,
              "first", np.argwhere(neq)[:3].tolist() if neq.any() else "")
