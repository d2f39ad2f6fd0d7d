#!/usr/bin/env python
"""Golden posterior moments from the CPU oracle (free-running restated pymc Metropolis,
numpy random tape): tests/golden/oracle_posterior_tac0.npz.  The GPU test compares its own
posterior means / SDs of the same TAC against these within 3 Monte-Carlo standard errors.

Run here (8 cores, ~1 min):  python tools/make_oracle_posterior.py
"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TUNE, DRAWS, CHAINS, TAC = 2000, 4000, 8, 0


def work(seed):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import mh
    from oracle.logp import Model
    g = os.path.join(ROOT, "tests", "golden")
    pr = np.load(os.path.join(g, "prior_stats_nROI48.npz"))
    ds = np.load(os.path.join(g, "dataset_s0.1.npz"))
    y = ds["tac_noisy_sampled"][TAC] / ds["dt"][None, :]
    m = Model(ds["time_vector"], ds["vartacref"][TAC], ds["vark2p"][TAC], y, ds["sigma_noise"],
              pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
    out = mh.run_chain_rng(m, TUNE, DRAWS, seed)
    return out["draws"][TUNE:].reshape(DRAWS, 96), out["scale"].reshape(96)


def main():
    from oracle import diagnostics as dg
    with mp.get_context("fork").Pool(min(CHAINS, os.cpu_count())) as pool:
        res = pool.map(work, [7000 + c for c in range(CHAINS)])
    x = np.stack([r[0] for r in res]).astype(np.float64)          # (chains, draws, 96)
    mean = x.mean(axis=(0, 1))
    sd = x.std(axis=(0, 1), ddof=1)
    mcse_mean = np.array([dg.mcse_mean(x[:, :, k]) for k in range(96)])
    mcse_sd = np.array([dg.mcse_sd(x[:, :, k]) for k in range(96)])
    rhat = np.array([dg.rhat_rank(x[:, :, k]) for k in range(96)])
    ess = np.array([dg.ess_bulk(x[:, :, k]) for k in range(96)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_posterior_tac0.npz"), mean=mean, sd=sd,
                        mcse_mean=mcse_mean, mcse_sd=mcse_sd, rhat=rhat, ess_bulk=ess, tune=TUNE, draws=DRAWS,
                        chains=CHAINS, tac=TAC, scale=np.stack([r[1] for r in res]))
    print("rhat max %.3f  ess_bulk min %.0f  sd median %.4f" % (rhat.max(), ess.min(), np.median(sd)))


if __name__ == "__main__":
    main()
