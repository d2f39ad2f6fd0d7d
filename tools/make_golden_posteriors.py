#!/usr/bin/env python
"""Golden inputs and posterior moments for the noise sweep of BASELINE configs[3] (sigma 0.05 / 0.1 / 0.2) and a second
TAC, from the CPU oracle only (no GPU):

  tests/golden/dataset_s0.05.npz, dataset_s0.2.npz     2 test-style TACs each (oracle/generator.py = restated
                                                       sample_sim_data.py, flag_testing_data = True)
  tests/golden/oracle_posterior_s{sigma}_tac{k}.npz    posterior mean / sd of DVR, R1 with their MCSEs from 16 free-running
                                                       chains of the fp64 C oracle (oracle/c/mh_oracle.c: restated PyMC
                                                       element-wise Metropolis), 5000 tune + 30000 draws each

tests/test_gpu_posterior.py compares the GPU sampler's moments on the same TACs with these (rms z in [0.7, 1.3] would be
exact calibration; the test bounds rms z and max |z|).  Run here (8 cores, ~6 min): python tools/make_golden_posteriors.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = os.path.join(ROOT, "tests", "golden")
TUNE, DRAWS, CHAINS = 5000, 30000, 16
CASES = (("0.05", 0), ("0.1", 2), ("0.2", 0), ("0.2", 1), ("0.1", 0))   # (the last one for tests/test_independent_posterior.py)


def dataset(sig):
    path = os.path.join(G, "dataset_s%s.npz" % sig)
    if not os.path.isfile(path):
        from oracle import generator
        pr = dict(np.load(os.path.join(G, "prior_stats_nROI48.npz")))
        ds = generator.generate(pr, 2, mean_sigma_noise=float(sig), test_style=True, seed=int(float(sig) * 1000),
                                reject_negative=False)   # (the fixtures predate the NaN rule of oracle.generator.mahalanobis_rule)
        keep = ("varDVR", "varR1", "vark2p", "vartacref", "tac_sampled", "tac_noisy_sampled", "mu_noise", "sigma_noise",
                "mean_sigma_noise", "time_vector", "dt", "seed")
        np.savez_compressed(path, **{k: np.asarray(ds[k]) for k in keep})
        print("wrote", path)
    return dict(np.load(path))


def main():
    from oracle import cmh, diagnostics as dg
    from oracle.logp import Model
    pr = np.load(os.path.join(G, "prior_stats_nROI48.npz"))
    only = [a for a in sys.argv[1:] if a != "--long"]     # e.g. "0.1:0" to (re)make one case
    long_run = "--long" in sys.argv[1:]                   # 64 chains instead of 16 -> oracle_posterior_long_*.npz (CPU-side checks only)
    chains = 64 if long_run else CHAINS
    for sig, tac in CASES:
        if only and "%s:%d" % (sig, tac) not in only:
            continue
        ds = dataset(sig)
        y = ds["tac_noisy_sampled"][tac] / ds["dt"][None, :]
        m = Model(ds["time_vector"], ds["vartacref"][tac], ds["vark2p"][tac], y, ds["sigma_noise"],
                  pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
        draws, _ = cmh.CModel(m).run_free(chains, TUNE, DRAWS, seed=4242 + tac + (77 if long_run else 0), keep=True)
        x = draws[:, TUNE:].reshape(chains, DRAWS, 96).astype(np.float64)
        out = dict(mean=x.mean(axis=(0, 1)), sd=x.std(axis=(0, 1), ddof=1),
                   mcse_mean=np.array([dg.mcse_mean(x[:, :, k]) for k in range(96)]),
                   mcse_sd=np.array([dg.mcse_sd(x[:, :, k]) for k in range(96)]),
                   rhat=np.array([dg.rhat_rank(x[:, :, k]) for k in range(96)]),
                   ess_bulk=np.array([dg.ess_bulk(x[:, :, k]) for k in range(96)]),
                   tune=TUNE, draws=DRAWS, chains=chains, tac=tac, sigma=float(sig))
        path = os.path.join(G, "oracle_posterior_%ss%s_tac%d.npz" % ("long_" if long_run else "", sig, tac))
        np.savez_compressed(path, **out)
        print("%s: rhat max %.3f  ess_bulk min %.0f  sd median %.4f" % (os.path.basename(path), out["rhat"].max(), out["ess_bulk"].min(),
                                                                         np.median(out["sd"])))


if __name__ == "__main__":
    main()
