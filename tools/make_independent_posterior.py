#!/usr/bin/env python
"""Posterior moments of the reference's model from a sampler that shares NOTHING with the path's algorithm -- a check that
the restated element-wise Metropolis (oracle/mh.py, oracle/c/mh_oracle.c) and therefore the GPU sampler draw from the
posterior that mcmc.py:147-155 defines, independent of PyMC's internals (tuning only changes efficiency: any valid
Metropolis kernel with frozen scaling has the model's posterior as its stationary distribution).

  target   log p(DVR, R1 | y) exactly as mcmc.py:147-155 states it, built from third-party / reference code only:
             scipy.stats.multivariate_normal(mu, Cov).logpdf          (pm.MvNormal, :148-149)
             the LIVE reference forward model kinetic_model.SRTM2(...).create_activity_curve(DVR, R1, k2p).T   (:27-39, :151)
             numpy.where(sn < 0, 1e-6, sn)                             (pt.switch, :152)
             scipy.stats.truncnorm(a=(0 - sn)/sigma, b=inf, loc=sn, scale=sigma).logpdf(y), sigma = sqrt(sn) sigma_noise   (:153-155)
           no formula of oracle/logp.py, no operator form, no cached terms.
  sampler  random-walk Metropolis on all 96 coordinates at once, x' = x + 2.38/sqrt(96) L xi with L L^T a FIXED pilot
           covariance (from a short run of the C oracle -- a proposal shape only: it cannot change the stationary law),
           one chain per core, over-dispersed starts, burn-in discarded, thinned.

Writes tests/golden/independent_posterior_s{sigma}_tac{k}.npz (mean, sd, their MCSEs, r_hat, ESS, acceptance rate).
tests/test_independent_posterior.py compares the oracle's golden posteriors with them; the GPU sampler is compared with the
oracle's in tests/test_gpu_posterior.py.  Needs /root/reference (build container only); ~25 min per case on 8 cores.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
G = os.path.join(ROOT, "tests", "golden")
CASES = (("0.1", 0, "oracle_posterior_s0.1_tac0.npz"), ("0.2", 0, "oracle_posterior_s0.2_tac0.npz"))
STEPS, BURN, THIN = 1030000, 30000, 50


def load_case(sig, tac):
    pr = np.load(os.path.join(G, "prior_stats_nROI48.npz"))
    ds = np.load(os.path.join(G, "dataset_s%s.npz" % sig))
    return pr, ds


def make_logpost(sig, tac):
    import kinetic_model as km                      # the live reference
    from scipy import stats
    pr, ds = load_case(sig, tac)
    t, dt = ds["time_vector"], ds["dt"]
    y = ds["tac_noisy_sampled"][tac] / dt[None, :]                                   # mcmc.py:79-80,109
    sigma_noise = ds["sigma_noise"]
    k2p = float(ds["vark2p"][tac])
    model = km.SRTM2(frame_time_list=t, frame_duration_list=dt, tac_reference=ds["vartacref"][tac])   # mcmc.py:133-134
    mv = (stats.multivariate_normal(pr["mu_DVR"], pr["Cov_DVR"]), stats.multivariate_normal(pr["mu_R1"], pr["Cov_R1"]))

    def logpost(x):
        sn = model.create_activity_curve(DVR=x[:48], R1=x[48:], k2p=k2p).T
        sn = np.where(sn < 0, 1e-6, sn)
        s = np.sqrt(sn) * sigma_noise
        with np.errstate(all="ignore"):
            ll = stats.truncnorm.logpdf(y, (0 - sn) / s, np.inf, loc=sn, scale=s).sum()
        return mv[0].logpdf(x[:48]) + mv[1].logpdf(x[48:]) + ll

    return logpost


def chain(args):
    sig, tac, seed, mean0, L = args
    logpost = make_logpost(sig, tac)
    rng = np.random.default_rng(seed)
    x = mean0 + 3.0 * (L @ rng.standard_normal(96))                 # over-dispersed start
    lp = logpost(x)
    while not np.isfinite(lp):
        x = mean0 + 3.0 * (L @ rng.standard_normal(96))
        lp = logpost(x)
    step = 2.38 / np.sqrt(96.0)
    keep = np.empty(((STEPS - BURN) // THIN, 96))
    nacc = 0
    for it in range(STEPS):
        xp = x + step * (L @ rng.standard_normal(96))
        lpp = logpost(xp)
        if np.isfinite(lpp) and np.log(rng.random()) < lpp - lp:
            x, lp = xp, lpp
            nacc += it >= BURN
        if it >= BURN and (it - BURN) % THIN == THIN - 1:
            keep[(it - BURN) // THIN] = x
    return keep, nacc / (STEPS - BURN)


def pilot(sig, tac):
    """Proposal shape: empirical covariance of a short run of the C oracle (8 chains x 8000 draws after 4000 tuning sweeps)."""
    from oracle import cmh
    from oracle.logp import Model
    pr, ds = load_case(sig, tac)
    y = ds["tac_noisy_sampled"][tac] / ds["dt"][None, :]
    m = Model(ds["time_vector"], ds["vartacref"][tac], ds["vark2p"][tac], y, ds["sigma_noise"],
              pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
    draws, _ = cmh.CModel(m).run_free(8, 4000, 8000, seed=99, keep=True)
    x = draws[:, 4000:].reshape(-1, 96).astype(np.float64)
    return x.mean(0), np.linalg.cholesky(np.cov(x, rowvar=False) + 1e-12 * np.eye(96))


def main():
    from oracle import diagnostics as dg
    ncore = os.cpu_count()
    for sig, tac, oracle_file in CASES:
        t0 = time.time()
        mean0, L = pilot(sig, tac)
        with mp.get_context("fork").Pool(ncore) as pool:
            res = pool.map(chain, [(sig, tac, 7000 + c, mean0, L) for c in range(ncore)])
        x = np.stack([r[0] for r in res])                           # (chains, kept, 96)
        acc = np.array([r[1] for r in res])
        out = dict(mean=x.mean(axis=(0, 1)), sd=x.std(axis=(0, 1), ddof=1),
                   mcse_mean=np.array([dg.mcse_mean(x[:, :, k]) for k in range(96)]),
                   mcse_sd=np.array([dg.mcse_sd(x[:, :, k]) for k in range(96)]),
                   rhat=np.array([dg.rhat_rank(x[:, :, k]) for k in range(96)]),
                   ess_bulk=np.array([dg.ess_bulk(x[:, :, k]) for k in range(96)]),
                   accept_rate=acc, chains=ncore, steps=STEPS, burn=BURN, thin=THIN, tac=tac, sigma=float(sig))
        path = os.path.join(G, "independent_posterior_s%s_tac%d.npz" % (sig, tac))
        np.savez_compressed(path, **out)
        ref = np.load(os.path.join(G, oracle_file))
        z = (out["mean"] - ref["mean"]) / np.sqrt(out["mcse_mean"] ** 2 + ref["mcse_mean"] ** 2)
        zs = (out["sd"] - ref["sd"]) / np.sqrt(out["mcse_sd"] ** 2 + ref["mcse_sd"] ** 2)
        print("%s: %.0f s, accept %.3f, rhat max %.3f, ess_bulk min %.0f | vs oracle: mean max|z| %.2f rms %.2f, sd max|z| %.2f rms %.2f"
              % (os.path.basename(path), time.time() - t0, acc.mean(), out["rhat"].max(), out["ess_bulk"].min(),
                 np.abs(z).max(), np.sqrt((z ** 2).mean()), np.abs(zs).max(), np.sqrt((zs ** 2).mean())), flush=True)


if __name__ == "__main__":
    main()
