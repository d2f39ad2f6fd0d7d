"""The oracle's posterior against a sampler that shares nothing with the path's algorithm.

tools/make_independent_posterior.py (build container) sampled the posterior of mcmc.py:147-155 with a full-covariance
random-walk Metropolis whose target is assembled from third-party / reference code only -- scipy's multivariate_normal and
truncnorm log-densities and the LIVE /root/reference kinetic_model.SRTM2 forward model -- and committed the moments as
tests/golden/independent_posterior_*.npz.  The golden posteriors of the restated PyMC element-wise Metropolis (fp64 C
oracle, tests/golden/oracle_posterior_*.npz), which the GPU sampler is held to in tests/test_gpu_posterior.py, agree with
them within Monte-Carlo error: the restated model AND the restated sampler draw from the posterior the reference defines.
What this does not pin is PyMC's efficiency-only behaviour (tuning schedule, visit order): tests/test_pymc_pin.py."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("0.1", 0, "oracle_posterior_s0.1_tac0.npz"), ("0.2", 0, "oracle_posterior_s0.2_tac0.npz")]    # 16 chains x 30 000 draws each


@pytest.mark.parametrize("sigma,tac,oracle_file", CASES)
def test_oracle_posterior_matches_independent_sampler(sigma, tac, oracle_file):
    """Two oracle runs per case against the independent sampler (8 chains x 1 M steps): the 16 x 30 000-draw golden the GPU
    sampler is compared with, and a 64 x 30 000-draw run (oracle_posterior_long_*) that resolves 1-2 % of a posterior SD."""
    ind = np.load(os.path.join(GOLDEN, "independent_posterior_s%s_tac%d.npz" % (sigma, tac)))
    # the independent run itself has converged and is long enough to resolve a shift of a fraction of a posterior SD
    assert ind["rhat"].max() < 1.01 and ind["ess_bulk"].min() > 20000
    assert 0.15 < ind["accept_rate"].mean() < 0.35                      # random-walk Metropolis near its optimum
    rms = lambda z: float(np.sqrt((z ** 2).mean()))
    for fname, res_max, mean_tol, sd_tol in ((oracle_file, 0.035, 0.08, 0.04), ("oracle_posterior_long_s%s_tac%d.npz" % (sigma, tac), 0.02, 0.05, 0.02)):
        ref = np.load(os.path.join(GOLDEN, fname))
        assert int(ref["tac"]) == tac == int(ind["tac"])
        res = np.sqrt(ind["mcse_mean"] ** 2 + ref["mcse_mean"] ** 2) / ref["sd"]
        assert res.max() < res_max, (fname, res.max())                 # combined MCSE as a fraction of the posterior SD
        z_mean = (ind["mean"] - ref["mean"]) / np.sqrt(ind["mcse_mean"] ** 2 + ref["mcse_mean"] ** 2)
        z_sd = (ind["sd"] - ref["sd"]) / np.sqrt(ind["mcse_sd"] ** 2 + ref["mcse_sd"] ** 2)
        assert np.abs(z_mean).max() < 4.0 and rms(z_mean) < 1.4, (fname, np.abs(z_mean).max(), rms(z_mean))
        assert np.abs(z_sd).max() < 4.0 and rms(z_sd) < 1.4, (fname, np.abs(z_sd).max(), rms(z_sd))
        # in absolute terms: every posterior mean within mean_tol posterior SDs, every SD within sd_tol
        assert (np.abs(ind["mean"] - ref["mean"]) / ref["sd"]).max() < mean_tol, fname
        assert np.abs(ind["sd"] / ref["sd"] - 1).max() < sd_tol, fname


def test_independent_target_equals_oracle_logp(models, dataset, prior):
    """The independent target, re-assembled here from scipy alone (forward model: the pinned restatement, the live reference
    does not travel), equals oracle.logp.Model.logp_full at random states to 1e-10 relative: the two samplers were given
    the same posterior by two routes."""
    from scipy import stats
    m = models[0]
    mv = (stats.multivariate_normal(prior["mu_DVR"], prior["Cov_DVR"]), stats.multivariate_normal(prior["mu_R1"], prior["Cov_R1"]))
    rng = np.random.default_rng(5)
    for _ in range(10):
        DVR = dataset["varDVR"][0] * (1 + 0.02 * rng.standard_normal(48))
        R1 = dataset["varR1"][0] * (1 + 0.02 * rng.standard_normal(48))
        sn = m._forward.srtm2_tac(m.t, m.c_r, DVR, R1, m.k2p).T
        sn = np.where(sn < 0, 1e-6, sn)
        s = np.sqrt(sn) * m.sigma_noise
        ref = mv[0].logpdf(DVR) + mv[1].logpdf(R1) + stats.truncnorm.logpdf(m.y, (0 - sn) / s, np.inf, loc=sn, scale=s).sum()
        got = m.logp_full(DVR, R1)
        assert abs(got - ref) <= 1e-10 * abs(ref), (got, ref)
