"""Restated PyMC log-probabilities vs scipy.stats (the only available cross-check: pymc is
third-party and absent -- 'parity unpinned' at that boundary, see oracle/__init__.py)."""
import numpy as np
from scipy import stats

from oracle import logp


def test_truncnormal_matches_scipy():
    rng = np.random.default_rng(0)
    mu = rng.uniform(1e-3, 8, 500)
    sig = rng.uniform(0.01, 2, 500)
    y = np.abs(mu + sig * rng.standard_normal(500))
    ref = stats.truncnorm.logpdf(y, (0 - mu) / sig, np.inf, loc=mu, scale=sig)
    got = logp.truncnormal_lower0_logpdf(y, mu, sig)
    assert np.abs(got - ref).max() < 1e-10 * (1 + np.abs(ref).max())
    assert logp.truncnormal_lower0_logpdf(np.array([-0.1]), np.array([1.0]), np.array([1.0]))[0] == -np.inf


def test_mvnormal_matches_scipy(prior):
    rng = np.random.default_rng(1)
    x = prior["mu_DVR"] + 0.05 * rng.standard_normal(48)
    ref = stats.multivariate_normal.logpdf(x, prior["mu_DVR"], prior["Cov_DVR"])
    assert abs(logp.mvnormal_logpdf(x, prior["mu_DVR"], prior["Cov_DVR"]) - ref) < 1e-6 * abs(ref)


def test_reduced_loglik_differs_by_constant(models):
    m = models[0]
    rng = np.random.default_rng(2)
    c = logp.loglik_constant(m.sigma_noise)
    for _ in range(3):
        DVR = m.mu[0] * (1 + 0.05 * rng.standard_normal(48))
        R1 = m.mu[1] * (1 + 0.05 * rng.standard_normal(48))
        sn = m._forward.srtm2_tac(m.t, m.c_r, DVR, R1, m.k2p).T
        full = logp.loglik_roi(m.y, sn, m.sigma_noise)
        red = m.ll_all(DVR, R1)
        assert np.abs(full - (red + c)).max() < 1e-9 * np.abs(full).max()


def test_clamp_and_nan_semantics():
    s = np.array([-1.0, 0.0, 2.0, np.nan])
    out = logp.clamp_tac(s)
    assert out[0] == 1e-6 and out[1] == 0.0 and out[2] == 2.0 and np.isnan(out[3])


def test_full_logp_equals_sum_of_parts(models, prior):
    m = models[1]
    DVR, R1 = m.mu[0] * 1.01, m.mu[1] * 0.99
    sn = m._forward.srtm2_tac(m.t, m.c_r, DVR, R1, m.k2p).T
    tot = logp.loglik_roi(m.y, sn, m.sigma_noise).sum() + logp.mvnormal_logpdf(DVR, m.mu[0], m.cov[0]) \
        + logp.mvnormal_logpdf(R1, m.mu[1], m.cov[1])
    assert abs(m.logp_full(DVR, R1) - tot) < 1e-9 * abs(tot)
