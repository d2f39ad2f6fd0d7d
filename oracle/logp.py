"""Log-probability of the reference PyMC model (mcmc.py:147-155), restated in numpy fp64.

PARITY UNPINNED at the PyMC boundary (pymc==5.12.0 is third-party and absent); formulas
are the published densities and are cross-checked against scipy.stats in tests/.

Model (mcmc.py:148-155):
    var_DVR ~ MvNormal(mu_DVR, Cov_DVR);  var_R1 ~ MvNormal(mu_R1, Cov_R1)   (unconstrained)
    sn  = SRTM2(DVR, R1, k2p).T ; sn = switch(sn < 0, 1e-6, sn)
    y_obs ~ TruncatedNormal(mu=sn, sigma=sqrt(sn)*sigma_noise, lower=0)
"""
import numpy as np
from scipy.special import erfc

LOG_SQRT_2PI = 0.5 * np.log(2.0 * np.pi)


def clamp_tac(sn):
    """pytensor switch(sn < 0, 1e-6, sn) (mcmc.py:152). NaN stays NaN."""
    return np.where(sn < 0, 1e-6, sn)


def truncnormal_lower0_logpdf(y, mu, sigma):
    """log pdf of Normal(mu, sigma) truncated to [0, inf): normal log-pdf minus
    log(1 - Phi((0-mu)/sigma)) = log1p(-erfc(mu/(sigma sqrt2))/2); -inf where y < 0."""
    with np.errstate(all="ignore"):
        z = (y - mu) / sigma
        out = -0.5 * z * z - np.log(sigma) - LOG_SQRT_2PI \
            - np.log1p(-0.5 * erfc(mu / (sigma * np.sqrt(2.0))))
        return np.where(y < 0, -np.inf, out)


def loglik_elements(y, sn, sigma_noise):
    """(R,T) element log-likelihoods, mcmc.py:152-155."""
    sn = clamp_tac(sn)
    with np.errstate(all="ignore"):
        return truncnormal_lower0_logpdf(y, sn, np.sqrt(sn) * sigma_noise)


def loglik_roi(y, sn, sigma_noise):
    """Per-ROI sums over frames, shape (R,)."""
    return loglik_elements(y, sn, sigma_noise).sum(axis=-1)


def loglik_roi_reduced(y, sn, sigma_noise):
    """Per-ROI log-likelihood WITHOUT the state-independent terms -log(sigma_noise)
    - log sqrt(2 pi) (they cancel in every Metropolis difference):
       sum_t -((y-s)^2)/(2 s sig^2) - log(s)/2 - log1p(-erfc(sqrt(s)/(sig sqrt2))/2)."""
    s = clamp_tac(sn)
    with np.errstate(all="ignore"):
        el = -0.5 * (y - s) ** 2 / (s * sigma_noise ** 2) - 0.5 * np.log(s) \
            - np.log1p(-0.5 * erfc(np.sqrt(s) / (sigma_noise * np.sqrt(2.0))))
        el = np.where(y < 0, -np.inf, el)
    return el.sum(axis=-1)


def loglik_constant(sigma_noise):
    """The dropped per-ROI constant: sum_t -log(sigma_noise) - log sqrt(2pi)."""
    return (-np.log(sigma_noise) - LOG_SQRT_2PI).sum(axis=-1)


def mvnormal_logpdf(x, mu, cov):
    """MvNormal log-density via Cholesky (what PyMC's MvNormal.logp does)."""
    L = np.linalg.cholesky(cov)
    d = np.linalg.solve(L, x - mu)
    return -0.5 * d @ d - np.log(np.diag(L)).sum() - 0.5 * len(mu) * np.log(2 * np.pi)


class Model:
    """Everything one test TAC's posterior needs (one loop body of mcmc.py:104-157)."""

    def __init__(self, t, c_r, k2p, y_obs, sigma_noise, mu_DVR, Cov_DVR, mu_R1, Cov_R1):
        from . import forward
        self.t = np.asarray(t, np.float64)
        self.c_r = np.asarray(c_r, np.float64)
        self.k2p = float(k2p)
        self.y = np.asarray(y_obs, np.float64)
        self.sigma_noise = np.asarray(sigma_noise, np.float64)
        self.mu = [np.asarray(mu_DVR, np.float64), np.asarray(mu_R1, np.float64)]
        self.cov = [np.asarray(Cov_DVR, np.float64), np.asarray(Cov_R1, np.float64)]
        self.P = [np.linalg.inv(c) for c in self.cov]
        self.P = [0.5 * (p + p.T) for p in self.P]
        self.M = forward.build_M(self.t, self.c_r)
        self._forward = forward

    # ---- reference-faithful full-model log-probability (what delta_logp evaluates) ----
    def logp_full(self, DVR, R1):
        with np.errstate(all="ignore"):
            sn = self._forward.srtm2_tac(self.t, self.c_r, DVR, R1, self.k2p).T
            ll = loglik_elements(self.y, sn, self.sigma_noise).sum()
        return ll + mvnormal_logpdf(DVR, self.mu[0], self.cov[0]) \
            + mvnormal_logpdf(R1, self.mu[1], self.cov[1])

    # ---- lean per-ROI pieces (operator form) ----
    def tac_roi(self, roi, dvr, r1):
        with np.errstate(all="ignore"):
            k2 = self.k2p * r1
            k2a = k2 / dvr
            e = np.exp(-k2a * self.t)
            return r1 * self.c_r + (k2 - r1 * k2a) * (self.M @ e)

    def ll_roi(self, roi, dvr, r1):
        s = self.tac_roi(roi, dvr, r1)
        return float(loglik_roi_reduced(self.y[roi], s, self.sigma_noise[roi]))

    def ll_all(self, DVR, R1):
        with np.errstate(all="ignore"):
            sn = self._forward.srtm2_tac_M(self.t, self.c_r, self.M, DVR, R1, self.k2p).T
        return loglik_roi_reduced(self.y, sn, self.sigma_noise)
