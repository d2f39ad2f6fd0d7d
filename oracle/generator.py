"""Synthetic test-data generator, restating sample_sim_data.py:128-224 and
helper_func.py:146-162 (numpy fp64, seeded `numpy.random.Generator` instead of the
reference's global RNG -- so streams differ from the reference's, distributions do not).

Produces the exact pickle schema of sample_sim_data.py:218-224.
"""
import numpy as np
from scipy import stats as spst
from . import forward, frames


def trunc_normal(rng, mean, std, low=0.0, size=None):
    """helper_func.trunc_normal (helper_func.py:146-150) with upp=None -> +inf."""
    mean = np.asarray(mean, np.float64)
    std = np.asarray(std, np.float64)
    a = (low - mean) / std
    return spst.truncnorm.rvs(a, np.inf, loc=mean, scale=std, size=size, random_state=rng)


def mahalanobis_rule(x, mu, cov_inv, dof, alpha, reject_negative=True):
    """cond_test of sample_sim_data.py:129-133: chi2.cdf(mahalanobis(mu, x, Cov_inv) ** 2, dof) < alpha.
    scipy's mahalanobis is sqrt(delta @ VI @ delta): where the quadratic form comes out NEGATIVE -- it does for about half
    of the draws of the reference TAC, whose covariance is rank-deficient (cond ~1e20) and whose np.linalg.inv is numerically
    indefinite -- the distance is NaN, the cdf is NaN and the comparison is False: the draw is rejected.
    reject_negative=False is the rule as this oracle had it in round 1 (chi2.cdf of a negative form is 0: accepted); it is
    kept only so that tools/make_golden*.py reproduce the committed fixtures, which were drawn with it.
    x: (d,) or (n, d)."""
    dlt = np.asarray(x, np.float64) - mu
    m = dlt @ cov_inv @ dlt if dlt.ndim == 1 else np.array([d @ cov_inv @ d for d in dlt])
    if not reject_negative:
        return spst.chi2.cdf(m, dof) < alpha
    with np.errstate(invalid="ignore"):
        d_square = np.sqrt(m) ** 2
        return spst.chi2.cdf(d_square, dof) < alpha


def truncnormal_samples(rng, mu, cov, cov_inv, n_samples, test_style=False, alpha=0.8, dof=None, reject_negative=True):
    """helper_func.truncnormal_samples (helper_func.py:153-162): rejection until all
    components >= 0 and, for the test set (sample_sim_data.py:128-133), the Mahalanobis
    d^2 satisfies chi2.cdf(d^2, dof) < alpha with dof = len(mu_DVR) = 48 for every
    variable (even the 54-dim reference TAC, sample_sim_data.py:132)."""
    out = []
    # numpy.random.multivariate_normal factorises with an SVD and tolerates the merely
    # positive-SEMI-definite Cov_tac_ref of the prior file; do the same.
    _, sv, vt = np.linalg.svd(cov)
    A = np.sqrt(sv)[:, None] * vt
    while len(out) < n_samples:
        x = mu + rng.standard_normal(mu.size) @ A
        if np.any(x < 0):
            continue
        if test_style and not mahalanobis_rule(x, mu, cov_inv, dof, alpha, reject_negative):
            continue
        out.append(x)
    return out


def generate(prior, n_samples, mean_sigma_noise=0.1, test_style=False, seed=0, alpha=0.8, reject_negative=True):
    """Returns the dict sample_sim_data.py pickles (lists of per-sample arrays)."""
    rng = np.random.default_rng(seed)
    t, dt = frames.frame_grid()
    n_roi = prior["mu_DVR"].size
    inv = {k: np.linalg.inv(prior["Cov_" + k]) for k in ("DVR", "R1", "tac_ref")}
    kw = dict(test_style=test_style, alpha=alpha, dof=n_roi, reject_negative=reject_negative)
    draw = lambda k, n: truncnormal_samples(rng, prior["mu_" + k], prior["Cov_" + k], inv[k], n, **kw)
    varDVR, varR1, vartacref = draw("DVR", n_samples), draw("R1", n_samples), draw("tac_ref", n_samples)
    k2p = float(prior["mu_k2p"])
    vark2p = [prior["mu_k2p"] for _ in range(n_samples)]
    tac_sampled = []
    for n_i in range(n_samples):                             # sample_sim_data.py:171-188
        while True:
            x = (forward.srtm2_tac(t, vartacref[n_i], varDVR[n_i], varR1[n_i], k2p) * dt[:, None]).T
            if not np.any(x < 0):
                break
            varDVR[n_i], varR1[n_i], vartacref[n_i] = draw("DVR", 1)[0], draw("R1", 1)[0], draw("tac_ref", 1)[0]
        tac_sampled.append(x)
    lam = np.log(2) / frames.MK_HALF_T                       # :193
    sigma_roi = trunc_normal(rng, mean_sigma_noise, 0.3 * mean_sigma_noise, low=0, size=n_roi)
    sigma_noise = sigma_roi[:, None] / np.sqrt(dt[None, :] * np.exp(-lam * t))
    mu_noise = np.zeros_like(sigma_noise)
    tac_noisy = []
    for x in tac_sampled:                                    # :205-215
        xn = x / dt[None, :]
        for r in range(n_roi):
            xn[r] += np.sqrt(xn[r]) * trunc_normal(rng, mu_noise[r], sigma_noise[r], low=-np.sqrt(xn[r]))
        tac_noisy.append(xn * dt[None, :])
    return {"varDVR": varDVR, "varR1": varR1, "vark2p": vark2p, "vartacref": vartacref,
            "tac_sampled": tac_sampled, "tac_noisy_sampled": tac_noisy,
            "mu_noise": mu_noise, "sigma_noise": sigma_noise,
            "mean_sigma_noise": mean_sigma_noise, "flag_mahalanobis": test_style,
            "target_ROI_names": [str(v) for v in prior["ROI_names"]],      # a list of names in prior_stats_nROI48.pik
            "time_vector": t, "dt": dt, "seed": seed}


def model_from_dataset(ds, prior, sample):
    """The per-sample setup of mcmc.py:79-80,106-112,133-134 -> oracle Model."""
    from .logp import Model
    dt = np.asarray(ds["dt"], np.float64)
    y = np.asarray(ds["tac_noisy_sampled"][sample], np.float64) / dt[None, :]
    return Model(ds["time_vector"], ds["vartacref"][sample], ds["vark2p"][sample], y,
                 ds["sigma_noise"], prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
