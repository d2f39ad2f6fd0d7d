"""Synthetic SRTM2 test data with the reference's generator semantics
(sample_sim_data.py:128-224, helper_func.py:146-162), vectorised: parameters are drawn on
the host (numpy), the forward simulation runs on the B200 through libpetmh.

Produces the reference's pickle schema (sample_sim_data.py:218-224).  Used by bench.py and
by ``python -m pet_posterior_distribution_b200.sample_sim_data`` to write
sim_data/nROI48/<timestamp>_{train,test}/data_nROI48_n<N>_s<sigma>.pik like the reference.
"""
import json
import os
import pickle
from datetime import datetime

import numpy as np
from scipy import stats as spst

from .frames import MK_HALF_T, frame_grid
from .sampler import MHSampler

# module-level configuration, same names and defaults as sample_sim_data.py:88-95 (the shipped defaults write the
# training set; mcmc.py reads a test set made with n_samples = 100, flag_testing_data = True)
n_samples = 100000
n_ROI = 48
flag_testing_data = False
mean_sigma_noise_save = 1e-1
alpha = 0.8


def _mvn_positive(rng, mu, cov, cov_inv, n, test_style, alpha_, dof):
    """helper_func.truncnormal_samples: reject draws with a negative component and, for the
    test set, with chi2.cdf(Mahalanobis^2, dof) >= alpha (sample_sim_data.py:128-133)."""
    _, sv, vt = np.linalg.svd(cov)
    A = np.sqrt(sv)[:, None] * vt
    out = np.empty((0, mu.size))
    while out.shape[0] < n:
        x = mu + rng.standard_normal((max(64, 2 * (n - out.shape[0])), mu.size)) @ A
        ok = (x >= 0).all(axis=1)
        if test_style:
            # scipy's mahalanobis is sqrt(d' VI d): a NEGATIVE quadratic form (about half of the reference-TAC draws: the
            # inverse of its rank-deficient covariance is numerically indefinite) is NaN there and fails the test
            d = x - mu
            d2 = np.einsum("ij,jk,ik->i", d, cov_inv, d)
            ok &= (d2 >= 0) & (spst.chi2.cdf(np.maximum(d2, 0.0), dof) < alpha_)
        out = np.concatenate([out, x[ok]])
    return out[:n]


def _roi_names(prior):
    """target_ROI_names as the reference pickles it: the plain list of prior_stats_nROI48.pik (sample_sim_data.py:103,221)."""
    names = prior.get("ROI_names")
    return None if names is None else [str(v) for v in names]


def _trunc_normal(rng, mean, std, low):
    a = (low - mean) / std
    return spst.truncnorm.rvs(a, np.inf, loc=mean, scale=std, random_state=rng)


def noise_table(rng, mean_sigma_noise, t, dt, nroi=48):
    """sigma_noise (48,54) of sample_sim_data.py:193-201: per-ROI level ~ TruncNormal(mean, 0.3 mean, low=0),
    scaled by 1/sqrt(dt * exp(-lambda t)), lambda = ln2 / 109.8 min."""
    lam = np.log(2) / MK_HALF_T
    sigma_roi = _trunc_normal(rng, np.full(nroi, mean_sigma_noise), 0.3 * mean_sigma_noise, 0.0)
    return sigma_roi[:, None] / np.sqrt(dt[None, :] * np.exp(-lam * t))


def generate_gpu(prior, n, mean_sigma_noise=0.1, seed=0, device=0, sampler=None, test_style=False, alpha_=0.8):
    """Data set generated entirely on the B200 (K4, petmh_synth): returns the reference's pickle schema.
    test_style adds the Mahalanobis rule of the reference's test set (sample_sim_data.py:128-133).  With `sampler`
    given, the batch stays bound to it as its data (no host round trip)."""
    rng = np.random.default_rng(seed)
    t, dt = frame_grid()
    sigma_noise = noise_table(rng, mean_sigma_noise, t, dt)
    own = sampler is None
    s = sampler or MHSampler(n_chains=1, max_tacs=n, device=device)
    if own:
        s.set_frames(t, dt)
        s.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
    s.synth_test_rule(alpha_ if test_style else None, prior["Cov_DVR"], prior["Cov_R1"], prior["Cov_tac_ref"])
    try:
        s.synth(n, seed, prior["mu_tac_ref"], prior["Cov_tac_ref"], float(prior["mu_k2p"]), sigma_noise)
        g = s.synth_get()
    finally:
        if own:
            s.close()
        else:
            s.synth_test_rule(None)
    return {"varDVR": list(g["DVR"].astype(np.float64)), "varR1": list(g["R1"].astype(np.float64)),
            "vark2p": [prior["mu_k2p"] for _ in range(n)], "vartacref": list(g["tac_ref"]),
            "tac_sampled": list(g["tac_clean"].astype(np.float64) * dt[None, None, :]),
            "tac_noisy_sampled": list(g["y"].astype(np.float64) * dt[None, None, :]),
            "mu_noise": np.zeros_like(sigma_noise), "sigma_noise": sigma_noise, "mean_sigma_noise": mean_sigma_noise,
            "flag_mahalanobis": bool(test_style), "target_ROI_names": _roi_names(prior), "time_vector": t, "dt": dt}


def generate(prior, n, mean_sigma_noise=0.1, test_style=False, seed=0, device=0, alpha_=0.8):
    """Host draws + GPU forward model; supports the test-style Mahalanobis rule (sample_sim_data.py:128-133)."""
    rng = np.random.default_rng(seed)
    t, dt = frame_grid()
    nroi = prior["mu_DVR"].size
    inv = {k: np.linalg.inv(prior["Cov_" + k]) for k in ("DVR", "R1", "tac_ref")}
    draw = lambda k, m: _mvn_positive(rng, prior["mu_" + k], prior["Cov_" + k], inv[k], m, test_style, alpha_, nroi)
    DVR, R1, cref = draw("DVR", n), draw("R1", n), draw("tac_ref", n)
    k2p = float(prior["mu_k2p"])
    fwd = MHSampler(n_chains=1, max_tacs=n, device=device)
    fwd.set_frames(t, dt)
    fwd.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
    tac = np.empty((n, nroi, t.size))
    todo = np.arange(n)
    while todo.size:                                   # sample_sim_data.py:171-188
        fwd.set_data(np.ones((n, nroi, t.size)), cref, np.full(n, k2p), np.ones((nroi, t.size)))
        for i in todo:
            tac[i] = fwd.forward(int(i), DVR[i], R1[i])
        bad = np.array([i for i in todo if (tac[i] < 0).any()], int)
        if bad.size:
            DVR[bad], R1[bad], cref[bad] = draw("DVR", bad.size), draw("R1", bad.size), draw("tac_ref", bad.size)
        todo = bad
    fwd.close()
    sigma_noise = noise_table(rng, mean_sigma_noise, t, dt, nroi)       # :193-201
    noisy = tac + np.sqrt(tac) * _trunc_normal(rng, np.zeros_like(tac), np.broadcast_to(sigma_noise, tac.shape),
                                               -np.sqrt(tac))                                  # :205-215
    return {"varDVR": list(DVR), "varR1": list(R1), "vark2p": [prior["mu_k2p"] for _ in range(n)],
            "vartacref": list(cref), "tac_sampled": list(tac * dt[None, None, :]),
            "tac_noisy_sampled": list(noisy * dt[None, None, :]), "mu_noise": np.zeros_like(sigma_noise),
            "sigma_noise": sigma_noise, "mean_sigma_noise": mean_sigma_noise, "flag_mahalanobis": test_style,
            "target_ROI_names": _roi_names(prior), "time_vector": t, "dt": dt}


def load_prior(path=None):
    """prior_stats_nROI48.pik (reference) or the .npz re-save shipped under tests/golden."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in ([path] if path else []) + [os.path.join(".", "prior_stats_nROI%d.pik" % n_ROI),
                                          os.path.join(here, "tests", "golden", "prior_stats_nROI48.npz")]:
        if p and os.path.isfile(p):
            if p.endswith(".npz"):
                z = np.load(p)
                return {k: z[k] for k in z.files}
            return pickle.load(open(p, "rb"))
    raise FileNotFoundError("prior_stats_nROI48 not found")


def main(seed=None):
    """Write sim_data/nROI48/<ts>_{train,test}/data_*.pik + args_*.txt (sample_sim_data.py:96-240): every step on the GPU
    (K4), the test set with the Mahalanobis rule of :128-133.  seed: default the current time, like an unseeded run."""
    prior = load_prior()
    now = datetime.now()
    sd = int(now.timestamp()) if seed is None else int(seed)
    ds = generate_gpu(prior, n_samples, mean_sigma_noise_save, seed=sd, test_style=flag_testing_data, alpha_=alpha)
    str_test = "_test" if flag_testing_data else "_train"
    str_noise = "_s{:.1e}".format(mean_sigma_noise_save)
    save_samples_dir = os.path.join("./sim_data", "nROI{}".format(n_ROI))
    d = os.path.join(save_samples_dir, now.strftime("%y-%m-%d_%H-%M-%S") + str_test)
    os.makedirs(d, exist_ok=True)
    pickle.dump(ds, open(os.path.join(d, "data_nROI{}_n{}{}.pik".format(n_ROI, n_samples, str_noise)), "wb"))
    names = ds["target_ROI_names"]
    with open(os.path.join(d, "args_nROI{}_n{}{}.txt".format(n_ROI, n_samples, str_noise)), "wt") as f:   # :228-240
        json.dump({"mean_sigma_noise": mean_sigma_noise_save, "target_ROI_names": None if names is None else [str(v) for v in names],
                   "MK_half_T": MK_HALF_T, "MK_lambda": np.log(2) / MK_HALF_T, "n_samples": n_samples, "n_ROI": n_ROI,
                   "save_samples_dir": save_samples_dir, "flag_mahalanobis": bool(flag_testing_data)}, f, indent=2, sort_keys=True)
    print("wrote", d)
    return d


if __name__ == "__main__":
    main()
