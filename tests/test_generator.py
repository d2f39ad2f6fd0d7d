"""Restated sample_sim_data.py generator (oracle side): schema, constraints, reproducibility."""
import numpy as np

from oracle import generator

KEYS = {"varDVR", "varR1", "vark2p", "vartacref", "tac_sampled", "tac_noisy_sampled", "mu_noise", "sigma_noise",
        "mean_sigma_noise", "flag_mahalanobis", "target_ROI_names", "time_vector", "dt"}


def test_schema_and_constraints(prior):
    ds = generator.generate(prior, 3, 0.1, test_style=True, seed=3)
    assert KEYS <= set(ds)                                            # sample_sim_data.py:218-224
    assert len(ds["varDVR"]) == 3 and ds["varDVR"][0].shape == (48,) and ds["vartacref"][0].shape == (54,)
    assert ds["tac_sampled"][0].shape == (48, 54) and ds["sigma_noise"].shape == (48, 54)
    for k in ("varDVR", "varR1", "vartacref", "tac_sampled", "tac_noisy_sampled"):
        assert all((np.asarray(v) >= 0).all() for v in ds[k]), k     # positivity rejection / truncated noise
    from scipy import stats
    inv = np.linalg.inv(prior["Cov_DVR"])
    for x in ds["varDVR"]:                                            # test-set Mahalanobis rule
        d = x - prior["mu_DVR"]
        assert stats.chi2.cdf(d @ inv @ d, 48) < 0.8
    assert float(ds["vark2p"][0]) == float(prior["mu_k2p"])


def test_reproducible_and_noise_scales(prior):
    a = generator.generate(prior, 2, 0.05, seed=9)
    b = generator.generate(prior, 2, 0.05, seed=9)
    assert np.array_equal(a["tac_noisy_sampled"][1], b["tac_noisy_sampled"][1])
    c = generator.generate(prior, 2, 0.2, seed=9)
    assert c["sigma_noise"].mean() > 2 * a["sigma_noise"].mean()
    # units: tac_sampled is concentration x dt (sample_sim_data.py:178); y_obs = noisy/dt (mcmc.py:79-80)
    m = generator.model_from_dataset(a, prior, 0)
    assert np.allclose(m.y * a["dt"][None, :], a["tac_noisy_sampled"][0])


def test_mahalanobis_rule_matches_scipy(prior):
    """The test-set rule (sample_sim_data.py:129-133) restated == the reference's own call,
    chi2.cdf(scipy.spatial.distance.mahalanobis(mu, x, Cov_inv) ** 2, 48) < alpha, including its NaN behaviour: the inverse
    of the rank-deficient Cov_tac_ref is numerically indefinite, about half of its quadratic forms are negative, sqrt gives
    NaN and the draw is rejected."""
    import warnings
    from scipy import stats
    from scipy.spatial.distance import mahalanobis
    rng = np.random.default_rng(0)
    for k in ("DVR", "R1", "tac_ref"):
        mu, cov = prior["mu_" + k], prior["Cov_" + k]
        inv = np.linalg.inv(cov)                                      # sample_sim_data.py:106,110,117
        _, sv, vt = np.linalg.svd(cov)
        x = mu + rng.standard_normal((300, mu.size)) @ (np.sqrt(sv)[:, None] * vt)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = np.array([stats.chi2.cdf(mahalanobis(mu, v, inv) ** 2, 48) < 0.8 for v in x])
        mine = generator.mahalanobis_rule(x, mu, inv, 48, 0.8)
        if k == "tac_ref":       # rounding noise of an ill-conditioned form: the summation order may flip a rare draw
            assert (ref == mine).mean() > 0.97 and 0.01 < mine.mean() < 0.25
            neg = np.einsum("ni,ij,nj->n", x - mu, inv, x - mu) < 0
            assert neg.mean() > 0.3 and not mine[neg].any()          # negative forms are rejected, as in the reference
        else:
            assert (ref == mine).all() and 0.6 < mine.mean() < 0.95


def test_product_host_draws_follow_the_test_set_rule(prior):
    """sample_sim_data._mvn_positive (the product's host-side draws for generate(test_style=True)): every kept vector is
    non-negative and passes the reference's rule, negative quadratic forms included (no GPU involved: pure numpy)."""
    from scipy import stats
    from pet_posterior_distribution_b200 import sample_sim_data as gen
    rng = np.random.default_rng(4)
    thr = stats.chi2.ppf(0.8, 48)
    for k, n in (("DVR", 200), ("tac_ref", 60)):
        mu, cov = prior["mu_" + k], prior["Cov_" + k]
        inv = np.linalg.inv(cov)
        x = gen._mvn_positive(rng, mu, cov, inv, n, True, 0.8, 48)
        assert x.shape == (n, mu.size) and (x >= 0).all()
        d2 = np.einsum("ij,jk,ik->i", x - mu, inv, x - mu)
        assert (d2 >= 0).all() and (d2 < thr).all(), k
        assert generator.mahalanobis_rule(x, mu, inv, 48, 0.8).mean() > 0.95      # (tac_ref: up to rounding of the ill-conditioned form)
    free = gen._mvn_positive(rng, prior["mu_DVR"], prior["Cov_DVR"], np.linalg.inv(prior["Cov_DVR"]), 300, False, 0.8, 48)
    d2 = np.einsum("ij,jk,ik->i", free - prior["mu_DVR"], np.linalg.inv(prior["Cov_DVR"]), free - prior["mu_DVR"])
    assert 0.08 < (d2 >= thr).mean() < 0.35                                       # the training set keeps what the rule would drop
