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
