"""K4: the batched GPU generator vs the reference generator's semantics
(sample_sim_data.py:141-215): constraints, forward consistency, distributions, reproducibility."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make(prior, dataset, n, seed, tac_gid0=0):
    from pet_posterior_distribution_b200 import MHSampler
    s = MHSampler(n_chains=2, max_tacs=n, seed=1, tac_gid0=tac_gid0)
    s.set_frames(dataset["time_vector"], dataset["dt"])
    s.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
    s.synth(n, seed, prior["mu_tac_ref"], prior["Cov_tac_ref"], float(prior["mu_k2p"]), dataset["sigma_noise"])
    return s


def test_constraints_and_forward_consistency(prior, dataset):
    from oracle import forward
    s = _make(prior, dataset, 256, seed=3)
    g = s.synth_get()
    for k in ("DVR", "R1", "tac_ref", "tac_clean", "y"):
        assert np.isfinite(g[k]).all() and (g[k] >= 0).all(), k        # positivity rejection, truncated noise
    assert (g["attempts"] >= 1).all() and g["attempts"].max() < 50
    t = dataset["time_vector"]
    for i in (0, 17, 255):                                               # clean TAC == reference forward model of the draws
        ref = forward.srtm2_tac(t, g["tac_ref"][i], g["DVR"][i].astype(np.float64), g["R1"][i].astype(np.float64),
                                float(prior["mu_k2p"])).T
        assert np.abs(g["tac_clean"][i] / ref - 1).max() < 1e-5
    # the generated batch is bound as the sampler's data: run a few sweeps on it
    s.run(draws=5, tune=100)
    assert np.isfinite(s.summary()[..., :2]).all()


def test_distributions(prior, dataset):
    n = 4096
    g = _make(prior, dataset, n, seed=11).synth_get()
    # against the CPU restatement of the reference generator (same rejection rules): means and SDs of the
    # accepted draws agree within Monte-Carlo error (the positivity / negative-TAC rejections shift both
    # away from the untruncated prior, identically on both sides)
    from oracle import generator
    m = 1500
    ref = generator.generate(prior, m, 0.1, test_style=False, seed=123)
    for key, okey in (("DVR", "varDVR"), ("R1", "varR1"), ("tac_ref", "vartacref")):
        a, b2 = g[key].astype(np.float64), np.asarray(ref[okey])
        se = np.sqrt(a.var(axis=0) / n + b2.var(axis=0) / m)
        zs = (a.mean(axis=0) - b2.mean(axis=0)) / se
        assert np.abs(zs).max() < 5.0, (key, np.abs(zs).max())
        ratio = a.std(axis=0) / b2.std(axis=0)
        assert np.abs(ratio - 1).max() < 0.12, (key, ratio.min(), ratio.max())
    # noise model: (y - x)/sqrt(x) ~ TruncNormal(0, sigma, low=-sqrt(x)); where sqrt(x) >> sigma it is N(0, sigma^2)
    x, y, sig = g["tac_clean"].astype(np.float64), g["y"].astype(np.float64), dataset["sigma_noise"]
    res = (y - x) / np.sqrt(x)
    hi = (np.sqrt(x) > 6 * sig[None]).mean(axis=0) > 0.99               # (roi, frame) cells with negligible truncation
    sd = res.std(axis=0)
    assert hi.sum() > 500
    assert np.abs(sd[hi] / sig[hi] - 1).max() < 0.08 and np.abs(res.mean(axis=0)[hi] / sig[hi]).max() < 0.1
    assert (res >= -np.sqrt(x) - 1e-4).all()


def test_reproducible_and_shard_independent(prior, dataset):
    a = _make(prior, dataset, 8, seed=5).synth_get()
    b = _make(prior, dataset, 8, seed=5).synth_get()
    assert np.array_equal(a["y"], b["y"]) and np.array_equal(a["tac_ref"], b["tac_ref"])
    c = _make(prior, dataset, 4, seed=5, tac_gid0=4).synth_get()         # second half generated on "another rank"
    assert np.array_equal(c["y"], a["y"][4:]) and np.array_equal(c["DVR"], a["DVR"][4:])
    d = _make(prior, dataset, 8, seed=6).synth_get()
    assert not np.array_equal(d["y"], a["y"])


def test_test_style_mahalanobis_rule(prior, dataset):
    """K4 with the reference's test-set rule (sample_sim_data.py:128-133): every kept DVR / R1 draw satisfies
    chi2.cdf(d^2, 48) < alpha (recomputed on the host with the same inverse), the rule binds (the training-style stream
    violates it for ~20 % of the draws), the kept draws are distributed like the host path's, and switching the rule off
    restores the training-style stream bit for bit."""
    from scipy import stats
    from pet_posterior_distribution_b200 import MHSampler
    from pet_posterior_distribution_b200 import sample_sim_data as gen
    n, alpha = 768, 0.8
    thr = stats.chi2.ppf(alpha, 48)
    s = MHSampler(n_chains=2, max_tacs=n, seed=1)
    s.set_frames(dataset["time_vector"], dataset["dt"])
    s.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
    args = (prior["mu_tac_ref"], prior["Cov_tac_ref"], float(prior["mu_k2p"]), dataset["sigma_noise"])
    s.synth_test_rule(alpha, prior["Cov_DVR"], prior["Cov_R1"], prior["Cov_tac_ref"])
    s.synth(n, 7, *args)
    g = s.synth_get()
    s.synth_test_rule(None)
    s.synth(n, 7, *args)
    tr = s.synth_get()

    def d2(x, k):
        d = x.astype(np.float64) - prior["mu_" + k]
        return np.einsum("ni,ij,nj->n", d, np.linalg.inv(prior["Cov_" + k]), d)

    for k in ("DVR", "R1", "tac_ref", "tac_clean", "y"):
        assert np.isfinite(g[k]).all() and (g[k] >= 0).all(), k
    assert (g["attempts"] >= 1).all()
    for k in ("DVR", "R1"):
        a, b = d2(g[k], k), d2(tr[k], k)
        assert a.min() >= 0 and a.max() < thr + 0.1, (k, a.max(), thr)   # (+0.1: the draws come back rounded to fp32)
        assert a.max() > 0.85 * thr                                      # ... and the bound is reached
        assert 0.08 < (b > thr).mean() < 0.35, (k, (b > thr).mean())     # the rule binds: 1 - alpha of the unfiltered draws
        assert g[k].std(axis=0).mean() < tr[k].std(axis=0).mean()        # truncation in Mahalanobis distance shrinks the spread
    # the host path of the same rule (numpy draws, sample_sim_data.generate): same distribution of the kept draws
    m = 400
    ref = gen.generate(prior, m, 0.1, test_style=True, seed=5, alpha_=alpha)
    for key, okey in (("DVR", "varDVR"), ("R1", "varR1")):
        a, b = g[key].astype(np.float64), np.asarray(ref[okey])
        se = np.sqrt(a.var(axis=0) / n + b.var(axis=0) / m)
        assert np.abs((a.mean(axis=0) - b.mean(axis=0)) / se).max() < 5.0, key
        assert np.abs(a.std(axis=0) / b.std(axis=0) - 1).max() < 0.2, key
    # rule off again == the training-style generator
    plain = _make(prior, dataset, n, seed=7).synth_get()
    assert np.array_equal(plain["y"], tr["y"]) and np.array_equal(plain["DVR"], tr["DVR"])
    assert not np.array_equal(plain["DVR"], g["DVR"])
    # the pickle schema of the GPU path carries the flag (sample_sim_data.py:221)
    ds = gen.generate_gpu(prior, 8, 0.1, seed=3, test_style=True)
    assert ds["flag_mahalanobis"] is True and len(ds["varDVR"]) == 8 and ds["tac_noisy_sampled"][0].shape == (48, 54)
    assert gen.generate_gpu(prior, 8, 0.1, seed=3)["flag_mahalanobis"] is False
    s.close()


def test_distributions_match_the_live_reference(prior, dataset):
    """K4 vs OUTPUTS OF THE REFERENCE ITSELF: tests/golden/reference_generated_stats.npz holds the per-coordinate mean / sd of the
    DVR, R1 and reference-TAC draws that /root/reference/sample_sim_data.py kept in a 3000-sample training-style and a
    1200-sample test-style run (tools/make_reference_generated.py), after its positivity, negative-TAC and Mahalanobis rules.
    The GPU generator's kept draws have the same moments within Monte-Carlo error."""
    import os
    from pet_posterior_distribution_b200 import MHSampler
    st = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_generated_stats.npz"))
    n = 4096
    s = MHSampler(n_chains=2, max_tacs=n, seed=1)
    s.set_frames(dataset["time_vector"], dataset["dt"])
    s.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
    args = (prior["mu_tac_ref"], prior["Cov_tac_ref"], float(prior["mu_k2p"]), dataset["sigma_noise"])
    for tag, alpha in (("train", None), ("test", 0.8)):
        s.synth_test_rule(alpha, prior["Cov_DVR"], prior["Cov_R1"], prior["Cov_tac_ref"])
        s.synth(n, 21, *args)
        g = s.synth_get()
        n_ref = int(st[tag + "_n"])
        for key, okey in (("DVR", "varDVR"), ("R1", "varR1"), ("tac_ref", "vartacref")):
            a = g[key].astype(np.float64)
            mu, sd = st["%s_%s_mean" % (tag, okey)], st["%s_%s_sd" % (tag, okey)]
            z = (a.mean(axis=0) - mu) / np.sqrt(sd ** 2 / n_ref + a.var(axis=0) / n)
            assert np.abs(z).max() < 5.0 and np.sqrt((z ** 2).mean()) < 1.7, (tag, key, np.abs(z).max(), np.sqrt((z ** 2).mean()))
            ratio = a.std(axis=0, ddof=1) / sd
            assert np.abs(ratio - 1).max() < (0.1 if tag == "train" else 0.15), (tag, key, ratio.min(), ratio.max())
    s.close()
