"""Pin of the generator rows on OUTPUTS OF THE LIVE REFERENCE: tools/make_reference_generated.py
exec'd /root/reference/sample_sim_data.py (unmodified source, its module constants set as a user sets them) in the build
container and committed what it pickled as tests/golden/reference_generated_*.npz.  Checked here, on the CPU:

  * oracle/generator.py and the product's host-side draws (pet_posterior_distribution_b200/sample_sim_data.py) write the
    same pickle schema, the same directory / file names, and satisfy the same rules on the reference's own draws;
  * oracle/forward.py reproduces the reference's clean TACs from the reference's drawn parameters (1e-12);
  * the accepted parameter draws and the added noise are distributed like the reference's (Monte-Carlo bounds; the
    reference draws from numpy's global RandomState, the restatements from seeded Generators: streams differ by design).

The GPU generator (K4) is compared with the same fixtures in tests/test_gpu_synth.py."""
import os

import numpy as np
import pytest
from scipy import stats as sp

from oracle import forward, frames, generator

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARAMS = (("DVR", "varDVR"), ("R1", "varR1"), ("tac_ref", "vartacref"))


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def ref_stats():
    return _load("reference_generated_stats.npz")


@pytest.mark.parametrize("style", ["test", "train"])
def test_pickle_schema_and_names(prior, style):
    """sample_sim_data.py:96-100,163-168,218-240: keys, container types, directory suffix and file names of the live run."""
    ref = _load("reference_generated_%s_s0.1.npz" % style)
    ds = generator.generate(prior, 2, 0.1, test_style=(style == "test"), seed=1)
    ds.pop("seed")                                                      # the oracle's own addition
    assert sorted(ds) == list(ref["pickle_keys"])
    for k, tname in zip(ref["pickle_keys"], ref["pickle_types"]):
        if k in ("mean_sigma_noise", "flag_mahalanobis"):
            assert type(ds[k]).__name__ == tname, k
        else:                                                           # lists of per-sample arrays vs plain arrays
            assert isinstance(ds[k], list) == (tname == "list"), (k, tname)
    assert bool(ref["flag_mahalanobis"]) == (style == "test") == ds["flag_mahalanobis"]
    assert str(ref["dir_suffix"]) == style
    assert str(ref["file_name"]) == "data_nROI48_n6_s1.0e-01.pik" and list(ref["args_file"]) == ["args_nROI48_n6_s1.0e-01.txt"]
    assert ref["varDVR"].shape == (6, 48) and ref["vartacref"].shape == (6, 54) and ref["vark2p"].shape == (6,)
    assert ref["tac_sampled"].shape == ref["tac_noisy_sampled"].shape == (6, 48, 54)
    assert ref["sigma_noise"].shape == ref["mu_noise"].shape == (48, 54) and not ref["mu_noise"].any()
    assert np.all(ref["vark2p"] == float(prior["mu_k2p"]))              # :150: k2p fixed at the population value
    assert list(ref["target_ROI_names"]) == [str(v) for v in prior["ROI_names"]]
    t, dt = frames.frame_grid()                                         # :29-86: the acquisition frames
    assert np.array_equal(ref["time_vector"], t) and np.array_equal(ref["dt"], dt)


@pytest.mark.parametrize("style", ["test", "train"])
def test_forward_and_rules_on_reference_draws(prior, style):
    ref = _load("reference_generated_%s_s0.1.npz" % style)
    t, dt = ref["time_vector"], ref["dt"]
    for s in range(6):                                                  # :171-188 tac_sampled = create_activity_curve * dt
        x = (forward.srtm2_tac(t, ref["vartacref"][s], ref["varDVR"][s], ref["varR1"][s], float(ref["vark2p"][s])) * dt[:, None]).T
        assert np.abs(x / ref["tac_sampled"][s] - 1).max() < 1e-12
    for k in ("varDVR", "varR1", "vartacref", "tac_sampled", "tac_noisy_sampled"):
        assert (ref[k] >= 0).all(), k                                   # helper_func.py:160, sample_sim_data.py:174,207-212
    if style == "test":                                                 # :129-133: every kept draw passes the restated rule
        for pk, dk in PARAMS:
            inv = np.linalg.inv(prior["Cov_" + pk])
            assert generator.mahalanobis_rule(ref[dk], prior["mu_" + pk], inv, 48, 0.8).all(), pk
    # :193-199 sigma_noise = sigma_roi / sqrt(dt exp(-lambda t)): one level per ROI
    lam = np.log(2) / frames.MK_HALF_T
    sr = ref["sigma_noise"] * np.sqrt(dt[None, :] * np.exp(-lam * t))
    assert np.ptp(sr, axis=1).max() < 1e-15 and (sr > 0).all()
    assert float(ref["mean_sigma_noise"]) == 0.1


def _compare_moments(draws, st, tag, n_ref, z_max=5.0, sd_tol=0.15):
    for pk, dk in PARAMS:
        b = np.asarray(draws[dk], np.float64)
        mu, sd = st["%s_%s_mean" % (tag, dk)], st["%s_%s_sd" % (tag, dk)]
        se = np.sqrt(sd ** 2 / n_ref + b.var(axis=0) / b.shape[0])
        z = (b.mean(axis=0) - mu) / se
        assert np.abs(z).max() < z_max and np.sqrt((z ** 2).mean()) < 1.6, (tag, dk, np.abs(z).max())
        ratio = b.std(axis=0, ddof=1) / sd
        assert np.abs(ratio - 1).max() < sd_tol, (tag, dk, ratio.min(), ratio.max())


def test_oracle_draws_distributed_like_the_reference(prior, ref_stats):
    """Accepted DVR / R1 / reference-TAC draws of the restated generator (positivity, negative-TAC redraw, and for the test set
    the Mahalanobis rule with scipy's NaN behaviour) vs 3000 training-style / 1200 test-style samples of the live script."""
    _compare_moments(generator.generate(prior, 600, 0.1, test_style=False, seed=123), ref_stats, "train", int(ref_stats["train_n"]))
    _compare_moments(generator.generate(prior, 250, 0.1, test_style=True, seed=5), ref_stats, "test", int(ref_stats["test_n"]), sd_tol=0.25)
    # the Mahalanobis rule trims the spread (about 1 % per coordinate in 48 dimensions): the reference's own two sets differ that way
    for _, dk in PARAMS[:2]:
        assert 0.97 < ref_stats["test_%s_sd" % dk].mean() / ref_stats["train_%s_sd" % dk].mean() < 1.0


def test_product_host_draws_distributed_like_the_reference(prior, ref_stats):
    """The product's vectorised host-side draws (sample_sim_data._mvn_positive) inside the redraw-while-negative loop of
    sample_sim_data.generate (:171-188; the forward model here is the oracle's, the product's runs on the GPU) vs the live
    script.  The redraw matters: about a fifth of the training-style parameter sets give a negative clean TAC somewhere, and
    dropping them moves the DVR means by ~10 standard errors of this comparison."""
    from pet_posterior_distribution_b200 import sample_sim_data as gen
    rng = np.random.default_rng(9)
    t, dt = frames.frame_grid()
    k2p = float(prior["mu_k2p"])
    for tag, test, m, tol in (("train", False, 2500, 0.12), ("test", True, 1000, 0.18)):
        inv = {pk: np.linalg.inv(prior["Cov_" + pk]) for pk, _ in PARAMS}
        draw = lambda pk, n: gen._mvn_positive(rng, prior["mu_" + pk], prior["Cov_" + pk], inv[pk], n, test, 0.8, 48)
        d = {dk: draw(pk, m) for pk, dk in PARAMS}
        todo, redrawn = np.arange(m), 0
        while todo.size:
            bad = np.array([i for i in todo if (forward.srtm2_tac(t, d["vartacref"][i], d["varDVR"][i], d["varR1"][i], k2p) < 0).any()], int)
            for pk, dk in PARAMS:
                if bad.size:
                    d[dk][bad] = draw(pk, bad.size)
            redrawn += bad.size
            todo = bad
        if not test:
            assert 0.05 < redrawn / m < 0.6, redrawn / m
        _compare_moments(d, ref_stats, tag, int(ref_stats[tag + "_n"]), sd_tol=tol)


def test_noise_model_of_the_reference_is_the_restated_law(ref_stats):
    """sample_sim_data.py:205-215 and helper_func.py:146-150: (noisy - clean) / (sqrt(clean) sigma) is a standard normal
    truncated at -sqrt(clean)/sigma.  The live script's residuals follow that law -- the one oracle.generator.trunc_normal
    and K4 implement -- far from the truncation (quantiles of 6.7 M residuals) and where it bites (KS on 4000 pairs)."""
    q = ref_stats["train_noise_z_far_quantiles"]
    assert np.abs(q - sp.norm.ppf([0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99])).max() < 5e-3
    assert int(ref_stats["train_noise_z_far_n"]) > 5e6
    for tag in ("train", "test"):
        low, z = ref_stats[tag + "_noise_near_low"], ref_stats[tag + "_noise_near_z"]
        assert (z >= low).all() and low.max() > -0.1 and len(z) >= 4000
        u = sp.truncnorm.cdf(z, low, np.inf)
        assert sp.kstest(u, "uniform").pvalue > 1e-3
        m, s = sp.truncnorm.mean(low, np.inf), sp.truncnorm.std(low, np.inf)
        assert abs(((z - m) / s).mean()) < 4 / np.sqrt(len(z))
        assert float(ref_stats[tag + "_min_noisy"]) >= 0 and float(ref_stats[tag + "_min_clean"]) >= 0
    # late frames (plain normal) have unit spread, early frames (small clean signal, truncation) less: both sides of :210
    sd = ref_stats["train_noise_z_sd_by_frame"]
    assert abs(sd[-10:].mean() - 1) < 5e-3 and sd[0] < 0.8


def test_noise_level_per_roi(ref_stats):
    """sample_sim_data.py:197: sigma_roi ~ TruncNormal(mean, 0.3 mean, low 0): 40 live runs pooled (1920 levels) vs the law, and
    the restatements' tables drawn from the same law."""
    sr = ref_stats["sigma_roi_pool"]
    law = lambda v: sp.truncnorm.cdf(v, (0 - 0.1) / 0.03, np.inf, loc=0.1, scale=0.03)
    assert sp.kstest(law(sr), "uniform").pvalue > 1e-3 and abs(sr.mean() - 0.1) < 4 * 0.03 / np.sqrt(len(sr))
    from pet_posterior_distribution_b200 import sample_sim_data as gen
    t, dt = frames.frame_grid()
    lam = np.log(2) / frames.MK_HALF_T
    mine = np.concatenate([(gen.noise_table(np.random.default_rng(s), 0.1, t, dt) * np.sqrt(dt * np.exp(-lam * t)))[:, 0] for s in range(40)])
    assert sp.kstest(law(mine), "uniform").pvalue > 1e-3
    assert sp.ks_2samp(mine, sr).pvalue > 1e-3
    orc = generator.trunc_normal(np.random.default_rng(3), 0.1, 0.03, low=0, size=1920)
    assert sp.ks_2samp(orc, sr).pvalue > 1e-3


@pytest.mark.skipif(not os.path.isfile("/root/reference/sample_sim_data.py"), reason="the live reference exists in the build container only")
def test_fixture_regenerates_from_the_live_reference():
    """Provenance: exec'ing /root/reference/sample_sim_data.py again (same numpy global seed) reproduces the committed
    test-style fixture bit for bit.  Child process: the runner installs stub modules for matplotlib / diffusion_model."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, io, contextlib, numpy as np
sys.path.insert(0, %r)
sys.path.insert(0, %r + "/tools")
import make_reference_generated as mk
ref = np.load(%r)
seed = int(ref["numpy_global_seed"])
with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
    d = mk.run_reference_script(6, True, 0.1, seed)
a = mk.as_arrays(d, seed)
for k in ("varDVR", "varR1", "vartacref", "tac_sampled", "tac_noisy_sampled", "sigma_noise"):
    assert np.array_equal(a[k], ref[k]), k
print("ok")
""" % (root, root, os.path.join(GOLDEN, "reference_generated_test_s0.1.npz"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-300:], r.stderr[-800:])
