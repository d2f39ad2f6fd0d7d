"""GPU tests of the round-2 library surface: extra pm.summary columns (hdi, mcse_sd) on the GPU, the batch-means ESS of
the moments mode, the full checkpoint (resume INSIDE a tuning window), explicit global ids (sample gaps, chain
sharding), ArviZ's odd-length split, the generator's rejection-cap error, and the vectorised draw store."""
import numpy as np
import pytest

from conftest import make_sampler

pytestmark = pytest.mark.gpu


def test_summary_ext_matches_oracle(dataset, prior):
    """hdi_3% / hdi_97% / mcse_sd / ess_sd from the GPU == numpy restatement of arviz.hdi / _mcse_sd / _ess_sd."""
    from oracle import diagnostics as dg
    s = make_sampler(dataset, prior, n_chains=4, max_draws=600, seed=5, tacs=[0, 2])
    s.run(draws=600, tune=1500)
    dvr, r1 = s.chains()
    ext = s.summary_ext()
    assert ext.shape == (2, 96, 4)
    worst = np.zeros(4)
    for tac in range(2):
        for coord in range(0, 96, 7):
            a = (dvr if coord < 48 else r1)[tac, :, :, coord % 48].astype(np.float64)
            lo, hi = dg.hdi(a)
            ref = np.array([lo, hi, dg.mcse_sd(a), dg.ess_sd(a)])
            got = ext[tac, coord].astype(np.float64)
            err = np.abs(got - ref) / np.abs(ref)
            err[:2] = np.abs(got[:2] - ref[:2])                 # hdi bounds are draws: exact up to fp32 storage
            worst = np.maximum(worst, err)
    print("worst err [hdi_lo hdi_hi (abs) mcse_sd ess_sd (rel)]:", worst)
    assert worst[0] < 1e-7 and worst[1] < 1e-7
    assert worst[2] < 5e-3 and worst[3] < 5e-3


def test_odd_number_of_draws_follows_arviz_split(dataset, prior):
    """ArviZ's _split_chains drops the middle draw of an odd-length chain for R-hat / ESS, pm.summary's mean / sd use
    every draw: both paths (stored draws, running moments) do the same."""
    from oracle import diagnostics as dg
    a = make_sampler(dataset, prior, n_chains=4, max_draws=301, seed=11, tacs=[1])
    a.run(draws=301, tune=600)
    dvr, r1 = a.chains()
    x = np.concatenate([dvr[0], r1[0]], axis=-1).astype(np.float64)      # (4, 301, 96)
    sm = a.summary()[0]
    assert np.abs(sm[:, 0] - x.mean(axis=(0, 1))).max() < 2e-6          # all 301 draws
    assert np.abs(sm[:, 1] / x.std(axis=(0, 1), ddof=1) - 1).max() < 1e-5
    for coord in (0, 50, 95):
        assert abs(sm[coord, 5] / dg.rhat_rank(x[:, :, coord]) - 1) < 1e-4
        assert abs(sm[coord, 3] / dg.ess_bulk(x[:, :, coord]) - 1) < 2e-3
    b = make_sampler(dataset, prior, n_chains=4, max_draws=0, seed=11, tacs=[1])
    b.run(draws=301, tune=600)
    mm = b.summary()[0]
    halves = np.concatenate([x[:, :150], x[:, 151:]], axis=0)            # middle draw (index 150) in neither half
    assert np.abs(mm[:, 0] - halves.mean(axis=(0, 1))).max() < 2e-6
    W = halves.var(axis=1, ddof=1).mean(axis=0)
    rhat = np.sqrt((149 / 150 * W + halves.mean(axis=1).var(axis=0, ddof=1)) / W)
    assert np.abs(mm[:, 5] / rhat - 1).max() < 1e-3


def test_moments_mode_batch_means_ess(dataset, prior):
    """The O(1)-memory summary: its batch-means ESS equals the numpy restatement on the same chains, and agrees with
    the stored-draw path's mean-ESS (ArviZ's estimator) within 15 % in the median and a factor 1.5 everywhere."""
    from oracle import diagnostics as dg
    n_draws, C = 4800, 16
    a = make_sampler(dataset, prior, n_chains=C, max_draws=n_draws, seed=31, tacs=[0])
    a.run(draws=n_draws, tune=2500)
    sr = a.summary()[0]
    dvr, r1 = a.chains()
    x = np.concatenate([dvr[0], r1[0]], axis=-1).astype(np.float64)
    b = make_sampler(dataset, prior, n_chains=C, max_draws=0, seed=31, tacs=[0])
    b.reset(); b.plan(draws=n_draws, tune=2500)
    for n in (700, 1800, 37, 1963, 2000, 800):                            # odd chunking: launches are cut at batch boundaries
        b.advance(n)
    sm = b.summary()[0]
    ref = np.array([dg.batch_means_ess(x[:, :, k]) for k in range(96)])
    assert np.abs(sm[:, 3] / ref - 1).max() < 2e-3, "batch-means ESS differs from its numpy restatement"
    ess_mean_rank_path = (sr[:, 1].astype(np.float64) / sr[:, 2]) ** 2    # mcse_mean = sd / sqrt(ess_mean)
    ratio = sm[:, 3] / ess_mean_rank_path
    print("batch-means ESS / ArviZ mean-ESS: median %.3f min %.3f max %.3f" % (np.median(ratio), ratio.min(), ratio.max()))
    # (batch means are biased high by ~ tau / B: here B = 300 draws, the slowest coordinates have tau ~ 100)
    assert abs(np.median(ratio) - 1) < 0.15 and ratio.min() > 1 / 1.6 and ratio.max() < 1.6
    assert np.abs(sm[:, 2] / (sm[:, 1] / np.sqrt(sm[:, 3])) - 1).max() < 1e-5
    assert np.abs(sm[:, 0] - sr[:, 0]).max() < 2e-6 and np.isnan(sm[:, 4]).all()


def test_checkpoint_resume_inside_a_tuning_window(dataset, prior):
    """A run cut at sweep 250 (mid tuning window: the PyMC accept counters matter) and at sweep 430 (mid draws: moments,
    accepted-move counters and stored draws matter), restored into FRESH handles, is bit-identical to the uninterrupted run."""
    from pet_posterior_distribution_b200 import PetmhError
    kw = dict(n_chains=6, max_draws=50, seed=17, tacs=[0, 3])
    a = make_sampler(dataset, prior, **kw)
    a.run(draws=100, tune=400, thin=2)
    ref = (a.chains(), a.summary(), a.state())
    b = make_sampler(dataset, prior, **kw)
    b.reset(); b.plan(draws=100, tune=400, thin=2)
    b.advance(250)
    blob = b.checkpoint()
    c = make_sampler(dataset, prior, **kw)
    c.restore(blob)
    c.advance(180)
    blob2 = c.checkpoint()
    d = make_sampler(dataset, prior, **kw)
    d.restore(blob2)
    d.advance(70)
    got = (d.chains(), d.summary(), d.state())
    assert np.array_equal(got[0][0], ref[0][0]) and np.array_equal(got[0][1], ref[0][1])
    assert np.array_equal(got[1], ref[1], equal_nan=True)
    assert np.array_equal(got[2][0], ref[2][0]) and np.array_equal(got[2][1], ref[2][1])
    # moments mode too (no stored draws)
    kw0 = dict(n_chains=4, max_draws=0, seed=17, tacs=[1])
    e = make_sampler(dataset, prior, **kw0)
    e.run(draws=160, tune=300)
    f = make_sampler(dataset, prior, **kw0)
    f.reset(); f.plan(draws=160, tune=300); f.advance(333)
    g = make_sampler(dataset, prior, **kw0)
    g.restore(f.checkpoint()); g.advance(127)
    # (a batch cut by the checkpoint is accumulated in two launches: same numbers, another rounding of the fp32 moments)
    assert np.allclose(g.summary(), e.summary(), rtol=2e-4, atol=0, equal_nan=True)
    # a blob that does not fit the handle is refused
    h = make_sampler(dataset, prior, n_chains=5, max_draws=50, seed=17, tacs=[0, 3])
    with pytest.raises(PetmhError):
        h.restore(blob)
    with pytest.raises(PetmhError):
        make_sampler(dataset, prior, n_chains=6, max_draws=50, seed=18, tacs=[0, 3]).restore(blob)


def test_global_ids_make_results_independent_of_batching_and_chain_sharding(dataset, prior):
    """Streams are keyed by (global TAC id, global chain id): a sample's chains do not depend on which other samples
    share the batch (mcmc.py's skip rule leaves gaps), and the chains of one TAC can be split over handles (ranks)."""
    full = make_sampler(dataset, prior, n_chains=8, max_draws=30, seed=9, tacs=[0, 1, 2, 3])
    full.run(draws=30, tune=120)
    dv, r1 = full.chains()
    part = make_sampler(dataset, prior, n_chains=8, max_draws=30, seed=9, tacs=[1, 3])      # samples 0 and 2 already done
    part.set_global_ids(np.array([1, 3], np.uint64))
    part.run(draws=30, tune=120)
    pv, p1 = part.chains()
    assert np.array_equal(pv, dv[[1, 3]]) and np.array_equal(p1, r1[[1, 3]])
    # chains 3..7 of sample 2 on "another rank"
    sh = make_sampler(dataset, prior, n_chains=5, max_draws=30, seed=9, tacs=[2])
    sh.set_global_ids(np.array([2], np.uint64), chain_gid0=3, chains_per_tac_global=8)
    sh.run(draws=30, tune=120)
    sv, s1 = sh.chains()
    assert np.array_equal(sv[0], dv[2, 3:]) and np.array_equal(s1[0], r1[2, 3:])


def test_gathered_summaries_equal_the_single_handle_ones(dataset, prior):
    """petmh_summary_from_draws_device / _from_moments_device (the chain-sharded multi-GPU path) on the state of one
    handle reproduce petmh_get_summary / petmh_get_summary_ext bit for bit."""
    import ctypes as C
    import torch
    from pet_posterior_distribution_b200 import _lib
    from pet_posterior_distribution_b200.distributed import _summary_inputs
    for max_draws in (200, 0):
        s = make_sampler(dataset, prior, n_chains=6, max_draws=max_draws, seed=4, tacs=[2])
        s.run(draws=200, tune=300)
        ref = s.summary()
        v = _summary_inputs(s, 0)
        ctr = v["counters"]
        out = torch.zeros((1, 96, 8), device="cuda:0")
        ext = torch.zeros((1, 96, 4), device="cuda:0")
        vp = lambda x: C.c_void_p(x.data_ptr())
        if max_draws:
            d = v["draws"].contiguous()
            rc = _lib.lib.petmh_summary_from_draws_device(0, vp(d), 1, 6, ctr[0], vp(v["nacc"]), vp(v["scale"]), ctr[7], vp(out), vp(ext), None)
            assert rc == 0
            torch.cuda.synchronize()
            assert np.array_equal(ext.cpu().numpy(), s.summary_ext(), equal_nan=True)
        else:
            nh, nb = (C.c_int * 2)(ctr[2], ctr[3]), (C.c_int * 2)(ctr[4], ctr[5])
            rc = _lib.lib.petmh_summary_from_moments_device(0, vp(v["mom"]), vp(v["mu"]), 1, 6, nh, nb, ctr[6], vp(v["nacc"]),
                                                            vp(v["scale"]), ctr[7], vp(out), None)
            assert rc == 0
            torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), ref, equal_nan=True)


def test_synth_reports_rejection_caps(prior, dataset):
    """A prior whose draws are almost never positive exhausts the generator's redraw caps: petmh_synth says so
    (PETMH_ESYNTH) instead of silently keeping an invalid draw, and marks the TACs in attempts[]."""
    from pet_posterior_distribution_b200 import MHSampler, PetmhError
    from pet_posterior_distribution_b200 import sample_sim_data as gen
    t, dt = gen.frame_grid()
    sig = gen.noise_table(np.random.default_rng(3), 0.1, t, dt)
    s = MHSampler(n_chains=2, max_tacs=3, seed=1)
    s.set_frames(t, dt)
    s.set_prior(prior["mu_DVR"] - 3.0, prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])      # DVR ~ -2 +- 0.35
    with pytest.raises(PetmhError) as e:
        s.synth(3, 5, prior["mu_tac_ref"], prior["Cov_tac_ref"], float(prior["mu_k2p"]), sig)
    assert e.value.code == -6 and "rejection cap" in str(e.value)
    assert (s.synth_get()["attempts"] < 0).all()
    s2 = MHSampler(n_chains=2, max_tacs=3, seed=1)
    s2.set_frames(t, dt)
    s2.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
    s2.synth(3, 5, prior["mu_tac_ref"], prior["Cov_tac_ref"], float(prior["mu_k2p"]), sig)
    assert (s2.synth_get()["attempts"] > 0).all()


def test_two_frame_grids_in_one_process(dataset, prior):
    """The frame-time table lives in the handle (no process-global constant memory): a second handle with a different
    (scaled) grid works next to the first, on the same operator sparsity."""
    a = make_sampler(dataset, prior, n_chains=2, max_draws=0, seed=1, tacs=[0])
    from pet_posterior_distribution_b200 import MHSampler
    b = MHSampler(n_chains=2, max_tacs=1, seed=1)
    b.set_frames(dataset["time_vector"] * 1.5, dataset["dt"] * 1.5)        # same pattern, other times
    b.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
    y = dataset["tac_noisy_sampled"][:1] / dataset["dt"][None, None, :]
    b.set_data(y, dataset["vartacref"][:1], dataset["vark2p"][:1], dataset["sigma_noise"])
    DVR, R1 = dataset["varDVR"][0], dataset["varR1"][0]
    fa, fb = a.forward(0, DVR, R1), b.forward(0, DVR, R1)
    from oracle import forward
    ra = forward.srtm2_tac(dataset["time_vector"], dataset["vartacref"][0], DVR, R1, float(dataset["vark2p"][0])).T
    rb = forward.srtm2_tac(dataset["time_vector"] * 1.5, dataset["vartacref"][0], DVR, R1, float(dataset["vark2p"][0])).T
    assert np.abs(fa / ra - 1).max() < 1e-5 and np.abs(fb / rb - 1).max() < 1e-5
    assert np.abs(fa - fb).max() > 1e-3


def test_input_scale_is_checked(dataset, prior):
    """Data far outside the supported units (or non-finite) are refused instead of silently freezing chains."""
    from pet_posterior_distribution_b200 import PetmhError
    s = make_sampler(dataset, prior, n_chains=2, tacs=[0])
    y = dataset["tac_noisy_sampled"][:1] / dataset["dt"][None, None, :]
    for bad in (y * 1e9, np.where(np.arange(y.size).reshape(y.shape) == 7, np.nan, y)):
        with pytest.raises(PetmhError) as e:
            s.set_data(bad, dataset["vartacref"][:1], dataset["vark2p"][:1], dataset["sigma_noise"])
        assert "scale" in str(e.value)
        with pytest.raises(PetmhError):
            s.run(draws=2, tune=2)                      # no data bound after the refusal
    s.set_data(y.astype(np.float32), dataset["vartacref"][:1].astype(np.float32), dataset["vark2p"][:1], dataset["sigma_noise"].astype(np.float32))
    s.run(draws=2, tune=2)
