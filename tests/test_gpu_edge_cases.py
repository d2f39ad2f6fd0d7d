"""Edge cases of the sampler on a real GPU: odd / single chain counts, thinning and capacity,
degenerate data, resume, error behaviour of the C ABI."""
import numpy as np
import pytest

from conftest import make_sampler

pytestmark = pytest.mark.gpu


def test_chain_streams_do_not_depend_on_chain_count(dataset, prior, monkeypatch):
    """Chain c of TAC 0 has Philox gid c whatever n_chains is: its draws are identical for
    n_chains = 1, 3, 4, 5, 6, 10, 12 (different CTA shapes, idle half-warps; 5/6/10/12 chains need 96/160/192
    threads, which the small-job CTA shrink must not halve to a fraction of a warp) -> the thread mapping is
    invisible.  Checked for the normal kernel (PETMH_WIDE=0) and for the automatic choice."""
    for wide in ("0", None):
        if wide is None:
            monkeypatch.delenv("PETMH_WIDE", raising=False)
        else:
            monkeypatch.setenv("PETMH_WIDE", wide)
        runs = {}
        for C in (1, 3, 4, 5, 6, 10, 12):
            s = make_sampler(dataset, prior, n_chains=C, max_draws=30, seed=5, tacs=[0])
            s.run(draws=30, tune=100)
            runs[C] = s.chains()
            s.close()
            assert runs[C][0].shape == (1, C, 30, 48) and np.isfinite(runs[C][0]).all()
        ref = runs[12]
        for C in (1, 3, 4, 5, 6, 10):
            assert np.array_equal(runs[C][0][0], ref[0][0, :C]) and np.array_equal(runs[C][1][0], ref[1][0, :C]), C
        assert not np.array_equal(ref[0][0, 0], ref[0][0, 1])
    # sub-wave jobs of several TACs pick odd CTA sizes (petmh.cu threads_per_cta: any multiple of 32 threads)
    monkeypatch.setenv("PETMH_WIDE", "0")
    big = {}
    for C in (22, 44, 52):
        s = make_sampler(dataset, prior, n_chains=C, max_draws=10, seed=5, tacs=[0, 1, 2, 3, 0, 1, 2])
        s.run(draws=10, tune=120)
        big[C] = s.chains()
        s.close()
    for C in (22, 44):   # (gid = TAC index * n_chains + chain: only TAC 0's streams are independent of n_chains)
        assert np.array_equal(big[C][0][0], big[52][0][0, :C]) and np.array_equal(big[C][1][0], big[52][1][0, :C]), C
        assert np.isfinite(big[C][0]).all()


def test_thinning_and_capacity(dataset, prior):
    full = make_sampler(dataset, prior, n_chains=2, max_draws=60, seed=9, tacs=[1])
    full.run(draws=60, tune=100)
    d_full = full.chains()[0]
    thin = make_sampler(dataset, prior, n_chains=2, max_draws=20, seed=9, tacs=[1])
    thin.run(draws=60, tune=100, thin=3)
    assert thin.n_stored == 20
    assert np.array_equal(thin.chains()[0], d_full[:, :, ::3])
    cap = make_sampler(dataset, prior, n_chains=2, max_draws=10, seed=9, tacs=[1])
    cap.run(draws=60, tune=100)                       # more draws than capacity: the first 10 are kept
    assert cap.n_stored == 10 and np.array_equal(cap.chains()[0], d_full[:, :, :10])
    none = make_sampler(dataset, prior, n_chains=2, max_draws=0, seed=9, tacs=[1])
    none.run(draws=60, tune=100)
    from pet_posterior_distribution_b200 import PetmhError
    with pytest.raises(PetmhError):
        none.chains()
    assert np.isfinite(none.summary()[..., :2]).all()


def test_negative_observation_freezes_the_chain(dataset, prior):
    """y < 0 is outside the truncated support: the reference's log-probability is -inf, every
    delta_logp is NaN and metrop_select rejects everything -- the chain stays at the prior mean."""
    from pet_posterior_distribution_b200 import MHSampler
    s = MHSampler(n_chains=2, max_tacs=2, max_draws=5, seed=1)
    s.set_frames(dataset["time_vector"], dataset["dt"])
    s.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
    y = dataset["tac_noisy_sampled"][:2] / dataset["dt"][None, None, :]
    y[0, 7, 20] = -0.5
    s.set_data(y, dataset["vartacref"][:2], dataset["vark2p"][:2], dataset["sigma_noise"])
    s.run(draws=5, tune=100)
    dvr, r1 = s.chains()
    assert (dvr[0] == prior["mu_DVR"].astype(np.float32)).all() and (r1[0] == prior["mu_R1"].astype(np.float32)).all()
    assert (dvr[1] != prior["mu_DVR"].astype(np.float32)).any()
    ll, _ = s.loglik(0, prior["mu_DVR"], prior["mu_R1"])
    assert np.isneginf(ll).all()


def test_state_roundtrip_resume(dataset, prior):
    a = make_sampler(dataset, prior, n_chains=4, max_draws=0, seed=3, tacs=[0, 1])
    a.run(draws=0, tune=300)
    q, sc = a.state()
    assert q.shape == (2, 4, 96) and (sc > 0).all() and (sc < 1).all()
    b = make_sampler(dataset, prior, n_chains=4, max_draws=0, seed=3, tacs=[0, 1])
    b.set_state(q, sc, sweep=300)
    b.plan(draws=50, tune=300)
    a.plan(draws=50, tune=300)
    a.advance(50)
    b.advance(50)
    qa, sa = a.state()
    qb, sb = b.state()
    assert np.array_equal(qa, qb) and np.array_equal(sa, sb)       # warm start == continued run
    assert np.array_equal(sa, sc)                                   # scaling frozen after tuning


def test_error_behaviour(dataset, prior):
    from pet_posterior_distribution_b200 import MHSampler, PetmhError
    s = MHSampler(n_chains=2, max_tacs=1)
    with pytest.raises(PetmhError) as e:
        s.run(10, 10)
    assert e.value.code == -1
    s.set_frames(dataset["time_vector"], dataset["dt"])
    s.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
    y = dataset["tac_noisy_sampled"][:2] / dataset["dt"][None, None, :]
    with pytest.raises(PetmhError):                                 # more TACs than capacity
        s.set_data(y, dataset["vartacref"][:2], dataset["vark2p"][:2], dataset["sigma_noise"])
    with pytest.raises(PetmhError):                                 # sigma_noise never given
        s.set_data(y[:1], dataset["vartacref"][:1], dataset["vark2p"][:1], None)
    with pytest.raises(PetmhError):                                 # not positive definite
        s.set_prior(prior["mu_DVR"], -np.eye(48), prior["mu_R1"], prior["Cov_R1"])
    ok = make_sampler(dataset, prior, n_chains=2, tacs=[0])
    ok.plan(10, 10)
    with pytest.raises(PetmhError):                                 # advance before reset / set_state
        ok.advance(5)
    ok.reset()
    ok.advance(5)
    with pytest.raises(PetmhError):
        MHSampler(n_chains=0)
    with pytest.raises(ValueError):
        s.set_frames(np.zeros(10), np.zeros(10))


def test_large_batch_shard_invariance_and_determinism(dataset, prior):
    """Size-independent properties at a batch large enough to fill the GPU many times over
    (8192 TACs x 16 chains = 2 M (chain, ROI) threads): same seed -> identical summaries; the batch cut
    into two shards with global TAC offsets (what multi-GPU sharding does) -> identical rows."""
    from pet_posterior_distribution_b200 import MHSampler
    S, C = 8192, 16
    k = dataset["varDVR"].shape[0]
    idx = np.arange(S) % k
    y = (dataset["tac_noisy_sampled"] / dataset["dt"][None, None, :]).astype(np.float32)[idx]
    cr = dataset["vartacref"].astype(np.float32)[idx]
    k2p = dataset["vark2p"].astype(np.float32)[idx]
    sig = dataset["sigma_noise"].astype(np.float32)

    def run(lo, hi):
        s = MHSampler(n_chains=C, max_tacs=hi - lo, seed=77, tac_gid0=lo)
        s.set_frames(dataset["time_vector"], dataset["dt"])
        s.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
        s.set_data(y[lo:hi], cr[lo:hi], k2p[lo:hi], sig)
        s.run(draws=20, tune=100)
        out = s.summary()
        s.close()
        return out

    full = run(0, S)
    assert np.isfinite(full[..., :2]).all()
    again = run(0, S)
    assert np.array_equal(full, again, equal_nan=True)
    a, b = run(0, 3000), run(3000, S)                          # ragged shards
    assert np.array_equal(np.concatenate([a, b]), full, equal_nan=True)
    # identical data, different global TAC index -> different Philox streams -> different chains
    assert not np.array_equal(full[0], full[k])


def test_wide_small_job_path_is_bit_identical(dataset, prior, monkeypatch):
    """Small jobs run the wide kernels (petmh_device.cuh WIDE: 1 = three warps per chain pair, one ROI slot each;
    2 = nine warps per pair, one row block of one slot each): same random numbers, same per-item arithmetic ->
    chains, summaries and resumed state equal the normal kernel's bit for bit.  Covers a ragged chain count (odd:
    half of the last pair idles) and CTAs of 1..4 triples."""
    for n_chains, tacs in [(5, [0, 1]), (64, [2]), (16, list(range(24))), (32, list(range(20)))]:
        out = {}
        for wide in ("0", "1", "2"):
            monkeypatch.setenv("PETMH_WIDE", wide)
            s = make_sampler(dataset, prior, n_chains=n_chains, max_draws=40, seed=11, tacs=[t % 4 for t in tacs])
            s.run(draws=40, tune=230)            # crosses two tuning boundaries and the 200-sweep launch chunk
            dvr, r1 = s.chains()
            q, sc = s.state()
            out[wide] = (dvr.copy(), r1.copy(), q.copy(), sc.copy(), s.summary().copy())
            s.close()
        for w in ("1", "2"):
            for a, b in zip(out["0"], out[w]):
                assert np.array_equal(a, b, equal_nan=True)
        assert np.isfinite(out["1"][0]).all()
