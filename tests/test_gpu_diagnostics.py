"""K3 on the GPU vs the numpy restatement of ArviZ (oracle/diagnostics.py): rank-normalised
split R-hat, bulk/tail ESS, MCSE, mean, sd computed from the kernel's own stored chains."""
import numpy as np
import pytest

from conftest import make_sampler

pytestmark = pytest.mark.gpu


def test_rank_summary_matches_oracle(dataset, prior):
    from oracle import diagnostics as dg
    s = make_sampler(dataset, prior, n_chains=4, max_draws=600, seed=5, tacs=[0, 2])
    s.run(draws=600, tune=1500)
    dvr, r1 = s.chains()                              # (2, 4, 600, 48)
    summ = s.summary()                                # (2, 96, 8)
    assert summ.shape == (2, 96, 8)
    worst = np.zeros(6)
    for tac in range(2):
        for coord in list(range(0, 96, 7)):
            a = (dvr if coord < 48 else r1)[tac, :, :, coord % 48].astype(np.float64)
            ref = dg.summary_row(a)
            got = summ[tac, coord, :6].astype(np.float64)
            rel = np.abs(got - ref) / np.abs(ref)
            worst = np.maximum(worst, rel)
    print("worst rel err [mean sd mcse ess_bulk ess_tail rhat]:", worst)
    assert worst[0] < 1e-6 and worst[1] < 1e-5         # mean, sd
    assert worst[5] < 1e-4                              # r_hat
    assert (worst[2:5] < 2e-3).all()                    # mcse / ess (fp32 z-scores, fp32 rho storage)
    acc = summ[..., 6]
    assert ((acc > 0.05) & (acc < 0.8)).all()
    assert (summ[..., 7] > 0).all() and (summ[..., 7] < 1).all()      # tuned scalings


def test_moments_summary_consistent_with_chains(dataset, prior):
    """moments-only mode (max_draws = 0) vs the same run's stored chains: mean, sd, classic split R-hat."""
    a = make_sampler(dataset, prior, n_chains=8, max_draws=400, seed=21, tacs=[1])
    a.run(draws=400, tune=1000)
    dvr, r1 = a.chains()
    b = make_sampler(dataset, prior, n_chains=8, max_draws=0, seed=21, tacs=[1])
    b.run(draws=400, tune=1000)
    sm = b.summary()[0]
    x = np.concatenate([dvr[0], r1[0]], axis=-1).astype(np.float64)     # (8, 400, 96)
    assert np.abs(sm[:, 0] - x.mean(axis=(0, 1))).max() < 2e-6
    assert np.abs(sm[:, 1] / x.std(axis=(0, 1), ddof=1) - 1).max() < 1e-4
    halves = np.concatenate([x[:, :200], x[:, 200:]], axis=0)            # (16, 200, 96)
    W = halves.var(axis=1, ddof=1).mean(axis=0)
    B_over_n = halves.mean(axis=1).var(axis=0, ddof=1)
    rhat = np.sqrt((199 / 200 * W + B_over_n) / W)
    assert np.abs(sm[:, 5] / rhat - 1).max() < 1e-3
    assert np.isnan(sm[:, 4]).all() and (sm[:, 3] > 1).all()


def test_posterior_recovers_truth(dataset, prior):
    """Sanity: with 16 tuned chains the truth lies within a few posterior SDs for every ROI and
    R-hat is close to 1 (converged)."""
    s = make_sampler(dataset, prior, n_chains=16, max_draws=500, seed=3, tacs=[0])
    s.run(draws=2000, tune=3000, thin=4)
    sm = s.summary()[0]
    truth = np.concatenate([dataset["varDVR"][0], dataset["varR1"][0]])
    zs = (sm[:, 0] - truth) / sm[:, 1]
    print("max |z| of truth:", np.abs(zs).max(), " max rhat:", sm[:, 5].max(), " min ess_bulk:", sm[:, 3].min())
    assert np.abs(zs).max() < 5.0
    assert sm[:, 5].max() < 1.2


def test_many_chains_config4_shape(dataset, prior):
    """BASELINE configs[3]: 1024 chains per TAC -- sampler, stored draws and rank diagnostics all handle it."""
    s = make_sampler(dataset, prior, n_chains=1024, max_draws=100, seed=8, tacs=[2])
    s.run(draws=6000, tune=3000, thin=60)
    sm = s.summary()[0]
    print("1024 chains: rhat max %.3f ess_bulk min %.0f" % (sm[:, 5].max(), sm[:, 3].min()))
    assert np.isfinite(sm[:, :6]).all() and (sm[:, 3] > 100).all() and sm[:, 5].max() < 1.5
    dvr, _ = s.chains()
    assert dvr.shape == (1, 1024, 100, 48)


def test_tfp_cross_chain_ess_matches_oracle(dataset, prior):
    """petmh_get_ess_cross_chain vs the numpy restatement of tfp.mcmc.effective_sample_size(cross_chain_dims)."""
    from oracle import diagnostics as dg
    s = make_sampler(dataset, prior, n_chains=4, max_draws=1500, seed=9, tacs=[0, 3])
    s.run(draws=1500, tune=1500)
    dvr, r1 = s.chains()
    ess = s.ess_cross_chain()
    assert ess.shape == (2, 96) and np.isfinite(ess).all() and (ess > 1).all() and (ess <= 4 * 1500 * 1.0001).all()
    worst = 0.0
    for tac in range(2):
        for coord in range(0, 96, 5):
            a = (dvr if coord < 48 else r1)[tac, :, :, coord % 48].astype(np.float64)
            ref = dg.tfp_ess_cross_chain(a)
            worst = max(worst, abs(ess[tac, coord] / ref - 1))
    print("worst rel err of cross-chain ESS:", worst)
    assert worst < 2e-3
    # one chain: single-chain formula
    s1 = make_sampler(dataset, prior, n_chains=1, max_draws=800, seed=9, tacs=[0])
    s1.run(draws=800, tune=1500)
    d1, _ = s1.chains()
    e1 = s1.ess_cross_chain()
    ref = dg.tfp_ess_cross_chain(d1[0, :, :, 7].astype(np.float64))
    assert abs(e1[0, 7] / ref - 1) < 2e-3
    # no stored draws -> loud error
    s0 = make_sampler(dataset, prior, n_chains=4, max_draws=0, seed=9, tacs=[0])
    s0.run(draws=50, tune=100)
    with pytest.raises(RuntimeError):
        s0.ess_cross_chain()


def test_posterior_cov_matches_numpy(dataset, prior):
    """petmh_get_posterior_cov == np.cov / np.corrcoef of the pooled stored draws (what main_script.py:717-738 computes from
    DVR_mcmc / R1_mcmc), with unused draw slots (n_stored < max_draws) and a ragged last tile."""
    from pet_posterior_distribution_b200 import PetmhError
    s = make_sampler(dataset, prior, n_chains=3, max_draws=400, seed=9, tacs=[1, 3])
    s.run(draws=331, tune=700)
    dvr, r1 = s.chains()                              # (2, 3, 331, 48): 993 pooled draws = 15 tiles of 64 + 33
    cov, corr = s.posterior_cov()
    assert cov.shape == (2, 2, 48, 48) and corr.shape == cov.shape
    for tac in range(2):
        for b, arr in enumerate((dvr, r1)):
            x = arr[tac].reshape(-1, 48).astype(np.float64)
            ref, rc = np.cov(x, rowvar=False), np.corrcoef(x, rowvar=False)
            sd = np.sqrt(np.diag(ref))
            assert np.abs(cov[tac, b] - ref).max() <= 1e-5 * np.outer(sd, sd).max()
            assert (np.abs(cov[tac, b] - ref) <= 1e-5 * np.outer(sd, sd) + 1e-300).all()
            assert np.abs(corr[tac, b] - rc).max() < 1e-5 and np.abs(np.diag(corr[tac, b]) - 1).max() < 1e-6
            assert np.array_equal(cov[tac, b], cov[tac, b].T)
    s.close()
    s0 = make_sampler(dataset, prior, n_chains=2, max_draws=0, tacs=[0])     # moments mode: no stored draws
    s0.run(draws=10, tune=10)
    with pytest.raises(PetmhError):
        s0.posterior_cov()
    s0.close()
