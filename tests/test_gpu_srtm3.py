"""SURVEY.md 8 f3 on the GPU: the k2-free SRTM as a sampled three-block model (petmh_srtm_sample) against oracle/srtm3.py:
teacher-forced accept/reject decisions on a shared tape, and a free run whose k2 posterior finds k2p R1."""
import numpy as np
import pytest

from conftest import make_sampler

pytestmark = pytest.mark.gpu


def _k2_prior(dataset, prior):
    k2p = float(dataset["vark2p"][0])
    mu = k2p * prior["mu_R1"]
    return mu, np.diag((0.3 * mu) ** 2)


def test_taped_decisions_match_oracle(dataset, prior, models):
    from oracle import srtm3
    s = make_sampler(dataset, prior, n_chains=2, max_draws=0, seed=3, tacs=[1])
    mu_k2, cov_k2 = _k2_prior(dataset, prior)
    m3 = srtm3.Model3(models[1], mu_k2, cov_k2)
    n_sweeps, tune = 230, 200
    rng = np.random.default_rng(5)
    tapes = [srtm3.random_tape(n_sweeps, rng) for _ in range(2)]
    out = s.sample_srtm(mu_k2, cov_k2, draws=n_sweeps - tune, tune=tune,
                        tape=tuple(np.stack([t[k] for t in tapes]) for k in range(3)), tac=0)
    total = agree = 0
    worst = 0.0
    for c in range(2):
        ref = srtm3.run_chain(m3, tapes[c], tune, n_sweeps - tune, forced_draws=out["draws"][c])
        dec = ~ref["undecidable"]
        total += int(dec.sum())
        agree += int((ref["accept"][dec] == ref["forced_accept"][dec]).sum())
        near = dec & np.isfinite(ref["delta"]) & (np.abs(ref["delta"]) < 50)
        worst = max(worst, float(np.abs(out["delta"][c][near] - ref["delta"][near]).max()))
    print("SRTM (k2 free): %d of %d taped decisions identical; worst |Delta_gpu - Delta_fp64| %.2e" % (agree, total, worst))
    assert total > 60000 and agree >= 0.9999 * total
    assert worst < 5e-3


def test_free_run_recovers_k2(dataset, prior):
    """Data simulated with k2 = k2p R1 (SRTM2): with a loose prior on k2 the posterior of k2 / R1 sits at k2p."""
    from oracle import diagnostics as dg
    s = make_sampler(dataset, prior, n_chains=8, max_draws=0, seed=11, tacs=[0])
    mu_k2, cov_k2 = _k2_prior(dataset, prior)
    out = s.sample_srtm(mu_k2, cov_k2, draws=3000, tune=3000, thin=3)
    assert out["k2"].shape == (1, 8, 1000, 48) and np.isfinite(out["DVR"]).all()
    assert ((out["accept_rate"] > 0.05) & (out["accept_rate"] < 0.8)).all()
    k2p = float(dataset["vark2p"][0])
    ratio = (out["k2"][0] / out["R1"][0]).reshape(-1, 48)
    z = (ratio.mean(0) - k2p) / ratio.std(0)
    rhat = np.array([dg.rhat_rank(out["k2"][0, :, :, i].astype(np.float64)) for i in range(0, 48, 6)])
    print("k2 / R1 vs k2p: max |z| %.2f; r_hat(k2) max %.3f" % (np.abs(z).max(), rhat.max()))
    assert np.abs(z).max() < 4.5 and rhat.max() < 1.3
    # same seed -> same chains; the SRTM2 state of the handle is untouched
    again = s.sample_srtm(mu_k2, cov_k2, draws=3000, tune=3000, thin=3)
    assert np.array_equal(again["k2"], out["k2"])
