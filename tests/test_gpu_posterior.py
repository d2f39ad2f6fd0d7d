"""Posterior DVR / R1 means and SDs of the GPU sampler vs the CPU oracle's free-running
chains (tests/golden/oracle_posterior_tac0.npz, tools/make_oracle_posterior.py):
agreement within 3 Monte-Carlo standard errors (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, make_sampler

pytestmark = pytest.mark.gpu


def test_moments_within_3_mcse_of_oracle(dataset, prior):
    from oracle import diagnostics as dg
    ref = np.load(os.path.join(GOLDEN, "oracle_posterior_tac0.npz"))
    s = make_sampler(dataset, prior, n_chains=64, max_draws=2000, seed=2024, tacs=[int(ref["tac"])])
    s.run(draws=4000, tune=2000, thin=2)
    sm = s.summary()[0]
    dvr, r1 = s.chains()
    x = np.concatenate([dvr[0], r1[0]], axis=-1).astype(np.float64)      # (64, 2000, 96)
    # means
    z_mean = (sm[:, 0] - ref["mean"]) / np.sqrt(sm[:, 2].astype(np.float64) ** 2 + ref["mcse_mean"] ** 2)
    # SDs: MCSE of the GPU SD from the same ArviZ formula the oracle's golden used
    mcse_sd_gpu = np.array([dg.mcse_sd(x[:, :, k]) for k in range(0, 96)])
    z_sd = (sm[:, 1] - ref["sd"]) / np.sqrt(mcse_sd_gpu ** 2 + ref["mcse_sd"] ** 2)
    print("mean: max|z| %.2f rms %.2f | sd: max|z| %.2f rms %.2f | gpu rhat max %.3f ess_bulk min %.0f"
          % (np.abs(z_mean).max(), np.sqrt((z_mean ** 2).mean()), np.abs(z_sd).max(), np.sqrt((z_sd ** 2).mean()),
             sm[:, 5].max(), sm[:, 3].min()))
    assert np.abs(z_mean).max() < 3.0, "posterior means differ from the oracle by more than 3 MCSE"
    assert np.abs(z_sd).max() < 3.0, "posterior SDs differ from the oracle by more than 3 MCSE"
    assert sm[:, 5].max() < 1.05
    # tuned proposal scales land in the same place (median over chains, factor 1.5)
    sc_gpu = np.median(s.state()[1][0], axis=0)
    sc_ref = np.median(ref["scale"], axis=0)
    assert (np.abs(np.log(sc_gpu / sc_ref)) < np.log(1.6)).all()
