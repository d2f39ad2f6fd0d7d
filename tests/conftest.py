import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def prior():
    z = np.load(os.path.join(GOLDEN, "prior_stats_nROI48.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def forward_golden():
    z = np.load(os.path.join(GOLDEN, "forward_golden.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def dataset():
    z = np.load(os.path.join(GOLDEN, "dataset_s0.1.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def models(dataset, prior):
    """oracle Models for the golden dataset's TACs (mcmc.py:79-80,106-112 set-up)."""
    from oracle.logp import Model
    out = []
    for s in range(dataset["varDVR"].shape[0]):
        y = dataset["tac_noisy_sampled"][s] / dataset["dt"][None, :]
        out.append(Model(dataset["time_vector"], dataset["vartacref"][s], dataset["vark2p"][s], y,
                         dataset["sigma_noise"], prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"]))
    return out


def make_sampler(dataset, prior, n_chains=4, max_draws=0, seed=1234, tacs=None, **kw):
    from pet_posterior_distribution_b200 import MHSampler
    idx = list(range(dataset["varDVR"].shape[0])) if tacs is None else list(tacs)
    s = MHSampler(n_chains=n_chains, max_tacs=len(idx), max_draws=max_draws, seed=seed, **kw)
    s.set_frames(dataset["time_vector"], dataset["dt"])
    s.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
    y = dataset["tac_noisy_sampled"][idx] / dataset["dt"][None, None, :]
    s.set_data(y, dataset["vartacref"][idx], dataset["vark2p"][idx], dataset["sigma_noise"])
    return s
