"""CPU oracle for the MH-MCMC / SRTM2 hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU (numpy fp64 + a small C file) restatement of the reference
algorithm of yanisdjebra/PET_posterior_distribution for the one path this repo
accelerates (mcmc.py Metropolis-Hastings over kinetic_model.SRTM2).

It is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it.  Nothing
under ``pet_posterior_distribution_b200/`` imports it; the product path raises if the
CUDA library is missing.

Parity status (see DESIGN.md "Oracle"):
  * forward model (kinetic_model.py:12-57,134-158) -- PINNED: checked against the live
    reference ``kinetic_model.py`` in the build container; golden vectors committed in
    ``tests/golden/forward_golden.npz`` (made by ``tools/make_golden.py``).
  * generator (sample_sim_data.py:96-224, helper_func.py:146-162) -- PINNED on outputs of the live reference script
    exec'd in the build container: ``tests/golden/reference_generated_*.npz`` (``tools/make_reference_generated.py``,
    ``tests/test_reference_generated.py``).
  * the posterior the path must output -- checked statistically against a sampler that shares nothing with the path's
    algorithm (full-covariance random-walk Metropolis over scipy's densities and the live reference forward model:
    ``tools/make_independent_posterior.py``, ``tests/test_independent_posterior.py``).
  * PyMC model log-probability / Metropolis semantics / ArviZ diagnostics
    (mcmc.py:147-157,181-187; pymc==5.12.0, arviz unpinned) -- PARITY UNPINNED: those
    packages are third-party, absent from /root/reference and not installable offline;
    the reference ships no tests and its chain pickles are Git-LFS pointers.  The
    restatement follows the published algorithms and is cross-checked against
    scipy.stats (truncnorm / multivariate_normal) only; ``tests/test_pymc_pin.py`` pins it wherever those packages exist.
"""
