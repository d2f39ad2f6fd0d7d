#!/usr/bin/env python
"""BASELINE configs[2] end to end through the drop-in entry point: a 100-sample test set (test-style generator =
restated sample_sim_data.py with the Mahalanobis rule) x 48 ROIs x 256 chains, sharded over the GPUs of the box:

  python tools/run_config2.py prepare /tmp/c2                   # writes sim_data/ + prior_stats_nROI48.pik (one GPU)
  torchrun --nproc-per-node 8 ... tools/run_config2.py run /tmp/c2 [--chains 256 --iter 20000 --burn 40000 --thin 100]

`run` calls pet_posterior_distribution_b200.mcmc exactly as `python -m ...mcmc` would (the reference's script seam) and
rank 0 reports seconds, chain-steps/s per GPU and the files written."""
import glob, os, pickle, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def prepare(root, n=100):
    from pet_posterior_distribution_b200 import sample_sim_data as gen
    prior = gen.load_prior()
    ds = gen.generate(prior, n, 0.1, test_style=True, seed=20251019)
    d = os.path.join(root, "sim_data", "nROI48", "26-01-01_00-00-00_test")
    os.makedirs(d, exist_ok=True)
    pickle.dump(ds, open(os.path.join(d, "data_nROI48_n100_s1.0e-01.pik"), "wb"))
    pickle.dump(prior, open(os.path.join(root, "prior_stats_nROI48.pik"), "wb"))
    print("wrote", d)


def run(root, argv):
    from pet_posterior_distribution_b200 import mcmc
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    t0 = time.time()
    written = mcmc._cli(argv + ["--data-dir", os.path.join(root, "sim_data"), "--prior", os.path.join(root, "prior_stats_nROI48.pik")])
    dt = time.time() - t0
    n_s = mcmc.sample_range.stop - mcmc.sample_range.start
    steps = n_s * mcmc.chains * 96 * (mcmc.iter_mcmc + mcmc.burn_mcmc)
    print("rank %d/%d: %d files written in %.1f s" % (rank, world, len(written), dt), flush=True)
    if rank == 0:
        files = glob.glob(os.path.join(root, "sim_data", "nROI48", "*_test", "MCMC_s1.0e-01", "*.pik"))
        rh = [ln for ln in open(os.path.join(os.path.dirname(files[0]), "rhat_less_than_102.txt"))] if files and os.path.isfile(
            os.path.join(os.path.dirname(files[0]), "rhat_less_than_102.txt")) else []
        print("configs[2]: %d samples x %d chains x (%d tune + %d draws, thin %d) on %d GPU(s): %.1f s wall incl. file output "
              "(%.3e chain-steps/s overall, %.3e per GPU); %d pickles on disk; %d samples logged with r_hat > 1.02"
              % (n_s, mcmc.chains, mcmc.burn_mcmc, mcmc.iter_mcmc, mcmc.thin, world, dt, steps / dt, steps / dt / world, len(files), len(rh)), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "prepare":
        prepare(sys.argv[2])
    else:
        run(sys.argv[2], sys.argv[3:])
