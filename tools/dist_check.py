#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU): TAC-sharded posterior summaries,
gathered with one NCCL all-gather, must equal the single-GPU result bit for bit (Philox streams
are keyed by the GLOBAL TAC index, so results do not depend on the number of GPUs)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
from pet_posterior_distribution_b200.distributed import run_sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
pr = np.load(os.path.join(g, "prior_stats_nROI48.npz")); ds = np.load(os.path.join(g, "dataset_s0.1.npz"))
S = 7                                                    # ragged over 2/4/8 ranks
idx = np.arange(S) % 4
y = (ds["tac_noisy_sampled"] / ds["dt"][None, None, :])[idx]
prior = {k: pr[k] for k in pr.files}
out = run_sharded(y, ds["vartacref"][idx], ds["vark2p"][idx], ds["sigma_noise"], ds["time_vector"], ds["dt"], prior,
                  draws=200, tune=300, n_chains=8, seed=77, max_draws=200, device=local)
assert out.shape == (S, 96, 8)
if rank == 0:
    with MHSampler(n_chains=8, max_tacs=S, max_draws=200, seed=77, device=local) as s:
        s.set_frames(ds["time_vector"], ds["dt"]); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
        s.set_data(y, ds["vartacref"][idx], ds["vark2p"][idx], ds["sigma_noise"])
        s.run(draws=200, tune=300)
        ref = s.summary()
    got = out.cpu().numpy()
    same = np.array_equal(np.nan_to_num(got, nan=-1.0), np.nan_to_num(ref, nan=-1.0))
    print("world %d: gathered summaries %s the single-GPU run (max |diff| %.3g)" % (
        world, "EQUAL" if same else "DIFFER FROM", np.nanmax(np.abs(got - ref))))
    assert same
dist.destroy_process_group()
