#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU): posterior summaries computed over N GPUs, TAC-sharded
(S >= N) or chain-sharded (S < N: the chains of a TAC split over ranks, draws / moments gathered to the TAC's owner),
must equal the single-GPU result bit for bit -- Philox streams are keyed by the GLOBAL TAC and chain index, so results
do not depend on the number of GPUs.  Cases: 7 TACs x 8 chains (ragged TAC shards); BASELINE configs[1]-like 1 TAC x 64
chains; configs[3]-like 3 TACs x 1024 chains (thinned); 1 TAC x 16 chains in moments mode."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
from pet_posterior_distribution_b200.distributed import run_sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
pr = np.load(os.path.join(g, "prior_stats_nROI48.npz")); ds = np.load(os.path.join(g, "dataset_s0.1.npz"))
prior = {k: pr[k] for k in pr.files}
CASES = [  # S, chains, draws, tune, thin, max_draws
    (7, 8, 200, 300, 1, 200),
    (1, 64, 400, 600, 1, 400),
    (3, 1024, 600, 600, 6, 100),
    (1, 16, 320, 300, 1, 0),
]
for S, C, draws, tune, thin, md in CASES:
    idx = np.arange(S) % 4
    y = (ds["tac_noisy_sampled"] / ds["dt"][None, None, :])[idx]
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    out = run_sharded(y, ds["vartacref"][idx], ds["vark2p"][idx], ds["sigma_noise"], ds["time_vector"], ds["dt"], prior,
                      draws=draws, tune=tune, n_chains=C, thin=thin, seed=77, max_draws=md, device=local)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    assert out.shape == (S, 96, 8)
    if rank == 0:
        with MHSampler(n_chains=C, max_tacs=S, max_draws=md, seed=77, device=local) as s:
            s.set_frames(ds["time_vector"], ds["dt"]); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
            s.set_data(y, ds["vartacref"][idx], ds["vark2p"][idx], ds["sigma_noise"])
            t1 = time.perf_counter(); s.run(draws=draws, tune=tune, thin=thin); ref = s.summary(); d1 = time.perf_counter() - t1
        got = out.cpu().numpy()
        same = np.array_equal(np.nan_to_num(got, nan=-1.0), np.nan_to_num(ref, nan=-1.0))
        print("world %d, %d TAC(s) x %d chains (%s, %s): %.3f s on %d GPUs vs %.3f s on one; gathered summaries %s the single-GPU run "
              "(max |diff| %.3g)" % (world, S, C, "TAC-sharded" if S >= world else "chain-sharded", "stored draws" if md else "moments",
                                    dt, world, d1, "EQUAL" if same else "DIFFER FROM", np.nanmax(np.abs(got - ref))), flush=True)
        assert same
dist.destroy_process_group()
