#!/usr/bin/env python
"""BASELINE configs[3] end to end on the GPUs of one box (run under torchrun): noise sweep sigma in {0.05, 0.1, 0.2},
synthetic SRTM2 TACs from the restated sample_sim_data.py priors (GPU generator K4, same seed on every rank -> same data),
1024 chains per TAC at the reference's length (40 000 tune + 20 000 draws, thinned to 1 000 stored draws per chain).
3 TACs < 8 ranks: the chains of every TAC are split over the ranks (distributed.run_sharded, chain-sharded), the thinned
draws gathered to the TAC's owner for the rank-normalised summary.  Rank 0 reports seconds, chain-steps/s, R-hat / ESS and
how many posterior SDs the truth lies from the posterior mean."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
from pet_posterior_distribution_b200 import sample_sim_data as gen
from pet_posterior_distribution_b200.distributed import run_sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S, C, TUNE, DRAWS, THIN = 3, 1024, 40000, 20000, 20
prior = gen.load_prior()
t, dtv = gen.frame_grid()
for sigma in (0.05, 0.1, 0.2):
    sig = gen.noise_table(np.random.default_rng(int(sigma * 1000)), sigma, t, dtv)
    with MHSampler(n_chains=1, max_tacs=S, seed=0, device=local) as g:
        g.set_frames(t, dtv); g.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
        g.synth(S, 777, prior["mu_tac_ref"], prior["Cov_tac_ref"], float(prior["mu_k2p"]), sig)
        d = g.synth_get()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    summ = run_sharded(d["y"].astype(np.float64), d["tac_ref"], np.full(S, float(prior["mu_k2p"])), sig, t, dtv, prior,
                       draws=DRAWS, tune=TUNE, n_chains=C, thin=THIN, seed=31, max_draws=DRAWS // THIN, device=local)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        sm = summ.cpu().numpy()
        truth = np.concatenate([d["DVR"], d["R1"]], axis=1)
        z = (sm[:, :, 0] - truth) / sm[:, :, 1]
        steps = S * C * 96 * (TUNE + DRAWS)
        print("configs[3] sigma %.2f: %d TACs x %d chains x (%d tune + %d draws, thin %d) on %d GPUs (chain-sharded): %.2f s "
              "(%.3e chain-steps/s); r_hat max %.4f, ess_bulk min %.0f of %d stored draws per TAC; truth within %.2f posterior SDs (max), rms %.2f"
              % (sigma, S, C, TUNE, DRAWS, THIN, world, dt, steps / dt, np.nanmax(sm[:, :, 5]), np.nanmin(sm[:, :, 3]), C * DRAWS // THIN,
                 np.abs(z).max(), np.sqrt((z ** 2).mean())), flush=True)
dist.destroy_process_group()
