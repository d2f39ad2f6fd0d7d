"""Multi-GPU: one process per GPU (torch.distributed, NCCL over NVLink), TAC-sharded.

The unit of independent work is (TAC, chain) (ROIs of a chain are coupled by the prior), so
TACs are cut into contiguous blocks, one per rank; sampling needs NO collective.  The only
exchange is one all-gather of the per-(TAC, coordinate) summaries at the end; each rank's
K3 kernel writes its rows straight into its slice of the gather buffer
(petmh_summary_device), so there is no staging copy.  Philox keys use the GLOBAL TAC index,
hence results do not depend on the number of GPUs.
"""
import numpy as np


def shard_bounds(n_items, world, rank):
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most one."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_items, world):
    return [shard_bounds(n_items, world, r)[1] - shard_bounds(n_items, world, r)[0] for r in range(world)]


def gather_summaries(local, n_total, group=None):
    """all-gather the (S_local, 96, 8) float32 summaries of every rank into (n_total, 96, 8).
    `local` is a torch tensor (CUDA for NCCL, CPU for gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_total, world)
    smax = max(sizes)
    pad = torch.zeros((smax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world, smax) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1, *local.shape[1:]), pad, group=group)
    return torch.cat([out[r, : sizes[r]] for r in range(world)], dim=0)


def run_sharded(y_obs, tac_ref, k2p, sigma_noise, time_vector, dt, prior, draws, tune, n_chains=4, thin=1,
                seed=0, max_draws=0, device=None):
    """Posterior summaries for ALL TACs on every rank: shard by TAC, sample locally, all-gather.
    Requires an initialised NCCL process group (one rank per GPU)."""
    import torch
    import torch.distributed as dist
    from .sampler import MHSampler
    rank, world = dist.get_rank(), dist.get_world_size()
    device = torch.cuda.current_device() if device is None else device
    S = y_obs.shape[0]
    lo, hi = shard_bounds(S, world, rank)
    sizes = shard_sizes(S, world)
    smax = max(sizes)
    gather = torch.zeros((world, smax, 96, 8), dtype=torch.float32, device="cuda:%d" % device)
    if hi > lo:
        with MHSampler(n_chains=n_chains, max_tacs=hi - lo, max_draws=max_draws, seed=seed, device=device,
                       tac_gid0=lo) as s:
            s.set_frames(time_vector, dt)
            s.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
            s.set_data(y_obs[lo:hi], tac_ref[lo:hi], np.asarray(k2p)[lo:hi], sigma_noise)
            s.run(draws=draws, tune=tune, thin=thin)
            # K3 writes this rank's rows directly into its slot of the all-gather buffer
            s.summary_into(gather[rank].data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    dist.all_gather_into_tensor(gather.view(world * smax, 96, 8), gather[rank].clone())
    return torch.cat([gather[r, : sizes[r]] for r in range(world)], dim=0)
