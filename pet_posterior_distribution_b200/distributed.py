"""Multi-GPU: one process per GPU (torch.distributed, NCCL over NVLink).

The unit of independent work is (TAC, chain) (ROIs of a chain are coupled by the prior):

* S >= world: TACs are cut into contiguous blocks, one per rank; every chain of a TAC is co-resident, R-hat / ESS
  are local and sampling needs NO collective.  The only exchange is one all-gather of the per-(TAC, coordinate)
  summaries at the end; each rank's K3 kernel writes its rows straight into its slice of the gather buffer
  (petmh_summary_device), so there is no staging copy.
* S < world (BASELINE configs[1]: one TAC x 64 chains; configs[3]: a few TACs x 1024 chains): the CHAINS of a TAC are
  split over the ranks assigned to it.  Sampling is still collective-free; afterwards the (thinned) draws -- or, in
  moments mode, the per-chain running moments -- of a TAC are all-gathered and its owner rank (the first of its group)
  runs the rank-normalised / batch-means summary over all of them (petmh_summary_from_draws_device /
  petmh_summary_from_moments_device), then the summaries are all-gathered as above.

Philox streams are keyed by the GLOBAL TAC index and the GLOBAL chain index (petmh_set_global_ids), hence every
draw, and every summary, is independent of the number of GPUs.
"""
import ctypes as C

import numpy as np


def shard_bounds(n_items, world, rank):
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most one."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_items, world):
    return [shard_bounds(n_items, world, r)[1] - shard_bounds(n_items, world, r)[0] for r in range(world)]


def chain_shards(n_tacs, n_chains, world):
    """Work assignment when there are fewer TACs than ranks: rank r works on TAC r % S; the ranks of a TAC (its
    "group", ascending) split its chains in contiguous blocks.  Returns one dict per rank:
    tac, c_lo, c_hi (may be empty when a group has more ranks than chains), group (ranks of the TAC), owner."""
    S, W = int(n_tacs), int(world)
    assert 0 < S < W
    out = []
    for r in range(W):
        t = r % S
        group = list(range(t, W, S))
        lo, hi = shard_bounds(n_chains, len(group), group.index(r))
        out.append(dict(tac=t, c_lo=lo, c_hi=hi, group=group, owner=group[0]))
    return out


def gather_summaries(local, n_total, group=None):
    """all-gather the (S_local, 96, K) float32 summaries of every rank into (n_total, 96, K).
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


def gather_padded(local, counts, group=None):
    """all-gather tensors whose first dimension differs per rank (counts[r] rows on rank r): returns the list of
    every rank's rows.  One collective on a buffer padded to max(counts)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    cmax = max(max(counts), 1)
    pad = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out.view((world, cmax) + tuple(local.shape[1:]))
    return [out[r, : counts[r]] for r in range(world)]


class _DevArray:
    """Zero-copy view of device memory owned by a libpetmh handle (CUDA array interface v2)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _summary_inputs(s, device):
    """torch views of the handle's summary inputs + the counters (see petmh_export_summary_inputs)."""
    import torch
    from . import _lib
    p = [C.c_void_p() for _ in range(5)]
    cnt = (C.c_int * 8)()
    s._ck(_lib.lib.petmh_export_summary_inputs(s._h, *[C.byref(x) for x in p], cnt))
    nc = s.n_tac * s.n_chains
    dev = "cuda:%d" % device
    wrap = lambda ptr, shape, ts: torch.as_tensor(_DevArray(ptr, shape, ts), device=dev)
    out = dict(counters=list(cnt),
               mom=wrap(p[1].value, (nc, 2, 96, 5), "<f4"), nacc=wrap(p[2].value, (nc, 96), "<i4"),
               scale=wrap(p[3].value, (nc, 96), "<f4"), mu=wrap(p[4].value, (96,), "<f8"))
    if p[0].value and cnt[1] > 0:
        out["draws"] = wrap(p[0].value, (nc, cnt[1], 96), "<f4")[:, : cnt[0]]
    return out


def _run_tac_sharded(data, prior, draws, tune, n_chains, thin, seed, max_draws, device, tac_ids, keep):
    import torch
    import torch.distributed as dist
    from .sampler import MHSampler
    rank, world = dist.get_rank(), dist.get_world_size()
    S = data["y_obs"].shape[0]
    lo, hi = shard_bounds(S, world, rank)
    sizes = shard_sizes(S, world)
    smax = max(sizes)
    gather = torch.zeros((world, smax, 96, 8), dtype=torch.float32, device="cuda:%d" % device)
    local = {}
    if hi > lo:
        with MHSampler(n_chains=n_chains, max_tacs=hi - lo, max_draws=max_draws, seed=seed, device=device) as s:
            s.set_frames(data["time_vector"], data["dt"])
            s.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
            s.set_data(data["y_obs"][lo:hi], data["tac_ref"][lo:hi], np.asarray(data["k2p"])[lo:hi], data["sigma_noise"])
            s.set_global_ids(np.asarray(tac_ids, np.uint64)[lo:hi])
            s.run(draws=draws, tune=tune, thin=thin)
            # K3 writes this rank's rows directly into its slot of the all-gather buffer
            s.summary_into(gather[rank].data_ptr(), torch.cuda.current_stream().cuda_stream)
            if keep and max_draws > 0:
                dvr, r1 = s.chains()
                ext = s.summary_ext() if s.n_stored >= 8 else None
                for j in range(hi - lo):
                    local[lo + j] = dict(dvr=dvr[j], r1=r1[j], ext=None if ext is None else ext[j])
            local["_kernel_ms"] = s.last_kernel_ms()[0]
    torch.cuda.synchronize()
    dist.all_gather_into_tensor(gather.view(world * smax, 96, 8), gather[rank].clone())
    summ = torch.cat([gather[r, : sizes[r]] for r in range(world)], dim=0)
    return summ, local


def _run_chain_sharded(data, prior, draws, tune, n_chains, thin, seed, max_draws, device, tac_ids, keep):
    import torch
    import torch.distributed as dist
    from . import _lib
    from .sampler import MHSampler
    rank, world = dist.get_rank(), dist.get_world_size()
    S = data["y_obs"].shape[0]
    plan = chain_shards(S, n_chains, world)
    me = plan[rank]
    t, c_lo, c_hi = me["tac"], me["c_lo"], me["c_hi"]
    n_loc = c_hi - c_lo
    dev = "cuda:%d" % device
    counts = [p["c_hi"] - p["c_lo"] for p in plan]
    local = {}
    n_store = min(max_draws, (draws + thin - 1) // thin) if max_draws > 0 else 0
    z = lambda *shape, dtype=torch.float32: torch.zeros(shape, dtype=dtype, device=dev)
    d_loc, mom_loc, nacc_loc, sc_loc, counters, mu = z(0, n_store, 96), z(0, 2, 96, 5), z(0, 96, dtype=torch.int32), z(0, 96), None, None
    s = None
    if n_loc > 0:
        s = MHSampler(n_chains=n_loc, max_tacs=1, max_draws=max_draws, seed=seed, device=device)
        s.set_frames(data["time_vector"], data["dt"])
        s.set_prior(prior["mu_DVR"], prior["Cov_DVR"], prior["mu_R1"], prior["Cov_R1"])
        s.set_data(data["y_obs"][t:t + 1], data["tac_ref"][t:t + 1], np.asarray(data["k2p"])[t:t + 1], data["sigma_noise"])
        s.set_global_ids(np.asarray(tac_ids, np.uint64)[t:t + 1], chain_gid0=c_lo, chains_per_tac_global=n_chains)
        s.run(draws=draws, tune=tune, thin=thin)
        v = _summary_inputs(s, device)
        counters, mu = v["counters"], v["mu"]
        mom_loc, nacc_loc, sc_loc = v["mom"], v["nacc"], v["scale"]
        if "draws" in v:
            d_loc = v["draws"].contiguous()
        local["_kernel_ms"] = s.last_kernel_ms()[0]
    # ---- one all-gather per array: the state the owner needs for the whole-TAC summary ----
    use_draws = n_store >= 8
    g_nacc = gather_padded(nacc_loc, counts)
    g_sc = gather_padded(sc_loc, counts)
    g_main = gather_padded(d_loc if use_draws else mom_loc, counts)
    ctr = torch.tensor(counters if counters is not None else [0] * 8, dtype=torch.int32, device=dev)
    dist.all_reduce(ctr, op=dist.ReduceOp.MAX)                      # identical on every rank that sampled
    ctr = [int(x) for x in ctr.cpu()]
    rows = torch.zeros((S, 96, 8), dtype=torch.float32, device=dev)
    ext = torch.zeros((1, 96, 4), dtype=torch.float32, device=dev)
    if rank == me["owner"]:
        grp = me["group"]
        cat = lambda parts: torch.cat([parts[r] for r in grp], dim=0).contiguous()
        nacc_all, sc_all, main_all = cat(g_nacc), cat(g_sc), cat(g_main)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        vp = lambda x: C.c_void_p(x.data_ptr())
        if use_draws:
            rc = _lib.lib.petmh_summary_from_draws_device(device, vp(main_all), 1, n_chains, ctr[0], vp(nacc_all), vp(sc_all), ctr[7],
                                                          vp(rows[t]), vp(ext), st)
        else:
            if mu is None:
                mu = torch.tensor(np.concatenate([prior["mu_DVR"], prior["mu_R1"]]), dtype=torch.float64, device=dev)
            nh, nb = (C.c_int * 2)(ctr[2], ctr[3]), (C.c_int * 2)(ctr[4], ctr[5])
            rc = _lib.lib.petmh_summary_from_moments_device(device, vp(main_all), vp(mu), 1, n_chains, nh, nb, ctr[6], vp(nacc_all),
                                                            vp(sc_all), ctr[7], vp(rows[t]), st)
        if rc:
            raise _lib.PetmhError(rc, (_lib.lib.petmh_last_error(None) or b"").decode())
        torch.cuda.synchronize()
        if keep and use_draws:
            d = main_all.cpu().numpy()
            local[t] = dict(dvr=d[..., :48].copy(), r1=d[..., 48:].copy(), ext=ext[0].cpu().numpy())
    if s is not None:
        s.close()
    dist.all_reduce(rows)                                            # every TAC's rows come from exactly one owner
    return rows, local


def run_sharded(y_obs, tac_ref, k2p, sigma_noise, time_vector, dt, prior, draws, tune, n_chains=4, thin=1,
                seed=0, max_draws=0, device=None, tac_ids=None, keep_chains=False):
    """Posterior summaries (S, 96, 8) for ALL TACs on every rank.  TAC-sharded when S >= world, chain-sharded
    otherwise.  Requires an initialised NCCL process group (one rank per GPU).  tac_ids: global ids of the TACs
    (default 0..S-1) -- they key the Philox streams.  keep_chains: also return {tac: dict(dvr, r1, ext)} for the
    TACs this rank owns (needs max_draws > 0), e.g. to write the reference's per-sample files."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    device = torch.cuda.current_device() if device is None else device
    S = y_obs.shape[0]
    data = dict(y_obs=y_obs, tac_ref=tac_ref, k2p=k2p, sigma_noise=sigma_noise, time_vector=time_vector, dt=dt)
    tac_ids = np.arange(S) if tac_ids is None else np.asarray(tac_ids)
    fn = _run_tac_sharded if S >= world else _run_chain_sharded
    summ, local = fn(data, prior, draws, tune, n_chains, thin, seed, max_draws, device, tac_ids, keep_chains)
    return (summ, local) if keep_chains else summ
