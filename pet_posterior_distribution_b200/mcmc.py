"""Drop-in for the reference's ``mcmc.py`` (the MH-MCMC baseline), B200-native.

Same module-level configuration names (mcmc.py:53-59), the same input discovery
(glob of ``sim_data/nROI48/*_test/data_nROI48_n100_s1.0e-01.pik``, latest; mcmc.py:62-71),
the same per-sample skip-if-exists rule (mcmc.py:116-128) and the same three outputs per
sample (mcmc.py:162-194): ``MCMC_s*/MH_MCMC_nROI48_it*_brn*_km_obs-*.pik`` with keys
idata / DVR_mcmc / k2p_mcmc / R1_mcmc / iter / burn / y_obs / km_obs / chains / elapsed_time,
its ``_summary.csv`` and ``rhat_less_than_102.txt``.

Differences by design: ``chains`` is honoured (the reference never passes it to pm.sample),
all pending test samples run as ONE batch on the GPU instead of a Python loop, and the
module does nothing at import time -- run ``python -m pet_posterior_distribution_b200.mcmc``
(from the directory that holds ``sim_data/`` and ``prior_stats_nROI48.pik``) or call main().
Under ``torchrun --nproc-per-node N -m pet_posterior_distribution_b200.mcmc`` the pending samples are sharded over
N GPUs (by sample when there are at least N of them, else by chain: distributed.run_sharded) and every rank
writes the files of the samples it owns.
"""
import glob
import os
import pickle
import time

import numpy as np

from . import diagnostics
from .kinetic_model import SRTM2
from .sampler import MHSampler

NP_DTYPE = np.float64
FLAG_PLOT = False            # mcmc.py:23; True writes the figures of mcmc.py:198-258 (plots.py; needs matplotlib)

# ---- configuration: same names and defaults as mcmc.py:45-59 ---------------------------------
CUR_DIR = './'
n_ROI_test = 48
n_samples_test = 100
mean_sigma_noise_load = 1e-1
iter_mcmc = 200
burn_mcmc = 400
chains = 4
sample_range = range(0, 10)  # mcmc.py:104 `for sample_plot in range(0, 10)`
seed = 0
thin = 1


class CreateTAC_SRTM2:
    """mcmc.py:27-39 without PyTensor: same constructor, ``perform(node, inputs, outputs)`` writes
    ``outputs[0][0] = create_activity_curve(DVR, R1, k2p).T`` ((48,54) float64), and it is callable."""
    __props__ = ()

    def __init__(self, k_srtm):
        self.k_srtm = k_srtm

    def perform(self, node, inputs, outputs, **kwargs):
        outputs[0][0] = self.k_srtm.create_activity_curve(DVR=inputs[0], R1=inputs[1], k2p=inputs[2]).T

    def __call__(self, DVR, R1, k2p):
        out = [[None]]
        self.perform(None, [DVR, R1, k2p], out)
        return out[0][0]


def find_test_file(data_dir=None):
    """mcmc.py:62-69: latest ``*_test`` directory holding the data pickle."""
    data_dir = data_dir or os.path.join(CUR_DIR, 'sim_data')
    str_noise = '_s{:.1e}'.format(mean_sigma_noise_load)
    pattern = os.path.join(data_dir, 'nROI{}'.format(n_ROI_test), '*_test',
                           'data_nROI{}_n{}{}.pik'.format(n_ROI_test, n_samples_test, str_noise))
    hits = sorted(glob.glob(pattern))
    if not hits:
        raise IndexError("no test data found: " + pattern)       # the reference raises IndexError here too
    return os.path.dirname(hits[-1]), os.path.basename(hits[-1])


def save_name(km_obs):
    """mcmc.py:119-123."""
    return 'MH_MCMC_nROI{}_it{:.1e}_brn{:.1e}_km_obs-{:.3f}-{:.3f}-{:.3f}.pik'.format(
        n_ROI_test, iter_mcmc, burn_mcmc, km_obs['DVR'][0], km_obs['R1'][0], km_obs['k2p'][0])


def _dist_context():
    """(rank, world, local_rank, initialised_here).  Under torchrun (WORLD_SIZE > 1) the pending samples are sharded over
    the ranks, one process per GPU (NCCL); otherwise a single process on one GPU."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return 0, 1, None, False
    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    here = False
    if not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        here = True
    return dist.get_rank(), dist.get_world_size(), local, here


def _write_sample(mcmc_roi_dir, fname, sample_plot, km_obs, dvr, r1, y_obs, summ, ext, elapsed, prior_mean=None):
    """The three per-sample outputs of mcmc.py:162-194 (+ the figures of :198-258 when FLAG_PLOT)."""
    DVR_mcmc = dvr.astype(NP_DTYPE)
    R1_mcmc = r1.astype(NP_DTYPE)
    k2p_mcmc = np.full(DVR_mcmc.shape[:2], km_obs['k2p'][0])
    save_mcmc_dic = {
        'idata': diagnostics.make_idata(DVR_mcmc, R1_mcmc, km_obs['k2p'], {'scaling': summ[:, 7], 'accept': summ[:, 6]}),
        'DVR_mcmc': DVR_mcmc, 'k2p_mcmc': k2p_mcmc, 'R1_mcmc': R1_mcmc,
        'iter': iter_mcmc, 'burn': burn_mcmc, 'y_obs': y_obs, 'km_obs': km_obs,
        'chains': chains, 'elapsed_time': elapsed,
    }
    pickle.dump(save_mcmc_dic, open(fname, 'wb'))
    with open(fname.replace('.pik', '_summary.csv'), 'w') as f:
        f.write(diagnostics.summary_csv(DVR_mcmc, R1_mcmc, km_obs['k2p'], summ, ext))
    diagnostics.append_rhat_log(mcmc_roi_dir, os.path.basename(fname), sample_plot, summ)
    if FLAG_PLOT:
        from . import plots
        plots.plot_sample(fname, {'DVR': DVR_mcmc, 'R1': R1_mcmc}, km_obs, prior_mean)


def main(data_dir=None, prior_path=None, device=0):
    load_km_dir, load_km_fname = find_test_file(data_dir)
    load_test_dict = pickle.load(open(os.path.join(load_km_dir, load_km_fname), 'rb'))
    time_vector = np.array(load_test_dict['time_vector'], dtype=NP_DTYPE)
    dt = np.array(load_test_dict['dt'], dtype=NP_DTYPE)
    DVR_load = np.array(load_test_dict['varDVR'], dtype=NP_DTYPE)
    R1_load = np.array(load_test_dict['varR1'], dtype=NP_DTYPE)
    k2p_load = np.array(load_test_dict['vark2p'], ndmin=2, dtype=NP_DTYPE).T
    tac_load = np.array(load_test_dict['tac_noisy_sampled'], dtype=NP_DTYPE) / dt[None, None, :]   # mcmc.py:79-80
    prior_path = prior_path or os.path.join(CUR_DIR, 'prior_stats_nROI{}.pik'.format(n_ROI_test))
    stats_dict = pickle.load(open(prior_path, 'rb'))
    sigma_noise = np.array(load_test_dict['sigma_noise'], dtype=NP_DTYPE)
    tacref_load = np.array(load_test_dict['vartacref'], dtype=NP_DTYPE)
    str_noise = '_s{:.1e}'.format(mean_sigma_noise_load)
    mcmc_roi_dir = os.path.join(load_km_dir, 'MCMC{}'.format(str_noise))
    rank, world, local, dist_here = _dist_context()
    if rank == 0:
        os.makedirs(mcmc_roi_dir, exist_ok=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()                                               # the skip rule below must see one state of the directory

    pending = []
    for sample_plot in sample_range:
        km_obs = {'DVR': DVR_load[sample_plot], 'R1': R1_load[sample_plot], 'k2p': k2p_load[sample_plot]}
        fname = os.path.join(mcmc_roi_dir, save_name(km_obs))
        if os.path.isfile(fname):                                   # mcmc.py:125-128
            if rank == 0:
                print('MCMC File already exists with these parameters (sample {})... Skipping.'.format(sample_plot))
            continue
        pending.append((sample_plot, km_obs, fname))
    if world > 1:
        dist.barrier()                                               # nobody writes before everybody has listed
    if not pending:
        if dist_here:
            dist.destroy_process_group()
        return []

    # A chain's Philox stream is keyed by its sample index (mcmc.py:104's loop variable) and its chain index: results do
    # not depend on which other samples are pending, on batching, or on the number of GPUs.
    idx = [p[0] for p in pending]
    n_store = (iter_mcmc + thin - 1) // thin
    tic = time.time()
    written = []
    if world == 1:
        with MHSampler(n_chains=chains, max_tacs=len(idx), max_draws=n_store, seed=seed, device=device) as s:
            s.set_frames(time_vector, dt)
            s.set_prior(stats_dict['mu_DVR'], stats_dict['Cov_DVR'], stats_dict['mu_R1'], stats_dict['Cov_R1'])
            s.set_data(tac_load[idx], tacref_load[idx], k2p_load[idx, 0], sigma_noise)
            s.set_global_ids(np.asarray(idx, np.uint64))
            s.run(draws=iter_mcmc, tune=burn_mcmc, thin=thin)       # pm.sample(draws, tune, step=Metropolis)
            dvr, r1 = s.chains()
            summ = s.summary()
            ext = s.summary_ext() if s.n_stored >= 8 else None
            kernel_ms, launches = s.last_kernel_ms()
        mine = {j: dict(dvr=dvr[j], r1=r1[j], ext=None if ext is None else ext[j]) for j in range(len(idx))}
    else:
        from .distributed import run_sharded
        summ_t, mine = run_sharded(tac_load[idx], tacref_load[idx], k2p_load[idx, 0], sigma_noise, time_vector, dt, stats_dict,
                                   draws=iter_mcmc, tune=burn_mcmc, n_chains=chains, thin=thin, seed=seed, max_draws=n_store,
                                   device=local, tac_ids=idx, keep_chains=True)
        summ = summ_t.cpu().numpy()
        kernel_ms = mine.pop('_kernel_ms', 0.0)
    elapsed_time = time.time() - tic
    print('elapsed time: {:.1f} sec ({} samples, {} chains, {} GPU(s); sweep kernels of rank {} {:.1f} ms)'.format(
        elapsed_time, len(idx), chains, world, rank, kernel_ms))
    for j, (sample_plot, km_obs, fname) in enumerate(pending):
        if j not in mine:
            continue                                                 # another rank owns this sample
        m = mine[j]
        _write_sample(mcmc_roi_dir, fname, sample_plot, km_obs, m['dvr'], m['r1'], tac_load[sample_plot].reshape([n_ROI_test, -1]),
                      summ[j], m['ext'], elapsed_time / len(idx), {'DVR': stats_dict['mu_DVR'], 'R1': stats_dict['mu_R1']})
        written.append(fname)
    if world > 1:
        dist.barrier()
        if dist_here:
            dist.destroy_process_group()
    return written


def _cli(argv=None):
    """The reference is configured by editing its module constants (mcmc.py:53-59); the same constants can be given
    on the command line here (an extension: the reference has no CLI)."""
    import argparse
    g = globals()
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--chains", type=int, default=chains)
    ap.add_argument("--iter", type=int, default=iter_mcmc, dest="iter_mcmc")
    ap.add_argument("--burn", type=int, default=burn_mcmc, dest="burn_mcmc")
    ap.add_argument("--thin", type=int, default=thin)
    ap.add_argument("--seed", type=int, default=seed)
    ap.add_argument("--samples", type=int, nargs=2, default=[sample_range.start, sample_range.stop], metavar=("FIRST", "STOP"))
    ap.add_argument("--data-dir", default=None)
    ap.add_argument("--prior", default=None)
    a = ap.parse_args(argv)
    g.update(chains=a.chains, iter_mcmc=a.iter_mcmc, burn_mcmc=a.burn_mcmc, thin=a.thin, seed=a.seed,
             sample_range=range(a.samples[0], a.samples[1]))
    return main(data_dir=a.data_dir, prior_path=a.prior)


if __name__ == "__main__":
    _cli()
