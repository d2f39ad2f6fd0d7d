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
FLAG_PLOT = False            # plots (mcmc.py:198-258) are out of scope

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
    str_noise = '_s{:.1e}'.format(mean_sigma_noise_load)
    mcmc_roi_dir = os.path.join(load_km_dir, 'MCMC{}'.format(str_noise))
    os.makedirs(mcmc_roi_dir, exist_ok=True)

    pending = []
    for sample_plot in sample_range:
        km_obs = {'DVR': DVR_load[sample_plot], 'R1': R1_load[sample_plot], 'k2p': k2p_load[sample_plot]}
        fname = os.path.join(mcmc_roi_dir, save_name(km_obs))
        if os.path.isfile(fname):                                   # mcmc.py:125-128
            print('MCMC File already exists with these parameters (sample {})... Skipping.'.format(sample_plot))
            continue
        pending.append((sample_plot, km_obs, fname))
    if not pending:
        return []

    idx = [p[0] for p in pending]
    n_store = (iter_mcmc + thin - 1) // thin
    tic = time.time()
    with MHSampler(n_chains=chains, max_tacs=len(idx), max_draws=n_store, seed=seed, device=device,
                   tac_gid0=min(idx)) as s:
        s.set_frames(time_vector, dt)
        s.set_prior(stats_dict['mu_DVR'], stats_dict['Cov_DVR'], stats_dict['mu_R1'], stats_dict['Cov_R1'])
        s.set_data(tac_load[idx], np.array(load_test_dict['vartacref'], dtype=NP_DTYPE)[idx], k2p_load[idx, 0], sigma_noise)
        s.run(draws=iter_mcmc, tune=burn_mcmc, thin=thin)           # pm.sample(draws, tune, step=Metropolis)
        dvr, r1 = s.chains()
        summ = s.summary()
        kernel_ms, launches = s.last_kernel_ms()
    elapsed_time = time.time() - tic
    print('elapsed time: {:.1f} sec ({} samples, {} chains; sweep kernels {:.1f} ms)'.format(
        elapsed_time, len(idx), chains, kernel_ms))

    written = []
    for j, (sample_plot, km_obs, fname) in enumerate(pending):
        DVR_mcmc = dvr[j].astype(NP_DTYPE)
        R1_mcmc = r1[j].astype(NP_DTYPE)
        k2p_mcmc = np.full(DVR_mcmc.shape[:2], km_obs['k2p'][0])
        y_obs = tac_load[sample_plot].reshape([n_ROI_test, -1])
        save_mcmc_dic = {
            'idata': diagnostics.make_idata(DVR_mcmc, R1_mcmc, km_obs['k2p'],
                                            {'scaling': summ[j][:, 7], 'accept': summ[j][:, 6]}),
            'DVR_mcmc': DVR_mcmc, 'k2p_mcmc': k2p_mcmc, 'R1_mcmc': R1_mcmc,
            'iter': iter_mcmc, 'burn': burn_mcmc, 'y_obs': y_obs, 'km_obs': km_obs,
            'chains': chains, 'elapsed_time': elapsed_time / len(idx),
        }
        pickle.dump(save_mcmc_dic, open(fname, 'wb'))
        with open(fname.replace('.pik', '_summary.csv'), 'w') as f:
            f.write(diagnostics.summary_csv(DVR_mcmc, R1_mcmc, km_obs['k2p'], summ[j]))
        diagnostics.append_rhat_log(mcmc_roi_dir, os.path.basename(fname), sample_plot, summ[j])
        written.append(fname)
    return written


if __name__ == "__main__":
    main()
