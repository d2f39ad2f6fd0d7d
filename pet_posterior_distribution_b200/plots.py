"""The reference's optional figures (mcmc.py:198-258, FLAG_PLOT): per parameter (DVR, R1) a histogram of the pooled
posterior draws of one ROI with the matching normal density and the posterior mean marked, and a grid of the
histograms of all 48 ROIs.  File names follow the reference: ``<pickle>_{DVR,R1}_ROI<k>.png`` and
``<pickle>_{DVR,R1}_ROI_all.png``.  matplotlib is imported lazily (it is not needed for sampling)."""
import numpy as np


def _normal_pdf(x, m, s):
    return np.exp(-0.5 * ((x - m) / s) ** 2) / (s * np.sqrt(2 * np.pi))


def plot_sample(save_dir_filename, chains, km_obs, prior_mean, roi_plot=0, bins=100):
    """chains: {'DVR': (chains, draws, 48), 'R1': ...}; km_obs: the sample's true parameters; prior_mean: {'DVR': mu_DVR,
    'R1': mu_R1}.  Returns the list of files written."""
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    written = []
    for name in ("DVR", "R1"):
        draws = np.asarray(chains[name])
        n_roi = draws.shape[-1]
        one = draws[:, :, roi_plot].ravel()
        m, s = one.mean(), one.std()
        fig, ax = plt.subplots(1, 1, figsize=(12, 5))
        ax.hist(one, bins=bins, color="red", density=True, label="MCMC " + name)
        lo, hi = ax.get_xlim()
        xs = np.arange(lo, hi, 1e-4)
        ax.plot(xs, _normal_pdf(xs, m, s), linewidth=3.0)
        ax.axvline(x=m, color="black", linestyle="--", label="mean")
        ax.set_title("prior: $\\mu = {:.4f}$ ({}_observed = {:.4f}) \nMCMC: $\\mu = {:.4f}$ - $\\sigma = {:.4f}$".format(
            prior_mean[name][roi_plot], name, km_obs[name][roi_plot], m, s))
        out = save_dir_filename.replace(".pik", "_{}_ROI{:d}.png".format(name, roi_plot))
        fig.savefig(out)
        plt.close(fig)
        written.append(out)
        side = int(np.ceil(np.sqrt(n_roi)))
        fig_all, axes = plt.subplots(side, side, figsize=(16, 10))
        axes = np.atleast_1d(axes).ravel()
        for k in range(n_roi):
            v = draws[:, :, k].ravel()
            axes[k].hist(v, density=True, label=str(k), bins=bins)
            xs = np.linspace(*axes[k].get_xlim(), num=500)
            axes[k].plot(xs, _normal_pdf(xs, v.mean(), v.std()), linewidth=3.0, color="black")
            axes[k].set_title("({}_obs = {:.2f}) $\\mu = {:.2f}$ - $\\sigma = {:.3f}$".format(name, km_obs[name][k], v.mean(), v.std()))
            axes[k].legend()
        out = save_dir_filename.replace(".pik", "_{}_ROI_all.png".format(name))
        fig_all.savefig(out)
        plt.close(fig_all)
        written.append(out)
    return written
