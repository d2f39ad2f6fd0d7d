/* petmh.h -- C ABI of libpetmh.so: B200-native batched Metropolis-Hastings posterior
 * sampler for the SRTM2 PET kinetic model.
 *
 * Drop-in boundary for ONE path of yanisdjebra/PET_posterior_distribution: the MCMC
 * baseline `mcmc.py` (element-wise Metropolis, mcmc.py:147-157) over the SRTM2 forward
 * model of `kinetic_model.py` (kinetic_model.py:12-57, 134-158).  Each entry point
 * names the reference interface it replaces.  The reference is pure Python, so the
 * reference-side binding is a ctypes stub.
 *
 * Conventions: every function returns 0 on success, a negative PETMH_E* code on error
 * (message via petmh_last_error).  The caller owns every buffer.  All pointers are HOST
 * pointers, row-major, unless the name says `_device`.  A handle is bound to one CUDA
 * device and one stream and is not thread-safe.  There is NO CPU fallback: create fails
 * with PETMH_ENODEVICE when no sm_100 device is present.
 *
 * Fixed shapes of the path: R = 48 ROIs, T = 54 frames (the reference's acquisition
 * grid, sample_sim_data.py:29-85), state = (DVR[48], R1[48]) per chain.
 */
#ifndef PETMH_H
#define PETMH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PETMH_N_ROI 48
#define PETMH_N_FRAMES 54
#define PETMH_N_COORD 96   /* DVR[48] then R1[48] */
#define PETMH_N_STATS 8    /* columns of the summary table, see petmh_get_summary */

#define PETMH_OK 0
#define PETMH_EINVAL (-1)     /* bad argument / call order */
#define PETMH_ENODEVICE (-2)  /* no sm_100 CUDA device (there is no CPU fallback) */
#define PETMH_ECUDA (-3)      /* CUDA runtime error */
#define PETMH_ENOMEM (-4)
#define PETMH_EGRID (-5)      /* frame grid's operator sparsity differs from the compiled schedule */
#define PETMH_ESYNTH (-6)     /* petmh_synth: some TACs hit a rejection cap (data bound, attempts[] < 0 marks them) */
#define PETMH_ESTATE (-7)     /* petmh_set_checkpoint: blob does not match this handle's configuration */

typedef struct petmh_handle petmh_t;

typedef struct {
    int32_t device;        /* CUDA ordinal */
    int32_t n_chains;      /* chains per TAC (mcmc.py:58 `chains`, honoured here)            */
    int32_t max_tacs;      /* capacity: test TACs resident at once                            */
    int32_t max_draws;     /* stored (thinned) draws per chain; 0 = running moments only      */
    uint64_t seed;         /* Philox key                                                      */
    uint64_t tac_gid0;     /* global index of local TAC 0 (multi-GPU shards keep their streams) */
} petmh_cfg;

/* ---- lifetime -------------------------------------------------------------------- */
int petmh_create(const petmh_cfg* cfg, petmh_t** out);
void petmh_destroy(petmh_t* h);
const char* petmh_last_error(const petmh_t* h); /* h may be NULL: error of a failed create */
int petmh_version(void);

/* ---- model inputs ------------------------------------------------------------------ */
/* replaces SRTM2.__init__(frame_time_list, frame_duration_list, .) kinetic_model.py:136-140
 * (frame END times and durations in minutes; mcmc.py:73-74). */
int petmh_set_frames(petmh_t* h, const double* t54, const double* dt54);
/* replaces pm.MvNormal("var_DVR", mu, cov) / ("var_R1", ...) mcmc.py:148-149 */
int petmh_set_prior(petmh_t* h, const double* mu_dvr48, const double* cov_dvr48x48,
                    const double* mu_r1_48, const double* cov_r1_48x48);
/* replaces the per-sample setup mcmc.py:106-112,133-134,153-155:
 *   y[n_tac][48][54]   = tac_noisy_sampled / dt  (mcmc.py:79-80,109)
 *   tac_ref[n_tac][54] = vartacref[sample]       (mcmc.py:133-134)
 *   k2p[n_tac]         = km_obs['k2p'][0]        (mcmc.py:150)
 *   sigma_noise[48][54] shared by all TACs       (mcmc.py:96,153)
 * Supported scale: |y| and |tac_ref| finite and <= 1e6 (the reference's data are O(1)); PETMH_EINVAL otherwise.  The fp32
 * likelihood takes one logarithm of the product of four frames' model values, so model TACs must stay within about
 * [3e-10, 4e9]: states that leave it evaluate to a non-finite log-likelihood and are rejected, as the reference rejects
 * non-finite log-probabilities -- rescale the activity units if the data themselves are outside. */
int petmh_set_data(petmh_t* h, int n_tac, const double* y, const double* tac_ref,
                   const double* k2p, const double* sigma_noise);
/* Global identity of the local TACs and chains, so that Philox streams -- and therefore every draw -- do not depend on
 * how a job is batched or sharded (replaces nothing in the reference: pm.sample seeds chains by position).  Chain c of
 * local TAC s draws from stream  gid = tac_gid(s) * chains_per_tac_global + chain_gid0 + c,  tac_gid(s) = tac_gids[s]
 * if given (e.g. the sample index of mcmc.py:104's loop when some samples are skipped), else cfg.tac_gid0 + s;
 * chains_per_tac_global = 0 means cfg.n_chains (this handle owns every chain of its TACs).  Splitting the chains of one
 * TAC over ranks: chain_gid0 = first local chain's global index, chains_per_tac_global = the total. */
int petmh_set_global_ids(petmh_t* h, int n_tac, const uint64_t* tac_gids, uint64_t chain_gid0,
                         uint64_t chains_per_tac_global);
/* same, float32 inputs (bulk path for large batches; pinned memory recommended);
 * sigma_noise may be NULL to keep the previous one. */
int petmh_set_data_f32(petmh_t* h, int n_tac, const float* y, const float* tac_ref,
                       const float* k2p, const float* sigma_noise);

/* ---- the module-level helpers of kinetic_model.py, any grid, fp64 (callers of the reference's functions outside mcmc.py;
 * the sampler itself evaluates the same arithmetic in operator form and never calls these) ----
 * replaces interp1d_linear_vec(x, xp, fp) kinetic_model.py:35-57 (dim = 0): out[nx][m] = W fp[np][m] with the reference's
 * weights -- searchsorted-left node hi, lo = hi - 1 (index -1 wraps for x <= xp[0], as numpy's does), |xp[lo] - x| on hi and
 * |xp[hi] - x| on lo, normalised.  x beyond xp[np-1] is PETMH_EINVAL (the reference raises IndexError). */
int petmh_interp1d_linear(petmh_t* h, int nx, const double* x, int np, const double* xp, const double* fp, int m, double* out);
/* replaces estimate_continuous_convolution(x, y0, y1, num_points_resample) kinetic_model.py:12-32 = SRTM.convolve /
 * SRTM2.convolve (:125-128, :199-201): y0[n] (np.interp) and y1[n][m] (interp1d_linear_vec) resampled onto num_points
 * (0: the default 2 n) uniform points on [x[0], x[n-1]], causal discrete convolution truncated to that length times the grid
 * spacing, interpolated back to x: out[n][m].  x strictly increasing. */
int petmh_continuous_convolution(petmh_t* h, int n, const double* x, const double* y0, const double* y1, int m, int num_points,
                                 double* out);
/* replaces SRTM.make_time_exponential(param, time_vector) kinetic_model.py:118-122 (= SRTM2's, :192-196) without the
 * optional scales: out[nt][np] = exp(param[j] * t[i]). */
int petmh_time_exponential(petmh_t* h, int np, const double* param, int nt, const double* t, double* out);

/* ---- synthetic inputs on the GPU (K4; replaces sample_sim_data.py:141-215 for the training-style set) ----
 * Draw, per TAC, DVR / R1 from the handle's priors and the reference TAC from N(mu_tacref, cov_tacref)
 * with positivity rejection (helper_func.py:153-162), forward-simulate, redraw while any clean TAC value is
 * negative (sample_sim_data.py:171-188), add the truncated signal-dependent noise (:205-215) and leave
 * y (= noisy concentration), tac_ref and k2p in the handle as if petmh_set_data had been called.
 * The reference loops until a draw passes; here a vector is redrawn at most 4000 times and the triple at most 1000
 * times: if any TAC exhausts a cap the call returns PETMH_ESYNTH (the data of the other TACs is valid and bound). */
int petmh_synth(petmh_t* h, int n_tac, uint64_t seed, const double* mu_tacref54, const double* cov_tacref54x54,
                double k2p, const double* sigma_noise48x54);
/* Test-style rule of the reference's generator for the following petmh_synth calls (sample_sim_data.py:128-133,
 * flag_testing_data = True): a drawn vector x of DVR, R1 or the reference TAC is also rejected unless
 *   0 <= d2 < d2_max,   d2 = (x - mu)^T cov_inv (x - mu),   d2_max = chi2.ppf(alpha, 48)
 * -- the reference's chi2.cdf(mahalanobis(mu, x, Cov_inv) ** 2, 48) < alpha with dof 48 for all three variables (:132)
 * and the caller's inverses (the reference: np.linalg.inv(Cov), :106,110,117).  The inverse of the rank-deficient
 * Cov_tac_ref is numerically indefinite: about half of its forms come out negative, which scipy's mahalanobis turns into
 * NaN and the reference's comparison rejects; so does this rule.  d2_max <= 0 or a NULL inverse switches it off
 * (the training-style set, the default). */
int petmh_synth_set_test_rule(petmh_t* h, double d2_max, const double* cov_inv_dvr48x48, const double* cov_inv_r1_48x48,
                              const double* cov_inv_tacref54x54);
/* what petmh_synth generated (any pointer may be NULL): dvr_r1[n][96] f32, tac_ref[n][54] f64,
 * tac_clean[n][48][54] f32 and y[n][48][54] f32 in concentration units, attempts[n] (triples drawn; negative
 * when the TAC hit a rejection cap). */
int petmh_synth_get(petmh_t* h, float* dvr_r1, double* tac_ref, float* tac_clean, float* y, int* attempts);

/* ---- parity hooks ------------------------------------------------------------------ */
/* replaces CreateTAC_SRTM2.perform (mcmc.py:38-39) == SRTM2.create_activity_curve(...).T
 * (kinetic_model.py:142-158): out[48][54], unclamped. */
int petmh_forward(petmh_t* h, int tac, const double* dvr48, const double* r1_48, double* out48x54);
/* replaces SRTM.forward_model(DVR, k2, R1, tac_ref).T (kinetic_model.py:69-84; k2 free per ROI): out[48][54] */
int petmh_forward_srtm(petmh_t* h, int tac, const double* dvr48, const double* k2_48, const double* r1_48,
                       double* out48x54);
/* replaces the model log-probability pieces (mcmc.py:148-155): ll48[i] = sum_t log
 * TruncatedNormal(y_it | mu=sn_it, sigma=sqrt(sn_it) sigma_noise_it, lower=0) including the
 * state-independent constants; logprior2 = {log MvNormal(DVR), log MvNormal(R1)}. */
int petmh_loglik(petmh_t* h, int tac, const double* dvr48, const double* r1_48, double* ll48,
                 double* logprior2);
/* the device-built exact convolution operator M (54x54) of estimate_continuous_convolution
 * (kinetic_model.py:12-32): conv = M @ exp(-k2a t). */
int petmh_get_operator(petmh_t* h, int tac, double* m54x54);
/* the Chebyshev-in-k2a form of the same operator that the sweep kernel evaluates:
 * conv[j] = sum_d A[j][d] T_d(s), s = (2 k2a - k2a_lo - k2a_hi) / (k2a_hi - k2a_lo), valid for k2a in [k2a_lo, k2a_hi]
 * (outside it the kernel uses M itself).  a[520] f32 is laid out [3 row blocks of 18 frames][ncols3[b] columns][20]
 * (18 rows + 2 pad).  Any output pointer may be NULL.  Replaces the same lines as petmh_get_operator. */
int petmh_get_cheb_operator(petmh_t* h, int tac, float* a520, int* ncols3, double* k2a_lo, double* k2a_hi);
/* raw Philox4x32-10 words the kernel draws for (chain gid, sweep, block): out[48][4] */
int petmh_philox_raw(petmh_t* h, uint64_t chain_gid, uint32_t sweep, uint32_t block, uint32_t* out48x4);

/* ---- sampling (replaces pm.sample(draws, tune, step=pm.Metropolis(...)) mcmc.py:156-157) --- */
/* Start all chains at the prior mean with scaling 1 (pymc defaults), sweep counter 0. */
int petmh_reset(petmh_t* h);
/* tune sweeps with pymc's scaling table every 100 sweeps, then draws sweeps with frozen
 * scaling; every thin-th draw is stored when max_draws > 0; running moments always.
 * One sweep = 96 chain-steps per chain.  Equivalent to reset + plan + advance(tune+draws). */
int petmh_run(petmh_t* h, int draws, int tune, int thin);
/* Lower level: declare the schedule (sweeps [0,tune) tune, [tune,tune+draws) are draws),
 * then continue every chain for n_sweeps at a time (checkpointable between calls). */
int petmh_plan(petmh_t* h, int draws, int tune, int thin);
int petmh_advance(petmh_t* h, int n_sweeps);
/* Teacher-forcing / decision-parity mode: like petmh_run for the first n_tape_chains chains
 * of TAC `tac` only, consuming an explicit random tape instead of Philox:
 *   normals[c][s][2][48] f32, logu[c][s][2][48] f32, rank[c][s][2][48] u8 (visit position)
 * and recording every sweep: draws_out[c][s][2][48] f32, optional delta_out (f32, the
 * log acceptance ratio at decision time) and accept_out (u8), scale_out[c][2][48] f32. */
int petmh_run_taped(petmh_t* h, int tac, int n_tape_chains, int n_sweeps, int tune,
                    const float* normals, const float* logu, const uint8_t* rank,
                    float* draws_out, float* delta_out, uint8_t* accept_out, float* scale_out);

/* ---- the k2-free SRTM as a SAMPLED model ------------------------------------------------------------
 * kinetic_model.py:62-84 SRTM.forward_model(DVR, k2, R1, tac_ref) with the reference's likelihood block (mcmc.py:151-155),
 * the handle's MvNormal priors on DVR / R1 (mcmc.py:148-149) and a caller-supplied MvNormal prior on k2[48] (the reference
 * ships none and mcmc.py never samples k2: this goes beyond it, with pm.Metropolis / CompoundStep semantics over the three
 * blocks DVR, R1, k2 in that order).  Chains start at the prior means with scaling 1.  Self-contained: uses the handle's
 * frames / prior / data, leaves its SRTM2 chain state untouched.  A simple kernel (one CTA per chain), not the hot path.
 *   free run:  tape_* = NULL; draws_out[n_tac][n_chains][ceil(draws/thin)][3][48] f32 (block order DVR, R1, k2);
 *              accept_rate_out[n_tac][n_chains][3][48] (may be NULL)
 *   taped run: n_tape_chains chains of TAC tape_tac off tape_normals / tape_logu / tape_rank [c][draws+tune][3][48]
 *              (f32, f32, u8 visit position); EVERY sweep is recorded: draws_out[c][draws+tune][3][48], optional
 *              delta_out (f32 log acceptance ratio at decision time) and accept_out (u8). */
int petmh_srtm_sample(petmh_t* h, const double* mu_k2_48, const double* cov_k2_48x48, int draws, int tune, int thin,
                      int n_tape_chains, int tape_tac, const float* tape_normals, const float* tape_logu,
                      const uint8_t* tape_rank, float* draws_out, float* delta_out, uint8_t* accept_out,
                      float* accept_rate_out);

/* ---- outputs (replaces idata.posterior[...] / pm.summary / pm.rhat mcmc.py:162-187) ---- */
int petmh_n_stored(const petmh_t* h); /* stored draws per chain so far */
/* dvr/r1: [n_tac][n_chains][n_stored][48] float32 (DVR_mcmc / R1_mcmc, mcmc.py:162-163) */
int petmh_get_chains(petmh_t* h, float* dvr, float* r1);
/* out[n_tac][96][PETMH_N_STATS] f32: mean, sd, mcse_mean, ess_bulk, ess_tail, r_hat,
 * accept_rate, scaling -- from stored draws when max_draws > 0 (rank-normalised split
 * R-hat, bulk / tail ESS and MCSE as ArviZ computes them), else from running split-half moments: the
 * ess_bulk column then holds a BATCH-MEANS effective sample size (8 batches per half chain; needs >= 4 closed
 * batches, else NaN), r_hat the CLASSIC (not rank-normalised) split R-hat, ess_tail NaN.  Split halves as ArviZ:
 * the first and last draws/2 draws of the plan; an odd middle draw counts for neither. */
int petmh_get_summary(petmh_t* h, float* out);
/* the remaining pm.summary columns (mcmc.py:181), from the stored draws (max_draws > 0, >= 8 stored):
 * out[n_tac][96][4] f32 = hdi_3%, hdi_97% (arviz.hdi, hdi_prob 0.94: narrowest interval of the pooled sorted draws),
 * mcse_sd, ess_sd (arviz _mcse_sd / _ess_sd).  Computed together with petmh_get_summary and cached. */
int petmh_get_summary_ext(petmh_t* h, float* out);
/* same, written to a DEVICE buffer (e.g. a slice of an NCCL all-gather buffer) on
 * `stream` (a cudaStream_t, 0 = the handle's). */
int petmh_summary_device(petmh_t* h, float* d_out, void* stream);
/* Summaries of state gathered from several handles -- the chains of ONE TAC sampled on several GPUs (SURVEY.md 8e:
 * "if S < G shard chains"); all pointers are DEVICE pointers on `device`, no handle involved (errors: petmh_last_error(NULL)).
 *   from_draws:   d_draws[n_tac*n_chains][n_stored][96] f32 (dense), d_nacc / d_scale [n_tac*n_chains][96]
 *                 -> d_out8[n_tac][96][8] (as petmh_get_summary) and, if not NULL, d_ext4[n_tac][96][4] (as petmh_get_summary_ext)
 *   from_moments: d_mom[n_tac*n_chains][2][96][5] f32 as the sweep kernel keeps them, d_mu96 = prior means (f64),
 *                 n_half2 / n_batch2 / batch_len / n_draw_sweeps as petmh_export_summary_inputs reports them.
 * petmh_export_summary_inputs hands out the handle's own device buffers and counters for such a gather. */
int petmh_summary_from_draws_device(int device, const float* d_draws, int n_tac, int n_chains, int n_stored,
                                    const uint32_t* d_nacc, const float* d_scale, int n_draw_sweeps,
                                    float* d_out8, float* d_ext4, void* stream);
int petmh_summary_from_moments_device(int device, const float* d_mom, const double* d_mu96, int n_tac, int n_chains,
                                      const int* n_half2, const int* n_batch2, int batch_len, const uint32_t* d_nacc,
                                      const float* d_scale, int n_draw_sweeps, float* d_out8, void* stream);
int petmh_export_summary_inputs(petmh_t* h, void** d_draws, void** d_mom, void** d_nacc, void** d_scale, void** d_mu96,
                                int* counters8);
/* out[n_tac][96] f32: cross-chain effective sample size of every coordinate from the stored draws,
 * as the consumer computes it with tfp.mcmc.effective_sample_size(..., cross_chain_dims=-1)
 * (main_script.py:807-810; TFP defaults: lags from the first negative autocorrelation on dropped).
 * Needs max_draws > 0 and >= 2 stored draws; one chain falls back to the single-chain formula. */
int petmh_get_ess_cross_chain(petmh_t* h, float* out);
/* cov[n_tac][2][48][48], corr[n_tac][2][48][48] f64 (either may be NULL; block 0 = DVR, 1 = R1): covariance (ddof = 1) and
 * correlation across ROIs of the pooled stored draws, as the consumer forms them for its tables with
 * np.cov(mcmc_res['DVR_mcmc'].reshape([-1, n_ROI]), rowvar=False) / np.corrcoef(...) (main_script.py:717-738).
 * Needs max_draws > 0 and >= 2 pooled draws. */
int petmh_get_posterior_cov(petmh_t* h, double* cov, double* corr);
int petmh_get_state(petmh_t* h, float* q /*[n_tac][n_chains][96]*/, float* scale /*same*/);
/* Warm start: overwrite every chain's position and scaling (either may be NULL), ZERO the tuning counters, the
 * accepted-move counters and the running moments, and set the sweep counter.  Not a resume -- see
 * petmh_get_checkpoint / petmh_set_checkpoint for that. */
int petmh_set_state(petmh_t* h, const float* q, const float* scale, int sweep);

/* Full checkpoint (replaces nothing: the reference can only skip finished samples, mcmc.py:125-128): positions, scalings,
 * PyMC tune counters, accepted-move counters, running moments, the stored draws and the schedule position, as one
 * host blob.  After petmh_set_data (same TACs) + petmh_set_checkpoint on a handle with the same n_chains / seed, the
 * run continues bit for bit as if it had never stopped -- also inside a 100-sweep tuning window.  Returns
 * PETMH_ESTATE when the blob does not fit the handle. */
int64_t petmh_checkpoint_bytes(const petmh_t* h);
int petmh_get_checkpoint(petmh_t* h, void* buf, int64_t nbytes);
int petmh_set_checkpoint(petmh_t* h, const void* buf, int64_t nbytes);

/* ---- timing / stream ------------------------------------------------------------- */
int petmh_set_stream(petmh_t* h, void* cuda_stream);
int petmh_synchronize(petmh_t* h);
/* CUDA-event time of the sweep-kernel launches of the last run/advance, and their count */
int petmh_last_kernel_ms(const petmh_t* h, float* ms, int* launches);

#ifdef __cplusplus
}
#endif
#endif /* PETMH_H */
