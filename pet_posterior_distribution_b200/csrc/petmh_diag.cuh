// petmh_diag.cuh -- on-GPU posterior summaries (K3) from the RUNNING MOMENTS of the sweep kernel: the
// O(1)-memory mode (max_draws = 0, e.g. BASELINE configs[4]: a million TACs cannot store their draws).
// Stands in for pm.summary / pm.rhat (mcmc.py:181,186-187) with estimators that need no stored draws:
//   mean, sd (ddof = 1)      pooled over every chain and counted draw
//   r_hat                    CLASSIC split R-hat (Gelman et al. 2013) over the 2C half chains -- not rank-normalised
//   ess_bulk, mcse_mean      BATCH-MEANS effective sample size: each half chain is cut into NB = 8 batches of B draws,
//                            sigma2_inf = B var(batch means about the grand mean), ESS = N var_plus / sigma2_inf,
//                            var_plus = (n-1)/n W + (between half-chain variance) as ArviZ's _ess uses it; the
//                            relative bias is ~ tau / B (autocorrelation time over batch length)
//   ess_tail                 NaN (needs order statistics)
// The stored-draw path (petmh_rankdiag.cuh) is the one that follows ArviZ; tests compare the two on the same run.
// ArviZ's split drops the middle draw of an odd-length chain: so do the moments (petmh_advance).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace petmh {

struct DiagParams {
    const float* mom;       // [S*C][2][96][MOMF]  mean-mu, M2, open batch sum, mean / M2 of the closed batch means
    const double* mu;       // [2][48]
    const uint32_t* nacc;   // [S*C][96]
    const float* scale;     // [S*C][96]
    int n_tacs, n_chains;
    int n_half[2];          // draws merged in each half (per chain)
    int n_batch[2];         // closed batches in each half (per chain)
    int batch_len;          // draws per batch
    int n_draw_sweeps;      // draw sweeps done (accept-rate denominator)
    float* out;             // [S][96][8]
};

// one thread per (tac, coord)
__global__ void summary_moments_kernel(const DiagParams d) {
    const size_t gi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= (size_t)d.n_tacs * 96) return;
    const int coord = (int)(gi % 96);
    const size_t tac = gi / 96;
    const int C = d.n_chains;
    const double mu = d.mu[coord];
    double tot_n = 0, tot_sum = 0, tot_ss = 0;       // pooled, about mu
    double w_sum = 0, cm_sum = 0, cm_sq = 0;
    double bm_k = 0, bm_sum = 0;                     // batch means: count, sum
    int m = 0;
    double acc = 0, sc = 0;
    for (int c = 0; c < C; c++) {
        const size_t cg = tac * C + c;
        for (int hf = 0; hf < 2; hf++) {
            const int n = d.n_half[hf];
            if (n == 0) continue;
            const float* mo = d.mom + ((cg * 2 + hf) * 96 + coord) * MOMF;
            const double mean = mo[0], M2 = mo[1];
            tot_n += n;
            tot_sum += mean * n;
            tot_ss += M2 + mean * mean * n;
            if (n > 1) w_sum += M2 / (n - 1);
            cm_sum += mean;
            cm_sq += mean * mean;
            bm_k += d.n_batch[hf];
            bm_sum += (double)mo[3] * d.n_batch[hf];
            m++;
        }
        acc += d.nacc[cg * 96 + coord];
        sc += d.scale[cg * 96 + coord];
    }
    float* o = d.out + gi * 8;
    const float nanv = CUDART_NAN_F;
    if (tot_n < 2) {
        for (int k = 0; k < 6; k++) o[k] = nanv;
    } else {
        const double gmean = tot_sum / tot_n;
        const double var = (tot_ss - tot_sum * gmean) / (tot_n - 1);
        const double sd = sqrt(fmax(var, 0.0));
        const double W = w_sum / m;
        double rhat = nanv, ess = nanv;
        // split statistics need both halves with the same length (always true once a run is complete)
        if (m > 1 && d.n_half[0] == d.n_half[1] && d.n_half[0] > 1 && W > 0) {
            const double n = d.n_half[0];
            const double B_over_n = (cm_sq - cm_sum * cm_sum / m) / (m - 1);
            const double var_plus = (n - 1) / n * W + B_over_n;
            rhat = sqrt(var_plus / W);
            if (bm_k >= 4) {
                const double g = bm_sum / bm_k;
                double ss = 0;
                for (int c = 0; c < C; c++)
                    for (int hf = 0; hf < 2; hf++) {
                        const float* mo = d.mom + (((tac * C + c) * 2 + hf) * 96 + coord) * MOMF;
                        const double dm = (double)mo[3] - g;
                        ss += (double)mo[4] + d.n_batch[hf] * dm * dm;
                    }
                const double s2inf = d.batch_len * ss / (bm_k - 1);
                if (s2inf > 0) ess = fmin(tot_n * var_plus / s2inf, tot_n * log10(tot_n));
            }
        }
        o[0] = (float)(gmean + mu);
        o[1] = (float)sd;
        o[2] = (float)(sd / sqrt(ess));
        o[3] = (float)ess;
        o[4] = nanv;
        o[5] = (float)rhat;
    }
    o[6] = d.n_draw_sweeps > 0 ? (float)(acc / ((double)C * d.n_draw_sweeps)) : nanv;
    o[7] = (float)(sc / C);
}

static inline int launch_summary(const DiagParams& d, cudaStream_t st) {
    const size_t n = (size_t)d.n_tacs * 96;
    summary_moments_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d);
    return (int)cudaGetLastError();
}

}  // namespace petmh
