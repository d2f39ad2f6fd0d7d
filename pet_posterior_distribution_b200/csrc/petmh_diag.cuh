// petmh_diag.cuh -- on-GPU posterior summaries (K3): replaces pm.summary / pm.rhat
// (mcmc.py:181,186-187; ArviZ semantics) for the columns
//   mean, sd(ddof=1), mcse_mean, ess_bulk, ess_tail, r_hat, accept_rate, scaling.
// Two sources:
//   * running split-half moments kept by the sweep kernel (always available, O(1) memory
//     per chain): classic split R-hat, AR(1) effective sample size;
//   * stored draws (max_draws > 0): rank-normalised split R-hat and Geyer ESS as ArviZ
//     computes them (see petmh_rank_diag.cuh).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace petmh {

struct DiagParams {
    const float* mom;       // [S*C][2][96][3]  mean-mu, M2, lag-1 co-moment
    const double* mu;       // [2][48]
    const uint32_t* nacc;   // [S*C][96]
    const float* scale;     // [S*C][96]
    const float* draws;     // [S*C][max_draws][96] or null
    int n_tacs, n_chains, max_draws, n_stored;
    int n_half[2];          // draws merged in each half
    int lag_terms[2];       // number of lag-1 products in each half (per chain)
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
    const int nh = (d.n_half[0] > 0) + (d.n_half[1] > 0);
    double tot_n = 0, tot_sum = 0, tot_ss = 0;       // pooled, about mu
    double w_sum = 0, cm_sum = 0, cm_sq = 0, c1_sum = 0, lagn = 0;
    int m = 0;
    double acc = 0, sc = 0;
    for (int c = 0; c < C; c++) {
        const size_t cg = tac * C + c;
        for (int hf = 0; hf < 2; hf++) {
            const int n = d.n_half[hf];
            if (n == 0) continue;
            const float* mo = d.mom + ((cg * 2 + hf) * 96 + coord) * 3;
            const double mean = mo[0], M2 = mo[1], C1 = mo[2];
            tot_n += n;
            tot_sum += mean * n;
            tot_ss += M2 + mean * mean * n;
            if (n > 1) w_sum += M2 / (n - 1);
            cm_sum += mean;
            cm_sq += mean * mean;
            c1_sum += C1;
            lagn += d.lag_terms[hf];
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
        double rhat = nanv;
        // classic split R-hat over the m = 2C half chains (equal length n when both halves exist)
        if (m > 1 && nh == 2 && d.n_half[0] == d.n_half[1]) {
            const double n = d.n_half[0];
            const double B_over_n = (cm_sq - cm_sum * cm_sum / m) / (m - 1);
            rhat = sqrt(((n - 1) / n * W + B_over_n) / W);
        }
        // AR(1) effective sample size from the pooled lag-1 autocorrelation
        double ess = nanv;
        if (lagn > 0 && W > 0) {
            double rho = (c1_sum / lagn) / (W);
            rho = fmin(fmax(rho, -0.999), 0.999999);
            ess = tot_n * (1.0 - rho) / (1.0 + rho);
            ess = fmin(ess, tot_n);
        }
        o[0] = (float)(gmean + mu);
        o[1] = (float)sd;
        o[2] = (float)(sd / sqrt(ess));
        o[3] = (float)ess;
        o[4] = nanv;
        o[5] = (float)rhat;
    }
    o[6] = tot_n > 0 ? (float)(acc / tot_n) : nanv;
    o[7] = (float)(sc / C);
}

static inline int launch_summary(const DiagParams& d, cudaStream_t st) {
    const size_t n = (size_t)d.n_tacs * 96;
    summary_moments_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d);
    return (int)cudaGetLastError();
}

}  // namespace petmh
