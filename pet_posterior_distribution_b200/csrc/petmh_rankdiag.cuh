// petmh_rankdiag.cuh -- K3, stored-draw diagnostics on the GPU: rank-normalised split R-hat,
// bulk / tail / mean effective sample sizes and MCSE as ArviZ computes them for
// pm.summary / pm.rhat (mcmc.py:181,186-187; Vehtari et al. 2021).  PARITY UNPINNED
// (arviz is third-party and absent): checked against oracle/diagnostics.py only.
//
// Per (TAC, coordinate) segment: the C chains x N stored draws are split in halves
// (2C chains of h = N/2; an odd middle draw is dropped), globally ranked with ties averaged
// (one 64-bit-key radix sort over all segments: key = segment << 32 | order-preserving
// float bits), mapped through the inverse normal cdf, and fed to the Geyer-truncated
// autocorrelation sum.  Not a hot path: clarity over speed.
#pragma once
#include <algorithm>
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace petmh {

struct RankDiagParams {
    const float* draws;   // [S*C][max_draws][96]
    int n_chains, max_draws, n_stored;
    int tac0, n_tac;      // TAC batch
    int h;                // half length
    int L;                // 2*C*h values per segment
};

__device__ __forceinline__ uint32_t f2sortable(float f) {
    uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float sortable2f(uint32_t u) {
    u ^= ((u >> 31) ? 0x80000000u : 0xffffffffu);
    return __uint_as_float(u);
}

// element k of a segment -> (chain, draw): split chains ordered [first halves..., second halves...]
__device__ __forceinline__ float seg_value(const RankDiagParams& p, int seg, int k) {
    const int tac = p.tac0 + seg / 96, coord = seg % 96;
    const int sc = k / p.h, i = k - sc * p.h;
    const int c = sc % p.n_chains;
    const int d = sc < p.n_chains ? i : p.n_stored - p.h + i;
    return p.draws[(((size_t)tac * p.n_chains + c) * p.max_draws + d) * 96 + coord];
}

// mode 0: x; mode 1: |x - median|
__global__ void rd_build_keys(const RankDiagParams p, int mode, const float* med, float* xs, unsigned long long* keys,
                              uint32_t* vals, size_t n) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int seg = (int)(e / p.L), k = (int)(e - (size_t)seg * p.L);
    float x = seg_value(p, seg, k);
    if (mode == 0) xs[e] = x;
    else x = fabsf(x - med[seg]);
    keys[e] = ((unsigned long long)seg << 32) | f2sortable(x);
    vals[e] = (uint32_t)k;
}

// sorted keys/vals -> z-scores scattered back to original positions (ties: average rank)
__global__ void rd_ranks(const unsigned long long* keys, const uint32_t* vals, float* z, int L, size_t n) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const size_t seg0 = (e / L) * (size_t)L;
    if (e != seg0 && keys[e - 1] == keys[e]) return;   // not a run start
    size_t b = e + 1;
    const size_t seg_end = seg0 + L;
    while (b < seg_end && keys[b] == keys[e]) b++;
    const double rank = 0.5 * ((double)(e - seg0) + (double)(b - 1 - seg0)) + 1.0;
    const float zz = (float)normcdfinv((rank - 0.375) / ((double)L + 0.25));
    for (size_t q = e; q < b; q++) z[seg0 + vals[q]] = zz;
}

// per segment: median, q05, q95 (numpy linear-interpolation quantiles) from the sorted keys
__global__ void rd_quantiles(const unsigned long long* keys, int L, int nseg, float* med, float* q05, float* q95) {
    const int seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= nseg) return;
    const unsigned long long* k = keys + (size_t)seg * L;
    auto at = [&](int i) { return (double)sortable2f((uint32_t)k[i]); };
    auto quant = [&](double q) {
        const double pos = q * (L - 1);
        const int lo = (int)floor(pos);
        const int hi = min(lo + 1, L - 1);
        const double fr = pos - lo;
        return (float)(at(lo) + (at(hi) - at(lo)) * fr);
    };
    med[seg] = quant(0.5);
    q05[seg] = quant(0.05);
    q95[seg] = quant(0.95);
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int tid = threadIdx.x;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) sh[tid >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sh[w];
    return t;
}

// series value of type ty at element e: 0 = z (bulk), 1 = x <= q05, 2 = x <= q95, 3 = x, 4 = x^2 (arviz _ess_sd)
constexpr int N_SERIES = 5;
__device__ __forceinline__ float series(int ty, const float* xs, const float* z, float q05, float q95, size_t e) {
    if (ty == 0) return z[e];
    if (ty == 3) return xs[e];
    if (ty == 4) return xs[e] * xs[e];
    const float x = xs[e];
    return (ty == 1 ? x <= q05 : x <= q95) ? 1.f : 0.f;
}

// rank-normalised split R-hat: max of bulk (z) and tail (z of folded); one CTA per segment
__global__ void rd_rhat(const float* z, const float* zf, int m, int h, float* rhat_out) {
    __shared__ double sh[32];
    const int seg = blockIdx.x;
    float best = 0.f;
    for (int which = 0; which < 2; which++) {
        const float* a = (which ? zf : z) + (size_t)seg * m * h;
        double sum_mean = 0, sum_mean2 = 0, sum_var = 0;
        for (int c = 0; c < m; c++) {
            double s = 0;
            for (int i = threadIdx.x; i < h; i += blockDim.x) s += a[(size_t)c * h + i];
            const double mean = block_sum(s, sh) / h;
            double q = 0;
            for (int i = threadIdx.x; i < h; i += blockDim.x) { const double d = a[(size_t)c * h + i] - mean; q += d * d; }
            const double var = block_sum(q, sh) / (h - 1);
            sum_mean += mean; sum_mean2 += mean * mean; sum_var += var;
        }
        const double between = (double)h * (sum_mean2 - sum_mean * sum_mean / m) / (m - 1);
        const double within = sum_var / m;
        const float r = (float)sqrt((between / within + h - 1) / h);
        best = which == 0 ? r : fmaxf(best, r);
    }
    if (threadIdx.x == 0) rhat_out[seg] = best;
}

// arviz _ess for one (segment, series type).  A batch of lags = ESS_LB / ng lags x ng chain groups; ng (4, 2 or 1) is
// the number of chains staged in dynamic shared memory at a time (ng x h floats, up to ESS_STAGE_BYTES): h = 10 000 (the
// reference's 20 000 draws) stages 4 chains, h <= 51 200 one; beyond that (staged = 0) the fp64 path reads global memory.
constexpr int ESS_LB = 256;
constexpr int ESS_STAGE_BYTES = 200 * 1024;
__global__ void __launch_bounds__(ESS_LB) rd_ess(const float* xs, const float* z, const float* q05, const float* q95, int m,
                                                 int h, float* rho_scratch /*[nseg*N_SERIES][h]*/, float* ess_out /*[nseg][N_SERIES]*/,
                                                 int staged /* dynamic smem holds ng x h floats */, int ng) {
    extern __shared__ float ess_stage[];   // [chain group][h]: the centred series of the chain the group is working on
    __shared__ double sh[32];
    __shared__ double cmean[2048];
    __shared__ int s_stop, s_t;
    __shared__ double s_even, s_odd;
    const int seg = blockIdx.x / N_SERIES, ty = blockIdx.x % N_SERIES, tid = threadIdx.x;
    const int ESS_LAGS = ESS_LB / ng;
    const size_t base = (size_t)seg * m * h;
    const float a05 = q05[seg], a95 = q95[seg];
    float* rho = rho_scratch + (size_t)blockIdx.x * h;
    // chain means, biased variances
    double sum_mean = 0, sum_mean2 = 0, acov0 = 0;
    for (int c = 0; c < m; c++) {
        double s = 0;
        for (int i = tid; i < h; i += blockDim.x) s += series(ty, xs, z, a05, a95, base + (size_t)c * h + i);
        const double mean = block_sum(s, sh) / h;
        double q = 0;
        for (int i = tid; i < h; i += blockDim.x) { const double d = series(ty, xs, z, a05, a95, base + (size_t)c * h + i) - mean; q += d * d; }
        const double v = block_sum(q, sh) / h;
        if (tid == 0 && c < 2048) cmean[c] = mean;
        sum_mean += mean; sum_mean2 += mean * mean; acov0 += v;
    }
    __syncthreads();
    const double mean_var = acov0 / m * h / (h - 1.0);
    double var_plus = mean_var * (h - 1.0) / h;
    if (m > 1) var_plus += (sum_mean2 - sum_mean * sum_mean / m) / (m - 1);
    const double ntot = (double)m * h;
    if (!(var_plus > 0.0) || h < 4) {
        if (tid == 0) ess_out[seg * N_SERIES + ty] = CUDART_NAN_F;
        return;
    }
    if (tid == 0) { s_stop = 0; s_t = 1; s_even = 1.0; s_odd = 0.0; rho[0] = 1.f; }
    __syncthreads();
    int have = 0;   // lags [1, have] available in rho as raw "1 - (mean_var - acov)/var_plus"
    int max_t = -1;
    // a batch = ESS_LAGS lags x ESS_LB / ESS_LAGS chain groups (the Geyer cut-off is usually well inside the first batch)
    __shared__ double part[ESS_LB];
    const int tl = tid % ESS_LAGS, grp = tid / ESS_LAGS;
    while (true) {
        // compute lags have+1 .. have+ESS_LAGS
        const int t = have + 1 + tl;
        double acc = 0.0;
        if (staged) {
            // Staged path: each group centres its chain once per batch into shared memory as fp32 (the draws are fp32:
            // rounding x - mean to fp32 is below their own quantisation) and forms the lag products with fp32 FMAs in
            // chunks of 8, added up in fp64 -- one LDS + one FFMA per product instead of two conversions, two fp64
            // subtractions and a DFMA (the loop was bound by the FP64 pipe).  Uniform trip count: barriers inside.
            const int NG = ng;
            float* cb = ess_stage + (size_t)grp * h;
            for (int c0 = 0; c0 < m; c0 += NG) {
                const int c = c0 + grp;
                __syncthreads();
                if (c < m) {
                    const double mu = cmean[min(c, 2047)];
                    const size_t o = base + (size_t)c * h;
                    for (int i = tl; i < h; i += ESS_LAGS) cb[i] = (float)((double)series(ty, xs, z, a05, a95, o + i) - mu);
                }
                __syncthreads();
                if (c < m && t < h) {
                    const int nn = h - t;
                    const float* ca = cb;
                    const float* cl = cb + t;
                    double a = 0.0;
                    int n = 0;
                    for (; n + 15 < nn; n += 16) {
                        float s0 = 0.f, s1 = 0.f;
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            s0 = fmaf(ca[n + u], cl[n + u], s0);
                            s1 = fmaf(ca[n + 8 + u], cl[n + 8 + u], s1);
                        }
                        a += (double)s0 + (double)s1;
                    }
                    for (; n < nn; n++) a += (double)(ca[n] * cl[n]);
                    acc += a / h;
                }
            }
        } else if (t < h) {
            for (int c = grp; c < m; c += ng) {
                const double mu = cmean[min(c, 2047)];
                const size_t o = base + (size_t)c * h;
                // four independent partial sums: the loop is bound by the latency of the fp64 accumulation chain
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                const int nn = h - t;
                int n = 0;
                for (; n + 3 < nn; n += 4) {
                    a0 += (series(ty, xs, z, a05, a95, o + n) - mu) * (series(ty, xs, z, a05, a95, o + n + t) - mu);
                    a1 += (series(ty, xs, z, a05, a95, o + n + 1) - mu) * (series(ty, xs, z, a05, a95, o + n + 1 + t) - mu);
                    a2 += (series(ty, xs, z, a05, a95, o + n + 2) - mu) * (series(ty, xs, z, a05, a95, o + n + 2 + t) - mu);
                    a3 += (series(ty, xs, z, a05, a95, o + n + 3) - mu) * (series(ty, xs, z, a05, a95, o + n + 3 + t) - mu);
                }
                for (; n < nn; n++) a0 += (series(ty, xs, z, a05, a95, o + n) - mu) * (series(ty, xs, z, a05, a95, o + n + t) - mu);
                acc += ((a0 + a1) + (a2 + a3)) / h;
            }
        }
        part[tid] = acc;
        __syncthreads();
        if (grp == 0 && t < h) {
            for (int g2 = 1; g2 < ng; g2++) acc += part[g2 * ESS_LAGS + tl];
            rho[t] = (float)(1.0 - (mean_var - acc / m) / var_plus);
        }
        __syncthreads();
        have = min(have + ESS_LAGS, h - 1);
        if (tid == 0) {
            // Geyer's initial positive sequence, resumed where it stopped
            int tt = s_t;
            double even = s_even, odd = s_odd;
            if (tt == 1 && s_odd == 0.0) { odd = rho[1]; }
            bool stop = false;
            while (true) {
                if (!(tt < (h - 3) && (even + odd) > 0.0)) { stop = true; break; }
                if (tt + 2 > have) break;                       // need more lags
                even = rho[tt + 1];
                odd = rho[tt + 2];
                if (!((even + odd) >= 0.0)) { rho[tt + 1] = 0.f; rho[tt + 2] = 0.f; }
                tt += 2;
            }
            s_t = tt; s_even = even; s_odd = odd; s_stop = stop ? 1 : 0;
        }
        __syncthreads();
        if (s_stop || have >= h - 1) break;
    }
    if (tid == 0) {
        const int tt = s_t;
        max_t = tt - 2;
        // raw values beyond the accepted prefix must read as zero (arviz's rho_hat_t is zero-initialised)
        const double even = s_even;
        if (even > 0 && max_t + 1 >= 0 && max_t + 1 < h) rho[max_t + 1] = (float)even;
        // initial monotone sequence
        int t2 = 1;
        while (t2 <= max_t - 2) {
            if ((rho[t2 + 1] + rho[t2 + 2]) > (rho[t2 - 1] + rho[t2])) {
                rho[t2 + 1] = (rho[t2 - 1] + rho[t2]) / 2.f;
                rho[t2 + 2] = rho[t2 + 1];
            }
            t2 += 2;
        }
        double ssum = 0.0;
        for (int i = 0; i <= max_t; i++) ssum += rho[i];
        double tau = -1.0 + 2.0 * ssum + ((max_t + 1 >= 0 && max_t + 1 < h) ? (double)rho[max_t + 1] : 0.0);
        tau = fmax(tau, 1.0 / log10(ntot));
        ess_out[seg * N_SERIES + ty] = (float)(ntot / tau);
    }
}

// arviz hdi (unimodal, hdi_prob): the narrowest interval holding floor(prob L) + 1 of the segment's sorted values;
// one CTA per segment, first minimum wins (numpy argmin)
__global__ void rd_hdi(const unsigned long long* keys, int L, double prob, float* lo_out, float* hi_out) {
    __shared__ float s_w[32];
    __shared__ int s_i[32];
    const int seg = blockIdx.x, tid = threadIdx.x;
    const unsigned long long* k = keys + (size_t)seg * L;
    const int inc = (int)floor(prob * L), n_int = L - inc;
    float best = CUDART_INF_F;
    int bi = 0x7fffffff;
    if (inc >= 1 && n_int >= 1) {
        for (int i = tid; i < n_int; i += blockDim.x) {
            const float w = (float)((double)sortable2f((uint32_t)k[i + inc]) - (double)sortable2f((uint32_t)k[i]));
            if (w < best) { best = w; bi = i; }      // (ascending i per thread: keeps the first minimum)
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float w2 = __shfl_down_sync(0xffffffffu, best, o);
        const int i2 = __shfl_down_sync(0xffffffffu, bi, o);
        if (w2 < best || (w2 == best && i2 < bi)) { best = w2; bi = i2; }
    }
    if ((tid & 31) == 0) { s_w[tid >> 5] = best; s_i[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++)
            if (s_w[w] < best || (s_w[w] == best && s_i[w] < bi)) { best = s_w[w]; bi = s_i[w]; }
        if (bi == 0x7fffffff) { lo_out[seg] = sortable2f((uint32_t)k[0]); hi_out[seg] = sortable2f((uint32_t)k[L - 1]); }
        else { lo_out[seg] = sortable2f((uint32_t)k[bi]); hi_out[seg] = sortable2f((uint32_t)k[bi + inc]); }
    }
}

// pooled mean / sd (ddof = 1) over ALL stored draws of a segment (pm.summary does not split; for an odd number of
// draws the split series drop the middle one) + assembly of the 8-column row and of the 4 extra pm.summary columns
__global__ void rd_finalize(const RankDiagParams p, const float* ess, const float* rhat, const float* hlo, const float* hhi,
                            const uint32_t* nacc, const float* scale, int n_draw_sweeps, float* out, float* ext) {
    __shared__ double sh[32];
    const int seg = blockIdx.x;
    const int tac = p.tac0 + seg / 96, coord = seg % 96;
    const int n = p.n_chains * p.n_stored;
    auto val = [&](int i) {
        const int c = i / p.n_stored, d = i - c * p.n_stored;
        return (double)p.draws[(((size_t)tac * p.n_chains + c) * p.max_draws + d) * 96 + coord];
    };
    double s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += val(i);
    const double mean = block_sum(s, sh) / n;
    double q = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const double d = val(i) - mean; q += d * d; }
    const double sd = sqrt(block_sum(q, sh) / (n - 1));
    if (threadIdx.x == 0) {
        double acc = 0, sc = 0;
        for (int c = 0; c < p.n_chains; c++) {
            acc += nacc[((size_t)tac * p.n_chains + c) * 96 + coord];
            sc += scale[((size_t)tac * p.n_chains + c) * 96 + coord];
        }
        const float* e = ess + (size_t)seg * N_SERIES;
        float* o = out + ((size_t)tac * 96 + coord) * 8;
        o[0] = (float)mean;
        o[1] = (float)sd;
        o[2] = (float)(sd / sqrt((double)e[3]));
        o[3] = e[0];
        o[4] = fminf(e[1], e[2]);
        o[5] = rhat[seg];
        o[6] = n_draw_sweeps > 0 ? (float)(acc / ((double)p.n_chains * n_draw_sweeps)) : CUDART_NAN_F;
        o[7] = (float)(sc / p.n_chains);
        if (ext) {   // hdi_3%, hdi_97%, mcse_sd, ess_sd (arviz _mcse_sd / _ess_sd)
            float* x = ext + ((size_t)tac * 96 + coord) * 4;
            const double esd = fmin((double)e[3], (double)e[4]);
            x[0] = hlo[seg];
            x[1] = hhi[seg];
            x[2] = (float)(sd * sqrt(exp(1.0) * pow(1.0 - 1.0 / esd, esd - 1.0) - 1.0));
            x[3] = (float)esd;
        }
    }
}

// Host driver.  Returns a cudaError_t (0 = ok).
static inline int launch_rank_summary(const float* d_draws, int n_tac_total, int n_chains, int max_draws, int n_stored,
                                      const uint32_t* nacc, const float* scale, int n_draw_sweeps, float* d_out,
                                      float* d_ext /*[n_tac][96][4] or null*/, cudaStream_t st) {
    const int h = n_stored / 2;
    const int m = 2 * n_chains;
    const int L = m * h;
    if (h < 4 || m > 2048) return (int)cudaErrorInvalidValue;   // up to 1024 chains per TAC (BASELINE configs[3])
    const size_t max_elems = (size_t)1 << 27;
    int tacs_per_batch = (int)std::max<size_t>(1, max_elems / ((size_t)96 * L));
    tacs_per_batch = std::min(tacs_per_batch, n_tac_total);
    const size_t nseg_b = (size_t)tacs_per_batch * 96, nmax = nseg_b * L;
    float *xs = nullptr, *z = nullptr, *zf = nullptr, *med = nullptr, *q05 = nullptr, *q95 = nullptr, *rho = nullptr, *ess = nullptr,
          *rhat = nullptr, *hlo = nullptr, *hhi = nullptr;
    unsigned long long *k0 = nullptr, *k1 = nullptr;
    uint32_t *v0 = nullptr, *v1 = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = cudaSuccess;
#define RD(call) do { e = (call); if (e != cudaSuccess) goto done; } while (0)
    // stream-ordered allocations: the handle keeps the pool cached, so repeated summaries do not pay cudaMalloc/cudaFree
    RD(cudaMallocAsync(&xs, nmax * 4, st)); RD(cudaMallocAsync(&z, nmax * 4, st)); RD(cudaMallocAsync(&zf, nmax * 4, st));
    RD(cudaMallocAsync(&k0, nmax * 8, st)); RD(cudaMallocAsync(&k1, nmax * 8, st)); RD(cudaMallocAsync(&v0, nmax * 4, st)); RD(cudaMallocAsync(&v1, nmax * 4, st));
    RD(cudaMallocAsync(&med, nseg_b * 4, st)); RD(cudaMallocAsync(&q05, nseg_b * 4, st)); RD(cudaMallocAsync(&q95, nseg_b * 4, st));
    RD(cudaMallocAsync(&rho, nseg_b * N_SERIES * (size_t)h * 4, st)); RD(cudaMallocAsync(&ess, nseg_b * N_SERIES * 4, st)); RD(cudaMallocAsync(&rhat, nseg_b * 4, st));
    RD(cudaMallocAsync(&hlo, nseg_b * 4, st)); RD(cudaMallocAsync(&hhi, nseg_b * 4, st));
    {
        int seg_bits = 1;
        while (((size_t)1 << seg_bits) < nseg_b) seg_bits++;
        RD(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, v0, v1, nmax, 0, 32 + seg_bits, st));
        RD(cudaMallocAsync(&tmp, tmp_bytes, st));
        for (int t0 = 0; t0 < n_tac_total; t0 += tacs_per_batch) {
            RankDiagParams p{d_draws, n_chains, max_draws, n_stored, t0, std::min(tacs_per_batch, n_tac_total - t0), h, L};
            const int nseg = p.n_tac * 96;
            const size_t n = (size_t)nseg * L;
            const unsigned gb = (unsigned)((n + 255) / 256);
            rd_build_keys<<<gb, 256, 0, st>>>(p, 0, nullptr, xs, k0, v0, n);
            RD(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, n, 0, 32 + seg_bits, st));
            rd_ranks<<<gb, 256, 0, st>>>(k1, v1, z, L, n);
            rd_quantiles<<<(nseg + 127) / 128, 128, 0, st>>>(k1, L, nseg, med, q05, q95);
            rd_hdi<<<nseg, 256, 0, st>>>(k1, L, 0.94, hlo, hhi);
            rd_build_keys<<<gb, 256, 0, st>>>(p, 1, med, xs, k0, v0, n);
            RD(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, n, 0, 32 + seg_bits, st));
            rd_ranks<<<gb, 256, 0, st>>>(k1, v1, zf, L, n);
            rd_rhat<<<nseg, 256, 0, st>>>(z, zf, m, h, rhat);
            {
                const size_t per_chain = (size_t)h * sizeof(float);
                const int ng = 4 * per_chain <= (size_t)ESS_STAGE_BYTES ? 4 : (2 * per_chain <= (size_t)ESS_STAGE_BYTES ? 2 : 1);
                const int staged = (size_t)ng * per_chain <= (size_t)ESS_STAGE_BYTES ? 1 : 0;
                const size_t dyn = staged ? (size_t)ng * per_chain : 0;
                if (dyn > 16 * 1024) RD(cudaFuncSetAttribute(rd_ess, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
                rd_ess<<<nseg * N_SERIES, ESS_LB, dyn, st>>>(xs, z, q05, q95, m, h, rho, ess, staged, staged ? ng : 4);
            }
            rd_finalize<<<nseg, 256, 0, st>>>(p, ess, rhat, hlo, hhi, nacc, scale, n_draw_sweeps, d_out, d_ext);
            RD(cudaGetLastError());
        }
        RD(cudaStreamSynchronize(st));
    }
done:
#undef RD
    {
        void* bufs[] = {xs, z, zf, k0, k1, v0, v1, med, q05, q95, rho, ess, rhat, hlo, hhi, tmp};
        for (void* b : bufs) if (b) cudaFreeAsync(b, st);
        cudaStreamSynchronize(st);
    }
    return (int)e;
}

// ---- posterior covariance / correlation across ROIs (consumer side: main_script.py:717-738) ----------------------
// np.cov(X, rowvar=False) (ddof = 1) and np.corrcoef(X, rowvar=False) of X = the pooled stored draws of one parameter block,
// (chains * draws, 48): one CTA per (TAC, block).  Draws are centred on the fp64 column means, the products of a tile of
// PC_TILE draws are summed with fp32 FMAs and the tiles in fp64 (as rd_ess does).
constexpr int PC_TILE = 64;
__global__ void __launch_bounds__(256) posterior_cov_kernel(const float* draws, int n_chains, int max_draws, int n_stored,
                                                            double* cov /*[n_tac][2][48][48] or null*/, double* corr /*same*/) {
    __shared__ double s_mean[48];
    __shared__ double s_part[5][48];
    __shared__ float s_x[PC_TILE][48];
    __shared__ double s_cov[48 * 48];
    const int tac = blockIdx.x >> 1, b = blockIdx.x & 1, tid = threadIdx.x;
    const size_t n = (size_t)n_chains * n_stored;
    auto row = [&](size_t k) {   // pooled draw k of this TAC: chain k / n_stored, stored slot k % n_stored
        const size_t c = k / n_stored, d = k - c * n_stored;
        return draws + (((size_t)tac * n_chains + c) * max_draws + d) * 96 + b * 48;
    };
    {   // column means: 5 groups of 48 threads stride over the draws
        const int i = tid % 48, g = tid / 48;
        if (g < 5) {
            double s = 0.0;
            for (size_t k = g; k < n; k += 5) s += (double)row(k)[i];
            s_part[g][i] = s;
        }
        __syncthreads();
        if (tid < 48) s_mean[tid] = ((s_part[0][tid] + s_part[1][tid]) + (s_part[2][tid] + s_part[3][tid]) + s_part[4][tid]) / (double)n;
        __syncthreads();
    }
    double acc[9];               // 48 * 48 = 9 * 256 matrix entries: entry q = tid + 256 p
#pragma unroll
    for (int p = 0; p < 9; p++) acc[p] = 0.0;
    for (size_t k0 = 0; k0 < n; k0 += PC_TILE) {
        const int nt = (int)(n - k0 < (size_t)PC_TILE ? n - k0 : (size_t)PC_TILE);
        for (int e = tid; e < nt * 48; e += 256) {
            const int d = e / 48, i = e - d * 48;
            s_x[d][i] = (float)((double)row(k0 + d)[i] - s_mean[i]);
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 9; p++) {
            const int q = tid + 256 * p, i = q / 48, j = q - i * 48;
            float a = 0.f;
            for (int d = 0; d < nt; d++) a = fmaf(s_x[d][i], s_x[d][j], a);
            acc[p] += (double)a;
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 9; p++) s_cov[tid + 256 * p] = acc[p] / (double)(n - 1);
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 9; p++) {
        const int q = tid + 256 * p, i = q / 48, j = q - i * 48;
        const size_t o = (size_t)blockIdx.x * (48 * 48) + q;
        if (cov) cov[o] = s_cov[q];
        if (corr) corr[o] = fmin(1.0, fmax(-1.0, s_cov[q] / sqrt(s_cov[i * 48 + i] * s_cov[j * 48 + j])));   // np.corrcoef clips too
    }
}

// ---- TFP-style cross-chain ESS (consumer side: main_script.py:807-810) ----------------------
// tfp.mcmc.effective_sample_size(x, cross_chain_dims) with its defaults, restated (PARITY UNPINNED:
// tensorflow_probability is third-party and absent; checked against oracle/diagnostics.py
// tfp_ess_cross_chain): unsplit chains, raw values, per-chain auto-covariance normalised by 1/(N-k),
// rho_k = 1 - (W - mean_c acov_k) / (W + B/N), lags from the first rho_k < 0 on dropped,
// ESS = C N / (-1 + 2 sum_k (N-k)/N rho_k).
__global__ void tfp_gather(const float* draws, int n_chains, int max_draws, int n_stored, int tac0, float* xc, size_t n) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const size_t per_seg = (size_t)n_chains * n_stored;
    const int seg = (int)(e / per_seg);
    const size_t k = e - (size_t)seg * per_seg;
    const int c = (int)(k / n_stored), d = (int)(k - (size_t)c * n_stored);
    const int tac = tac0 + seg / 96, coord = seg % 96;
    xc[e] = draws[(((size_t)tac * n_chains + c) * max_draws + d) * 96 + coord];
}

constexpr int TFP_LB = 256;   // lags per batch
__global__ void __launch_bounds__(TFP_LB) tfp_ess(float* xc /*[nseg][C][N], centred in place*/, int C, int N, float* out) {
    __shared__ double sh[32];
    __shared__ int s_stop;
    __shared__ double s_sum;
    const int seg = blockIdx.x, tid = threadIdx.x;
    float* x = xc + (size_t)seg * C * N;
    double w_biased = 0.0, sm = 0.0, sm2 = 0.0;
    for (int c = 0; c < C; c++) {
        float* xr = x + (size_t)c * N;
        double s = 0;
        for (int i = tid; i < N; i += blockDim.x) s += xr[i];
        const double mean = block_sum(s, sh) / N;
        double q = 0;
        for (int i = tid; i < N; i += blockDim.x) { const float d = (float)((double)xr[i] - mean); xr[i] = d; q += (double)d * d; }
        w_biased += block_sum(q, sh) / N;
        sm += mean; sm2 += mean * mean;
    }
    __syncthreads();
    w_biased /= C;
    const double b_div_n = C > 1 ? (sm2 - sm * sm / C) / (C - 1) : 0.0;
    const double approx = w_biased + b_div_n;
    if (!(w_biased > 0.0) || N < 2) {
        if (tid == 0) out[seg] = CUDART_NAN_F;
        return;
    }
    if (tid == 0) { s_stop = 0; s_sum = 0.0; }
    __syncthreads();
    for (int k0 = 0; k0 < N; k0 += TFP_LB) {
        const int k = k0 + tid;
        double rho = 0.0;
        if (k < N) {
            double acov = 0.0;
            for (int c = 0; c < C; c++) {
                const float* xr = x + (size_t)c * N;
                double a = 0.0;
                for (int n = 0; n + k < N; n++) a += (double)xr[n] * (double)xr[n + k];
                acov += a / (N - k);
            }
            acov /= C;
            rho = C > 1 ? 1.0 - (w_biased - acov) / approx : acov / w_biased;
        }
        // first negative lag of the batch (lags beyond N count as negative)
        const unsigned neg = __ballot_sync(0xffffffffu, !(k < N) || rho < 0.0);
        __shared__ int first_neg[TFP_LB / 32];
        if ((tid & 31) == 0) first_neg[tid >> 5] = neg ? (tid + __ffs(neg) - 1) : TFP_LB;
        __syncthreads();
        int fn = TFP_LB;
        for (int w = 0; w < TFP_LB / 32; w++) fn = min(fn, first_neg[w]);
        const double term = tid < fn ? (double)(N - k) / N * rho : 0.0;
        const double bs = block_sum(term, sh);
        if (tid == 0) { s_sum += bs; if (fn < TFP_LB) s_stop = 1; }
        __syncthreads();
        if (s_stop) break;
    }
    if (tid == 0) out[seg] = (float)((double)C * N / (-1.0 + 2.0 * s_sum));
}

// Host driver: out[n_tac][96] (device).  Returns a cudaError_t (0 = ok).
static inline int launch_tfp_ess(const float* d_draws, int n_tac_total, int n_chains, int max_draws, int n_stored,
                                 float* d_out, cudaStream_t st) {
    if (n_stored < 2) return (int)cudaErrorInvalidValue;
    const size_t per_seg = (size_t)n_chains * n_stored;
    const size_t max_elems = (size_t)1 << 28;
    int tacs_per_batch = (int)std::max<size_t>(1, max_elems / ((size_t)96 * per_seg));
    tacs_per_batch = std::min(tacs_per_batch, n_tac_total);
    float* xc = nullptr;
    cudaError_t e = cudaMallocAsync(&xc, (size_t)tacs_per_batch * 96 * per_seg * 4, st);
    if (e != cudaSuccess) return (int)e;
    for (int t0 = 0; t0 < n_tac_total; t0 += tacs_per_batch) {
        const int nt = std::min(tacs_per_batch, n_tac_total - t0);
        const size_t n = (size_t)nt * 96 * per_seg;
        tfp_gather<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_draws, n_chains, max_draws, n_stored, t0, xc, n);
        tfp_ess<<<nt * 96, TFP_LB, 0, st>>>(xc, n_chains, n_stored, d_out + (size_t)t0 * 96);
    }
    e = cudaGetLastError();
    cudaFreeAsync(xc, st);
    const cudaError_t e2 = cudaStreamSynchronize(st);
    return (int)(e != cudaSuccess ? e : e2);
}

}  // namespace petmh
