// petmh_synth.cuh -- K4: batched synthetic-data generator on the GPU, the step
// immediately before the hot path.  Restates sample_sim_data.py:128-215 + helper_func.py:146-162: per TAC draw
// DVR, R1 (48) and the reference TAC (54) from the prior MvNormals with positivity rejection -- and, for the
// test-style set (flag_testing_data = True, :128-133), the Mahalanobis rule chi2.cdf(d^2, 48) < alpha --,
// forward-simulate through the production routine (eval3), redraw the triple while any clean TAC value is negative
// (sample_sim_data.py:171-188), add the signal-dependent truncated-Gaussian noise (:205-215), and leave
// y = noisy concentration, c_r and k2p in the handle's input buffers -- no host round trip.
// Distributional, not bitwise, parity with the reference (it uses numpy's global generator).
#pragma once
#include "petmh_device.cuh"

namespace petmh {

struct SynthParams {
    const double* AT[3];     // factors of Cov (transposed: AT[k][j]), A A^T = Cov: DVR, R1, tac_ref
    const double* mu3[3];
    int dim[3], rank[3];
    const float* sigma;      // [48][54] sigma_noise
    float k2p;
    unsigned long long seed, tac_gid0;
    const unsigned long long* tac_gids;   // optional explicit global TAC ids
    int n_tac;
    int* n_capped;           // number of TACs that hit a rejection cap (their data is NOT a valid draw)
    // test-style rule (sample_sim_data.py:128-133), off when d2_max <= 0: a drawn vector x of variable v is kept only if
    // 0 <= (x - mu)^T cinv[v] (x - mu) < d2_max = chi2.ppf(alpha, 48) (a negative form is NaN in scipy's mahalanobis and
    // fails the reference's comparison as well)
    const double* cinv[3];   // caller-supplied inverses [dim][dim] (the reference: np.linalg.inv(Cov))
    double d2_max;
    // outputs
    float* y;                // [S][48][54]
    double* cref;            // [S][54]
    float* k2p_out;          // [S]
    float* truth;            // [S][96] DVR, R1 (may be null)
    float* clean;            // [S][48][54] clean TAC, concentration units (may be null)
    int* attempts;           // [S] number of (DVR, R1, tac_ref) triples drawn
};

__device__ __forceinline__ float2 philox_normal2(unsigned long long seed, unsigned long long gid, uint32_t a, uint32_t b) {
    const uint4 x = philox4x32_10(make_uint4(a, b, (uint32_t)gid, (uint32_t)(gid >> 32)),
                                  make_uint2((uint32_t)seed ^ 0x5eed5eedu, (uint32_t)(seed >> 32)));
    const float r = sqrtf(-2.f * logf(u01(x.x)));
    float sn, cs;
    sincosf(6.283185307179586f * u01(x.y), &sn, &cs);
    return make_float2(r * cs, r * sn);
}

// blockDim.x = 64, one CTA per TAC
__global__ void __launch_bounds__(64) synth_kernel(const SweepParams p, const SynthParams sp) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double z[64];
    __shared__ double xv[3][64];
    __shared__ double red[2];
    const int tid = threadIdx.x, tac = blockIdx.x;
    const unsigned long long gid = sp.tac_gids ? sp.tac_gids[tac] : sp.tac_gid0 + tac;
    bool capped = false;
    float* scratch = reinterpret_cast<float*>(smem + SM_STATE);   // [32 lanes][3][54] clean TAC by lane
    int attempt = 0;
    while (true) {
        // ---- the three positive MvNormal draws (helper_func.truncnormal_samples) ----
        for (int v = 0; v < 3; v++) {
            int tries = 0;
            while (true) {
                const float2 n2 = philox_normal2(sp.seed, gid, (uint32_t)(tid >> 1), (uint32_t)(((attempt * 4 + v) << 12) + tries));
                if (tid < 64) z[tid] = (tid & 1) ? n2.y : n2.x;
                __syncthreads();
                double acc = 0.0;
                if (tid < sp.dim[v]) {
                    acc = sp.mu3[v][tid];
                    for (int k = 0; k < sp.rank[v]; k++) acc = fma(sp.AT[v][k * sp.dim[v] + tid], z[k], acc);
                    xv[v][tid] = acc;
                }
                bool bad = __syncthreads_or(tid < sp.dim[v] && acc < 0.0);   // (the barrier also publishes xv[v])
                if (!bad && sp.d2_max > 0.0) {                             // CTA-uniform: the Mahalanobis rule of the test set
                    double part = 0.0;
                    if (tid < sp.dim[v]) {
                        const double* ci = sp.cinv[v] + (size_t)tid * sp.dim[v];
                        double t = 0.0;
                        for (int j = 0; j < sp.dim[v]; j++) t = fma(ci[j], xv[v][j] - sp.mu3[v][j], t);
                        part = (acc - sp.mu3[v][tid]) * t;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    if ((tid & 31) == 0) red[tid >> 5] = part;
                    __syncthreads();
                    const double d2 = red[0] + red[1];
                    bad = !(d2 >= 0.0 && d2 < sp.d2_max);
                    // (red is rewritten two barriers later at the earliest: no trailing barrier needed)
                }
                tries++;
                if (!bad) break;
                if (tries > 4000) { capped = true; break; }
            }
        }
        if (capped) break;   // (CTA-uniform) a vector could not be drawn positive: give up on this TAC
        // ---- forward simulation through the production routine ----
        if (tid < NT) sp.cref[(size_t)tac * NT + tid] = xv[2][tid];
        if (tid == 0) sp.k2p_out[tac] = sp.k2p;
        for (int i = tid; i < 48 * NT; i += 64) sp.y[(size_t)tac * 48 * NT + i] = 1.0f;   // placeholder observations
        __threadfence_block();
        __syncthreads();
        load_tac_image(p, tac, smem, tid, 64);
        if (tid < 32) {
            const int l16 = tid & 15;
            eval3<0, true>(l16, (float)xv[0][l16], (float)xv[0][l16 + 16], (float)xv[0][l16 + 32], (float)xv[1][l16],
                     (float)xv[1][l16 + 16], (float)xv[1][l16 + 32], scratch + tid * SLOTS * NT);
        }
        __syncthreads();
        int neg = 0;
        for (int i = tid; i < 48 * NT; i += 64) {
            const int r = i / NT, j = i - r * NT;
            neg |= scratch[((r & 15) * SLOTS + (r >> 4)) * NT + j] < 0.f;
        }
        neg = __syncthreads_or(neg);
        attempt++;
        if (!neg) break;
        if (attempt > 1000) { capped = true; break; }
    }
    // ---- noise: x + sqrt(x) * TruncNormal(0, sigma, low = -sqrt(x))  (sample_sim_data.py:205-215) ----
    for (int i = tid; i < 48 * NT; i += 64) {
        const int r = i / NT, j = i - r * NT;
        const float x = scratch[((r & 15) * SLOTS + (r >> 4)) * NT + j];
        const float sq = sqrtf(fmaxf(x, 0.f)), sg = sp.sigma[i];
        float n = 0.f;
        for (int t = 0; t < 64; t++) {
            const float2 n2 = philox_normal2(sp.seed, gid, (uint32_t)(0x10000 + i), (uint32_t)(0x40000000 + t));
            n = sg * n2.x;
            if (n >= -sq) break;
            n = sg * n2.y;
            if (n >= -sq) break;
            n = 0.f;
        }
        sp.y[(size_t)tac * 48 * NT + i] = fmaf(sq, n, x);
        if (sp.clean) sp.clean[(size_t)tac * 48 * NT + i] = x;
    }
    if (sp.truth && tid < 48) {
        sp.truth[(size_t)tac * 96 + tid] = (float)xv[0][tid];
        sp.truth[(size_t)tac * 96 + 48 + tid] = (float)xv[1][tid];
    }
    // (capped is CTA-uniform: it derives from __syncthreads_or results)
    if (tid == 0) {
        sp.attempts[tac] = capped ? -(attempt + 1) : attempt;   // negative: a rejection cap was hit
        if (capped) atomicAdd(sp.n_capped, 1);
    }
}

}  // namespace petmh
