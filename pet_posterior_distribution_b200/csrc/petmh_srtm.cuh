// petmh_srtm.cuh -- SURVEY.md 8 f3: the k2-free SRTM (kinetic_model.py:62-84, SRTM.forward_model(DVR, k2, R1, tac_ref))
// as a SAMPLED model: element-wise Metropolis over three 48-coordinate blocks (DVR, R1, k2) with the reference's
// likelihood block (mcmc.py:151-155) and MvNormal priors on all three (mcmc.py:148-149 for DVR / R1; the k2 prior is
// supplied by the caller: the reference ships none and never samples k2, so this goes beyond mcmc.py -- same sampler
// semantics as pm.Metropolis / CompoundStep, checked against oracle/srtm3.py).
//
// Not a hot path: clarity over speed.  One 64-thread CTA per chain, thread i < 48 owns ROI i; forward model and
// likelihood through the production routines (load_tac_image + the exact-operator exact_block); the prior coupling is
// resolved by a plain serial scan in visit order (48 steps per block sweep), r = P (q - mu) in fp64.
#pragma once
#include "petmh_device.cuh"

namespace petmh {

struct SrtmParams {
    const double* P3;        // [3][48][48] prior precisions: DVR, R1, k2
    const double* mu3;       // [3][48]
    float* q;                // [n_tac * n_chains][3][48] state (in / out)
    float* scale;            // same
    int* cnt;                // same: accepted moves since the last tune
    unsigned* nacc;          // same: accepted moves in draw sweeps
    float* draws;            // [n_tac * n_chains][n_out][3][48]
    int n_out, n_chains, sweep0, n_sweeps, tune_until, thin;
    unsigned long long seed, tac_gid0;
    // taped / debug mode (null otherwise): every sweep is recorded
    const float* tape_n;     // [n_chains][n_sweeps_total][3][48]
    const float* tape_logu;
    const uint8_t* tape_rank;
    float* dbg_delta;
    uint8_t* dbg_accept;
    int tape_tac, tape_sweeps;
};

__device__ __forceinline__ float srtm_loglik(const int roi, const float dvr, const float r1, const float k2) {
    extern __shared__ __align__(16) unsigned char smem[];
    float v = 0.f;
#pragma unroll 1
    for (int blk = 0; blk < NBLK; blk++) v += exact_block<false, true>(roi, dvr, r1, blk, nullptr, false, k2);
    return (smem + SM_BAD)[roi] ? -INFINITY : v;
}

__global__ void __launch_bounds__(64) srtm_sweep_kernel(const SweepParams p, const SrtmParams sp) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ float sq[3][48];
    __shared__ uint32_t skey[48];
    __shared__ int sorder[48];
    __shared__ int s_acc;
    __shared__ double s_d;
    const int tid = threadIdx.x;
    const bool taped = sp.tape_n != nullptr;
    const int tac = taped ? sp.tape_tac : (int)(blockIdx.x / sp.n_chains);
    const int chain = taped ? (int)blockIdx.x : (int)(blockIdx.x % sp.n_chains);
    const size_t cg = (taped ? 0 : (size_t)tac * sp.n_chains) + chain;
    const unsigned long long gid = (sp.tac_gid0 + (unsigned long long)tac) * (unsigned long long)sp.n_chains + chain;
    const bool act = tid < 48;
    const int roi = act ? tid : 47;                       // (both warps run every warp-wide vote of the likelihood code)
    load_tac_image(p, tac, smem, tid, 64);
    float q[3], sc[3];
    int cn[3];
    unsigned na[3] = {0u, 0u, 0u};
    double r[3];
#pragma unroll
    for (int b = 0; b < 3; b++) {
        const size_t o = cg * 144 + b * 48 + roi;
        q[b] = sp.q[o]; sc[b] = sp.scale[o]; cn[b] = sp.cnt[o];
        if (act) sq[b][roi] = q[b];
    }
    __syncthreads();
#pragma unroll 1
    for (int b = 0; b < 3; b++) {
        double acc = 0.0;
        const double* Pr = sp.P3 + ((size_t)b * 48 + roi) * 48;
        for (int j = 0; j < 48; j++) acc = fma(Pr[j], (double)sq[b][j] - sp.mu3[b * 48 + j], acc);
        r[b] = acc;
    }
    float ll_old = srtm_loglik(roi, q[0], q[1], q[2]);
#pragma unroll 1
    for (int it = 0; it < sp.n_sweeps; it++) {
        const int sweep = sp.sweep0 + it;
        const bool tuning = sweep < sp.tune_until;
#pragma unroll 1
        for (int b = 0; b < 3; b++) {
            if (tuning && sweep > 0 && (sweep % TUNE_INTERVAL) == 0) {
                sc[b] = __fmul_rn(sc[b], tune_factor(cn[b]));
                cn[b] = 0;
            }
            float nrm, logu;
            uint32_t key;
            if (taped) {
                const size_t o = (((size_t)chain * sp.tape_sweeps + sweep) * 3 + b) * 48 + roi;
                nrm = sp.tape_n[o]; logu = sp.tape_logu[o];
                key = (((uint32_t)sp.tape_rank[o] + 1u) << 6) | (uint32_t)roi;
            } else {
                const int ctr = 3 * sweep + b;            // draw_randoms counts 2 * sweep + block: feed it this kernel's 3 * sweep + b
                draw_randoms(sp.seed, gid, ctr >> 1, ctr & 1, roi, nrm, logu, key);
            }
            const float qo = q[b], qn = __fadd_rn(qo, __fmul_rn(nrm, sc[b]));   // q' = fl32(q + fl32(n * scale))
            const float ll_new = srtm_loglik(roi, b == 0 ? qn : q[0], b == 1 ? qn : q[1], b == 2 ? qn : q[2]);
            const double d = (double)qn - (double)qo;
            const float dll = ll_new - ll_old;
            const double pre = ((double)logu - (double)dll) + 0.5 * d * d * sp.P3[((size_t)b * 48 + roi) * 48 + roi];
            const bool finite = fabsf(dll) <= 3.0e38f;                           // metrop_select's isfinite guard
            if (act) skey[roi] = key;
            __syncthreads();
            if (act) {                                                           // visit position = number of smaller keys
                int rank = 0;
                for (int j = 0; j < 48; j++) rank += skey[j] < key;
                sorder[rank] = roi;
            }
            __syncthreads();
#pragma unroll 1
            for (int v = 0; v < 48; v++) {                                       // the reference's sequential scan, verbatim
                const int w = sorder[v];
                if (tid == w) {
                    const double t = fma(d, r[b], pre);                          // log u - Delta
                    const bool acc = finite && t < 0.0;
                    s_acc = acc ? 1 : 0;
                    s_d = d;
                    if (taped && sp.dbg_delta) {
                        const size_t o = (((size_t)chain * sp.tape_sweeps + sweep) * 3 + b) * 48 + roi;
                        sp.dbg_delta[o] = finite ? (float)((double)logu - t) : CUDART_NAN_F;
                        sp.dbg_accept[o] = acc ? 1 : 0;
                    }
                    if (acc) {
                        q[b] = qn; ll_old = ll_new; cn[b]++;
                        if (!tuning) na[b]++;
                    }
                }
                __syncthreads();
                if (s_acc) r[b] = fma(sp.P3[((size_t)b * 48 + roi) * 48 + w], s_d, r[b]);   // P is symmetric
                __syncthreads();
            }
        }
        // ---- record ----
        const bool rec_all = taped;
        const unsigned di = tuning ? 0u : (unsigned)(sweep - sp.tune_until), slot = rec_all ? (unsigned)sweep : di / (unsigned)sp.thin;
        if (act && sp.draws && (rec_all || (!tuning && slot * (unsigned)sp.thin == di)) && slot < (unsigned)sp.n_out) {
            float* dst = sp.draws + (cg * sp.n_out + slot) * 144 + roi;
            dst[0] = q[0]; dst[48] = q[1]; dst[96] = q[2];
        }
    }
    if (act) {
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const size_t o = cg * 144 + b * 48 + roi;
            sp.q[o] = q[b]; sp.scale[o] = sc[b]; sp.cnt[o] = cn[b]; sp.nacc[o] += na[b];
        }
    }
}

__global__ void srtm_init_kernel(float* q, float* scale, int* cnt, unsigned* nacc, const double* mu3, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { q[i] = (float)mu3[i % 144]; scale[i] = 1.0f; cnt[i] = 0; nacc[i] = 0u; }
}

}  // namespace petmh
