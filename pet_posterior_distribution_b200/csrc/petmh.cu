// petmh.cu -- host side of libpetmh.so: the C ABI declared in include/petmh.h.
// Device code: petmh_device.cuh (sweep kernel) and petmh_diag.cuh (diagnostics).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/petmh.h"
#include "petmh_device.cuh"
#include "petmh_diag.cuh"
#include "petmh_rankdiag.cuh"
#include "petmh_synth.cuh"
#include "petmh_srtm.cuh"
#include "petmh_conv.cuh"

using namespace petmh;

static std::string g_create_error;
// dynamic shared memory of the one-CTA hook / generator kernels: the TAC image + max(prologue scratch, 32 lanes x 3 x 54 TAC values)
static constexpr int HOOK_SMEM = SM_STATE + (TMP_BYTES > 32 * SLOTS * NT * 4 ? TMP_BYTES : 32 * SLOTS * NT * 4);

struct petmh_handle {
    petmh_cfg cfg{};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    // model
    bool have_frames = false, have_prior = false, have_data = false, have_noise = false;
    FrameTables ft_host{};
    FrameTables* d_ft = nullptr;
    double t[NT]{}, dt[NT]{};
    double mu[2][48]{}, logdet[2]{};
    std::vector<double> P;   // [2][48][48]
    double* d_P = nullptr;
    double* d_mu = nullptr;
    float* d_cc = nullptr;
    double ll_const[48]{};   // sum_t -log(sigma_noise) - log sqrt(2 pi)
    // data
    int n_tac = 0;
    float* d_y = nullptr;
    double* d_cref = nullptr;
    float* d_k2p = nullptr;
    // state
    float *d_q = nullptr, *d_scale = nullptr, *d_mom = nullptr, *d_draws = nullptr;
    float2* d_momw = nullptr;
    float* d_summary = nullptr;   // [max_tacs][96][8], allocated on first petmh_get_summary
    float* d_summary_ext = nullptr;   // [max_tacs][96][4]: hdi_3%, hdi_97%, mcse_sd, ess_sd (stored-draw path)
    int ext_sweep = -1;               // sweep counter the cached extra columns belong to
    // K4 generator
    double cov[2][48 * 48]{};
    double* d_synth_f64 = nullptr;   // factors + means
    float *d_synth_truth = nullptr, *d_synth_clean = nullptr;
    int* d_synth_attempts = nullptr;
    int* d_synth_capped = nullptr;
    double* d_synth_cinv = nullptr;  // test-style rule: caller's inverse covariances [48x48 | 48x48 | 54x54]
    double synth_d2_max = 0.0;       // <= 0: rule off (training-style set)
    // global ids of the Philox streams (petmh_set_global_ids)
    unsigned long long* d_tac_gids = nullptr;
    bool have_tac_gids = false;
    unsigned long long chain_gid0 = 0, chain_stride = 0;   // stride 0 = n_chains
    uint8_t* d_cnt = nullptr;
    uint32_t* d_nacc = nullptr;
    // schedule
    int plan_draws = 0, plan_tune = 0, plan_thin = 1;
    int sweep = 0;
    bool state_ready = false;     // petmh_reset / petmh_set_state called
    int mom_n[2] = {0, 0};        // draws merged per half
    int mom_batches[2] = {0, 0};  // closed batches per half (batch-means ESS of the moments mode)
    // scratch for hooks
    float* d_scratch = nullptr;   // >= 48*54 + 48 + 96 floats
    double* d_scratch64 = nullptr;
    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    int last_launches = 0;
    int launch_sweeps = 200;
    bool launch_sweeps_env = false;   // PETMH_LAUNCH_SWEEPS given: no automatic choice
    int wide = -1;                // -1 auto, 0 never, 1 always three warps per chain pair, 2 always nine (PETMH_WIDE)
};

static int fail(petmh_t* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(h, PETMH_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// stream-ordered device buffer, released on every exit path (the CU() macro returns from the middle of a function)
template <class T>
struct Staged {
    T* p = nullptr;
    cudaStream_t st = nullptr;
    Staged() = default;
    Staged(const Staged&) = delete;
    Staged& operator=(const Staged&) = delete;
    ~Staged() { if (p) cudaFreeAsync(p, st); }
    cudaError_t alloc(size_t n, cudaStream_t s) {
        st = s;
        return cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), s);
    }
    cudaError_t upload(const T* src, size_t n, cudaStream_t s) {
        const cudaError_t e = alloc(n, s);
        return e != cudaSuccess ? e : cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};
using StagedF64 = Staged<double>;

extern "C" int petmh_version(void) { return 101; }

extern "C" const char* petmh_last_error(const petmh_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

// ------------------------------------------------------------------------------------
// frame tables: linear-interpolation weights exactly as interp1d_linear_vec
// (kinetic_model.py:41-49): searchsorted-left index hi, lo = hi-1 (wraps at -1), weight
// |xp[lo]-x| on hi and |xp[hi]-x| on lo, normalised.
// ------------------------------------------------------------------------------------
static void two_tap(const double* xp, int np, double x, int& ia, double& wa, int& ib, double& wb) {
    int hi = 0;
    while (hi < np && xp[hi] < x) hi++;          // searchsorted(side='left')
    int lo = hi - 1;
    if (lo < 0) lo += np;                         // numpy negative index wraps
    if (hi >= np) hi = np - 1;                    // cannot happen for x <= max(xp)
    double w_hi = std::fabs(xp[lo] - x), w_lo = std::fabs(xp[hi] - x);
    if (lo == hi) { w_hi = 1.0; w_lo = 0.0; }
    const double s = w_hi + w_lo;
    ia = lo; wa = w_lo / s;
    ib = hi; wb = w_hi / s;
}

// Modified Bessel function I_d(x), x >= 0 moderate (h t <= ~3): ascending series, fp64.
static double bessel_i(int d, double x) {
    const double hx = 0.5 * x, q = hx * hx;
    double term = 1.0;
    for (int k = 1; k <= d; k++) term *= hx / k;          // (x/2)^d / d!
    double sum = term;
    for (int m = 1; m < 200; m++) {
        term *= q / ((double)m * (double)(m + d));
        sum += term;
        if (term <= 1e-18 * sum) break;
    }
    return sum;
}

static int build_frame_tables(petmh_t* h, const double* t) {
    FrameTables& ft = h->ft_host;
    double tmin = t[0], tmax = t[0];
    for (int i = 0; i < NT; i++) {
        if (i && !(t[i] > t[i - 1])) return fail(h, PETMH_EINVAL, "frame times must be strictly increasing");
        tmin = std::min(tmin, t[i]);
        tmax = std::max(tmax, t[i]);
    }
    double x[NGRID];                              // np.linspace(min, max, 2*54)
    const double step = (tmax - tmin) / (NGRID - 1);
    for (int i = 0; i < NGRID; i++) x[i] = tmin + i * step;
    x[NGRID - 1] = tmax;
    ft.dx = x[1] - x[0];
    for (int i = 0; i < NGRID; i++) two_tap(t, NT, x[i], ft.fa[i], ft.fwa[i], ft.fb[i], ft.fwb[i]);
    for (int j = 0; j < NT; j++) two_tap(x, NGRID, t[j], ft.ba[j], ft.bwa[j], ft.bb[j], ft.bwb[j]);
    for (int f = 0; f < NT; f++) {
        ft.klo[f] = NGRID; ft.khi[f] = -1;
        for (int k = 0; k < NGRID; k++) {
            const bool hit = (ft.fa[k] == f && ft.fwa[k] != 0.0) || (ft.fb[k] == f && ft.fwb[k] != 0.0);
            if (hit) { ft.klo[f] = std::min(ft.klo[f], k); ft.khi[f] = std::max(ft.khi[f], k); }
        }
    }
    // structural pattern of M and comparison with the compiled schedule
    static const int acol_c[NCOL] = PETMH_ACTIVE_COLS;
    static const int nrow_c[NT] = PETMH_NROW_PREFIX;
    static const short pack_c[MPACK] = PETMH_MPACK_SRC;
    bool nz[NT][NT];
    for (int j = 0; j < NT; j++)
        for (int f = 0; f < NT; f++) {
            bool any = false;
            for (int side = 0; side < 2 && !any; side++) {
                const int i = side ? ft.bb[j] : ft.ba[j];
                const double w = side ? ft.bwb[j] : ft.bwa[j];
                if (w == 0.0) continue;
                if (ft.klo[f] <= std::min(ft.khi[f], i)) any = true;
            }
            nz[j][f] = any;
        }
    for (int j = 0; j < NT; j++) {
        int cnt = 0;
        for (int f = 0; f < NT; f++) cnt += nz[j][f];
        bool ok = cnt == nrow_c[j];
        for (int c = 0; c < nrow_c[j] && ok; c++) ok = nz[j][acol_c[c]];
        if (!ok)
            return fail(h, PETMH_EGRID,
                        "frame grid not supported: sparsity of the convolution operator (row %d) differs from the "
                        "compiled 54-frame schedule (tools/gen_schedule.py)", j);
    }
    for (int c = 0; c < NCOL + 3; c++) ft.tcol[c] = 0.f;
    for (int c = 0; c < NCOL; c++) { ft.acol[c] = acol_c[c]; ft.tcol[c] = (float)t[acol_c[c]]; }
    for (int j = 0; j < NT; j++) ft.nrow[j] = nrow_c[j];
    // Chebyshev table C[f][d] = c_d (-1)^d e^{-kmid t_f} I_d(h t_f) (c_0 = 1, c_d = 2) for the range the kernel's fp32
    // s = k2a * inv_h - C0 maps to [-1, 1]: k2a t_last in [CHEB_KT_LO, CHEB_KT_HI]
    {
        const double h0 = 0.5 * (CHEB_KT_HI - CHEB_KT_LO) / tmax;
        ft.inv_h = (float)(1.0 / h0);
        const double hh = 1.0 / (double)ft.inv_h, kmid = (double)CHEB_C0 * hh;
        for (int c = 0; c < NCOL; c++) {
            const double tf = t[acol_c[c]], e = std::exp(-kmid * tf);
            for (int d = 0; d < NCHMAX; d++) {
                const double v = e * bessel_i(d, hh * tf);
                ft.cheb_c[c][d] = (d == 0 ? 1.0 : 2.0) * ((d & 1) ? -v : v);
            }
        }
    }
    for (int i = 0; i < MPACK; i++) ft.pack_src[i] = pack_c[i];
    return PETMH_OK;
}

// symmetric positive-definite inverse + log-determinant via Cholesky (fp64)
static bool spd_inverse(const double* a, int n, double* inv, double* logdet) {
    std::vector<double> L(n * n, 0.0);
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = a[i * n + j];
            for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
            if (i == j) {
                if (!(s > 0.0)) return false;
                L[i * n + i] = std::sqrt(s);
            } else
                L[i * n + j] = s / L[j * n + j];
        }
    *logdet = 0.0;
    for (int i = 0; i < n; i++) *logdet += 2.0 * std::log(L[i * n + i]);
    std::vector<double> Li(n * n, 0.0);   // L^-1
    for (int c = 0; c < n; c++) {
        for (int i = c; i < n; i++) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; k++) s -= L[i * n + k] * Li[k * n + c];
            Li[i * n + c] = s / L[i * n + i];
        }
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = 0.0;
            for (int k = i; k < n; k++) s += Li[k * n + i] * Li[k * n + j];
            inv[i * n + j] = inv[j * n + i] = s;
        }
    return true;
}

// ------------------------------------------------------------------------------------
extern "C" int petmh_create(const petmh_cfg* cfg, petmh_t** out) {
    petmh_t* h = nullptr;
    if (!cfg || !out) return fail(nullptr, PETMH_EINVAL, "null argument");
    if (cfg->n_chains < 1 || cfg->max_tacs < 1 || cfg->max_draws < 0)
        return fail(nullptr, PETMH_EINVAL, "n_chains and max_tacs must be >= 1, max_draws >= 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, PETMH_ENODEVICE, "no CUDA device: libpetmh has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, PETMH_EINVAL, "device ordinal out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10)
        return fail(nullptr, PETMH_ENODEVICE, "device %d is not sm_100 (Blackwell B200): kernels are built for sm_100a only",
                    cfg->device);
    h = new petmh_handle();
    h->cfg = *cfg;
    if (const char* e = getenv("PETMH_LAUNCH_SWEEPS")) { h->launch_sweeps = std::max(1, atoi(e)); h->launch_sweeps_env = true; }
    if (const char* e = getenv("PETMH_WIDE")) h->wide = std::max(0, std::min(2, atoi(e)));
    auto bail = [&](int code) { petmh_destroy(h); return code; };
#define CUC(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            fail(nullptr, e_ == cudaErrorMemoryAllocation ? PETMH_ENOMEM : PETMH_ECUDA, "%s failed: %s", #call, \
                 cudaGetErrorString(e_));                                                             \
            return bail(e_ == cudaErrorMemoryAllocation ? PETMH_ENOMEM : PETMH_ECUDA);                \
        }                                                                                             \
    } while (0)
    CUC(cudaSetDevice(cfg->device));
    CUC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {   // keep freed staging buffers cached in the stream-ordered pool (a trim at every sync cost ~0.6 s per set_data)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, cfg->device) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    h->own_stream = true;
    CUC(cudaEventCreate(&h->ev0));
    CUC(cudaEventCreate(&h->ev1));
    const size_t S = cfg->max_tacs, NC = S * (size_t)cfg->n_chains;
    CUC(cudaMalloc(&h->d_ft, sizeof(FrameTables)));
    CUC(cudaMalloc(&h->d_P, 2 * 48 * 48 * sizeof(double)));
    CUC(cudaMalloc(&h->d_mu, 2 * 48 * sizeof(double)));
    CUC(cudaMalloc(&h->d_cc, 48 * NT * sizeof(float)));
    CUC(cudaMalloc(&h->d_y, S * 48 * NT * sizeof(float)));
    CUC(cudaMalloc(&h->d_cref, S * NT * sizeof(double)));
    CUC(cudaMalloc(&h->d_k2p, S * sizeof(float)));
    CUC(cudaMalloc(&h->d_q, NC * 96 * sizeof(float)));
    CUC(cudaMalloc(&h->d_scale, NC * 96 * sizeof(float)));
    CUC(cudaMalloc(&h->d_cnt, NC * 96));
    CUC(cudaMalloc(&h->d_nacc, NC * 96 * sizeof(uint32_t)));
    CUC(cudaMalloc(&h->d_mom, NC * 96 * 2 * MOMF * sizeof(float)));
    CUC(cudaMalloc(&h->d_momw, NC * 96 * sizeof(float2)));
    if (cfg->max_draws > 0) CUC(cudaMalloc(&h->d_draws, NC * (size_t)cfg->max_draws * 96 * sizeof(float)));
    CUC(cudaMalloc(&h->d_scratch, (48 * NT + 256) * sizeof(float)));
    CUC(cudaMalloc(&h->d_scratch64, NT * NT * sizeof(double)));
    CUC(cudaFuncSetAttribute(mh_sweep_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(256)));
    CUC(cudaFuncSetAttribute(mh_sweep_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(256)));
    CUC(cudaFuncSetAttribute(mh_sweep_kernel<0, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_wide()));
    CUC(cudaFuncSetAttribute(mh_sweep_kernel<0, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_wide()));
    CUC(cudaFuncSetAttribute(synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HOOK_SMEM));
    CUC(cudaFuncSetAttribute(forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HOOK_SMEM));
    CUC(cudaFuncSetAttribute(cheb_operator_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HOOK_SMEM));
    CUC(cudaFuncSetAttribute(srtm_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HOOK_SMEM));
#undef CUC
    *out = h;
    return PETMH_OK;
}

extern "C" void petmh_destroy(petmh_t* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    void* bufs[] = {h->d_ft, h->d_P, h->d_mu, h->d_cc, h->d_y, h->d_cref, h->d_k2p, h->d_q, h->d_scale,
                    h->d_cnt, h->d_nacc, h->d_mom, h->d_draws, h->d_scratch, h->d_scratch64, h->d_momw, h->d_summary, h->d_summary_ext, h->d_synth_f64, h->d_synth_truth,
                    h->d_synth_clean, h->d_synth_attempts, h->d_synth_capped, h->d_synth_cinv, h->d_tac_gids};
    for (void* b : bufs) if (b) cudaFree(b);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int petmh_set_stream(petmh_t* h, void* s) {
    if (!h) return PETMH_EINVAL;
    if (h->own_stream && h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    h->stream = (cudaStream_t)s;
    h->own_stream = false;
    return PETMH_OK;
}
extern "C" int petmh_synchronize(petmh_t* h) {
    if (!h) return PETMH_EINVAL;
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}
extern "C" int petmh_last_kernel_ms(const petmh_t* h, float* ms, int* launches) {
    if (!h) return PETMH_EINVAL;
    if (ms) *ms = h->last_ms;
    if (launches) *launches = h->last_launches;
    return PETMH_OK;
}

extern "C" int petmh_set_frames(petmh_t* h, const double* t54, const double* dt54) {
    if (!h || !t54 || !dt54) return fail(h, PETMH_EINVAL, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    int rc = build_frame_tables(h, t54);
    if (rc) return rc;
    memcpy(h->t, t54, sizeof h->t);
    memcpy(h->dt, dt54, sizeof h->dt);
    CU(cudaMemcpyAsync(h->d_ft, &h->ft_host, sizeof(FrameTables), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->have_frames = true;
    return PETMH_OK;
}

extern "C" int petmh_set_prior(petmh_t* h, const double* mu_dvr, const double* cov_dvr, const double* mu_r1,
                               const double* cov_r1) {
    if (!h || !mu_dvr || !cov_dvr || !mu_r1 || !cov_r1) return fail(h, PETMH_EINVAL, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    h->P.assign(2 * 48 * 48, 0.0);
    const double* covs[2] = {cov_dvr, cov_r1};
    const double* mus[2] = {mu_dvr, mu_r1};
    for (int b = 0; b < 2; b++) {
        memcpy(h->mu[b], mus[b], 48 * sizeof(double));
        memcpy(h->cov[b], covs[b], 48 * 48 * sizeof(double));
        if (!spd_inverse(covs[b], 48, h->P.data() + b * 48 * 48, &h->logdet[b]))
            return fail(h, PETMH_EINVAL, "prior covariance %d is not positive definite", b);
    }
    CU(cudaMemcpyAsync(h->d_P, h->P.data(), 2 * 48 * 48 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_mu, h->mu, 2 * 48 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->have_prior = true;
    return PETMH_OK;
}

// Supported scale of the inputs (documented in include/petmh.h): |y| and |c_r| at most 1e6 (the reference's data are O(1):
// activity concentration divided by the frame duration).  Checked on the device: one pass over the batch.
static int check_data_range(petmh_t* h, int n_tac) {
    unsigned* d = reinterpret_cast<unsigned*>(h->d_scratch);
    CU(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned), h->stream));
    const size_t ny = (size_t)n_tac * 48 * NT, nc = (size_t)n_tac * NT;
    data_range_kernel<<<(unsigned)std::min<size_t>(1184, (ny + 255) / 256), 256, 0, h->stream>>>(h->d_y, ny, h->d_cref, nc, d);
    CU(cudaGetLastError());
    unsigned r[2];
    CU(cudaMemcpyAsync(r, d, sizeof r, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float my, mc;
    memcpy(&my, &r[0], 4);
    memcpy(&mc, &r[1], 4);
    if (!(my <= 1e6f) || !(mc <= 1e6f))
        return fail(h, PETMH_EINVAL, "input scale not supported: max |y| = %g, max |tac_ref| = %g (must be finite and <= 1e6; the fp32 "
                    "likelihood multiplies four frames' model values before taking a logarithm -- rescale the activity units)", (double)my, (double)mc);
    return PETMH_OK;
}

static int upload_noise(petmh_t* h, const double* sig) {
    std::vector<float> cc(48 * NT);
    for (int r = 0; r < 48; r++) {
        double c = 0.0;
        for (int j = 0; j < NT; j++) {
            const double s = sig[r * NT + j];
            if (!(s > 0.0)) return fail(h, PETMH_EINVAL, "sigma_noise must be > 0");
            cc[r * NT + j] = (float)(1.0 / (s * std::sqrt(2.0)));
            c += -std::log(s) - 0.5 * std::log(2.0 * M_PI);
        }
        h->ll_const[r] = c;
    }
    CU(cudaMemcpyAsync(h->d_cc, cc.data(), cc.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->have_noise = true;
    return PETMH_OK;
}

extern "C" int petmh_set_global_ids(petmh_t* h, int n_tac, const uint64_t* tac_gids, uint64_t chain_gid0,
                                    uint64_t chains_per_tac_global) {
    if (!h) return PETMH_EINVAL;
    if (n_tac < 0 || n_tac > h->cfg.max_tacs) return fail(h, PETMH_EINVAL, "n_tac %d outside [0, max_tacs=%d]", n_tac, h->cfg.max_tacs);
    if (chains_per_tac_global && chain_gid0 + (uint64_t)h->cfg.n_chains > chains_per_tac_global)
        return fail(h, PETMH_EINVAL, "chain_gid0 + n_chains exceeds chains_per_tac_global");
    CU(cudaSetDevice(h->cfg.device));
    if (tac_gids && n_tac > 0) {
        if (!h->d_tac_gids) CU(cudaMalloc(&h->d_tac_gids, (size_t)h->cfg.max_tacs * sizeof(unsigned long long)));
        CU(cudaMemcpyAsync(h->d_tac_gids, tac_gids, (size_t)n_tac * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        h->have_tac_gids = true;
    } else {
        h->have_tac_gids = false;
    }
    h->chain_gid0 = chain_gid0;
    h->chain_stride = chains_per_tac_global;
    return PETMH_OK;
}

extern "C" int petmh_set_data(petmh_t* h, int n_tac, const double* y, const double* tac_ref, const double* k2p,
                              const double* sigma_noise) {
    if (!h || !y || !tac_ref || !k2p) return fail(h, PETMH_EINVAL, "null argument");
    if (n_tac < 1 || n_tac > h->cfg.max_tacs) return fail(h, PETMH_EINVAL, "n_tac %d outside [1, max_tacs=%d]", n_tac, h->cfg.max_tacs);
    CU(cudaSetDevice(h->cfg.device));
    if (sigma_noise) { int rc = upload_noise(h, sigma_noise); if (rc) return rc; }
    if (!h->have_noise) return fail(h, PETMH_EINVAL, "sigma_noise never set");
    const size_t ny = (size_t)n_tac * 48 * NT, nc = (size_t)n_tac * NT, nk = n_tac;
    {
        StagedF64 dy, dc, dk;
        CU(dy.upload(y, ny, h->stream));
        CU(dc.upload(tac_ref, nc, h->stream));
        CU(dk.upload(k2p, nk, h->stream));
        convert_data_kernel<<<(unsigned)((ny + 255) / 256), 256, 0, h->stream>>>(dy.p, dc.p, dk.p, h->d_y, h->d_cref, h->d_k2p, ny, nc, nk);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(h->stream));
    { int rc = check_data_range(h, n_tac); if (rc) { h->have_data = false; return rc; } }
    h->n_tac = n_tac;
    h->have_data = true;
    return PETMH_OK;
}

extern "C" int petmh_set_data_f32(petmh_t* h, int n_tac, const float* y, const float* tac_ref, const float* k2p,
                                  const float* sigma_noise) {
    if (!h || !y || !tac_ref || !k2p) return fail(h, PETMH_EINVAL, "null argument");
    if (n_tac < 1 || n_tac > h->cfg.max_tacs) return fail(h, PETMH_EINVAL, "n_tac %d outside [1, max_tacs=%d]", n_tac, h->cfg.max_tacs);
    CU(cudaSetDevice(h->cfg.device));
    if (sigma_noise) {
        std::vector<double> s(48 * NT);
        for (int i = 0; i < 48 * NT; i++) s[i] = sigma_noise[i];
        int rc = upload_noise(h, s.data());
        if (rc) return rc;
    }
    if (!h->have_noise) return fail(h, PETMH_EINVAL, "sigma_noise never set");
    const size_t ny = (size_t)n_tac * 48 * NT, nc = (size_t)n_tac * NT;
    {
        Staged<float> dc;
        CU(cudaMemcpyAsync(h->d_y, y, ny * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        CU(dc.upload(tac_ref, nc, h->stream));
        CU(cudaMemcpyAsync(h->d_k2p, k2p, (size_t)n_tac * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        convert_data_f32_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, h->stream>>>(dc.p, h->d_cref, nc);
        CU(cudaGetLastError());
    }
    { int rc = check_data_range(h, n_tac); if (rc) { h->have_data = false; return rc; } }
    h->n_tac = n_tac;
    h->have_data = true;
    return PETMH_OK;
}

static int check_ready(petmh_t* h) {
    if (!h) return PETMH_EINVAL;
    if (!h->have_frames) return fail(h, PETMH_EINVAL, "petmh_set_frames not called");
    if (!h->have_prior) return fail(h, PETMH_EINVAL, "petmh_set_prior not called");
    if (!h->have_data) return fail(h, PETMH_EINVAL, "petmh_set_data not called");
    return PETMH_OK;
}

static SweepParams base_params(petmh_t* h) {
    SweepParams p{};
    p.ft = h->d_ft;
    p.P = h->d_P;
    p.mu = h->d_mu;
    p.cc = h->d_cc;
    p.y = h->d_y;
    p.cref = h->d_cref;
    p.k2p = h->d_k2p;
    p.q = h->d_q;
    p.scale = h->d_scale;
    p.cnt = h->d_cnt;
    p.draws = h->d_draws;
    p.mom = h->d_mom;
    p.nacc = h->d_nacc;
    p.momw = h->d_momw;
    p.max_draws = h->cfg.max_draws;
    p.n_tacs = h->n_tac;
    p.n_chains = h->cfg.n_chains;
    p.thin = 1;
    p.seed = h->cfg.seed;
    p.tac_gid0 = h->cfg.tac_gid0;
    p.tac_gids = h->have_tac_gids ? h->d_tac_gids : nullptr;
    p.chain_gid0 = h->chain_gid0;
    p.chain_stride = h->chain_stride ? h->chain_stride : (unsigned long long)h->cfg.n_chains;
    return p;
}

// ---- parity hooks ----------------------------------------------------------------------
static int run_forward(petmh_t* h, int tac, const double* dvr, const double* r1, std::vector<float>& tac_out,
                       std::vector<float>& ll_out) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (tac < 0 || tac >= h->n_tac) return fail(h, PETMH_EINVAL, "tac index out of range");
    CU(cudaSetDevice(h->cfg.device));
    float in[96];
    for (int i = 0; i < 48; i++) { in[i] = (float)dvr[i]; in[48 + i] = (float)r1[i]; }
    float* d_in = h->d_scratch + 48 * NT + 64;
    CU(cudaMemcpyAsync(d_in, in, sizeof in, cudaMemcpyHostToDevice, h->stream));
    SweepParams p = base_params(h);
    forward_kernel<<<1, 64, HOOK_SMEM, h->stream>>>(p, tac, d_in, d_in + 48, h->d_scratch, h->d_scratch + 48 * NT);
    CU(cudaGetLastError());
    tac_out.resize(48 * NT);
    ll_out.resize(48);
    CU(cudaMemcpyAsync(tac_out.data(), h->d_scratch, 48 * NT * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(ll_out.data(), h->d_scratch + 48 * NT, 48 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}

extern "C" int petmh_forward(petmh_t* h, int tac, const double* dvr48, const double* r1_48, double* out48x54) {
    if (!h || !dvr48 || !r1_48 || !out48x54) return fail(h, PETMH_EINVAL, "null argument");
    std::vector<float> t, l;
    int rc = run_forward(h, tac, dvr48, r1_48, t, l);
    if (rc) return rc;
    for (int i = 0; i < 48 * NT; i++) out48x54[i] = t[i];
    return PETMH_OK;
}

extern "C" int petmh_loglik(petmh_t* h, int tac, const double* dvr48, const double* r1_48, double* ll48,
                            double* logprior2) {
    if (!h || !dvr48 || !r1_48 || !ll48) return fail(h, PETMH_EINVAL, "null argument");
    std::vector<float> t, l;
    int rc = run_forward(h, tac, dvr48, r1_48, t, l);
    if (rc) return rc;
    for (int i = 0; i < 48; i++) ll48[i] = (double)l[i] + h->ll_const[i];
    if (logprior2) {
        const double* q[2] = {dvr48, r1_48};
        for (int b = 0; b < 2; b++) {
            double quad = 0.0;
            for (int i = 0; i < 48; i++) {
                double s = 0.0;
                for (int j = 0; j < 48; j++) s += h->P[(b * 48 + i) * 48 + j] * (q[b][j] - h->mu[b][j]);
                quad += (q[b][i] - h->mu[b][i]) * s;
            }
            logprior2[b] = -0.5 * quad - 0.5 * h->logdet[b] - 24.0 * std::log(2.0 * M_PI);
        }
    }
    return PETMH_OK;
}

extern "C" int petmh_forward_srtm(petmh_t* h, int tac, const double* dvr48, const double* k2_48, const double* r1_48,
                                  double* out48x54) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!dvr48 || !k2_48 || !r1_48 || !out48x54) return fail(h, PETMH_EINVAL, "null argument");
    if (tac < 0 || tac >= h->n_tac) return fail(h, PETMH_EINVAL, "tac index out of range");
    CU(cudaSetDevice(h->cfg.device));
    float in[144];
    for (int i = 0; i < 48; i++) { in[i] = (float)dvr48[i]; in[48 + i] = (float)k2_48[i]; in[96 + i] = (float)r1_48[i]; }
    float* d_in = h->d_scratch + 48 * NT + 64;
    CU(cudaMemcpyAsync(d_in, in, sizeof in, cudaMemcpyHostToDevice, h->stream));
    SweepParams p = base_params(h);
    forward_srtm_kernel<<<1, 256, 0, h->stream>>>(p, tac, d_in, d_in + 48, d_in + 96, h->d_scratch);
    CU(cudaGetLastError());
    std::vector<float> t(48 * NT);
    CU(cudaMemcpyAsync(t.data(), h->d_scratch, 48 * NT * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 48 * NT; i++) out48x54[i] = t[i];
    return PETMH_OK;
}

extern "C" int petmh_get_operator(petmh_t* h, int tac, double* m) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!m || tac < 0 || tac >= h->n_tac) return fail(h, PETMH_EINVAL, "bad argument");
    CU(cudaSetDevice(h->cfg.device));
    SweepParams p = base_params(h);
    operator_kernel<<<1, 256, 0, h->stream>>>(p, tac, h->d_scratch64);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(m, h->d_scratch64, NT * NT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}

extern "C" int petmh_get_cheb_operator(petmh_t* h, int tac, float* a, int* ncols3, double* k2a_lo, double* k2a_hi) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (tac < 0 || tac >= h->n_tac) return fail(h, PETMH_EINVAL, "bad argument");
    CU(cudaSetDevice(h->cfg.device));
    if (a) {
        SweepParams p = base_params(h);
        cheb_operator_kernel<<<1, 64, HOOK_SMEM, h->stream>>>(p, tac, h->d_scratch);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(a, h->d_scratch, (NCH0 + NCH1 + NCH2) * RSTRIDE * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    if (ncols3) { ncols3[0] = NCH0; ncols3[1] = NCH1; ncols3[2] = NCH2; }
    // the range the kernel's fp32 s = k2a * inv_h - C0 maps to [-1, 1]
    const double hh = 1.0 / (double)h->ft_host.inv_h, kmid = (double)CHEB_C0 * hh;
    if (k2a_lo) *k2a_lo = kmid - hh;
    if (k2a_hi) *k2a_hi = kmid + hh;
    return PETMH_OK;
}

extern "C" int petmh_philox_raw(petmh_t* h, uint64_t gid, uint32_t sweep, uint32_t block, uint32_t* out) {
    if (!h || !out) return fail(h, PETMH_EINVAL, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    uint32_t* d = reinterpret_cast<uint32_t*>(h->d_scratch);
    philox_kernel<<<1, 64, 0, h->stream>>>(h->cfg.seed, gid, sweep, block, d);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d, 48 * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}

// ---- the module-level helpers of kinetic_model.py on general grids (fp64; petmh_conv.cuh) ---------------------------
namespace {
int col_threads(int m) { return std::min(256, (m + 31) / 32 * 32); }
}  // namespace

extern "C" int petmh_interp1d_linear(petmh_t* h, int nx, const double* x, int np, const double* xp, const double* fp, int m,
                                     double* out) {
    if (!h || !x || !xp || !fp || !out) return fail(h, PETMH_EINVAL, "null argument");
    if (nx < 1 || np < 2 || m < 1) return fail(h, PETMH_EINVAL, "need nx >= 1, np >= 2, m >= 1");
    for (int i = 0; i < nx; i++)
        if (!(x[i] <= xp[np - 1]))
            return fail(h, PETMH_EINVAL, "x[%d] = %g lies beyond xp[-1] = %g (or is NaN): the reference indexes out of bounds there "
                        "(IndexError, kinetic_model.py:46)", i, x[i], xp[np - 1]);
    CU(cudaSetDevice(h->cfg.device));
    StagedF64 dx, dxp, dfp, dout;
    CU(dx.upload(x, nx, h->stream));
    CU(dxp.upload(xp, np, h->stream));
    CU(dfp.upload(fp, (size_t)np * m, h->stream));
    CU(dout.alloc((size_t)nx * m, h->stream));
    interp_rows_kernel<<<nx, col_threads(m), 0, h->stream>>>(dx.p, nx, dxp.p, np, dfp.p, m, dout.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout.p, (size_t)nx * m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}

extern "C" int petmh_continuous_convolution(petmh_t* h, int n, const double* x, const double* y0, const double* y1, int m,
                                            int num_points, double* out) {
    if (!h || !x || !y0 || !y1 || !out) return fail(h, PETMH_EINVAL, "null argument");
    if (n < 2 || m < 1 || num_points < 0 || num_points == 1) return fail(h, PETMH_EINVAL, "need n >= 2, m >= 1, num_points 0 (= 2 n) or >= 2");
    for (int i = 1; i < n; i++)
        if (!(x[i] > x[i - 1])) return fail(h, PETMH_EINVAL, "x must be strictly increasing (x[%d] = %g, x[%d] = %g)", i - 1, x[i - 1], i, x[i]);
    const int N = num_points ? num_points : 2 * n;                  // kinetic_model.py:13-16 (np.unique(x).size = n here)
    CU(cudaSetDevice(h->cfg.device));
    StagedF64 dx, dy0, dy1, dxrs, dy0rs, dy1rs, dconv, dout;
    CU(dx.upload(x, n, h->stream));
    CU(dy0.upload(y0, n, h->stream));
    CU(dy1.upload(y1, (size_t)n * m, h->stream));
    CU(dxrs.alloc(N, h->stream));
    CU(dy0rs.alloc(N, h->stream));
    CU(dy1rs.alloc((size_t)N * m, h->stream));
    CU(dconv.alloc((size_t)N * m, h->stream));
    CU(dout.alloc((size_t)n * m, h->stream));
    const int T = col_threads(m);
    conv_resample_kernel<<<N, T, 0, h->stream>>>(dx.p, n, dy0.p, dy1.p, m, N, dxrs.p, dy0rs.p, dy1rs.p);
    CU(cudaGetLastError());
    conv_discrete_kernel<<<N, T, 0, h->stream>>>(dxrs.p, dy0rs.p, dy1rs.p, m, N, dconv.p);
    CU(cudaGetLastError());
    conv_back_kernel<<<n, T, 0, h->stream>>>(dx.p, n, dxrs.p, N, dconv.p, m, dout.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout.p, (size_t)n * m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}

extern "C" int petmh_time_exponential(petmh_t* h, int np, const double* param, int nt, const double* t, double* out) {
    if (!h || !param || !t || !out) return fail(h, PETMH_EINVAL, "null argument");
    if (np < 1 || nt < 1) return fail(h, PETMH_EINVAL, "need np >= 1, nt >= 1");
    CU(cudaSetDevice(h->cfg.device));
    StagedF64 dp, dt, dout;
    CU(dp.upload(param, np, h->stream));
    CU(dt.upload(t, nt, h->stream));
    const size_t n = (size_t)nt * np;
    CU(dout.alloc(n, h->stream));
    time_exponential_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(dp.p, np, dt.p, nt, dout.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}

// ---- sampling --------------------------------------------------------------------------
extern "C" int petmh_reset(petmh_t* h) {
    if (!h) return PETMH_EINVAL;
    if (!h->have_prior) return fail(h, PETMH_EINVAL, "petmh_set_prior not called");
    CU(cudaSetDevice(h->cfg.device));
    const size_t NC = (size_t)h->cfg.max_tacs * h->cfg.n_chains;
    const size_t n = NC * 96 * 2 * MOMF;
    init_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_q, h->d_scale, h->d_cnt, h->d_nacc, h->d_mom, h->d_mu, NC);
    CU(cudaGetLastError());
    h->sweep = 0;
    h->ext_sweep = -1;
    h->state_ready = true;
    h->mom_n[0] = h->mom_n[1] = 0;
    h->mom_batches[0] = h->mom_batches[1] = 0;
    return PETMH_OK;
}

extern "C" int petmh_plan(petmh_t* h, int draws, int tune, int thin) {
    if (!h) return PETMH_EINVAL;
    if (draws < 0 || tune < 0 || thin < 1) return fail(h, PETMH_EINVAL, "draws/tune must be >= 0 and thin >= 1");
    h->plan_draws = draws;
    h->plan_tune = tune;
    h->plan_thin = thin;
    return PETMH_OK;
}

// batch-means ESS of the moments mode: every split half (draws / 2 draws) holds up to MOM_NBATCH batches of this length
// (the estimator is biased high by ~ tau / B for an autocorrelation time tau: long batches matter more than many)
constexpr int MOM_NBATCH = 8;
static int moments_batch_len(int plan_draws) {
    const int n_half = plan_draws / 2;
    return n_half >= 2 ? std::max(1, n_half / MOM_NBATCH) : 0;
}

static int threads_per_cta(const petmh_t* h) {
    int t = h->cfg.n_chains * 16;
    t = (t + 31) / 32 * 32;
    t = std::min(t, 256);
    // Jobs of less than one wave (2 CTAs x 148 SMs; e.g. 12 TACs x 256 chains, or one TAC x 64 chains when the wide
    // kernels are off): every CTA is resident at once, so the time is that of the busiest SM -- the sweep loop is
    // latency-bound per warp.  Pick the CTA size (whole warps = chain pairs, any multiple of 32 threads) that minimises
    // ceil(CTAs / 148) x warps per CTA while all CTAs still fit one wave; ties go to the smaller CTA (more SMs in use).
    auto groups = [&](int tt) { return (size_t)h->n_tac * ((h->cfg.n_chains * 16 + tt - 1) / tt); };
    if (groups(t) < 2 * 148) {
        auto cost = [&](int tt) { return (long)((groups(tt) + 147) / 148) * tt; };
        int best = t;
        long best_cost = cost(t);
        for (int tt = t - 32; tt >= 32; tt -= 32) {
            if (groups(tt) > 2 * 148) break;
            if (cost(tt) <= best_cost) { best = tt; best_cost = cost(tt); }
        }
        t = best;
    }
    return t;
}

extern "C" int petmh_advance(petmh_t* h, int n_sweeps) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (n_sweeps < 0) return fail(h, PETMH_EINVAL, "n_sweeps < 0");
    if (!h->state_ready) return fail(h, PETMH_EINVAL, "chain state not initialised: call petmh_reset, petmh_run or petmh_set_state first");
    CU(cudaSetDevice(h->cfg.device));
    // small jobs: the wide kernels when even they leave SMs to spare -- nine warps per chain pair (one pair per CTA)
    // while every pair gets its own SM, three warps per pair (up to four pairs per CTA) while one wave holds them all
    const size_t pairs = (size_t)h->n_tac * ((h->cfg.n_chains + 1) / 2);
    int wide = 0;
    if (h->wide > 0) wide = h->wide;
    else if (h->wide < 0) wide = pairs <= 148 ? 2 : (pairs <= (size_t)WIDE_MAX_TRIPLES * 148 ? 1 : 0);
    int nthr, chains_per_cta;
    if (wide == 2) {
        nthr = 288;
        chains_per_cta = 2;
    } else if (wide) {
        int triples = 1;   // per CTA: one unless the CTAs would outnumber the SMs
        while (triples < WIDE_MAX_TRIPLES && (size_t)h->n_tac * ((h->cfg.n_chains + 2 * triples - 1) / (2 * triples)) > 148) triples *= 2;
        nthr = 96 * triples;
        chains_per_cta = 2 * triples;
    } else {
        nthr = threads_per_cta(h);
        chains_per_cta = nthr / 16;
    }
    const int groups = (h->cfg.n_chains + chains_per_cta - 1) / chains_per_cta;
    const unsigned grid = (unsigned)((size_t)h->n_tac * groups);
    // Split halves as ArviZ's _split_chains: the first and the last n_half = draws / 2 draws of a chain (an odd
    // middle draw belongs to neither).  Each half is cut into MOM_NBATCH batches of blen draws for the batch-means ESS
    // of the moments mode; launches never straddle the tune / half / batch boundaries.
    const int n_half = h->plan_draws / 2, half1_at = h->plan_draws - n_half;
    const int blen = moments_batch_len(h->plan_draws);
    // sweeps per launch: 200 keeps a launch of a full GPU near a second; a job of at most one wave that stores its draws runs
    // 1000 (every launch rebuilds the per-TAC operators in its prologue: 1.7 % of a one-posterior job at 200 --
    // profiles/r02_variant_probes.txt section 14).  Chains, states and stored-draw summaries do not depend on the chunking;
    // the fp32 running moments do in their last bits (another merge order), so runs whose summary comes from the moments keep
    // ONE chunking whatever their size: sharded and single-GPU results stay bit-identical there too.
    const bool summary_from_draws =
        h->cfg.max_draws > 0 && std::min(h->cfg.max_draws, (h->plan_draws + h->plan_thin - 1) / h->plan_thin) >= 8;
    const int launch_sweeps = (h->launch_sweeps_env || grid > 2u * 148u || !summary_from_draws) ? h->launch_sweeps : 1000;
    int left = n_sweeps;
    h->last_launches = 0;
    CU(cudaEventRecord(h->ev0, h->stream));
    while (left > 0) {
        int n = std::min(left, launch_sweeps);
        int half = -1, b_len = 0, b_idx = 0, b_end = 0;
        if (h->sweep < h->plan_tune) {
            n = std::min(n, h->plan_tune - h->sweep);
        } else {
            const int di = h->sweep - h->plan_tune;                       // draw index
            int pos = -1;
            if (di < n_half) { half = 0; pos = di; }
            else if (di < half1_at) n = 1;                                // the dropped middle draw
            else if (di < h->plan_draws) { half = 1; pos = di - half1_at; }
            if (half >= 0) {
                b_idx = blen > 0 ? pos / blen : 0;
                if (blen > 0 && b_idx < MOM_NBATCH && (b_idx + 1) * blen <= n_half) {
                    b_len = blen;
                    n = std::min(n, (b_idx + 1) * blen - pos);
                    b_end = pos + n == (b_idx + 1) * blen;
                } else {
                    n = std::min(n, n_half - pos);                        // leftover draws of the half: in no batch
                }
            }
        }
        SweepParams p = base_params(h);
        p.sweep0 = h->sweep;
        p.n_sweeps = n;
        p.tune_until = h->plan_tune;
        p.thin = h->plan_thin;
        p.mom_half = half;
        p.mom_n_before = half >= 0 ? h->mom_n[half] : 0;
        p.batch_len = b_len;
        p.batch_idx = b_idx;
        p.batch_end = b_end;
        if (wide == 2) mh_sweep_kernel<0, false, 2><<<grid, nthr, smem_bytes_wide(), h->stream>>>(p);
        else if (wide) mh_sweep_kernel<0, false, 1><<<grid, nthr, smem_bytes_wide(), h->stream>>>(p);
        else mh_sweep_kernel<0, false><<<grid, nthr, smem_bytes(256), h->stream>>>(p);
        CU(cudaGetLastError());
        if (half >= 0) {
            h->mom_n[half] += n;
            if (b_end) h->mom_batches[half] = b_idx + 1;
        }
        h->sweep += n;
        left -= n;
        h->last_launches++;
    }
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaEventSynchronize(h->ev1));
    CU(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return PETMH_OK;
}

extern "C" int petmh_run(petmh_t* h, int draws, int tune, int thin) {
    int rc = petmh_reset(h);
    if (rc) return rc;
    rc = petmh_plan(h, draws, tune, thin);
    if (rc) return rc;
    return petmh_advance(h, draws + tune);
}

extern "C" int petmh_run_taped(petmh_t* h, int tac, int n_tape_chains, int n_sweeps, int tune, const float* normals,
                               const float* logu, const uint8_t* rank, float* draws_out, float* delta_out,
                               uint8_t* accept_out, float* scale_out) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!normals || !logu || !rank || !draws_out) return fail(h, PETMH_EINVAL, "null argument");
    if (tac < 0 || tac >= h->n_tac || n_tape_chains < 1 || n_sweeps < 1 || tune < 0 || tune > n_sweeps)
        return fail(h, PETMH_EINVAL, "bad taped-run arguments");
    CU(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)n_tape_chains * n_sweeps * 96;
    float *dn = nullptr, *dl = nullptr, *dd = nullptr, *ddelta = nullptr, *dscale = nullptr;
    uint8_t *dr = nullptr, *dacc = nullptr;
    auto body = [&]() -> int {
        CU(cudaMalloc(&dn, n * 4)); CU(cudaMalloc(&dl, n * 4)); CU(cudaMalloc(&dr, n));
        CU(cudaMalloc(&dd, n * 4)); CU(cudaMalloc(&ddelta, n * 4)); CU(cudaMalloc(&dacc, n));
        CU(cudaMalloc(&dscale, (size_t)n_tape_chains * 96 * 4));
        CU(cudaMemcpyAsync(dn, normals, n * 4, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(dl, logu, n * 4, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(dr, rank, n, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemsetAsync(ddelta, 0, n * 4, h->stream));
        CU(cudaMemsetAsync(dacc, 0, n, h->stream));
        SweepParams p = base_params(h);
        p.n_chains = n_tape_chains;
        p.scale = dscale;
        p.draws = nullptr;
        p.sweep0 = 0;
        p.n_sweeps = n_sweeps;
        p.tune_until = tune;
        p.tape_n = dn; p.tape_logu = dl; p.tape_rank = dr;
        p.dbg_draws = dd; p.dbg_delta = ddelta; p.dbg_accept = dacc;
        p.tape_tac = tac; p.tape_sweeps = n_sweeps;
        const int nthr = std::min(256, (n_tape_chains * 16 + 31) / 32 * 32);
        const int cpc = nthr / 16;
        const unsigned grid = (unsigned)((n_tape_chains + cpc - 1) / cpc);
        mh_sweep_kernel<0, true><<<grid, nthr, smem_bytes(256), h->stream>>>(p);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(draws_out, dd, n * 4, cudaMemcpyDeviceToHost, h->stream));
        if (delta_out) CU(cudaMemcpyAsync(delta_out, ddelta, n * 4, cudaMemcpyDeviceToHost, h->stream));
        if (accept_out) CU(cudaMemcpyAsync(accept_out, dacc, n, cudaMemcpyDeviceToHost, h->stream));
        if (scale_out) CU(cudaMemcpyAsync(scale_out, dscale, (size_t)n_tape_chains * 96 * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return PETMH_OK;
    };
    rc = body();
    cudaFree(dn); cudaFree(dl); cudaFree(dr); cudaFree(dd); cudaFree(ddelta); cudaFree(dacc); cudaFree(dscale);   // also on error paths
    return rc;
}

// ---- outputs ---------------------------------------------------------------------------
extern "C" int petmh_n_stored(const petmh_t* h) {
    if (!h || h->cfg.max_draws == 0) return 0;
    const int done = std::max(0, std::min(h->sweep, h->plan_tune + h->plan_draws) - h->plan_tune);
    return std::min(h->cfg.max_draws, (done + h->plan_thin - 1) / h->plan_thin);
}

extern "C" int petmh_get_chains(petmh_t* h, float* dvr, float* r1) {
    if (!h || !dvr || !r1) return fail(h, PETMH_EINVAL, "null argument");
    if (!h->d_draws) return fail(h, PETMH_EINVAL, "handle created with max_draws = 0: no stored chains");
    CU(cudaSetDevice(h->cfg.device));
    const int ns = petmh_n_stored(h);
    const size_t NC = (size_t)h->n_tac * h->cfg.n_chains;
    if (ns == 0) return PETMH_OK;
    // device layout [chain][max_draws][2][48] -> two host arrays [chain][ns][48]
    for (int b = 0; b < 2; b++) {
        float* dst = b ? r1 : dvr;
        if (ns == h->cfg.max_draws) {
            CU(cudaMemcpy2DAsync(dst, 48 * sizeof(float), h->d_draws + b * 48, 96 * sizeof(float), 48 * sizeof(float),
                                 NC * (size_t)ns, cudaMemcpyDeviceToHost, h->stream));
        } else {
            for (size_t c = 0; c < NC; c++)   // skip the unused slots of every chain
                CU(cudaMemcpy2DAsync(dst + c * ns * 48, 48 * sizeof(float),
                                     h->d_draws + c * (size_t)h->cfg.max_draws * 96 + b * 48, 96 * sizeof(float),
                                     48 * sizeof(float), (size_t)ns, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}

extern "C" int petmh_get_state(petmh_t* h, float* q, float* scale) {
    if (!h) return PETMH_EINVAL;
    CU(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)h->n_tac * h->cfg.n_chains * 96;
    if (q) CU(cudaMemcpyAsync(q, h->d_q, n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (scale) CU(cudaMemcpyAsync(scale, h->d_scale, n * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}

extern "C" int petmh_set_state(petmh_t* h, const float* q, const float* scale, int sweep) {
    if (!h || sweep < 0) return fail(h, PETMH_EINVAL, "bad argument");
    if (h->n_tac < 1) return fail(h, PETMH_EINVAL, "petmh_set_data not called");
    CU(cudaSetDevice(h->cfg.device));
    const size_t NCall = (size_t)h->cfg.max_tacs * h->cfg.n_chains;
    const size_t n = (size_t)h->n_tac * h->cfg.n_chains * 96;
    if (q) CU(cudaMemcpyAsync(h->d_q, q, n * 4, cudaMemcpyHostToDevice, h->stream));
    if (scale) CU(cudaMemcpyAsync(h->d_scale, scale, n * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(h->d_cnt, 0, NCall * 96, h->stream));
    CU(cudaMemsetAsync(h->d_nacc, 0, NCall * 96 * sizeof(uint32_t), h->stream));
    CU(cudaMemsetAsync(h->d_mom, 0, NCall * 96 * 2 * MOMF * sizeof(float), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->sweep = sweep;
    h->state_ready = true;
    h->mom_n[0] = h->mom_n[1] = 0;
    h->mom_batches[0] = h->mom_batches[1] = 0;
    return PETMH_OK;
}

// ---- full checkpoint: everything a run needs to continue exactly where it stopped ----------
struct CkptHeader {
    uint32_t magic, version;
    int32_t n_tac, n_chains, max_draws, n_stored, plan_draws, plan_tune, plan_thin, sweep;
    int32_t mom_n[2], mom_batches[2];
    uint64_t seed;
};
static constexpr uint32_t CKPT_MAGIC = 0x50544d48u;   // "HMTP"

extern "C" int64_t petmh_checkpoint_bytes(const petmh_t* h) {
    if (!h || h->n_tac < 1) return 0;
    const int64_t nc = (int64_t)h->n_tac * h->cfg.n_chains, ns = petmh_n_stored(h);
    return (int64_t)sizeof(CkptHeader) + nc * 96 * (4 + 4 + 1 + 4 + 2 * MOMF * 4) + nc * ns * 96 * 4;
}

extern "C" int petmh_get_checkpoint(petmh_t* h, void* buf, int64_t nbytes) {
    if (!h || !buf) return fail(h, PETMH_EINVAL, "null argument");
    if (!h->state_ready || h->n_tac < 1) return fail(h, PETMH_EINVAL, "nothing to checkpoint: no data or chain state");
    if (nbytes < petmh_checkpoint_bytes(h)) return fail(h, PETMH_EINVAL, "buffer too small: need %lld bytes", (long long)petmh_checkpoint_bytes(h));
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    const size_t nc = (size_t)h->n_tac * h->cfg.n_chains, n = nc * 96;
    const int ns = petmh_n_stored(h);
    CkptHeader hd{CKPT_MAGIC, 2, h->n_tac, h->cfg.n_chains, h->cfg.max_draws, ns, h->plan_draws, h->plan_tune, h->plan_thin, h->sweep,
                  {h->mom_n[0], h->mom_n[1]}, {h->mom_batches[0], h->mom_batches[1]}, h->cfg.seed};
    unsigned char* o = static_cast<unsigned char*>(buf);
    memcpy(o, &hd, sizeof hd); o += sizeof hd;
    CU(cudaMemcpyAsync(o, h->d_q, n * 4, cudaMemcpyDeviceToHost, h->stream)); o += n * 4;
    CU(cudaMemcpyAsync(o, h->d_scale, n * 4, cudaMemcpyDeviceToHost, h->stream)); o += n * 4;
    CU(cudaMemcpyAsync(o, h->d_nacc, n * 4, cudaMemcpyDeviceToHost, h->stream)); o += n * 4;
    CU(cudaMemcpyAsync(o, h->d_mom, n * 2 * MOMF * 4, cudaMemcpyDeviceToHost, h->stream)); o += n * 2 * MOMF * 4;
    CU(cudaMemcpyAsync(o, h->d_cnt, n, cudaMemcpyDeviceToHost, h->stream)); o += n;
    if (ns > 0)   // the stored slots of every chain: device [chain][max_draws][96] -> blob [chain][n_stored][96]
        CU(cudaMemcpy2DAsync(o, (size_t)ns * 96 * 4, h->d_draws, (size_t)h->cfg.max_draws * 96 * 4, (size_t)ns * 96 * 4, nc,
                             cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}

extern "C" int petmh_set_checkpoint(petmh_t* h, const void* buf, int64_t nbytes) {
    if (!h || !buf) return fail(h, PETMH_EINVAL, "null argument");
    if (nbytes < (int64_t)sizeof(CkptHeader)) return fail(h, PETMH_ESTATE, "checkpoint truncated");
    CkptHeader hd;
    memcpy(&hd, buf, sizeof hd);
    if (hd.magic != CKPT_MAGIC || hd.version != 2) return fail(h, PETMH_ESTATE, "not a petmh checkpoint (or another version)");
    if (hd.n_tac != h->n_tac || hd.n_chains != h->cfg.n_chains)
        return fail(h, PETMH_ESTATE, "checkpoint holds %d TACs x %d chains, the handle has %d x %d bound (call petmh_set_data first)",
                    hd.n_tac, hd.n_chains, h->n_tac, h->cfg.n_chains);
    if (hd.n_stored > h->cfg.max_draws) return fail(h, PETMH_ESTATE, "checkpoint stores %d draws per chain, max_draws is %d", hd.n_stored, h->cfg.max_draws);
    if (hd.seed != h->cfg.seed) return fail(h, PETMH_ESTATE, "checkpoint was taken with another seed: the continuation would not be the same run");
    const size_t nc = (size_t)hd.n_tac * hd.n_chains, n = nc * 96;
    const int64_t need = (int64_t)sizeof(CkptHeader) + (int64_t)n * (4 + 4 + 1 + 4 + 2 * MOMF * 4) + (int64_t)nc * hd.n_stored * 96 * 4;
    if (nbytes < need) return fail(h, PETMH_ESTATE, "checkpoint truncated: %lld of %lld bytes", (long long)nbytes, (long long)need);
    CU(cudaSetDevice(h->cfg.device));
    const unsigned char* o = static_cast<const unsigned char*>(buf) + sizeof hd;
    CU(cudaMemcpyAsync(h->d_q, o, n * 4, cudaMemcpyHostToDevice, h->stream)); o += n * 4;
    CU(cudaMemcpyAsync(h->d_scale, o, n * 4, cudaMemcpyHostToDevice, h->stream)); o += n * 4;
    CU(cudaMemcpyAsync(h->d_nacc, o, n * 4, cudaMemcpyHostToDevice, h->stream)); o += n * 4;
    CU(cudaMemcpyAsync(h->d_mom, o, n * 2 * MOMF * 4, cudaMemcpyHostToDevice, h->stream)); o += n * 2 * MOMF * 4;
    CU(cudaMemcpyAsync(h->d_cnt, o, n, cudaMemcpyHostToDevice, h->stream)); o += n;
    if (hd.n_stored > 0)
        CU(cudaMemcpy2DAsync(h->d_draws, (size_t)h->cfg.max_draws * 96 * 4, o, (size_t)hd.n_stored * 96 * 4, (size_t)hd.n_stored * 96 * 4, nc,
                             cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->plan_draws = hd.plan_draws; h->plan_tune = hd.plan_tune; h->plan_thin = hd.plan_thin;
    h->sweep = hd.sweep;
    h->mom_n[0] = hd.mom_n[0]; h->mom_n[1] = hd.mom_n[1];
    h->mom_batches[0] = hd.mom_batches[0]; h->mom_batches[1] = hd.mom_batches[1];
    h->state_ready = true;
    return PETMH_OK;
}

extern "C" int petmh_summary_device(petmh_t* h, float* d_out, void* stream) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!d_out) return fail(h, PETMH_EINVAL, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    if (st != h->stream) CU(cudaStreamSynchronize(h->stream));
    const int draw_sweeps = std::max(0, h->sweep - h->plan_tune);   // every non-tuning sweep counts accepted moves
    if (h->d_draws && petmh_n_stored(h) >= 8) {
        // stored draws: rank-normalised split R-hat, bulk/tail ESS, MCSE (ArviZ semantics)
        if (!h->d_summary_ext) CU(cudaMalloc(&h->d_summary_ext, (size_t)h->cfg.max_tacs * 96 * 4 * sizeof(float)));
        h->ext_sweep = -1;
        rc = launch_rank_summary(h->d_draws, h->n_tac, h->cfg.n_chains, h->cfg.max_draws, petmh_n_stored(h), h->d_nacc, h->d_scale,
                                 draw_sweeps, d_out, h->d_summary_ext, st);
        if (rc) return fail(h, PETMH_ECUDA, "rank-diagnostics failed: %s", cudaGetErrorString((cudaError_t)rc));
        h->ext_sweep = h->sweep;
        return PETMH_OK;
    }
    DiagParams dp{};
    dp.mom = h->d_mom;
    dp.mu = h->d_mu;
    dp.nacc = h->d_nacc;
    dp.scale = h->d_scale;
    dp.n_tacs = h->n_tac;
    dp.n_chains = h->cfg.n_chains;
    dp.n_half[0] = h->mom_n[0];
    dp.n_half[1] = h->mom_n[1];
    dp.n_batch[0] = h->mom_batches[0];
    dp.n_batch[1] = h->mom_batches[1];
    dp.batch_len = moments_batch_len(h->plan_draws);
    dp.n_draw_sweeps = draw_sweeps;
    dp.out = d_out;
    rc = launch_summary(dp, st);
    if (rc) return fail(h, PETMH_ECUDA, "summary kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    return PETMH_OK;
}

extern "C" int petmh_get_summary(petmh_t* h, float* out) {
    if (!h || !out) return fail(h, PETMH_EINVAL, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)h->n_tac * 96 * PETMH_N_STATS;
    if (!h->d_summary) CU(cudaMalloc(&h->d_summary, (size_t)h->cfg.max_tacs * 96 * PETMH_N_STATS * sizeof(float)));
    int rc = petmh_summary_device(h, h->d_summary, nullptr);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, h->d_summary, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}

extern "C" int petmh_get_summary_ext(petmh_t* h, float* out) {
    if (!h || !out) return fail(h, PETMH_EINVAL, "null argument");
    if (!h->d_draws || petmh_n_stored(h) < 8)
        return fail(h, PETMH_EINVAL, "hdi / mcse_sd need stored draws (max_draws > 0 and >= 8 draws stored; have %d)", petmh_n_stored(h));
    CU(cudaSetDevice(h->cfg.device));
    if (h->ext_sweep != h->sweep || !h->d_summary_ext) {      // not computed for the current state yet
        if (!h->d_summary) CU(cudaMalloc(&h->d_summary, (size_t)h->cfg.max_tacs * 96 * PETMH_N_STATS * sizeof(float)));
        int rc = petmh_summary_device(h, h->d_summary, nullptr);
        if (rc) return rc;
    }
    CU(cudaMemcpyAsync(out, h->d_summary_ext, (size_t)h->n_tac * 96 * 4 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}

// ---- handle-free summaries of gathered state (chains of one TAC sampled on several GPUs) -------------------------
extern "C" int petmh_summary_from_draws_device(int device, const float* d_draws, int n_tac, int n_chains, int n_stored,
                                               const uint32_t* d_nacc, const float* d_scale, int n_draw_sweeps,
                                               float* d_out8, float* d_ext4, void* stream) {
    if (!d_draws || !d_nacc || !d_scale || !d_out8 || n_tac < 1 || n_chains < 1 || n_stored < 8) {
        g_create_error = "petmh_summary_from_draws_device: bad argument (needs >= 8 stored draws)";
        return PETMH_EINVAL;
    }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return PETMH_ECUDA; }
    const int rc = launch_rank_summary(d_draws, n_tac, n_chains, n_stored, n_stored, d_nacc, d_scale, n_draw_sweeps, d_out8, d_ext4,
                                       (cudaStream_t)stream);
    if (rc) { g_create_error = std::string("rank-diagnostics failed: ") + cudaGetErrorString((cudaError_t)rc); return PETMH_ECUDA; }
    return PETMH_OK;
}

extern "C" int petmh_summary_from_moments_device(int device, const float* d_mom, const double* d_mu96, int n_tac, int n_chains,
                                                 const int* n_half2, const int* n_batch2, int batch_len, const uint32_t* d_nacc,
                                                 const float* d_scale, int n_draw_sweeps, float* d_out8, void* stream) {
    if (!d_mom || !d_mu96 || !n_half2 || !n_batch2 || !d_nacc || !d_scale || !d_out8 || n_tac < 1 || n_chains < 1) {
        g_create_error = "petmh_summary_from_moments_device: bad argument";
        return PETMH_EINVAL;
    }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return PETMH_ECUDA; }
    DiagParams dp{};
    dp.mom = d_mom; dp.mu = d_mu96; dp.nacc = d_nacc; dp.scale = d_scale;
    dp.n_tacs = n_tac; dp.n_chains = n_chains;
    dp.n_half[0] = n_half2[0]; dp.n_half[1] = n_half2[1];
    dp.n_batch[0] = n_batch2[0]; dp.n_batch[1] = n_batch2[1];
    dp.batch_len = batch_len; dp.n_draw_sweeps = n_draw_sweeps; dp.out = d_out8;
    const int rc = launch_summary(dp, (cudaStream_t)stream);
    if (rc) { g_create_error = std::string("summary kernel failed: ") + cudaGetErrorString((cudaError_t)rc); return PETMH_ECUDA; }
    return PETMH_OK;
}

// device pointers and counters of the handle's summary inputs, for gathering them across ranks (device memory owned by
// the handle; valid until it is destroyed)
extern "C" int petmh_export_summary_inputs(petmh_t* h, void** d_draws, void** d_mom, void** d_nacc, void** d_scale, void** d_mu96,
                                           int* counters /*[8]: n_stored, max_draws, n_half0, n_half1, n_batch0, n_batch1, batch_len, n_draw_sweeps*/) {
    if (!h) return PETMH_EINVAL;
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    if (d_draws) *d_draws = h->d_draws;
    if (d_mom) *d_mom = h->d_mom;
    if (d_nacc) *d_nacc = h->d_nacc;
    if (d_scale) *d_scale = h->d_scale;
    if (d_mu96) *d_mu96 = h->d_mu;
    if (counters) {
        counters[0] = petmh_n_stored(h); counters[1] = h->cfg.max_draws;
        counters[2] = h->mom_n[0]; counters[3] = h->mom_n[1];
        counters[4] = h->mom_batches[0]; counters[5] = h->mom_batches[1];
        counters[6] = moments_batch_len(h->plan_draws);
        counters[7] = std::max(0, h->sweep - h->plan_tune);
    }
    return PETMH_OK;
}

extern "C" int petmh_get_ess_cross_chain(petmh_t* h, float* out) {
    if (!h || !out) return fail(h, PETMH_EINVAL, "null argument");
    int rc = check_ready(h);
    if (rc) return rc;
    const int n_stored = petmh_n_stored(h);
    if (h->cfg.max_draws <= 0 || n_stored < 2)
        return fail(h, PETMH_EINVAL, "cross-chain ESS needs stored draws (max_draws > 0, >= 2 draws stored; have %d)", n_stored);
    CU(cudaSetDevice(h->cfg.device));
    Staged<float> d_out;
    CU(d_out.alloc((size_t)h->n_tac * 96, h->stream));
    const int e = launch_tfp_ess(h->d_draws, h->n_tac, h->cfg.n_chains, h->cfg.max_draws, n_stored, d_out.p, h->stream);
    if (e) return fail(h, PETMH_ECUDA, "tfp_ess: %s", cudaGetErrorString((cudaError_t)e));
    CU(cudaMemcpyAsync(out, d_out.p, (size_t)h->n_tac * 96 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}

extern "C" int petmh_get_posterior_cov(petmh_t* h, double* cov, double* corr) {
    if (!h || (!cov && !corr)) return fail(h, PETMH_EINVAL, "null argument");
    int rc = check_ready(h);
    if (rc) return rc;
    const int ns = petmh_n_stored(h);
    if (!h->d_draws || (size_t)h->cfg.n_chains * ns < 2)
        return fail(h, PETMH_EINVAL, "the posterior covariance needs stored draws (max_draws > 0, >= 2 pooled draws; have %d per chain)", ns);
    CU(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)h->n_tac * 2 * 48 * 48;
    StagedF64 dcov, dcorr;
    CU(dcov.alloc(n, h->stream));
    CU(dcorr.alloc(n, h->stream));
    posterior_cov_kernel<<<(unsigned)h->n_tac * 2, 256, 0, h->stream>>>(h->d_draws, h->cfg.n_chains, h->cfg.max_draws, ns, dcov.p, dcorr.p);
    CU(cudaGetLastError());
    if (cov) CU(cudaMemcpyAsync(cov, dcov.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (corr) CU(cudaMemcpyAsync(corr, dcorr.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}

// ---- the k2-free SRTM as a sampled three-block model ------------------------------------------------
extern "C" int petmh_srtm_sample(petmh_t* h, const double* mu_k2, const double* cov_k2, int draws, int tune, int thin,
                                 int n_tape_chains, int tape_tac, const float* tape_normals, const float* tape_logu,
                                 const uint8_t* tape_rank, float* draws_out, float* delta_out, uint8_t* accept_out,
                                 float* accept_rate_out) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!mu_k2 || !cov_k2 || !draws_out || draws < 0 || tune < 0 || thin < 1 || draws + tune < 1)
        return fail(h, PETMH_EINVAL, "bad argument");
    const bool taped = tape_normals != nullptr;
    if (taped && (!tape_logu || !tape_rank || n_tape_chains < 1 || tape_tac < 0 || tape_tac >= h->n_tac))
        return fail(h, PETMH_EINVAL, "bad taped-run arguments");
    CU(cudaSetDevice(h->cfg.device));
    std::vector<double> P3(3 * 48 * 48), mu3(3 * 48);
    memcpy(P3.data(), h->P.data(), 2 * 48 * 48 * sizeof(double));              // DVR, R1 as set by petmh_set_prior
    memcpy(mu3.data(), h->mu, 2 * 48 * sizeof(double));
    memcpy(mu3.data() + 96, mu_k2, 48 * sizeof(double));
    double logdet;
    if (!spd_inverse(cov_k2, 48, P3.data() + 2 * 48 * 48, &logdet)) return fail(h, PETMH_EINVAL, "k2 prior covariance is not positive definite");
    const int n_chains = taped ? n_tape_chains : h->cfg.n_chains;
    const int n_tac = taped ? 1 : h->n_tac;
    const size_t nc = (size_t)n_tac * n_chains, n = nc * 144;
    const int total = draws + tune, n_out = taped ? total : (draws + thin - 1) / thin;
    double *dP = nullptr, *dmu = nullptr;
    float *dq = nullptr, *dsc = nullptr, *dd = nullptr, *dtn = nullptr, *dtl = nullptr, *ddel = nullptr;
    int* dcnt = nullptr;
    unsigned* dna = nullptr;
    uint8_t *dtr = nullptr, *dacc = nullptr;
    auto body = [&]() -> int {
        CU(cudaMalloc(&dP, P3.size() * 8)); CU(cudaMalloc(&dmu, mu3.size() * 8));
        CU(cudaMalloc(&dq, n * 4)); CU(cudaMalloc(&dsc, n * 4)); CU(cudaMalloc(&dcnt, n * 4)); CU(cudaMalloc(&dna, n * 4));
        CU(cudaMalloc(&dd, std::max<size_t>(1, nc * (size_t)n_out * 144) * 4));
        CU(cudaMemcpyAsync(dP, P3.data(), P3.size() * 8, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(dmu, mu3.data(), mu3.size() * 8, cudaMemcpyHostToDevice, h->stream));
        srtm_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(dq, dsc, dcnt, dna, dmu, n);
        CU(cudaGetLastError());
        SrtmParams sp{};
        sp.P3 = dP; sp.mu3 = dmu; sp.q = dq; sp.scale = dsc; sp.cnt = dcnt; sp.nacc = dna; sp.draws = dd;
        sp.n_out = n_out; sp.n_chains = n_chains; sp.tune_until = tune; sp.thin = thin;
        sp.seed = h->cfg.seed; sp.tac_gid0 = h->cfg.tac_gid0;
        if (taped) {
            const size_t nt = (size_t)n_chains * total * 144;
            CU(cudaMalloc(&dtn, nt * 4)); CU(cudaMalloc(&dtl, nt * 4)); CU(cudaMalloc(&dtr, nt));
            CU(cudaMalloc(&ddel, nt * 4)); CU(cudaMalloc(&dacc, nt));
            CU(cudaMemcpyAsync(dtn, tape_normals, nt * 4, cudaMemcpyHostToDevice, h->stream));
            CU(cudaMemcpyAsync(dtl, tape_logu, nt * 4, cudaMemcpyHostToDevice, h->stream));
            CU(cudaMemcpyAsync(dtr, tape_rank, nt, cudaMemcpyHostToDevice, h->stream));
            CU(cudaMemsetAsync(ddel, 0, nt * 4, h->stream));
            CU(cudaMemsetAsync(dacc, 0, nt, h->stream));
            sp.tape_n = dtn; sp.tape_logu = dtl; sp.tape_rank = dtr; sp.dbg_delta = ddel; sp.dbg_accept = dacc;
            sp.tape_tac = tape_tac; sp.tape_sweeps = total;
        }
        SweepParams p = base_params(h);
        for (int s0 = 0; s0 < total; s0 += 500) {                                // (chunked: no launch runs for minutes)
            sp.sweep0 = s0;
            sp.n_sweeps = std::min(500, total - s0);
            srtm_sweep_kernel<<<(unsigned)nc, 64, HOOK_SMEM, h->stream>>>(p, sp);
            CU(cudaGetLastError());
        }
        CU(cudaMemcpyAsync(draws_out, dd, nc * (size_t)n_out * 144 * 4, cudaMemcpyDeviceToHost, h->stream));
        if (taped && delta_out) CU(cudaMemcpyAsync(delta_out, ddel, (size_t)n_chains * total * 144 * 4, cudaMemcpyDeviceToHost, h->stream));
        if (taped && accept_out) CU(cudaMemcpyAsync(accept_out, dacc, (size_t)n_chains * total * 144, cudaMemcpyDeviceToHost, h->stream));
        std::vector<unsigned> na(n);
        CU(cudaMemcpyAsync(na.data(), dna, n * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (accept_rate_out)
            for (size_t i = 0; i < n; i++) accept_rate_out[i] = draws > 0 ? (float)na[i] / (float)draws : 0.f;
        return PETMH_OK;
    };
    rc = body();
    void* bufs[] = {dP, dmu, dq, dsc, dcnt, dna, dd, dtn, dtl, dtr, ddel, dacc};
    for (void* b : bufs) if (b) cudaFree(b);
    return rc;
}

// ---- K4: synthetic data on the GPU ---------------------------------------------------------
// pivoted Cholesky of a symmetric positive SEMI-definite matrix: cov ~= A A^T, A is n x rank
// (numpy.random.multivariate_normal tolerates the rank-deficient Cov_tac_ref through an SVD; any
// factor with A A^T = cov gives the same distribution).  Returns A transposed (AT[k][j]).
static int psd_factor_T(const double* cov, int n, std::vector<double>& AT) {
    std::vector<double> a(cov, cov + n * n), L(n * n, 0.0);
    std::vector<int> piv(n);
    for (int i = 0; i < n; i++) piv[i] = i;
    double dmax0 = 0.0;
    for (int i = 0; i < n; i++) dmax0 = std::max(dmax0, a[i * n + i]);
    int rank = 0;
    for (int k = 0; k < n; k++) {
        int best = k;
        for (int i = k + 1; i < n; i++) if (a[piv[i] * n + piv[i]] > a[piv[best] * n + piv[best]]) best = i;
        std::swap(piv[k], piv[best]);
        const int pk = piv[k];
        const double d = a[pk * n + pk];
        if (!(d > 1e-13 * dmax0)) break;
        const double l = std::sqrt(d);
        L[pk * n + k] = l;
        for (int i = k + 1; i < n; i++) {
            const int pi = piv[i];
            L[pi * n + k] = a[pi * n + pk] / l;
        }
        for (int i = k + 1; i < n; i++)
            for (int j = k + 1; j < n; j++) {
                const int pi = piv[i], pj = piv[j];
                a[pi * n + pj] -= L[pi * n + k] * L[pj * n + k];
            }
        rank++;
    }
    AT.assign((size_t)n * n, 0.0);
    for (int k = 0; k < rank; k++)
        for (int j = 0; j < n; j++) AT[(size_t)k * n + j] = L[j * n + k];
    return rank;
}

extern "C" int petmh_synth_set_test_rule(petmh_t* h, double d2_max, const double* cov_inv_dvr48x48, const double* cov_inv_r1_48x48,
                                         const double* cov_inv_tacref54x54) {
    if (!h) return PETMH_EINVAL;
    if (!(d2_max > 0.0) || !cov_inv_dvr48x48 || !cov_inv_r1_48x48 || !cov_inv_tacref54x54) {   // off: the training-style set
        h->synth_d2_max = 0.0;
        return PETMH_OK;
    }
    if (!std::isfinite(d2_max)) return fail(h, PETMH_EINVAL, "d2_max must be finite");
    CU(cudaSetDevice(h->cfg.device));
    constexpr size_t n48 = 48 * 48, n54 = (size_t)NT * NT;
    if (!h->d_synth_cinv) CU(cudaMalloc(&h->d_synth_cinv, (2 * n48 + n54) * sizeof(double)));
    CU(cudaMemcpyAsync(h->d_synth_cinv, cov_inv_dvr48x48, n48 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_synth_cinv + n48, cov_inv_r1_48x48, n48 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_synth_cinv + 2 * n48, cov_inv_tacref54x54, n54 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->synth_d2_max = d2_max;
    return PETMH_OK;
}

extern "C" int petmh_synth(petmh_t* h, int n_tac, uint64_t seed, const double* mu_tacref54, const double* cov_tacref54x54,
                           double k2p, const double* sigma_noise48x54) {
    if (!h || !mu_tacref54 || !cov_tacref54x54 || !sigma_noise48x54) return fail(h, PETMH_EINVAL, "null argument");
    if (!h->have_frames || !h->have_prior) return fail(h, PETMH_EINVAL, "petmh_set_frames / petmh_set_prior not called");
    if (n_tac < 1 || n_tac > h->cfg.max_tacs) return fail(h, PETMH_EINVAL, "n_tac %d outside [1, max_tacs=%d]", n_tac, h->cfg.max_tacs);
    CU(cudaSetDevice(h->cfg.device));
    int rc = upload_noise(h, sigma_noise48x54);
    if (rc) return rc;
    // factors and means: [AT_dvr 48x48 | AT_r1 48x48 | AT_ref 54x54 | mu_dvr 48 | mu_r1 48 | mu_ref 54]
    std::vector<double> buf, at;
    SynthParams sp{};
    const double* covs[3] = {h->cov[0], h->cov[1], cov_tacref54x54};
    const double* mus[3] = {h->mu[0], h->mu[1], mu_tacref54};
    const int dims[3] = {48, 48, 54};
    size_t off[3], moff[3];
    for (int v = 0; v < 3; v++) {
        sp.dim[v] = dims[v];
        sp.rank[v] = psd_factor_T(covs[v], dims[v], at);
        if (sp.rank[v] < 1) return fail(h, PETMH_EINVAL, "covariance %d has no positive direction", v);
        off[v] = buf.size();
        buf.insert(buf.end(), at.begin(), at.end());
    }
    for (int v = 0; v < 3; v++) { moff[v] = buf.size(); buf.insert(buf.end(), mus[v], mus[v] + dims[v]); }
    std::vector<float> sig(48 * NT);
    for (int i = 0; i < 48 * NT; i++) sig[i] = (float)sigma_noise48x54[i];
    if (!h->d_synth_f64) CU(cudaMalloc(&h->d_synth_f64, (48 * 48 * 2 + 54 * 54 + 48 * 2 + 54) * sizeof(double) + 48 * NT * sizeof(float)));
    const size_t S = h->cfg.max_tacs;
    if (!h->d_synth_truth) CU(cudaMalloc(&h->d_synth_truth, S * 96 * sizeof(float)));
    if (!h->d_synth_clean) CU(cudaMalloc(&h->d_synth_clean, S * 48 * NT * sizeof(float)));
    if (!h->d_synth_attempts) CU(cudaMalloc(&h->d_synth_attempts, S * sizeof(int)));
    if (!h->d_synth_capped) CU(cudaMalloc(&h->d_synth_capped, sizeof(int)));
    CU(cudaMemsetAsync(h->d_synth_capped, 0, sizeof(int), h->stream));
    CU(cudaMemcpyAsync(h->d_synth_f64, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    float* d_sig = reinterpret_cast<float*>(h->d_synth_f64 + buf.size());
    CU(cudaMemcpyAsync(d_sig, sig.data(), sig.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    for (int v = 0; v < 3; v++) { sp.AT[v] = h->d_synth_f64 + off[v]; sp.mu3[v] = h->d_synth_f64 + moff[v]; }
    sp.sigma = d_sig;
    sp.k2p = (float)k2p;
    sp.seed = seed;
    sp.tac_gid0 = h->cfg.tac_gid0;
    sp.tac_gids = h->have_tac_gids ? h->d_tac_gids : nullptr;
    sp.n_capped = h->d_synth_capped;
    sp.d2_max = h->d_synth_cinv ? h->synth_d2_max : 0.0;
    sp.cinv[0] = h->d_synth_cinv;
    sp.cinv[1] = h->d_synth_cinv ? h->d_synth_cinv + 48 * 48 : nullptr;
    sp.cinv[2] = h->d_synth_cinv ? h->d_synth_cinv + 2 * 48 * 48 : nullptr;
    sp.n_tac = n_tac;
    sp.y = h->d_y; sp.cref = h->d_cref; sp.k2p_out = h->d_k2p;
    sp.truth = h->d_synth_truth; sp.clean = h->d_synth_clean; sp.attempts = h->d_synth_attempts;
    h->n_tac = n_tac;
    SweepParams p = base_params(h);
    synth_kernel<<<n_tac, 64, HOOK_SMEM, h->stream>>>(p, sp);
    CU(cudaGetLastError());
    int capped = 0;
    CU(cudaMemcpyAsync(&capped, h->d_synth_capped, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->have_data = true;
    if (capped > 0)
        return fail(h, PETMH_ESYNTH, "%d of %d synthetic TACs hit a rejection cap (4000 positivity / Mahalanobis redraws of one vector or 1000 "
                    "negative-TAC redraws of the triple): their data is not a valid draw; petmh_synth_get's attempts[] is negative "
                    "for them", capped, n_tac);
    return PETMH_OK;
}

extern "C" int petmh_synth_get(petmh_t* h, float* dvr_r1 /*[n][96]*/, double* tac_ref /*[n][54]*/, float* tac_clean /*[n][48][54]*/,
                               float* y /*[n][48][54]*/, int* attempts /*[n]*/) {
    if (!h || !h->d_synth_truth) return fail(h, PETMH_EINVAL, "petmh_synth not called");
    CU(cudaSetDevice(h->cfg.device));
    const size_t n = h->n_tac;
    if (dvr_r1) CU(cudaMemcpyAsync(dvr_r1, h->d_synth_truth, n * 96 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (tac_ref) CU(cudaMemcpyAsync(tac_ref, h->d_cref, n * NT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (tac_clean) CU(cudaMemcpyAsync(tac_clean, h->d_synth_clean, n * 48 * NT * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (y) CU(cudaMemcpyAsync(y, h->d_y, n * 48 * NT * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (attempts) CU(cudaMemcpyAsync(attempts, h->d_synth_attempts, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PETMH_OK;
}
