/* mh_oracle.c -- CPU oracle of the MH-MCMC / SRTM2 hot path in plain C (fp64), TEST INFRASTRUCTURE ONLY.
 *
 * The same "lean" algorithm as oracle/mh.py: exact operator form of the reference's
 * resample-convolve-interpolate convolution (kinetic_model.py:12-32) conv = M exp(-k2a t), truncated-normal
 * log-likelihood of mcmc.py:152-155, MvNormal priors through r = P (q - mu), pymc's element-wise Metropolis
 * with its tune table (mcmc.py:156-157; pymc 5.12 semantics -- PARITY UNPINNED, see oracle/__init__.py).
 * Chain state and proposal arithmetic are float32 exactly like the CUDA kernel; every log-probability is fp64.
 *
 * Two uses: (1) a fast checker that consumes the same explicit tape as oracle/mh.py and the kernel's taped
 * mode; (2) the "restructured CPU" baseline of bench.py (one thread per chain group from Python, internal xoshiro generator).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it.
 *
 * Build: make -C oracle/c   (gcc -O3 -ffp-contract=off -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NR 48
#define NT 54

typedef struct {
    const double* M;      /* [54][54] dense operator */
    const int* ncol;      /* [54] number of leading active columns used by row j */
    const int* acol;      /* [n_active] active column -> frame index */
    const double* t;      /* [54] */
    const double* cr;     /* [54] */
    double k2p;
    const double* y;      /* [48][54] */
    const double* sig;    /* [48][54] */
    const double* mu;     /* [2][48] */
    const double* P;      /* [2][48][48] */
} model_t;

/* reduced per-ROI log-likelihood (state-independent constants dropped), SURVEY.md A.3 */
static double ll_roi(const model_t* m, int roi, double dvr, double r1) {
    const double k2 = m->k2p * r1, k2a = k2 / dvr, coef = k2 - r1 * k2a;
    double e[NT];
    for (int f = 0; f < NT; f++) e[f] = exp(-k2a * m->t[f]);
    double ll = 0.0;
    const double* y = m->y + roi * NT;
    const double* sg = m->sig + roi * NT;
    for (int j = 0; j < NT; j++) {
        double conv = 0.0;
        const double* row = m->M + j * NT;
        for (int c = 0; c < m->ncol[j]; c++) conv += row[m->acol[c]] * e[m->acol[c]];
        double s = r1 * m->cr[j] + coef * conv;
        if (s < 0) s = 1e-6;                                     /* mcmc.py:152 */
        if (y[j] < 0) return -INFINITY;
        const double d = y[j] - s;
        ll += -0.5 * d * d / (s * sg[j] * sg[j]) - 0.5 * log(s) - log1p(-0.5 * erfc(sqrt(s) / (sg[j] * M_SQRT2)));
    }
    return ll;
}

static float tune_factor(int c) {
    if (c < 1) return 0.1f;
    if (c < 5) return 0.5f;
    if (c < 20) return 0.9f;
    if (c > 95) return 10.0f;
    if (c > 75) return 2.0f;
    if (c > 50) return 1.1f;
    return 1.0f;
}

/* xoshiro256** for the tape-free throughput mode */
typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(rng_t* r) {
    const uint64_t res = rotl(r->s[1] * 5, 7) * 9, t = r->s[1] << 17;
    r->s[2] ^= r->s[0]; r->s[3] ^= r->s[1]; r->s[1] ^= r->s[2]; r->s[0] ^= r->s[3]; r->s[2] ^= t; r->s[3] = rotl(r->s[3], 45);
    return res;
}
static double rng_u01(rng_t* r) { return ((rng_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static void rng_seed(rng_t* r, uint64_t seed) {
    for (int i = 0; i < 4; i++) { seed += 0x9E3779B97F4A7C15ull; uint64_t z = seed; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; r->s[i] = z ^ (z >> 31); }
}

/* One chain.  tape pointers (normals/logu [n_sweeps][2][48] f32, rank u8) may be NULL -> internal generator.
 * Outputs may be NULL.  Returns the number of accepted moves. */
/* forced != NULL: teacher forcing -- decisions are evaluated (accept[], delta[]) but the state follows the
 * trajectory `forced` [n_sweeps][2][48] of another implementation; forced_acc[] receives the decision that
 * trajectory took, undecidable[] marks proposals equal to the state. */
static long run_chain(const model_t* m, int n_sweeps, int n_tune, const float* normals, const float* logu,
                      const uint8_t* rank, uint64_t seed, float* draws, uint8_t* accept, double* delta, float* scale_out,
                      const float* forced, uint8_t* forced_acc, uint8_t* undecidable) {
    float q[2][NR], scale[2][NR];
    int cnt[2][NR];
    double ll[NR], r[NR];
    rng_t rng;
    rng_seed(&rng, seed);
    for (int b = 0; b < 2; b++) for (int i = 0; i < NR; i++) { q[b][i] = (float)m->mu[b * NR + i]; scale[b][i] = 1.0f; cnt[b][i] = 0; }
    for (int i = 0; i < NR; i++) ll[i] = ll_roi(m, i, q[0][i], q[1][i]);
    long nacc = 0;
    for (int s = 0; s < n_sweeps; s++) {
        for (int b = 0; b < 2; b++) {
            if (s < n_tune && s > 0 && s % 100 == 0)
                for (int i = 0; i < NR; i++) { scale[b][i] = scale[b][i] * tune_factor(cnt[b][i]); cnt[b][i] = 0; }
            const double* P = m->P + b * NR * NR;
            for (int i = 0; i < NR; i++) {
                double a = 0.0;
                for (int j = 0; j < NR; j++) a += P[i * NR + j] * ((double)q[b][j] - m->mu[b * NR + j]);
                r[i] = a;
            }
            int order[NR];
            float nrm[NR], lu[NR];
            if (normals) {
                const size_t o = ((size_t)s * 2 + b) * NR;
                for (int i = 0; i < NR; i++) { nrm[i] = normals[o + i]; lu[i] = logu[o + i]; order[rank[o + i]] = i; }
            } else {
                for (int i = 0; i < NR; i += 2) {
                    const double rr = sqrt(-2.0 * log(rng_u01(&rng))), th = 6.283185307179586 * rng_u01(&rng);
                    nrm[i] = (float)(rr * cos(th)); nrm[i + 1] = (float)(rr * sin(th));
                }
                for (int i = 0; i < NR; i++) { lu[i] = (float)log(rng_u01(&rng)); order[i] = i; }
                for (int i = NR - 1; i > 0; i--) { const int j = (int)(rng_next(&rng) % (uint64_t)(i + 1)); const int tt = order[i]; order[i] = order[j]; order[j] = tt; }
            }
            for (int v = 0; v < NR; v++) {
                const int i = order[v];
                const float step = nrm[i] * scale[b][i];          /* fl32(n * scale) */
                const float qn = q[b][i] + step;                  /* fl32(q + step)  */
                const double d = (double)qn - (double)q[b][i];
                const double lln = b == 0 ? ll_roi(m, i, qn, q[1][i]) : ll_roi(m, i, q[0][i], qn);
                const double dl = (lln - ll[i]) - (d * r[i] + 0.5 * d * d * P[i * NR + i]);
                const int acc = isfinite(dl) && (double)lu[i] < dl;
                const size_t o = ((size_t)s * 2 + b) * NR + i;
                if (delta) delta[o] = dl;
                if (accept) accept[o] = (uint8_t)acc;
                int apply = acc;
                if (forced) {
                    const float fv = forced[o];
                    if (qn == q[b][i]) { undecidable[o] = 1; forced_acc[o] = (uint8_t)acc; }
                    else { undecidable[o] = 0; apply = (fv == qn); forced_acc[o] = (uint8_t)apply; }
                }
                if (apply) {
                    q[b][i] = qn; ll[i] = lln; cnt[b][i]++; nacc++;
                    for (int j = 0; j < NR; j++) r[j] += P[j * NR + i] * d;
                }
            }
            if (draws) memcpy(draws + ((size_t)s * 2 + b) * NR, q[b], NR * sizeof(float));
        }
    }
    if (scale_out) memcpy(scale_out, scale, sizeof scale);
    return nacc;
}

static model_t make_model(const double* M, const int* ncol, const int* acol, const double* t, const double* cr, double k2p,
                          const double* y, const double* sig, const double* mu, const double* P) {
    model_t m = {M, ncol, acol, t, cr, k2p, y, sig, mu, P};
    return m;
}

/* taped chain (checker) */
long mh_oracle_taped(const double* M, const int* ncol, const int* acol, const double* t, const double* cr, double k2p,
                     const double* y, const double* sig, const double* mu, const double* P, int n_sweeps, int n_tune,
                     const float* normals, const float* logu, const uint8_t* rank, float* draws, uint8_t* accept,
                     double* delta, float* scale_out) {
    const model_t m = make_model(M, ncol, acol, t, cr, k2p, y, sig, mu, P);
    return run_chain(&m, n_sweeps, n_tune, normals, logu, rank, 0, draws, accept, delta, scale_out, NULL, NULL, NULL);
}

/* teacher-forced replay of another implementation's trajectory on the same tape */
long mh_oracle_forced(const double* M, const int* ncol, const int* acol, const double* t, const double* cr, double k2p,
                      const double* y, const double* sig, const double* mu, const double* P, int n_sweeps, int n_tune,
                      const float* normals, const float* logu, const uint8_t* rank, const float* forced, uint8_t* accept,
                      uint8_t* forced_acc, uint8_t* undecidable, double* delta, float* scale_out) {
    const model_t m = make_model(M, ncol, acol, t, cr, k2p, y, sig, mu, P);
    return run_chain(&m, n_sweeps, n_tune, normals, logu, rank, 0, NULL, accept, delta, scale_out, forced, forced_acc, undecidable);
}

/* free-running chains in parallel (baseline / statistics): draws_out [n_chains][n_sweeps][2][48] or NULL */
long mh_oracle_free(const double* M, const int* ncol, const int* acol, const double* t, const double* cr, double k2p,
                    const double* y, const double* sig, const double* mu, const double* P, int n_chains, int n_sweeps,
                    int n_tune, uint64_t seed, float* draws_out) {
    const model_t m = make_model(M, ncol, acol, t, cr, k2p, y, sig, mu, P);
    long total = 0;
/* chains run serially here; callers parallelise over threads (ctypes releases the GIL): libgomp is not in this image */
    for (int c = 0; c < n_chains; c++)
        total += run_chain(&m, n_sweeps, n_tune, NULL, NULL, NULL, seed + 1000003ull * (uint64_t)c,
                           draws_out ? draws_out + (size_t)c * n_sweeps * 2 * NR : NULL, NULL, NULL, NULL, NULL, NULL, NULL);
    return total;
}

double mh_oracle_ll_roi(const double* M, const int* ncol, const int* acol, const double* t, const double* cr, double k2p,
                        const double* y, const double* sig, int roi, double dvr, double r1) {
    const model_t m = make_model(M, ncol, acol, t, cr, k2p, y, sig, NULL, NULL);
    return ll_roi(&m, roi, dvr, r1);
}
