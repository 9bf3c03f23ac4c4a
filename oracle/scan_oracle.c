/*
 * oracle/scan_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the selective-scan arithmetic that nnUZoo's Mamba blocks
 * reach through `selective_scan_fn`.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (nnuzoo_b200/) never does.
 *
 * Forward follows the reference's pure-PyTorch statement
 *   nnunetv2/nets/seg_mamba/selective_scan_interface.py:86-152 (selective_scan_ref)
 * line by line (citations inline).  The backward is the reverse-time adjoint of
 * that forward; the reference has no closed-form
 * CPU backward (it relies on autograd through :86-152), so the adjoint here is
 * pinned against autograd of the reference itself by oracle/gen_golden.py and
 * tests/test_oracle_golden.py.
 *
 * Parity status: PINNED against outputs of the reference's own
 * selective_scan_ref (forward) and torch.autograd through it (backward), run in
 * the build container and committed under tests/golden/ (the reference ships no
 * golden vectors of its own: SURVEY.md section 4).
 *
 * Two arithmetic flavours are built from the same source:
 *   acc_t = float  : same precision/operation order as the reference (fp32,
 *                    separate multiply and add; built with -ffp-contract=off)
 *   acc_t = double : for error budgeting of the fp32 CUDA kernels.
 *
 * Layouts (all C-contiguous, L innermost), same as the reference docstring :88-100
 *   u, delta, z, out : (batch, dim, L)
 *   A                : (dim, dstate)
 *   B, C             : (batch, ngroups, dstate, L)   (3-D B/C == ngroups 1)
 *   D, delta_bias    : (dim)
 *   last_state       : (batch, dim, dstate)
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* F.softplus with the default threshold of 20 (selective_scan_interface.py:107) */
#define DEFINE_SCAN(SUFFIX, acc_t, EXP, LOG1P)                                              \
                                                                                            \
static inline acc_t softplus_##SUFFIX(acc_t x) { return x > (acc_t)20 ? x : LOG1P(EXP(x)); } \
static inline acc_t sigmoid_##SUFFIX(acc_t x) { return (acc_t)1 / ((acc_t)1 + EXP(-x)); }   \
                                                                                            \
/* one (batch, dim) row of the forward; optionally records h_t for the adjoint */           \
static void row_fwd_##SUFFIX(const float *u, const float *delta, const float *A,            \
                             const float *Bm, const float *Cm, const float *Dp,             \
                             const float *z, const float *bias, int softplus,               \
                             int dstate, long L, float *out, float *last_state,             \
                             acc_t *dl_store, acc_t *h_store, acc_t *y_store)               \
{                                                                                           \
    acc_t x[64]; /* dstate <= 64 */                                                         \
    for (int n = 0; n < dstate; ++n) x[n] = 0;          /* :119 x = zeros */                \
    for (long i = 0; i < L; ++i) {                      /* :133 loop over L */              \
        acc_t dl = (acc_t)delta[i];                     /* :103 */                          \
        if (bias) dl = dl + (acc_t)bias[0];             /* :105 */                          \
        if (softplus) dl = softplus_##SUFFIX(dl);       /* :107 */                          \
        acc_t uu = (acc_t)u[i];                         /* :102 */                          \
        acc_t y = 0;                                                                        \
        for (int n = 0; n < dstate; ++n) {                                                  \
            acc_t dA = EXP(dl * (acc_t)A[n]);           /* :121 deltaA */                   \
            acc_t dBu = dl * (acc_t)Bm[(long)n * L + i] * uu; /* :129 deltaB_u */           \
            x[n] = dA * x[n] + dBu;                     /* :134 */                          \
            y += x[n] * (acc_t)Cm[(long)n * L + i];     /* :141 */                          \
            if (h_store) h_store[i * dstate + n] = x[n];                                    \
        }                                                                                   \
        if (dl_store) dl_store[i] = dl;                                                     \
        acc_t o = y;                                                                        \
        if (Dp) o = y + uu * (acc_t)Dp[0];              /* :148 */                          \
        if (y_store) y_store[i] = o;                                                        \
        if (z) { acc_t zz = (acc_t)z[i]; o = o * (zz * sigmoid_##SUFFIX(zz)); } /* :150 */  \
        if (out) out[i] = (float)o;                     /* :151 cast */                     \
    }                                                                                       \
    if (last_state)                                                                         \
        for (int n = 0; n < dstate; ++n) last_state[n] = (float)x[n]; /* :142-143 */        \
}                                                                                           \
                                                                                            \
int nzo_scan_fwd_##SUFFIX(const float *u, const float *delta, const float *A,               \
                          const float *B, const float *C, const float *D, const float *z,   \
                          const float *delta_bias, int delta_softplus, int batch, int dim,  \
                          int dstate, long L, int ngroups, float *out, float *last_state)   \
{                                                                                           \
    if (dstate > 64 || ngroups < 1 || dim % ngroups) return 1;                              \
    const int dpg = dim / ngroups;                      /* :128 repeat G -> G*H */          \
    _Pragma("omp parallel for collapse(2) schedule(dynamic)")                               \
    for (int b = 0; b < batch; ++b)                                                         \
        for (int d = 0; d < dim; ++d) {                                                     \
            const long row = (long)b * dim + d;                                             \
            const int g = d / dpg;                                                          \
            const float *Bm = B + ((long)b * ngroups + g) * dstate * L;                     \
            const float *Cm = C + ((long)b * ngroups + g) * dstate * L;                     \
            row_fwd_##SUFFIX(u + row * L, delta + row * L, A + (long)d * dstate, Bm, Cm,    \
                             D ? D + d : NULL, z ? z + row * L : NULL,                      \
                             delta_bias ? delta_bias + d : NULL, delta_softplus, dstate, L, \
                             out + row * L, last_state ? last_state + row * dstate : NULL,  \
                             NULL, NULL, NULL);                                             \
        }                                                                                   \
    return 0;                                                                               \
}                                                                                           \
                                                                                            \
/* Reverse-time adjoint of the forward above.  All gradient  */            \
/* outputs are overwritten.  Any of dD, dz, ddelta_bias may be NULL.          */            \
int nzo_scan_bwd_##SUFFIX(const float *u, const float *delta, const float *A,               \
                          const float *B, const float *C, const float *D, const float *z,   \
                          const float *delta_bias, int delta_softplus, const float *dout,   \
                          int batch, int dim, int dstate, long L, int ngroups,              \
                          float *du, float *ddelta, float *dA, float *dB, float *dC,        \
                          float *dD, float *dz, float *ddelta_bias)                         \
{                                                                                           \
    if (dstate > 64 || ngroups < 1 || dim % ngroups) return 1;                              \
    const int dpg = dim / ngroups;                                                          \
    const long nbg = (long)batch * ngroups;                                                 \
    /* per-batch partials for the (dim)-shaped reductions keep the result    */             \
    /* independent of the thread schedule                                    */             \
    acc_t *dA_part = (acc_t *)calloc((size_t)batch * dim * dstate, sizeof(acc_t));          \
    acc_t *dD_part = (acc_t *)calloc((size_t)batch * dim, sizeof(acc_t));                   \
    acc_t *db_part = (acc_t *)calloc((size_t)batch * dim, sizeof(acc_t));                   \
    if (!dA_part || !dD_part || !db_part) return 2;                                         \
    int fail = 0;                                                                           \
    _Pragma("omp parallel for schedule(dynamic)")                                           \
    for (long bg = 0; bg < nbg; ++bg) {                                                     \
        const int b = (int)(bg / ngroups), g = (int)(bg % ngroups);                         \
        const float *Bm = B + bg * dstate * L;                                              \
        const float *Cm = C + bg * dstate * L;                                              \
        acc_t *dBa = (acc_t *)calloc((size_t)dstate * L, sizeof(acc_t));                    \
        acc_t *dCa = (acc_t *)calloc((size_t)dstate * L, sizeof(acc_t));                    \
        acc_t *hs = (acc_t *)malloc((size_t)dstate * L * sizeof(acc_t));                    \
        acc_t *dls = (acc_t *)malloc((size_t)L * sizeof(acc_t));                            \
        acc_t *ys = (acc_t *)malloc((size_t)L * sizeof(acc_t));                             \
        if (!dBa || !dCa || !hs || !dls || !ys) { fail = 1; }                               \
        else for (int dd = 0; dd < dpg; ++dd) {                                             \
            const int d = g * dpg + dd;                                                     \
            const long row = (long)b * dim + d;                                             \
            const float *ur = u + row * L, *Ar = A + (long)d * dstate;                      \
            const float *zr = z ? z + row * L : NULL;                                       \
            const float *gor = dout + row * L;                                              \
            row_fwd_##SUFFIX(ur, delta + row * L, Ar, Bm, Cm, D ? D + d : NULL, NULL,       \
                             delta_bias ? delta_bias + d : NULL, delta_softplus, dstate, L, \
                             NULL, NULL, dls, hs, ys);                                      \
            acc_t dh[64], anext[64];                                                        \
            for (int n = 0; n < dstate; ++n) { dh[n] = 0; anext[n] = 0; }                   \
            acc_t dDacc = 0, dbacc = 0;                                                     \
            for (long i = L - 1; i >= 0; --i) {                                             \
                acc_t go = (acc_t)gor[i], dy = go, uu = (acc_t)ur[i], dl = dls[i];          \
                if (zr) {                                                                   \
                    acc_t zz = (acc_t)zr[i], sg = sigmoid_##SUFFIX(zz);                     \
                    dy = go * (zz * sg);                                                    \
                    if (dz) dz[row * L + i] =                                               \
                        (float)(go * ys[i] * (sg * ((acc_t)1 + zz * ((acc_t)1 - sg))));     \
                }                                                                           \
                acc_t dui = D ? dy * (acc_t)D[d] : (acc_t)0;                                \
                dDacc += dy * uu;                                                           \
                acc_t ddl = 0, sB = 0;                                                      \
                for (int n = 0; n < dstate; ++n) {                                          \
                    acc_t a = EXP(dl * (acc_t)Ar[n]);                                       \
                    acc_t Bv = (acc_t)Bm[(long)n * L + i], Cv = (acc_t)Cm[(long)n * L + i]; \
                    acc_t hprev = i > 0 ? hs[(i - 1) * dstate + n] : (acc_t)0;              \
                    dh[n] = Cv * dy + anext[n] * dh[n];                                     \
                    anext[n] = a;                                                           \
                    dCa[(long)n * L + i] += dy * hs[i * dstate + n];                        \
                    dBa[(long)n * L + i] += dh[n] * dl * uu;                                \
                    sB += dh[n] * Bv;                                                       \
                    acc_t gq = dh[n] * a * hprev;                                           \
                    ddl += dh[n] * Bv * uu + (acc_t)Ar[n] * gq;                             \
                    dA_part[((long)b * dim + d) * dstate + n] += dl * gq;                   \
                }                                                                           \
                dui += dl * sB;                                                             \
                du[row * L + i] = (float)dui;                                               \
                if (delta_softplus) {                                                       \
                    acc_t raw = (acc_t)delta[row * L + i] +                                 \
                                (delta_bias ? (acc_t)delta_bias[d] : (acc_t)0);             \
                    if (!(raw > (acc_t)20)) ddl = ddl * sigmoid_##SUFFIX(raw);              \
                }                                                                           \
                ddelta[row * L + i] = (float)ddl;                                           \
                dbacc += ddl;                                                               \
            }                                                                               \
            dD_part[(long)b * dim + d] = dDacc;                                             \
            db_part[(long)b * dim + d] = dbacc;                                             \
        }                                                                                   \
        if (!fail) for (long k = 0; k < (long)dstate * L; ++k) {                            \
            dB[bg * dstate * L + k] = (float)dBa[k];                                        \
            dC[bg * dstate * L + k] = (float)dCa[k];                                        \
        }                                                                                   \
        free(dBa); free(dCa); free(hs); free(dls); free(ys);                                \
    }                                                                                       \
    for (long k = 0; k < (long)dim * dstate; ++k) {                                         \
        acc_t s = 0;                                                                        \
        for (int b = 0; b < batch; ++b) s += dA_part[(long)b * dim * dstate + k];           \
        dA[k] = (float)s;                                                                   \
    }                                                                                       \
    for (int d = 0; d < dim; ++d) {                                                         \
        acc_t s = 0, t = 0;                                                                 \
        for (int b = 0; b < batch; ++b) { s += dD_part[(long)b * dim + d];                  \
                                          t += db_part[(long)b * dim + d]; }                \
        if (dD) dD[d] = (float)s;                                                           \
        if (ddelta_bias) ddelta_bias[d] = (float)t;                                         \
    }                                                                                       \
    free(dA_part); free(dD_part); free(db_part);                                            \
    return fail ? 2 : 0;                                                                    \
}

DEFINE_SCAN(f32, float, expf, log1pf)
DEFINE_SCAN(f64, double, exp, log1p)

int nzo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void nzo_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
