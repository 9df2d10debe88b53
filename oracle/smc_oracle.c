/* ORACLE — test infrastructure only.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product never does.
 *
 * CPU restatement of the reference's particle-filter path, one function per reference function,
 * with the un-pinned third-party pieces (RNG, Normal sampler, categorical sampler) replaced by the
 * frozen definitions of docs/SPEC.md.  Structure deliberately mirrors the Julia code: a
 * per-particle loop, resampling on every step, fresh temporaries on every step.
 *
 *   smco_normalize          <- normalize            /root/reference/src/particles.jl:5-15
 *   smco_ancestors          <- resample             /root/reference/src/particles.jl:17-19 (SPEC §5)
 *   smco_bootstrap_init     <- bootstrap_filter     /root/reference/src/particles.jl:87-105
 *   smco_bootstrap_step     <- bootstrap_filter!    /root/reference/src/particles.jl:107-129
 *   smco_log_likelihood     <- log_likelihood       /root/reference/src/particles.jl:132-147
 *   smco_batch_log_likelihood <- the Threads.@threads loops /root/reference/src/smc_samplers.jl:112-121,223-229
 *   model_*                 <- LinearModel / UCSV methods /root/reference/src/state_space_models.jl:74-109,215-259
 *   smco_kalman_*           <- kalman_filter / log_likelihood /root/reference/src/kalman_filter.jl:29-70
 *   smco_guided_*           <- particle_filter / particle_filter! /root/reference/src/particles.jl:28-84 (SPEC §10)
 *   smco_kalman_mv_*        <- kalman_filter (matrix)       /root/reference/src/kalman_filter.jl:3-27
 *   smco_simulate           <- simulate             /root/reference/src/state_space_models.jl:11-26
 *
 * PARITY UNPINNED: the reference has no tests, fixtures or golden vectors and
 * cannot be run here (Julia absent, module does not load).  What pins this file instead: Philox
 * KATs, libm agreement of the det-math, the Kalman likelihood on LG models, closed-form normalize
 * cases (tests/test_oracle.py).
 */
#include <stdio.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "det_math.h"

enum { KIND_LG1D = 0, KIND_SV = 1, KIND_UCSV = 2 };
enum { RS_MULTINOMIAL = 0, RS_STRATIFIED = 1, RS_SYSTEMATIC = 2 };
static int g_arith_f32 = 0;   /* SPEC §9b tier switch (set by smco_set_arith_f32 below) */
enum { P_INIT = 1, P_TRANS = 2, P_RESAMPLE = 3, P_SIMULATE = 8, P_RESAMPLE_CELL = 9 };

int smco_state_dim(int kind) { return kind >= 3 ? kind - 1 : (kind == KIND_UCSV ? 3 : 1); }   /* kinds 3..5: multivariate LG, d = 2..4 */

/* ------------------------------------------------------------------ vector math for the tests */
void smco_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { o_philox(ctr, key, out); }
void smco_exp(const double *x, double *y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = o_exp(x[i]); }
void smco_log(const double *x, double *y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = o_log(x[i]); }
void smco_expf(const double *x, double *y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = (double)of_exp((float)x[i]); }
void smco_logf(const double *x, double *y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = (double)of_log((float)x[i]); }
void smco_sincos2pif(const double *u, double *s, double *c, int64_t n) {
  for (int64_t i = 0; i < n; ++i) { float a, b; of_sincos2pi((float)u[i], &a, &b); s[i] = (double)a; c[i] = (double)b; }
}
void smco_quantf(const double *x, int S, uint64_t *q, int64_t n) { for (int64_t i = 0; i < n; ++i) q[i] = of_quant((float)x[i], S); }
void smco_sincos2pi(const double *u, double *s, double *c, int64_t n) {
  for (int64_t i = 0; i < n; ++i) o_sincos2pi(u[i], &s[i], &c[i]);
}
void smco_quant(const double *x, int S, uint64_t *q, int64_t n) { for (int64_t i = 0; i < n; ++i) q[i] = o_quant(x[i], S); }
void smco_normals(uint64_t seed, uint32_t epoch, uint32_t stream, uint32_t t, uint32_t kind, uint32_t comp,
                  int64_t n, double *out) {
  for (int64_t i = 0; i < n; ++i) out[i] = o_normal(seed, epoch, (uint32_t)i, stream, t, kind, comp);
}
void smco_uniforms64(uint64_t seed, uint32_t epoch, uint32_t stream, uint32_t t, uint32_t kind, int64_t n,
                     uint64_t *out) {
  for (int64_t i = 0; i < n; ++i) out[i] = o_uniform64(seed, epoch, (uint32_t)i, stream, t, kind);
}
int smco_quant_shift(int64_t n) {
  int c = 0;
  while (((int64_t)1 << c) < n) ++c;
  return 61 - c;
}

/* ------------------------------------------------------------------ models (SPEC §4) */
/* lower Cholesky factor of a positive SEMI-definite matrix (SPEC §4b): a non-positive pivot gives a zero column */
static void chol_psd(const double *A, int d, double *L) {
  for (int i = 0; i < d * d; ++i) L[i] = 0.0;
  for (int j = 0; j < d; ++j) {
    double s = A[j * d + j];
    for (int k = 0; k < j; ++k) s = s - L[j * d + k] * L[j * d + k];
    if (!(s > 0.0)) continue;
    L[j * d + j] = sqrt(s);
    for (int i = j + 1; i < d; ++i) {
      double v = A[i * d + j];
      for (int k = 0; k < j; ++k) v = v - L[i * d + k] * L[j * d + k];
      L[i * d + j] = v / L[j * d + j];
    }
  }
}
/* D holds 64 doubles; multivariate LG (state_space_models.jl:137-189): block A, B, Q, R, x0, Σ0 -> A | B | chol(Q) | 1/sqrt(R) | c | x0 | chol(Σ0) */
void smco_derive(int kind, const double *P, double *D) {
  for (int i = 0; i < 64; ++i) D[i] = 0.0;
  if (kind >= 3) {
    int d = kind - 1;
    const double *A = P, *B = P + d * d, *Q = B + d, *x0 = Q + d * d + 1, *S0 = x0 + d;
    double R = Q[d * d], sr = sqrt(R), *o = D;
    memcpy(o, A, sizeof(double) * d * d); o += d * d;
    memcpy(o, B, sizeof(double) * d); o += d;
    chol_psd(Q, d, o); o += d * d;
    o[0] = 1.0 / sr; o[1] = -(o_log(sr) + O_HALF_LOG_2PI); o += 2;
    memcpy(o, x0, sizeof(double) * d); o += d;
    chol_psd(S0, d, o);
    return;
  }
  if (kind == KIND_LG1D) { /* A,B,Q,R,x0,s0 : state_space_models.jl:74-109 */
    double sr = sqrt(P[3]);
    D[0] = P[0]; D[1] = P[1]; D[2] = sqrt(P[2]); D[3] = P[4]; D[4] = sqrt(P[5]);
    D[5] = 1.0 / sr;
    D[6] = -(o_log(sr) + O_HALF_LOG_2PI);
    D[7] = o_log(D[2]);                           /* log sd of the transition: guided weights (SPEC §10) */
  } else if (kind == KIND_SV) {
    D[0] = P[0]; D[1] = P[1]; D[2] = P[2];
    D[3] = P[2] / sqrt(1.0 - P[1] * P[1]);
    D[4] = o_log(P[2]);                           /* log sd of the transition: guided weights (SPEC §10) */
  } else { /* UCSV: state_space_models.jl:215-259 */
    D[0] = P[0]; D[1] = P[1]; D[2] = P[2]; D[3] = P[3]; D[4] = P[4];
    D[5] = o_exp(0.5 * P[3]);
  }
}

/* x: the d state components of one particle; z: d standard normals */
static void model_init(int kind, const double *D, const double *z, double *x) {
  if (kind >= 3) {                                /* MvNormal(x0, Σ0) :185-189 */
    int d = kind - 1;
    const double *x0 = D + 2 * d * d + d + 2, *L0 = x0 + d;
    for (int i = 0; i < d; ++i) {
      double acc = x0[i];
      for (int j = 0; j <= i; ++j) acc = fma(L0[i * d + j], z[j], acc);
      x[i] = acc;
    }
    return;
  }
  if (kind == KIND_LG1D) {
    x[0] = fma(D[4], z[0], D[3]);                 /* Normal(x0, sqrt(s0)) :105-109 */
  } else if (kind == KIND_SV) {
    x[0] = fma(D[3], z[0], D[0]);
  } else {
    x[0] = fma(D[5], z[0], D[2]);                 /* Normal(x0, exp(0.5 lse0)) :255 */
    x[1] = fma(D[0], z[1], D[3]);                 /* Normal(lse0, ge) :256 */
    x[2] = fma(D[1], z[2], D[4]);                 /* Normal(lsn0, gn) :257 */
  }
}
static void model_transition(int kind, const double *D, const double *z, const double *xp, double *x) {
  if (kind >= 3) {                                /* MvNormal(A x, Q) :163-171 */
    int d = kind - 1;
    const double *A = D, *LQ = D + d * d + d;
    for (int i = 0; i < d; ++i) {
      double acc = A[i * d] * xp[0];
      for (int j = 1; j < d; ++j) acc = fma(A[i * d + j], xp[j], acc);
      for (int j = 0; j <= i; ++j) acc = fma(LQ[i * d + j], z[j], acc);
      x[i] = acc;
    }
    return;
  }
  if (kind == KIND_LG1D) {
    x[0] = fma(D[2], z[0], D[0] * xp[0]);         /* Normal(A x, sqrt(Q)) :87-94 */
  } else if (kind == KIND_SV) {
    x[0] = fma(D[2], z[0], fma(D[1], xp[0] - D[0], D[0]));
  } else {
    double sd = o_exp(0.5 * xp[1]);               /* previous log-vol :238 */
    x[0] = fma(sd, z[0], xp[0]);
    x[1] = fma(D[0], z[1], xp[1]);
    x[2] = fma(D[1], z[2], xp[2]);
  }
}
static double model_logweight(int kind, const double *D, const double *x, double y) {
  if (kind >= 3) {                                /* Normal(B x, sqrt(R)) — R a variance as in kalman_filter.jl:16 (D9) */
    int d = kind - 1;
    const double *B = D + d * d, *tail = D + 2 * d * d + d;
    double m = B[0] * x[0];
    for (int j = 1; j < d; ++j) m = fma(B[j], x[j], m);
    double v = (y - m) * tail[0];
    return fma(-0.5 * v, v, tail[1]);
  }
  if (kind == KIND_LG1D) {                        /* logpdf(Normal(B x, sqrt(R)), y) :96-103 */
    double v = (y - D[1] * x[0]) * D[5];
    return fma(-0.5 * v, v, D[6]);
  } else if (kind == KIND_SV) {
    return fma(-0.5 * (y * y), o_exp(-x[0]), -(fma(0.5, x[0], O_HALF_LOG_2PI)));
  } else {                                        /* Normal(x, exp(0.5 lsn)) with current lsn :244-247 */
    double d = y - x[0];
    return fma(-0.5 * (d * d), o_exp(-x[2]), -(fma(0.5, x[2], O_HALF_LOG_2PI)));
  }
}

/* ------------------------------------------------------------------ a1 normalize */
void smco_normalize(const double *logw, int64_t n, double *logmu, double *w, double *ess) {
  double maxw = -INFINITY;
  for (int64_t i = 0; i < n; ++i) if (logw[i] > maxw) maxw = logw[i];      /* :6 */
  double *e = (double *)malloc(sizeof(double) * (size_t)n);
  double sumw = 0.0;
  for (int64_t i = 0; i < n; ++i) { e[i] = o_exp(logw[i] - maxw); sumw += e[i]; } /* :7-8 */
  *logmu = maxw + log(sumw) - log((double)n);                              /* :10 */
  double s2 = 0.0;
  for (int64_t i = 0; i < n; ++i) { double wi = e[i] / sumw; if (w) w[i] = wi; s2 += wi * wi; } /* :11-12 */
  *ess = 1.0 / s2;
  free(e);
}

/* ------------------------------------------------------------------ a2 resample (SPEC §5) */
static void ancestors_from_q_n(const uint64_t *q, int64_t n, int64_t n_out, int resampler, uint64_t seed, uint32_t epoch,
                               uint32_t stream, uint32_t t, uint32_t purpose, int64_t *anc);
static void ancestors_from_q(const uint64_t *q, int64_t n, int resampler, uint64_t seed, uint32_t epoch,
                             uint32_t stream, uint32_t t, uint32_t purpose, int64_t *anc) {
  ancestors_from_q_n(q, n, n, resampler, seed, epoch, stream, t, purpose, anc);
}
/* n_out draws from n weights (resample(w, N), particles.jl:17): thresholds i = 0 .. n_out-1 with strata of width floor((2^64-1)/n_out) */
static void ancestors_from_q_n(const uint64_t *q, int64_t n, int64_t n_out, int resampler, uint64_t seed, uint32_t epoch,
                               uint32_t stream, uint32_t t, uint32_t purpose, int64_t *anc) {
  uint64_t *C = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n);
  uint64_t run = 0;
  for (int64_t j = 0; j < n; ++j) { run += q[j]; C[j] = run; }
  uint64_t Q = run;
  uint64_t R = 0xFFFFFFFFFFFFFFFFull / (uint64_t)n_out;
  uint64_t U0 = o_uniform64(seed, epoch, 0, stream, t, purpose);
  for (int64_t i = 0; i < n_out; ++i) {
    if (Q == 0) { anc[i] = i < n ? i : n - 1; continue; }
    uint64_t F;
    if (resampler == RS_MULTINOMIAL) F = o_uniform64(seed, epoch, (uint32_t)i, stream, t, purpose);
    else if (resampler == RS_STRATIFIED) F = (uint64_t)i * R + o_mulhi(o_uniform64(seed, epoch, (uint32_t)i, stream, t, purpose), R);
    else F = (uint64_t)i * R + o_mulhi(U0, R);
    uint64_t tau = o_mulhi(F, Q);
    int64_t lo = 0, hi = n;                   /* a = #{j : C_j <= tau} */
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (C[mid] <= tau) lo = mid + 1; else hi = mid; }
    anc[i] = lo;
  }
  free(C);
}

/* SPEC §5c: multinomial resampling of a LARGE cloud (n > 8192) in two levels.  The reference draws n i.i.d. categorical
 * indices (particles.jl:18); their offspring counts are multinomial and the resampled particles exchangeable, so
 *   level 1: the n thresholds tau_i = floor(U32(i) Q / 2^32) are counted per cell of 4096 consecutive particles -> K_c;
 *   level 2: output position g of cell c (positions O_c .. O_c + K_c - 1, O = exclusive prefix of K) draws V32(g) from a second
 *            Philox purpose and takes the in-cell ancestor  c*4096 + #{ j in cell : Cloc_j <= floor(V32(g) W_c / 2^32) };
 *   (32-bit uniforms, four per Philox block: a threshold only has to pick one of 4096 cells / one of 4096 particles)
 *   order:   every chunk of 8192 consecutive output positions of a cell is written in ascending ancestor order. */
#define MN_CELL 4096
#define MN_CHUNK 8192
#define MN_LEGACY_MAX 8192
/* 32-bit uniform of SPEC §5c: word (i & 3) of the Philox block at index i >> 2; tau = floor(u * v / 2^32) */
static uint32_t o_uniform32(uint64_t seed, uint32_t epoch, uint32_t i, uint32_t stream, uint32_t t, uint32_t purpose) {
  uint32_t r[4];
  uint32_t ctr[4] = {i >> 2, stream, t, (purpose << 24) | (epoch & 0xFFFFFFu)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  o_philox(ctr, key, r);
  return r[i & 3];
}
static uint64_t o_mul32(uint32_t u, uint64_t v) { return (uint64_t)(((unsigned __int128)u * v) >> 32); }
static int cmp_i64(const void *a, const void *b) {
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return (x > y) - (x < y);
}
static void ancestors_two_level(const uint64_t *q, int64_t n, uint64_t seed, uint32_t epoch, uint32_t stream, uint32_t t,
                                int64_t *anc) {
  int64_t ncells = (n + MN_CELL - 1) / MN_CELL;
  uint64_t *C = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n);
  uint64_t *cellC = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)ncells);
  int64_t *K = (int64_t *)calloc((size_t)ncells, sizeof(int64_t));
  uint64_t run = 0;
  for (int64_t j = 0; j < n; ++j) { run += q[j]; C[j] = run; }
  uint64_t Q = run;
  if (Q == 0) {
    for (int64_t i = 0; i < n; ++i) anc[i] = i;
    free(C); free(cellC); free(K);
    return;
  }
  for (int64_t c = 0; c < ncells; ++c) {
    int64_t last = (c + 1) * MN_CELL - 1;
    if (last > n - 1) last = n - 1;
    cellC[c] = C[last];
  }
  for (int64_t i = 0; i < n; ++i) {                      /* level 1 */
    uint64_t tau = o_mul32(o_uniform32(seed, epoch, (uint32_t)i, stream, t, P_RESAMPLE), Q);
    int64_t lo = 0, hi = ncells;                         /* c = #{c : cellC_c <= tau} */
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (cellC[mid] <= tau) lo = mid + 1; else hi = mid; }
    K[lo] += 1;
  }
  int64_t O = 0;
  for (int64_t c = 0; c < ncells; ++c) {                 /* level 2 */
    int64_t j0 = c * MN_CELL, len = (n - j0 < MN_CELL) ? n - j0 : MN_CELL;
    uint64_t base = c ? cellC[c - 1] : 0, W = cellC[c] - base;
    for (int64_t g0 = O; g0 < O + K[c]; g0 += MN_CHUNK) {
      int64_t g1 = (g0 + MN_CHUNK < O + K[c]) ? g0 + MN_CHUNK : O + K[c];
      for (int64_t g = g0; g < g1; ++g) {
        uint64_t tau = o_mul32(o_uniform32(seed, epoch, (uint32_t)g, stream, t, P_RESAMPLE_CELL), W);
        int64_t lo = 0, hi = len;                        /* #{ j in cell : C_j - base <= tau } */
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (C[j0 + mid] - base <= tau) lo = mid + 1; else hi = mid; }
        anc[g] = j0 + lo;
      }
      qsort(anc + g0, (size_t)(g1 - g0), sizeof(int64_t), cmp_i64);
    }
    O += K[c];
  }
  free(C); free(cellC); free(K);
}

void smco_ancestors(const double *logw, int64_t n, int resampler, uint64_t seed, uint32_t epoch, uint32_t stream,
                    uint32_t t, int64_t *anc) {
  int S = smco_quant_shift(n);
  double mx = -INFINITY;
  for (int64_t i = 0; i < n; ++i) if (logw[i] > mx) mx = logw[i];
  uint64_t *q = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n);
  if (g_arith_f32) {
    for (int64_t i = 0; i < n; ++i) q[i] = of_quant((float)logw[i] - (float)mx, S);
  } else {
    for (int64_t i = 0; i < n; ++i) q[i] = o_quant(logw[i] - mx, S);
  }
  if (resampler == RS_MULTINOMIAL && n > MN_LEGACY_MAX) ancestors_two_level(q, n, seed, epoch, stream, t, anc);
  else ancestors_from_q(q, n, resampler, seed, epoch, stream, t, P_RESAMPLE, anc);
  free(q);
}

/* standalone resample(w): q_i = trunc((w_i / max w) 2^S) */
void smco_resample_w(const double *w, int64_t n, int resampler, uint64_t seed, uint32_t epoch, uint32_t stream,
                     uint32_t t, uint32_t purpose, int64_t *anc) {
  int S = smco_quant_shift(n);
  double mx = 0.0;
  for (int64_t i = 0; i < n; ++i) if (w[i] > mx) mx = w[i];
  uint64_t *q = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n);
  double scale = o_from_bits((uint64_t)(1023 + S) << 52);
  for (int64_t i = 0; i < n; ++i) q[i] = (mx > 0.0 && w[i] > 0.0) ? (uint64_t)((w[i] / mx) * scale) : 0;
  ancestors_from_q(q, n, resampler, seed, epoch, stream, t, purpose, anc);
  free(q);
}
/* resample(w, N) with N != length(w)   particles.jl:17-19 */
void smco_resample_w_n(const double *w, int64_t n, int64_t n_out, int resampler, uint64_t seed, uint32_t epoch, uint32_t stream,
                       uint32_t t, uint32_t purpose, int64_t *anc) {
  int S = smco_quant_shift(n);
  double mx = 0.0;
  for (int64_t i = 0; i < n; ++i) if (w[i] > mx) mx = w[i];
  uint64_t *q = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n);
  double scale = o_from_bits((uint64_t)(1023 + S) << 52);
  for (int64_t i = 0; i < n; ++i) q[i] = (mx > 0.0 && w[i] > 0.0) ? (uint64_t)((w[i] / mx) * scale) : 0;
  ancestors_from_q_n(q, n, n_out, resampler, seed, epoch, stream, t, purpose, anc);
  free(q);
}

/* SPEC §9: the binary32 STATE tier.  When set, every state component is rounded to binary32 right
 * after it is drawn (initial draw and transition) — it is what the device stores — and everything
 * downstream (log-weight, quantisation, resampling, the next transition) sees the rounded value,
 * all arithmetic staying binary64.  Process-global: tests set it, run, and reset it. */
static int g_state_f32 = 0;
void smco_set_state_f32(int on) { g_state_f32 = on ? 1 : 0; }
int smco_get_state_f32(void) { return g_state_f32; }
static inline void round_state(double *xi, int d) {
  if (g_state_f32)
    for (int k = 0; k < d; ++k) xi[k] = (double)(float)xi[k];
}

/* SPEC §9b: the binary32 ARITHMETIC tier.  When set, the normals (four per Philox block), the model arithmetic, the log-weights
 * and the exponential behind the fixed-point weights are binary32; states and log-weights are binary32 values (held in the
 * double arrays of this interface), the CDF / thresholds / ancestors machinery is unchanged.  Process-global like the state tier. */
void smco_set_arith_f32(int on) { g_arith_f32 = on ? 1 : 0; }
int smco_get_arith_f32(void) { return g_arith_f32; }
static void modelf_init(int kind, const float *D, const float *z, float *x) {
  if (kind == KIND_LG1D) x[0] = fmaf(D[4], z[0], D[3]);
  else if (kind == KIND_SV) x[0] = fmaf(D[3], z[0], D[0]);
  else { x[0] = fmaf(D[5], z[0], D[2]); x[1] = fmaf(D[0], z[1], D[3]); x[2] = fmaf(D[1], z[2], D[4]); }
}
static void modelf_transition(int kind, const float *D, const float *z, const float *xp, float *x) {
  if (kind == KIND_LG1D) x[0] = fmaf(D[2], z[0], D[0] * xp[0]);
  else if (kind == KIND_SV) x[0] = fmaf(D[2], z[0], fmaf(D[1], xp[0] - D[0], D[0]));
  else {
    float sd = of_exp(0.5f * xp[1]);
    x[0] = fmaf(sd, z[0], xp[0]);
    x[1] = fmaf(D[0], z[1], xp[1]);
    x[2] = fmaf(D[1], z[2], xp[2]);
  }
}
static float modelf_logweight(int kind, const float *D, const float *x, float y) {
  if (kind == KIND_LG1D) {
    float v = (y - D[1] * x[0]) * D[5];
    return fmaf(-0.5f * v, v, D[6]);
  } else if (kind == KIND_SV) {
    return fmaf(-0.5f * (y * y), of_exp(-x[0]), -(fmaf(0.5f, x[0], OF_HALF_LOG_2PI)));
  } else {
    float d = y - x[0];
    return fmaf(-0.5f * (d * d), of_exp(-x[2]), -(fmaf(0.5f, x[2], OF_HALF_LOG_2PI)));
  }
}
static void derive_f(int kind, const double *P, float *Df) {   /* the binary64 derived block rounded once to binary32 */
  double D[64];
  smco_derive(kind, P, D);
  for (int i = 0; i < 8; ++i) Df[i] = (float)D[i];
}
/* normalize() in the tier: e_i = expf(logw_i - max) in binary32, the sums in binary64 */
void smco_normalize_f32(const double *logw, int64_t n, double *logmu, double *w, double *ess) {
  float maxw = -INFINITY;
  for (int64_t i = 0; i < n; ++i) if ((float)logw[i] > maxw) maxw = (float)logw[i];
  double sumw = 0.0, sum2 = 0.0;
  for (int64_t i = 0; i < n; ++i) { double e = (double)of_exp((float)logw[i] - maxw); sumw += e; sum2 += e * e; }
  *logmu = (double)maxw + log(sumw) - log((double)n);
  *ess = (sumw * sumw) / sum2;
  if (w) for (int64_t i = 0; i < n; ++i) w[i] = (double)of_exp((float)logw[i] - maxw) / sumw;
}

/* ------------------------------------------------------------------ a3 bootstrap_filter */
/* x is SoA [d][n]. */
void smco_bootstrap_init(int kind, const double *P, int64_t n, double y, uint64_t seed, uint32_t epoch,
                         uint32_t stream, double *x, double *logw) {
  double D[64];
  smco_derive(kind, P, D);
  int d = smco_state_dim(kind);
  if (g_arith_f32) {
    float Df[8];
    derive_f(kind, P, Df);
    for (int64_t i = 0; i < n; ++i) {
      float z[4] = {0, 0, 0, 0}, xi[4];
      for (int k = 0; k < d; ++k) z[k] = of_normal(seed, epoch, (uint32_t)i, stream, 0, P_INIT, (uint32_t)k);
      modelf_init(kind, Df, z, xi);
      for (int k = 0; k < d; ++k) x[(int64_t)k * n + i] = (double)xi[k];
      logw[i] = (double)modelf_logweight(kind, Df, xi, (float)y);
    }
    return;
  }
  for (int64_t i = 0; i < n; ++i) {                                         /* :96-99 */
    double z[4] = {0, 0, 0, 0}, xi[4];
    for (int k = 0; k < d; ++k) z[k] = o_normal(seed, epoch, (uint32_t)i, stream, 0, P_INIT, (uint32_t)k);
    model_init(kind, D, z, xi);
    round_state(xi, d);
    for (int k = 0; k < d; ++k) x[(int64_t)k * n + i] = xi[k];
    logw[i] = model_logweight(kind, D, xi, y);
  }
}

/* ------------------------------------------------------------------ a4 bootstrap_filter! */
void smco_bootstrap_step(int kind, const double *P, int64_t n, double y, uint32_t t, int resampler, uint64_t seed,
                         uint32_t epoch, uint32_t stream, double *x, double *logw, int64_t *anc_out) {
  double D[64];
  smco_derive(kind, P, D);
  int d = smco_state_dim(kind);
  int64_t *a = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);              /* a = resample(weights) :117 */
  smco_ancestors(logw, n, resampler, seed, epoch, stream, t, a);
  double *xp = (double *)malloc(sizeof(double) * (size_t)(n * d));         /* xp = deepcopy(x[a]) :119 */
  for (int k = 0; k < d; ++k)
    for (int64_t i = 0; i < n; ++i) xp[(int64_t)k * n + i] = x[(int64_t)k * n + a[i]];
  if (g_arith_f32) {
    float Df[8];
    derive_f(kind, P, Df);
    for (int64_t i = 0; i < n; ++i) {
      float z[4] = {0, 0, 0, 0}, par[4] = {0, 0, 0, 0}, xi[4];
      for (int k = 0; k < d; ++k) {
        z[k] = of_normal(seed, epoch, (uint32_t)i, stream, t, P_TRANS, (uint32_t)k);
        par[k] = (float)xp[(int64_t)k * n + i];
      }
      modelf_transition(kind, Df, z, par, xi);
      for (int k = 0; k < d; ++k) x[(int64_t)k * n + i] = (double)xi[k];
      logw[i] = (double)modelf_logweight(kind, Df, xi, (float)y);
    }
    if (anc_out) memcpy(anc_out, a, sizeof(int64_t) * (size_t)n);
    free(xp);
    free(a);
    return;
  }
  for (int64_t i = 0; i < n; ++i) {                                         /* :122-125 */
    double z[4] = {0, 0, 0, 0}, par[4] = {0, 0, 0, 0}, xi[4];
    for (int k = 0; k < d; ++k) {
      z[k] = o_normal(seed, epoch, (uint32_t)i, stream, t, P_TRANS, (uint32_t)k);
      par[k] = xp[(int64_t)k * n + i];
    }
    model_transition(kind, D, z, par, xi);
    round_state(xi, d);
    for (int k = 0; k < d; ++k) x[(int64_t)k * n + i] = xi[k];
    logw[i] = model_logweight(kind, D, xi, y);
  }
  if (anc_out) memcpy(anc_out, a, sizeof(int64_t) * (size_t)n);
  free(xp);
  free(a);
}

/* ------------------------------------------------------------------ N3 particle_filter! with a proposal */
/* particles.jl:55-84 for the one-dimensional models and the affine-Gaussian proposal family of SPEC §10:
 * prop = (c0, c1, c2), x' ~ Normal(c0 + c1 xp, c2).  Same resampling and the same Philox normal as the
 * bootstrap step; logw = logpdf(observation(x'), y) + logpdf(transition(xp), x') - logpdf(proposal(xp), x'). */
/* UCSV (SPEC §10b): the two log-volatilities move by the transition, the trend by the conditionally optimal Gaussian tempered by
 * κ = prop[0] in [0, 1]:  r = σε²/ση'² with σε = exp(le/2) (previous, state_space_models.jl:238) and ση'² = exp(ln') (current, :246);
 * g = κ r/(1 + r); x' ~ Normal(x + g (y − x), (1 − g) σε²); κ = 0 is the bootstrap move, κ = 1 the locally optimal one. */
static void guided_step_ucsv(const double *D, int64_t n, double y, uint32_t t, uint64_t seed, uint32_t epoch, uint32_t stream,
                             const double *prop, const int64_t *a, double *x, double *logw) {
  const double kappa = prop[0];
  double *xp = (double *)malloc(sizeof(double) * (size_t)(3 * n));
  for (int k = 0; k < 3; ++k)
    for (int64_t i = 0; i < n; ++i) xp[(int64_t)k * n + i] = x[(int64_t)k * n + a[i]];
  for (int64_t i = 0; i < n; ++i) {
    double z[3], par[3], xi[3];
    for (int k = 0; k < 3; ++k) {
      z[k] = o_normal(seed, epoch, (uint32_t)i, stream, t, P_TRANS, (uint32_t)k);
      par[k] = xp[(int64_t)k * n + i];
    }
    xi[1] = fma(D[0], z[1], par[1]);
    xi[2] = fma(D[1], z[2], par[2]);
    double sd = o_exp(0.5 * par[1]);
    double r = (sd * sd) * o_exp(-xi[2]);
    double g = kappa * (r / (1.0 + r));
    double omg = 1.0 - g;
    double mq = fma(g, y - par[0], par[0]);
    double c2 = sd * sqrt(omg);
    xi[0] = fma(c2, z[0], mq);                                              /* x[i] = rand(proposal(model, xp[i])) :73 */
    double zt = (xi[0] - par[0]) / sd;
    double corr = fma(-0.5 * zt, zt, 0.5 * (z[0] * z[0])) + 0.5 * o_log(omg);  /* logpdf(transition) − logpdf(proposal) :77-78 */
    for (int k = 0; k < 3; ++k) x[(int64_t)k * n + i] = xi[k];
    logw[i] = model_logweight(KIND_UCSV, D, xi, y) + corr;                  /* :74 */
  }
  free(xp);
}

int smco_guided_step(int kind, const double *P, int64_t n, double y, uint32_t t, int resampler, uint64_t seed,
                     uint32_t epoch, uint32_t stream, const double *prop, double *x, double *logw, int64_t *anc_out) {
  if (kind > KIND_UCSV) return -1;
  double D[64];
  smco_derive(kind, P, D);
  int64_t *a = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);              /* a = resample(weights) :66 */
  smco_ancestors(logw, n, resampler, seed, epoch, stream, t, a);
  if (kind == KIND_UCSV) {
    guided_step_ucsv(D, n, y, t, seed, epoch, stream, prop, a, x, logw);
    if (anc_out) memcpy(anc_out, a, sizeof(int64_t) * (size_t)n);
    free(a);
    return 0;
  }
  const double c0 = prop[0], c1 = prop[1], c2 = prop[2], lc2 = o_log(prop[2]), ic2 = 1.0 / prop[2];
  const double isdf = 1.0 / D[2], lsdf = (kind == KIND_LG1D) ? D[7] : D[4];
  double *xp = (double *)malloc(sizeof(double) * (size_t)n);               /* xp = deepcopy(x[a]) :68 */
  for (int64_t i = 0; i < n; ++i) xp[i] = x[a[i]];
  for (int64_t i = 0; i < n; ++i) {                                         /* :72-80 */
    double z = o_normal(seed, epoch, (uint32_t)i, stream, t, P_TRANS, 0);
    double mq = fma(c1, xp[i], c0);
    double xi = fma(c2, z, mq);                                             /* x[i] = rand(proposal(model, xp[i])) :73 */
    round_state(&xi, 1);                                                    /* SPEC §9: the stored state is what every density sees */
    double mf = (kind == KIND_LG1D) ? D[0] * xp[i] : fma(D[1], xp[i] - D[0], D[0]);
    double zt = (xi - mf) * isdf;
    double zq = (xi - mq) * ic2;
    double lf = fma(-0.5 * zt, zt, -lsdf);                                  /* logpdf(transition(model, xp[i]), x[i]) :77 */
    double lq = fma(-0.5 * zq, zq, -lc2);                                   /* logpdf(proposal(xp[i]), x[i]) :78 */
    x[i] = xi;
    logw[i] = model_logweight(kind, D, &xi, y) + (lf - lq);                 /* :74 */
  }
  if (anc_out) memcpy(anc_out, a, sizeof(int64_t) * (size_t)n);
  free(xp);
  free(a);
  return 0;
}

/* whole series: bootstrap initial step (particles.jl:40-42; the unmatched "+ logpdf(initial_dist)" of :44 is ruled a
 * defect, SPEC §10), then guided steps with prop[t] = (c0, c1, c2) of time t (row 0 unused). */
double smco_guided_log_likelihood(int kind, const double *P, int64_t n, const double *y, int64_t T, int resampler,
                                  uint64_t seed, uint32_t epoch, uint32_t stream, const double *prop, int64_t prop_stride,
                                  double *x, double *logw, double *logmu_out, double *ess_out) {
  double *xl = x ? x : (double *)malloc(sizeof(double) * (size_t)(n * smco_state_dim(kind)));
  double *lw = logw ? logw : (double *)malloc(sizeof(double) * (size_t)n);
  double logZ = 0.0, lm, es;
  smco_bootstrap_init(kind, P, n, y[0], seed, epoch, stream, xl, lw);
  smco_normalize(lw, n, &lm, NULL, &es);
  logZ = lm;
  if (logmu_out) logmu_out[0] = lm;
  if (ess_out) ess_out[0] = es;
  for (int64_t t = 1; t < T; ++t) {
    smco_guided_step(kind, P, n, y[t], (uint32_t)t, resampler, seed, epoch, stream, prop + t * prop_stride, xl, lw, NULL);
    smco_normalize(lw, n, &lm, NULL, &es);
    logZ += lm;
    if (logmu_out) logmu_out[t] = lm;
    if (ess_out) ess_out[t] = es;
  }
  if (!x) free(xl);
  if (!logw) free(lw);
  return logZ;
}

/* M guided filters; prop is [T][M][3] */
void smco_batch_guided_log_likelihood(int kind, const double *P, const uint8_t *active, int64_t M, int64_t n,
                                      const double *y, int64_t T, int resampler, uint64_t seed, uint32_t epoch,
                                      uint32_t stream0, const double *prop, double *logZ, double *x, double *logw) {
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < M; ++m) {
    if (active && !active[m]) { logZ[m] = -INFINITY; continue; }
    logZ[m] = smco_guided_log_likelihood(kind, P + 8 * m, n, y, T, resampler, seed, epoch, stream0 + (uint32_t)m,
                                         prop + 3 * m, 3 * M, x ? x + m * n * smco_state_dim(kind) : NULL, logw ? logw + m * n : NULL, NULL, NULL);
  }
}

/* ------------------------------------------------------------------ a5 log_likelihood */
double smco_log_likelihood(int kind, const double *P, int64_t n, const double *y, int64_t T, int resampler,
                           uint64_t seed, uint32_t epoch, uint32_t stream, double *x, double *logw,
                           double *logmu_out, double *ess_out, int64_t *anc_out) {
  int d = smco_state_dim(kind);
  double *xl = x ? x : (double *)malloc(sizeof(double) * (size_t)(n * d));
  double *lw = logw ? logw : (double *)malloc(sizeof(double) * (size_t)n);
  double logZ = 0.0, lm, es;
  smco_bootstrap_init(kind, P, n, y[0], seed, epoch, stream, xl, lw);       /* :139 */
  (g_arith_f32 ? smco_normalize_f32 : smco_normalize)(lw, n, &lm, NULL, &es);
  logZ = lm;
  if (logmu_out) logmu_out[0] = lm;
  if (ess_out) ess_out[0] = es;
  for (int64_t t = 1; t < T; ++t) {                                         /* :141-144 */
    smco_bootstrap_step(kind, P, n, y[t], (uint32_t)t, resampler, seed, epoch, stream, xl, lw,
                        anc_out ? anc_out + t * n : NULL);
    (g_arith_f32 ? smco_normalize_f32 : smco_normalize)(lw, n, &lm, NULL, &es);
    logZ += lm;
    if (logmu_out) logmu_out[t] = lm;
    if (ess_out) ess_out[t] = es;
  }
  if (!x) free(xl);
  if (!logw) free(lw);
  return logZ;
}

/* M independent filters, threaded over theta like Threads.@threads (smc_samplers.jl:112,223).
 * Filter m uses Philox stream (stream0 + m); inactive filters (out-of-support proposals, :116) are
 * skipped and report logZ = -inf. x: [M][d][n], logw: [M][n]; either may be NULL. */
void smco_batch_log_likelihood(int kind, const double *P, const uint8_t *active, int64_t M, int64_t n,
                               const double *y, int64_t T, int resampler, uint64_t seed, uint32_t epoch,
                               uint32_t stream0, double *logZ, double *x, double *logw) {
  int d = smco_state_dim(kind);
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < M; ++m) {
    if (active && !active[m]) { logZ[m] = -INFINITY; continue; }
    logZ[m] = smco_log_likelihood(kind, P + 8 * m, n, y, T, resampler, seed, epoch, stream0 + (uint32_t)m,
                                  x ? x + m * d * n : NULL, logw ? logw + m * n : NULL, NULL, NULL, NULL);
  }
}

int smco_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ a16 Kalman (scalar) */
/* kalman_filter.jl:29-53: predict, then update, returns the step log-likelihood */
double smco_kalman_step(const double *P, double *xt, double *St, double yt, int predict) {
  double A = P[0], B = P[1], Q = P[2], R = P[3];
  double x = *xt, S = *St;
  if (predict) { x = A * x; S = (A * A) * S + Q; }
  double sig = (B * B) * S + R;
  double dy = yt - B * x;
  x = x + (S * B) * (1.0 / sig) * dy;
  S = S - ((S * B) * (S * B)) * (1.0 / sig);
  *xt = x; *St = S;
  return -0.5 * (log(2.0 * M_PI) + log(sig) + (dy / sig * dy));
}
/* kalman_filter.jl:55-70. matched_init=1 skips the first predict so the target equals the
 * particle filter's (x1 ~ N(x0, s0) with no transition). */
double smco_kalman_loglik(const double *P, const double *y, int64_t T, int matched_init, double *xT, double *ST) {
  double x = P[4], S = P[5], ll = 0.0;
  for (int64_t t = 0; t < T; ++t) ll += smco_kalman_step(P, &x, &S, y[t], !(matched_init && t == 0));
  if (xT) *xT = x;
  if (ST) *ST = S;
  return ll;
}

/* ------------------------------------------------------------------ N4 Kalman (matrix, scalar observation) */
/* kalman_filter.jl:3-27.  Model block, row-major: A[d][d], B[d], Q[d][d], R, x0[d], S0[d][d]; d <= 4. */
double smco_kalman_mv_step(int d, const double *P, double *x, double *S, double yt, int predict) {
  const double *A = P, *B = P + d * d, *Q = P + d * d + d;
  const double R = P[2 * d * d + d];
  double xn[4], AS[16], K[4];
  if (predict) {
    for (int i = 0; i < d; ++i) {                      /* xt = A*xt :13 */
      double acc = 0.0;
      for (int j = 0; j < d; ++j) acc += A[i * d + j] * x[j];
      xn[i] = acc;
      for (int k = 0; k < d; ++k) {
        double u = 0.0;
        for (int j = 0; j < d; ++j) u += A[i * d + j] * S[j * d + k];
        AS[i * d + k] = u;
      }
    }
    for (int i = 0; i < d; ++i) {                      /* St = A*St*A' + Q :14 */
      x[i] = xn[i];
      for (int k = 0; k < d; ++k) {
        double u = 0.0;
        for (int j = 0; j < d; ++j) u += AS[i * d + j] * A[k * d + j];
        S[i * d + k] = u + Q[i * d + k];
      }
    }
  }
  double bx = 0.0, sig = 0.0;
  for (int i = 0; i < d; ++i) {
    double u = 0.0;
    for (int j = 0; j < d; ++j) u += S[i * d + j] * B[j];
    K[i] = u;
    bx += B[i] * x[i];
  }
  for (int i = 0; i < d; ++i) sig += B[i] * K[i];
  sig += R;                                            /* :16 */
  double dy = yt - bx;                                 /* :17 */
  double inv = 1.0 / sig;
  for (int i = 0; i < d; ++i) {
    double g = K[i] * inv;
    x[i] = x[i] + g * dy;                              /* :20 */
    for (int j = 0; j < d; ++j) S[i * d + j] = S[i * d + j] - g * K[j];   /* :21 */
  }
  return -0.5 * (log(2.0 * M_PI) + log(sig) + (dy * inv * dy));          /* :24-26 */
}
/* log_likelihood(y, model) kalman_filter.jl:55-70 */
double smco_kalman_mv_loglik(int d, const double *P, const double *y, int64_t T, int matched_init, double *xT, double *ST) {
  double x[4], S[16], ll = 0.0;
  for (int i = 0; i < d; ++i) x[i] = P[2 * d * d + d + 1 + i];
  for (int i = 0; i < d * d; ++i) S[i] = P[2 * d * d + 2 * d + 1 + i];
  for (int64_t t = 0; t < T; ++t) ll += smco_kalman_mv_step(d, P, x, S, y[t], !(matched_init && t == 0));
  if (xT) for (int i = 0; i < d; ++i) xT[i] = x[i];
  if (ST) for (int i = 0; i < d * d; ++i) ST[i] = S[i];
  return ll;
}

/* ------------------------------------------------------------------ simulate */
/* state_space_models.jl:11-26 with Philox purpose 8: stream 0 = state noise, stream 1 = obs noise.
 * The observation draw inverts the model's weight function: y = mean + sd * z. */
void smco_simulate(int kind, const double *P, int64_t T, uint64_t seed, double *x, double *y) {
  double D[64];
  smco_derive(kind, P, D);
  int d = smco_state_dim(kind);
  double cur[4] = {0, 0, 0, 0}, nxt[4];
  for (int64_t t = 0; t < T; ++t) {
    double z[4] = {0, 0, 0, 0};
    for (int k = 0; k < d; ++k) z[k] = o_normal(seed, 0, (uint32_t)t, 0, 0, P_SIMULATE, (uint32_t)k);
    if (t == 0) model_init(kind, D, z, nxt); else model_transition(kind, D, z, cur, nxt);
    for (int k = 0; k < d; ++k) { cur[k] = nxt[k]; x[(int64_t)k * T + t] = cur[k]; }
    double zo = o_normal(seed, 0, (uint32_t)t, 1, 0, P_SIMULATE, 0);
    double mean, sd;
    if (kind >= 3) {
      const double *B = D + d * d;
      mean = B[0] * cur[0];
      for (int j = 1; j < d; ++j) mean = fma(B[j], cur[j], mean);
      sd = sqrt(P[2 * d * d + d]);
    }
    else if (kind == KIND_LG1D) { mean = D[1] * cur[0]; sd = sqrt(P[3]); }
    else if (kind == KIND_SV) { mean = 0.0; sd = o_exp(0.5 * cur[0]); }
    else { mean = cur[0]; sd = o_exp(0.5 * cur[2]); }
    y[t] = fma(sd, zo, mean);
  }
}

/* ------------------------------------------------------------------ reference-STYLE timing arm
 * What /root/reference/src/particles.jl:87-147 executes for a univariate LinearModel, with the costs the Julia code has:
 *   - resample(w) = sample(1:n, Weights(w), n) (:17-19): StatsBase's alias method — a Walker/Vose alias table built on EVERY step
 *     (four fresh arrays), then two uniforms and two random reads per draw; unsorted ancestors;
 *   - xp = deepcopy(x[a]) (:119): a fresh array and a random gather;
 *   - rand(Normal(A xp, sqrt(Q))) and logpdf(Normal(B x, sqrt(R)), y) per particle (:122-125, state_space_models.jl:87-103): the sqrt
 *     and the log σ are recomputed for every particle, as the reference's distribution objects do;
 *   - normalize (:5-15): max, exp, sum, a fresh weight vector, ess.
 * Random numbers: xoshiro256++ (the generator behind Julia's default RNG) and the polar method for normals (Julia uses a ziggurat).
 * NOT part of the parity oracle — its stream is its own; it is checked distributionally (E[Ẑ] = Z against the Kalman filter) and
 * used only as the timed CPU arm. */
typedef struct { uint64_t s[4]; double spare; int has; } rs_rng;
static inline uint64_t rs_rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rs_next(rs_rng *g) {
  uint64_t *s = g->s, r = rs_rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rs_rotl(s[3], 45);
  return r;
}
static void rs_seed(rs_rng *g, uint64_t seed) {
  for (int i = 0; i < 4; ++i) {                          /* splitmix64 */
    uint64_t z = (seed += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    g->s[i] = z ^ (z >> 31);
  }
  g->has = 0;
}
static inline double rs_uniform(rs_rng *g) { return (double)(rs_next(g) >> 11) * 0x1p-53; }
static inline double rs_randn(rs_rng *g) {
  if (g->has) { g->has = 0; return g->spare; }
  double u, v, s;
  do { u = 2.0 * rs_uniform(g) - 1.0; v = 2.0 * rs_uniform(g) - 1.0; s = u * u + v * v; } while (s >= 1.0 || s == 0.0);
  double f = sqrt(-2.0 * log(s) / s);
  g->spare = v * f; g->has = 1;
  return u * f;
}
static inline double rs_normal_logpdf(double mu, double sigma, double y) {   /* logpdf(Normal(mu, sigma), y) */
  double z = (y - mu) / sigma;
  return -0.5 * z * z - log(sigma) - 0.9189385332046727;
}
/* normalize(logw) -> (logμ, w, ess)  particles.jl:5-15; w is a fresh vector */
static double *rs_normalize(const double *logw, int64_t n, double *logmu, double *ess) {
  double maxw = -INFINITY, sumw = 0.0, s2 = 0.0;
  for (int64_t i = 0; i < n; ++i) if (logw[i] > maxw) maxw = logw[i];
  double *w = (double *)malloc(sizeof(double) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) { w[i] = exp(logw[i] - maxw); sumw += w[i]; }
  *logmu = maxw + log(sumw) - log((double)n);
  for (int64_t i = 0; i < n; ++i) { w[i] /= sumw; s2 += w[i] * w[i]; }
  *ess = 1.0 / s2;
  return w;
}
/* sample(1:n, Weights(w), n): make_alias_table! + alias_sample! */
static int64_t *rs_alias_sample(rs_rng *g, const double *w, int64_t n) {
  double *ap = (double *)malloc(sizeof(double) * (size_t)n);
  int64_t *alias = (int64_t *)malloc(sizeof(int64_t) * (size_t)n), *larges = (int64_t *)malloc(sizeof(int64_t) * (size_t)n),
          *smalls = (int64_t *)malloc(sizeof(int64_t) * (size_t)n), *a = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
  int64_t kl = 0, ks = 0;
  for (int64_t i = 0; i < n; ++i) {
    ap[i] = w[i] * (double)n;
    alias[i] = i;
    if (ap[i] > 1.0) larges[kl++] = i; else if (ap[i] < 1.0) smalls[ks++] = i;
  }
  while (kl > 0 && ks > 0) {
    int64_t s = smalls[--ks], l = larges[--kl];
    alias[s] = l;
    ap[l] = (ap[l] + ap[s]) - 1.0;
    if (ap[l] > 1.0) larges[kl++] = l; else if (ap[l] < 1.0) smalls[ks++] = l;
  }
  for (int64_t i = 0; i < n; ++i) {
    int64_t j = (int64_t)(rs_uniform(g) * (double)n);
    if (j >= n) j = n - 1;
    a[i] = (rs_uniform(g) < ap[j]) ? j : alias[j];
  }
  free(ap); free(alias); free(larges); free(smalls);
  return a;
}
double smco_reference_style_log_likelihood(const double *P, int64_t n, const double *y, int64_t T, uint64_t seed) {
  const double A = P[0], B = P[1], Q = P[2], R = P[3], x0 = P[4], s0 = P[5];
  rs_rng g;
  rs_seed(&g, seed);
  double *x = (double *)malloc(sizeof(double) * (size_t)n), *logw = (double *)malloc(sizeof(double) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) {                      /* bootstrap_filter :96-99 */
    x[i] = x0 + sqrt(s0) * rs_randn(&g);
    logw[i] = rs_normal_logpdf(B * x[i], sqrt(R), y[0]);
  }
  double logmu, ess, logZ;
  double *w = rs_normalize(logw, n, &logmu, &ess);
  free(logw);
  logZ = logmu;
  for (int64_t t = 1; t < T; ++t) {                      /* bootstrap_filter! :107-129 */
    logw = (double *)malloc(sizeof(double) * (size_t)n);  /* logw = similar(weights) */
    int64_t *a = rs_alias_sample(&g, w, n);               /* a = resample(weights) */
    double *xp = (double *)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) xp[i] = x[a[i]];      /* xp = deepcopy(x[a]) */
    for (int64_t i = 0; i < n; ++i) {
      x[i] = A * xp[i] + sqrt(Q) * rs_randn(&g);          /* rand(transition(model, xp[i])) */
      logw[i] = rs_normal_logpdf(B * x[i], sqrt(R), y[t]);
    }
    free(w); free(a); free(xp);
    w = rs_normalize(logw, n, &logmu, &ess);
    free(logw);
    logZ += logmu;
  }
  free(w); free(x);
  return logZ;
}

/* the reference-style arm for the UCSV model (state_space_models.jl:215-259): array-of-3-vectors state as the reference keeps it
 * (x[i] is a 3-vector; kept contiguous here — the reference allocates a fresh heap vector per particle and step, which this port
 * does NOT charge), alias-table resampling rebuilt every step, libm exp / log, per-step allocations. */
double smco_reference_style_log_likelihood_ucsv(const double *P, int64_t n, const double *y, int64_t T, uint64_t seed) {
  const double ge = P[0], gn = P[1], x0 = P[2], lse0 = P[3], lsn0 = P[4];
  rs_rng g;
  rs_seed(&g, seed);
  double *x = (double *)malloc(sizeof(double) * 3 * (size_t)n), *logw = (double *)malloc(sizeof(double) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) {                      /* initial_dist :249-259 */
    double *s = x + 3 * i;
    s[0] = x0 + exp(0.5 * lse0) * rs_randn(&g);
    s[1] = lse0 + ge * rs_randn(&g);
    s[2] = lsn0 + gn * rs_randn(&g);
    logw[i] = rs_normal_logpdf(s[0], exp(0.5 * s[2]), y[0]);   /* observation :244-247 */
  }
  double logmu, ess, logZ;
  double *w = rs_normalize(logw, n, &logmu, &ess);
  free(logw);
  logZ = logmu;
  for (int64_t t = 1; t < T; ++t) {
    logw = (double *)malloc(sizeof(double) * (size_t)n);
    int64_t *a = rs_alias_sample(&g, w, n);
    double *xp = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    for (int64_t i = 0; i < n; ++i) memcpy(xp + 3 * i, x + 3 * a[i], sizeof(double) * 3);   /* xp = deepcopy(x[a]) */
    for (int64_t i = 0; i < n; ++i) {
      const double *p = xp + 3 * i;
      double *s = x + 3 * i;
      s[0] = p[0] + exp(0.5 * p[1]) * rs_randn(&g);      /* transition :233-242 (previous log σε) */
      s[1] = p[1] + ge * rs_randn(&g);
      s[2] = p[2] + gn * rs_randn(&g);
      logw[i] = rs_normal_logpdf(s[0], exp(0.5 * s[2]), y[t]);
    }
    free(w); free(a); free(xp);
    w = rs_normalize(logw, n, &logmu, &ess);
    free(logw);
    logZ += logmu;
  }
  free(w); free(x);
  return logZ;
}

/* M reference-style filters, `Threads.@threads for m` (smc_samplers.jl:112): OpenMP over θ; kind 0 (LG1D) or 2 (UCSV) */
int smco_reference_style_batch(int kind, const double *P, int64_t M, int64_t n, const double *y, int64_t T, uint64_t seed, double *logZ) {
  if (kind != KIND_LG1D && kind != KIND_UCSV) return -1;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t m = 0; m < M; ++m)
    logZ[m] = (kind == KIND_LG1D) ? smco_reference_style_log_likelihood(P + 8 * m, n, y, T, seed + (uint64_t)m)
                                  : smco_reference_style_log_likelihood_ucsv(P + 8 * m, n, y, T, seed + (uint64_t)m);
  return 0;
}
