/* ORACLE — test infrastructure only (never linked into the product).
 *
 * Plain-C restatement of docs/SPEC.md §1-§3: Philox4x32-10, the deterministic exp/log/sincos and
 * the Box-Muller normal pair that stand in for Julia's global RNG + Distributions.jl `Normal`
 * (call sites /root/reference/src/particles.jl:97-98,123-124; src/state_space_models.jl:93,102,108).
 * Written from the SPEC text with coefficient tables and loops; the CUDA side
 * (sequential_monte_carlo_b200/csrc/smcb_detmath.cuh) is a separate, unrolled implementation and
 * tests/test_detmath.py checks the two agree bit-for-bit.
 * Compile with -ffp-contract=off: only the fma() calls below may fuse.
 * PARITY UNPINNED: the reference ships no tests or golden vectors; this file is
 * pinned by the Philox known-answer vectors, by libm (<= 2 ulp) and by exact-rational spot checks.
 */
#ifndef SMC_ORACLE_DET_MATH_H
#define SMC_ORACLE_DET_MATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static const double O_MAGIC = 0x1.8p52;
static const double O_LN2_HI = 0x1.62e42fee00000p-1;
static const double O_LN2_LO = 0x1.a39ef35793c76p-33;
static const double O_LOG2E = 0x1.71547652b82fep+0;
static const double O_HALF_LOG_2PI = 0x1.d67f1c864beb5p-1;
static const double O_SQRT2 = 0x1.6a09e667f3bcdp+0;

static const double O_EXP_E[10] = {
    0x1.0000000000001p-1,  0x1.5555555555556p-3,  0x1.5555555553d63p-5,  0x1.11111111109b3p-7,
    0x1.6c16c1788bd90p-10, 0x1.a01a01a7c41d5p-13, 0x1.a019b90d2ae7ap-16, 0x1.71de0dae63bb3p-19,
    0x1.289185613a3d6p-22, 0x1.af38a9b0ec855p-26};
static const double O_LOG_R[7] = {0x1.5555555555558p-1, 0x1.99999999952d7p-2, 0x1.2492492df281ap-2,
                                  0x1.c71c62e3f11e6p-3, 0x1.7462b51cb66b1p-3, 0x1.39fe51a7c18f9p-3,
                                  0x1.2b5900de53b32p-3};
static const double O_SINQ_S[7] = {0x1.921fb54442d18p+0,  -0x1.4abbce625be41p-1, 0x1.466bc677587f3p-4,
                                   -0x1.32d2cce2e55bfp-8, 0x1.50782fd9b7104p-13, -0x1.e300715607854p-19,
                                   0x1.e3f362f896ffep-25};
static const double O_COSQ_C[8] = {0x1.0000000000000p+0,  -0x1.3bd3cc9be45dep+0, 0x1.03c1f081b5aacp-2,
                                   -0x1.55d3c7e3c90f2p-6, 0x1.e1f50683554a4p-11, -0x1.a6d1ec788deb9p-16,
                                   0x1.f9cc40b4d973bp-22, -0x1.b2649ccb4360dp-28};

static inline double o_horner(const double *c, int n, double z) {
  double acc = c[n - 1];
  for (int k = n - 2; k >= 0; --k) acc = fma(acc, z, c[k]);
  return acc;
}

static inline double o_from_bits(uint64_t b) {
  double d;
  memcpy(&d, &b, sizeof d);
  return d;
}
static inline uint64_t o_to_bits(double d) {
  uint64_t b;
  memcpy(&b, &d, sizeof b);
  return b;
}
static inline uint64_t o_mulhi(uint64_t a, uint64_t b) {
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
}

/* ---- SPEC §1 ---- */
static inline void o_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  uint32_t k[2] = {key[0], key[1]};
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n[4];
    n[0] = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    n[1] = (uint32_t)p1;
    n[2] = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    n[3] = (uint32_t)p0;
    memcpy(c, n, sizeof c);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  memcpy(out, c, sizeof c);
}

/* ---- SPEC §3 ---- */
static inline void o_exp_parts(double x, double *p, int *k) {
  double kf = (x * O_LOG2E + O_MAGIC) - O_MAGIC;
  double r = fma(-kf, O_LN2_HI, x);
  r = fma(-kf, O_LN2_LO, r);
  double E = o_horner(O_EXP_E, 10, r);
  *p = 1.0 + fma(r * r, E, r);
  *k = (int)kf;
}
static inline double o_ldexp_bits(double p, int n) {
  return o_from_bits(o_to_bits(p) + ((uint64_t)(int64_t)n << 52));
}
static inline double o_exp(double x) {
  if (x < -700.0) return 0.0;
  if (x > 700.0) return INFINITY;
  double p;
  int k;
  o_exp_parts(x, &p, &k);
  return o_ldexp_bits(p, k);
}
static inline uint64_t o_quant(double x, int S) {
  if (!(x >= -700.0)) return 0;
  double p;
  int k;
  o_exp_parts(x, &p, &k);
  if (k + S < 0) return 0;
  uint64_t v = (uint64_t)o_ldexp_bits(p, k + S);
  uint64_t cap = (uint64_t)1 << S;
  return v < cap ? v : cap;
}
static inline double o_log(double u) {
  uint64_t b = o_to_bits(u);
  int e = (int)((b >> 52) & 0x7FF) - 1023;
  double m = o_from_bits((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
  if (m > O_SQRT2) {
    m = m * 0.5;
    e = e + 1;
  }
  double f = m - 1.0;
  double s = f / (2.0 + f);
  double z = s * s;
  double lm = fma(s * z, o_horner(O_LOG_R, 7, z), s + s);
  return fma((double)e, O_LN2_HI, fma((double)e, O_LN2_LO, lm));
}
static inline void o_sincos2pi(double u, double *sn, double *cs) {
  double a = 4.0 * u;
  double nf = (a + O_MAGIC) - O_MAGIC;
  double r = a - nf;
  int n = ((int)nf) & 3;
  double z = r * r;
  double sr = r * o_horner(O_SINQ_S, 7, z);
  double cr = o_horner(O_COSQ_C, 8, z);
  switch (n) {
    case 0: *sn = sr; *cs = cr; break;
    case 1: *sn = cr; *cs = -sr; break;
    case 2: *sn = -sr; *cs = -cr; break;
    default: *sn = -cr; *cs = sr; break;
  }
}

/* ---- SPEC §2 ---- */
static inline uint32_t o_purpose(uint32_t kind, uint32_t comp, uint32_t epoch) {
  return ((kind | (comp << 4)) << 24) | (epoch & 0xFFFFFFu);
}
static inline double o_u52(uint32_t hi, uint32_t lo) {
  uint64_t k = ((uint64_t)hi << 20) | (lo >> 12);
  return (double)(2 * k + 1) * 0x1p-53;
}
/* normals for particles 2*pair and 2*pair+1 */
static inline void o_normal_pair(uint64_t seed, uint32_t epoch, uint32_t pair, uint32_t stream,
                                 uint32_t t, uint32_t kind, uint32_t comp, double *z0, double *z1) {
  uint32_t ctr[4] = {pair, stream, t, o_purpose(kind, comp, epoch)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  o_philox(ctr, key, r);
  double u1 = o_u52(r[0], r[1]);
  double u2 = o_u52(r[2], r[3]);
  double rho = sqrt(-2.0 * o_log(u1));
  double sn, cs;
  o_sincos2pi(u2, &sn, &cs);
  *z0 = rho * cs;
  *z1 = rho * sn;
}
static inline double o_normal(uint64_t seed, uint32_t epoch, uint32_t i, uint32_t stream, uint32_t t,
                              uint32_t kind, uint32_t comp) {
  double z0, z1;
  o_normal_pair(seed, epoch, i >> 1, stream, t, kind, comp, &z0, &z1);
  return (i & 1) ? z1 : z0;
}
static inline uint64_t o_uniform64(uint64_t seed, uint32_t epoch, uint32_t i, uint32_t stream,
                                   uint32_t t, uint32_t kind) {
  uint32_t ctr[4] = {i >> 1, stream, t, o_purpose(kind, 0, epoch)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  o_philox(ctr, key, r);
  return (i & 1) ? (((uint64_t)r[2] << 32) | r[3]) : (((uint64_t)r[0] << 32) | r[1]);
}


/* ---- SPEC §9b: the binary32 ARITHMETIC tier (every operation below is an IEEE binary32 operation; only fmaf fuses) ---- */
static const float OF_MAGIC = 0x1.8p23f;
static const float OF_LN2_HI = 0x1.62e4p-1f;
static const float OF_LN2_LO = 0x1.7f7d1cp-20f;
static const float OF_LOG2E = 0x1.715476p+0f;
static const float OF_HALF_LOG_2PI = 0x1.d67f1cp-1f;
static const float OF_SQRT2 = 0x1.6a09e6p+0f;
/* lowest degree first */
static const float OF_EXP_E[5] = {0x1.0p-1f, 0x1.5554dcp-3f, 0x1.55551ap-5f, 0x1.120b6ep-7f, 0x1.6d110ap-10f};
static const float OF_LOG_R[3] = {0x1.55555cp-1f, 0x1.997c2ep-2f, 0x1.2ee656p-2f};
static const float OF_SINQ_S[4] = {0x1.921fb6p+0f, -0x1.4abbbap-1f, 0x1.465ec2p-4f, -0x1.2d9b1ep-8f};
static const float OF_COSQ_C[4] = {0x1.0p+0f, -0x1.3bd392p+0f, 0x1.03af58p-2f, -0x1.4e5dd4p-6f};

static inline float of_horner(const float *c, int n, float z) {
  float acc = c[n - 1];
  for (int k = n - 2; k >= 0; --k) acc = fmaf(acc, z, c[k]);
  return acc;
}
static inline float of_from_bits(uint32_t b) { float f; memcpy(&f, &b, sizeof f); return f; }
static inline uint32_t of_to_bits(float f) { uint32_t b; memcpy(&b, &f, sizeof b); return b; }
static inline void of_exp_parts(float x, float *p, int *k) {
  float kf = (x * OF_LOG2E + OF_MAGIC) - OF_MAGIC;
  float r = fmaf(-kf, OF_LN2_HI, x);
  r = fmaf(-kf, OF_LN2_LO, r);
  float E = of_horner(OF_EXP_E, 5, r);
  *p = 1.0f + fmaf(r * r, E, r);
  *k = (int)kf;
}
static inline float of_exp(float x) {
  if (x < -86.0f) return 0.0f;
  if (x > 87.0f) return INFINITY;
  float p;
  int k;
  of_exp_parts(x, &p, &k);
  return of_from_bits(of_to_bits(p) + ((uint32_t)k << 23));
}
/* q = min(trunc(exp(x) 2^S), 2^S), x <= 0: exact integer value of p 2^(k+S) for the 24-bit mantissa of p */
static inline uint64_t of_quant(float x, int S) {
  if (!(x >= -86.0f)) return 0;
  float p;
  int k;
  of_exp_parts(x, &p, &k);
  int ep;
  float fr = frexpf(p, &ep);                       /* p = fr 2^ep, fr in [1/2, 1) */
  uint64_t M = (uint64_t)ldexpf(fr, 24);           /* 24-bit integer mantissa */
  int sh = ep - 24 + k + S;
  uint64_t v = sh >= 0 ? (M << sh) : (sh > -64 ? (M >> (-sh)) : 0);
  uint64_t cap = (uint64_t)1 << S;
  return v < cap ? v : cap;
}
static inline float of_log(float u) {
  uint32_t b = of_to_bits(u);
  int e = (int)((b >> 23) & 0xFF) - 127;
  float m = of_from_bits((b & 0x007FFFFFu) | 0x3F800000u);
  if (m > OF_SQRT2) {
    m = m * 0.5f;
    e = e + 1;
  }
  float f = m - 1.0f;
  float s = f / (2.0f + f);
  float z = s * s;
  float lm = fmaf(s * z, of_horner(OF_LOG_R, 3, z), s + s);
  return fmaf((float)e, OF_LN2_HI, fmaf((float)e, OF_LN2_LO, lm));
}
static inline void of_sincos2pi(float u, float *sn, float *cs) {
  float a = 4.0f * u;
  float nf = (a + OF_MAGIC) - OF_MAGIC;
  float r = a - nf;
  int n = ((int)nf) & 3;
  float z = r * r;
  float sr = r * of_horner(OF_SINQ_S, 4, z);
  float cr = of_horner(OF_COSQ_C, 4, z);
  switch (n) {
    case 0: *sn = sr; *cs = cr; break;
    case 1: *sn = cr; *cs = -sr; break;
    case 2: *sn = -sr; *cs = -cr; break;
    default: *sn = -cr; *cs = sr; break;
  }
}
/* particle i uses the Philox block at index i >> 2: words (r0, r1) -> the Box-Muller pair of particles 4q, 4q+1, (r2, r3) -> 4q+2, 4q+3 */
static inline float of_normal(uint64_t seed, uint32_t epoch, uint32_t i, uint32_t stream, uint32_t t, uint32_t kind, uint32_t comp) {
  uint32_t ctr[4] = {i >> 2, stream, t, o_purpose(kind, comp, epoch)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  o_philox(ctr, key, r);
  int h = (i >> 1) & 1;
  float u1 = (float)(2u * (r[2 * h] >> 9) + 1u) * 0x1p-24f;
  float u2 = (float)(2u * (r[2 * h + 1] >> 9) + 1u) * 0x1p-24f;
  float rho = sqrtf(-2.0f * of_log(u1));
  float sn, cs;
  of_sincos2pi(u2, &sn, &cs);
  return (i & 1) ? rho * sn : rho * cs;
}

#endif
