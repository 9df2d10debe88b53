// Deterministic binary64 math + Philox4x32-10 for the particle-filter hot path (docs/SPEC.md §1-§3).
//
// Every function here is a fixed sequence of IEEE-754 operations (compiled with -fmad=false, so
// only the fma() calls written out are fused).  The same sequences are restated independently in
// oracle/det_math.h; tests compare the two bit-for-bit.  Replaces the un-pinned Distributions.jl
// Normal sampler / logpdf used at /root/reference/src/particles.jl:97-98,123-124.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define SMCB_HD __host__ __device__ __forceinline__
#else
#define SMCB_HD inline
#endif

namespace smcb {

// ---------------------------------------------------------------------------------------------
// bit casts
SMCB_HD double u64_as_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double d;
  __builtin_memcpy(&d, &b, 8);
  return d;
#endif
}
SMCB_HD uint64_t double_as_u64(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t b;
  __builtin_memcpy(&b, &d, 8);
  return b;
#endif
}
SMCB_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (SPEC §1)
struct Philox4 {
  uint32_t r0, r1, r2, r3;
};

SMCB_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umulhi(a, b);
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}

SMCB_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                              uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
    philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 o;
  o.r0 = c0;
  o.r1 = c1;
  o.r2 = c2;
  o.r3 = c3;
  return o;
}

// purposes (SPEC §2): low nibble kind, high nibble state component
enum : uint32_t {
  PURPOSE_INIT = 1,
  PURPOSE_TRANSITION = 2,
  PURPOSE_RESAMPLE = 3,
  PURPOSE_PRIOR = 4,
  PURPOSE_THETA_RESAMPLE = 5,
  PURPOSE_MH_PROPOSAL = 6,
  PURPOSE_MH_ACCEPT = 7,
  PURPOSE_SIMULATE = 8,
  PURPOSE_RESAMPLE_CELL = 9,  // in-cell thresholds of the two-level multinomial resampler (SPEC §5c)
};

struct RngKey {
  uint32_t k0, k1;    // seed halves
  uint32_t epoch;     // 24-bit sweep ordinal
  // the ten round keys (k0 + r W0, k1 + r W1), derived once on the host: as a kernel argument they sit
  // in the constant bank and feed the round's XOR directly (no per-thread key schedule)
  uint32_t rk[10][2];
};

SMCB_HD RngKey make_rng_key(uint32_t k0, uint32_t k1, uint32_t epoch) {
  RngKey key;
  key.k0 = k0;
  key.k1 = k1;
  key.epoch = epoch;
  for (int r = 0; r < 10; ++r) {
    key.rk[r][0] = k0 + (uint32_t)r * 0x9E3779B9u;
    key.rk[r][1] = k1 + (uint32_t)r * 0xBB67AE85u;
  }
  return key;
}

SMCB_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const RngKey& key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
    philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ key.rk[r][0];
    uint32_t n2 = hi0 ^ c3 ^ key.rk[r][1];
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
  }
  Philox4 o;
  o.r0 = c0;
  o.r1 = c1;
  o.r2 = c2;
  o.r3 = c3;
  return o;
}

SMCB_HD uint32_t purpose_word(uint32_t kind, uint32_t comp, uint32_t epoch) {
  return ((kind | (comp << 4)) << 24) | (epoch & 0xFFFFFFu);
}

// ---------------------------------------------------------------------------------------------
// constants (SPEC §3)
#define SMCB_MAGIC 0x1.8p52
#define SMCB_LN2_HI 0x1.62e42fee00000p-1
#define SMCB_LN2_LO 0x1.a39ef35793c76p-33
#define SMCB_LOG2E 0x1.71547652b82fep+0
#define SMCB_HALF_LOG_2PI 0x1.d67f1c864beb5p-1
#define SMCB_SQRT2 0x1.6a09e667f3bcdp+0

// Polynomial coefficients, highest degree first.  On the device they live in constant memory so
// that DFMA takes them as c[bank][offset] operands; as literals ptxas re-materialises every one
// with two UMOVs per use (11-19 % of all issued instructions in the first ncu captures).
#if defined(__CUDA_ARCH__)
#define SMCB_COEF_DECL static __constant__ double
#else
#define SMCB_COEF_DECL static const double
#endif
SMCB_COEF_DECL kExpC[10] = {0x1.af38a9b0ec855p-26, 0x1.289185613a3d6p-22, 0x1.71de0dae63bb3p-19, 0x1.a019b90d2ae7ap-16,
                            0x1.a01a01a7c41d5p-13, 0x1.6c16c1788bd90p-10, 0x1.11111111109b3p-7,  0x1.5555555553d63p-5,
                            0x1.5555555555556p-3,  0x1.0000000000001p-1};
SMCB_COEF_DECL kLogC[7] = {0x1.2b5900de53b32p-3, 0x1.39fe51a7c18f9p-3, 0x1.7462b51cb66b1p-3, 0x1.c71c62e3f11e6p-3,
                           0x1.2492492df281ap-2, 0x1.99999999952d7p-2, 0x1.5555555555558p-1};
SMCB_COEF_DECL kSinC[7] = {0x1.e3f362f896ffep-25, -0x1.e300715607854p-19, 0x1.50782fd9b7104p-13, -0x1.32d2cce2e55bfp-8,
                           0x1.466bc677587f3p-4,  -0x1.4abbce625be41p-1,  0x1.921fb54442d18p+0};
SMCB_COEF_DECL kCosC[8] = {-0x1.b2649ccb4360dp-28, 0x1.f9cc40b4d973bp-22, -0x1.a6d1ec788deb9p-16, 0x1.e1f50683554a4p-11,
                           -0x1.55d3c7e3c90f2p-6,  0x1.03c1f081b5aacp-2,  -0x1.3bd3cc9be45dep+0,  0x1.0000000000000p+0};
SMCB_COEF_DECL kMathC[6] = {SMCB_MAGIC, SMCB_LN2_HI, SMCB_LN2_LO, SMCB_LOG2E, SMCB_HALF_LOG_2PI, SMCB_SQRT2};

// exp(x) = p * 2^k, p in about [0.707, 1.415]
SMCB_HD void det_exp_parts(double x, double& p, int& k) {
  // t = x*log2(e) + 1.5*2^52 holds round(x*log2 e) in its low mantissa bits: k comes from the bit
  // pattern (no F2I), kf = t - MAGIC is the same integer as a double (SPEC §3: k = int(kf)).
  const double t = x * kMathC[3] + kMathC[0];
  const double kf = t - kMathC[0];
  k = (int)(uint32_t)double_as_u64(t);
  double r = fma(-kf, kMathC[1], x);
  r = fma(-kf, kMathC[2], r);
  double e = kExpC[0];
#pragma unroll
  for (int i = 1; i < 10; ++i) e = fma(e, r, kExpC[i]);
  p = 1.0 + fma(r * r, e, r);
}

SMCB_HD double scale_pow2(double p, int n) {  // p * 2^n for normal results
  return u64_as_double(double_as_u64(p) + ((uint64_t)(int64_t)n << 52));
}

SMCB_HD double det_exp(double x) {
  if (x < -700.0) return 0.0;
  if (x > 700.0) return INFINITY;
  double p;
  int k;
  det_exp_parts(x, p, k);
  return scale_pow2(p, k);
}

// e = exp(x) and q = min(trunc(exp(x) * 2^S), 2^S) for x <= 0  (SPEC §3 det_quant).
// Branch-free (selects only) so that independent evaluations interleave in the FP64 pipe.
SMCB_HD void det_exp_quant(double x, int S, double& e, uint64_t& q) {
  const bool valid = (x >= -700.0);             // false for x < -700, -inf and NaN
  const double xc = valid ? x : -700.0;
  double p;
  int k;
  det_exp_parts(xc, p, k);
  const double ev = scale_pow2(p, k);
  e = valid ? ev : ((x < -700.0) ? 0.0 : x);    // NaN propagates into the float sums; q = 0
  const int ks = k + S;
  const uint64_t v = (uint64_t)scale_pow2(p, ks < 0 ? 0 : ks);
  const uint64_t cap = (uint64_t)1 << S;
  const uint64_t vq = v < cap ? v : cap;
  q = (valid && ks >= 0) ? vq : 0;
}

#if defined(__CUDACC__)
// det_exp_quant for the streaming sum kernel, x = logw - max <= 0 (or -inf / NaN): the same q bit for bit;
// e differs from det_exp only below x = -700, where it is exp(-700) ~ 1e-304 instead of 0 (it only
// enters the sums of SPEC §6, compared at 1e-10).  The x < -700 / -inf cases are clamped with one
// integer compare on the high word, NaN is carried by one predicate.
__device__ __forceinline__ void det_exp_quant_stream(double x, int S, double& e, uint64_t& q) {
  const uint32_t hi = (uint32_t)(double_as_u64(x) >> 32);
  const bool isnan = (x != x);
  const double xc = (hi >= 0xC085E000u) ? -700.0 : x;  // x <= -700, -inf (and negative-sign NaN: handled below)
  double p;
  int k;
  det_exp_parts(xc, p, k);
  const double ev = scale_pow2(p, k);
  e = isnan ? x : ev;
  const int ks = k + S;
  // SPEC §3 caps q at 2^S.  For x <= 0 the cap never binds: p 2^k <= 1 (k <= -1 gives p 2^k <= 1.4143 / 2; k = 0 means
  // r = x in (-0.347, 0], where p = 1 + (r + r^2 E) with r + r^2 E <= 0), so the truncation is <= 2^S already — and x is
  // logw - max(logw) with the EXACT max, never positive.
  const uint64_t v = (uint64_t)scale_pow2(p, ks < 0 ? 0 : ks);
  q = (ks >= 0 && !isnan) ? v : 0;
}
#endif

SMCB_HD double det_log(double u) {
  uint64_t b = double_as_u64(u);
  int e = (int)((b >> 52) & 0x7FF) - 1023;
  double m = u64_as_double((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
  if (m > kMathC[5]) {
    m *= 0.5;
    e += 1;
  }
  double f = m - 1.0;
  double s = f / (2.0 + f);
  double z = s * s;
  double r = kLogC[0];
#pragma unroll
  for (int i = 1; i < 7; ++i) r = fma(r, z, kLogC[i]);
  double lm = fma(s * z, r, s + s);
  double ef = (double)e;
  return fma(ef, kMathC[1], fma(ef, kMathC[2], lm));
}

// sin(2 pi u), cos(2 pi u), u in [0,1)
SMCB_HD void det_sincos2pi(double u, double& sn, double& cs) {
  double a = 4.0 * u;
  const double tn = a + kMathC[0];
  double nf = tn - kMathC[0];
  double r = a - nf;
  int n = (int)(uint32_t)double_as_u64(tn) & 3;   // low mantissa bits of a + 1.5*2^52 = round(a)
  double z = r * r;
  double s = kSinC[0];
#pragma unroll
  for (int i = 1; i < 7; ++i) s = fma(s, z, kSinC[i]);
  double sr = r * s;
  double c = kCosC[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) c = fma(c, z, kCosC[i]);
  double s1 = (n & 1) ? c : sr;
  double c1 = (n & 1) ? sr : c;
  // n=0:(sr,cr) 1:(cr,-sr) 2:(-sr,-cr) 3:(-cr,sr)
  sn = (n & 2) ? -s1 : s1;
  cs = ((n + 1) & 2) ? -c1 : c1;
}

// 52-bit uniforms of a Philox block as exact doubles in (0,1): (2k+1) * 2^-53
SMCB_HD double u52_open(uint32_t hi, uint32_t lo) {
  uint64_t k = ((uint64_t)hi << 20) | (uint64_t)(lo >> 12);
  // 1.k in [1,2) minus (1 - 2^-53) = k*2^-52 + 2^-53, exact
  return u64_as_double(0x3FF0000000000000ull | k) - 0x1.fffffffffffffp-1;
}

// Box-Muller pair from one Philox block (SPEC §2)
SMCB_HD void normal_pair(const Philox4& b, double& z0, double& z1) {
  double u1 = u52_open(b.r0, b.r1);
  double u2 = u52_open(b.r2, b.r3);
  double rho = sqrt(-2.0 * det_log(u1));
  double sn, cs;
  det_sincos2pi(u2, sn, cs);
  z0 = rho * cs;
  z1 = rho * sn;
}

SMCB_HD void normal_pair_at(const RngKey& key, uint32_t pair, uint32_t stream, uint32_t t,
                            uint32_t kind, uint32_t comp, double& z0, double& z1) {
  Philox4 b = philox4x32_10(pair, stream, t, purpose_word(kind, comp, key.epoch), key);
  normal_pair(b, z0, z1);
}

SMCB_HD uint64_t uniform64_of(const Philox4& b, uint32_t i) {
  return (i & 1) ? (((uint64_t)b.r2 << 32) | b.r3) : (((uint64_t)b.r0 << 32) | b.r1);
}

SMCB_HD uint64_t uniform64_at(const RngKey& key, uint32_t i, uint32_t stream, uint32_t t,
                              uint32_t kind) {
  Philox4 b = philox4x32_10(i >> 1, stream, t, purpose_word(kind, 0, key.epoch), key);
  return uniform64_of(b, i);
}

// ---------------------------------------------------------------------------------------------
// resampling thresholds (SPEC §5)
enum : int { RESAMPLE_MULTINOMIAL = 0, RESAMPLE_STRATIFIED = 1, RESAMPLE_SYSTEMATIC = 2 };

SMCB_HD int quant_shift(uint64_t n) {  // S = 61 - ceil(log2 n)
  int c = 0;
  while (((uint64_t)1 << c) < n) ++c;
  return 61 - c;
}
SMCB_HD uint64_t strata_width(uint64_t n) { return 0xFFFFFFFFFFFFFFFFull / n; }

}  // namespace smcb
