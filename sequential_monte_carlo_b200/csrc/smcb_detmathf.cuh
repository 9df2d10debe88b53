// The binary32 ARITHMETIC tier (docs/SPEC.md §9b; BASELINE.json north star: "fp32 ... float4-vectorised", tolerance 1e-4):
// Philox-addressed normals four per block, deterministic expf / logf / sincos2pif as fixed sequences of IEEE-754 binary32
// operations (only the written fmaf fuse: nvcc -fmad=false, gcc -ffp-contract=off), float model functors.  The oracle restates
// every function independently (oracle/det_math.h) and the two are compared bit for bit.  The weights are still quantised to
// the uint64 fixed point of SPEC §5, so CDF, thresholds and ancestors are exactly those of the binary64 tiers' machinery.
// Coefficients: tools/gen_coeffs_f32.py (Chebyshev fits at 200 bits, rounded once to binary32).
#pragma once
#include "smcb_detmath.cuh"
#include "smcb_models.cuh"

namespace smcb {

SMCB_HD float u32_as_float(uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}
SMCB_HD uint32_t float_as_u32(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t b;
  memcpy(&b, &f, 4);
  return b;
#endif
}

#define SMCB_MAGICF 0x1.8p23f
#define SMCB_LN2F_HI 0x1.62e4p-1f
#define SMCB_LN2F_LO 0x1.7f7d1cp-20f
#define SMCB_LOG2EF 0x1.715476p+0f
#define SMCB_HALF_LOG_2PIF 0x1.d67f1cp-1f
#define SMCB_SQRT2F 0x1.6a09e6p+0f

#if defined(__CUDA_ARCH__)
#define SMCB_COEFF_DECL static __constant__ float
#else
#define SMCB_COEFF_DECL static const float
#endif
// highest degree first
SMCB_COEFF_DECL kExpFC[5] = {0x1.6d110ap-10f, 0x1.120b6ep-7f, 0x1.55551ap-5f, 0x1.5554dcp-3f, 0x1.0p-1f};
SMCB_COEFF_DECL kLogFC[3] = {0x1.2ee656p-2f, 0x1.997c2ep-2f, 0x1.55555cp-1f};
SMCB_COEFF_DECL kSinFC[4] = {-0x1.2d9b1ep-8f, 0x1.465ec2p-4f, -0x1.4abbbap-1f, 0x1.921fb6p+0f};
SMCB_COEFF_DECL kCosFC[4] = {-0x1.4e5dd4p-6f, 0x1.03af58p-2f, -0x1.3bd392p+0f, 0x1.0p+0f};

// exp(x) = p 2^k, p in about [0.707, 1.415]
SMCB_HD void det_expf_parts(float x, float& p, int& k) {
  const float t = x * SMCB_LOG2EF + SMCB_MAGICF;   // the low mantissa bits of t hold round(x log2 e)
  const float kf = t - SMCB_MAGICF;
  k = (int)(float_as_u32(t) & 0x7FFFFFu) - 0x400000;
  float r = fmaf(-kf, SMCB_LN2F_HI, x);
  r = fmaf(-kf, SMCB_LN2F_LO, r);
  float e = kExpFC[0];
#pragma unroll
  for (int i = 1; i < 5; ++i) e = fmaf(e, r, kExpFC[i]);
  p = 1.0f + fmaf(r * r, e, r);
}
SMCB_HD float scale_pow2f(float p, int n) { return u32_as_float(float_as_u32(p) + ((uint32_t)n << 23)); }

SMCB_HD float det_expf(float x) {
  if (x < -86.0f) return 0.0f;
  if (x > 87.0f) return INFINITY;
  float p;
  int k;
  det_expf_parts(x, p, k);
  return scale_pow2f(p, k);
}

// e = exp(x) and q = min(trunc(exp(x) 2^S), 2^S) for x <= 0, the fixed-point weight of SPEC §5 from a binary32 exponential:
// p = M 2^(ep - 23) with the 24-bit integer mantissa M, so p 2^(k + S) is an integer shift of M (no floating-point scaling)
SMCB_HD void det_exp_quantf(float x, int S, float& e, uint64_t& q) {
  const bool valid = (x >= -86.0f);  // false for x < -86, -inf and NaN
  const float xc = valid ? x : -86.0f;
  float p;
  int k;
  det_expf_parts(xc, p, k);
  e = valid ? scale_pow2f(p, k) : ((x < -86.0f) ? 0.0f : x);  // NaN propagates into the sums; q = 0
  const uint32_t pb = float_as_u32(p);
  const uint64_t M = (uint64_t)((pb & 0x7FFFFFu) | 0x800000u);
  const int sh = (int)(pb >> 23) - 127 - 23 + k + S;
  const uint64_t v = sh >= 0 ? (M << (sh > 40 ? 40 : sh)) : (sh > -64 ? (M >> (-sh)) : 0ull);
  const uint64_t cap = (uint64_t)1 << S;
  q = valid ? (v < cap ? v : cap) : 0ull;
}

SMCB_HD float det_logf(float u) {  // u positive, normal
  const uint32_t b = float_as_u32(u);
  int e = (int)((b >> 23) & 0xFF) - 127;
  float m = u32_as_float((b & 0x007FFFFFu) | 0x3F800000u);
  if (m > SMCB_SQRT2F) {
    m *= 0.5f;
    e += 1;
  }
  const float f = m - 1.0f;
  const float s = f / (2.0f + f);
  const float z = s * s;
  float r = kLogFC[0];
#pragma unroll
  for (int i = 1; i < 3; ++i) r = fmaf(r, z, kLogFC[i]);
  const float lm = fmaf(s * z, r, s + s);
  const float ef = (float)e;
  return fmaf(ef, SMCB_LN2F_HI, fmaf(ef, SMCB_LN2F_LO, lm));
}

// sin(2 pi u), cos(2 pi u), u in [0,1)
SMCB_HD void det_sincos2pif(float u, float& sn, float& cs) {
  const float a = 4.0f * u;
  const float tn = a + SMCB_MAGICF;
  const float nf = tn - SMCB_MAGICF;
  const float r = a - nf;
  const int n = (int)(float_as_u32(tn) & 3u);
  const float z = r * r;
  float s = kSinFC[0];
#pragma unroll
  for (int i = 1; i < 4; ++i) s = fmaf(s, z, kSinFC[i]);
  const float sr = r * s;
  float c = kCosFC[0];
#pragma unroll
  for (int i = 1; i < 4; ++i) c = fmaf(c, z, kCosFC[i]);
  const float s1 = (n & 1) ? c : sr;
  const float c1 = (n & 1) ? sr : c;
  sn = (n & 2) ? -s1 : s1;
  cs = ((n + 1) & 2) ? -c1 : c1;
}

// Four standard normals from ONE Philox block (SPEC §9b): particles 4q .. 4q+3 use the block at index q; words (r0, r1) give
// the Box-Muller pair of particles 4q, 4q+1, words (r2, r3) that of 4q+2, 4q+3; u = (2 (r >> 9) + 1) 2^-24 in (0, 1), exact.
SMCB_HD void normal_quadf(const Philox4& b, float z[4]) {
  const uint32_t w[4] = {b.r0, b.r1, b.r2, b.r3};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = (float)(2u * (w[2 * h] >> 9) + 1u) * 0x1p-24f;
    const float u2 = (float)(2u * (w[2 * h + 1] >> 9) + 1u) * 0x1p-24f;
    const float rho = sqrtf(-2.0f * det_logf(u1));
    float sn, cs;
    det_sincos2pif(u2, sn, cs);
    z[2 * h] = rho * cs;
    z[2 * h + 1] = rho * sn;
  }
}
SMCB_HD void normal_quadf_at(const RngKey& key, uint32_t quad, uint32_t stream, uint32_t t, uint32_t kind, uint32_t comp, float z[4]) {
  const Philox4 b = philox4x32_10(quad, stream, t, purpose_word(kind, comp, key.epoch), key);
  normal_quadf(b, z);
}

// ---- float model functors: the derived block of smcb_models.cuh rounded once to binary32 (SPEC §9b) -------------------------
struct DerivedF {
  float d[kParamStride];
};
SMCB_HD void derive_params_f(int kind, const double* P, float* Df) {
  double D[kParamStride];
  derive_params(kind, P, D);
  for (int i = 0; i < kParamStride; ++i) Df[i] = (float)D[i];
}

struct ModelLG1Df {
  static constexpr int KIND = KIND_LG1D;
  static constexpr int D = 1;
  float A, B, sq, x0, s0, ir, c;
  SMCB_HD void load(const float* d) { A = d[0]; B = d[1]; sq = d[2]; x0 = d[3]; s0 = d[4]; ir = d[5]; c = d[6]; }
  SMCB_HD void init(const float* z, float* x) const { x[0] = fmaf(s0, z[0], x0); }
  SMCB_HD void transition(const float* z, const float* xp, float* x) const { x[0] = fmaf(sq, z[0], A * xp[0]); }
  SMCB_HD float logweight(const float* x, float y) const {
    const float v = (y - B * x[0]) * ir;
    return fmaf(-0.5f * v, v, c);
  }
};
struct ModelSVf {
  static constexpr int KIND = KIND_SV;
  static constexpr int D = 1;
  float mu, rho, sigma, s0;
  SMCB_HD void load(const float* d) { mu = d[0]; rho = d[1]; sigma = d[2]; s0 = d[3]; }
  SMCB_HD void init(const float* z, float* x) const { x[0] = fmaf(s0, z[0], mu); }
  SMCB_HD void transition(const float* z, const float* xp, float* x) const { x[0] = fmaf(sigma, z[0], fmaf(rho, xp[0] - mu, mu)); }
  SMCB_HD float logweight(const float* x, float y) const {
    return fmaf(-0.5f * (y * y), det_expf(-x[0]), -(fmaf(0.5f, x[0], SMCB_HALF_LOG_2PIF)));
  }
};
struct ModelUCSVf {
  static constexpr int KIND = KIND_UCSV;
  static constexpr int D = 3;
  float ge, gn, x0, lse0, lsn0, e0;
  SMCB_HD void load(const float* d) { ge = d[0]; gn = d[1]; x0 = d[2]; lse0 = d[3]; lsn0 = d[4]; e0 = d[5]; }
  SMCB_HD void init(const float* z, float* x) const {
    x[0] = fmaf(e0, z[0], x0);
    x[1] = fmaf(ge, z[1], lse0);
    x[2] = fmaf(gn, z[2], lsn0);
  }
  SMCB_HD void transition(const float* z, const float* xp, float* x) const {
    const float sd = det_expf(0.5f * xp[1]);
    x[0] = fmaf(sd, z[0], xp[0]);
    x[1] = fmaf(ge, z[1], xp[1]);
    x[2] = fmaf(gn, z[2], xp[2]);
  }
  SMCB_HD float logweight(const float* x, float y) const {
    const float d = y - x[0];
    return fmaf(-0.5f * (d * d), det_expf(-x[2]), -(fmaf(0.5f, x[2], SMCB_HALF_LOG_2PIF)));
  }
};

}  // namespace smcb
