// State-space models as device functors (docs/SPEC.md §4).
//
//   LG1D  <- LinearModel univariate methods  /root/reference/src/state_space_models.jl:74-109
//            (also unobserved_components :119-128, which is LG1D with A=B=1, Q=σε, R=ση, σ0=σε)
//   UCSV  <- UCSV methods                    /root/reference/src/state_space_models.jl:215-259
//   SV    <- absent from the reference's src/ (SURVEY.md F6); canonical stochastic volatility
//
// The reference rebuilds Normal(μ, sqrt(Q)) and re-takes log σ for every particle (:93,:102); here
// the square roots, 1/σ and log σ are derived once per θ on the host (derive_params) and the per
// particle work is 2-3 fma.
#pragma once
#include "smcb_detmath.cuh"

namespace smcb {

enum : int { KIND_LG1D = 0, KIND_SV = 1, KIND_UCSV = 2, KIND_COUNT = 3 };
constexpr int kParamStride = 8;  // doubles per θ in both the raw and the derived block

struct Derived {
  double d[kParamStride];
};

SMCB_HD int state_dim(int kind) { return kind == KIND_UCSV ? 3 : 1; }

// host + device so that the same derived block is produced wherever it is computed
SMCB_HD void derive_params(int kind, const double* P, double* D) {
  for (int i = 0; i < kParamStride; ++i) D[i] = 0.0;
  if (kind == KIND_LG1D) {
    double sr = sqrt(P[3]);
    D[0] = P[0];
    D[1] = P[1];
    D[2] = sqrt(P[2]);
    D[3] = P[4];
    D[4] = sqrt(P[5]);
    D[5] = 1.0 / sr;
    D[6] = -(det_log(sr) + SMCB_HALF_LOG_2PI);
    D[7] = det_log(D[2]);  // log sd of the transition kernel: guided weights only (SPEC §10)
  } else if (kind == KIND_SV) {
    D[0] = P[0];
    D[1] = P[1];
    D[2] = P[2];
    D[3] = P[2] / sqrt(1.0 - P[1] * P[1]);
    D[4] = det_log(P[2]);  // log sd of the transition kernel: guided weights only (SPEC §10)
  } else {
    D[0] = P[0];
    D[1] = P[1];
    D[2] = P[2];
    D[3] = P[3];
    D[4] = P[4];
    D[5] = det_exp(0.5 * P[3]);
  }
}

struct ModelLG1D {
  static constexpr int KIND = KIND_LG1D;
  static constexpr int D = 1;
  double A, B, sq, x0, s0, ir, c;
  SMCB_HD void load(const double* d) {
    A = d[0]; B = d[1]; sq = d[2]; x0 = d[3]; s0 = d[4]; ir = d[5]; c = d[6];
  }
  SMCB_HD void init(const double* z, double* x) const { x[0] = fma(s0, z[0], x0); }
  SMCB_HD void transition(const double* z, const double* xp, double* x) const {
    x[0] = fma(sq, z[0], A * xp[0]);
  }
  SMCB_HD double logweight(const double* x, double y) const {
    double v = (y - B * x[0]) * ir;
    return fma(-0.5 * v, v, c);
  }
};

struct ModelSV {
  static constexpr int KIND = KIND_SV;
  static constexpr int D = 1;
  double mu, rho, sigma, s0;
  SMCB_HD void load(const double* d) {
    mu = d[0]; rho = d[1]; sigma = d[2]; s0 = d[3];
  }
  SMCB_HD void init(const double* z, double* x) const { x[0] = fma(s0, z[0], mu); }
  SMCB_HD void transition(const double* z, const double* xp, double* x) const {
    x[0] = fma(sigma, z[0], fma(rho, xp[0] - mu, mu));
  }
  SMCB_HD double logweight(const double* x, double y) const {
    return fma(-0.5 * (y * y), det_exp(-x[0]), -(fma(0.5, x[0], SMCB_HALF_LOG_2PI)));
  }
};

struct ModelUCSV {
  static constexpr int KIND = KIND_UCSV;
  static constexpr int D = 3;
  double ge, gn, x0, lse0, lsn0, e0;
  SMCB_HD void load(const double* d) {
    ge = d[0]; gn = d[1]; x0 = d[2]; lse0 = d[3]; lsn0 = d[4]; e0 = d[5];
  }
  SMCB_HD void init(const double* z, double* x) const {
    x[0] = fma(e0, z[0], x0);
    x[1] = fma(ge, z[1], lse0);
    x[2] = fma(gn, z[2], lsn0);
  }
  SMCB_HD void transition(const double* z, const double* xp, double* x) const {
    double sd = det_exp(0.5 * xp[1]);
    x[0] = fma(sd, z[0], xp[0]);
    x[1] = fma(ge, z[1], xp[1]);
    x[2] = fma(gn, z[2], xp[2]);
  }
  SMCB_HD double logweight(const double* x, double y) const {
    double d = y - x[0];
    return fma(-0.5 * (d * d), det_exp(-x[2]), -(fma(0.5, x[2], SMCB_HALF_LOG_2PI)));
  }
};

// ---- guided particle filter (particle_filter! with a proposal, particles.jl:66-80; SPEC §10) ----------------
// The transition kernel of the D = 1 models as a density: mean, sd, log sd.  Kept out of the model structs so
// that the bootstrap kernels carry no extra members.
template <class Model>
struct TransDensity;
template <>
struct TransDensity<ModelLG1D> {
  double isd, lsd;  // 1 / sd (one IEEE division per thread, not per particle), log sd
  SMCB_HD void load(const double* d) { isd = 1.0 / d[2]; lsd = d[7]; }
  SMCB_HD double mean(const ModelLG1D& m, double xp) const { return m.A * xp; }
};
template <>
struct TransDensity<ModelSV> {
  double isd, lsd;
  SMCB_HD void load(const double* d) { isd = 1.0 / d[2]; lsd = d[4]; }
  SMCB_HD double mean(const ModelSV& m, double xp) const { return fma(m.rho, xp - m.mu, m.mu); }
};

constexpr int kProposalStride = 5;  // c0, c1, c2, det_log(c2), 1 / c2: the proposal x' ~ N(c0 + c1 xp, c2^2) of one (t, θ)

struct ProposalCoef {  // by-value kernel argument of the grid-wide guided step
  double c[kProposalStride];
};

// logpdf(transition(xp), x') - logpdf(proposal(xp), x')   (particles.jl:77-78; the two 0.5 log 2π cancel); mq is the
// proposal mean fma(c1, xp, c0) that produced x'
template <class Model>
SMCB_HD double guided_correction(const Model& mdl, const TransDensity<Model>& f, const double* pc, double mq, double xp, double x) {
  const double zt = (x - f.mean(mdl, xp)) * f.isd;
  const double zq = (x - mq) * pc[4];
  const double lf = fma(-0.5 * zt, zt, -f.lsd);
  const double lq = fma(-0.5 * zq, zq, -pc[3]);
  return lf - lq;
}

// x' = rand(proposal(xp)); logw = logpdf(observation(x'), y) + logpdf(transition(xp), x') - logpdf(proposal(xp), x')
// (particles.jl:73-78).  z is the standard normal the bootstrap transition would consume.
template <class Model>
SMCB_HD double guided_move(const Model& mdl, const TransDensity<Model>& f, const double* pc, double z, double xp, double y,
                           double* x) {
  const double mq = fma(pc[1], xp, pc[0]);
  x[0] = fma(pc[2], z, mq);
  return mdl.logweight(x, y) + guided_correction(mdl, f, pc, mq, xp, x[0]);
}

}  // namespace smcb
