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

// kinds 0..2 run on every engine; kinds 3..5 = MultivariateLinearGaussian with a d = 2, 3, 4 dimensional state and a scalar
// observation (state_space_models.jl:137-189, hodrick_prescott :193-202): the large-N single filter only
enum : int { KIND_LG1D = 0, KIND_SV = 1, KIND_UCSV = 2, KIND_COUNT = 3, KIND_MVLG2 = 3, KIND_MVLG3 = 4, KIND_MVLG4 = 5, KIND_ALL = 6 };
constexpr int kParamStride = 8;  // doubles per θ in both the raw and the derived block

struct Derived {
  double d[kParamStride];
};

struct DerivedMV {  // by-value kernel argument of the multivariate kinds
  double d[64];
};
constexpr int kMvStride = 64;

SMCB_HD int state_dim(int kind) { return kind >= KIND_MVLG2 ? kind - 1 : (kind == KIND_UCSV ? 3 : 1); }
SMCB_HD bool is_mv_kind(int kind) { return kind >= KIND_MVLG2 && kind < KIND_ALL; }

// lower Cholesky factor (row-major d×d) of a symmetric positive SEMI-definite matrix by the plain Cholesky–Banachiewicz recursion:
// a pivot that is not positive gives a zero column (hodrick_prescott's Q = diag(1/λ, 0), state_space_models.jl:197: the reference's
// MvNormal(A x, Q) would throw PosDefException there; a deterministic second component is what the model means)
SMCB_HD void chol_psd(const double* A, int d, double* L) {
  for (int i = 0; i < d * d; ++i) L[i] = 0.0;
  for (int j = 0; j < d; ++j) {
    double s = A[j * d + j];
    for (int k = 0; k < j; ++k) s = s - L[j * d + k] * L[j * d + k];
    if (!(s > 0.0)) continue;
    const double ljj = sqrt(s);
    L[j * d + j] = ljj;
    for (int i = j + 1; i < d; ++i) {
      double v = A[i * d + j];
      for (int k = 0; k < j; ++k) v = v - L[i * d + k] * L[j * d + k];
      L[i * d + j] = v / ljj;
    }
  }
}

// block (include/smcb200.h, the layout of smcb_kalman_mv_*): A[d][d], B[d], Q[d][d], R, x0[d], Σ0[d][d] row-major  ->
// derived: A[d²] | B[d] | LQ[d²] | ir = 1/sqrt(R) | c = -(det_log(sqrt(R)) + ½ log 2π) | x0[d] | L0[d²]
// R is a VARIANCE as in kalman_filter.jl:16 and in the univariate methods (:102); the multivariate `observation` passes it to Normal
// un-square-rooted (:178), an inconsistency of the reference ruled a defect — the filter then targets the
// likelihood the reference's own Kalman filter computes.
SMCB_HD void derive_params_mv(int d, const double* P, double* D) {
  for (int i = 0; i < kMvStride; ++i) D[i] = 0.0;
  const double* A = P;
  const double* B = P + d * d;
  const double* Q = B + d;
  const double R = Q[d * d];
  const double* x0 = Q + d * d + 1;
  const double* S0 = x0 + d;
  double* o = D;
  for (int i = 0; i < d * d; ++i) o[i] = A[i];
  o += d * d;
  for (int i = 0; i < d; ++i) o[i] = B[i];
  o += d;
  chol_psd(Q, d, o);
  o += d * d;
  const double sr = sqrt(R);
  o[0] = 1.0 / sr;
  o[1] = -(det_log(sr) + SMCB_HALF_LOG_2PI);
  o += 2;
  for (int i = 0; i < d; ++i) o[i] = x0[i];
  o += d;
  chol_psd(S0, d, o);
}

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
  using DV = Derived;
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
  using DV = Derived;
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
  using DV = Derived;
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

// MultivariateLinearGaussian (state_space_models.jl:137-189): x' ~ MvNormal(A x, Q), y ~ Normal(B x, R), x1 ~ MvNormal(x0, Σ0).
// docs/SPEC.md §4b: x'_i = fma-chain of A[i][·] x then of LQ[i][0..i] z; log-weight from v = (y − B x') / sqrt(R).
template <int DIM>
struct ModelMVLG {
  using DV = DerivedMV;
  static constexpr int KIND = KIND_MVLG2 + DIM - 2;
  static constexpr int D = DIM;
  double A[DIM][DIM], B[DIM], LQ[DIM][DIM], ir, c, x0[DIM], L0[DIM][DIM];
  SMCB_HD void load(const double* d) {
    const double* o = d;
    for (int i = 0; i < DIM; ++i)
      for (int j = 0; j < DIM; ++j) A[i][j] = o[i * DIM + j];
    o += DIM * DIM;
    for (int i = 0; i < DIM; ++i) B[i] = o[i];
    o += DIM;
    for (int i = 0; i < DIM; ++i)
      for (int j = 0; j < DIM; ++j) LQ[i][j] = o[i * DIM + j];
    o += DIM * DIM;
    ir = o[0];
    c = o[1];
    o += 2;
    for (int i = 0; i < DIM; ++i) x0[i] = o[i];
    o += DIM;
    for (int i = 0; i < DIM; ++i)
      for (int j = 0; j < DIM; ++j) L0[i][j] = o[i * DIM + j];
  }
  SMCB_HD void init(const double* z, double* x) const {
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      double acc = x0[i];
#pragma unroll
      for (int j = 0; j <= i; ++j) acc = fma(L0[i][j], z[j], acc);
      x[i] = acc;
    }
  }
  SMCB_HD void transition(const double* z, const double* xp, double* x) const {
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      double acc = A[i][0] * xp[0];
#pragma unroll
      for (int j = 1; j < DIM; ++j) acc = fma(A[i][j], xp[j], acc);
#pragma unroll
      for (int j = 0; j <= i; ++j) acc = fma(LQ[i][j], z[j], acc);
      x[i] = acc;
    }
  }
  SMCB_HD double logweight(const double* x, double y) const {
    double m = B[0] * x[0];
#pragma unroll
    for (int j = 1; j < DIM; ++j) m = fma(B[j], x[j], m);
    const double v = (y - m) * ir;
    return fma(-0.5 * v, v, c);
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

template <>
struct TransDensity<ModelUCSV> {  // (the UCSV move of SPEC §10b derives its densities from the state: nothing per θ)
  SMCB_HD void load(const double*) {}
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

// UCSV (docs/SPEC.md §10b): the log-volatilities move by the transition, the trend by the conditionally optimal Gaussian move
// tempered by κ in [0, 1]: with σε = exp(le/2) of the PARENT (state_space_models.jl:238) and ση'² = exp(ln') of the NEW state (:246),
// r = σε²/ση'², g = κ r/(1 + r): x' ~ N(x + g (y − x), (1 − g) σε²).  κ = 0 is the bootstrap move, κ = 1 p(x' | x, le, ln', y).
// logw = logpdf(observation(x'), y) + logpdf(transition(xp), x') − logpdf(proposal(xp), x')   (particles.jl:73-78)
SMCB_HD double guided_move_ucsv(const ModelUCSV& mdl, double kappa, const double* z, const double* xp, double y, double* x) {
  x[1] = fma(mdl.ge, z[1], xp[1]);
  x[2] = fma(mdl.gn, z[2], xp[2]);
  const double sd = det_exp(0.5 * xp[1]);
  const double r = (sd * sd) * det_exp(-x[2]);
  const double g = kappa * (r / (1.0 + r));
  const double omg = 1.0 - g;
  const double mq = fma(g, y - xp[0], xp[0]);
  const double c2 = sd * sqrt(omg);
  x[0] = fma(c2, z[0], mq);
  const double zt = (x[0] - xp[0]) / sd;
  const double corr = fma(-0.5 * zt, zt, 0.5 * (z[0] * z[0])) + 0.5 * det_log(omg);
  return mdl.logweight(x, y) + corr;
}

}  // namespace smcb
