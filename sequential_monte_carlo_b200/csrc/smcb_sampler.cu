// Device-resident θ-level samplers (see smcb_sampler.cuh).  The M-length control vectors never leave the GPU: per
// smc²! step the host enqueues one batched filter step, one in-place ncclAllGather of M/G doubles, one single-CTA
// kernel (logω += logμ, logZ += logμ, normalise, ESS) and reads back 64 bytes.  The arithmetic of the θ level is
// frozen in docs/SPEC.md §11 so that the oracle (oracle/samplers.py) reproduces θ bit for bit.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "smcb_sampler.cuh"

namespace smcb {

namespace {

constexpr int kTB = 1024;  // threads of the single-CTA θ kernels

// ---- block-wide helpers (one CTA of kTB threads; sh = 32 doubles of shared scratch) ---------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0;
  return warp_sum(r);
}
__device__ __forceinline__ double block_max(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : -INFINITY;
  return warp_max(r);
}

// normalize(logω) (particles.jl:5-15, called as `reweight` at the θ level: SURVEY F3) for lw[0..M) in shared memory.
// Writes ω; returns ess and log Σ exp(lw) (block-uniform).
__device__ void normalize_block(const double* lw, int M, double* __restrict__ omega, double* sh, double& ess, double& logsum) {
  double mx = -INFINITY;
  for (int i = threadIdx.x; i < M; i += blockDim.x) mx = lw[i] > mx ? lw[i] : mx;
  mx = block_max(mx, sh);
  double se = 0.0, se2 = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const double e = det_exp(lw[i] - mx);
    se += e;
    se2 += e * e;
  }
  se = block_sum(se, sh);
  se2 = block_sum(se2, sh);
  if (omega)
    for (int i = threadIdx.x; i < M; i += blockDim.x) omega[i] = det_exp(lw[i] - mx) / se;
  ess = (se * se) / se2;
  logsum = mx + log(se);
}

// ---- prior: a product of univariate laws (priors.py; README.md:81-85, examples/inflation_example.jl:234-239) -------------
enum : int { PRIOR_NORMAL = 0, PRIOR_LOGNORMAL = 1, PRIOR_UNIFORM = 2, PRIOR_TRUNCNORMAL = 3 };
#define SMCB_PRIOR_HALF_LOG_2PI 0x1.d67f1c864beb4p-1  // 0.5 * log(2π) as the host languages round it (priors.py; one ulp below SPEC §3's)

// insupport(prior, θ) and logpdf(prior, θ)   smc_samplers.jl:116,123-126.  row = kind, p0, p1, lo, hi, c0, c1:
//   Normal(μ, σ): p0 = μ, p1 = σ, c0 = log σ;  LogNormal(μ, σ): the same on log x;  Uniform(a, b): lo, hi, c0 = -log(b - a);
//   TruncatedNormal(μ, σ, lo, hi): c0 = log σ, c1 = log of the mass of [lo, hi]
SMCB_HD bool prior_eval(const PriorTable& pr, const double* th, double& lp) {
  bool ok = true;
  double s = 0.0;
  for (int k = 0; k < pr.d; ++k) {
    const double* r = pr.row[k];
    const int kind = (int)r[0];
    const double x = th[k];
    double l = 0.0;
    bool in = true;
    if (kind == PRIOR_NORMAL) {
      in = isfinite(x);
      const double z = (x - r[1]) / r[2];
      l = -0.5 * z * z - r[5] - SMCB_PRIOR_HALF_LOG_2PI;
    } else if (kind == PRIOR_LOGNORMAL) {
      in = isfinite(x) && x > 0.0;
      const double lx = log(in ? x : 1.0);
      const double z = (lx - r[1]) / r[2];
      l = -lx - r[5] - SMCB_PRIOR_HALF_LOG_2PI - 0.5 * z * z;
    } else if (kind == PRIOR_UNIFORM) {
      in = (x >= r[3]) && (x <= r[4]);
      l = r[5];
    } else {
      in = (x >= r[3]) && (x <= r[4]);
      const double z = (x - r[1]) / r[2];
      l = -0.5 * z * z - r[5] - SMCB_PRIOR_HALF_LOG_2PI - r[6];
    }
    ok = ok && in;
    s = s + l;
  }
  lp = ok ? s : -INFINITY;
  return ok;
}

// model(θ): the parameter block of the state-space model of one θ-particle, then its derived block (smcb_models.cuh)
SMCB_HD void params_from_theta(int kind, const ParamMap& map, const double* th, double* derived) {
  double P[kParamStride];
  for (int k = 0; k < kParamStride; ++k) P[k] = map.src[k] >= 0 ? th[map.src[k]] : map.cst[k];
  derive_params(kind, P, derived);
}

// lp[m] = logpdf(prior, θ_m), derived[m] = derive(model(θ_m)) for the initial θ-particles
__global__ void theta_prepare_kernel(const double* __restrict__ theta, int M, int d, int kind, PriorTable prior, ParamMap map,
                                     double* __restrict__ lp, double* __restrict__ derived) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double th[kMaxThetaDim];
  for (int k = 0; k < d; ++k) th[k] = theta[m * d + k];
  double l;
  prior_eval(prior, th, l);
  lp[m] = l;
  params_from_theta(kind, map, th, derived + (int64_t)m * kParamStride);
}

// mode 0: smc²      logZ = logμ;           ω, ess = reweight(logμ)                       smc_samplers.jl:297-298
// mode 1: smc²!     logω = log ω + logμ;   logZ += logμ;  ω, ess = reweight(logω)        :324,333-338
// mode 2: exchange! ω, ess = reweight(new_logZ − logZ);  logZ = new_logZ                 :183-185
__global__ void __launch_bounds__(kTB) theta_step_kernel(const double* __restrict__ inc, double* __restrict__ omega,
                                                         double* __restrict__ logz, ThetaScalars* __restrict__ scal, int M, int mode) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* lw = reinterpret_cast<double*>(smem_raw);
  __shared__ double sh[32];
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const double v = inc[i];
    if (mode == 0) {
      lw[i] = v;
      logz[i] = v;
    } else if (mode == 1) {
      lw[i] = log(omega[i]) + v;
      logz[i] = logz[i] + v;
    } else {
      lw[i] = v - logz[i];
      logz[i] = v;
    }
  }
  __syncthreads();
  double ess, logsum;
  normalize_block(lw, M, omega, sh, ess, logsum);
  if (threadIdx.x == 0) {
    scal->ess = ess;
    scal->logsum = logsum;
  }
}

// The tempering bisection of density_tempered (smc_samplers.jl:237-266) for one stage, entirely on the device:
// find ξ' in (ξ, 2] with ESS((ξ' − ξ)·logZ) ≈ ess_min, clamp to 1 on the last stage.
__global__ void __launch_bounds__(kTB) theta_bisect_kernel(const double* __restrict__ logz, double* __restrict__ omega,
                                                           ThetaScalars* __restrict__ scal, int M, double xi, double ess_min) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* lw = reinterpret_cast<double*>(smem_raw);
  __shared__ double sh[32];
  double lower = xi, upper = 2.0, newxi = xi, ess = 0.0, logsum = 0.0;
  const double oldxi = xi;
  int flag = 1;
  while (upper - lower > 1.0e-6) {
    newxi = (upper + lower) / 2.0;
    __syncthreads();
    for (int i = threadIdx.x; i < M; i += blockDim.x) lw[i] = (newxi - oldxi) * logz[i];
    __syncthreads();
    normalize_block(lw, M, nullptr, sh, ess, logsum);
    if (ess == ess_min) break;
    else if (ess < ess_min) upper = newxi;
    else lower = newxi;
  }
  if (newxi >= 1.0) {  // corner solution: the last stage never resamples                  :261-266
    flag = 0;
    newxi = 1.0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < M; i += blockDim.x) lw[i] = (newxi - oldxi) * logz[i];
  __syncthreads();
  normalize_block(lw, M, omega, sh, ess, logsum);
  if (threadIdx.x == 0) {
    scal->ess = ess;
    scal->xi = newxi;
    scal->logsum = logsum;
    scal->resample_flag = flag;
  }
}

// exclusive block scan of per-thread totals; returns the exclusive prefix of this thread and the grand total
__device__ __forceinline__ unsigned long long block_excl_scan_u64(unsigned long long v, unsigned long long* shw, unsigned long long& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned long long inc = warp_scan_u64(v, lane);
  __syncthreads();
  if (lane == 31) shw[warp] = inc;
  __syncthreads();
  const unsigned long long wv = (lane < nw) ? shw[lane] : 0ull;
  const unsigned long long winc = warp_scan_u64(wv, lane);
  const unsigned long long wexcl = __shfl_sync(kFullMask, winc - wv, warp);
  total = __shfl_sync(kFullMask, winc, 31);
  return wexcl + (inc - v);
}

// a = resample(ω) then sort!(a)   (smc_samplers.jl:75, docs/SPEC.md §5/§5b): fixed-point CDF of ω in shared memory, one
// threshold per slot, a counting sort of the ancestors (offspring counts -> prefix sum -> runs of equal parents).
__global__ void __launch_bounds__(kTB) theta_resample_kernel(const double* __restrict__ omega, int M, int resampler, RngKey key, uint32_t t,
                                                             int32_t* __restrict__ anc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* cdf = reinterpret_cast<unsigned long long*>(smem_raw);
  int* cnt = reinterpret_cast<int*>(smem_raw + sizeof(unsigned long long) * (size_t)M);
  __shared__ double sh[32];
  __shared__ unsigned long long shw[32];
  const int tid = threadIdx.x;
  const int chunk = (M + kTB - 1) / kTB;
  const int lo = min(tid * chunk, M), hi = min(lo + chunk, M);
  double mx = 0.0;
  for (int i = tid; i < M; i += kTB) mx = omega[i] > mx ? omega[i] : mx;
  mx = block_max(mx, sh);
  mx = mx > 0.0 ? mx : 0.0;
  const int S = quant_shift((uint64_t)M);
  const double scale = u64_as_double((uint64_t)(1023 + S) << 52);
  unsigned long long run = 0;
  for (int i = lo; i < hi; ++i) {
    const double w = omega[i];
    const unsigned long long q = (mx > 0.0 && w > 0.0) ? (unsigned long long)((w / mx) * scale) : 0ull;
    run += q;
    cdf[i] = run;  // thread-local inclusive prefix, offset below
    cnt[i] = 0;
  }
  unsigned long long Q;
  const unsigned long long excl = block_excl_scan_u64(run, shw, Q);
  for (int i = lo; i < hi; ++i) cdf[i] += excl;
  __syncthreads();
  const uint64_t R = strata_width((uint64_t)M);
  const uint64_t U0 = uniform64_at(key, 0u, 0u, t, PURPOSE_THETA_RESAMPLE);
  for (int i = tid; i < M; i += kTB) {
    int a = i;
    if (Q != 0) {
      uint64_t F;
      if (resampler == RESAMPLE_MULTINOMIAL) F = uniform64_at(key, (uint32_t)i, 0u, t, PURPOSE_THETA_RESAMPLE);
      else if (resampler == RESAMPLE_STRATIFIED) F = (uint64_t)i * R + mulhi64(uniform64_at(key, (uint32_t)i, 0u, t, PURPOSE_THETA_RESAMPLE), R);
      else F = (uint64_t)i * R + mulhi64(U0, R);
      const uint64_t tau = mulhi64(F, Q);
      int l = 0, h = M - 1;
      while (l < h) {
        const int mid = (l + h) >> 1;
        if (cdf[mid] <= tau) l = mid + 1;
        else h = mid;
      }
      a = l;
    }
    atomicAdd(&cnt[a], 1);
  }
  __syncthreads();
  unsigned long long c = 0;
  for (int i = lo; i < hi; ++i) c += (unsigned long long)cnt[i];
  unsigned long long tot;
  unsigned long long pos = block_excl_scan_u64(c, shw, tot);
  for (int i = lo; i < hi; ++i) {
    const int n = cnt[i];
    for (int k = 0; k < n; ++k) anc[pos + k] = i;
    pos += n;
  }
}

// θ = θ[a]; logZ = logZ[a]; (log-prior and parameter blocks follow); ω uniform       smc_samplers.jl:78-83
__global__ void theta_gather_kernel(const int32_t* __restrict__ anc, int M, int d, const double* __restrict__ th_in, double* __restrict__ th_out,
                                    const double* __restrict__ lz_in, double* __restrict__ lz_out, const double* __restrict__ lp_in,
                                    double* __restrict__ lp_out, const double* __restrict__ dv_in, double* __restrict__ dv_out,
                                    double* __restrict__ omega) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int a = anc[m];
  for (int k = 0; k < d; ++k) th_out[m * d + k] = th_in[a * d + k];
  lz_out[m] = lz_in[a];
  lp_out[m] = lp_in[a];
  for (int k = 0; k < kParamStride; ++k) dv_out[(int64_t)m * kParamStride + k] = dv_in[(int64_t)a * kParamStride + k];
  omega[m] = 1.0 / (double)M;
}

// θ' = rand(pmmh_kernel(θ_m, scale_c)); insupport(prior, θ'); model(θ')                           smc_samplers.jl:114-117
// θ'_j = θ_j + Σ_{k<=j} z_k L[j][k] (k ascending, docs/SPEC.md §11); z_k = element m of the host-level normal stream k
__global__ void theta_propose_kernel(const double* __restrict__ theta, int M, int d, int kind, CholFactor L, RngKey key, uint32_t c,
                                     PriorTable prior, ParamMap map, double* __restrict__ theta_prop, double* __restrict__ lp_prop,
                                     uint8_t* __restrict__ ok_out, double* __restrict__ derived_prop) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double th[kMaxThetaDim], z[kMaxThetaDim], tp[kMaxThetaDim];
  for (int k = 0; k < d; ++k) {
    th[k] = theta[m * d + k];
    double z0, z1;
    normal_pair_at(key, (uint32_t)(m >> 1), (uint32_t)k, c, PURPOSE_MH_PROPOSAL, 0u, z0, z1);
    z[k] = (m & 1) ? z1 : z0;
  }
  for (int j = 0; j < d; ++j) {
    double acc = z[0] * L.l[j][0];
    for (int k = 1; k <= j; ++k) acc = acc + z[k] * L.l[j][k];
    tp[j] = th[j] + acc;
    theta_prop[m * d + j] = tp[j];
  }
  double lp;
  const bool ok = prior_eval(prior, tp, lp);
  lp_prop[m] = lp;
  ok_out[m] = ok ? 1 : 0;
  params_from_theta(kind, map, ok ? tp : th, derived_prop + (int64_t)m * kParamStride);
}

// acc_ratio = ξ (logZ' − logZ) + logpdf(prior, θ') − logpdf(prior, θ); accept iff finite target and log u < acc_ratio   :123-135
__global__ void theta_accept_kernel(int M, int d, double xi, RngKey key, uint32_t c, const double* __restrict__ theta_prop,
                                    const double* __restrict__ lp_prop, const uint8_t* __restrict__ ok, const double* __restrict__ logz_prop,
                                    const double* __restrict__ derived_prop, double* __restrict__ theta, double* __restrict__ lp,
                                    double* __restrict__ logz, double* __restrict__ derived, uint8_t* __restrict__ accept,
                                    uint8_t* __restrict__ acc_any) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const double zp = logz_prop[m], lpp = lp_prop[m];
  const double ratio = xi * (zp - logz[m]) + (lpp - lp[m]);
  const uint64_t u64 = uniform64_at(key, (uint32_t)m, 0u, c, PURPOSE_MH_ACCEPT);
  const double u = (double)(u64 >> 11) * 0x1.0p-53;
  const bool acc = ok[m] && (zp + lpp > -INFINITY) && (log(u) < ratio);
  accept[m] = acc ? 1 : 0;
  if (acc) {
    logz[m] = zp;
    lp[m] = lpp;
    for (int k = 0; k < d; ++k) theta[m * d + k] = theta_prop[m * d + k];
    for (int k = 0; k < kParamStride; ++k) derived[(int64_t)m * kParamStride + k] = derived_prop[(int64_t)m * kParamStride + k];
    acc_any[m] = 1;
  }
}

// ω[m] = 1.0 (normalised); acc_ratio = Σ acc_array / M                                               :139-142
__global__ void __launch_bounds__(kTB) theta_finish_kernel(int M, const uint8_t* __restrict__ acc_any, double* __restrict__ omega,
                                                           ThetaScalars* __restrict__ scal) {
  __shared__ double sh[32];
  double c = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    c += acc_any[i] ? 1.0 : 0.0;
    omega[i] = 1.0 / (double)M;
  }
  c = block_sum(c, sh);
  if (threadIdx.x == 0) scal->acc_count = c;
}

template <class T>
void dev_alloc(T*& p, size_t n) {
  SMCB_CUDA_TRY(cudaMalloc(&p, sizeof(T) * n));
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host-only pieces
void make_exchange_plan(const int32_t* parents, int64_t M, int rank, int nranks, ExchangePlan& plan) {
  const int64_t Mloc = M / nranks, lo = (int64_t)rank * Mloc;
  plan.local_parents.resize((size_t)Mloc);
  plan.send_peer.clear(); plan.send_slot.clear(); plan.recv_peer.clear(); plan.recv_slot.clear();
  for (int64_t j = 0; j < Mloc; ++j) {
    const int64_t p = parents[lo + j];
    plan.local_parents[(size_t)j] = (p / Mloc == rank) ? (int32_t)(p - lo) : (int32_t)j;
  }
  // grouped by peer rank, increasing global slot inside a group: both sides enumerate the same order
  for (int peer = 0; peer < nranks; ++peer) {
    if (peer == rank) continue;
    for (int64_t m = peer * Mloc; m < (peer + 1) * Mloc; ++m)  // slots of `peer` whose parent is mine
      if (parents[m] / Mloc == rank) {
        plan.send_peer.push_back(peer);
        plan.send_slot.push_back((int32_t)(parents[m] - lo));
      }
    for (int64_t m = lo; m < lo + Mloc; ++m)  // my slots whose parent lives on `peer`
      if (parents[m] / Mloc == peer) {
        plan.recv_peer.push_back(peer);
        plan.recv_slot.push_back((int32_t)(m - lo));
      }
  }
}

void random_walk_sigma(const double* theta, int64_t M, int d, double* out) {
  // docs/SPEC.md §11: means and cross products accumulated sequentially over m = 0 .. M-1
  double mean[kMaxThetaDim];
  for (int k = 0; k < d; ++k) {
    double s = 0.0;
    for (int64_t m = 0; m < M; ++m) s = s + theta[m * d + k];
    mean[k] = s / (double)M;
  }
  double cov[kMaxThetaDim][kMaxThetaDim];
  for (int j = 0; j < d; ++j)
    for (int k = j; k < d; ++k) {
      double s = 0.0;
      for (int64_t m = 0; m < M; ++m) s = s + (theta[m * d + j] - mean[j]) * (theta[m * d + k] - mean[k]);
      cov[j][k] = cov[k][j] = s / (double)(M - 1);
    }
  const double dth2 = 2.83 * 2.83;  // the reference's constant
  if (d == 1) {  // smc_samplers.jl:87-92: σ used as a standard deviation
    out[0] = std::fabs(cov[0][0]) < 1.0e-8 ? 1.0e-2 : dth2 * cov[0][0] + 1.0e-10;
    return;
  }
  const double dth = dth2 / (double)d;  // :97
  double fro = 0.0;
  for (int j = 0; j < d; ++j)
    for (int k = 0; k < d; ++k) fro = fro + cov[j][k] * cov[j][k];
  const bool tiny = std::sqrt(fro) < 1.0e-8;  // norm(cov) < 1e-8   :98
  for (int j = 0; j < d; ++j)
    for (int k = 0; k < d; ++k) {
      if (tiny) out[j * d + k] = (j == k) ? 1.0e-2 : 0.0;
      else out[j * d + k] = dth * cov[j][k] + ((j == k) ? 1.0e-10 : 0.0);
    }
}

bool cholesky_lower(const double* A, int d, double scale, double* L) {
  for (int i = 0; i < d * d; ++i) L[i] = 0.0;
  for (int j = 0; j < d; ++j) {
    double s = scale * A[j * d + j];
    for (int k = 0; k < j; ++k) s = s - L[j * d + k] * L[j * d + k];
    if (!(s > 0.0)) return false;
    const double ljj = std::sqrt(s);
    L[j * d + j] = ljj;
    for (int i = j + 1; i < d; ++i) {
      double v = scale * A[i * d + j];
      for (int k = 0; k < j; ++k) v = v - L[i * d + k] * L[j * d + k];
      L[i * d + j] = v / ljj;
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------------ ThetaSampler
ThetaSampler::ThetaSampler(int device, cudaStream_t stream, const Comm& comm, const smcb_sampler_config& cfg, const double* theta0)
    : device_(device), stream_(stream), comm_(comm) {
  kind_ = cfg.kind;
  d_ = cfg.d_theta;
  N_ = cfg.N;
  M_ = cfg.M;
  chain_ = cfg.chain;
  resampler_ = cfg.resampler;
  theta_resampler_ = cfg.theta_resampler;
  seed_ = cfg.seed;
  if (kind_ < 0 || kind_ >= KIND_COUNT) throw Error{SMCB_ERR_BAD_ARG, "sampler: unknown model kind"};
  if (d_ < 1 || d_ > kMaxThetaDim) throw Error{SMCB_ERR_BAD_ARG, "sampler: d_theta must be in [1, 8]"};
  if (M_ < 2 || M_ > kMaxThetaParticles) throw Error{SMCB_ERR_BAD_ARG, "sampler: M must be in [2, 16384]"};
  if (M_ % comm_.nranks) throw Error{SMCB_ERR_BAD_ARG, "sampler: M must be divisible by the number of ranks"};
  if (chain_ < 1 || chain_ > 64) throw Error{SMCB_ERR_BAD_ARG, "sampler: chain must be in [1, 64]"};
  if (resampler_ < 0 || resampler_ > 2 || theta_resampler_ < 0 || theta_resampler_ > 2) throw Error{SMCB_ERR_BAD_ARG, "sampler: unknown resampler"};
  if (!theta0) throw Error{SMCB_ERR_BAD_ARG, "sampler: theta0 is null"};
  Mloc_ = M_ / comm_.nranks;
  lo_ = (int64_t)comm_.rank * Mloc_;
  ess_min_ = (double)M_ * cfg.ess_threshold;  // :46
  acc_threshold_ = cfg.min_ar;
  ess_ = (double)M_;
  prior_.d = d_;
  for (int k = 0; k < d_; ++k) {
    const int pk = (int)cfg.prior[k][0];
    if (pk < 0 || pk > 3) throw Error{SMCB_ERR_BAD_ARG, "sampler: unknown prior family"};
    for (int j = 0; j < kPriorStride; ++j) prior_.row[k][j] = cfg.prior[k][j];
  }
  for (int k = 0; k < kParamStride; ++k) {
    if (cfg.map_src[k] >= d_) throw Error{SMCB_ERR_BAD_ARG, "sampler: parameter map refers to a θ component that does not exist"};
    map_.src[k] = cfg.map_src[k];
    map_.cst[k] = cfg.map_const[k];
  }
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  cur_.reset(new BatchFilter(device_, stream_, kind_, Mloc_, N_));
  prop_.reset(new BatchFilter(device_, stream_, kind_, Mloc_, N_));  // the proposals' clouds: allocated here, not inside the first rejuvenation
  if (comm_.active()) {
    // Staging buffers of the cloud exchange, sized once for "every local slot receives / sends one cloud".  They used to grow inside
    // resample() by cudaFree + cudaMalloc whenever a θ-resample moved more clouds than any before it — a driver call that takes
    // process-wide (and, on a shared box, machine-wide) locks and was measured to stall single rejuvenations by 0.1-1.1 s at random
    // (profiles/r2_c5_steps_n4_before_preallocation.jsonl).  A resample that needs more (one parent with children all over the
    // other ranks) still grows them.
    for (int i = 0; i < 2; ++i) {
      const int64_t bytes = Mloc_ * cur_->cloud_bytes();
      SMCB_CUDA_TRY(cudaMalloc(&xbuf_[i], (size_t)bytes));
      xbuf_cap_[i] = bytes;
    }
  }
  const size_t M = (size_t)M_;
  for (int i = 0; i < 2; ++i) {
    dev_alloc(theta_[i], M * d_);
    dev_alloc(logz_[i], M);
    dev_alloc(lp_[i], M);
    dev_alloc(derived_[i], M * kParamStride);
  }
  dev_alloc(omega_, M);
  dev_alloc(theta_prop_, M * d_);
  dev_alloc(lp_prop_, M);
  dev_alloc(derived_prop_, M * kParamStride);
  dev_alloc(logz_prop_, M);
  dev_alloc(logmu_, M);
  dev_alloc(ok_, M);
  dev_alloc(accept_, M);
  dev_alloc(acc_any_, M);
  dev_alloc(anc_, M);
  dev_alloc(slots_dev_, 3 * M);
  dev_alloc(scal_dev_, 1);
  SMCB_CUDA_TRY(cudaMallocHost(&scal_host_, sizeof(ThetaScalars)));
  SMCB_CUDA_TRY(cudaMallocHost(&anc_host_, sizeof(int32_t) * M));
  SMCB_CUDA_TRY(cudaMallocHost(&slots_host_, sizeof(int32_t) * 3 * M));
  SMCB_CUDA_TRY(cudaMallocHost(&theta_host_, sizeof(double) * M * d_));
  std::memset(scal_host_, 0, sizeof(ThetaScalars));
  std::memcpy(theta_host_, theta0, sizeof(double) * M * d_);
  SMCB_CUDA_TRY(cudaMemsetAsync(scal_dev_, 0, sizeof(ThetaScalars), stream_));
  SMCB_CUDA_TRY(cudaMemcpyAsync(theta_[0], theta_host_, sizeof(double) * M * d_, cudaMemcpyHostToDevice, stream_));
  SMCB_CUDA_TRY(cudaMemsetAsync(logz_[0], 0, sizeof(double) * M, stream_));  // :44
  theta_prepare_kernel<<<(unsigned)((M_ + 127) / 128), 128, 0, stream_>>>(theta_[0], (int)M_, d_, kind_, prior_, map_, lp_[0], derived_[0]);
  SMCB_CUDA_TRY(cudaGetLastError());
  ++n_theta_launches_;
  theta_finish_kernel<<<1, kTB, 0, stream_>>>((int)M_, accept_, omega_, scal_dev_);  // ω = 1/M   :39 (the count it writes is not read)
  SMCB_CUDA_TRY(cudaGetLastError());
  ++n_theta_launches_;
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
  SMCB_CUDA_TRY(cudaFuncSetAttribute(theta_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * kMaxThetaParticles)));
  SMCB_CUDA_TRY(cudaFuncSetAttribute(theta_bisect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * kMaxThetaParticles)));
  SMCB_CUDA_TRY(cudaFuncSetAttribute(theta_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(12 * kMaxThetaParticles)));
}

ThetaSampler::~ThetaSampler() {
  cudaSetDevice(device_);
  cudaStreamSynchronize(stream_);
  for (int i = 0; i < 2; ++i) {
    cudaFree(theta_[i]); cudaFree(logz_[i]); cudaFree(lp_[i]); cudaFree(derived_[i]); cudaFree(xbuf_[i]);
  }
  cudaFree(omega_); cudaFree(theta_prop_); cudaFree(lp_prop_); cudaFree(derived_prop_); cudaFree(logz_prop_); cudaFree(logmu_);
  cudaFree(ok_); cudaFree(accept_); cudaFree(acc_any_); cudaFree(anc_); cudaFree(slots_dev_); cudaFree(scal_dev_); cudaFree(y_dev_);
  cudaFreeHost(scal_host_); cudaFreeHost(anc_host_); cudaFreeHost(slots_host_); cudaFreeHost(theta_host_);
  for (auto& m : marks_) { cudaEventDestroy(m.e0); cudaEventDestroy(m.e1); }
  for (auto e : ev_free_) cudaEventDestroy(e);
  for (auto e : span_)
    if (e) cudaEventDestroy(e);
}

void ThetaSampler::mark(int klass, bool start) {
  if (!profiling_) return;
  auto get = [&]() {
    cudaEvent_t e;
    if (!ev_free_.empty()) { e = ev_free_.back(); ev_free_.pop_back(); }
    else SMCB_CUDA_TRY(cudaEventCreate(&e));
    return e;
  };
  if (start) {
    Mark m{klass, get(), get()};
    SMCB_CUDA_TRY(cudaEventRecord(m.e0, stream_));
    marks_.push_back(m);
  } else {
    SMCB_CUDA_TRY(cudaEventRecord(marks_.back().e1, stream_));
  }
}

void ThetaSampler::resolve_marks() {  // call after a stream synchronisation
  for (auto& m : marks_) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, m.e0, m.e1) == cudaSuccess) ms_[m.klass] += ms;
    ev_free_.push_back(m.e0);
    ev_free_.push_back(m.e1);
  }
  marks_.clear();
}

void ThetaSampler::set_data(const double* y, int64_t T) {
  if (!y || T < 1) throw Error{SMCB_ERR_BAD_ARG, "sampler: y must be non-null and T >= 1"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  if (y_cap_ < T) {
    SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
    cudaFree(y_dev_);
    y_dev_ = nullptr;
    y_cap_ = 0;
    dev_alloc(y_dev_, (size_t)T);
    y_cap_ = T;
  }
  SMCB_CUDA_TRY(cudaMemcpyAsync(y_dev_, y, sizeof(double) * T, cudaMemcpyHostToDevice, stream_));
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));  // the caller's buffer is not retained
  T_ = T;
}

void ThetaSampler::all_gather(double* all) {
  if (!comm_.active()) return;
  mark(SK_ALLGATHER, true);
  SMCB_NCCL_TRY(nccl_api().AllGather(all + lo_, all, (size_t)Mloc_, kNcclFloat64, comm_.comm, stream_));
  mark(SK_ALLGATHER, false);
}

void ThetaSampler::read_scalars() {
  SMCB_CUDA_TRY(cudaMemcpyAsync(scal_host_, scal_dev_, sizeof(ThetaScalars), cudaMemcpyDeviceToHost, stream_));
  if (span_open_) SMCB_CUDA_TRY(cudaEventRecord(span_[1], stream_));
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
  ++n_syncs_;
  if (span_open_) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, span_[0], span_[1]) == cudaSuccess) span_ms_ = ms;
  }
  resolve_marks();
}

void ThetaSampler::open_span() {  // device-timeline span of a run: from its first enqueued operation to its last completed one
  for (int i = 0; i < 2; ++i)
    if (!span_[i]) SMCB_CUDA_TRY(cudaEventCreate(&span_[i]));
  SMCB_CUDA_TRY(cudaEventRecord(span_[0], stream_));
  span_open_ = true;
  span_ms_ = 0.0;
}

void ThetaSampler::smc2_init() {
  if (T_ < 1) throw Error{SMCB_ERR_STATE, "sampler: set_data before smc2_init"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  open_span();
  const uint32_t e = next_epoch();
  mark(SK_FILTER, true);
  cur_->init_dev(derived_[tcur_] + lo_ * kParamStride, nullptr, y_dev_, key(e), (uint32_t)lo_, logmu_ + lo_, nullptr);
  mark(SK_FILTER, false);
  n_particle_updates_ += Mloc_ * N_;
  all_gather(logmu_);
  mark(SK_THETA, true);
  theta_step_kernel<<<1, kTB, sizeof(double) * M_, stream_>>>(logmu_, omega_, logz_[tcur_], scal_dev_, (int)M_, 0);
  SMCB_CUDA_TRY(cudaGetLastError());
  ++n_theta_launches_;
  mark(SK_THETA, false);
  read_scalars();
  ess_ = scal_host_->ess;
  started_ = true;
}

void ThetaSampler::smc2_step(int64_t t, double* ess, int* rejuvenated) {
  if (!started_) throw Error{SMCB_ERR_STATE, "sampler: smc2_step before smc2_init / density_tempered"};
  if (t < 1 || t >= T_) throw Error{SMCB_ERR_BAD_ARG, "sampler: t must be in [1, T)"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  int rj = 0;
  if (ess_ < ess_min_) {  // :312
    resample();           // :314
    rejuvenate(t, 1.0);   // rejuvenate!(smc, y[1:t-1])   :317
    exchange(t);          // :320
    rj = 1;
  }
  mark(SK_FILTER, true);
  cur_->step_dev(derived_[tcur_] + lo_ * kParamStride, y_dev_ + t, (uint32_t)t, resampler_, logmu_ + lo_, nullptr);  // :325-331
  mark(SK_FILTER, false);
  n_particle_updates_ += Mloc_ * N_;
  ++n_steps_;
  all_gather(logmu_);
  mark(SK_THETA, true);
  theta_step_kernel<<<1, kTB, sizeof(double) * M_, stream_>>>(logmu_, omega_, logz_[tcur_], scal_dev_, (int)M_, 1);  // :333-338
  SMCB_CUDA_TRY(cudaGetLastError());
  ++n_theta_launches_;
  mark(SK_THETA, false);
  read_scalars();
  ess_ = scal_host_->ess;
  if (ess) *ess = ess_;
  if (rejuvenated) *rejuvenated = rj;
}

void ThetaSampler::resample() {
  const int o = tcur_ ^ 1;
  mark(SK_THETA, true);
  theta_resample_kernel<<<1, kTB, 12 * (size_t)M_, stream_>>>(omega_, (int)M_, theta_resampler_, key(0u), n_resample_, anc_);
  SMCB_CUDA_TRY(cudaGetLastError());
  ++n_theta_launches_;
  ++n_resample_;
  theta_gather_kernel<<<(unsigned)((M_ + 127) / 128), 128, 0, stream_>>>(anc_, (int)M_, d_, theta_[tcur_], theta_[o], logz_[tcur_], logz_[o],
                                                                         lp_[tcur_], lp_[o], derived_[tcur_], derived_[o], omega_);
  SMCB_CUDA_TRY(cudaGetLastError());
  ++n_theta_launches_;
  mark(SK_THETA, false);
  tcur_ = o;
  SMCB_CUDA_TRY(cudaMemcpyAsync(theta_host_, theta_[tcur_], sizeof(double) * M_ * d_, cudaMemcpyDeviceToHost, stream_));
  if (!comm_.active()) {
    mark(SK_EXCHANGE, true);
    cur_->gather_dev(anc_);
    mark(SK_EXCHANGE, false);
    SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));  // θ mirror for the proposal covariance
    ++n_syncs_;
    return;
  }
  SMCB_CUDA_TRY(cudaMemcpyAsync(anc_host_, anc_, sizeof(int32_t) * M_, cudaMemcpyDeviceToHost, stream_));
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
  ++n_syncs_;
  ExchangePlan plan;
  make_exchange_plan(anc_host_, M_, comm_.rank, comm_.nranks, plan);
  const int64_t ns = (int64_t)plan.send_slot.size(), nr = (int64_t)plan.recv_slot.size();
  const int64_t cb = cur_->cloud_bytes();
  const int64_t need[2] = {ns * cb, nr * cb};
  for (int i = 0; i < 2; ++i)
    if (xbuf_cap_[i] < need[i]) {
      cudaFree(xbuf_[i]);
      xbuf_[i] = nullptr;
      xbuf_cap_[i] = 0;
      SMCB_CUDA_TRY(cudaMalloc(&xbuf_[i], (size_t)need[i]));
      xbuf_cap_[i] = need[i];
    }
  std::memcpy(slots_host_, plan.local_parents.data(), sizeof(int32_t) * Mloc_);
  if (ns) std::memcpy(slots_host_ + M_, plan.send_slot.data(), sizeof(int32_t) * ns);
  if (nr) std::memcpy(slots_host_ + 2 * M_, plan.recv_slot.data(), sizeof(int32_t) * nr);
  SMCB_CUDA_TRY(cudaMemcpyAsync(slots_dev_, slots_host_, sizeof(int32_t) * 3 * M_, cudaMemcpyHostToDevice, stream_));
  mark(SK_EXCHANGE, true);
  cur_->pack_dev(slots_dev_ + M_, ns, xbuf_[0], true);  // reads the pre-gather clouds
  const NcclApi& nc = nccl_api();
  SMCB_NCCL_TRY(nc.GroupStart());
  for (int64_t i = 0; i < ns;) {
    int64_t j = i;
    while (j < ns && plan.send_peer[(size_t)j] == plan.send_peer[(size_t)i]) ++j;
    SMCB_NCCL_TRY(nc.Send(static_cast<char*>(xbuf_[0]) + i * cb, (size_t)((j - i) * cb), kNcclUint8, plan.send_peer[(size_t)i], comm_.comm, stream_));
    i = j;
  }
  for (int64_t i = 0; i < nr;) {
    int64_t j = i;
    while (j < nr && plan.recv_peer[(size_t)j] == plan.recv_peer[(size_t)i]) ++j;
    SMCB_NCCL_TRY(nc.Recv(static_cast<char*>(xbuf_[1]) + i * cb, (size_t)((j - i) * cb), kNcclUint8, plan.recv_peer[(size_t)i], comm_.comm, stream_));
    i = j;
  }
  SMCB_NCCL_TRY(nc.GroupEnd());
  cur_->gather_dev(slots_dev_);
  cur_->pack_dev(slots_dev_ + 2 * M_, nr, xbuf_[1], false);
  mark(SK_EXCHANGE, false);
  n_clouds_moved_ += nr;
}

void ThetaSampler::rejuvenate(int64_t t_len, double xi) {
  const uint32_t ordinal = n_rejuv_++;
  double Sigma[kMaxThetaDim * kMaxThetaDim];
  random_walk_sigma(theta_host_, M_, d_, Sigma);  // pmmh_kernel = smc.kernel(smc.θ)   :107
  if (!prop_) prop_.reset(new BatchFilter(device_, stream_, kind_, Mloc_, N_));
  SMCB_CUDA_TRY(cudaMemsetAsync(acc_any_, 0, (size_t)M_, stream_));  // acc_array = zeros(Int64, M)   :104
  const RngKey hkey = key(ordinal);
  const unsigned grid = (unsigned)((M_ + 127) / 128);
  for (int c = 0; c < chain_; ++c) {
    const double scale = 0.5 * (double)(chain_ - c);  // scales = 0.5*reverse(1:chain)   :108
    CholFactor L;
    std::memset(&L, 0, sizeof(L));
    if (d_ == 1) L.l[0][0] = scale * Sigma[0];
    else {
      double Lf[kMaxThetaDim * kMaxThetaDim];
      if (!cholesky_lower(Sigma, d_, scale, Lf)) throw Error{SMCB_ERR_STATE, "rejuvenate: proposal covariance is not positive definite"};
      for (int j = 0; j < d_; ++j)
        for (int k = 0; k < d_; ++k) L.l[j][k] = Lf[j * d_ + k];
    }
    mark(SK_THETA, true);
    theta_propose_kernel<<<grid, 128, 0, stream_>>>(theta_[tcur_], (int)M_, d_, kind_, L, hkey, (uint32_t)c, prior_, map_, theta_prop_, lp_prop_,
                                                   ok_, derived_prop_);
    SMCB_CUDA_TRY(cudaGetLastError());
  ++n_theta_launches_;
    mark(SK_THETA, false);
    const uint32_t e = next_epoch();
    mark(SK_FILTER, true);
    prop_->run_dev(derived_prop_ + lo_ * kParamStride, ok_ + lo_, y_dev_, t_len, resampler_, key(e), (uint32_t)lo_, logz_prop_ + lo_);  // :117-121
    mark(SK_FILTER, false);
    ++n_sweeps_;
    n_particle_updates_ += Mloc_ * N_ * t_len;
    all_gather(logz_prop_);
    mark(SK_THETA, true);
    theta_accept_kernel<<<grid, 128, 0, stream_>>>((int)M_, d_, xi, hkey, (uint32_t)c, theta_prop_, lp_prop_, ok_, logz_prop_, derived_prop_,
                                                  theta_[tcur_], lp_[tcur_], logz_[tcur_], derived_[tcur_], accept_, acc_any_);
    SMCB_CUDA_TRY(cudaGetLastError());
  ++n_theta_launches_;
    mark(SK_THETA, false);
    mark(SK_EXCHANGE, true);
    cur_->accept_dev(*prop_, accept_ + lo_);  // :130-133
    mark(SK_EXCHANGE, false);
  }
  mark(SK_THETA, true);
  theta_finish_kernel<<<1, kTB, 0, stream_>>>((int)M_, acc_any_, omega_, scal_dev_);
  SMCB_CUDA_TRY(cudaGetLastError());
  ++n_theta_launches_;
  mark(SK_THETA, false);
  SMCB_CUDA_TRY(cudaMemcpyAsync(theta_host_, theta_[tcur_], sizeof(double) * M_ * d_, cudaMemcpyDeviceToHost, stream_));
  read_scalars();
  acc_ratio_ = scal_host_->acc_count / (double)M_;  // :142
  ++n_rejuv_done_;
}

void ThetaSampler::exchange(int64_t t_len) {
  if (!(acc_ratio_ < acc_threshold_)) return;  // :165
  if (N_ > 4096) return;                       // "[cannot exceed max state particles]"   :187
  N_ *= 2;                                     // :167
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
  n_batch_launches_retired_ += cur_->launches() + (prop_ ? prop_->launches() : 0);
  prop_.reset();
  cur_.reset(new BatchFilter(device_, stream_, kind_, Mloc_, N_));
  prop_.reset(new BatchFilter(device_, stream_, kind_, Mloc_, N_));
  const uint32_t e = next_epoch();
  mark(SK_FILTER, true);
  cur_->run_dev(derived_[tcur_] + lo_ * kParamStride, nullptr, y_dev_, t_len, resampler_, key(e), (uint32_t)lo_, logmu_ + lo_);  // :174-180
  mark(SK_FILTER, false);
  ++n_sweeps_;
  n_particle_updates_ += Mloc_ * N_ * t_len;
  all_gather(logmu_);
  mark(SK_THETA, true);
  theta_step_kernel<<<1, kTB, sizeof(double) * M_, stream_>>>(logmu_, omega_, logz_[tcur_], scal_dev_, (int)M_, 2);  // :183-185
  SMCB_CUDA_TRY(cudaGetLastError());
  ++n_theta_launches_;
  mark(SK_THETA, false);
  read_scalars();
  ess_ = scal_host_->ess;
}

int ThetaSampler::density_tempered(double* schedule, int cap) {
  if (T_ < 1) throw Error{SMCB_ERR_STATE, "sampler: set_data before density_tempered"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  open_span();
  const uint32_t e = next_epoch();
  mark(SK_FILTER, true);
  cur_->run_dev(derived_[tcur_] + lo_ * kParamStride, nullptr, y_dev_, T_, resampler_, key(e), (uint32_t)lo_, logz_[tcur_] + lo_);  // :223-229
  mark(SK_FILTER, false);
  ++n_sweeps_;
  n_particle_updates_ += Mloc_ * N_ * T_;
  all_gather(logz_[tcur_]);
  double xi = 0.0;
  int stages = 0;
  while (xi < 1.0) {  // :235
    mark(SK_THETA, true);
    theta_bisect_kernel<<<1, kTB, sizeof(double) * M_, stream_>>>(logz_[tcur_], omega_, scal_dev_, (int)M_, xi, ess_min_);  // :237-266
    SMCB_CUDA_TRY(cudaGetLastError());
  ++n_theta_launches_;
    mark(SK_THETA, false);
    read_scalars();
    xi = scal_host_->xi;
    ess_ = scal_host_->ess;
    double* row = (schedule && stages < cap) ? schedule + 3 * stages : nullptr;
    if (row) {
      row[0] = xi;
      row[1] = ess_;
      row[2] = -1.0;
    }
    ++stages;
    if (scal_host_->resample_flag) {
      resample();             // :272
      rejuvenate(T_, xi);     // :275
      if (row) row[2] = acc_ratio_;
    }
    if (stages > 100000) throw Error{SMCB_ERR_STATE, "density_tempered: the tempering schedule does not advance"};
  }
  started_ = true;
  return stages;
}

void ThetaSampler::get(double* theta, double* omega, double* logZ, double* ess, double* acc_ratio, int64_t* N) {
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  if (theta) SMCB_CUDA_TRY(cudaMemcpyAsync(theta, theta_[tcur_], sizeof(double) * M_ * d_, cudaMemcpyDeviceToHost, stream_));
  if (omega) SMCB_CUDA_TRY(cudaMemcpyAsync(omega, omega_, sizeof(double) * M_, cudaMemcpyDeviceToHost, stream_));
  if (logZ) SMCB_CUDA_TRY(cudaMemcpyAsync(logZ, logz_[tcur_], sizeof(double) * M_, cudaMemcpyDeviceToHost, stream_));
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
  if (ess) *ess = ess_;
  if (acc_ratio) *acc_ratio = acc_ratio_;
  if (N) *N = N_;
}

void ThetaSampler::stats(double ms[8], int64_t counts[8]) {
  for (int i = 0; i < 8; ++i) { ms[i] = 0.0; counts[i] = 0; }
  for (int i = 0; i < SK_COUNT; ++i) ms[i] = ms_[i];
  ms[4] = span_ms_;
  counts[0] = n_sweeps_;
  counts[1] = n_steps_;
  counts[2] = n_rejuv_done_;
  counts[3] = n_clouds_moved_;
  counts[4] = n_particle_updates_;
  counts[5] = n_syncs_;
  counts[6] = n_theta_launches_ + n_batch_launches_retired_ + (cur_ ? cur_->launches() : 0) + (prop_ ? prop_->launches() : 0);
  counts[7] = n_resample_;
}

}  // namespace smcb
