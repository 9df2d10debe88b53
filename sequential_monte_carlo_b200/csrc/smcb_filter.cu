// Grid-wide kernels of one large-N bootstrap particle filter (docs/SPEC.md §5-§7).
//
// One time step = two launches (reference: bootstrap_filter!, /root/reference/src/particles.jl:107-129)
//
//   scan_kernel  normalize() + the CDF of resample():  reads logw, quantises exp(logw - max) to
//                fixed point, single-pass decoupled look-back prefix sum in uint64 (exact, so
//                ancestors do not depend on the tiling), writes the CDF, reduces Σe, Σe² (-> logμ,
//                ess of the previous step) in a fixed order.             particles.jl:5-15,117
//   prop_kernel  resample + gather + transition + observation logpdf fused: each CTA owns a
//                contiguous particle range, finds the CDF window of its first/last threshold with
//                a 32-ary warp search, stages the window in shared memory, every thread
//                binary-searches its two ancestors there, gathers the parents, draws the
//                transition with Philox/Box-Muller, writes x', logw' with 16-byte stores and
//                contributes to the exact max(logw') for the next scan.  particles.jl:117-125
//
// HBM traffic per particle-update (LG1D fp64): logw 8 R + cdf 8 W | cdf 8 R + x 8 R + x' 8 W +
// logw' 8 W = 48 B (the ancestor vector is never materialised unless recording is on).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "smcb_filter.cuh"

namespace smcb {

namespace {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048 log-weights per CTA

constexpr int kPropThreads = 256;
constexpr int kPropPairs = 2;                                  // pairs per thread
constexpr int kPropParticles = kPropThreads * kPropPairs * 2;  // 1024 particles per CTA
constexpr int kStageCap = 4096;                                // CDF entries staged per CTA (32 KB)

constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagInc = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ double block_max(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = (threadIdx.x < nw) ? sh[threadIdx.x] : -INFINITY;
  if (warp == 0) r = warp_max(r);
  return r;  // valid in warp 0
}

__global__ void reset_ctrl_kernel(FilterCtrl* ctrl) {
  ctrl->maxslot[0] = encode_ordered(-INFINITY);
  ctrl->maxslot[1] = encode_ordered(-INFINITY);
  ctrl->total = 0;
  ctrl->scan_ticket = 0;
  ctrl->scan_done = 0;
}

// ------------------------------------------------------------------------------------------------
// bootstrap_filter: x_i ~ initial_dist, logw_i = logpdf(observation(x_i), y)   particles.jl:96-99
template <class Model>
__global__ void __launch_bounds__(256) init_kernel(Derived dv, double y0, int64_t N, int64_t ld, RngKey key,
                                                    uint32_t stream, double* __restrict__ x,
                                                    double* __restrict__ logw, FilterCtrl* ctrl) {
  constexpr int D = Model::D;
  __shared__ double sh[32];
  Model mdl;
  mdl.load(dv.d);
  const int64_t npairs = (N + 1) >> 1;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double vmax = -INFINITY;
  if (p < npairs) {
    double za[D], zb[D], xa[D], xb[D];
#pragma unroll
    for (int k = 0; k < D; ++k) normal_pair_at(key, (uint32_t)p, stream, 0u, PURPOSE_INIT, (uint32_t)k, za[k], zb[k]);
    mdl.init(za, xa);
    mdl.init(zb, xb);
    const double la = mdl.logweight(xa, y0);
    const double lb = mdl.logweight(xb, y0);
    const int64_t i = 2 * p;
    if (i + 1 < N) {
#pragma unroll
      for (int k = 0; k < D; ++k) *reinterpret_cast<double2*>(x + k * ld + i) = make_double2(xa[k], xb[k]);
      *reinterpret_cast<double2*>(logw + i) = make_double2(la, lb);
      if (la > vmax) vmax = la;  // `>` ignores NaN like the CPU loop
      if (lb > vmax) vmax = lb;
    } else {
#pragma unroll
      for (int k = 0; k < D; ++k) x[k * ld + i] = xa[k];
      logw[i] = la;
      if (la > vmax) vmax = la;
    }
  }
  double bm = block_max(vmax, sh);
  if (threadIdx.x == 0) atomicMax(&ctrl->maxslot[0], encode_ordered(bm));
}

// ------------------------------------------------------------------------------------------------
// normalize() + CDF.  WRITE_CDF=false is the stats-only variant used after the last step.
template <bool WRITE_CDF, bool FROM_W>
__global__ void __launch_bounds__(kScanThreads)
    scan_kernel(const double* __restrict__ logw, uint64_t* __restrict__ cdf, int64_t N, int S, FilterCtrl* ctrl,
                unsigned long long* desc_cur, unsigned long long* desc_next, double* psum, double* psum2,
                StepStats* stats_out, int slot, unsigned ntiles) {
  __shared__ unsigned s_tile;
  __shared__ unsigned long long s_wq[kScanThreads / 32];
  __shared__ double s_we[kScanThreads / 32], s_we2[kScanThreads / 32];
  __shared__ unsigned long long s_excl;
  __shared__ bool s_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(&ctrl->scan_ticket, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  const double mx = decode_ordered(ctrl->maxslot[slot]);

  const int64_t base = (int64_t)tile * kScanTile + (int64_t)tid * kScanItems;
  double lw[kScanItems];
  if (base + kScanItems <= N) {
    const double2 a = __ldcs(reinterpret_cast<const double2*>(logw + base));
    const double2 b = __ldcs(reinterpret_cast<const double2*>(logw + base + 2));
    lw[0] = a.x; lw[1] = a.y; lw[2] = b.x; lw[3] = b.y;
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) lw[k] = (base + k < N) ? logw[base + k] : -INFINITY;
  }
  unsigned long long q[kScanItems];
  double se = 0.0, se2 = 0.0;
  unsigned long long tq = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    double e;
    uint64_t qq;
    if (base + k < N) {
      if (FROM_W) {  // standalone resample(w): q = trunc((w / max w) 2^S)   (SPEC §5)
        e = lw[k];
        qq = (mx > 0.0 && e > 0.0) ? (uint64_t)((e / mx) * u64_as_double((uint64_t)(1023 + S) << 52)) : 0ull;
      } else {
        det_exp_quant(lw[k] - mx, S, e, qq);
      }
    } else {
      e = 0.0;
      qq = 0;
    }
    se += e;
    se2 += e * e;
    tq += qq;
    q[k] = tq;  // inclusive within the thread
  }
  // block-level scan of the thread totals
  const unsigned long long winc = warp_scan_u64(tq, lane);
  const double wse = warp_sum(se), wse2 = warp_sum(se2);
  if (lane == 31) s_wq[warp] = winc;
  if (lane == 0) {
    s_we[warp] = wse;
    s_we2[warp] = wse2;
  }
  __syncthreads();
  if (warp == 0) {
    constexpr int NW = kScanThreads / 32;
    unsigned long long v = (lane < NW) ? s_wq[lane] : 0ull;
    const unsigned long long vinc = warp_scan_u64(v, lane);
    const unsigned long long agg = __shfl_sync(kFullMask, vinc, NW - 1);
    if (lane < NW) s_wq[lane] = vinc - v;  // exclusive warp offsets
    double e1 = (lane < NW) ? s_we[lane] : 0.0, e2 = (lane < NW) ? s_we2[lane] : 0.0;
    e1 = warp_sum(e1);
    e2 = warp_sum(e2);
    if (lane == 0) {
      psum[tile] = e1;
      psum2[tile] = e2;
    }
    unsigned long long excl = 0;
    if (WRITE_CDF) {
      if (lane == 0) st_volatile_u64(&desc_cur[tile], (tile == 0 ? kFlagInc : kFlagAgg) | agg);
      if (tile > 0) {
        int64_t look = (int64_t)tile - 1;
        while (true) {
          const int64_t idx = look - lane;
          const unsigned long long d = (idx >= 0) ? ld_volatile_u64(&desc_cur[idx]) : kFlagInc;
          const unsigned invalid = __ballot_sync(kFullMask, (d >> 62) == 0ull);
          const unsigned inc = __ballot_sync(kFullMask, (d >> 62) >= 2ull);
          const int first_inc = inc ? (__ffs(inc) - 1) : 32;
          const unsigned need = (first_inc >= 31) ? 0xFFFFFFFFu : ((2u << first_inc) - 1u);
          if (invalid & need) continue;  // a predecessor we need has not published yet
          const unsigned long long v2 = (lane <= first_inc) ? (d & kValueMask) : 0ull;
          excl += warp_sum_u64(v2);
          if (first_inc < 32) break;
          look -= 32;
        }
        if (lane == 0) st_volatile_u64(&desc_cur[tile], kFlagInc | (excl + agg));
      }
      if (lane == 0) {
        desc_next[tile] = 0ull;  // ready for the next step's scan
        if (tile == ntiles - 1) ctrl->total = excl + agg;
      }
    }
    if (lane == 0) s_excl = excl;
  }
  __syncthreads();
  if (WRITE_CDF) {
    const unsigned long long off = s_excl + s_wq[warp] + (winc - tq);
    if (base + kScanItems <= N) {
      ulonglong2 o0 = make_ulonglong2(off + q[0], off + q[1]);
      ulonglong2 o1 = make_ulonglong2(off + q[2], off + q[3]);
      *reinterpret_cast<ulonglong2*>(cdf + base) = o0;
      *reinterpret_cast<ulonglong2*>(cdf + base + 2) = o1;
    } else {
#pragma unroll
      for (int k = 0; k < kScanItems; ++k)
        if (base + k < N) cdf[base + k] = off + q[k];
    }
  }
  // last CTA out: fixed-order reduction of the per-tile partials -> deterministic Σe, Σe²
  __threadfence();
  if (tid == 0) s_last = (atomicAdd(&ctrl->scan_done, 1u) == ntiles - 1);
  __syncthreads();
  if (s_last) {
    double a = 0.0, b = 0.0;
    for (unsigned j = tid; j < ntiles; j += kScanThreads) {
      a += __ldcg(&psum[j]);
      b += __ldcg(&psum2[j]);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();
    if (lane == 0) {
      s_we[warp] = a;
      s_we2[warp] = b;
    }
    __syncthreads();
    if (warp == 0) {
      constexpr int NW = kScanThreads / 32;
      double e1 = (lane < NW) ? s_we[lane] : 0.0, e2 = (lane < NW) ? s_we2[lane] : 0.0;
      e1 = warp_sum(e1);
      e2 = warp_sum(e2);
      if (lane == 0) {
        stats_out->mx = mx;
        stats_out->sum = e1;
        stats_out->sum2 = e2;
        ctrl->scan_ticket = 0;
        ctrl->scan_done = 0;
        ctrl->maxslot[slot ^ 1] = encode_ordered(-INFINITY);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 32-ary cooperative search: #{ j in [0,n) : C_j <= tau } with one warp (5 rounds at n = 2^24)
__device__ __forceinline__ int64_t warp_lower_count(const uint64_t* __restrict__ C, int64_t n, uint64_t tau,
                                                    int lane) {
  int64_t lo = 0, hi = n;
  while (hi > lo) {
    const int64_t len = hi - lo;
    const int64_t step = (len + 31) >> 5;
    const int64_t p = lo + (int64_t)lane * step + (step - 1);
    const bool le = (p < hi) && (__ldg(&C[p]) <= tau);
    const unsigned m = __ballot_sync(kFullMask, le);
    const int c = __popc(m);  // monotone -> leading ones
    lo += (int64_t)c * step;
    if (c == 32) break;
    const int64_t nh = lo + step - 1;
    hi = nh < hi ? nh : hi;
  }
  return lo;
}

__device__ __forceinline__ uint64_t threshold_of(int resampler, uint64_t i, uint64_t Rw, uint64_t u,
                                                 uint64_t Q) {
  // SPEC §5: F_i then tau_i = mulhi(F_i, Q); u = U(i) (multinomial/stratified) or mulhi(U(0),R) (systematic)
  uint64_t F;
  if (resampler == RESAMPLE_MULTINOMIAL) F = u;
  else if (resampler == RESAMPLE_STRATIFIED) F = i * Rw + mulhi64(u, Rw);
  else F = i * Rw + u;
  return mulhi64(F, Q);
}

// bootstrap_filter!: a = resample(w); x_i ~ transition(x[a_i]); logw_i = logpdf(observation(x_i), y)
template <class Model>
__global__ void __launch_bounds__(kPropThreads)
    prop_kernel(Derived dv, double y, int64_t N, int64_t ld, int resampler, uint64_t Rw, RngKey key,
                uint32_t stream, uint32_t t, const uint64_t* __restrict__ cdf, const double* __restrict__ xprev,
                double* __restrict__ xnew, double* __restrict__ logw, int32_t* __restrict__ anc_out,
                FilterCtrl* ctrl) {
  constexpr int D = Model::D;
  __shared__ uint64_t s_cdf[kStageCap];
  __shared__ int64_t s_bound[2];
  __shared__ double sh[32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Model mdl;
  mdl.load(dv.d);
  const uint64_t Q = ctrl->total;
  const int64_t npairs = (N + 1) >> 1;
  const int64_t pair0 = (int64_t)blockIdx.x * (kPropThreads * kPropPairs);
  const int64_t i_first = 2 * pair0;
  int64_t i_end = i_first + kPropParticles;
  if (i_end > N) i_end = N;
  const bool sorted = (resampler != RESAMPLE_MULTINOMIAL);

  uint64_t sys_off = 0;
  if (resampler == RESAMPLE_SYSTEMATIC) sys_off = mulhi64(uniform64_at(key, 0u, stream, t, PURPOSE_RESAMPLE), Rw);

  // warps 0/1 locate the CDF window of this CTA while the others already draw their normals
  if (sorted && Q != 0 && warp < 2) {
    const int64_t ib = (warp == 0) ? i_first : (i_end - 1);
    uint64_t u = sys_off;
    if (resampler == RESAMPLE_STRATIFIED) u = uniform64_at(key, (uint32_t)ib, stream, t, PURPOSE_RESAMPLE);
    const uint64_t tau = threshold_of(resampler, (uint64_t)ib, Rw, u, Q);
    const int64_t a = warp_lower_count(cdf, N - 1, tau, lane);
    if (lane == 0) s_bound[warp] = a;
  }

  double za[kPropPairs][D], zb[kPropPairs][D];
  uint64_t ua[kPropPairs], ub[kPropPairs];
#pragma unroll
  for (int r = 0; r < kPropPairs; ++r) {
    const int64_t p = pair0 + (int64_t)r * kPropThreads + tid;
    ua[r] = ub[r] = sys_off;
    if (p < npairs) {
#pragma unroll
      for (int k = 0; k < D; ++k)
        normal_pair_at(key, (uint32_t)p, stream, t, PURPOSE_TRANSITION, (uint32_t)k, za[r][k], zb[r][k]);
      if (resampler != RESAMPLE_SYSTEMATIC) {
        const Philox4 b = philox4x32_10((uint32_t)p, stream, t, purpose_word(PURPOSE_RESAMPLE, 0, key.epoch), key.k0, key.k1);
        ua[r] = uniform64_of(b, 0);
        ub[r] = uniform64_of(b, 1);
      }
    }
  }
  __syncthreads();

  int64_t a_lo = 0, a_hi = N - 1;
  bool staged = false;
  if (sorted && Q != 0) {
    a_lo = s_bound[0];
    a_hi = s_bound[1];
    const int64_t span = a_hi - a_lo;
    staged = span <= kStageCap;
    if (staged) {
      for (int64_t j = tid; j < span; j += kPropThreads) s_cdf[j] = __ldcs(&cdf[a_lo + j]);
    }
  }
  __syncthreads();

  double vmax = -INFINITY;
#pragma unroll
  for (int r = 0; r < kPropPairs; ++r) {
    const int64_t p = pair0 + (int64_t)r * kPropThreads + tid;
    if (p >= npairs) continue;
    const int64_t i = 2 * p;
    const bool two = (i + 1 < N);
    int64_t a0 = i, a1 = i + 1;
    if (Q != 0) {
      const uint64_t t0 = threshold_of(resampler, (uint64_t)i, Rw, ua[r], Q);
      const uint64_t t1 = threshold_of(resampler, (uint64_t)(i + 1), Rw, ub[r], Q);
      if (staged) {
        int lo = 0, hi = (int)(a_hi - a_lo);
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (s_cdf[mid] <= t0) lo = mid + 1;
          else hi = mid;
        }
        a0 = a_lo + lo;
        hi = (int)(a_hi - a_lo);  // a1 >= a0 for sorted thresholds: keep lo
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (s_cdf[mid] <= t1) lo = mid + 1;
          else hi = mid;
        }
        a1 = a_lo + lo;
      } else {
        a0 = lower_count(cdf, a_lo, a_hi, t0);
        a1 = two ? lower_count(cdf, sorted ? a0 : a_lo, a_hi, t1) : a0;
      }
    }
    double xpa[D], xpb[D], xa[D], xb[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      xpa[k] = __ldg(&xprev[k * ld + a0]);
      xpb[k] = two ? __ldg(&xprev[k * ld + a1]) : xpa[k];
    }
    mdl.transition(za[r], xpa, xa);
    mdl.transition(zb[r], xpb, xb);
    const double la = mdl.logweight(xa, y);
    const double lb = mdl.logweight(xb, y);
    if (two) {
#pragma unroll
      for (int k = 0; k < D; ++k) *reinterpret_cast<double2*>(xnew + k * ld + i) = make_double2(xa[k], xb[k]);
      *reinterpret_cast<double2*>(logw + i) = make_double2(la, lb);
      if (anc_out) *reinterpret_cast<int2*>(anc_out + i) = make_int2((int)a0, (int)a1);
      if (la > vmax) vmax = la;
      if (lb > vmax) vmax = lb;
    } else {
#pragma unroll
      for (int k = 0; k < D; ++k) xnew[k * ld + i] = xa[k];
      logw[i] = la;
      if (anc_out) anc_out[i] = (int)a0;
      if (la > vmax) vmax = la;
    }
  }
  const double bm = block_max(vmax, sh);
  if (tid == 0) atomicMax(&ctrl->maxslot[t & 1u], encode_ordered(bm));
}

// w_i = exp(logw_i - max) / Σe   (normalize, particles.jl:11) — only when the caller fetches w
__global__ void weights_kernel(const double* __restrict__ logw, double* __restrict__ w, int64_t N, StepStats st) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) w[i] = det_exp(logw[i] - st.mx) / st.sum;
}

// utilities (normalize / resample on caller-supplied vectors; θ-level sizes, not hot)
__global__ void max_kernel(const double* __restrict__ v, int64_t n, FilterCtrl* ctrl) {
  __shared__ double sh[32];
  double m = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double a = v[i];
    if (a > m) m = a;
  }
  m = block_max(m, sh);
  if (threadIdx.x == 0) atomicMax(&ctrl->maxslot[0], encode_ordered(m));
}

__global__ void ancestor_kernel(const uint64_t* __restrict__ cdf, int64_t N, int resampler, uint64_t Rw, RngKey key,
                                uint32_t stream, uint32_t t, uint32_t purpose, const FilterCtrl* ctrl,
                                int64_t* __restrict__ anc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint64_t Q = ctrl->total;
  if (Q == 0) {
    anc[i] = i;
    return;
  }
  uint64_t u;
  if (resampler == RESAMPLE_SYSTEMATIC) u = mulhi64(uniform64_at(key, 0u, stream, t, purpose), Rw);
  else u = uniform64_at(key, (uint32_t)i, stream, t, purpose);
  const uint64_t tau = threshold_of(resampler, (uint64_t)i, Rw, u, Q);
  anc[i] = lower_count(cdf, (int64_t)0, N - 1, tau);
}

template <class F>
void dispatch_model(int kind, F&& f) {
  switch (kind) {
    case KIND_LG1D: f(ModelLG1D{}); break;
    case KIND_SV: f(ModelSV{}); break;
    case KIND_UCSV: f(ModelUCSV{}); break;
    default: throw Error{SMCB_ERR_BAD_ARG, "unknown model kind"};
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
SingleFilter::~SingleFilter() {
  release();
  for (auto e : ev_pool_) cudaEventDestroy(e);
  for (auto e : ev_call_)
    if (e) cudaEventDestroy(e);
}

void SingleFilter::release() {
  cudaFree(x_[0]); cudaFree(x_[1]); cudaFree(logw_); cudaFree(w_tmp_); cudaFree(cdf_); cudaFree(anc_);
  cudaFree(ctrl_); cudaFree(desc_); cudaFree(psum_); cudaFree(psum2_); cudaFree(stats_dev_);
  x_[0] = x_[1] = logw_ = w_tmp_ = psum_ = psum2_ = nullptr;
  cdf_ = nullptr; anc_ = nullptr; ctrl_ = nullptr; desc_ = nullptr; stats_dev_ = nullptr;
  cap_N_ = cap_d_ = cap_stats_ = cap_anc_rows_ = ntiles_cap_ = 0;
}

void SingleFilter::ensure_capacity(int kind, int64_t N, int64_t anc_rows) {
  const int d = state_dim(kind);
  const int64_t ld = (N + 31) & ~int64_t(31);
  if (ld > cap_N_ || d > cap_d_) {
    cudaFree(x_[0]); cudaFree(x_[1]); cudaFree(logw_); cudaFree(w_tmp_); cudaFree(cdf_);
    cudaFree(desc_); cudaFree(psum_); cudaFree(psum2_);
    x_[0] = x_[1] = logw_ = w_tmp_ = psum_ = psum2_ = nullptr; cdf_ = nullptr; desc_ = nullptr;
    cap_N_ = 0;
    const int64_t cd = std::max<int64_t>(d, cap_d_);
    SMCB_CUDA_TRY(cudaMalloc(&x_[0], sizeof(double) * ld * cd));
    SMCB_CUDA_TRY(cudaMalloc(&x_[1], sizeof(double) * ld * cd));
    SMCB_CUDA_TRY(cudaMalloc(&logw_, sizeof(double) * ld));
    SMCB_CUDA_TRY(cudaMalloc(&cdf_, sizeof(uint64_t) * ld));
    ntiles_cap_ = (ld + kScanTile - 1) / kScanTile;
    SMCB_CUDA_TRY(cudaMalloc(&desc_, sizeof(unsigned long long) * 2 * ntiles_cap_));
    SMCB_CUDA_TRY(cudaMalloc(&psum_, sizeof(double) * ntiles_cap_));
    SMCB_CUDA_TRY(cudaMalloc(&psum2_, sizeof(double) * ntiles_cap_));
    cap_N_ = ld;
    cap_d_ = cd;
    cudaFree(anc_); anc_ = nullptr; cap_anc_rows_ = 0;
  }
  if (!ctrl_) SMCB_CUDA_TRY(cudaMalloc(&ctrl_, sizeof(FilterCtrl)));
  if (anc_rows > cap_anc_rows_ || !anc_) {
    cudaFree(anc_); anc_ = nullptr;
    const int64_t rows = std::max<int64_t>(anc_rows, 1);
    SMCB_CUDA_TRY(cudaMalloc(&anc_, sizeof(int32_t) * rows * cap_N_));
    cap_anc_rows_ = rows;
  }
}

void SingleFilter::mark(int klass, bool start) {
  launches_[klass] += start ? 1 : 0;
  if (!profiling_) return;
  if (ev_used_ == ev_pool_.size()) {
    cudaEvent_t e;
    SMCB_CUDA_TRY(cudaEventCreate(&e));
    ev_pool_.push_back(e);
  }
  SMCB_CUDA_TRY(cudaEventRecord(ev_pool_[ev_used_], stream_));
  if (start) marks_.push_back(Mark{klass, ev_used_, 0});
  else marks_.back().e1 = ev_used_;
  ++ev_used_;
}

void SingleFilter::begin_call() {
  for (int i = 0; i < 2; ++i)
    if (!ev_call_[i]) SMCB_CUDA_TRY(cudaEventCreate(&ev_call_[i]));
  for (int k = 0; k < TK_COUNT; ++k) { ms_[k] = 0; launches_[k] = 0; }
  marks_.clear();
  ev_used_ = 0;
  SMCB_CUDA_TRY(cudaEventRecord(ev_call_[0], stream_));
}

void SingleFilter::end_call() {
  SMCB_CUDA_TRY(cudaEventRecord(ev_call_[1], stream_));
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
  float ms = 0;
  SMCB_CUDA_TRY(cudaEventElapsedTime(&ms, ev_call_[0], ev_call_[1]));
  ms_[TK_TOTAL] = ms;
  for (const auto& m : marks_) {
    SMCB_CUDA_TRY(cudaEventElapsedTime(&ms, ev_pool_[m.e0], ev_pool_[m.e1]));
    ms_[m.klass] += ms;
  }
  launches_[TK_TOTAL] = launches_[TK_SCAN] + launches_[TK_PROP] + launches_[TK_INIT] + launches_[TK_STATS];
}

void SingleFilter::timing(double ms[TK_COUNT], int64_t launches[TK_COUNT]) const {
  for (int k = 0; k < TK_COUNT; ++k) { ms[k] = ms_[k]; launches[k] = launches_[k]; }
}

void SingleFilter::launch_init(double y0) {
  SMCB_CUDA_TRY(cudaMemsetAsync(desc_, 0, sizeof(unsigned long long) * 2 * ntiles_cap_, stream_));
  reset_ctrl_kernel<<<1, 1, 0, stream_>>>(ctrl_);
  launches_[TK_INIT] += 1;
  const int64_t npairs = (N_ + 1) / 2;
  const unsigned grid = (unsigned)((npairs + 255) / 256);
  mark(TK_INIT, true);
  dispatch_model(kind_, [&](auto m) {
    using M = decltype(m);
    init_kernel<M><<<grid, 256, 0, stream_>>>(dv_, y0, N_, ld_, key_, stream_id_, x_[cur_], logw_, ctrl_);
  });
  mark(TK_INIT, false);
  SMCB_CUDA_TRY(cudaGetLastError());
}

void SingleFilter::launch_scan(int64_t stat_index, bool write_cdf) {
  const unsigned ntiles = (unsigned)((N_ + kScanTile - 1) / kScanTile);
  const int slot = (int)(t_ & 1u);  // max of the weights produced at time t_
  unsigned long long* dc = desc_ + (size_t)(t_ & 1u) * ntiles_cap_;
  unsigned long long* dn = desc_ + (size_t)((t_ + 1) & 1u) * ntiles_cap_;
  const int klass = write_cdf ? TK_SCAN : TK_STATS;
  mark(klass, true);
  StepStats* so = stats_dev_ + stat_index;
  if (from_w_)
    scan_kernel<true, true><<<ntiles, kScanThreads, 0, stream_>>>(logw_, cdf_, N_, S_, ctrl_, dc, dn, psum_, psum2_, so, slot, ntiles);
  else if (write_cdf)
    scan_kernel<true, false><<<ntiles, kScanThreads, 0, stream_>>>(logw_, cdf_, N_, S_, ctrl_, dc, dn, psum_, psum2_, so, slot, ntiles);
  else
    scan_kernel<false, false><<<ntiles, kScanThreads, 0, stream_>>>(logw_, cdf_, N_, S_, ctrl_, dc, dn, psum_, psum2_, so, slot, ntiles);
  mark(klass, false);
  SMCB_CUDA_TRY(cudaGetLastError());
}

void SingleFilter::launch_prop(double y, int resampler) {
  const int64_t npairs = (N_ + 1) / 2;
  const unsigned grid = (unsigned)((npairs + kPropThreads * kPropPairs - 1) / (kPropThreads * kPropPairs));
  int32_t* anc = nullptr;
  if (record_anc_) {
    const int64_t row = std::min<int64_t>(anc_rows_, cap_anc_rows_ - 1);
    anc = anc_ + row * cap_N_;
    anc_rows_ = row + 1;
  }
  const uint32_t t = t_ + 1;
  mark(TK_PROP, true);
  dispatch_model(kind_, [&](auto m) {
    using M = decltype(m);
    prop_kernel<M><<<grid, kPropThreads, 0, stream_>>>(dv_, y, N_, ld_, resampler, R_, key_, stream_id_, t, cdf_,
                                                       x_[cur_], x_[cur_ ^ 1], logw_, anc, ctrl_);
  });
  mark(TK_PROP, false);
  SMCB_CUDA_TRY(cudaGetLastError());
  cur_ ^= 1;
  t_ = t;
}

static void check_args(int kind, int64_t N, int resampler) {
  if (kind < 0 || kind >= KIND_COUNT) throw Error{SMCB_ERR_BAD_ARG, "unknown model kind"};
  if (N < 1 || N > (int64_t(1) << 31) - 64) throw Error{SMCB_ERR_BAD_ARG, "N must be in [1, 2^31-64]"};
  if (resampler < 0 || resampler > 2) throw Error{SMCB_ERR_BAD_ARG, "unknown resampler"};
}

void SingleFilter::init(int kind, const double* params, int64_t N, double y0, const RngKey& key,
                        uint32_t stream_id, StepStats* st) {
  check_args(kind, N, 0);
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  N_ = 0;
  ensure_capacity(kind, N, 1);
  if (cap_stats_ < 2) {
    cudaFree(stats_dev_); stats_dev_ = nullptr;
    SMCB_CUDA_TRY(cudaMalloc(&stats_dev_, sizeof(StepStats) * 2));
    cap_stats_ = 2;
  }
  kind_ = kind; d_ = state_dim(kind); N_ = N; ld_ = cap_N_;
  S_ = quant_shift((uint64_t)N); R_ = strata_width((uint64_t)N);
  key_ = key; stream_id_ = stream_id; t_ = 0; cur_ = 0; anc_rows_ = 0;
  derive_params(kind, params, dv_.d);
  begin_call();
  launch_init(y0);
  launch_scan(0, false);
  SMCB_CUDA_TRY(cudaMemcpyAsync(&last_, stats_dev_, sizeof(StepStats), cudaMemcpyDeviceToHost, stream_));
  end_call();
  if (st) *st = last_;
}

void SingleFilter::step(const double* params, double y, int resampler, StepStats* st) {
  if (!live()) throw Error{SMCB_ERR_STATE, "bootstrap_step before bootstrap_init / log_likelihood"};
  check_args(kind_, N_, resampler);
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  if (params) derive_params(kind_, params, dv_.d);
  if (!record_anc_) anc_rows_ = 0;
  begin_call();
  launch_scan(0, true);
  launch_prop(y, resampler);
  launch_scan(0, false);
  SMCB_CUDA_TRY(cudaMemcpyAsync(&last_, stats_dev_, sizeof(StepStats), cudaMemcpyDeviceToHost, stream_));
  end_call();
  if (st) *st = last_;
}

void SingleFilter::run(int kind, const double* params, int64_t N, const double* y, int64_t T, int resampler,
                       const RngKey& key, uint32_t stream_id, StepStats* stats_out) {
  check_args(kind, N, resampler);
  if (T < 1) throw Error{SMCB_ERR_BAD_ARG, "T must be >= 1"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  N_ = 0;
  ensure_capacity(kind, N, record_anc_ ? std::max<int64_t>(T - 1, 1) : 1);
  if (cap_stats_ < T) {
    cudaFree(stats_dev_); stats_dev_ = nullptr;
    SMCB_CUDA_TRY(cudaMalloc(&stats_dev_, sizeof(StepStats) * T));
    cap_stats_ = T;
  }
  kind_ = kind; d_ = state_dim(kind); N_ = N; ld_ = cap_N_;
  S_ = quant_shift((uint64_t)N); R_ = strata_width((uint64_t)N);
  key_ = key; stream_id_ = stream_id; t_ = 0; cur_ = 0; anc_rows_ = 0;
  derive_params(kind, params, dv_.d);
  begin_call();
  launch_init(y[0]);
  for (int64_t t = 1; t < T; ++t) {
    launch_scan(t - 1, true);  // stats of time t-1 + CDF
    launch_prop(y[t], resampler);
  }
  launch_scan(T - 1, false);
  std::vector<StepStats> tmp;
  SMCB_CUDA_TRY(cudaMemcpyAsync(stats_out, stats_dev_, sizeof(StepStats) * T, cudaMemcpyDeviceToHost, stream_));
  end_call();
  last_ = stats_out[T - 1];
}

void SingleFilter::fetch(double* x_host, double* w_host, double* logw_host) {
  if (!live()) throw Error{SMCB_ERR_STATE, "no filter state to fetch"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  if (x_host)
    SMCB_CUDA_TRY(cudaMemcpy2DAsync(x_host, sizeof(double) * N_, x_[cur_], sizeof(double) * ld_, sizeof(double) * N_,
                                    d_, cudaMemcpyDeviceToHost, stream_));
  if (logw_host)
    SMCB_CUDA_TRY(cudaMemcpyAsync(logw_host, logw_, sizeof(double) * N_, cudaMemcpyDeviceToHost, stream_));
  if (w_host) {
    if (!w_tmp_) SMCB_CUDA_TRY(cudaMalloc(&w_tmp_, sizeof(double) * cap_N_));
    weights_kernel<<<(unsigned)((N_ + 255) / 256), 256, 0, stream_>>>(logw_, w_tmp_, N_, last_);
    SMCB_CUDA_TRY(cudaGetLastError());
    SMCB_CUDA_TRY(cudaMemcpyAsync(w_host, w_tmp_, sizeof(double) * N_, cudaMemcpyDeviceToHost, stream_));
  }
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
}

int64_t SingleFilter::fetch_ancestors(int64_t* anc_host, int64_t rows_cap) {
  if (!live()) throw Error{SMCB_ERR_STATE, "no filter state to fetch"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  const int64_t rows = std::min<int64_t>(anc_rows_, rows_cap);
  std::vector<int32_t> tmp((size_t)N_);
  for (int64_t r = 0; r < rows; ++r) {
    SMCB_CUDA_TRY(cudaMemcpyAsync(tmp.data(), anc_ + r * cap_N_, sizeof(int32_t) * N_, cudaMemcpyDeviceToHost, stream_));
    SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
    for (int64_t i = 0; i < N_; ++i) anc_host[r * N_ + i] = tmp[(size_t)i];
  }
  return rows;
}

// ---- utilities on caller vectors: normalize(logw) (particles.jl:5-15), resample(w) (:17-19) -----
void SingleFilter::load_vector(const double* host, int64_t n, bool is_log) {
  if (n < 1 || n > (int64_t(1) << 31) - 64) throw Error{SMCB_ERR_BAD_ARG, "n must be in [1, 2^31-64]"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  N_ = 0;
  ensure_capacity(KIND_LG1D, n, 1);
  if (cap_stats_ < 2) {
    cudaFree(stats_dev_); stats_dev_ = nullptr;
    SMCB_CUDA_TRY(cudaMalloc(&stats_dev_, sizeof(StepStats) * 2));
    cap_stats_ = 2;
  }
  kind_ = KIND_LG1D; d_ = 0; N_ = n; ld_ = cap_N_;
  S_ = quant_shift((uint64_t)n); R_ = strata_width((uint64_t)n);
  t_ = 0; cur_ = 0; anc_rows_ = 0; from_w_ = !is_log;
  SMCB_CUDA_TRY(cudaMemsetAsync(desc_, 0, sizeof(unsigned long long) * 2 * ntiles_cap_, stream_));
  reset_ctrl_kernel<<<1, 1, 0, stream_>>>(ctrl_);
  SMCB_CUDA_TRY(cudaMemcpyAsync(logw_, host, sizeof(double) * n, cudaMemcpyHostToDevice, stream_));
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 1184);
  max_kernel<<<grid, 256, 0, stream_>>>(logw_, n, ctrl_);
  SMCB_CUDA_TRY(cudaGetLastError());
}

void SingleFilter::normalize_vector(const double* logw_host, int64_t n, StepStats* st, double* w_host) {
  load_vector(logw_host, n, true);
  begin_call();
  launch_scan(0, false);
  SMCB_CUDA_TRY(cudaMemcpyAsync(&last_, stats_dev_, sizeof(StepStats), cudaMemcpyDeviceToHost, stream_));
  end_call();
  *st = last_;
  if (w_host) fetch(nullptr, w_host, nullptr);
  N_ = 0;
}

void SingleFilter::resample_vector(const double* w_host, int64_t n, int resampler, const RngKey& key, uint32_t stream_id,
                                   uint32_t t, uint32_t purpose, int64_t* anc_host) {
  if (resampler < 0 || resampler > 2) throw Error{SMCB_ERR_BAD_ARG, "unknown resampler"};
  load_vector(w_host, n, false);
  begin_call();
  launch_scan(0, true);
  from_w_ = false;
  int64_t* anc_dev = nullptr;
  SMCB_CUDA_TRY(cudaMalloc(&anc_dev, sizeof(int64_t) * n));
  ancestor_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream_>>>(cdf_, n, resampler, R_, key, stream_id, t, purpose,
                                                                    ctrl_, anc_dev);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(anc_host, anc_dev, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, stream_);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
  cudaFree(anc_dev);
  N_ = 0;
  SMCB_CUDA_TRY(e);
  end_call();
}

}  // namespace smcb
