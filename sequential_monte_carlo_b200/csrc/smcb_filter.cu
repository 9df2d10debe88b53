// Grid-wide kernels of one large-N bootstrap particle filter (docs/SPEC.md §5-§7).
//
// One time step of bootstrap_filter! (/root/reference/src/particles.jl:107-129):
//
//   sorted resamplers (stratified / systematic) — four launches chained by programmatic dependent
//   launch, see the block comment further down:   sum_kernel -> bounds_kernel -> anc_hist_kernel -> move_kernel
//
//   multinomial (the reference's i.i.d. law, unsorted thresholds) — two launches:
//   scan_kernel  normalize() + the CDF of resample():  reads logw, quantises exp(logw - max) to
//                fixed point, single-pass decoupled look-back prefix sum in uint64 (exact, so
//                ancestors do not depend on the tiling), writes the CDF, reduces Σe, Σe² (-> logμ,
//                ess of the previous step) in a fixed order.             particles.jl:5-15,117
//   prop_kernel  resample + gather + transition + observation logpdf fused: every thread
//                binary-searches the global CDF for its two ancestors, gathers the parents, draws
//                the transition with Philox/Box-Muller, writes x', logw' with 16-byte stores and
//                contributes to the exact max(logw') for the next scan.  particles.jl:117-125
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "smcb_filter.cuh"
#include "smcb_detmathf.cuh"

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
  ctrl->sys_off = 0;
  ctrl->rq_lo = ctrl->rq_hi = 0;
  ctrl->inv_rq = 0.0;
  ctrl->scan_ticket = 0;
  ctrl->scan_done = 0;
}

// State storage type XT (docs/SPEC.md §9): double, or float = every state component rounded to
// binary32 right after it is drawn; all arithmetic stays binary64.
template <class XT> struct XVec2;
template <> struct XVec2<double> { using type = double2; static __device__ __forceinline__ double2 make(double a, double b) { return make_double2(a, b); } };
template <> struct XVec2<float> { using type = float2; static __device__ __forceinline__ float2 make(double a, double b) { return make_float2((float)a, (float)b); } };
template <class XT, int D>
__device__ __forceinline__ void round_state(double (&x)[D]) {
  if (sizeof(XT) == 4) {
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = (double)(float)x[k];
  }
}

// ------------------------------------------------------------------------------------------------
// bootstrap_filter: x_i ~ initial_dist, logw_i = logpdf(observation(x_i), y)   particles.jl:96-99
template <class Model, class XT>
__global__ void __launch_bounds__(256) init_kernel(typename Model::DV dv, double y0, int64_t N, int64_t ld, RngKey key,
                                                    uint32_t stream, XT* __restrict__ x,
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
    round_state<XT>(xa);
    round_state<XT>(xb);
    const double la = mdl.logweight(xa, y0);
    const double lb = mdl.logweight(xb, y0);
    const int64_t i = 2 * p;
    if (i + 1 < N) {
#pragma unroll
      for (int k = 0; k < D; ++k) *reinterpret_cast<typename XVec2<XT>::type*>(x + k * ld + i) = XVec2<XT>::make(xa[k], xb[k]);
      *reinterpret_cast<double2*>(logw + i) = make_double2(la, lb);
      if (la > vmax) vmax = la;  // `>` ignores NaN like the CPU loop
      if (lb > vmax) vmax = lb;
    } else {
#pragma unroll
      for (int k = 0; k < D; ++k) x[k * ld + i] = (XT)xa[k];
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
    prop_kernel(typename Model::DV dv, double y, int64_t N, int64_t ld, int resampler, uint64_t Rw, RngKey key,
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
        const Philox4 b = philox4x32_10((uint32_t)p, stream, t, purpose_word(PURPOSE_RESAMPLE, 0, key.epoch), key);
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

// ================================================================================================
// Sorted resamplers (stratified / systematic): the four-launch step  sum -> bounds -> anc -> move.
//
// sum_kernel quantises the weights and writes the TILE-LOCAL inclusive CDF cl (one u64 per particle;
// one warp walks one tile, so the running prefix never leaves its registers) plus one total per tile;
// every CTA publishes the prefixes of its 8 tiles and one CTA total, the last CTA scans the <= 1024
// CTA totals (exact integers, so every ancestor equals the one a sequential CDF would give).
// bounds_kernel finds, for every ancestor CTA, the ancestor of its first particle (32-ary warp
// searches of the tile index and of one tile).  anc_hist_kernel reads exactly the CDF entries between
// its own and the next CTA's bound and turns each into "the first particle whose threshold reaches
// it" by arithmetic (the thresholds are an arithmetic progression, systematic, or one per stratum,
// stratified) — a histogram and a prefix sum give the ancestors, no search and no walk (see the block
// comment at the kernel).  move_kernel gathers the parents through the sorted int32 ancestor
// vector, draws the transition (Philox + Box-Muller), weights, and feeds the exact max.
// HBM traffic per particle-update (LG1D fp64): logw 8 R + cl 8 W | cl 8 R + anc 4 W | anc 4 R + x 8 R
// + x' 8 W + logw' 8 W = 56 B, the algorithmic figure of SURVEY.md §8(d).
constexpr int kChunk = 128;                            // particles per warp trip of sum_kernel (32 lanes x 4)
constexpr int kSumThreads = 256;                       // 8 warps, one tile per warp, no block barrier in the main loop
constexpr int kSumWarps = kSumThreads / 32;
constexpr int kSumCtasPerSm = 4;
constexpr int kMaxTiles = 8192;                        // tiles of one step (one warp each)
constexpr int kMaxSumCtas = kMaxTiles / kSumWarps;     // CTA totals scanned by the last CTA of sum_kernel
constexpr int kSysThreads = 128;                       // anc_hist_kernel: 128 threads x 16 particles (256 x 16: 45 us, 128 x 8: 45 us, 128 x 16: 39 us)
constexpr int kSysPer = 16;                            // particles per thread: kSysPer / 4 runs of 4 consecutive ones
constexpr int kSysParticles = kSysThreads * kSysPer;   // 2048 particles per ancestor CTA


// AT: arithmetic of the weights — double, or float = the binary32-arithmetic tier of SPEC §9b (the elements are then binary32 too)
template <class XT, class AT = double>  // XT: the element type behind `logw` — double (log-weights, or binary64 states when from_x) or float (binary32 states / log-weights)
__global__ void __launch_bounds__(kSumThreads, kSumCtasPerSm)
    sum_kernel(const XT* __restrict__ logw, unsigned long long* __restrict__ cl, int64_t N, int S, FilterCtrl* ctrl,
               StepIndex ix, double* psum, double* psum2, StepStats* stats_out, int slot, int resampler, uint64_t Rw,
               RngKey key, uint32_t stream, uint32_t t, int from_x, Derived dv, double ycur) {
  // from_x: `logw` points at the LG1D states x and the log-weight logpdf(observation(x), ycur) is
  // recomputed here (2 fma, bit-identical to what move_kernel maximised over) instead of being
  // written by move_kernel and read back: 8 B per particle-update less through HBM.
  // (the constants stay in the kernel-argument bank: ModelLG1D::logweight with B = d[1], ir = d[5], c = d[6])
  auto lg_logweight = [&](double x) {
    const double v = (ycur - dv.d[1] * x) * dv.d[5];
    return fma(-0.5 * v, v, dv.d[6]);
  };
  __shared__ double s_we[kSumWarps], s_we2[kSumWarps];
  __shared__ unsigned long long s_scan[kSumWarps];
  __shared__ unsigned long long s_cta_excl[kMaxSumCtas];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x * kSumWarps + warp;
  pdl_launch_dependents();
  pdl_wait();  // logw and its max come from the previous step's move kernel
  const double mx = decode_ordered(ctrl->maxslot[slot]);
  if (tile < ix.ntiles) {
    // one warp walks its tile chunk by chunk: the running prefix stays in registers
    unsigned long long running = 0;
    double se = 0.0, se2 = 0.0;
    const int64_t tile0 = (int64_t)tile * ix.tile_items;
    int nch = ix.chunks_per_tile;
    {
      const int64_t left = N - tile0;
      const int64_t need = (left + kChunk - 1) / kChunk;
      if (need < nch) nch = (int)need;
    }
    // chunks that lie entirely inside [0, N) take a body without any bounds logic; at most one partial chunk follows
    int nfull = (int)((N - tile0) / kChunk);
    if (nfull > nch) nfull = nch;
    auto process = [&](auto full_tag, int c, double (&lw)[4]) {
      constexpr bool FULL = decltype(full_tag)::value;
      const int64_t base = tile0 + (int64_t)c * kChunk + lane * 4;
      unsigned long long q[4];
      if constexpr (std::is_same<AT, float>::value) {
        // binary32 arithmetic (SPEC §9b): the elements are binary32 values (carried here as doubles, exactly); the log-weight
        // of an LG1D state, the shift by the max and the exponential are binary32 operations, the two sums binary64
        const float fB = (float)dv.d[1], fir = (float)dv.d[5], fc = (float)dv.d[6], fy = (float)ycur, fmx = (float)mx;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float lf = (float)lw[k];
          if (from_x) {
            const float v = (fy - fB * lf) * fir;
            lf = (FULL || base + k < N) ? fmaf(-0.5f * v, v, fc) : -INFINITY;
          }
          float ef;
          uint64_t qq;
          det_exp_quantf(lf - fmx, S, ef, qq);
          const double e = (double)ef;
          se += e;
          se2 += e * e;
          q[k] = qq;
        }
      } else {
        if (from_x) {  // (partial chunk: out-of-range items were loaded as -inf and their "weight" must stay -inf)
#pragma unroll
          for (int k = 0; k < 4; ++k) lw[k] = (FULL || base + k < N) ? lg_logweight(lw[k]) : -INFINITY;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // out-of-range items carry logw = -inf: e = 0, q = 0
          double e;
          uint64_t qq;
          det_exp_quant_stream(lw[k] - mx, S, e, qq);
          se += e;
          se2 += e * e;
          q[k] = qq;
        }
      }
      q[1] += q[0];
      q[2] += q[1];
      q[3] += q[2];
      const unsigned long long winc = warp_scan_u64(q[3], lane);
      const unsigned long long off = running + (winc - q[3]);
      if (FULL || base + 4 <= N) {
        *reinterpret_cast<ulonglong2*>(cl + base) = make_ulonglong2(off + q[0], off + q[1]);
        *reinterpret_cast<ulonglong2*>(cl + base + 2) = make_ulonglong2(off + q[2], off + q[3]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (base + k < N) cl[base + k] = off + q[k];
      }
      running += __shfl_sync(kFullMask, winc, 31);
    };
    // software pipeline: the next chunk's log-weights are in flight while this one is quantised
    if (sizeof(XT) == 8) {
      const double2* src = reinterpret_cast<const double2*>(logw + tile0 + lane * 4);  // chunk c: src[c * 64], src[c * 64 + 1]
      double2 na = make_double2(0.0, 0.0), nb = na;
      if (nfull > 0) {
        na = __ldcs(src);
        nb = __ldcs(src + 1);
      }
#pragma unroll 1
      for (int c = 0; c < nfull; ++c) {
        double lw[4] = {na.x, na.y, nb.x, nb.y};
        if (c + 1 < nfull) {
          na = __ldcs(src + (c + 1) * (kChunk / 2));
          nb = __ldcs(src + (c + 1) * (kChunk / 2) + 1);
        }
        process(std::true_type{}, c, lw);
      }
    } else {  // binary32 states: one 16-byte load per lane and chunk
      const float4* src = reinterpret_cast<const float4*>(logw + tile0 + lane * 4);  // chunk c: src[c * 32]
      float4 nf = make_float4(0.f, 0.f, 0.f, 0.f);
      if (nfull > 0) nf = __ldcs(src);
#pragma unroll 1
      for (int c = 0; c < nfull; ++c) {
        double lw[4] = {(double)nf.x, (double)nf.y, (double)nf.z, (double)nf.w};
        if (c + 1 < nfull) nf = __ldcs(src + (c + 1) * (kChunk / 4));
        process(std::true_type{}, c, lw);
      }
    }
    if (nfull < nch) {  // the partial chunk at the end of the array
      const int64_t base = tile0 + (int64_t)nfull * kChunk + lane * 4;
      double lw[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) lw[k] = (base + k < N) ? (double)logw[base + k] : -INFINITY;
      process(std::false_type{}, nfull, lw);
    }
    se = warp_sum(se);
    se2 = warp_sum(se2);
    if (lane == 0) {
      s_we[warp] = se;
      s_we2[warp] = se2;
      s_scan[warp] = running;
    }
  } else if (lane == 0) {
    s_we[warp] = 0.0;
    s_we2[warp] = 0.0;
    s_scan[warp] = 0ull;
  }
  __syncthreads();
  if (tid == 0) {
    // the CTA's 8 consecutive tiles: CTA-local exclusive prefixes and one total per CTA, so that the
    // last CTA only has to scan gridDim.x (<= 1024) values; Σe, Σe² in a fixed order (deterministic)
    unsigned long long acc = 0;
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int w = 0; w < kSumWarps; ++w) {
      const int tl = blockIdx.x * kSumWarps + w;
      if (tl < ix.ntiles) {
        ix.tile_lexcl[tl] = acc;
        ix.tile_tot[tl] = s_scan[w];
      }
      acc += s_scan[w];
      a += s_we[w];
      b += s_we2[w];
    }
    ix.cta_tot[blockIdx.x] = acc;
    psum[blockIdx.x] = a;
    psum2[blockIdx.x] = b;
    __threadfence();
    s_last = (atomicAdd(&ctrl->scan_done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  // ---- last CTA out: scan of the CTA totals, global tile offsets, Q, systematic offset, Σe, Σe²
  constexpr int PER = kMaxSumCtas / kSumThreads;  // 4 contiguous CTA totals per thread
  const int nctas = (int)gridDim.x;
  double a = 0.0, b = 0.0;
  unsigned long long cv[PER], run = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int j = tid * PER + k;
    cv[k] = 0;
    if (j < nctas) {
      a += __ldcg(&psum[j]);
      b += __ldcg(&psum2[j]);
      cv[k] = __ldcg(&ix.cta_tot[j]);
    }
    run += cv[k];
  }
  a = warp_sum(a);
  b = warp_sum(b);
  const unsigned long long winc = warp_scan_u64(run, lane);
  __syncthreads();  // s_we / s_scan are reused
  if (lane == 0) {
    s_we[warp] = a;
    s_we2[warp] = b;
  }
  if (lane == 31) s_scan[warp] = winc;
  __syncthreads();
  const unsigned long long wv = (lane < kSumWarps) ? s_scan[lane] : 0ull;
  const unsigned long long wvinc = warp_scan_u64(wv, lane);
  const unsigned long long wexcl = __shfl_sync(kFullMask, wvinc - wv, warp);
  const unsigned long long Q = __shfl_sync(kFullMask, wvinc, kSumWarps - 1);
  unsigned long long acc = wexcl + (winc - run);
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    s_cta_excl[tid * PER + k] = acc;
    acc += cv[k];
  }
  __syncthreads();
  // global tile offsets: independent, coalesced, and issued 8 deep — one L2 round trip per 2048 tiles, not per 256
  // (the rest of the GPU idles while this CTA finishes)
  constexpr int UNR = 8;
  for (int j0 = 0; j0 < ix.ntiles; j0 += kSumThreads * UNR) {
    unsigned long long le[UNR], tt[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + u * kSumThreads + tid;
      le[u] = (j < ix.ntiles) ? __ldcg(&ix.tile_lexcl[j]) : 0ull;
      tt[u] = (j < ix.ntiles) ? __ldcg(&ix.tile_tot[j]) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + u * kSumThreads + tid;
      if (j < ix.ntiles) {
        const unsigned long long ex = s_cta_excl[j / kSumWarps] + le[u];
        ix.tile_excl[j] = ex;
        ix.tile_incl[j] = ex + tt[u];
      }
    }
  }
  if (warp == 0) {
    double e1 = (lane < kSumWarps) ? s_we[lane] : 0.0, e2 = (lane < kSumWarps) ? s_we2[lane] : 0.0;
    e1 = warp_sum(e1);
    e2 = warp_sum(e2);
    if (lane == 0) {
      stats_out->mx = mx;
      stats_out->sum = e1;
      stats_out->sum2 = e2;
      ctrl->total = Q;
      ctrl->rq_lo = Rw * Q;
      ctrl->rq_hi = mulhi64(Rw, Q);
      ctrl->inv_rq = 1.0 / ((double)mulhi64(Rw, Q) + (double)(Rw * Q) * 0x1p-64);
      ctrl->sys_off = mulhi64(uniform64_at(key, 0u, stream, t, PURPOSE_RESAMPLE), Rw);  // used by the systematic resampler only
      ctrl->scan_done = 0;
      ctrl->maxslot[slot ^ 1] = encode_ordered(-INFINITY);
    }
  }
}

// two independent searches in lock step (one warp): the dependent round trips of the two overlap
__device__ __forceinline__ void warp_count_le2(const unsigned long long* __restrict__ v0, const unsigned long long* __restrict__ v1, int n0,
                                               int n1, uint64_t tau0, uint64_t tau1, int lane, int& r0, int& r1) {
  int lo0 = 0, hi0 = n0, lo1 = 0, hi1 = n1;
  while (hi0 > lo0 || hi1 > lo1) {
    const int st0 = (hi0 - lo0 + 31) >> 5, st1 = (hi1 - lo1 + 31) >> 5;
    const int p0 = lo0 + lane * st0 + (st0 - 1), p1 = lo1 + lane * st1 + (st1 - 1);
    const bool in0 = (hi0 > lo0) && (p0 < hi0), in1 = (hi1 > lo1) && (p1 < hi1);
    unsigned long long e0 = 0, e1 = 0;
    if (in0) e0 = __ldg(&v0[p0]);
    if (in1) e1 = __ldg(&v1[p1]);
    const int c0 = __popc(__ballot_sync(kFullMask, in0 && e0 <= tau0));
    const int c1 = __popc(__ballot_sync(kFullMask, in1 && e1 <= tau1));
    if (hi0 > lo0) {
      lo0 += c0 * st0;
      if (c0 == 32) hi0 = lo0;  // (only when every probe was in range and <= tau: the rest is empty)
      else { const int nh = lo0 + st0 - 1; hi0 = nh < hi0 ? nh : hi0; }
    }
    if (hi1 > lo1) {
      lo1 += c1 * st1;
      if (c1 == 32) hi1 = lo1;
      else { const int nh = lo1 + st1 - 1; hi1 = nh < hi1 ? nh : hi1; }
    }
  }
  r0 = lo0;
  r1 = lo1;
}

// One warp per TWO propagate-CTA boundaries (b and b + half), searched in lock step so that the whole
// grid is one wave of dependent round trips instead of two.  A separate launch so that these round
// trips are not on the critical path of every ancestor CTA.
__global__ void __launch_bounds__(256)
    bounds_kernel(StepIndex ix, const unsigned long long* __restrict__ cl, const FilterCtrl* ctrl, int N, int resampler,
                  uint64_t Rw, RngKey key, uint32_t stream, uint32_t t, int nbounds, int block_particles) {
  const int lane = threadIdx.x & 31;
  const int half = (nbounds + 1) >> 1;
  const int b0 = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b0 >= half) return;
  const int b1 = b0 + half;
  const bool has1 = b1 < nbounds;
  pdl_launch_dependents();
  pdl_wait();  // the tile index and Q come from sum_kernel
  const uint64_t Q = ctrl->total;
  if (Q == 0) return;
  uint64_t tau[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    int64_t i = (int64_t)(s ? (has1 ? b1 : b0) : b0) * block_particles;
    if (i > N - 1) i = N - 1;  // the last boundary is the last particle
    uint64_t u = ctrl->sys_off;
    if (resampler == RESAMPLE_STRATIFIED) u = uniform64_at(key, (uint32_t)i, stream, t, PURPOSE_RESAMPLE);
    tau[s] = threshold_of(resampler, (uint64_t)i, Rw, u, Q);
  }
  int T0, T1;
  warp_count_le2(ix.tile_incl, ix.tile_incl, ix.ntiles, ix.ntiles, tau[0], tau[1], lane, T0, T1);
  if (T0 > ix.ntiles - 1) T0 = ix.ntiles - 1;
  if (T1 > ix.ntiles - 1) T1 = ix.ntiles - 1;
  const uint64_t rem0 = tau[0] - __ldg(&ix.tile_excl[T0]), rem1 = tau[1] - __ldg(&ix.tile_excl[T1]);
  const int t0 = T0 * ix.tile_items, t1 = T1 * ix.tile_items;
  int n0 = N - t0, n1 = N - t1;
  if (n0 > ix.tile_items) n0 = ix.tile_items;
  if (n1 > ix.tile_items) n1 = ix.tile_items;
  int c0, c1;
  warp_count_le2(cl + t0, cl + t1, n0, n1, rem0, rem1, lane, c0, c1);
  if (c0 > n0 - 1) c0 = n0 - 1;
  if (c1 > n1 - 1) c1 = n1 - 1;
  if (lane == 0) {
    ix.bound_pos[b0] = t0 + c0;
    ix.bound_tile[b0] = T0;
    if (has1) {
      ix.bound_pos[b1] = t1 + c1;
      ix.bound_tile[b1] = T1;
    }
  }
}

// Systematic thresholds are an arithmetic progression in 128-bit fixed point: tau_i = hi64(A + i D),
// A = F_first Q, D = R Q.  So the FIRST particle of the CTA whose threshold reaches a CDF entry C,
//     o(C) = min{ i >= 0 : tau_i >= C } = ceil((C 2^64 - A) / D),
// needs no search: a double estimate of the quotient (absolute error < 1e-12 because D / 2^64 >= 1/2:
// the largest weight is 2^S exactly) decides it unless it falls within 1e-9 of an integer, and then the
// three candidates are compared exactly.  The ancestor of particle i is a_lo + #{ entries : o <= i }:
// a histogram of o over the CTA's window followed by a prefix sum — no window in shared memory, no
// dependent shared-memory chain, work proportional to entries + particles.
struct SysProgression {
  unsigned long long a_lo, a_hi, d_lo, d_hi;
  double inv, k;
};
__device__ __forceinline__ uint64_t sys_tau_exact(unsigned long long a_lo, unsigned long long a_hi, unsigned long long d_lo,
                                                  unsigned long long d_hi, int i) {
  const unsigned long long m = (unsigned long long)i;
  const unsigned long long plo = m * d_lo;
  const unsigned long long phi = m * d_hi + mulhi64(m, d_lo);
  const unsigned long long slo = a_lo + plo;
  return a_hi + phi + (slo < plo ? 1ull : 0ull);
}
// exact o(C) given that it lies in {n-1, ..., n+2}
__device__ __noinline__ int sys_first_exact(unsigned long long a_lo, unsigned long long a_hi, unsigned long long d_lo,
                                            unsigned long long d_hi, unsigned long long C, int n) {
  int o = n - 1;
#pragma unroll
  for (int i = -1; i <= 1; ++i) o += (n + i < 0 || sys_tau_exact(a_lo, a_hi, d_lo, d_hi, n + i) < C) ? 1 : 0;
  return o;
}
// exact floor((C 2^64 - A) / D) given that it lies in {n-1, n, n+1}: the last i with A + i D <= C 2^64
__device__ __noinline__ int strat_floor_exact(unsigned long long a_lo, unsigned long long a_hi, unsigned long long d_lo,
                                              unsigned long long d_hi, unsigned long long C, int n) {
  int is = n - 2;
#pragma unroll
  for (int i = -1; i <= 1; ++i) {
    bool le = n + i < 0;
    if (!le) {
      const unsigned long long m = (unsigned long long)(n + i);
      const unsigned long long plo = m * d_lo;
      const unsigned long long slo = a_lo + plo;
      const unsigned long long shi = a_hi + m * d_hi + mulhi64(m, d_lo) + (slo < plo ? 1ull : 0ull);
      le = shi < C || (shi == C && slo == 0);
    }
    is += le ? 1 : 0;
  }
  return is;
}
// estimate of o(C) = ceil(quotient): o = round(quotient + 1/2); `near` when the quotient is within eps of an
// integer (then o may be off by one and sys_first_exact(…, o - 1) decides).  sp.k = 1/2 - a_frac * inv.
__device__ __forceinline__ int sys_estimate(const SysProgression& sp, unsigned long long C, double near_thr, bool& near) {
  const double r05 = fma(__ull2double_rn(C - sp.a_hi), sp.inv, sp.k);
  const double tt = r05 + 0x1.8p52;
  const double g = r05 - (tt - 0x1.8p52);  // in [-1/2, 1/2]: distance of quotient + 1/2 from its nearest integer
  near = fabs(g) > near_thr;               // near_thr = 1/2 - eps
  return __double2loint(tt);
}
// ---- the ancestor kernel of the sorted resamplers (particles.jl:117 with SPEC §5's thresholds).
// One CTA resolves kSysParticles consecutive particles.  It reads the CDF entries [a_lo, a_hi) between its
// own first ancestor and the next CTA's (bounds_kernel), adds 1 to hist[o(C)] for each, o(C) = the first
// local particle whose threshold reaches C, and the ancestor of local particle i is
// a_lo + sum_{o <= i} hist[o].  No search, no window in shared memory.
//   systematic: o(C) = ceil((C 2^64 - A) / D), A = F_first Q, D = R Q (SysProgression above);
//   stratified: the thresholds are one per stratum of width R: with s = floor((C 2^64 - A) / D), A = first R Q,
//               every earlier stratum's threshold is below C, every later one's is not, and stratum s
//               itself is decided by its own uniform: o(C) = s + [tau_s < C]  (one Philox block per entry).
// Very uneven weights make some windows long and mostly dead (no threshold falls into them).  Whole tiles
// of such a window are accounted for with ONE add when o(first) == o(last) — read off the tile index,
// not the entries — so a CTA streams at most the tiles that contain one of its thresholds.
//
// hist[o] += 1 when pred, as ONE predicated instruction on a shared-space address (an if around atomicAdd
// compiles to a divergence region per call and re-derives the shared window base each time)
__device__ __forceinline__ void hist_inc(uint32_t hist_saddr, int o, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(hist_saddr + 4u * (uint32_t)o), "r"((int)pred) : "memory");
}
template <int RESAMPLER>
__global__ void __launch_bounds__(kSysThreads, 12)
    anc_hist_kernel(int N, uint64_t Rw, RngKey key, uint32_t stream, uint32_t t, StepIndex ix, const unsigned long long* __restrict__ cl,
                    int32_t* __restrict__ anc_out, const FilterCtrl* __restrict__ ctrl, double eps) {
  constexpr bool STRAT = RESAMPLER == RESAMPLE_STRATIFIED;
  constexpr int NW = kSysThreads / 32, NRUN = kSysPer / 4;
  __shared__ __align__(16) int s_cnt[kSysParticles + 4];
  __shared__ int s_wtot[NRUN][NW];
  __shared__ unsigned s_skip[kMaxTiles / 32];  // wide windows only: tiles already accounted for wholesale
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t hist;
  asm volatile("mov.u32 %0, %1;" : "=r"(hist) : "r"((uint32_t)__cvta_generic_to_shared(s_cnt)));  // opaque: kept in a register
#pragma unroll
  for (int r = 0; r < NRUN; ++r) reinterpret_cast<int4*>(s_cnt)[r * kSysThreads + tid] = make_int4(0, 0, 0, 0);
  if (tid == 0) reinterpret_cast<int4*>(s_cnt)[NRUN * kSysThreads] = make_int4(0, 0, 0, 0);
  pdl_launch_dependents();
  pdl_wait();  // the window bounds come from bounds_kernel
  // everything the CTA needs from the two producer kernels in ONE round trip (no load behind a branch)
  const uint64_t Q = ctrl->total;
  const int a_lo = __ldg(&ix.bound_pos[blockIdx.x]);
  const int a_hi = __ldg(&ix.bound_pos[blockIdx.x + 1]);  // ancestor of the next CTA's first particle >= all of mine
  const int T0 = __ldg(&ix.bound_tile[blockIdx.x]), T1 = __ldg(&ix.bound_tile[blockIdx.x + 1]);
  const uint64_t c_off = ctrl->sys_off, c_rq_lo = ctrl->rq_lo, c_rq_hi = ctrl->rq_hi;
  const double c_inv = ctrl->inv_rq;
  const int base_i = blockIdx.x * kSysParticles;
  const bool full_cta = (int64_t)(blockIdx.x + 1) * kSysParticles <= (int64_t)N;

  int v[NRUN][4];
  int off[NRUN];
  if (Q == 0) {  // every particle is its own ancestor (SPEC §5)
#pragma unroll
    for (int r = 0; r < NRUN; ++r) {
      off[r] = base_i + r * 4 * kSysThreads + 4 * tid;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[r][k] = k;
    }
  } else {
    SysProgression sp;
    {
      const uint64_t F0 = (uint64_t)base_i * Rw + (STRAT ? 0ull : c_off);
      sp.a_lo = F0 * Q;
      sp.a_hi = mulhi64(F0, Q);
      sp.d_lo = c_rq_lo;
      sp.d_hi = c_rq_hi;
      sp.inv = c_inv;
      sp.k = (STRAT ? -0.5 : 0.5) - __ull2double_rn(sp.a_lo) * 0x1p-64 * c_inv;  // round(q - 1/2) = floor, round(q + 1/2) = ceil
    }
    // o(C) from the estimate e (and, stratified, the stratum's own threshold); valid entries only
    auto first_reaching = [&](unsigned long long C, int e, bool& in_range) -> int {
      if (!STRAT) {
        in_range = (unsigned)e < (unsigned)kSysParticles;
        return e;
      }
      in_range = false;
      if ((unsigned)e >= (unsigned)kSysParticles || base_i + e >= N) return 0;  // a stratum of a later CTA (or past the last one)
      const uint64_t i = (uint64_t)(base_i + e);
      const uint64_t tau = threshold_of(RESAMPLE_STRATIFIED, i, Rw, uniform64_at(key, (uint32_t)i, stream, t, PURPOSE_RESAMPLE), Q);
      const int o = e + (tau < C ? 1 : 0);
      in_range = o < kSysParticles;
      return o;
    };
    auto exact = [&](unsigned long long C, int e) -> int {
      return STRAT ? strat_floor_exact(sp.a_lo, sp.a_hi, sp.d_lo, sp.d_hi, C, e) : sys_first_exact(sp.a_lo, sp.a_hi, sp.d_lo, sp.d_hi, C, e - 1);
    };
    const double near_thr = 0.5 - eps;
    const bool wide = T1 - T0 >= 3;
    if (wide) {
      for (int w = tid; w < kMaxTiles / 32; w += kSysThreads) s_skip[w] = 0u;
    }
    __syncthreads();  // the histogram (and the bitmap) is zero
    if (wide) {
      for (int T = T0 + 1 + tid; T < T1; T += kSysThreads) {  // tiles strictly inside the window: full tiles, all entries in [a_lo, a_hi)
        int o2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const unsigned long long C = __ldg(e ? &ix.tile_incl[T] : &ix.tile_excl[T]);  // = CDF of a real entry of the window, both
          bool nr, in;
          int est = sys_estimate(sp, C, near_thr, nr);
          if (nr) est = exact(C, est);
          o2[e] = first_reaching(C, est, in);
          if (!in) o2[e] = kSysParticles;  // (every entry of the window is reached by the next CTA's first particle at the latest)
        }
        if (o2[0] == o2[1]) {
          if (o2[1] < kSysParticles) atomicAdd(&s_cnt[o2[1]], ix.tile_items);
          atomicOr(&s_skip[T >> 5], 1u << (T & 31));
        }
      }
      __syncthreads();
    }
    const int s0 = a_lo & ~1;
    bool near = false;  // an estimate too close to an integer to decide: resolved exactly in a second pass
    auto sweep = [&](auto exact_pass) {
      constexpr bool EXACT = decltype(exact_pass)::value;
      int T = T0, tlo = s0;
      while (tlo < a_hi) {
        const int tend_full = (T + 1) * ix.tile_items;
        const int thi = tend_full < a_hi ? tend_full : a_hi;
        if (!(wide && ((s_skip[T >> 5] >> (T & 31)) & 1u))) {
          const unsigned long long base = __ldg(&ix.tile_excl[T]);
          auto tally = [&](bool valid, unsigned long long C) {
            bool nr, in;
            const int est = sys_estimate(sp, C, near_thr, nr);
            if (EXACT) {
              if (valid && nr) {
                const int o = first_reaching(C, exact(C, est), in);
                if (in) atomicAdd(&s_cnt[o], 1);
              }
            } else {
              near |= valid && nr;
              if (STRAT) {
                if (valid && !nr) {
                  const int o = first_reaching(C, est, in);
                  hist_inc(hist, o, in);
                }
              } else {
                hist_inc(hist, est, valid && !nr && (unsigned)est < (unsigned)kSysParticles);
              }
            }
          };
          for (int j = tlo + 2 * tid; j < thi; j += 4 * kSysThreads) {  // tlo even, tile_items even; two loads in flight
            const int j1 = j + 2 * kSysThreads;
            const bool has1 = j1 < thi;
            const ulonglong2 v0 = __ldcs(reinterpret_cast<const ulonglong2*>(cl + j));
            ulonglong2 v1 = make_ulonglong2(0, 0);
            if (has1) v1 = __ldcs(reinterpret_cast<const ulonglong2*>(cl + j1));
            tally(j >= a_lo, v0.x + base);
            tally(j + 1 < thi, v0.y + base);
            tally(has1, v1.x + base);
            tally(has1 && j1 + 1 < thi, v1.y + base);
          }
        }
        tlo = thi;
        ++T;
      }
    };
    sweep(std::false_type{});
    if (__syncthreads_or(near)) {
      sweep(std::true_type{});
      __syncthreads();
    }
    // prefix sums: run r of thread tid covers local particles r * 4 * kSysThreads + 4 tid .. + 3 (conflict-free
    // 16-byte reads of the histogram, fully coalesced 16-byte stores of the ancestors)
    int incl[NRUN];
#pragma unroll
    for (int r = 0; r < NRUN; ++r) {
      const int4 c = reinterpret_cast<const int4*>(s_cnt)[r * kSysThreads + tid];
      v[r][0] = c.x;
      v[r][1] = c.x + c.y;
      v[r][2] = c.x + c.y + c.z;
      v[r][3] = c.x + c.y + c.z + c.w;
      incl[r] = v[r][3];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
      for (int r = 0; r < NRUN; ++r) {
        const int u = __shfl_up_sync(kFullMask, incl[r], o);
        if (lane >= o) incl[r] += u;
      }
    }
    if (lane == 31) {
#pragma unroll
      for (int r = 0; r < NRUN; ++r) s_wtot[r][warp] = incl[r];
    }
    __syncthreads();
    int run_base = a_lo;
#pragma unroll
    for (int r = 0; r < NRUN; ++r) {
      off[r] = run_base + incl[r] - v[r][3];
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const int tw = s_wtot[r][w];
        if (w < warp) off[r] += tw;
        run_base += tw;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NRUN; ++r) {
    const int i = base_i + r * 4 * kSysThreads + 4 * tid;
    if (full_cta) {
      *reinterpret_cast<int4*>(anc_out + i) = make_int4(off[r] + v[r][0], off[r] + v[r][1], off[r] + v[r][2], off[r] + v[r][3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i + k < N) anc_out[i + k] = off[r] + v[r][k];
    }
  }
}

// ================================================================================================
// Multinomial resampling of a large cloud (N > kMnLegacyMax) — the reference's law, `sample(1:n, Weights(w), n)`
// (particles.jl:17-19), without N random walks through a CDF that does not fit the L2 (docs/SPEC.md §5c).
// The offspring counts of i.i.d. thresholds are multinomial whatever order the thresholds are looked at in, and
// the particles of a resampled cloud are exchangeable, so the draw is organised in two levels:
//   level 1  the N thresholds tau_i = mulhi(U(i), Q) are only COUNTED per cell of kMnCell consecutive particles
//            (mn_count_kernel: an interpolation guess + a short gallop in the <= 32 KB cell index, a shared-memory
//            histogram per CTA) -> K_c, and their prefix O_c;
//   level 2  given K_c, the thresholds that fell into cell c are i.i.d. uniform on the cell's mass: output
//            position g in [O_c, O_c + K_c) draws its own V(g) (a second Philox purpose), searches the cell-local
//            CDF held in shared memory, and every chunk of kMnChunk consecutive output positions of a cell is
//            written in ascending ancestor order (mn_cell_kernel: histogram + prefix sum, no sort).
// The ancestor vector is piecewise sorted (cells in order), so move_kernel's gather is as coalesced as for the
// systematic resampler; every step is defined by particle / output indices only, not by the launch geometry, so the
// oracle reproduces the ancestors bit for bit.  sum_kernel (tile-local CDF, Q, normalize statistics) is shared
// with the sorted resamplers:  sum -> mn_prep -> mn_count -> mn_cell -> move.
constexpr int kMnCell = 4096;
constexpr int kMnChunk = 8192;
constexpr int kMnLegacyMax = 8192;      // N <= this: per-particle search of the materialised CDF, unsorted ancestors (SPEC §5 row 0)
constexpr int kMnCountThreads = 512;
constexpr int kMnSmemCells = 8192;      // cells whose counters fit a CTA's shared memory (N <= 2^25)
constexpr int kMnCellThreads = 512;

struct MnIndex {
  unsigned long long* cellC;  // [ncells] global CDF at the last particle of every cell
  int32_t* K;                 // [ncells] thresholds that fell into the cell
  int32_t* O;                 // [ncells] exclusive prefix of K: first output position of the cell
  int32_t* part_start;        // [ncells + 1] exclusive prefix of max(1, ceil(K_c / kMnChunk)): first work item of the cell
  int32_t* lut;               // [kMnLutMax + 1] bucket of a threshold (its top bits) -> first cell whose boundary is not below the bucket
  unsigned int* ticket;       // last-CTA-out counter of mn_count_kernel
  int ncells;
};
constexpr int kMnLutBits = 13;
constexpr int kMnLutMax = 1 << kMnLutBits;  // buckets of the level-1 lookup table
constexpr int kMnSubBits = 12;
constexpr int kMnSub = 1 << kMnSubBits;     // outputs per sub-pass of mn_cell_kernel = buckets of its counting sort

// floor(u * v / 2^32) for a 32-bit uniform u and v < 2^61 (exact: the two partial products fit 64 bits)
__device__ __forceinline__ uint64_t mulhi32_64(uint32_t u, uint64_t v) {
  return (uint64_t)u * (v >> 32) + (((uint64_t)u * (v & 0xFFFFFFFFull)) >> 32);
}

// number of low bits dropped so that (v >> shift) < 2^bits for every v <= vmax
__device__ __forceinline__ int mn_shift(unsigned long long vmax, int bits) {
  const int len = 64 - __clzll((long long)(vmax | 1ull));
  return len > bits ? len - bits : 0;
}

// cellC[c] = C at the last particle of cell c (tile-local CDF + tile offset); K[c] = 0; and the level-1 lookup table:
// lut[b] = #{ c : (cellC_c >> s) < b } — every one of those cells ends below any threshold of bucket b
__global__ void mn_prep_kernel(MnIndex mn, StepIndex ix, const unsigned long long* __restrict__ cl, const FilterCtrl* __restrict__ ctrl, int N) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();  // the tile-local CDF, the tile offsets and Q come from sum_kernel
  if (c >= mn.ncells) return;
  auto cell_cdf = [&](int cc) -> unsigned long long {
    int64_t last = (int64_t)(cc + 1) * kMnCell - 1;
    if (last > N - 1) last = N - 1;
    return __ldg(&ix.tile_excl[(int)(last / ix.tile_items)]) + __ldg(&cl[last]);
  };
  const unsigned long long C = cell_cdf(c);
  mn.cellC[c] = C;
  mn.K[c] = 0;
  const unsigned long long Q = ctrl->total;
  if (Q == 0) return;
  const int s = mn_shift(Q, kMnLutBits);
  const int b_hi = (int)(C >> s);
  const int b_lo = c ? (int)(cell_cdf(c - 1) >> s) + 1 : 0;
  for (int b = b_lo; b <= b_hi; ++b) mn.lut[b] = c;
}

// idx / tile_items without an integer division: tile_items = 128 cpt, so the quotient is (idx >> 7) / cpt, estimated in binary32
// (idx >> 7 < 2^24 is exact in a float) and corrected by at most one
__device__ __forceinline__ int mn_tile_of(int idx, int cpt, float inv_cpt) {
  const int x = idx >> 7;
  int q = (int)((float)x * inv_cpt);
  const int r = x - q * cpt;
  q += (r >= cpt) ? 1 : 0;
  q -= (r < 0) ? 1 : 0;
  return q;
}

// level 1: K[c] = #{ i : tau_i in cell c }; the last CTA out turns K into output offsets and work items.
// A threshold's cell is read off a lookup table indexed by its top bits (about two buckets per cell) and fixed by at most a step or
// two through the cell boundaries; table, boundaries and counters all live in shared memory.
__global__ void __launch_bounds__(kMnCountThreads)
    mn_count_kernel(MnIndex mn, const FilterCtrl* __restrict__ ctrl, int N, RngKey key, uint32_t stream, uint32_t t) {
  extern __shared__ __align__(16) unsigned char mn_cnt_smem[];
  unsigned long long* s_cell = reinterpret_cast<unsigned long long*>(mn_cnt_smem);   // [ncells] when ncells <= kMnSmemCells
  const int ncells = mn.ncells;
  const bool in_smem = ncells <= kMnSmemCells;
  int* s_lut = reinterpret_cast<int*>(mn_cnt_smem + (in_smem ? sizeof(unsigned long long) * (size_t)ncells : 0));  // [kMnLutMax + 1]
  int* s_hist = s_lut + kMnLutMax + 1;                                                // [ncells] when ncells <= kMnSmemCells
  __shared__ bool s_last;
  __shared__ int s_wtot[2][kMnCountThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (in_smem)
    for (int c = tid; c < ncells; c += kMnCountThreads) s_hist[c] = 0;
  pdl_launch_dependents();
  pdl_wait();  // cellC, the lookup table and K = 0 come from mn_prep_kernel, Q from sum_kernel
  const uint64_t Q = ctrl->total;
  const int s = mn_shift(Q, kMnLutBits);
  if (Q != 0) {
    const int nb = (int)(Q >> s) + 1;
    for (int b = tid; b < nb; b += kMnCountThreads) s_lut[b] = __ldg(&mn.lut[b]);
    if (in_smem)
      for (int c = tid; c < ncells; c += kMnCountThreads) s_cell[c] = __ldg(&mn.cellC[c]);
  }
  __syncthreads();
  const int nquads = (N + 3) >> 2;
  for (int p = blockIdx.x * kMnCountThreads + tid; p < nquads; p += gridDim.x * kMnCountThreads) {
    // SPEC §5c level 1: threshold i uses word (i & 3) of the Philox block at index i >> 2 (a 32-bit uniform is ample for a count per cell)
    const Philox4 b = philox4x32_10((uint32_t)p, stream, t, purpose_word(PURPOSE_RESAMPLE, 0, key.epoch), key);
    const uint32_t u4[4] = {b.r0, b.r1, b.r2, b.r3};
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int i = 4 * p + h;
      if (i >= N) break;
      int c;
      if (Q != 0) {
        const uint64_t tau = mulhi32_64(u4[h], Q);
        c = s_lut[(int)(tau >> s)];
        // cells whose boundary shares the threshold's bucket (cellC[ncells - 1] = Q > tau ends the walk)
        if (in_smem) {
          while (s_cell[c] <= tau) ++c;
        } else {
          while (__ldg(&mn.cellC[c]) <= tau) ++c;
        }
      } else {
        c = i / kMnCell;  // no mass at all: every particle is its own ancestor (SPEC §5)
      }
      if (in_smem) atomicAdd(&s_hist[c], 1);
      else atomicAdd(&mn.K[c], 1);
    }
  }
  __syncthreads();
  if (in_smem) {
    // every CTA starts its flush at a different cell so that the CTAs do not queue up on the same counters
    const int rot = (int)(((unsigned)blockIdx.x * 2654435761u) % (unsigned)ncells);
    for (int k = tid; k < ncells; k += kMnCountThreads) {
      int c = k + rot;
      c -= (c >= ncells) ? ncells : 0;
      const int v = s_hist[c];
      if (v) atomicAdd(&mn.K[c], v);
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(mn.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  // ---- last CTA out: O = exclusive prefix of K, part_start = exclusive prefix of the work items per cell
  __threadfence();
  int run_k = 0, run_p = 0;
  for (int c0 = 0; c0 < ncells; c0 += kMnCountThreads) {
    const int c = c0 + tid;
    const int k = c < ncells ? __ldcg(&mn.K[c]) : 0;
    const int np = c < ncells ? (k > kMnChunk ? (k + kMnChunk - 1) / kMnChunk : 1) : 0;
    int ik = k, ip = np;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int uk = __shfl_up_sync(kFullMask, ik, o), up = __shfl_up_sync(kFullMask, ip, o);
      if (lane >= o) { ik += uk; ip += up; }
    }
    __syncthreads();
    if (lane == 31) { s_wtot[0][warp] = ik; s_wtot[1][warp] = ip; }
    __syncthreads();
    int wk = 0, wp = 0, tk = 0, tp = 0;
#pragma unroll
    for (int w = 0; w < kMnCountThreads / 32; ++w) {
      const int a = s_wtot[0][w], b2 = s_wtot[1][w];
      if (w < warp) { wk += a; wp += b2; }
      tk += a; tp += b2;
    }
    if (c < ncells) {
      mn.O[c] = run_k + wk + ik - k;
      mn.part_start[c] = run_p + wp + ip - np;
    }
    run_k += tk;
    run_p += tp;
  }
  if (tid == 0) {
    mn.part_start[ncells] = run_p;
    *mn.ticket = 0u;
  }
}

// level 2: one CTA per (cell, chunk of kMnChunk output positions).  The cell-local CDF sits in shared memory together with a
// lookup table over its top bits (about one bucket per particle): a threshold finds its particle with one table read and a step
// or two along the CDF — no binary search — and bumps the particle's offspring counter.  The inclusive prefix R_j of the counters
// then says that outputs R_{j-1} .. R_j - 1 belong to particle j: head flags + a max-scan write the chunk in ascending order.
// Thread `tid` owns the 8 consecutive particles 8 tid .. 8 tid + 7 throughout.
__global__ void __launch_bounds__(kMnCellThreads, 3)
    mn_cell_kernel(MnIndex mn, StepIndex ix, const unsigned long long* __restrict__ cl, const FilterCtrl* __restrict__ ctrl, int N, RngKey key,
                   uint32_t stream, uint32_t t, int32_t* __restrict__ anc) {
  extern __shared__ __align__(16) unsigned char mn_smem[];
  unsigned long long* s_c = reinterpret_cast<unsigned long long*>(mn_smem);             // [kMnCell] cell-local inclusive CDF
  int* s_head = reinterpret_cast<int*>(mn_smem);                                        // [kMnChunk] (after the counting) head flags
  int* s_lut = reinterpret_cast<int*>(mn_smem + sizeof(unsigned long long) * kMnCell);  // [kMnSub + 1]
  int* s_hist = s_lut + kMnSub + 4;                                                     // [kMnCell] offspring counts (16-byte aligned)
  __shared__ int s_w[kMnCellThreads / 32];
  __shared__ int s_item[2];
  constexpr int PER = kMnCell / kMnCellThreads;  // 8 particles per thread
  constexpr int NW = kMnCellThreads / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_launch_dependents();
  pdl_wait();  // K, O, part_start come from mn_count_kernel
  const int ncells = mn.ncells;
  const int item = blockIdx.x;
  int c = item, part = 0;
  if (item >= ncells) {  // work items beyond the first chunk of every cell exist only when some K_c > kMnChunk
    const int total = __ldg(&mn.part_start[ncells]);
    if (item >= total) return;
    const int want = item - ncells;  // the want-th extra chunk: cell c, part >= 1 with (part_start[c] - c) + part - 1 = want
    if (tid == 0) {
      int lo = 0, hi = ncells - 1;  // last c with (part_start[c] - c) <= want
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&mn.part_start[mid]) - mid <= want) lo = mid;
        else hi = mid - 1;
      }
      s_item[0] = lo;
      s_item[1] = want - (__ldg(&mn.part_start[lo]) - lo) + 1;
    }
    __syncthreads();
    c = s_item[0];
    part = s_item[1];
  }
  const int Kc = __ldg(&mn.K[c]), Oc = __ldg(&mn.O[c]);
  const int g0 = Oc + part * kMnChunk;
  const int g1 = (Oc + Kc < g0 + kMnChunk) ? Oc + Kc : g0 + kMnChunk;
  const int n_out = g1 - g0;
  if (n_out <= 0) return;
  const int j0 = c * kMnCell;
  const int len = (N - j0 < kMnCell) ? N - j0 : kMnCell;
  const uint64_t Q = ctrl->total;
  if (Q == 0) {  // identity ancestors
    for (int r = tid; r < n_out; r += kMnCellThreads) anc[g0 + r] = j0 + part * kMnChunk + r;
    return;
  }
  const unsigned long long base = c ? __ldg(&mn.cellC[c - 1]) : 0ull;
  const unsigned long long W = __ldg(&mn.cellC[c]) - base;
  const int s = mn_shift(W, kMnSubBits);
  // ---- the cell-local CDF of this thread's 8 particles (past the end of a ragged cell: W, so that they own no threshold)
  const int jf = tid * PER;
  unsigned long long C[PER];
  {
    const int i_first = j0 + jf;                     // multiple of 8: the 8 entries never straddle more than one tile boundary
    const int cpt = ix.tile_items >> 7;
    const int T0 = mn_tile_of(i_first < N ? i_first : N - 1, cpt, 1.0f / (float)cpt);
    const int tile_end = (T0 + 1) * ix.tile_items;
    const unsigned long long e0 = __ldg(&ix.tile_excl[T0]) - base;
    const unsigned long long e1 = (tile_end < i_first + PER && tile_end < N) ? __ldg(&ix.tile_excl[T0 + 1]) - base : e0;
    if (jf + PER <= len) {
#pragma unroll
      for (int k = 0; k < PER; k += 2) {
        const ulonglong2 v = __ldcs(reinterpret_cast<const ulonglong2*>(cl + i_first + k));
        C[k] = v.x + (i_first + k < tile_end ? e0 : e1);
        C[k + 1] = v.y + (i_first + k + 1 < tile_end ? e0 : e1);
      }
    } else {
#pragma unroll
      for (int k = 0; k < PER; ++k) C[k] = (jf + k < len) ? __ldcs(&cl[i_first + k]) + (i_first + k < tile_end ? e0 : e1) : W;
    }
#pragma unroll
    for (int k = 0; k < PER; k += 2) reinterpret_cast<ulonglong2*>(s_c)[(jf + k) >> 1] = make_ulonglong2(C[k], C[k + 1]);
    reinterpret_cast<int4*>(s_hist)[2 * tid] = make_int4(0, 0, 0, 0);
    reinterpret_cast<int4*>(s_hist)[2 * tid + 1] = make_int4(0, 0, 0, 0);
    reinterpret_cast<int4*>(s_lut)[2 * tid] = make_int4(0, 0, 0, 0);
    reinterpret_cast<int4*>(s_lut)[2 * tid + 1] = make_int4(0, 0, 0, 0);
  }
  __syncthreads();
  // ---- lookup table: lut[b] = #{ j : (C_j >> s) < b } = the first particle whose bucket C_j >> s is not below b.  Particle j OWNS
  // the buckets (C_{j-1} >> s, C_j >> s]: it writes its index into the first one it owns (no two particles share a first bucket)
  // and a max-scan over the buckets hands it the rest — no per-particle loop over buckets, so neither the divergence of a loop
  // whose length varies from lane to lane nor a special path for a particle that holds most of the cell's mass.
  // (the table was zeroed with the counters above; thread `tid` scans the 8 buckets 8 tid .. 8 tid + 7)
  if (jf < len) {
    int bprev = jf ? (int)(s_c[jf - 1] >> s) : -1;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      if (jf + k < len) {
        const int b_hi = (int)(C[k] >> s);
        if (b_hi > bprev) s_lut[bprev + 1] = jf + k;
        bprev = b_hi;
      }
    }
  }
  __syncthreads();
  {
    static_assert(kMnSub == kMnCellThreads * 8, "one thread scans 8 buckets");
    const int4 l0 = reinterpret_cast<const int4*>(s_lut)[2 * tid], l1 = reinterpret_cast<const int4*>(s_lut)[2 * tid + 1];
    int v[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
    for (int k = 1; k < 8; ++k) v[k] = v[k] > v[k - 1] ? v[k] : v[k - 1];
    int inc = v[7];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(kFullMask, inc, o);
      if (lane >= o) inc = u > inc ? u : inc;
    }
    const int wprev = __shfl_up_sync(kFullMask, inc, 1);
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int carry = lane ? wprev : 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
      if (w < warp) carry = s_w[w] > carry ? s_w[w] : carry;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = v[k] > carry ? v[k] : carry;
    reinterpret_cast<int4*>(s_lut)[2 * tid] = make_int4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<int4*>(s_lut)[2 * tid + 1] = make_int4(v[4], v[5], v[6], v[7]);
  }
  __syncthreads();  // (also: s_w is reused by the prefix sum below)
  // ---- in-cell thresholds (SPEC §5c level 2): output position g uses word (g & 3) of the Philox block at index g >> 2
  for (int p = (g0 >> 2) + tid; 4 * p < g1; p += kMnCellThreads) {
    const Philox4 blk = philox4x32_10((uint32_t)p, stream, t, purpose_word(PURPOSE_RESAMPLE_CELL, 0, key.epoch), key);
    const uint32_t v4[4] = {blk.r0, blk.r1, blk.r2, blk.r3};
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int g = 4 * p + h;
      if (g >= g0 && g < g1) {
        const uint64_t tau = mulhi32_64(v4[h], W);
        int j = s_lut[(int)(tau >> s)];
        while (s_c[j] <= tau) ++j;  // (C of the cell's last particle = W > tau ends the walk)
        atomicAdd(&s_hist[j], 1);
      }
    }
  }
  __syncthreads();
  // ---- R_j = inclusive prefix of the offspring counts
  int R[PER], Rm1;
  {
    const int4 h0 = reinterpret_cast<const int4*>(s_hist)[2 * tid], h1 = reinterpret_cast<const int4*>(s_hist)[2 * tid + 1];
    R[0] = h0.x; R[1] = R[0] + h0.y; R[2] = R[1] + h0.z; R[3] = R[2] + h0.w;
    R[4] = R[3] + h1.x; R[5] = R[4] + h1.y; R[6] = R[5] + h1.z; R[7] = R[6] + h1.w;
    int inc = R[7];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(kFullMask, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
      if (w < warp) woff += s_w[w];
    Rm1 = woff + inc - R[7];
#pragma unroll
    for (int k = 0; k < PER; ++k) R[k] += Rm1;
  }
  // ---- head flags (they reuse the CDF's shared memory: every thread is past the counting) and the max-scan
  for (int r = tid; r < n_out; r += kMnCellThreads) s_head[r] = 0;
  __syncthreads();
  {
    int prev = Rm1;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      if (R[k] > prev) s_head[prev] = jf + k;
      prev = R[k];
    }
  }
  __syncthreads();
  {
    constexpr int OUT = kMnChunk / kMnCellThreads;  // 16 consecutive outputs per thread
    const int r0 = tid * OUT;
    int v[OUT], run = 0;
    if (r0 < n_out) {
#pragma unroll
      for (int q = 0; q < OUT / 4; ++q) {
        const int4 hv = reinterpret_cast<const int4*>(s_head)[tid * (OUT / 4) + q];  // (entries past n_out may hold stale CDF bits: masked)
        const int e[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int val = (r0 + 4 * q + k < n_out) ? e[k] : 0;
          run = val > run ? val : run;
          v[4 * q + k] = run;
        }
      }
    }
    int inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(kFullMask, inc, o);
      if (lane >= o) inc = u > inc ? u : inc;
    }
    const int wprev = __shfl_up_sync(kFullMask, inc, 1);
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int carry = lane ? wprev : 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
      if (w < warp) carry = s_w[w] > carry ? s_w[w] : carry;
    if (r0 + OUT <= n_out && ((g0 + r0) & 3) == 0) {
#pragma unroll
      for (int q = 0; q < OUT / 4; ++q) {
        int4 o4;
        o4.x = j0 + (v[4 * q] > carry ? v[4 * q] : carry);
        o4.y = j0 + (v[4 * q + 1] > carry ? v[4 * q + 1] : carry);
        o4.z = j0 + (v[4 * q + 2] > carry ? v[4 * q + 2] : carry);
        o4.w = j0 + (v[4 * q + 3] > carry ? v[4 * q + 3] : carry);
        *reinterpret_cast<int4*>(anc + g0 + r0 + 4 * q) = o4;
      }
    } else if (r0 < n_out) {
#pragma unroll
      for (int k = 0; k < OUT; ++k)
        if (r0 + k < n_out) anc[g0 + r0 + k] = j0 + (v[k] > carry ? v[k] : carry);
    }
  }
}

// x_i ~ transition(x[a_i]); logw_i = logpdf(observation(x_i), y)   (particles.jl:119-125): a streaming
// pass, kMovePairs pairs (4 consecutive particles) per thread for two independent Philox / Box-Muller
// chains; parents gathered through the (sorted, hence near-coalesced) ancestor vector.
constexpr int kMoveThreads = 256;
constexpr int kMovePairs = 2;
template <class Model, bool FULL, class XT>
__device__ __forceinline__ double move_particles(const Model& mdl, double y, int N, int64_t ld, const RngKey& key, uint32_t stream, uint32_t t,
                                                 int i0, const int32_t* __restrict__ anc, const XT* __restrict__ xprev,
                                                 XT* __restrict__ xnew, double* __restrict__ logw /* null: not stored */) {
  constexpr int D = Model::D;
  constexpr int PER = 2 * kMovePairs;
  int a[PER];
  if (FULL) {
#pragma unroll
    for (int q = 0; q < PER / 2; ++q) {
      const int2 v = __ldcs(reinterpret_cast<const int2*>(anc + i0) + q);
      a[2 * q] = v.x; a[2 * q + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < PER; ++k) a[k] = (i0 + k < N) ? anc[i0 + k] : 0;
  }
  double xp[PER][D];
  if (D == 1) {  // all parents requested before the first normal is drawn
#pragma unroll
    for (int k = 0; k < PER; ++k) xp[k][0] = (double)__ldg(&xprev[a[k]]);
  }
  double vmax = -INFINITY;
#pragma unroll
  for (int r = 0; r < kMovePairs; ++r) {
    const int i = i0 + 2 * r;
    if (!FULL && i >= N) break;
    if (D > 1) {
#pragma unroll
      for (int c = 0; c < D; ++c) {
        xp[2 * r][c] = (double)__ldg(&xprev[c * ld + a[2 * r]]);
        xp[2 * r + 1][c] = (double)__ldg(&xprev[c * ld + a[2 * r + 1]]);
      }
    }
    double za[D], zb[D], xa[D], xb[D];
#pragma unroll
    for (int c = 0; c < D; ++c) normal_pair_at(key, (uint32_t)(i >> 1), stream, t, PURPOSE_TRANSITION, (uint32_t)c, za[c], zb[c]);
    mdl.transition(za, xp[2 * r], xa);
    mdl.transition(zb, xp[2 * r + 1], xb);
    round_state<XT>(xa);
    round_state<XT>(xb);
    const double la = mdl.logweight(xa, y);
    const double lb = mdl.logweight(xb, y);
    if (FULL || i + 1 < N) {
#pragma unroll
      for (int c = 0; c < D; ++c) *reinterpret_cast<typename XVec2<XT>::type*>(xnew + c * ld + i) = XVec2<XT>::make(xa[c], xb[c]);
      if (logw) *reinterpret_cast<double2*>(logw + i) = make_double2(la, lb);
      if (la > vmax) vmax = la;  // `>` ignores NaN like the CPU loop
      if (lb > vmax) vmax = lb;
    } else {
#pragma unroll
      for (int c = 0; c < D; ++c) xnew[c * ld + i] = (XT)xa[c];
      if (logw) logw[i] = la;
      if (la > vmax) vmax = la;
    }
  }
  return vmax;
}

template <class Model, class XT>
__global__ void __launch_bounds__(kMoveThreads, (Model::D == 1) ? 6 : 3)
    move_kernel(typename Model::DV dv, double y, int N, int64_t ld, RngKey key, uint32_t stream, uint32_t t, const int32_t* __restrict__ anc,
                const XT* __restrict__ xprev, XT* __restrict__ xnew, double* __restrict__ logw, FilterCtrl* ctrl) {
  constexpr int PER = 2 * kMovePairs;
  constexpr int NW = kMoveThreads / 32;
  __shared__ unsigned long long s_max[NW];
  const int tid = threadIdx.x;
  Model mdl;
  mdl.load(dv.d);
  const int i0 = (blockIdx.x * kMoveThreads + tid) * PER;
  pdl_launch_dependents();
  pdl_wait();  // the ancestors come from anc_hist_kernel
  double vmax;
  if ((int64_t)(blockIdx.x + 1) * (kMoveThreads * PER) <= (int64_t)N)
    vmax = move_particles<Model, true, XT>(mdl, y, N, ld, key, stream, t, i0, anc, xprev, xnew, logw);
  else
    vmax = move_particles<Model, false, XT>(mdl, y, N, ld, key, stream, t, i0, anc, xprev, xnew, logw);
  // exact max(logw) for the next normalize(): ordered encoding, REDUX per warp, one atomic per CTA
  const unsigned long long wm = warp_max_ordered(encode_ordered(vmax));
  if ((tid & 31) == 0) s_max[tid >> 5] = wm;
  __syncthreads();
  if (tid == 0) {
    unsigned long long m = s_max[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) m = s_max[w] > m ? s_max[w] : m;
    atomicMax(&ctrl->maxslot[t & 1u], m);
  }
}

// Guided move (particle_filter!, particles.jl:66-80; docs/SPEC.md §10) of the one-dimensional models: x_i ~ N(c0 + c1 x[a_i], c2²)
// from the SAME Philox normal, logw_i = logpdf(observation(x_i), y) + logpdf(transition(x[a_i]), x_i) − logpdf(proposal(x[a_i]), x_i),
// always stored (it is not a function of x_i alone).  Same streaming layout as move_kernel; in the binary32-state tier the
// drawn state is rounded before any density sees it (SPEC §9).
template <class Model, bool FULL, class XT>
__device__ __forceinline__ double guided_move_particles(const Model& mdl, const TransDensity<Model>& f, const ProposalCoef& pc, double y, int N,
                                                        const RngKey& key, uint32_t stream, uint32_t t, int i0, const int32_t* __restrict__ anc,
                                                        const XT* __restrict__ xprev, XT* __restrict__ xnew, double* __restrict__ logw) {
  static_assert(Model::D == 1, "guided proposals are defined for the one-dimensional models");
  constexpr int PER = 2 * kMovePairs;
  int a[PER];
  if (FULL) {
#pragma unroll
    for (int q = 0; q < PER / 2; ++q) {
      const int2 v = __ldcs(reinterpret_cast<const int2*>(anc + i0) + q);
      a[2 * q] = v.x; a[2 * q + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < PER; ++k) a[k] = (i0 + k < N) ? anc[i0 + k] : 0;
  }
  double xp[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) xp[k] = (double)__ldg(&xprev[a[k]]);
  double vmax = -INFINITY;
#pragma unroll
  for (int r = 0; r < kMovePairs; ++r) {
    const int i = i0 + 2 * r;
    if (!FULL && i >= N) break;
    double za, zb;
    normal_pair_at(key, (uint32_t)(i >> 1), stream, t, PURPOSE_TRANSITION, 0u, za, zb);
    const double mqa = fma(pc.c[1], xp[2 * r], pc.c[0]), mqb = fma(pc.c[1], xp[2 * r + 1], pc.c[0]);
    double xa[1] = {fma(pc.c[2], za, mqa)}, xb[1] = {fma(pc.c[2], zb, mqb)};
    round_state<XT>(xa);
    round_state<XT>(xb);
    const double la = mdl.logweight(xa, y) + guided_correction(mdl, f, pc.c, mqa, xp[2 * r], xa[0]);
    const double lb = mdl.logweight(xb, y) + guided_correction(mdl, f, pc.c, mqb, xp[2 * r + 1], xb[0]);
    if (FULL || i + 1 < N) {
      *reinterpret_cast<typename XVec2<XT>::type*>(xnew + i) = XVec2<XT>::make(xa[0], xb[0]);
      *reinterpret_cast<double2*>(logw + i) = make_double2(la, lb);
      if (la > vmax) vmax = la;
      if (lb > vmax) vmax = lb;
    } else {
      xnew[i] = (XT)xa[0];
      logw[i] = la;
      if (la > vmax) vmax = la;
    }
  }
  return vmax;
}

template <class Model, class XT>
__global__ void __launch_bounds__(kMoveThreads, 4)
    guided_move_kernel(Derived dv, ProposalCoef pc, double y, int N, RngKey key, uint32_t stream, uint32_t t, const int32_t* __restrict__ anc,
                       const XT* __restrict__ xprev, XT* __restrict__ xnew, double* __restrict__ logw, FilterCtrl* ctrl) {
  constexpr int PER = 2 * kMovePairs;
  constexpr int NW = kMoveThreads / 32;
  __shared__ unsigned long long s_max[NW];
  const int tid = threadIdx.x;
  Model mdl;
  mdl.load(dv.d);
  TransDensity<Model> f;
  f.load(dv.d);
  const int i0 = (blockIdx.x * kMoveThreads + tid) * PER;
  pdl_launch_dependents();
  pdl_wait();  // the ancestors come from anc_hist_kernel
  double vmax;
  if ((int64_t)(blockIdx.x + 1) * (kMoveThreads * PER) <= (int64_t)N)
    vmax = guided_move_particles<Model, true, XT>(mdl, f, pc, y, N, key, stream, t, i0, anc, xprev, xnew, logw);
  else
    vmax = guided_move_particles<Model, false, XT>(mdl, f, pc, y, N, key, stream, t, i0, anc, xprev, xnew, logw);
  const unsigned long long wm = warp_max_ordered(encode_ordered(vmax));
  if ((tid & 31) == 0) s_max[tid >> 5] = wm;
  __syncthreads();
  if (tid == 0) {
    unsigned long long m = s_max[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) m = s_max[w] > m ? s_max[w] : m;
    atomicMax(&ctrl->maxslot[t & 1u], m);
  }
}

// The guided move of UCSV (docs/SPEC.md §10b; guided_move_ucsv in smcb_models.cuh): one pair of particles per thread, three
// component gathers through the sorted ancestor vector, binary64 storage, log-weights stored.
__global__ void __launch_bounds__(kMoveThreads, 3)
    guided_move_ucsv_kernel(Derived dv, double kappa, double y, int N, int64_t ld, RngKey key, uint32_t stream, uint32_t t, const int32_t* __restrict__ anc,
                            const double* __restrict__ xprev, double* __restrict__ xnew, double* __restrict__ logw, FilterCtrl* ctrl) {
  constexpr int NW = kMoveThreads / 32;
  __shared__ unsigned long long s_max[NW];
  const int tid = threadIdx.x;
  ModelUCSV mdl;
  mdl.load(dv.d);
  const int p = blockIdx.x * kMoveThreads + tid;
  const int i = 2 * p;
  pdl_launch_dependents();
  pdl_wait();  // the ancestors come from anc_hist_kernel
  double vmax = -INFINITY;
  if (i < N) {
    const bool two = i + 1 < N;
    const int a0 = anc[i], a1 = two ? anc[i + 1] : a0;
    double xpa[3], xpb[3], za[3], zb[3], xa[3], xb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      xpa[k] = __ldg(&xprev[k * ld + a0]);
      xpb[k] = __ldg(&xprev[k * ld + a1]);
      normal_pair_at(key, (uint32_t)p, stream, t, PURPOSE_TRANSITION, (uint32_t)k, za[k], zb[k]);
    }
    const double la = guided_move_ucsv(mdl, kappa, za, xpa, y, xa);
#pragma unroll
    for (int k = 0; k < 3; ++k) xnew[k * ld + i] = xa[k];
    logw[i] = la;
    vmax = la;
    if (two) {
      const double lb = guided_move_ucsv(mdl, kappa, zb, xpb, y, xb);
#pragma unroll
      for (int k = 0; k < 3; ++k) xnew[k * ld + i + 1] = xb[k];
      logw[i + 1] = lb;
      if (lb > vmax) vmax = lb;
    }
  }
  const unsigned long long wm = warp_max_ordered(encode_ordered(vmax));
  if ((tid & 31) == 0) s_max[tid >> 5] = wm;
  __syncthreads();
  if (tid == 0) {
    unsigned long long m = s_max[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) m = s_max[w] > m ? s_max[w] : m;
    atomicMax(&ctrl->maxslot[t & 1u], m);
  }
}

// ================================================================================================
// The binary32-ARITHMETIC tier (docs/SPEC.md §9b): four particles per thread = one Philox block per state component, float
// Box-Muller / model arithmetic / log-weights, 16-byte float4 loads and stores.  Everything between the weights and the
// ancestors (sum_kernel<float, float> -> bounds / anc_hist or the two-level multinomial kernels) is shared with the other tiers.
template <class ModelF>
__global__ void __launch_bounds__(256) init_kernel_f(DerivedF dv, float y0, int64_t N, int64_t ld, RngKey key, uint32_t stream,
                                                      float* __restrict__ x, float* __restrict__ logw, FilterCtrl* ctrl) {
  constexpr int D = ModelF::D;
  __shared__ double sh[32];
  ModelF mdl;
  mdl.load(dv.d);
  const int64_t nquads = (N + 3) >> 2;
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double vmax = -INFINITY;
  if (q < nquads) {
    float z[D][4];
#pragma unroll
    for (int k = 0; k < D; ++k) normal_quadf_at(key, (uint32_t)q, stream, 0u, PURPOSE_INIT, (uint32_t)k, z[k]);
    float xo[D][4], lw[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float zi[D], xi[D];
#pragma unroll
      for (int k = 0; k < D; ++k) zi[k] = z[k][h];
      mdl.init(zi, xi);
      lw[h] = mdl.logweight(xi, y0);
#pragma unroll
      for (int k = 0; k < D; ++k) xo[k][h] = xi[k];
    }
    const int64_t i = 4 * q;
    if (i + 3 < N) {
#pragma unroll
      for (int k = 0; k < D; ++k) *reinterpret_cast<float4*>(x + k * ld + i) = make_float4(xo[k][0], xo[k][1], xo[k][2], xo[k][3]);
      *reinterpret_cast<float4*>(logw + i) = make_float4(lw[0], lw[1], lw[2], lw[3]);
#pragma unroll
      for (int h = 0; h < 4; ++h)
        if ((double)lw[h] > vmax) vmax = (double)lw[h];
    } else {
      for (int h = 0; h < 4 && i + h < N; ++h) {
#pragma unroll
        for (int k = 0; k < D; ++k) x[k * ld + i + h] = xo[k][h];
        logw[i + h] = lw[h];
        if ((double)lw[h] > vmax) vmax = (double)lw[h];
      }
    }
  }
  const double bm = block_max(vmax, sh);
  if (threadIdx.x == 0) atomicMax(&ctrl->maxslot[0], encode_ordered(bm));
}

template <class ModelF>
__global__ void __launch_bounds__(kMoveThreads, (ModelF::D == 1) ? 6 : 3)
    move_kernel_f(DerivedF dv, float y, int N, int64_t ld, RngKey key, uint32_t stream, uint32_t t, const int32_t* __restrict__ anc,
                  const float* __restrict__ xprev, float* __restrict__ xnew, float* __restrict__ logw /* null: not stored */, FilterCtrl* ctrl) {
  constexpr int D = ModelF::D;
  constexpr int NW = kMoveThreads / 32;
  __shared__ unsigned long long s_max[NW];
  const int tid = threadIdx.x;
  ModelF mdl;
  mdl.load(dv.d);
  const int q = blockIdx.x * kMoveThreads + tid;
  const int i0 = 4 * q;
  pdl_launch_dependents();
  pdl_wait();  // the ancestors come from anc_hist_kernel / mn_cell_kernel
  double vmax = -INFINITY;
  if (i0 < N) {
    const bool full = i0 + 3 < N;
    int a[4];
    if (full) {
      const int4 v = __ldcs(reinterpret_cast<const int4*>(anc + i0));
      a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
    } else {
#pragma unroll
      for (int h = 0; h < 4; ++h) a[h] = (i0 + h < N) ? anc[i0 + h] : 0;
    }
    float xp[D][4];
#pragma unroll
    for (int k = 0; k < D; ++k)
#pragma unroll
      for (int h = 0; h < 4; ++h) xp[k][h] = __ldg(&xprev[k * ld + a[h]]);
    float z[D][4];
#pragma unroll
    for (int k = 0; k < D; ++k) normal_quadf_at(key, (uint32_t)q, stream, t, PURPOSE_TRANSITION, (uint32_t)k, z[k]);
    float xo[D][4], lw[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float zi[D], pi[D], xi[D];
#pragma unroll
      for (int k = 0; k < D; ++k) { zi[k] = z[k][h]; pi[k] = xp[k][h]; }
      mdl.transition(zi, pi, xi);
      lw[h] = mdl.logweight(xi, y);
#pragma unroll
      for (int k = 0; k < D; ++k) xo[k][h] = xi[k];
    }
    if (full) {
#pragma unroll
      for (int k = 0; k < D; ++k) *reinterpret_cast<float4*>(xnew + k * ld + i0) = make_float4(xo[k][0], xo[k][1], xo[k][2], xo[k][3]);
      if (logw) *reinterpret_cast<float4*>(logw + i0) = make_float4(lw[0], lw[1], lw[2], lw[3]);
#pragma unroll
      for (int h = 0; h < 4; ++h)
        if ((double)lw[h] > vmax) vmax = (double)lw[h];  // `>` ignores NaN like the CPU loop
    } else {
      for (int h = 0; h < 4 && i0 + h < N; ++h) {
#pragma unroll
        for (int k = 0; k < D; ++k) xnew[k * ld + i0 + h] = xo[k][h];
        if (logw) logw[i0 + h] = lw[h];
        if ((double)lw[h] > vmax) vmax = (double)lw[h];
      }
    }
  }
  const unsigned long long wm = warp_max_ordered(encode_ordered(vmax));
  if ((tid & 31) == 0) s_max[tid >> 5] = wm;
  __syncthreads();
  if (tid == 0) {
    unsigned long long m = s_max[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) m = s_max[w] > m ? s_max[w] : m;
    atomicMax(&ctrl->maxslot[t & 1u], m);
  }
}

// LG1D log-weights of the tier, materialised when a caller fetches them; and w_i = expf(logw_i - max) / Σe
__global__ void logw_kernel_f(DerivedF dv, float y, int64_t N, const float* __restrict__ x, float* __restrict__ logw) {
  ModelLG1Df mdl;
  mdl.load(dv.d);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    const float xi = x[i];
    logw[i] = mdl.logweight(&xi, y);
  }
}
__global__ void weights_kernel_f(const float* __restrict__ logw, double* __restrict__ w, int64_t N, StepStats st) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) w[i] = (double)det_expf(logw[i] - (float)st.mx) / st.sum;
}

// logw_i = logpdf(observation(x_i), y): materialises the log-weights that the LG1D step does not store
template <class XT>
__global__ void logw_kernel(Derived dv, double y, int64_t N, const XT* __restrict__ x, double* __restrict__ logw) {
  ModelLG1D mdl;
  mdl.load(dv.d);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    const double xi = (double)x[i];
    logw[i] = mdl.logweight(&xi, y);
  }
}
// binary32 states -> binary64 for the host-facing fetch
__global__ void widen_kernel(const float* __restrict__ src, double* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)src[i];
}

// w_i = exp(logw_i - max) / Σe   (normalize, particles.jl:11) — only when the caller fetches w
__global__ void weights_kernel(const double* __restrict__ logw, double* __restrict__ w, int64_t N, StepStats st) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) w[i] = det_exp(logw[i] - st.mx) / st.sum;
}

// ------------------------------------------------------------------------------------------------
// On-device weighted summaries of the current cloud (docs/SPEC.md §8): the per-step
// `quantile(x, weights(w), p)` / `var(x, weights(w))` of README.md:41,51 and
// examples/inflation_example.jl:39-55 without reading 16 B per particle back to the host.
// Weights are the fixed-point q_i of SPEC §5 (recovered from the tile-local CDF that sum_kernel left
// behind), so every count is an exact integer and the quantiles are bit-reproducible.
constexpr int kSumThreadsW = 256;
constexpr int kMaxProbs = 16;

__device__ __forceinline__ unsigned long long q_of(const unsigned long long* __restrict__ cl, int64_t i, int tile_items, bool weighted) {
  if (!weighted) return 1ull;
  const unsigned long long c = cl[i];
  return (i % tile_items == 0) ? c : c - cl[i - 1];
}

// mode 0: part[b] = Σ q x ; mode 1: part[b] = Σ q (x - mean)^2   (per-block partials, fixed order)
template <class XT>
__global__ void __launch_bounds__(kSumThreadsW)
    wmoment_kernel(const XT* __restrict__ x, const unsigned long long* __restrict__ cl, int64_t N, int tile_items, int weighted,
                   int mode, double mean, double* __restrict__ part) {
  __shared__ double sh[kSumThreadsW / 32];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kSumThreadsW + threadIdx.x; i < N; i += (int64_t)gridDim.x * kSumThreadsW) {
    const double qd = (double)q_of(cl, i, tile_items, weighted != 0);
    const double v = (double)x[i];
    if (qd != 0.0) acc += mode == 0 ? qd * v : qd * ((v - mean) * (v - mean));  // q = 0: an infinite state must not poison the sum
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kSumThreadsW / 32; ++w) t += sh[w];
    part[blockIdx.x] = t;
  }
}

// one radix-select pass (8 bits, most significant first) for np probabilities at once:
// hist[j][digit] += q_i for the particles whose key agrees with prefix[j] in the digits already fixed
template <class XT>
__global__ void __launch_bounds__(kSumThreadsW)
    rsel_hist_kernel(const XT* __restrict__ x, const unsigned long long* __restrict__ cl, int64_t N, int tile_items, int weighted,
                     int pass, int np, const unsigned long long* __restrict__ prefix, unsigned long long* __restrict__ hist) {
  __shared__ unsigned long long s_hist[kMaxProbs * 256];
  __shared__ unsigned long long s_prefix[kMaxProbs];
  for (int k = threadIdx.x; k < np * 256; k += kSumThreadsW) s_hist[k] = 0ull;
  if (threadIdx.x < np) s_prefix[threadIdx.x] = prefix[threadIdx.x];
  __syncthreads();
  const int shift = 56 - 8 * pass;
  for (int64_t i = (int64_t)blockIdx.x * kSumThreadsW + threadIdx.x; i < N; i += (int64_t)gridDim.x * kSumThreadsW) {
    const unsigned long long q = q_of(cl, i, tile_items, weighted != 0);
    if (q == 0ull) continue;
    const unsigned long long key = encode_ordered((double)x[i]);
    const unsigned digit = (unsigned)(key >> shift) & 255u;
    for (int j = 0; j < np; ++j) {
      const bool match = (pass == 0) || ((key ^ s_prefix[j]) >> (shift + 8)) == 0ull;
      if (match) atomicAdd(&s_hist[j * 256 + digit], q);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < np * 256; k += kSumThreadsW)
    if (s_hist[k]) atomicAdd(&hist[k], s_hist[k]);
}

// picks the digit holding rank[j] (0-based mass offset inside the current prefix), fixes it, clears the histogram
__global__ void rsel_pick_kernel(int pass, int np, unsigned long long* __restrict__ prefix, unsigned long long* __restrict__ rank,
                                 unsigned long long* __restrict__ hist) {
  const int j = threadIdx.x;
  if (j < np) {
    unsigned long long r = rank[j], cum = 0;
    int d = 255;
    for (int b = 0; b < 256; ++b) {
      const unsigned long long h = hist[j * 256 + b];
      if (cum + h > r) { d = b; break; }
      cum += h;
    }
    rank[j] = r - cum;
    prefix[j] |= (unsigned long long)d << (56 - 8 * pass);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < np * 256; k += blockDim.x) hist[k] = 0ull;
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

__global__ void ancestor_kernel(const uint64_t* __restrict__ cdf, int64_t N, int64_t n_out, int resampler, uint64_t Rw, RngKey key,
                                uint32_t stream, uint32_t t, uint32_t purpose, const FilterCtrl* ctrl,
                                int64_t* __restrict__ anc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const uint64_t Q = ctrl->total;
  if (Q == 0) {
    anc[i] = i < N ? i : N - 1;
    return;
  }
  uint64_t u;
  if (resampler == RESAMPLE_SYSTEMATIC) u = mulhi64(uniform64_at(key, 0u, stream, t, purpose), Rw);
  else u = uniform64_at(key, (uint32_t)i, stream, t, purpose);
  const uint64_t tau = threshold_of(resampler, (uint64_t)i, Rw, u, Q);
  anc[i] = lower_count(cdf, (int64_t)0, N - 1, tau);
}

// state storage type of the live filter (SPEC §9)
template <class F>
static void dispatch_xt(int prec, F&& f) {
  if (prec) f(float{});
  else f(double{});
}
template <class F>
void dispatch_model(int kind, F&& f) {
  switch (kind) {
    case KIND_LG1D: f(ModelLG1D{}); break;
    case KIND_SV: f(ModelSV{}); break;
    case KIND_UCSV: f(ModelUCSV{}); break;
    case KIND_MVLG2: f(ModelMVLG<2>{}); break;
    case KIND_MVLG3: f(ModelMVLG<3>{}); break;
    case KIND_MVLG4: f(ModelMVLG<4>{}); break;
    default: throw Error{SMCB_ERR_BAD_ARG, "unknown model kind"};
  }
}

}  // namespace

// the derived block a model's kernels take by value: 8 doubles for the univariate kinds, 64 for the multivariate ones
static inline const Derived& pick_dv(const Derived& a, const DerivedMV&, const Derived*) { return a; }
static inline const DerivedMV& pick_dv(const Derived&, const DerivedMV& b, const DerivedMV*) { return b; }
#define SMCB_DV(M) pick_dv(dv_, dvmv_, (const typename M::DV*)nullptr)

static DerivedF to_float(const Derived& d) {
  DerivedF f;
  for (int i = 0; i < kParamStride; ++i) f.d[i] = (float)d.d[i];
  return f;
}
template <class F>
static void dispatch_model_f(int kind, F&& f) {
  switch (kind) {
    case KIND_LG1D: f(ModelLG1Df{}); break;
    case KIND_SV: f(ModelSVf{}); break;
    case KIND_UCSV: f(ModelUCSVf{}); break;
    default: throw Error{SMCB_ERR_BAD_ARG, "unknown model kind"};
  }
}

// ------------------------------------------------------------------------------------------------
SingleFilter::~SingleFilter() {
  release();
  for (auto e : ev_pool_) cudaEventDestroy(e);
  for (auto e : ev_call_)
    if (e) cudaEventDestroy(e);
}

void SingleFilter::release() {
  cudaFree(x_[0]); cudaFree(x_[1]); cudaFree(logw_[0]); cudaFree(logw_[1]); cudaFree(w_tmp_); cudaFree(cdf_); cudaFree(anc_);
  cudaFree(ctrl_); cudaFree(desc_); cudaFree(psum_); cudaFree(psum2_); cudaFree(stats_dev_);
  cudaFree(tile_arrays_); cudaFree(bound_arrays_); cudaFree(summary_dev_); cudaFree(mn_arrays_); cudaFree(util_anc_);
  util_anc_ = nullptr; util_cap_ = 0;
  tile_arrays_ = nullptr; bound_arrays_ = nullptr; bound_cap_ = 0; summary_dev_ = nullptr; summary_cap_ = 0; mn_arrays_ = nullptr; mn_cap_ = 0;
  x_[0] = x_[1] = logw_[0] = logw_[1] = w_tmp_ = psum_ = psum2_ = nullptr;
  cdf_ = nullptr; anc_ = nullptr; ctrl_ = nullptr; desc_ = nullptr; stats_dev_ = nullptr;
  cap_N_ = cap_d_ = cap_stats_ = cap_anc_rows_ = ntiles_cap_ = 0;
}

void SingleFilter::ensure_capacity(int kind, int64_t N, int64_t anc_rows) {
  const int d = state_dim(kind);
  const int64_t ld = (N + 31) & ~int64_t(31);
  if (ld > cap_N_ || d > cap_d_) {
    cudaFree(x_[0]); cudaFree(x_[1]); cudaFree(logw_[0]); cudaFree(logw_[1]); cudaFree(w_tmp_); cudaFree(cdf_);
    cudaFree(desc_); cudaFree(psum_); cudaFree(psum2_);
    x_[0] = x_[1] = logw_[0] = logw_[1] = w_tmp_ = psum_ = psum2_ = nullptr; cdf_ = nullptr; desc_ = nullptr;
    cap_N_ = 0;
    const int64_t cd = std::max<int64_t>(d, cap_d_);
    SMCB_CUDA_TRY(cudaMalloc(&x_[0], sizeof(double) * ld * cd));
    SMCB_CUDA_TRY(cudaMalloc(&x_[1], sizeof(double) * ld * cd));
    SMCB_CUDA_TRY(cudaMalloc(&logw_[0], sizeof(double) * ld));
    SMCB_CUDA_TRY(cudaMalloc(&logw_[1], sizeof(double) * ld));
    ntiles_cap_ = (ld + kScanTile - 1) / kScanTile;
    SMCB_CUDA_TRY(cudaMalloc(&desc_, sizeof(unsigned long long) * 2 * ntiles_cap_));
    const int64_t npart = std::max<int64_t>(ntiles_cap_, kMaxTiles);  // per-tile partial sums of either scan flavour
    SMCB_CUDA_TRY(cudaMalloc(&psum_, sizeof(double) * npart));
    SMCB_CUDA_TRY(cudaMalloc(&psum2_, sizeof(double) * npart));
    cap_N_ = ld;
    cap_d_ = cd;
    cudaFree(anc_); anc_ = nullptr; cap_anc_rows_ = 0;
  }
  if (!ctrl_) SMCB_CUDA_TRY(cudaMalloc(&ctrl_, sizeof(FilterCtrl)));
  if (!tile_arrays_) SMCB_CUDA_TRY(cudaMalloc(&tile_arrays_, sizeof(unsigned long long) * 5 * kMaxTiles));
  {
    const int64_t nb = cap_N_ / kSysParticles + 3;
    if (nb > bound_cap_) {
      cudaFree(bound_arrays_); bound_arrays_ = nullptr; bound_cap_ = 0;
      SMCB_CUDA_TRY(cudaMalloc(&bound_arrays_, sizeof(int32_t) * 2 * nb));
      bound_cap_ = nb;
    }
  }
  {
    const int64_t nc = (cap_N_ + kMnCell - 1) / kMnCell + 1;
    if (nc > mn_cap_) {  // cell index of the two-level multinomial resampler (allocated with the rest: 20 B per 4096 particles)
      cudaFree(mn_arrays_); mn_arrays_ = nullptr; mn_cap_ = 0;
      const size_t bytes = (size_t)(sizeof(unsigned long long) * nc + sizeof(int32_t) * (3 * nc + 8 + (1 << 13) + 1));
      SMCB_CUDA_TRY(cudaMalloc(&mn_arrays_, bytes));
      SMCB_CUDA_TRY(cudaMemsetAsync(mn_arrays_, 0, bytes, stream_));
      mn_cap_ = nc;
    }
  }
  if (num_sms_ == 0) {
    int v = 0;
    SMCB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device_));
    num_sms_ = v > 0 ? v : 148;
    if (const char* e = std::getenv("SMCB_ANC_FORCE_EXACT")) anc_eps_ = (e[0] && e[0] != '0') ? 2.0 : 1e-9;
  }
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
  launches_[TK_TOTAL] = launches_[TK_SCAN] + launches_[TK_PROP] + launches_[TK_INIT] + launches_[TK_STATS] + launches_[TK_BOUNDS] + launches_[TK_ANC];
}

void SingleFilter::timing(double ms[TK_COUNT], int64_t launches[TK_COUNT]) const {
  for (int k = 0; k < TK_COUNT; ++k) { ms[k] = ms_[k]; launches[k] = launches_[k]; }
}

void SingleFilter::ensure_logw() {
  if (logw_valid_ || !live()) return;
  if (prec_ == 2) {
    logw_kernel_f<<<(unsigned)((N_ + 255) / 256), 256, 0, stream_>>>(to_float(dv_w_), (float)y_cur_, N_, reinterpret_cast<const float*>(x_[cur_]),
                                                                   reinterpret_cast<float*>(logw_[cur_]));
    SMCB_CUDA_TRY(cudaGetLastError());
    logw_valid_ = true;
    return;
  }
  dispatch_xt(prec_, [&](auto tag) {
    using XT = decltype(tag);
    logw_kernel<XT><<<(unsigned)((N_ + 255) / 256), 256, 0, stream_>>>(dv_w_, y_cur_, N_, reinterpret_cast<const XT*>(x_[cur_]), logw_[cur_]);
  });
  SMCB_CUDA_TRY(cudaGetLastError());
  logw_valid_ = true;
}

void SingleFilter::launch_init(double y0) {
  logw_valid_ = true;
  sum_done_ = false;
  SMCB_CUDA_TRY(cudaMemsetAsync(desc_, 0, sizeof(unsigned long long) * 2 * ntiles_cap_, stream_));
  desc_clean_t_ = 0;
  reset_ctrl_kernel<<<1, 1, 0, stream_>>>(ctrl_);
  launches_[TK_INIT] += 1;
  const int64_t npairs = (N_ + 1) / 2;
  const unsigned grid = (unsigned)((npairs + 255) / 256);
  mark(TK_INIT, true);
  if (prec_ == 2 && is_mv_kind(kind_)) throw Error{SMCB_ERR_UNSUPPORTED, "the binary32-arithmetic tier is built for the univariate kinds and UCSV (docs/SPEC.md §9b)"};
  if (prec_ == 2) {
    const unsigned qgrid = (unsigned)(((N_ + 3) / 4 + 255) / 256);
    dispatch_model_f(kind_, [&](auto m) {
      using M = decltype(m);
      init_kernel_f<M><<<qgrid, 256, 0, stream_>>>(to_float(dv_), (float)y0, N_, ld_, key_, stream_id_, reinterpret_cast<float*>(x_[cur_]),
                                                  reinterpret_cast<float*>(logw_[cur_]), ctrl_);
    });
  } else
  dispatch_model(kind_, [&](auto m) {
    using M = decltype(m);
    dispatch_xt(prec_, [&](auto tag) {
      using XT = decltype(tag);
      init_kernel<M, XT><<<grid, 256, 0, stream_>>>(SMCB_DV(M), y0, N_, ld_, key_, stream_id_, reinterpret_cast<XT*>(x_[cur_]), logw_[cur_], ctrl_);
    });
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
  if ((write_cdf || from_w_) && !cdf_) SMCB_CUDA_TRY(cudaMalloc(&cdf_, sizeof(uint64_t) * cap_N_));  // multinomial / utilities only
  if (write_cdf || from_w_) {
    // The look-back descriptors of this scan must start cleared.  A WRITE_CDF scan at time t clears the half the scan at
    // t + 1 will use — but sorted-resampler steps advance t_ without scanning, so a multinomial step after them would find
    // the descriptors of an older scan of the same parity, still flagged inclusive, and could accept a stale prefix.
    if (desc_clean_t_ != (int64_t)t_) SMCB_CUDA_TRY(cudaMemsetAsync(dc, 0, sizeof(unsigned long long) * ntiles, stream_));
    desc_clean_t_ = (int64_t)t_ + 1;  // this launch clears dn
  }
  ensure_logw();
  const double* lw = logw_[cur_];
  mark(klass, true);
  StepStats* so = stats_dev_ + stat_index;
  if (from_w_)
    scan_kernel<true, true><<<ntiles, kScanThreads, 0, stream_>>>(lw, cdf_, N_, S_, ctrl_, dc, dn, psum_, psum2_, so, slot, ntiles);
  else if (write_cdf)
    scan_kernel<true, false><<<ntiles, kScanThreads, 0, stream_>>>(lw, cdf_, N_, S_, ctrl_, dc, dn, psum_, psum2_, so, slot, ntiles);
  else
    scan_kernel<false, false><<<ntiles, kScanThreads, 0, stream_>>>(lw, cdf_, N_, S_, ctrl_, dc, dn, psum_, psum2_, so, slot, ntiles);
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
    prop_kernel<M><<<grid, kPropThreads, 0, stream_>>>(SMCB_DV(M), y, N_, ld_, resampler, R_, key_, stream_id_, t, cdf_,
                                                       x_[cur_], x_[cur_ ^ 1], logw_[cur_ ^ 1], anc, ctrl_);
  });
  mark(TK_PROP, false);
  SMCB_CUDA_TRY(cudaGetLastError());
  cur_ ^= 1;
  t_ = t;
  logw_valid_ = true;
  y_cur_ = y;
  sum_done_ = false;
}

// tile geometry of the sorted-resampler step: one warp per tile, one full wave of resident warps when N
// allows, at most kMaxTiles tiles; returns the tile-local CDF buffer
unsigned long long* SingleFilter::step_index(StepIndex& ix) {
  const int64_t nchunks = (N_ + kChunk - 1) / kChunk;
  const int64_t resident = (int64_t)num_sms_ * kSumCtasPerSm * kSumWarps;
  int64_t cpt = std::max<int64_t>((nchunks + resident - 1) / resident, (nchunks + kMaxTiles - 1) / kMaxTiles);
  cpt = std::max<int64_t>(cpt, 1);
  ix.chunks_per_tile = (int)cpt;
  ix.tile_items = (int)(cpt * kChunk);
  ix.ntiles = (int)((nchunks + cpt - 1) / cpt);
  ix.tile_tot = tile_arrays_;
  ix.tile_excl = tile_arrays_ + kMaxTiles;
  ix.tile_incl = tile_arrays_ + 2 * kMaxTiles;
  ix.tile_lexcl = tile_arrays_ + 3 * kMaxTiles;
  ix.cta_tot = tile_arrays_ + 4 * kMaxTiles;
  ix.bound_pos = bound_arrays_;
  ix.bound_tile = bound_arrays_ + bound_cap_;
  if (!cdf_) SMCB_CUDA_TRY(cudaMalloc(&cdf_, sizeof(uint64_t) * cap_N_));  // holds the tile-local CDF here (the global CDF on the multinomial path)
  return reinterpret_cast<unsigned long long*>(cdf_);
}

// normalize() ingredients of the CURRENT weights (-> stats_dev_[stat_index]) and everything the next
// sorted-resampler step needs (tile-local CDF, tile index, Q, systematic offset).  A stepping caller
// runs it right after a step to read logμ / ess; the next step then starts at bounds_kernel.
void SingleFilter::launch_sum(int64_t stat_index) {
  StepIndex ix;
  unsigned long long* cl = step_index(ix);
  const uint32_t t = t_ + 1;
  mark(TK_SCAN, true);
  const int from_x = logw_valid_ ? 0 : 1;  // the previous LG1D step kept its log-weights implicit in x
  if (prec_ == 2)  // binary32 arithmetic: the states AND the stored log-weights are binary32
    SMCB_CUDA_TRY(launch_pdl(sum_kernel<float, float>, dim3((ix.ntiles + kSumWarps - 1) / kSumWarps), dim3(kSumThreads), stream_,
                             from_x ? reinterpret_cast<const float*>(x_[cur_]) : reinterpret_cast<const float*>(logw_[cur_]), cl, N_, S_, ctrl_, ix, psum_,
                             psum2_, stats_dev_ + stat_index, (int)(t_ & 1u), (int)RESAMPLE_SYSTEMATIC, R_, key_, stream_id_, t, from_x, dv_w_, y_cur_));
  else if (from_x && prec_)
    SMCB_CUDA_TRY(launch_pdl(sum_kernel<float>, dim3((ix.ntiles + kSumWarps - 1) / kSumWarps), dim3(kSumThreads), stream_,
                             reinterpret_cast<const float*>(x_[cur_]), cl, N_, S_, ctrl_, ix, psum_, psum2_, stats_dev_ + stat_index, (int)(t_ & 1u),
                             (int)RESAMPLE_SYSTEMATIC, R_, key_, stream_id_, t, from_x, dv_w_, y_cur_));
  else
    SMCB_CUDA_TRY(launch_pdl(sum_kernel<double>, dim3((ix.ntiles + kSumWarps - 1) / kSumWarps), dim3(kSumThreads), stream_,
                             from_x ? (const double*)x_[cur_] : (const double*)logw_[cur_], cl, N_, S_, ctrl_, ix, psum_, psum2_,
                             stats_dev_ + stat_index, (int)(t_ & 1u), (int)RESAMPLE_SYSTEMATIC, R_, key_, stream_id_, t, from_x, dv_w_, y_cur_));
  mark(TK_SCAN, false);
  SMCB_CUDA_TRY(cudaGetLastError());
  sum_done_ = true;
}

// one bootstrap_filter! step: stats of the current weights (-> stats_dev_[stat_index]) and the move to t_+1
void SingleFilter::launch_step(int64_t stat_index, double y, int resampler, const double* proposal) {
  ProposalCoef pc{};
  if (proposal) {  // guided step (docs/SPEC.md §10): sorted resamplers, one-dimensional models
    if (prec_ == 2) throw Error{SMCB_ERR_UNSUPPORTED, "guided proposals are not built for the binary32-arithmetic tier (docs/SPEC.md §9b)"};
    if (kind_ > KIND_UCSV) throw Error{SMCB_ERR_UNSUPPORTED, "guided proposals are defined for LG1D, SV (docs/SPEC.md §10) and UCSV (§10b)"};
    if (kind_ == KIND_UCSV && prec_ != 0) throw Error{SMCB_ERR_UNSUPPORTED, "the guided UCSV move keeps binary64 states (docs/SPEC.md §10b)"};
    if (kind_ == KIND_UCSV && !(proposal[0] >= 0.0 && proposal[0] <= 1.0))
      throw Error{SMCB_ERR_BAD_ARG, "proposal (UCSV, docs/SPEC.md §10b): the triple is (kappa, 0, 1) with kappa in [0, 1]"};
    if (legacy_multinomial(resampler))
      throw Error{SMCB_ERR_UNSUPPORTED, "guided single filter with N <= 8192: stratified or systematic resampling (multinomial guided filters of that size run on the batched engine)"};
    if (!(proposal[2] > 0.0) || !std::isfinite(proposal[2]) || !std::isfinite(proposal[0]) || !std::isfinite(proposal[1]))
      throw Error{SMCB_ERR_BAD_ARG, "proposal: coefficients must be finite and the standard deviation c2 > 0"};
    pc.c[0] = proposal[0]; pc.c[1] = proposal[1]; pc.c[2] = proposal[2]; pc.c[3] = det_log(proposal[2]); pc.c[4] = 1.0 / proposal[2];
  }
  if (legacy_multinomial(resampler)) {  // a small cloud: materialised CDF + per-particle search, unsorted ancestors (SPEC §5 row 0)
    launch_scan(stat_index, true);
    launch_prop(y, resampler);
    return;
  }
  StepIndex ix;
  unsigned long long* cl = step_index(ix);
  if (!sum_done_) launch_sum(stat_index);  // (a stepping caller already ran it to read the statistics of the current weights)
  const int block_particles = kSysParticles;
  const unsigned nblocks = (unsigned)((N_ + block_particles - 1) / block_particles);
  const uint32_t t = t_ + 1;
  const int nbounds = (int)nblocks + 1;
  int32_t* anc = anc_;  // row 0 doubles as the scratch ancestor vector when nothing is recorded
  if (record_anc_) {
    const int64_t row = std::min<int64_t>(anc_rows_, cap_anc_rows_ - 1);
    anc = anc_ + row * cap_N_;
    anc_rows_ = row + 1;
  }
  if (resampler == RESAMPLE_MULTINOMIAL) {  // large cloud: two-level multinomial draw (SPEC §5c), three launches
    MnIndex mn;
    mn.ncells = (int)((N_ + kMnCell - 1) / kMnCell);
    mn.cellC = reinterpret_cast<unsigned long long*>(mn_arrays_);
    mn.K = reinterpret_cast<int32_t*>(mn.cellC + mn_cap_);
    mn.O = mn.K + mn_cap_;
    mn.part_start = mn.O + mn_cap_;
    mn.lut = mn.part_start + mn_cap_ + 1;
    mn.ticket = reinterpret_cast<unsigned int*>(mn.lut + kMnLutMax + 1);
    mark(TK_BOUNDS, true);
    SMCB_CUDA_TRY(launch_pdl(mn_prep_kernel, dim3((mn.ncells + 255) / 256), dim3(256), stream_, mn, ix, (const unsigned long long*)cl,
                             (const FilterCtrl*)ctrl_, (int)N_));
    const int64_t want = (N_ / 4 + kMnCountThreads * 4 - 1) / (kMnCountThreads * 4);  // >= 16 thresholds per thread
    const unsigned cgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms_ * 2));
    const size_t csmem = sizeof(int) * (size_t)(kMnLutMax + 1) + (mn.ncells <= kMnSmemCells ? (sizeof(int) + sizeof(unsigned long long)) * (size_t)mn.ncells : 0);
    if (!mn_attr_set_)
      SMCB_CUDA_TRY(cudaFuncSetAttribute(mn_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(int) * (kMnLutMax + 1) + (sizeof(int) + sizeof(unsigned long long)) * kMnSmemCells)));
    SMCB_CUDA_TRY(launch_pdl_smem(mn_count_kernel, dim3(cgrid), dim3(kMnCountThreads), csmem, stream_, mn, (const FilterCtrl*)ctrl_, (int)N_, key_,
                                  stream_id_, t));
    mark(TK_BOUNDS, false);
    SMCB_CUDA_TRY(cudaGetLastError());
    const unsigned items = (unsigned)(mn.ncells + (N_ + kMnChunk - 1) / kMnChunk + 1);  // >= Σ_c max(1, ceil(K_c / kMnChunk))
    mark(TK_ANC, true);
    constexpr size_t kCellSmem = sizeof(unsigned long long) * kMnCell + sizeof(int) * (kMnSub + 4 + kMnCell);  // 64 KB: CDF (later the head flags) + table + counters
    if (!mn_attr_set_) {
      SMCB_CUDA_TRY(cudaFuncSetAttribute(mn_cell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCellSmem));
      mn_attr_set_ = true;
    }
    SMCB_CUDA_TRY(launch_pdl_smem(mn_cell_kernel, dim3(items), dim3(kMnCellThreads), kCellSmem, stream_, mn, ix, (const unsigned long long*)cl,
                                  (const FilterCtrl*)ctrl_, (int)N_, key_, stream_id_, t, anc));
    mark(TK_ANC, false);
    SMCB_CUDA_TRY(cudaGetLastError());
  } else {
  mark(TK_BOUNDS, true);
  SMCB_CUDA_TRY(launch_pdl(bounds_kernel, dim3(((nbounds + 1) / 2 + 7) / 8), dim3(256), stream_, ix, cl, ctrl_, (int)N_, resampler, R_, key_,
                           stream_id_, t, nbounds, block_particles));
  mark(TK_BOUNDS, false);
  SMCB_CUDA_TRY(cudaGetLastError());
  mark(TK_ANC, true);
  if (resampler == RESAMPLE_SYSTEMATIC)
    SMCB_CUDA_TRY(launch_pdl(anc_hist_kernel<RESAMPLE_SYSTEMATIC>, dim3(nblocks), dim3(kSysThreads), stream_, (int)N_, R_, key_, stream_id_, t, ix, cl,
                             anc, ctrl_, anc_eps_));
  else
    SMCB_CUDA_TRY(launch_pdl(anc_hist_kernel<RESAMPLE_STRATIFIED>, dim3(nblocks), dim3(kSysThreads), stream_, (int)N_, R_, key_, stream_id_, t, ix, cl,
                             anc, ctrl_, anc_eps_));
  mark(TK_ANC, false);
  SMCB_CUDA_TRY(cudaGetLastError());
  }
  const unsigned mblocks = (unsigned)((N_ + kMoveThreads * 2 * kMovePairs - 1) / (kMoveThreads * 2 * kMovePairs));
  // LG1D: the new log-weights are two fma of the new states; they are not stored (sum_kernel and, when a caller asks
  // for them, logw_kernel recompute them bit for bit).  SV / UCSV weights cost an exp: stored as before.
  const bool implicit_logw = (kind_ == KIND_LG1D) && !proposal;
  mark(TK_PROP, true);
  if (proposal && kind_ == KIND_UCSV) {
    const unsigned ublocks = (unsigned)(((N_ + 1) / 2 + kMoveThreads - 1) / kMoveThreads);
    SMCB_CUDA_TRY(launch_pdl(guided_move_ucsv_kernel, dim3(ublocks), dim3(kMoveThreads), stream_, dv_, pc.c[0], y, (int)N_, ld_, key_, stream_id_, t,
                             (const int32_t*)anc, (const double*)x_[cur_], (double*)x_[cur_ ^ 1], logw_[cur_ ^ 1], ctrl_));
  } else if (proposal) {
    dispatch_xt(prec_, [&](auto tag) {
      using XT = decltype(tag);
      if (kind_ == KIND_LG1D)
        SMCB_CUDA_TRY(launch_pdl(guided_move_kernel<ModelLG1D, XT>, dim3(mblocks), dim3(kMoveThreads), stream_, dv_, pc, y, (int)N_, key_, stream_id_, t,
                                 anc, reinterpret_cast<const XT*>(x_[cur_]), reinterpret_cast<XT*>(x_[cur_ ^ 1]), logw_[cur_ ^ 1], ctrl_));
      else
        SMCB_CUDA_TRY(launch_pdl(guided_move_kernel<ModelSV, XT>, dim3(mblocks), dim3(kMoveThreads), stream_, dv_, pc, y, (int)N_, key_, stream_id_, t,
                                 anc, reinterpret_cast<const XT*>(x_[cur_]), reinterpret_cast<XT*>(x_[cur_ ^ 1]), logw_[cur_ ^ 1], ctrl_));
    });
  } else if (prec_ == 2) {
    const unsigned qblocks = (unsigned)(((N_ + 3) / 4 + kMoveThreads - 1) / kMoveThreads);
    dispatch_model_f(kind_, [&](auto m) {
      using M = decltype(m);
      SMCB_CUDA_TRY(launch_pdl(move_kernel_f<M>, dim3(qblocks), dim3(kMoveThreads), stream_, to_float(dv_), (float)y, (int)N_, ld_, key_, stream_id_, t,
                               (const int32_t*)anc, reinterpret_cast<const float*>(x_[cur_]), reinterpret_cast<float*>(x_[cur_ ^ 1]),
                               implicit_logw ? (float*)nullptr : reinterpret_cast<float*>(logw_[cur_ ^ 1]), ctrl_));
    });
  } else
  dispatch_model(kind_, [&](auto m) {
    using M = decltype(m);
    dispatch_xt(prec_, [&](auto tag) {
      using XT = decltype(tag);
      SMCB_CUDA_TRY(launch_pdl(move_kernel<M, XT>, dim3(mblocks), dim3(kMoveThreads), stream_, SMCB_DV(M), y, (int)N_, ld_, key_, stream_id_, t, anc,
                               reinterpret_cast<const XT*>(x_[cur_]), reinterpret_cast<XT*>(x_[cur_ ^ 1]),
                               implicit_logw ? (double*)nullptr : logw_[cur_ ^ 1], ctrl_));
    });
  });
  mark(TK_PROP, false);
  SMCB_CUDA_TRY(cudaGetLastError());
  cur_ ^= 1;
  t_ = t;
  logw_valid_ = !implicit_logw;
  y_cur_ = y;
  dv_w_ = dv_;
  sum_done_ = false;  // the parameters these weights were computed with (step() may be handed new ones)
}

bool SingleFilter::legacy_multinomial(int resampler) const { return resampler == RESAMPLE_MULTINOMIAL && N_ <= kMnLegacyMax; }

void SingleFilter::set_params(int kind, const double* params) {
  if (is_mv_kind(kind)) {
    const int d = state_dim(kind);
    const double R = params[2 * d * d + d];
    if (!(R > 0.0) || !std::isfinite(R)) throw Error{SMCB_ERR_BAD_ARG, "multivariate linear model: the observation variance R must be positive"};
    derive_params_mv(d, params, dvmv_.d);
  } else {
    derive_params(kind, params, dv_.d);
  }
}

static void check_args(int kind, int64_t N, int resampler) {
  if (kind < 0 || kind >= KIND_ALL) throw Error{SMCB_ERR_BAD_ARG, "unknown model kind"};
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
  kind_ = kind; d_ = state_dim(kind); N_ = N; ld_ = cap_N_; prec_ = next_prec_;
  S_ = quant_shift((uint64_t)N); R_ = strata_width((uint64_t)N);
  key_ = key; stream_id_ = stream_id; t_ = 0; cur_ = 0; anc_rows_ = 0;
  set_params(kind, params);
  begin_call();
  launch_init(y0);
  launch_sum(0);
  SMCB_CUDA_TRY(cudaMemcpyAsync(&last_, stats_dev_, sizeof(StepStats), cudaMemcpyDeviceToHost, stream_));
  end_call();
  if (st) *st = last_;
}

void SingleFilter::step(const double* params, double y, int resampler, StepStats* st, const double* proposal) {
  if (!live()) throw Error{SMCB_ERR_STATE, "bootstrap_step before bootstrap_init / log_likelihood"};
  check_args(kind_, N_, resampler);
  if (prec_ && legacy_multinomial(resampler)) throw Error{SMCB_ERR_BAD_ARG, "binary32 states with N <= 8192: sorted resamplers only (docs/SPEC.md §9)"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  if (params) set_params(kind_, params);
  if (!record_anc_) anc_rows_ = 0;
  begin_call();
  launch_step(1, y, resampler, proposal);  // (the statistics of the old weights, if it has to recompute them, go to slot 1)
  if (legacy_multinomial(resampler)) launch_scan(0, false);
  else launch_sum(0);            // statistics of the new weights now; the next sorted step reuses everything else
  SMCB_CUDA_TRY(cudaMemcpyAsync(&last_, stats_dev_, sizeof(StepStats), cudaMemcpyDeviceToHost, stream_));
  end_call();
  if (st) *st = last_;
}

void SingleFilter::run(int kind, const double* params, int64_t N, const double* y, int64_t T, int resampler,
                       const RngKey& key, uint32_t stream_id, StepStats* stats_out, const double* proposal) {
  check_args(kind, N, resampler);
  if (next_prec_ && resampler == RESAMPLE_MULTINOMIAL && N <= kMnLegacyMax) throw Error{SMCB_ERR_BAD_ARG, "binary32 states with N <= 8192: sorted resamplers only (docs/SPEC.md §9)"};
  if (T < 1) throw Error{SMCB_ERR_BAD_ARG, "T must be >= 1"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  N_ = 0;
  ensure_capacity(kind, N, record_anc_ ? std::max<int64_t>(T - 1, 1) : 1);
  if (cap_stats_ < T) {
    cudaFree(stats_dev_); stats_dev_ = nullptr;
    SMCB_CUDA_TRY(cudaMalloc(&stats_dev_, sizeof(StepStats) * T));
    cap_stats_ = T;
  }
  kind_ = kind; d_ = state_dim(kind); N_ = N; ld_ = cap_N_; prec_ = next_prec_;
  S_ = quant_shift((uint64_t)N); R_ = strata_width((uint64_t)N);
  key_ = key; stream_id_ = stream_id; t_ = 0; cur_ = 0; anc_rows_ = 0;
  set_params(kind, params);
  begin_call();
  launch_init(y[0]);
  for (int64_t t = 1; t < T; ++t) {
    launch_step(t - 1, y[t], resampler, proposal ? proposal + 3 * t : nullptr);  // stats of time t-1, then the step to time t
  }
  if (legacy_multinomial(resampler)) launch_scan(T - 1, false);
  else launch_sum(T - 1);
  std::vector<StepStats> tmp;
  SMCB_CUDA_TRY(cudaMemcpyAsync(stats_out, stats_dev_, sizeof(StepStats) * T, cudaMemcpyDeviceToHost, stream_));
  end_call();
  last_ = stats_out[T - 1];
}

// weighted (or plain) mean, variance and quantiles of every state component of the current cloud
void SingleFilter::summary(const double* probs, int np, bool weighted, double* mean_out, double* var_out, double* q_out) {
  if (!live()) throw Error{SMCB_ERR_STATE, "no filter state to summarise"};
  if (np < 0 || np > kMaxProbs) throw Error{SMCB_ERR_BAD_ARG, "at most 16 probabilities per call"};
  for (int j = 0; j < np; ++j)
    if (!(probs[j] >= 0.0 && probs[j] <= 1.0)) throw Error{SMCB_ERR_BAD_ARG, "probabilities must lie in [0, 1]"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  StepIndex ix;
  unsigned long long* cl = step_index(ix);
  unsigned long long Q = (unsigned long long)N_;
  if (weighted) {
    if (cap_stats_ < 2) {  // slot 1 is scratch for the statistics sum_kernel writes
      StepStats* nd = nullptr;
      SMCB_CUDA_TRY(cudaMalloc(&nd, sizeof(StepStats) * 2));
      if (stats_dev_) SMCB_CUDA_TRY(cudaMemcpyAsync(nd, stats_dev_, sizeof(StepStats) * cap_stats_, cudaMemcpyDeviceToDevice, stream_));
      SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
      cudaFree(stats_dev_);
      stats_dev_ = nd;
      cap_stats_ = 2;
    }
    if (!sum_done_) launch_sum(1);  // the tile-local CDF of the current weights (and Q)
    FilterCtrl hc;
    SMCB_CUDA_TRY(cudaMemcpyAsync(&hc, ctrl_, sizeof(FilterCtrl), cudaMemcpyDeviceToHost, stream_));
    SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
    Q = hc.total;
  }
  const int nblk = (int)std::min<int64_t>((N_ + kSumThreadsW - 1) / kSumThreadsW, (int64_t)num_sms_ * 8);
  const size_t scratch_words = (size_t)nblk + 2 * kMaxProbs + (size_t)kMaxProbs * 256;
  if (summary_cap_ < scratch_words) {
    cudaFree(summary_dev_); summary_dev_ = nullptr; summary_cap_ = 0;
    SMCB_CUDA_TRY(cudaMalloc(&summary_dev_, sizeof(unsigned long long) * scratch_words));
    summary_cap_ = scratch_words;
  }
  double* part = reinterpret_cast<double*>(summary_dev_);
  unsigned long long* prefix = summary_dev_ + nblk;
  unsigned long long* rank = prefix + kMaxProbs;
  unsigned long long* hist = rank + kMaxProbs;
  std::vector<double> hpart((size_t)nblk);
  for (int c = 0; c < d_; ++c) {
    const double* xc = x_[cur_] + (int64_t)c * ld_;
    const float* xcf = reinterpret_cast<const float*>(x_[cur_]) + (int64_t)c * ld_;
    double mean = NAN, var = NAN;
    if (Q != 0 && (mean_out || var_out)) {
      for (int mode = 0; mode < (var_out ? 2 : 1); ++mode) {
        if (prec_) wmoment_kernel<float><<<nblk, kSumThreadsW, 0, stream_>>>(xcf, cl, N_, ix.tile_items, weighted ? 1 : 0, mode, mean, part);
        else wmoment_kernel<double><<<nblk, kSumThreadsW, 0, stream_>>>(xc, cl, N_, ix.tile_items, weighted ? 1 : 0, mode, mean, part);
        SMCB_CUDA_TRY(cudaGetLastError());
        SMCB_CUDA_TRY(cudaMemcpyAsync(hpart.data(), part, sizeof(double) * nblk, cudaMemcpyDeviceToHost, stream_));
        SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
        double t = 0.0;
        for (int b = 0; b < nblk; ++b) t += hpart[(size_t)b];
        if (mode == 0) mean = t / (double)Q; else var = t / (double)Q;
      }
    }
    if (mean_out) mean_out[c] = mean;
    if (var_out) var_out[c] = var;
    if (np > 0 && q_out) {
      if (Q == 0) {
        for (int j = 0; j < np; ++j) q_out[(size_t)c * np + j] = NAN;
        continue;
      }
      unsigned long long hr[kMaxProbs], hp[kMaxProbs];
      for (int j = 0; j < np; ++j) {  // SPEC §8: r = min(floor(p Q), Q - 1); the quantile is the smallest x with mass(<= x) > r
        unsigned long long r = (unsigned long long)(probs[j] * (double)Q);
        hr[j] = r > Q - 1 ? Q - 1 : r;
        hp[j] = 0ull;
      }
      SMCB_CUDA_TRY(cudaMemcpyAsync(rank, hr, sizeof(unsigned long long) * np, cudaMemcpyHostToDevice, stream_));
      SMCB_CUDA_TRY(cudaMemcpyAsync(prefix, hp, sizeof(unsigned long long) * np, cudaMemcpyHostToDevice, stream_));
      SMCB_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * kMaxProbs * 256, stream_));
      for (int pass = 0; pass < 8; ++pass) {
        if (prec_) rsel_hist_kernel<float><<<nblk, kSumThreadsW, 0, stream_>>>(xcf, cl, N_, ix.tile_items, weighted ? 1 : 0, pass, np, prefix, hist);
        else rsel_hist_kernel<double><<<nblk, kSumThreadsW, 0, stream_>>>(xc, cl, N_, ix.tile_items, weighted ? 1 : 0, pass, np, prefix, hist);
        rsel_pick_kernel<<<1, 256, 0, stream_>>>(pass, np, prefix, rank, hist);
      }
      SMCB_CUDA_TRY(cudaGetLastError());
      SMCB_CUDA_TRY(cudaMemcpyAsync(hp, prefix, sizeof(unsigned long long) * np, cudaMemcpyDeviceToHost, stream_));
      SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
      for (int j = 0; j < np; ++j) {
        const unsigned long long e = hp[j];
        const unsigned long long b = (e >> 63) ? (e & 0x7FFFFFFFFFFFFFFFull) : ~e;  // decode_ordered on the host
        double v;
        std::memcpy(&v, &b, sizeof v);
        q_out[(size_t)c * np + j] = v;
      }
    }
  }
}

void SingleFilter::fetch(double* x_host, double* w_host, double* logw_host) {
  if (!live()) throw Error{SMCB_ERR_STATE, "no filter state to fetch"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  if (x_host && !prec_)
    SMCB_CUDA_TRY(cudaMemcpy2DAsync(x_host, sizeof(double) * N_, x_[cur_], sizeof(double) * ld_, sizeof(double) * N_,
                                    d_, cudaMemcpyDeviceToHost, stream_));
  if (x_host && prec_) {  // binary32 states are widened on the device: the host-facing layout stays [d][N] doubles
    if (!w_tmp_) SMCB_CUDA_TRY(cudaMalloc(&w_tmp_, sizeof(double) * cap_N_));
    for (int c = 0; c < d_; ++c) {
      widen_kernel<<<(unsigned)((N_ + 255) / 256), 256, 0, stream_>>>(reinterpret_cast<const float*>(x_[cur_]) + (int64_t)c * ld_, w_tmp_, N_);
      SMCB_CUDA_TRY(cudaGetLastError());
      SMCB_CUDA_TRY(cudaMemcpyAsync(x_host + (int64_t)c * N_, w_tmp_, sizeof(double) * N_, cudaMemcpyDeviceToHost, stream_));
    }
  }
  if (logw_host || w_host) ensure_logw();
  if (logw_host && prec_ == 2) {  // binary32 log-weights, widened like the states
    if (!w_tmp_) SMCB_CUDA_TRY(cudaMalloc(&w_tmp_, sizeof(double) * cap_N_));
    widen_kernel<<<(unsigned)((N_ + 255) / 256), 256, 0, stream_>>>(reinterpret_cast<const float*>(logw_[cur_]), w_tmp_, N_);
    SMCB_CUDA_TRY(cudaGetLastError());
    SMCB_CUDA_TRY(cudaMemcpyAsync(logw_host, w_tmp_, sizeof(double) * N_, cudaMemcpyDeviceToHost, stream_));
  } else if (logw_host)
    SMCB_CUDA_TRY(cudaMemcpyAsync(logw_host, logw_[cur_], sizeof(double) * N_, cudaMemcpyDeviceToHost, stream_));
  if (w_host) {
    if (!w_tmp_) SMCB_CUDA_TRY(cudaMalloc(&w_tmp_, sizeof(double) * cap_N_));
    if (prec_ == 2) weights_kernel_f<<<(unsigned)((N_ + 255) / 256), 256, 0, stream_>>>(reinterpret_cast<const float*>(logw_[cur_]), w_tmp_, N_, last_);
    else
    weights_kernel<<<(unsigned)((N_ + 255) / 256), 256, 0, stream_>>>(logw_[cur_], w_tmp_, N_, last_);
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
  t_ = 0; cur_ = 0; anc_rows_ = 0; from_w_ = !is_log; logw_valid_ = true; sum_done_ = false;
  SMCB_CUDA_TRY(cudaMemsetAsync(desc_, 0, sizeof(unsigned long long) * 2 * ntiles_cap_, stream_));
  desc_clean_t_ = 0;
  reset_ctrl_kernel<<<1, 1, 0, stream_>>>(ctrl_);
  SMCB_CUDA_TRY(cudaMemcpyAsync(logw_[cur_], host, sizeof(double) * n, cudaMemcpyHostToDevice, stream_));
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 1184);
  max_kernel<<<grid, 256, 0, stream_>>>(logw_[cur_], n, ctrl_);
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
                                   uint32_t t, uint32_t purpose, int64_t* anc_host, int64_t n_out) {
  if (resampler < 0 || resampler > 2) throw Error{SMCB_ERR_BAD_ARG, "unknown resampler"};
  if (n_out <= 0) n_out = n;
  load_vector(w_host, n, false);
  begin_call();
  launch_scan(0, true);
  from_w_ = false;
  if (util_cap_ < n_out) {  // scratch of the utility: kept across calls (a cudaMalloc / cudaFree pair per call cost more than the kernels)
    cudaFree(util_anc_);
    util_anc_ = nullptr;
    util_cap_ = 0;
    SMCB_CUDA_TRY(cudaMalloc(&util_anc_, sizeof(int64_t) * n_out));
    util_cap_ = n_out;
  }
  int64_t* anc_dev = util_anc_;
  ancestor_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, stream_>>>(cdf_, n, n_out, resampler, strata_width((uint64_t)n_out), key, stream_id, t,
                                                                        purpose, ctrl_, anc_dev);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(anc_host, anc_dev, sizeof(int64_t) * n_out, cudaMemcpyDeviceToHost, stream_);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
  N_ = 0;
  SMCB_CUDA_TRY(e);
  end_call();
}

}  // namespace smcb
