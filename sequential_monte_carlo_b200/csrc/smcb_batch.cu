// Batched bootstrap filters: one CTA per θ-particle, cloud + CDF resident in shared memory for the
// whole series, so a sweep of T observations is ONE launch and touches HBM only for the final
// clouds (docs/SPEC.md §5-§7; same arithmetic as the grid-wide kernels of smcb_filter.cu, so the
// ancestors are bit-identical to the oracle's).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "smcb_batch.cuh"

namespace smcb {

namespace {

struct BatchArgs {
  const double* derived;   // [M][8]
  const uint8_t* active;   // [M] or null
  const double* y;         // observations consumed by this launch, y[0] is for time t_begin
  const double* prop;      // guided kernels only: [t_end - t_begin + 1][M][5] proposal coefficients (SPEC §10), row 0 is for t_begin
  double* x;               // [M][d][ld]
  double* logw;            // [M][ld]
  StepStats* stats;        // [M]
  double* logz_out;        // [M]  Σ logμ over the steps this launch accounts for
  double* ess_out;         // [M]
  int64_t N, ld;
  int S;
  uint64_t R;
  RngKey key;
  uint32_t stream0;
  uint32_t t_begin, t_end;  // inclusive range of time indices assimilated by this launch
  int resampler;
  int from_init;            // t_begin == 0 draws the cloud; else continue from (x, logw, stats)
  int x_in_smem;
  int64_t npad;             // N rounded up to even
  int64_t M;                // θ-particles of this launch (the grid of the static kernel)
  // dynamic (θ, chunk) scheduling (DYN kernels): the series is cut into nchunks chunks of `chunk` steps, the resident CTAs claim
  // units u = c·M + m from sched[0] in order, and sched[1 + m] counts the finished chunks of θ-particle m
  unsigned* sched;
  uint32_t chunk, nchunks;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// the transition density of the guided kernels; empty for the bootstrap ones
template <class Model, bool GUIDED>
struct GuidedDensity : TransDensity<Model> {};
template <class Model>
struct GuidedDensity<Model, false> {
  __device__ __forceinline__ void load(const double*) {}
};

__device__ __forceinline__ int smem_lower_count(const uint64_t* C, int lo, int hi, uint64_t tau) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (C[mid] <= tau) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
// the same count for a threshold known to be >= the one that gave `lo` (sorted resamplers: the second particle of a pair):
// its ancestor is lo or a few entries further, so gallop from lo instead of bisecting [lo, hi) from scratch
__device__ __forceinline__ int smem_lower_count_from(const uint64_t* C, int lo, int hi, uint64_t tau) {
  int step = 1, top = lo;
  while (top < hi && C[top] <= tau) {
    lo = top + 1;
    top = lo + step - 1;
    step <<= 1;
    if (top > hi) top = hi;
  }
  // invariant: every entry below lo is <= tau; C[top] > tau or top == hi
  return smem_lower_count(C, lo, top, tau);
}

// GUIDED: the move of every step t >= 1 draws from the affine-Gaussian proposal of (t, θ) and the weight carries
// transition / proposal (particle_filter!, particles.jl:66-80; SPEC §10); D = 1 models only.
// One unit of work: θ-particle m, observations tb..te (the whole launch for the static kernel, one chunk for the dynamic one).
// from_init: tb == 0 draws the cloud; else the unit continues from (x, logw, stats) in global memory, and with `carry` also
// from the running Σ logμ in logz_out[m] — the sum is continued in the same order, so a chunked series gives the bits of an
// unchunked one.
template <class Model, int PAIRS, bool GUIDED>
__device__ __forceinline__ void batch_unit(const BatchArgs& a, const int64_t m, const uint32_t tb, const uint32_t te, const bool from_init,
                                           const bool carry, unsigned char* smem_raw, unsigned long long (*s_wq)[32], double (*s_f)[32]) {
  constexpr int D = Model::D;
  static_assert(!GUIDED || D == 1 || Model::KIND == KIND_UCSV, "guided proposals: affine-Gaussian for D = 1 (SPEC §10), the tempered trend move for UCSV (§10b)");
  __shared__ unsigned long long s_sys;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  if (a.active && !a.active[m]) {
    if (tid == 0) {
      a.logz_out[m] = -INFINITY;
      a.ess_out[m] = 0.0;
    }
    return;
  }
  const int N = (int)a.N;
  const int64_t ld = a.ld;
  const int npairs = (N + 1) >> 1;
  uint64_t* s_cdf = reinterpret_cast<uint64_t*>(smem_raw);
  double* xs = a.x_in_smem ? reinterpret_cast<double*>(smem_raw + sizeof(uint64_t) * a.npad) : (a.x + m * D * ld);
  const int64_t ldx = a.x_in_smem ? a.npad : ld;
  double* gx = a.x + m * D * ld;
  double* glw = a.logw + m * ld;

  Model mdl;
  mdl.load(a.derived + m * kParamStride);
  GuidedDensity<Model, GUIDED> fdens;
  fdens.load(a.derived + m * kParamStride);
  const uint32_t stream = a.stream0 + (uint32_t)m;
  const double logN = log((double)N);

  double lw[PAIRS][2];
  double mx = -INFINITY;
  double logz = carry ? __ldcg(a.logz_out + m) : 0.0;
  bool account = true;

  // block-uniform max of the thread-local log-weights (exact, order-independent)
  auto block_max_all = [&](double v) -> double {
    v = warp_max(v);
    if (lane == 0) s_f[2][warp] = v;
    __syncthreads();
    double r = (lane < nwarps) ? s_f[2][lane] : -INFINITY;
    return warp_max(r);
  };

  if (!from_init) {
    // continue from the stored cloud (L2 loads: under dynamic scheduling another SM may have written it)
#pragma unroll
    for (int r = 0; r < PAIRS; ++r) {
      const int p = r * nthreads + tid;
      const int i = 2 * p;
      lw[r][0] = (i < N) ? __ldcg(glw + i) : -INFINITY;
      lw[r][1] = (i + 1 < N) ? __ldcg(glw + i + 1) : -INFINITY;
      if (a.x_in_smem && p < npairs) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
          xs[k * ldx + i] = __ldcg(gx + k * ld + i);
          if (i + 1 < N) xs[k * ldx + i + 1] = __ldcg(gx + k * ld + i + 1);
        }
      }
    }
    mx = __ldcg(&a.stats[m].mx);
    account = false;  // the stored weights' logμ was reported by the launch that produced them
    __syncthreads();
  }

  for (uint32_t t = tb; t <= te; ++t) {
    const double y = a.y[t - a.t_begin];
    if (t == 0 && from_init) {
      // bootstrap_filter: x_i ~ initial_dist; logw_i = logpdf(observation(x_i), y1)   particles.jl:96-99
#pragma unroll
      for (int r = 0; r < PAIRS; ++r) {
        const int p = r * nthreads + tid;
        const int i = 2 * p;
        lw[r][0] = lw[r][1] = -INFINITY;
        if (p < npairs) {
          double za[D], zb[D], xa[D], xb[D];
#pragma unroll
          for (int k = 0; k < D; ++k) normal_pair_at(a.key, (uint32_t)p, stream, 0u, PURPOSE_INIT, (uint32_t)k, za[k], zb[k]);
          mdl.init(za, xa);
          mdl.init(zb, xb);
          lw[r][0] = mdl.logweight(xa, y);
#pragma unroll
          for (int k = 0; k < D; ++k) xs[k * ldx + i] = xa[k];
          if (i + 1 < N) {
            lw[r][1] = mdl.logweight(xb, y);
#pragma unroll
            for (int k = 0; k < D; ++k) xs[k * ldx + i + 1] = xb[k];
          }
        }
      }
    } else {
      // the proposal of this step is requested first: its L2 / HBM round trip hides behind the normalisation and the search
      double pc[kProposalStride] = {0.0, 0.0, 0.0, 0.0, 0.0};
      if constexpr (GUIDED) {
        const double* g = a.prop + ((int64_t)(t - a.t_begin) * a.M + m) * kProposalStride;
#pragma unroll
        for (int k = 0; k < kProposalStride; ++k) pc[k] = __ldg(g + k);
      }
      // ---- normalize(previous logw) and its fixed-point CDF                         particles.jl:5-15,117
      unsigned long long q0[PAIRS], tq[PAIRS], winc[PAIRS];
      double se = 0.0;  // (Σe² is only needed for the ess of the LAST step of the launch: the final normalisation below computes it)
#pragma unroll
      for (int r = 0; r < PAIRS; ++r) {
        double e0 = 0.0, e1 = 0.0;
        uint64_t qa = 0, qb = 0;
        const int i = 2 * (r * nthreads + tid);
        if (i < N) det_exp_quant(lw[r][0] - mx, a.S, e0, qa);
        if (i + 1 < N) det_exp_quant(lw[r][1] - mx, a.S, e1, qb);
        se += e0;
        se += e1;
        q0[r] = qa;
        tq[r] = qa + qb;
        winc[r] = warp_scan_u64(tq[r], lane);
        if (lane == 31) s_wq[r][warp] = winc[r];
      }
      se = warp_sum(se);
      if (lane == 0) s_f[0][warp] = se;
      // the one uniform of the systematic resampler (SPEC §5): a Philox block for thread 0, not for every thread of the CTA
      if (tid == 0 && a.resampler == RESAMPLE_SYSTEMATIC)
        s_sys = mulhi64(uniform64_at(a.key, 0u, stream, t, PURPOSE_RESAMPLE), a.R);
      __syncthreads();  // A
      unsigned long long segbase = 0;
#pragma unroll
      for (int r = 0; r < PAIRS; ++r) {
        const unsigned long long v = (lane < nwarps) ? s_wq[r][lane] : 0ull;
        const unsigned long long vinc = warp_scan_u64(v, lane);
        const unsigned long long wexcl = __shfl_sync(kFullMask, vinc - v, warp);
        const unsigned long long segtot = __shfl_sync(kFullMask, vinc, 31);
        const unsigned long long excl = segbase + wexcl + (winc[r] - tq[r]);
        const int i = 2 * (r * nthreads + tid);
        if (i < a.npad) {
          s_cdf[i] = excl + q0[r];
          s_cdf[i + 1] = excl + tq[r];
        }
        segbase += segtot;
      }
      const unsigned long long Q = segbase;
      if (warp == 0) {  // logμ of the previous step (particles.jl:10): thread 0 carries Σ logμ, nobody else needs it
        double e1 = (lane < nwarps) ? s_f[0][lane] : 0.0;
        e1 = warp_sum(e1);
        if (account && tid == 0) logz += mx + log(e1) - logN;
      }
      account = true;
      __syncthreads();  // B: CDF complete
      // ---- a = resample(w); xp = x[a]                                                particles.jl:117-119
      const uint64_t sys_off = (a.resampler == RESAMPLE_SYSTEMATIC) ? s_sys : 0ull;
      double xpa[PAIRS][D], xpb[PAIRS][D];
#pragma unroll
      for (int r = 0; r < PAIRS; ++r) {
        const int p = r * nthreads + tid;
        const int i = 2 * p;
        if (p < npairs) {
          int a0 = i, a1 = (i + 1 < N) ? i + 1 : i;
          if (Q != 0) {
            uint64_t ua = sys_off, ub = sys_off;
            if (a.resampler != RESAMPLE_SYSTEMATIC) {
              const Philox4 b = philox4x32_10((uint32_t)p, stream, t, purpose_word(PURPOSE_RESAMPLE, 0, a.key.epoch),
                                              a.key);
              ua = uniform64_of(b, 0);
              ub = uniform64_of(b, 1);
            }
            uint64_t F0, F1;
            if (a.resampler == RESAMPLE_MULTINOMIAL) {
              F0 = ua;
              F1 = ub;
            } else if (a.resampler == RESAMPLE_STRATIFIED) {
              F0 = (uint64_t)i * a.R + mulhi64(ua, a.R);
              F1 = (uint64_t)(i + 1) * a.R + mulhi64(ub, a.R);
            } else {
              F0 = (uint64_t)i * a.R + ua;
              F1 = (uint64_t)(i + 1) * a.R + ub;
            }
            const uint64_t t0 = mulhi64(F0, Q), t1 = mulhi64(F1, Q);
            a0 = smem_lower_count(s_cdf, 0, N - 1, t0);
            a1 = (a.resampler == RESAMPLE_MULTINOMIAL) ? smem_lower_count(s_cdf, 0, N - 1, t1) : smem_lower_count_from(s_cdf, a0, N - 1, t1);
          }
#pragma unroll
          for (int k = 0; k < D; ++k) {
            xpa[r][k] = xs[k * ldx + a0];
            xpb[r][k] = xs[k * ldx + a1];
          }
        }
      }
      __syncthreads();  // C: every parent read before the cloud is overwritten
      // ---- x_i ~ transition(xp_i); logw_i = logpdf(observation(x_i), y)              particles.jl:122-125
      // (guided: x_i ~ proposal(xp_i); logw_i += logpdf(transition(xp_i), x_i) - logpdf(proposal(xp_i), x_i)   :73-78)
#pragma unroll
      for (int r = 0; r < PAIRS; ++r) {
        const int p = r * nthreads + tid;
        const int i = 2 * p;
        lw[r][0] = lw[r][1] = -INFINITY;
        if (p < npairs) {
          double za[D], zb[D], xa[D], xb[D];
#pragma unroll
          for (int k = 0; k < D; ++k)
            normal_pair_at(a.key, (uint32_t)p, stream, t, PURPOSE_TRANSITION, (uint32_t)k, za[k], zb[k]);
          if constexpr (GUIDED && D == 1) lw[r][0] = guided_move(mdl, fdens, pc, za[0], xpa[r][0], y, xa);
          else if constexpr (GUIDED) lw[r][0] = guided_move_ucsv(mdl, pc[0], za, xpa[r], y, xa);
          else {
            mdl.transition(za, xpa[r], xa);
            lw[r][0] = mdl.logweight(xa, y);
          }
#pragma unroll
          for (int k = 0; k < D; ++k) xs[k * ldx + i] = xa[k];
          if (i + 1 < N) {
            if constexpr (GUIDED && D == 1) lw[r][1] = guided_move(mdl, fdens, pc, zb[0], xpb[r][0], y, xb);
            else if constexpr (GUIDED) lw[r][1] = guided_move_ucsv(mdl, pc[0], zb, xpb[r], y, xb);
            else {
              mdl.transition(zb, xpb[r], xb);
              lw[r][1] = mdl.logweight(xb, y);
            }
#pragma unroll
            for (int k = 0; k < D; ++k) xs[k * ldx + i + 1] = xb[k];
          }
        }
      }
    }
    double v = -INFINITY;
#pragma unroll
    for (int r = 0; r < PAIRS; ++r) {
      if (lw[r][0] > v) v = lw[r][0];
      if (lw[r][1] > v) v = lw[r][1];
    }
    mx = block_max_all(v);  // sync D
  }

  // ---- normalize(final logw): logμ, ess, stats for the next launch
  {
    double se = 0.0, se2 = 0.0;
#pragma unroll
    for (int r = 0; r < PAIRS; ++r) {
      const int i = 2 * (r * nthreads + tid);
      if (i < N) { const double e = det_exp(lw[r][0] - mx); se += e; se2 += e * e; }
      if (i + 1 < N) { const double e = det_exp(lw[r][1] - mx); se += e; se2 += e * e; }
    }
    se = warp_sum(se);
    se2 = warp_sum(se2);
    __syncthreads();
    if (lane == 0) {
      s_f[0][warp] = se;
      s_f[1][warp] = se2;
    }
    __syncthreads();
    double e1 = (lane < nwarps) ? s_f[0][lane] : 0.0, e2 = (lane < nwarps) ? s_f[1][lane] : 0.0;
    e1 = warp_sum(e1);
    e2 = warp_sum(e2);
    if (account && tid == 0) logz += mx + log(e1) - logN;
    if (tid == 0) {
      a.stats[m].mx = mx;
      a.stats[m].sum = e1;
      a.stats[m].sum2 = e2;
      a.logz_out[m] = logz;
      a.ess_out[m] = (e1 * e1) / e2;
    }
  }
#pragma unroll
  for (int r = 0; r < PAIRS; ++r) {
    const int p = r * nthreads + tid;
    const int i = 2 * p;
    if (p < npairs) {
      if (i + 1 < N) {
        *reinterpret_cast<double2*>(glw + i) = make_double2(lw[r][0], lw[r][1]);
        if (a.x_in_smem) {
#pragma unroll
          for (int k = 0; k < D; ++k)
            *reinterpret_cast<double2*>(gx + k * ld + i) = make_double2(xs[k * ldx + i], xs[k * ldx + i + 1]);
        }
      } else {
        glw[i] = lw[r][0];
        if (a.x_in_smem) {
#pragma unroll
          for (int k = 0; k < D; ++k) gx[k * ld + i] = xs[k * ldx + i];
        }
      }
    }
  }
}

// Static kernel: one CTA per θ-particle, the whole series.  DYN: a persistent grid (one wave of resident CTAs) claims
// (chunk, θ) units in order — 512 θ on 148 one-CTA SMs are 3.46 waves of whole series (4 with one CTA per θ, 13.5 % of the
// launch idle: the θ-sharded config 5 on 8 GPUs), but 3.46·C waves of chunks.  Unit (c, m) needs unit (c − 1, m), which was
// claimed M units earlier by a CTA that is running: waiting on it cannot deadlock.
// (the scheduling loop of the small-CTA kernels of the one-component models must not cost them a resident CTA: same register cap as
// the static kernels, which fit 1024 / MAXT CTAs of 64-register threads)
template <class Model, int PAIRS, int MAXT, bool GUIDED, bool DYN>
__global__ void __launch_bounds__(MAXT, (DYN && Model::D == 1 && PAIRS <= 2 && !GUIDED) ? 1024 / MAXT : 1) batch_kernel(const BatchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned long long s_wq[PAIRS][32];
  __shared__ double s_f[3][32];
  if constexpr (!DYN) {
    batch_unit<Model, PAIRS, GUIDED>(a, (int64_t)blockIdx.x, a.t_begin, a.t_end, a.from_init != 0, false, smem_raw, s_wq, s_f);
  } else {
    __shared__ unsigned s_unit;
    const unsigned M = (unsigned)a.M, nunits = M * a.nchunks;
    for (;;) {
      if (threadIdx.x == 0) s_unit = atomicAdd(a.sched, 1u);
      __syncthreads();
      const unsigned u = s_unit;
      if (u >= nunits) break;
      const unsigned c = u / M, m = u - c * M;
      const bool act = !(a.active && !a.active[m]);
      if (act && c > 0) {
        if (threadIdx.x == 0) {
          while (ld_acquire_u32(a.sched + 1 + m) < c) __nanosleep(100);
          __threadfence();
        }
        __syncthreads();  // the acquire above orders every thread's reads of the cloud after the producer's writes
      }
      if (act || c == 0) {
        const uint32_t tb = a.t_begin + c * a.chunk;
        const uint32_t te = (tb + a.chunk - 1 < a.t_end) ? tb + a.chunk - 1 : a.t_end;
        batch_unit<Model, PAIRS, GUIDED>(a, (int64_t)m, tb, te, a.from_init != 0 && c == 0, c > 0, smem_raw, s_wq, s_f);
      }
      __threadfence();
      __syncthreads();  // also: everybody has read s_unit and is done with the shared scratch
      if (act && threadIdx.x == 0) st_release_u32(a.sched + 1 + m, c + 1);
    }
  }
}

// rows = whole clouds: dst slot <- src slot, 16 bytes per thread per trip
__global__ void copy_clouds_kernel(double* __restrict__ dx, const double* __restrict__ sx, double* __restrict__ dlw,
                                   const double* __restrict__ slw, StepStats* __restrict__ dst, const StepStats* __restrict__ sst,
                                   const int32_t* __restrict__ dslot, const int32_t* __restrict__ sslot,
                                   const uint8_t* __restrict__ mask, int64_t xrow, int64_t wrow) {
  const int64_t j = blockIdx.x;
  if (mask && !mask[j]) return;
  const int64_t d = dslot ? dslot[j] : j, s = sslot ? sslot[j] : j;
  const double2* s2 = reinterpret_cast<const double2*>(sx + s * xrow);
  double2* d2 = reinterpret_cast<double2*>(dx + d * xrow);
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < xrow / 2; i += (int64_t)gridDim.y * blockDim.x) d2[i] = s2[i];
  const double2* w2 = reinterpret_cast<const double2*>(slw + s * wrow);
  double2* e2 = reinterpret_cast<double2*>(dlw + d * wrow);
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < wrow / 2; i += (int64_t)gridDim.y * blockDim.x) e2[i] = w2[i];
  if (blockIdx.y == 0 && threadIdx.x == 0) dst[d] = sst[s];
}

// pack / unpack whole clouds into a flat device buffer: per slot [x: d*ld][logw: ld][stats: 4 doubles]
__global__ void pack_clouds_kernel(double* __restrict__ x, double* __restrict__ lw, StepStats* __restrict__ st,
                                   const int32_t* __restrict__ slots, double* __restrict__ buf, int64_t xrow, int64_t wrow,
                                   int to_buffer) {
  const int64_t j = blockIdx.x;
  const int64_t s = slots[j];
  double* b = buf + j * (xrow + wrow + 4);
  double* gx = x + s * xrow;
  double* gw = lw + s * wrow;
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < xrow; i += (int64_t)gridDim.y * blockDim.x) {
    if (to_buffer) b[i] = gx[i];
    else gx[i] = b[i];
  }
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < wrow; i += (int64_t)gridDim.y * blockDim.x) {
    if (to_buffer) b[xrow + i] = gw[i];
    else gw[i] = b[xrow + i];
  }
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    double* sb = b + xrow + wrow;
    if (to_buffer) { sb[0] = st[s].mx; sb[1] = st[s].sum; sb[2] = st[s].sum2; sb[3] = 0.0; }
    else { st[s].mx = sb[0]; st[s].sum = sb[1]; st[s].sum2 = sb[2]; }
  }
}

__global__ void batch_weights_kernel(const double* __restrict__ logw, const StepStats* __restrict__ st, double* __restrict__ w,
                                     int64_t N, int64_t ld) {
  const int64_t m = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) w[m * N + i] = det_exp(logw[m * ld + i] - st[m].mx) / st[m].sum;
}

// mean[m][c] = w[m]' x[m][c]  (plotting_utils.jl:116-124,150: the weighted state mean of every θ-particle's cloud);
// one CTA per (m, c), fixed summation order
__global__ void __launch_bounds__(256)
    batch_wmean_kernel(const double* __restrict__ x, const double* __restrict__ logw, const StepStats* __restrict__ st, double* __restrict__ mean,
                       int64_t N, int64_t ld, int d) {
  __shared__ double sh[8];
  const int64_t m = blockIdx.x;
  const int c = blockIdx.y;
  const double mx = st[m].mx;
  const double* xc = x + (m * d + c) * ld;
  const double* lw = logw + m * ld;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += 256) {
    const double e = det_exp(lw[i] - mx);
    if (e != 0.0) acc += e * xc[i];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    mean[m * d + c] = t / st[m].sum;
  }
}

// mean[m][c] as above and var[m][c] = Σ_i w_i (x_i − mean)²  (var(x[m], weights(w[m])), the per-θ counterpart of
// examples/inflation_example.jl:46: population variance, corrected = false); one CTA per (m, c), two passes over
// the cloud, fixed summation order
__global__ void __launch_bounds__(256)
    batch_wmoment_kernel(const double* __restrict__ x, const double* __restrict__ logw, const StepStats* __restrict__ st,
                         double* __restrict__ mean, double* __restrict__ var, int64_t N, int64_t ld, int d) {
  __shared__ double sh[8];
  __shared__ double s_mean;
  const int64_t m = blockIdx.x;
  const int c = blockIdx.y;
  const double mx = st[m].mx;
  const double* xc = x + (m * d + c) * ld;
  const double* lw = logw + m * ld;
  for (int pass = 0; pass < 2; ++pass) {
    const double mu = pass ? s_mean : 0.0;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += 256) {
      const double e = det_exp(lw[i] - mx);
      if (e != 0.0) {
        const double dx = xc[i] - mu;
        acc += pass ? e * (dx * dx) : e * dx;
      }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += sh[w];
      t /= st[m].sum;
      if (pass) var[m * d + c] = t;
      else { mean[m * d + c] = t; s_mean = t; }
    }
    __syncthreads();
  }
}

// q[m][c][j] = weighted (or plain) lower empirical quantile p_j of component c of θ-particle m's cloud (SPEC §8):
// the per-θ bands of get_quantiles_uc / get_quantiles_ucsv (examples/inflation_example.jl:39-55,241-253).  One
// CTA per (m, c): fixed-point weights q_i (SPEC §5), exact total, then an 8-pass radix select (one byte per
// pass, most significant first) on the order-preserving bit pattern of x for all np probabilities at once.
constexpr int kBatchMaxProbs = 16;
__global__ void __launch_bounds__(256)
    batch_wquantile_kernel(const double* __restrict__ x, const double* __restrict__ logw, const StepStats* __restrict__ st,
                           const double* __restrict__ probs, int np, int weighted, int S, double* __restrict__ out, int64_t N, int64_t ld, int d) {
  __shared__ unsigned long long s_hist[kBatchMaxProbs * 256];
  __shared__ unsigned long long s_prefix[kBatchMaxProbs], s_rank[kBatchMaxProbs];
  __shared__ unsigned long long s_red[8];
  const int64_t m = blockIdx.x;
  const int c = blockIdx.y, tid = threadIdx.x;
  const double mx = st[m].mx;
  const double* xc = x + (m * d + c) * ld;
  const double* lw = logw + m * ld;
  auto weight = [&](int64_t i) -> unsigned long long {
    if (!weighted) return 1ull;
    double e;
    uint64_t q;
    det_exp_quant(lw[i] - mx, S, e, q);
    return q;
  };
  unsigned long long tot = 0;
  for (int64_t i = tid; i < N; i += 256) tot += weight(i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFullMask, tot, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = tot;
  __syncthreads();
  unsigned long long Q = 0;
  for (int w = 0; w < 8; ++w) Q += s_red[w];
  if (Q == 0) {  // no mass (all weights underflowed / NaN): undefined
    if (tid < np) out[(m * d + c) * np + tid] = NAN;
    return;
  }
  if (tid < np) {
    unsigned long long r = (unsigned long long)(probs[tid] * (double)Q);  // r = min(floor(p Q), Q - 1)
    s_rank[tid] = r > Q - 1 ? Q - 1 : r;
    s_prefix[tid] = 0ull;
  }
  for (int pass = 0; pass < 8; ++pass) {
    for (int k = tid; k < np * 256; k += 256) s_hist[k] = 0ull;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    for (int64_t i = tid; i < N; i += 256) {
      const unsigned long long q = weight(i);
      if (q == 0ull) continue;
      const unsigned long long key = encode_ordered(xc[i]);
      const unsigned digit = (unsigned)(key >> shift) & 255u;
      for (int j = 0; j < np; ++j) {
        const bool match = (pass == 0) || ((key ^ s_prefix[j]) >> (shift + 8)) == 0ull;
        if (match) atomicAdd(&s_hist[j * 256 + digit], q);
      }
    }
    __syncthreads();
    if (tid < np) {
      unsigned long long r = s_rank[tid], cum = 0;
      int dg = 255;
      for (int b = 0; b < 256; ++b) {
        const unsigned long long h = s_hist[tid * 256 + b];
        if (cum + h > r) { dg = b; break; }
        cum += h;
      }
      s_rank[tid] = r - cum;
      s_prefix[tid] |= (unsigned long long)dg << shift;
    }
    __syncthreads();
  }
  if (tid < np) out[(m * d + c) * np + tid] = decode_ordered(s_prefix[tid]);
}

// scalar Kalman recursion, one thread per model                                 kalman_filter.jl:29-70
__global__ void kalman_kernel(const double* __restrict__ params, const uint8_t* __restrict__ active, int64_t M,
                              const double* __restrict__ y, int64_t T, int predict_first, double* __restrict__ loglik,
                              double* __restrict__ xio, double* __restrict__ sio, int use_state) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  if (active && !active[m]) {
    loglik[m] = -INFINITY;
    return;
  }
  const double* P = params + m * kParamStride;
  const double A = P[0], B = P[1], Q = P[2], R = P[3];
  double x = use_state ? xio[m] : P[4];
  double S = use_state ? sio[m] : P[5];
  double ll = 0.0;
  for (int64_t t = 0; t < T; ++t) {
    if (predict_first || t > 0) {  // :33-34
      x = A * x;
      S = (A * A) * S + Q;
    }
    const double sig = (B * B) * S + R;       // :37
    const double dy = y[t] - B * x;           // :38
    const double inv = 1.0 / sig;
    x = x + (S * B) * inv * dy;               // :41
    S = S - ((S * B) * (S * B)) * inv;        // :42
    ll += -0.5 * (log(2.0 * M_PI) + log(sig) + (dy / sig * dy));  // :45
  }
  loglik[m] = ll;
  if (xio) xio[m] = x;
  if (sio) sio[m] = S;
}

// matrix Kalman recursion with a scalar observation, one thread per model          kalman_filter.jl:3-27
// model block (row-major): A[D][D], B[D], Q[D][D], R, x0[D], Σ0[D][D]  = 3D² + 2D + 1 doubles
template <int D>
__global__ void kalman_mv_kernel(const double* __restrict__ models, const uint8_t* __restrict__ active, int64_t M,
                                 const double* __restrict__ y, int64_t T, int predict_first, double* __restrict__ loglik,
                                 double* __restrict__ xio, double* __restrict__ sio, int use_state) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  if (active && !active[m]) {
    loglik[m] = -INFINITY;
    return;
  }
  constexpr int STRIDE = 3 * D * D + 2 * D + 1;
  const double* P = models + m * STRIDE;
  double A[D][D], B[D], Q[D][D], x[D], S[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    B[i] = P[D * D + i];
    x[i] = use_state ? xio[m * D + i] : P[2 * D * D + D + 1 + i];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      A[i][j] = P[i * D + j];
      Q[i][j] = P[D * D + D + i * D + j];
      S[i][j] = use_state ? sio[(m * D + i) * D + j] : P[2 * D * D + 2 * D + 1 + i * D + j];
    }
  }
  const double R = P[2 * D * D + D];
  double ll = 0.0;
  for (int64_t t = 0; t < T; ++t) {
    if (predict_first || t > 0) {
      double xn[D], AS[D][D];
#pragma unroll
      for (int i = 0; i < D; ++i) {  // xt = A*xt                                       :13
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc += A[i][j] * x[j];
        xn[i] = acc;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double u = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) u += A[i][j] * S[j][k];
          AS[i][k] = u;
        }
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {  // Σt = A*Σt*A' + Q                                :14
        x[i] = xn[i];
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double u = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) u += AS[i][j] * A[k][j];
          S[i][k] = u + Q[i][k];
        }
      }
    }
    double K[D], bx = 0.0, sig = 0.0;  // K = Σt*B'
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double u = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) u += S[i][j] * B[j];
      K[i] = u;
      bx += B[i] * x[i];
    }
#pragma unroll
    for (int i = 0; i < D; ++i) sig += B[i] * K[i];
    sig += R;                       // σt = B*Σt*B' + R                                  :16
    const double dy = y[t] - bx;    // Δyt = yt - B*xt                                   :17
    const double inv = 1.0 / sig;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const double g = K[i] * inv;
      x[i] = x[i] + g * dy;         // xt + (Σt*B')*inv(σt)*Δyt                          :20
#pragma unroll
      for (int j = 0; j < D; ++j) S[i][j] = S[i][j] - g * K[j];  // Σt - (Σt*B')*inv(σt)*(B*Σt')  :21
    }
    ll += -0.5 * (log(2.0 * M_PI) + log(sig) + (dy * inv * dy));  // :24-26
  }
  loglik[m] = ll;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    if (xio) xio[m * D + i] = x[i];
#pragma unroll
    for (int j = 0; j < D; ++j)
      if (sio) sio[(m * D + i) * D + j] = S[i][j];
  }
}

// (c0, c1, c2) -> (c0, c1, c2, det_log(c2), 1 / c2) for every (t, θ) of a guided launch
__global__ void proposal_prepare_kernel(const double* __restrict__ raw, double* __restrict__ out, int64_t n) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double c0 = raw[3 * j], c1 = raw[3 * j + 1], c2 = raw[3 * j + 2];
  double* q = out + kProposalStride * j;
  q[0] = c0;
  q[1] = c1;
  q[2] = c2;
  q[3] = det_log(c2);
  q[4] = 1.0 / c2;
}

}  // namespace

// Steps per chunk of the dynamically scheduled launch, 0 = one CTA per θ for the whole series (host-only; smcb_batch_chunk_plan).
// Whole series, one CTA per θ, waste two ways (measured: profiles/r2_chunk_*.jsonl).  More θ than resident CTAs: ceil(M / slots)
// waves for M / slots waves of work (UCSV 4096: 512 θ on 148 slots, SV 2048: 1024 θ on 296 slots — 3.46 waves run as 4).
// All resident but unevenly spread over SMs whose warps are saturated: the SMs that hold ceil(M / SMs) CTAs finish last
// (LG1D 1024: 512 CTAs of 256 threads on 148 SMs).  Chunks pay when either wastes more than 3 %; a chunk boundary costs
// 5-7 µs (publish, acquire, reload), so chunks are at least 8 steps of a cloud of >= 1024 particles.
// masked: the launch carries an `active` mask (the proposals of rejuvenate! that fell outside the prior's support do not run,
// smc_samplers.jl:116): how many of the M CTAs do any work is only known on the device, so the waves of a one-CTA-per-θ launch are
// ceil(M_active / slots) for an M_active nobody planned for (config 5 on 4 GPUs lost 7 % that way) — such sweeps are chunked
// whenever there are more θ than resident CTAs.
int64_t plan_batch_chunk(int64_t M, int64_t N, int64_t steps, int threads, int64_t slots, int num_sms, bool masked) {
  if (M <= num_sms || slots < 1 || num_sms < 1) return 0;
  const double waves = (double)M / (double)slots, load = (double)M / (double)num_sms;
  double eff = 1.0;
  if (M > slots) eff = masked ? 0.0 : waves / std::ceil(waves);
  else if ((int64_t)threads * (int64_t)std::ceil(load) >= 1024) eff = load / std::ceil(load);
  const int64_t min_chunk = std::max<int64_t>(8, 8192 / std::max<int64_t>(N, 1));
  if (!(eff < 0.97) || steps < 2 * min_chunk) return 0;
  const int64_t want = (33 * slots + M - 1) / M;              // chunks per θ for >= 33 waves of units
  const int64_t nch = std::max<int64_t>(1, std::min<int64_t>(want, steps / min_chunk));
  const int64_t chunk = (steps + nch - 1) / nch;
  return chunk >= steps ? 0 : chunk;
}

namespace {

struct DynPlan {   // host side of the dynamic scheduling
  unsigned* sched;  // [1 + M]
  int num_sms;
  int force_chunk;  // SMCB_BATCH_CHUNK: -1 automatic, 0 never, K > 0 chunks of K steps
};

template <class Model, int PAIRS, int MAXT, bool GUIDED>
void launch_batch(BatchArgs a, int64_t M, int threads, size_t smem, cudaStream_t stream, const DynPlan& dp) {
  auto kern = batch_kernel<Model, PAIRS, MAXT, GUIDED, false>;
  auto kdyn = batch_kernel<Model, PAIRS, MAXT, GUIDED, true>;
  const int64_t steps = (int64_t)a.t_end - (int64_t)a.t_begin + 1;
  int64_t chunk = 0, slots = 0;
  if (dp.force_chunk != 0 && steps >= 2 && dp.sched) {
    SMCB_CUDA_TRY(cudaFuncSetAttribute(kdyn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SMCB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kdyn, threads, smem));
    slots = (int64_t)std::max(occ, 1) * dp.num_sms;
    if (dp.force_chunk > 0) chunk = dp.force_chunk;
    else chunk = plan_batch_chunk(M, a.N, steps, threads, slots, dp.num_sms, a.active != nullptr);
    if (chunk >= steps) chunk = 0;
  }
  if (chunk > 0) {
    a.sched = dp.sched;
    a.chunk = (uint32_t)chunk;
    a.nchunks = (uint32_t)((steps + chunk - 1) / chunk);
    SMCB_CUDA_TRY(cudaMemsetAsync(dp.sched, 0, sizeof(unsigned) * (size_t)(M + 1), stream));
    kdyn<<<(unsigned)std::min<int64_t>(M, slots), threads, smem, stream>>>(a);
  } else {
    SMCB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)M, threads, smem, stream>>>(a);
  }
  SMCB_CUDA_TRY(cudaGetLastError());
}

template <class Model, bool GUIDED = false>
void launch_batch_model(const BatchArgs& a, int64_t M, int pairs, int threads, size_t smem, cudaStream_t stream, const DynPlan& dp) {
  if (pairs == 2) {
    if (threads <= 256) launch_batch<Model, 2, 256, GUIDED>(a, M, threads, smem, stream, dp);
    else if (threads <= 512) launch_batch<Model, 2, 512, GUIDED>(a, M, threads, smem, stream, dp);
    else launch_batch<Model, 2, 1024, GUIDED>(a, M, threads, smem, stream, dp);
  } else if (pairs == 1) {
    if (threads <= 512) launch_batch<Model, 1, 512, GUIDED>(a, M, threads, smem, stream, dp);
    else launch_batch<Model, 1, 1024, GUIDED>(a, M, threads, smem, stream, dp);
  } else {
    if (threads <= 512) launch_batch<Model, 4, 512, GUIDED>(a, M, threads, smem, stream, dp);
    else launch_batch<Model, 4, 1024, GUIDED>(a, M, threads, smem, stream, dp);
  }
}

constexpr size_t kSmemBudget = 227 * 1024 - 2048;  // dynamic part; static scratch is < 2 KB

}  // namespace

// ------------------------------------------------------------------------------------------------
BatchFilter::BatchFilter(int device, cudaStream_t stream, int kind, int64_t M, int64_t N)
    : device_(device), stream_(stream), kind_(kind), d_(0), M_(M), N_(N) {
  if (is_mv_kind(kind)) throw Error{SMCB_ERR_UNSUPPORTED, "multivariate linear models run on the single filter (smcb_bootstrap_* / smcb_log_likelihood), not on the batched engine"};
  if (kind < 0 || kind >= KIND_COUNT) throw Error{SMCB_ERR_BAD_ARG, "unknown model kind"};
  if (M < 1 || M > (1 << 24)) throw Error{SMCB_ERR_BAD_ARG, "M must be in [1, 2^24]"};
  if (N < 1 || N > 8192) throw Error{SMCB_ERR_UNSUPPORTED, "batched filters support N in [1, 8192] (use the single filter above)"};
  d_ = state_dim(kind);
  ld_ = (N + 31) & ~int64_t(31);
  S_ = quant_shift((uint64_t)N);
  R_ = strata_width((uint64_t)N);
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  for (int i = 0; i < 2; ++i) {
    SMCB_CUDA_TRY(cudaMalloc(&x_[i], sizeof(double) * M * d_ * ld_));
    SMCB_CUDA_TRY(cudaMalloc(&logw_[i], sizeof(double) * M * ld_));
    SMCB_CUDA_TRY(cudaMalloc(&stats_[i], sizeof(StepStats) * M));
    SMCB_CUDA_TRY(cudaMemsetAsync(x_[i], 0, sizeof(double) * M * d_ * ld_, stream_));
    SMCB_CUDA_TRY(cudaMemsetAsync(logw_[i], 0, sizeof(double) * M * ld_, stream_));
    SMCB_CUDA_TRY(cudaMemsetAsync(stats_[i], 0, sizeof(StepStats) * M, stream_));
    SMCB_CUDA_TRY(cudaEventCreate(&ev_[i]));
  }
  SMCB_CUDA_TRY(cudaMalloc(&derived_, sizeof(double) * M * kParamStride));
  SMCB_CUDA_TRY(cudaMalloc(&active_, M));
  SMCB_CUDA_TRY(cudaMalloc(&out_dev_, sizeof(double) * 2 * M));
  SMCB_CUDA_TRY(cudaMalloc(&sched_, sizeof(unsigned) * (size_t)(M + 1)));
  host_tmp_.resize((size_t)(M * kParamStride));
}

BatchFilter::~BatchFilter() {
  cudaSetDevice(device_);
  for (int i = 0; i < 2; ++i) {
    cudaFree(x_[i]); cudaFree(logw_[i]); cudaFree(stats_[i]);
    if (ev_[i]) cudaEventDestroy(ev_[i]);
  }
  cudaFree(derived_); cudaFree(active_); cudaFree(y_dev_); cudaFree(out_dev_); cudaFree(w_tmp_); cudaFree(slots_dev_);
  cudaFree(prop_dev_);
  cudaFree(sched_);
}

void BatchFilter::begin_call() {
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  SMCB_CUDA_TRY(cudaEventRecord(ev_[0], stream_));
}

void BatchFilter::end_call() {
  SMCB_CUDA_TRY(cudaEventRecord(ev_[1], stream_));
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
  float ms = 0;
  SMCB_CUDA_TRY(cudaEventElapsedTime(&ms, ev_[0], ev_[1]));
  last_ms_ = ms;
}

void BatchFilter::upload_params(const double* params, const uint8_t* active) {
  if (params) {
    for (int64_t m = 0; m < M_; ++m) derive_params(kind_, params + m * kParamStride, host_tmp_.data() + m * kParamStride);
    SMCB_CUDA_TRY(cudaMemcpyAsync(derived_, host_tmp_.data(), sizeof(double) * M_ * kParamStride, cudaMemcpyHostToDevice, stream_));
    SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));  // host_tmp_ is reused
    has_params_ = true;
  }
  use_active_ = active != nullptr;
  if (active) SMCB_CUDA_TRY(cudaMemcpyAsync(active_, active, M_, cudaMemcpyHostToDevice, stream_));
}

// proposal: [rows][M][3] host coefficients (c0, c1, c2) of x' ~ N(c0 + c1 xp, c2^2); the device block is [rows][M][5] with
// det_log(c2) and 1 / c2 appended by proposal_prepare_kernel (SPEC §10; the device det_log is the host's bit for bit)
void BatchFilter::upload_proposal(const double* proposal, int64_t rows) {
  const int64_t n = rows * M_;
  for (int64_t j = 0; j < n; ++j) {
    const double c2 = proposal[3 * j + 2];
    if (!(c2 > 0.0) || !std::isfinite(c2) || !std::isfinite(proposal[3 * j]) || !std::isfinite(proposal[3 * j + 1]))
      throw Error{SMCB_ERR_BAD_ARG, "proposal: coefficients must be finite and the standard deviation c2 > 0"};
    if (kind_ == KIND_UCSV && !(proposal[3 * j] >= 0.0 && proposal[3 * j] <= 1.0))
      throw Error{SMCB_ERR_BAD_ARG, "proposal (UCSV, docs/SPEC.md §10b): the triple is (kappa, 0, 1) with kappa in [0, 1]"};
  }
  if (prop_cap_ < n) {
    cudaFree(prop_dev_);
    prop_dev_ = nullptr;
    prop_cap_ = 0;
    SMCB_CUDA_TRY(cudaMalloc(&prop_dev_, sizeof(double) * (kProposalStride + 3) * n));  // derived block, then the raw copy
    prop_cap_ = n;
  }
  double* raw = prop_dev_ + kProposalStride * prop_cap_;
  SMCB_CUDA_TRY(cudaMemcpyAsync(raw, proposal, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, stream_));
  proposal_prepare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream_>>>(raw, prop_dev_, n);
  SMCB_CUDA_TRY(cudaGetLastError());
  ++launches_;
}

void BatchFilter::launch(const IO& io, bool from_init, uint32_t t_begin, uint32_t t_end, int resampler, bool guided) {
  const int64_t npairs = (N_ + 1) / 2;
  int pairs = 2;
  if ((npairs + pairs - 1) / pairs > 1024) pairs = 4;
  // Few θ-particles on this GPU (θ sharded over many GPUs: 512 θ / 8 GPUs = 64 CTAs on 148 SMs): every CTA has an SM to itself
  // and the sweep time is the per-step dependent chain of ONE cloud — one pair per thread shortens it (N = 1024, T = 100:
  // 4.25 against 5.77 µs per step for M <= 128, 6.7 against 7.1 at M = 256; from M = 512 on two pairs per thread win on
  // throughput: profiles/r2_batch_occupancy_lg1024.jsonl).
  if (num_sms_ == 0) {
    int v = 0;
    SMCB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device_));
    num_sms_ = v > 0 ? v : 148;
  }
  if (pairs == 2 && npairs <= 1024 && M_ <= 2 * (int64_t)num_sms_) pairs = 1;
  if (const char* e = std::getenv("SMCB_BATCH_PAIRS")) {
    const int v = std::atoi(e);
    if ((v == 1 || v == 2 || v == 4) && (npairs + v - 1) / v <= 1024) pairs = v;
  }
  int threads = (int)(((npairs + pairs - 1) / pairs + 31) / 32 * 32);
  const int64_t npad = (N_ + 1) & ~int64_t(1);
  const size_t cdf_bytes = sizeof(uint64_t) * (size_t)npad;
  const size_t x_bytes = sizeof(double) * (size_t)npad * d_;
  bool x_in_smem = cdf_bytes + x_bytes <= kSmemBudget;
  // Placement of the clouds, by measurement (profiles/r1_c5_clouds_in_l2_v22.jsonl, r1_batch_cloud_placement_probe_v22.jsonl):
  // a 3-component cloud of more than ~3600 particles (UCSV, config 5: 4096 θ × 4096) runs 10 % faster when only the CDF
  // is in shared memory and the cloud stays in global memory (the resident CTAs' clouds live in L2): 294 -> 266 ms of device
  // time at T = 100, results bit-identical.  One-component clouds of 8192 particles show no difference (5.96 ms either way)
  // and keep the shared-memory placement.  The cause was not profiled; the CTAs per SM do not change (1024 threads at 64
  // registers fill the register file either way), so the candidates are the three strided component gathers hitting the
  // same banks and the larger L1 carve-out.  SMCB_BATCH_X_SMEM=0/1 forces either placement.
  constexpr size_t kHalfSm = 113 * 1024;
  if (x_in_smem && d_ == 3 && cdf_bytes + x_bytes > kHalfSm && cdf_bytes <= kHalfSm && threads <= 1024) x_in_smem = false;
  if (const char* e = std::getenv("SMCB_BATCH_X_SMEM")) {
    if (std::atoi(e) == 0) x_in_smem = false;
    else x_in_smem = cdf_bytes + x_bytes <= kSmemBudget;
  }
  const size_t smem = cdf_bytes + (x_in_smem ? x_bytes : 0);
  if (smem > kSmemBudget) throw Error{SMCB_ERR_UNSUPPORTED, "CDF does not fit in shared memory"};
  BatchArgs a;
  a.derived = io.derived;
  a.active = io.active;
  a.y = io.y;
  a.prop = guided ? prop_dev_ : nullptr;
  a.x = x_[cur_];
  a.logw = logw_[cur_];
  a.stats = stats_[cur_];
  a.logz_out = io.logz_out;
  a.ess_out = io.ess_out ? io.ess_out : out_dev_ + M_;
  a.N = N_; a.ld = ld_; a.S = S_; a.R = R_;
  a.key = key_; a.stream0 = stream0_;
  a.t_begin = t_begin; a.t_end = t_end;
  a.resampler = resampler;
  a.from_init = from_init ? 1 : 0;
  a.x_in_smem = x_in_smem ? 1 : 0;
  a.npad = npad;
  a.M = M_;
  a.sched = nullptr;
  a.chunk = 0;
  a.nchunks = 1;
  DynPlan dp{sched_, num_sms_, -1};
  if (const char* e = std::getenv("SMCB_BATCH_CHUNK")) dp.force_chunk = std::atoi(e);
  if (guided) {
    if (kind_ == KIND_LG1D) launch_batch_model<ModelLG1D, true>(a, M_, pairs, threads, smem, stream_, dp);
    else if (kind_ == KIND_SV) launch_batch_model<ModelSV, true>(a, M_, pairs, threads, smem, stream_, dp);
    else launch_batch_model<ModelUCSV, true>(a, M_, pairs, threads, smem, stream_, dp);
    ++launches_;
    return;
  }
  switch (kind_) {
    case KIND_LG1D: launch_batch_model<ModelLG1D>(a, M_, pairs, threads, smem, stream_, dp); break;
    case KIND_SV: launch_batch_model<ModelSV>(a, M_, pairs, threads, smem, stream_, dp); break;
    default: launch_batch_model<ModelUCSV>(a, M_, pairs, threads, smem, stream_, dp); break;
  }
  ++launches_;
}

static void ensure_y(double*& y_dev, int64_t& cap, int64_t T) {
  if (cap < T) {
    cudaFree(y_dev);
    y_dev = nullptr;
    cap = 0;
    SMCB_CUDA_TRY(cudaMalloc(&y_dev, sizeof(double) * T));
    cap = T;
  }
}

void BatchFilter::init(const double* params, const uint8_t* active, double y0, const RngKey& key, uint32_t stream0,
                       double* logmu, double* ess) {
  begin_call();
  upload_params(params, active);
  ensure_y(y_dev_, y_cap_, 1);
  SMCB_CUDA_TRY(cudaMemcpyAsync(y_dev_, &y0, sizeof(double), cudaMemcpyHostToDevice, stream_));
  key_ = key; stream0_ = stream0; t_ = 0;
  launch(own_io(), true, 0, 0, RESAMPLE_SYSTEMATIC);
  if (logmu) SMCB_CUDA_TRY(cudaMemcpyAsync(logmu, out_dev_, sizeof(double) * M_, cudaMemcpyDeviceToHost, stream_));
  if (ess) SMCB_CUDA_TRY(cudaMemcpyAsync(ess, out_dev_ + M_, sizeof(double) * M_, cudaMemcpyDeviceToHost, stream_));
  end_call();
  live_ = true;
}

void BatchFilter::step(const double* params, double y, int resampler, double* logmu, double* ess, const double* proposal) {
  if (!live_) throw Error{SMCB_ERR_STATE, "batch_step before batch_init / batch_log_likelihood"};
  if (resampler < 0 || resampler > 2) throw Error{SMCB_ERR_BAD_ARG, "unknown resampler"};
  begin_call();
  upload_params(params, nullptr);
  ensure_y(y_dev_, y_cap_, 1);
  SMCB_CUDA_TRY(cudaMemcpyAsync(y_dev_, &y, sizeof(double), cudaMemcpyHostToDevice, stream_));
  if (proposal) upload_proposal(proposal, 1);
  launch(own_io(), false, t_ + 1, t_ + 1, resampler, proposal != nullptr);
  t_ += 1;
  if (logmu) SMCB_CUDA_TRY(cudaMemcpyAsync(logmu, out_dev_, sizeof(double) * M_, cudaMemcpyDeviceToHost, stream_));
  if (ess) SMCB_CUDA_TRY(cudaMemcpyAsync(ess, out_dev_ + M_, sizeof(double) * M_, cudaMemcpyDeviceToHost, stream_));
  end_call();
}

void BatchFilter::run(const double* params, const uint8_t* active, const double* y, int64_t T, int resampler,
                      const RngKey& key, uint32_t stream0, double* logZ, const double* proposal) {
  if (resampler < 0 || resampler > 2) throw Error{SMCB_ERR_BAD_ARG, "unknown resampler"};
  begin_call();
  upload_params(params, active);
  ensure_y(y_dev_, y_cap_, T);
  SMCB_CUDA_TRY(cudaMemcpyAsync(y_dev_, y, sizeof(double) * T, cudaMemcpyHostToDevice, stream_));
  key_ = key; stream0_ = stream0;
  if (proposal) upload_proposal(proposal, T);  // row 0 (the initial draw is the bootstrap one) is not read
  launch(own_io(), true, 0, (uint32_t)(T - 1), resampler, proposal != nullptr);
  t_ = (uint32_t)(T - 1);
  SMCB_CUDA_TRY(cudaMemcpyAsync(logZ, out_dev_, sizeof(double) * M_, cudaMemcpyDeviceToHost, stream_));
  end_call();
  live_ = true;
}

void BatchFilter::upload_slots(const int32_t* a, const int32_t* b, int64_t n) {
  if (slot_cap_ < n) {
    cudaFree(slots_dev_);
    slots_dev_ = nullptr;
    slot_cap_ = 0;
    SMCB_CUDA_TRY(cudaMalloc(&slots_dev_, sizeof(int32_t) * 2 * n));
    slot_cap_ = n;
  }
  if (a) SMCB_CUDA_TRY(cudaMemcpyAsync(slots_dev_, a, sizeof(int32_t) * n, cudaMemcpyHostToDevice, stream_));
  if (b) SMCB_CUDA_TRY(cudaMemcpyAsync(slots_dev_ + slot_cap_, b, sizeof(int32_t) * n, cudaMemcpyHostToDevice, stream_));
}

void BatchFilter::gather(const int32_t* parents) {
  if (!live_) throw Error{SMCB_ERR_STATE, "batch_gather before any filtering"};
  for (int64_t m = 0; m < M_; ++m)
    if (parents[m] < 0 || parents[m] >= M_) throw Error{SMCB_ERR_BAD_ARG, "batch_gather: parent index out of range"};
  begin_call();
  upload_slots(parents, nullptr, M_);
  const int64_t xrow = d_ * ld_, wrow = ld_;
  dim3 grid((unsigned)M_, (unsigned)std::max<int64_t>(1, std::min<int64_t>(8, xrow / 2 / 256)));
  copy_clouds_kernel<<<grid, 256, 0, stream_>>>(x_[cur_ ^ 1], x_[cur_], logw_[cur_ ^ 1], logw_[cur_], stats_[cur_ ^ 1],
                                                stats_[cur_], nullptr, slots_dev_, nullptr, xrow, wrow);
  SMCB_CUDA_TRY(cudaGetLastError());
  ++launches_;
  cur_ ^= 1;
  end_call();
}

void BatchFilter::accept_from(const BatchFilter& prop, const uint8_t* accept) {
  if (prop.M_ != M_ || prop.N_ != N_ || prop.kind_ != kind_ || prop.device_ != device_)
    throw Error{SMCB_ERR_BAD_ARG, "batch_accept: batches differ in shape, model or device"};
  if (!prop.live_) throw Error{SMCB_ERR_STATE, "batch_accept: proposal batch holds no clouds"};
  begin_call();
  SMCB_CUDA_TRY(cudaMemcpyAsync(active_, accept, M_, cudaMemcpyHostToDevice, stream_));
  use_active_ = false;
  const int64_t xrow = d_ * ld_, wrow = ld_;
  dim3 grid((unsigned)M_, (unsigned)std::max<int64_t>(1, std::min<int64_t>(8, xrow / 2 / 256)));
  copy_clouds_kernel<<<grid, 256, 0, stream_>>>(x_[cur_], prop.x_[prop.cur_], logw_[cur_], prop.logw_[prop.cur_], stats_[cur_],
                                                prop.stats_[prop.cur_], nullptr, nullptr, active_, xrow, wrow);
  SMCB_CUDA_TRY(cudaGetLastError());
  ++launches_;
  if (!live_) { t_ = prop.t_; key_ = prop.key_; stream0_ = prop.stream0_; live_ = true; }
  end_call();
}

void BatchFilter::fetch(double* x_host, double* w_host, double* logw_host) {
  if (!live_) throw Error{SMCB_ERR_STATE, "no batch state to fetch"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  if (x_host)
    SMCB_CUDA_TRY(cudaMemcpy2DAsync(x_host, sizeof(double) * N_, x_[cur_], sizeof(double) * ld_, sizeof(double) * N_, M_ * d_,
                                    cudaMemcpyDeviceToHost, stream_));
  if (logw_host)
    SMCB_CUDA_TRY(cudaMemcpy2DAsync(logw_host, sizeof(double) * N_, logw_[cur_], sizeof(double) * ld_, sizeof(double) * N_, M_,
                                    cudaMemcpyDeviceToHost, stream_));
  if (w_host) {
    ensure_scratch((size_t)(M_ * std::max<int64_t>(N_, d_)));
    dim3 grid((unsigned)((N_ + 255) / 256), (unsigned)M_);
    batch_weights_kernel<<<grid, 256, 0, stream_>>>(logw_[cur_], stats_[cur_], w_tmp_, N_, ld_);
    SMCB_CUDA_TRY(cudaGetLastError());
    SMCB_CUDA_TRY(cudaMemcpyAsync(w_host, w_tmp_, sizeof(double) * M_ * N_, cudaMemcpyDeviceToHost, stream_));
  }
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
}

void BatchFilter::weighted_mean(double* mean_host) {
  if (!live_) throw Error{SMCB_ERR_STATE, "no batch state to summarise"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  ensure_scratch((size_t)(M_ * d_));
  dim3 grid((unsigned)M_, (unsigned)d_);
  batch_wmean_kernel<<<grid, 256, 0, stream_>>>(x_[cur_], logw_[cur_], stats_[cur_], w_tmp_, N_, ld_, d_);
  SMCB_CUDA_TRY(cudaGetLastError());
  SMCB_CUDA_TRY(cudaMemcpyAsync(mean_host, w_tmp_, sizeof(double) * M_ * d_, cudaMemcpyDeviceToHost, stream_));
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
}

void BatchFilter::weighted_moments(double* mean_host, double* var_host) {
  if (!live_) throw Error{SMCB_ERR_STATE, "no batch state to summarise"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  const size_t words = (size_t)M_ * d_;
  ensure_scratch(2 * words);  // mean[M][d] then var[M][d]; the scratch is kept across calls (no cudaMalloc / cudaFree per call)
  double* scratch = w_tmp_;
  dim3 grid((unsigned)M_, (unsigned)d_);
  batch_wmoment_kernel<<<grid, 256, 0, stream_>>>(x_[cur_], logw_[cur_], stats_[cur_], scratch, scratch + words, N_, ld_, d_);
  SMCB_CUDA_TRY(cudaGetLastError());
  if (mean_host) SMCB_CUDA_TRY(cudaMemcpyAsync(mean_host, scratch, sizeof(double) * words, cudaMemcpyDeviceToHost, stream_));
  if (var_host) SMCB_CUDA_TRY(cudaMemcpyAsync(var_host, scratch + words, sizeof(double) * words, cudaMemcpyDeviceToHost, stream_));
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
}

void BatchFilter::ensure_scratch(size_t words) {
  if (w_cap_ >= words) return;
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
  cudaFree(w_tmp_);
  w_tmp_ = nullptr;
  w_cap_ = 0;
  SMCB_CUDA_TRY(cudaMalloc(&w_tmp_, sizeof(double) * words));
  w_cap_ = words;
}

void BatchFilter::weighted_quantiles(const double* probs, int np, bool weighted, double* q_host) {
  if (!live_) throw Error{SMCB_ERR_STATE, "no batch state to summarise"};
  if (np < 1 || np > kBatchMaxProbs || !probs || !q_host) throw Error{SMCB_ERR_BAD_ARG, "batch_weighted_quantiles: 1..16 probabilities"};
  for (int j = 0; j < np; ++j)
    if (!(probs[j] >= 0.0 && probs[j] <= 1.0)) throw Error{SMCB_ERR_BAD_ARG, "batch_weighted_quantiles: probabilities must lie in [0, 1]"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  const size_t words = (size_t)np + (size_t)M_ * d_ * np;  // probs[np] then out[M][d][np]
  ensure_scratch(words);
  double* scratch = w_tmp_;
  SMCB_CUDA_TRY(cudaMemcpyAsync(scratch, probs, sizeof(double) * np, cudaMemcpyHostToDevice, stream_));
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));  // the caller's probs buffer is not retained
  dim3 grid((unsigned)M_, (unsigned)d_);
  batch_wquantile_kernel<<<grid, 256, 0, stream_>>>(x_[cur_], logw_[cur_], stats_[cur_], scratch, np, weighted ? 1 : 0, S_, scratch + np, N_, ld_, d_);
  SMCB_CUDA_TRY(cudaGetLastError());
  SMCB_CUDA_TRY(cudaMemcpyAsync(q_host, scratch + np, sizeof(double) * M_ * d_ * np, cudaMemcpyDeviceToHost, stream_));
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream_));
}

// ---- device-resident variants: enqueue only (no host copies, no synchronisation) --------------------------------
void BatchFilter::init_dev(const double* derived_dev, const uint8_t* active_dev, const double* y0_dev, const RngKey& key,
                           uint32_t stream0, double* logmu_dev, double* ess_dev) {
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  key_ = key; stream0_ = stream0; t_ = 0;
  launch(IO{derived_dev, active_dev, y0_dev, logmu_dev, ess_dev}, true, 0, 0, RESAMPLE_SYSTEMATIC);
  live_ = true;
}

void BatchFilter::step_dev(const double* derived_dev, const double* y_dev, uint32_t t, int resampler, double* logmu_dev,
                           double* ess_dev) {
  if (!live_) throw Error{SMCB_ERR_STATE, "batch step before init / log_likelihood"};
  if (resampler < 0 || resampler > 2) throw Error{SMCB_ERR_BAD_ARG, "unknown resampler"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  launch(IO{derived_dev, nullptr, y_dev, logmu_dev, ess_dev}, false, t, t, resampler);
  t_ = t;
}

void BatchFilter::run_dev(const double* derived_dev, const uint8_t* active_dev, const double* y_dev, int64_t T, int resampler,
                          const RngKey& key, uint32_t stream0, double* logz_dev) {
  if (resampler < 0 || resampler > 2) throw Error{SMCB_ERR_BAD_ARG, "unknown resampler"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  key_ = key; stream0_ = stream0;
  launch(IO{derived_dev, active_dev, y_dev, logz_dev, nullptr}, true, 0, (uint32_t)(T - 1), resampler);
  t_ = (uint32_t)(T - 1);
  live_ = true;
}

void BatchFilter::gather_dev(const int32_t* parents_dev) {
  if (!live_) throw Error{SMCB_ERR_STATE, "batch_gather before any filtering"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  const int64_t xrow = d_ * ld_, wrow = ld_;
  dim3 grid((unsigned)M_, (unsigned)std::max<int64_t>(1, std::min<int64_t>(8, xrow / 2 / 256)));
  copy_clouds_kernel<<<grid, 256, 0, stream_>>>(x_[cur_ ^ 1], x_[cur_], logw_[cur_ ^ 1], logw_[cur_], stats_[cur_ ^ 1],
                                                stats_[cur_], nullptr, parents_dev, nullptr, xrow, wrow);
  SMCB_CUDA_TRY(cudaGetLastError());
  ++launches_;
  cur_ ^= 1;
}

void BatchFilter::accept_dev(const BatchFilter& prop, const uint8_t* mask_dev) {
  if (prop.M_ != M_ || prop.N_ != N_ || prop.kind_ != kind_ || prop.device_ != device_)
    throw Error{SMCB_ERR_BAD_ARG, "batch_accept: batches differ in shape, model or device"};
  if (!prop.live_) throw Error{SMCB_ERR_STATE, "batch_accept: proposal batch holds no clouds"};
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  const int64_t xrow = d_ * ld_, wrow = ld_;
  dim3 grid((unsigned)M_, (unsigned)std::max<int64_t>(1, std::min<int64_t>(8, xrow / 2 / 256)));
  copy_clouds_kernel<<<grid, 256, 0, stream_>>>(x_[cur_], prop.x_[prop.cur_], logw_[cur_], prop.logw_[prop.cur_], stats_[cur_],
                                                prop.stats_[prop.cur_], nullptr, nullptr, mask_dev, xrow, wrow);
  SMCB_CUDA_TRY(cudaGetLastError());
  ++launches_;
  if (!live_) { t_ = prop.t_; key_ = prop.key_; stream0_ = prop.stream0_; live_ = true; }
}

void BatchFilter::pack_dev(const int32_t* slots_dev, int64_t n, void* buf_dev, bool to_buffer) {
  if (n == 0) return;
  SMCB_CUDA_TRY(cudaSetDevice(device_));
  const int64_t xrow = d_ * ld_, wrow = ld_;
  dim3 grid((unsigned)n, (unsigned)std::max<int64_t>(1, std::min<int64_t>(8, xrow / 256)));
  pack_clouds_kernel<<<grid, 256, 0, stream_>>>(x_[cur_], logw_[cur_], stats_[cur_], slots_dev, static_cast<double*>(buf_dev),
                                                xrow, wrow, to_buffer ? 1 : 0);
  SMCB_CUDA_TRY(cudaGetLastError());
  ++launches_;
  if (!to_buffer) live_ = true;
}

int64_t BatchFilter::cloud_bytes() const { return (int64_t)sizeof(double) * ((d_ + 1) * ld_ + 4); }

void BatchFilter::pack(const int32_t* slots, int64_t n, void* buf_dev, bool to_buffer) {
  if (n == 0) return;
  for (int64_t j = 0; j < n; ++j)
    if (slots[j] < 0 || slots[j] >= M_) throw Error{SMCB_ERR_BAD_ARG, "batch_pack: slot out of range"};
  begin_call();
  upload_slots(slots, nullptr, n);
  const int64_t xrow = d_ * ld_, wrow = ld_;
  dim3 grid((unsigned)n, (unsigned)std::max<int64_t>(1, std::min<int64_t>(8, xrow / 256)));
  pack_clouds_kernel<<<grid, 256, 0, stream_>>>(x_[cur_], logw_[cur_], stats_[cur_], slots_dev_, static_cast<double*>(buf_dev),
                                                xrow, wrow, to_buffer ? 1 : 0);
  SMCB_CUDA_TRY(cudaGetLastError());
  ++launches_;
  if (!to_buffer) live_ = true;
  end_call();
}

// ------------------------------------------------------------------------------------------------
void* DeviceScratch::get(size_t bytes) {
  if (cap < bytes) {
    cudaFree(p);
    p = nullptr;
    cap = 0;
    SMCB_CUDA_TRY(cudaMalloc(&p, bytes));
    cap = bytes;
  }
  return p;
}
DeviceScratch::~DeviceScratch() { cudaFree(p); }

namespace {
// carve 256-byte aligned pieces out of one scratch allocation
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  static size_t up(size_t n) { return (n + 255) & ~size_t(255); }
  template <class T>
  T* take(size_t n) {
    T* r = reinterpret_cast<T*>(base + off);
    off += up(sizeof(T) * n);
    return r;
  }
};
}  // namespace

void kalman_batch(int device, cudaStream_t stream, DeviceScratch& scratch, const double* params, const uint8_t* active, int64_t M,
                  const double* y, int64_t T, bool predict_first, double* loglik, double* x, double* sigma,
                  bool use_state) {
  SMCB_CUDA_TRY(cudaSetDevice(device));
  const size_t total = Carver::up(sizeof(double) * M * kParamStride) + Carver::up(sizeof(double) * T) + 3 * Carver::up(sizeof(double) * M) + Carver::up((size_t)M);
  Carver c(scratch.get(total));
  double* dp = c.take<double>((size_t)M * kParamStride);
  double* dy = c.take<double>((size_t)T);
  double* dl = c.take<double>((size_t)M);
  double* dx = c.take<double>((size_t)M);
  double* ds = c.take<double>((size_t)M);
  uint8_t* da = active ? c.take<uint8_t>((size_t)M) : nullptr;
  SMCB_CUDA_TRY(cudaMemcpyAsync(dp, params, sizeof(double) * M * kParamStride, cudaMemcpyHostToDevice, stream));
  SMCB_CUDA_TRY(cudaMemcpyAsync(dy, y, sizeof(double) * T, cudaMemcpyHostToDevice, stream));
  if (active) SMCB_CUDA_TRY(cudaMemcpyAsync(da, active, M, cudaMemcpyHostToDevice, stream));
  if (use_state) {
    SMCB_CUDA_TRY(cudaMemcpyAsync(dx, x, sizeof(double) * M, cudaMemcpyHostToDevice, stream));
    SMCB_CUDA_TRY(cudaMemcpyAsync(ds, sigma, sizeof(double) * M, cudaMemcpyHostToDevice, stream));
  }
  kalman_kernel<<<(unsigned)((M + 127) / 128), 128, 0, stream>>>(dp, da, M, dy, T, predict_first ? 1 : 0, dl, dx, ds, use_state ? 1 : 0);
  SMCB_CUDA_TRY(cudaGetLastError());
  SMCB_CUDA_TRY(cudaMemcpyAsync(loglik, dl, sizeof(double) * M, cudaMemcpyDeviceToHost, stream));
  if (x) SMCB_CUDA_TRY(cudaMemcpyAsync(x, dx, sizeof(double) * M, cudaMemcpyDeviceToHost, stream));
  if (sigma) SMCB_CUDA_TRY(cudaMemcpyAsync(sigma, ds, sizeof(double) * M, cudaMemcpyDeviceToHost, stream));
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream));
}

void kalman_mv_batch(int device, cudaStream_t stream, DeviceScratch& scratch, int d, const double* models, const uint8_t* active, int64_t M,
                     const double* y, int64_t T, bool predict_first, double* loglik, double* x, double* sigma,
                     bool use_state) {
  if (d < 1 || d > 4) throw Error{SMCB_ERR_UNSUPPORTED, "matrix Kalman filter: state dimension must be in [1, 4]"};
  SMCB_CUDA_TRY(cudaSetDevice(device));
  const int64_t stride = 3 * d * d + 2 * d + 1;
  const size_t total = Carver::up(sizeof(double) * M * stride) + Carver::up(sizeof(double) * T) + Carver::up(sizeof(double) * M) +
                       Carver::up(sizeof(double) * M * d) + Carver::up(sizeof(double) * M * d * d) + Carver::up((size_t)M);
  Carver c(scratch.get(total));
  double* dp = c.take<double>((size_t)(M * stride));
  double* dy = c.take<double>((size_t)T);
  double* dl = c.take<double>((size_t)M);
  double* dx = c.take<double>((size_t)(M * d));
  double* ds = c.take<double>((size_t)(M * d * d));
  uint8_t* da = active ? c.take<uint8_t>((size_t)M) : nullptr;
  SMCB_CUDA_TRY(cudaMemcpyAsync(dp, models, sizeof(double) * M * stride, cudaMemcpyHostToDevice, stream));
  SMCB_CUDA_TRY(cudaMemcpyAsync(dy, y, sizeof(double) * T, cudaMemcpyHostToDevice, stream));
  if (active) SMCB_CUDA_TRY(cudaMemcpyAsync(da, active, M, cudaMemcpyHostToDevice, stream));
  if (use_state) {
    SMCB_CUDA_TRY(cudaMemcpyAsync(dx, x, sizeof(double) * M * d, cudaMemcpyHostToDevice, stream));
    SMCB_CUDA_TRY(cudaMemcpyAsync(ds, sigma, sizeof(double) * M * d * d, cudaMemcpyHostToDevice, stream));
  }
  const unsigned grid = (unsigned)((M + 127) / 128);
  const int pf = predict_first ? 1 : 0, us = use_state ? 1 : 0;
  switch (d) {
    case 1: kalman_mv_kernel<1><<<grid, 128, 0, stream>>>(dp, da, M, dy, T, pf, dl, dx, ds, us); break;
    case 2: kalman_mv_kernel<2><<<grid, 128, 0, stream>>>(dp, da, M, dy, T, pf, dl, dx, ds, us); break;
    case 3: kalman_mv_kernel<3><<<grid, 128, 0, stream>>>(dp, da, M, dy, T, pf, dl, dx, ds, us); break;
    default: kalman_mv_kernel<4><<<grid, 128, 0, stream>>>(dp, da, M, dy, T, pf, dl, dx, ds, us); break;
  }
  SMCB_CUDA_TRY(cudaGetLastError());
  SMCB_CUDA_TRY(cudaMemcpyAsync(loglik, dl, sizeof(double) * M, cudaMemcpyDeviceToHost, stream));
  if (x) SMCB_CUDA_TRY(cudaMemcpyAsync(x, dx, sizeof(double) * M * d, cudaMemcpyDeviceToHost, stream));
  if (sigma) SMCB_CUDA_TRY(cudaMemcpyAsync(sigma, ds, sizeof(double) * M * d * d, cudaMemcpyDeviceToHost, stream));
  SMCB_CUDA_TRY(cudaStreamSynchronize(stream));
}

}  // namespace smcb
