// Shared device helpers for the particle-filter kernels: ordered-double encoding for an exact,
// order-independent atomic max, warp/block reductions, cache-hinted loads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <utility>

#include "../../include/smcb200.h"  // smcb_status error codes
#include "smcb_detmath.cuh"

namespace smcb {

struct Error {
  int code;
  std::string msg;
};

#define SMCB_CUDA_TRY(expr)                                                                     \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      throw ::smcb::Error{_e == cudaErrorMemoryAllocation ? ::SMCB_ERR_OOM                \
                                                          : ::SMCB_ERR_CUDA,              \
                          std::string(#expr) + ": " + cudaGetErrorString(_e)};                  \
    }                                                                                           \
  } while (0)

#if defined(__CUDACC__)

constexpr unsigned kFullMask = 0xFFFFFFFFu;

// Programmatic dependent launch (sm_90+): the four kernels of a filter step form a strict chain, so
// every kernel lets its successor's CTAs become resident at once (they fill the SM slots freed by
// this grid's tail) and blocks until its predecessor has completed and flushed before it reads
// anything the predecessor wrote.  Hides launch latency and ramp-up at every kernel boundary.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// launch with the programmatic-stream-serialization attribute (the kernel must call pdl_wait())
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl_smem(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// monotone map double -> uint64 so that atomicMax on the image is max on the doubles
__device__ __forceinline__ unsigned long long encode_ordered(double d) {
  unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double decode_ordered(unsigned long long e) {
  unsigned long long b = (e >> 63) ? (e & 0x7FFFFFFFFFFFFFFFull) : ~e;
  return __longlong_as_double((long long)b);
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double t = __shfl_xor_sync(kFullMask, v, o);
    v = (t > v) ? t : v;
  }
  return v;
}
// max over the warp of ordered-encoded doubles with two 32-bit REDUX instead of five 64-bit shuffle rounds
__device__ __forceinline__ unsigned long long warp_max_ordered(unsigned long long e) {
  const unsigned hi = (unsigned)(e >> 32), lo = (unsigned)e;
  const unsigned mh = __reduce_max_sync(kFullMask, hi);
  const unsigned ml = __reduce_max_sync(kFullMask, hi == mh ? lo : 0u);
  return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
// inclusive scan across the lanes of a warp
__device__ __forceinline__ unsigned long long warp_scan_u64(unsigned long long v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

__device__ __forceinline__ double ld_cg(const double* p) { return __ldcg(p); }

// number of entries C[lo..hi) that are <= tau, plus lo  (ancestor index, SPEC §5)
template <class Ptr>
__device__ __forceinline__ int64_t lower_count(Ptr C, int64_t lo, int64_t hi, uint64_t tau) {
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (C[mid] <= tau) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

#endif  // __CUDACC__

}  // namespace smcb
