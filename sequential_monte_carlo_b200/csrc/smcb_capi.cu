// extern "C" boundary of libsmcb200 (include/smcb200.h).  No C++ types or exceptions cross it.
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../include/smcb200.h"
#include "smcb_batch.cuh"
#include "smcb_filter.cuh"
#include "smcb_sampler.cuh"
#include "smcb_detmathf.cuh"

using namespace smcb;

struct smcb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint64_t seed = 0;
  uint32_t next_epoch = 0;
  std::string err;
  std::unique_ptr<SingleFilter> filter;
  std::unique_ptr<SingleFilter> scratch;  // normalize / resample utilities
  std::vector<StepStats> stats;
  Comm comm;                              // θ-sharding across GPUs (smcb_comm_init); nranks == 1 without it
  DeviceScratch kalman_scratch;           // device staging of the Kalman entry points, kept across calls
  double* comm_buf = nullptr;             // device staging of smcb_comm_all_gather
  int64_t comm_buf_cap = 0;
  RngKey key(uint32_t epoch) const { return make_rng_key((uint32_t)seed, (uint32_t)(seed >> 32), epoch & 0xFFFFFFu); }
};

struct smcb_batch {
  smcb_ctx* ctx = nullptr;
  BatchFilter* impl = nullptr;
  bool owned = true;   // false: a view of a sampler's live clouds (smcb_sampler_clouds)
  ~smcb_batch() { if (owned) delete impl; }
};

struct smcb_sampler {
  smcb_ctx* ctx = nullptr;
  std::unique_ptr<ThetaSampler> impl;
  smcb_batch view;     // borrowed handle of impl->clouds()
};

namespace {

thread_local std::string g_create_error;

template <class F>
int guarded(smcb_ctx* ctx, F&& f) {
  try {
    f();
    return SMCB_OK;
  } catch (const Error& e) {
    if (ctx) ctx->err = e.msg; else g_create_error = e.msg;
    return e.code;
  } catch (const std::bad_alloc&) {
    if (ctx) ctx->err = "host allocation failed"; else g_create_error = "host allocation failed";
    return SMCB_ERR_OOM;
  } catch (const std::exception& e) {
    if (ctx) ctx->err = e.what(); else g_create_error = e.what();
    return SMCB_ERR_CUDA;
  }
}

inline void need(bool ok, const char* what) {
  if (!ok) throw Error{SMCB_ERR_BAD_ARG, what};
}

inline void stats_to(const StepStats& s, int64_t N, double* logmu, double* ess) {
  // normalize(): logμ = max + log Σe − log N ; ess = (Σe)² / Σe²      particles.jl:10-13
  if (logmu) *logmu = s.mx + std::log(s.sum) - std::log((double)N);
  if (ess) *ess = (s.sum * s.sum) / s.sum2;
}

}  // namespace

extern "C" {

int smcb_version(void) { return 100; }

int smcb_state_dim(int kind) { return (kind >= 0 && kind < KIND_ALL) ? state_dim(kind) : SMCB_ERR_BAD_ARG; }

int smcb_create(int device, uint64_t seed, smcb_ctx** out) {
  if (!out) return SMCB_ERR_BAD_ARG;
  *out = nullptr;
  return guarded(nullptr, [&] {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      throw Error{SMCB_ERR_CUDA, std::string("no CUDA device: libsmcb200 has no CPU fallback (") +
                                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") + ")"};
    need(device >= 0 && device < count, "device index out of range");
    SMCB_CUDA_TRY(cudaSetDevice(device));
    std::unique_ptr<smcb_ctx> c(new smcb_ctx);
    c->device = device;
    c->seed = seed;
    SMCB_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->filter.reset(new SingleFilter(device, c->stream));
    c->scratch.reset(new SingleFilter(device, c->stream));
    *out = c.release();
  });
}

int smcb_destroy(smcb_ctx* ctx) {
  if (!ctx) return SMCB_OK;
  cudaSetDevice(ctx->device);
  smcb_comm_destroy(ctx);
  ctx->filter.reset();
  ctx->scratch.reset();
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SMCB_OK;
}

const char* smcb_last_error(const smcb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int smcb_set_rng(smcb_ctx* ctx, uint64_t seed, uint32_t epoch) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  ctx->seed = seed;
  ctx->next_epoch = epoch & 0xFFFFFFu;
  return SMCB_OK;
}

int smcb_get_epoch(const smcb_ctx* ctx, uint32_t* next_epoch) {
  if (!ctx || !next_epoch) return SMCB_ERR_BAD_ARG;
  *next_epoch = ctx->next_epoch;
  return SMCB_OK;
}

int smcb_record_ancestors(smcb_ctx* ctx, int enable) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  ctx->filter->set_record_ancestors(enable != 0);
  return SMCB_OK;
}

int smcb_set_profiling(smcb_ctx* ctx, int enable) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  ctx->filter->set_profiling(enable != 0);
  return SMCB_OK;
}

int smcb_set_precision(smcb_ctx* ctx, int precision) {
  if (!ctx || precision < 0 || precision > 2) return SMCB_ERR_BAD_ARG;
  ctx->filter->set_precision(precision);
  return SMCB_OK;
}

int smcb_get_timing(const smcb_ctx* ctx, double ms[7], int64_t launches[7]) {
  if (!ctx || !ms || !launches) return SMCB_ERR_BAD_ARG;
  ctx->filter->timing(ms, launches);
  return SMCB_OK;
}

int smcb_synchronize(smcb_ctx* ctx) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    SMCB_CUDA_TRY(cudaSetDevice(ctx->device));
    SMCB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  });
}

int smcb_alloc_pinned(smcb_ctx* ctx, int64_t bytes, void** out) {
  if (!ctx || !out || bytes <= 0) return SMCB_ERR_BAD_ARG;
  *out = nullptr;
  return guarded(ctx, [&] {
    SMCB_CUDA_TRY(cudaSetDevice(ctx->device));
    SMCB_CUDA_TRY(cudaMallocHost(out, (size_t)bytes));
  });
}

int smcb_free_pinned(smcb_ctx* ctx, void* ptr) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    if (ptr) SMCB_CUDA_TRY(cudaFreeHost(ptr));
  });
}

// ------------------------------------------------------------------ utilities
int smcb_normalize(smcb_ctx* ctx, const double* logw, int64_t n, double* logmu, double* w, double* ess) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(logw && n >= 1, "normalize: logw must be non-null and n >= 1");
    StepStats st;
    ctx->scratch->normalize_vector(logw, n, &st, w);
    stats_to(st, n, logmu, ess);
  });
}

int smcb_resample(smcb_ctx* ctx, const double* w, int64_t n, int resampler, uint32_t stream, uint32_t t,
                  uint32_t purpose, int64_t* ancestors) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(w && ancestors && n >= 1, "resample: w, ancestors must be non-null and n >= 1");
    need(purpose >= 1 && purpose <= 15, "resample: purpose must be in [1, 15]");
    ctx->scratch->resample_vector(w, n, resampler, ctx->key(ctx->next_epoch), stream, t, purpose, ancestors);
  });
}

int smcb_resample_n(smcb_ctx* ctx, const double* w, int64_t n, int64_t n_out, int resampler, uint32_t stream, uint32_t t,
                    uint32_t purpose, int64_t* ancestors) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(w && ancestors && n >= 1 && n_out >= 1, "resample_n: w, ancestors must be non-null and n, n_out >= 1");
    need(purpose >= 1 && purpose <= 15, "resample_n: purpose must be in [1, 15]");
    ctx->scratch->resample_vector(w, n, resampler, ctx->key(ctx->next_epoch), stream, t, purpose, ancestors, n_out);
  });
}

// ------------------------------------------------------------------ one filter
int smcb_bootstrap_init(smcb_ctx* ctx, int kind, const double* params, int64_t N, double y, uint32_t stream,
                        double* logmu, double* ess) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(params != nullptr, "bootstrap_init: params is null");
    StepStats st;
    ctx->filter->init(kind, params, N, y, ctx->key(ctx->next_epoch), stream, &st);
    ctx->next_epoch = (ctx->next_epoch + 1) & 0xFFFFFFu;
    stats_to(st, N, logmu, ess);
  });
}

int smcb_bootstrap_step(smcb_ctx* ctx, const double* params, double y, int resampler, double* logmu, double* ess) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    StepStats st;
    ctx->filter->step(params, y, resampler, &st);
    stats_to(st, ctx->filter->N(), logmu, ess);
  });
}

int smcb_log_likelihood(smcb_ctx* ctx, int kind, const double* params, int64_t N, const double* y, int64_t T,
                        int resampler, uint32_t stream, double* logZ, double* logmu_out, double* ess_out) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(params && y && T >= 1, "log_likelihood: params, y must be non-null and T >= 1");
    ctx->stats.resize((size_t)T);
    ctx->filter->run(kind, params, N, y, T, resampler, ctx->key(ctx->next_epoch), stream, ctx->stats.data());
    ctx->next_epoch = (ctx->next_epoch + 1) & 0xFFFFFFu;
    double z = 0.0;
    for (int64_t t = 0; t < T; ++t) {
      double lm, es;
      stats_to(ctx->stats[(size_t)t], N, &lm, &es);
      z += lm;  // logZ += logμ      particles.jl:143
      if (logmu_out) logmu_out[t] = lm;
      if (ess_out) ess_out[t] = es;
    }
    if (logZ) *logZ = z;
  });
}

int smcb_guided_step(smcb_ctx* ctx, const double* params, double y, int resampler, const double* proposal, double* logmu,
                     double* ess) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(proposal != nullptr, "guided_step: proposal is null (use smcb_bootstrap_step for the bootstrap filter)");
    StepStats st;
    ctx->filter->step(params, y, resampler, &st, proposal);
    stats_to(st, ctx->filter->N(), logmu, ess);
  });
}

int smcb_guided_log_likelihood(smcb_ctx* ctx, int kind, const double* params, int64_t N, const double* y, int64_t T,
                               int resampler, uint32_t stream, const double* proposal, double* logZ, double* logmu_out,
                               double* ess_out) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(params && y && proposal && T >= 1, "guided_log_likelihood: params, y, proposal must be non-null and T >= 1");
    if (kind != KIND_LG1D && kind != KIND_SV && kind != KIND_UCSV)
      throw Error{SMCB_ERR_UNSUPPORTED, "guided proposals are defined for LG1D, SV (docs/SPEC.md §10) and UCSV (§10b)"};
    if (resampler == RESAMPLE_MULTINOMIAL && T > 1)
      throw Error{SMCB_ERR_UNSUPPORTED, "guided single filter: stratified or systematic resampling (multinomial guided filters run on the batched engine, N <= 8192)"};
    ctx->stats.resize((size_t)T);
    ctx->filter->run(kind, params, N, y, T, resampler, ctx->key(ctx->next_epoch), stream, ctx->stats.data(), proposal);
    ctx->next_epoch = (ctx->next_epoch + 1) & 0xFFFFFFu;
    double z = 0.0;
    for (int64_t t = 0; t < T; ++t) {
      double lm, es;
      stats_to(ctx->stats[(size_t)t], N, &lm, &es);
      z += lm;
      if (logmu_out) logmu_out[t] = lm;
      if (ess_out) ess_out[t] = es;
    }
    if (logZ) *logZ = z;
  });
}

int smcb_fetch_state(smcb_ctx* ctx, double* x, double* w, double* logw) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] { ctx->filter->fetch(x, w, logw); });
}

int smcb_weighted_summary(smcb_ctx* ctx, const double* probs, int nprobs, int weighted, double* mean, double* var, double* quantiles) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(nprobs == 0 || (probs && quantiles), "weighted_summary: probabilities without an output buffer");
    ctx->filter->summary(probs, nprobs, weighted != 0, mean, var, quantiles);
  });
}

int smcb_fetch_ancestors(smcb_ctx* ctx, int64_t* ancestors, int64_t rows_cap, int64_t* rows_out) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(ancestors && rows_cap >= 0, "fetch_ancestors: bad buffer");
    const int64_t r = ctx->filter->fetch_ancestors(ancestors, rows_cap);
    if (rows_out) *rows_out = r;
  });
}

int smcb_device_state(smcb_ctx* ctx, const double** x_dev, const double** logw_dev, int64_t* ld) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    if (!ctx->filter->live()) throw Error{SMCB_ERR_STATE, "no filter state"};
    if (x_dev) *x_dev = ctx->filter->dev_x();
    if (logw_dev) *logw_dev = ctx->filter->dev_logw();
    if (ld) *ld = ctx->filter->ld();
  });
}

// ------------------------------------------------------------------ batch
int smcb_batch_create(smcb_ctx* ctx, int kind, int64_t M, int64_t N, smcb_batch** out) {
  if (!ctx || !out) return SMCB_ERR_BAD_ARG;
  *out = nullptr;
  return guarded(ctx, [&] {
    std::unique_ptr<smcb_batch> b(new smcb_batch);
    b->ctx = ctx;
    b->impl = new BatchFilter(ctx->device, ctx->stream, kind, M, N);
    *out = b.release();
  });
}

int smcb_batch_destroy(smcb_batch* b) {
  if (!b || !b->owned) return SMCB_OK;  // a sampler's view dies with the sampler
  cudaSetDevice(b->ctx->device);
  delete b;
  return SMCB_OK;
}

int smcb_batch_init(smcb_batch* b, const double* params, const uint8_t* active, double y, uint32_t stream0,
                    double* logmu, double* ess) {
  if (!b) return SMCB_ERR_BAD_ARG;
  smcb_ctx* ctx = b->ctx;
  return guarded(ctx, [&] {
    need(params != nullptr, "batch_init: params is null");
    b->impl->init(params, active, y, ctx->key(ctx->next_epoch), stream0, logmu, ess);
    ctx->next_epoch = (ctx->next_epoch + 1) & 0xFFFFFFu;
  });
}

int smcb_batch_step(smcb_batch* b, const double* params, double y, int resampler, double* logmu, double* ess) {
  if (!b) return SMCB_ERR_BAD_ARG;
  return guarded(b->ctx, [&] { b->impl->step(params, y, resampler, logmu, ess); });
}

int smcb_batch_log_likelihood(smcb_batch* b, const double* params, const uint8_t* active, const double* y,
                              int64_t T, int resampler, uint32_t stream0, double* logZ) {
  if (!b) return SMCB_ERR_BAD_ARG;
  smcb_ctx* ctx = b->ctx;
  return guarded(ctx, [&] {
    need(params && y && T >= 1 && logZ, "batch_log_likelihood: params, y, logZ must be non-null and T >= 1");
    b->impl->run(params, active, y, T, resampler, ctx->key(ctx->next_epoch), stream0, logZ);
    ctx->next_epoch = (ctx->next_epoch + 1) & 0xFFFFFFu;
  });
}

int smcb_batch_step_guided(smcb_batch* b, const double* params, double y, int resampler, const double* proposal,
                           double* logmu, double* ess) {
  if (!b) return SMCB_ERR_BAD_ARG;
  return guarded(b->ctx, [&] {
    need(proposal != nullptr, "batch_step_guided: proposal is null (use smcb_batch_step for the bootstrap filter)");
    b->impl->step(params, y, resampler, logmu, ess, proposal);
  });
}

int smcb_batch_log_likelihood_guided(smcb_batch* b, const double* params, const uint8_t* active, const double* y,
                                     int64_t T, int resampler, uint32_t stream0, const double* proposal, double* logZ) {
  if (!b) return SMCB_ERR_BAD_ARG;
  smcb_ctx* ctx = b->ctx;
  return guarded(ctx, [&] {
    need(params && y && T >= 1 && logZ && proposal, "batch_log_likelihood_guided: params, y, proposal, logZ must be non-null and T >= 1");
    b->impl->run(params, active, y, T, resampler, ctx->key(ctx->next_epoch), stream0, logZ, proposal);
    ctx->next_epoch = (ctx->next_epoch + 1) & 0xFFFFFFu;
  });
}

int smcb_batch_gather(smcb_batch* b, const int32_t* parents) {
  if (!b) return SMCB_ERR_BAD_ARG;
  return guarded(b->ctx, [&] {
    need(parents != nullptr, "batch_gather: parents is null");
    b->impl->gather(parents);
  });
}

int smcb_batch_accept(smcb_batch* current, const smcb_batch* proposal, const uint8_t* accept) {
  if (!current || !proposal) return SMCB_ERR_BAD_ARG;
  return guarded(current->ctx, [&] {
    need(accept != nullptr, "batch_accept: accept is null");
    current->impl->accept_from(*proposal->impl, accept);
  });
}

int smcb_batch_fetch(smcb_batch* b, double* x, double* w, double* logw) {
  if (!b) return SMCB_ERR_BAD_ARG;
  return guarded(b->ctx, [&] { b->impl->fetch(x, w, logw); });
}

int smcb_batch_weighted_quantiles(smcb_batch* b, const double* probs, int nprobs, int weighted, double* quantiles) {
  if (!b || !probs || !quantiles) return SMCB_ERR_BAD_ARG;
  return guarded(b->ctx, [&] { b->impl->weighted_quantiles(probs, nprobs, weighted != 0, quantiles); });
}

int smcb_batch_weighted_mean(smcb_batch* b, double* mean) {
  if (!b || !mean) return SMCB_ERR_BAD_ARG;
  return guarded(b->ctx, [&] { b->impl->weighted_mean(mean); });
}

int smcb_batch_weighted_moments(smcb_batch* b, double* mean, double* var) {
  if (!b || (!mean && !var)) return SMCB_ERR_BAD_ARG;
  return guarded(b->ctx, [&] { b->impl->weighted_moments(mean, var); });
}

int64_t smcb_batch_cloud_bytes(const smcb_batch* b) { return b ? b->impl->cloud_bytes() : 0; }

int smcb_batch_pack(smcb_batch* b, const int32_t* slots, int64_t n, void* buf_dev) {
  if (!b) return SMCB_ERR_BAD_ARG;
  return guarded(b->ctx, [&] {
    need(n >= 0 && (n == 0 || (slots && buf_dev)), "batch_pack: bad arguments");
    b->impl->pack(slots, n, buf_dev, true);
  });
}

int smcb_batch_unpack(smcb_batch* b, const int32_t* slots, int64_t n, const void* buf_dev) {
  if (!b) return SMCB_ERR_BAD_ARG;
  return guarded(b->ctx, [&] {
    need(n >= 0 && (n == 0 || (slots && buf_dev)), "batch_unpack: bad arguments");
    b->impl->pack(slots, n, const_cast<void*>(buf_dev), false);
  });
}

int smcb_batch_get_timing(const smcb_batch* b, double* ms_last_call, int64_t* launches_total) {
  if (!b) return SMCB_ERR_BAD_ARG;
  if (ms_last_call) *ms_last_call = b->impl->last_ms();
  if (launches_total) *launches_total = b->impl->launches();
  return SMCB_OK;
}

// ------------------------------------------------------------------ multi-GPU (one process per GPU)
int smcb_comm_unique_id(uint8_t id[128]) {
  if (!id) return SMCB_ERR_BAD_ARG;
  return guarded(nullptr, [&] {
    NcclUniqueId u;
    SMCB_NCCL_TRY(nccl_api().GetUniqueId(&u));
    std::memcpy(id, u.internal, 128);
  });
}

int smcb_comm_init(smcb_ctx* ctx, int rank, int nranks, const uint8_t id[128]) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(nranks >= 1 && rank >= 0 && rank < nranks, "comm_init: rank must be in [0, nranks)");
    if (ctx->comm.comm) throw Error{SMCB_ERR_STATE, "comm_init: the context already has a communicator"};
    if (nranks == 1) { ctx->comm.rank = 0; ctx->comm.nranks = 1; return; }
    need(id != nullptr, "comm_init: id is null");
    SMCB_CUDA_TRY(cudaSetDevice(ctx->device));
    NcclUniqueId u;
    std::memcpy(u.internal, id, 128);
    NcclComm c = nullptr;
    SMCB_NCCL_TRY(nccl_api().CommInitRank(&c, nranks, u, rank));
    ctx->comm.comm = c;
    ctx->comm.rank = rank;
    ctx->comm.nranks = nranks;
    // build the rings / channels now, outside anybody's timed region
    double* tmp = nullptr;
    SMCB_CUDA_TRY(cudaMalloc(&tmp, sizeof(double) * nranks));
    cudaError_t e = cudaMemsetAsync(tmp, 0, sizeof(double) * nranks, ctx->stream);
    int r = nccl_api().AllGather(tmp + rank, tmp, 1, kNcclFloat64, c, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    SMCB_CUDA_TRY(e);
    SMCB_NCCL_TRY(r);
  });
}

int smcb_comm_rank(const smcb_ctx* ctx, int* rank, int* nranks) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  if (rank) *rank = ctx->comm.rank;
  if (nranks) *nranks = ctx->comm.nranks;
  return SMCB_OK;
}

int smcb_comm_all_gather(smcb_ctx* ctx, const double* local, int64_t n_local, double* all) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(local && all && n_local >= 1, "comm_all_gather: bad arguments");
    const int G = ctx->comm.nranks;
    if (G == 1) { std::memcpy(all, local, sizeof(double) * n_local); return; }
    SMCB_CUDA_TRY(cudaSetDevice(ctx->device));
    if (ctx->comm_buf_cap < n_local * G) {
      cudaFree(ctx->comm_buf);
      ctx->comm_buf = nullptr;
      ctx->comm_buf_cap = 0;
      SMCB_CUDA_TRY(cudaMalloc(&ctx->comm_buf, sizeof(double) * n_local * G));
      ctx->comm_buf_cap = n_local * G;
    }
    double* mine = ctx->comm_buf + (int64_t)ctx->comm.rank * n_local;
    SMCB_CUDA_TRY(cudaMemcpyAsync(mine, local, sizeof(double) * n_local, cudaMemcpyHostToDevice, ctx->stream));
    SMCB_NCCL_TRY(nccl_api().AllGather(mine, ctx->comm_buf, (size_t)n_local, kNcclFloat64, ctx->comm.comm, ctx->stream));
    SMCB_CUDA_TRY(cudaMemcpyAsync(all, ctx->comm_buf, sizeof(double) * n_local * G, cudaMemcpyDeviceToHost, ctx->stream));
    SMCB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  });
}

int smcb_comm_destroy(smcb_ctx* ctx) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  if (ctx->comm.comm) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nccl_api().CommDestroy(ctx->comm.comm);
  }
  cudaFree(ctx->comm_buf);
  ctx->comm_buf = nullptr;
  ctx->comm_buf_cap = 0;
  ctx->comm = Comm{};
  return SMCB_OK;
}

int smcb_batch_chunk_plan(int64_t M, int64_t N, int64_t steps, int threads, int64_t slots, int num_sms, int masked, int64_t* chunk) {
  if (!chunk || M < 1 || N < 1 || steps < 1 || threads < 1) return SMCB_ERR_BAD_ARG;
  *chunk = smcb::plan_batch_chunk(M, N, steps, threads, slots, num_sms, masked != 0);
  return SMCB_OK;
}

int smcb_exchange_plan(const int32_t* parents, int64_t M, int rank, int nranks, int32_t* local_parents, int32_t* send_peer,
                       int32_t* send_slot, int64_t* n_send, int32_t* recv_peer, int32_t* recv_slot, int64_t* n_recv) {
  if (!parents || M < 1 || nranks < 1 || rank < 0 || rank >= nranks || M % nranks || !local_parents || !n_send || !n_recv)
    return SMCB_ERR_BAD_ARG;
  for (int64_t m = 0; m < M; ++m)
    if (parents[m] < 0 || parents[m] >= M) return SMCB_ERR_BAD_ARG;
  return guarded(nullptr, [&] {
    ExchangePlan plan;
    make_exchange_plan(parents, M, rank, nranks, plan);
    std::memcpy(local_parents, plan.local_parents.data(), sizeof(int32_t) * plan.local_parents.size());
    *n_send = (int64_t)plan.send_slot.size();
    *n_recv = (int64_t)plan.recv_slot.size();
    for (size_t i = 0; i < plan.send_slot.size(); ++i) {
      if (send_peer) send_peer[i] = plan.send_peer[i];
      if (send_slot) send_slot[i] = plan.send_slot[i];
    }
    for (size_t i = 0; i < plan.recv_slot.size(); ++i) {
      if (recv_peer) recv_peer[i] = plan.recv_peer[i];
      if (recv_slot) recv_slot[i] = plan.recv_slot[i];
    }
  });
}

int smcb_random_walk_sigma(const double* theta, int64_t M, int d, double* sigma) {
  if (!theta || !sigma || M < 2 || d < 1 || d > kMaxThetaDim) return SMCB_ERR_BAD_ARG;
  random_walk_sigma(theta, M, d, sigma);
  return SMCB_OK;
}

int smcb_cholesky_lower(const double* A, int d, double scale, double* L) {
  if (!A || !L || d < 1 || d > kMaxThetaDim) return SMCB_ERR_BAD_ARG;
  return cholesky_lower(A, d, scale, L) ? SMCB_OK : SMCB_ERR_STATE;
}

// ------------------------------------------------------------------ device-resident θ-level samplers
int smcb_sampler_create(smcb_ctx* ctx, const smcb_sampler_config* cfg, const double* theta0, smcb_sampler** out) {
  if (!ctx || !out) return SMCB_ERR_BAD_ARG;
  *out = nullptr;
  return guarded(ctx, [&] {
    need(cfg && theta0, "sampler_create: cfg, theta0 must be non-null");
    std::unique_ptr<smcb_sampler> s(new smcb_sampler);
    s->ctx = ctx;
    s->impl.reset(new ThetaSampler(ctx->device, ctx->stream, ctx->comm, *cfg, theta0));
    s->view.ctx = ctx;
    s->view.owned = false;
    *out = s.release();
  });
}

int smcb_sampler_destroy(smcb_sampler* s) {
  if (!s) return SMCB_OK;
  cudaSetDevice(s->ctx->device);
  delete s;
  return SMCB_OK;
}

int smcb_sampler_set_data(smcb_sampler* s, const double* y, int64_t T) {
  if (!s) return SMCB_ERR_BAD_ARG;
  return guarded(s->ctx, [&] { s->impl->set_data(y, T); });
}

int smcb_sampler_smc2_init(smcb_sampler* s) {
  if (!s) return SMCB_ERR_BAD_ARG;
  return guarded(s->ctx, [&] { s->impl->smc2_init(); });
}

int smcb_sampler_smc2_step(smcb_sampler* s, int64_t t, double* ess, int* rejuvenated) {
  if (!s) return SMCB_ERR_BAD_ARG;
  return guarded(s->ctx, [&] { s->impl->smc2_step(t, ess, rejuvenated); });
}

int smcb_sampler_density_tempered(smcb_sampler* s, double* schedule, int cap, int* n_stages) {
  if (!s) return SMCB_ERR_BAD_ARG;
  return guarded(s->ctx, [&] {
    const int n = s->impl->density_tempered(schedule, cap);
    if (n_stages) *n_stages = n;
  });
}

int smcb_sampler_get(smcb_sampler* s, double* theta, double* omega, double* logZ, double* ess, double* acc_ratio, int64_t* N) {
  if (!s) return SMCB_ERR_BAD_ARG;
  return guarded(s->ctx, [&] { s->impl->get(theta, omega, logZ, ess, acc_ratio, N); });
}

int smcb_sampler_clouds(smcb_sampler* s, smcb_batch** clouds) {
  if (!s || !clouds) return SMCB_ERR_BAD_ARG;
  s->view.impl = s->impl->clouds();
  *clouds = &s->view;
  return SMCB_OK;
}

int smcb_sampler_set_profiling(smcb_sampler* s, int enable) {
  if (!s) return SMCB_ERR_BAD_ARG;
  s->impl->set_profiling(enable != 0);
  return SMCB_OK;
}

int smcb_sampler_stats(smcb_sampler* s, double ms[8], int64_t counts[8]) {
  if (!s || !ms || !counts) return SMCB_ERR_BAD_ARG;
  return guarded(s->ctx, [&] { s->impl->stats(ms, counts); });
}

// ------------------------------------------------------------------ Kalman
int smcb_kalman_batch_step(smcb_ctx* ctx, const double* params, int64_t M, double y, double* x, double* sigma,
                           double* loglik) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(params && x && sigma && loglik && M >= 1, "kalman_batch_step: bad arguments");
    kalman_batch(ctx->device, ctx->stream, ctx->kalman_scratch, params, nullptr, M, &y, 1, /*predict_first=*/true, loglik, x, sigma,
                 /*use_state=*/true);
  });
}

int smcb_kalman_batch_loglik(smcb_ctx* ctx, const double* params, const uint8_t* active, int64_t M, const double* y,
                             int64_t T, int matched_init, double* loglik, double* x, double* sigma) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(params && y && loglik && M >= 1 && T >= 1, "kalman_batch_loglik: bad arguments");
    kalman_batch(ctx->device, ctx->stream, ctx->kalman_scratch, params, active, M, y, T, /*predict_first=*/!matched_init, loglik, x, sigma,
                 /*use_state=*/false);
  });
}

int smcb_kalman_mv_batch_step(smcb_ctx* ctx, int d, const double* models, int64_t M, double y, double* x, double* sigma,
                              double* loglik) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(models && x && sigma && loglik && M >= 1, "kalman_mv_batch_step: bad arguments");
    kalman_mv_batch(ctx->device, ctx->stream, ctx->kalman_scratch, d, models, nullptr, M, &y, 1, /*predict_first=*/true, loglik, x, sigma,
                    /*use_state=*/true);
  });
}

int smcb_kalman_mv_batch_loglik(smcb_ctx* ctx, int d, const double* models, const uint8_t* active, int64_t M,
                                const double* y, int64_t T, int matched_init, double* loglik, double* x, double* sigma) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(models && y && loglik && M >= 1 && T >= 1, "kalman_mv_batch_loglik: bad arguments");
    kalman_mv_batch(ctx->device, ctx->stream, ctx->kalman_scratch, d, models, active, M, y, T, /*predict_first=*/!matched_init, loglik, x, sigma,
                    /*use_state=*/false);
  });
}

// ------------------------------------------------------------------ host-side helpers
int smcb_rng_normals(uint64_t seed, uint32_t epoch, uint32_t stream, uint32_t t, uint32_t purpose, uint32_t comp,
                     int64_t n, double* out) {
  if (!out || n < 0 || purpose > 15 || comp > 15) return SMCB_ERR_BAD_ARG;
  const RngKey key = make_rng_key((uint32_t)seed, (uint32_t)(seed >> 32), epoch & 0xFFFFFFu);
  for (int64_t p = 0; 2 * p < n; ++p) {
    double z0, z1;
    normal_pair_at(key, (uint32_t)p, stream, t, purpose, comp, z0, z1);
    out[2 * p] = z0;
    if (2 * p + 1 < n) out[2 * p + 1] = z1;
  }
  return SMCB_OK;
}

int smcb_rng_uniforms64(uint64_t seed, uint32_t epoch, uint32_t stream, uint32_t t, uint32_t purpose, int64_t n,
                        uint64_t* out) {
  if (!out || n < 0 || purpose > 15) return SMCB_ERR_BAD_ARG;
  const RngKey key = make_rng_key((uint32_t)seed, (uint32_t)(seed >> 32), epoch & 0xFFFFFFu);
  for (int64_t i = 0; i < n; ++i) out[i] = uniform64_at(key, (uint32_t)i, stream, t, purpose);
  return SMCB_OK;
}

}  // extern "C"

namespace {
template <class Model>
void simulate_model(const double* D, const double* P, int64_t T, uint64_t seed, double* x, double* y) {
  // state noise: purpose SIMULATE stream 0 component k, index t; observation noise: stream 1
  const RngKey key = make_rng_key((uint32_t)seed, (uint32_t)(seed >> 32), 0u);
  Model mdl;
  mdl.load(D);
  constexpr int DIM = Model::D;
  double cur[DIM], nxt[DIM], z[DIM];
  for (int64_t t = 0; t < T; ++t) {
    for (int k = 0; k < DIM; ++k) {
      double z0, z1;
      normal_pair_at(key, (uint32_t)(t >> 1), 0u, 0u, PURPOSE_SIMULATE, (uint32_t)k, z0, z1);
      z[k] = (t & 1) ? z1 : z0;
    }
    if (t == 0) mdl.init(z, nxt); else mdl.transition(z, cur, nxt);
    for (int k = 0; k < DIM; ++k) { cur[k] = nxt[k]; x[(int64_t)k * T + t] = cur[k]; }
    double z0, z1;
    normal_pair_at(key, (uint32_t)(t >> 1), 1u, 0u, PURPOSE_SIMULATE, 0u, z0, z1);
    const double zo = (t & 1) ? z1 : z0;
    double mean, sd;
    if (Model::KIND >= KIND_MVLG2) {
      const double* B = D + DIM * DIM;
      mean = B[0] * cur[0];
      for (int j = 1; j < DIM; ++j) mean = fma(B[j], cur[j], mean);
      sd = sqrt(P[2 * DIM * DIM + DIM]);
    }
    else if (Model::KIND == KIND_LG1D) { mean = D[1] * cur[0]; sd = sqrt(P[3]); }
    else if (Model::KIND == KIND_SV) { mean = 0.0; sd = det_exp(0.5 * cur[0]); }
    else { mean = cur[0]; sd = det_exp(0.5 * cur[DIM - 1]); }
    y[t] = fma(sd, zo, mean);
  }
}

__global__ void selftest_kernel(int fn, const double* in, int64_t n, double aux, double* o0, double* o1) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = in[i];
  if (fn == 0) o0[i] = det_exp(v);
  else if (fn == 1) o0[i] = det_log(v);
  else if (fn == 2) { double s, c; det_sincos2pi(v, s, c); o0[i] = s; o1[i] = c; }
  else if (fn == 3) { double e; uint64_t q; det_exp_quant(v, (int)aux, e, q); o0[i] = u64_as_double(q); o1[i] = e; }
  // the binary32 functions of SPEC §9b (inputs are rounded to binary32 first, outputs widened exactly)
  else if (fn == 4) o0[i] = (double)det_expf((float)v);
  else if (fn == 5) o0[i] = (double)det_logf((float)v);
  else if (fn == 6) { float s, c; det_sincos2pif((float)v, s, c); o0[i] = (double)s; o1[i] = (double)c; }
  else { float e; uint64_t q; det_exp_quantf((float)v, (int)aux, e, q); o0[i] = u64_as_double(q); o1[i] = (double)e; }
}
}  // namespace

extern "C" {

int smcb_simulate(int kind, const double* params, int64_t T, uint64_t seed, double* x, double* y) {
  if (!params || !x || !y || T < 1 || kind < 0 || kind >= KIND_ALL) return SMCB_ERR_BAD_ARG;
  double D[kMvStride];
  if (is_mv_kind(kind)) {
    derive_params_mv(state_dim(kind), params, D);
    if (kind == KIND_MVLG2) simulate_model<ModelMVLG<2>>(D, params, T, seed, x, y);
    else if (kind == KIND_MVLG3) simulate_model<ModelMVLG<3>>(D, params, T, seed, x, y);
    else simulate_model<ModelMVLG<4>>(D, params, T, seed, x, y);
    return SMCB_OK;
  }
  derive_params(kind, params, D);
  if (kind == KIND_LG1D) simulate_model<ModelLG1D>(D, params, T, seed, x, y);
  else if (kind == KIND_SV) simulate_model<ModelSV>(D, params, T, seed, x, y);
  else simulate_model<ModelUCSV>(D, params, T, seed, x, y);
  return SMCB_OK;
}

int smcb_selftest_math(smcb_ctx* ctx, int fn, const double* in, int64_t n, double aux, double* out0, double* out1) {
  if (!ctx) return SMCB_ERR_BAD_ARG;
  return guarded(ctx, [&] {
    need(in && out0 && n >= 1 && fn >= 0 && fn <= 7, "selftest_math: bad arguments");
    SMCB_CUDA_TRY(cudaSetDevice(ctx->device));
    double *di = nullptr, *d0 = nullptr, *d1 = nullptr;
    cudaError_t e = cudaMalloc(&di, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMalloc(&d0, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMalloc(&d1, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(di, in, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
      selftest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(fn, di, n, aux, d0, d1);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out0, d0, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && out1 && (fn == 2 || fn == 3 || fn >= 6)) e = cudaMemcpyAsync(out1, d1, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(di); cudaFree(d0); cudaFree(d1);
    SMCB_CUDA_TRY(e);
  });
}

}  // extern "C"
