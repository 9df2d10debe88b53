// M independent bootstrap filters of N particles each, one CTA per θ-particle, the whole time
// loop inside one launch (the "particle of filters" of SMC² and the PMMH sweeps of rejuvenate!).
// Replaces the loops over θ at /root/reference/src/smc_samplers.jl:112-121,223-229,289-293,325-335.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "smcb_common.cuh"
#include "smcb_filter.cuh"
#include "smcb_models.cuh"

namespace smcb {

class BatchFilter {
 public:
  BatchFilter(int device, cudaStream_t stream, int kind, int64_t M, int64_t N);
  ~BatchFilter();
  BatchFilter(const BatchFilter&) = delete;
  BatchFilter& operator=(const BatchFilter&) = delete;

  // M × bootstrap_filter(N, y, model(θ_m))                       smc_samplers.jl:289-293
  void init(const double* params, const uint8_t* active, double y0, const RngKey& key, uint32_t stream0,
            double* logmu, double* ess);
  // M × bootstrap_filter!(x[m], w[m], y, model(θ_m))              smc_samplers.jl:325-335
  // proposal != null: M × particle_filter!(x[m], w[m], y, model(θ_m), proposal_m)  particles.jl:55-84 (SPEC §10);
  // proposal is [M][3] = (c0, c1, c2) of x' ~ N(c0 + c1 xp, c2²)
  void step(const double* params, double y, int resampler, double* logmu, double* ess, const double* proposal = nullptr);
  // M × log_likelihood(N, y, model(θ_m))                          smc_samplers.jl:117-121,223-229
  void run(const double* params, const uint8_t* active, const double* y, int64_t T, int resampler,
           const RngKey& key, uint32_t stream0, double* logZ, const double* proposal = nullptr);  // proposal: [T][M][3]

  void gather(const int32_t* parents);                             // smc_samplers.jl:74-84
  void accept_from(const BatchFilter& prop, const uint8_t* accept);  // smc_samplers.jl:130-133
  void fetch(double* x_host, double* w_host, double* logw_host);
  // [M][d][np]: lower empirical quantiles of every cloud under its weights (weighted) or counting particles once (SPEC §8)
  void weighted_quantiles(const double* probs, int np, bool weighted, double* q_host);
  void weighted_mean(double* mean_host);  // [M][d]: w[m]' x[m], computed on the device (plotting_utils.jl:116-124,150)
  // [M][d] each: mean and population variance of every cloud under its weights (var(x, weights(w)), inflation_example.jl:46)
  void weighted_moments(double* mean_host, double* var_host);
  int64_t cloud_bytes() const;
  void pack(const int32_t* slots, int64_t n, void* buf_dev, bool to_buffer);

  // ---- device-resident variants (the θ-level engine, smcb_sampler.cu): every pointer is a DEVICE pointer, nothing is
  // copied from or to the host and nothing synchronises — the calls only enqueue work on the stream.
  // derived_dev: [M][8] DERIVED parameter blocks (derive_params); active_dev: [M] or null; y_dev: the observations this
  // call consumes; logz_dev / logmu_dev / ess_dev: [M] outputs (ess_dev may be null).
  void init_dev(const double* derived_dev, const uint8_t* active_dev, const double* y0_dev, const RngKey& key, uint32_t stream0,
                double* logmu_dev, double* ess_dev);
  void step_dev(const double* derived_dev, const double* y_dev, uint32_t t, int resampler, double* logmu_dev, double* ess_dev);
  void run_dev(const double* derived_dev, const uint8_t* active_dev, const double* y_dev, int64_t T, int resampler,
               const RngKey& key, uint32_t stream0, double* logz_dev);
  void gather_dev(const int32_t* parents_dev);                              // slot m <- slot parents_dev[m]
  void accept_dev(const BatchFilter& prop, const uint8_t* mask_dev);        // slots with mask_dev[m] != 0 <- prop's slot m
  void pack_dev(const int32_t* slots_dev, int64_t n, void* buf_dev, bool to_buffer);
  int kind() const { return kind_; }
  int state_dim_() const { return d_; }
  int64_t ld() const { return ld_; }

  double last_ms() const { return last_ms_; }
  int64_t launches() const { return launches_; }
  int64_t M() const { return M_; }
  int64_t N() const { return N_; }

 private:
  void upload_params(const double* params, const uint8_t* active);
  struct IO {  // device pointers one launch reads its inputs from and writes its per-θ results to
    const double* derived;
    const uint8_t* active;
    const double* y;
    double* logz_out;
    double* ess_out;
  };
  IO own_io() const { return IO{derived_, use_active_ ? active_ : nullptr, y_dev_, out_dev_, out_dev_ + M_}; }
  void launch(const IO& io, bool from_init, uint32_t t_begin, uint32_t t_end, int resampler, bool guided = false);
  void upload_proposal(const double* proposal, int64_t rows);
  void begin_call();
  void end_call();
  void upload_slots(const int32_t* a, const int32_t* b, int64_t n);
  void ensure_scratch(size_t words);

  int device_;
  cudaStream_t stream_;
  int kind_, d_;
  int num_sms_ = 0;
  int64_t M_, N_, ld_;
  int S_;
  uint64_t R_;
  RngKey key_{};
  uint32_t stream0_ = 0;
  uint32_t t_ = 0;
  bool live_ = false;
  bool has_params_ = false;

  double* x_[2] = {nullptr, nullptr};     // [M][d][ld], [1] is the gather target
  double* logw_[2] = {nullptr, nullptr};  // [M][ld]
  StepStats* stats_[2] = {nullptr, nullptr};
  int cur_ = 0;
  double* derived_ = nullptr;  // [M][8]
  uint8_t* active_ = nullptr;  // [M]
  bool use_active_ = false;
  double* y_dev_ = nullptr;
  int64_t y_cap_ = 0;
  double* out_dev_ = nullptr;  // [2][M]: Σ logμ, ess
  double* w_tmp_ = nullptr;    // scratch of fetch / the summaries, grown on demand and kept
  size_t w_cap_ = 0;
  int32_t* slots_dev_ = nullptr;  // [2][slot_cap_]
  int64_t slot_cap_ = 0;
  std::vector<double> host_tmp_;
  double* prop_dev_ = nullptr;  // [cap][5] proposal coefficients of the guided launches, then [cap][3] raw upload
  int64_t prop_cap_ = 0;
  unsigned* sched_ = nullptr;   // [1 + M]: unit counter and per-θ chunk counters of the dynamically scheduled launches

  cudaEvent_t ev_[2] = {nullptr, nullptr};
  double last_ms_ = 0;
  int64_t launches_ = 0;
};

// steps per chunk of the dynamically scheduled batched launch (0: one CTA per θ-particle for the whole series); host-only
int64_t plan_batch_chunk(int64_t M, int64_t N, int64_t steps, int threads, int64_t slots, int num_sms, bool masked);

// grow-only device scratch of the Kalman entry points (owned by the context: IBIS calls them once per observation, and a
// cudaMalloc / cudaFree per call is a driver round trip that takes process-wide locks)
struct DeviceScratch {
  void* p = nullptr;
  size_t cap = 0;
  void* get(size_t bytes);   // >= bytes, 256-byte aligned; contents are not preserved across calls
  ~DeviceScratch();
};

// kalman_filter / log_likelihood(y, model) for M LG1D models (kalman_filter.jl:29-70).
// use_state: start from x/sigma given by the caller (one-step API); else from (x0, σ0) of params.
void kalman_batch(int device, cudaStream_t stream, DeviceScratch& scratch, const double* params, const uint8_t* active, int64_t M,
                  const double* y, int64_t T, bool predict_first, double* loglik, double* x, double* sigma,
                  bool use_state);


// kalman_filter / log_likelihood(y, model) for M multivariate LinearModels with a scalar observation
// (MultivariateLinearGaussian, hodrick_prescott: state_space_models.jl:137-202; kalman_filter.jl:3-27,55-70).
// models: [M][3d² + 2d + 1] = A[d][d], B[d], Q[d][d], R, x0[d], Σ0[d][d] row-major; x: [M][d], sigma: [M][d][d].
void kalman_mv_batch(int device, cudaStream_t stream, DeviceScratch& scratch, int d, const double* models, const uint8_t* active, int64_t M,
                     const double* y, int64_t T, bool predict_first, double* loglik, double* x, double* sigma,
                     bool use_state);

}  // namespace smcb
