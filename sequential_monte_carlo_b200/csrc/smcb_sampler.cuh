// Device-resident θ-level samplers: SMC² and density-tempered SMC with every M-length vector (θ, ω, logZ,
// log-prior, parameter blocks) living on the GPU, replicated on every rank, and the M inner particle filters
// sharded over the ranks of an NCCL communicator.
//
// Replaces, behind one C handle (include/smcb200.h, smcb_sampler_*):
//   SMC(...)                 /root/reference/src/smc_samplers.jl:29-59
//   resample!(smc)           :74-84      theta_resample_kernel + cloud exchange (device gather / ncclSend+Recv)
//   random_walk_kernel       :87-101     host, d <= 8 (fixed summation order, docs/SPEC.md §11)
//   rejuvenate!              :103-148    theta_propose_kernel -> batch_kernel sweep -> all-gather -> theta_accept_kernel
//   exchange!                :163-189    new batches at 2N, one sweep, reweight
//   density_tempered         :222-281    tempering bisection in ONE single-CTA kernel per stage
//   smc² / smc²!             :288-340    batch_kernel step -> ncclAllGather -> theta_step_kernel -> 32-byte D2H
// The reference runs these loops on the host with one CPU particle filter per θ; the host mirror
// (smc_samplers.py, the Julia shim) used to keep the M-vectors in numpy and paid an H2D/D2H round trip per vector.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>
#include <vector>

#include "../../include/smcb200.h"
#include "smcb_batch.cuh"
#include "smcb_nccl.cuh"

namespace smcb {

constexpr int kMaxThetaDim = 8;
constexpr int kPriorStride = 8;  // kind, p0, p1, lo, hi, c0, c1, (unused)
constexpr int64_t kMaxThetaParticles = 16384;

struct PriorTable {  // by-value kernel argument
  double row[kMaxThetaDim][kPriorStride];
  int d;
};
struct ParamMap {  // params[k] = src[k] >= 0 ? θ[src[k]] : cst[k]
  int src[kParamStride];
  double cst[kParamStride];
};
struct CholFactor {  // lower-triangular factor of scale·Σ of one chain step (univariate: l[0][0] = scale·σ)
  double l[kMaxThetaDim][kMaxThetaDim];
};

// what the host reads back after a step: 64 bytes in page-locked memory
struct ThetaScalars {
  double ess;
  double xi;          // tempering exponent chosen by the bisection
  double acc_count;   // θ-particles that accepted at least one move in the last rejuvenation
  double logsum;      // log Σ exp(logω) of the last normalisation (evidence increment)
  int32_t resample_flag;
  int32_t not_pd;
  double pad[3];
};

// host-only: who sends which cloud where after a θ-resample (replicated, deterministic)
struct ExchangePlan {
  std::vector<int32_t> local_parents;               // [Mloc]: rank-local parent slot, or the slot itself when the parent is remote
  std::vector<int32_t> send_peer, send_slot;        // clouds this rank sends, grouped by destination rank, increasing global slot
  std::vector<int32_t> recv_peer, recv_slot;        // clouds this rank receives, grouped by source rank, increasing global slot
};
void make_exchange_plan(const int32_t* parents, int64_t M, int rank, int nranks, ExchangePlan& plan);

// host-only: Σ of random_walk_kernel (smc_samplers.jl:87-101) with the summation order of docs/SPEC.md §11; returns false when
// the reference would take the "small covariance" branch; out is d×d row-major (univariate: out[0] = σ, a standard deviation)
void random_walk_sigma(const double* theta, int64_t M, int d, double* out);
// lower Cholesky factor of scale·Σ (row-major d×d, plain Cholesky–Banachiewicz); false if not positive definite
bool cholesky_lower(const double* A, int d, double scale, double* L);

enum { SK_FILTER = 0, SK_ALLGATHER = 1, SK_EXCHANGE = 2, SK_THETA = 3, SK_COUNT = 4 };

class ThetaSampler {
 public:
  ThetaSampler(int device, cudaStream_t stream, const Comm& comm, const smcb_sampler_config& cfg, const double* theta0);
  ~ThetaSampler();
  ThetaSampler(const ThetaSampler&) = delete;
  ThetaSampler& operator=(const ThetaSampler&) = delete;

  void set_data(const double* y, int64_t T);
  void smc2_init();                                          // smc²(smc, y)
  void smc2_step(int64_t t, double* ess, int* rejuvenated);  // smc²!(smc, y, t)   (t 0-based)
  int density_tempered(double* schedule, int cap);           // returns the number of stages; schedule [cap][3] = (ξ, ess, acceptance ratio of the stage's rejuvenation or -1)
  void get(double* theta, double* omega, double* logZ, double* ess, double* acc_ratio, int64_t* N);
  BatchFilter* clouds() { return cur_.get(); }
  void set_profiling(bool on) { profiling_ = on; }
  void stats(double ms[8], int64_t counts[8]);

  int64_t M() const { return M_; }
  int64_t Mloc() const { return Mloc_; }
  int d_theta() const { return d_; }

 private:
  RngKey key(uint32_t epoch) const { return make_rng_key((uint32_t)seed_, (uint32_t)(seed_ >> 32), epoch & 0xFFFFFFu); }
  uint32_t next_epoch() { return epoch_++; }
  void all_gather(double* all);          // in place: every rank's slice [lo, lo + Mloc) of `all` -> all M entries everywhere
  void read_scalars();                   // 64-byte D2H + the one stream synchronisation of a step
  void resample();                       // resample!(smc)
  void rejuvenate(int64_t t_len, double xi);   // rejuvenate!(smc, y[1:t_len], ξ)
  void exchange(int64_t t_len);          // exchange!(smc, y[1:t_len])
  void mark(int klass, bool start);
  void resolve_marks();
  void open_span();

  int device_;
  cudaStream_t stream_;
  Comm comm_;
  int kind_, d_;
  int64_t N_, M_, Mloc_, lo_;
  int chain_, resampler_, theta_resampler_;
  double ess_min_, acc_threshold_;
  uint64_t seed_;
  PriorTable prior_{};
  ParamMap map_{};
  uint32_t epoch_ = 1;     // ordinal of the next batched sweep (device Philox epoch)
  uint32_t n_resample_ = 0, n_rejuv_ = 0;
  double ess_ = 0.0, acc_ratio_ = 0.0;
  bool started_ = false;

  std::unique_ptr<BatchFilter> cur_, prop_;

  // replicated device state, all [M] (θ: [M][d])
  double* theta_[2] = {nullptr, nullptr};
  double* logz_[2] = {nullptr, nullptr};
  double* lp_[2] = {nullptr, nullptr};
  double* derived_[2] = {nullptr, nullptr};   // [M][8] derived parameter blocks of the current θ
  int tcur_ = 0;
  double* omega_ = nullptr;
  double* theta_prop_ = nullptr;
  double* lp_prop_ = nullptr;
  double* derived_prop_ = nullptr;
  double* logz_prop_ = nullptr;     // [M] all-gathered log-likelihoods of the proposals
  double* logmu_ = nullptr;         // [M] all-gathered step increments
  uint8_t* ok_ = nullptr;
  uint8_t* accept_ = nullptr;
  uint8_t* acc_any_ = nullptr;
  int32_t* anc_ = nullptr;
  int32_t* slots_dev_ = nullptr;    // [3][M]: local parents, send slots, recv slots
  ThetaScalars* scal_dev_ = nullptr;
  double* y_dev_ = nullptr;
  int64_t T_ = 0, y_cap_ = 0;
  void* xbuf_[2] = {nullptr, nullptr};   // packed clouds: send, recv
  int64_t xbuf_cap_[2] = {0, 0};

  // page-locked host mirrors
  ThetaScalars* scal_host_ = nullptr;
  int32_t* anc_host_ = nullptr;
  int32_t* slots_host_ = nullptr;   // [3][M]
  double* theta_host_ = nullptr;    // [M][d] mirror of θ (refreshed after every rejuvenation; the covariance is a host job)
  std::vector<double> theta_tmp_;

  // instrumentation
  bool profiling_ = false;
  struct Mark { int klass; cudaEvent_t e0, e1; };
  std::vector<Mark> marks_;
  std::vector<cudaEvent_t> ev_free_;
  double ms_[SK_COUNT] = {0, 0, 0, 0};
  cudaEvent_t span_[2] = {nullptr, nullptr};
  bool span_open_ = false;
  double span_ms_ = 0.0;   // CUDA-event time from the start of smc² / density_tempered to the last completed step
  int64_t n_sweeps_ = 0, n_steps_ = 0, n_rejuv_done_ = 0, n_clouds_moved_ = 0, n_particle_updates_ = 0, n_syncs_ = 0, n_theta_launches_ = 0,
          n_batch_launches_retired_ = 0;
};

}  // namespace smcb
