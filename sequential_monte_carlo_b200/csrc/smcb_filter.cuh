// One large-N bootstrap particle filter resident on one GPU (grid-wide kernels).
// Host-side driver for bootstrap_filter / bootstrap_filter! / log_likelihood
// (/root/reference/src/particles.jl:87-147).  Kernels are in smcb_filter.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "smcb_common.cuh"
#include "smcb_models.cuh"

namespace smcb {

struct FilterCtrl {
  unsigned long long maxslot[2];  // ordered-encoded max(logw), slot = t & 1
  unsigned long long total;       // Q = C[N-1] of the last scan
  unsigned long long sys_off;     // mulhi(U(0), R) of the systematic resampler for the coming step
  unsigned long long rq_lo, rq_hi;  // R * Q as a 128-bit value: the increment of (i R + u) Q per particle
  double inv_rq;                  // 2^64 / (R Q) in double: ESTIMATE of 1 / (threshold spacing), see anc_hist_kernel
  unsigned int scan_ticket;       // dynamic tile ids of the look-back scan
  unsigned int scan_done;
};

// index of the sorted-resampler step (device pointers + tile geometry), passed to the kernels by value
struct StepIndex {
  unsigned long long* tile_tot;   // [ntiles]
  unsigned long long* tile_lexcl; // [ntiles] exclusive prefix of the tile inside its sum_kernel CTA (8 tiles)
  unsigned long long* cta_tot;    // [ntiles / 8] total of each sum_kernel CTA
  unsigned long long* tile_excl;  // [ntiles] global exclusive prefix of the tile
  unsigned long long* tile_incl;  // [ntiles]
  int32_t* bound_pos;             // [nblocks + 1] ancestor of each propagate CTA's first particle (last: of particle N-1)
  int32_t* bound_tile;            // [nblocks + 1] its tile
  int ntiles;
  int chunks_per_tile;
  int tile_items;
};

struct StepStats {  // normalize() ingredients for one time step (SPEC §6)
  double mx, sum, sum2;
};

enum { TK_TOTAL = 0, TK_SCAN = 1, TK_PROP = 2, TK_INIT = 3, TK_STATS = 4, TK_BOUNDS = 5, TK_ANC = 6, TK_COUNT = 7 };

class SingleFilter {
 public:
  SingleFilter(int device, cudaStream_t stream) : device_(device), stream_(stream) {}
  ~SingleFilter();

  // bootstrap_filter(N, y, model): draws the cloud and weights it against y0
  void init(int kind, const double* params, int64_t N, double y0, const RngKey& key, uint32_t stream_id,
            StepStats* st);
  // bootstrap_filter!(x, w, y, model); params may be null (keep)
  // proposal != null: particle_filter!(x, w, y, model, proposal) with proposal = (c0, c1, c2) of x' ~ N(c0 + c1 xp, c2²)
  // (particles.jl:55-84, docs/SPEC.md §10): one-dimensional models, sorted resamplers
  void step(const double* params, double y, int resampler, StepStats* st, const double* proposal = nullptr);
  // log_likelihood(N, y, model): init + T-1 steps, one host sync at the end; stats_out[T]
  void run(int kind, const double* params, int64_t N, const double* y, int64_t T, int resampler,
           const RngKey& key, uint32_t stream_id, StepStats* stats_out, const double* proposal = nullptr);  // proposal: [T][3], row 0 unused

  // normalize(logw) / resample(w) on caller vectors (scratch use of this object; clobbers its state)
  void normalize_vector(const double* logw_host, int64_t n, StepStats* st, double* w_host);
  // n_out > 0: draw n_out ancestors from the n weights (resample(w, N), particles.jl:17)
  void resample_vector(const double* w_host, int64_t n, int resampler, const RngKey& key, uint32_t stream_id,
                       uint32_t t, uint32_t purpose, int64_t* anc_host, int64_t n_out = 0);

  void fetch(double* x_host, double* w_host, double* logw_host);
  // weighted mean / variance / quantiles of each state component, computed on the device (SPEC §8)
  void summary(const double* probs, int np, bool weighted, double* mean_out, double* var_out, double* q_out);
  int64_t fetch_ancestors(int64_t* anc_host, int64_t rows_cap);

  void set_record_ancestors(bool on) { record_anc_ = on; }
  void set_profiling(bool on) { profiling_ = on; }
  // SPEC §9: 0 = binary64, 1 = binary32 STATES (binary64 arithmetic), 2 = binary32 ARITHMETIC (§9b); takes effect at the next init / run
  void set_precision(int p) { next_prec_ = p; }
  int precision() const { return prec_; }
  void timing(double ms[TK_COUNT], int64_t launches[TK_COUNT]) const;

  int64_t N() const { return N_; }
  int64_t ld() const { return ld_; }
  int kind() const { return kind_; }
  uint32_t t() const { return t_; }
  bool live() const { return N_ > 0; }
  const double* dev_x() const { return x_[cur_]; }
  const double* dev_logw() { ensure_logw(); return logw_[cur_]; }

 private:
  void ensure_capacity(int kind, int64_t N, int64_t anc_rows);
  void set_params(int kind, const double* params);  // derive_params / derive_params_mv into dv_ / dvmv_
  void load_vector(const double* host, int64_t n, bool is_log);
  void begin_call();
  void end_call();
  void ensure_logw();  // materialise the log-weights an LG1D step keeps implicit in x
  void launch_init(double y0);
  void launch_prop(double y, int resampler);
  void launch_step(int64_t stat_index, double y, int resampler, const double* proposal = nullptr);  // one bootstrap_filter! / guided step
  void launch_scan(int64_t stat_index, bool write_cdf);
  void launch_sum(int64_t stat_index);
  unsigned long long* step_index(StepIndex& ix);
  bool legacy_multinomial(int resampler) const;  // small clouds keep the per-particle search with unsorted ancestors (SPEC §5 row 0 vs §5c)
  void mark(int klass, bool start);
  void release();

  int device_;
  cudaStream_t stream_;
  int kind_ = -1;
  int d_ = 0;
  int64_t N_ = 0, ld_ = 0, cap_N_ = 0, cap_d_ = 0, cap_stats_ = 0, cap_anc_rows_ = 0;
  int S_ = 0;
  uint64_t R_ = 0;
  uint32_t t_ = 0;
  uint32_t stream_id_ = 0;
  RngKey key_{};
  Derived dv_{};
  DerivedMV dvmv_{};   // derived block of the multivariate kinds (KIND_MVLG2..4)
  bool record_anc_ = false;
  bool profiling_ = false;
  bool from_w_ = false;
  int64_t anc_rows_ = 0;  // rows currently valid in anc_
  int cur_ = 0;
  StepStats last_{};

  double* x_[2] = {nullptr, nullptr};
  double* logw_[2] = {nullptr, nullptr};  // ping-pong with x_: the fused step reads the old weights while writing the new
  double* w_tmp_ = nullptr;
  uint64_t* cdf_ = nullptr;
  int32_t* anc_ = nullptr;
  FilterCtrl* ctrl_ = nullptr;
  unsigned long long* desc_ = nullptr;  // [2][ntiles_cap]
  int64_t desc_clean_t_ = -1;           // the descriptor half of time index desc_clean_t_ is known to be cleared (launch_scan)
  // tile index of the sorted-resampler step (sum -> bounds -> prop2)
  unsigned long long* tile_arrays_ = nullptr;  // [5][kMaxTiles]: tot, excl, incl, lexcl, cta_tot
  unsigned long long* summary_dev_ = nullptr;  // scratch of summary(): block partials, radix-select prefixes / ranks / histograms
  size_t summary_cap_ = 0;
  int64_t* util_anc_ = nullptr;                // scratch of resample_vector, kept across calls
  int64_t util_cap_ = 0;
  void* mn_arrays_ = nullptr;                  // cell index of the two-level multinomial resampler: cellC u64[cap], K, O i32[cap], part_start i32[cap + 1], ticket
  int64_t mn_cap_ = 0;
  bool mn_attr_set_ = false;
  int32_t* bound_arrays_ = nullptr;            // [2][bound_cap_]: ancestor of each propagate CTA's first particle, its tile
  int64_t bound_cap_ = 0;
  int num_sms_ = 0;
  int prec_ = 0, next_prec_ = 0;  // state storage: 0 double, 1 float (the x_ allocations are sized for double either way)
  double anc_eps_ = 1e-9;  // anc_hist_kernel: quotient estimates closer than this to an integer are decided exactly (test hook: SMCB_ANC_FORCE_EXACT)
  bool sum_done_ = false;   // sum_kernel already ran for the current weights (statistics read by a stepping caller)
  bool logw_valid_ = true;  // logw_[cur_] holds the current log-weights (false: implicit in x_[cur_] and y_cur_)
  double y_cur_ = 0.0;      // observation the current weights were computed against
  Derived dv_w_{};          // ... and the derived parameters they were computed with
  double* psum_ = nullptr;
  double* psum2_ = nullptr;
  StepStats* stats_dev_ = nullptr;
  int64_t ntiles_cap_ = 0;

  // timing
  cudaEvent_t ev_call_[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> ev_pool_;
  struct Mark {
    int klass;
    size_t e0, e1;
  };
  std::vector<Mark> marks_;
  size_t ev_used_ = 0;
  double ms_[TK_COUNT] = {0, 0, 0, 0, 0, 0, 0};
  int64_t launches_[TK_COUNT] = {0, 0, 0, 0, 0, 0, 0};
};

}  // namespace smcb
