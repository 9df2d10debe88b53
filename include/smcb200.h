/* libsmcb200 — C ABI of the B200-native particle-filter hot path.
 *
 * Drop-in boundary for SequentialMonteCarlo.jl's particle-filter path (the reference is pure Julia
 * and has no FFI layer of its own; each entry point below names the Julia function it replaces,
 * file:line under /root/reference).  Julia binds these with `ccall((:sym, "libsmcb200"), Cint, ...)`
 * — see INTEGRATION.md; the Python host mirror binds them with ctypes
 * (sequential_monte_carlo_b200/_lib.py).
 *
 * Conventions
 *   - every function returns 0 on success or a negative smcb_status; the message of the last
 *     failure is smcb_last_error(ctx) (ctx may be NULL for a failure of smcb_create);
 *   - all pointers are HOST pointers owned by the caller, copied in/out synchronously, never
 *     retained, unless the parameter name ends in `_dev` (device pointers on the context's GPU);
 *   - all indices are 0-based (the Julia shim adds 1 to ancestors);
 *   - state clouds are SoA: x[k*N + i] is component k of particle i (UCSV: k=0 x, 1 log σε, 2 log ση);
 *   - a model is (kind, 8 doubles): see smcb_model_kind;
 *   - randomness is the counter-based Philox stream of docs/SPEC.md: (seed, epoch, stream) name a
 *     sweep; the context hands out a fresh epoch to every init / log_likelihood call, or the caller
 *     pins it with smcb_set_rng for reproducibility;
 *   - there is no CPU fallback: without a CUDA device smcb_create fails with SMCB_ERR_CUDA.
 *   - a context is not thread-safe; distinct contexts may be used from distinct threads.
 */
#ifndef SMCB200_H
#define SMCB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smcb_ctx smcb_ctx;     /* one GPU, one stream, one single-filter slot */
typedef struct smcb_batch smcb_batch; /* M independent filters of N particles (θ-level samplers) */
typedef struct smcb_sampler smcb_sampler; /* a device-resident SMC² / density-tempered sampler (mutable struct SMC) */

typedef enum {
  SMCB_OK = 0,
  SMCB_ERR_BAD_ARG = -1,
  SMCB_ERR_CUDA = -2,
  SMCB_ERR_OOM = -3,
  SMCB_ERR_STATE = -4, /* e.g. step before init */
  SMCB_ERR_UNSUPPORTED = -5,
  SMCB_ERR_NCCL = -6 /* NCCL missing or a collective failed (multi-GPU entry points only) */
} smcb_status;

/* params[8] per model (unused trailing entries ignored):
 *   LG1D  A, B, Q, R, x0, σ0   (Q, R, σ0 are variances)   state_space_models.jl:46-58,74-109
 *         unobserved_components(σε,ση,x0) is LG1D(1,1,σε,ση,x0,σε)              :119-128
 *   SV    μ, ρ, σ              x1~N(μ,σ²/(1-ρ²)), x'~N(μ+ρ(x-μ),σ), y~N(0,exp(x/2))  (absent upstream)
 *   UCSV  γε, γη, x0, logσε0, logση0                       state_space_models.jl:215-259 */
typedef enum { SMCB_LG1D = 0, SMCB_SV = 1, SMCB_UCSV = 2, SMCB_MVLG2 = 3, SMCB_MVLG3 = 4, SMCB_MVLG4 = 5 } smcb_model_kind;
/* SMCB_MVLG2..4: MultivariateLinearGaussian with a d = 2, 3, 4 dimensional state and a scalar observation
 *   x' ~ MvNormal(A x, Q), y ~ Normal(B x, R), x1 ~ MvNormal(x0, Σ0)        state_space_models.jl:137-189, hodrick_prescott :193-202
 * `params` is then the block of smcb_kalman_mv_* below (3d² + 2d + 1 doubles, row-major: A, B, Q, R, x0, Σ0), NOT 8 doubles.  Q and Σ0
 * may be positive SEMI-definite (hodrick_prescott's Q has a zero pivot); R is a variance as in kalman_filter.jl:16.  Engine: the large-N
 * single filter (smcb_bootstrap_init / _step, smcb_log_likelihood, smcb_fetch_state, smcb_weighted_summary, smcb_simulate), binary64
 * arithmetic; the batched / θ-level entry points answer SMCB_ERR_UNSUPPORTED (IBIS over the matrix Kalman filter covers them). */

/* MULTINOMIAL is the reference's distribution (i.i.d., unsorted; particles.jl:17-19);
 * STRATIFIED / SYSTEMATIC give sorted ancestors and are the bandwidth-optimal modes. */
typedef enum { SMCB_MULTINOMIAL = 0, SMCB_STRATIFIED = 1, SMCB_SYSTEMATIC = 2 } smcb_resampler;

#define SMCB_PARAM_STRIDE 8

/* ------------------------------------------------------------------ context */
int smcb_create(int device, uint64_t seed, smcb_ctx** out);
int smcb_destroy(smcb_ctx* ctx);
const char* smcb_last_error(const smcb_ctx* ctx);
int smcb_version(void);
int smcb_state_dim(int kind);
/* next sweep uses (seed, epoch); later sweeps epoch+1, ... */
int smcb_set_rng(smcb_ctx* ctx, uint64_t seed, uint32_t epoch);
int smcb_get_epoch(const smcb_ctx* ctx, uint32_t* next_epoch);
/* keep every step's ancestor vector of the single filter (for parity tests; costs 4 B/particle/step) */
int smcb_record_ancestors(smcb_ctx* ctx, int enable);
/* per-kernel CUDA-event timing of the single filter (bench.py roofline); off by default */
int smcb_set_profiling(smcb_ctx* ctx, int enable);
/* State storage of the single filter from the next smcb_bootstrap_init / smcb_log_likelihood on:
 * 0 = binary64 (default), 1 = binary32 states (docs/SPEC.md §9: every state component is rounded to
 * binary32 where it is stored, arithmetic stays binary64), 2 = binary32 ARITHMETIC (docs/SPEC.md §9b: float Philox
 * normals four per block, float model arithmetic and log-weights, float4 loads / stores; the CDF stays uint64).
 * Tiers 1 and 2: sorted resamplers at any N, multinomial for N > 8192; tier 2 has no guided step.  The reference is
 * Float64 only (src/particles.jl:87-147); these are the north star's fp32 tier.  Host-facing arrays stay double. */
int smcb_set_precision(smcb_ctx* ctx, int precision);
/* device time of the last sweep / step, and per-kernel-class totals when profiling is on:
 * ms[0] whole call, ms[1] sum/scan (quantise + prefix sums + Σe, Σe²), ms[2] gather+propagate+weight
 * (plus the ancestor search on the multinomial path), ms[3] init, ms[4] stats-only, ms[5] window bounds,
 * ms[6] ancestor search of the sorted resamplers; launches[0..6] the matching launch counts. */
int smcb_get_timing(const smcb_ctx* ctx, double ms[7], int64_t launches[7]);
int smcb_synchronize(smcb_ctx* ctx);
/* page-locked host buffers for fast fetches (PCIe D2H at link rate instead of pageable staging);
 * the Python / Julia shims keep one per result array and wrap it as an array view */
int smcb_alloc_pinned(smcb_ctx* ctx, int64_t bytes, void** out);
int smcb_free_pinned(smcb_ctx* ctx, void* ptr);

/* ------------------------------------------------------------------ utilities */
/* normalize(logw) -> (logμ, w, ess)                                   particles.jl:5-15
 * w may be NULL. */
int smcb_normalize(smcb_ctx* ctx, const double* logw, int64_t n, double* logmu, double* w, double* ess);
/* resample(w, n) -> ancestors                                          particles.jl:17-19
 * draws come from Philox (seed, epoch) of the context at (stream, t, purpose). */
int smcb_resample(smcb_ctx* ctx, const double* w, int64_t n, int resampler, uint32_t stream, uint32_t t,
                  uint32_t purpose, int64_t* ancestors);

/* resample(w, N) with N != length(w): n_out ancestors (0-based indices into w) from the n weights                 particles.jl:17 */
int smcb_resample_n(smcb_ctx* ctx, const double* w, int64_t n, int64_t n_out, int resampler, uint32_t stream, uint32_t t,
                    uint32_t purpose, int64_t* ancestors);

/* ------------------------------------------------------------------ one filter (large N) */
/* bootstrap_filter(N, y, model) -> (x, w, logμ)                        particles.jl:87-105
 * the cloud stays on the device; read it back with smcb_fetch_state. */
int smcb_bootstrap_init(smcb_ctx* ctx, int kind, const double* params, int64_t N, double y, uint32_t stream,
                        double* logmu, double* ess);
/* bootstrap_filter!(states, weights, y, model) -> (logμ, w, ess)       particles.jl:107-129
 * params may be NULL (keep the model given to init). */
int smcb_bootstrap_step(smcb_ctx* ctx, const double* params, double y, int resampler, double* logmu,
                        double* ess);
/* log_likelihood(N, y, model) -> (x, w, logZ)                          particles.jl:132-147
 * one call for the whole series; logmu_out / ess_out (length T) may be NULL. */
int smcb_log_likelihood(smcb_ctx* ctx, int kind, const double* params, int64_t N, const double* y, int64_t T,
                        int resampler, uint32_t stream, double* logZ, double* logmu_out, double* ess_out);
/* particle_filter!(states, weights, y, model, proposal) -> (logμ, w, ess)          particles.jl:55-84
 * and the loop of examples/inflation_example.jl:164-171 with a proposal, for ONE large-N filter (docs/SPEC.md §10):
 * proposal = (c0, c1, c2) of x' ~ N(c0 + c1·xp, c2²) for this step ([T][3] for a whole series, row 0 not read — the
 * initial step is smcb_bootstrap_init's).  LG1D, SV: both storage tiers of smcb_set_precision.  UCSV (docs/SPEC.md §10b): the triple
 * is (kappa, 0, 1), kappa in [0, 1] tempers the conditionally optimal move of the trend (0 = bootstrap, 1 = p(x' | x, le, ln', y));
 * the log-volatilities move by the transition; binary64 states.  Sorted resamplers; multinomial resampling ->
 * SMCB_ERR_UNSUPPORTED (the batched entry points below take multinomial, N <= 8192). */
int smcb_guided_step(smcb_ctx* ctx, const double* params, double y, int resampler, const double* proposal, double* logmu,
                     double* ess);
int smcb_guided_log_likelihood(smcb_ctx* ctx, int kind, const double* params, int64_t N, const double* y, int64_t T,
                               int resampler, uint32_t stream, const double* proposal, double* logZ, double* logmu_out,
                               double* ess_out);
/* x [d*N] SoA, w [N] normalised weights, logw [N] unnormalised log-weights; any may be NULL */
int smcb_fetch_state(smcb_ctx* ctx, double* x, double* w, double* logw);
/* Summaries of the current cloud computed ON THE DEVICE — what the reference's per-step
 * `quantile(x, [0.25,0.5,0.75])` (README.md:41,51) and `quantile(x, weights(w), p)` / `var(x, weights(w))`
 * (examples/inflation_example.jl:39-55) need, without the 16 B/particle read-back.  weighted != 0 uses the
 * current weights (their fixed-point values, docs/SPEC.md §8), else every particle counts once.
 * mean [d], var [d] (population variance), quantiles [d][nprobs] (lower empirical quantile: the smallest
 * x whose cumulative weight exceeds p); any output may be NULL, nprobs <= 16. */
int smcb_weighted_summary(smcb_ctx* ctx, const double* probs, int nprobs, int weighted, double* mean, double* var,
                          double* quantiles);
/* ancestors of steps t = 1 .. T-1 ([T-1][N], row t-1 = step t) if recording was on, else the
 * last step's only when stepping with bootstrap_step (rows = 1); 0 rows when recording is off.
 * rows_cap = rows the buffer can take. */
int smcb_fetch_ancestors(smcb_ctx* ctx, int64_t* ancestors, int64_t rows_cap, int64_t* rows_out);
/* device views of the current cloud (valid until the next call on ctx) */
int smcb_device_state(smcb_ctx* ctx, const double** x_dev, const double** logw_dev, int64_t* ld);

/* ------------------------------------------------------------------ M filters at once (θ-level) */
/* Replaces the Threads.@threads / serial loops over θ-particles:
 *   smc_samplers.jl:112-121 (rejuvenate!), :174-180 (exchange!), :223-229 (density_tempered),
 *   :289-293 (smc²), :325-335 (smc²!).
 * Filter m uses Philox stream (stream0 + m): give the GLOBAL θ index so that results do not depend
 * on how θ is sharded across GPUs.  params is [M][8]; active[m]==0 skips filter m (out-of-support
 * proposals, :116) and reports logμ = logZ = -inf.  active may be NULL (all active). */
int smcb_batch_create(smcb_ctx* ctx, int kind, int64_t M, int64_t N, smcb_batch** out);
int smcb_batch_destroy(smcb_batch* b);
int smcb_batch_init(smcb_batch* b, const double* params, const uint8_t* active, double y, uint32_t stream0,
                    double* logmu, double* ess);
int smcb_batch_step(smcb_batch* b, const double* params, double y, int resampler, double* logmu, double* ess);
/* whole series for every θ in ONE launch; the final clouds stay in b */
int smcb_batch_log_likelihood(smcb_batch* b, const double* params, const uint8_t* active, const double* y,
                              int64_t T, int resampler, uint32_t stream0, double* logZ);
/* Guided filters — particle_filter / particle_filter! with a proposal      particles.jl:28-84 (docs/SPEC.md §10)
 * The reference's proposal is a Julia closure (model, xp) -> distribution; the device evaluates the affine-Gaussian
 * family  x' ~ N(c0 + c1·xp, c2²)  (c2 > 0; a closure over y_t picks the coefficients per step, e.g. the locally
 * optimal proposal of an LG model), for the one-dimensional models (LG1D, SV); for UCSV the triple is (kappa, 0, 1), the tempered
 * optimal trend move of docs/SPEC.md §10b.  Every step t >= 1 draws x' from the
 * proposal and weights by  logpdf(observation(x'), y) + logpdf(transition(xp), x') - logpdf(proposal(xp), x')  (:73-78);
 * the initial step draws from initial_dist and weights by the observation density (:40-42; the reference's
 * "+ logpdf(initial_dist, x)" at :44 lacks its "- logpdf(proposal)" partner, commented out at :45 — ruled a defect).
 * proposal: [M][3] for one step, [T][M][3] for a whole series (row 0 is not read). */
int smcb_batch_step_guided(smcb_batch* b, const double* params, double y, int resampler, const double* proposal,
                           double* logmu, double* ess);
int smcb_batch_log_likelihood_guided(smcb_batch* b, const double* params, const uint8_t* active, const double* y,
                                     int64_t T, int resampler, uint32_t stream0, const double* proposal, double* logZ);
/* θ-resample: slot m takes a deep copy of slot parents[m] (x, log-weights, rng identity)
 *   smc_samplers.jl:74-84 (with the w-permutation / aliasing defects D3, D4 fixed) */
int smcb_batch_gather(smcb_batch* b, const int32_t* parents);
/* MH accept: slots with accept[m]!=0 take the cloud of the same slot of `proposal`
 *   smc_samplers.jl:130-133 */
int smcb_batch_accept(smcb_batch* current, const smcb_batch* proposal, const uint8_t* accept);
/* x [M][d][N], w [M][N] normalised, logw [M][N]; any may be NULL */
int smcb_batch_fetch(smcb_batch* b, double* x, double* w, double* logw);
/* mean [M][d]: the weighted state mean w[m]' * x[m] of every θ-particle's cloud, computed on the device — what
 * estimated_trend(smc) and quantile(smc, p) integrate over θ (plotting_utils.jl:116-124,140-157) */
int smcb_batch_weighted_mean(smcb_batch* b, double* mean);
/* mean, var [M][d] (either may be NULL): mean as above and the population variance Σ w_i (x_i - mean)² of every cloud
 * under its own weights — var(x, weights(w)) per θ-particle (examples/inflation_example.jl:46) */
int smcb_batch_weighted_moments(smcb_batch* b, double* mean, double* var);
/* quantiles [M][d][nprobs] (nprobs <= 16): the lower empirical quantiles of every θ-particle's cloud under its own
 * weights (weighted != 0) or counting every particle once, computed on the device (docs/SPEC.md §8) — the per-θ
 * bands of get_quantiles_uc / get_quantiles_ucsv, examples/inflation_example.jl:39-55,241-253 */
int smcb_batch_weighted_quantiles(smcb_batch* b, const double* probs, int nprobs, int weighted, double* quantiles);
/* cross-GPU moves of whole clouds (θ-resample across ranks): pack slot m into / unpack from a
 * device buffer of smcb_batch_cloud_bytes(b) bytes that the caller sends with NCCL / P2P */
int64_t smcb_batch_cloud_bytes(const smcb_batch* b);
int smcb_batch_pack(smcb_batch* b, const int32_t* slots, int64_t n, void* buf_dev);
int smcb_batch_unpack(smcb_batch* b, const int32_t* slots, int64_t n, const void* buf_dev);
int smcb_batch_get_timing(const smcb_batch* b, double* ms_last_call, int64_t* launches_total);

/* ------------------------------------------------------------------ multi-GPU: one process per GPU, θ sharded
 * The reference's only parallelism is Threads.@threads over θ (smc_samplers.jl:112,174,223); here rank r of G owns the
 * θ-particles [r·M/G, (r+1)·M/G) with their whole state clouds.  NCCL is bound at run time (dlopen of libnccl.so.2, or
 * $SMCB_NCCL_LIB); single-GPU users never need it.  Rank 0 calls smcb_comm_unique_id, the host language carries the 128
 * bytes to the other ranks (MPI.bcast, a file, torch.distributed ...), then EVERY rank calls smcb_comm_init (collective).
 * Samplers created from the context afterwards are sharded; their results do not depend on G. */
int smcb_comm_unique_id(uint8_t id[128]);
int smcb_comm_init(smcb_ctx* ctx, int rank, int nranks, const uint8_t id[128]);
int smcb_comm_rank(const smcb_ctx* ctx, int* rank, int* nranks);
/* all[r·n_local + i] = rank r's local[i] on every rank (host vectors; ncclAllGather on the context's stream) */
int smcb_comm_all_gather(smcb_ctx* ctx, const double* local, int64_t n_local, double* all);
int smcb_comm_destroy(smcb_ctx* ctx);
/* host-only (no GPU): how a whole-series sweep of M filters of N particles over `steps` observations is scheduled on a GPU that keeps
 * `slots` CTAs of `threads` threads resident on `num_sms` SMs: *chunk = 0 — one CTA per θ-particle for the whole series (the loops over θ
 * of smc_samplers.jl:112-121,223-229 as one launch); *chunk = K > 0 — a persistent grid claims (chunk of K steps, θ) units in order
 * (used when whole series would leave > 3 % of the launch idle, and — masked != 0 — for every sweep that carries an `active` mask over
 * more θ than resident CTAs: how many of them run is known only on the device; results are bit-identical either way).  SMCB_BATCH_CHUNK in the
 * environment overrides the plan (0, or a chunk length). */
int smcb_batch_chunk_plan(int64_t M, int64_t N, int64_t steps, int threads, int64_t slots, int num_sms, int masked, int64_t* chunk);
/* host-only (no GPU): who sends which cloud where after a θ-resample with (sorted) parents[M].  local_parents [M/G]:
 * rank-local parent slot (the slot itself where the parent is remote); send_* / recv_*: the clouds this rank sends /
 * receives as (peer rank, local slot), grouped by peer, increasing global slot — both sides enumerate the same order.
 * send_* / recv_* may be NULL (counts only); capacity M entries for send_* (one cloud may have children in every slot of the
 * other ranks), M/G for recv_*. */
int smcb_exchange_plan(const int32_t* parents, int64_t M, int rank, int nranks, int32_t* local_parents, int32_t* send_peer,
                       int32_t* send_slot, int64_t* n_send, int32_t* recv_peer, int32_t* recv_slot, int64_t* n_recv);

/* ------------------------------------------------------------------ device-resident θ-level samplers
 * mutable struct SMC + SMC(N, M, model, prior, chain, ess_threshold, min_ar)     smc_samplers.jl:5-59
 * θ, ω, logZ, the log-prior and the parameter blocks of all M θ-particles live on the GPU (replicated on every rank of
 * the communicator); the M inner particle filters are sharded.  Per smc²! the host enqueues one batched filter step, one
 * in-place all-gather of M/G doubles and one single-CTA kernel (logω += logμ, normalise, ESS) and reads back 64 bytes.
 *   prior: product of d_theta univariate laws, row k = (family, p0, p1, lo, hi, c0, c1, 0):
 *     0 Normal(μ=p0, σ=p1), c0 = log σ        1 LogNormal(μ=p0, σ=p1), c0 = log σ
 *     2 Uniform(lo, hi), c0 = -log(hi - lo)    3 TruncatedNormal(μ=p0, σ=p1, lo, hi), c0 = log σ, c1 = log mass of [lo, hi]
 *   model: θ -> StateSpaceModel as a selection map, params[k] = map_src[k] >= 0 ? θ[map_src[k]] : map_const[k]
 *     (README.md:75-78 lg_mod: src = {0,-1,1,2,-1,-1,..}, const = {·,1,·,·,0,1}; examples/inflation_example.jl:229-232)
 *   theta0 [M][d_theta]: the prior draws θ = map(m -> rand(prior), 1:M) (:38), drawn by the host language.
 * The arithmetic of the θ level (covariance, Cholesky, proposal, accept) is frozen in docs/SPEC.md §11. */
typedef struct {
  int32_t kind;            /* smcb_model_kind of model(θ) */
  int32_t d_theta;         /* 1..8 */
  int64_t N;               /* state particles per θ (<= 8192; exchange! doubles it) */
  int64_t M;               /* θ-particles (2..16384, divisible by the number of ranks) */
  int32_t chain;           /* PMMH moves per rejuvenation */
  int32_t resampler;       /* smcb_resampler of the inner filters */
  int32_t theta_resampler; /* smcb_resampler of resample!(smc) */
  int32_t reserved;
  double ess_threshold;    /* ess_min = M · ess_threshold                       :46 */
  double min_ar;           /* acc_threshold of exchange! (-1.0 disables it)     :35 */
  uint64_t seed;
  double prior[8][8];
  int32_t map_src[8];
  double map_const[8];
} smcb_sampler_config;
int smcb_sampler_create(smcb_ctx* ctx, const smcb_sampler_config* cfg, const double* theta0, smcb_sampler** out);
int smcb_sampler_destroy(smcb_sampler* s);
/* the observations y[0..T) the following calls refer to (copied to the device once) */
int smcb_sampler_set_data(smcb_sampler* s, const double* y, int64_t T);
/* smc²(smc, y)                                                              smc_samplers.jl:288-301 */
int smcb_sampler_smc2_init(smcb_sampler* s);
/* smc²!(smc, y, t): assimilates y[t] (0-based; Julia's t is t+1); resample! / rejuvenate! / exchange! first when
 * ess < ess_min.  ess, rejuvenated may be NULL.                              smc_samplers.jl:308-340 */
int smcb_sampler_smc2_step(smcb_sampler* s, int64_t t, double* ess, int* rejuvenated);
/* density_tempered(smc, y); schedule [cap][3] receives (ξ, ess, acceptance ratio of the stage's rejuvenation or -1 when the
 * stage did not resample) of every stage (may be NULL)                         smc_samplers.jl:222-281 */
int smcb_sampler_density_tempered(smcb_sampler* s, double* schedule, int cap, int* n_stages);
/* public fields of the struct: θ [M][d_theta], ω [M], logZ [M], ess, acc_ratio, N; any may be NULL   :5-27 */
int smcb_sampler_get(smcb_sampler* s, double* theta, double* omega, double* logZ, double* ess, double* acc_ratio, int64_t* N);
/* smc.x, smc.w: a view of this rank's M/G live clouds for smcb_batch_fetch / _weighted_* (owned by the sampler; valid
 * until the next call that may run exchange!) */
int smcb_sampler_clouds(smcb_sampler* s, smcb_batch** clouds);
/* ms[0..3]: device time in the inner filters / all-gathers / cloud moves (θ-resample exchange + accept copies) / θ-level
 * kernels (CUDA events, only while profiling is on); ms[4]: CUDA-event time on the context's stream from the start of the
 * last smc² / density_tempered call to the last completed step (always measured); counts: whole-series sweeps, smc²! steps, rejuvenations, clouds
 * received from other ranks, particle-updates of this rank, stream synchronisations, kernel launches, θ-resamples */
int smcb_sampler_set_profiling(smcb_sampler* s, int enable);
int smcb_sampler_stats(smcb_sampler* s, double ms[8], int64_t counts[8]);
/* host-only (no GPU): Σ of random_walk_kernel(θ) for θ [M][d] (d×d row-major; d = 1: the σ the reference uses as a standard
 * deviation) and the lower Cholesky factor of scale·A — the fixed-order arithmetic of docs/SPEC.md §11, exposed so that a
 * host-language sampler (arbitrary model / prior closures) proposes exactly what the device sampler proposes.
 *                                                                           smc_samplers.jl:87-101 */
int smcb_random_walk_sigma(const double* theta, int64_t M, int d, double* sigma);
int smcb_cholesky_lower(const double* A, int d, double scale, double* L);

/* Kalman filter for LG1D, M models at once                             kalman_filter.jl:29-70
 * params [M][8]; x, sigma [M] in/out; loglik [M] out (step log-likelihoods) */
int smcb_kalman_batch_step(smcb_ctx* ctx, const double* params, int64_t M, double y, double* x, double* sigma,
                           double* loglik);
/* log_likelihood(y, model::LinearModel); matched_init=1 skips the first predict */
int smcb_kalman_batch_loglik(smcb_ctx* ctx, const double* params, const uint8_t* active, int64_t M,
                             const double* y, int64_t T, int matched_init, double* loglik, double* x,
                             double* sigma);

/* Matrix Kalman filter, M multivariate LinearModels with a scalar observation at once    kalman_filter.jl:3-27,55-70
 * (MultivariateLinearGaussian, hodrick_prescott: state_space_models.jl:137-202).  State dimension
 * 1 <= d <= 4.  models [M][3d²+2d+1], row-major: A[d][d], B[d], Q[d][d], R, x0[d], Σ0[d][d].
 * step: x [M][d], sigma [M][d][d] in/out, loglik [M] out.  loglik: starts from (x0, Σ0) of every model, x / sigma
 * receive the final filtered moments (may be NULL); matched_init=1 skips the first predict as above. */
int smcb_kalman_mv_batch_step(smcb_ctx* ctx, int d, const double* models, int64_t M, double y, double* x, double* sigma,
                              double* loglik);
int smcb_kalman_mv_batch_loglik(smcb_ctx* ctx, int d, const double* models, const uint8_t* active, int64_t M,
                                const double* y, int64_t T, int matched_init, double* loglik, double* x, double* sigma);

/* ------------------------------------------------------------------ host-side helpers (no GPU) */
/* The θ-level samplers draw priors, MH proposals and accept uniforms on the host from the same
 * Philox stream and the same deterministic Box-Muller as the device (docs/SPEC.md §2-§3), so a run
 * is reproducible bit-for-bit across GPU counts.  These run on the CPU and need no context. */
int smcb_rng_normals(uint64_t seed, uint32_t epoch, uint32_t stream, uint32_t t, uint32_t purpose, uint32_t comp,
                     int64_t n, double* out);
int smcb_rng_uniforms64(uint64_t seed, uint32_t epoch, uint32_t stream, uint32_t t, uint32_t purpose, int64_t n,
                        uint64_t* out);
/* simulate(model, T) -> (x [d][T], y [T])                               state_space_models.jl:11-28 */
int smcb_simulate(int kind, const double* params, int64_t T, uint64_t seed, double* x, double* y);
/* device self-test of the deterministic math (parity tests): fn 0 exp, 1 log, 2 sincos2pi (out0=sin,
 * out1=cos), 3 quantise with shift S = (int)aux -> out0 holds uint64 bit patterns; 4-7 the same four in the binary32
 * arithmetic of docs/SPEC.md §9b (inputs rounded to binary32 first, outputs widened exactly) */
int smcb_selftest_math(smcb_ctx* ctx, int fn, const double* in, int64_t n, double aux, double* out0, double* out1);

#ifdef __cplusplus
}
#endif
#endif /* SMCB200_H */
