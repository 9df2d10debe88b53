/* A plain-C client of include/smcb200.h — what a foreign-function binding (Julia's ccall) sees:
 * the header compiled as C (no C++, no CUDA headers), the library opened with dlopen, every entry point
 * resolved by name.  Built and run by tests/test_abi.py (CPU part) and tests/test_widen_guided_kalman.py
 * (GPU part); prints one "key value" line per result for the test to check against the oracle.
 *
 *   abi_client <libsmcb200.so> symbols <name>...    every name must resolve; prints "resolved <n>"
 *   abi_client <libsmcb200.so> host                 host-side helpers only (no GPU): version, state dims, simulate
 *   abi_client <libsmcb200.so> gpu                  one bootstrap filter, one guided batch, one matrix Kalman run, one smc² run of the
 *                                                   device-resident sampler
 */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/smcb200.h"

#define LOAD(name)                                                       \
  __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name);     \
  if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

#define CHECK(call)                                                                          \
  do {                                                                                       \
    int rc_ = (call);                                                                        \
    if (rc_ != SMCB_OK) {                                                                    \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, p_smcb_last_error(ctx));                 \
      return 3;                                                                              \
    }                                                                                        \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: abi_client <lib> symbols|host|gpu ...\n"); return 1; }
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }

  if (strcmp(argv[2], "symbols") == 0) {
    int n = 0;
    for (int i = 3; i < argc; ++i) {
      if (!dlsym(lib, argv[i])) { fprintf(stderr, "missing symbol %s\n", argv[i]); return 2; }
      ++n;
    }
    printf("resolved %d\n", n);
    return 0;
  }

  LOAD(smcb_version) LOAD(smcb_state_dim) LOAD(smcb_simulate) LOAD(smcb_last_error)
  const double lg[SMCB_PARAM_STRIDE] = {0.5, 1.0, 0.9, 0.8, 0.0, 1.0, 0, 0};
  enum { T = 40 };
  double xs[T], y[T];
  if (p_smcb_simulate(SMCB_LG1D, lg, T, 1998, xs, y) != SMCB_OK) return 3;

  if (strcmp(argv[2], "host") == 0) {
    printf("version %d\n", p_smcb_version());
    printf("dims %d %d %d %d\n", p_smcb_state_dim(SMCB_LG1D), p_smcb_state_dim(SMCB_SV), p_smcb_state_dim(SMCB_UCSV), p_smcb_state_dim(7));
    for (int t = 0; t < 3; ++t) printf("y%d %.17g\n", t, y[t]);
    return 0;
  }

  if (strcmp(argv[2], "gpu") != 0) return 1;
  LOAD(smcb_create) LOAD(smcb_destroy) LOAD(smcb_set_rng) LOAD(smcb_log_likelihood) LOAD(smcb_fetch_state)
  LOAD(smcb_batch_create) LOAD(smcb_batch_destroy) LOAD(smcb_batch_log_likelihood_guided) LOAD(smcb_batch_weighted_moments)
  LOAD(smcb_kalman_mv_batch_loglik) LOAD(smcb_kalman_batch_loglik)
  smcb_ctx* ctx = NULL;
  if (p_smcb_create(0, 7, &ctx) != SMCB_OK) { fprintf(stderr, "smcb_create: %s\n", p_smcb_last_error(NULL)); return 3; }

  /* log_likelihood(N, y, model): particles.jl:132-147 */
  enum { N = 2048 };
  double logZ = 0.0;
  CHECK(p_smcb_set_rng(ctx, 7, 1));
  CHECK(p_smcb_log_likelihood(ctx, SMCB_LG1D, lg, N, y, T, SMCB_SYSTEMATIC, 0, &logZ, NULL, NULL));
  double* x = (double*)malloc(sizeof(double) * N);
  CHECK(p_smcb_fetch_state(ctx, x, NULL, NULL));
  printf("pf_logZ %.17g\npf_x0 %.17g\npf_xlast %.17g\n", logZ, x[0], x[N - 1]);
  free(x);

  /* M guided filters with the locally optimal proposal: particles.jl:55-84, docs/SPEC.md §10 */
  enum { M = 3, NB = 512 };
  double params[M * SMCB_PARAM_STRIDE], prop[T * M * 3], z[M], mean[M], var[M];
  for (int m = 0; m < M; ++m) memcpy(params + m * SMCB_PARAM_STRIDE, lg, sizeof lg);
  const double s2 = 1.0 / (1.0 / lg[2] + lg[1] * lg[1] / lg[3]);
  for (int t = 0; t < T; ++t)
    for (int m = 0; m < M; ++m) {
      double* c = prop + (t * M + m) * 3;
      c[0] = s2 * lg[1] * y[t] / lg[3];
      c[1] = s2 * lg[0] / lg[2];
      c[2] = sqrt(s2);
    }
  smcb_batch* b = NULL;
  CHECK(p_smcb_batch_create(ctx, SMCB_LG1D, M, NB, &b));
  CHECK(p_smcb_set_rng(ctx, 7, 2));
  CHECK(p_smcb_batch_log_likelihood_guided(b, params, NULL, y, T, SMCB_SYSTEMATIC, 10, prop, z));
  CHECK(p_smcb_batch_weighted_moments(b, mean, var));
  for (int m = 0; m < M; ++m) printf("guided_logZ%d %.17g\nguided_mean%d %.17g\nguided_var%d %.17g\n", m, z[m], m, mean[m], m, var[m]);
  CHECK(p_smcb_batch_destroy(b));

  /* Kalman: the scalar model as a d = 1 block and through the scalar entry point; kalman_filter.jl:3-70 */
  const double blk[6] = {lg[0], lg[1], lg[2], lg[3], lg[4], lg[5]}; /* A, B, Q, R, x0, S0 for d = 1 */
  double ll_mv = 0.0, ll_sc = 0.0;
  CHECK(p_smcb_kalman_mv_batch_loglik(ctx, 1, blk, NULL, 1, y, T, 0, &ll_mv, NULL, NULL));
  CHECK(p_smcb_kalman_batch_loglik(ctx, lg, NULL, 1, y, T, 0, &ll_sc, NULL, NULL));
  printf("kalman_mv %.17g\nkalman_scalar %.17g\n", ll_mv, ll_sc);

  /* the device-resident θ-level sampler (mutable struct SMC, smc², smc²!: smc_samplers.jl:5-59,288-340): lg_mod(θ) with the README's
   * lg_prior, 32 θ-particles × 128 state particles, chain 2.  θ0 is a fixed grid inside the prior's support (the host language draws
   * it; any values do for the check: the test feeds the same θ0 to the Python-driven sampler and to the oracle). */
  LOAD(smcb_sampler_create) LOAD(smcb_sampler_destroy) LOAD(smcb_sampler_set_data) LOAD(smcb_sampler_smc2_init)
  LOAD(smcb_sampler_smc2_step) LOAD(smcb_sampler_get)
  enum { MS = 32, NS = 128, DT = 3 };
  smcb_sampler_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.kind = SMCB_LG1D; cfg.d_theta = DT; cfg.N = NS; cfg.M = MS; cfg.chain = 2;
  cfg.resampler = SMCB_SYSTEMATIC; cfg.theta_resampler = SMCB_MULTINOMIAL;
  cfg.ess_threshold = 0.5; cfg.min_ar = -1.0; cfg.seed = 11;
  /* TruncatedNormal(0, 1, -1, 1), LogNormal(), LogNormal(): rows (family, p0, p1, lo, hi, c0, c1, 0) */
  const double tn[8] = {3, 0.0, 1.0, -1.0, 1.0, 0.0, log(erf(1.0 / sqrt(2.0))), 0}, ln_[8] = {1, 0.0, 1.0, 0, 0, 0.0, 0, 0};
  memcpy(cfg.prior[0], tn, sizeof tn); memcpy(cfg.prior[1], ln_, sizeof ln_); memcpy(cfg.prior[2], ln_, sizeof ln_);
  const int32_t src[8] = {0, -1, 1, 2, -1, -1, -1, -1};            /* LinearGaussian(θ1, 1.0, θ2, θ3, 0.0), σ0 = 1 */
  const double cst[8] = {0, 1.0, 0, 0, 0.0, 1.0, 0, 0};
  memcpy(cfg.map_src, src, sizeof src); memcpy(cfg.map_const, cst, sizeof cst);
  double theta0[MS * DT], th[MS * DT], om[MS], lz[MS], ess = 0, acc = 0;
  for (int m = 0; m < MS; ++m) {
    theta0[m * DT + 0] = -0.9 + 1.8 * (m + 0.5) / MS;
    theta0[m * DT + 1] = 0.4 + 0.05 * ((m * 7) % MS);
    theta0[m * DT + 2] = 0.5 + 0.04 * ((m * 11) % MS);
  }
  smcb_sampler* sp = NULL;
  CHECK(p_smcb_sampler_create(ctx, &cfg, theta0, &sp));
  CHECK(p_smcb_sampler_set_data(sp, y, T));
  CHECK(p_smcb_sampler_smc2_init(sp));
  int nrej = 0;
  for (int t = 1; t < T; ++t) {
    int rj = 0;
    CHECK(p_smcb_sampler_smc2_step(sp, t, &ess, &rj));
    nrej += rj;
  }
  int64_t Nfin = 0;
  CHECK(p_smcb_sampler_get(sp, th, om, lz, &ess, &acc, &Nfin));
  double sth = 0, slz = 0, som = 0;
  for (int m = 0; m < MS; ++m) { slz += lz[m]; som += om[m]; for (int k = 0; k < DT; ++k) sth += th[m * DT + k]; }
  printf("smc2_rejuvenations %d\nsmc2_ess %.17g\nsmc2_theta_sum %.17g\nsmc2_logZ_sum %.17g\nsmc2_omega_sum %.17g\nsmc2_N %lld\n", nrej, ess, sth, slz, som,
         (long long)Nfin);
  CHECK(p_smcb_sampler_destroy(sp));
  CHECK(p_smcb_destroy(ctx));
  return 0;
}
