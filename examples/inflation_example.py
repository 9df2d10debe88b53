"""The reference's worked example (/root/reference/examples/inflation_example.jl) on this library, same flow and names:

  1. the unobserved-components model `uc_mod(θ) = StateSpaceModel(UC(θ...), (1,1))` with `uc_prior`, SMC² over the series,
     and after every observation the ω-mixture of the per-θ quartile bands of the trend and the cycle and the variance
     of the trend (`get_quantiles_uc`, :39-55) — computed on the device, no cloud is read back;
  2. a particle filter at the posterior mean (`get_latent_states_uc`, :145-172) through `particle_filter` /
     `particle_filter!` — bootstrap as in the reference, or guided (UC: the locally optimal proposal; UC-SV: the optimal trend move);
  3. the UC-SV model `ucsv_mod(θ) = StateSpaceModel(UCSV(θ[1],θ[2],(θ[3],θ[4])), (3,1))` with `ucsv_prior` (:229-253).

The reference downloads the PCE inflation series from FRED (:12-19); there is no network here, so the series is
simulated from the UC-SV model at the parameters the reference's figures show (seed 1998 as in :57,255).  Plots are
out of scope; the bands are returned as arrays.

    python examples/inflation_example.py [--N 1024] [--M 512] [--T 241]        (needs a CUDA device)
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequential_monte_carlo_b200 as smc  # noqa: E402

QUARTILES = (0.25, 0.5, 0.75)


def uc_mod(θ):                                                    # :28-31
    return smc.StateSpaceModel(smc.UC(θ[0], θ[1], θ[2]), (1, 1))


uc_prior = smc.product_distribution([smc.Normal(3.0, 2.0), smc.Uniform(0.0, 4.0), smc.Uniform(0.0, 4.0)])   # :33-37


def ucsv_mod(θ):                                                  # :229-232
    return smc.StateSpaceModel(smc.UCSV(θ[0], θ[1], (θ[2], θ[3])), (3, 1))


ucsv_prior = smc.product_distribution([smc.Uniform(0.0, 1.0), smc.Normal(3.0, 2.0), smc.Uniform(0.0, 2.0), smc.Uniform(0.0, 2.0)])   # :234-239


def get_quantiles(s, yt):
    """get_quantiles_uc / get_quantiles_ucsv (:39-55, :241-253): (trend quartiles, cycle quartiles, trend variance), each the
    ω-mixture over the θ-particles of the per-cloud summaries"""
    xq, cq = smc.get_quantiles(s, yt, QUARTILES)
    _, var = smc.state_variances(s)
    return xq, cq, float(np.sum(s.ω * var[:, 0]))


def run_smc2(model, prior, y, N, M, chain, ess_threshold=0.5, *, seed=1998, ctx=None, verbose=False):
    """the loop of :57-75 / :255-270: smc², then for every t the bands of the current clouds and smc²!"""
    T = len(y)
    s = smc.SMC(N, M, model, prior, chain, ess_threshold, seed=seed, ctx=ctx)
    xqs, cqs, variances = np.zeros((T, 3)), np.zeros((T, 3)), np.zeros(T)
    smc.smc2(s, y)
    for t in range(1, T):
        xqs[t - 1], cqs[t - 1], variances[t - 1] = get_quantiles(s, y[t - 1])
        smc.smc2_step(s, y, t, verbose=verbose)
    xqs[T - 1], cqs[T - 1], variances[T - 1] = get_quantiles(s, y[T - 1])
    return s, xqs, cqs, variances


def get_latent_states(N, y, model, proposal=None, *, ctx=None):
    """get_latent_states_uc (:145-172): quartile bands of the trend and the cycle from one particle filter at fixed θ"""
    T = len(y)
    xq, cq, variances = np.zeros((T, 3)), np.zeros((T, 3)), np.zeros(T)
    x, w, _ = smc.particle_filter(N, y[0], model, proposal, ctx=ctx)
    for t in range(T):
        if t > 0:
            _, w, _ = smc.particle_filter_(x, w, y[t], model, proposal, resampler="systematic")
        q, qc = np.asarray(smc.quantile(x, w, QUARTILES)), np.asarray(smc.quantile(x, w, [1 - p for p in QUARTILES]))
        xq[t] = q if q.ndim == 1 else q[:, 0]                               # UCSV: the trend is the first state component
        cq[t] = y[t] - (qc if qc.ndim == 1 else qc[:, 0])                   # quantile(y[t] .- x, p) = y[t] - quantile(x, 1-p)
        variances[t] = np.ravel(smc.weighted_mean_var(x, w)[1])[0]
    return xq, cq, variances


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--M", type=int, default=512)
    ap.add_argument("--T", type=int, default=241)     # 1960Q1-2020Q1
    args = ap.parse_args()
    _, y = smc.simulate(ucsv_mod([0.2, 3.0, 1.0, 1.0]), args.T, seed=1998)

    uc, xqs, cqs, _ = run_smc2(uc_mod, uc_prior, y, args.N, args.M, 3)                        # SMC(1024,512,uc_mod,uc_prior,3,0.5)  :58
    θ_uc = smc.expected_parameters(uc).ravel()
    print("UC   E[θ] = (x0, σε, ση) =", np.round(θ_uc, 3), " ess =", round(uc.ess, 1))
    print("     trend quartiles at T:", np.round(xqs[-1], 3), " cycle quartiles at T:", np.round(cqs[-1], 3))

    pred = uc_mod(θ_uc)                                                                        # :141-142
    for name, proposal in (("bootstrap", None), ("guided", smc.locally_optimal_proposal)):
        xq, _, _ = get_latent_states(args.N, y, pred, proposal)
        print(f"     particle filter at E[θ] ({name}): trend quartiles at T:", np.round(xq[-1], 3))

    ucsv, xqs, cqs, _ = run_smc2(ucsv_mod, ucsv_prior, y, args.N, args.M, 3)
    print("UCSV E[θ] = (γ, x0, log σε, log ση) =", np.round(smc.expected_parameters(ucsv).ravel(), 3), " ess =", round(ucsv.ess, 1))
    print("     trend quartiles at T:", np.round(xqs[-1], 3), " cycle quartiles at T:", np.round(cqs[-1], 3))
    pred_sv = ucsv_mod(smc.expected_parameters(ucsv).ravel())                                  # get_latent_states_ucsv  :241-260
    for name, proposal in (("bootstrap", None), ("guided: optimal trend move, docs/SPEC.md §10b", smc.UCSVTrendProposal(1.0))):
        xq, _, _ = get_latent_states(args.N, y, pred_sv, proposal)
        print(f"     particle filter at E[θ] ({name}): trend quartiles at T:", np.round(xq[-1], 3))


if __name__ == "__main__":
    main()
