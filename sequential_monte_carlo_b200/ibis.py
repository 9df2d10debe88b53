"""IBIS — host mirror of /root/reference/src/ibis.jl: the θ-level sampler of SMC² with the exact
Kalman filter as inner filter (no state particles).  The M recursions run as one device launch
(csrc/smcb_batch.cu): kalman_kernel for univariate LinearModels, kalman_mv_kernel for multivariate ones
(MultivariateLinearGaussian / hodrick_prescott; the reference's IBIS is generic in the state type XT, ΣT, ibis.jl:3-52)."""
import math
import sys

import numpy as np

from . import _lib
from .particles import default_context, resampler_id
from .smc_samplers import _propose, random_walk_kernel
from .state_space_models import MultivariateLinearModel


class IBIS:
    """mutable struct IBIS + constructor (ibis.jl:3-52).  Fields θ, ω, x, Σ, ess, ess_min, M, chain,
    logZ, model, prior, kernel, acc_threshold, acc_ratio as in the reference."""

    def __init__(self, M, model, prior, chain, ess_threshold, min_ar=-1.0, *, seed=1998, theta_resampler="multinomial", ctx=None):
        self.M, self.chain, self.model, self.prior = int(M), int(chain), model, prior
        self.seed = int(seed)
        self.theta_resampler = resampler_id(theta_resampler)
        self.ctx = ctx or default_context()
        self.kernel = random_walk_kernel
        self.θ = prior.sample(self.M, self.seed)                     # :34
        self.ω = np.full(self.M, 1.0 / self.M)                       # :35
        self.d = None                                                # state dimension of a multivariate model, else None
        P = self._params(self.θ)
        if self.d is None:
            self.x = P[:, 4].copy()                                  # [mod.x0 for mod in mods]   :38
            self.Σ = P[:, 5].copy()                                  # [mod.σ0 for mod in mods]   :39
        else:
            d = self.d
            self.x = P[:, 2 * d * d + d + 1: 2 * d * d + 2 * d + 1].copy()
            self.Σ = P[:, 2 * d * d + 2 * d + 1:].reshape(self.M, d, d).copy()
        self.logZ = np.zeros(self.M)
        self.ess, self.ess_min = 1.0 * self.M, self.M * float(ess_threshold)
        self.acc_threshold, self.acc_ratio = float(min_ar), 0.0
        self._n_resample = self._n_rejuv = 0

    def _params(self, θ):
        ms = [self.model(th) for th in θ]
        if ms and isinstance(ms[0], MultivariateLinearModel):
            self.d = ms[0].state_dim
            return np.stack([m.block() for m in ms])
        if ms and ms[0].kind != _lib.LG1D:
            raise TypeError("IBIS needs a LinearModel (the inner filter is the Kalman filter, ibis.jl:100,172)")
        return np.stack([m.params8() for m in ms])

    def _kalman_loglik(self, P, y, active):
        """M × log_likelihood(y, model(θ_m))  (ibis.jl:100) -> (logZ [M], x_T, Σ_T)"""
        if self.d is None:
            return self.ctx.kalman_loglik(P, y, matched_init=False, active=active)
        return self.ctx.kalman_mv_loglik(self.d, P, y, matched_init=False, active=active)

    def _kalman_step(self, P, y):
        """M × kalman_filter(model(θ_m), x_m, Σ_m, y)  (ibis.jl:172-177) -> (x, Σ, step log-likelihoods)"""
        if self.d is None:
            return self.ctx.kalman_step(P, self.x, self.Σ, y)
        return self.ctx.kalman_mv_step(self.d, P, self.x, self.Σ, y)

    def _where(self, accept, new, old):
        return np.where(accept.reshape((-1,) + (1,) * (np.ndim(old) - 1)), new, old)

    def __repr__(self):                                              # Base.show(io, ibis)   ibis.jl:66-71
        return f"ess     = {round(self.ess, 3)}\nmean(θ) = {expected_parameters(self).ravel()}"


def resample_(ibis):
    """resample!(ibis) (ibis.jl:72-84) — permutes θ, x, Σ, logZ."""
    ibis.ctx.set_rng(ibis.seed, 0)
    a = ibis.ctx.resample(ibis.ω, ibis.theta_resampler, stream=0, t=ibis._n_resample, purpose=_lib.P_THETA_RESAMPLE)
    ibis._n_resample += 1
    a = np.sort(a)   # docs/SPEC.md §5b: ascending θ-ancestors (same convention as SMC, where it keeps clouds rank-local)
    ibis.θ, ibis.x, ibis.Σ, ibis.logZ = ibis.θ[a], ibis.x[a], ibis.Σ[a], ibis.logZ[a]
    ibis.ω = np.full(ibis.M, 1.0 / ibis.M)
    return a


def rejuvenate_(ibis, y, ξ=1.0, verbose=False):
    """rejuvenate!(ibis, y, ξ, verbose) (ibis.jl:86-125)"""
    y = np.ascontiguousarray(y, np.float64)
    M, d = ibis.θ.shape
    acc = np.zeros(M, bool)
    Σk, univariate = ibis.kernel(ibis.θ)
    scales = 0.5 * np.arange(ibis.chain, 0, -1)
    if verbose:
        sys.stdout.write("\t[rejuvenating]")
    ordinal = ibis._n_rejuv
    ibis._n_rejuv += 1
    lp_cur = np.array([ibis.prior.logpdf(th) for th in ibis.θ])
    for c in range(ibis.chain):
        z = np.stack([_lib.rng_normals(ibis.seed, ordinal, k, c, _lib.P_MH_PROPOSAL, 0, M) for k in range(d)], axis=1)
        θ_prop = _propose(ibis.θ, Σk, univariate, scales[c], z)
        ok = np.array([ibis.prior.insupport(th) for th in θ_prop])
        P = ibis._params(np.where(ok[:, None], θ_prop, ibis.θ))
        logZ_prop, x_prop, Σ_prop = ibis._kalman_loglik(P, y, ok.astype(np.uint8))                                   # :100
        lp_prop = np.array([ibis.prior.logpdf(th) if o else -math.inf for th, o in zip(θ_prop, ok)])
        with np.errstate(invalid="ignore", divide="ignore"):
            ratio = ξ * (logZ_prop - ibis.logZ) + (lp_prop - lp_cur)
            u = _lib.rng_uniforms01(ibis.seed, ordinal, 0, c, _lib.P_MH_ACCEPT, M)
            accept = ok & (logZ_prop + lp_prop > -math.inf) & (np.log(u) < ratio)
        ibis.logZ = np.where(accept, logZ_prop, ibis.logZ)
        ibis.θ = np.where(accept[:, None], θ_prop, ibis.θ)
        ibis.x = ibis._where(accept, x_prop, ibis.x)
        ibis.Σ = ibis._where(accept, Σ_prop, ibis.Σ)
        lp_cur = np.where(accept, lp_prop, lp_cur)
        acc |= accept
    ibis.ω = np.full(M, 1.0 / M)
    ibis.acc_ratio = float(acc.sum()) / M
    if verbose:
        sys.stdout.write("\tacc_rate: %1.5f" % ibis.acc_ratio)
    return ibis


def smc2(ibis, y):
    """smc²(ibis, y) (ibis.jl:128-147)"""
    y = np.ascontiguousarray(y, np.float64)
    ibis.x, ibis.Σ, ll = ibis._kalman_step(ibis._params(ibis.θ), y[0])
    ibis.logZ = ll.copy()
    _, ibis.ω, ibis.ess = ibis.ctx.normalize(ll)
    return ibis


def smc2_step(ibis, y, t, verbose=True):
    """smc²!(ibis, y, t) (ibis.jl:154-189); t 0-based."""
    y = np.ascontiguousarray(y, np.float64)
    if verbose:
        sys.stdout.write("t = %4d\tess = %4.3f" % (t, ibis.ess))
    ibis.rejuvenated = False
    if ibis.ess < ibis.ess_min:
        resample_(ibis)
        rejuvenate_(ibis, y[:t], 1.0, verbose)
        ibis.rejuvenated = True
    with np.errstate(divide="ignore"):
        logω = np.log(ibis.ω)
    ibis.x, ibis.Σ, ll = ibis._kalman_step(ibis._params(ibis.θ), y[t])                      # :172-177
    logω = logω + ll
    ibis.logZ = ibis.logZ + ll
    _, ibis.ω, ibis.ess = ibis.ctx.normalize(logω)
    if verbose:
        sys.stdout.write("\n")
    return ibis


def expected_parameters(ibis, reference_style=False):
    """ibis.jl:54-58 (same D6 quirk as SMC's)"""
    ω = ibis.ω
    if reference_style:
        e = np.exp(ω - ω.max())
        ω = e / e.sum()
    return (ibis.θ * ω[:, None]).sum(axis=0)[:, None]


def observation_dist(ibis):
    """observation_dist(ibis) (plotting_utils.jl:96-113): the ω-mixture of the predicted measurement B x_m and of its
    variance B Σ_m B' + R over the θ-particles (filtered moments, no predict step — as the reference)."""
    P = ibis._params(ibis.θ)
    if ibis.d is None:
        ym, Sm = P[:, 1] * ibis.x, (P[:, 1] ** 2) * ibis.Σ + P[:, 3]
    else:
        d = ibis.d
        B, R = P[:, d * d: d * d + d], P[:, 2 * d * d + d]
        ym = np.einsum("mi,mi->m", B, ibis.x)
        Sm = np.einsum("mi,mij,mj->m", B, ibis.Σ, B) + R
    return float(np.sum(ibis.ω * ym)), float(np.sum(ibis.ω * Sm))


def estimated_trend(ibis):
    """estimated_trend(ibis::IBIS) (plotting_utils.jl:114)"""
    return observation_dist(ibis)[0]


def quantile(ibis, p):
    """quantile(ibis, p) (plotting_utils.jl:126-137): quantiles of Normal(y, sqrt(Σ)) with (y, Σ) = observation_dist(ibis)"""
    from statistics import NormalDist
    p = np.sort(np.atleast_1d(np.asarray(p, np.float64)))            # sort!(p)  :130
    y, S = observation_dist(ibis)
    return np.array([NormalDist(y, math.sqrt(S)).inv_cdf(float(v)) for v in p])

