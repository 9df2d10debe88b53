"""ORACLE — test infrastructure only.  CPU restatement of the reference's θ-level samplers,
function by function, on top of the C oracle's particle filters (oracle/smc_oracle.c):

    OSMC                <- mutable struct SMC + SMC(...)    /root/reference/src/smc_samplers.jl:5-59
    o_resample          <- resample!                        :74-84
    o_random_walk_kernel<- random_walk_kernel               :87-101
    o_rejuvenate        <- rejuvenate!                      :103-148
    o_exchange          <- exchange!                        :163-189
    o_density_tempered  <- density_tempered                 :222-281
    o_smc2 / o_smc2_step<- smc² / smc²!                     :288-340
    o_expected_parameters <- expected_parameters            :61-65
    OIBIS, o_ibis_*     <- IBIS + methods                   /root/reference/src/ibis.jl:3-189 (scalar and matrix Kalman inner filter)
    OSMC(proposal=...)  <- extension, not in the reference: guided inner filters (docs/SPEC.md §10) in every sweep / step
    priors              <- the Distributions.jl subset used by README.md:81-85 and
                           examples/inflation_example.jl:234-239 (un-vendored, un-pinned: SURVEY F8)

The structure follows the Julia code: a loop over θ-particles with one full CPU particle filter per
θ (threaded by the C oracle like `Threads.@threads`).  Randomness follows docs/SPEC.md §2/§8 so
that the CUDA path can be compared draw for draw.  PARITY UNPINNED: the only
recorded reference output is the docstring trace at smc_samplers.jl:207-219 (unknown data/seed),
used as a plausibility anchor in tests/test_theta_level.py.
"""
import math

import numpy as np

from . import oracle as o

HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------- priors
class ONormal:
    def __init__(self, mu=0.0, sigma=1.0):
        self.mu, self.sigma = float(mu), float(sigma)

    def insupport(self, x):
        return bool(np.isfinite(x))

    def logpdf(self, x):
        z = (x - self.mu) / self.sigma
        return -0.5 * z * z - math.log(self.sigma) - HALF_LOG_2PI

    def sample(self, M, seed, k):
        return self.mu + self.sigma * o.normals(seed, 0, k, 0, o.P_PRIOR, 0, M)


class OLogNormal:
    def __init__(self, mu=0.0, sigma=1.0):
        self.mu, self.sigma = float(mu), float(sigma)

    def insupport(self, x):
        return bool(np.isfinite(x) and x > 0.0)

    def logpdf(self, x):
        if not self.insupport(x):
            return -math.inf
        lx = math.log(x)
        z = (lx - self.mu) / self.sigma
        return -lx - math.log(self.sigma) - HALF_LOG_2PI - 0.5 * z * z

    def sample(self, M, seed, k):
        return np.exp(self.mu + self.sigma * o.normals(seed, 0, k, 0, o.P_PRIOR, 0, M))


class OUniform:
    def __init__(self, a=0.0, b=1.0):
        self.a, self.b = float(a), float(b)

    def insupport(self, x):
        return bool(self.a <= x <= self.b)

    def logpdf(self, x):
        return -math.log(self.b - self.a) if self.insupport(x) else -math.inf

    def sample(self, M, seed, k):
        return self.a + (self.b - self.a) * o.uniforms01(seed, 0, k, 0, o.P_PRIOR, M)


class OTruncatedNormal:
    def __init__(self, mu, sigma, lo, hi):
        self.mu, self.sigma, self.lo, self.hi = float(mu), float(sigma), float(lo), float(hi)
        cdf = lambda z: 0.5 * (1.0 + math.erf(z / math.sqrt(2.0)))
        self.logmass = math.log(cdf((self.hi - self.mu) / self.sigma) - cdf((self.lo - self.mu) / self.sigma))

    def insupport(self, x):
        return bool(self.lo <= x <= self.hi)

    def logpdf(self, x):
        if not self.insupport(x):
            return -math.inf
        z = (x - self.mu) / self.sigma
        return -0.5 * z * z - math.log(self.sigma) - HALF_LOG_2PI - self.logmass

    def sample(self, M, seed, k):
        out, todo, attempt = np.empty(M), np.ones(M, bool), 0
        while todo.any():
            x = self.mu + self.sigma * o.normals(seed, 0, k, attempt, o.P_PRIOR, 0, M)
            ok = todo & (x >= self.lo) & (x <= self.hi)
            out[ok] = x[ok]
            todo &= ~ok
            attempt += 1
        return out


class OProduct:
    def __init__(self, comps):
        self.comps = list(comps)

    def insupport(self, th):
        return all(c.insupport(float(v)) for c, v in zip(self.comps, th))

    def logpdf(self, th):
        s = 0.0
        for c, v in zip(self.comps, th):
            s += c.logpdf(float(v))
        return s

    def sample(self, M, seed):
        return np.stack([c.sample(M, seed, k) for k, c in enumerate(self.comps)], axis=1)


# ----------------------------------------------------------------------------- SMC
class OSMC:
    """SMC(N, M, model, prior, chain, ess_threshold, min_ar)  smc_samplers.jl:29-59.
    `model(θ)` returns (kind, params) for the C oracle."""

    def __init__(self, N, M, model, prior, chain, ess_threshold, min_ar=-1.0, seed=1998, resampler=o.MULTINOMIAL,
                 theta_resampler=o.MULTINOMIAL, proposal=None):
        self.N, self.M, self.chain = N, M, chain
        self.proposal = proposal                           # extension: proposal(P [M,8], y) -> [M,3] guides every inner filter step (SPEC §10)
        self.model, self.prior = model, prior
        self.seed, self.resampler, self.theta_resampler = seed, resampler, theta_resampler
        self.theta = prior.sample(M, seed)                 # θ = map(m -> rand(prior), 1:M)      :38
        self.omega = np.full(M, 1.0 / M)                   # :39
        self.logZ = np.zeros(M)                            # :44
        self.ess = 1.0 * M                                 # :45
        self.ess_min = M * ess_threshold                   # :46
        self.acc_threshold, self.acc_ratio = min_ar, 0.0
        self.kind = model(self.theta[0])[0]
        self.d = o.state_dim(self.kind)
        self.x = np.zeros((M, self.d, N))                  # :41 (one array per θ; the reference aliases, harmlessly)
        self.logw = np.zeros((M, N))                       # unnormalised log-weights of each cloud (w = normalize(logw))
        self.epoch, self.n_resample, self.n_rejuv = 1, 0, 0
        self.cloud_epoch = 0                               # Philox epoch the live clouds step in

    def params(self, theta):
        return np.stack([o.params8(self.model(th)[1]) for th in theta])

    def sweep(self, P, active, y, epoch):
        """M × log_likelihood(N, y, model(θ_m)) — guided when a proposal is set"""
        if self.proposal is None:
            return o.batch_log_likelihood(self.kind, P, active, self.N, y, self.resampler, self.seed, epoch, 0)
        prop = np.stack([np.asarray(self.proposal(P, float(yt)), np.float64) for yt in y])
        return o.batch_guided_log_likelihood(self.kind, P, active, self.N, y, self.resampler, prop, self.seed, epoch, 0)


def o_expected_parameters(smc, reference_style=False):
    w = smc.omega
    if reference_style:                                    # _,ω,_ = reweight(smc.ω)             :62
        _, w, _ = o.normalize(w)
    return (smc.theta * w[:, None]).sum(axis=0)[:, None]   # :63-64


def o_resample(smc):
    a = o.resample_w(smc.omega, smc.theta_resampler, smc.seed, 0, 0, smc.n_resample, purpose=o.P_THETA_RESAMPLE)   # :75
    smc.n_resample += 1
    a = np.sort(a)                                         # SPEC §5b: θ-ancestors in ascending order (slots are exchangeable)
    smc.theta = smc.theta[a]                               # :78
    smc.omega = np.full(smc.M, 1.0 / smc.M)                # ω[a], made uniform by rejuvenate! :139
    smc.x = smc.x[a].copy()                                # :82 (deep copy: SURVEY D4)
    smc.logw = smc.logw[a].copy()                          # the weights follow their cloud
    smc.logZ = smc.logZ[a]                                 # :83
    return a


def _seq_sum(v):
    """left-to-right sum (np.cumsum accumulates sequentially; np.sum is pairwise)"""
    return np.cumsum(np.ascontiguousarray(v, np.float64))[-1]


def o_cholesky(A):
    """lower Cholesky factor by the plain Cholesky–Banachiewicz recursion (docs/SPEC.md §11: LAPACK's blocked / fused
    operation order is not reproducible on a device)"""
    d = A.shape[0]
    L = np.zeros((d, d))
    for j in range(d):
        s = float(A[j, j])
        for k in range(j):
            s = s - L[j, k] * L[j, k]
        if not s > 0.0:
            raise np.linalg.LinAlgError("Matrix is not positive definite")
        L[j, j] = math.sqrt(s)
        for i in range(j + 1, d):
            v = float(A[i, j])
            for k in range(j):
                v = v - L[i, k] * L[j, k]
            L[i, j] = v / L[j, j]
    return L


def o_propose(theta, L, z):
    """θ'_j = θ_j + Σ_{k<=j} z_k L[j][k], k ascending (docs/SPEC.md §11)"""
    out = np.empty_like(theta)
    for j in range(theta.shape[1]):
        acc = z[:, 0] * L[j, 0]
        for k in range(1, j + 1):
            acc = acc + z[:, k] * L[j, k]
        out[:, j] = theta[:, j] + acc
    return out


def o_random_walk_kernel(theta):
    M, d = theta.shape
    mean = np.array([_seq_sum(theta[:, k]) for k in range(d)]) / M     # cov(Θ'): sums in slot order (docs/SPEC.md §11)
    c = theta - mean
    cov = np.empty((d, d))
    for j in range(d):
        for k in range(j, d):
            cov[j, k] = cov[k, j] = _seq_sum(c[:, j] * c[:, k]) / (M - 1)
    if d == 1:                                             # :87-92
        dth = 2.83 * 2.83
        sig = 1.0e-2 if abs(cov[0, 0]) < 1.0e-8 else dth * cov[0, 0] + 1.0e-10
        return np.array([[sig]]), True
    dth = (2.83 * 2.83) / d                                # :97
    if math.sqrt(float(_seq_sum((cov * cov).ravel()))) < 1.0e-8:       # norm(cov) < 1e-8        :98
        return 1.0e-2 * np.eye(d), False
    return dth * cov + 1.0e-10 * np.eye(d), False


def o_rejuvenate(smc, y, xi=1.0):
    y = np.ascontiguousarray(y, np.float64)
    M, d = smc.theta.shape
    acc = np.zeros(M, bool)                                # acc_array = zeros(Int64, M)          :104
    Sigma, uni = o_random_walk_kernel(smc.theta)           # :107
    scales = 0.5 * np.arange(smc.chain, 0, -1)             # :108
    ordinal = smc.n_rejuv
    smc.n_rejuv += 1
    lp_cur = np.array([smc.prior.logpdf(th) for th in smc.theta])
    for c in range(smc.chain):                             # for c in 1:chain (batched over m)    :113
        z = np.stack([o.normals(smc.seed, ordinal, k, c, o.P_MH_PROPOSAL, 0, M) for k in range(d)], axis=1)
        if uni:
            prop = smc.theta + (scales[c] * Sigma[0, 0]) * z
        else:
            prop = o_propose(smc.theta, o_cholesky(scales[c] * Sigma), z)   # rand(MvNormal(θ[m], scales[c]·Σ))    :114
        ok = np.array([smc.prior.insupport(th) for th in prop])       # :116
        P = smc.params(np.where(ok[:, None], prop, smc.theta))
        epoch = smc.epoch
        smc.epoch += 1
        zprop, xprop, lwprop = smc.sweep(P, ok.astype(np.uint8), y, epoch)   # log_likelihood(N, y, model(θ_prop))  :117-121
        lp_prop = np.array([smc.prior.logpdf(th) if k else -math.inf for th, k in zip(prop, ok)])
        with np.errstate(invalid="ignore", divide="ignore"):
            ratio = xi * (zprop - smc.logZ) + (lp_prop - lp_cur)      # :123-127
            u = o.uniforms01(smc.seed, ordinal, 0, c, o.P_MH_ACCEPT, M)
            accept = ok & (zprop + lp_prop > -math.inf) & (np.log(u) < ratio)   # :129
        for m in np.flatnonzero(accept):                   # :130-135
            smc.logZ[m] = zprop[m]
            smc.theta[m] = prop[m]
            smc.x[m] = xprop[m]
            smc.logw[m] = lwprop[m]
            lp_cur[m] = lp_prop[m]
            acc[m] = True
    smc.omega = np.full(M, 1.0 / M)                        # ω[m] = 1.0                            :139
    smc.acc_ratio = float(acc.sum()) / M                   # :142
    return smc


def o_exchange(smc, y):
    """exchange!(smc, y)  smc_samplers.jl:163-189: double N when the acceptance ratio fell below acc_threshold"""
    if not (smc.acc_ratio < smc.acc_threshold):            # :164
        return
    if smc.N > 4096:                                       # :166,186-187
        return
    smc.N *= 2                                             # :167
    epoch = smc.epoch
    smc.epoch += 1
    smc.cloud_epoch = epoch
    new_logZ, smc.x, smc.logw = smc.sweep(smc.params(smc.theta), None, np.ascontiguousarray(y, np.float64), epoch)   # :174-180
    _, smc.omega, smc.ess = o.normalize(new_logZ - smc.logZ)   # :183
    smc.logZ = new_logZ                                    # :184


def o_density_tempered(smc, y):
    y = np.ascontiguousarray(y, np.float64)
    epoch = smc.epoch
    smc.epoch += 1
    smc.cloud_epoch = epoch
    smc.logZ, smc.x, smc.logw = smc.sweep(smc.params(smc.theta), None, y, epoch)   # :223-229
    _, smc.omega, smc.ess = o.normalize(smc.logZ)          # :232
    xi = 0.0
    smc.schedule = []
    while xi < 1.0:                                        # :235
        resample_flag = True
        lower = oldxi = xi
        upper = 2.0
        newxi = None
        while upper - lower > 1.0e-6:                      # :246
            newxi = (upper + lower) / 2.0
            logw = (newxi - oldxi) * smc.logZ
            _, smc.omega, smc.ess = o.normalize(logw)
            if smc.ess == smc.ess_min:
                break
            elif smc.ess < smc.ess_min:
                upper = newxi
            else:
                lower = newxi
        if newxi >= 1.0:                                   # :261
            resample_flag = False
            newxi = 1.0
            logw = (newxi - oldxi) * smc.logZ
            _, smc.omega, smc.ess = o.normalize(logw)
        xi = newxi
        smc.schedule.append((xi, smc.ess))
        if resample_flag:
            o_resample(smc)                                # :272
            o_rejuvenate(smc, y, xi)                       # :275
    return smc


def o_smc2(smc, y):
    epoch = smc.epoch
    smc.epoch += 1
    smc.cloud_epoch = epoch
    P = smc.params(smc.theta)
    logmu = np.empty(smc.M)
    for m in range(smc.M):                                 # :289-295
        smc.x[m], smc.logw[m] = o.bootstrap_init(smc.kind, P[m], smc.N, y[0], smc.seed, epoch, m)
        logmu[m], _, _ = o.normalize(smc.logw[m])
    smc.logZ = logmu.copy()                                # :297
    _, smc.omega, smc.ess = o.normalize(logmu)             # :298
    return smc


def o_smc2_step(smc, y, t):
    """t: 0-based index of the new observation (Julia's t-1)."""
    smc.rejuvenated = False
    if smc.ess < smc.ess_min:                              # :312
        o_resample(smc)                                    # :314
        o_rejuvenate(smc, y[:t], 1.0)                      # :317
        o_exchange(smc, y[:t])                             # :320
        smc.rejuvenated = True
    with np.errstate(divide="ignore"):
        logw = np.log(smc.omega)                           # :324
    P = smc.params(smc.theta)
    prop_t = None if smc.proposal is None else np.asarray(smc.proposal(P, float(y[t])), np.float64)
    for m in range(smc.M):                                 # :325-335
        if smc.proposal is None:
            o.bootstrap_step(smc.kind, P[m], smc.x[m], smc.logw[m], y[t], t, smc.resampler, smc.seed, smc.cloud_epoch, m)
        else:
            o.guided_step(smc.kind, P[m], smc.x[m], smc.logw[m], y[t], t, smc.resampler, prop_t[m], smc.seed, smc.cloud_epoch, m)
        lm, _, _ = o.normalize(smc.logw[m])
        logw[m] += lm
        smc.logZ[m] += lm
    _, smc.omega, smc.ess = o.normalize(logw)              # :338
    return smc


# ----------------------------------------------------------------------------- IBIS (ibis.jl)
class OIBIS:
    """IBIS(M, model, prior, chain, ess_threshold, min_ar)  ibis.jl:26-52: the θ-level machinery of
    SMC with the Kalman filter as the (exact) inner filter.  `model(θ)` returns (kind, params) for a univariate
    LinearModel or (("mv", d), block) for a multivariate one (block = A, B, Q, R, x0, Σ0 row-major)."""

    def __init__(self, M, model, prior, chain, ess_threshold, min_ar=-1.0, seed=1998, theta_resampler=o.MULTINOMIAL):
        self.M, self.chain, self.model, self.prior, self.seed = M, chain, model, prior, seed
        self.theta_resampler = theta_resampler
        self.theta = prior.sample(M, seed)
        self.omega = np.full(M, 1.0 / M)
        self.logZ = np.zeros(M)
        self.ess, self.ess_min = 1.0 * M, M * ess_threshold
        self.acc_threshold, self.acc_ratio = min_ar, 0.0
        kind0 = self.model(self.theta[0])[0]
        self.d = kind0[1] if isinstance(kind0, tuple) else None
        P = self.params(self.theta)
        if self.d is None:
            self.x = P[:, 4].copy()                        # model(θ).x0   ibis.jl:39
            self.Sigma = P[:, 5].copy()                    # model(θ).σ0   ibis.jl:40
        else:
            d = self.d
            self.x = P[:, 2 * d * d + d + 1: 2 * d * d + 2 * d + 1].copy()
            self.Sigma = P[:, 2 * d * d + 2 * d + 1:].reshape(M, d, d).copy()
        self.n_resample, self.n_rejuv = 0, 0

    def params(self, theta):
        if self.d is not None:
            return np.stack([np.asarray(self.model(th)[1], np.float64) for th in theta])
        return np.stack([o.params8(self.model(th)[1]) for th in theta])

    def kalman_loglik(self, Pm, y):
        if self.d is None:
            return o.kalman_loglik(Pm, y, matched_init=False)
        return o.kalman_mv_loglik(self.d, Pm, y, matched_init=False)

    def kalman_step(self, Pm, x, S, y):
        if self.d is None:
            return o.kalman_step(Pm, x, S, y)
        return o.kalman_mv_step(self.d, Pm, x, S, y)


def o_ibis_resample(s):
    a = o.resample_w(s.omega, s.theta_resampler, s.seed, 0, 0, s.n_resample, purpose=o.P_THETA_RESAMPLE)   # ibis.jl:75
    s.n_resample += 1
    a = np.sort(a)                                                                 # SPEC §5b
    s.theta, s.x, s.Sigma, s.logZ = s.theta[a], s.x[a], s.Sigma[a], s.logZ[a]     # ibis.jl:78-84
    s.omega = np.full(s.M, 1.0 / s.M)
    return a


def o_ibis_rejuvenate(s, y):
    y = np.ascontiguousarray(y, np.float64)
    M, d = s.theta.shape
    acc = np.zeros(M, bool)
    Sigma, uni = o_random_walk_kernel(s.theta)             # ibis.jl:90
    scales = 0.5 * np.arange(s.chain, 0, -1)
    ordinal = s.n_rejuv
    s.n_rejuv += 1
    lp_cur = np.array([s.prior.logpdf(th) for th in s.theta])
    for c in range(s.chain):                               # ibis.jl:95-119
        z = np.stack([o.normals(s.seed, ordinal, k, c, o.P_MH_PROPOSAL, 0, M) for k in range(d)], axis=1)
        prop = s.theta + (scales[c] * Sigma[0, 0]) * z if uni else o_propose(s.theta, o_cholesky(scales[c] * Sigma), z)
        ok = np.array([s.prior.insupport(th) for th in prop])
        P = s.params(np.where(ok[:, None], prop, s.theta))
        zprop, xprop, Sprop = np.full(M, -math.inf), np.zeros_like(s.x), np.zeros_like(s.Sigma)
        for m in np.flatnonzero(ok):
            xprop[m], Sprop[m], zprop[m] = s.kalman_loglik(P[m], y)                       # log_likelihood(y, model(θ_prop))  ibis.jl:100
        lp_prop = np.array([s.prior.logpdf(th) if k else -math.inf for th, k in zip(prop, ok)])
        with np.errstate(invalid="ignore", divide="ignore"):
            ratio = (zprop - s.logZ) + (lp_prop - lp_cur)
            u = o.uniforms01(s.seed, ordinal, 0, c, o.P_MH_ACCEPT, M)
            accept = ok & (zprop + lp_prop > -math.inf) & (np.log(u) < ratio)
        s.logZ = np.where(accept, zprop, s.logZ)
        s.theta = np.where(accept[:, None], prop, s.theta)
        s.x = np.where(accept.reshape((-1,) + (1,) * (s.x.ndim - 1)), xprop, s.x)
        s.Sigma = np.where(accept.reshape((-1,) + (1,) * (s.Sigma.ndim - 1)), Sprop, s.Sigma)
        lp_cur = np.where(accept, lp_prop, lp_cur)
        acc |= accept
    s.omega = np.full(M, 1.0 / M)
    s.acc_ratio = float(acc.sum()) / M
    return s


def o_ibis_init(s, y):
    """smc²(ibis, y)  ibis.jl:128-147: one Kalman step per θ at the first observation."""
    P = s.params(s.theta)
    ll = np.empty(s.M)
    for m in range(s.M):
        s.x[m], s.Sigma[m], ll[m] = s.kalman_step(P[m], s.x[m], s.Sigma[m], y[0])
    s.logZ = ll.copy()
    _, s.omega, s.ess = o.normalize(ll)
    return s


def o_ibis_step(s, y, t):
    """smc²!(ibis, y, t)  ibis.jl:154-189 (t 0-based)."""
    s.rejuvenated = False
    if s.ess < s.ess_min:
        o_ibis_resample(s)
        o_ibis_rejuvenate(s, y[:t])
        s.rejuvenated = True
    with np.errstate(divide="ignore"):
        logw = np.log(s.omega)
    P = s.params(s.theta)
    for m in range(s.M):
        s.x[m], s.Sigma[m], ll = s.kalman_step(P[m], s.x[m], s.Sigma[m], y[t])     # ibis.jl:172-177
        logw[m] += ll
        s.logZ[m] += ll
    _, s.omega, s.ess = o.normalize(logw)
    return s
