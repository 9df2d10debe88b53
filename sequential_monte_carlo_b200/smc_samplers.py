"""θ-level samplers — host mirror of /root/reference/src/smc_samplers.jl.

    SMC(N, M, model, prior, chain, ess_threshold, min_ar=-1.0)        smc_samplers.jl:5-59
    smc2(smc, y)                       (Julia: smc²)                  :288-301
    smc2_step(smc, y, t, verbose)      (Julia: smc²!)                 :308-340
    density_tempered(smc, y, verbose)                                 :222-281
    expected_parameters(smc)                                          :61-65
    resample_ / rejuvenate_ / random_walk_kernel                      :74-148

The reference loops over θ-particles on the host (`Threads.@threads`, or serially in smc²!) and runs
one CPU particle filter per θ.  Here every such loop is ONE batched launch over all θ
(`_lib.Batch`, csrc/smcb_batch.cu); the host keeps only the M-length control flow.  With a
communicator the θ-particles are sharded across GPUs (one process per GPU): every rank holds whole
state clouds for its slice, the M-length vectors are replicated through all-gathers, and the
clouds of resampled parents move between GPUs.  Results do not depend on the
number of GPUs: Philox streams are indexed by the global θ index.

Python spelling: `t` is the 0-based index of the observation being assimilated (Julia's t-1);
ancestors are 0-based.  Decisions on the reference's defects D3–D7 are in DESIGN.md.
"""
import math
import sys

import numpy as np

from . import _lib
from .particles import default_context, resampler_id
from .state_space_models import params_of


# ----------------------------------------------------------------------------- communicator
class LocalComm:
    """world of one GPU"""
    rank, world = 0, 1

    def all_gather(self, local):
        return np.ascontiguousarray(local)

    def exchange(self, send, recv_counts, nbytes):  # pragma: no cover - never called with world == 1
        raise RuntimeError("no peers")


class TorchComm:
    """One process per GPU over torch.distributed (NCCL on GPUs; gloo in the CPU tests).
    all_gather moves the replicated M-length vectors; exchange moves packed clouds point-to-point."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                                          if dist.get_backend() == "nccl" else torch.device("cpu"))
        self.warm_up()

    def warm_up(self):
        """NCCL builds its all-gather rings and every point-to-point channel lazily (hundreds of ms on
        first use): touch them once here, outside anybody's timed region."""
        self.all_gather(np.zeros(1))
        if self.world > 1:
            one = {r: self.torch.zeros(8, dtype=self.torch.uint8, device=self.device) for r in range(self.world) if r != self.rank}
            self.exchange(one, {r: 1 for r in one}, 8)

    def all_gather(self, local):
        t = self.torch.from_numpy(np.ascontiguousarray(local, np.float64)).to(self.device)
        out = self.torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
        self.dist.all_gather_into_tensor(out, t)   # rank-major concatenation along dim 0
        return out.cpu().numpy()

    def exchange(self, send, recv_counts, nbytes):
        """send: {dst_rank: uint8 tensor of k*nbytes}; recv_counts: {src_rank: k}.  Returns {src: tensor}."""
        ops, recv = [], {}
        for src, k in sorted(recv_counts.items()):
            recv[src] = self.torch.empty(k * nbytes, dtype=self.torch.uint8, device=self.device)
            ops.append(self.dist.P2POp(self.dist.irecv, recv[src], src))
        for dst, buf in sorted(send.items()):
            ops.append(self.dist.P2POp(self.dist.isend, buf, dst))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
            if self.device.type == "cuda":
                self.torch.cuda.synchronize(self.device)
        return recv


class NcclComm:
    """θ-sharding over the library's own NCCL communicator (smcb_comm_init, include/smcb200.h): one process per GPU,
    rank r owns θ-particles [r·M/G, (r+1)·M/G).  The host language only carries the 128-byte NCCL id from rank 0 to the
    other ranks — here through torch.distributed (any backend), in Julia through MPI or a file; every M-length vector
    then stays on the GPUs (all-gathers and cloud moves are issued by the library on its own stream)."""

    def __init__(self, ctx, rank, world, unique_id):
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        if ctx.comm_rank()[1] == 1 and self.world > 1:
            ctx.comm_init(self.rank, self.world, unique_id)

    @classmethod
    def from_torch(cls, ctx):
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        if ctx.comm_rank()[1] > 1 or world == 1:
            return cls(ctx, rank, world, None)
        box = [_lib.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(ctx, rank, world, box[0])

    def all_gather(self, local):
        return self.ctx.comm_all_gather(local)


_PRIOR_FAMILIES = None


def prior_descriptor(prior):
    """[d, 8] rows (family, p0, p1, lo, hi, c0, c1, 0) of a product of univariate priors for smcb_sampler_config, or None when
    a component is not one of the device families (Normal, LogNormal, Uniform, TruncatedNormal)"""
    from . import priors as pr
    comps = getattr(prior, "components", None)
    if comps is None:
        comps = [prior] if isinstance(prior, pr.Distribution) else None
    if not comps or len(comps) > _lib.MAX_THETA_DIM:
        return None
    rows = np.zeros((len(comps), 8))
    for k, c in enumerate(comps):
        if type(c) is pr.Normal:
            rows[k, :6] = (_lib.PRIOR_NORMAL, c.μ, c.σ, 0.0, 0.0, math.log(c.σ))
        elif type(c) is pr.LogNormal:
            rows[k, :6] = (_lib.PRIOR_LOGNORMAL, c.μ, c.σ, 0.0, 0.0, math.log(c.σ))
        elif type(c) is pr.Uniform:
            rows[k, :6] = (_lib.PRIOR_UNIFORM, 0.0, 0.0, c.a, c.b, -math.log(c.b - c.a))
        elif type(c) is pr.TruncatedNormal:
            rows[k, :7] = (_lib.PRIOR_TRUNCNORMAL, c.μ, c.σ, c.lo, c.hi, math.log(c.σ), c._logmass)
        else:
            return None
    return rows


def parameter_map(model, prior, d):
    """(kind, src [8], const [8]) when model(θ) only SELECTS components of θ and constants into the model's parameter block
    (every model closure of the reference's README and example: lg_mod, uc_mod, ucsv_mod), found by probing the closure on
    prior draws; None for anything else (the host-language sampler then runs the control flow)."""
    try:
        θ = np.asarray(prior.sample(6, 0xC0FFEE), np.float64).reshape(6, d)
        P, m0 = params_of(model, θ)
    except Exception:
        return None
    src, cst = np.full(8, -1, np.int32), np.zeros(8)
    for k in range(8):
        col = P[:, k]
        if np.all(col == col[0]):
            hits = [j for j in range(d) if np.all(θ[:, j] == col)]
            if hits:
                src[k] = hits[0]
            else:
                cst[k] = col[0]
            continue
        hits = [j for j in range(d) if np.array_equal(θ[:, j], col)]
        if not hits:
            return None
        src[k] = hits[0]
    if not hasattr(m0, "kind") or m0.kind not in (_lib.LG1D, _lib.SV, _lib.UCSV):
        return None
    return int(m0.kind), src, cst


def exchange_plan(parents, rank, world):
    """Who sends which cloud where after a θ-resample (replicated, deterministic).
    Returns (local_parents[Mloc] int32, send {dst: [src local slots]}, recv {src: [dst local slots]})."""
    parents = np.asarray(parents, np.int64)
    M = parents.size
    Mloc = M // world
    owner = parents // Mloc
    dst_rank = np.arange(M) // Mloc
    lo = rank * Mloc
    local_parents = np.arange(Mloc, dtype=np.int32)
    mine = slice(lo, lo + Mloc)
    is_local = owner[mine] == rank
    local_parents[is_local] = (parents[mine][is_local] - lo).astype(np.int32)
    send, recv = {}, {}
    for m in np.flatnonzero(owner != dst_rank):  # increasing m on both sides => matching order
        if owner[m] == rank:
            send.setdefault(int(dst_rank[m]), []).append(int(parents[m] - lo))
        if dst_rank[m] == rank:
            recv.setdefault(int(owner[m]), []).append(int(m - lo))
    return local_parents, send, recv


# ----------------------------------------------------------------------------- proposal kernel
def random_walk_kernel(θ):
    """smc_samplers.jl:87-101.  θ: [M, d].  Returns Σ (d×d) such that proposals are
    MvNormal(x, scale·Σ); for d = 1 the reference uses σ = 2.83²·var + 1e-10 as a standard
    deviation (Normal(x, scale·σ)) — kept, and encoded as Σ = σ² with scale applied to σ."""
    θ = np.ascontiguousarray(θ, np.float64)
    # the sums run in slot order and the 2.83²/d scaling, the 1e-10 jitter and the small-covariance branch are applied
    # by the library's host routine (docs/SPEC.md §11), the one the device-resident sampler uses: both propose the same θ'
    return _lib.random_walk_sigma(θ), θ.shape[1] == 1


def _propose(θ, Σ, univariate, scale, z):
    """rand(MvNormal(θ_m, scale·Σ)) for every m: θ'_j = θ_j + Σ_{k<=j} z_k L[j][k] with L the plain lower Cholesky factor
    of scale·Σ and k ascending (docs/SPEC.md §11; LAPACK / BLAS operation orders are not reproducible on a device)"""
    if univariate:
        return θ + (scale * Σ[0, 0]) * z
    L = _lib.cholesky_lower(Σ, scale)
    out = np.empty_like(θ)
    for j in range(θ.shape[1]):
        acc = z[:, 0] * L[j, 0]
        for k in range(1, j + 1):
            acc = acc + z[:, k] * L[j, k]
        out[:, j] = θ[:, j] + acc
    return out


def lg_optimal_proposals(P, y):
    """[M, 3] locally optimal proposals p(x' | xp, y) of M univariate LinearModels given their parameter blocks
    P [M, 8] = (A, B, Q, R, x0, σ0, ·, ·): 1/c2² = 1/Q + B²/R, c1 = c2²·A/Q, c0 = c2²·B·y/R (docs/SPEC.md §10)."""
    P = np.asarray(P, np.float64)
    A, B, Q, R = P[:, 0], P[:, 1], P[:, 2], P[:, 3]
    s2 = 1.0 / (1.0 / Q + B * B / R)
    return np.stack([s2 * B * float(y) / R, s2 * A / Q, np.sqrt(s2)], axis=1)


# ----------------------------------------------------------------------------- the sampler
class SMC:
    """mutable struct SMC (smc_samplers.jl:5-27) + constructor (:29-59).

    N state particles per θ, M θ-particles.  Public fields as in the reference:
    θ [M, d], ω [M], ess, ess_min, N, M, chain, logZ [M], model, prior, kernel, acc_threshold,
    acc_ratio; `x` and `w` are read from the device on access ([M_local, d, N] / [M_local, N]).
    """

    def __init__(self, N, M, model, prior, chain, ess_threshold, min_ar=-1.0, *, seed=1998, resampler="multinomial",
                 theta_resampler="multinomial", ctx=None, comm=None, proposal=None, engine="auto"):
        """engine: "device" keeps θ, ω, logZ and the whole control flow on the GPU(s) (smcb_sampler_*, csrc/smcb_sampler.cu) —
        possible when the prior is a product of the device families and model(θ) selects components of θ into the parameter
        block; "host" runs the control flow below in numpy with one batched launch per loop over θ (any closures, guided
        inner filters); "auto" picks "device" whenever it is possible."""
        self.N, self.M, self.chain = int(N), int(M), int(chain)
        self.model, self.prior = model, prior
        self.kernel = random_walk_kernel
        self.acc_threshold, self.acc_ratio = float(min_ar), 0.0
        # Extension (not in the reference's SMC): guided inner filters.  proposal(P [M, 8], y) -> [M, 3] gives every
        # θ-particle's affine-Gaussian proposal (c0, c1, c2) for the step that assimilates y (docs/SPEC.md §10), e.g.
        # lg_optimal_proposals; every inner filter step after the first then runs as particle_filter! instead of
        # bootstrap_filter!.  None (default) is the reference's algorithm.
        self.proposal = proposal
        self.seed = int(seed)
        self.resampler = resampler_id(resampler)
        self.theta_resampler = resampler_id(theta_resampler)
        self.comm = comm or LocalComm()
        if self.M % self.comm.world:
            raise ValueError(f"M={self.M} must be divisible by the number of GPUs ({self.comm.world})")
        self.Mloc = self.M // self.comm.world
        self.lo = self.comm.rank * self.Mloc
        self.ctx = ctx or default_context()
        self._θ = prior.sample(self.M, self.seed)                      # θ = map(m -> rand(prior), 1:M)      :38
        self._ω = np.full(self.M, 1.0 / self.M)                        # :39
        self._logZ = np.zeros(self.M)                                  # :44
        self.ess = 1.0 * self.M                                        # :45
        self.ess_min = self.M * float(ess_threshold)                   # :46
        self._P, m0 = params_of(model, self._θ)     # [M, 8] parameter blocks, kept in step with θ
        self.kind, self.d = m0.kind, m0.state_dim
        self._eng, self._stale, self._y_dev = None, False, None
        if engine not in ("auto", "device", "host"):
            raise ValueError("engine must be 'auto', 'device' or 'host'")
        if engine != "host":
            why = self._try_device_engine(ess_threshold)
            if why and engine == "device":
                raise ValueError("engine='device' is not possible here: " + why)
        self.engine = "device" if self._eng is not None else "host"
        self._cur = self.ctx.batch(self.kind, self.Mloc, self.N) if self._eng is None else None
        self._prop = None
        self._epoch = 1          # ordinal of the next batched sweep (device Philox epoch)
        self._n_resample = 0     # ordinal of the next θ-resample
        self._n_rejuv = 0        # ordinal of the next rejuvenation (host Philox epoch)
        self._params_dirty = True  # device copy of the parameter blocks is stale
        self.stats = {"sweeps": 0, "particle_updates": 0, "device_ms": 0.0, "clouds_moved": 0}

    # -- the device-resident engine
    def _try_device_engine(self, ess_threshold):
        """create the smcb_sampler; returns None on success, else the reason the host path is used"""
        if not isinstance(self.ctx, _lib.Context):
            return "the context is not a CUDA context"
        if self.proposal is not None:
            return "guided inner filters run on the host-language path"
        if not (2 <= self.M <= _lib.MAX_THETA_PARTICLES) or self.N > 8192:
            return "M or N out of the device sampler's range"
        d = self._θ.shape[1]
        rows = prior_descriptor(self.prior)
        if rows is None or rows.shape[0] != d:
            return "the prior is not a product of Normal / LogNormal / Uniform / TruncatedNormal"
        pm = parameter_map(self.model, self.prior, d)
        if pm is None or pm[0] != self.kind:
            return "model(θ) is not a selection of θ components and constants"
        if isinstance(self.comm, TorchComm):
            if self.comm.world > 1 and self.comm.device.type != "cuda":
                return "the communicator is not on CUDA devices"
            self.comm = NcclComm.from_torch(self.ctx)
        elif not isinstance(self.comm, (LocalComm, NcclComm)):
            return "unknown communicator type"
        cfg = _lib.SamplerConfig()
        cfg.kind, cfg.d_theta, cfg.N, cfg.M, cfg.chain = int(self.kind), d, self.N, self.M, self.chain
        cfg.resampler, cfg.theta_resampler = int(self.resampler), int(self.theta_resampler)
        cfg.ess_threshold, cfg.min_ar, cfg.seed = float(ess_threshold), float(self.acc_threshold), self.seed & (2 ** 64 - 1)
        for k in range(d):
            for j in range(8):
                cfg.prior[k][j] = float(rows[k, j])
        for k in range(8):
            cfg.map_src[k], cfg.map_const[k] = int(pm[1][k]), float(pm[2][k])
        self._eng = self.ctx.sampler(cfg, self._θ)
        return None

    def _engine_data(self, y):
        """the observations the engine's calls refer to live on the device; re-sent only when they change"""
        y = np.ascontiguousarray(y, np.float64)
        if self._y_dev is None or self._y_dev.shape != y.shape or not np.array_equal(self._y_dev, y):
            self._eng.set_data(y)
            self._y_dev = y.copy()

    def _refresh(self):
        if self._eng is not None and self._stale:
            self._θ, self._ω, self._logZ, self.ess, self.acc_ratio, self.N = self._eng.get()
            self._P = self._params(self._θ)
            self._stale = False

    # θ, ω, logZ: public fields of the reference's struct (smc_samplers.jl:5-27); with the device engine they are read
    # back from the GPU when somebody looks at them
    θ = property(lambda self: (self._refresh(), self._θ)[1], lambda self, v: setattr(self, "_θ", v))
    ω = property(lambda self: (self._refresh(), self._ω)[1], lambda self, v: setattr(self, "_ω", v))
    logZ = property(lambda self: (self._refresh(), self._logZ)[1], lambda self, v: setattr(self, "_logZ", v))

    def _clouds(self):
        return self._cur if self._eng is None else self._eng.clouds()

    # -- helpers
    def _params(self, θ):
        return params_of(self.model, θ)[0]

    def _prior_v(self, θ):
        """(insupport [M], logpdf [M]) of the prior for θ [M, d] — vectorised when the prior offers it"""
        if hasattr(self.prior, "insupport_v"):
            ok = self.prior.insupport_v(θ)
            with np.errstate(invalid="ignore", divide="ignore"):
                return ok, np.where(ok, self.prior.logpdf_v(θ), -math.inf)
        ok = np.array([self.prior.insupport(th) for th in θ])
        return ok, np.array([self.prior.logpdf(th) if o else -math.inf for th, o in zip(θ, ok)])

    def _local(self, v):
        return v[self.lo: self.lo + self.Mloc]

    def _proposals(self, P_local, y):
        """None, or the [len(y), Mloc, 3] proposal coefficients of the local θ-particles for every observation of y"""
        if self.proposal is None:
            return None
        return np.stack([np.asarray(self.proposal(P_local, float(yt)), np.float64).reshape(P_local.shape[0], 3) for yt in np.atleast_1d(y)])

    def _next_epoch(self):
        e = self._epoch
        self._epoch += 1
        self.ctx.set_rng(self.seed, e)
        return e

    def _account(self, batch, steps):
        ms, _ = batch.timing()
        self.stats["sweeps"] += 1
        self.stats["particle_updates"] += self.Mloc * self.N * steps
        self.stats["device_ms"] += ms

    @property
    def x(self):
        return self._clouds().fetch(want_x=True, want_w=False)[0]

    @property
    def w(self):
        return self._clouds().fetch(want_x=False, want_w=True)[1]

    def close(self):
        for b in (self._cur, self._prop):
            if b is not None:
                b.close()
        self._cur = self._prop = None
        if self._eng is not None:
            self._refresh()
            self._eng.close()
            self._eng = None

    def __repr__(self):
        return f"ess     = {round(self.ess, 3)}\nmean(θ) = {expected_parameters(self).ravel()}"


def expected_parameters(smc, reference_style=False):
    """Σ ω_m θ_m as a d×1 matrix (smc_samplers.jl:61-65).  The reference passes the already
    normalised ω through `reweight` (a softmax of ω itself), which is almost the
    unweighted mean; reference_style=True reproduces that."""
    ω = smc.ω
    if reference_style:
        e = np.exp(ω - ω.max())
        ω = e / e.sum()
    return (smc.θ * ω[:, None]).sum(axis=0)[:, None]


def state_means(smc):
    """[M, d]: the weighted state mean `smc.w[m]' * smc.x[m]` of every θ-particle's cloud
    (plotting_utils.jl:120,150), computed on the device(s); the clouds are not read back."""
    return smc.comm.all_gather(smc._clouds().weighted_mean())


def state_variances(smc):
    """([M, d] means, [M, d] variances): `mean(smc.x[m], weights(smc.w[m]))`, `var(smc.x[m], weights(smc.w[m]))` of every
    θ-particle's cloud (the per-θ counterpart of examples/inflation_example.jl:46; population variance), computed on the
    device(s); the clouds are not read back."""
    mean, var = smc._clouds().weighted_moments()
    return smc.comm.all_gather(mean), smc.comm.all_gather(var)


def state_quantiles(smc, p, weighted=True):
    """[M, d, len(p)]: `quantile(smc.x[m], weights(smc.w[m]), p)` (weighted) or `quantile(smc.x[m], p)` of every
    θ-particle's cloud (examples/inflation_example.jl:44,250), computed on the device(s) by a per-cloud radix
    select (docs/SPEC.md §8); the clouds are not read back."""
    return smc.comm.all_gather(smc._clouds().weighted_quantiles(np.atleast_1d(np.asarray(p, np.float64)), weighted))


def get_quantiles(smc, yt, p=(0.25, 0.5, 0.75), weighted=True, component=0):
    """get_quantiles_uc / get_quantiles_ucsv (examples/inflation_example.jl:39-55,241-253): the ω-mixture over θ of
    the per-cloud quantiles of the trend x and of the cycle yt − x.  Returns (xquantiles, cquantiles).
    The cycle's p-quantile is yt minus the trend's (1 − p)-quantile (lower empirical quantiles on both sides)."""
    p = np.atleast_1d(np.asarray(p, np.float64))
    q = state_quantiles(smc, np.concatenate([p, 1.0 - p]), weighted)[:, component, :]
    xq = (smc.ω[:, None] * q[:, : p.size]).sum(axis=0)
    cq = (smc.ω[:, None] * (float(yt) - q[:, p.size:])).sum(axis=0)
    return xq, cq


def _observation_mean_sd(smc):
    """mean and sd of observation(model(θ_m), x̄_m) for every m (state_space_models.jl:96-103,244-247)."""
    xm, P = state_means(smc), (smc._refresh(), smc._P)[1]
    if smc.kind == _lib.LG1D:
        return P[:, 1] * xm[:, 0], np.sqrt(P[:, 3])                    # Normal(B x, sqrt(R))
    if smc.kind == _lib.SV:
        return np.zeros(smc.M), np.exp(0.5 * xm[:, 0])                  # Normal(0, exp(x/2))
    return xm[:, 0], np.exp(0.5 * xm[:, 2])                             # UCSV: Normal(x, exp(lση/2))


def estimated_trend(smc):
    """estimated_trend(smc) (plotting_utils.jl:116-124): Σ_m ω_m mean(observation(model(θ_m), w_m' x_m)); for an IBIS sampler
    the mixture of the Kalman filters' predicted measurements (:114)."""
    if not hasattr(smc, "_cur"):
        from . import ibis as _ibis
        return _ibis.estimated_trend(smc)
    mu, _ = _observation_mean_sd(smc)
    return float(np.sum(smc.ω * mu))


def quantile_smc(smc, p):
    """quantile(smc, p) (plotting_utils.jl:140-157): Σ_m ω_m quantile(observation(model(θ_m), w_m' x_m), p),
    the states integrated out through their weighted means as in the reference."""
    from statistics import NormalDist
    p = np.sort(np.atleast_1d(np.asarray(p, np.float64)))             # sort!(p)  :144
    z = np.array([NormalDist().inv_cdf(float(v)) for v in p])
    mu, sd = _observation_mean_sd(smc)
    return (smc.ω[:, None] * (mu[:, None] + sd[:, None] * z[None, :])).sum(axis=0)


def resample_(smc):
    """resample!(smc) (smc_samplers.jl:74-84): multinomial ancestors on ω; θ, logZ and the state
    clouds (x AND w: D3; deep copies: D4) follow their parents; ω becomes uniform (D5)."""
    smc.ctx.set_rng(smc.seed, 0)
    a = smc.ctx.resample(smc.ω, smc.theta_resampler, stream=0, t=smc._n_resample, purpose=_lib.P_THETA_RESAMPLE)
    smc._n_resample += 1
    # docs/SPEC.md §5b: the ancestors are put in ascending order.  After a resample the θ-particles are exchangeable
    # (uniform ω, every later operation is a mean over slots or a per-slot move with its own random stream), so the slot
    # order carries no information — but slot m lives on rank m // Mloc, and with sorted ancestors the parent of a slot
    # is almost always on the same rank: 4–6 % of the clouds cross GPUs at G = 8 instead of 87 %.
    a = np.sort(a)
    smc.θ = smc.θ[a]
    smc._P = smc._P[a]
    smc.logZ = smc.logZ[a]
    smc.ω = np.full(smc.M, 1.0 / smc.M)
    smc.stats["clouds_moved"] += redistribute(_BatchStore(smc._cur, smc.comm), a, smc.comm)
    smc._params_dirty = True
    return a


class _BatchStore:
    """adapter: the clouds of a device batch as seen by redistribute()"""

    def __init__(self, batch, comm):
        self.batch, self.comm = batch, comm
        self.nbytes = batch.cloud_bytes()

    def pack(self, slots):
        torch = self.comm.torch
        buf = torch.empty(len(slots) * self.nbytes, dtype=torch.uint8, device=self.comm.device)
        self.batch.pack(slots, buf.data_ptr())
        return buf

    def gather(self, local_parents):
        self.batch.gather(local_parents)

    def unpack(self, slots, buf):
        self.batch.unpack(slots, buf.data_ptr())


def redistribute(store, parents, comm):
    """Move whole clouds so that global slot m holds a deep copy of the cloud of parents[m]:
    rank-local parents by an on-device gather, remote parents by pack -> point-to-point -> unpack
    (smc_samplers.jl:82 `smc.x = smc.x[a]`, across GPUs).  Returns the number of clouds received."""
    if comm.world == 1:
        store.gather(np.asarray(parents, np.int32))
        return 0
    local_parents, send, recv = exchange_plan(parents, comm.rank, comm.world)
    sendbufs = {dst: store.pack(slots) for dst, slots in send.items()}   # read the pre-gather clouds
    got = comm.exchange(sendbufs, {src: len(s) for src, s in recv.items()}, store.nbytes)
    store.gather(local_parents)
    moved = 0
    for src, slots in recv.items():
        store.unpack(slots, got[src])
        moved += len(slots)
    return moved


def rejuvenate_(smc, y, ξ=1.0, verbose=False):
    """rejuvenate!(smc, y, ξ, verbose) (smc_samplers.jl:103-148): `chain` PMMH moves per θ-particle;
    each move is one full particle filter over y — all M of them in one batched launch."""
    y = np.ascontiguousarray(y, np.float64)
    M, d = smc.θ.shape
    acc = np.zeros(M, bool)
    Σ, univariate = smc.kernel(smc.θ)                                  # pmmh_kernel = smc.kernel(smc.θ)   :107
    scales = 0.5 * np.arange(smc.chain, 0, -1)                         # 0.5*reverse(1:chain)              :108
    if verbose:
        sys.stdout.write("\t[rejuvenating]")
    if smc._prop is None:
        smc._prop = smc.ctx.batch(smc.kind, smc.Mloc, smc.N)
    ordinal = smc._n_rejuv
    smc._n_rejuv += 1
    _, lp_cur = smc._prior_v(smc.θ)
    for c in range(smc.chain):
        z = np.stack([_lib.rng_normals(smc.seed, ordinal, k, c, _lib.P_MH_PROPOSAL, 0, M) for k in range(d)], axis=1)
        θ_prop = _propose(smc.θ, Σ, univariate, scales[c], z)          # rand(pmmh_kernel(θ[m], scales[c]))  :114
        ok, lp_prop = smc._prior_v(θ_prop)                             # insupport(prior, θ_prop)            :116
        params = smc._params(np.where(ok[:, None], θ_prop, smc.θ))
        smc._next_epoch()
        z_loc = smc._prop.log_likelihood(smc._local(params), y, smc.resampler, stream0=smc.lo,
                                         active=smc._local(ok).astype(np.uint8),
                                         proposal=smc._proposals(smc._local(params), y))   # log_likelihood(N, y, model(θ_prop)) :117-121
        smc._account(smc._prop, y.size)
        logZ_prop = smc.comm.all_gather(z_loc)
        with np.errstate(invalid="ignore"):
            acc_ratio = ξ * (logZ_prop - smc.logZ) + (lp_prop - lp_cur)                # :123-127
            u = _lib.rng_uniforms01(smc.seed, ordinal, 0, c, _lib.P_MH_ACCEPT, M)
            with np.errstate(divide="ignore"):
                accept = ok & (logZ_prop + lp_prop > -math.inf) & (np.log(u) < acc_ratio)   # :129
        smc.logZ = np.where(accept, logZ_prop, smc.logZ)                                # :130-133
        smc.θ = np.where(accept[:, None], θ_prop, smc.θ)
        smc._P = np.where(accept[:, None], params, smc._P)
        lp_cur = np.where(accept, lp_prop, lp_cur)
        smc._cur.accept(smc._prop, smc._local(accept).astype(np.uint8))
        acc |= accept
    smc.ω = np.full(M, 1.0 / M)                                         # ω[m] = 1.0 (then normalised)       :139
    smc.acc_ratio = float(acc.sum()) / M                                # :142
    smc._params_dirty = True
    if verbose:
        sys.stdout.write("\tacc_rate: %1.5f" % smc.acc_ratio)
    return smc


def exchange_(smc, y, verbose=False):
    """exchange!(smc, y, verbose) (smc_samplers.jl:163-189): if the acceptance ratio fell below
    acc_threshold and N <= 4096, double N, re-filter every θ and reweight by exp(new_logZ - logZ)."""
    if not (smc.acc_ratio < smc.acc_threshold):
        return
    if smc.N > 4096:
        sys.stdout.write("\n\t[cannot exceed max state particles]")
        return
    smc.N *= 2
    if verbose:
        sys.stdout.write("\t%d particles added" % smc.N)
    y = np.ascontiguousarray(y, np.float64)
    old, oldp = smc._cur, smc._prop
    smc._cur = smc.ctx.batch(smc.kind, smc.Mloc, smc.N)
    smc._prop = None
    old.close()
    if oldp is not None:
        oldp.close()
    smc._next_epoch()
    z_loc = smc._cur.log_likelihood(smc._local(smc._P), y, smc.resampler, stream0=smc.lo, proposal=smc._proposals(smc._local(smc._P), y))
    smc._account(smc._cur, y.size)
    new_logZ = smc.comm.all_gather(z_loc)
    _, smc.ω, smc.ess = smc.ctx.normalize(new_logZ - smc.logZ)
    smc.logZ = new_logZ


def density_tempered(smc, y, verbose=True):
    """density_tempered(smc, y) (smc_samplers.jl:222-281), Duan & Fulop's density-tempered SMC."""
    y = np.ascontiguousarray(y, np.float64)
    if smc._eng is not None:                                            # the whole loop below, on the device(s)
        smc._engine_data(y)
        stages = smc._eng.density_tempered()
        smc.schedule = [(ξ, ess) for ξ, ess, _ in stages]
        smc.acceptance = [ar for _, _, ar in stages if ar >= 0.0]     # acc_rate of every rejuvenation, as the reference prints it
        smc._stale = True
        smc._refresh()
        if verbose:                                                     # the trace format of smc_samplers.jl:207-214
            for ξ, ess, ar in stages:
                sys.stdout.write("ξ = %1.5f\tess = %4.3f" % (ξ, ess) + ("\t[rejuvenating]\tacc_rate: %1.5f\n" % ar if ar >= 0.0 else "\n"))
        return smc
    smc._next_epoch()
    z_loc = smc._cur.log_likelihood(smc._local(smc._P), y, smc.resampler, stream0=smc.lo,
                                    proposal=smc._proposals(smc._local(smc._P), y))          # :223-229
    smc._account(smc._cur, y.size)
    smc.logZ = smc.comm.all_gather(z_loc)
    _, smc.ω, smc.ess = smc.ctx.normalize(smc.logZ)                    # :232
    ξ = 0.0
    smc.schedule, smc.acceptance = [], []
    while ξ < 1.0:
        resample_flag = True
        lower = oldξ = ξ
        upper = 2.0
        newξ = None
        while upper - lower > 1.0e-6:                                   # bisection for ξ                    :240-258
            newξ = (upper + lower) / 2.0
            logω = (newξ - oldξ) * smc.logZ
            _, smc.ω, smc.ess = smc.ctx.normalize(logω)
            if smc.ess == smc.ess_min:
                break
            elif smc.ess < smc.ess_min:
                upper = newξ
            else:
                lower = newξ
        if newξ >= 1.0:                                                 # corner solution                    :261-266
            resample_flag = False
            newξ = 1.0
            logω = (newξ - oldξ) * smc.logZ
            _, smc.ω, smc.ess = smc.ctx.normalize(logω)
        ξ = newξ
        smc.schedule.append((ξ, smc.ess))
        if verbose:
            sys.stdout.write("ξ = %1.5f\tess = %4.3f" % (ξ, smc.ess))
        if resample_flag:
            resample_(smc)                                              # :272
            rejuvenate_(smc, y, ξ, verbose)                             # :275
            smc.acceptance.append(smc.acc_ratio)
        if verbose:
            sys.stdout.write("\n")
    return smc


def smc2(smc, y):
    """smc²(smc, y) (smc_samplers.jl:288-301): M bootstrap filters at the first observation."""
    y = np.ascontiguousarray(y, np.float64)
    if smc._eng is not None:
        smc._engine_data(y)
        smc._eng.smc2_init()
        smc._stale = True
        return smc
    smc._next_epoch()
    lm_loc, _ = smc._cur.init(smc._local(smc._P), y[0], stream0=smc.lo)
    smc._account(smc._cur, 1)
    logmu = smc.comm.all_gather(lm_loc)
    smc.logZ = logmu.copy()                                             # smc.logZ = smc.ω                   :297
    _, smc.ω, smc.ess = smc.ctx.normalize(logmu)                        # :298
    return smc


def smc2_step(smc, y, t, verbose=True):
    """smc²!(smc, y, t) (smc_samplers.jl:308-340); t is the 0-based index of the new observation."""
    if smc._eng is not None:                                            # one call: (resample!, rejuvenate!, exchange!,) M filter steps, reweight
        smc._engine_data(y)
        if verbose:
            sys.stdout.write("t = %4d\tess = %4.3f" % (t, smc.ess))
        smc.ess, smc.rejuvenated = smc._eng.smc2_step(t)
        smc._stale = True
        if smc.rejuvenated:
            _, _, _, _, smc.acc_ratio, smc.N = smc._eng.get(False, False, False)
            if verbose:
                sys.stdout.write("\t[rejuvenating]\tacc_rate: %1.5f" % smc.acc_ratio)
        if verbose:
            sys.stdout.write("\n")
        return smc
    y = np.ascontiguousarray(y, np.float64)
    if verbose:
        sys.stdout.write("t = %4d\tess = %4.3f" % (t, smc.ess))
    smc.rejuvenated = False
    if smc.ess < smc.ess_min:                                           # :312
        resample_(smc)                                                  # :314
        rejuvenate_(smc, y[:t], 1.0, verbose)                           # rejuvenate!(smc, y[1:t-1])         :317
        exchange_(smc, y[:t], verbose)                                  # :320
        smc.rejuvenated = True
    with np.errstate(divide="ignore"):
        logω = np.log(smc.ω)                                            # :324
    prop = smc._proposals(smc._local(smc._P), y[t])
    lm_loc, _ = smc._cur.step(y[t], smc.resampler, params=smc._local(smc._P) if smc._params_dirty else None,
                              proposal=None if prop is None else prop[0])                  # M × bootstrap_filter! (particle_filter! when guided)  :325-331
    smc._params_dirty = False
    smc._account(smc._cur, 1)
    logmu = smc.comm.all_gather(lm_loc)
    logω = logω + logmu                                                 # :333
    smc.logZ = smc.logZ + logmu                                         # :334
    _, smc.ω, smc.ess = smc.ctx.normalize(logω)                         # :338
    if verbose:
        sys.stdout.write("\n")
    return smc
