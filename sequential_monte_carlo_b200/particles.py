"""Particle filters — host mirror of /root/reference/src/particles.jl over libsmcb200.

    normalize(logw)                    -> (logμ, w, ess)        particles.jl:5-15
    resample(w, N=len(w))              -> ancestors (0-based)   particles.jl:17-19
    bootstrap_filter(N, y, model)      -> (x, w, logμ)          particles.jl:87-105
    bootstrap_filter_(x, w, y, model)  -> (logμ, w, ess)        particles.jl:107-129  (Julia: bootstrap_filter!)
    log_likelihood(N, y, model)        -> (x, w, logZ)          particles.jl:132-147
    particle_filter(N, y, model, proposal) / particle_filter_(x, w, y, model, proposal)   particles.jl:28-84
                                       guided filter, affine-Gaussian proposals on the device (docs/SPEC.md §10)

`x` and `w` are handles on the device-resident cloud: they behave like numpy arrays
(np.asarray, indexing, quantiles) and copy to the host only when read.  All compute is CUDA; there
is no CPU fallback.
"""
import numpy as np

from . import _lib

_RESAMPLERS = {"multinomial": _lib.MULTINOMIAL, "stratified": _lib.STRATIFIED, "systematic": _lib.SYSTEMATIC,
               _lib.MULTINOMIAL: _lib.MULTINOMIAL, _lib.STRATIFIED: _lib.STRATIFIED, _lib.SYSTEMATIC: _lib.SYSTEMATIC}
_DEFAULT = None
_SMALL_N_MAX = 8192     # clouds up to this size run block-resident (csrc/smcb_batch.cu) when a whole series is filtered in one call


def resampler_id(r):
    try:
        return _RESAMPLERS[r]
    except KeyError:
        raise ValueError(f"unknown resampler {r!r}: multinomial (reference), stratified or systematic") from None


def default_context():
    """The process-wide context (GPU 0, seed 1998 as in examples/inflation_example.jl:57) used when no ctx is given."""
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = _lib.Context(0, 1998)
    return _DEFAULT


def set_default_context(ctx):
    global _DEFAULT
    _DEFAULT = ctx


class _DeviceArray:
    """Lazy host view of the filter state living in ctx (valid until the filter is re-initialised)."""

    def __init__(self, ctx, gen, which):
        self._ctx, self._gen, self._which, self._host, self._host_t = ctx, gen, which, None, -1
        self._T0 = getattr(ctx, "_T", 0)    # the step a weights handle belongs to (bootstrap_filter! returns a NEW w every step)

    def _stale(self):
        return getattr(self._ctx, "_gen", 0) != self._gen

    def numpy(self):
        if self._host is not None and (self._stale() or self._host_t == self._ctx._T):
            return self._host
        if self._stale():
            raise RuntimeError("this particle cloud was replaced by a later bootstrap_filter / log_likelihood on the same context")
        if self._which == "w" and self._ctx._T != self._T0:
            if self._host is not None:
                return self._host          # the weights of its own step, read while they were current
            raise RuntimeError("these weights belong to an earlier step of the filter and were not read while they were current: "
                               "use the w returned by the latest bootstrap_filter! / particle_filter!")
        if self._which == "x":
            x, _, _ = self._ctx.fetch_state(want_x=True, want_w=False)
            self._host = x[0] if x.shape[0] == 1 else np.ascontiguousarray(x.T)   # UCSV: N rows of 3 (state_space_models.jl:229-231)
        else:
            _, w, _ = self._ctx.fetch_state(want_x=False, want_w=True)
            self._host = w
        self._host_t = self._ctx._T
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, i):
        return self.numpy()[i]

    def __len__(self):
        return self._ctx._N

    @property
    def shape(self):
        return self.numpy().shape

    def __repr__(self):
        return f"<device {self._which} of {self._ctx._N} particles on cuda:{self._ctx.device}>"


def _handles(ctx):
    ctx._gen = getattr(ctx, "_gen", 0) + 1
    return _DeviceArray(ctx, ctx._gen, "x"), _DeviceArray(ctx, ctx._gen, "w")


def normalize(logw, ctx=None):
    """(logμ, w, ess) = normalize(logw)  — particles.jl:5-15 (also the undefined `reweight`)."""
    ctx = ctx or default_context()
    return ctx.normalize(np.asarray(logw, np.float64))


reweight = normalize


def resample(w, N=None, *, resampler="multinomial", ctx=None, stream=0, t=0, purpose=3):
    """ancestors = resample(w)  — particles.jl:17-19.  0-based indices (Julia's are 1-based)."""
    ctx = ctx or default_context()
    w = np.asarray(w, np.float64)
    return ctx.resample(w, resampler_id(resampler), stream=stream, t=t, purpose=purpose, n_out=N)


def bootstrap_filter(N, y, model, *, ctx=None, stream=0):
    """x, w, logμ = bootstrap_filter(N, y[1], model)  — particles.jl:87-105"""
    ctx = ctx or default_context()
    logmu, _ = ctx.bootstrap_init(model.kind, model.params(), int(N), float(y), stream)
    x, w = _handles(ctx)
    return x, w, logmu


def bootstrap_filter_(states, weights, y, model, *, resampler="multinomial"):
    """logμ, w, ess = bootstrap_filter!(x, w, y[t], model)  — particles.jl:107-129.
    `states` is advanced in place on the device; as in the reference the caller rebinds w."""
    ctx = states._ctx
    if states._stale():
        raise RuntimeError("stale particle cloud")
    logmu, ess = ctx.bootstrap_step(float(y), resampler_id(resampler), model.params())
    return logmu, _DeviceArray(ctx, ctx._gen, "w"), ess


class AffineGaussianProposal:
    """The proposal family the device evaluates (docs/SPEC.md §10): x' ~ N(c0 + c1·xp, c2²), c2 > 0.  Stands in for the
    reference's closure `proposal(model, xp) -> distribution` (particles.jl:73,78).  Called as proposal(model, y) it
    returns the coefficients (c0, c1, c2) of the step that assimilates y; subclass or pass any callable with that
    signature for a proposal that looks at the observation."""

    def __init__(self, c0, c1, c2):
        self.c0, self.c1, self.c2 = float(c0), float(c1), float(c2)

    def __call__(self, model, y):
        return self.c0, self.c1, self.c2


def locally_optimal_proposal(model, y):
    """p(x' | xp, y) of a univariate LinearModel, the variance-optimal proposal: precision 1/Q + B²/R, mean
    s²(A·xp/Q + B·y/R).  Use as `particle_filter_(x, w, y, model, locally_optimal_proposal)`."""
    A, B, Q, R = (float(v) for v in np.asarray(model.params(), np.float64).ravel()[:4])
    s2 = 1.0 / (1.0 / Q + B * B / R)
    return s2 * B * float(y) / R, s2 * A / Q, float(np.sqrt(s2))


class UCSVTrendProposal:
    """The guided move of the UCSV model (docs/SPEC.md §10b): the two log-volatilities move by the transition and the trend by
    the conditionally optimal Gaussian move, tempered by κ in [0, 1] — x' ~ N(x + g·(y − x), (1 − g)·σε²) with
    g = κ·σε²/(σε² + ση'²), σε = exp(le/2) of the parent, ση'² = exp(ln') of the new state.  κ = 0 is the bootstrap move,
    κ = 1 is p(x' | x, le, ln', y), after which the weight no longer depends on x'.  Use as
    `particle_filter_(x, w, y, model, UCSVTrendProposal(1.0))`."""

    def __init__(self, kappa=1.0):
        if not 0.0 <= float(kappa) <= 1.0:
            raise ValueError("kappa must lie in [0, 1]")
        self.kappa = float(kappa)

    def __call__(self, model, y):
        return self.kappa, 0.0, 1.0


_GUIDED_KINDS = (_lib.LG1D, _lib.SV, _lib.UCSV)
_GUIDED_BATCH_MAX = 8192   # guided filters up to this size run on the batched engine (any resampler), larger ones on the single filter


class _GuidedCloud:
    """x or w of a guided filter: the cloud lives in a one-θ batch on the device (guided filters run on the batched
    engine for N <= 8192; larger guided filters use the single filter and its _DeviceArray handles) and is copied to the
    host only when read."""

    def __init__(self, batch, which):
        self._batch, self._which, self._host, self._host_t = batch, which, None, -1

    def numpy(self):
        t = getattr(self._batch, "_t", 0)
        if self._host is None or self._host_t != t:
            x, w, _ = self._batch.fetch(want_x=self._which == "x", want_w=self._which == "w")
            if self._which == "x":
                self._host = x[0, 0] if x.shape[1] == 1 else np.ascontiguousarray(x[0].T)   # UCSV: N rows of 3
            else:
                self._host = w[0]
            self._host_t = t
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, i):
        return self.numpy()[i]

    def __len__(self):
        return self._batch.N

    @property
    def shape(self):
        return (self._batch.N,) if self._which == "w" or self._batch.d == 1 else (self._batch.N, self._batch.d)


def _check_family(model, proposal):
    """the device evaluates one proposal family per model: affine-Gaussian for the one-dimensional models (docs/SPEC.md §10), the
    tempered optimal trend move for UCSV (§10b); anything else is refused, never run as a bootstrap filter"""
    if model.kind not in _GUIDED_KINDS:
        raise NotImplementedError("guided proposals are built for LinearModel, StochasticVolatility (affine-Gaussian) and UCSV (UCSVTrendProposal)")
    if (model.kind == _lib.UCSV) != isinstance(proposal, UCSVTrendProposal):
        raise NotImplementedError("UCSV takes a UCSVTrendProposal (docs/SPEC.md §10b); the one-dimensional models take the affine-Gaussian family (§10)")


def _proposal_coefficients(proposal, model, y):
    c = proposal(model, y) if callable(proposal) else proposal
    c = np.asarray(c, np.float64).ravel()
    if c.size != 3:
        raise TypeError("proposal must be an AffineGaussianProposal or a callable (model, y) -> (c0, c1, c2) "
                        "describing x' ~ N(c0 + c1·xp, c2²) (docs/SPEC.md §10); arbitrary closures cannot run on the device")
    return c


def particle_filter(N, y, model, proposal=None, *, ctx=None, stream=0):
    """x, w, logμ = particle_filter(N, y[1], model, proposal)  — particles.jl:28-51.  With proposal = nothing
    (the form the reference's own example uses, examples/inflation_example.jl:164,178) this is bootstrap_filter.
    With a proposal the initial step is STILL the bootstrap one — it draws from initial_dist and weights by the
    observation density (:40-42; the "+ logpdf(initial_dist, x)" of :44 has lost its "- logpdf(proposal)" partner,
    commented out at :45, and is ruled a defect, docs/SPEC.md §10) — but the cloud is placed on the batched engine
    so that particle_filter_ can continue it with guided moves (one-dimensional models, N <= 8192)."""
    if proposal is None:
        return bootstrap_filter(N, y, model, ctx=ctx, stream=stream)
    ctx = ctx or default_context()
    _check_family(model, proposal)
    if int(N) > _GUIDED_BATCH_MAX:      # large clouds: the grid-wide single filter (guided steps need a sorted resampler there)
        return bootstrap_filter(N, y, model, ctx=ctx, stream=stream)
    b = ctx.batch(model.kind, 1, int(N))
    logmu, _ = b.init(np.asarray(model.params(), np.float64).reshape(1, -1), float(y), stream0=stream)
    b._t = 0
    return _GuidedCloud(b, "x"), _GuidedCloud(b, "w"), float(logmu[0])


def particle_filter_(states, weights, y, model, proposal=None, *, resampler="multinomial"):
    """logμ, w, ess = particle_filter!(x, w, y[t], model, proposal)  — particles.jl:53-84.  proposal = nothing is
    bootstrap_filter! (the reference would call `nothing(model, x)` there, :73).  Otherwise x' ~ proposal and
    logw = logpdf(observation(x'), y) + logpdf(transition(xp), x') - logpdf(proposal(xp), x')  (:73-78), evaluated on the
    device for the affine-Gaussian family (AffineGaussianProposal, locally_optimal_proposal)."""
    if proposal is None:
        if isinstance(states, _GuidedCloud):
            b = states._batch
            lm, es = b.step(float(y), resampler_id(resampler), np.asarray(model.params(), np.float64).reshape(1, -1))
            b._t += 1
            return float(lm[0]), _GuidedCloud(b, "w"), float(es[0])
        return bootstrap_filter_(states, weights, y, model, resampler=resampler)
    if isinstance(states, (_DeviceArray, _GuidedCloud)):
        _check_family(model, proposal)
    c = _proposal_coefficients(proposal, model, y)
    if isinstance(states, _DeviceArray):   # the large-N single filter: guided move kernel of the grid-wide path
        ctx = states._ctx
        if states._stale():
            raise RuntimeError("stale particle cloud")
        logmu, ess = ctx.guided_step(float(y), c, resampler_id(resampler), model.params())
        return logmu, _DeviceArray(ctx, ctx._gen, "w"), ess
    if not isinstance(states, _GuidedCloud):
        raise RuntimeError("guided steps continue a cloud created by particle_filter(N, y, model, proposal)")
    b = states._batch
    lm, es = b.step(float(y), resampler_id(resampler), np.asarray(model.params(), np.float64).reshape(1, -1), proposal=c.reshape(1, 3))
    b._t += 1
    return float(lm[0]), _GuidedCloud(b, "w"), float(es[0])


def guided_log_likelihood(N, y, model, proposal, *, resampler="multinomial", ctx=None, stream=0):
    """x, w, logZ of the guided filter over the whole series in ONE launch (the loop of
    examples/inflation_example.jl:164-171 with a proposal): bootstrap initial step, guided steps for t >= 2."""
    ctx = ctx or default_context()
    y = np.asarray(y, np.float64)
    _check_family(model, proposal)
    prop = np.stack([_proposal_coefficients(proposal, model, yt) for yt in y]).reshape(y.size, 1, 3)
    if int(N) > _GUIDED_BATCH_MAX:
        logZ = ctx.guided_log_likelihood(model.kind, model.params(), int(N), y, prop.reshape(-1, 3), resampler_id(resampler), stream)
        x, w = _handles(ctx)
        return x, w, logZ
    b = ctx.batch(model.kind, 1, int(N))
    z = b.log_likelihood(np.asarray(model.params(), np.float64).reshape(1, -1), y, resampler_id(resampler), stream0=stream, proposal=prop)
    b._t = y.size - 1
    return _GuidedCloud(b, "x"), _GuidedCloud(b, "w"), float(z[0])


def quantile(x, w_or_p, p=None):
    """quantile(x, p) (README.md:41,51: every particle counts once) or quantile(x, weights(w), p)
    (examples/inflation_example.jl:44) of a cloud that lives on the device — computed there, nothing is
    read back.  Lower empirical quantile (no interpolation); one row per state component."""
    if hasattr(x, "θ"):                      # quantile(smc::SMC, p) / quantile(ibis::IBIS, p)  plotting_utils.jl:126-157
        if not hasattr(x, "_cur"):
            from . import ibis as _ibis
            return _ibis.quantile(x, w_or_p)
        from .smc_samplers import quantile_smc
        return quantile_smc(x, w_or_p)
    weighted = p is not None
    probs = np.atleast_1d(np.asarray(p if weighted else w_or_p, np.float64))
    if isinstance(x, _GuidedCloud):           # a guided filter's cloud lives in a one-θ batch: per-cloud radix select there
        return x._batch.weighted_quantiles(probs, weighted=weighted)[0, 0]
    if x._stale():
        raise RuntimeError("stale particle cloud")
    q = x._ctx.summary(probs, weighted=weighted)[2]
    return q[0] if q.shape[0] == 1 else q.T


def weighted_mean_var(x, w=None):
    """(mean, var) of the cloud x under the weights w (mean(x, weights(w)), var(x, weights(w)):
    examples/inflation_example.jl:46), on the device."""
    if isinstance(x, _GuidedCloud):
        if w is None:
            raise NotImplementedError("unweighted moments of a guided filter's cloud: pass its weights (the resampled cloud is weighted)")
        m, v = x._batch.weighted_moments()
        return m[0, 0], v[0, 0]
    if x._stale():
        raise RuntimeError("stale particle cloud")
    m, v, _ = x._ctx.summary((), weighted=w is not None)
    return (m[0], v[0]) if m.size == 1 else (m, v)


def log_likelihood(N, y, model, *, resampler="multinomial", ctx=None, stream=0):
    """x, w, logZ = log_likelihood(N, y, model)  — particles.jl:132-147, one call for the whole series."""
    ctx = ctx or default_context()
    y = np.asarray(y, np.float64)
    if int(N) <= _SMALL_N_MAX and y.size > 1 and model.kind <= _lib.UCSV:    # (multivariate linear models run on the single filter)
        # a small cloud (BASELINE configs[0]: N = 1024, T = 100) lives in one CTA for the whole series: ONE launch of the
        # block-resident engine instead of two to four grid-wide launches per observation (same arithmetic, same x, w, logZ)
        b = ctx.batch(model.kind, 1, int(N))
        z = b.log_likelihood(np.asarray(model.params(), np.float64).reshape(1, -1), y, resampler_id(resampler), stream0=stream)
        b._t = y.size - 1
        ctx._gen = getattr(ctx, "_gen", 0) + 1          # as on the single filter: earlier handles of this context go stale
        return _GuidedCloud(b, "x"), _GuidedCloud(b, "w"), float(z[0])
    logZ = ctx.log_likelihood(model.kind, model.params(), int(N), y, resampler_id(resampler), stream)
    x, w = _handles(ctx)
    return x, w, logZ
