"""Particle filters — host mirror of /root/reference/src/particles.jl over libsmcb200.

    normalize(logw)                    -> (logμ, w, ess)        particles.jl:5-15
    resample(w, N=len(w))              -> ancestors (0-based)   particles.jl:17-19
    bootstrap_filter(N, y, model)      -> (x, w, logμ)          particles.jl:87-105
    bootstrap_filter_(x, w, y, model)  -> (logμ, w, ess)        particles.jl:107-129  (Julia: bootstrap_filter!)
    log_likelihood(N, y, model)        -> (x, w, logZ)          particles.jl:132-147

`x` and `w` are handles on the device-resident cloud: they behave like numpy arrays
(np.asarray, indexing, quantiles) and copy to the host only when read.  All compute is CUDA; there
is no CPU fallback.
"""
import numpy as np

from . import _lib

_RESAMPLERS = {"multinomial": _lib.MULTINOMIAL, "stratified": _lib.STRATIFIED, "systematic": _lib.SYSTEMATIC,
               _lib.MULTINOMIAL: _lib.MULTINOMIAL, _lib.STRATIFIED: _lib.STRATIFIED, _lib.SYSTEMATIC: _lib.SYSTEMATIC}
_DEFAULT = None


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

    def _stale(self):
        return getattr(self._ctx, "_gen", 0) != self._gen

    def numpy(self):
        if self._host is not None and (self._stale() or self._host_t == self._ctx._T):
            return self._host
        if self._stale():
            raise RuntimeError("this particle cloud was replaced by a later bootstrap_filter / log_likelihood on the same context")
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
    if N is not None and int(N) != w.size:
        raise NotImplementedError("resample(w, N) with N != length(w) is not used anywhere on the reference's path")
    return ctx.resample(w, resampler_id(resampler), stream=stream, t=t, purpose=purpose)


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


def particle_filter(N, y, model, proposal=None, *, ctx=None, stream=0):
    """x, w, logμ = particle_filter(N, y[1], model, proposal)  — particles.jl:28-51.  With proposal = nothing
    (the only form the reference's own example uses, examples/inflation_example.jl:164,178) the initial step IS
    the bootstrap one: it draws from initial_dist and weights by the observation density.  A guided proposal
    functor (particles.jl:73-78, inconsistent in the reference: `proposal(model, xp)` vs `proposal(xp)`) is not
    built: NotImplementedError, never a silent bootstrap run."""
    if proposal is not None:
        raise NotImplementedError("guided proposals (particles.jl:73-78) are not built; pass proposal=None for the bootstrap filter")
    return bootstrap_filter(N, y, model, ctx=ctx, stream=stream)


def particle_filter_(states, weights, y, model, proposal=None, *, resampler="multinomial"):
    """logμ, w, ess = particle_filter!(x, w, y[t], model, proposal)  — particles.jl:53-84; proposal = nothing is
    bootstrap_filter! (the reference would call `nothing(model, x)` there, particles.jl:73)."""
    if proposal is not None:
        raise NotImplementedError("guided proposals (particles.jl:73-78) are not built; pass proposal=None for the bootstrap filter")
    return bootstrap_filter_(states, weights, y, model, resampler=resampler)


def quantile(x, w_or_p, p=None):
    """quantile(x, p) (README.md:41,51: every particle counts once) or quantile(x, weights(w), p)
    (examples/inflation_example.jl:44) of a cloud that lives on the device — computed there, nothing is
    read back.  Lower empirical quantile (no interpolation); one row per state component."""
    if hasattr(x, "θ"):                      # quantile(smc::SMC, p)  plotting_utils.jl:140-157
        from .smc_samplers import quantile_smc
        return quantile_smc(x, w_or_p)
    weighted = p is not None
    probs = np.atleast_1d(np.asarray(p if weighted else w_or_p, np.float64))
    if x._stale():
        raise RuntimeError("stale particle cloud")
    q = x._ctx.summary(probs, weighted=weighted)[2]
    return q[0] if q.shape[0] == 1 else q.T


def weighted_mean_var(x, w=None):
    """(mean, var) of the cloud x under the weights w (mean(x, weights(w)), var(x, weights(w)):
    examples/inflation_example.jl:46), on the device."""
    if x._stale():
        raise RuntimeError("stale particle cloud")
    m, v, _ = x._ctx.summary((), weighted=w is not None)
    return (m[0], v[0]) if m.size == 1 else (m, v)


def log_likelihood(N, y, model, *, resampler="multinomial", ctx=None, stream=0):
    """x, w, logZ = log_likelihood(N, y, model)  — particles.jl:132-147, one call for the whole series."""
    ctx = ctx or default_context()
    logZ = ctx.log_likelihood(model.kind, model.params(), int(N), np.asarray(y, np.float64), resampler_id(resampler), stream)
    x, w = _handles(ctx)
    return x, w, logZ
