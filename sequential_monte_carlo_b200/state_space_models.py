"""State-space models — host mirror of /root/reference/src/state_space_models.jl.

A model is (kind, 8 parameters) handed to the device functors of csrc/smcb_models.cuh; nothing is
evaluated on the host.  Both spellings of the reference are provided: the on-disk
`UnivariateLinearGaussian(A=…, B=…, Q=…, R=…)`, `unobserved_components`, `UCSV`, and the README /
example spelling `StateSpaceModel(LinearGaussian(A,B,Q,R,x0),(1,1))`, `UC(...)`.
"""
import numpy as np

from . import _lib


class ParamColumn(np.ndarray):
    """One component of θ for ALL θ-particles at once (length M).  The samplers call the user's
    `model(θ)` once with θ[k] = ParamColumn instead of M times with scalars (the reference's
    `model.(θ)` broadcast, ibis.jl:37); model constructors accept these and keep arrays."""

    def __new__(cls, a):
        return np.asarray(a, dtype=np.float64).view(cls)


def _field(name, v):
    if isinstance(v, ParamColumn):
        return v
    if np.ndim(v) != 0:
        raise NotImplementedError(f"{name}: multivariate linear models are outside the accelerated path (SURVEY §8f N4)")
    return float(v)


def _stack_params(vals):
    """[k] scalars / ParamColumns -> [8] list or [M, 8] array"""
    cols = [v for v in vals if isinstance(v, ParamColumn)]
    if not cols:
        return [float(v) for v in vals]
    M = len(cols[0])
    out = np.zeros((M, len(vals)))
    for k, v in enumerate(vals):
        out[:, k] = np.asarray(v)
    return out


class _ModelMeta(type):
    def __call__(cls, *args, **kwargs):
        if cls is StateSpaceModel:  # README constructor: StateSpaceModel(spec, dims) -> spec
            spec = args[0] if args else kwargs.get("spec")
            dims = args[1] if len(args) > 1 else kwargs.get("dims")
            if not isinstance(spec, StateSpaceModel):
                raise TypeError("StateSpaceModel(spec, dims): spec must be a model such as LinearGaussian(...) or UCSV(...)")
            if dims is not None and tuple(dims) != (spec.state_dim, 1):
                raise ValueError(f"dims {tuple(dims)} do not match the model's ({spec.state_dim}, 1)")
            return spec
        return super().__call__(*args, **kwargs)


class StateSpaceModel(metaclass=_ModelMeta):
    """abstract type StateSpaceModel (state_space_models.jl:9).  `StateSpaceModel(spec, dims)` — the
    README constructor (README.md:12-15) — returns `spec` itself after checking `dims`."""
    kind = None
    state_dim = 1

    def params(self):
        raise NotImplementedError

    def params8(self):
        return _lib.params8(self.params())


class LinearModel(StateSpaceModel):
    """struct LinearModel (state_space_models.jl:46-58), univariate methods (:74-109).
    x[t] ~ N(A x[t-1], Q), y[t] ~ N(B x[t], R), x[1] ~ N(x0, σ0); Q, R, σ0 are variances."""
    kind = _lib.LG1D

    def __init__(self, A, B, Q, R, x0=0.0, σ0=1.0):
        self.A, self.B, self.Q, self.R, self.x0, self.σ0 = (_field(n, v) for n, v in (("A", A), ("B", B), ("Q", Q), ("R", R),
                                                                                        ("x0", x0), ("σ0", σ0)))

    def params(self):
        return _stack_params([self.A, self.B, self.Q, self.R, self.x0, self.σ0])

    def __repr__(self):
        return f"LinearModel(A={self.A}, B={self.B}, Q={self.Q}, R={self.R}, x0={self.x0}, σ0={self.σ0})"


def UnivariateLinearGaussian(*, A, B, Q, R, x0=0.0, σ0=1.0):
    """state_space_models.jl:74-77"""
    return LinearModel(A, B, Q, R, x0, σ0)


def LinearGaussian(A, B, Q, R, x0=0.0, σ0=1.0):
    """README spelling: LinearGaussian(θ[1], 1.0, θ[2], θ[3], 0.0)  (README.md:12-15)"""
    return LinearModel(A, B, Q, R, x0, σ0)


def unobserved_components(σε=None, ση=None, x0=0.0, **kw):
    """local level model (state_space_models.jl:119-128): LG1D with A=B=1, Q=σε, R=ση, σ0=σε"""
    σε = kw.get("sigma_eps", σε)
    ση = kw.get("sigma_eta", ση)
    return LinearModel(1.0, 1.0, σε, ση, x0, σε)


def UC(x0, σε, ση):
    """Example spelling `UC(θ...)` (examples/inflation_example.jl:28-31).  `UC` exists nowhere on disk; its positional order
    is fixed by the example's prior `[Normal(3,2), Uniform(0,4), Uniform(0,4)]` (:33-37): the level first, then the two
    variances (a Normal(3,2) draw can be negative, so it cannot be a variance)."""
    return unobserved_components(σε, ση, x0)


class MultivariateLinearModel(StateSpaceModel):
    """LinearModel with a vector state and a scalar observation (state_space_models.jl:137-186):
    x[t] ~ N(A x[t-1], Q), y[t] ~ N(B x[t], R), x[0] ~ N(X0, Σ0).  Served by the matrix Kalman filter on the device
    (kalman_filter.kalman_filter / log_likelihood, kalman_filter.jl:3-27,55-70; state dimension <= 4) and, for d = 2..4, by the
    particle filters of the large-N single filter (bootstrap_filter / bootstrap_filter! / log_likelihood: device functor
    ModelMVLG<d>, csrc/smcb_models.cuh; docs/SPEC.md §4b).  Q and Σ0 may be positive semi-definite (hodrick_prescott's Q has a
    zero pivot, where the reference's own MvNormal(A*x, Q) throws); R is a variance, as in the reference's Kalman filter."""

    def __init__(self, A, B, Q, R, X0=None, Σ0=None):
        self.A = np.atleast_2d(np.asarray(A, np.float64))
        d = self.A.shape[0]
        if self.A.shape != (d, d) or not 1 <= d <= 4:
            raise ValueError("A must be square with 1 <= d <= 4")
        self.B = np.asarray(B, np.float64).reshape(1, d)
        self.Q = np.asarray(Q, np.float64).reshape(d, d)
        self.R = np.asarray(R, np.float64).reshape(1)
        self.x0 = np.zeros(d) if X0 is None else np.asarray(X0, np.float64).reshape(d)
        self.σ0 = np.eye(d) if Σ0 is None else np.asarray(Σ0, np.float64).reshape(d, d)
        self.state_dim = d
        self.kind = (None, None, _lib.MVLG2, _lib.MVLG3, _lib.MVLG4)[d]    # d = 1: use LinearGaussian (the univariate kind)

    def block(self):
        """[3d² + 2d + 1] row-major block A, B, Q, R, x0, Σ0 (include/smcb200.h, smcb_kalman_mv_batch_*)"""
        return np.concatenate([self.A.ravel(), self.B.ravel(), self.Q.ravel(), self.R, self.x0, self.σ0.ravel()])

    def params(self):
        """the parameter block the particle filters take for the multivariate kinds (include/smcb200.h: SMCB_MVLG2..4)"""
        if self.kind is None:
            raise NotImplementedError("a one-dimensional multivariate model: use LinearGaussian(A, B, Q, R, x0, σ0)")
        return self.block()

    def __repr__(self):
        return f"MultivariateLinearModel(d={self.state_dim})"


def MultivariateLinearGaussian(*, A, B, Q, R, X0=None, Σ0=None):
    """state_space_models.jl:137-154"""
    return MultivariateLinearModel(A, B, Q, R, X0, Σ0)


def hodrick_prescott(*, λ, y, init_cov=1000.0):
    """Hodrick–Prescott trend model (state_space_models.jl:187-202): x[t] ~ N(2x[t-1] - x[t-2], 1/λ), y[t] ~ N(x[t], 1)"""
    y = np.asarray(y, np.float64)
    return MultivariateLinearGaussian(A=[[2.0, -1.0], [1.0, 0.0]], B=[1.0, 0.0], Q=[[1.0 / float(λ), 0.0], [0.0, 0.0]], R=[1.0],
                                      X0=[3 * y[0] - 2 * y[1], 2 * y[0] - y[1]], Σ0=float(init_cov) * np.eye(2))


class UCSV(StateSpaceModel):
    """struct UCSV (state_space_models.jl:215-259): state (x, log σε, log ση).
    `UCSV(γ, x0, (log_σε, log_ση))`; a scalar γ is used for both volatilities (example spelling,
    examples/inflation_example.jl:229-232)."""
    kind = _lib.UCSV
    state_dim = 3

    def __init__(self, γ, x0, log_σ0):
        if isinstance(γ, ParamColumn) or np.ndim(γ) == 0:
            g = (_field("γ", γ), _field("γ", γ))
        else:
            g = (_field("γε", γ[0]), _field("γη", γ[1]))
        self.γ, self.x0, self.log_σ0 = g, _field("x0", x0), (_field("log_σε", log_σ0[0]), _field("log_ση", log_σ0[1]))

    def params(self):
        return _stack_params([self.γ[0], self.γ[1], self.x0, self.log_σ0[0], self.log_σ0[1]])

    def __repr__(self):
        return f"UCSV(γ={self.γ}, x0={self.x0}, log_σ0={self.log_σ0})"


def unobserved_components_stochastic_volatility(*, x0, γε, γη, log_σε, log_ση):
    """state_space_models.jl:225-227"""
    return UCSV((γε, γη), x0, (log_σε, log_ση))


class StochasticVolatility(StateSpaceModel):
    """Canonical SV model (absent from the reference's src/; shown in
    visuals/stochastic_volatility_animation_1000.gif with panels μ, ρ, σ):
    x[1] ~ N(μ, σ²/(1-ρ²)), x[t] ~ N(μ + ρ (x[t-1]-μ), σ²), y[t] ~ N(0, exp(x[t]))."""
    kind = _lib.SV

    def __init__(self, μ, ρ, σ):
        self.μ, self.ρ, self.σ = _field("μ", μ), _field("ρ", ρ), _field("σ", σ)

    def params(self):
        return _stack_params([self.μ, self.ρ, self.σ])

    def __repr__(self):
        return f"StochasticVolatility(μ={self.μ}, ρ={self.ρ}, σ={self.σ})"


SV = StochasticVolatility


def params_of(model, θ):
    """[M, 8] parameter blocks of model(θ_m) for every row of θ [M, d]: one vectorised call with
    ParamColumns when the user's constructor allows it, else M scalar calls."""
    θ = np.asarray(θ, np.float64)
    try:
        m = model([ParamColumn(θ[:, k]) for k in range(θ.shape[1])])
        P = np.asarray(m.params(), np.float64)
        if P.ndim == 2 and P.shape[0] == θ.shape[0]:
            return _lib.params8(P), m
    except Exception:
        pass
    ms = [model(th) for th in θ]
    return np.stack([mm.params8() for mm in ms]), ms[0]


# ---- the model methods the reference exports (state_space_models.jl:1: transition, observation, initial_dist) ---------------
# Host-side descriptions of the densities the device functors of csrc/smcb_models.cuh evaluate: what a user of the
# reference gets from `transition(model, x)` etc.  Univariate states return a `priors.Normal` (μ, σ with σ a standard
# deviation, as Distributions.jl's Normal); UCSV returns the tuple of its three independent Normals (the reference's
# `TupleProduct`, state_space_models.jl:237,254); multivariate linear models return `MvNormal(μ, Σ)`.
class MvNormal:
    """mean vector and covariance matrix (Distributions.jl's MvNormal(μ, Σ)) of a multivariate linear model's densities"""

    def __init__(self, μ, Σ):
        self.μ, self.Σ = np.asarray(μ, np.float64), np.asarray(Σ, np.float64)

    def logpdf(self, x):
        d = np.asarray(x, np.float64) - self.μ
        _, logdet = np.linalg.slogdet(self.Σ)
        return float(-0.5 * (d @ np.linalg.solve(self.Σ, d) + logdet + self.μ.size * np.log(2 * np.pi)))


def _scalar_model(model):
    if isinstance(model.params(), np.ndarray):
        raise TypeError("transition / observation / initial_dist take ONE model, not a ParamColumn batch")


def initial_dist(model):
    """initial_dist(model)  (state_space_models.jl:105-109, 249-259; SV: the stationary law)"""
    from .priors import Normal
    if isinstance(model, MultivariateLinearModel):
        return MvNormal(model.x0, model.σ0)                                   # :181-185
    _scalar_model(model)
    if model.kind == _lib.LG1D:
        return Normal(model.x0, np.sqrt(model.σ0))
    if model.kind == _lib.SV:
        return Normal(model.μ, model.σ / np.sqrt(1.0 - model.ρ ** 2))
    return (Normal(model.x0, np.exp(0.5 * model.log_σ0[0])), Normal(model.log_σ0[0], model.γ[0]), Normal(model.log_σ0[1], model.γ[1]))


def transition(model, x):
    """transition(model, x)  (state_space_models.jl:87-94, 233-242): the law of x[t] given x[t-1] = x"""
    from .priors import Normal
    if isinstance(model, MultivariateLinearModel):
        return MvNormal(model.A @ np.asarray(x, np.float64), model.Q)          # :163-170
    _scalar_model(model)
    if model.kind == _lib.LG1D:
        return Normal(model.A * float(x), np.sqrt(model.Q))
    if model.kind == _lib.SV:
        return Normal(model.μ + model.ρ * (float(x) - model.μ), model.σ)
    x = np.asarray(x, np.float64)
    return (Normal(x[0], np.exp(0.5 * x[1])), Normal(x[1], model.γ[0]), Normal(x[2], model.γ[1]))   # previous log σε  :238


def observation(model, x):
    """observation(model, x)  (state_space_models.jl:96-103, 244-247): the law of y[t] given x[t] = x"""
    from .priors import Normal
    if isinstance(model, MultivariateLinearModel):
        return Normal(float((model.B @ np.asarray(x, np.float64))[0]), float(np.sqrt(model.R[0])))   # R a variance, as in the Kalman filter
    _scalar_model(model)
    if model.kind == _lib.LG1D:
        return Normal(model.B * float(x), np.sqrt(model.R))
    if model.kind == _lib.SV:
        return Normal(0.0, np.exp(0.5 * float(x)))
    x = np.asarray(x, np.float64)
    return Normal(x[0], np.exp(0.5 * x[2]))                                    # current log ση  :246


def simulate(*args, seed=1998):
    """simulate([rng,] model, T) -> (x, y)  (state_space_models.jl:11-28).  The rng argument of the
    reference becomes a Philox seed: an int, or a numpy Generator from which one is drawn.
    x has shape [T] (or [T, 3] for UCSV, one row per period)."""
    if len(args) == 3:                                   # simulate(rng, model, T)  :11-26
        rng, model, T = args
        seed = int(rng.integers(0, 2 ** 63)) if hasattr(rng, "integers") else int(rng)
    elif len(args) == 2:                                 # simulate(model, T) = simulate(Random.GLOBAL_RNG, model, T)  :28
        model, T = args
    else:
        raise TypeError("simulate([rng,] model, T)")
    x, y = _lib.simulate(model.kind, model.params(), int(T), int(seed))
    return (x[0] if model.state_dim == 1 else np.ascontiguousarray(x.T)), y


def preallocate(model, N):
    """preallocate(model, N) (state_space_models.jl:80-85, 229-231)"""
    return np.zeros(N) if model.state_dim == 1 else np.zeros((N, model.state_dim))
