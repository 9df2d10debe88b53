"""Kalman filter for linear Gaussian models — mirror of /root/reference/src/kalman_filter.jl (scalar methods :29-53,
matrix methods :3-27, log_likelihood :55-70), evaluated on the device for M models at once (the inner filter of IBIS,
ibis.jl:95-105,172-177)."""
import numpy as np

from . import _lib
from .particles import default_context
from .state_space_models import MultivariateLinearModel


def kalman_filter(model, x, Σ, y, *, ctx=None):
    """x, Σ, loglik = kalman_filter(model, x, Σ, y)  — one predict + update (kalman_filter.jl:29-53; vector state :3-27)"""
    ctx = ctx or default_context()
    if isinstance(model, MultivariateLinearModel):
        d = model.state_dim
        xs, ss, ll = ctx.kalman_mv_step(d, model.block(), np.asarray(x, np.float64).reshape(1, d),
                                        np.asarray(Σ, np.float64).reshape(1, d, d), float(y))
        return xs[0], ss[0], float(ll[0])
    xs, ss, ll = ctx.kalman_step(model.params(), x, Σ, float(y))
    return float(xs[0]), float(ss[0]), float(ll[0])


def log_likelihood(y, model, *, matched_init=False, ctx=None):
    """log_likelihood(y, model::LinearModel)  (kalman_filter.jl:55-70).  The reference predicts before
    the first update, i.e. treats (x0, σ0) as the law of x[0]; the particle filter draws x[1] from it
    (SURVEY.md D1).  matched_init=True gives the likelihood the particle filter targets."""
    ctx = ctx or default_context()
    if isinstance(model, MultivariateLinearModel):
        ll, _, _ = ctx.kalman_mv_loglik(model.state_dim, model.block(), np.asarray(y, np.float64), matched_init)
        return float(ll[0])
    ll, _, _ = ctx.kalman_loglik(model.params(), np.asarray(y, np.float64), matched_init)
    return float(ll[0])


def filtered_moments(y, model, *, matched_init=False, ctx=None):
    """(x_T, Σ_T, logZ) — what the reference's log_likelihood(y, model) returns in full (kalman_filter.jl:69)"""
    ctx = ctx or default_context()
    if isinstance(model, MultivariateLinearModel):
        ll, x, s = ctx.kalman_mv_loglik(model.state_dim, model.block(), np.asarray(y, np.float64), matched_init)
        return x[0], s[0], float(ll[0])
    ll, x, s = ctx.kalman_loglik(model.params(), np.asarray(y, np.float64), matched_init)
    return float(x[0]), float(s[0]), float(ll[0])
