"""Kalman filter for univariate linear Gaussian models — mirror of /root/reference/src/kalman_filter.jl:29-70,
evaluated on the device for M models at once (the inner filter of IBIS, ibis.jl:95-105,172-177)."""
import numpy as np

from . import _lib
from .particles import default_context


def kalman_filter(model, x, Σ, y, *, ctx=None):
    """x, Σ, loglik = kalman_filter(model, x, Σ, y)  — one predict + update (kalman_filter.jl:29-53)"""
    ctx = ctx or default_context()
    xs, ss, ll = ctx.kalman_step(model.params(), x, Σ, float(y))
    return float(xs[0]), float(ss[0]), float(ll[0])


def log_likelihood(y, model, *, matched_init=False, ctx=None):
    """log_likelihood(y, model::LinearModel)  (kalman_filter.jl:55-70).  The reference predicts before
    the first update, i.e. treats (x0, σ0) as the law of x[0]; the particle filter draws x[1] from it
    (SURVEY.md D1).  matched_init=True gives the likelihood the particle filter targets."""
    ctx = ctx or default_context()
    ll, _, _ = ctx.kalman_loglik(model.params(), np.asarray(y, np.float64), matched_init)
    return float(ll[0])
