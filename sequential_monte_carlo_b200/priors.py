"""Priors over θ — the slice of Distributions.jl the reference's samplers use
(README.md:81-85, examples/inflation_example.jl:234-239, src/smc_samplers.jl:38,116,123-126):
`rand`, `logpdf`, `insupport` of a product of univariate distributions.  Host-side, M×d numbers.

Draws come from the Philox host stream (docs/SPEC.md §8): component k uses stream k, attempt a uses
time index a, purpose PRIOR; particle m reads element m of the vector.
"""
import math

import numpy as np

from . import _lib

_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def _phi(z):
    return 0.5 * (1.0 + math.erf(z / math.sqrt(2.0)))


class Distribution:
    def insupport(self, x):
        return bool(np.isfinite(x))

    # vectorised forms over M values (used by the samplers; same formulas with numpy ufuncs)
    def insupport_v(self, x):
        return np.isfinite(x)

    def logpdf_v(self, x):
        return np.array([self.logpdf(float(v)) for v in x])

    def sample(self, M, seed, k):
        raise NotImplementedError


class Normal(Distribution):
    def __init__(self, μ=0.0, σ=1.0):
        self.μ, self.σ = float(μ), float(σ)

    def logpdf(self, x):
        z = (x - self.μ) / self.σ
        return -0.5 * z * z - math.log(self.σ) - _HALF_LOG_2PI

    def logpdf_v(self, x):
        z = (x - self.μ) / self.σ
        return -0.5 * z * z - math.log(self.σ) - _HALF_LOG_2PI

    def sample(self, M, seed, k):
        return self.μ + self.σ * _lib.rng_normals(seed, 0, k, 0, _lib.P_PRIOR, 0, M)


class LogNormal(Distribution):
    def __init__(self, μ=0.0, σ=1.0):
        self.μ, self.σ = float(μ), float(σ)

    def insupport(self, x):
        return bool(np.isfinite(x) and x > 0.0)

    def logpdf(self, x):
        if not self.insupport(x):
            return -math.inf
        lx = math.log(x)
        z = (lx - self.μ) / self.σ
        return -lx - math.log(self.σ) - _HALF_LOG_2PI - 0.5 * z * z

    def insupport_v(self, x):
        return np.isfinite(x) & (x > 0.0)

    def logpdf_v(self, x):
        ok = self.insupport_v(x)
        lx = np.log(np.where(ok, x, 1.0))
        z = (lx - self.μ) / self.σ
        return np.where(ok, -lx - math.log(self.σ) - _HALF_LOG_2PI - 0.5 * z * z, -math.inf)

    def sample(self, M, seed, k):
        return np.exp(self.μ + self.σ * _lib.rng_normals(seed, 0, k, 0, _lib.P_PRIOR, 0, M))


class Uniform(Distribution):
    def __init__(self, a=0.0, b=1.0):
        self.a, self.b = float(a), float(b)

    def insupport(self, x):
        return bool(self.a <= x <= self.b)

    def logpdf(self, x):
        return -math.log(self.b - self.a) if self.insupport(x) else -math.inf

    def insupport_v(self, x):
        return (x >= self.a) & (x <= self.b)

    def logpdf_v(self, x):
        return np.where(self.insupport_v(x), -math.log(self.b - self.a), -math.inf)

    def sample(self, M, seed, k):
        return self.a + (self.b - self.a) * _lib.rng_uniforms01(seed, 0, k, 0, _lib.P_PRIOR, M)


class TruncatedNormal(Distribution):
    """TruncatedNormal(μ, σ, lower, upper) (README.md:82)"""

    def __init__(self, μ, σ, lower, upper):
        self.μ, self.σ, self.lo, self.hi = float(μ), float(σ), float(lower), float(upper)
        self._logmass = math.log(_phi((self.hi - self.μ) / self.σ) - _phi((self.lo - self.μ) / self.σ))

    def insupport(self, x):
        return bool(self.lo <= x <= self.hi)

    def logpdf(self, x):
        if not self.insupport(x):
            return -math.inf
        z = (x - self.μ) / self.σ
        return -0.5 * z * z - math.log(self.σ) - _HALF_LOG_2PI - self._logmass

    def insupport_v(self, x):
        return (x >= self.lo) & (x <= self.hi)

    def logpdf_v(self, x):
        z = (x - self.μ) / self.σ
        return np.where(self.insupport_v(x), -0.5 * z * z - math.log(self.σ) - _HALF_LOG_2PI - self._logmass, -math.inf)

    def sample(self, M, seed, k):
        out = np.empty(M)
        todo = np.ones(M, bool)
        attempt = 0
        while todo.any():  # rejection from the parent normal; attempt a reads time index a
            x = self.μ + self.σ * _lib.rng_normals(seed, 0, k, attempt, _lib.P_PRIOR, 0, M)
            ok = todo & (x >= self.lo) & (x <= self.hi)
            out[ok] = x[ok]
            todo &= ~ok
            attempt += 1
            if attempt > 10000:
                raise RuntimeError("TruncatedNormal: support has negligible mass")
        return out


class Product(Distribution):
    """product_distribution([...])"""

    def __init__(self, components):
        self.components = list(components)

    def __len__(self):
        return len(self.components)

    def insupport(self, θ):
        return all(c.insupport(float(v)) for c, v in zip(self.components, θ))

    def logpdf(self, θ):
        s = 0.0
        for c, v in zip(self.components, θ):
            s += c.logpdf(float(v))
        return s

    def insupport_v(self, θ):
        """[M] bool for θ [M, d]"""
        ok = np.ones(θ.shape[0], bool)
        for k, c in enumerate(self.components):
            ok &= c.insupport_v(θ[:, k])
        return ok

    def logpdf_v(self, θ):
        s = np.zeros(θ.shape[0])
        for k, c in enumerate(self.components):
            s = s + c.logpdf_v(θ[:, k])
        return s

    def sample(self, M, seed):
        """[M, d] draws: θ = map(m -> rand(prior), 1:M)  (smc_samplers.jl:38)"""
        return np.stack([c.sample(M, seed, k) for k, c in enumerate(self.components)], axis=1)


def product_distribution(components):
    return Product(components)
