"""sequential_monte_carlo_b200 — B200-native particle-filter hot path behind the API of
SequentialMonteCarlo.jl (bootstrap_filter / bootstrap_filter! / log_likelihood / SMC / smc² /
density_tempered).  Python host mirror over the C ABI of include/smcb200.h; all compute runs in
hand-written sm_100a CUDA (csrc/).  No CPU fallback.
"""
from . import _lib
from ._lib import MULTINOMIAL, STRATIFIED, SYSTEMATIC, Context, Batch, SMCBError  # noqa: F401
from ._lib import LG1D as KIND_LG1D, SV as KIND_SV, UCSV as KIND_UCSV  # noqa: F401

__version__ = "0.1.0"

from .state_space_models import (StateSpaceModel, LinearModel, UnivariateLinearGaussian, LinearGaussian,  # noqa: E402,F401
                                 unobserved_components, UC, UCSV, MultivariateLinearModel, MultivariateLinearGaussian, hodrick_prescott, unobserved_components_stochastic_volatility,
                                 StochasticVolatility, SV, simulate, preallocate, transition, observation, initial_dist, MvNormal)
from .particles import (normalize, reweight, resample, bootstrap_filter, bootstrap_filter_, particle_filter, particle_filter_, log_likelihood,
                        AffineGaussianProposal, UCSVTrendProposal, locally_optimal_proposal, guided_log_likelihood,  # noqa: E402,F401
                        quantile, weighted_mean_var, default_context, set_default_context)
from .priors import Normal, LogNormal, Uniform, TruncatedNormal, product_distribution  # noqa: E402,F401
from .smc_samplers import (SMC, smc2, smc2_step, density_tempered, expected_parameters, random_walk_kernel,  # noqa: E402,F401
                           lg_optimal_proposals, estimated_trend, state_means, state_variances, state_quantiles, get_quantiles, LocalComm, TorchComm, NcclComm)
from .ibis import IBIS  # noqa: E402,F401
from . import kalman_filter, ibis, smc_samplers, particles, state_space_models, priors  # noqa: E402,F401
