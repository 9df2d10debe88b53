"""sequential_monte_carlo_b200 — B200-native particle-filter hot path behind the API of
SequentialMonteCarlo.jl (bootstrap_filter / bootstrap_filter! / log_likelihood / SMC / smc² /
density_tempered).  Python host mirror over the C ABI of include/smcb200.h; all compute runs in
hand-written sm_100a CUDA (csrc/).  No CPU fallback.
"""
from . import _lib
from ._lib import LG1D, SV, UCSV, MULTINOMIAL, STRATIFIED, SYSTEMATIC, Context, Batch, SMCBError  # noqa: F401

__version__ = "0.1.0"
