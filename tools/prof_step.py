"""Short single-filter run for ncu: LG1D N=2^24, T=5, systematic (3 kernels per step)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import sequential_monte_carlo_b200 as smc
ctx = smc.Context(0, 1998)
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
y = smc._lib.simulate(smc.KIND_LG1D, [0.5, 1.0, 0.9, 0.8, 0.0, 1.0], 5, 1998)[1]
z = ctx.log_likelihood(smc.KIND_LG1D, [0.5, 1.0, 0.9, 0.8, 0.0, 1.0], 1 << logn, y, smc.SYSTEMATIC)
print("logZ", z, ctx.timing())
