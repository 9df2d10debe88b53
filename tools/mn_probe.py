"""one short multinomial sweep of the single filter (two-level draw, docs/SPEC.md §5c) for ncu: python tools/mn_probe.py logn T"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequential_monte_carlo_b200 as smc
logn, T = int(sys.argv[1]), int(sys.argv[2])
ctx = smc.Context(0, 1998)
P = [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]
y = smc._lib.simulate(0, P, T, 1998)[1]
z = ctx.log_likelihood(0, P, 1 << logn, y, smc.MULTINOMIAL)
print("logZ", z, "ms", ctx.timing()[0]["total"])
