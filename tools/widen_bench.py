"""The "widen" leg of bench.py on its own (guided vs bootstrap sweeps, guided single filter, matrix Kalman): prints its JSON."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import sequential_monte_carlo_b200 as smc

ctx = smc.Context(0, 1998)
print(json.dumps(bench.widen_leg(ctx)))
