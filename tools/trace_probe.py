"""Stage structure of density_tempered at the docstring trace's shape (smc_samplers.jl:198-219) for several series lengths T
(the docstring does not record T).  python tools/trace_probe.py T1,T2,... [seeds]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sequential_monte_carlo_b200 as smc

def lg_mod(th):
    return smc.StateSpaceModel(smc.LinearGaussian(th[0], 1.0, th[1], th[2], 0.0), (1, 1))

pg = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.LogNormal()])
ctx = smc.Context(0, 1998)
Ts = [int(v) for v in sys.argv[1].split(",")]
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
for T in Ts:
    for seed in range(seeds):
        y = smc.simulate(lg_mod([0.5, 0.9, 0.8]), T, seed=1000 + seed)[1]
        g = smc.SMC(1024, 512, lg_mod, pg, 3, 0.5, seed=seed, ctx=ctx, engine="device")
        g._engine_data(y)
        st = g._eng.density_tempered()
        g._stale = True
        m = smc.expected_parameters(g).ravel()
        print(json.dumps({"T": T, "seed": seed, "stages": len(st), "xi": [round(s[0], 5) for s in st], "ess": [round(s[1], 3) for s in st],
                          "acc": [round(s[2], 4) for s in st], "mean": [round(float(v), 4) for v in m]}), flush=True)
        g.close()
