"""One batched bootstrap sweep (M θ × N particles × T) timed on the device; SMCB_BATCH_X_SMEM=0/1 picks where the clouds live.
python tools/batch_xsmem_probe.py kind M N T"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sequential_monte_carlo_b200 as smc

kind, M, N, T = (int(v) for v in sys.argv[1:5])
true = {0: [0.5, 1.0, 0.9, 0.8, 0.0, 1.0], 1: [-1.0, 0.9, 0.3], 2: [0.2, 0.2, 3.0, 1.0, 1.0]}[kind]
y = smc._lib.simulate(kind, true, T, 1998)[1]
ctx = smc.Context(0, 1998)
P = np.tile(smc._lib.params8(true), (M, 1))
b = ctx.batch(kind, M, N)
out = {"kind": kind, "M": M, "N": N, "T": T, "x_smem_env": os.environ.get("SMCB_BATCH_X_SMEM")}
for rs, name in ((smc.SYSTEMATIC, "systematic"), (smc.MULTINOMIAL, "multinomial")):
    ctx.set_rng(1, 1)
    b.log_likelihood(P, y, rs, 0)
    ctx.set_rng(1, 1)
    z = b.log_likelihood(P, y, rs, 0)
    out[name] = {"ms": b.timing()[0], "logZ_sum": float(z.sum())}
print(json.dumps(out))
