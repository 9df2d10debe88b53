"""How the batched engine behaves when a GPU holds FEW θ-particles (the regime of θ-sharding over 8 GPUs):
device time of one whole-series sweep for M in a list, pairs per thread 1 / 2 / 4 (SMCB_BATCH_PAIRS).
python tools/batch_occupancy_probe.py kind N T M1,M2,..."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 5:   # child: one (pairs) setting, prints a json line per M
    import numpy as np
    import sequential_monte_carlo_b200 as smc
    kind, N, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    Ms = [int(v) for v in sys.argv[4].split(",")]
    true = {0: [0.5, 1.0, 0.9, 0.8, 0.0, 1.0], 1: [-1.0, 0.9, 0.3], 2: [0.2, 0.2, 3.0, 1.0, 1.0]}[kind]
    y = smc._lib.simulate(kind, true, T, 1998)[1]
    ctx = smc.Context(0, 1998)
    for M in Ms:
        P = np.tile(smc._lib.params8(true), (M, 1))
        b = ctx.batch(kind, M, N)
        row = {"kind": kind, "N": N, "T": T, "M": M, "pairs": os.environ.get("SMCB_BATCH_PAIRS", "auto"), "cluster": os.environ.get("SMCB_BATCH_CLUSTER", "auto")}
        for rs, name in ((smc.SYSTEMATIC, "systematic"), (smc.MULTINOMIAL, "multinomial")):
            b.log_likelihood(P, y, rs, 0)
            ms = []
            for _ in range(3):
                b.log_likelihood(P, y, rs, 0)
                ms.append(b.timing()[0])
            row[name + "_ms"] = min(ms)
            row[name + "_us_per_step"] = 1e3 * min(ms) / T
        b.close()
        print(json.dumps(row), flush=True)
else:
    for pairs in ("1", "2", "4"):
        env = dict(os.environ, SMCB_BATCH_PAIRS=pairs)
        subprocess.run([sys.executable, __file__] + sys.argv[1:5] + ["child"], env=env)
