"""Dynamic (θ, chunk) scheduling of the batched engine (SMCB_BATCH_CHUNK = 0 static / K steps per chunk / unset = automatic):
device time of one whole-series sweep and bit-identity of logZ and of the final clouds against the static launch.
python tools/chunk_probe.py kind N T M1,M2,... [chunks, default 0,4,8,16,32,auto]"""
import hashlib, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 6:   # child: one chunk setting, one json line per M
    import numpy as np
    import sequential_monte_carlo_b200 as smc
    kind, N, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    Ms = [int(v) for v in sys.argv[4].split(",")]
    true = {0: [0.5, 1.0, 0.9, 0.8, 0.0, 1.0], 1: [-1.0, 0.9, 0.3], 2: [0.2, 0.2, 3.0, 1.0, 1.0]}[kind]
    y = smc._lib.simulate(kind, true, T, 1998)[1]
    ctx = smc.Context(0, 1998)
    for M in Ms:
        P = np.tile(smc._lib.params8(true), (M, 1))
        P[:, 0] *= np.linspace(0.9, 1.1, M)
        b = ctx.batch(kind, M, N)
        row = {"kind": kind, "N": N, "T": T, "M": M, "chunk": os.environ.get("SMCB_BATCH_CHUNK", "auto")}
        for rs, name in ((smc.SYSTEMATIC, "systematic"), (smc.MULTINOMIAL, "multinomial")):
            z = b.log_likelihood(P, y, rs, 0)
            ms = []
            for _ in range(3):
                z = b.log_likelihood(P, y, rs, 0)
                ms.append(b.timing()[0])
            x, _, lw = b.fetch(want_w=False, want_logw=True)
            h = hashlib.sha256(np.ascontiguousarray(z).tobytes() + x.tobytes() + lw.tobytes()).hexdigest()[:16]
            row[name + "_ms"] = min(ms)
            row[name + "_sha"] = h
        b.close()
        print(json.dumps(row), flush=True)
else:
    chunks = sys.argv[5].split(",") if len(sys.argv) > 5 else ["0", "4", "8", "16", "32", "auto"]
    for c in chunks:
        env = dict(os.environ)
        env.pop("SMCB_BATCH_CHUNK", None)
        if c != "auto":
            env["SMCB_BATCH_CHUNK"] = c
        subprocess.run([sys.executable, __file__] + sys.argv[1:5] + ["x", "child"], env=env)
