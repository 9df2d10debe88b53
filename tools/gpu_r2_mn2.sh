#!/bin/bash
# two-level multinomial v2 (bucket counting sort): parity tests, multinomial timing, batch occupancy probe
set -x
python -m pytest tests/test_gpu_filter.py -m gpu -x -q -k "two_level or ragged or medium or mixed or after_sorted or f32" 2>&1 | tail -8
python - <<'PY'
import json, numpy as np, sequential_monte_carlo_b200 as smc
ctx = smc.Context(0, 1998)
P = [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]
y = smc._lib.simulate(0, P, 60, 1998)[1]
for logn in (20, 22, 24):
    N = 1 << logn
    for rs, name in ((smc.SYSTEMATIC, "systematic"), (smc.MULTINOMIAL, "multinomial")):
        ctx.log_likelihood(0, P, N, y, rs)
        ctx.set_profiling(True)
        ctx.log_likelihood(0, P, N, y, rs)
        ms, n = ctx.timing()
        ctx.set_profiling(False)
        ctx.log_likelihood(0, P, N, y, rs)
        tot = ctx.timing()[0]["total"]
        print(json.dumps({"logn": logn, "resampler": name, "us_per_step": 1e3 * tot / 60, "Gpups": N * 60 / tot / 1e6,
                          "per_launch_us": {k: round(1e3 * ms[k] / max(n[k], 1), 1) for k in ("scan", "bounds", "anc", "prop")}}), flush=True)
PY
python tools/batch_occupancy_probe.py 0 1024 100 32,64,128,256,512,1024 2>&1 | tee gpurun_out/r2_batch_occupancy_lg1024.jsonl
python tools/batch_occupancy_probe.py 2 4096 40 64,148,296,512,592 2>&1 | tee gpurun_out/r2_batch_occupancy_ucsv4096.jsonl
