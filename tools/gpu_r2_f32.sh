#!/bin/bash
set -x
python -m pytest tests/test_gpu_filter.py -m gpu -x -q -k "f32 or resample_with or stale_weights" 2>&1 | tail -15
python -m pytest tests/test_theta_level.py -m gpu -x -q -k "docstring" 2>&1 | tail -15
python - <<'PY'
import json, sequential_monte_carlo_b200 as smc
ctx = smc.Context(0, 1998)
P = [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]
y = smc._lib.simulate(0, P, 60, 1998)[1]
for prec in ("f64", "f32", "f32_arith"):
    ctx.set_precision(prec)
    for logn in (22, 24):
        N = 1 << logn
        for rs, name in ((smc.SYSTEMATIC, "systematic"), (smc.MULTINOMIAL, "multinomial")):
            z = ctx.log_likelihood(0, P, N, y, rs)
            ctx.set_profiling(True)
            ctx.log_likelihood(0, P, N, y, rs)
            ms, n = ctx.timing()
            ctx.set_profiling(False)
            ctx.log_likelihood(0, P, N, y, rs)
            tot = ctx.timing()[0]["total"]
            print(json.dumps({"prec": prec, "logn": logn, "resampler": name, "us_per_step": round(1e3 * tot / 60, 1), "Gpups": round(N * 60 / tot / 1e6, 1), "logZ": z,
                              "per_launch_us": {k: round(1e3 * ms[k] / max(n[k], 1), 1) for k in ("scan", "bounds", "anc", "prop")}}), flush=True)
PY
