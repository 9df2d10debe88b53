"""Per-call cost of the stepping API (bootstrap_filter + bootstrap_filter! one observation at a time, README usage)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import sequential_monte_carlo_b200 as smc
ctx = smc.Context(0, 1998)
LG = [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]
for logn in (20, 24):
    N = 1 << logn
    y = smc._lib.simulate(smc.KIND_LG1D, LG, 40, 1998)[1]
    ctx.bootstrap_init(smc.KIND_LG1D, LG, N, float(y[0]), 0)
    for t in range(1, 8):
        ctx.bootstrap_step(float(y[t]), smc.SYSTEMATIC, LG)
    ctx.synchronize()
    t0 = time.perf_counter()
    dev = 0.0
    for t in range(8, 40):
        ctx.bootstrap_step(float(y[t]), smc.SYSTEMATIC, LG)
        dev += ctx.timing()[0]["total"]
    wall = (time.perf_counter() - t0) / 32
    print(f"N=2^{logn}: bootstrap_filter! {1e6 * wall:.1f} us wall per call, {1e3 * dev / 32:.1f} us device", flush=True)
    # on-device summaries after a step (docs/SPEC.md §8) against reading the cloud back
    t0 = time.perf_counter()
    for _ in range(5):
        m, v, q = ctx.summary([0.25, 0.5, 0.75])
    t_dev = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    for _ in range(3):
        x, w, _ = ctx.fetch_state()
    t_host = (time.perf_counter() - t0) / 3
    print(f"N=2^{logn}: weighted mean/var/3 quantiles on the device {1e3 * t_dev:.2f} ms; reading x, w back (pinned) {1e3 * t_host:.2f} ms "
          f"(+ the host-side sort)", flush=True)
