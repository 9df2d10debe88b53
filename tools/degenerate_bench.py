"""Per-kernel time of the sorted-resampler step at N = 2^24 as the likelihood sharpens (observation
variance R from 0.8 down to 1e-10): the CDF windows of the ancestor CTAs go from ~2048 entries each to a
few huge, mostly dead ones.  Scratch timing (profiling mode, per-launch CUDA events)."""
import sys, json
import numpy as np
sys.path.insert(0, ".")
import sequential_monte_carlo_b200 as smc

ctx = smc.Context(0, 1998)
rng = np.random.default_rng(0)
N, T = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24), 12
y = rng.normal(size=T)
for R in (0.8, 1e-2, 1e-4, 1e-6, 1e-8, 1e-10):
    P = [0.5, 1.0, 0.9, R, 0.0, 1.0]
    for rs, name in ((smc.SYSTEMATIC, "systematic"), (smc.STRATIFIED, "stratified"), (smc.MULTINOMIAL, "multinomial")):
        ctx.log_likelihood(smc.KIND_LG1D, P, N, y[:3], rs)
        ctx.set_profiling(True)
        _, _, ess = ctx.log_likelihood(smc.KIND_LG1D, P, N, y, rs, per_step=True)
        msp, npf = ctx.timing()
        ctx.set_profiling(False)
        print(json.dumps(dict(R=R, resampler=name, ess_min=float(ess.min()), ess_med=float(np.median(ess)),
                              **{k + "_us": round(1e3 * msp[k] / max(npf[k], 1), 2) for k in ("scan", "bounds", "anc", "prop")})), flush=True)
