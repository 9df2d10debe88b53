"""per-step time and per-launch split of the large-N single filter under the three resamplers (multinomial = SPEC §5c)
python tools/mn_time.py [logn ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequential_monte_carlo_b200 as smc
ctx = smc.Context(0, 1998)
P = [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]
T = 60
y = smc._lib.simulate(0, P, T, 1998)[1]
for logn in [int(v) for v in sys.argv[1:]] or [22, 24]:
    N = 1 << logn
    for rs, name in ((smc.SYSTEMATIC, "systematic"), (smc.STRATIFIED, "stratified"), (smc.MULTINOMIAL, "multinomial")):
        z = ctx.log_likelihood(0, P, N, y, rs)
        ctx.set_profiling(True)
        ctx.log_likelihood(0, P, N, y, rs)
        ms, n = ctx.timing()
        ctx.set_profiling(False)
        ctx.log_likelihood(0, P, N, y, rs)
        tot = ctx.timing()[0]["total"]
        print(json.dumps({"logn": logn, "resampler": name, "us_per_step": round(1e3 * tot / T, 1), "Gpups": round(N * T / tot / 1e6, 1),
                          "frac56": round(56 * N * T / (tot * 1e-3) / 6551.4e9, 3), "logZ": z,
                          "per_launch_us": {k: round(1e3 * ms[k] / max(n[k], 1), 1) for k in ("scan", "bounds", "anc", "prop")}}), flush=True)
