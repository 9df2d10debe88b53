"""Scratch timing of the single-filter and batched paths (not the contract bench; see bench.py)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import sequential_monte_carlo_b200 as smc

LG = [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]
ctx = smc.Context(0, 1998)
rng = np.random.default_rng(0)
out = []
for logn in ([20, 22, 24] if len(sys.argv) < 2 else [int(a) for a in sys.argv[1:]]):
    N, T = 1 << logn, 60
    y = rng.normal(size=T)
    for rs, name in ((smc.SYSTEMATIC, "systematic"), (smc.STRATIFIED, "stratified"), (smc.MULTINOMIAL, "multinomial")):
        if rs == smc.MULTINOMIAL and logn > 22:
            TT = 12
        else:
            TT = T
        ctx.log_likelihood(smc.KIND_LG1D, LG, N, y[:5], rs)  # warm-up
        ctx.set_profiling(False)
        z = ctx.log_likelihood(smc.KIND_LG1D, LG, N, y[:TT], rs)
        ms, n = ctx.timing()
        ctx.set_profiling(True)
        ctx.log_likelihood(smc.KIND_LG1D, LG, N, y[:TT], rs)
        msp, npf = ctx.timing()
        ctx.set_profiling(False)
        pups = N * TT / (ms["total"] * 1e-3)
        rec = dict(N=N, T=TT, resampler=name, total_ms=ms["total"], gpups=pups / 1e9, frac56=pups * 56 / 6551.4e9,
                   scan_us=1e3 * msp["scan"] / max(npf["scan"], 1), bounds_us=1e3 * msp["bounds"] / max(npf["bounds"], 1),
                   anc_us=1e3 * msp["anc"] / max(npf["anc"], 1), prop_us=1e3 * msp["prop"] / max(npf["prop"], 1),
                   init_us=1e3 * msp["init"] / max(npf["init"], 1), logZ=z)
        print(json.dumps(rec), flush=True)
        out.append(rec)
# the binary32-state tier (docs/SPEC.md §9), systematic only; roofline figure at SURVEY §8d's 40 B per particle-update
ctx.set_precision("f32")
for logn in ([20, 22, 24] if len(sys.argv) < 2 else [int(a) for a in sys.argv[1:]]):
    N, T = 1 << logn, 60
    y = rng.normal(size=T)
    ctx.log_likelihood(smc.KIND_LG1D, LG, N, y[:5], smc.SYSTEMATIC)
    z = ctx.log_likelihood(smc.KIND_LG1D, LG, N, y, smc.SYSTEMATIC)
    ms, n = ctx.timing()
    ctx.set_profiling(True)
    ctx.log_likelihood(smc.KIND_LG1D, LG, N, y, smc.SYSTEMATIC)
    msp, npf = ctx.timing()
    ctx.set_profiling(False)
    pups = N * T / (ms["total"] * 1e-3)
    print(json.dumps(dict(N=N, T=T, precision="f32 states", resampler="systematic", total_ms=ms["total"], gpups=pups / 1e9,
                          frac40=pups * 40 / 6551.4e9, **{k + "_us": 1e3 * msp[k] / max(npf[k], 1) for k in ("scan", "bounds", "anc", "prop", "init")},
                          logZ=z)), flush=True)
ctx.set_precision("f64")
# batched: config 3 shape (512 x 1024, T=100), config 4 (1024 x 2048 SV, T=500), config 5 per-GPU (512 x 4096 UCSV, T=241)
for kind, M, N, T, P in ((smc.KIND_LG1D, 512, 1024, 100, LG), (smc.KIND_SV, 1024, 2048, 500, [-1.0, 0.9, 0.3]),
                         (smc.KIND_UCSV, 512, 4096, 241, [0.2, 0.2, 3.0, 1.0, 1.0]), (smc.KIND_LG1D, 4096, 1024, 100, LG)):
    b = ctx.batch(kind, M, N)
    Pm = np.tile(smc._lib.params8(P), (M, 1))
    y = rng.normal(size=T)
    for rs, name in ((smc.SYSTEMATIC, "systematic"), (smc.MULTINOMIAL, "multinomial")):
        b.log_likelihood(Pm, y[:10], rs)
        z = b.log_likelihood(Pm, y, rs)
        ms, _ = b.timing()
        rec = dict(batch=True, kind=kind, M=M, N=N, T=T, resampler=name, ms=ms, gpups=M * N * T / ms / 1e6, logZ_mean=float(z.mean()))
        print(json.dumps(rec), flush=True)
    t0 = time.perf_counter()
    lm, es = b.step(y[0], smc.MULTINOMIAL)
    ms, _ = b.timing()
    print(json.dumps(dict(batch_step=True, kind=kind, M=M, N=N, ms=ms, wall_ms=1e3 * (time.perf_counter() - t0))), flush=True)
    b.close()
