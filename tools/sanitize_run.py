"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
single filter (three resamplers, ragged sizes, degenerate weights, two-level multinomial, binary32 tiers, multivariate LG, guided
moves), batched engine (static and dynamically scheduled, guided LG1D / UCSV), device-resident sampler (smc², density-tempered)."""
import os
import sys
sys.path.insert(0, ".")
import numpy as np
import sequential_monte_carlo_b200 as smc
ctx = smc.Context(0, 1998)
LG = [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]
UC = [0.2, 0.2, 3.0, 1.0, 1.0]
y = smc._lib.simulate(smc.KIND_LG1D, LG, 6, 1998)[1]
for N in (1, 31, 5000, 70001, (1 << 16) + 3):
    for rs in (smc.SYSTEMATIC, smc.STRATIFIED, smc.MULTINOMIAL):
        ctx.log_likelihood(smc.KIND_LG1D, LG, N, y, rs)
# degenerate weights: wide-window path of anc_hist_kernel, heavy cells of the two-level multinomial draw
ctx.log_likelihood(smc.KIND_LG1D, [0.5, 1.0, 0.9, 1e-9, 0.0, 1.0], 1 << 16, y, smc.SYSTEMATIC)
ctx.log_likelihood(smc.KIND_LG1D, [0.5, 1.0, 0.9, 1e-9, 0.0, 1.0], 1 << 16, y, smc.MULTINOMIAL)
ctx.log_likelihood(smc.KIND_UCSV, UC, 4099, y, smc.SYSTEMATIC)
ctx.log_likelihood(smc.KIND_SV, [-1.0, 0.9, 0.3], 2049, y, smc.STRATIFIED)
for prec in ("f32", "f32_arith", "f64"):
    ctx.set_precision(prec)
    ctx.log_likelihood(smc.KIND_LG1D, LG, 20011, y, smc.SYSTEMATIC)
    ctx.log_likelihood(smc.KIND_UCSV, UC, 9001, y, smc.MULTINOMIAL)
blk = smc.MultivariateLinearGaussian(A=[[0.7, 0.2], [-0.1, 0.5]], B=[1.0, 0.5], Q=[[0.5, 0.1], [0.1, 0.3]], R=[0.8]).block()
ctx.log_likelihood(smc._lib.MVLG2, blk, 9001, y, smc.SYSTEMATIC)
prop = np.array([smc.locally_optimal_proposal(smc.LinearGaussian(*LG[:5]), yt) for yt in y])
ctx.guided_log_likelihood(smc.KIND_LG1D, LG, 20011, y, prop, smc.SYSTEMATIC)
ctx.guided_log_likelihood(smc.KIND_UCSV, UC, 20011, y, np.tile([0.7, 0.0, 1.0], (y.size, 1)), smc.STRATIFIED)
x, w, lw = ctx.fetch_state(want_logw=True)
ctx.summary((0.25, 0.5, 0.75))
P = np.tile(smc._lib.params8(UC), (8, 1))
b = ctx.batch(smc.KIND_UCSV, 8, 1024)
b.log_likelihood(P, y, smc.SYSTEMATIC)
b.log_likelihood(P, y, smc.MULTINOMIAL)
b.log_likelihood(P, y, smc.SYSTEMATIC, proposal=np.tile([1.0, 0.0, 1.0], (y.size, 8, 1)))
b.weighted_moments()
b.weighted_quantiles((0.1, 0.9))
b.close()
# the dynamically scheduled kernel: more θ than resident CTAs would need M > 148; forced chunks exercise the same code on a few θ
y2 = smc._lib.simulate(smc.KIND_LG1D, LG, 13, 1998)[1]
for chunk in ("1", "4"):
    os.environ["SMCB_BATCH_CHUNK"] = chunk
    for kind, par, N in ((smc.KIND_LG1D, LG, 300), (smc.KIND_UCSV, UC, 640), (smc.KIND_UCSV, UC, 4096)):
        M = 5
        b = ctx.batch(kind, M, N)
        act = np.ones(M, np.uint8)
        act[2] = 0
        b.log_likelihood(np.tile(smc._lib.params8(par), (M, 1)), y2, smc.SYSTEMATIC, 0, act)
        b.step(0.3, smc.STRATIFIED)
        b.close()
os.environ.pop("SMCB_BATCH_CHUNK")
# device-resident θ-level samplers
lg_mod = lambda th: smc.StateSpaceModel(smc.LinearGaussian(th[0], 1.0, th[1], th[2], 0.0), (1, 1))     # noqa: E731
prior = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.LogNormal()])
yl = smc._lib.simulate(smc.KIND_LG1D, LG, 12, 1998)[1]
s = smc.SMC(64, 32, lg_mod, prior, 2, 0.5, seed=3, ctx=ctx, engine="device", resampler="systematic")
smc.smc2(s, yl)
for t in range(1, yl.size):
    smc.smc2_step(s, yl, t, verbose=False)
s.close()
s = smc.SMC(64, 32, lg_mod, prior, 2, 0.5, seed=3, ctx=ctx, engine="device")
smc.density_tempered(s, yl, verbose=False)
s.close()
print("sanitize run ok", float(w.sum()))
