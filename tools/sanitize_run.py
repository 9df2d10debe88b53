"""Small single-filter and batched runs for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import sequential_monte_carlo_b200 as smc
ctx = smc.Context(0, 1998)
LG = [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]
y = smc._lib.simulate(smc.KIND_LG1D, LG, 6, 1998)[1]
for N in (1, 31, 5000, 70001, (1 << 16) + 3):
    for rs in (smc.SYSTEMATIC, smc.STRATIFIED, smc.MULTINOMIAL):
        ctx.log_likelihood(smc.KIND_LG1D, LG, N, y, rs)
# degenerate weights: wide-window path of anc_hist_kernel
ctx.log_likelihood(smc.KIND_LG1D, [0.5, 1.0, 0.9, 1e-9, 0.0, 1.0], 1 << 16, y, smc.SYSTEMATIC)
ctx.log_likelihood(smc.KIND_UCSV, [0.2, 0.2, 3.0, 1.0, 1.0], 4099, y, smc.SYSTEMATIC)
ctx.log_likelihood(smc.KIND_SV, [-1.0, 0.9, 0.3], 2049, y, smc.STRATIFIED)
x, w, lw = ctx.fetch_state(want_logw=True)
b = ctx.batch(smc.KIND_UCSV, 8, 1024)
P = np.tile(smc._lib.params8([0.2, 0.2, 3.0, 1.0, 1.0]), (8, 1))
b.log_likelihood(P, y, smc.SYSTEMATIC)
b.log_likelihood(P, y, smc.MULTINOMIAL)
b.close()
print("sanitize run ok", float(w.sum()))
