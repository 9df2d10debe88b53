import sys, time
sys.path.insert(0, ".")
import numpy as np
import sequential_monte_carlo_b200 as smc
ctx = smc.Context(0, 1)
N = 1 << 24
y = np.zeros(5)
ctx.log_likelihood(smc.KIND_LG1D, [0.5, 1, 0.9, 0.8, 0, 1], N, y, smc.SYSTEMATIC)
for rep in range(3):
    t0 = time.perf_counter(); x, _, _ = ctx.fetch_state(want_x=True, want_w=False); t1 = time.perf_counter()
    _, w, _ = ctx.fetch_state(want_x=False, want_w=True); t2 = time.perf_counter()
    print(f"fetch x {1e3*(t1-t0):.1f} ms, fetch w {1e3*(t2-t1):.1f} ms", x.flags['OWNDATA'], flush=True)
a = np.empty(N)
import ctypes as C
t0 = time.perf_counter(); ctx._check(ctx._lib.smcb_fetch_state(ctx._h, a.ctypes.data_as(C.c_void_p), None, None)); print("pageable x", 1e3*(time.perf_counter()-t0))
