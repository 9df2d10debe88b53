"""ctypes front-end of oracle/liboracle.so plus numpy restatements of the reference's host-level
functions.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED: the reference (/root/reference, Julia) has no tests or golden vectors and cannot
run in this image; see oracle/smc_oracle.c for what pins this oracle instead.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KIND_LG1D, KIND_SV, KIND_UCSV = 0, 1, 2
MULTINOMIAL, STRATIFIED, SYSTEMATIC = 0, 1, 2
P_INIT, P_TRANS, P_RESAMPLE, P_PRIOR, P_THETA_RESAMPLE, P_MH_PROPOSAL, P_MH_ACCEPT, P_SIMULATE = 1, 2, 3, 4, 5, 6, 7, 8
P_RESAMPLE_CELL = 9   # in-cell thresholds of the two-level multinomial resampler (docs/SPEC.md §5c)
MN_CELL, MN_CHUNK, MN_LEGACY_MAX = 4096, 8192, 8192

_dp = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("smc_oracle.c", "det_math.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.smco_log_likelihood.restype = C.c_double
        L.smco_kalman_step.restype = C.c_double
        L.smco_kalman_loglik.restype = C.c_double
        L.smco_guided_log_likelihood.restype = C.c_double
        L.smco_reference_style_log_likelihood.restype = C.c_double
        L.smco_kalman_mv_step.restype = C.c_double
        L.smco_kalman_mv_loglik.restype = C.c_double
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def params8(p):
    """8-double parameter block of the univariate kinds / UCSV; a multivariate LG block (3d² + 2d + 1 doubles) passes through"""
    p = np.asarray(p, dtype=np.float64).ravel()
    if p.size > 8:
        return np.ascontiguousarray(p)
    out = np.zeros(8)
    out[: p.size] = p
    return out


def state_dim(kind):
    return kind - 1 if kind >= 3 else (3 if kind == KIND_UCSV else 1)   # kinds 3..5: multivariate LG, d = 2..4


def quant_shift(n):
    return int(lib().smco_quant_shift(C.c_int64(n)))


# ---------------------------------------------------------------- det math / RNG primitives
def philox(ctr, key):
    out = np.zeros(4, np.uint32)
    lib().smco_philox(_p(np.asarray(ctr, np.uint32)), _p(np.asarray(key, np.uint32)), _p(out))
    return out


def det_exp(x):
    x = np.ascontiguousarray(x, np.float64)
    y = np.empty_like(x)
    lib().smco_exp(_p(x), _p(y), C.c_int64(x.size))
    return y


def det_log(x):
    x = np.ascontiguousarray(x, np.float64)
    y = np.empty_like(x)
    lib().smco_log(_p(x), _p(y), C.c_int64(x.size))
    return y


def det_sincos2pi(u):
    u = np.ascontiguousarray(u, np.float64)
    s, c = np.empty_like(u), np.empty_like(u)
    lib().smco_sincos2pi(_p(u), _p(s), _p(c), C.c_int64(u.size))
    return s, c


def det_quant(x, S):
    x = np.ascontiguousarray(x, np.float64)
    q = np.empty(x.shape, np.uint64)
    lib().smco_quant(_p(x), C.c_int(S), _p(q), C.c_int64(x.size))
    return q


def normals(seed, epoch, stream, t, kind, comp, n):
    out = np.empty(n)
    lib().smco_normals(C.c_uint64(seed), C.c_uint32(epoch), C.c_uint32(stream), C.c_uint32(t), C.c_uint32(kind),
                       C.c_uint32(comp), C.c_int64(n), _p(out))
    return out


def uniforms64(seed, epoch, stream, t, kind, n):
    out = np.empty(n, np.uint64)
    lib().smco_uniforms64(C.c_uint64(seed), C.c_uint32(epoch), C.c_uint32(stream), C.c_uint32(t), C.c_uint32(kind),
                          C.c_int64(n), _p(out))
    return out


def uniforms32(seed, epoch, stream, t, kind, n):
    """32-bit uniforms of docs/SPEC.md §5c: element i is word (i & 3) of the Philox block at index i >> 2.  Built on the 64-bit
    stream: U64 of index 2q is (r0 << 32) | r1 of block q, of index 2q + 1 is (r2 << 32) | r3."""
    u64 = uniforms64(seed, epoch, stream, t, kind, 2 * ((n + 3) // 4))
    out = np.empty(4 * ((n + 3) // 4), np.uint32)
    out[0::4] = (u64[0::2] >> np.uint64(32)).astype(np.uint32)
    out[1::4] = (u64[0::2] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    out[2::4] = (u64[1::2] >> np.uint64(32)).astype(np.uint32)
    out[3::4] = (u64[1::2] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return out[:n]


def uniforms01(seed, epoch, stream, t, kind, n):
    """(U64 >> 11) * 2^-53 in [0,1) — used by the host-level MH accept draw."""
    return (uniforms64(seed, epoch, stream, t, kind, n) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


# ---------------------------------------------------------------- a1 / a2
def normalize(logw):
    """normalize(logw) -> (logmu, w, ess)   /root/reference/src/particles.jl:5-15"""
    logw = np.ascontiguousarray(logw, np.float64)
    w = np.empty_like(logw)
    lm, es = C.c_double(), C.c_double()
    lib().smco_normalize(_p(logw), C.c_int64(logw.size), C.byref(lm), _p(w), C.byref(es))
    return lm.value, w, es.value


def normalize_numpy(logw):
    """Line-by-line numpy restatement of normalize with libm exp (tolerance cross-check)."""
    maxw = np.max(logw)
    w = np.exp(logw - maxw)
    sumw = np.sum(w)
    logmu = maxw + np.log(sumw) - np.log(len(logw))
    w = w / sumw
    return logmu, w, 1.0 / np.sum(w ** 2)


def ancestors(logw, resampler, seed, epoch, stream, t):
    """SPEC §5 ancestors from unnormalised log-weights (replaces resample, particles.jl:17-19)."""
    logw = np.ascontiguousarray(logw, np.float64)
    a = np.empty(logw.size, np.int64)
    lib().smco_ancestors(_p(logw), C.c_int64(logw.size), C.c_int(resampler), C.c_uint64(seed), C.c_uint32(epoch),
                         C.c_uint32(stream), C.c_uint32(t), _p(a))
    return a


def ancestors_numpy(logw, resampler, seed, epoch, stream, t):
    """Independent numpy restatement of SPEC §5 (python ints for the 128-bit products)."""
    n = len(logw)
    S = quant_shift(n)
    q = det_quant(np.asarray(logw) - np.max(logw), S)
    Cs = np.cumsum(q, dtype=np.uint64)
    Q = int(Cs[-1])
    R = (2 ** 64 - 1) // n
    U = [int(v) for v in uniforms64(seed, epoch, stream, t, P_RESAMPLE, n)]
    out = np.empty(n, np.int64)
    if resampler == MULTINOMIAL and n > MN_LEGACY_MAX and Q != 0:
        # SPEC §5c, vectorised differently from the C oracle: level-1 cells by one searchsorted, level-2 per cell
        ncells = (n + MN_CELL - 1) // MN_CELL
        last = np.minimum((np.arange(ncells) + 1) * MN_CELL - 1, n - 1)
        cellC = Cs[last]
        U32 = uniforms32(seed, epoch, stream, t, P_RESAMPLE, n)         # word (i & 3) of the Philox block at index i >> 2
        tau1 = np.array([(int(u) * Q) >> 32 for u in U32], dtype=np.uint64)
        K = np.bincount(np.searchsorted(cellC, tau1, side="right"), minlength=ncells)
        V = [int(v) for v in uniforms32(seed, epoch, stream, t, P_RESAMPLE_CELL, n)]
        O = 0
        for c in range(ncells):
            j0, j1 = c * MN_CELL, min((c + 1) * MN_CELL, n)
            base = int(cellC[c - 1]) if c else 0
            W = int(cellC[c]) - base
            loc = Cs[j0:j1] - np.uint64(base)
            for g0 in range(O, O + int(K[c]), MN_CHUNK):
                g1 = min(g0 + MN_CHUNK, O + int(K[c]))
                tau2 = np.array([(V[g] * W) >> 32 for g in range(g0, g1)], dtype=np.uint64)
                out[g0:g1] = np.sort(j0 + np.searchsorted(loc, tau2, side="right"))
            O += int(K[c])
        return out
    for i in range(n):
        if Q == 0:
            out[i] = i
            continue
        if resampler == MULTINOMIAL:
            F = U[i]
        elif resampler == STRATIFIED:
            F = i * R + ((U[i] * R) >> 64)
        else:
            F = i * R + ((U[0] * R) >> 64)
        tau = (F * Q) >> 64
        out[i] = np.searchsorted(Cs, np.uint64(tau), side="right")
    return out


def resample_w(w, resampler, seed, epoch, stream, t, purpose=P_RESAMPLE, n_out=None):
    w = np.ascontiguousarray(w, np.float64)
    if n_out is not None and int(n_out) != w.size:            # resample(w, N) with N != length(w)   particles.jl:17
        a = np.empty(int(n_out), np.int64)
        lib().smco_resample_w_n(_p(w), C.c_int64(w.size), C.c_int64(int(n_out)), C.c_int(resampler), C.c_uint64(seed), C.c_uint32(epoch),
                                C.c_uint32(stream), C.c_uint32(t), C.c_uint32(purpose), _p(a))
        return a
    a = np.empty(w.size, np.int64)
    lib().smco_resample_w(_p(w), C.c_int64(w.size), C.c_int(resampler), C.c_uint64(seed), C.c_uint32(epoch),
                          C.c_uint32(stream), C.c_uint32(t), C.c_uint32(purpose), _p(a))
    return a


def detf(fn, x, S=0):
    """the binary32 functions of docs/SPEC.md §9b on values rounded to binary32: fn in 'exp', 'log', 'sincos', 'quant'"""
    x = np.ascontiguousarray(x, np.float64)
    if fn == "sincos":
        s, c = np.empty_like(x), np.empty_like(x)
        lib().smco_sincos2pif(_p(x), _p(s), _p(c), C.c_int64(x.size))
        return s, c
    if fn == "quant":
        q = np.empty(x.size, np.uint64)
        lib().smco_quantf(_p(x), C.c_int(S), _p(q), C.c_int64(x.size))
        return q
    out = np.empty_like(x)
    getattr(lib(), "smco_expf" if fn == "exp" else "smco_logf")(_p(x), _p(out), C.c_int64(x.size))
    return out


class arith_f32:
    """`with oracle.arith_f32(): ...` runs the filters in SPEC §9b's binary32-ARITHMETIC tier (float normals four per Philox block,
    float model arithmetic, float exponential behind the fixed-point weights); states / log-weights come back as binary32 values."""

    def __init__(self, on=True):
        self.on = bool(on)

    def __enter__(self):
        self.prev = lib().smco_get_arith_f32()
        lib().smco_set_arith_f32(C.c_int(1 if self.on else 0))
        return self

    def __exit__(self, *exc):
        lib().smco_set_arith_f32(C.c_int(self.prev))
        return False


def normalize_f32(logw):
    """normalize() of the binary32-arithmetic tier: expf in binary32, sums in binary64"""
    logw = np.ascontiguousarray(logw, np.float64)
    w = np.empty_like(logw)
    lm, es = C.c_double(), C.c_double()
    lib().smco_normalize_f32(_p(logw), C.c_int64(logw.size), C.byref(lm), _p(w), C.byref(es))
    return lm.value, w, es.value


class state_f32:
    """`with oracle.state_f32(): ...` runs the filters in SPEC §9's binary32-state tier (states rounded to
    binary32 where the device stores them; arithmetic unchanged)."""

    def __init__(self, on=True):
        self.on = bool(on)

    def __enter__(self):
        self.prev = lib().smco_get_state_f32()
        lib().smco_set_state_f32(C.c_int(1 if self.on else 0))
        return self

    def __exit__(self, *exc):
        lib().smco_set_state_f32(C.c_int(self.prev))
        return False


# ---------------------------------------------------------------- a3 / a4 / a5
def bootstrap_init(kind, params, n, y0, seed, epoch=0, stream=0):
    """bootstrap_filter(N, y, model) -> (x [d,n], logw)   particles.jl:87-105 (weights left unnormalised)"""
    x = np.empty((state_dim(kind), n))
    logw = np.empty(n)
    lib().smco_bootstrap_init(C.c_int(kind), _p(params8(params)), C.c_int64(n), C.c_double(y0), C.c_uint64(seed),
                              C.c_uint32(epoch), C.c_uint32(stream), _p(x), _p(logw))
    return x, logw


def bootstrap_step(kind, params, x, logw, y, t, resampler, seed, epoch=0, stream=0):
    """bootstrap_filter!(x, w, y, model): mutates x, logw in place; returns the ancestors.  particles.jl:107-129"""
    n = logw.size
    anc = np.empty(n, np.int64)
    lib().smco_bootstrap_step(C.c_int(kind), _p(params8(params)), C.c_int64(n), C.c_double(y), C.c_uint32(t),
                              C.c_int(resampler), C.c_uint64(seed), C.c_uint32(epoch), C.c_uint32(stream), _p(x),
                              _p(logw), _p(anc))
    return anc


def log_likelihood(kind, params, n, y, resampler, seed, epoch=0, stream=0, want_anc=False):
    """log_likelihood(N, y, model)   particles.jl:132-147.  Returns dict(x, logw, logZ, logmu[T], ess[T], anc)."""
    y = np.ascontiguousarray(y, np.float64)
    T = y.size
    x = np.empty((state_dim(kind), n))
    logw = np.empty(n)
    logmu, ess = np.empty(T), np.empty(T)
    anc = np.zeros((T, n), np.int64) if want_anc else None
    logZ = lib().smco_log_likelihood(C.c_int(kind), _p(params8(params)), C.c_int64(n), _p(y), C.c_int64(T),
                                     C.c_int(resampler), C.c_uint64(seed), C.c_uint32(epoch), C.c_uint32(stream),
                                     _p(x), _p(logw), _p(logmu), _p(ess), _p(anc))
    return dict(x=x, logw=logw, logZ=logZ, logmu=logmu, ess=ess, anc=anc)


def batch_log_likelihood(kind, params, active, n, y, resampler, seed, epoch, stream0=0, want_state=True):
    """M filters threaded over theta (smc_samplers.jl:112-121,223-229).  params: [M,8]."""
    params = np.ascontiguousarray(params, np.float64)
    M = params.shape[0]
    y = np.ascontiguousarray(y, np.float64)
    d = state_dim(kind)
    logZ = np.empty(M)
    x = np.empty((M, d, n)) if want_state else None
    logw = np.empty((M, n)) if want_state else None
    act = None if active is None else np.ascontiguousarray(active, np.uint8)
    lib().smco_batch_log_likelihood(C.c_int(kind), _p(params), _p(act), C.c_int64(M), C.c_int64(n), _p(y),
                                    C.c_int64(y.size), C.c_int(resampler), C.c_uint64(seed), C.c_uint32(epoch),
                                    C.c_uint32(stream0), _p(logZ), _p(x), _p(logw))
    return logZ, x, logw


def reference_style_log_likelihood(params, n, y, seed):
    """The timed CPU arm of BASELINE.md §3 (LG1D): particles.jl:87-147 with the reference's own costs — alias-table multinomial
    resampling rebuilt on every step, fresh allocations per step, per-particle sqrt / log σ; its own RNG (xoshiro256++, polar
    normals), so it is checked distributionally only.  Returns logZ."""
    y = np.ascontiguousarray(y, np.float64)
    p = np.ascontiguousarray(np.asarray(params, np.float64).ravel()[:6])
    return float(lib().smco_reference_style_log_likelihood(_p(p), C.c_int64(int(n)), _p(y), C.c_int64(y.size), C.c_uint64(int(seed))))


def reference_style_batch(kind, params, n, y, seed):
    """M reference-style filters threaded over θ like `Threads.@threads for m` (smc_samplers.jl:112); kind 0 (LG1D) or 2 (UCSV);
    params [M, 8].  Returns logZ [M].  Timing arm only (its own RNG)."""
    params = np.ascontiguousarray(params, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    z = np.empty(params.shape[0])
    rc = lib().smco_reference_style_batch(C.c_int(int(kind)), _p(params), C.c_int64(params.shape[0]), C.c_int64(int(n)), _p(y), C.c_int64(y.size),
                                          C.c_uint64(int(seed)), _p(z))
    if rc != 0:
        raise ValueError("reference-style arm: LG1D and UCSV only")
    return z


# ---------------------------------------------------------------- N3 guided filter (SPEC §10)
def guided_step(kind, params, x, logw, y, t, resampler, prop, seed, epoch=0, stream=0):
    """particle_filter!(x, w, y, model, proposal) with proposal = Normal(c0 + c1 xp, c2): mutates x [1,n] / logw in
    place, returns the ancestors.  particles.jl:55-84"""
    n = logw.size
    anc = np.empty(n, np.int64)
    prop = np.ascontiguousarray(prop, np.float64)
    rc = lib().smco_guided_step(C.c_int(kind), _p(params8(params)), C.c_int64(n), C.c_double(y), C.c_uint32(t),
                                C.c_int(resampler), C.c_uint64(seed), C.c_uint32(epoch), C.c_uint32(stream), _p(prop),
                                _p(x), _p(logw), _p(anc))
    if rc != 0:
        raise ValueError("guided proposals are defined for LG1D, SV (affine-Gaussian, SPEC §10) and UCSV (tempered optimal trend move, §10b)")
    return anc


def guided_log_likelihood(kind, params, n, y, resampler, prop, seed, epoch=0, stream=0):
    """bootstrap initial step + guided steps; prop: [T, 3] (row 0 unused; UCSV: (κ, ·, ·), SPEC §10b).  dict(x, logw, logZ, logmu[T], ess[T])."""
    y = np.ascontiguousarray(y, np.float64)
    T = y.size
    prop = np.ascontiguousarray(prop, np.float64).reshape(T, 3)
    x, logw = np.empty((state_dim(kind), n)), np.empty(n)
    logmu, ess = np.empty(T), np.empty(T)
    logZ = lib().smco_guided_log_likelihood(C.c_int(kind), _p(params8(params)), C.c_int64(n), _p(y), C.c_int64(T),
                                            C.c_int(resampler), C.c_uint64(seed), C.c_uint32(epoch), C.c_uint32(stream),
                                            _p(prop), C.c_int64(3), _p(x), _p(logw), _p(logmu), _p(ess))
    return dict(x=x, logw=logw, logZ=logZ, logmu=logmu, ess=ess)


def batch_guided_log_likelihood(kind, params, active, n, y, resampler, prop, seed, epoch, stream0=0):
    """M guided filters threaded over theta; params [M,8], prop [T,M,3].  Returns (logZ[M], x[M,1,n], logw[M,n])."""
    params = np.ascontiguousarray(params, np.float64)
    M = params.shape[0]
    y = np.ascontiguousarray(y, np.float64)
    prop = np.ascontiguousarray(prop, np.float64).reshape(y.size, M, 3)
    logZ, x, logw = np.empty(M), np.empty((M, state_dim(kind), n)), np.empty((M, n))
    act = None if active is None else np.ascontiguousarray(active, np.uint8)
    lib().smco_batch_guided_log_likelihood(C.c_int(kind), _p(params), _p(act), C.c_int64(M), C.c_int64(n), _p(y),
                                           C.c_int64(y.size), C.c_int(resampler), C.c_uint64(seed), C.c_uint32(epoch),
                                           C.c_uint32(stream0), _p(prop), _p(logZ), _p(x), _p(logw))
    return logZ, x, logw


def optimal_proposal_lg(params, y):
    """(c0, c1, c2) of the locally optimal proposal p(x' | xp, y) of an LG1D model (A,B,Q,R,..): precision
    1/Q + B²/R, mean s²(A xp / Q + B y / R)."""
    A, B, Q, R = (float(v) for v in np.asarray(params, np.float64)[:4])
    s2 = 1.0 / (1.0 / Q + B * B / R)
    return np.array([s2 * B * float(y) / R, s2 * A / Q, np.sqrt(s2)])


def weighted_summary(x, logw, probs, weighted=True):
    """SPEC §8 restated with numpy: fixed-point weights q_i of the current log-weights (SPEC §5), population mean /
    variance, and lower empirical quantiles = the smallest x whose cumulative q exceeds r = min(floor(p Q), Q - 1).
    x: [d, n].  Stands in for quantile(x, weights(w), p) / var(x, weights(w)) of examples/inflation_example.jl:44-46
    (StatsBase, un-pinned: SURVEY F8) and for quantile(x, p) of README.md:41,51 when weighted=False."""
    x = np.atleast_2d(np.asarray(x, np.float64))
    n = x.shape[1]
    if weighted:
        lw = np.asarray(logw, np.float64)
        q = det_quant(lw - lw.max(), quant_shift(n)).astype(object)    # exact Python integers
    else:
        q = np.ones(n, dtype=object)
    Q = int(q.sum())
    qd = np.array([float(v) for v in q])
    probs = np.atleast_1d(np.asarray(probs, np.float64))
    mean, var, quant = np.empty(x.shape[0]), np.empty(x.shape[0]), np.empty((x.shape[0], probs.size))
    for c in range(x.shape[0]):
        keep = qd != 0
        mean[c] = np.sum(qd[keep] * x[c][keep]) / float(Q)
        var[c] = np.sum(qd[keep] * (x[c][keep] - mean[c]) ** 2) / float(Q)
        order = np.argsort(x[c], kind="stable")
        cum = np.cumsum(q[order])                                       # object dtype: exact
        for j, p in enumerate(probs):
            r = min(int(np.uint64(np.float64(p) * np.float64(Q))) if p * float(Q) < 1.8e19 else Q - 1, Q - 1)
            k = next(i for i, cv in enumerate(cum) if cv > r)
            quant[c, j] = x[c][order][k]
    return mean, var, quant


def num_threads():
    return int(lib().smco_num_threads())


# ---------------------------------------------------------------- Kalman / simulate
def kalman_step(params, x, S, y, predict=True):
    xs, Ss = C.c_double(x), C.c_double(S)
    ll = lib().smco_kalman_step(_p(params8(params)), C.byref(xs), C.byref(Ss), C.c_double(y), C.c_int(int(predict)))
    return xs.value, Ss.value, ll


def kalman_loglik(params, y, matched_init=False):
    y = np.ascontiguousarray(y, np.float64)
    xT, ST = C.c_double(), C.c_double()
    ll = lib().smco_kalman_loglik(_p(params8(params)), _p(y), C.c_int64(y.size), C.c_int(int(matched_init)),
                                  C.byref(xT), C.byref(ST))
    return xT.value, ST.value, ll


def mv_block(A, B, Q, R, x0, S0):
    """row-major model block A[d][d], B[d], Q[d][d], R, x0[d], Σ0[d][d] of a multivariate LinearModel"""
    return np.concatenate([np.asarray(A, np.float64).ravel(), np.asarray(B, np.float64).ravel(), np.asarray(Q, np.float64).ravel(),
                           np.asarray(R, np.float64).ravel()[:1], np.asarray(x0, np.float64).ravel(), np.asarray(S0, np.float64).ravel()])


def kalman_mv_step(d, block, x, S, y, predict=True):
    """kalman_filter(model, xt, Σt, yt) for a multivariate LinearModel   kalman_filter.jl:3-27"""
    block = np.ascontiguousarray(block, np.float64)
    x = np.array(x, np.float64).ravel().copy()
    S = np.array(S, np.float64).reshape(d, d).copy()
    ll = lib().smco_kalman_mv_step(C.c_int(d), _p(block), _p(x), _p(S), C.c_double(y), C.c_int(int(predict)))
    return x, S, ll


def kalman_mv_loglik(d, block, y, matched_init=False):
    block = np.ascontiguousarray(block, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    x, S = np.empty(d), np.empty((d, d))
    ll = lib().smco_kalman_mv_loglik(C.c_int(d), _p(block), _p(y), C.c_int64(y.size), C.c_int(int(matched_init)), _p(x), _p(S))
    return x, S, ll


def simulate(kind, params, T, seed):
    x = np.empty((state_dim(kind), T))
    y = np.empty(T)
    lib().smco_simulate(C.c_int(kind), _p(params8(params)), C.c_int64(T), C.c_uint64(seed), _p(x), _p(y))
    return x, y
