"""ctypes binding of libsmcb200.so (include/smcb200.h) — the only way the Python host code reaches
the GPU.  There is no CPU fallback: a missing library or a missing CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

LG1D, SV, UCSV = 0, 1, 2
MVLG2, MVLG3, MVLG4 = 3, 4, 5     # MultivariateLinearGaussian, d = 2..4 (single filter only); params = the A, B, Q, R, x0, Σ0 block
MULTINOMIAL, STRATIFIED, SYSTEMATIC = 0, 1, 2
PARAM_STRIDE = 8
# Philox purposes of the host-level streams (docs/SPEC.md §2)
P_PRIOR, P_THETA_RESAMPLE, P_MH_PROPOSAL, P_MH_ACCEPT, P_SIMULATE = 4, 5, 6, 7, 8

_c_ctx = C.c_void_p
_c_batch = C.c_void_p
_c_sampler = C.c_void_p
PRIOR_NORMAL, PRIOR_LOGNORMAL, PRIOR_UNIFORM, PRIOR_TRUNCNORMAL = 0, 1, 2, 3
MAX_THETA_DIM, MAX_THETA_PARTICLES = 8, 16384


class SamplerConfig(C.Structure):
    """smcb_sampler_config (include/smcb200.h)"""
    _fields_ = [("kind", C.c_int32), ("d_theta", C.c_int32), ("N", C.c_int64), ("M", C.c_int64), ("chain", C.c_int32),
                ("resampler", C.c_int32), ("theta_resampler", C.c_int32), ("reserved", C.c_int32), ("ess_threshold", C.c_double),
                ("min_ar", C.c_double), ("seed", C.c_uint64), ("prior", (C.c_double * 8) * 8), ("map_src", C.c_int32 * 8),
                ("map_const", C.c_double * 8)]

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol include/smcb200.h declares
SIGNATURES = {
    "smcb_create": (C.c_int, [C.c_int, C.c_uint64, C.POINTER(_c_ctx)]),
    "smcb_destroy": (C.c_int, [_c_ctx]),
    "smcb_last_error": (C.c_char_p, [_c_ctx]),
    "smcb_version": (C.c_int, []),
    "smcb_state_dim": (C.c_int, [C.c_int]),
    "smcb_set_rng": (C.c_int, [_c_ctx, C.c_uint64, C.c_uint32]),
    "smcb_get_epoch": (C.c_int, [_c_ctx, C.POINTER(C.c_uint32)]),
    "smcb_record_ancestors": (C.c_int, [_c_ctx, C.c_int]),
    "smcb_set_profiling": (C.c_int, [_c_ctx, C.c_int]),
    "smcb_set_precision": (C.c_int, [_c_ctx, C.c_int]),
    "smcb_get_timing": (C.c_int, [_c_ctx, _dp, _i64p]),
    "smcb_synchronize": (C.c_int, [_c_ctx]),
    "smcb_alloc_pinned": (C.c_int, [_c_ctx, C.c_int64, C.POINTER(C.c_void_p)]),
    "smcb_free_pinned": (C.c_int, [_c_ctx, C.c_void_p]),
    "smcb_normalize": (C.c_int, [_c_ctx, C.c_void_p, C.c_int64, _dp, C.c_void_p, _dp]),
    "smcb_resample": (C.c_int, [_c_ctx, C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    "smcb_resample_n": (C.c_int, [_c_ctx, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    "smcb_bootstrap_init": (C.c_int, [_c_ctx, C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_uint32, _dp, _dp]),
    "smcb_bootstrap_step": (C.c_int, [_c_ctx, C.c_void_p, C.c_double, C.c_int, _dp, _dp]),
    "smcb_log_likelihood": (C.c_int, [_c_ctx, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_uint32,
                                      _dp, C.c_void_p, C.c_void_p]),
    "smcb_guided_step": (C.c_int, [_c_ctx, C.c_void_p, C.c_double, C.c_int, C.c_void_p, _dp, _dp]),
    "smcb_guided_log_likelihood": (C.c_int, [_c_ctx, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_uint32,
                                             C.c_void_p, _dp, C.c_void_p, C.c_void_p]),
    "smcb_fetch_state": (C.c_int, [_c_ctx, C.c_void_p, C.c_void_p, C.c_void_p]),
    "smcb_fetch_ancestors": (C.c_int, [_c_ctx, C.c_void_p, C.c_int64, _i64p]),
    "smcb_device_state": (C.c_int, [_c_ctx, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _i64p]),
    "smcb_batch_create": (C.c_int, [_c_ctx, C.c_int, C.c_int64, C.c_int64, C.POINTER(_c_batch)]),
    "smcb_batch_destroy": (C.c_int, [_c_batch]),
    "smcb_batch_init": (C.c_int, [_c_batch, C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_void_p, C.c_void_p]),
    "smcb_batch_step": (C.c_int, [_c_batch, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]),
    "smcb_batch_log_likelihood": (C.c_int, [_c_batch, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_uint32,
                                            C.c_void_p]),
    "smcb_batch_step_guided": (C.c_int, [_c_batch, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "smcb_batch_log_likelihood_guided": (C.c_int, [_c_batch, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_uint32,
                                                   C.c_void_p, C.c_void_p]),
    "smcb_batch_gather": (C.c_int, [_c_batch, C.c_void_p]),
    "smcb_batch_accept": (C.c_int, [_c_batch, _c_batch, C.c_void_p]),
    "smcb_batch_fetch": (C.c_int, [_c_batch, C.c_void_p, C.c_void_p, C.c_void_p]),
    "smcb_batch_weighted_mean": (C.c_int, [_c_batch, C.c_void_p]),
    "smcb_batch_weighted_moments": (C.c_int, [_c_batch, C.c_void_p, C.c_void_p]),
    "smcb_batch_weighted_quantiles": (C.c_int, [_c_batch, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "smcb_batch_cloud_bytes": (C.c_int64, [_c_batch]),
    "smcb_batch_pack": (C.c_int, [_c_batch, C.c_void_p, C.c_int64, C.c_void_p]),
    "smcb_batch_unpack": (C.c_int, [_c_batch, C.c_void_p, C.c_int64, C.c_void_p]),
    "smcb_batch_get_timing": (C.c_int, [_c_batch, _dp, _i64p]),
    "smcb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "smcb_comm_init": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p]),
    "smcb_comm_rank": (C.c_int, [_c_ctx, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "smcb_comm_all_gather": (C.c_int, [_c_ctx, C.c_void_p, C.c_int64, C.c_void_p]),
    "smcb_comm_destroy": (C.c_int, [_c_ctx]),
    "smcb_batch_chunk_plan": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_int, _i64p]),
    "smcb_exchange_plan": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, _i64p,
                                     C.c_void_p, C.c_void_p, _i64p]),
    "smcb_sampler_create": (C.c_int, [_c_ctx, C.POINTER(SamplerConfig), C.c_void_p, C.POINTER(_c_sampler)]),
    "smcb_sampler_destroy": (C.c_int, [_c_sampler]),
    "smcb_sampler_set_data": (C.c_int, [_c_sampler, C.c_void_p, C.c_int64]),
    "smcb_sampler_smc2_init": (C.c_int, [_c_sampler]),
    "smcb_sampler_smc2_step": (C.c_int, [_c_sampler, C.c_int64, _dp, C.POINTER(C.c_int)]),
    "smcb_sampler_density_tempered": (C.c_int, [_c_sampler, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "smcb_sampler_get": (C.c_int, [_c_sampler, C.c_void_p, C.c_void_p, C.c_void_p, _dp, _dp, _i64p]),
    "smcb_sampler_clouds": (C.c_int, [_c_sampler, C.POINTER(_c_batch)]),
    "smcb_sampler_set_profiling": (C.c_int, [_c_sampler, C.c_int]),
    "smcb_sampler_stats": (C.c_int, [_c_sampler, _dp, _i64p]),
    "smcb_random_walk_sigma": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "smcb_cholesky_lower": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_void_p]),
    "smcb_kalman_batch_step": (C.c_int, [_c_ctx, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "smcb_kalman_batch_loglik": (C.c_int, [_c_ctx, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "smcb_kalman_mv_batch_step": (C.c_int, [_c_ctx, C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "smcb_kalman_mv_batch_loglik": (C.c_int, [_c_ctx, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "smcb_rng_normals": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int64, C.c_void_p]),
    "smcb_rng_uniforms64": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int64, C.c_void_p]),
    "smcb_simulate": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "smcb_weighted_summary": (C.c_int, [_c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "smcb_selftest_math": (C.c_int, [_c_ctx, C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p]),
}

_LIB = None


class SMCBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsmcb200 error {code}: {msg}")
        self.code = code


def library_path():
    return _build.LIB_PATH


def load():
    """dlopen libsmcb200.so (rebuilding it first when a source or header is newer than the binary: a stale
    library would silently break the bit-for-bit agreement of host, device and oracle) and bind every symbol."""
    global _LIB
    if _LIB is None:
        path = _build.build_library()   # returns at once when the binary is up to date
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def params8(p):
    """[..., k<=8] -> contiguous [..., 8] float64 parameter block(s); a multivariate LG block (more than 8 doubles) passes through."""
    p = np.asarray(p, dtype=np.float64)
    if p.shape[-1] > PARAM_STRIDE:
        return np.ascontiguousarray(p)
    out = np.zeros(p.shape[:-1] + (PARAM_STRIDE,), dtype=np.float64)
    out[..., : p.shape[-1]] = p
    return np.ascontiguousarray(out)


def state_dim(kind):
    return kind - 1 if kind >= MVLG2 else (3 if kind == UCSV else 1)


def rng_normals(seed, epoch, stream, t, purpose, comp, n):
    """n standard normals of the host-level Philox stream (docs/SPEC.md §2); CPU only."""
    out = np.empty(int(n))
    rc = load().smcb_rng_normals(C.c_uint64(int(seed) & (2 ** 64 - 1)), int(epoch), int(stream), int(t), int(purpose), int(comp),
                                 int(n), _ptr(out))
    if rc != 0:
        raise SMCBError(rc, "smcb_rng_normals: bad arguments")
    return out


def rng_uniforms64(seed, epoch, stream, t, purpose, n):
    out = np.empty(int(n), np.uint64)
    rc = load().smcb_rng_uniforms64(C.c_uint64(int(seed) & (2 ** 64 - 1)), int(epoch), int(stream), int(t), int(purpose), int(n),
                                    _ptr(out))
    if rc != 0:
        raise SMCBError(rc, "smcb_rng_uniforms64: bad arguments")
    return out


def rng_uniforms01(seed, epoch, stream, t, purpose, n):
    """(U64 >> 11) * 2^-53 in [0, 1)."""
    return (rng_uniforms64(seed, epoch, stream, t, purpose, n) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def random_walk_sigma(θ):
    """Σ of random_walk_kernel(θ) for θ [M, d] in the fixed summation order of docs/SPEC.md §11 (host-only)."""
    θ = np.ascontiguousarray(θ, np.float64)
    M, d = θ.shape
    out = np.zeros((d, d))
    rc = load().smcb_random_walk_sigma(_ptr(θ), M, d, _ptr(out))
    if rc != 0:
        raise SMCBError(rc, "smcb_random_walk_sigma: bad arguments")
    return out


def cholesky_lower(A, scale=1.0):
    """lower Cholesky factor of scale·A by the plain Cholesky–Banachiewicz recursion (docs/SPEC.md §11, host-only)."""
    A = np.ascontiguousarray(A, np.float64)
    d = A.shape[0]
    L = np.zeros((d, d))
    rc = load().smcb_cholesky_lower(_ptr(A), d, float(scale), _ptr(L))
    if rc != 0:
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    return L


def batch_chunk_plan(M, N, steps, threads, slots, num_sms=148, masked=False):
    """smcb_batch_chunk_plan: steps per chunk of the dynamically scheduled batched sweep, 0 = one CTA per θ (host-only)"""
    out = C.c_int64(0)
    rc = load().smcb_batch_chunk_plan(int(M), int(N), int(steps), int(threads), int(slots), int(num_sms), int(bool(masked)), C.byref(out))
    if rc != 0:
        raise SMCBError(rc, "smcb_batch_chunk_plan: bad arguments")
    return int(out.value)


def exchange_plan(parents, rank, world):
    """smcb_exchange_plan: (local_parents [M/G], [(peer, slot)] to send, [(peer, slot)] to receive) — host-only"""
    a = np.ascontiguousarray(parents, np.int32)
    Mloc = a.size // world
    lp = np.empty(Mloc, np.int32)
    sp, ss = np.empty(a.size, np.int32), np.empty(a.size, np.int32)   # one cloud may go to every slot of every other rank
    rp, rs = np.empty(Mloc, np.int32), np.empty(Mloc, np.int32)
    ns, nr = C.c_int64(), C.c_int64()
    rc = load().smcb_exchange_plan(_ptr(a), a.size, int(rank), int(world), _ptr(lp), _ptr(sp), _ptr(ss), C.byref(ns), _ptr(rp), _ptr(rs),
                                   C.byref(nr))
    if rc != 0:
        raise SMCBError(rc, "smcb_exchange_plan: bad arguments")
    return lp, list(zip(sp[: ns.value].tolist(), ss[: ns.value].tolist())), list(zip(rp[: nr.value].tolist(), rs[: nr.value].tolist()))


def comm_unique_id():
    """128 bytes identifying a new NCCL communicator (call on ONE rank, carry to the others, then Context.comm_init)"""
    buf = (C.c_uint8 * 128)()
    rc = load().smcb_comm_unique_id(buf)
    if rc != 0:
        msg = load().smcb_last_error(None)
        raise SMCBError(rc, msg.decode() if msg else "smcb_comm_unique_id failed")
    return bytes(buf)


def simulate(kind, params, T, seed):
    """simulate(model, T) -> (x [d, T], y [T])   /root/reference/src/state_space_models.jl:11-28; CPU only."""
    p = params8(np.asarray(params, np.float64).ravel())
    x = np.empty((state_dim(kind), int(T)))
    y = np.empty(int(T))
    rc = load().smcb_simulate(int(kind), _ptr(p), int(T), C.c_uint64(int(seed) & (2 ** 64 - 1)), _ptr(x), _ptr(y))
    if rc != 0:
        raise SMCBError(rc, "smcb_simulate: bad arguments")
    return x, y


class Context:
    """One GPU + one stream + one single-filter slot (smcb_ctx)."""

    def __init__(self, device=0, seed=0):
        self._lib = load()
        self._h = _c_ctx()
        rc = self._lib.smcb_create(int(device), C.c_uint64(int(seed) & (2 ** 64 - 1)), C.byref(self._h))
        if rc != 0:
            msg = self._lib.smcb_last_error(None)
            raise SMCBError(rc, msg.decode() if msg else "smcb_create failed")
        self.device = int(device)
        self.seed = int(seed)
        self._N = 0
        self._kind = LG1D
        self._T = 0
        self._pinned = {}

    def close(self):
        if getattr(self, "_h", None):
            for ptr, _ in getattr(self, "_pinned", {}).values():
                self._lib.smcb_free_pinned(self._h, ptr)
            self._pinned = {}
            self._lib.smcb_destroy(self._h)
            self._h = None

    def pinned_array(self, name, shape):
        """float64 numpy view of a page-locked buffer owned by this context (reused by later calls
        with the same name: copy it if you need the values after the next fetch)."""
        n = int(np.prod(shape))
        ent = self._pinned.get(name)
        if ent is None or ent[1] < n:
            if ent is not None:
                self._lib.smcb_free_pinned(self._h, ent[0])
            ptr = C.c_void_p()
            self._check(self._lib.smcb_alloc_pinned(self._h, 8 * n, C.byref(ptr)))
            ent = (ptr, n)
            self._pinned[name] = ent
        buf = (C.c_double * n).from_address(ent[0].value)
        return np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.smcb_last_error(self._h)
            raise SMCBError(rc, msg.decode() if msg else "")

    # ---- rng / instrumentation
    def set_rng(self, seed, epoch=0):
        self.seed = int(seed)
        self._check(self._lib.smcb_set_rng(self._h, C.c_uint64(int(seed) & (2 ** 64 - 1)), int(epoch)))

    def next_epoch(self):
        e = C.c_uint32()
        self._check(self._lib.smcb_get_epoch(self._h, C.byref(e)))
        return e.value

    def record_ancestors(self, on=True):
        self._check(self._lib.smcb_record_ancestors(self._h, int(bool(on))))

    def set_profiling(self, on=True):
        self._check(self._lib.smcb_set_profiling(self._h, int(bool(on))))

    def set_precision(self, precision):
        """State storage of the single filter from the next bootstrap_init / log_likelihood on: "f64" (default)
        or "f32" (docs/SPEC.md §9: states rounded to binary32 where stored, arithmetic binary64, sorted resamplers)."""
        p = {"f64": 0, "f32": 1, 0: 0, 1: 1, 2: 2, "float64": 0, "float32": 1, "f32_states": 1, "f32_arith": 2, "f32_arithmetic": 2}[precision]
        self._check(self._lib.smcb_set_precision(self._h, p))

    def timing(self):
        ms = (C.c_double * 7)()
        n = (C.c_int64 * 7)()
        self._check(self._lib.smcb_get_timing(self._h, ms, n))
        keys = ("total", "scan", "prop", "init", "stats", "bounds", "anc")
        return {k: float(ms[i]) for i, k in enumerate(keys)}, {k: int(n[i]) for i, k in enumerate(keys)}

    def synchronize(self):
        self._check(self._lib.smcb_synchronize(self._h))

    # ---- utilities
    def normalize(self, logw, want_w=True):
        logw = np.ascontiguousarray(logw, np.float64)
        w = np.empty_like(logw) if want_w else None
        lm, es = C.c_double(), C.c_double()
        self._check(self._lib.smcb_normalize(self._h, _ptr(logw), logw.size, C.byref(lm), _ptr(w), C.byref(es)))
        return lm.value, w, es.value

    def resample(self, w, resampler=MULTINOMIAL, stream=0, t=0, purpose=3, n_out=None):
        w = np.ascontiguousarray(w, np.float64)
        if n_out is not None and int(n_out) != w.size:
            a = np.empty(int(n_out), np.int64)
            self._check(self._lib.smcb_resample_n(self._h, _ptr(w), w.size, int(n_out), int(resampler), int(stream), int(t), int(purpose), _ptr(a)))
            return a
        a = np.empty(w.size, np.int64)
        self._check(self._lib.smcb_resample(self._h, _ptr(w), w.size, int(resampler), int(stream), int(t), int(purpose), _ptr(a)))
        return a

    # ---- one filter
    def bootstrap_init(self, kind, params, N, y, stream=0):
        p = params8(params)
        lm, es = C.c_double(), C.c_double()
        self._check(self._lib.smcb_bootstrap_init(self._h, int(kind), _ptr(p), int(N), float(y), int(stream), C.byref(lm), C.byref(es)))
        self._N, self._kind, self._T = int(N), int(kind), 1
        return lm.value, es.value

    def bootstrap_step(self, y, resampler=MULTINOMIAL, params=None):
        p = None if params is None else params8(params)
        lm, es = C.c_double(), C.c_double()
        self._check(self._lib.smcb_bootstrap_step(self._h, _ptr(p), float(y), int(resampler), C.byref(lm), C.byref(es)))
        self._T += 1
        return lm.value, es.value

    def log_likelihood(self, kind, params, N, y, resampler=MULTINOMIAL, stream=0, per_step=False):
        p = params8(params)
        y = np.ascontiguousarray(y, np.float64)
        T = y.size
        z = C.c_double()
        lm = np.empty(T) if per_step else None
        es = np.empty(T) if per_step else None
        self._check(self._lib.smcb_log_likelihood(self._h, int(kind), _ptr(p), int(N), _ptr(y), T, int(resampler), int(stream),
                                                  C.byref(z), _ptr(lm), _ptr(es)))
        self._N, self._kind, self._T = int(N), int(kind), T
        return (z.value, lm, es) if per_step else z.value

    def guided_step(self, y, proposal, resampler=SYSTEMATIC, params=None):
        """particle_filter! with proposal = (c0, c1, c2) on the single large-N filter (docs/SPEC.md §10; sorted resamplers)"""
        p = None if params is None else params8(params)
        q = np.ascontiguousarray(proposal, np.float64).reshape(3)
        lm, es = C.c_double(), C.c_double()
        self._check(self._lib.smcb_guided_step(self._h, _ptr(p), float(y), int(resampler), _ptr(q), C.byref(lm), C.byref(es)))
        self._T += 1
        return lm.value, es.value

    def guided_log_likelihood(self, kind, params, N, y, proposal, resampler=SYSTEMATIC, stream=0, per_step=False):
        """whole series of the guided single filter in one call; proposal [T, 3] (row 0 unused)"""
        p = params8(params)
        y = np.ascontiguousarray(y, np.float64)
        T = y.size
        q = np.ascontiguousarray(proposal, np.float64).reshape(T, 3)
        z = C.c_double()
        lm = np.empty(T) if per_step else None
        es = np.empty(T) if per_step else None
        self._check(self._lib.smcb_guided_log_likelihood(self._h, int(kind), _ptr(p), int(N), _ptr(y), T, int(resampler), int(stream),
                                                         _ptr(q), C.byref(z), _ptr(lm), _ptr(es)))
        self._N, self._kind, self._T = int(N), int(kind), T
        return (z.value, lm, es) if per_step else z.value

    def fetch_state(self, want_x=True, want_w=True, want_logw=False):
        d = state_dim(self._kind)
        big = self._N >= (1 << 17)   # large clouds land in page-locked buffers (views, reused by the next fetch)
        x = (self.pinned_array("x", (d, self._N)) if big else np.empty((d, self._N))) if want_x else None
        w = (self.pinned_array("w", (self._N,)) if big else np.empty(self._N)) if want_w else None
        lw = (self.pinned_array("logw", (self._N,)) if big else np.empty(self._N)) if want_logw else None
        self._check(self._lib.smcb_fetch_state(self._h, _ptr(x), _ptr(w), _ptr(lw)))
        return x, w, lw

    def summary(self, probs=(), weighted=True):
        """(mean [d], var [d], quantiles [d, len(probs)]) of the current cloud, computed on the device."""
        d = state_dim(self._kind)
        p = np.ascontiguousarray(probs, np.float64).ravel()
        mean, var, q = np.empty(d), np.empty(d), np.empty((d, p.size))
        self._check(self._lib.smcb_weighted_summary(self._h, _ptr(p) if p.size else None, int(p.size), int(bool(weighted)),
                                                    _ptr(mean), _ptr(var), _ptr(q) if p.size else None))
        return mean, var, q

    def fetch_ancestors(self, rows):
        a = np.empty((max(int(rows), 1), self._N), np.int64)
        got = C.c_int64()
        self._check(self._lib.smcb_fetch_ancestors(self._h, _ptr(a), a.shape[0], C.byref(got)))
        return a[: got.value]

    def selftest_math(self, fn, values, aux=0.0):
        v = np.ascontiguousarray(values, np.float64)
        o0, o1 = np.empty_like(v), np.empty_like(v)
        self._check(self._lib.smcb_selftest_math(self._h, int(fn), _ptr(v), v.size, float(aux), _ptr(o0), _ptr(o1)))
        return o0, o1

    # ---- Kalman
    def kalman_step(self, params, x, sigma, y):
        p = params8(params).reshape(-1, PARAM_STRIDE)
        M = p.shape[0]
        x = np.ascontiguousarray(np.broadcast_to(np.asarray(x, np.float64), (M,))).copy()
        s = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma, np.float64), (M,))).copy()
        ll = np.empty(M)
        self._check(self._lib.smcb_kalman_batch_step(self._h, _ptr(p), M, float(y), _ptr(x), _ptr(s), _ptr(ll)))
        return x, s, ll

    def kalman_loglik(self, params, y, matched_init=False, active=None):
        p = params8(params).reshape(-1, PARAM_STRIDE)
        M = p.shape[0]
        y = np.ascontiguousarray(y, np.float64)
        act = None if active is None else np.ascontiguousarray(active, np.uint8)
        ll, x, s = np.empty(M), np.empty(M), np.empty(M)
        self._check(self._lib.smcb_kalman_batch_loglik(self._h, _ptr(p), _ptr(act), M, _ptr(y), y.size, int(bool(matched_init)),
                                                       _ptr(ll), _ptr(x), _ptr(s)))
        return ll, x, s

    @staticmethod
    def _mv_blocks(d, models):
        stride = 3 * d * d + 2 * d + 1
        b = np.ascontiguousarray(models, np.float64).reshape(-1, stride)
        return b, b.shape[0]

    def kalman_mv_step(self, d, models, x, sigma, y):
        """M × kalman_filter(model, x, Σ, y) for multivariate LinearModels (kalman_filter.jl:3-27); models [M, 3d²+2d+1]
        = A, B, Q, R, x0, Σ0 row-major.  Returns x [M, d], Σ [M, d, d], step log-likelihoods [M]."""
        b, M = self._mv_blocks(d, models)
        x = np.ascontiguousarray(np.broadcast_to(np.asarray(x, np.float64), (M, d))).copy()
        s = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma, np.float64), (M, d, d))).copy()
        ll = np.empty(M)
        self._check(self._lib.smcb_kalman_mv_batch_step(self._h, int(d), _ptr(b), M, float(y), _ptr(x), _ptr(s), _ptr(ll)))
        return x, s, ll

    def kalman_mv_loglik(self, d, models, y, matched_init=False, active=None):
        b, M = self._mv_blocks(d, models)
        y = np.ascontiguousarray(y, np.float64)
        act = None if active is None else np.ascontiguousarray(active, np.uint8)
        ll, x, s = np.empty(M), np.empty((M, d)), np.empty((M, d, d))
        self._check(self._lib.smcb_kalman_mv_batch_loglik(self._h, int(d), _ptr(b), _ptr(act), M, _ptr(y), y.size,
                                                          int(bool(matched_init)), _ptr(ll), _ptr(x), _ptr(s)))
        return ll, x, s

    def batch(self, kind, M, N):
        return Batch(self, kind, M, N)

    # ---- multi-GPU (one process per GPU)
    def comm_init(self, rank, nranks, unique_id):
        """join the NCCL communicator named by unique_id (collective: every rank calls it); samplers created from
        this context afterwards shard their θ-particles over the ranks"""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._check(self._lib.smcb_comm_init(self._h, int(rank), int(nranks), buf))

    def comm_rank(self):
        r, n = C.c_int(), C.c_int()
        self._check(self._lib.smcb_comm_rank(self._h, C.byref(r), C.byref(n)))
        return r.value, n.value

    def comm_all_gather(self, local):
        local = np.ascontiguousarray(local, np.float64)
        _, n = self.comm_rank()
        out = np.empty((n * local.shape[0],) + local.shape[1:])
        self._check(self._lib.smcb_comm_all_gather(self._h, _ptr(local), local.size, _ptr(out)))
        return out

    def comm_destroy(self):
        self._check(self._lib.smcb_comm_destroy(self._h))

    def sampler(self, config, theta0):
        return Sampler(self, config, theta0)


class Sampler:
    """smcb_sampler: a device-resident SMC² / density-tempered sampler (csrc/smcb_sampler.cu)"""

    def __init__(self, ctx, config, theta0):
        self.ctx, self._lib = ctx, ctx._lib
        self.cfg = config
        self.M, self.d = int(config.M), int(config.d_theta)
        θ0 = np.ascontiguousarray(theta0, np.float64).reshape(self.M, self.d)
        self._h = _c_sampler()
        ctx._check(self._lib.smcb_sampler_create(ctx._h, C.byref(config), _ptr(θ0), C.byref(self._h)))
        self._rank, self._world = ctx.comm_rank()

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self._lib.smcb_sampler_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_data(self, y):
        y = np.ascontiguousarray(y, np.float64)
        self.ctx._check(self._lib.smcb_sampler_set_data(self._h, _ptr(y), y.size))

    def smc2_init(self):
        self.ctx._check(self._lib.smcb_sampler_smc2_init(self._h))

    def smc2_step(self, t):
        ess, rj = C.c_double(), C.c_int()
        self.ctx._check(self._lib.smcb_sampler_smc2_step(self._h, int(t), C.byref(ess), C.byref(rj)))
        return ess.value, bool(rj.value)

    def density_tempered(self, cap=4096):
        """[(ξ, ess, acceptance ratio or -1)] of every tempering stage"""
        sched = np.zeros((cap, 3))
        n = C.c_int()
        self.ctx._check(self._lib.smcb_sampler_density_tempered(self._h, _ptr(sched), cap, C.byref(n)))
        return [(float(a), float(b), float(c)) for a, b, c in sched[: min(n.value, cap)]]

    def get(self, want_theta=True, want_omega=True, want_logZ=True):
        θ = np.empty((self.M, self.d)) if want_theta else None
        ω = np.empty(self.M) if want_omega else None
        z = np.empty(self.M) if want_logZ else None
        ess, ar, N = C.c_double(), C.c_double(), C.c_int64()
        self.ctx._check(self._lib.smcb_sampler_get(self._h, _ptr(θ), _ptr(ω), _ptr(z), C.byref(ess), C.byref(ar), C.byref(N)))
        return θ, ω, z, ess.value, ar.value, N.value

    def clouds(self):
        """a Batch view of this rank's live clouds (owned by the sampler)"""
        h = _c_batch()
        self.ctx._check(self._lib.smcb_sampler_clouds(self._h, C.byref(h)))
        _, _, _, _, _, N = self.get(False, False, False)
        return Batch._view(self.ctx, int(self.cfg.kind), self.M // self._world, N, h)

    def set_profiling(self, on=True):
        self.ctx._check(self._lib.smcb_sampler_set_profiling(self._h, int(bool(on))))

    def stats(self):
        ms = (C.c_double * 8)()
        n = (C.c_int64 * 8)()
        self.ctx._check(self._lib.smcb_sampler_stats(self._h, ms, n))
        keys_ms = ("filter_ms", "allgather_ms", "exchange_ms", "theta_ms", "span_ms")
        keys_n = ("sweeps", "steps", "rejuvenations", "clouds_moved", "particle_updates", "syncs", "launches", "theta_resamples")
        out = {k: float(ms[i]) for i, k in enumerate(keys_ms)}
        out.update({k: int(n[i]) for i, k in enumerate(keys_n)})
        return out


class Batch:
    """M filters of N particles (smcb_batch): the particle-of-filters of SMC² / PMMH sweeps."""

    def __init__(self, ctx, kind, M, N):
        self.ctx = ctx
        self._lib = ctx._lib
        self.kind, self.M, self.N = int(kind), int(M), int(N)
        self.d = state_dim(self.kind)
        self._h = _c_batch()
        self._owned = True
        ctx._check(self._lib.smcb_batch_create(ctx._h, self.kind, self.M, self.N, C.byref(self._h)))

    @classmethod
    def _view(cls, ctx, kind, M, N, handle):
        b = cls.__new__(cls)
        b.ctx, b._lib = ctx, ctx._lib
        b.kind, b.M, b.N = int(kind), int(M), int(N)
        b.d = state_dim(b.kind)
        b._h, b._owned = handle, False
        return b

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None) and getattr(self, "_owned", True):
            self._lib.smcb_batch_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _params(self, params):
        if params is None:
            return None
        p = params8(params).reshape(-1, PARAM_STRIDE)
        if p.shape[0] != self.M:
            raise ValueError(f"params must have {self.M} rows")
        return p

    @staticmethod
    def _mask(m):
        return None if m is None else np.ascontiguousarray(m, np.uint8)

    def init(self, params, y, stream0=0, active=None):
        p, act = self._params(params), self._mask(active)
        lm, es = np.empty(self.M), np.empty(self.M)
        self.ctx._check(self._lib.smcb_batch_init(self._h, _ptr(p), _ptr(act), float(y), int(stream0), _ptr(lm), _ptr(es)))
        return lm, es

    def _proposal(self, proposal, rows):
        q = np.ascontiguousarray(np.broadcast_to(np.asarray(proposal, np.float64), (rows, self.M, 3)))
        return q

    def step(self, y, resampler=MULTINOMIAL, params=None, proposal=None):
        """one step of every filter; proposal [M, 3] = (c0, c1, c2) makes it the guided step of docs/SPEC.md §10"""
        p = self._params(params)
        lm, es = np.empty(self.M), np.empty(self.M)
        if proposal is None:
            self.ctx._check(self._lib.smcb_batch_step(self._h, _ptr(p), float(y), int(resampler), _ptr(lm), _ptr(es)))
        else:
            q = self._proposal(proposal, 1)
            self.ctx._check(self._lib.smcb_batch_step_guided(self._h, _ptr(p), float(y), int(resampler), _ptr(q), _ptr(lm), _ptr(es)))
        return lm, es

    def log_likelihood(self, params, y, resampler=MULTINOMIAL, stream0=0, active=None, proposal=None):
        """whole series for every θ in one launch; proposal [T, M, 3] (row 0 unused) runs the guided filter"""
        p, act = self._params(params), self._mask(active)
        y = np.ascontiguousarray(y, np.float64)
        z = np.empty(self.M)
        if proposal is None:
            self.ctx._check(self._lib.smcb_batch_log_likelihood(self._h, _ptr(p), _ptr(act), _ptr(y), y.size, int(resampler),
                                                                int(stream0), _ptr(z)))
        else:
            q = self._proposal(proposal, y.size)
            self.ctx._check(self._lib.smcb_batch_log_likelihood_guided(self._h, _ptr(p), _ptr(act), _ptr(y), y.size, int(resampler),
                                                                       int(stream0), _ptr(q), _ptr(z)))
        return z

    def gather(self, parents):
        a = np.ascontiguousarray(parents, np.int32)
        if a.size != self.M:
            raise ValueError("parents must have M entries")
        self.ctx._check(self._lib.smcb_batch_gather(self._h, _ptr(a)))

    def accept(self, proposal, accept):
        m = self._mask(accept)
        self.ctx._check(self._lib.smcb_batch_accept(self._h, proposal._h, _ptr(m)))

    def fetch(self, want_x=True, want_w=True, want_logw=False):
        x = np.empty((self.M, self.d, self.N)) if want_x else None
        w = np.empty((self.M, self.N)) if want_w else None
        lw = np.empty((self.M, self.N)) if want_logw else None
        self.ctx._check(self._lib.smcb_batch_fetch(self._h, _ptr(x), _ptr(w), _ptr(lw)))
        return x, w, lw

    def weighted_mean(self):
        """[M, d] weighted state means w[m]' x[m], computed on the device."""
        out = np.empty((self.M, self.d))
        self.ctx._check(self._lib.smcb_batch_weighted_mean(self._h, _ptr(out)))
        return out

    def weighted_moments(self):
        """([M, d] means, [M, d] population variances) of every cloud under its weights, computed on the device
        (var(x, weights(w)) per θ-particle, examples/inflation_example.jl:46)."""
        mean, var = np.empty((self.M, self.d)), np.empty((self.M, self.d))
        self.ctx._check(self._lib.smcb_batch_weighted_moments(self._h, _ptr(mean), _ptr(var)))
        return mean, var

    def weighted_quantiles(self, probs, weighted=True):
        """[M, d, len(probs)] lower empirical quantiles of every cloud under its own weights (or counting every
        particle once), computed on the device (docs/SPEC.md §8; examples/inflation_example.jl:39-55)."""
        p = np.ascontiguousarray(np.atleast_1d(probs), np.float64)
        out = np.empty((self.M, self.d, p.size))
        self.ctx._check(self._lib.smcb_batch_weighted_quantiles(self._h, _ptr(p), int(p.size), int(bool(weighted)), _ptr(out)))
        return out

    def cloud_bytes(self):
        return int(self._lib.smcb_batch_cloud_bytes(self._h))

    def pack(self, slots, buf_dev_ptr):
        s = np.ascontiguousarray(slots, np.int32)
        self.ctx._check(self._lib.smcb_batch_pack(self._h, _ptr(s), s.size, C.c_void_p(int(buf_dev_ptr))))

    def unpack(self, slots, buf_dev_ptr):
        s = np.ascontiguousarray(slots, np.int32)
        self.ctx._check(self._lib.smcb_batch_unpack(self._h, _ptr(s), s.size, C.c_void_p(int(buf_dev_ptr))))

    def timing(self):
        ms, n = C.c_double(), C.c_int64()
        self.ctx._check(self._lib.smcb_batch_get_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value
