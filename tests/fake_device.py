"""TEST INFRASTRUCTURE: an oracle-backed stand-in for the device objects of sequential_monte_carlo_b200._lib
(Context, Batch) so that the PRODUCT's host-side sampler logic (smc_samplers.py: sharding over θ, the replicated
control flow, the point-to-point cloud exchange) can run on CPU ranks under gloo and be compared with the
single-process oracle sampler (oracle/samplers.py).  Mirrors the semantics of csrc/smcb_batch.cu:

  * a batch has ONE Philox identity (seed, epoch) — taken from the context when init / log_likelihood runs, the
    context then moves to the next epoch — and slot m draws from stream (stream0 + m) at its own time index t;
  * gather / accept / pack / unpack move whole clouds (states + unnormalised log-weights).

Nothing here is product code; the product never imports it (it needs the CUDA library, there is no CPU fallback).
"""
import ctypes

import numpy as np

from oracle import oracle as o


class FakeBatch:
    def __init__(self, ctx, kind, M, N):
        self.ctx, self.kind, self.M, self.N = ctx, int(kind), int(M), int(N)
        self.d = o.state_dim(self.kind)
        self.x = np.zeros((self.M, self.d, self.N))
        self.lw = np.zeros((self.M, self.N))
        self.P = None
        self.seed = self.epoch = self.stream0 = self.t = 0
        self.live = False

    # -- filters
    def _take_identity(self, stream0):
        self.seed, self.epoch, self.stream0 = self.ctx.seed, self.ctx.epoch, int(stream0)
        self.ctx.epoch += 1

    def init(self, params, y, stream0=0, active=None):
        self.P = np.array(params, np.float64)
        self._take_identity(stream0)
        lm, es = np.empty(self.M), np.empty(self.M)
        for m in range(self.M):
            self.x[m], self.lw[m] = o.bootstrap_init(self.kind, self.P[m], self.N, float(y), self.seed, self.epoch, self.stream0 + m)
            lm[m], _, es[m] = o.normalize(self.lw[m])
        self.t, self.live = 0, True
        return lm, es

    def step(self, y, resampler=0, params=None, proposal=None):
        assert self.live
        if params is not None:
            self.P = np.array(params, np.float64)
        self.t += 1
        lm, es = np.empty(self.M), np.empty(self.M)
        for m in range(self.M):
            if proposal is None:
                o.bootstrap_step(self.kind, self.P[m], self.x[m], self.lw[m], float(y), self.t, resampler, self.seed, self.epoch,
                                 self.stream0 + m)
            else:
                o.guided_step(self.kind, self.P[m], self.x[m], self.lw[m], float(y), self.t, resampler, np.asarray(proposal)[m],
                              self.seed, self.epoch, self.stream0 + m)
            lm[m], _, es[m] = o.normalize(self.lw[m])
        return lm, es

    def log_likelihood(self, params, y, resampler=0, stream0=0, active=None, proposal=None):
        self.P = np.array(params, np.float64)
        self._take_identity(stream0)
        y = np.ascontiguousarray(y, np.float64)
        if proposal is None:
            z, x, lw = o.batch_log_likelihood(self.kind, self.P, active, self.N, y, resampler, self.seed, self.epoch, self.stream0)
        else:
            z, x, lw = o.batch_guided_log_likelihood(self.kind, self.P, active, self.N, y, resampler, proposal, self.seed, self.epoch,
                                                     self.stream0)
        on = np.ones(self.M, bool) if active is None else np.asarray(active, bool)
        self.x[on], self.lw[on] = x[on], lw[on]
        self.t, self.live = y.size - 1, True
        return z

    # -- whole clouds
    def gather(self, parents):
        a = np.asarray(parents, np.int64)
        self.x, self.lw = self.x[a].copy(), self.lw[a].copy()

    def accept(self, proposal, accept):
        m = np.asarray(accept, bool)
        self.x[m], self.lw[m] = proposal.x[m], proposal.lw[m]
        if not self.live:
            self.seed, self.epoch, self.stream0, self.t, self.live = proposal.seed, proposal.epoch, proposal.stream0, proposal.t, True

    def cloud_bytes(self):
        return 8 * (self.d + 1) * self.N

    def _rows(self, slots):
        return np.concatenate([self.x[slots].reshape(len(slots), -1), self.lw[slots]], axis=1)

    def pack(self, slots, ptr):
        rows = np.ascontiguousarray(self._rows(np.asarray(slots, np.int64)))
        ctypes.memmove(int(ptr), rows.ctypes.data, rows.nbytes)

    def unpack(self, slots, ptr):
        slots = np.asarray(slots, np.int64)
        rows = np.empty((len(slots), (self.d + 1) * self.N))
        ctypes.memmove(rows.ctypes.data, int(ptr), rows.nbytes)
        self.x[slots] = rows[:, : self.d * self.N].reshape(len(slots), self.d, self.N)
        self.lw[slots] = rows[:, self.d * self.N:]

    def fetch(self, want_x=True, want_w=True, want_logw=False):
        w = np.stack([o.normalize(l)[1] for l in self.lw]) if want_w else None
        return (self.x.copy() if want_x else None), w, (self.lw.copy() if want_logw else None)

    # -- per-cloud summaries (SPEC §8)
    def weighted_mean(self):
        return self.weighted_moments()[0]

    def weighted_moments(self):
        mean, var = np.empty((self.M, self.d)), np.empty((self.M, self.d))
        for m in range(self.M):
            w = o.normalize(self.lw[m])[1]
            mean[m] = self.x[m] @ w
            var[m] = ((self.x[m] - mean[m][:, None]) ** 2) @ w
        return mean, var

    def weighted_quantiles(self, probs, weighted=True):
        return np.stack([o.weighted_summary(self.x[m], self.lw[m], probs, weighted=weighted)[2] for m in range(self.M)])

    def timing(self):
        return 0.0, 0

    def close(self):
        pass


class FakeContext:
    """the calls smc_samplers.py makes on a _lib.Context"""

    def __init__(self, seed=0):
        self.seed, self.epoch = int(seed), 0

    def set_rng(self, seed, epoch=0):
        self.seed, self.epoch = int(seed), int(epoch)

    def batch(self, kind, M, N):
        return FakeBatch(self, kind, M, N)

    def normalize(self, logw, want_w=True):
        return o.normalize(np.asarray(logw, np.float64))

    def resample(self, w, resampler=0, stream=0, t=0, purpose=3):
        return o.resample_w(np.asarray(w, np.float64), resampler, self.seed, self.epoch, stream, t, purpose=purpose)

    # -- Kalman entry points (ibis.py)
    def kalman_step(self, params, x, sigma, y):
        P = np.asarray(params, np.float64).reshape(-1, 8)
        M = P.shape[0]
        x = np.broadcast_to(np.asarray(x, np.float64), (M,)).copy()
        s = np.broadcast_to(np.asarray(sigma, np.float64), (M,)).copy()
        ll = np.empty(M)
        for m in range(M):
            x[m], s[m], ll[m] = o.kalman_step(P[m], x[m], s[m], float(y))
        return x, s, ll

    def kalman_loglik(self, params, y, matched_init=False, active=None):
        P = np.asarray(params, np.float64).reshape(-1, 8)
        M = P.shape[0]
        ll, x, s = np.full(M, -np.inf), np.zeros(M), np.zeros(M)
        for m in range(M):
            if active is None or active[m]:
                x[m], s[m], ll[m] = o.kalman_loglik(P[m], y, matched_init)
        return ll, x, s

    def kalman_mv_step(self, d, models, x, sigma, y):
        B = np.asarray(models, np.float64).reshape(-1, 3 * d * d + 2 * d + 1)
        M = B.shape[0]
        x = np.broadcast_to(np.asarray(x, np.float64), (M, d)).copy()
        s = np.broadcast_to(np.asarray(sigma, np.float64), (M, d, d)).copy()
        ll = np.empty(M)
        for m in range(M):
            x[m], s[m], ll[m] = o.kalman_mv_step(d, B[m], x[m], s[m], float(y))
        return x, s, ll

    def kalman_mv_loglik(self, d, models, y, matched_init=False, active=None):
        B = np.asarray(models, np.float64).reshape(-1, 3 * d * d + 2 * d + 1)
        M = B.shape[0]
        ll, x, s = np.full(M, -np.inf), np.zeros((M, d)), np.zeros((M, d, d))
        for m in range(M):
            if active is None or active[m]:
                x[m], s[m], ll[m] = o.kalman_mv_loglik(d, B[m], y, matched_init)
        return ll, x, s

    # -- the single filter (particles.py): bootstrap_filter / bootstrap_filter! / log_likelihood / guided steps / summaries
    device = -1

    def _epoch_for_sweep(self):
        e = self.epoch
        self.epoch += 1
        return e

    def bootstrap_init(self, kind, params, N, y, stream=0):
        self._kind, self._N, self._T, self._P = int(kind), int(N), 1, np.asarray(params, np.float64)
        self._id = (self.seed, self._epoch_for_sweep(), int(stream))
        self._x, self._lw = o.bootstrap_init(kind, self._P, N, float(y), *self._id)
        self._t = 0
        lm, _, es = o.normalize(self._lw)
        return lm, es

    def bootstrap_step(self, y, resampler=0, params=None):
        if params is not None:
            self._P = np.asarray(params, np.float64)
        self._t += 1
        self._T += 1
        o.bootstrap_step(self._kind, self._P, self._x, self._lw, float(y), self._t, resampler, *self._id)
        lm, _, es = o.normalize(self._lw)
        return lm, es

    def guided_step(self, y, proposal, resampler=2, params=None):
        if params is not None:
            self._P = np.asarray(params, np.float64)
        self._t += 1
        self._T += 1
        o.guided_step(self._kind, self._P, self._x, self._lw, float(y), self._t, resampler, proposal, *self._id)
        lm, _, es = o.normalize(self._lw)
        return lm, es

    def log_likelihood(self, kind, params, N, y, resampler=0, stream=0, per_step=False):
        y = np.ascontiguousarray(y, np.float64)
        self._kind, self._N, self._T, self._P = int(kind), int(N), y.size, np.asarray(params, np.float64)
        self._id = (self.seed, self._epoch_for_sweep(), int(stream))
        r = o.log_likelihood(kind, self._P, N, y, resampler, *self._id)
        self._x, self._lw, self._t = r["x"], r["logw"], y.size - 1
        return (r["logZ"], r["logmu"], r["ess"]) if per_step else r["logZ"]

    def fetch_state(self, want_x=True, want_w=True, want_logw=False):
        return (self._x.copy() if want_x else None), (o.normalize(self._lw)[1] if want_w else None), (self._lw.copy() if want_logw else None)

    def summary(self, probs=(), weighted=True):
        return o.weighted_summary(self._x, self._lw, np.asarray(probs, np.float64), weighted=weighted)
