"""GPU parity of the batched filters (one CTA per θ-particle) against the CPU oracle's loop over θ
(/root/reference/src/smc_samplers.jl:112-121,223-229,289-293,325-335), through the C ABI."""
import numpy as np
import pytest

import sequential_monte_carlo_b200 as smc

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _thetas(kind, M, rng):
    if kind == smc.KIND_LG1D:
        return np.stack([rng.uniform(-0.9, 0.9, M), np.ones(M), rng.uniform(0.3, 2, M), rng.uniform(0.3, 2, M),
                         np.zeros(M), np.ones(M)], 1)
    if kind == smc.KIND_SV:
        return np.stack([rng.normal(-1, 0.5, M), rng.uniform(0.5, 0.98, M), rng.uniform(0.1, 0.6, M)], 1)
    return np.stack([rng.uniform(0.05, 0.5, M), rng.uniform(0.05, 0.5, M), rng.normal(3, 1, M), rng.uniform(0, 2, M),
                     rng.uniform(0, 2, M)], 1)


TRUE = {smc.KIND_LG1D: [0.5, 1.0, 0.9, 0.8, 0.0, 1.0], smc.KIND_SV: [-1.0, 0.9, 0.3], smc.KIND_UCSV: [0.2, 0.2, 3.0, 1.0, 1.0]}


@pytest.mark.parametrize("kind,N", [(smc.KIND_LG1D, 1024), (smc.KIND_LG1D, 1000), (smc.KIND_LG1D, 33), (smc.KIND_LG1D, 8192), (smc.KIND_SV, 2048),
                                     (smc.KIND_UCSV, 4096), (smc.KIND_UCSV, 777), (smc.KIND_UCSV, 8192), (smc.KIND_SV, 1)])
def test_batch_log_likelihood_bit_exact(ctx, oracle, kind, N):
    M, T = 12, 30 if N <= 4096 else 8
    rng = np.random.default_rng(N)
    _, y = oracle.simulate(kind, TRUE[kind], T, 1998)
    P = smc._lib.params8(_thetas(kind, M, rng))
    active = np.ones(M, np.uint8)
    active[[3, 7]] = 0
    for resampler in (smc.MULTINOMIAL, smc.STRATIFIED, smc.SYSTEMATIC):
        seed, epoch, stream0 = 42, 17 + resampler, 100
        zo, xo, lwo = oracle.batch_log_likelihood(kind, P, active, N, y, resampler, seed, epoch, stream0)
        b = ctx.batch(kind, M, N)
        ctx.set_rng(seed, epoch)
        z = b.log_likelihood(P, y, resampler, stream0, active)
        x, w, lw = b.fetch(want_logw=True)
        on = active.astype(bool)
        assert np.all(np.isneginf(z[~on])) and np.all(np.isneginf(zo[~on]))
        np.testing.assert_allclose(z[on], zo[on], rtol=RTOL, atol=0)
        np.testing.assert_array_equal(x[on], xo[on])          # bit-exact clouds => bit-exact ancestors at every step
        np.testing.assert_array_equal(lw[on], lwo[on])
        for m in np.flatnonzero(on):
            _, wo, _ = oracle.normalize(lwo[m])
            np.testing.assert_allclose(w[m], wo, rtol=RTOL, atol=0)
        b.close()


def test_batch_matches_single_filter(ctx, oracle):
    """the same (seed, epoch, stream) gives the same cloud from the grid-wide and the CTA-resident kernels"""
    kind, N, T, M = smc.KIND_LG1D, 4096, 25, 4
    _, y = oracle.simulate(kind, TRUE[kind], T, 7)
    P = smc._lib.params8(_thetas(kind, M, np.random.default_rng(3)))
    b = ctx.batch(kind, M, N)
    ctx.set_rng(5, 9)
    z = b.log_likelihood(P, y, smc.SYSTEMATIC, stream0=50)
    xb, wb, _ = b.fetch()
    for m in range(M):
        ctx.set_rng(5, 9)
        zs = ctx.log_likelihood(kind, P[m], N, y, smc.SYSTEMATIC, stream=50 + m)
        x, w, _ = ctx.fetch_state()
        np.testing.assert_array_equal(x, xb[m])
        np.testing.assert_allclose(w, wb[m], rtol=RTOL)
        assert abs(zs - z[m]) <= RTOL * abs(zs)
    b.close()


@pytest.mark.parametrize("kind", [smc.KIND_LG1D, smc.KIND_UCSV])
def test_batch_init_step_gather_accept(ctx, oracle, kind):
    """smc² / smc²! skeleton: init at y1, steps, θ-resample gather, accept from a proposal batch."""
    M, N, T = 10, 1024, 12
    rng = np.random.default_rng(5)
    _, y = oracle.simulate(kind, TRUE[kind], T, 3)
    P = smc._lib.params8(_thetas(kind, M, rng))
    seed, epoch, s0 = 77, 2, 1000
    cur = ctx.batch(kind, M, N)
    ctx.set_rng(seed, epoch)
    lm, es = cur.init(P, y[0], stream0=s0)
    xo, lwo = [], []
    for m in range(M):
        x_, lw_ = oracle.bootstrap_init(kind, P[m], N, y[0], seed, epoch, s0 + m)
        xo.append(x_)
        lwo.append(lw_)
        lmo, _, eso = oracle.normalize(lw_)
        assert abs(lm[m] - lmo) <= RTOL * abs(lmo) and abs(es[m] - eso) <= RTOL * eso

    def ostep(t):
        for m in range(M):
            oracle.bootstrap_step(kind, P[m], xo[m], lwo[m], y[t], t, oracle.MULTINOMIAL, seed, epoch, s0 + m)

    for t in range(1, 5):
        lm, es = cur.step(y[t], smc.MULTINOMIAL)
        ostep(t)
        for m in range(M):
            lmo, _, eso = oracle.normalize(lwo[m])
            assert abs(lm[m] - lmo) <= RTOL * abs(lmo) and abs(es[m] - eso) <= RTOL * eso
    # θ-resample: deep copies, duplicates then evolve independently (own Philox stream)
    parents = np.array([0, 0, 0, 3, 3, 5, 9, 9, 9, 9], np.int32)
    cur.gather(parents)
    xo = [xo[a].copy() for a in parents]
    lwo = [lwo[a].copy() for a in parents]
    # a proposal sweep over y[0:5] with new θ; accept some
    P2 = smc._lib.params8(_thetas(kind, M, rng))
    prop = ctx.batch(kind, M, N)
    ctx.set_rng(seed, epoch + 1)
    act = np.ones(M, np.uint8)
    act[4] = 0
    zp = prop.log_likelihood(P2, y[:5], smc.MULTINOMIAL, stream0=s0, active=act)
    zo, xpo, lwpo = oracle.batch_log_likelihood(kind, P2, act, N, y[:5], oracle.MULTINOMIAL, seed, epoch + 1, s0)
    np.testing.assert_allclose(zp[act > 0], zo[act > 0], rtol=RTOL)
    accept = np.array([1, 0, 1, 0, 0, 1, 0, 0, 1, 0], np.uint8)
    cur.accept(prop, accept)
    for m in np.flatnonzero(accept):
        xo[m], lwo[m], P[m] = xpo[m].copy(), lwpo[m].copy(), P2[m]
    x, _, lw = cur.fetch(want_w=False, want_logw=True)
    np.testing.assert_array_equal(x, np.stack(xo))
    np.testing.assert_array_equal(lw, np.stack(lwo))
    for t in range(5, T):
        lm, es = cur.step(y[t], smc.MULTINOMIAL, params=P)
        ostep(t)
    x, _, lw = cur.fetch(want_w=False, want_logw=True)
    np.testing.assert_array_equal(x, np.stack(xo))
    np.testing.assert_array_equal(lw, np.stack(lwo))
    assert not np.array_equal(x[0], x[1])  # duplicated parents diverged
    cur.close()
    prop.close()


def test_batch_pack_unpack_roundtrip(ctx, oracle):
    import torch
    kind, M, N = smc.KIND_UCSV, 6, 500
    _, y = oracle.simulate(kind, TRUE[kind], 5, 3)
    P = smc._lib.params8(_thetas(kind, M, np.random.default_rng(8)))
    a, b = ctx.batch(kind, M, N), ctx.batch(kind, M, N)
    ctx.set_rng(1, 1)
    a.log_likelihood(P, y, smc.SYSTEMATIC)
    ctx.set_rng(2, 2)
    b.log_likelihood(P, y, smc.SYSTEMATIC)
    xa, wa, lwa = a.fetch(want_logw=True)
    xb, wb, lwb = b.fetch(want_logw=True)
    buf = torch.empty(2 * a.cloud_bytes(), dtype=torch.uint8, device="cuda:0")
    a.pack([1, 4], buf.data_ptr())
    torch.cuda.synchronize()
    b.unpack([0, 5], buf.data_ptr())
    x2, w2, lw2 = b.fetch(want_logw=True)
    np.testing.assert_array_equal(x2[0], xa[1])
    np.testing.assert_array_equal(x2[5], xa[4])
    np.testing.assert_array_equal(w2[5], wa[4])
    np.testing.assert_array_equal(x2[1:5], xb[1:5])
    np.testing.assert_array_equal(lw2[0], lwa[1])
    a.close()
    b.close()


def test_batch_errors(ctx):
    with pytest.raises(smc.SMCBError) as e:
        ctx.batch(smc.KIND_LG1D, 4, 1 << 15)
    assert e.value.code == -5
    b = ctx.batch(smc.KIND_LG1D, 4, 64)
    with pytest.raises(smc.SMCBError):
        b.step(0.0)
    with pytest.raises(smc.SMCBError):
        b.gather([0, 1, 2, 9])
    b.close()


def test_kalman_batch(ctx, oracle):
    rng = np.random.default_rng(2)
    M, T = 37, 60
    P = smc._lib.params8(_thetas(smc.KIND_LG1D, M, rng))
    _, y = oracle.simulate(smc.KIND_LG1D, TRUE[smc.KIND_LG1D], T, 4)
    act = np.ones(M, np.uint8)
    act[5] = 0
    for matched in (False, True):
        ll, x, s = ctx.kalman_loglik(P, y, matched, act)
        for m in range(M):
            if not act[m]:
                assert np.isneginf(ll[m])
                continue
            xo, so, lo = oracle.kalman_loglik(P[m], y, matched)
            assert abs(ll[m] - lo) <= 1e-12 * abs(lo)
            assert abs(x[m] - xo) <= 1e-12 * max(1.0, abs(xo)) and abs(s[m] - so) <= 1e-12 * so
    x0, s0 = np.zeros(M), np.ones(M)
    x1, s1, l1 = ctx.kalman_step(P, x0, s0, y[0])
    for m in range(M):
        xo, so, lo = oracle.kalman_step(P[m], 0.0, 1.0, y[0])
        assert abs(x1[m] - xo) <= 1e-13 * max(1.0, abs(xo)) and abs(l1[m] - lo) <= 1e-13 * abs(lo)


def test_batch_weighted_quantiles(ctx, oracle):
    """smcb_batch_weighted_quantiles: [M, d, np] per-cloud quantiles (SPEC §8) for a 3-component model, ragged N,
    against the oracle cloud by cloud; argument errors."""
    kind, M, N, T = smc.KIND_UCSV, 6, 777, 9
    P = np.tile(smc._lib.params8([0.2, 0.2, 3.0, 1.0, 1.0]), (M, 1))
    P[:, 0] = np.linspace(0.1, 0.6, M)
    _, y = oracle.simulate(kind, [0.2, 0.2, 3.0, 1.0, 1.0], T, 5)
    b = ctx.batch(kind, M, N)
    ctx.set_rng(9, 4)
    b.log_likelihood(P, y, smc.SYSTEMATIC, stream0=2)
    _, xo, lwo = oracle.batch_log_likelihood(kind, P, None, N, y, smc.SYSTEMATIC, 9, 4, 2)
    ps = [0.0, 0.05, 0.5, 0.95, 1.0]
    for weighted in (True, False):
        q = b.weighted_quantiles(ps, weighted=weighted)
        assert q.shape == (M, 3, len(ps))
        for m in range(M):
            np.testing.assert_array_equal(q[m], oracle.weighted_summary(xo[m], lwo[m], ps, weighted=weighted)[2])
    with pytest.raises(smc.SMCBError):
        b.weighted_quantiles([1.5])
    with pytest.raises(smc.SMCBError):
        b.weighted_quantiles(np.linspace(0, 1, 17))
    b.close()


@pytest.mark.parametrize("kind,N,M", [(smc.KIND_LG1D, 1024, 333), (smc.KIND_UCSV, 4096, 170), (smc.KIND_SV, 2048, 301), (smc.KIND_UCSV, 640, 200)])
def test_dynamic_chunk_scheduling_is_bit_identical(ctx, oracle, kind, N, M, monkeypatch):
    """the persistent-grid launch that claims (chunk, θ) units dynamically (smcb_batch.cu, DYN kernels) carries the cloud, the
    statistics and the running Σ logμ through global memory between chunks: logZ, states and log-weights have to be the bits of
    the one-CTA-per-θ launch — forced chunk lengths (ragged last chunk, chunk = 1), the automatic choice, some θ inactive,
    all three resamplers, and the oracle on a few θ"""
    T = 29
    rng = np.random.default_rng(N + M)
    _, y = oracle.simulate(kind, TRUE[kind], T, 1998)
    P = smc._lib.params8(_thetas(kind, M, rng))
    active = np.ones(M, np.uint8)
    active[rng.integers(0, M, 9)] = 0
    b = ctx.batch(kind, M, N)
    for resampler in (smc.MULTINOMIAL, smc.STRATIFIED, smc.SYSTEMATIC):
        out = {}
        for chunk in ("0", "1", "5", "8", None):
            if chunk is None:
                monkeypatch.delenv("SMCB_BATCH_CHUNK", raising=False)
            else:
                monkeypatch.setenv("SMCB_BATCH_CHUNK", chunk)
            ctx.set_rng(9, 20 + resampler)
            z = b.log_likelihood(P, y, resampler, 7, active)
            x, _, lw = b.fetch(want_w=False, want_logw=True)
            out[chunk] = (z.copy(), x.copy(), lw.copy())
        on = active.astype(bool)
        for chunk in ("1", "5", "8", None):
            for ref, got in zip(out["0"], out[chunk]):
                np.testing.assert_array_equal(got[on], ref[on])
            assert np.all(np.isneginf(out[chunk][0][~on]))
        sel = np.flatnonzero(on)[[0, len(np.flatnonzero(on)) // 2, -1]]
        for m in sel:                                        # (the oracle numbers its streams from stream0 by position)
            zo1, xo1, lwo1 = oracle.batch_log_likelihood(kind, P[m:m + 1], None, N, y, resampler, 9, 20 + resampler, 7 + int(m))
            np.testing.assert_array_equal(out["5"][1][m], xo1[0])
            np.testing.assert_array_equal(out["5"][2][m], lwo1[0])
            assert abs(out["5"][0][m] - zo1[0]) <= RTOL * abs(zo1[0])
    monkeypatch.delenv("SMCB_BATCH_CHUNK", raising=False)
    # the stepping calls after a chunked sweep continue from its stored state
    ctx.set_rng(9, 31)
    monkeypatch.setenv("SMCB_BATCH_CHUNK", "4")
    b.log_likelihood(P, y[:20], smc.SYSTEMATIC, 7)
    monkeypatch.delenv("SMCB_BATCH_CHUNK", raising=False)
    lm1, _ = b.step(float(y[20]), smc.SYSTEMATIC)
    x1, _, _ = b.fetch(want_w=False)
    ctx.set_rng(9, 31)
    monkeypatch.setenv("SMCB_BATCH_CHUNK", "0")
    b.log_likelihood(P, y[:20], smc.SYSTEMATIC, 7)
    lm0, _ = b.step(float(y[20]), smc.SYSTEMATIC)
    x0, _, _ = b.fetch(want_w=False)
    np.testing.assert_array_equal(x1, x0)
    np.testing.assert_array_equal(lm1, lm0)
    b.close()
