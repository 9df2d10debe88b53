"""GPU parity of the single-filter path (bootstrap_filter / bootstrap_filter! / log_likelihood,
/root/reference/src/particles.jl:87-147) against the CPU oracle, through the C ABI.

Bar: ancestors, states and log-weights bit-exact; logμ / ess / logZ / w within rel 1e-10 (fp64).
"""
import numpy as np
import pytest

import sequential_monte_carlo_b200 as smc

pytestmark = pytest.mark.gpu

RTOL = 1e-10  # fp64 tolerance of the north star for logZ, weights, quantiles

MODELS = {
    smc.KIND_LG1D: [0.5, 1.0, 0.9, 0.8, 0.0, 1.0],       # lg_mod([0.5,0.9,0.8])  README.md:12-22
    smc.KIND_SV: [-1.0, 0.9, 0.3],
    smc.KIND_UCSV: [0.2, 0.2, 3.0, 1.0, 1.0],            # examples/inflation_example.jl:229-239 shape
}


def _data(oracle, kind, T, seed=1998):
    _, y = oracle.simulate(kind, MODELS[kind], T, seed)
    return y


def _compare_run(ctx, oracle, kind, N, T, resampler, seed=7, epoch=3, stream=0):
    y = _data(oracle, kind, T)
    ref = oracle.log_likelihood(kind, MODELS[kind], N, y, resampler, seed, epoch, stream, want_anc=True)
    ctx.set_rng(seed, epoch)
    ctx.record_ancestors(True)
    logZ, logmu, ess = ctx.log_likelihood(kind, MODELS[kind], N, y, resampler, stream, per_step=True)
    x, w, logw = ctx.fetch_state(want_logw=True)
    anc = ctx.fetch_ancestors(T - 1)
    ctx.record_ancestors(False)
    assert ctx.next_epoch() == epoch + 1
    if T > 1:
        assert anc.shape == (T - 1, N)
        np.testing.assert_array_equal(anc, ref["anc"][1:])          # bit-exact ancestors, every step
    np.testing.assert_array_equal(x, ref["x"])                      # bit-exact states
    np.testing.assert_array_equal(logw, ref["logw"])                # bit-exact log-weights
    np.testing.assert_allclose(logmu, ref["logmu"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(ess, ref["ess"], rtol=RTOL, atol=0)
    assert abs(logZ - ref["logZ"]) <= RTOL * abs(ref["logZ"])
    _, wref, _ = oracle.normalize(ref["logw"])
    np.testing.assert_allclose(w, wref, rtol=RTOL, atol=0)
    # quantiles of the weighted cloud (what README.md:39-53 computes from x, w)
    for k in range(x.shape[0]):
        o = np.argsort(x[k], kind="stable")
        cg, cr = np.cumsum(w[o]), np.cumsum(wref[o])
        for p in (0.05, 0.5, 0.95):
            assert x[k][o][np.searchsorted(cg, p)] == ref["x"][k][o][np.searchsorted(cr, p)]
    return logZ


@pytest.mark.parametrize("resampler", [smc.MULTINOMIAL, smc.STRATIFIED, smc.SYSTEMATIC])
@pytest.mark.parametrize("kind", [smc.KIND_LG1D, smc.KIND_SV, smc.KIND_UCSV])
def test_config1_bit_exact(ctx, oracle, kind, resampler):
    """BASELINE config 1 shape: N=1024, T=100."""
    _compare_run(ctx, oracle, kind, 1024, 100, resampler)


@pytest.mark.parametrize("N", [1, 2, 3, 31, 1001, 2048, 2049, 5000, 70001])
def test_ragged_sizes(ctx, oracle, N):
    for resampler in (smc.MULTINOMIAL, smc.SYSTEMATIC):
        _compare_run(ctx, oracle, smc.KIND_LG1D, N, 12, resampler, seed=11, epoch=N % 97)
    _compare_run(ctx, oracle, smc.KIND_UCSV, N, 6, smc.STRATIFIED, seed=5, epoch=1, stream=9)


def test_single_observation(ctx, oracle):
    _compare_run(ctx, oracle, smc.KIND_LG1D, 777, 1, smc.SYSTEMATIC)


def test_medium_n_many_tiles(ctx, oracle):
    """2^18 particles: 128 scan tiles, 256 propagate CTAs — exercises the look-back chain."""
    _compare_run(ctx, oracle, smc.KIND_LG1D, 1 << 18, 8, smc.SYSTEMATIC)
    _compare_run(ctx, oracle, smc.KIND_SV, 1 << 17, 5, smc.MULTINOMIAL)


def test_above_2_pow_24_multi_trip_tiles(ctx, oracle):
    """N > 2^24: the sum kernel's CTAs make several trips per tile (tile index capped at 8192 tiles)."""
    _compare_run(ctx, oracle, smc.KIND_LG1D, (1 << 24) + 4097, 3, smc.SYSTEMATIC)


@pytest.mark.parametrize("kind,N,T", [(smc.KIND_LG1D, 8193, 6), (smc.KIND_SV, 12289 + 4096 * 3 + 17, 5), (smc.KIND_UCSV, (1 << 16) + 5, 4),
                                      (smc.KIND_LG1D, (1 << 20) + 3, 4)])
def test_two_level_multinomial_bit_exact(ctx, oracle, kind, N, T):
    """SPEC §5c: the reference's multinomial law for clouds above 8192 particles (counts per 4096-particle cell, in-cell
    thresholds, ascending chunks): ancestors of every step, states and log-weights bit for bit against the oracle."""
    _compare_run(ctx, oracle, kind, N, T, smc.MULTINOMIAL, seed=17, epoch=N % 89)


def test_two_level_multinomial_degenerate_weights(ctx, oracle):
    """A sharp likelihood puts nearly all the mass on a few particles: a few cells receive almost every output (many chunks
    per cell, most cells empty) — and the all-(-inf) cloud is its own ancestor."""
    kind, N, T = smc.KIND_LG1D, (1 << 17) + 99, 5
    y = _data(oracle, kind, T)
    for R in (1e-6, 1e-10):
        params = [0.5, 1.0, 0.9, R, 0.0, 1.0]
        ref = oracle.log_likelihood(kind, params, N, y, smc.MULTINOMIAL, 3, 0, 0, want_anc=True)
        ctx.set_rng(3, 0)
        ctx.record_ancestors(True)
        logZ, logmu, ess = ctx.log_likelihood(kind, params, N, y, smc.MULTINOMIAL, 0, per_step=True)
        anc = ctx.fetch_ancestors(T - 1)
        x, _, _ = ctx.fetch_state(want_w=False)
        ctx.record_ancestors(False)
        np.testing.assert_array_equal(anc, ref["anc"][1:])
        np.testing.assert_array_equal(x, ref["x"])
        np.testing.assert_allclose(logmu, ref["logmu"], rtol=RTOL)
        assert ess.min() < 500
    # the stepping API: multinomial steps between sorted ones at a size above the legacy limit, statistics after every step
    N = 50000
    ctx.set_rng(21, 1)
    ctx.bootstrap_init(kind, MODELS[kind], N, y[0], stream=2)
    xo, lwo = oracle.bootstrap_init(kind, MODELS[kind], N, y[0], 21, 1, 2)
    for t, rs in zip(range(1, T), (smc.MULTINOMIAL, smc.SYSTEMATIC, smc.MULTINOMIAL, smc.MULTINOMIAL)):
        lm, es = ctx.bootstrap_step(y[t], rs)
        oracle.bootstrap_step(kind, MODELS[kind], xo, lwo, y[t], t, rs, 21, 1, 2)
        lmo, _, eso = oracle.normalize(lwo)
        assert abs(lm - lmo) <= RTOL * abs(lmo) and abs(es - eso) <= RTOL * eso
    x, _, lw = ctx.fetch_state(want_w=False, want_logw=True)
    np.testing.assert_array_equal(x, xo)
    np.testing.assert_array_equal(lw, lwo)


def test_degenerate_weights_window_fallback(ctx, oracle):
    """A sharp likelihood (tiny R) concentrates the weight on few particles, so some CTAs see CDF
    windows wider than the staging buffer and take the global-search path."""
    kind, N, T = smc.KIND_LG1D, 1 << 16, 6
    y = _data(oracle, kind, T)
    for resampler, R in ((smc.SYSTEMATIC, 1e-6), (smc.STRATIFIED, 1e-9), (smc.SYSTEMATIC, 1e-9)):
        params = [0.5, 1.0, 0.9, R, 0.0, 1.0]
        ref = oracle.log_likelihood(kind, params, N, y, resampler, 3, 0, 0, want_anc=True)
        ctx.set_rng(3, 0)
        ctx.record_ancestors(True)
        logZ, logmu, ess = ctx.log_likelihood(kind, params, N, y, resampler, 0, per_step=True)
        anc = ctx.fetch_ancestors(T - 1)
        x, _, logw = ctx.fetch_state(want_w=False, want_logw=True)
        ctx.record_ancestors(False)
        np.testing.assert_array_equal(anc, ref["anc"][1:])
        np.testing.assert_array_equal(x, ref["x"])
        np.testing.assert_allclose(logmu, ref["logmu"], rtol=RTOL)
        assert ess.min() < 500


def test_stepping_api_matches_whole_series(ctx, oracle):
    """bootstrap_filter + bootstrap_filter! one observation at a time (README.md:33-61 usage)."""
    kind, N, T = smc.KIND_LG1D, 4096, 20
    y = _data(oracle, kind, T)
    ctx.set_rng(21, 5)
    logZ, logmu, ess = ctx.log_likelihood(kind, MODELS[kind], N, y, smc.STRATIFIED, 2, per_step=True)
    x_all, w_all, _ = ctx.fetch_state()
    ctx.set_rng(21, 5)
    lm0, es0 = ctx.bootstrap_init(kind, MODELS[kind], N, y[0], stream=2)
    lms, ess2 = [lm0], [es0]
    for t in range(1, T):
        lm, es = ctx.bootstrap_step(y[t], smc.STRATIFIED)
        lms.append(lm)
        ess2.append(es)
    x, w, _ = ctx.fetch_state()
    np.testing.assert_array_equal(x, x_all)
    np.testing.assert_array_equal(w, w_all)
    # the per-step statistics come from two kernels with different (fixed) summation orders
    np.testing.assert_allclose(np.array(lms), logmu, rtol=1e-13)
    np.testing.assert_allclose(np.array(ess2), ess, rtol=1e-13)
    # and against the oracle's step function
    xo, lwo = oracle.bootstrap_init(kind, MODELS[kind], N, y[0], 21, 5, 2)
    for t in range(1, T):
        oracle.bootstrap_step(kind, MODELS[kind], xo, lwo, y[t], t, oracle.STRATIFIED, 21, 5, 2)
    np.testing.assert_array_equal(x, xo)


def test_step_before_init_is_an_error():
    c = smc.Context(0, 1)
    with pytest.raises(smc.SMCBError) as e:
        c.bootstrap_step(0.0)
    assert e.value.code == -4
    with pytest.raises(smc.SMCBError):
        c.log_likelihood(7, [0.0], 10, [0.0])
    with pytest.raises(smc.SMCBError):
        c.log_likelihood(smc.KIND_LG1D, MODELS[smc.KIND_LG1D], 0, [0.0])
    c.close()


def test_pf_vs_kalman(ctx, oracle):
    """log_likelihood(N,y,model) vs log_likelihood(y,model) (kalman_filter.jl:55-70) on LG1D.
    The PF's target is the matched-init Kalman likelihood."""
    kind, T = smc.KIND_LG1D, 100
    y = _data(oracle, kind, T)
    _, _, kf_matched = oracle.kalman_loglik(MODELS[kind], y, matched_init=True)
    _, _, kf_ref = oracle.kalman_loglik(MODELS[kind], y, matched_init=False)
    N, reps = 16384, 24
    vals = []
    for r in range(reps):
        ctx.set_rng(100 + r, 0)
        vals.append(ctx.log_likelihood(kind, MODELS[kind], N, y, smc.MULTINOMIAL))
    vals = np.array(vals)
    lme = np.log(np.mean(np.exp(vals - vals.max()))) + vals.max()  # E[Ẑ] = Z (unbiased)
    sd = vals.std(ddof=1)
    assert sd < 0.3
    assert abs(lme - kf_matched) < 4 * sd / np.sqrt(reps) + 0.01
    assert abs(kf_ref - kf_matched) < 0.2  # reference-style (predict-first) value, reported for D1
    ll, _, _ = ctx.kalman_loglik(MODELS[kind], y, matched_init=True)
    assert abs(ll[0] - kf_matched) <= 1e-12 * abs(kf_matched)
    ll, xT, sT = ctx.kalman_loglik(MODELS[kind], y, matched_init=False)
    xo, so, _ = oracle.kalman_loglik(MODELS[kind], y, matched_init=False)
    assert abs(ll[0] - kf_ref) <= 1e-12 * abs(kf_ref)
    assert abs(xT[0] - xo) <= 1e-12 * abs(xo) and abs(sT[0] - so) <= 1e-12 * so


def test_resample_with_a_different_output_size(ctx, oracle):
    """resample(w, N) with N != length(w) (particles.jl:17-19): N ancestors drawn from length(w) weights"""
    rng = np.random.default_rng(5)
    w = rng.random(3000)
    w /= w.sum()
    for rs in (smc.MULTINOMIAL, smc.STRATIFIED, smc.SYSTEMATIC):
        for n_out in (1, 77, 3000, 10001):
            ctx.set_rng(9, 4)
            a = smc.resample(w, n_out, resampler=rs, ctx=ctx, stream=2, t=3)
            np.testing.assert_array_equal(a, oracle.resample_w(w, rs, 9, 4, 2, 3, n_out=n_out))
            assert a.shape == (n_out,) and a.min() >= 0 and a.max() < w.size


def test_stale_weights_handle(ctx, oracle):
    """bootstrap_filter! returns a NEW w every step (particles.jl:128): a handle of an earlier step keeps its own values when
    they were read in time, and refuses to hand out a later step's weights otherwise"""
    lg = smc.LinearGaussian(0.5, 1.0, 0.9, 0.8, 0.0)
    y = _data(oracle, smc.KIND_LG1D, 4)
    x, w0, _ = smc.bootstrap_filter(20000, y[0], lg, ctx=ctx)
    first = np.array(w0)
    _, w1, _ = smc.bootstrap_filter_(x, w0, y[1], lg, resampler="systematic")
    np.testing.assert_array_equal(np.asarray(w0), first)            # read in time: keeps its values
    _, w2, _ = smc.bootstrap_filter_(x, w1, y[2], lg, resampler="systematic")
    with pytest.raises(RuntimeError):
        np.asarray(w1)                                              # never read while current
    assert abs(np.asarray(w2).sum() - 1.0) < 1e-12


def test_normalize_utility(ctx, oracle):
    rng = np.random.default_rng(0)
    for n in (1, 2, 1000, 2048, 100003):
        logw = rng.normal(size=n) * 3 - 500.0
        lm, w, es = ctx.normalize(logw)
        lo, wo, eo = oracle.normalize(logw)
        assert abs(lm - lo) <= RTOL * abs(lo)
        assert abs(es - eo) <= RTOL * eo
        np.testing.assert_allclose(w, wo, rtol=RTOL)
    lm, w, es = ctx.normalize(np.full(64, -3.25))          # equal weights: logμ = logw, ess = N
    assert lm == pytest.approx(-3.25, abs=1e-14) and es == pytest.approx(64.0, rel=1e-14)
    onehot = np.full(100, -np.inf)
    onehot[17] = 2.0
    lm, w, es = ctx.normalize(onehot)                       # one-hot: ess = 1
    assert es == pytest.approx(1.0) and w[17] == 1.0 and lm == pytest.approx(2.0 - np.log(100))
    lm, w, es = ctx.normalize(np.full(8, -np.inf))          # all -Inf -> NaN like the reference
    assert np.isnan(lm)


def test_resample_utility(ctx, oracle):
    rng = np.random.default_rng(1)
    for n in (1, 5, 512, 4096, 50000):
        w = rng.random(n) ** 4
        w /= w.sum()
        for rs in (smc.MULTINOMIAL, smc.STRATIFIED, smc.SYSTEMATIC):
            ctx.set_rng(9, 4)
            a = ctx.resample(w, rs, stream=3, t=11, purpose=5)
            ao = oracle.resample_w(w, rs, 9, 4, 3, 11, purpose=5)
            np.testing.assert_array_equal(a, ao)
    # offspring counts are unbiased: E[count_j] = n w_j (multinomial), chi-square over bins
    n = 4096
    w = rng.random(n)
    w /= w.sum()
    counts = np.zeros(n)
    reps = 200
    for r in range(reps):
        ctx.set_rng(1234, r)
        counts += np.bincount(ctx.resample(w, smc.MULTINOMIAL), minlength=n)
    bins = counts.reshape(64, -1).sum(1)
    expect = (w * n * reps).reshape(64, -1).sum(1)
    chi2 = ((bins - expect) ** 2 / expect).sum()
    assert chi2 < 63 + 6 * np.sqrt(2 * 63)


def test_device_detmath_bit_exact(ctx, oracle):
    """The deterministic math of docs/SPEC.md §3 evaluated ON THE DEVICE equals the oracle bit for bit:
    exp, log, sincos2pi and the quantiser."""
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-720, 10, 50000), -rng.exponential(3.0, 50000), [0.0, -0.0, -700.0, -745.0, -np.inf]])
    np.testing.assert_array_equal(ctx.selftest_math(0, x)[0], oracle.det_exp(x))
    u = np.concatenate([rng.random(100000), 2.0 ** -rng.uniform(1, 53, 100000), [2.0 ** -53, 1 - 2.0 ** -53, 0.5, 0.685546875]])
    np.testing.assert_array_equal(ctx.selftest_math(1, u)[0], oracle.det_log(u))
    s, c = ctx.selftest_math(2, u)
    so, co = oracle.det_sincos2pi(u)
    np.testing.assert_array_equal(s, so)
    np.testing.assert_array_equal(c, co)
    S = oracle.quant_shift(1 << 24)
    q, _ = ctx.selftest_math(3, x[x <= 0], aux=S)
    np.testing.assert_array_equal(q.view(np.uint64), oracle.det_quant(x[x <= 0], S))


def test_stepping_with_mixed_resamplers_and_changing_parameters(ctx, oracle):
    """The stepping API across every resampler transition and with new model parameters handed to a
    step: the weights being resampled are those computed under the OLD parameters (the LG1D path keeps
    them implicit in x), the transition and the new weights use the NEW ones."""
    kind, N, T = smc.KIND_LG1D, 6000, 13
    y = _data(oracle, kind, T)
    seq = [smc.SYSTEMATIC, smc.SYSTEMATIC, smc.MULTINOMIAL, smc.STRATIFIED, smc.SYSTEMATIC, smc.MULTINOMIAL,
           smc.MULTINOMIAL, smc.STRATIFIED, smc.STRATIFIED, smc.SYSTEMATIC, smc.MULTINOMIAL, smc.SYSTEMATIC]
    P0, P1 = MODELS[kind], [0.7, 1.0, 0.5, 1.3, 0.0, 1.0]
    ctx.set_rng(33, 4)
    ctx.bootstrap_init(kind, P0, N, y[0], stream=1)
    xo, lwo = oracle.bootstrap_init(kind, P0, N, y[0], 33, 4, 1)
    for t in range(1, T):
        P = P1 if t >= 6 else P0
        lm, es = ctx.bootstrap_step(y[t], seq[t - 1], P)
        oracle.bootstrap_step(kind, P, xo, lwo, y[t], t, seq[t - 1], 33, 4, 1)
        lmo, _, eso = oracle.normalize(lwo)
        assert abs(lm - lmo) <= RTOL * abs(lmo) and abs(es - eso) <= RTOL * eso
        if t in (3, 6, 12):
            x, w, lw = ctx.fetch_state(want_logw=True)
            np.testing.assert_array_equal(x, xo)
            np.testing.assert_array_equal(lw, lwo)


def test_multinomial_after_sorted_steps_at_many_tiles(ctx, oracle):
    """A multinomial step that follows sorted-resampler steps must not see the look-back descriptors of an older scan of
    the same parity (they are flagged "inclusive" with stale prefixes; the sorted steps advance t without scanning).  Many
    tiles (N = 2^19 + 77: 257 scan tiles), multinomial at t = 1, 5, 9, 11 with systematic / stratified in between; states
    and log-weights bit for bit against the oracle after every multinomial step."""
    kind, N, T = smc.KIND_SV, (1 << 19) + 77, 12
    y = _data(oracle, kind, T)
    seq = [smc.MULTINOMIAL, smc.SYSTEMATIC, smc.STRATIFIED, smc.SYSTEMATIC, smc.MULTINOMIAL, smc.SYSTEMATIC, smc.SYSTEMATIC,
           smc.STRATIFIED, smc.MULTINOMIAL, smc.SYSTEMATIC, smc.MULTINOMIAL]
    for rep in range(3):            # the hazard is a race: repeat it
        ctx.set_rng(91 + rep, 2)
        ctx.bootstrap_init(kind, MODELS[kind], N, y[0], stream=0)
        xo, lwo = oracle.bootstrap_init(kind, MODELS[kind], N, y[0], 91 + rep, 2, 0)
        for t in range(1, T):
            ctx.bootstrap_step(y[t], seq[t - 1])
            oracle.bootstrap_step(kind, MODELS[kind], xo, lwo, y[t], t, seq[t - 1], 91 + rep, 2, 0)
            if seq[t - 1] == smc.MULTINOMIAL:
                x, _, lw = ctx.fetch_state(want_w=False, want_logw=True)
                np.testing.assert_array_equal(x, xo)
                np.testing.assert_array_equal(lw, lwo)


@pytest.mark.parametrize("kind", [smc.KIND_LG1D, smc.KIND_UCSV])
def test_on_device_summaries(ctx, oracle, kind):
    """docs/SPEC.md §8: weighted / plain mean, variance and quantiles of the cloud computed on the device
    (README.md:41,51; examples/inflation_example.jl:39-55) against the oracle's sort-and-cumulate restatement:
    quantiles bit-exact, moments at the fp64 tolerance."""
    N, T = 20011, 9
    y = _data(oracle, kind, T)
    probs = [0.0, 0.05, 0.25, 0.5, 0.75, 0.95, 1.0]
    for resampler in (smc.SYSTEMATIC, smc.MULTINOMIAL):
        ref = oracle.log_likelihood(kind, MODELS[kind], N, y, resampler, 17, 2, 0)
        ctx.set_rng(17, 2)
        ctx.log_likelihood(kind, MODELS[kind], N, y, resampler, 0)
        for weighted in (True, False):
            m, v, q = ctx.summary(probs, weighted=weighted)
            mo, vo, qo = oracle.weighted_summary(ref["x"], ref["logw"], probs, weighted=weighted)
            np.testing.assert_array_equal(q, qo)
            np.testing.assert_allclose(m, mo, rtol=RTOL, atol=1e-13)
            np.testing.assert_allclose(v, vo, rtol=RTOL)
        assert np.all(np.diff(q, axis=1) >= 0)
    # the stepping API: summaries after every step, the cloud never leaves the device
    ctx.set_rng(17, 3)
    ctx.bootstrap_init(kind, MODELS[kind], N, y[0], 0)
    xo, lwo = oracle.bootstrap_init(kind, MODELS[kind], N, y[0], 17, 3, 0)
    for t in range(1, 4):
        ctx.bootstrap_step(y[t], smc.SYSTEMATIC)
        oracle.bootstrap_step(kind, MODELS[kind], xo, lwo, y[t], t, oracle.SYSTEMATIC, 17, 3, 0)
        _, _, q = ctx.summary([0.25, 0.5, 0.75])
        np.testing.assert_array_equal(q, oracle.weighted_summary(xo, lwo, [0.25, 0.5, 0.75])[2])


def test_fuzz_sizes_and_weight_profiles(ctx, oracle):
    """40 seeded random configurations of the sorted-resampler step against the oracle: particle counts
    around every granularity of the kernels (128-particle chunks, 1024-particle ancestor CTAs, tiles),
    likelihoods from flat to very sharp (CDF windows from one entry to many passes), both sorted
    resamplers.  Ancestors of every step, final states and log-weights bit-exact."""
    rng = np.random.default_rng(20261018)
    kind, T = smc.KIND_LG1D, 5
    specials = [127, 128, 129, 1023, 1024, 1025, 2047, 2049, 3583, 3584, 3585, 7168, 28672, 28673, 65535, 65537, 131071]
    for trial in range(40):
        N = int(specials[trial]) if trial < len(specials) else int(rng.integers(1, 300000))
        R = float(10.0 ** rng.uniform(-8, 1))            # observation variance: 1e-8 (degenerate weights) .. 10 (flat)
        A = float(rng.uniform(-0.95, 0.95))
        params = [A, 1.0, float(10.0 ** rng.uniform(-2, 0.5)), R, 0.0, 1.0]
        resampler = smc.SYSTEMATIC if trial % 3 else smc.STRATIFIED
        _, y = oracle.simulate(kind, [A, 1.0, params[2], max(R, 0.3), 0.0, 1.0], T, 100 + trial)
        ref = oracle.log_likelihood(kind, params, N, y, resampler, 5, trial, 3, want_anc=True)
        ctx.set_rng(5, trial)
        ctx.record_ancestors(True)
        logZ = ctx.log_likelihood(kind, params, N, y, resampler, 3)
        anc = ctx.fetch_ancestors(T - 1)
        x, _, lw = ctx.fetch_state(want_w=False, want_logw=True)
        ctx.record_ancestors(False)
        tag = f"trial {trial}: N={N} R={R:.3g} resampler={resampler}"
        assert np.array_equal(anc, ref["anc"][1:]), tag
        assert np.array_equal(x, ref["x"]) and np.array_equal(lw, ref["logw"]), tag
        assert abs(logZ - ref["logZ"]) <= RTOL * abs(ref["logZ"]) or (np.isnan(logZ) and np.isnan(ref["logZ"])), tag


def test_systematic_exact_decision_path(oracle, monkeypatch):
    """anc_hist_kernel decides `first particle whose threshold reaches a CDF entry` from a double estimate and
    falls back to exact 128-bit comparisons when the estimate is within 1e-9 of an integer — a path
    random inputs almost never take.  SMCB_ANC_FORCE_EXACT makes every entry take it; the ancestors
    must not change."""
    monkeypatch.setenv("SMCB_ANC_FORCE_EXACT", "1")
    c = smc.Context(0, seed=11)
    try:
        kind, T = smc.KIND_LG1D, 4
        for trial, (N, R) in enumerate([(1000, 0.8), (1024, 1e-6), (5000, 0.05), (70001, 3.0), (262144, 0.8)]):
            params = [0.7, 1.0, 0.9, R, 0.0, 1.0]
            _, y = oracle.simulate(kind, [0.7, 1.0, 0.9, 0.8, 0.0, 1.0], T, 40 + trial)
            ref = oracle.log_likelihood(kind, params, N, y, smc.SYSTEMATIC, 11, trial, 0, want_anc=True)
            c.set_rng(11, trial)
            c.record_ancestors(True)
            c.log_likelihood(kind, params, N, y, smc.SYSTEMATIC, 0)
            anc = c.fetch_ancestors(T - 1)
            x, _, _ = c.fetch_state(want_w=False)
            c.record_ancestors(False)
            assert np.array_equal(anc, ref["anc"][1:]), (N, R)
            assert np.array_equal(x, ref["x"]), (N, R)
    finally:
        c.close()


@pytest.mark.parametrize("kind", [smc.KIND_LG1D, smc.KIND_SV, smc.KIND_UCSV])
def test_f32_state_tier(ctx, oracle, kind):
    """docs/SPEC.md §9: binary32 states, binary64 arithmetic.  Bit-exact against the oracle run in the same
    tier (ancestors, states, log-weights), every state representable in binary32, and logZ within the north
    star's fp32 tolerance (rel 1e-4) of the binary64 tier."""
    T = 16
    y = _data(oracle, kind, T)
    ctx.set_precision("f32")
    try:
        for N, resampler in ((1024, smc.SYSTEMATIC), (5001, smc.STRATIFIED), (70001, smc.SYSTEMATIC)):
            with oracle.state_f32():
                ref = oracle.log_likelihood(kind, MODELS[kind], N, y, resampler, 13, 2, 1, want_anc=True)
            ref64 = oracle.log_likelihood(kind, MODELS[kind], N, y, resampler, 13, 2, 1)
            ctx.set_rng(13, 2)
            ctx.record_ancestors(True)
            logZ, logmu, ess = ctx.log_likelihood(kind, MODELS[kind], N, y, resampler, 1, per_step=True)
            anc = ctx.fetch_ancestors(T - 1)
            x, w, logw = ctx.fetch_state(want_logw=True)
            ctx.record_ancestors(False)
            np.testing.assert_array_equal(anc, ref["anc"][1:])
            np.testing.assert_array_equal(x, ref["x"])
            np.testing.assert_array_equal(logw, ref["logw"])
            assert np.array_equal(x, x.astype(np.float32).astype(np.float64))
            np.testing.assert_allclose(logmu, ref["logmu"], rtol=RTOL, atol=0)
            np.testing.assert_allclose(ess, ref["ess"], rtol=RTOL, atol=0)
            assert abs(logZ - ref["logZ"]) <= RTOL * abs(ref["logZ"])
            # fp32 tier against the fp64 tier.  Before the first resampling the difference is pure rounding; later
            # the two runs share ancestors only until the first threshold that falls between the two CDFs, after
            # which they are two estimates of the same Z (Monte-Carlo error, not rounding error).
            assert abs(logmu[0] - ref64["logmu"][0]) <= 1e-6 * abs(ref64["logmu"][0])
            tol = 1e-4 if kind == smc.KIND_LG1D else 5e-3
            assert abs(logZ - ref64["logZ"]) <= tol * abs(ref64["logZ"])
            assert not np.array_equal(x, ref64["x"])                            # (it IS a different tier)
            # on-device summaries read the binary32 states
            m, v, q = ctx.summary([0.1, 0.5, 0.9])
            mo, vo, qo = oracle.weighted_summary(ref["x"], ref["logw"], [0.1, 0.5, 0.9])
            np.testing.assert_allclose(m, mo, rtol=RTOL)
            np.testing.assert_array_equal(q, qo)
        # stepping API in the same tier
        ctx.set_rng(13, 2)
        ctx.bootstrap_init(kind, MODELS[kind], 70001, y[0], stream=1)
        for t in range(1, T):
            ctx.bootstrap_step(y[t], smc.SYSTEMATIC)
        xs, _, _ = ctx.fetch_state(want_w=False)
        np.testing.assert_array_equal(xs, ref["x"])
        # multinomial in this tier: the two-level draw of SPEC §5c (N > 8192) shares the sorted path's kernels and works; the
        # per-particle search of small clouds does not exist in binary32 and is refused
        with oracle.state_f32():
            refm = oracle.log_likelihood(kind, MODELS[kind], 70001, y, smc.MULTINOMIAL, 13, 3, 1)
        ctx.set_rng(13, 3)
        ctx.log_likelihood(kind, MODELS[kind], 70001, y, smc.MULTINOMIAL, stream=1)
        xm, _, _ = ctx.fetch_state(want_w=False)
        np.testing.assert_array_equal(xm, refm["x"])
        with pytest.raises(smc.SMCBError):
            ctx.log_likelihood(kind, MODELS[kind], 1024, y, smc.MULTINOMIAL)
    finally:
        ctx.set_precision("f64")
    # and the binary64 tier is untouched afterwards
    _compare_run(ctx, oracle, kind, 1001, 6, smc.SYSTEMATIC)


def test_f32_arithmetic_math_bit_exact(ctx, oracle):
    """docs/SPEC.md §9b: the device's binary32 exp / log / sincos2pi / quantiser against the oracle's independent restatement,
    bit for bit (the ancestors of the tier depend on every bit of the exponential)"""
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.uniform(-90, 3, 100000), [-0.0, 0.0, -86.0, -86.000001, -1e-9, -np.inf]]).astype(np.float32).astype(np.float64)
    np.testing.assert_array_equal(ctx.selftest_math(4, x)[0], oracle.detf("exp", x))
    u = np.concatenate([(2 * rng.integers(0, 1 << 23, 100000) + 1) * 2.0 ** -24, [2.0 ** -24, 1 - 2.0 ** -24, 0.5, 0.70710677, 0.7071068]])
    np.testing.assert_array_equal(ctx.selftest_math(5, u)[0], oracle.detf("log", u))
    s, c = ctx.selftest_math(6, u)
    so, co = oracle.detf("sincos", u)
    np.testing.assert_array_equal(s, so)
    np.testing.assert_array_equal(c, co)
    for S in (37, 47, 51):
        xq = x[x <= 0]
        q, _ = ctx.selftest_math(7, xq, aux=S)
        np.testing.assert_array_equal(q.view(np.uint64), oracle.detf("quant", xq, S))


@pytest.mark.parametrize("kind", [smc.KIND_LG1D, smc.KIND_SV, smc.KIND_UCSV])
def test_f32_arithmetic_tier(ctx, oracle, kind):
    """docs/SPEC.md §9b: binary32 ARITHMETIC (float Philox normals four per block, float model arithmetic and log-weights, the
    uint64 CDF kept).  Ancestors, states and log-weights bit-exact against the oracle run in the same tier for the sorted
    resamplers and the two-level multinomial draw; logμ / ess at 1e-10 (the sums are binary64 in both)."""
    T = 12
    y = _data(oracle, kind, T)
    ctx.set_precision("f32_arith")
    try:
        for N, resampler in ((1024, smc.SYSTEMATIC), (5003, smc.STRATIFIED), (70001, smc.SYSTEMATIC), (40002, smc.MULTINOMIAL)):
            with oracle.arith_f32():
                ref = oracle.log_likelihood(kind, MODELS[kind], N, y, resampler, 13, 2, 1, want_anc=True)
            ctx.set_rng(13, 2)
            ctx.record_ancestors(True)
            logZ, logmu, ess = ctx.log_likelihood(kind, MODELS[kind], N, y, resampler, 1, per_step=True)
            anc = ctx.fetch_ancestors(T - 1)
            x, w, logw = ctx.fetch_state(want_logw=True)
            ctx.record_ancestors(False)
            np.testing.assert_array_equal(anc, ref["anc"][1:])
            np.testing.assert_array_equal(x, ref["x"])
            np.testing.assert_array_equal(logw, ref["logw"])
            assert np.array_equal(x, x.astype(np.float32).astype(np.float64)) and np.array_equal(logw, logw.astype(np.float32).astype(np.float64))
            np.testing.assert_allclose(logmu, ref["logmu"], rtol=RTOL, atol=0)
            np.testing.assert_allclose(ess, ref["ess"], rtol=RTOL, atol=0)
            with oracle.arith_f32():
                np.testing.assert_allclose(w, oracle.normalize_f32(ref["logw"])[1], rtol=RTOL, atol=0)
            if kind == smc.KIND_LG1D and N == 70001:   # summaries read binary32 states and the uint64 weights: bit-exact as in the other tiers
                m, v, q = ctx.summary([0.1, 0.5, 0.9])
                mo, vo, qo = oracle.weighted_summary(ref["x"], ref["logw"], [0.1, 0.5, 0.9])
                np.testing.assert_allclose(m, mo, rtol=1e-6)     # (the oracle's summary quantises the weights in binary64)
        # the stepping API in the tier
        ctx.set_rng(13, 2)
        ctx.bootstrap_init(kind, MODELS[kind], 70001, y[0], stream=1)
        for t in range(1, T):
            ctx.bootstrap_step(y[t], smc.SYSTEMATIC)
        with oracle.arith_f32():
            ref = oracle.log_likelihood(kind, MODELS[kind], 70001, y, smc.SYSTEMATIC, 13, 2, 1)
        np.testing.assert_array_equal(ctx.fetch_state(want_w=False)[0], ref["x"])
        with pytest.raises(smc.SMCBError):
            ctx.log_likelihood(kind, MODELS[kind], 1024, y, smc.MULTINOMIAL)          # small-cloud multinomial: binary64 only
        if kind != smc.KIND_UCSV:
            with pytest.raises(smc.SMCBError):
                ctx.guided_log_likelihood(kind, MODELS[kind], 20000, y, np.tile([0.0, 0.5, 1.0], (T, 1)), smc.SYSTEMATIC)
    finally:
        ctx.set_precision("f64")


def test_f32_arithmetic_tier_agrees_with_binary64(ctx, oracle):
    """BASELINE.json north star: logZ of the fp32 tier within rel 1e-4 of fp64.  The tier draws its normals four per Philox block,
    so it is another Monte-Carlo run of the same filter: at N = 2^22 the two estimates of log Z agree to the north star's
    tolerance, and both agree with the matched-init Kalman likelihood (the exact value)."""
    N, T = 1 << 22, 100
    y = _data(oracle, smc.KIND_LG1D, T)
    z64 = ctx.log_likelihood(smc.KIND_LG1D, MODELS[smc.KIND_LG1D], N, y, smc.SYSTEMATIC)
    ctx.set_precision("f32_arith")
    try:
        z32 = ctx.log_likelihood(smc.KIND_LG1D, MODELS[smc.KIND_LG1D], N, y, smc.SYSTEMATIC)
        z32m = ctx.log_likelihood(smc.KIND_LG1D, MODELS[smc.KIND_LG1D], N, y, smc.MULTINOMIAL)
    finally:
        ctx.set_precision("f64")
    kal = float(ctx.kalman_loglik(MODELS[smc.KIND_LG1D], y, matched_init=True)[0][0])
    assert abs(z32 - z64) <= 1e-4 * abs(z64), (z32, z64)
    assert abs(z32m - z64) <= 1e-4 * abs(z64), (z32m, z64)
    assert abs(z32 - kal) <= 1e-4 * abs(kal) and abs(z64 - kal) <= 1e-4 * abs(kal), (z32, z64, kal)


def _mv_model(d, rng):
    """a stable d-dimensional MultivariateLinearGaussian with a full-rank Q (state_space_models.jl:137-154)"""
    A = 0.7 * np.eye(d) + 0.1 * rng.normal(size=(d, d)) / d
    G = rng.normal(size=(d, d))
    S0 = rng.normal(size=(d, d))
    return smc.MultivariateLinearGaussian(A=A, B=rng.normal(size=d), Q=0.3 * (G @ G.T) / d + 0.05 * np.eye(d), R=[0.6],
                                          X0=0.3 * rng.normal(size=d), Σ0=(S0 @ S0.T) / d + 0.2 * np.eye(d))


@pytest.mark.parametrize("d", [2, 3, 4])
def test_multivariate_linear_gaussian_particle_filter(ctx, oracle, d):
    """the particle-filter functor of MultivariateLinearGaussian (state_space_models.jl:156-189; MvNormal(A x, Q)
    transition through the Cholesky factor of Q, scalar observation): ancestors, states, log-weights bit-exact against the oracle
    for every resampler incl. the two-level multinomial draw, and the likelihood against the reference's own exact filter —
    the matrix Kalman filter (kalman_filter.jl:3-27), on the device and in the oracle."""
    rng = np.random.default_rng(40 + d)
    m = _mv_model(d, rng)
    T = 10
    x_sim, y = smc.simulate(m, T, seed=5)
    xo, yo = oracle.simulate(m.kind, m.block(), T, 5)
    np.testing.assert_array_equal(y, yo)
    np.testing.assert_array_equal(x_sim, xo.T)
    for N, rs in ((1024, smc.SYSTEMATIC), (5001, smc.MULTINOMIAL), (20011, smc.STRATIFIED), (30000, smc.MULTINOMIAL)):
        ref = oracle.log_likelihood(m.kind, m.block(), N, y, rs, 21, 1, 0, want_anc=True)
        ctx.set_rng(21, 1)
        ctx.record_ancestors(True)
        logZ, logmu, ess = ctx.log_likelihood(m.kind, m.block(), N, y, rs, 0, per_step=True)
        anc = ctx.fetch_ancestors(T - 1)
        x, w, logw = ctx.fetch_state(want_logw=True)
        ctx.record_ancestors(False)
        assert x.shape == (d, N)
        np.testing.assert_array_equal(anc, ref["anc"][1:])
        np.testing.assert_array_equal(x, ref["x"])
        np.testing.assert_array_equal(logw, ref["logw"])
        np.testing.assert_allclose(logmu, ref["logmu"], rtol=RTOL, atol=0)
        assert abs(logZ - ref["logZ"]) <= RTOL * abs(ref["logZ"])
    # the public API and the exact answer: PF log-mean-exp over seeds against the matched-init matrix Kalman filter
    kal = float(ctx.kalman_mv_loglik(d, m.block(), y, matched_init=True)[0][0])
    zs = []
    for seed in range(6):
        ctx.set_rng(100 + seed, 0)
        xh, wh, z = smc.log_likelihood(1 << 17, y, m, resampler="systematic", ctx=ctx)
        zs.append(z)
    assert np.asarray(xh).shape == (1 << 17, d) and abs(np.asarray(wh).sum() - 1) < 1e-12
    zs = np.array(zs)
    lme = zs.max() + np.log(np.mean(np.exp(zs - zs.max())))
    assert abs(lme - kal) < 4 * max(zs.std(), 1e-3) / np.sqrt(len(zs)) + 5e-3, (lme, kal, zs.std())
    mean, var, q = ctx.summary([0.5])
    assert mean.shape == (d,) and q.shape == (d, 1)
    with pytest.raises(smc.SMCBError):
        ctx.batch(m.kind, 4, 128)                              # batched / θ-level engines: univariate kinds and UCSV only


def test_hodrick_prescott_particle_filter(ctx, oracle):
    """hodrick_prescott (state_space_models.jl:193-202): Q = diag(1/λ, 0) is singular — the reference's MvNormal(A x, Q) throws —
    and the second state component is the lagged first; the functor's semi-definite Cholesky makes that exact."""
    y = smc.simulate(smc.LinearGaussian(0.9, 1.0, 0.3, 1.0, 2.0), 40, seed=3)[1]
    hp = smc.hodrick_prescott(λ=50.0, y=y, init_cov=4.0)
    N = 50000
    ref = oracle.log_likelihood(hp.kind, hp.block(), N, y, smc.SYSTEMATIC, 8, 0, 0)
    ctx.set_rng(8, 0)
    z = ctx.log_likelihood(hp.kind, hp.block(), N, y, smc.SYSTEMATIC)
    x, _, _ = ctx.fetch_state(want_w=False)
    np.testing.assert_array_equal(x, ref["x"])
    assert abs(z - ref["logZ"]) <= RTOL * abs(ref["logZ"])
    kal = float(ctx.kalman_mv_loglik(2, hp.block(), y, matched_init=True)[0][0])
    assert abs(z - kal) < 0.5, (z, kal)


def test_readme_loop_through_the_host_mirror(oracle):
    """The README's online loop (README.md:33-61) spelled with the reference's function names:
    bootstrap_filter, then bootstrap_filter! and quantile per observation; particle_filter / particle_filter!
    with proposal = nothing (particles.jl:28-84, examples/inflation_example.jl:164,178) is the same filter."""
    from sequential_monte_carlo_b200 import particles
    kind, N, T = smc.KIND_LG1D, 3000, 12
    y = _data(oracle, kind, T)
    model = smc.StateSpaceModel(smc.LinearGaussian(0.5, 1.0, 0.9, 0.8, 0.0), (1, 1))
    c = smc.Context(0, seed=77)
    particles.set_default_context(c)
    try:
        runs = []
        for init, step in ((smc.bootstrap_filter, smc.bootstrap_filter_),
                           (lambda n, y0, m: smc.particle_filter(n, y0, m, None), lambda x, w, yt, m, **k: smc.particle_filter_(x, w, yt, m, None, **k))):
            c.set_rng(77, 0)
            x, w, logmu = init(N, y[0], model)
            logZ, qs = logmu, []
            for t in range(1, T):
                logmu, w, ess = step(x, w, y[t], model, resampler="systematic")
                logZ += logmu
                qs.append(smc.quantile(x, [0.25, 0.5, 0.75]))
            runs.append((np.array(x.numpy()), np.array(w.numpy()), logZ, np.array(qs)))
        for a, b in zip(runs[0], runs[1]):
            np.testing.assert_array_equal(a, b)
        ref = oracle.log_likelihood(kind, MODELS[kind], N, y, smc.SYSTEMATIC, 77, 0, 0)
        np.testing.assert_array_equal(runs[0][0].reshape(-1), ref["x"].reshape(-1))
        assert abs(runs[0][2] - ref["logZ"]) <= RTOL * abs(ref["logZ"])
        xs = np.sort(ref["x"][0])
        np.testing.assert_array_equal(runs[0][3][-1], xs[np.minimum((np.array([0.25, 0.5, 0.75]) * N).astype(int), N - 1)])
    finally:
        particles.set_default_context(None)
        c.close()
