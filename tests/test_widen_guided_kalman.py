"""SURVEY §8(f) rows N3 / N4 and the per-θ variances:

  * guided particle filter (particle_filter / particle_filter! with a proposal, /root/reference/src/particles.jl:28-84;
    docs/SPEC.md §10) — the oracle's restatement pinned on the CPU (a proposal equal to the transition IS the
    bootstrap filter; the locally optimal proposal targets the matched-init Kalman likelihood with a smaller
    variance), the CUDA path bit-exact against it on the GPU through the C ABI;
  * matrix Kalman filter (MultivariateLinearGaussian / hodrick_prescott, state_space_models.jl:137-202;
    kalman_filter.jl:3-27,55-70) — oracle against a numpy restatement on the CPU, CUDA against the oracle on the GPU;
  * per-θ weighted means / variances of the clouds (var(x, weights(w)), examples/inflation_example.jl:46).

Also here: the plain-C client of the ABI on the GPU, IBIS over a multivariate model, and the CUDA path against the
committed golden bit patterns.  The file sorts last; all of it ran green on a B200 (profiles/r1_pytest_gpu_widen_rows.log).
"""
import numpy as np
import pytest

import sequential_monte_carlo_b200 as smc

LG = [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]
SVP = [-1.0, 0.9, 0.3]
RTOL = 1e-10


def _lg_thetas(M, rng):
    return np.stack([rng.uniform(-0.9, 0.9, M), np.ones(M), rng.uniform(0.3, 2, M), rng.uniform(0.3, 2, M), np.zeros(M), np.ones(M)], 1)


def _sv_thetas(M, rng):
    return np.stack([rng.normal(-1, 0.5, M), rng.uniform(0.5, 0.98, M), rng.uniform(0.1, 0.6, M)], 1)


def _stable(rng, d):
    A = rng.standard_normal((d, d))
    return 0.9 * A / np.linalg.norm(A, 2)          # spectral norm 0.9: a stable transition


def _numpy_kalman(A, B, Q, R, x, S, y, matched_init=False):
    A, B, Q = np.asarray(A, float), np.asarray(B, float).reshape(1, -1), np.asarray(Q, float)
    x, S, ll = np.asarray(x, float).copy(), np.asarray(S, float).copy(), 0.0
    for t, yt in enumerate(y):
        if not (matched_init and t == 0):
            x, S = A @ x, A @ S @ A.T + Q
        sig = (B @ S @ B.T)[0, 0] + R
        dy = yt - (B @ x)[0]
        K = (S @ B.T)[:, 0]
        x, S = x + K / sig * dy, S - np.outer(K, K) / sig
        ll += -0.5 * (np.log(2 * np.pi) + np.log(sig) + dy * dy / sig)
    return x, S, ll


# ------------------------------------------------------------------------------------------------ CPU: oracle pinned
def test_oracle_guided_with_transition_proposal_is_bootstrap(oracle):
    """q = f: same draws, same states, and transition − proposal cancels exactly (SPEC §10)"""
    for kind, th, coef in ((oracle.KIND_LG1D, LG, [0.0, 0.5, np.sqrt(0.9)]), (oracle.KIND_SV, SVP, [-1.0 * (1 - 0.9), 0.9, 0.3])):
        _, y = oracle.simulate(kind, th, 40, 11)
        prop = np.tile(coef, (y.size, 1))
        g = oracle.guided_log_likelihood(kind, th, 600, y, oracle.SYSTEMATIC, prop, 3, 1, 2)
        b = oracle.log_likelihood(kind, th, 600, y, oracle.SYSTEMATIC, 3, 1, 2)
        if kind == oracle.KIND_LG1D:
            np.testing.assert_array_equal(g["x"], b["x"])
            assert g["logZ"] == b["logZ"]
        else:  # SV: μ + ρ(xp − μ) and c0 + c1 xp round differently; the clouds agree to rounding until a resample splits them
            np.testing.assert_allclose(g["logmu"][:3], b["logmu"][:3], rtol=1e-9)


def test_oracle_guided_optimal_proposal_targets_kalman(oracle):
    """E[Ẑ] = Z for any proposal; the locally optimal one has a much smaller variance than the bootstrap filter"""
    _, y = oracle.simulate(oracle.KIND_LG1D, LG, 100, 1998)
    prop = np.array([oracle.optimal_proposal_lg(LG, yt) for yt in y])
    _, _, kll = oracle.kalman_loglik(LG, y, matched_init=True)
    g = np.array([oracle.guided_log_likelihood(0, LG, 1024, y, oracle.SYSTEMATIC, prop, s)["logZ"] for s in range(24)])
    b = np.array([oracle.log_likelihood(0, LG, 1024, y, oracle.SYSTEMATIC, s)["logZ"] for s in range(24)])
    assert g.std() < 0.5 * b.std()
    lme = np.log(np.mean(np.exp(g - g.max()))) + g.max()          # log-mean-exp of the unbiased estimates
    assert abs(lme - kll) < 4 * g.std() / np.sqrt(g.size) + 0.02


def test_oracle_guided_step_against_a_numpy_restatement(oracle):
    """particle_filter! (particles.jl:66-80) restated with numpy / libm on the oracle's own ancestors and normals:
    x' = c0 + c1 xp + c2 z, logw = logpdf(obs) + logpdf(N(x'; μf, σf)) − logpdf(N(x'; c0 + c1 xp, c2))"""
    def lognorm(v, mu, sd):
        return -0.5 * ((v - mu) / sd) ** 2 - np.log(sd) - 0.5 * np.log(2 * np.pi)
    n, t, seed, epoch, stream = 500, 3, 9, 2, 4
    for kind, th, prop in ((oracle.KIND_LG1D, LG, [0.2, 0.3, 0.7]), (oracle.KIND_SV, SVP, [-0.4, 0.6, 0.5])):
        x, lw = oracle.bootstrap_init(kind, th, n, 0.4, seed, epoch, stream)
        x0, lw0 = x.copy(), lw.copy()
        a = oracle.guided_step(kind, th, x, lw, -0.3, t, oracle.SYSTEMATIC, prop, seed, epoch, stream)
        np.testing.assert_array_equal(a, oracle.ancestors(lw0, oracle.SYSTEMATIC, seed, epoch, stream, t))
        xp = x0[0][a]
        z = oracle.normals(seed, epoch, stream, t, oracle.P_TRANS, 0, n)
        xn = prop[0] + prop[1] * xp + prop[2] * z
        np.testing.assert_allclose(x[0], xn, rtol=1e-14, atol=1e-15)
        if kind == oracle.KIND_LG1D:
            A, B, Q, R = th[:4]
            lobs, lf = lognorm(-0.3, B * xn, np.sqrt(R)), lognorm(xn, A * xp, np.sqrt(Q))
        else:
            mu, rho, sig = th
            lobs, lf = lognorm(-0.3, 0.0, np.exp(0.5 * xn)), lognorm(xn, mu + rho * (xp - mu), sig)
        np.testing.assert_allclose(lw, lobs + lf - lognorm(xn, prop[0] + prop[1] * xp, prop[2]), rtol=1e-11, atol=1e-12)


def test_oracle_matrix_kalman(oracle):
    """smco_kalman_mv_* against numpy matrix algebra; d = 1 reproduces the scalar recursion; Hodrick–Prescott block"""
    _, y = oracle.simulate(oracle.KIND_LG1D, LG, 80, 5)
    blk = oracle.mv_block([[0.5]], [1.0], [[0.9]], [0.8], [0.0], [[1.0]])
    for matched in (False, True):
        assert oracle.kalman_mv_loglik(1, blk, y, matched)[2] == pytest.approx(oracle.kalman_loglik(LG, y, matched)[2], rel=1e-13)
    rng = np.random.default_rng(0)
    for d in (2, 3, 4):
        A = _stable(rng, d)
        G = rng.standard_normal((d, d))
        Q, S0 = G @ G.T + 0.1 * np.eye(d), np.eye(d) * 2.0
        B, R, x0 = rng.standard_normal(d), 0.7, rng.standard_normal(d)
        blk = oracle.mv_block(A, B, Q, [R], x0, S0)
        for matched in (False, True):
            x, S, ll = oracle.kalman_mv_loglik(d, blk, y, matched)
            xn, Sn, lln = _numpy_kalman(A, B, Q, R, x0, S0, y, matched)
            assert ll == pytest.approx(lln, rel=1e-11)
            np.testing.assert_allclose(x, xn, rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(S, Sn, rtol=1e-9, atol=1e-12)
    hp = smc.hodrick_prescott(λ=1600.0, y=y)
    x, S, ll = oracle.kalman_mv_loglik(2, hp.block(), y)
    xn, Sn, lln = _numpy_kalman(hp.A, hp.B, hp.Q, hp.R[0], hp.x0, hp.σ0, y)
    assert ll == pytest.approx(lln, rel=1e-11)
    np.testing.assert_allclose(x, xn, rtol=1e-8)


def test_host_mirror_of_the_new_models_and_proposals():
    y = np.arange(6.0) ** 1.5
    hp = smc.hodrick_prescott(λ=1600.0, y=y)
    assert hp.state_dim == 2 and hp.block().shape == (17,)
    np.testing.assert_array_equal(hp.A, [[2.0, -1.0], [1.0, 0.0]])                 # state_space_models.jl:195
    np.testing.assert_array_equal(hp.x0, [3 * y[0] - 2 * y[1], 2 * y[0] - y[1]])    # :199
    assert hp.Q[0, 0] == 1 / 1600.0 and hp.σ0[1, 1] == 1000.0
    m = smc.MultivariateLinearGaussian(A=np.eye(3), B=[1, 0, 0], Q=np.eye(3), R=[2.0])
    np.testing.assert_array_equal(m.x0, np.zeros(3))                               # X0 = zeros, Σ0 = I  :137
    np.testing.assert_array_equal(m.σ0, np.eye(3))
    assert m.kind == smc._lib.MVLG3 and np.array_equal(m.params(), m.block())          # the particle-filter functor's block (SPEC §4b)
    with pytest.raises(NotImplementedError):
        smc.MultivariateLinearGaussian(A=[[0.5]], B=[1.0], Q=[[1.0]], R=[1.0]).params()  # d = 1: LinearGaussian is the univariate kind
    lg = smc.LinearGaussian(0.5, 1.0, 0.9, 0.8)
    c0, c1, c2 = smc.locally_optimal_proposal(lg, 0.3)
    s2 = 1 / (1 / 0.9 + 1 / 0.8)
    assert (c0, c1, c2) == pytest.approx((s2 * 0.3 / 0.8, s2 * 0.5 / 0.9, np.sqrt(s2)))
    assert smc.AffineGaussianProposal(1, 2, 3)(lg, 9.0) == (1.0, 2.0, 3.0)


UCSVP = [0.2, 0.2, 3.0, 1.0, 1.0]


def test_oracle_guided_ucsv_move(oracle):
    """SPEC §10b: κ = 0 IS the bootstrap filter (same states; the weight correction −½zt² + ½z² vanishes to rounding), κ = 1 is
    the conditionally optimal trend move: logZ estimates the same likelihood with a much smaller scatter"""
    N, T = 512, 60
    _, y = oracle.simulate(2, UCSVP, T, 5)
    z0, z1, zb = [], [], []
    for seed in range(24):
        b = oracle.log_likelihood(2, UCSVP, N, y, oracle.SYSTEMATIC, seed, 0, 0)
        g0 = oracle.guided_log_likelihood(2, UCSVP, N, y, oracle.SYSTEMATIC, np.tile([0.0, 0.0, 1.0], (T, 1)), seed)
        g1 = oracle.guided_log_likelihood(2, UCSVP, N, y, oracle.SYSTEMATIC, np.tile([1.0, 0.0, 1.0], (T, 1)), seed)
        if seed == 0:
            np.testing.assert_array_equal(g0["x"], b["x"])
            np.testing.assert_allclose(g0["logw"], b["logw"], rtol=0, atol=1e-12)
            assert g1["x"].shape == (3, N) and np.all(np.isfinite(g1["logw"]))
        zb.append(b["logZ"]); z0.append(g0["logZ"]); z1.append(g1["logZ"])
    np.testing.assert_allclose(z0, zb, rtol=1e-12)
    assert np.std(z1) < 0.6 * np.std(zb)
    lme = lambda z: np.log(np.mean(np.exp(np.array(z) - np.max(z)))) + np.max(z)      # noqa: E731
    assert abs(lme(z1) - lme(zb)) < 3 * np.std(zb) / np.sqrt(24) + 0.1


# ------------------------------------------------------------------------------------------------ GPU: parity
@pytest.mark.gpu
@pytest.mark.parametrize("N", [1024, 777, 4096, 6])
def test_guided_ucsv_batch_bit_exact(ctx, oracle, N):
    """the tempered optimal trend move of UCSV (SPEC §10b) in the batched engine (clouds in shared memory, and in global memory at
    N = 4096), κ per (t, θ): clouds and log-weights bit-exact against the oracle, logZ to 1e-10"""
    M, T = 5, 12 if N <= 1024 else 5
    rng = np.random.default_rng(N)
    _, y = oracle.simulate(2, UCSVP, T, 1998)
    th = np.stack([rng.uniform(0.1, 0.4, M), rng.uniform(0.1, 0.4, M), rng.normal(3, 0.3, M), rng.normal(1, 0.2, M), rng.normal(1, 0.2, M)], 1)
    P = smc._lib.params8(th)
    prop = np.zeros((T, M, 3))
    prop[:, :, 0] = rng.uniform(0, 1, (T, M))
    prop[:, 0, 0], prop[:, 1, 0] = 1.0, 0.0
    prop[:, :, 2] = 1.0
    active = np.ones(M, np.uint8)
    active[3] = 0
    for resampler in (smc.MULTINOMIAL, smc.STRATIFIED, smc.SYSTEMATIC):
        seed, epoch, stream0 = 78, 3 + resampler, 41
        zo, xo, lwo = oracle.batch_guided_log_likelihood(2, P, active, N, y, resampler, prop, seed, epoch, stream0)
        b = ctx.batch(smc.KIND_UCSV, M, N)
        ctx.set_rng(seed, epoch)
        z = b.log_likelihood(P, y, resampler, stream0, active, proposal=prop)
        x, _, lw = b.fetch(want_w=False, want_logw=True)
        on = active.astype(bool)
        assert np.all(np.isneginf(z[~on]))
        np.testing.assert_array_equal(x[on], xo[on])
        np.testing.assert_array_equal(lw[on], lwo[on])
        np.testing.assert_allclose(z[on], zo[on], rtol=RTOL, atol=0)
        b.close()
    b = ctx.batch(smc.KIND_UCSV, M, N)
    bad = prop.copy()
    bad[1, 2, 0] = 1.2
    with pytest.raises(smc.SMCBError):
        b.log_likelihood(P, y, smc.SYSTEMATIC, 0, proposal=bad)
    b.close()


@pytest.mark.gpu
def test_guided_ucsv_single_filter_and_python_mirror(ctx, oracle):
    """the same move on the grid-wide path (guided_move_ucsv_kernel) for one large cloud, whole series and stepping API, and through
    particle_filter / particle_filter! with UCSVTrendProposal (small cloud: batched engine; large: single filter)"""
    N, T = 20011, 10
    _, y = oracle.simulate(2, UCSVP, T, 3)
    prop = np.tile([1.0, 0.0, 1.0], (T, 1))
    prop[::2, 0] = 0.6
    for resampler in (smc.STRATIFIED, smc.SYSTEMATIC):
        seed, epoch, stream = 32, 2 + resampler, 8
        ref = oracle.guided_log_likelihood(2, UCSVP, N, y, resampler, prop, seed, epoch, stream)
        ctx.set_rng(seed, epoch)
        z, lm, es = ctx.guided_log_likelihood(smc.KIND_UCSV, UCSVP, N, y, prop, resampler, stream, per_step=True)
        x, _, lw = ctx.fetch_state(want_w=False, want_logw=True)
        np.testing.assert_array_equal(x, ref["x"])
        np.testing.assert_array_equal(lw, ref["logw"])
        np.testing.assert_allclose(lm, ref["logmu"], rtol=RTOL, atol=0)
        assert abs(z - ref["logZ"]) <= RTOL * abs(ref["logZ"])
    ucsv = smc.UCSV((UCSVP[0], UCSVP[1]), UCSVP[2], (UCSVP[3], UCSVP[4]))
    for n in (1500, 30000):
        ctx.set_rng(13, 5)
        xs, w, logmu = smc.particle_filter(n, y[0], ucsv, smc.UCSVTrendProposal(1.0), ctx=ctx, stream=6)
        xo, lwo = oracle.bootstrap_init(2, UCSVP, n, y[0], 13, 5, 6)
        for t in range(1, 6):
            logmu, w, ess = smc.particle_filter_(xs, w, y[t], ucsv, smc.UCSVTrendProposal(1.0), resampler="systematic")
            oracle.guided_step(2, UCSVP, xo, lwo, y[t], t, oracle.SYSTEMATIC, [1.0, 0.0, 1.0], 13, 5, 6)
            lmo, wo, esso = oracle.normalize(lwo)
            assert abs(logmu - lmo) <= RTOL * abs(lmo) and abs(ess - esso) <= 1e-9 * esso
        np.testing.assert_array_equal(np.asarray(xs), xo.T)
        np.testing.assert_allclose(np.asarray(w), wo, rtol=RTOL)
    with pytest.raises(ValueError):
        smc.UCSVTrendProposal(1.5)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N", [(smc.KIND_LG1D, 1024), (smc.KIND_LG1D, 777), (smc.KIND_SV, 2048), (smc.KIND_LG1D, 8192), (smc.KIND_SV, 5)])
def test_guided_batch_bit_exact(ctx, oracle, kind, N):
    """M guided filters, whole series in one launch: clouds and log-weights bit-exact against the oracle, logZ to 1e-10"""
    M, T = 9, 25 if N <= 4096 else 6
    rng = np.random.default_rng(N)
    true = LG if kind == smc.KIND_LG1D else SVP
    _, y = oracle.simulate(kind, true, T, 1998)
    th = _lg_thetas(M, rng) if kind == smc.KIND_LG1D else _sv_thetas(M, rng)
    P = smc._lib.params8(th)
    if kind == smc.KIND_LG1D:   # every θ's own locally optimal proposal, per step
        prop = np.array([[oracle.optimal_proposal_lg(th[m], yt) for m in range(M)] for yt in y])
    else:                       # SV: a damped, widened version of the transition that leans on |y|
        prop = np.array([[[th[m, 0] * (1 - 0.8 * th[m, 1]) + 0.05 * np.log(yt * yt + 1e-3), 0.8 * th[m, 1], 1.3 * th[m, 2]]
                          for m in range(M)] for yt in y])
    active = np.ones(M, np.uint8)
    active[4] = 0
    for resampler in (smc.MULTINOMIAL, smc.STRATIFIED, smc.SYSTEMATIC):
        seed, epoch, stream0 = 77, 3 + resampler, 40
        zo, xo, lwo = oracle.batch_guided_log_likelihood(kind, P, active, N, y, resampler, prop, seed, epoch, stream0)
        b = ctx.batch(kind, M, N)
        ctx.set_rng(seed, epoch)
        z = b.log_likelihood(P, y, resampler, stream0, active, proposal=prop)
        x, _, lw = b.fetch(want_w=False, want_logw=True)
        on = active.astype(bool)
        assert np.all(np.isneginf(z[~on]))
        np.testing.assert_array_equal(x[on], xo[on])
        np.testing.assert_array_equal(lw[on], lwo[on])
        np.testing.assert_allclose(z[on], zo[on], rtol=RTOL, atol=0)
        b.close()


@pytest.mark.gpu
def test_guided_stepping_api_and_python_mirror(ctx, oracle):
    """particle_filter / particle_filter! with a proposal (particles.jl:28-84) one observation per call == the oracle's
    step-by-step run; guided_log_likelihood == the same series in one launch; proposal=None continues as bootstrap"""
    N, T = 1500, 20
    lg = smc.LinearGaussian(0.5, 1.0, 0.9, 0.8)
    _, y = oracle.simulate(0, LG, T, 21)
    ctx.set_rng(13, 5)
    x, w, logmu = smc.particle_filter(N, y[0], lg, smc.locally_optimal_proposal, ctx=ctx, stream=6)
    xo, lwo = oracle.bootstrap_init(0, LG, N, y[0], 13, 5, 6)
    lmo, _, _ = oracle.normalize(lwo)
    assert abs(logmu - lmo) <= RTOL * abs(lmo)
    logZ = logmu
    for t in range(1, T):
        logmu, w, ess = smc.particle_filter_(x, w, y[t], lg, smc.locally_optimal_proposal, resampler="systematic")
        oracle.guided_step(0, LG, xo, lwo, y[t], t, oracle.SYSTEMATIC, oracle.optimal_proposal_lg(LG, y[t]), 13, 5, 6)
        lmo, wo, esso = oracle.normalize(lwo)
        assert abs(logmu - lmo) <= RTOL * abs(lmo) and abs(ess - esso) <= 1e-9 * esso
        logZ += logmu
    np.testing.assert_array_equal(np.asarray(x), xo[0])
    np.testing.assert_allclose(np.asarray(w), wo, rtol=RTOL)
    # summaries of the guided cloud on the device: quantile(x, weights(w), p), mean / var(x, weights(w))
    ps = [0.25, 0.5, 0.75]
    mo, vo, qo = oracle.weighted_summary(xo, lwo, ps, weighted=True)
    np.testing.assert_array_equal(smc.quantile(x, w, ps), qo[0])
    np.testing.assert_array_equal(smc.quantile(x, ps), oracle.weighted_summary(xo, lwo, ps, weighted=False)[2][0])
    m, v = smc.weighted_mean_var(x, w)
    assert abs(m - float(xo[0] @ wo)) <= 1e-10 * max(1.0, abs(m)) and abs(v - float(((xo[0] - m) ** 2) @ wo)) <= 1e-9 * v
    ctx.set_rng(13, 5)
    x2, _, logZ2 = smc.guided_log_likelihood(N, y, lg, smc.locally_optimal_proposal, resampler="systematic", ctx=ctx, stream=6)
    np.testing.assert_array_equal(np.asarray(x2), xo[0])
    assert abs(logZ2 - logZ) <= RTOL * abs(logZ)
    # a bootstrap step on the same cloud (proposal=None) == the oracle's bootstrap step
    logmu, w, _ = smc.particle_filter_(x, w, 0.25, lg, None, resampler="systematic")
    oracle.bootstrap_step(0, LG, xo, lwo, 0.25, T, oracle.SYSTEMATIC, 13, 5, 6)
    np.testing.assert_array_equal(np.asarray(x), xo[0])
    # errors: UCSV takes (κ, 0, 1) with κ in [0, 1] (SPEC §10b); a non-positive proposal sd is refused; coefficients must be three numbers
    b = ctx.batch(smc.KIND_UCSV, 2, 64)
    b.init(np.tile(smc._lib.params8([0.2, 0.2, 3.0, 1.0, 1.0]), (2, 1)), 1.0)
    with pytest.raises(smc.SMCBError):
        b.step(1.0, smc.SYSTEMATIC, proposal=np.tile([1.5, 0.0, 1.0], (2, 1)))
    b.close()
    b = ctx.batch(smc.KIND_LG1D, 2, 64)
    b.init(np.tile(smc._lib.params8(LG), (2, 1)), 1.0)
    with pytest.raises(smc.SMCBError):
        b.step(1.0, smc.SYSTEMATIC, proposal=np.tile([0.0, 1.0, 0.0], (2, 1)))
    b.close()
    with pytest.raises(TypeError):
        smc.particle_filter_(x, w, 0.1, lg, lambda model, yt: (1.0, 2.0))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,f32", [(smc.KIND_LG1D, 20011, False), (smc.KIND_SV, 9000, False), (smc.KIND_LG1D, 1 << 16, True),
                                         (smc.KIND_SV, 12345, True), (smc.KIND_LG1D, 7, False)])
def test_guided_single_filter_bit_exact(ctx, oracle, kind, N, f32):
    """the grid-wide guided step (guided_move_kernel after the sorted-resampler kernels), whole series and stepping API,
    binary64 and binary32-state tiers: states and log-weights bit-exact against the oracle, logZ / logμ to 1e-10"""
    T = 14
    true = LG if kind == smc.KIND_LG1D else SVP
    _, y = oracle.simulate(kind, true, T, 3)
    if kind == smc.KIND_LG1D:
        prop = np.array([oracle.optimal_proposal_lg(true, yt) for yt in y])
    else:
        prop = np.array([[-1.0 * (1 - 0.8 * 0.9) + 0.05 * np.log(yt * yt + 1e-3), 0.8 * 0.9, 1.3 * 0.3] for yt in y])
    try:
        ctx.set_precision("f32" if f32 else "f64")
        for resampler in (smc.STRATIFIED, smc.SYSTEMATIC):
            seed, epoch, stream = 31, 2 + resampler, 9
            with oracle.state_f32(f32):
                ref = oracle.guided_log_likelihood(kind, true, N, y, resampler, prop, seed, epoch, stream)
            ctx.set_rng(seed, epoch)
            z, lm, es = ctx.guided_log_likelihood(kind, true, N, y, prop, resampler, stream, per_step=True)
            x, _, lw = ctx.fetch_state(want_w=False, want_logw=True)
            np.testing.assert_array_equal(x, ref["x"])
            np.testing.assert_array_equal(lw, ref["logw"])
            np.testing.assert_allclose(lm, ref["logmu"], rtol=RTOL, atol=0)
            np.testing.assert_allclose(es, ref["ess"], rtol=1e-9, atol=0)
            assert abs(z - ref["logZ"]) <= RTOL * abs(ref["logZ"])
            # stepping API: bootstrap init, guided steps mixed with one bootstrap step in the middle
            ctx.set_rng(seed, epoch)
            ctx.bootstrap_init(kind, true, N, y[0], stream)
            with oracle.state_f32(f32):
                xo, lwo = oracle.bootstrap_init(kind, true, N, y[0], seed, epoch, stream)
                for t in range(1, 6):
                    if t == 3:
                        lmu, _ = ctx.bootstrap_step(y[t], resampler)
                        oracle.bootstrap_step(kind, true, xo, lwo, y[t], t, resampler, seed, epoch, stream)
                    else:
                        lmu, _ = ctx.guided_step(y[t], prop[t], resampler)
                        oracle.guided_step(kind, true, xo, lwo, y[t], t, resampler, prop[t], seed, epoch, stream)
                    lmo, _, _ = oracle.normalize(lwo)
                    assert abs(lmu - lmo) <= RTOL * abs(lmo)
            x, _, lw = ctx.fetch_state(want_w=False, want_logw=True)
            np.testing.assert_array_equal(x, xo)
            np.testing.assert_array_equal(lw, lwo)
    finally:
        ctx.set_precision("f64")


@pytest.mark.gpu
def test_guided_single_filter_errors_and_python_mirror(ctx, oracle):
    """bad coefficients and multinomial resampling are refused on the grid-wide guided path; particle_filter / particle_filter! route
    clouds above 8192 particles to it"""
    _, y = oracle.simulate(0, LG, 6, 3)
    ctx.bootstrap_init(smc.KIND_UCSV, [0.2, 0.2, 3.0, 1.0, 1.0], 4096, 1.0)
    with pytest.raises(smc.SMCBError):
        ctx.guided_step(1.0, [1.5, 0.0, 1.0], smc.SYSTEMATIC)      # UCSV (SPEC §10b): κ outside [0, 1]
    ctx.bootstrap_init(smc.KIND_LG1D, LG, 4096, y[0])
    with pytest.raises(smc.SMCBError):
        ctx.guided_step(y[1], [0.0, 1.0, 1.0], smc.MULTINOMIAL)
    with pytest.raises(smc.SMCBError):
        ctx.guided_step(y[1], [0.0, 1.0, -1.0], smc.SYSTEMATIC)
    lg, N = smc.LinearGaussian(0.5, 1.0, 0.9, 0.8), 30000
    ctx.set_rng(4, 4)
    x, w, logmu = smc.particle_filter(N, y[0], lg, smc.locally_optimal_proposal, ctx=ctx, stream=1)
    xo, lwo = oracle.bootstrap_init(0, LG, N, y[0], 4, 4, 1)
    for t in range(1, 6):
        logmu, w, ess = smc.particle_filter_(x, w, y[t], lg, smc.locally_optimal_proposal, resampler="systematic")
        oracle.guided_step(0, LG, xo, lwo, y[t], t, oracle.SYSTEMATIC, oracle.optimal_proposal_lg(LG, y[t]), 4, 4, 1)
    np.testing.assert_array_equal(np.asarray(x), xo[0])
    _, wo, _ = oracle.normalize(lwo)
    np.testing.assert_allclose(np.asarray(w), wo, rtol=RTOL)
    ctx.set_rng(4, 4)
    x2, _, _ = smc.guided_log_likelihood(N, y, lg, smc.locally_optimal_proposal, resampler="systematic", ctx=ctx, stream=1)
    np.testing.assert_array_equal(np.asarray(x2), xo[0])


@pytest.mark.gpu
def test_guided_variance_reduction_on_device(ctx, oracle):
    """the point of a proposal: at equal N the locally optimal proposal's logZ scatters far less around the matched-init
    Kalman likelihood than the bootstrap filter's (LG1D, N = 512, 32 independent streams in one batch each)"""
    M, N, T = 32, 512, 100
    _, y = oracle.simulate(0, LG, T, 1998)
    P = np.tile(smc._lib.params8(LG), (M, 1))
    prop = np.array([[oracle.optimal_proposal_lg(LG, yt)] * M for yt in y])
    b = ctx.batch(smc.KIND_LG1D, M, N)
    ctx.set_rng(5, 1)
    zg = b.log_likelihood(P, y, smc.SYSTEMATIC, 0, proposal=prop)
    ctx.set_rng(5, 1)
    zb = b.log_likelihood(P, y, smc.SYSTEMATIC, 0)
    b.close()
    kll = ctx.kalman_loglik(LG, y, matched_init=True)[0][0]
    assert zg.std() < 0.5 * zb.std()
    assert abs(np.log(np.mean(np.exp(zg - zg.max()))) + zg.max() - kll) < 4 * zg.std() / np.sqrt(M) + 0.02


@pytest.mark.gpu
def test_kalman_mv_batch(ctx, oracle):
    """smcb_kalman_mv_batch_loglik / _step, d = 1..4, against the oracle model by model; the Python mirror on
    hodrick_prescott; d = 1 equals the scalar kernel"""
    rng = np.random.default_rng(1)
    _, y = oracle.simulate(0, LG, 70, 9)
    for d in (1, 2, 3, 4):
        M = 21
        blocks = []
        for m in range(M):
            A = _stable(rng, d)
            G = rng.standard_normal((d, d))
            blocks.append(oracle.mv_block(A, rng.standard_normal(d), G @ G.T + 0.1 * np.eye(d), [rng.uniform(0.3, 2)],
                                          rng.standard_normal(d), 2.0 * np.eye(d)))
        blocks = np.stack(blocks)
        act = np.ones(M, np.uint8)
        act[3] = 0
        for matched in (False, True):
            ll, x, S = ctx.kalman_mv_loglik(d, blocks, y, matched, act)
            for m in range(M):
                if not act[m]:
                    assert np.isneginf(ll[m])
                    continue
                xo, So, lo = oracle.kalman_mv_loglik(d, blocks[m], y, matched)
                assert abs(ll[m] - lo) <= 1e-12 * abs(lo)
                np.testing.assert_allclose(x[m], xo, rtol=1e-11, atol=1e-13)
                np.testing.assert_allclose(S[m], So, rtol=1e-11, atol=1e-13)
        x1, S1, l1 = ctx.kalman_mv_step(d, blocks, np.zeros((M, d)), np.broadcast_to(np.eye(d), (M, d, d)), y[0])
        for m in range(M):
            xo, So, lo = oracle.kalman_mv_step(d, blocks[m], np.zeros(d), np.eye(d), y[0])
            assert abs(l1[m] - lo) <= 1e-12 * abs(lo)
            np.testing.assert_allclose(x1[m], xo, rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(S1[m], So, rtol=1e-12, atol=1e-14)
    blk = oracle.mv_block([[0.5]], [1.0], [[0.9]], [0.8], [0.0], [[1.0]])
    assert ctx.kalman_mv_loglik(1, blk, y)[0][0] == pytest.approx(ctx.kalman_loglik(LG, y)[0][0], rel=1e-13)
    hp = smc.hodrick_prescott(λ=1600.0, y=y)
    ll = smc.kalman_filter.log_likelihood(y, hp, ctx=ctx)
    assert ll == pytest.approx(oracle.kalman_mv_loglik(2, hp.block(), y)[2], rel=1e-12)
    xT, ST, ll2 = smc.kalman_filter.filtered_moments(y, hp, ctx=ctx)
    assert ll2 == ll and xT.shape == (2,) and ST.shape == (2, 2)
    x1, S1, l1 = smc.kalman_filter.kalman_filter(hp, hp.x0, hp.σ0, y[0], ctx=ctx)
    xo, So, lo = oracle.kalman_mv_step(2, hp.block(), hp.x0, hp.σ0, y[0])
    np.testing.assert_allclose(x1, xo, rtol=1e-12)
    assert l1 == pytest.approx(lo, rel=1e-12)
    with pytest.raises(smc.SMCBError):
        ctx.kalman_mv_loglik(5, np.zeros((1, 86)), y)


@pytest.mark.gpu
def test_batch_weighted_moments(ctx, oracle):
    """smcb_batch_weighted_moments: per-θ mean and population variance of the clouds under their weights"""
    kind, M, N, T = smc.KIND_UCSV, 7, 1001, 8
    P = np.tile(smc._lib.params8([0.2, 0.2, 3.0, 1.0, 1.0]), (M, 1))
    P[:, 0] = np.linspace(0.1, 0.6, M)
    _, y = oracle.simulate(kind, [0.2, 0.2, 3.0, 1.0, 1.0], T, 5)
    b = ctx.batch(kind, M, N)
    ctx.set_rng(9, 4)
    b.log_likelihood(P, y, smc.SYSTEMATIC, stream0=2)
    _, xo, lwo = oracle.batch_log_likelihood(kind, P, None, N, y, smc.SYSTEMATIC, 9, 4, 2)
    mean, var = b.weighted_moments()
    np.testing.assert_allclose(mean, b.weighted_mean(), rtol=1e-13)
    for m in range(M):
        _, w, _ = oracle.normalize(lwo[m])
        mo = xo[m] @ w
        vo = ((xo[m] - mo[:, None]) ** 2) @ w
        np.testing.assert_allclose(mean[m], mo, rtol=1e-10)
        np.testing.assert_allclose(var[m], vo, rtol=1e-9)
    b.close()


@pytest.mark.gpu
def test_plain_c_client_runs_the_path(oracle):
    """tests/abi_client.c (plain C, dlopen): one bootstrap filter, three guided filters with per-θ moments, the
    Kalman entry points and a whole smc² run of the device-resident sampler, checked against the oracle — the boundary works without Python or C++ on the caller's side"""
    from tests.test_abi import run_c_client
    got = {k: float(v) for k, v in run_c_client("gpu").items()}
    _, y = oracle.simulate(0, LG, 40, 1998)
    ref = oracle.log_likelihood(0, LG, 2048, y, oracle.SYSTEMATIC, 7, 1, 0)
    assert abs(got["pf_logZ"] - ref["logZ"]) <= RTOL * abs(ref["logZ"])
    assert got["pf_x0"] == ref["x"][0, 0] and got["pf_xlast"] == ref["x"][0, -1]
    # the C client derives the coefficients with its own arithmetic; feed the oracle the very same numbers
    s2 = 1.0 / (1.0 / LG[2] + LG[1] * LG[1] / LG[3])
    prop = np.array([[[s2 * LG[1] * yt / LG[3], s2 * LG[0] / LG[2], np.sqrt(s2)]] * 3 for yt in y])
    P = np.tile(smc._lib.params8(LG), (3, 1))
    zo, xo, lwo = oracle.batch_guided_log_likelihood(0, P, None, 512, y, oracle.SYSTEMATIC, prop, 7, 2, 10)
    for m in range(3):
        assert abs(got[f"guided_logZ{m}"] - zo[m]) <= RTOL * abs(zo[m])
        _, w, _ = oracle.normalize(lwo[m])
        mo = float(xo[m, 0] @ w)
        assert abs(got[f"guided_mean{m}"] - mo) <= 1e-10 * max(1.0, abs(mo))
        assert abs(got[f"guided_var{m}"] - float(((xo[m, 0] - mo) ** 2) @ w)) <= 1e-9
    lo = oracle.kalman_loglik(LG, y)[2]
    assert abs(got["kalman_mv"] - lo) <= 1e-12 * abs(lo) and abs(got["kalman_scalar"] - lo) <= 1e-12 * abs(lo)
    # the device-resident sampler driven from C: the same θ0 through the oracle's smc² (theta overridden after construction)
    from oracle import samplers as S
    MS, NS = 32, 128
    m_ = np.arange(MS)
    th0 = np.stack([-0.9 + 1.8 * (m_ + 0.5) / MS, 0.4 + 0.05 * ((m_ * 7) % MS), 0.5 + 0.04 * ((m_ * 11) % MS)], 1)
    po = S.OProduct([S.OTruncatedNormal(0, 1, -1, 1), S.OLogNormal(), S.OLogNormal()])
    o_ = S.OSMC(NS, MS, lambda θ: (0, [θ[0], 1.0, θ[1], θ[2], 0.0, 1.0]), po, 2, 0.5, seed=11, resampler=oracle.SYSTEMATIC)
    o_.theta = th0.copy()
    S.o_smc2(o_, y)
    nrej = 0
    for t in range(1, len(y)):
        S.o_smc2_step(o_, y, t)
        nrej += int(o_.rejuvenated)
    assert int(got["smc2_rejuvenations"]) == nrej and nrej >= 1 and int(got["smc2_N"]) == NS
    assert got["smc2_theta_sum"] == pytest.approx(float(o_.theta.sum()), rel=1e-13)       # θ bit-identical: only the summation differs
    assert got["smc2_logZ_sum"] == pytest.approx(float(o_.logZ.sum()), rel=1e-10)
    assert got["smc2_ess"] == pytest.approx(o_.ess, rel=1e-9) and got["smc2_omega_sum"] == pytest.approx(1.0, abs=1e-12)


@pytest.mark.gpu
def test_ibis_multivariate_hodrick_prescott(ctx, oracle):
    """IBIS (ibis.jl:3-189) is generic in the state type: with a multivariate LinearModel the inner filter is the matrix
    Kalman filter.  Hodrick–Prescott with an unknown λ ~ U(1, 2000): the device sampler against the oracle's, step by step"""
    from oracle import samplers as S
    from sequential_monte_carlo_b200 import ibis as ib
    rng = np.random.default_rng(0)
    T, M = 80, 96
    y = np.cumsum(np.cumsum(rng.normal(0, 0.05, T))) + rng.normal(0, 1, T)
    g = smc.IBIS(M, lambda θ: smc.hodrick_prescott(λ=θ[0], y=y), smc.product_distribution([smc.Uniform(1.0, 2000.0)]), 2, 0.5, seed=3, ctx=ctx)
    o_ = S.OIBIS(M, lambda th: (("mv", 2), smc.hodrick_prescott(λ=th[0], y=y).block()), S.OProduct([S.OUniform(1.0, 2000.0)]), 2, 0.5, seed=3)
    assert g.d == 2 and g.x.shape == (M, 2) and g.Σ.shape == (M, 2, 2)
    ib.smc2(g, y)
    S.o_ibis_init(o_, y)
    n = 0
    for t in range(1, T):
        ib.smc2_step(g, y, t, verbose=False)
        S.o_ibis_step(o_, y, t)
        assert g.rejuvenated == o_.rejuvenated
        n += g.rejuvenated
    assert n >= 1
    np.testing.assert_array_equal(g.θ, o_.theta)
    np.testing.assert_allclose(g.logZ, o_.logZ, rtol=1e-11)
    np.testing.assert_allclose(g.x, o_.x, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(g.Σ, o_.Sigma, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(g.ω, o_.omega, rtol=RTOL, atol=1e-300)


@pytest.mark.gpu
def test_cuda_path_against_committed_golden_vectors(ctx):
    """tests/golden/*.json hold bit patterns written by tools/gen_golden.py from the oracle; here the CUDA path alone (no
    oracle in the loop) must reproduce them: states and ancestors bit for bit, logZ to 1e-10"""
    import json
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden")
    for v in json.load(open(os.path.join(gold, "oracle_vectors.json"))):
        _, y = smc._lib.simulate(v["kind"], v["params"], v["T"], v["data_seed"])
        assert [float(a).hex() for a in y[:4]] == v["y_hex"]
        ctx.set_rng(v["seed"], v["epoch"])
        ctx.record_ancestors(True)
        try:
            z = ctx.log_likelihood(v["kind"], v["params"], v["N"], y, v["resampler"], v["stream"])
            x, _, _ = ctx.fetch_state(want_w=False)
            anc = ctx.fetch_ancestors(v["T"] - 1)
        finally:
            ctx.record_ancestors(False)
        zg = float.fromhex(v["logZ_hex"])
        assert abs(z - zg) <= RTOL * abs(zg)
        assert [float(a).hex() for a in x[:, -1]] == v["x_last_hex"] and float(np.sum(x)).hex() == v["x_sum_hex"]
        assert [int(a) for a in anc[0][:16]] == v["anc_t1_head"]
        assert int(np.sum(anc * (np.arange(v["N"]) + 1)) % (2 ** 61 - 1)) == v["anc_checksum"]
    for v in json.load(open(os.path.join(gold, "widen_vectors.json"))):
        if v["what"] == "guided":
            _, y = smc._lib.simulate(v["kind"], v["params"], v["T"], v["data_seed"])
            prop = np.array([[float.fromhex(c) for c in row] for row in v["prop_hex"]])
            b = ctx.batch(v["kind"], 1, v["N"])
            ctx.set_rng(v["seed"], v["epoch"])
            z = b.log_likelihood(smc._lib.params8(v["params"]).reshape(1, -1), y, v["resampler"], v["stream"], proposal=prop[:, None, :])
            x, _, lw = b.fetch(want_w=False, want_logw=True)
            b.close()
            zg = float.fromhex(v["logZ_hex"])
            assert abs(z[0] - zg) <= RTOL * abs(zg)
            assert float(np.sum(x[0])).hex() == v["x_sum_hex"] and float(np.sum(lw[0])).hex() == v["logw_sum_hex"]
            assert float(x[0, 0, -1]).hex() == v["x_last_hex"]
            if v["resampler"] != smc.MULTINOMIAL:      # the grid-wide guided path takes the sorted resamplers
                ctx.set_rng(v["seed"], v["epoch"])
                ctx.guided_log_likelihood(v["kind"], v["params"], v["N"], y, prop, v["resampler"], v["stream"])
                xs, _, lws = ctx.fetch_state(want_w=False, want_logw=True)
                assert float(np.sum(xs)).hex() == v["x_sum_hex"] and float(np.sum(lws)).hex() == v["logw_sum_hex"]
        else:
            _, y = smc._lib.simulate(0, LG, v["T"], v["data_seed"])
            blk = np.array([float.fromhex(c) for c in v["block_hex"]])
            ll, x, S = ctx.kalman_mv_loglik(v["d"], blk, y, v["matched_init"])
            lg = float.fromhex(v["ll_hex"])
            assert abs(ll[0] - lg) <= 1e-12 * abs(lg)
            np.testing.assert_allclose(x[0], [float.fromhex(c) for c in v["x_hex"]], rtol=1e-11, atol=1e-13)
            np.testing.assert_allclose(S[0].ravel(), [float.fromhex(c) for c in v["S_hex"]], rtol=1e-11, atol=1e-13)


@pytest.mark.gpu
def test_cuda_path_against_committed_golden_vectors_of_round_2(ctx):
    """tests/golden/round2_vectors.json: the two-level multinomial draw of a large cloud (SPEC §5c), the guided UCSV move on both
    engines (§10b) and the multivariate LG particle filter (§4b) — the CUDA path alone against the committed bit patterns"""
    import json
    import os
    for v in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "round2_vectors.json"))):
        zg = float.fromhex(v["logZ_hex"])
        ctx.set_rng(v["seed"], v["epoch"])
        if v["what"] == "two_level_multinomial":
            _, y = smc._lib.simulate(v["kind"], v["params"], v["T"], v["data_seed"])
            ctx.record_ancestors(True)
            try:
                z = ctx.log_likelihood(v["kind"], v["params"], v["N"], y, v["resampler"], v["stream"])
                x, _, _ = ctx.fetch_state(want_w=False)
                anc = ctx.fetch_ancestors(v["T"] - 1)
            finally:
                ctx.record_ancestors(False)
            assert abs(z - zg) <= RTOL * abs(zg) and float(np.sum(x)).hex() == v["x_sum_hex"]
            assert [int(a) for a in anc[0][:16]] == v["anc_t1_head"]
            assert int(np.sum(anc * (np.arange(v["N"]) + 1)) % (2 ** 61 - 1)) == v["anc_checksum"]
        elif v["what"] == "guided_ucsv":
            _, y = smc._lib.simulate(v["kind"], v["params"], v["T"], v["data_seed"])
            kap = np.array([float.fromhex(c) for c in v["kappa_hex"]])
            prop = np.stack([kap, np.zeros(kap.size), np.ones(kap.size)], 1)
            b = ctx.batch(v["kind"], 1, v["N"])
            z = b.log_likelihood(smc._lib.params8(v["params"]).reshape(1, -1), y, v["resampler"], v["stream"], proposal=prop[:, None, :])
            x, _, lw = b.fetch(want_w=False, want_logw=True)
            b.close()
            assert abs(z[0] - zg) <= RTOL * abs(zg)
            assert float(np.sum(x[0])).hex() == v["x_sum_hex"] and float(np.sum(lw[0])).hex() == v["logw_sum_hex"]
            assert [float(a).hex() for a in x[0, :, -1]] == v["x_last_hex"]
            if v["resampler"] != smc.MULTINOMIAL:
                ctx.set_rng(v["seed"], v["epoch"])
                ctx.guided_log_likelihood(v["kind"], v["params"], v["N"], y, prop, v["resampler"], v["stream"])
                xs, _, lws = ctx.fetch_state(want_w=False, want_logw=True)
                assert float(np.sum(xs)).hex() == v["x_sum_hex"] and float(np.sum(lws)).hex() == v["logw_sum_hex"]
        else:
            _, y = smc._lib.simulate(0, LG, v["T"], v["data_seed"])
            blk = np.array([float.fromhex(c) for c in v["block_hex"]])
            z = ctx.log_likelihood(v["kind"], blk, v["N"], y, v["resampler"], v["stream"])
            x, _, _ = ctx.fetch_state(want_w=False)
            assert abs(z - zg) <= RTOL * abs(zg) and float(np.sum(x)).hex() == v["x_sum_hex"]
            assert [float(a).hex() for a in x[:, -1]] == v["x_last_hex"]


@pytest.mark.gpu
def test_smc2_with_guided_inner_filters(ctx, oracle):
    """extension of SMC (not in the reference): every inner filter step guided by the θ-particle's own locally optimal
    proposal (SMC(..., proposal=lg_optimal_proposals)); smc² / smc²! with rejuvenations against the oracle's sampler run
    the same way — θ and clouds bit for bit.  (Host logic also covered on CPU ranks: test_sharded_sampler_logic_over_gloo.)"""
    from oracle import samplers as S
    N, M, T, chain = 64, 48, 40, 2
    _, y = oracle.simulate(0, LG, T, 1998)
    pg = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.LogNormal()])
    po = S.OProduct([S.OTruncatedNormal(0, 1, -1, 1), S.OLogNormal(), S.OLogNormal()])
    g = smc.SMC(N, M, lambda θ: smc.StateSpaceModel(smc.LinearGaussian(θ[0], 1.0, θ[1], θ[2], 0.0), (1, 1)), pg, chain, 0.5, seed=11,
                resampler="systematic", ctx=ctx, proposal=smc.lg_optimal_proposals)
    ref = S.OSMC(N, M, lambda θ: (0, [θ[0], 1.0, θ[1], θ[2], 0.0, 1.0]), po, chain, 0.5, seed=11, resampler=oracle.SYSTEMATIC,
                 proposal=smc.lg_optimal_proposals)
    smc.smc2(g, y)
    S.o_smc2(ref, y)
    n = 0
    for t in range(1, T):
        smc.smc2_step(g, y, t, verbose=False)
        S.o_smc2_step(ref, y, t)
        assert g.rejuvenated == ref.rejuvenated
        n += g.rejuvenated
    assert n >= 1
    np.testing.assert_array_equal(g.θ, ref.theta)
    np.testing.assert_allclose(g.logZ, ref.logZ, rtol=RTOL, atol=0)
    np.testing.assert_allclose(g.ω, ref.omega, rtol=1e-9, atol=1e-300)
    np.testing.assert_array_equal(g.x, ref.x)
    g.close()
