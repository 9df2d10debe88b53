"""GPU parity of the θ-level samplers (smc², smc²!, density_tempered, IBIS;
/root/reference/src/smc_samplers.jl, ibis.jl) against the oracle restatement (oracle/samplers.py).

Both sides draw priors, proposals, accept uniforms and θ-ancestors from the same Philox streams, so
θ-particles and every state cloud must come out bit-identical; logZ, ω, ess agree to rel 1e-10."""
import numpy as np
import pytest

import sequential_monte_carlo_b200 as smc
from sequential_monte_carlo_b200 import smc_samplers as ss, ibis as ib

pytestmark = pytest.mark.gpu
RTOL = 1e-9   # ω / ess of M-vectors built from logZ values that themselves agree to 1e-10 relative (|logZ| ~ 1e2)

LG_TRUE = [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]


def lg_mod(θ):                                                        # README.md:75-78
    return smc.StateSpaceModel(smc.LinearGaussian(θ[0], 1.0, θ[1], θ[2], 0.0), (1, 1))


def lg_mod_o(θ):
    return 0, [θ[0], 1.0, θ[1], θ[2], 0.0, 1.0]


def lg_priors():
    from oracle import samplers as S
    return (smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.LogNormal()]),     # README.md:81-85
            S.OProduct([S.OTruncatedNormal(0, 1, -1, 1), S.OLogNormal(), S.OLogNormal()]))


def _check_same(g, o_, oracle, clouds=True):
    np.testing.assert_array_equal(g.θ, o_.theta)
    np.testing.assert_allclose(g.logZ, o_.logZ, rtol=1e-10, atol=0)
    np.testing.assert_allclose(g.ω, o_.omega, rtol=RTOL, atol=1e-300)
    assert abs(g.ess - o_.ess) <= RTOL * o_.ess
    if clouds:
        np.testing.assert_array_equal(g.x, o_.x)
        w = g.w
        for m in range(0, g.M, max(1, g.M // 8)):
            _, wo, _ = oracle.normalize(o_.logw[m])
            np.testing.assert_allclose(w[m], wo, rtol=1e-10, atol=0)


ENGINES = ["device", "host"]   # the device-resident sampler (smcb_sampler_*) and the host-language control flow


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("resampler", ["multinomial", "systematic"])
def test_smc2_lg(ctx, oracle, resampler, engine):
    """BASELINE config 3 shape, scaled down: SMC(N, M, lg_mod, lg_prior, 3, 0.5), smc² then smc²! for t = 2..T."""
    from oracle import samplers as S
    N, M, T, chain = 128, 64, 40, 3
    _, y = oracle.simulate(0, LG_TRUE, T, 1998)
    pg, po = lg_priors()
    g = smc.SMC(N, M, lg_mod, pg, chain, 0.5, seed=11, resampler=resampler, ctx=ctx, engine=engine)
    assert g.engine == engine
    o_ = S.OSMC(N, M, lg_mod_o, po, chain, 0.5, seed=11, resampler=ss.resampler_id(resampler))
    smc.smc2(g, y)
    S.o_smc2(o_, y)
    _check_same(g, o_, oracle)
    n_rejuv = 0
    for t in range(1, T):
        smc.smc2_step(g, y, t, verbose=False)
        S.o_smc2_step(o_, y, t)
        assert g.rejuvenated == o_.rejuvenated
        n_rejuv += g.rejuvenated
        if g.rejuvenated or t == T - 1:
            _check_same(g, o_, oracle)
            assert g.acc_ratio == o_.acc_ratio
    assert n_rejuv >= 2
    np.testing.assert_array_equal(smc.expected_parameters(g), S.o_expected_parameters(o_) * 1.0) if False else None
    np.testing.assert_allclose(smc.expected_parameters(g), S.o_expected_parameters(o_), rtol=1e-9)
    np.testing.assert_allclose(smc.expected_parameters(g, reference_style=True), S.o_expected_parameters(o_, True), rtol=1e-9)
    # estimated_trend / quantile(smc, p) (plotting_utils.jl:116-124,140-157): the per-θ weighted state means come
    # from the device; restated here with numpy on the oracle's clouds
    xm = smc.state_means(g)
    xo = np.stack([(oracle.normalize(o_.logw[m])[1] * o_.x[m]).sum(axis=-1) for m in range(M)])
    np.testing.assert_allclose(xm, xo.reshape(M, -1), rtol=1e-10, atol=1e-13)
    mu, sd = g._P[:, 1] * xo.reshape(M, -1)[:, 0], np.sqrt(g._P[:, 3])
    np.testing.assert_allclose(smc.estimated_trend(g), np.sum(g.ω * mu), rtol=1e-10)
    from statistics import NormalDist
    z = np.array([NormalDist().inv_cdf(v) for v in (0.05, 0.5, 0.95)])
    np.testing.assert_allclose(smc.quantile(g, [0.95, 0.05, 0.5]), (g.ω[:, None] * (mu[:, None] + sd[:, None] * z)).sum(axis=0), rtol=1e-10)
    # per-θ quantile bands (get_quantiles_uc, examples/inflation_example.jl:39-55): radix select per cloud on the
    # device against the oracle's sort + cumulative sum, bit for bit
    ps = [0.25, 0.5, 0.75]
    for weighted in (True, False):
        qg = smc.state_quantiles(g, ps, weighted=weighted)
        assert qg.shape == (M, 1, 3)
        for m in range(0, M, 7):
            np.testing.assert_array_equal(qg[m], oracle.weighted_summary(o_.x[m], o_.logw[m], ps, weighted=weighted)[2])
    xq, cq = smc.get_quantiles(g, float(y[-1]), ps)
    qall = smc.state_quantiles(g, [0.25, 0.5, 0.75, 0.75, 0.5, 0.25])[:, 0, :]
    np.testing.assert_allclose(xq, (g.ω[:, None] * qall[:, :3]).sum(axis=0), rtol=1e-13)
    np.testing.assert_allclose(cq, (g.ω[:, None] * (float(y[-1]) - qall[:, 3:])).sum(axis=0), rtol=1e-13)
    assert np.all(np.diff(xq) >= 0) and np.all(np.diff(cq) >= 0)
    g.close()


@pytest.mark.parametrize("engine", ENGINES)
def test_smc2_ucsv(ctx, oracle, engine):
    """BASELINE config 5 shape, scaled down: the 4-parameter UCSV of examples/inflation_example.jl:229-239."""
    from oracle import samplers as S
    N, M, T, chain = 96, 32, 24, 2
    _, y = oracle.simulate(2, [0.2, 0.2, 3.0, 1.0, 1.0], T, 1998)
    pg = smc.product_distribution([smc.Uniform(0, 1), smc.Normal(3, 2), smc.Uniform(0, 2), smc.Uniform(0, 2)])
    po = S.OProduct([S.OUniform(0, 1), S.ONormal(3, 2), S.OUniform(0, 2), S.OUniform(0, 2)])
    g = smc.SMC(N, M, lambda θ: smc.StateSpaceModel(smc.UCSV(θ[0], θ[1], (θ[2], θ[3])), (3, 1)), pg, chain, 0.5, seed=3, ctx=ctx,
                engine=engine)
    assert g.engine == engine
    o_ = S.OSMC(N, M, lambda θ: (2, [θ[0], θ[0], θ[1], θ[2], θ[3]]), po, chain, 0.5, seed=3)
    smc.smc2(g, y)
    S.o_smc2(o_, y)
    for t in range(1, T):
        smc.smc2_step(g, y, t, verbose=False)
        S.o_smc2_step(o_, y, t)
        assert g.rejuvenated == o_.rejuvenated
    _check_same(g, o_, oracle)
    g.close()


@pytest.mark.parametrize("engine", ENGINES)
def test_density_tempered_lg(ctx, oracle, capsys, engine):
    """BASELINE config 4 algorithm on the README's LG example, scaled down."""
    from oracle import samplers as S
    N, M, T, chain = 128, 96, 50, 3
    _, y = oracle.simulate(0, LG_TRUE, T, 1998)
    pg, po = lg_priors()
    g = smc.SMC(N, M, lg_mod, pg, chain, 0.5, seed=5, ctx=ctx, engine=engine)
    o_ = S.OSMC(N, M, lg_mod_o, po, chain, 0.5, seed=5)
    smc.density_tempered(g, y, verbose=True)
    S.o_density_tempered(o_, y)
    out = capsys.readouterr().out
    assert out.count("ξ = ") == len(g.schedule) and "[rejuvenating]" in out and "acc_rate:" in out   # smc_samplers.jl:207-214 format
    assert len(g.schedule) == len(o_.schedule) >= 3
    for (xg, eg), (xo, eo) in zip(g.schedule, o_.schedule):
        assert abs(xg - xo) <= 1e-12 and abs(eg - eo) <= RTOL * eo
    assert g.schedule[-1][0] == 1.0
    assert all(abs(e - g.ess_min) < 0.5 for _, e in g.schedule[:-1])      # bisection lands on ess_min (docstring trace: 255.99 / 256)
    _check_same(g, o_, oracle)
    g.close()


@pytest.mark.parametrize("engine", ENGINES)
def test_density_tempered_sv(ctx, oracle, engine):
    """BASELINE config 4 model: stochastic volatility, θ = (μ, ρ, σ)."""
    from oracle import samplers as S
    N, M, T = 128, 48, 60
    _, y = oracle.simulate(1, [-1.0, 0.9, 0.3], T, 1998)
    pg = smc.product_distribution([smc.Normal(0, 2), smc.Uniform(-1, 1), smc.LogNormal(-1, 1)])
    po = S.OProduct([S.ONormal(0, 2), S.OUniform(-1, 1), S.OLogNormal(-1, 1)])
    g = smc.SMC(N, M, lambda θ: smc.SV(θ[0], θ[1], θ[2]), pg, 2, 0.5, seed=8, resampler="stratified", ctx=ctx, engine=engine)
    o_ = S.OSMC(N, M, lambda θ: (1, [θ[0], θ[1], θ[2]]), po, 2, 0.5, seed=8, resampler=1)
    smc.density_tempered(g, y, verbose=False)
    S.o_density_tempered(o_, y)
    assert len(g.schedule) == len(o_.schedule)
    _check_same(g, o_, oracle)
    g.close()


@pytest.mark.parametrize("engine", ENGINES)
def test_exchange_doubles_state_particles(ctx, oracle, engine):
    """exchange! (smc_samplers.jl:163-189): with min_ar above any acceptance rate N doubles after a rejuvenation, every θ is
    re-filtered with the doubled cloud and ω ∝ exp(new logZ − logZ) — θ, the doubled clouds, ω and logZ against the oracle's
    restatement (o_exchange) at the doubling step and one step later."""
    from oracle import samplers as S
    N, M, T = 64, 32, 30
    _, y = oracle.simulate(0, LG_TRUE, T, 1998)
    pg, po = lg_priors()
    g = smc.SMC(N, M, lg_mod, pg, 1, 0.9, 2.0, seed=2, ctx=ctx, engine=engine)
    o_ = S.OSMC(N, M, lg_mod_o, po, 1, 0.9, 2.0, seed=2)
    smc.smc2(g, y)
    S.o_smc2(o_, y)
    for t in range(1, T):
        smc.smc2_step(g, y, t, verbose=False)
        S.o_smc2_step(o_, y, t)
        assert g.rejuvenated == o_.rejuvenated
        if g.rejuvenated:
            break
    assert g.rejuvenated and g.N == 128 == o_.N and g.x.shape == (M, 1, 128)
    _check_same(g, o_, oracle)
    assert np.isfinite(g.logZ).all() and abs(g.ω.sum() - 1) < 1e-12
    smc.smc2_step(g, y, t + 1, verbose=False)
    S.o_smc2_step(o_, y, t + 1)
    _check_same(g, o_, oracle)
    g.close()


def test_smc2_config3_shape_against_the_oracle(ctx, oracle):
    """BASELINE config 3 at its own shape: SMC(1024, 512, lg_mod, lg_prior, 3, 0.5), T = 100 (README.md:88-103), the device
    sampler against the OpenMP oracle: same rejuvenation times, θ and clouds bit for bit, logZ to 1e-10."""
    from oracle import samplers as S
    N, M, T, chain = 1024, 512, 100, 3
    _, y = oracle.simulate(0, LG_TRUE, T, 1998)
    pg, po = lg_priors()
    g = smc.SMC(N, M, lg_mod, pg, chain, 0.5, seed=1998, resampler="systematic", ctx=ctx, engine="device")
    o_ = S.OSMC(N, M, lg_mod_o, po, chain, 0.5, seed=1998, resampler=2)
    smc.smc2(g, y)
    S.o_smc2(o_, y)
    n = 0
    for t in range(1, T):
        smc.smc2_step(g, y, t, verbose=False)
        S.o_smc2_step(o_, y, t)
        assert g.rejuvenated == o_.rejuvenated, t
        n += g.rejuvenated
    assert n >= 3
    _check_same(g, o_, oracle)
    m = smc.expected_parameters(g).ravel()
    assert abs(m[0] - 0.5) < 0.25 and 0.4 < m[1] < 1.8 and 0.3 < m[2] < 1.6
    g.close()


def test_density_tempered_config4_shape_reduced_T(ctx, oracle):
    """BASELINE config 4 at its own M × N (1024 θ × 2048 state particles, SV) with T cut to 40 so that the oracle finishes
    in seconds: schedule, θ and clouds against the oracle."""
    from oracle import samplers as S
    N, M, T = 2048, 1024, 40
    _, y = oracle.simulate(1, [-1.0, 0.9, 0.3], T, 1998)
    pg = smc.product_distribution([smc.Normal(0, 2), smc.Uniform(-1, 1), smc.LogNormal(-1, 1)])
    po = S.OProduct([S.ONormal(0, 2), S.OUniform(-1, 1), S.OLogNormal(-1, 1)])
    g = smc.SMC(N, M, lambda θ: smc.SV(θ[0], θ[1], θ[2]), pg, 3, 0.5, seed=4, resampler="systematic", ctx=ctx, engine="device")
    o_ = S.OSMC(N, M, lambda θ: (1, [θ[0], θ[1], θ[2]]), po, 3, 0.5, seed=4, resampler=2)
    smc.density_tempered(g, y, verbose=False)
    S.o_density_tempered(o_, y)
    assert len(g.schedule) == len(o_.schedule) >= 2
    for (xg, eg), (xo, eo) in zip(g.schedule, o_.schedule):
        assert abs(xg - xo) <= 1e-12 and abs(eg - eo) <= RTOL * eo
    _check_same(g, o_, oracle, clouds=False)
    x = g.x
    for m in range(0, M, 97):
        np.testing.assert_array_equal(x[m], o_.x[m])
    g.close()


def test_smc2_config5_shape_reduced_T(ctx, oracle):
    """BASELINE config 5's inner shape (UCSV, 4096 state particles per θ) with M = 256 θ-particles and T = 12: the large
    clouds take the L2 placement of batch_kernel; θ, logZ and sampled clouds against the oracle."""
    from oracle import samplers as S
    N, M, T, chain = 4096, 256, 14, 2
    _, y = oracle.simulate(2, [0.2, 0.2, 3.0, 1.0, 1.0], T, 1998)
    pg = smc.product_distribution([smc.Uniform(0, 1), smc.Normal(3, 2), smc.Uniform(0, 2), smc.Uniform(0, 2)])
    po = S.OProduct([S.OUniform(0, 1), S.ONormal(3, 2), S.OUniform(0, 2), S.OUniform(0, 2)])
    g = smc.SMC(N, M, lambda θ: smc.StateSpaceModel(smc.UCSV(θ[0], θ[1], (θ[2], θ[3])), (3, 1)), pg, chain, 0.9, seed=3,
                resampler="systematic", ctx=ctx, engine="device")
    o_ = S.OSMC(N, M, lambda θ: (2, [θ[0], θ[0], θ[1], θ[2], θ[3]]), po, chain, 0.9, seed=3, resampler=2)
    smc.smc2(g, y)
    S.o_smc2(o_, y)
    n = 0
    for t in range(1, T):
        smc.smc2_step(g, y, t, verbose=False)
        S.o_smc2_step(o_, y, t)
        assert g.rejuvenated == o_.rejuvenated, t
        n += g.rejuvenated
    assert n >= 1
    _check_same(g, o_, oracle, clouds=False)
    x = g.x
    for m in range(0, M, 37):
        np.testing.assert_array_equal(x[m], o_.x[m])
    g.close()


def test_device_sampler_through_ctypes_only(ctx, oracle):
    """the new entry points called directly (no Python sampler class in between): smcb_sampler_create / set_data / smc2_init /
    smc2_step / get / stats / clouds, on a single-rank communicator set up through smcb_comm_init"""
    import ctypes as C
    from oracle import samplers as S
    lib = smc._lib.load()
    r, n = C.c_int(-1), C.c_int(-1)
    assert lib.smcb_comm_rank(ctx._h, C.byref(r), C.byref(n)) == 0 and (r.value, n.value) == (0, 1)
    N, M, T = 64, 32, 30
    _, y = oracle.simulate(0, LG_TRUE, T, 1998)
    pg, po = lg_priors()
    cfg = smc._lib.SamplerConfig()
    cfg.kind, cfg.d_theta, cfg.N, cfg.M, cfg.chain, cfg.resampler, cfg.theta_resampler = 0, 3, N, M, 2, 0, 0
    cfg.ess_threshold, cfg.min_ar, cfg.seed = 0.5, -1.0, 21
    rows = ss.prior_descriptor(pg)
    for k in range(3):
        for j in range(8):
            cfg.prior[k][j] = rows[k, j]
    for k, (src, cst) in enumerate([(0, 0.0), (-1, 1.0), (1, 0.0), (2, 0.0), (-1, 0.0), (-1, 1.0), (-1, 0.0), (-1, 0.0)]):
        cfg.map_src[k], cfg.map_const[k] = src, cst
    θ0 = np.ascontiguousarray(pg.sample(M, 21))
    h = C.c_void_p()
    assert lib.smcb_sampler_create(ctx._h, C.byref(cfg), θ0.ctypes.data_as(C.c_void_p), C.byref(h)) == 0, lib.smcb_last_error(ctx._h)
    assert lib.smcb_sampler_smc2_init(h) == smc._lib.SMCBError(-4, "").code            # no data yet: SMCB_ERR_STATE
    assert lib.smcb_sampler_set_data(h, y.ctypes.data_as(C.c_void_p), T) == 0
    assert lib.smcb_sampler_smc2_init(h) == 0
    o_ = S.OSMC(N, M, lg_mod_o, po, 2, 0.5, seed=21)
    S.o_smc2(o_, y)
    ess, rj, nrj = C.c_double(), C.c_int(), 0
    for t in range(1, T):
        assert lib.smcb_sampler_smc2_step(h, t, C.byref(ess), C.byref(rj)) == 0, lib.smcb_last_error(ctx._h)
        S.o_smc2_step(o_, y, t)
        assert bool(rj.value) == o_.rejuvenated and abs(ess.value - o_.ess) <= RTOL * o_.ess
        nrj += rj.value
    assert nrj >= 1
    assert lib.smcb_sampler_smc2_step(h, T, None, None) == -1                           # t out of range: SMCB_ERR_BAD_ARG
    θ, ω, z = np.empty((M, 3)), np.empty(M), np.empty(M)
    ar, Nn = C.c_double(), C.c_int64()
    assert lib.smcb_sampler_get(h, θ.ctypes.data_as(C.c_void_p), ω.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), C.byref(ess),
                                C.byref(ar), C.byref(Nn)) == 0
    np.testing.assert_array_equal(θ, o_.theta)
    np.testing.assert_allclose(z, o_.logZ, rtol=1e-10)
    np.testing.assert_allclose(ω, o_.omega, rtol=RTOL, atol=1e-300)
    assert ar.value == o_.acc_ratio and Nn.value == N
    b = C.c_void_p()
    assert lib.smcb_sampler_clouds(h, C.byref(b)) == 0
    x = np.empty((M, 1, N))
    assert lib.smcb_batch_fetch(b, x.ctypes.data_as(C.c_void_p), None, None) == 0
    np.testing.assert_array_equal(x, o_.x)
    ms, cnt = (C.c_double * 8)(), (C.c_int64 * 8)()
    assert lib.smcb_sampler_stats(h, ms, cnt) == 0
    assert cnt[1] == T - 1 and cnt[2] == nrj and cnt[0] == 2 * nrj and cnt[4] > 0
    assert lib.smcb_sampler_destroy(h) == 0


def test_posterior_is_plausible(ctx, oracle):
    """Statistical anchor: density-tempered posterior mean on LG data near the truth (0.5, 0.9, 0.8), in the
    range of the reference's docstring run (0.503, 1.025, 0.975; smc_samplers.jl:215-219, unknown data seed)."""
    y = smc.simulate(lg_mod([0.5, 0.9, 0.8]), 200, seed=1998)[1]
    pg, _ = lg_priors()
    g = smc.SMC(512, 512, lg_mod, pg, 3, 0.5, seed=1998, resampler="systematic", ctx=ctx)
    smc.density_tempered(g, y, verbose=False)
    m = smc.expected_parameters(g).ravel()
    assert 3 <= len(g.schedule) <= 12 and 0.05 < g.acc_ratio < 0.6
    assert abs(m[0] - 0.5) < 0.2 and 0.5 < m[1] < 1.6 and 0.4 < m[2] < 1.4
    # the PF-based logZ of the posterior mean agrees with Kalman's (kalman_filter.jl cross-check of the north star)
    ll = smc.kalman_filter.log_likelihood(y, lg_mod(m), matched_init=True, ctx=ctx)
    _, _, z = smc.log_likelihood(16384, y, lg_mod(m), ctx=ctx)
    assert abs(z - ll) < 1.0
    g.close()


def test_ibis(ctx, oracle):
    from oracle import samplers as S
    M, T = 128, 60
    _, y = oracle.simulate(0, LG_TRUE, T, 1998)
    pg, po = lg_priors()
    g = smc.IBIS(M, lg_mod, pg, 3, 0.5, seed=4, ctx=ctx)
    o_ = S.OIBIS(M, lg_mod_o, po, 3, 0.5, seed=4)
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
    np.testing.assert_allclose(g.x, o_.x, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(g.Σ, o_.Sigma, rtol=1e-11)
    np.testing.assert_allclose(g.ω, o_.omega, rtol=RTOL, atol=1e-300)
    with pytest.raises(TypeError):
        smc.IBIS(8, lambda θ: smc.SV(θ[0], θ[1], θ[2]), pg, 1, 0.5, ctx=ctx)


@pytest.mark.parametrize("which", ["lg_dyn", "ucsv_dyn"])
def test_device_sampler_with_dynamic_scheduling_equals_static(ctx, which, monkeypatch):
    """whole sampler runs whose rejuvenation sweeps are cut into dynamically scheduled (chunk, θ) units (more θ than resident CTAs,
    some proposals inactive) against the same runs with one CTA per θ (SMCB_BATCH_CHUNK=0): θ, ω, logZ and every cloud bit for bit"""
    from tests.dist_worker import build_sampler, run
    out = {}
    for mode in ("0", None):
        if mode is None:
            monkeypatch.delenv("SMCB_BATCH_CHUNK", raising=False)
        else:
            monkeypatch.setenv("SMCB_BATCH_CHUNK", mode)
        if which == "lg_dyn":
            s, y, algo = build_sampler(smc, "lg_dyn", ctx, None)
        else:
            pg = smc.product_distribution([smc.Uniform(0, 1), smc.Normal(3, 2), smc.Uniform(0, 2), smc.Uniform(0, 2)])
            model = lambda θ: smc.StateSpaceModel(smc.UCSV(θ[0], θ[1], (θ[2], θ[3])), (3, 1))      # noqa: E731
            y = smc.simulate(model([0.2, 3.0, 1.0, 1.0]), 40, seed=1998)[1]
            s, algo = smc.SMC(4096, 300, model, pg, 2, 0.5, seed=3, resampler="systematic", ctx=ctx, engine="device"), "smc2"
        rejuv = run(smc, s, y, algo)
        out[mode] = (np.array(rejuv), s.θ.copy(), s.ω.copy(), s.logZ.copy(), np.array(s.x), np.array(s.w))
        s.close()
    monkeypatch.delenv("SMCB_BATCH_CHUNK", raising=False)
    assert len(out["0"][0]) >= 2 and out["0"][0].max() >= 16          # at least one rejuvenation long enough to be chunked
    for a, b in zip(out["0"], out[None]):
        np.testing.assert_array_equal(a, b)


def test_reference_docstring_trace_is_a_plausible_draw(ctx):
    """The only output the reference records for this path (smc_samplers.jl:207-219; README.md:88-91): density_tempered with
    512 θ-particles × 1024 state particles, chain 3, ESS 0.5 on lg_mod — six stages ξ = 0.00825, 0.03895, 0.11587, 0.27741,
    0.67719, 1.0, per-stage ESS 255.986 … 256.000, acc_rate 0.157 – 0.211, final ESS 415.0.  The docstring does not say how long
    the series was or which commit printed it, and Julia's global RNG cannot be reproduced, so the pin is
    distributional, with two nuisance parameters settled by experiment (tools/trace_probe.py, profiles/r2_trace_probe.jsonl):
      * T: the first ξ is inversely proportional to the series length (0.08-0.11 at T = 100, 0.013-0.015 at T = 800);
        0.00825 and six stages need T ≈ 1300 — not the README's 100 periods;
      * the proposal scaling: with the on-disk multivariate kernel dθ = 2.83²/d (smc_samplers.jl:97) a rejuvenation moves
        40-60 % of the θ-particles; the recorded 16-21 % is what the UNdivided dθ = 2.83² of the univariate method (:89) gives,
        so the trace predates the division by d.
    With T = 1320 and the kernel field set to the undivided scaling (smc.kernel is a public field of the reference's struct)
    the recorded stage count, first ξ, overall growth of ξ, per-stage ESS, acceptance rates and final ESS have to lie inside
    the range our sampler produces over 10 data / sampler seeds; with the on-disk
    kernel the ξ schedule is the same law and only the acceptance differs (asserted as such).  Parity with the original stays
    "unpinned" in the bit-for-bit sense."""
    from sequential_monte_carlo_b200 import smc_samplers as ss
    ref_xi = [0.00825, 0.03895, 0.11587, 0.27741, 0.67719, 1.0]
    ref_acc = [0.18594, 0.21055, 0.16055, 0.17656, 0.15664]
    pg, _ = lg_priors()

    def undivided(θ):                                                  # random_walk_kernel with dθ = 2.83² for d > 1
        Σ, uni = ss.random_walk_kernel(θ)
        return (Σ if uni else Σ * θ.shape[1]), uni

    def runs(kernel, engine, seeds):
        out = []
        for seed in seeds:
            y = smc.simulate(lg_mod([0.5, 0.9, 0.8]), 1320, seed=1000 + seed)[1]
            g = smc.SMC(1024, 512, lg_mod, pg, 3, 0.5, seed=seed, ctx=ctx, engine=engine)   # on-disk argument order: N, M
            if kernel is not None:
                g.kernel = kernel
            smc.density_tempered(g, y, verbose=False)
            out.append(([s[0] for s in g.schedule], [s[1] for s in g.schedule], list(g.acceptance)))
            g.close()
        return out

    old = runs(undivided, "host", range(10))
    stages = [len(xi) for xi, _, _ in old]
    assert min(stages) <= 6 <= max(stages), stages
    first = [xi[0] for xi, _, _ in old]
    assert min(first) <= ref_xi[0] <= max(first), (min(first), max(first))
    # the schedule grows geometrically: overall ξ₅/ξ₁ of the recorded trace (82) inside the sampled range of the six-stage runs,
    # and its stage-to-stage factors (4.7, 3.0, 2.4, 2.4) within a quarter of ours (its first factor is larger than any of ours:
    # another data set)
    six = [xi for xi, _, _ in old if len(xi) == 6]
    span = [xi[4] / xi[0] for xi in six]
    assert min(span) <= ref_xi[4] / ref_xi[0] <= max(span), (ref_xi[4] / ref_xi[0], min(span), max(span))
    growth = [b / a for xi, _, _ in old for a, b in zip(xi[:-2], xi[1:-1])]   # (the last stage is cut at 1)
    for a, b in zip(ref_xi[:-2], ref_xi[1:-1]):
        assert min(growth) / 1.25 <= b / a <= 1.25 * max(growth), (b / a, min(growth), max(growth))
    assert max(abs(e - 256.0) for _, ess, _ in old for e in ess[:-1]) < 0.05   # the bisection lands on ess_min (trace: 255.986 … 256.000)
    accs = [a for _, _, acc in old for a in acc]
    assert min(accs) <= min(ref_acc) and max(ref_acc) <= max(accs), (min(accs), max(accs))
    fin = [ess[-1] for _, ess, _ in old]
    assert min(fin) <= 415.016 <= max(fin), (min(fin), max(fin))
    # the on-disk kernel (device engine): the same tempering schedule, a rejuvenation moves two to three times as many particles
    new = runs(None, "device", range(4))
    assert all(5 <= len(xi) <= 7 for xi, _, _ in new)
    assert all(0.006 < xi[0] < 0.012 for xi, _, _ in new)
    acc_new = [a for _, _, acc in new for a in acc]
    assert min(acc_new) > max(ref_acc) and 0.3 < np.median(acc_new) < 0.7, (min(acc_new), max(acc_new))
