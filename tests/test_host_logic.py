"""CPU tests of the host-side θ-level logic: priors, proposal kernel, the cloud-exchange plan of a
θ-resample and the torch.distributed plumbing (gloo, world_size 2) that the multi-GPU path uses."""
import math
import os
import socket

import numpy as np
import pytest

import sequential_monte_carlo_b200 as smc
from sequential_monte_carlo_b200 import smc_samplers as ss


def test_priors_match_oracle_restatement(oracle):
    from oracle import samplers as S
    pr = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.Uniform(0, 2), smc.Normal(3, 2)])
    po = S.OProduct([S.OTruncatedNormal(0, 1, -1, 1), S.OLogNormal(), S.OUniform(0, 2), S.ONormal(3, 2)])
    a, b = pr.sample(300, 1998), po.sample(300, 1998)
    np.testing.assert_array_equal(a, b)
    assert np.all(np.abs(a[:, 0]) <= 1) and np.all(a[:, 1] > 0) and np.all((a[:, 2] >= 0) & (a[:, 2] <= 2))
    for th in a[:20]:
        assert pr.logpdf(th) == po.logpdf(th) and pr.insupport(th)
    assert not pr.insupport([1.5, 1.0, 1.0, 0.0]) and pr.logpdf([0.0, -1.0, 1.0, 0.0]) == -math.inf
    # densities integrate to one (trapezoid) — pins the closed forms
    for d, lo, hi in ((smc.TruncatedNormal(0.3, 0.7, -1, 1), -1, 1), (smc.LogNormal(0.2, 0.5), 1e-9, 60), (smc.Normal(3, 2), -20, 26)):
        g = np.linspace(lo, hi, 200001)
        assert np.trapezoid(np.exp([d.logpdf(v) for v in g]), g) == pytest.approx(1.0, abs=1e-6)


def test_random_walk_kernel(oracle):
    from oracle import samplers as S
    rng = np.random.default_rng(0)
    th = rng.normal(size=(200, 3)) * [0.1, 1.0, 3.0]
    Sg, uni = ss.random_walk_kernel(th)
    So, unio = S.o_random_walk_kernel(th)
    np.testing.assert_array_equal(Sg, So)
    assert not uni and not unio
    np.testing.assert_allclose(Sg, 2.83 ** 2 / 3 * np.cov(th.T) + 1e-10 * np.eye(3), rtol=1e-12)      # smc_samplers.jl:97-98
    Sg, _ = ss.random_walk_kernel(np.ones((50, 2)))
    np.testing.assert_array_equal(Sg, 1e-2 * np.eye(2))                                               # degenerate branch
    S1, uni = ss.random_walk_kernel(th[:, :1])
    assert uni and S1[0, 0] == pytest.approx(2.83 ** 2 * np.var(th[:, 0], ddof=1) + 1e-10)            # :87-92


class _NumpyStore:
    """clouds as rows of a numpy array; what _BatchStore does on the device"""

    def __init__(self, rows, torch):
        self.rows, self.torch, self.nbytes = rows.copy(), torch, rows.shape[1] * 8

    def pack(self, slots):
        return self.torch.from_numpy(self.rows[slots].copy().view(np.uint8).reshape(-1))

    def gather(self, local_parents):
        self.rows = self.rows[local_parents].copy()

    def unpack(self, slots, buf):
        self.rows[slots] = buf.numpy().view(np.float64).reshape(len(slots), -1)


class _FakeComm:
    """in-process stand-in that runs all ranks' plans against each other"""

    def __init__(self, rank, world, mailbox):
        self.rank, self.world, self.mailbox = rank, world, mailbox


@pytest.mark.parametrize("world", [2, 4, 8])
def test_exchange_plan_is_consistent(world):
    rng = np.random.default_rng(world)
    M = 8 * world
    for _ in range(20):
        parents = np.sort(rng.integers(0, M, M)) if rng.random() < 0.5 else rng.integers(0, M, M)
        plans = [ss.exchange_plan(parents, r, world) for r in range(world)]
        Mloc = M // world
        for r, (lp, send, recv) in enumerate(plans):
            assert lp.shape == (Mloc,) and lp.min() >= 0 and lp.max() < Mloc
            for dst, slots in send.items():
                assert dst != r and len(slots) == len(plans[dst][2][r])        # what r sends to dst, dst expects from r
            for src, slots in recv.items():
                assert src != r and len(slots) == len(plans[src][1][r])
        # emulate: global cloud id = its value; after redistribution slot m must hold parents[m]
        clouds = [np.arange(r * Mloc, (r + 1) * Mloc, dtype=np.float64) for r in range(world)]
        packed = {(r, dst): clouds[r][slots].copy() for r, (lp, send, recv) in enumerate(plans) for dst, slots in send.items()}
        new = []
        for r, (lp, send, recv) in enumerate(plans):
            c = clouds[r][lp].copy()
            for src, slots in recv.items():
                c[slots] = packed[(src, r)]
            new.append(c)
        np.testing.assert_array_equal(np.concatenate(new), parents.astype(np.float64))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = ss.TorchComm()
        assert comm.device.type == "cpu" and comm.rank == rank and comm.world == world
        M, width = 12, 5
        Mloc = M // world
        g = comm.all_gather(np.arange(Mloc, dtype=np.float64) + 100 * rank)         # the replicated M-vectors
        assert g.shape == (M,) and g[Mloc] == 100.0 * (1 if world > 1 else 0)
        g2 = comm.all_gather(np.full((Mloc, 3), float(rank)))
        assert g2.shape == (M, 3) and g2[-1, 0] == world - 1
        rows = (np.arange(rank * Mloc, (rank + 1) * Mloc, dtype=np.float64)[:, None] * 10 + np.arange(width)[None, :])
        store = _NumpyStore(rows, torch)
        parents = np.array([11, 0, 0, 7, 7, 7, 1, 2, 6, 6, 5, 0])
        moved = ss.redistribute(store, parents, comm)
        want = parents[rank * Mloc:(rank + 1) * Mloc, None] * 10.0 + np.arange(width)[None, :]
        np.testing.assert_array_equal(store.rows, want)
        out.put((rank, int(moved)))
    finally:
        dist.destroy_process_group()


def test_sharded_redistribution_over_gloo():
    """world_size 2, gloo: all-gather of the M-vectors + point-to-point cloud moves after a θ-resample."""
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    out = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_gloo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(out.get(timeout=5) for _ in range(2))
    assert got[0] == 4 and got[1] == 4   # clouds whose parent lives on the other rank


def _sampler_worker(rank, world, port, out, algo):
    """the PRODUCT's θ-sharded sampler logic on CPU ranks (oracle-backed fake device, tests/fake_device.py) against the
    single-process oracle sampler"""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as o, samplers as S
        from tests.fake_device import FakeContext
        N, M, T, chain = 48, 16, 30, 2
        _, y = o.simulate(0, [0.5, 1.0, 0.9, 0.8, 0.0, 1.0], T, 1998)
        pg = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.LogNormal()])
        po = S.OProduct([S.OTruncatedNormal(0, 1, -1, 1), S.OLogNormal(), S.OLogNormal()])
        guided = ss.lg_optimal_proposals if algo.endswith("_guided") else None     # guided inner filters (extension, SPEC §10)
        g = smc.SMC(N, M, lambda θ: smc.StateSpaceModel(smc.LinearGaussian(θ[0], 1.0, θ[1], θ[2], 0.0), (1, 1)), pg, chain, 0.5,
                    seed=5, ctx=FakeContext(5), comm=ss.TorchComm(), proposal=guided)
        ref = S.OSMC(N, M, lambda θ: (0, [θ[0], 1.0, θ[1], θ[2], 0.0, 1.0]), po, chain, 0.5, seed=5, proposal=guided)
        n_rejuv = 0
        if algo.startswith("smc2"):
            smc.smc2(g, y)
            S.o_smc2(ref, y)
            for t in range(1, T):
                smc.smc2_step(g, y, t, verbose=False)
                S.o_smc2_step(ref, y, t)
                assert g.rejuvenated == ref.rejuvenated
                n_rejuv += g.rejuvenated
        else:
            smc.density_tempered(g, y, verbose=False)
            S.o_density_tempered(ref, y)
            n_rejuv = len(g.schedule) - 1
        assert n_rejuv >= 1
        np.testing.assert_array_equal(g.θ, ref.theta)                       # replicated on every rank
        np.testing.assert_allclose(g.logZ, ref.logZ, rtol=1e-12, atol=0)
        np.testing.assert_allclose(g.ω, ref.omega, rtol=1e-10, atol=1e-300)
        lo, hi = rank * (M // world), (rank + 1) * (M // world)
        np.testing.assert_array_equal(g._cur.x, ref.x[lo:hi])               # this rank's clouds == the oracle's slots [lo, hi)
        np.testing.assert_array_equal(g._cur.lw, ref.logw[lo:hi])
        out.put((rank, int(g.stats["clouds_moved"]), n_rejuv))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("algo", ["smc2", "density_tempered", "smc2_guided", "density_tempered_guided"])
def test_sharded_sampler_logic_over_gloo(algo, oracle):
    """world_size 2, gloo: smc² / smc²! and density_tempered of the product run θ-sharded on two CPU ranks over an
    oracle-backed fake device and land on the single-process oracle sampler bit for bit (θ, clouds) — the sharding by
    GLOBAL θ index, the replicated control flow and the cloud exchange after every θ-resample"""
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    out = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_sampler_worker, args=(r, 2, port, out, algo)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    got = [out.get(timeout=5) for _ in range(2)]
    assert {g[0] for g in got} == {0, 1} and got[0][2] == got[1][2] >= 1
    print("clouds received per rank:", sorted(got))
    assert sum(g[1] for g in got) >= 1        # at least one cloud crossed ranks (sorted θ-ancestors keep most parents local)


@pytest.mark.parametrize("which", ["lg", "hodrick_prescott"])
def test_ibis_host_logic_on_the_fake_device(oracle, which):
    """ibis.py (IBIS constructor, smc², smc²!, resample!, rejuvenate!: ibis.jl:26-189) driven on the CPU over the
    oracle-backed fake device, scalar and matrix Kalman inner filters, against the oracle's IBIS step by step"""
    from oracle import samplers as S
    from sequential_monte_carlo_b200 import ibis as ib
    from tests.fake_device import FakeContext
    if which == "lg":
        T, M = 50, 64
        _, y = oracle.simulate(0, [0.5, 1.0, 0.9, 0.8, 0.0, 1.0], T, 1998)
        pg = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.LogNormal()])
        po = S.OProduct([S.OTruncatedNormal(0, 1, -1, 1), S.OLogNormal(), S.OLogNormal()])
        g = smc.IBIS(M, lambda θ: smc.LinearGaussian(θ[0], 1.0, θ[1], θ[2], 0.0), pg, 3, 0.5, seed=4, ctx=FakeContext(4))
        ref = S.OIBIS(M, lambda θ: (0, [θ[0], 1.0, θ[1], θ[2], 0.0, 1.0]), po, 3, 0.5, seed=4)
    else:
        rng = np.random.default_rng(0)
        T, M = 60, 48
        y = np.cumsum(np.cumsum(rng.normal(0, 0.05, T))) + rng.normal(0, 1, T)
        g = smc.IBIS(M, lambda θ: smc.hodrick_prescott(λ=θ[0], y=y), smc.product_distribution([smc.Uniform(1.0, 2000.0)]), 2, 0.5,
                     seed=3, ctx=FakeContext(3))
        ref = S.OIBIS(M, lambda th: (("mv", 2), smc.hodrick_prescott(λ=th[0], y=y).block()), S.OProduct([S.OUniform(1.0, 2000.0)]), 2, 0.5, seed=3)
        assert g.d == 2 and g.x.shape == (M, 2) and g.Σ.shape == (M, 2, 2)
    ib.smc2(g, y)
    S.o_ibis_init(ref, y)
    n = 0
    for t in range(1, T):
        ib.smc2_step(g, y, t, verbose=False)
        S.o_ibis_step(ref, y, t)
        assert g.rejuvenated == ref.rejuvenated
        n += g.rejuvenated
    assert n >= 1
    np.testing.assert_array_equal(g.θ, ref.theta)
    np.testing.assert_array_equal(g.logZ, ref.logZ)
    np.testing.assert_array_equal(g.x, ref.x)
    np.testing.assert_array_equal(g.Σ, ref.Sigma)
    np.testing.assert_array_equal(g.ω, ref.omega)
    np.testing.assert_allclose(ib.expected_parameters(g), (ref.theta * ref.omega[:, None]).sum(axis=0)[:, None], rtol=1e-12)
    # observation_dist / estimated_trend / quantile(ibis, p)  (plotting_utils.jl:96-137), restated per θ-particle
    ymix = Smix = 0.0
    for m in range(M):
        mod = g.model(g.θ[m])
        if which == "lg":
            ym, Sm = mod.B * g.x[m], mod.B * g.Σ[m] * mod.B + mod.R
        else:
            ym, Sm = (mod.B @ g.x[m])[0], (mod.B @ g.Σ[m] @ mod.B.T)[0, 0] + mod.R[0]
        ymix, Smix = ymix + g.ω[m] * ym, Smix + g.ω[m] * Sm
    yo, So = ib.observation_dist(g)
    assert yo == pytest.approx(ymix, rel=1e-12) and So == pytest.approx(Smix, rel=1e-12)
    assert smc.estimated_trend(g) == yo
    q = smc.quantile(g, [0.95, 0.5, 0.05])
    assert q[1] == pytest.approx(yo, abs=1e-12) and q[2] - q[1] == pytest.approx(1.6448536269514722 * np.sqrt(So), rel=1e-9) and q[0] < q[1] < q[2]


@pytest.mark.parametrize("guided", [False, True])
def test_exchange_host_logic_on_the_fake_device(oracle, guided):
    """exchange! (smc_samplers.jl:163-189) on the CPU stand-in: with min_ar above any acceptance rate N doubles after a
    rejuvenation, every θ is re-filtered with the doubled cloud and ω ∝ exp(new logZ − logZ)"""
    from tests.fake_device import FakeContext
    N, M, T = 32, 16, 30
    _, y = oracle.simulate(0, [0.5, 1.0, 0.9, 0.8, 0.0, 1.0], T, 1998)
    pg = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.LogNormal()])
    prop = ss.lg_optimal_proposals if guided else None
    g = smc.SMC(N, M, lambda θ: smc.LinearGaussian(θ[0], 1.0, θ[1], θ[2], 0.0), pg, 1, 0.9, 2.0, seed=2, ctx=FakeContext(2), proposal=prop)
    smc.smc2(g, y)
    for t in range(1, T):
        smc.smc2_step(g, y, t, verbose=False)
        if g.rejuvenated:
            break
    assert g.rejuvenated and g.N == 64 and g.x.shape == (M, 1, 64)
    assert np.isfinite(g.logZ).all() and abs(g.ω.sum() - 1) < 1e-12
    # the doubled clouds: a fresh 64-particle filter over y[:t] on the new batch's Philox identity, then this call's own step to y[t]
    if guided:
        pr = np.stack([ss.lg_optimal_proposals(g._P, yt) for yt in y[:t]])
        _, x, lw = oracle.batch_guided_log_likelihood(0, g._P, None, 64, y[:t], 0, pr, g._cur.seed, g._cur.epoch, 0)
    else:
        _, x, lw = oracle.batch_log_likelihood(0, g._P, None, 64, y[:t], 0, g._cur.seed, g._cur.epoch, 0)
    pt = ss.lg_optimal_proposals(g._P, y[t])
    for m in range(M):
        if guided:
            oracle.guided_step(0, g._P[m], x[m], lw[m], y[t], t, 0, pt[m], g._cur.seed, g._cur.epoch, m)
        else:
            oracle.bootstrap_step(0, g._P[m], x[m], lw[m], y[t], t, 0, g._cur.seed, g._cur.epoch, m)
    np.testing.assert_array_equal(g._cur.x, x)
    assert g._cur.t == t
    smc.smc2_step(g, y, t + 1, verbose=False)
    assert g._cur.t == t + 1


def test_inflation_example_flow_on_the_fake_device(oracle):
    """examples/inflation_example.py (the reference's examples/inflation_example.jl on this library): the SMC² loop with
    per-step quartile bands and variances, UC and UC-SV models, at toy sizes on the CPU stand-in"""
    import importlib.util
    from tests.fake_device import FakeContext
    spec = importlib.util.spec_from_file_location("inflation_example", os.path.join(os.path.dirname(os.path.dirname(__file__)), "examples",
                                                                                     "inflation_example.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    _, y = oracle.simulate(2, [0.2, 0.2, 3.0, 1.0, 1.0], 25, 1998)
    for model, prior, dθ in ((ex.uc_mod, ex.uc_prior, 3), (ex.ucsv_mod, ex.ucsv_prior, 4)):
        s, xqs, cqs, variances = ex.run_smc2(model, prior, y, 64, 24, 2, seed=1998, ctx=FakeContext(1998))
        assert xqs.shape == (25, 3) and np.all(np.diff(xqs, axis=1) >= 0) and np.all(np.diff(cqs, axis=1) >= 0)
        assert np.all(variances > 0) and smc.expected_parameters(s).shape == (dθ, 1)
        np.testing.assert_allclose(xqs[:, 1] + cqs[:, 1], y, atol=np.abs(xqs[:, 2] - xqs[:, 0]).max())   # trend + cycle ≈ y
        assert abs(s.ω.sum() - 1) < 1e-12 and np.isfinite(s.logZ).all()


def test_readme_particle_filter_loop_on_the_fake_device(oracle):
    """particles.py host logic on the CPU stand-in — the README loop (README.md:33-61): bootstrap_filter, then
    bootstrap_filter! per observation with quantile(x, p) and logZ += logμ; the device-resident handles (lazy copies,
    staleness after a re-initialisation); log_likelihood; particle_filter with and without a proposal"""
    from sequential_monte_carlo_b200 import particles as pt
    from tests.fake_device import FakeContext
    ctx = FakeContext(1998)
    lg_example = smc.StateSpaceModel(smc.LinearGaussian(0.5, 1.0, 0.9, 0.8, 0.0), (1, 1))
    _, y = smc.simulate(lg_example, 30)
    ctx.set_rng(1998, 3)
    x, w, logμ = smc.bootstrap_filter(256, y[0], lg_example, ctx=ctx)
    xq, logZ = [smc.quantile(x, [0.25, 0.5, 0.75])], logμ
    for t in range(1, len(y)):
        logμ, w, ess = smc.bootstrap_filter_(x, w, y[t], lg_example)
        assert 1.0 <= ess <= 256.0
        xq.append(smc.quantile(x, [0.25, 0.5, 0.75]))
        logZ += logμ
    ref = oracle.log_likelihood(0, lg_example.params(), 256, y, oracle.MULTINOMIAL, 1998, 3, 0)
    assert logZ == pytest.approx(ref["logZ"], rel=1e-12)
    np.testing.assert_array_equal(np.asarray(x), ref["x"][0])
    np.testing.assert_allclose(np.asarray(w), oracle.normalize(ref["logw"])[1], rtol=1e-12)
    assert len(x) == 256 and x.shape == (256,) and np.all(np.diff(np.array(xq), axis=1) >= 0)
    m, v = smc.weighted_mean_var(x, w)
    assert m == pytest.approx(float(ref["x"][0] @ np.asarray(w)), rel=1e-9) and v > 0
    ctx.set_rng(1998, 3)
    x2, w2, logZ2 = smc.log_likelihood(256, y, lg_example, ctx=ctx)           # the whole series in one call
    assert logZ2 == pytest.approx(logZ, rel=1e-12)
    with pytest.raises(RuntimeError):                                          # the first cloud was replaced on this context ...
        smc.bootstrap_filter_(x, w, y[0], lg_example)
    assert np.asarray(x).shape == (256,)                                       # ... its last host copy stays readable
    np.testing.assert_array_equal(np.asarray(x2), ref["x"][0])
    # particle_filter: proposal = nothing is the bootstrap filter; a guided cloud above the batched engine's size continues on the single filter
    pt._GUIDED_BATCH_MAX, keep = 128, pt._GUIDED_BATCH_MAX
    try:
        ctx.set_rng(7, 1)
        xg, wg, _ = smc.particle_filter(256, y[0], lg_example, smc.locally_optimal_proposal, ctx=ctx)
        xo, lwo = oracle.bootstrap_init(0, lg_example.params(), 256, y[0], 7, 1, 0)
        for t in range(1, 6):
            _, wg, _ = smc.particle_filter_(xg, wg, y[t], lg_example, smc.locally_optimal_proposal, resampler="systematic")
            oracle.guided_step(0, lg_example.params(), xo, lwo, y[t], t, oracle.SYSTEMATIC, smc.locally_optimal_proposal(lg_example, y[t]), 7, 1, 0)
        np.testing.assert_array_equal(np.asarray(xg), xo[0])
        _, wg, _ = smc.particle_filter_(xg, wg, y[6], lg_example, None, resampler="systematic")     # bootstrap step on the same cloud
        oracle.bootstrap_step(0, lg_example.params(), xo, lwo, y[6], 6, oracle.SYSTEMATIC, 7, 1, 0)
        np.testing.assert_array_equal(np.asarray(xg), xo[0])
    finally:
        pt._GUIDED_BATCH_MAX = keep


def test_model_constructors():
    m = smc.StateSpaceModel(smc.LinearGaussian(0.5, 1.0, 0.9, 0.8, 0.0), (1, 1))       # README.md:12-15
    assert m.params() == [0.5, 1.0, 0.9, 0.8, 0.0, 1.0] and m.kind == smc.KIND_LG1D
    u = smc.UnivariateLinearGaussian(A=0.5, B=1.0, Q=0.9, R=0.8)                        # state_space_models.jl:74-77
    assert u.params() == m.params()
    uc = smc.unobserved_components(0.3, 0.7, 2.0)                                       # :119-128
    assert uc.params() == [1.0, 1.0, 0.3, 0.7, 2.0, 0.3]
    assert smc.UC(2.0, 0.3, 0.7).params() == uc.params()                                # UC(θ...): level first (example prior order, :33-37)
    v = smc.StateSpaceModel(smc.UCSV(0.2, 3.0, (1.0, 0.5)), (3, 1))                     # examples/inflation_example.jl:229-232
    assert v.params() == [0.2, 0.2, 3.0, 1.0, 0.5] and v.state_dim == 3
    w = smc.unobserved_components_stochastic_volatility(x0=3.0, γε=0.1, γη=0.2, log_σε=1.0, log_ση=0.5)
    assert w.params() == [0.1, 0.2, 3.0, 1.0, 0.5]
    with pytest.raises(ValueError):
        smc.StateSpaceModel(smc.UCSV(0.2, 3.0, (1.0, 0.5)), (1, 1))
    with pytest.raises(NotImplementedError):
        smc.LinearGaussian(np.eye(2), np.ones((1, 2)), np.eye(2), 1.0)
    x, y = smc.simulate(v, 50, seed=3)
    assert x.shape == (50, 3) and y.shape == (50,)
    x2, y2 = smc.simulate(3, v, 50)                                                      # simulate(rng, model, T)  :11
    np.testing.assert_array_equal(y2, y)
    assert smc.simulate(np.random.default_rng(0), m, 7)[1].shape == (7,)


def test_exported_model_methods_describe_what_the_device_evaluates(oracle):
    """transition / observation / initial_dist (exported by state_space_models.jl:1) against the oracle's functors: the
    log-weight of every initial particle is logpdf(observation(model, x_i), y), and one oracle transition step has the
    mean and standard deviation transition(model, x) states"""
    y = 0.37
    models = {0: smc.LinearGaussian(0.5, 1.3, 0.9, 0.8, 0.2, 1.5), 1: smc.SV(-1.0, 0.9, 0.3), 2: smc.UCSV((0.2, 0.3), 3.0, (1.0, 0.5))}
    for kind, m in models.items():
        x, lw = oracle.bootstrap_init(kind, m.params(), 64, y, 3, 0, 0)
        z = np.stack([oracle.normals(3, 0, 0, 0, oracle.P_INIT, k, 64) for k in range(m.state_dim)])
        i0 = smc.initial_dist(m)
        i0 = i0 if isinstance(i0, tuple) else (i0,)
        for k, dist in enumerate(i0):
            np.testing.assert_allclose(x[k], dist.μ + dist.σ * z[k], rtol=1e-13, atol=1e-15)
        for i in range(0, 64, 7):
            xi = x[:, i] if m.state_dim > 1 else x[0, i]
            assert smc.observation(m, xi).logpdf(y) == pytest.approx(lw[i], rel=1e-12, abs=1e-13)
        x1, lw1 = x.copy(), lw.copy()
        a = oracle.bootstrap_step(kind, m.params(), x1, lw1, y, 1, oracle.SYSTEMATIC, 3, 0, 0)
        zt = np.stack([oracle.normals(3, 0, 0, 1, oracle.P_TRANS, k, 64) for k in range(m.state_dim)])
        for i in range(0, 64, 9):
            xp = x[:, a[i]] if m.state_dim > 1 else x[0, a[i]]
            tr = smc.transition(m, xp)
            tr = tr if isinstance(tr, tuple) else (tr,)
            for k, dist in enumerate(tr):
                assert x1[k, i] == pytest.approx(dist.μ + dist.σ * zt[k, i], rel=1e-12, abs=1e-14)
    hp = smc.hodrick_prescott(λ=1600.0, y=np.array([1.0, 1.5, 2.5]))
    tr = smc.transition(hp, [2.0, 1.0])
    np.testing.assert_array_equal(tr.μ, [3.0, 2.0])                       # (2x[t-1] - x[t-2], x[t-1])
    assert smc.observation(hp, [2.0, 1.0]).μ == 2.0 and smc.initial_dist(hp).Σ[0, 0] == 1000.0
    mv = smc.MvNormal([0.0, 0.0], np.eye(2))
    assert mv.logpdf([0.0, 0.0]) == pytest.approx(-np.log(2 * np.pi))


def test_particle_filter_refuses_what_it_cannot_run_on_the_device():
    """particle_filter / particle_filter! (particles.jl:28-84) exist with the reference's signature.  Guided moves are
    built for the affine-Gaussian family on the one-dimensional models (docs/SPEC.md §10); anything else is refused
    loudly rather than run as a bootstrap filter — before any device work."""
    ucsv = smc.UCSV(0.2, 3.0, (1.0, 1.0))
    with pytest.raises(NotImplementedError):
        smc.particle_filter(16, 0.1, ucsv, proposal=smc.AffineGaussianProposal(0.0, 1.0, 1.0), ctx=object())
    m = smc.StateSpaceModel(smc.LinearGaussian(0.5, 1.0, 0.9, 0.8, 0.0), (1, 1))
    with pytest.raises(RuntimeError):      # a guided step continues a cloud made by particle_filter(..., proposal)
        smc.particle_filter_(None, None, 0.1, m, proposal=smc.locally_optimal_proposal)
    from sequential_monte_carlo_b200.particles import _proposal_coefficients
    with pytest.raises(TypeError):         # an arbitrary closure returning a distribution cannot run on the device
        _proposal_coefficients(lambda model, y: (0.0, 1.0), m, 0.1)
    np.testing.assert_allclose(_proposal_coefficients(smc.AffineGaussianProposal(1, 2, 3), m, 0.1), [1, 2, 3])


def test_exchange_plan_of_the_library_equals_the_python_plan():
    """smcb_exchange_plan (host-only C ABI entry, used by the device-resident sampler) against exchange_plan above"""
    from sequential_monte_carlo_b200 import _lib
    rng = np.random.default_rng(0)
    for _ in range(200):
        G, Mloc = int(rng.choice([1, 2, 4, 8])), int(rng.integers(1, 9))
        M = G * Mloc
        a = np.sort(rng.integers(0, M, M))
        for r in range(G):
            lp, send, recv = ss.exchange_plan(a, r, G)
            lp2, s2, r2 = _lib.exchange_plan(a, r, G)
            local = a[r * Mloc:(r + 1) * Mloc] // Mloc == r
            np.testing.assert_array_equal(lp[local], lp2[local])
            np.testing.assert_array_equal(lp2[~local], np.arange(Mloc)[~local])       # remote parents: the slot itself (overwritten by the unpack)
            assert [(d, sl) for d in sorted(send) for sl in send[d]] == s2
            assert [(d, sl) for d in sorted(recv) for sl in recv[d]] == r2
    # every cloud that crosses ranks is sent exactly once and received exactly once, in matching order
    a = np.sort(rng.integers(0, 64, 64))
    sent = {(r, d): [] for r in range(4) for d in range(4)}
    for r in range(4):
        _, s2, r2 = _lib.exchange_plan(a, r, 4)
        for d, sl in s2:
            sent[(r, d)].append(sl + r * 16)
    for r in range(4):
        _, _, r2 = _lib.exchange_plan(a, r, 4)
        got = {}
        for src, sl in r2:
            got.setdefault(src, []).append(sl)
        for src, slots in got.items():
            assert [int(a[r * 16 + sl]) for sl in slots] == sent[(src, r)]


def test_device_sampler_descriptors():
    """prior_descriptor / parameter_map: what decides between the device-resident sampler and the host-language control flow"""
    pg = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(0.5, 2.0), smc.Uniform(-1, 3), smc.Normal(3, 2)])
    rows = ss.prior_descriptor(pg)
    assert rows.shape == (4, 8) and list(rows[:, 0]) == [3, 1, 2, 0]
    assert rows[0, 6] == pg.components[0]._logmass and rows[1, 5] == np.log(2.0) and rows[2, 5] == -np.log(4.0) and rows[3, 1] == 3.0

    class Odd(smc.Normal):
        pass
    assert ss.prior_descriptor(smc.product_distribution([Odd(0, 1)])) is None          # unknown family -> host path
    pl = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.LogNormal()])
    kind, src, cst = ss.parameter_map(lambda θ: smc.StateSpaceModel(smc.LinearGaussian(θ[0], 1.0, θ[1], θ[2], 0.0), (1, 1)), pl, 3)
    assert kind == 0 and list(src[:6]) == [0, -1, 1, 2, -1, -1] and list(cst[:6]) == [0, 1, 0, 0, 0, 1]
    pu = smc.product_distribution([smc.Uniform(0, 1), smc.Normal(3, 2), smc.Uniform(0, 2), smc.Uniform(0, 2)])
    kind, src, cst = ss.parameter_map(lambda θ: smc.StateSpaceModel(smc.UCSV(θ[0], θ[1], (θ[2], θ[3])), (3, 1)), pu, 4)
    assert kind == 2 and list(src[:5]) == [0, 0, 1, 2, 3]
    # a closure that transforms θ is not a selection: host path
    assert ss.parameter_map(lambda θ: smc.LinearGaussian(θ[0], 1.0, θ[1] * 2.0, θ[2], 0.0), pl, 3) is None


def test_proposal_arithmetic_is_the_frozen_one(oracle):
    """docs/SPEC.md §11: the library's host routines (used by both product samplers) against the oracle's numpy restatement,
    bit for bit, and against LAPACK / numpy to rounding"""
    from oracle import samplers as S
    from sequential_monte_carlo_b200 import _lib
    rng = np.random.default_rng(3)
    for d in (2, 3, 4, 5):
        th = rng.normal(size=(300, d)) * rng.uniform(0.1, 3.0, d) + rng.normal(size=d)
        Sg, So = _lib.random_walk_sigma(th), S.o_random_walk_kernel(th)[0]
        np.testing.assert_array_equal(Sg, So)
        np.testing.assert_allclose(Sg, (2.83 * 2.83 / d) * np.cov(th.T) + 1e-10 * np.eye(d), rtol=1e-11)
        for scale in (1.5, 1.0, 0.5):
            L = _lib.cholesky_lower(Sg, scale)
            np.testing.assert_array_equal(L, S.o_cholesky(scale * So))
            np.testing.assert_allclose(L, np.linalg.cholesky(scale * Sg), rtol=1e-12, atol=1e-15)
            z = rng.normal(size=(300, d))
            np.testing.assert_array_equal(ss._propose(th, Sg, False, scale, z), S.o_propose(th, S.o_cholesky(scale * So), z))
            np.testing.assert_allclose(ss._propose(th, Sg, False, scale, z), th + z @ np.linalg.cholesky(scale * Sg).T, rtol=1e-12, atol=1e-14)
    with pytest.raises(np.linalg.LinAlgError):
        _lib.cholesky_lower(np.array([[1.0, 2.0], [2.0, 1.0]]))


def test_batch_chunk_plan_of_the_library():
    """smcb_batch_chunk_plan (host-only C ABI entry): when the batched sweep is cut into dynamically scheduled (chunk, θ) units —
    the shapes measured on the B200 (profiles/r2_chunk_auto.jsonl) and the cases that must stay one CTA per θ"""
    from sequential_monte_carlo_b200 import _lib
    plan = _lib.batch_chunk_plan
    # config 5 per GPU at 8 GPUs: 512 UCSV clouds of 4096 particles, 1024-thread CTAs, one per SM: 3.46 waves -> chunks
    k = plan(512, 4096, 240, 1024, 148)
    assert 8 <= k <= 40 and -(-240 // k) >= 6
    # the same at 1, 2, 4 GPUs fills its waves to within 1.2 %: left alone
    assert plan(4096, 4096, 240, 1024, 148) == plan(2048, 4096, 240, 1024, 148) == plan(1024, 4096, 240, 1024, 148) == 0
    # config 4 on one GPU: 1024 SV clouds of 2048 particles, two 512-thread CTAs per SM
    assert plan(1024, 2048, 499, 512, 296) > 0
    # config 3 on one GPU: all 512 CTAs (256 threads) resident, 3 or 4 to an SM whose warps are saturated
    assert plan(512, 1024, 99, 256, 592) > 0
    # a sweep with an `active` mask over more θ than resident CTAs is always chunked (the number of θ that run is known only on the
    # device): the rejuvenation sweeps of config 5 at 1, 2, 4 GPUs; without a mask, or with every CTA resident, the rule above holds
    for M in (4096, 2048, 1024):
        k = plan(M, 4096, 240, 1024, 148, masked=True)
        assert 8 <= k < 240
    assert plan(148, 4096, 240, 1024, 148, masked=True) == 0 and plan(512, 4096, 12, 1024, 148, masked=True) == 0
    # every θ has an SM to itself; a short series; small clouds that do not saturate an SM; a single step
    assert plan(148, 4096, 240, 1024, 148) == 0 and plan(64, 1024, 99, 512, 296) == 0
    assert plan(512, 4096, 12, 1024, 148) == 0
    assert plan(333, 301, 37, 160, 1184) == 0
    assert plan(512, 4096, 1, 1024, 148) == 0
    # chunks are at least 8 steps (more for small clouds) and never the whole series
    for M, N, steps, thr, slots in ((300, 4096, 100, 1024, 148), (600, 1024, 100, 256, 592), (200, 8192, 40, 1024, 148), (5000, 64, 400, 32, 2368)):
        k = plan(M, N, steps, thr, slots)
        assert k == 0 or (max(8, 8192 // N) <= k < steps)
    with pytest.raises(_lib.SMCBError):
        plan(0, 1024, 10, 256, 148)
