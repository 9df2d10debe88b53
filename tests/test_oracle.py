"""CPU tests of the oracle itself (no GPU): what pins it, given that the reference ships no tests
or golden vectors (SURVEY.md §4, §8c — parity unpinned): Philox known answers, libm agreement of
the deterministic math, closed-form normalize cases, an independent numpy restatement of the
resampler, the Kalman likelihood on LG models, and the committed regression vectors."""
import json
import math
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
LG = [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]


def ulp_diff(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.spacing(np.maximum(np.abs(b), np.finfo(float).tiny))


def test_philox_known_answers(oracle):
    for v in json.load(open(os.path.join(GOLD, "philox_kat.json"))):
        out = oracle.philox(v["ctr"], v["key"])
        assert ["%08x" % int(w) for w in out] == v["out"]


def test_detmath_vs_libm(oracle):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-700, 700, 20000), rng.uniform(-2, 2, 20000), [0.0, -0.0, 1.0, -1.0]])
    assert ulp_diff(oracle.det_exp(x), np.exp(x)).max() <= 2
    assert oracle.det_exp(np.array([-701.0, -np.inf]))[0] == 0.0 and oracle.det_exp(np.array([-np.inf]))[0] == 0.0
    assert np.isinf(oracle.det_exp(np.array([701.0]))[0])
    u = np.concatenate([rng.random(20000), 2.0 ** -rng.uniform(1, 53, 20000), [2.0 ** -53, 1 - 2.0 ** -53, 0.5, 1.0]])
    assert ulp_diff(oracle.det_log(u), np.log(u)).max() <= 2
    v = rng.random(40000)
    s, c = oracle.det_sincos2pi(v)
    # absolute error: near the zeros of sin/cos the reduced argument is exact, the result tiny
    # (np.sin(2*pi*v) itself carries up to 2*pi*v*2^-53 = 7e-16 of argument-rounding error)
    assert np.abs(s - np.sin(2 * np.pi * v)).max() < 1.2e-15 and np.abs(c - np.cos(2 * np.pi * v)).max() < 1.2e-15
    assert np.abs(s * s + c * c - 1).max() < 5e-16
    s, c = oracle.det_sincos2pi(np.array([0.0, 0.25, 0.5, 0.75]))
    np.testing.assert_array_equal(s, [0.0, 1.0, -0.0, -1.0])
    np.testing.assert_array_equal(c, [1.0, 0.0, -1.0, -0.0])


def test_detmath_golden_bits(oracle):
    g = json.load(open(os.path.join(GOLD, "detmath_vectors.json")))
    for name, fn in (("exp", oracle.det_exp), ("log", oracle.det_log)):
        xs = np.array([float.fromhex(k) for k in g[name]])
        assert [float(v).hex() for v in fn(xs)] == list(g[name].values())
    us = np.array([float.fromhex(k) for k in g["sin2pi"]])
    s, c = oracle.det_sincos2pi(us)
    assert [float(v).hex() for v in s] == list(g["sin2pi"].values())
    assert [float(v).hex() for v in c] == list(g["cos2pi"].values())
    assert [float(v).hex() for v in oracle.normals(1998, 1, 2, 3, 2, 0, 8)] == g["normals_head"]
    assert [int(v) for v in oracle.uniforms64(1998, 1, 2, 3, 3, 8)] == g["uniforms_head"]


def test_normals_are_standard_normal(oracle):
    z = oracle.normals(5, 0, 0, 0, 2, 0, 400000)
    assert abs(z.mean()) < 4 / math.sqrt(z.size)
    assert abs(z.var() - 1) < 0.01
    assert abs(np.mean(z ** 3)) < 0.02 and abs(np.mean(z ** 4) - 3) < 0.06
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 0.01     # the two Box-Muller outputs
    u = oracle.uniforms01(5, 0, 0, 0, 3, 200000)
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.005


def test_quantiser(oracle):
    S = oracle.quant_shift(1024)
    assert S == 51 and oracle.quant_shift(1025) == 50 and oracle.quant_shift(1) == 61
    q = oracle.det_quant(np.array([0.0, -math.log(2), -800.0, -np.inf, np.nan]), S)
    assert q[0] == 2 ** S and abs(int(q[1]) - 2 ** (S - 1)) <= 4 and q[2] == 0 and q[3] == 0 and q[4] == 0
    x = -np.random.default_rng(1).uniform(0, 60, 1000)
    q = oracle.det_quant(x, S).astype(np.float64)
    np.testing.assert_allclose(q, np.floor(np.exp(x) * 2.0 ** S), rtol=1e-12, atol=1)


def test_normalize_closed_forms(oracle):
    lm, w, ess = oracle.normalize(np.full(100, -7.5))
    assert lm == pytest.approx(-7.5, abs=1e-14) and ess == pytest.approx(100.0) and np.allclose(w, 0.01)
    lw = np.full(50, -np.inf)
    lw[3] = 1.25
    lm, w, ess = oracle.normalize(lw)
    assert ess == 1.0 and w[3] == 1.0 and lm == pytest.approx(1.25 - math.log(50))
    rng = np.random.default_rng(2)
    lw = rng.normal(size=5000) * 4
    a, b = oracle.normalize(lw), oracle.normalize_numpy(lw)   # C (det exp) vs numpy (libm) restatement
    assert a[0] == pytest.approx(b[0], rel=1e-13) and a[2] == pytest.approx(b[2], rel=1e-12)
    np.testing.assert_allclose(a[1], b[1], rtol=1e-13)
    assert np.isnan(oracle.normalize(np.full(4, -np.inf))[0])     # reference behaviour (no guard), particles.jl:5-15


@pytest.mark.parametrize("resampler", [0, 1, 2])
def test_ancestors_c_vs_numpy(oracle, resampler):
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 500, 1024):
        lw = rng.normal(size=n) * 3
        a = oracle.ancestors(lw, resampler, 11, 2, 5, 9)
        b = oracle.ancestors_numpy(lw, resampler, 11, 2, 5, 9)
        np.testing.assert_array_equal(a, b)
        assert a.min() >= 0 and a.max() < n
        if resampler != 0:
            assert np.all(np.diff(a) >= 0)                      # stratified / systematic are sorted
    lw = np.full(64, -np.inf)
    np.testing.assert_array_equal(oracle.ancestors(lw, resampler, 1, 0, 0, 1), np.arange(64))   # Q = 0
    lw = np.full(64, -np.inf)
    lw[40] = 0.0
    assert np.all(oracle.ancestors(lw, resampler, 1, 0, 0, 1) == 40)


def test_two_level_multinomial_restatements_agree(oracle):
    """SPEC §5c (multinomial for n > 8192): the C oracle against the independent numpy restatement, ragged sizes, a dead
    stretch of weights (empty cells), one dominant particle (one cell takes almost every output: many chunks), no mass."""
    rng = np.random.default_rng(7)
    for n in (8193, 12289, 40000):
        lw = rng.normal(size=n) * 2.0
        if n == 12289:
            lw[100:9000] = -800.0
        if n == 40000:
            lw[23456] = 25.0
        a = oracle.ancestors(lw, 0, 5, 2, 1, 7)
        np.testing.assert_array_equal(a, oracle.ancestors_numpy(lw, 0, 5, 2, 1, 7))
        assert a.min() >= 0 and a.max() < n
        cells = a // oracle.MN_CELL
        assert np.all(np.diff(cells) >= 0)                              # cells in order
        if n == 40000:
            assert np.mean(a == 23456) > 0.9
    lw = np.full(9000, -np.inf)
    np.testing.assert_array_equal(oracle.ancestors(lw, 0, 1, 0, 0, 1), np.arange(9000))          # Q = 0
    # n <= 8192 keeps SPEC §5 row 0 (unsorted, one search per particle)
    a = oracle.ancestors(rng.normal(size=8192), 0, 5, 2, 1, 7)
    assert np.any(np.diff(a) < 0)


def test_two_level_multinomial_has_the_multinomial_law(oracle):
    """The offspring counts of the two-level draw (SPEC §5c) are Multinomial(n, w) like those of the one-level draw (§5 row 0,
    StatsBase `sample(1:n, Weights(w), n)`, particles.jl:18): z-scores and χ² of the pooled counts of the heaviest
    particles against n·w, the count VARIANCE of single particles against n·w·(1 − w) (a low-variance scheme would fail
    it), and a two-sample comparison with the one-level draw on the same weights."""
    rng = np.random.default_rng(8)
    n, reps = 10000, 60
    lw = rng.normal(size=n) * 1.2
    w = np.exp(lw - lw.max())
    w /= w.sum()
    top = np.argsort(w)[-200:]
    c2 = np.stack([np.bincount(oracle.ancestors(lw, 0, 123, r, 0, 1), minlength=n)[top] for r in range(reps)])          # two-level
    lw_pad = lw[:8192]                                                                                                   # one-level law on the largest legacy size
    w_pad = np.exp(lw_pad - lw_pad.max())
    w_pad /= w_pad.sum()
    top1 = np.argsort(w_pad)[-200:]
    c1 = np.stack([np.bincount(oracle.ancestors(lw_pad, 0, 123, r, 0, 1), minlength=8192)[top1] for r in range(reps)])
    for c, ww, nn in ((c2, w[top], n), (c1, w_pad[top1], 8192)):
        mean, var = nn * ww, nn * ww * (1 - ww)
        z = (c.sum(axis=0) - reps * mean) / np.sqrt(reps * var)
        assert np.abs(z).max() < 4.5 and abs(z.mean()) < 0.35
        chi2 = float(np.sum(z * z))
        assert 120 < chi2 < 290                                          # χ²(200): mean 200, sd 20
        ratio = c.var(axis=0, ddof=1) / var                              # multinomial: ≈ 1 for every particle
        assert 0.8 < ratio.mean() < 1.2
    # two-sample: standardised counts of the two schemes have the same spread
    s2 = ((c2 - n * w[top]) / np.sqrt(n * w[top] * (1 - w[top]))).ravel()
    s1 = ((c1 - 8192 * w_pad[top1]) / np.sqrt(8192 * w_pad[top1] * (1 - w_pad[top1]))).ravel()
    assert abs(s2.std() - s1.std()) < 0.05 and abs(s2.mean() - s1.mean()) < 0.05


def test_resampler_is_unbiased(oracle):
    rng = np.random.default_rng(4)
    n = 256
    lw = rng.normal(size=n) * 1.5
    w = np.exp(lw - lw.max())
    w /= w.sum()
    for resampler in (0, 1, 2):
        counts = np.zeros(n)
        reps = 400
        for r in range(reps):
            counts += np.bincount(oracle.ancestors(lw, resampler, 99, r, 0, 1), minlength=n)
        z = (counts - reps * n * w) / np.sqrt(reps * n * w * (1 - w) + 1e-12)
        assert np.abs(z).max() < 5.5 and abs(z.mean()) < 0.5
        if resampler:  # low-variance schemes: every count within 1 of n w per draw
            assert np.all(np.abs(counts / reps - n * w) < 1.0)


def test_kalman_matches_closed_form(oracle):
    # one step by hand: kalman_filter.jl:29-53
    A, B, Q, R, x0, s0 = 0.5, 1.0, 0.9, 0.8, 0.0, 1.0
    y = 0.7
    xp, Sp = A * x0, A * A * s0 + Q
    sig = B * B * Sp + R
    x1 = xp + Sp * B / sig * (y - B * xp)
    S1 = Sp - (Sp * B) ** 2 / sig
    ll = -0.5 * (math.log(2 * math.pi) + math.log(sig) + (y - B * xp) ** 2 / sig)
    x, S, l = oracle.kalman_step(LG, x0, s0, y)
    assert x == pytest.approx(x1, rel=1e-14) and S == pytest.approx(S1, rel=1e-14) and l == pytest.approx(ll, rel=1e-14)


def test_reference_style_cpu_arm_targets_the_same_likelihood(oracle):
    """the timed CPU arm of BASELINE.md §3 (alias-table multinomial resampling rebuilt every step, fresh allocations, its own
    xoshiro256++ stream) is a bootstrap filter of the same model: E[Ẑ] = Z against the matched-init Kalman likelihood, and the
    scatter of logZ of a multinomial filter (larger than the systematic filter's, same order)"""
    _, y = oracle.simulate(0, LG, 60, 1998)
    _, _, kf = oracle.kalman_loglik(LG, y, matched_init=True)
    z = np.array([oracle.reference_style_log_likelihood(LG, 4096, y, s) for s in range(48)])
    lme = np.log(np.mean(np.exp(z - z.max()))) + z.max()
    assert abs(lme - kf) < 4 * z.std() / np.sqrt(z.size) + 0.02
    zs = np.array([oracle.log_likelihood(0, LG, 4096, y, oracle.MULTINOMIAL, s)["logZ"] for s in range(48)])
    assert 0.5 < z.std() / zs.std() < 2.0
    assert oracle.reference_style_log_likelihood(LG, 1, y[:3], 5) < 0.0          # a single particle runs
    # the UCSV arm, threaded over θ like Threads.@threads: same likelihood estimate as the parity oracle's filters (two samples)
    UC = [0.2, 0.2, 3.0, 1.0, 1.0]
    _, yu = oracle.simulate(2, UC, 40, 1998)
    P = np.tile(oracle.params8(UC), (48, 1))
    zr = oracle.reference_style_batch(2, P, 1024, yu, 3)
    zo, _, _ = oracle.batch_log_likelihood(2, P, None, 1024, yu, oracle.MULTINOMIAL, 3, 0, 0, want_state=False)
    lme = lambda z: np.log(np.mean(np.exp(z - z.max()))) + z.max()      # noqa: E731
    assert abs(lme(zr) - lme(zo)) < 4 * np.hypot(zr.std(), zo.std()) / np.sqrt(48) + 0.05 and 0.5 < zr.std() / zo.std() < 2.0
    with pytest.raises(ValueError):
        oracle.reference_style_batch(1, P, 64, yu, 3)


def test_particle_filter_targets_kalman_likelihood(oracle):
    """E[Ẑ] = Z: log-mean-exp of PF estimates vs the matched-init Kalman likelihood."""
    _, y = oracle.simulate(0, LG, 60, 1998)
    _, _, kf = oracle.kalman_loglik(LG, y, matched_init=True)
    _, _, kf_ref = oracle.kalman_loglik(LG, y, matched_init=False)
    for resampler in (0, 2):
        vals = np.array([oracle.log_likelihood(0, LG, 2048, y, resampler, 500 + r)["logZ"] for r in range(30)])
        lme = np.log(np.mean(np.exp(vals - vals.max()))) + vals.max()
        assert abs(lme - kf) < 4 * vals.std(ddof=1) / math.sqrt(30) + 0.02
    assert 0 < abs(kf - kf_ref) < 0.3


def test_multivariate_lg_particle_filter_targets_the_matrix_kalman_likelihood(oracle):
    """SPEC §4b: the oracle's multivariate-LG particle filter (state_space_models.jl:156-189) is unbiased for the likelihood of the
    reference's own exact filter, the matrix Kalman recursion (kalman_filter.jl:3-27, matched initial condition: SURVEY D1) — for a
    full-rank model and for hodrick_prescott's singular Q."""
    rng = np.random.default_rng(2)
    A = np.array([[0.8, 0.1], [0.0, 0.5]])
    blk = np.concatenate([A.ravel(), [1.0, 0.5], [0.5, 0.1, 0.1, 0.3], [0.7], [0.5, -0.2], [1.0, 0.2, 0.2, 0.5]])
    _, y = oracle.simulate(3, blk, 40, 7)
    _, _, kal = oracle.kalman_mv_loglik(2, blk, y, matched_init=True)
    zs = np.array([oracle.log_likelihood(3, blk, 3000, y, 2, s, 0, 0)["logZ"] for s in range(12)])
    lme = zs.max() + np.log(np.mean(np.exp(zs - zs.max())))
    assert abs(lme - kal) < 4 * zs.std() / np.sqrt(12) + 0.02, (lme, kal, zs.std())
    hp = np.concatenate([[2.0, -1.0, 1.0, 0.0], [1.0, 0.0], [1 / 50.0, 0.0, 0.0, 0.0], [1.0], [3 * y[0] - 2 * y[1], 2 * y[0] - y[1]], [4.0, 0.0, 0.0, 4.0]])
    _, _, kalhp = oracle.kalman_mv_loglik(2, hp, y, matched_init=True)
    r = oracle.log_likelihood(3, hp, 20000, y, 2, 1, 0, 0)
    assert abs(r["logZ"] - kalhp) < 0.6, (r["logZ"], kalhp)
    # a 4-dimensional state uses all of z[4]: states are finite, the filter runs
    A4 = 0.6 * np.eye(4)
    blk4 = np.concatenate([A4.ravel(), np.ones(4), (0.2 * np.eye(4)).ravel(), [0.5], np.zeros(4), np.eye(4).ravel()])
    _, y4 = oracle.simulate(5, blk4, 20, 3)
    r4 = oracle.log_likelihood(5, blk4, 2000, y4, 0, 1, 0, 0)
    assert r4["x"].shape == (4, 2000) and np.isfinite(r4["logZ"])


def test_oracle_regression_vectors(oracle):
    for v in json.load(open(os.path.join(GOLD, "oracle_vectors.json"))):
        _, y = oracle.simulate(v["kind"], v["params"], v["T"], v["data_seed"])
        assert [float(a).hex() for a in y[:4]] == v["y_hex"]
        r = oracle.log_likelihood(v["kind"], v["params"], v["N"], y, v["resampler"], v["seed"], v["epoch"], v["stream"],
                                  want_anc=True)
        assert float(r["logZ"]).hex() == v["logZ_hex"]
        assert [float(a).hex() for a in r["x"][:, -1]] == v["x_last_hex"]
        assert float(np.sum(r["x"])).hex() == v["x_sum_hex"]
        assert [int(a) for a in r["anc"][1][:16]] == v["anc_t1_head"]
        assert int(np.sum(r["anc"][1:] * (np.arange(v["N"]) + 1)) % (2 ** 61 - 1)) == v["anc_checksum"]


def test_oracle_regression_vectors_of_the_later_rows(oracle):
    """guided filter (docs/SPEC.md §10) and matrix Kalman filter: committed bit patterns of the oracle (tools/gen_golden.py)"""
    for v in json.load(open(os.path.join(GOLD, "widen_vectors.json"))):
        _, y = oracle.simulate(0 if v["what"] == "kalman_mv" else v["kind"], LG if v["what"] == "kalman_mv" else v["params"], v["T"], v["data_seed"])
        if v["what"] == "guided":
            prop = np.array([[float.fromhex(c) for c in row] for row in v["prop_hex"]])
            r = oracle.guided_log_likelihood(v["kind"], v["params"], v["N"], y, v["resampler"], prop, v["seed"], v["epoch"], v["stream"])
            assert float(r["logZ"]).hex() == v["logZ_hex"]
            assert float(np.sum(r["x"])).hex() == v["x_sum_hex"] and float(np.sum(r["logw"])).hex() == v["logw_sum_hex"]
            assert float(r["x"][0, -1]).hex() == v["x_last_hex"]
        else:
            blk = np.array([float.fromhex(c) for c in v["block_hex"]])
            x, S, ll = oracle.kalman_mv_loglik(v["d"], blk, y, v["matched_init"])
            assert float(ll).hex() == v["ll_hex"]
            assert [float(a).hex() for a in x] == v["x_hex"] and [float(a).hex() for a in S.ravel()] == v["S_hex"]


def test_oracle_regression_vectors_of_round_2(oracle):
    """two-level multinomial draw (SPEC §5c), guided UCSV move (§10b), multivariate LG particle filter (§4b): committed bit patterns"""
    for v in json.load(open(os.path.join(GOLD, "round2_vectors.json"))):
        if v["what"] == "two_level_multinomial":
            _, y = oracle.simulate(v["kind"], v["params"], v["T"], v["data_seed"])
            r = oracle.log_likelihood(v["kind"], v["params"], v["N"], y, v["resampler"], v["seed"], v["epoch"], v["stream"], want_anc=True)
            assert float(r["logZ"]).hex() == v["logZ_hex"] and float(np.sum(r["x"])).hex() == v["x_sum_hex"]
            assert [int(a) for a in r["anc"][1][:16]] == v["anc_t1_head"]
            assert int(np.sum(r["anc"][1:] * (np.arange(v["N"]) + 1)) % (2 ** 61 - 1)) == v["anc_checksum"]
        elif v["what"] == "guided_ucsv":
            _, y = oracle.simulate(v["kind"], v["params"], v["T"], v["data_seed"])
            kap = np.array([float.fromhex(c) for c in v["kappa_hex"]])
            prop = np.stack([kap, np.zeros(kap.size), np.ones(kap.size)], 1)
            r = oracle.guided_log_likelihood(v["kind"], v["params"], v["N"], y, v["resampler"], prop, v["seed"], v["epoch"], v["stream"])
            assert float(r["logZ"]).hex() == v["logZ_hex"] and float(np.sum(r["x"])).hex() == v["x_sum_hex"]
            assert float(np.sum(r["logw"])).hex() == v["logw_sum_hex"] and [float(a).hex() for a in r["x"][:, -1]] == v["x_last_hex"]
        else:
            _, y = oracle.simulate(0, LG, v["T"], v["data_seed"])
            blk = np.array([float.fromhex(c) for c in v["block_hex"]])
            r = oracle.log_likelihood(v["kind"], blk, v["N"], y, v["resampler"], v["seed"], v["epoch"], v["stream"])
            assert float(r["logZ"]).hex() == v["logZ_hex"] and float(np.sum(r["x"])).hex() == v["x_sum_hex"]
            assert [float(a).hex() for a in r["x"][:, -1]] == v["x_last_hex"]


def test_batch_oracle_equals_loop(oracle):
    rng = np.random.default_rng(6)
    M, N, T = 5, 100, 12
    P = np.zeros((M, 8))
    P[:, :6] = np.stack([rng.uniform(-0.9, 0.9, M), np.ones(M), rng.uniform(0.3, 2, M), rng.uniform(0.3, 2, M), np.zeros(M),
                         np.ones(M)], 1)
    _, y = oracle.simulate(0, LG, T, 3)
    act = np.array([1, 1, 0, 1, 1], np.uint8)
    z, x, lw = oracle.batch_log_likelihood(0, P, act, N, y, 0, 8, 4, 20)
    for m in range(M):
        if not act[m]:
            assert np.isneginf(z[m])
            continue
        r = oracle.log_likelihood(0, P[m], N, y, 0, 8, 4, 20 + m)
        assert r["logZ"] == z[m]
        np.testing.assert_array_equal(r["x"], x[m])


def test_weighted_summary_restatement(oracle):
    """SPEC §8 on a case small enough to check by hand, and against numpy on equal weights."""
    x = np.array([[3.0, -1.0, 2.0, 10.0]])
    lw = np.log(np.array([0.25, 0.25, 0.5, 1e-300]))
    m, v, q = oracle.weighted_summary(x, lw, [0.0, 0.2, 0.3, 0.6, 1.0])
    np.testing.assert_allclose(m, [1.5], rtol=1e-12)
    np.testing.assert_allclose(v, [0.25 * 2.25 + 0.25 * 6.25 + 0.5 * 0.25], rtol=1e-12)
    np.testing.assert_array_equal(q, [[-1.0, -1.0, 2.0, 2.0, 3.0]])          # cumulative weights: -1: .25, 2: .75, 3: 1
    rng = np.random.default_rng(0)
    z = rng.normal(size=(1, 1001))
    _, _, q = oracle.weighted_summary(z, None, [0.5, 0.25], weighted=False)
    s = np.sort(z[0])
    assert q[0, 0] == s[500] and q[0, 1] == s[250]


def test_state_f32_tier_restatement():
    """docs/SPEC.md §9: the binary32-state tier is the binary64 filter with ONE extra rounding of every drawn
    state; it is a process-global switch of the C oracle that must not leak."""
    from oracle import oracle as o
    P = [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]
    _, y = o.simulate(0, P, 30, 1998)
    r64 = o.log_likelihood(0, P, 2048, y, 2, 7, 0, 0, want_anc=True)
    with o.state_f32():
        r32 = o.log_likelihood(0, P, 2048, y, 2, 7, 0, 0, want_anc=True)
        # one step restated by hand from the binary64 pieces: round the init draw, weight the rounded state
        x0, lw0 = o.bootstrap_init(0, P, 64, y[0], 7)
    x0_64, _ = o.bootstrap_init(0, P, 64, y[0], 7)
    np.testing.assert_array_equal(x0, x0_64.astype(np.float32).astype(np.float64))
    v = (y[0] - 1.0 * x0[0]) * (1.0 / np.sqrt(0.8))
    np.testing.assert_allclose(lw0, -0.5 * v * v - (np.log(np.sqrt(0.8)) + 0.5 * np.log(2 * np.pi)), rtol=1e-13)
    assert np.array_equal(r32["x"], r32["x"].astype(np.float32).astype(np.float64))
    assert not np.array_equal(r32["x"], r64["x"])
    assert abs(r32["logZ"] - r64["logZ"]) <= 1e-4 * abs(r64["logZ"])
    again = o.log_likelihood(0, P, 2048, y, 2, 7, 0, 0)
    assert again["logZ"] == r64["logZ"]                      # the switch did not leak
