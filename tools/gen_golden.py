"""Writes tests/golden/*.json: Philox known answers (Random123 / cuRAND constants, verified in
SURVEY.md §7.1) and regression vectors of the CPU oracle at small sizes.  The reference is Julia,
cannot run in this image and ships no golden vectors (parity unpinned), so these
fixtures pin the ORACLE (and through it the CUDA path) against silent drift, not the Julia code.

    python tools/gen_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
os.makedirs(G, exist_ok=True)

kat = [
    dict(ctr=[0, 0, 0, 0], key=[0, 0], out=["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]),
    dict(ctr=[0xFFFFFFFF] * 4, key=[0xFFFFFFFF] * 2, out=["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]),
    dict(ctr=[0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], key=[0xA4093822, 0x299F31D0],
         out=["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]),
]
json.dump(kat, open(os.path.join(G, "philox_kat.json"), "w"), indent=1)

MODELS = {0: [0.5, 1.0, 0.9, 0.8, 0.0, 1.0], 1: [-1.0, 0.9, 0.3], 2: [0.2, 0.2, 3.0, 1.0, 1.0]}
vec = []
for kind, P in MODELS.items():
    _, y = o.simulate(kind, P, 40, 1998)
    for rs in (0, 1, 2):
        r = o.log_likelihood(kind, P, 257, y, rs, seed=7, epoch=3, stream=2, want_anc=True)
        vec.append(dict(kind=kind, params=P, N=257, T=40, data_seed=1998, resampler=rs, seed=7, epoch=3, stream=2,
                        y_hex=[float(v).hex() for v in y[:4]],
                        logZ_hex=float(r["logZ"]).hex(),
                        x_last_hex=[float(v).hex() for v in r["x"][:, -1]],
                        x_sum_hex=float(np.sum(r["x"])).hex(),
                        anc_t1_head=[int(v) for v in r["anc"][1][:16]],
                        anc_checksum=int(np.sum(r["anc"][1:] * (np.arange(257) + 1)) % (2 ** 61 - 1))))
json.dump(vec, open(os.path.join(G, "oracle_vectors.json"), "w"), indent=1)

# the rows built after the hot path: guided filter (docs/SPEC.md §10) and the matrix Kalman filter
wid = []
for kind, P in ((0, MODELS[0]), (1, MODELS[1])):
    _, y = o.simulate(kind, P, 30, 1998)
    if kind == 0:
        prop = np.array([o.optimal_proposal_lg(P, yt) for yt in y])
    else:
        prop = np.array([[P[0] * (1 - 0.8 * P[1]) + 0.05 * np.log(yt * yt + 1e-3), 0.8 * P[1], 1.3 * P[2]] for yt in y])
    for rs in (0, 2):
        r = o.guided_log_likelihood(kind, P, 257, y, rs, prop, 7, 3, 2)
        wid.append(dict(what="guided", kind=kind, params=P, N=257, T=30, data_seed=1998, resampler=rs, seed=7, epoch=3, stream=2,
                        prop_hex=[[float(v).hex() for v in row] for row in prop],
                        logZ_hex=float(r["logZ"]).hex(), x_sum_hex=float(np.sum(r["x"])).hex(),
                        logw_sum_hex=float(np.sum(r["logw"])).hex(), x_last_hex=float(r["x"][0, -1]).hex()))
_, y = o.simulate(0, MODELS[0], 60, 1998)
hp = o.mv_block([[2, -1], [1, 0]], [1, 0], [[1 / 1600.0, 0], [0, 0]], [1.0], [3 * y[0] - 2 * y[1], 2 * y[0] - y[1]], 1000 * np.eye(2))
for matched in (0, 1):
    x, S, ll = o.kalman_mv_loglik(2, hp, y, matched)
    wid.append(dict(what="kalman_mv", d=2, block_hex=[float(v).hex() for v in hp], T=60, data_seed=1998, matched_init=matched,
                    ll_hex=float(ll).hex(), x_hex=[float(v).hex() for v in x], S_hex=[float(v).hex() for v in S.ravel()]))
json.dump(wid, open(os.path.join(G, "widen_vectors.json"), "w"), indent=1)

# round 2: the two-level multinomial draw of a large cloud (SPEC §5c, sorted in-cell thresholds), the guided UCSV move (§10b),
# the particle-filter functor of the multivariate linear models (§4b)
r2 = []
_, y = o.simulate(0, MODELS[0], 6, 1998)
r = o.log_likelihood(0, MODELS[0], 20011, y, 0, seed=7, epoch=3, stream=2, want_anc=True)
r2.append(dict(what="two_level_multinomial", kind=0, params=MODELS[0], N=20011, T=6, data_seed=1998, resampler=0, seed=7, epoch=3, stream=2,
               logZ_hex=float(r["logZ"]).hex(), x_sum_hex=float(np.sum(r["x"])).hex(), anc_t1_head=[int(v) for v in r["anc"][1][:16]],
               anc_checksum=int(np.sum(r["anc"][1:] * (np.arange(20011) + 1)) % (2 ** 61 - 1))))
_, y = o.simulate(2, MODELS[2], 20, 1998)
kap = np.linspace(0.0, 1.0, 20)
prop = np.stack([kap, np.zeros(20), np.ones(20)], 1)
for rs in (0, 2):
    r = o.guided_log_likelihood(2, MODELS[2], 257, y, rs, prop, 7, 3, 2)
    r2.append(dict(what="guided_ucsv", kind=2, params=MODELS[2], N=257, T=20, data_seed=1998, resampler=rs, seed=7, epoch=3, stream=2,
                   kappa_hex=[float(v).hex() for v in kap], logZ_hex=float(r["logZ"]).hex(), x_sum_hex=float(np.sum(r["x"])).hex(),
                   logw_sum_hex=float(np.sum(r["logw"])).hex(), x_last_hex=[float(v).hex() for v in r["x"][:, -1]]))
_, y = o.simulate(0, MODELS[0], 30, 1998)
blk = o.mv_block([[0.7, 0.2], [-0.1, 0.5]], [1.0, 0.5], [[0.5, 0.1], [0.1, 0.3]], [0.8], [0.0, 0.0], np.eye(2))
r = o.log_likelihood(3, blk, 9001, y, 2, seed=7, epoch=3, stream=2)
r2.append(dict(what="mvlg_filter", kind=3, block_hex=[float(v).hex() for v in blk], N=9001, T=30, data_seed=1998, resampler=2, seed=7, epoch=3, stream=2,
               logZ_hex=float(r["logZ"]).hex(), x_sum_hex=float(np.sum(r["x"])).hex(), x_last_hex=[float(v).hex() for v in r["x"][:, -1]]))
json.dump(r2, open(os.path.join(G, "round2_vectors.json"), "w"), indent=1)

# det-math spot values (bit patterns) — any change of a coefficient or operation order shows up here
xs = [-700.0, -37.25, -1.0, -1e-3, 0.0, 0.5, 1.0, 10.125, 700.0]
us = [2.0 ** -53, 1e-9, 0.1, 0.5, 0.75, 1 - 2.0 ** -53]
dm = dict(exp={float(v).hex(): float(r).hex() for v, r in zip(xs, o.det_exp(xs))},
          log={float(v).hex(): float(r).hex() for v, r in zip(us, o.det_log(us))},
          sin2pi={float(v).hex(): float(r).hex() for v, r in zip(us, o.det_sincos2pi(us)[0])},
          cos2pi={float(v).hex(): float(r).hex() for v, r in zip(us, o.det_sincos2pi(us)[1])},
          normals_head=[float(v).hex() for v in o.normals(1998, 1, 2, 3, 2, 0, 8)],
          uniforms_head=[int(v) for v in o.uniforms64(1998, 1, 2, 3, 3, 8)])
json.dump(dm, open(os.path.join(G, "detmath_vectors.json"), "w"), indent=1)
print("wrote", os.listdir(G))
