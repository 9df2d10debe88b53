#!/usr/bin/env python3
"""Coefficient tables of the binary32 deterministic math (docs/SPEC.md §9b): Chebyshev fits at 200-bit precision (mpmath), rounded
once to binary32; the printed hex floats are normative and pasted into csrc/smcb_detmathf.cuh and oracle/det_math.h.
Run: python tools/gen_coeffs_f32.py"""
import struct

import mpmath as mp

mp.mp.prec = 240


def f32(x):
    return struct.unpack("<f", struct.pack("<f", float(x)))[0]


def fit(f, a, b, deg):
    c = mp.chebyfit(f, [a, b], deg + 1)
    return [f32(x) for x in c[::-1]]


def horner(c, z):
    acc = mp.mpf(c[-1])
    for k in range(len(c) - 2, -1, -1):
        acc = acc * z + mp.mpf(c[k])
    return acc


def max_err(fa, fe, a, b, n=4001):
    worst = mp.mpf(0)
    for i in range(n):
        z = a + (b - a) * mp.mpf(i) / (n - 1)
        ex = fe(z)
        if ex != 0:
            worst = max(worst, abs((fa(z) - ex) / ex))
    return worst


def report(name, c, err):
    print(f"// {name}: max rel err (exact arithmetic) = {mp.nstr(err, 3)}")
    print("  " + ", ".join(float.hex(v) + "f" for v in c))


L = mp.log(2) / 2 * mp.mpf("1.0001")
E_exact = lambda r: mp.mpf(1) / 2 if abs(r) < mp.mpf("1e-30") else (mp.exp(r) - 1 - r) / (r * r)
cE = fit(E_exact, -L, L, 4)
report("EXPF_E (exp(r) = 1 + r + r^2 E(r), |r| <= ln2/2, degree 4)", cE, max_err(lambda r: 1 + r + r * r * horner(cE, r), mp.exp, -L, L))

smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
zmax = smax * smax * mp.mpf("1.0001")
R_exact = lambda z: mp.mpf(2) / 3 if z < mp.mpf("1e-40") else (2 * mp.atanh(mp.sqrt(z)) - 2 * mp.sqrt(z)) / (mp.sqrt(z) * z)
cR = fit(R_exact, 0, zmax, 2)
report("LOGF_R (log(m) = 2s + s z R(z), z = s^2, degree 2)", cR,
       max_err(lambda s: 2 * s + s * s * s * horner(cR, s * s), lambda s: 2 * mp.atanh(s), mp.mpf("1e-6"), smax))

zq = mp.mpf("0.25") * mp.mpf("1.0001")
S_exact = lambda z: mp.pi / 2 if z < mp.mpf("1e-40") else mp.sin(mp.pi / 2 * mp.sqrt(z)) / mp.sqrt(z)
C_exact = lambda z: mp.cos(mp.pi / 2 * mp.sqrt(z))
cS = fit(S_exact, 0, zq, 3)
cC = fit(C_exact, 0, zq, 3)
report("SINQF_S (sin(pi/2 r) = r S(r^2), |r| <= 1/2, degree 3)", cS, max_err(lambda r: r * horner(cS, r * r), lambda r: mp.sin(mp.pi / 2 * r), mp.mpf("1e-6"), mp.mpf("0.5")))
report("COSQF_C (cos(pi/2 r) = C(r^2), |r| <= 1/2, degree 3)", cC, max_err(lambda r: horner(cC, r * r), lambda r: mp.cos(mp.pi / 2 * r), 0, mp.mpf("0.5")))

ln2 = mp.log(2)
hi = struct.unpack("<I", struct.pack("<f", float(ln2)))[0] & ~((1 << 9) - 1)   # low 9 mantissa bits zero: k * LN2_HI exact for |k| < 2^9
ln2_hi = struct.unpack("<f", struct.pack("<I", hi))[0]
print("LN2F_HI =", float.hex(ln2_hi) + "f", " LN2F_LO =", float.hex(f32(ln2 - mp.mpf(ln2_hi))) + "f", " LOG2EF =", float.hex(f32(1 / ln2)) + "f",
      " HALF_LOG_2PIF =", float.hex(f32(mp.log(2 * mp.pi) / 2)) + "f", " SQRT2F =", float.hex(f32(mp.sqrt(2))) + "f")
