#!/usr/bin/env python3
"""Generate the polynomial coefficient tables of the deterministic math spec (docs/SPEC.md §3).

Every transcendental on the particle-filter hot path (exp for the weights, log / sinpi / cospi for
the Box-Muller normals) is a fixed sequence of IEEE-754 binary64 fma/add/mul operations so that
the CUDA kernels and the CPU oracle produce bit-identical results.  The coefficients are Chebyshev
fits computed here at 200-bit precision with mpmath and rounded once to binary64; the rounded
values (as hex floats) are the normative ones and are pasted into
  sequential_monte_carlo_b200/csrc/smcb_detmath.cuh   (product)
  oracle/det_math.h                                    (oracle, restated independently)

Run:  python tools/gen_coeffs.py
"""
import mpmath as mp

mp.mp.prec = 240


def fit(f, a, b, deg):
    # chebyfit returns highest-degree coefficient first; we print lowest first.
    c = mp.chebyfit(f, [a, b], deg + 1)
    return [float(x) for x in c[::-1]]


def horner(c, z):
    acc = mp.mpf(c[-1])
    for k in range(len(c) - 2, -1, -1):
        acc = acc * z + mp.mpf(c[k])
    return acc


def report(name, coeffs, err):
    print(f"// {name}: max rel err of the rounded polynomial (exact arithmetic) = {mp.nstr(err, 3)}")
    for k, c in enumerate(coeffs):
        print(f"  /* c{k:<2d} */ {float.hex(c)},   // {c!r}")
    print()


def max_err(fapprox, fexact, a, b, n=4001):
    worst = mp.mpf(0)
    for i in range(n):
        z = a + (b - a) * mp.mpf(i) / (n - 1)
        ex = fexact(z)
        if ex == 0:
            continue
        e = abs((fapprox(z) - ex) / ex)
        worst = max(worst, e)
    return worst


# --- exp(r) = 1 + r + r^2 * E(r),  |r| <= ln2/2 ------------------------------------------
L = mp.log(2) / 2 * mp.mpf("1.0001")


def E_exact(r):
    if abs(r) < mp.mpf("1e-30"):
        return mp.mpf(1) / 2
    return (mp.exp(r) - 1 - r) / (r * r)


cE = fit(E_exact, -L, L, 9)
errE = max_err(lambda r: 1 + r + r * r * horner(cE, r), mp.exp, -L, L)
report("EXP_E (exp(r) = 1 + r + r^2*E(r), |r|<=ln2/2, degree 9)", cE, errE)

# --- log(m) = 2s + s*z*R(z),  s=(m-1)/(m+1), z=s^2, m in [sqrt(1/2), sqrt(2)] --------------
smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
zmax = smax * smax * mp.mpf("1.0001")


def R_exact(z):
    if z < mp.mpf("1e-40"):
        return mp.mpf(2) / 3
    s = mp.sqrt(z)
    return (2 * mp.atanh(s) - 2 * s) / (s * z)


cR = fit(R_exact, 0, zmax, 6)


def log_approx_s(s):
    z = s * s
    return 2 * s + s * z * horner(cR, z)


errR = max_err(log_approx_s, lambda s: 2 * mp.atanh(s), mp.mpf("1e-6"), smax)
report("LOG_R (log(m) = 2s + s*z*R(z), z=s^2<=0.0295, degree 6)", cR, errR)

# --- sin(pi/2 * r) = r * S(r^2), cos(pi/2 * r) = C(r^2), |r| <= 1/2 ------------------------
zq = mp.mpf("0.25") * mp.mpf("1.0001")


def S_exact(z):
    if z < mp.mpf("1e-40"):
        return mp.pi / 2
    r = mp.sqrt(z)
    return mp.sin(mp.pi / 2 * r) / r


def C_exact(z):
    return mp.cos(mp.pi / 2 * mp.sqrt(z))


cS = fit(S_exact, 0, zq, 6)
cC = fit(C_exact, 0, zq, 7)
errS = max_err(lambda r: r * horner(cS, r * r), lambda r: mp.sin(mp.pi / 2 * r), mp.mpf("1e-6"), mp.mpf("0.5"))
errC = max_err(lambda r: horner(cC, r * r), lambda r: mp.cos(mp.pi / 2 * r), 0, mp.mpf("0.5"))
report("SINQ_S (sin(pi/2 r) = r*S(r^2), |r|<=1/2, degree 6 in r^2)", cS, errS)
report("COSQ_C (cos(pi/2 r) = C(r^2), |r|<=1/2, degree 7 in r^2)", cC, errC)

# --- scalar constants ----------------------------------------------------------------------
ln2 = mp.log(2)
# LN2_HI has its low 21 mantissa bits zero so that k*LN2_HI is exact for |k| < 2^20.
import struct

hi_bits = struct.unpack("<Q", struct.pack("<d", float(ln2)))[0] & ~((1 << 21) - 1)
ln2_hi = struct.unpack("<d", struct.pack("<Q", hi_bits))[0]
ln2_lo = float(ln2 - mp.mpf(ln2_hi))
print("LN2_HI  =", float.hex(ln2_hi), repr(ln2_hi))
print("LN2_LO  =", float.hex(ln2_lo), repr(ln2_lo))
print("LOG2E   =", float.hex(float(1 / ln2)), repr(float(1 / ln2)))
print("HALF_LOG_2PI =", float.hex(float(mp.log(2 * mp.pi) / 2)), repr(float(mp.log(2 * mp.pi) / 2)))
print("SQRT_HALF    =", float.hex(float(mp.sqrt(mp.mpf(1) / 2))), repr(float(mp.sqrt(mp.mpf(1) / 2))))
