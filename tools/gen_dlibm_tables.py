"""Fits the polynomial coefficients of mathmap_b200/csrc/runtime/mm_dlibm.h in 60-digit arithmetic (mpmath) and prints
them as C hex-float literals with the approximation error of each polynomial.
Usage: python tools/gen_dlibm_tables.py"""
import random

import mpmath as mp

mp.mp.dps = 60


def cheb_fit(f, a, b, deg):
    """Polynomial of degree `deg` interpolating f at the Chebyshev nodes of [a, b] (near-minimax), monomial basis."""
    n = deg + 1
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A, y = mp.matrix(n, n), mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [float(c[i]) for i in range(n)]


def horner(cs, z):
    r = mp.mpf(0)
    for c in reversed(cs):
        r = r * z + mp.mpf(c)
    return r


def f_sin(z):  # sin(r) = r + r z S(z)
    if z == 0:
        return -mp.mpf(1) / 6
    s = mp.sqrt(z)
    return (mp.sin(s) / s - 1) / z


def f_cos(z):  # cos(r) = 1 - z/2 + z^2 C(z)
    if z == 0:
        return mp.mpf(1) / 24
    s = mp.sqrt(z)
    return (mp.cos(s) - 1 + z / 2) / (z * z)


def f_asin(z):  # asin(x) = x + x z A(z)
    if z == 0:
        return mp.mpf(1) / 6
    s = mp.sqrt(z)
    return (mp.asin(s) / s - 1) / z


def f_atan(z):  # atan(x) = x + x z T(z)
    if z == 0:
        return -mp.mpf(1) / 3
    s = mp.sqrt(z)
    return (mp.atan(s) / s - 1) / z


def f_exp(r):  # exp(r) = 1 + r + r^2 E(r)
    if r == 0:
        return mp.mpf(1) / 2
    return (mp.exp(r) - 1 - r) / (r * r)


def f_log(s):  # log((1+f)/(1-f)) = 2 f + 2 f s L(s), s = f^2
    if s == 0:
        return mp.mpf(1) / 3
    f = mp.sqrt(s)
    return (mp.log((1 + f) / (1 - f)) / (2 * f) - 1) / s


def show(name, cs):
    print("%s = {%s}" % (name, ", ".join(c.hex() for c in cs)))


def main():
    random.seed(1)
    lim = (mp.pi / 4) ** 2 * mp.mpf("1.0001")
    S, C = cheb_fit(f_sin, 0, lim, 5), cheb_fit(f_cos, 0, lim, 5)
    es = ec = mp.mpf(0)
    for _ in range(20000):
        r = mp.mpf(random.uniform(-0.7854, 0.7854))
        z = r * r
        if r != 0:
            es = max(es, abs(r + r * z * horner(S, z) - mp.sin(r)) / abs(mp.sin(r)))
        ec = max(ec, abs(1 - z / 2 + z * z * horner(C, z) - mp.cos(r)) / abs(mp.cos(r)))
    show("sin S", S)
    show("cos C", C)
    print("relative error in units of 2^-53: sin %.3f cos %.3f" % (es * 2 ** 53, ec * 2 ** 53))
    A = cheb_fit(f_asin, 0, mp.mpf("0.2501"), 11)
    ea = mp.mpf(0)
    for _ in range(20000):
        x = mp.mpf(random.uniform(1e-9, 0.5))
        ea = max(ea, abs(x + x * x * x * horner(A, x * x) - mp.asin(x)) / mp.asin(x))
    show("asin A", A)
    print("relative error in units of 2^-53: asin %.3f" % (ea * 2 ** 53))
    # atan on |x| <= tan(pi/16) after the fdlibm-style argument reduction
    for deg in (8, 9, 10):
        lim_t = mp.tan(mp.pi / 16) ** 2 * mp.mpf("1.001")
        T = cheb_fit(f_atan, 0, lim_t, deg)
        et = mp.mpf(0)
        for _ in range(5000):
            x = mp.mpf(random.uniform(1e-9, float(mp.tan(mp.pi / 16))))
            et = max(et, abs(x + x * x * x * horner(T, x * x) - mp.atan(x)) / mp.atan(x))
        print("atan degree %d: relative error %.3f" % (deg, et * 2 ** 53))
        if deg == 9:
            show("atan T", T)
    # exp on |r| <= ln2/2
    for deg in (9, 10, 11):
        h = mp.log(2) / 2 * mp.mpf("1.0001")
        E = cheb_fit(f_exp, -h, h, deg)
        ee = mp.mpf(0)
        for _ in range(5000):
            r = mp.mpf(random.uniform(-0.34658, 0.34658))
            ee = max(ee, abs(1 + r + r * r * horner(E, r) - mp.exp(r)) / mp.exp(r))
        print("exp degree %d: relative error %.3f" % (deg, ee * 2 ** 53))
        if deg == 10:
            show("exp E", E)
    # log: m in [sqrt(1/2), sqrt(2)), f = (m-1)/(m+1), |f| <= 0.1716
    for deg in (6, 7, 8):
        fl = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
        L = cheb_fit(f_log, 0, fl * fl * mp.mpf("1.001"), deg)
        el = mp.mpf(0)
        for _ in range(5000):
            f = mp.mpf(random.uniform(1e-6, float(fl)))
            v = 2 * f + 2 * f * f * f * horner(L, f * f)
            t = mp.log((1 + f) / (1 - f))
            el = max(el, abs(v - t) / t)
        print("log degree %d: relative error %.3f" % (deg, el * 2 ** 53))
        if deg == 7:
            show("log L", L)


if __name__ == "__main__":
    main()
