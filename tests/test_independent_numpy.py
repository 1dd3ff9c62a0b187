"""a cross-check of the oracle that does not go through the product's front end.

The oracle (oracle/oracle.py) evaluates IR that mathmap_b200's own parser, overload resolution and passes produced, so a
front-end bug is invisible to every CUDA-vs-oracle test.  Here four filters (the BASELINE configs' Ident, Twirl, Sea and
Mandelbrot) are evaluated FROM THEIR .mm TEXT BY HAND in numpy float32, following the reference's semantics as SURVEY.md
Appendix A records them (opmacros.h:156-157 coordinates, mathmap_common.c:59-72 / compiler.c:2339-2420 coordinate systems,
compiler.c:1710-1773 resize factors, builtins.c:133-245 samplers, builtins.lisp:533-550,732-736,1328-1346 quaternion
product, norm and polar conversion, new_template.c.in:272-293 quantisation) -- at non-square sizes, t != 0 and with both
samplers, i.e. away from the reference's own 256x256 goldens.  The oracle must reproduce them bit for bit."""
import math

import numpy as np
import pytest

import mathmap_b200 as mb
from conftest import filter_source, synthetic_rgba
from oracle.oracle import OracleFilter

F = np.float32


def f32(x):
    return np.asarray(x, dtype=np.float64).astype(np.float32)


def libm(fn, *args):
    """float arguments promoted to double, the double result narrowed"""
    return fn(*[np.asarray(a, dtype=np.float64) for a in args]).astype(np.float32)


def virtual_coords(W, H):
    cols = np.arange(W, dtype=np.float64)
    rows = np.arange(H, dtype=np.float64)
    xu = ((cols - (W - 1) / 2.0) / ((W - 1) / 2.0)).astype(np.float32)
    yu = ((-rows + (H - 1) / 2.0) / ((H - 1) / 2.0)).astype(np.float32)
    m = max(W, H)
    X = F(F(W) / F(m))
    Y = F(F(H) / F(m))
    x = (xu * X)[None, :].repeat(H, 0)
    y = (yu * Y)[:, None].repeat(W, 1)
    return x.astype(np.float32), y.astype(np.float32), X, Y


def sample(img, x, y, bilinear):
    """origVal on a default-flag image of a default-flag filter: resize factors max/w, max/h, then builtins.c:133-245"""
    h, w = img.shape[:2]
    m = max(w, h)
    x = (x * F(F(m) / F(w))).astype(np.float32)
    y = (y * F(F(m) / F(h))).astype(np.float32)
    sx, sy = F((w - 1) / 2.0), F((h - 1) / 2.0)
    px = ((x + F(1.0)) * sx).astype(np.float32)
    py = (-((y - F(1.0)) * sy)).astype(np.float32)
    texels = img.astype(np.float32)

    def texel(ix, iy):
        inside = (ix >= 0) & (ix < w) & (iy >= 0) & (iy < h)
        v = texels[np.clip(iy, 0, h - 1), np.clip(ix, 0, w - 1)]
        return np.where(inside[..., None], v, F(0.0))  # edge colour (0, 0, 0, 0)

    if not bilinear:
        ix = np.floor((px.astype(np.float64) + 0.5).astype(np.float32)).astype(np.int64)
        iy = np.floor((py.astype(np.float64) + 0.5).astype(np.float32)).astype(np.int64)
        q = texel(ix, iy)
    else:
        x1 = np.floor(px).astype(np.int64)
        y1 = np.floor(py).astype(np.int64)
        x2f = (px - x1.astype(np.float32)).astype(np.float32)
        y2f = (py - y1.astype(np.float32)).astype(np.float32)
        x1f = (F(1.0) - x2f).astype(np.float32)
        y1f = (F(1.0) - y2f).astype(np.float32)
        p1, p2, p3, p4 = (x1f * y1f)[..., None], (x1f * y2f)[..., None], (x2f * y1f)[..., None], (x2f * y2f)[..., None]
        c1, c2, c3, c4 = texel(x1, y1), texel(x1, y1 + 1), texel(x1 + 1, y1), texel(x1 + 1, y1 + 1)
        s = (((c1 * p1).astype(np.float32) + (c2 * p2).astype(np.float32)).astype(np.float32) + (c3 * p3).astype(np.float32)).astype(np.float32)
        s = (s + (c4 * p4).astype(np.float32)).astype(np.float32)
        q = np.rint(s)  # round half to even
    return (q.astype(np.float64) / 255.0).astype(np.float32)


def quantise(t):
    v = np.where(1.0 < t, F(1.0), t)          # MIN(1, t): NaN stays
    v = np.where(0.0 < v, v, F(0.0))          # MAX(0, v): NaN -> 0
    return np.floor(v.astype(np.float64) * 255.0).astype(np.uint8)


def ident(img, W, H, t, bilinear):
    x, y, _, _ = virtual_coords(W, H)
    return quantise(sample(img, x, y, bilinear))


def twirl(img, W, H, t, bilinear):
    """in(ra+ra:[0,(r/R-1)*(t-0.5)*4*pi])"""
    x, y, _, _ = virtual_coords(W, H)
    r = libm(np.hypot, x, y)
    with np.errstate(invalid="ignore", divide="ignore"):
        a = libm(np.arccos, (x / r).astype(np.float32))
    a = np.where(y < 0, (F(2 * math.pi) - a).astype(np.float32), a)
    a = np.where(r == 0, F(0.0), a).astype(np.float32)
    R = F(math.sqrt(2.0))
    e = ((r / R).astype(np.float32) - F(1.0)).astype(np.float32)
    e = (e * F(F(t) - F(0.5))).astype(np.float32)
    e = (e * F(4.0)).astype(np.float32)
    e = (e * F(math.pi)).astype(np.float32)
    r2 = (r + F(0.0)).astype(np.float32)
    a2 = (a + e).astype(np.float32)
    sx = (libm(np.cos, a2) * r2).astype(np.float32)
    sy = (libm(np.sin, a2) * r2).astype(np.float32)
    return quantise(sample(img, sx, sy, bilinear))


def sea(img, W, H, t, bilinear, amp1=0.03, amp2=0.01, wv=5.0):
    """s=sin(t*2*pi+wv*(Y-y+0.1)^-1); in(xy+xy:[amp1*s,amp2*s])"""
    x, y, _, Y = virtual_coords(W, H)
    base = ((Y - y).astype(np.float32) + F(0.1)).astype(np.float32)
    inv = np.where(base == 0, F(0.0), libm(np.power, base, np.full_like(base, -1.0)))  # pow guard: pow(0, b <= 0) = 0
    arg = ((F(F(t) * F(2.0)) * F(math.pi)).astype(np.float32) + (F(wv) * inv).astype(np.float32)).astype(np.float32)
    s = libm(np.sin, arg)
    sx = (x + (F(amp1) * s).astype(np.float32)).astype(np.float32)
    sy = (y + (F(amp2) * s).astype(np.float32)).astype(np.float32)
    return quantise(sample(img, sx, sy, bilinear))


def mandelbrot(W, H, n_iter):
    """p = quat:[x, y, 0, 0]; c = quat:[0, 0, 0, 0]; while abs(c) < 2 && iter < n-1: c = c*c + p; gray iter / n"""
    x, y, _, _ = virtual_coords(W, H)
    p = [x, y, np.zeros_like(x), np.zeros_like(x)]
    c = [np.zeros_like(x) for _ in range(4)]
    it = np.zeros(x.shape, np.int32)
    m = lambda u, v: (u * v).astype(np.float32)
    a3 = lambda u, v: (u + v).astype(np.float32)
    neg = lambda u: (-u).astype(np.float32)
    for _ in range(n_iter):
        n2 = a3(a3(a3(m(c[0], c[0]), m(c[1], c[1])), m(c[2], c[2])), m(c[3], c[3]))
        active = (libm(np.sqrt, n2) < 2) & (it < n_iter - 1)
        if not active.any():
            break
        a, b = c, c
        q0 = a3(a3(a3(m(a[0], b[0]), neg(m(a[1], b[1]))), neg(m(a[2], b[2]))), neg(m(a[3], b[3])))
        q1 = a3(a3(a3(m(a[0], b[1]), m(a[1], b[0])), m(a[2], b[3])), neg(m(a[3], b[2])))
        q2 = a3(a3(a3(m(a[0], b[2]), m(a[2], b[0])), neg(m(a[1], b[3]))), m(a[3], b[1]))
        q3 = a3(a3(a3(m(a[0], b[3]), m(a[3], b[0])), m(a[1], b[2])), neg(m(a[2], b[1])))
        new = [a3(q, pp) for q, pp in zip((q0, q1, q2, q3), p)]
        c = [np.where(active, nv, ov) for nv, ov in zip(new, c)]
        it = np.where(active, it + 1, it)
    g = (it.astype(np.float32) / F(n_iter)).astype(np.float32)
    return quantise(np.stack([g, g, g, np.ones_like(g)], axis=-1))


CASES = [("ident", "examples/Utilities/Ident.mm", ident), ("twirl", "examples/Distorts/Twirl.mm", twirl), ("sea", "examples/Distorts/Sea.mm", sea)]


@pytest.mark.parametrize("name,path,fn", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("bilinear", [False, True], ids=["nearest", "bilinear"])
@pytest.mark.parametrize("size,insize,t", [((211, 96), (211, 96), 0.3), ((90, 140), (123, 77), 0.85)])
def test_oracle_matches_hand_evaluation(name, path, fn, bilinear, size, insize, t):
    W, H = size
    img = synthetic_rgba(insize[0], insize[1], seed=7)
    m = mb.Module(source=filter_source(path))
    got = OracleFilter(m.ir).render(W, H, {"in": img}, t=t, antialiasing=bilinear)
    want = fn(img, W, H, t, bilinear)
    diff = np.abs(got.astype(int) - want.astype(int))
    assert np.array_equal(got, want), "%d pixels differ, max %d" % (int((diff.max(axis=2) > 0).sum()), int(diff.max()))


@pytest.mark.parametrize("size,n", [((160, 97), 32), ((75, 120), 100)])
def test_oracle_matches_hand_evaluation_mandelbrot(size, n):
    W, H = size
    m = mb.Module(source=filter_source("examples/Render/Mandelbrot.mm"))
    got = OracleFilter(m.ir).render(W, H, {"num_iterations": n}, t=0.0)
    want = mandelbrot(W, H, n)
    assert np.array_equal(got, want), "%d pixels differ" % int((np.abs(got.astype(int) - want.astype(int)).max(axis=2) > 0).sum())


# ---- edge behaviours and supersampling by hand (builtins.c:41-131, mathmap_common.c:874-936) --------------------------------
def apply_edge(ix, iy, w, h, mode_x, mode_y):
    """apply_edge_behaviour on int64 arrays of in-range-of-int32 values (finite coordinates only here)."""
    ix, iy = ix.copy(), iy.copy()

    def crem(a, b):  # C remainder: sign of the dividend
        return np.sign(a) * (np.abs(a) % b)

    if mode_x == 1:
        ix = np.where(ix < 0, crem(ix, w) + w, np.where(ix >= w, ix % w, ix))
    elif mode_x == 2:
        ix = np.where(ix < 0, crem(-ix, w), np.where(ix >= w, (w - 1) - (ix % w), ix))
    elif mode_x == 3:
        flip = (ix < 0) | (ix >= w)
        ix = np.where(ix < 0, crem(-ix, w), np.where(ix >= w, (w - 1) - (ix % w), ix))
        iy = np.where(flip, (h - 1) - iy, iy)
    if mode_y == 1:
        iy = np.where(iy < 0, crem(iy, h) + h, np.where(iy >= h, iy % h, iy))
    elif mode_y == 2:
        iy = np.where(iy < 0, crem(-iy, h), np.where(iy >= h, (h - 1) - (iy % h), iy))
    elif mode_y == 3:
        flip = (iy < 0) | (iy >= h)
        iy = np.where(iy < 0, crem(-iy, h), np.where(iy >= h, (h - 1) - (iy % h), iy))
        ix = np.where(flip, (w - 1) - ix, ix)
    return ix, iy


def sample_edges(img, x, y, bilinear, modes, colors, supersampling=False):
    """origVal with edge behaviours and edge colours (packed R<<24|G<<16|B<<8|A; x is tested first, builtins.c:121-131);
    under supersampling the nearest sampler does not add 0.5 (builtins.c:155-159)."""
    h, w = img.shape[:2]
    m = max(w, h)
    x = (x * F(F(m) / F(w))).astype(np.float32)
    y = (y * F(F(m) / F(h))).astype(np.float32)
    px = ((x + F(1.0)) * F((w - 1) / 2.0)).astype(np.float32)
    py = (-((y - F(1.0)) * F((h - 1) / 2.0))).astype(np.float32)
    texels = img.astype(np.float32)

    def unpack(c):
        return np.array([(c >> 24) & 255, (c >> 16) & 255, (c >> 8) & 255, c & 255], dtype=np.float32)

    def texel(ix, iy):
        ix, iy = apply_edge(ix, iy, w, h, modes[0], modes[1])
        out_x = (ix < 0) | (ix >= w)
        out_y = (iy < 0) | (iy >= h)
        v = texels[np.clip(iy, 0, h - 1), np.clip(ix, 0, w - 1)]
        v = np.where(out_y[..., None], unpack(colors[1]), v)
        return np.where(out_x[..., None], unpack(colors[0]), v)

    if not bilinear:
        if not supersampling:
            px = (px.astype(np.float64) + 0.5).astype(np.float32)
            py = (py.astype(np.float64) + 0.5).astype(np.float32)
        q = texel(np.floor(px).astype(np.int64), np.floor(py).astype(np.int64))
    else:
        x1 = np.floor(px).astype(np.int64)
        y1 = np.floor(py).astype(np.int64)
        x2f = (px - x1.astype(np.float32)).astype(np.float32)
        y2f = (py - y1.astype(np.float32)).astype(np.float32)
        x1f = (F(1.0) - x2f).astype(np.float32)
        y1f = (F(1.0) - y2f).astype(np.float32)
        p1, p2, p3, p4 = (x1f * y1f)[..., None], (x1f * y2f)[..., None], (x2f * y1f)[..., None], (x2f * y2f)[..., None]
        c1, c2, c3, c4 = texel(x1, y1), texel(x1, y1 + 1), texel(x1 + 1, y1), texel(x1 + 1, y1 + 1)
        s = (((c1 * p1).astype(np.float32) + (c2 * p2).astype(np.float32)).astype(np.float32) + (c3 * p3).astype(np.float32)).astype(np.float32)
        s = (s + (c4 * p4).astype(np.float32)).astype(np.float32)
        q = np.rint(s)
    return (q.astype(np.float64) / 255.0).astype(np.float32)


def coords_of(cols, rows, W, H, off):
    """CALC_VIRTUAL_X/Y (opmacros.h:156-157) for column / row index arrays and a sampling offset, times the filter's X, Y"""
    xu = ((cols.astype(np.float64) - (W - 1) / 2.0 + off) / ((W - 1) / 2.0)).astype(np.float32)
    yu = ((-rows.astype(np.float64) + (H - 1) / 2.0 - off) / ((H - 1) / 2.0)).astype(np.float32)
    m = max(W, H)
    x = (xu * F(F(W) / F(m)))[None, :].repeat(len(rows), 0)
    y = (yu * F(F(H) / F(m)))[:, None].repeat(len(cols), 1)
    return x.astype(np.float32), y.astype(np.float32)


@pytest.mark.parametrize("bilinear", [False, True], ids=["nearest", "bilinear"])
@pytest.mark.parametrize("modes", [(1, 1), (2, 2), (3, 3), (3, 1), (0, 2), (2, 0)])
def test_oracle_edge_behaviours_match_hand_evaluation(modes, bilinear):
    """Geometry/Zoom `in(xy * factor)` with factor 2.7: most samples fall outside and come back through wrap / reflect / rotate."""
    W, H = 83, 61
    img = synthetic_rgba(57, 44, seed=11)
    colors = (0x11223344, 0xA5667788)
    m = mb.Module(source=filter_source("examples/Geometry/Zoom.mm"))
    got = OracleFilter(m.ir).render(W, H, {"in": img, "factor": 2.7}, t=0.0, antialiasing=bilinear, edge_behaviour=modes, edge_colors=colors)
    x, y = coords_of(np.arange(W), np.arange(H), W, H, 0.0)
    want = quantise(sample_edges(img, (x * F(2.7)).astype(np.float32), (y * F(2.7)).astype(np.float32), bilinear, modes, colors))
    assert np.array_equal(got, want), "%d pixels differ" % int((np.abs(got.astype(int) - want.astype(int)).max(axis=2) > 0).sum())


@pytest.mark.parametrize("bilinear", [False, True], ids=["nearest", "bilinear"])
def test_oracle_supersampling_matches_hand_evaluation(bilinear):
    """-o (mathmap_common.c:874-936): per row the short line and two lines of a slice one column wider sampled at (-0.5, -0.5),
    combined (l1[c] + l1[c+1] + 2 l2[c] + l3[c] + l3[c+1]) / 6 in integers; the row below the last one is clamped away and l3
    keeps the previous line there.  Twirl at t = 0.3, one band (one thread)."""
    W, H, t = 47, 38, 0.3
    img = synthetic_rgba(W, H, seed=3)
    m = mb.Module(source=filter_source("examples/Distorts/Twirl.mm"))
    got = OracleFilter(m.ir).render(W, H, {"in": img}, t=t, antialiasing=bilinear, supersampling=True)

    def twirl_at(x, y):
        r = libm(np.hypot, x, y)
        with np.errstate(invalid="ignore", divide="ignore"):
            a = libm(np.arccos, (x / r).astype(np.float32))
        a = np.where(y < 0, (F(2 * math.pi) - a).astype(np.float32), a)
        a = np.where(r == 0, F(0.0), a).astype(np.float32)
        e = ((r / F(math.sqrt(2.0))).astype(np.float32) - F(1.0)).astype(np.float32)
        e = (((e * F(F(t) - F(0.5))).astype(np.float32) * F(4.0)).astype(np.float32) * F(math.pi)).astype(np.float32)
        a2 = (a + e).astype(np.float32)
        r2 = (r + F(0.0)).astype(np.float32)
        sx = (libm(np.cos, a2) * r2).astype(np.float32)
        sy = (libm(np.sin, a2) * r2).astype(np.float32)
        return quantise(sample_edges(img, sx, sy, bilinear, (0, 0), (0, 0), supersampling=True)).astype(np.int64)

    short = twirl_at(*coords_of(np.arange(W), np.arange(H), W, H, 0.0))
    longs = twirl_at(*coords_of(np.arange(W + 1), np.arange(H), W, H, -0.5))  # rows 0 .. H-1 of the wider slice; row H is clamped away
    want = np.empty((H, W, 4), np.uint8)
    for row in range(H):
        l1 = longs[row]
        l3 = longs[row + 1] if row + 1 < H else longs[row]
        want[row] = ((l1[:-1] + l1[1:] + 2 * short[row] + l3[:-1] + l3[1:]) // 6).astype(np.uint8)
    assert np.array_equal(got, want), "%d pixels differ" % int((np.abs(got.astype(int) - want.astype(int)).max(axis=2) > 0).sum())


# ---- the Gaussian blur's recursive filter by hand (native-filters/gauss.c:37-262) -------------------------------------------
def iir_constants(std_dev):
    """find_iir_constants (gauss.c:37-115) in Python doubles (math.* is the C library); std_dev is a float promoted to double."""
    sd = float(np.float32(std_dev))
    div = math.sqrt(2 * math.pi) * sd
    x0, x1, x2, x3 = -1.783 / sd, -1.723 / sd, 0.6318 / sd, 1.997 / sd
    x4, x5, x6, x7 = 1.6803 / div, 3.735 / div, -0.6803 / div, -0.2598 / div
    e, s, c = math.exp, math.sin, math.cos
    n_p = [x4 + x6,
           (e(x1) * (x7 * s(x3) - (x6 + 2 * x4) * c(x3)) + e(x0) * (x5 * s(x2) - (2 * x6 + x4) * c(x2))),
           (2 * e(x0 + x1) * ((x4 + x6) * c(x3) * c(x2) - x5 * c(x3) * s(x2) - x7 * c(x2) * s(x3)) + x6 * e(2 * x0) + x4 * e(2 * x1)),
           (e(x1 + 2 * x0) * (x7 * s(x3) - x6 * c(x3)) + e(x0 + 2 * x1) * (x5 * s(x2) - x4 * c(x2))),
           0.0]
    d_p = [0.0,
           -2 * e(x1) * c(x3) - 2 * e(x0) * c(x2),
           4 * c(x3) * c(x2) * e(x0 + x1) + e(2 * x1) + e(2 * x0),
           -2 * c(x2) * e(x0 + 2 * x1) - 2 * c(x3) * e(x1 + 2 * x0),
           e(2 * x0 + 2 * x1)]
    n_m = [0.0] + [n_p[i] - d_p[i] * n_p[0] for i in range(1, 5)]
    sum_n_p = sum_n_m = sum_d = 0.0
    for i in range(5):
        sum_n_p += n_p[i]
        sum_n_m += n_m[i]
        sum_d += d_p[i]
    a, b = sum_n_p / (1.0 + sum_d), sum_n_m / (1.0 + sum_d)
    return n_p, n_m, d_p, [d * a for d in d_p], [d * b for d in d_p]


def iir_lines(lines, std_dev):
    """One pass of gauss_iir over `lines` [count, n] float32 (gauss.c:160-196): causal and anticausal 4th-order recursions with
    double accumulators, every product and sum rounded on its own, in the reference's order of terms; float(vp + vm)."""
    n_p, n_m, d_p, bd_p, bd_m = iir_constants(std_dev)
    d_m = d_p
    cnt, n = lines.shape
    s = lines.astype(np.float64)
    vp = np.zeros((cnt, n))
    vm = np.zeros((cnt, n))
    init_p, init_m = s[:, 0], s[:, n - 1]
    for k in range(n):
        terms = min(k, 4)
        acc_p = np.zeros(cnt)
        acc_m = np.zeros(cnt)
        kp, km = k, n - 1 - k
        for i in range(terms + 1):
            prev_p = acc_p if i == 0 else vp[:, kp - i]
            prev_m = acc_m if i == 0 else vm[:, km + i]
            acc_p = acc_p + (n_p[i] * s[:, kp - i] - d_p[i] * prev_p)
            acc_m = acc_m + (n_m[i] * s[:, km + i] - d_m[i] * prev_m)
        for j in range(terms + 1, 5):
            acc_p = acc_p + (n_p[j] - bd_p[j]) * init_p
            acc_m = acc_m + (n_m[j] - bd_m[j]) * init_m
        vp[:, kp] = acc_p
        vm[:, km] = acc_m
    return (vp + vm).astype(np.float32)


@pytest.mark.parametrize("w,h,sh,sv", [(37, 23, 2.5, 1.7), (16, 41, 0.8, 6.0), (5, 3, 1.0, 1.0), (64, 9, 12.0, 0.6)])
def test_oracle_gaussian_iir_matches_hand_evaluation(w, h, sh, sv):
    """The oracle's gauss.c restatement (what the CUDA blur is compared with, raw float bits) against a numpy evaluation written
    from the reference's source: columns first, then rows, per channel; equal bits."""
    import ctypes
    rng = np.random.default_rng(w * 100 + h)
    data = rng.random((h, w, 4), dtype=np.float32)
    data[rng.random((h, w)) < 0.25] = 0.0
    got = np.ascontiguousarray(data.copy())
    olib = OracleFilter(mb.Module(source="filter f () rgba:[1,0,0,1] end").ir).lib
    olib.mmo_gaussian_blur_floats.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float]
    olib.mmo_gaussian_blur_floats.restype = None
    olib.mmo_gaussian_blur_floats(got.ctypes.data, w, h, sh, sv)
    want = data.copy()
    for ch in range(4):
        want[:, :, ch] = iir_lines(np.ascontiguousarray(want[:, :, ch].T), sv).T  # the vertical pass first (gauss.c:155)
    for ch in range(4):
        want[:, :, ch] = iir_lines(np.ascontiguousarray(want[:, :, ch]), sh)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "max abs difference %g" % np.abs(got - want).max()


# ---- output formats by hand (new_template.c.in:270-293) ---------------------------------------------------------------------
@pytest.mark.parametrize("bpp", [1, 2, 3])
def test_oracle_output_formats_match_hand_evaluation(bpp):
    """1 and 2 bytes per pixel store (R * 0.299 + G * 0.587 + B * 0.114) * 255.0 of the clamped channels, computed in double
    and truncated; 2 and 4 carry alpha in the last byte; 3 is RGB.  Twirl at t = 0.7 on a non-square frame, bilinear."""
    W, H, t = 90, 61, 0.7
    img = synthetic_rgba(W, H, seed=21)
    m = mb.Module(source=filter_source("examples/Distorts/Twirl.mm"))
    got = OracleFilter(m.ir).render(W, H, {"in": img}, t=t, antialiasing=True, bpp=bpp)
    x, y, _, _ = virtual_coords(W, H)
    r = libm(np.hypot, x, y)
    with np.errstate(invalid="ignore", divide="ignore"):
        a = libm(np.arccos, (x / r).astype(np.float32))
    a = np.where(y < 0, (F(2 * math.pi) - a).astype(np.float32), a)
    a = np.where(r == 0, F(0.0), a).astype(np.float32)
    e = ((r / F(math.sqrt(2.0))).astype(np.float32) - F(1.0)).astype(np.float32)
    e = (((e * F(F(t) - F(0.5))).astype(np.float32) * F(4.0)).astype(np.float32) * F(math.pi)).astype(np.float32)
    a2 = (a + e).astype(np.float32)
    tup = sample(img, (libm(np.cos, a2) * (r + F(0.0)).astype(np.float32)).astype(np.float32),
                 (libm(np.sin, a2) * (r + F(0.0)).astype(np.float32)).astype(np.float32), True)
    c = np.where(1.0 < tup, F(1.0), tup)
    c = np.where(0.0 < c, c, F(0.0)).astype(np.float64)  # CLAMP01, then promoted to double by the double literals
    gray = np.floor((c[..., 0] * 0.299 + c[..., 1] * 0.587 + c[..., 2] * 0.114) * 255.0).astype(np.uint8)
    alpha = np.floor(c[..., 3] * 255.0).astype(np.uint8)
    if bpp == 1:
        want = gray[..., None]
    elif bpp == 2:
        want = np.stack([gray, alpha], axis=-1)
    else:
        want = np.floor(c[..., :3] * 255.0).astype(np.uint8)
    assert np.array_equal(got, want), "%d pixels differ" % int((np.abs(got.astype(int) - want.astype(int)).max(axis=2) > 0).sum())


# ---- coordinate systems by hand (mathmap_common.c:59-72, compiler.c:1710-1773, 2311-2420) -----------------------------------
def sample_edges_factors(img, x, y, fx, fy, bilinear):
    """origVal with the resize factors as arguments (None: a stretched image, no factors); edge colour (0, 0, 0, 0)"""
    h, w = img.shape[:2]
    if fx is not None:
        x = (x * fx).astype(np.float32)
        y = (y * fy).astype(np.float32)
    px = ((x + F(1.0)) * F((w - 1) / 2.0)).astype(np.float32)
    py = (-((y - F(1.0)) * F((h - 1) / 2.0))).astype(np.float32)
    texels = img.astype(np.float32)

    def texel(ix, iy):
        inside = (ix >= 0) & (ix < w) & (iy >= 0) & (iy < h)
        v = texels[np.clip(iy, 0, h - 1), np.clip(ix, 0, w - 1)]
        return np.where(inside[..., None], v, F(0.0))

    if not bilinear:
        q = texel(np.floor((px.astype(np.float64) + 0.5).astype(np.float32)).astype(np.int64),
                  np.floor((py.astype(np.float64) + 0.5).astype(np.float32)).astype(np.int64))
    else:
        x1 = np.floor(px).astype(np.int64)
        y1 = np.floor(py).astype(np.int64)
        x2f = (px - x1.astype(np.float32)).astype(np.float32)
        y2f = (py - y1.astype(np.float32)).astype(np.float32)
        x1f = (F(1.0) - x2f).astype(np.float32)
        y1f = (F(1.0) - y2f).astype(np.float32)
        p1, p2, p3, p4 = (x1f * y1f)[..., None], (x1f * y2f)[..., None], (x2f * y1f)[..., None], (x2f * y2f)[..., None]
        c1, c2, c3, c4 = texel(x1, y1), texel(x1, y1 + 1), texel(x1 + 1, y1), texel(x1 + 1, y1 + 1)
        s = (((c1 * p1).astype(np.float32) + (c2 * p2).astype(np.float32)).astype(np.float32) + (c3 * p3).astype(np.float32)).astype(np.float32)
        s = (s + (c4 * p4).astype(np.float32)).astype(np.float32)
        q = np.rint(s)
    return (q.astype(np.float64) / 255.0).astype(np.float32)


def unit_coords(W, H):
    cols = np.arange(W, dtype=np.float64)
    rows = np.arange(H, dtype=np.float64)
    xu = ((cols - (W - 1) / 2.0) / ((W - 1) / 2.0)).astype(np.float32)
    yu = ((-rows + (H - 1) / 2.0) / ((H - 1) / 2.0)).astype(np.float32)
    return xu[None, :].repeat(H, 0), yu[:, None].repeat(W, 1)


COORDINATE_FILTERS = {
    # stretched filter: X = Y = 1 (W = H = 2); stretched image: no resize factors
    "stretched": ("stretched filter s (stretched image in)\n  in(xy * 0.8 + xy:[0.1, -0.05] + xy:[W, H] * 0.01)\nend\n",
                  lambda xu, yu, W, H, w, h: ((((xu * F(1.0)).astype(np.float32) * F(0.8)).astype(np.float32) + F(0.1)).astype(np.float32) + (F(2.0) * F(0.01)),
                                              (((yu * F(1.0)).astype(np.float32) * F(0.8)).astype(np.float32) + F(-0.05)).astype(np.float32) + (F(2.0) * F(0.01)),
                                              None, None)),
    # pixel filter: X = (W - 1) / 2, Y = (H - 1) / 2 in pixels; pixel image: factors 2 / w, 2 / h
    "pixel": ("pixel filter p (pixel image in)\n  in(xy + xy:[3.5, -2.25])\nend\n",
              lambda xu, yu, W, H, w, h: (((xu * (F(W - 1) / F(2))).astype(np.float32) + F(3.5)).astype(np.float32),
                                          ((yu * (F(H - 1) / F(2))).astype(np.float32) + F(-2.25)).astype(np.float32),
                                          F(2) / F(w), F(2) / F(h))),
    # default filter (unit, square: X = W / max, Y = H / max) sampling a pixel image
    "default_on_pixel_image": ("filter m (pixel image in)\n  in(xy * 20)\nend\n",
                               lambda xu, yu, W, H, w, h: (((xu * (F(W) / F(max(W, H)))).astype(np.float32) * F(20)).astype(np.float32),
                                                           ((yu * (F(H) / F(max(W, H)))).astype(np.float32) * F(20)).astype(np.float32),
                                                           F(2) / F(w), F(2) / F(h))),
}


@pytest.mark.parametrize("name", sorted(COORDINATE_FILTERS))
@pytest.mark.parametrize("bilinear", [False, True], ids=["nearest", "bilinear"])
def test_oracle_coordinate_systems_match_hand_evaluation(name, bilinear):
    src, coords = COORDINATE_FILTERS[name]
    W, H, w, h = 71, 40, 53, 64
    img = synthetic_rgba(w, h, seed=17)
    m = mb.Module(source=src)
    got = OracleFilter(m.ir).render(W, H, {"in": img}, t=0.0, antialiasing=bilinear)
    xu, yu = unit_coords(W, H)
    sx, sy, fx, fy = coords(xu, yu, W, H, w, h)
    want = quantise(sample_edges_factors(img, np.asarray(sx, np.float32), np.asarray(sy, np.float32), fx, fy, bilinear))
    diff = np.abs(got.astype(int) - want.astype(int)).max(axis=2)
    assert np.array_equal(got, want), "%d pixels differ, max %d" % (int((diff > 0).sum()), int(diff.max()))


# ---- the builtins' guards by hand (builtins.lisp:635-669, 804-817, 848-857, 873-876) ----------------------------------------
def test_oracle_builtin_guards_match_hand_evaluation():
    """x / 0 = 0, x % 0 = 0, asin outside [-1, 1] = 0, log(x <= 0) = 0; `%` is fmod in double; floor() is an int."""
    src = ("filter g ()\n  k = floor(y * 3);\n"
           "  rgba:[x / k / 8 + 0.5, asin(x * 2) / 3 + 0.5, log(x) / 6 + 0.5, (x * 4) % k / 8 + 0.5]\nend\n")
    W, H = 97, 64
    m = mb.Module(source=src)
    got = OracleFilter(m.ir).render(W, H, {}, t=0.0)
    x, y, _, _ = virtual_coords(W, H)
    k = np.floor((y * F(3)).astype(np.float32)).astype(np.int32).astype(np.float32)  # int, converted back where a float is needed
    with np.errstate(all="ignore"):
        c0 = np.where(k == 0, F(0.0), (x / k).astype(np.float32))
        x2 = (x * F(2)).astype(np.float32)
        c1 = np.where((x2 < -1) | (x2 > 1), F(0.0), libm(np.arcsin, x2))
        c2 = np.where(x <= 0, F(0.0), libm(np.log, x))
        x4 = (x * F(4)).astype(np.float32)
        c3 = np.where(k == 0, F(0.0), np.fmod(x4.astype(np.float64), k.astype(np.float64)).astype(np.float32))
    ch = [((c0 / F(8)).astype(np.float32) + F(0.5)).astype(np.float32), ((c1 / F(3)).astype(np.float32) + F(0.5)).astype(np.float32),
          ((c2 / F(6)).astype(np.float32) + F(0.5)).astype(np.float32), ((c3 / F(8)).astype(np.float32) + F(0.5)).astype(np.float32)]
    want = quantise(np.stack(ch, axis=-1).astype(np.float32))
    diff = np.abs(got.astype(int) - want.astype(int)).max(axis=2)
    assert np.array_equal(got, want), "%d pixels differ, max %d" % (int((diff > 0).sum()), int(diff.max()))


# ---- vector, matrix and interpolation builtins by hand (builtins.lisp:671-735, 1026-1090, 1173-1228) ------------------------
def test_oracle_vector_builtins_match_hand_evaluation():
    """crossp, normalize, dotp, det (2x2, 3x3), gray, pmod, clamp, lerp, scale, sign, inintv: every product and sum a float op,
    sums and products of several terms taken left to right."""
    src = ("filter b (image in)\n  p = in(xy);\n  v = v3:[x, y, 0.5];\n  w = v3:[0.2, x * y, 1];\n  c = crossp(v, w);\n  n = normalize(v);\n"
           "  d3 = det(m3x3:[x, 0.3, y, 0.1, y, 0.2, 0.7, 0.4, x]);\n  d2 = det(m2x2:[x, y, 0.5, 0.25]);\n  g = gray(p);\n"
           "  rgba:[scale(dotp(c, n), -1, 1, 0, 1), lerp(g, pmod(x * 3, 0.7), clamp(d3, -0.2, 0.6)), sign(d2) * 0.25 + 0.5, inintv(g, 0.3, 0.6) * 0.5 + 0.25]\nend\n")
    W, H = 101, 75
    img = synthetic_rgba(W, H, seed=9)
    m = mb.Module(source=src)
    got = OracleFilter(m.ir).render(W, H, {"in": img}, t=0.0)
    x, y, _, _ = virtual_coords(W, H)
    f = lambda v: np.asarray(v, dtype=np.float32)
    mul = lambda u, v: (f(u) * f(v)).astype(np.float32)
    add = lambda u, v: (f(u) + f(v)).astype(np.float32)
    sub = lambda u, v: (f(u) - f(v)).astype(np.float32)
    div = lambda u, v: (f(u) / f(v)).astype(np.float32)
    p = sample(img, x, y, False)
    half = np.full_like(x, 0.5)
    one = np.ones_like(x)
    xy_ = mul(x, y)
    v = [x, y, half]
    w = [np.full_like(x, F(0.2)), xy_, one]
    c = [sub(mul(v[1], w[2]), mul(v[2], w[1])), sub(mul(v[2], w[0]), mul(v[0], w[2])), sub(mul(v[0], w[1]), mul(v[1], w[0]))]
    l = add(add(mul(x, x), mul(y, y)), mul(half, half))
    root = libm(np.sqrt, l)
    n = [np.where(l == 0, F(0), div(q, root)) for q in v]
    dot = add(add(mul(c[0], n[0]), mul(c[1], n[1])), mul(c[2], n[2]))
    ch0 = add(mul(div(sub(dot, F(-1)), sub(F(1), F(-1))), sub(F(1), F(0))), F(0))
    a = [x, np.full_like(x, F(0.3)), y, np.full_like(x, F(0.1)), y, np.full_like(x, F(0.2)), np.full_like(x, F(0.7)), np.full_like(x, F(0.4)), x]
    t3 = lambda i, j, k: mul(mul(a[i], a[j]), a[k])
    d3 = sub(add(add(t3(0, 4, 8), t3(1, 5, 6)), t3(2, 3, 7)), add(add(t3(2, 4, 6), t3(0, 5, 7)), t3(1, 3, 8)))
    d2 = sub(mul(x, F(0.25)), mul(y, F(0.5)))
    g = add(add(mul(F(0.299), p[..., 0]), mul(F(0.587), p[..., 1])), mul(F(0.114), p[..., 2]))
    x3 = mul(x, F(3))
    mod = np.fmod(x3.astype(np.float64), np.float64(F(0.7))).astype(np.float32)
    pm = np.where(x3 < 0, add(mod, F(0.7)), mod)
    cl = np.where(d3 < F(-0.2), F(-0.2), np.where(F(0.6) < d3, F(0.6), d3)).astype(np.float32)
    ch1 = add(mul(sub(F(1), g), pm), mul(g, cl))
    sg = np.where(d2 < 0, F(-1), np.where(0 < d2, F(1), F(0))).astype(np.float32)
    ch2 = add(mul(sg, F(0.25)), F(0.5))
    ch3 = add(mul(np.where((F(0.3) <= g) & (g <= F(0.6)), F(1), F(0)).astype(np.float32), F(0.5)), F(0.25))
    want = quantise(np.stack([ch0, ch1, ch2, ch3], axis=-1).astype(np.float32))
    diff = np.abs(got.astype(int) - want.astype(int))
    assert np.array_equal(got, want), "pixels differing per channel %r, max %d" % ((diff > 0).sum(axis=(0, 1)).tolist(), int(diff.max()))


# ---- colour conversions by hand (builtins.lisp:1251-1322) -------------------------------------------------------------------
def test_oracle_hsv_conversions_match_hand_evaluation():
    """toHSVA, a hue rotation and desaturation depending on t, toRGBA: max / min with int constants, `= r max` branch order,
    h / 6.0, floor(h * 6) as an int and the six sectors."""
    src = "filter c (image in)\n  h = toHSVA(in(xy));\n  toRGBA(hsva:[h[0] + t, h[1] * 0.75, h[2], h[3]])\nend\n"
    W, H, t = 88, 66, 0.37
    img = synthetic_rgba(W, H, seed=13)
    img[:8] = img[:8, :, :1]  # grey rows: delta = 0 (division guard) and s = 0
    img[8:12] = 0             # black rows: max = 0
    m = mb.Module(source=src)
    got = OracleFilter(m.ir).render(W, H, {"in": img}, t=t)
    x, y, _, _ = virtual_coords(W, H)
    f = lambda v: np.asarray(v, dtype=np.float32)
    mul = lambda u, v: (f(u) * f(v)).astype(np.float32)
    add = lambda u, v: (f(u) + f(v)).astype(np.float32)
    sub = lambda u, v: (f(u) - f(v)).astype(np.float32)
    with np.errstate(all="ignore"):
        div = lambda u, v: np.where(f(v) == 0, F(0), (f(u) / f(v))).astype(np.float32)  # the `/` builtin's guard
        p = sample(img, x, y, False)
        clamp01 = lambda v: np.maximum(F(0), np.minimum(F(1), v)).astype(np.float32)
        r, g, b, al = clamp01(p[..., 0]), clamp01(p[..., 1]), clamp01(p[..., 2]), clamp01(p[..., 3])
        mx = np.maximum(r, np.maximum(g, b))
        mn = np.minimum(r, np.minimum(g, b))
        delta = sub(mx, mn)
        sat = div(delta, mx)
        hh = np.where(r == mx, div(sub(g, b), delta), np.where(g == mx, add(F(2), div(sub(b, r), delta)), add(F(4), div(sub(r, g), delta))))
        hh = div(hh, F(6.0))
        hue = np.where(hh < 0, add(hh, F(1)), hh)
        hue = np.where(mx == 0, F(0), hue).astype(np.float32)
        sat = np.where(mx == 0, F(0), sat).astype(np.float32)
        # the filter's own arithmetic
        h2, s2, v2, a2 = add(hue, F(t)), mul(sat, F(0.75)), mx, al
        # toRGBA
        s = clamp01(s2)
        v = clamp01(v2)
        a_out = clamp01(a2)
        hcl = np.maximum(F(0), h2).astype(np.float32)
        h6 = np.where(F(1) <= hcl, F(0), mul(hcl, F(6))).astype(np.float32)
        i = np.floor(h6).astype(np.int32)
        fr = sub(h6, i.astype(np.float32))
        pp = mul(v, sub(F(1), s))
        qq = mul(v, sub(F(1), mul(s, fr)))
        tt = mul(v, sub(F(1), mul(s, sub(F(1), fr))))
        sel = lambda c0, c1, c2, c3, c4, c5: np.where(i == 0, c0, np.where(i == 1, c1, np.where(i == 2, c2, np.where(i == 3, c3, np.where(i == 4, c4, c5)))))
        R = np.where(s == 0, v, sel(v, qq, pp, pp, tt, v))
        G = np.where(s == 0, v, sel(tt, v, v, qq, pp, pp))
        B = np.where(s == 0, v, sel(pp, pp, tt, v, v, qq))
    want = quantise(np.stack([R, G, B, a_out], axis=-1).astype(np.float32))
    diff = np.abs(got.astype(int) - want.astype(int))
    assert np.array_equal(got, want), "pixels differing per channel %r, max %d" % ((diff > 0).sum(axis=(0, 1)).tolist(), int(diff.max()))


# ---- curves and gradients by hand (opmacros.h:128, 192-194) -----------------------------------------------------------------
def test_oracle_curve_and_gradient_lookup_match_hand_evaluation():
    """values[(int)(CLAMP01(p) * 1023)]: a float product, truncated; gradient colours packed R<<24|G<<16|B<<8|A, channels / 255.0."""
    src = "filter cg (curve c, gradient g)\n  v = c(x * 0.6 + 0.5);\n  g(v * 0.8 + y * 0.3)\nend\n"
    W, H = 120, 90
    curve = (np.sqrt(np.arange(1024, dtype=np.float64) / 1023.0)).astype(np.float32)
    grad = np.random.RandomState(5).randint(0, 1 << 32, size=1024, dtype=np.uint64).astype(np.uint32)
    m = mb.Module(source=src)
    got = OracleFilter(m.ir).render(W, H, {"c": curve, "g": grad}, t=0.0)
    x, y, _, _ = virtual_coords(W, H)
    clamp01 = lambda v: np.where(0.0 < np.where(1.0 < v, F(1.0), v), np.where(1.0 < v, F(1.0), v), F(0.0)).astype(np.float32)
    index = lambda v: (clamp01(v) * F(1023)).astype(np.float32).astype(np.int64)
    p1 = ((x * F(0.6)).astype(np.float32) + F(0.5)).astype(np.float32)
    v = curve[index(p1)]
    p2 = ((v * F(0.8)).astype(np.float32) + (y * F(0.3)).astype(np.float32)).astype(np.float32)
    col = grad[index(p2)].astype(np.uint64)
    tup = np.stack([(col >> 24) & 255, (col >> 16) & 255, (col >> 8) & 255, col & 255], axis=-1).astype(np.float64)
    want = quantise((tup / 255.0).astype(np.float32))
    diff = np.abs(got.astype(int) - want.astype(int))
    assert np.array_equal(got, want), "pixels differing per channel %r, max %d" % ((diff > 0).sum(axis=(0, 1)).tolist(), int(diff.max()))


# ---- closures and filter calls by hand (opmacros.h:199-216, new_template.c.in:375-422) --------------------------------------
def test_oracle_closures_and_filter_calls_match_hand_evaluation():
    """A filter applied to an image is an image (closure): sampling it runs the filter's body at the sample's coordinates, the
    result stays a float tuple (nothing is quantised in between); a call with explicit coordinates is the same evaluation.
    The coordinates pass through two conversions on the way in: the closure is a unit-square image whose pixel size the
    optimiser takes from its first image argument (compopt/simplify.c:28-45), so they are multiplied by max/w, max/h of THAT
    image (compiler.c:2219, 1747-1760), and the callee then binds its x, y as argument * X, * Y (compiler.c:2406-2420).  All
    goldens are square, where every one of these factors is 1: this non-square case is what pins them."""
    src = ("filter inner (image in, float k: 0-2 (1))\n  in(xy * k)\nend\n\n"
           "filter outer (image in)\n  half = inner(in, 0.5);\n  half(xy + xy:[0.1, 0]) * 0.5 + inner(in, 1.5, xy * 0.5) * 0.25\nend\n")
    W, H, w, h = 93, 58, 70, 81
    img = synthetic_rgba(w, h, seed=23)
    m = mb.Module(source=src)
    x, y, X, Y = virtual_coords(W, H)
    mul = lambda u, v: (np.asarray(u, np.float32) * np.asarray(v, np.float32)).astype(np.float32)
    add = lambda u, v: (np.asarray(u, np.float32) + np.asarray(v, np.float32)).astype(np.float32)
    fx, fy = F(F(max(w, h)) / F(w)), F(F(max(w, h)) / F(h))

    def through_inner(cx, cy, k, bilinear):
        cx, cy = mul(mul(cx, fx), X), mul(mul(cy, fy), Y)
        return sample(img, mul(cx, F(k)), mul(cy, F(k)), bilinear)

    for bilinear in (False, True):
        got = OracleFilter(m.ir).render(W, H, {"in": img}, t=0.0, antialiasing=bilinear)
        first = through_inner(add(x, F(0.1)), add(y, F(0)), 0.5, bilinear)
        second = through_inner(mul(x, F(0.5)), mul(y, F(0.5)), 1.5, bilinear)
        want = quantise(add(mul(first, F(0.5)), mul(second, F(0.25))))
        diff = np.abs(got.astype(int) - want.astype(int))
        assert np.array_equal(got, want), "bilinear=%s: pixels differing per channel %r, max %d" % (bilinear, (diff > 0).sum(axis=(0, 1)).tolist(), int(diff.max()))


# ---- render() of a closure by hand (builtins.c:249-345) ---------------------------------------------------------------------
def test_oracle_render_of_a_closure_matches_hand_evaluation():
    """render(closure): a floatmap of the render size, each pixel the closure evaluated at fx = ((float)x - bx) / ax (float
    arithmetic, ax = bx = (W - 1) / 2, ay = -by) through ORIG_VAL with frame argument 0.0 -- so the callee's t is 0 whatever
    the frame's t is; the lookup afterwards is get_floatmap_pixel's lrintf(ax * x + bx), without resize factors."""
    src = ("filter inner (image in, float k: 0-2 (1))\n  in(xy * k + xy:[t, 0])\nend\n\n"
           "filter outer (image in)\n  rr = render(inner(in, 0.8));\n  rr(xy * 0.9 + xy:[t * 0.1, 0])\nend\n")
    W, H, w, h, t = 77, 52, 64, 90, 0.3
    img = synthetic_rgba(w, h, seed=29)
    m = mb.Module(source=src)
    x, y, X, Y = virtual_coords(W, H)
    mul = lambda u, v: (np.asarray(u, np.float32) * np.asarray(v, np.float32)).astype(np.float32)
    add = lambda u, v: (np.asarray(u, np.float32) + np.asarray(v, np.float32)).astype(np.float32)
    sub = lambda u, v: (np.asarray(u, np.float32) - np.asarray(v, np.float32)).astype(np.float32)
    div = lambda u, v: (np.asarray(u, np.float32) / np.asarray(v, np.float32)).astype(np.float32)
    ax = bx = F((W - 1) / 2.0)
    by = F((H - 1) / 2.0)
    ay = F(-by)
    fx, fy = F(F(max(w, h)) / F(w)), F(F(max(w, h)) / F(h))
    for bilinear in (False, True):
        got = OracleFilter(m.ir).render(W, H, {"in": img}, t=t, antialiasing=bilinear)
        gx = div(sub(np.arange(W, dtype=np.float32), bx), ax)[None, :].repeat(H, 0)
        gy = div(sub(np.arange(H, dtype=np.float32), by), ay)[:, None].repeat(W, 1)
        cx, cy = mul(mul(gx, fx), X), mul(mul(gy, fy), Y)          # the resize wrapper's factors, then the callee's x = arg * X
        rendered = sample(img, add(mul(cx, F(0.8)), F(0.0)), add(mul(cy, F(0.8)), F(0)), bilinear)  # t is 0 inside the rendering
        lx = add(mul(x, F(0.9)), mul(F(t), F(0.1)))
        ly = add(mul(y, F(0.9)), F(0))
        ix = np.rint(add(mul(ax, lx), bx)).astype(np.int64)
        iy = np.rint(add(mul(ay, ly), by)).astype(np.int64)
        inside = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
        tup = np.where(inside[..., None], rendered[np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)], F(0))
        want = quantise(tup.astype(np.float32))
        diff = np.abs(got.astype(int) - want.astype(int))
        assert np.array_equal(got, want), "bilinear=%s: pixels differing per channel %r, max %d" % (bilinear, (diff > 0).sum(axis=(0, 1)).tolist(), int(diff.max()))


# ---- complex arithmetic by hand (builtins.lisp:485-496, 599-611, 880-886) ---------------------------------------------------
def test_oracle_complex_arithmetic_matches_hand_evaluation():
    """Complex product, quotient (with its all-zero guard; inner `/` is the raw float division), conjugate, magnitude (hypot in
    double) -- Droste's building blocks, lowered to float ops in the order the builtins spell them."""
    src = ("filter z ()\n  p = ri:[x * 2, y * 3];\n  q = ri:[floor(x * 3), floor(y * 2)];\n  m = p * q;\n  d = p / q;\n  c = conj(d) * 0.25;\n"
           "  rgba:[m[0] / 16 + 0.5, m[1] / 16 + 0.5, c[0] + c[1] + 0.5, abs(p) / 4]\nend\n")
    W, H = 110, 70
    m = mb.Module(source=src)
    got = OracleFilter(m.ir).render(W, H, {}, t=0.0)
    x, y, _, _ = virtual_coords(W, H)
    f = lambda v: np.asarray(v, dtype=np.float32)
    mul = lambda u, v: (f(u) * f(v)).astype(np.float32)
    add = lambda u, v: (f(u) + f(v)).astype(np.float32)
    sub = lambda u, v: (f(u) - f(v)).astype(np.float32)
    with np.errstate(all="ignore"):
        rdiv = lambda u, v: (f(u) / f(v)).astype(np.float32)
        gdiv = lambda u, v: np.where(f(v) == 0, F(0), rdiv(u, v)).astype(np.float32)  # the language-level `/`
        p0, p1 = mul(x, F(2)), mul(y, F(3))
        q0 = np.floor(mul(x, F(3))).astype(np.int32).astype(np.float32)
        q1 = np.floor(mul(y, F(2))).astype(np.int32).astype(np.float32)
        m0 = sub(mul(p0, q0), mul(p1, q1))
        m1 = add(mul(p0, q1), mul(q0, p1))
        cc = add(mul(q0, q0), mul(q1, q1))
        d0 = rdiv(add(mul(p0, q0), mul(p1, q1)), cc)
        d1 = rdiv(add(mul((-p0).astype(np.float32), q1), mul(q0, p1)), cc)
        zero = (q0 == 0) & (q1 == 0)
        d0 = np.where(zero, F(0), d0).astype(np.float32)
        d1 = np.where(zero, F(0), d1).astype(np.float32)
        c0, c1 = mul(d0, F(0.25)), mul((-d1).astype(np.float32), F(0.25))
        ch = [add(gdiv(m0, F(16)), F(0.5)), add(gdiv(m1, F(16)), F(0.5)), add(add(c0, c1), F(0.5)), gdiv(libm(np.hypot, p0, p1), F(4))]
    want = quantise(np.stack(ch, axis=-1).astype(np.float32))
    diff = np.abs(got.astype(int) - want.astype(int))
    assert np.array_equal(got, want), "pixels differing per channel %r, max %d" % ((diff > 0).sum(axis=(0, 1)).tolist(), int(diff.max()))


# ---- the Gaussian blur's run-length FIR by hand (native-filters/gauss.c:264-633) --------------------------------------------
def rle_curve(sigma):
    """make_rle_curve: float curve and its running sums; note that `total` leaves the curve's last tap out."""
    sigma = float(np.float32(sigma))
    sigma2 = 2 * sigma * sigma
    l = math.sqrt(-sigma2 * math.log(1.0 / 255.0))
    n = int(math.ceil(l)) * 2
    if n % 2 == 0:
        n += 1
    length = n // 2
    curve = {0: F(1.0)}
    for i in range(1, length + 1):
        curve[i] = curve[-i] = F(math.exp(-(i * i) / sigma2))
    sums = [F(0)]
    for i in range(1, 2 * length + 1):
        sums.append(F(curve[i - length - 1] + sums[i - 1]))
    csum = {k - length: sums[k] for k in range(2 * length + 1)}
    return curve, csum, length, F(csum[length] - csum[-length])


def fir_line(line, sigma):
    """One line through run_length_encode + do_encoded_lre / do_full_lre, in float32 scalars, in the reference's order."""
    n = len(line)
    curve, csum, length, total = rle_curve(sigma)
    pix, rle = {}, {}
    idx = n + length - 1
    last, count, same = line[n - 1], 0, 0
    for _ in range(length):                     # the 'end' border
        count += 1
        pix[idx], rle[idx] = last, count
        idx -= 1
    for k in range(n - 1, -1, -1):              # the real pixels, from the right
        c = line[k]
        if c == last:
            count += 1
            same += 1
        else:
            count, last = 1, c
        pix[idx], rle[idx] = last, count
        idx -= 1
    for _ in range(length):                     # the start border
        count += 1
        pix[idx], rle[idx] = last, count
        idx -= 1
    out = np.empty(n, np.float32)
    if same > (3 * n) // 4:
        ctotal = int(total)                     # do_encoded_lre takes `int ctotal`
        for col in range(n):
            val = F(0.0)
            pos = col - length
            s1 = csum[-length]
            nb = rle[pos]
            i = -length + nb
            while i <= length:
                s2 = int(csum[i])               # `int s2 = csum[i]`
                val = F(val + F(pix[pos] * F(F(s2) - s1)))
                s1 = F(s2)
                pos += nb
                nb = rle[pos]
                i += nb
            val = F(val + F(pix[pos] * F(csum[length] - s1)))
            out[col] = F(val / F(ctotal))
    else:
        for col in range(n):
            val = F(F(0.0) + F(pix[col] * curve[0]))
            for i in range(1, length + 1):
                val = F(val + F(F(pix[col + i] + pix[col - i]) * curve[i]))
            out[col] = F(val / total)
    return out


@pytest.mark.parametrize("w,h,sh,sv,runs", [(23, 17, 0.45, 0.3, False), (19, 26, 0.3, 3.0, False), (40, 12, 2.0, 0.4, True), (31, 9, 0.0, 0.45, False)])
def test_oracle_gaussian_fir_matches_hand_evaluation(w, h, sh, sv, runs):
    """Either sigma below half a pixel sends both axes through gauss_rle (gauss.c:659-662): per line a run-length encoding with
    replicated borders, then the plain FIR -- or, when more than three quarters of the samples repeat their right neighbour, the
    run-length variant with its int-truncated running sums and int total.  A sigma of 0 skips its pass.  Raw float bits."""
    import ctypes
    rng = np.random.default_rng(w * 100 + h)
    data = rng.random((h, w, 4), dtype=np.float32)
    if runs:
        data[:, 3:] = data[:, 2:3]              # long runs along the rows: the encoded variant in the horizontal pass
        data[:, :, 1] = np.float32(200.0)       # values above 1: the int truncations show
    got = np.ascontiguousarray(data.copy())
    olib = OracleFilter(mb.Module(source="filter f () rgba:[1,0,0,1] end").ir).lib
    olib.mmo_gaussian_blur_floats.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float]
    olib.mmo_gaussian_blur_floats.restype = None
    olib.mmo_gaussian_blur_floats(got.ctypes.data, w, h, sh, sv)
    want = data.copy()
    if sv > 0.0:
        for col in range(w):
            for ch in range(4):
                want[:, col, ch] = fir_line(want[:, col, ch].copy(), sv)
    if sh > 0.0:
        for row in range(h):
            for ch in range(4):
                want[row, :, ch] = fir_line(want[row, :, ch].copy(), sh)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "max abs difference %g" % np.abs(got - want).max()


# ---- config 4 end to end by hand: Blur/Gaussian Blur.mm on a non-square frame (gauss.c:643-670, builtins.c:269-345) ----------
@pytest.mark.parametrize("bilinear", [False, True], ids=["nearest", "bilinear"])
def test_oracle_gaussian_blur_filter_matches_hand_evaluation(bilinear):
    """The whole filter: sigmas from dev * maxdim / size (guarded `/`), the drawable resampled into a floatmap of the render size
    with the NEAREST sampler whatever -i says (builtins.c:306) at fx = ((float)x - bx) / ax, sigma_px = |s * ax|, |s * ay|, the
    recursive filter columns first, the lrintf lookup at the stretched filter's own coordinates, quantisation."""
    W, H, w, h, dev = 75, 48, 60, 70, 0.05
    img = synthetic_rgba(w, h, seed=31)
    m = mb.Module(source=filter_source("examples/Blur/Gaussian Blur.mm"))
    got = OracleFilter(m.ir).render(W, H, {"in": img, "dev": dev}, t=0.0, antialiasing=bilinear)
    f = lambda v: np.asarray(v, dtype=np.float32)
    maxdim = F(max(w, h))
    hs = F(F(F(dev) * maxdim) / F(w))
    vs = F(F(F(dev) * maxdim) / F(h))
    ax = bx = F((W - 1) / 2.0)
    by = F((H - 1) / 2.0)
    ay = F(-by)
    gx = ((np.arange(W, dtype=np.float32) - bx) / ax).astype(np.float32)[None, :].repeat(H, 0)
    gy = ((np.arange(H, dtype=np.float32) - by) / ay).astype(np.float32)[:, None].repeat(W, 1)
    fm = sample_edges_factors(img, gx, gy, None, None, False)              # stretched image: no factors; nearest
    sigma_h, sigma_v = abs(F(hs * ax)), abs(F(vs * ay))
    assert sigma_h >= 0.5 and sigma_v >= 0.5
    for ch in range(4):
        fm[:, :, ch] = iir_lines(np.ascontiguousarray(fm[:, :, ch].T), sigma_v).T
    for ch in range(4):
        fm[:, :, ch] = iir_lines(np.ascontiguousarray(fm[:, :, ch]), sigma_h)
    xu, yu = unit_coords(W, H)                                             # a stretched filter's x, y are the unit coordinates
    ix = np.rint((f(ax * xu) + bx).astype(np.float32)).astype(np.int64)
    iy = np.rint((f(ay * yu) + by).astype(np.float32)).astype(np.int64)
    inside = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
    tup = np.where(inside[..., None], fm[np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)], F(0))
    want = quantise(tup.astype(np.float32))
    diff = np.abs(got.astype(int) - want.astype(int))
    assert np.array_equal(got, want), "pixels differing per channel %r, max %d" % ((diff > 0).sum(axis=(0, 1)).tolist(), int(diff.max()))
