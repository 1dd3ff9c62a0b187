// Double-precision sin / cos / asin / acos / atan / atan2 / exp / log for FLOAT arguments, for the precise math mode.
//
// The reference's ops call the host's double libm on a float argument and store the result in a float
// (ops.lisp:126-147: `sin((double)x)` etc.), so what the device has to reproduce is RN_float(f(x)) for float x; any
// double evaluation of f with an error of a few double ulps gives that float except when f(x) lies within ~2^-50
// (relative) of the midpoint of two floats, i.e. for about one argument in 10^7.  CUDA's own double sin/cos/acos meet
// that bar, but their coefficients are literals in the code: every double literal costs two UMOV/MOV issue slots in
// SASS, and these kernels are bound by instruction issue (Twirl: 59 UMOV per pixel out of 451 instructions).  The
// functions below read their coefficients from __constant__ tables instead (one LDCU.128 per TWO coefficients), do the
// quadrant selection after narrowing to float, and skip the generic entry checks that float arguments cannot trigger.
//
// Coefficients: near-minimax polynomials fitted in 60-digit arithmetic by tools/gen_dlibm_tables.py (Chebyshev-node
// interpolation); approximation errors in units of 2^-53 relative: sin 0.18, cos 0.012, asin 0.50, atan 0.18, exp 0.14, log 0.005.
// PINNED: tests/tools/dlibm_check.cpp compiles this header for the host and compares RN_float of every function with
// the host's double libm narrowed to float (tests/test_dlibm.py runs a prime-stride sample of all float bit patterns; over
// all of them sin, cos, acos, atan, exp and log never differ, asin for 2 arguments and atan2 for 2 of 8.6e9 pairs by one ulp).
#pragma once

#ifndef __CUDACC_RTC__
#include <cmath>
#include <cstdint>
#include <cstring>
#endif

#ifdef __CUDACC__
#define MM_D_FN __device__ __forceinline__
#define MM_D_TABLE static __constant__
#define MM_D_FMA(a, b, c) __fma_rn(a, b, c)
#define MM_D_LO(d) __double2loint(d)
#define MM_D_HI(d) __double2hiint(d)
#define MM_D_MAKE(hi, lo) __hiloint2double(hi, lo)
#else
#define MM_D_FN static inline
#define MM_D_TABLE static const
#define MM_D_FMA(a, b, c) fma(a, b, c)
static inline int mm_d_lo_host(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u & 0xffffffffu); }
#define MM_D_LO(d) mm_d_lo_host(d)
static inline int mm_d_hi_host(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u >> 32); }
static inline double mm_d_make_host(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &u, 8); return d; }
#define MM_D_HI(d) mm_d_hi_host(d)
#define MM_D_MAKE(hi, lo) mm_d_make_host(hi, lo)
#endif

// [0..3]   2/pi, and pi/2 split in three (Cody-Waite; the products are exact inside the fmas)
// [4..9]   sin:  sin(r) = r + r z S(z), z = r^2, |r| <= pi/4
// [10..15] cos:  cos(r) = 1 - z/2 + z^2 C(z)
MM_D_TABLE double mm_d_trig[16] = {
    0x1.45f306dc9c883p-1, 0x1.921fb54442d18p+0, 0x1.1a62633145c00p-54, 0x1.b839a252049c0p-104,
    -0x1.5555555555555p-3, 0x1.1111111110bb2p-7, -0x1.a01a019e83816p-13, 0x1.71de379654304p-19, -0x1.ae600aca92c64p-26, 0x1.5e0b05bff20f2p-33,
    0x1.5555555555555p-5, -0x1.6c16c16c16967p-10, 0x1.a01a019f4e9b4p-16, -0x1.27e4fa17bf139p-22, 0x1.1eeb68cd22f56p-29, -0x1.907d8f29fe831p-37};

// asin(x) = x + x z A(z), z = x^2 <= 1/4; [12], [13] = pi/2 hi, lo
MM_D_TABLE double mm_d_asin[14] = {
    0x1.555555555554fp-3, 0x1.3333333336df1p-4, 0x1.6db6db6844effp-5, 0x1.f1c71f9893729p-6, 0x1.6e8b2ac4e2b69p-6, 0x1.1c5945b94f7fdp-6,
    0x1.c871802719808p-7, 0x1.8529ff2c9d7d9p-7, 0x1.ff173d5fcbafcp-8, 0x1.06f2e84e31c1bp-6, -0x1.61041ccc64457p-7, 0x1.cdd3a521b6ed7p-6,
    0x1.921fb54442d18p+0, 0x1.1a62633145c07p-54};

// sin and cos of a float with 2^-27 <= |x| < 2^31, both narrowed to float.
MM_D_FN void mm_d_sincos_core(float xf, float &s, float &c) {
    const double x = (double)xf;
    // j = rint(x * 2/pi) through the 1.5 * 2^52 constant: the low word of t holds j as an integer
    const double t = MM_D_FMA(x, mm_d_trig[0], 6755399441055744.0);
    const int q = MM_D_LO(t);
    const double j = t - 6755399441055744.0;
    double r = MM_D_FMA(-j, mm_d_trig[1], x);
    r = MM_D_FMA(-j, mm_d_trig[2], r);
    r = MM_D_FMA(-j, mm_d_trig[3], r);
    const double z = r * r;
    double ps = MM_D_FMA(z, mm_d_trig[9], mm_d_trig[8]);
    double pc = MM_D_FMA(z, mm_d_trig[15], mm_d_trig[14]);
    ps = MM_D_FMA(z, ps, mm_d_trig[7]);
    pc = MM_D_FMA(z, pc, mm_d_trig[13]);
    ps = MM_D_FMA(z, ps, mm_d_trig[6]);
    pc = MM_D_FMA(z, pc, mm_d_trig[12]);
    ps = MM_D_FMA(z, ps, mm_d_trig[5]);
    pc = MM_D_FMA(z, pc, mm_d_trig[11]);
    ps = MM_D_FMA(z, ps, mm_d_trig[4]);
    pc = MM_D_FMA(z, pc, mm_d_trig[10]);
    const double sr = MM_D_FMA(r * z, ps, r);
    const double cr = MM_D_FMA(z * z, pc, MM_D_FMA(z, -0.5, 1.0));
    // quadrant: swap for odd j, then the signs; exact on the narrowed values (negation commutes with rounding)
    const float fs = (float)sr, fc = (float)cr;
    const float a = (q & 1) ? fc : fs, b = (q & 1) ? fs : fc;
    s = (q & 2) ? -a : a;
    c = ((q + 1) & 2) ? -b : b;
}

// One of the two: only the polynomial the quadrant asks for is evaluated (neighbouring pixels mostly agree on it).
MM_D_FN float mm_d_sin_or_cos(float xf, int shift) {  // shift 0: sin, 1: cos
    const double x = (double)xf;
    const double t = MM_D_FMA(x, mm_d_trig[0], 6755399441055744.0);
    const int q = MM_D_LO(t) + shift;
    const double j = t - 6755399441055744.0;
    double r = MM_D_FMA(-j, mm_d_trig[1], x);
    r = MM_D_FMA(-j, mm_d_trig[2], r);
    r = MM_D_FMA(-j, mm_d_trig[3], r);
    const double z = r * r;
    double v;
    if (q & 1) {
        double pc = MM_D_FMA(z, mm_d_trig[15], mm_d_trig[14]);
        pc = MM_D_FMA(z, pc, mm_d_trig[13]);
        pc = MM_D_FMA(z, pc, mm_d_trig[12]);
        pc = MM_D_FMA(z, pc, mm_d_trig[11]);
        pc = MM_D_FMA(z, pc, mm_d_trig[10]);
        v = MM_D_FMA(z * z, pc, MM_D_FMA(z, -0.5, 1.0));
    } else {
        double ps = MM_D_FMA(z, mm_d_trig[9], mm_d_trig[8]);
        ps = MM_D_FMA(z, ps, mm_d_trig[7]);
        ps = MM_D_FMA(z, ps, mm_d_trig[6]);
        ps = MM_D_FMA(z, ps, mm_d_trig[5]);
        ps = MM_D_FMA(z, ps, mm_d_trig[4]);
        v = MM_D_FMA(r * z, ps, r);
    }
    const float f = (float)v;
    return (q & 2) ? -f : f;
}
MM_D_FN float mm_d_sin_core(float xf) { return mm_d_sin_or_cos(xf, 0); }
MM_D_FN float mm_d_cos_core(float xf) { return mm_d_sin_or_cos(xf, 1); }

// acos / asin of a float in [-1, 1], narrowed to float.  |x| < 0.5: the asin series; else through
// asin(sqrt((1 - |x|) / 2)) like fdlibm's e_acos.c / e_asin.c, with sqrt correctly rounded.
MM_D_FN double mm_d_asin_poly(double z) {
    double p = MM_D_FMA(z, mm_d_asin[11], mm_d_asin[10]);
    p = MM_D_FMA(z, p, mm_d_asin[9]);
    p = MM_D_FMA(z, p, mm_d_asin[8]);
    p = MM_D_FMA(z, p, mm_d_asin[7]);
    p = MM_D_FMA(z, p, mm_d_asin[6]);
    p = MM_D_FMA(z, p, mm_d_asin[5]);
    p = MM_D_FMA(z, p, mm_d_asin[4]);
    p = MM_D_FMA(z, p, mm_d_asin[3]);
    p = MM_D_FMA(z, p, mm_d_asin[2]);
    p = MM_D_FMA(z, p, mm_d_asin[1]);
    p = MM_D_FMA(z, p, mm_d_asin[0]);
    return p;
}
MM_D_FN float mm_d_acos_core(float xf) {  // |xf| <= 1
    const double x = (double)xf, ax = fabs(x);
    if (ax < 0.5) {
        const double z = x * x;
        const double as = MM_D_FMA(x * z, mm_d_asin_poly(z), x);  // asin(x)
        return (float)((mm_d_asin[12] - as) + mm_d_asin[13]);
    }
    const double z = (1.0 - ax) * 0.5;  // exact
    const double sq = sqrt(z);
    const double as = MM_D_FMA(sq * z, mm_d_asin_poly(z), sq);  // asin(sqrt(z)) = acos(|x|) / 2
    const double r = as + as;
    return (float)(xf < 0.0f ? (mm_d_asin[12] - r) + (mm_d_asin[12] + mm_d_asin[13] * 2.0) : r);
}
MM_D_FN float mm_d_asin_core(float xf) {  // |xf| <= 1
    const double x = (double)xf, ax = fabs(x);
    if (ax < 0.5) {
        const double z = x * x;
        return (float)MM_D_FMA(x * z, mm_d_asin_poly(z), x);
    }
    const double z = (1.0 - ax) * 0.5;
    const double sq = sqrt(z);
    const double as = MM_D_FMA(sq * z, mm_d_asin_poly(z), sq);
    const double r = (mm_d_asin[12] - (as + as)) + mm_d_asin[13];  // pi/2 - 2 asin(sqrt(z))
    return (float)(xf < 0.0f ? -r : r);
}

// exp:  [0] log2(e), [1], [2] ln 2 hi, lo, [3..12]: exp(r) = 1 + r + r^2 E(r), |r| <= ln2 / 2 (0.14 units of 2^-53)
MM_D_TABLE double mm_d_exp[14] = {
    0x1.71547652b82fep+0, 0x1.62e42fefa39efp-1, 0x1.abc9e3b39803fp-56,
    0x1.0000000000001p-1, 0x1.5555555555556p-3, 0x1.5555555553d63p-5, 0x1.11111111109b3p-7, 0x1.6c16c1788bd90p-10, 0x1.a01a01a7c41d5p-13,
    0x1.a019b90d2ae7ap-16, 0x1.71de0dae63bb3p-19, 0x1.289185613a3d6p-22, 0x1.af38a9b0ec855p-26, 0.0};
// exp of a float in [-104, 89] (or NaN), narrowed to float; outside that range the float result is 0 or +inf
MM_D_FN float mm_d_exp_core(float xf) {
    const double x = (double)xf;
    const double t = MM_D_FMA(x, mm_d_exp[0], 6755399441055744.0);
    const int k = MM_D_LO(t);
    const double kd = t - 6755399441055744.0;
    double r = MM_D_FMA(-kd, mm_d_exp[1], x);
    r = MM_D_FMA(-kd, mm_d_exp[2], r);
    double e = MM_D_FMA(r, mm_d_exp[12], mm_d_exp[11]);
    e = MM_D_FMA(r, e, mm_d_exp[10]);
    e = MM_D_FMA(r, e, mm_d_exp[9]);
    e = MM_D_FMA(r, e, mm_d_exp[8]);
    e = MM_D_FMA(r, e, mm_d_exp[7]);
    e = MM_D_FMA(r, e, mm_d_exp[6]);
    e = MM_D_FMA(r, e, mm_d_exp[5]);
    e = MM_D_FMA(r, e, mm_d_exp[4]);
    e = MM_D_FMA(r, e, mm_d_exp[3]);
    const double p = MM_D_FMA(r * r, e, r) + 1.0;
    return (float)(p * MM_D_MAKE((k + 1023) << 20, 0));  // times 2^k, exact in double; the narrowing rounds once
}

// log:  [0], [1] ln 2 hi, lo, [2..9]: log((1+f)/(1-f)) = 2 f + 2 f s L(s), s = f^2, |f| <= 0.1716 (0.005 units of 2^-53)
MM_D_TABLE double mm_d_log[10] = {
    0x1.62e42fefa39efp-1, 0x1.abc9e3b39803fp-56,
    0x1.5555555555555p-2, 0x1.9999999999a39p-3, 0x1.2492492476a1ap-3, 0x1.c71c7201a55d7p-4, 0x1.745cf8e4bba1bp-4, 0x1.3b1c3c1c81c8fp-4,
    0x1.0fbde0f4ad17bp-4, 0x1.0c0aff044a970p-4};
// log of a positive finite float, narrowed to float
MM_D_FN float mm_d_log_core(float xf) {
    const double x = (double)xf;  // a normal double also for subnormal floats
    int hi = MM_D_HI(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;  // m in [1, 2)
    if (hi >= 0x3ff6a09f) { hi -= 0x00100000; e += 1; }  // m > 1.41421 (between floats): halve it, m in [0.7071, 1.41421]
    const double m = MM_D_MAKE(hi, MM_D_LO(x));
    const double f = (m - 1.0) / (m + 1.0);  // both exact: m carries a float's 24 bits
    const double s = f * f;
    double l = MM_D_FMA(s, mm_d_log[9], mm_d_log[8]);
    l = MM_D_FMA(s, l, mm_d_log[7]);
    l = MM_D_FMA(s, l, mm_d_log[6]);
    l = MM_D_FMA(s, l, mm_d_log[5]);
    l = MM_D_FMA(s, l, mm_d_log[4]);
    l = MM_D_FMA(s, l, mm_d_log[3]);
    l = MM_D_FMA(s, l, mm_d_log[2]);
    const double lm = MM_D_FMA(f * s, l, f);  // log(m) / 2
    const double ed = (double)e;
    return (float)MM_D_FMA(ed, mm_d_log[0], MM_D_FMA(ed, mm_d_log[1], lm + lm));
}

// atan:  [0..10]: atan(t) = t + t z T(z), z = t^2, |t| <= 7/16 (0.18 units of 2^-53); [11..18]: atan(0.5), atan(1), atan(1.5), pi/2 as hi, lo
MM_D_TABLE double mm_d_atan[20] = {
    -0x1.5555555555554p-2, 0x1.999999999883ap-3, -0x1.24924923aec39p-3, 0x1.c71c713562334p-4, -0x1.745cff3f33e2cp-4, 0x1.3b115eae24108p-4,
    -0x1.10ecfa0fc7230p-4, 0x1.df095242d93dep-5, -0x1.9c566526c99f1p-5, 0x1.35c0ac6fb5058p-5, -0x1.1f3fc77b05525p-6,
    0x1.dac670561bb4fp-2, 0x1.a2b7f222f65e2p-56, 0x1.921fb54442d18p-1, 0x1.1a62633145c07p-55, 0x1.f730bd281f69bp-1, 0x1.007887af0cbbdp-56,
    0x1.921fb54442d18p+0, 0x1.1a62633145c07p-54, 0.0};
// atan of a non-negative double (the argument reduction of fdlibm's s_atan.c); infinity gives pi/2
MM_D_FN double mm_d_atan_pos(double ax) {
    double t = ax, hi = 0.0, lo = 0.0;
    if (ax >= 0.4375) {
        double num, den;
        if (ax < 1.1875) {
            if (ax < 0.6875) { num = ax + ax - 1.0; den = 2.0 + ax; hi = mm_d_atan[11]; lo = mm_d_atan[12]; }
            else { num = ax - 1.0; den = ax + 1.0; hi = mm_d_atan[13]; lo = mm_d_atan[14]; }
        } else if (ax < 2.4375) { num = ax - 1.5; den = MM_D_FMA(1.5, ax, 1.0); hi = mm_d_atan[15]; lo = mm_d_atan[16]; }
        else { num = -1.0; den = ax; hi = mm_d_atan[17]; lo = mm_d_atan[18]; }
        t = num / den;
    }
    const double z = t * t;
    double p = MM_D_FMA(z, mm_d_atan[10], mm_d_atan[9]);
    p = MM_D_FMA(z, p, mm_d_atan[8]);
    p = MM_D_FMA(z, p, mm_d_atan[7]);
    p = MM_D_FMA(z, p, mm_d_atan[6]);
    p = MM_D_FMA(z, p, mm_d_atan[5]);
    p = MM_D_FMA(z, p, mm_d_atan[4]);
    p = MM_D_FMA(z, p, mm_d_atan[3]);
    p = MM_D_FMA(z, p, mm_d_atan[2]);
    p = MM_D_FMA(z, p, mm_d_atan[1]);
    p = MM_D_FMA(z, p, mm_d_atan[0]);
    const double r = MM_D_FMA(t * z, p, t);
    return hi + (r + lo);
}
// atan of any float but NaN, narrowed to float
MM_D_FN float mm_d_atan_core(float xf) {
    const float r = (float)mm_d_atan_pos(fabs((double)xf));
    return xf < 0.0f ? -r : (xf == 0.0f ? xf : r);
}
// atan2 of finite non-zero floats, narrowed to float (the quadrant rules of fdlibm's e_atan2.c)
MM_D_FN float mm_d_atan2_core(float yf, float xf) {
    const double z = mm_d_atan_pos(fabs((double)yf / (double)xf));
    double r = z;
    if (xf < 0.0f) r = (mm_d_atan[17] - z) + (mm_d_atan[17] + (mm_d_atan[18] + mm_d_atan[18]));  // pi - z
    const float f = (float)r;
    return yf < 0.0f ? -f : f;
}
