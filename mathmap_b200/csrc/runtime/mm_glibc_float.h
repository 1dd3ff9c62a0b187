// The float libm primitives inside glibc's float complex functions.
//
// The reference evaluates complex ops with glibc's csinf, ctanf, clogf, cargf ... (ops.lisp:196-213).  Those are
// built on glibc's FLOAT primitives, several of which are not correctly rounded (sinhf agrees with the correctly
// rounded value for 74 % of arguments, coshf 78 %, atan2f 84 %, log1pf 92 %), so evaluating them in double and
// narrowing does not reproduce the host's bits.  glibc's versions of these five are the classic fdlibm float
// routines (sysdeps/ieee754/flt-32: s_atanf.c, e_atan2f.c, s_expm1f.c, e_sinhf.c, e_coshf.c, s_log1pf.c; glibc is a
// third-party dependency of the reference, the host here has 2.39).  They are restated below operation by operation
// in float arithmetic, without FMA contraction.  PINNED: tests/tools/glibc_float_check.cpp compiles this very header
// for the host and compares it with the host's libm bit for bit; run over all 2^32 arguments, atanf, expm1f, sinhf,
// coshf and log1pf have no mismatch, atan2f none on 6*10^8 random pairs (tests/test_glibc_float.py runs a sample).
// sinhf/coshf call expf for large arguments; glibc's expf is correctly rounded for 99.94 % of arguments, the device
// uses the correctly rounded value.
#pragma once

#ifndef __CUDACC_RTC__
#include <cmath>
#include <cstdint>
#include <cstring>
#endif
#ifdef __CUDACC__
#define MM_G_FN __host__ __device__ inline
#else
#define MM_G_FN inline
#endif

#if defined(__CUDA_ARCH__)
#define MM_G_BITS(f) ((unsigned)__float_as_int(f))
#define MM_G_FLOAT(u) (__int_as_float((int)(u)))
#define MM_G_EXPF(x) ((float)exp((double)(x)))
#else
static inline unsigned mm_g_bits_host(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float mm_g_float_host(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
#define MM_G_BITS(f) mm_g_bits_host(f)
#define MM_G_FLOAT(u) mm_g_float_host(u)
#define MM_G_EXPF(x) expf(x)
#endif

// s_atanf.c
MM_G_FN float mm_g_atanf(float x) {
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f, -7.6918758452e-02f,
                          6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
    const int hx = (int)MM_G_BITS(x), ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c000000) {  // |x| >= 2^25
        if (ix > 0x7f800000) return x + x;
        return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
    }
    if (ix < 0x3ee00000) {  // |x| < 0.4375
        if (ix < 0x31000000) return x;
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {  // |x| < 1.1875
            if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else { id = 3; x = -1.0f / x; }
        }
    }
    float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return hx < 0 ? -z : z;
}

// e_atan2f.c
MM_G_FN float mm_g_atan2f(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int hx = (int)MM_G_BITS(x), ix = hx & 0x7fffffff, hy = (int)MM_G_BITS(y), iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return mm_g_atanf(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);  // 2*sign(x) + sign(y)
    if (iy == 0) {
        switch (m) {
        case 0: case 1: return y;
        case 2: return pi + tiny;
        default: return -pi - tiny;
        }
    }
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
            case 0: return pi_o_4 + tiny;
            case 1: return -pi_o_4 - tiny;
            case 2: return 3.0f * pi_o_4 + tiny;
            default: return -3.0f * pi_o_4 - tiny;
            }
        }
        switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return pi + tiny;
        default: return -pi - tiny;
        }
    }
    if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = mm_g_atanf(fabsf(y / x));
    switch (m) {
    case 0: return z;
    case 1: return MM_G_FLOAT(MM_G_BITS(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

// s_expm1f.c
MM_G_FN float mm_g_expm1f(float x) {
    const float one = 1.0f, huge = 1.0e+30f, tiny = 1.0e-30f, o_threshold = 8.8721679688e+01f, ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f,
                invln2 = 1.4426950216e+00f, Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f, Q3 = -7.9365076090e-05f, Q4 = 4.0082177293e-06f,
                Q5 = -2.0109921195e-07f;
    float y, hi, lo, c = 0.0f, t, e;
    int k;
    unsigned hx = MM_G_BITS(x);
    const unsigned xsb = hx & 0x80000000u;
    hx &= 0x7fffffffu;
    if (hx >= 0x4195b844u) {  // |x| >= 27 ln2
        if (hx >= 0x42b17218u) {
            if (hx > 0x7f800000u) return x + x;
            if (hx == 0x7f800000u) return xsb == 0 ? x : -1.0f;
            if (x > o_threshold) return huge * huge;
        }
        if (xsb != 0 && x + tiny < 0.0f) return tiny - one;
    }
    if (hx > 0x3eb17218u) {  // |x| > 0.5 ln2
        if (hx < 0x3F851592u) {
            if (xsb == 0) { hi = x - ln2_hi; lo = ln2_lo; k = 1; }
            else { hi = x + ln2_hi; lo = -ln2_lo; k = -1; }
        } else {
            k = (int)(invln2 * x + (xsb == 0 ? 0.5f : -0.5f));
            t = (float)k;
            hi = x - t * ln2_hi;
            lo = t * ln2_lo;
        }
        x = hi - lo;
        c = (hi - x) - lo;
    } else if (hx < 0x33000000u) {  // |x| < 2^-25
        t = huge + x;
        return x - (t - (huge + x));
    } else
        k = 0;
    const float hfx = 0.5f * x, hxs = x * hfx;
    const float r1 = one + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
    t = 3.0f - r1 * hfx;
    e = hxs * ((r1 - t) / (6.0f - x * t));
    if (k == 0) return x - (x * e - hxs);
    e = x * (e - c) - c;
    e -= hxs;
    if (k == -1) return 0.5f * (x - e) - 0.5f;
    if (k == 1) return x < -0.25f ? -2.0f * (e - (x + 0.5f)) : one + 2.0f * (x - e);
    if (k <= -2 || k > 56) {
        y = one - (e - x);
        y = MM_G_FLOAT(MM_G_BITS(y) + ((unsigned)k << 23));
        return y - one;
    }
    if (k < 23) {
        t = MM_G_FLOAT(0x3f800000u - (0x1000000u >> k));  // 1 - 2^-k
        y = t - (e - x);
        y = MM_G_FLOAT(MM_G_BITS(y) + ((unsigned)k << 23));
    } else {
        t = MM_G_FLOAT((unsigned)(0x7f - k) << 23);  // 2^-k
        y = x - (e + t);
        y += one;
        y = MM_G_FLOAT(MM_G_BITS(y) + ((unsigned)k << 23));
    }
    return y;
}

// e_sinhf.c
MM_G_FN float mm_g_sinhf(float x) {
    const int jx = (int)MM_G_BITS(x), ix = jx & 0x7fffffff;
    if (ix >= 0x7f800000) return x + x;
    const float h = jx < 0 ? -0.5f : 0.5f;
    if (ix < 0x41b00000) {  // |x| < 22
        if (ix < 0x31800000) return x;
        const float t = mm_g_expm1f(fabsf(x));
        if (ix < 0x3f800000) return h * (2.0f * t - t * t / (t + 1.0f));
        return h * (t + t / (t + 1.0f));
    }
    if (ix < 0x42b17180) return h * MM_G_EXPF(fabsf(x));
    if (ix <= 0x42b2d4fc) {
        const float w = MM_G_EXPF(0.5f * fabsf(x)), t = h * w;
        return t * w;
    }
    return x * 1.0e37f;
}

// e_coshf.c
MM_G_FN float mm_g_coshf(float x) {
    const int ix = (int)MM_G_BITS(x) & 0x7fffffff;
    if (ix >= 0x7f800000) return x * x;
    if (ix < 0x3eb17218) {  // |x| < 0.5 ln2
        const float t = mm_g_expm1f(fabsf(x)), w = 1.0f + t;
        if (ix < 0x24000000) return w;
        return 1.0f + (t * t) / (w + w);
    }
    if (ix < 0x41b00000) {  // |x| < 22
        const float t = MM_G_EXPF(fabsf(x));
        return 0.5f * t + 0.5f / t;
    }
    if (ix < 0x42b17180) return 0.5f * MM_G_EXPF(fabsf(x));
    if (ix <= 0x42b2d4fc) {
        const float w = MM_G_EXPF(0.5f * fabsf(x)), t = 0.5f * w;
        return t * w;
    }
    return 1.0e30f * 1.0e30f;
}

// s_log1pf.c
MM_G_FN float mm_g_log1pf(float x) {
    const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f, Lp1 = 6.6666668653e-01f, Lp2 = 4.0000000596e-01f, Lp3 = 2.8571429849e-01f,
                Lp4 = 2.2222198546e-01f, Lp5 = 1.8183572590e-01f, Lp6 = 1.5313838422e-01f, Lp7 = 1.4798198640e-01f;
    float f = 0.0f, c = 0.0f, u;
    const int hx = (int)MM_G_BITS(x), ax = hx & 0x7fffffff;
    int k = 1, hu = 0;
    if (hx < 0x3ed413d7) {  // x < 0.41422
        if (ax >= 0x3f800000) {  // x <= -1
            if (x == -1.0f) return -1.0f / 0.0f;
            return (x - x) / (x - x);
        }
        if (ax < 0x31000000) {  // |x| < 2^-29
            if (ax < 0x24800000) return x;
            return x - x * x * 0.5f;
        }
        if (hx > 0 || hx <= (int)0xbe95f61f) { k = 0; f = x; hu = 1; }  // -0.2929 < x < 0.41422
    }
    if (hx >= 0x7f800000) return x + x;
    if (k != 0) {
        if (hx < 0x5a000000) {
            u = 1.0f + x;
            hu = (int)MM_G_BITS(u);
            k = (hu >> 23) - 127;
            c = k > 0 ? 1.0f - (u - x) : x - (u - 1.0f);
            c /= u;
        } else {
            u = x;
            hu = (int)MM_G_BITS(u);
            k = (hu >> 23) - 127;
            c = 0.0f;
        }
        hu &= 0x007fffff;
        if (hu < 0x3504f7) u = MM_G_FLOAT((unsigned)hu | 0x3f800000u);
        else {
            k += 1;
            u = MM_G_FLOAT((unsigned)hu | 0x3f000000u);
            hu = (0x00800000 - hu) >> 2;
        }
        f = u - 1.0f;
    }
    const float hfsq = 0.5f * f * f;
    float R;
    if (hu == 0) {  // |f| < 2^-20
        if (f == 0.0f) {
            if (k == 0) return 0.0f;
            c += (float)k * ln2_lo;
            return (float)k * ln2_hi + c;
        }
        R = hfsq * (1.0f - 0.66666666666666666f * f);
        if (k == 0) return f - R;
        return (float)k * ln2_hi - ((R - ((float)k * ln2_lo + c)) - f);
    }
    const float s = f / (2.0f + f), z = s * s;
    R = z * (Lp1 + z * (Lp2 + z * (Lp3 + z * (Lp4 + z * (Lp5 + z * (Lp6 + z * Lp7))))));
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return (float)k * ln2_hi - ((hfsq - (s * (hfsq + R) + ((float)k * ln2_lo + c))) - f);
}
