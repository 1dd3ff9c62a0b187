// mathmap_b200 device runtime (sm_100a).  Hand-written device functions the
// NVRTC-generated filter kernels are compiled together with.  It replaces the
// C runtime the reference's generated code links against:
//   op semantics ................. reference opmacros.h:30-216, ops.lisp:112-253
//   sampling, edge behaviour ..... reference builtins/builtins.c:41-265, color.h:36-54
//   output quantisation .......... reference new_template.c.in:272-293
// Numeric contract: every IR op result is rounded to float32 (the reference
// stores each op in a float C variable); the translation unit is compiled with
// --fmad=false so + - * are never contracted; libm calls follow MM_PRECISE
// (1: evaluated in double and narrowed, like the host's double libm on float
// arguments; 0: CUDA's float libm, <= 2 ulp).  Integer work (texel addressing,
// rounding of bilinear sums to 8 bits, truncating output quantisation) is
// bit-exact in both modes.
//
// Compile-time configuration (set by backend/nvrtc_module.cpp):
//   MM_AA            1: bilinear sampler (cmdline -i), 0: nearest
//   MM_SUPERSAMPLING 1: nearest sampler omits the +0.5 (builtins.c:155-159)
//   MM_EDGE_X/Y      0 colour, 1 wrap, 2 reflect, 3 rotate
//   MM_PRECISE       see above
#pragma once

#ifndef MM_AA
#define MM_AA 0
#endif
#ifndef MM_SUPERSAMPLING
#define MM_SUPERSAMPLING 0
#endif
#ifndef MM_EDGE_X
#define MM_EDGE_X 0
#endif
#ifndef MM_EDGE_Y
#define MM_EDGE_Y 0
#endif
#ifndef MM_PRECISE
#define MM_PRECISE 0
#endif

#define MM_DEV __device__ __forceinline__
// Filters that call filters which cannot be inlined (recursion) run as device functions, one stack frame per level (about
// 100-250 bytes).  The host raises the device stack limit to 16 KB per thread for such modules; a call nested deeper than
// this returns transparent black and sets mm_call_overflow, which the host turns into an error -- instead of a stack
// overrun that would take the CUDA context with it.  (The deepest range an example declares is IFS Functional's 1-32.)
// MM_MAX_CALL_DEPTH: mm_types.h (the host names it in its error message)

// Pixel-grid launch geometry: 1-D blocks of 256 threads, each block renders a
// 32 x 8 pixel tile; MM_WARP_W selects the footprint of one warp inside the tile
// (32: one 128-byte output row segment per warp store; 8: an 8 x 4 patch, which
// keeps divergent filters -- escape-time loops, data-dependent taps -- more
// coherent and gathers more local).
#ifndef MM_WARP_W
#define MM_WARP_W 32
#endif
#define MM_BLOCK_W 32
#define MM_BLOCK_H 8
// A block renders `rows` (mm_params::rows, chosen per launch) vertically adjacent tiles one after the other (row,
// row + 8, ...), so that what depends on the column only (coordinate load and scaling, constant loads) is done once
// per thread.
MM_DEV void mm_pixel_coords(int &col, int &row, int rows) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#if MM_WARP_W == 32
    col = blockIdx.x * MM_BLOCK_W + lane;
    row = blockIdx.y * (MM_BLOCK_H * rows) + warp;
#elif MM_WARP_W == 16
    col = blockIdx.x * MM_BLOCK_W + (warp & 1) * 16 + (lane & 15);
    row = blockIdx.y * (MM_BLOCK_H * rows) + (warp >> 1) * 2 + (lane >> 4);
#else
    col = blockIdx.x * MM_BLOCK_W + (warp & 3) * 8 + (lane & 7);
    row = blockIdx.y * (MM_BLOCK_H * rows) + (warp >> 2) * 4 + (lane >> 3);
#endif
}
#include "mm_types.h"
#include "mm_elliptic.h"
#include "mm_glibc_float.h"
#include "mm_dlibm.h"

template <int N> struct mm_tup { float v[N]; };

// Tree vectors (tree_vectors.c:83-146): a tuple indexed by a computed subscript.  The reference keeps a persistent
// tree so that an update shares structure; the semantics are those of a value: get clamps the index into
// [0, length), set returns a copy with one element replaced.
template <int N> MM_DEV float mm_tree_vector_nth(int index, const mm_tup<N> &tv) {
    index = index < 0 ? 0 : (index >= N ? N - 1 : index);
    return tv.v[index];
}
template <int N> MM_DEV mm_tup<N> mm_set_tree_vector_nth(int index, mm_tup<N> tv, float value) {
    index = index < 0 ? 0 : (index >= N ? N - 1 : index);
    tv.v[index] = value;
    return tv;
}

// ---------------------------------------------------------------- conversions
// float/double -> int with x86 cvttss2si semantics (out of range and NaN give
// INT_MIN), which is what the reference's C casts do on its host.
MM_DEV int mm_f2i(float f) {
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return (int)0x80000000;
    return __float2int_rz(f);
}
MM_DEV int mm_d2i(double d) {
    if (!(d > -2147483649.0 && d < 2147483648.0)) return (int)0x80000000;
    return __double2int_rz(d);
}

// ------------------------------------------------------------------- scalar ops
// MIN/MAX are the template's macros: (a<b)?a:b and (a<b)?b:a (new_template.c.in:51-56)
MM_DEV float mm_min(float a, float b) { return (a < b) ? a : b; }
MM_DEV float mm_max(float a, float b) { return (a < b) ? b : a; }
MM_DEV int mm_min(int a, int b) { return (a < b) ? a : b; }
MM_DEV int mm_max(int a, int b) { return (a < b) ? b : a; }
MM_DEV float mm_clamp01(float x) { return mm_max(0.0f, mm_min(1.0f, x)); }
MM_DEV float mm_div(float a, float b) { return __fdiv_rn(a, b); }
MM_DEV float mm_mod(float a, float b) { return fmodf(a, b); }  // fmod is exact: float result == double result
MM_DEV float mm_sqrt(float a) { return __fsqrt_rn(a); }        // == (float)sqrt((double)a)
MM_DEV float mm_hypot(float a, float b) {
    // (float)hypot((double)a,(double)b): squares are exact in double
    double s = __dadd_rn(__dmul_rn((double)a, (double)a), __dmul_rn((double)b, (double)b));
    return (float)__dsqrt_rn(s);
}
MM_DEV float mm_sqr(float a) { return __fmul_rn(a, a); }  // == (float)pow((double)a, 2.0), see cuda_emit.cpp OP_POW
MM_DEV int mm_floor(float a) { return mm_f2i(floorf(a)); }
MM_DEV int mm_ceil(float a) { return mm_f2i(ceilf(a)); }
MM_DEV float mm_abs(float a) { return fabsf(a); }
MM_DEV int mm_abs(int a) { return mm_d2i(fabs((double)a)); }

#if MM_PRECISE
#define MM_LIBM1(name, fn) MM_DEV float name(float a) { return (float)fn((double)a); }
#define MM_LIBM2(name, fn) MM_DEV float name(float a, float b) { return (float)fn((double)a, (double)b); }
#else
#define MM_LIBM1(name, fn) MM_DEV float name(float a) { return fn##f(a); }
#define MM_LIBM2(name, fn) MM_DEV float name(float a, float b) { return fn##f(a, b); }
#endif
MM_LIBM1(mm_tan, tan)
MM_LIBM2(mm_pow, pow)
MM_LIBM1(mm_sinh, sinh)
MM_LIBM1(mm_cosh, cosh)
MM_LIBM1(mm_tanh, tanh)
MM_LIBM1(mm_asinh, asinh)
MM_LIBM1(mm_acosh, acosh)
MM_LIBM1(mm_atanh, atanh)
#if MM_PRECISE
// sin, cos, asin, acos: mm_dlibm.h (coefficients in constant memory) for the arguments it covers, CUDA's double
// functions for the rest (|x| >= 2^31, NaN).  Below 2^-27 the narrowed results are x and 1.
MM_DEV void mm_sincos(float a, float &s, float &c) {
    const float ax = fabsf(a);
    if (ax < 7.450580596923828125e-9f) { s = a; c = 1.0f; return; }
    if (ax < 2147483648.0f) { mm_d_sincos_core(a, s, c); return; }
    double ds, dc;
    sincos((double)a, &ds, &dc);
    s = (float)ds;
    c = (float)dc;
}
MM_DEV float mm_sin(float a) {
    const float ax = fabsf(a);
    if (ax < 7.450580596923828125e-9f) return a;
    if (ax < 2147483648.0f) return mm_d_sin_core(a);
    return (float)sin((double)a);
}
MM_DEV float mm_cos(float a) {
    const float ax = fabsf(a);
    if (ax < 7.450580596923828125e-9f) return 1.0f;
    if (ax < 2147483648.0f) return mm_d_cos_core(a);
    return (float)cos((double)a);
}
// RN_float(exp(x)) is +inf above 88.73 and 0 below -103.98
MM_DEV float mm_exp(float a) { return a > 89.0f ? __int_as_float(0x7f800000) : (a < -104.0f ? 0.0f : mm_d_exp_core(a)); }
MM_DEV float mm_log(float a) { return (a > 0.0f && a < __int_as_float(0x7f800000)) ? mm_d_log_core(a) : (float)log((double)a); }
MM_DEV float mm_atan(float a) { return a == a ? mm_d_atan_core(a) : (float)atan((double)a); }
MM_DEV float mm_atan2(float y, float x) {
    const float ay = fabsf(y), ax = fabsf(x), inf = __int_as_float(0x7f800000);
    if (ay > 0.0f && ay < inf && ax > 0.0f && ax < inf) return mm_d_atan2_core(y, x);
    return (float)atan2((double)y, (double)x);  // zeros, infinities, NaN: the special cases of the standard
}
MM_DEV float mm_asin(float a) { return fabsf(a) <= 1.0f ? mm_d_asin_core(a) : (float)asin((double)a); }
MM_DEV float mm_acos(float a) { return fabsf(a) <= 1.0f ? mm_d_acos_core(a) : (float)acos((double)a); }
#else
MM_LIBM1(mm_sin, sin)
MM_LIBM1(mm_cos, cos)
MM_LIBM1(mm_exp, exp)
MM_LIBM1(mm_log, log)
MM_LIBM1(mm_atan, atan)
MM_LIBM2(mm_atan2, atan2)
MM_LIBM1(mm_asin, asin)
MM_LIBM1(mm_acos, acos)
MM_DEV void mm_sincos(float a, float &s, float &c) { sincosf(a, &s, &c); }
#endif
// GSL's gsl_sf_gamma / gsl_sf_beta are third-party and absent; tgamma/lgamma based (parity unpinned)
MM_DEV float mm_gamma(float a) { return ((double)a > 171.0) ? 0.0f : (float)tgamma((double)a); }
MM_DEV float mm_beta(float a, float b) { return (float)exp(lgamma((double)a) + lgamma((double)b) - lgamma((double)a + (double)b)); }

// rand(a, b): the reference draws from glib's global Mersenne twister (g_random_double_range,
// opmacros.h:127), which is not reproducible across runs or threads; PARITY UNPINNED.  Here: a
// counter-based generator seeded per pixel (PCG output function), uniform in [a, b).
MM_DEV unsigned mm_rng_seed(int a, int b, int c) {
    unsigned h = (unsigned)a * 0x9E3779B1u ^ ((unsigned)b * 0x85EBCA77u + 0x7F4A7C15u) ^ ((unsigned)c * 0xC2B2AE3Du);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}
MM_DEV float mm_rand(unsigned &state, float a, float b) {
    state = state * 747796405u + 2891336453u;
    unsigned w = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    w = (w >> 22u) ^ w;
    double u = (double)w * (1.0 / 4294967296.0);
    return (float)((double)a + ((double)b - (double)a) * u);
}
// GSL elliptic functions (opmacros.h:101-125), see mm_elliptic.h; double inside, float at the boundary
MM_DEV float mm_ell_int_k_comp(float k) { return (float)mm_ellint_kcomp((double)k); }
MM_DEV float mm_ell_int_e_comp(float k) { return (float)mm_ellint_ecomp((double)k); }
MM_DEV float mm_ell_int_f(float phi, float k) { return (float)mm_ellint_f((double)phi, (double)k); }
MM_DEV float mm_ell_int_e(float phi, float k) { return (float)mm_ellint_e((double)phi, (double)k); }
MM_DEV float mm_ell_int_p(float phi, float k, float n) { return (float)mm_ellint_p((double)phi, (double)k, (double)n); }
MM_DEV float mm_ell_int_d(float phi, float k) { return (float)mm_ellint_d((double)phi, (double)k); }
MM_DEV float mm_ell_int_rc(float x, float y) { return (float)mm_ellint_rc((double)x, (double)y); }
MM_DEV float mm_ell_int_rd(float x, float y, float z) { return (float)mm_ellint_rd((double)x, (double)y, (double)z); }
MM_DEV float mm_ell_int_rf(float x, float y, float z) { return (float)mm_ellint_rf((double)x, (double)y, (double)z); }
MM_DEV float mm_ell_int_rj(float x, float y, float z, float p) { return (float)mm_ellint_rj((double)x, (double)y, (double)z, (double)p); }
MM_DEV mm_tup<3> mm_ell_jac(float u, float m) {
    double sn, cn, dn;
    mm_elljac((double)u, (double)m, &sn, &cn, &dn);
    mm_tup<3> r;
    r.v[0] = (float)sn; r.v[1] = (float)cn; r.v[2] = (float)dn;
    return r;
}

// gsl_linalg_HH_solve (GSL, third-party, absent) solves A x = b by Householder; restated with Cramer's rule in
// double.  PARITY UNPINNED (no reference test divides by a matrix).  A singular matrix gives 0.
MM_DEV mm_tup<2> mm_solve_linear_2(mm_tup<4> m, mm_tup<2> v) {
    double a = m.v[0], b = m.v[1], c = m.v[2], d = m.v[3], det = a * d - b * c;
    mm_tup<2> r;
    if (det == 0.0) { r.v[0] = r.v[1] = 0.f; return r; }
    r.v[0] = (float)(((double)v.v[0] * d - b * (double)v.v[1]) / det);
    r.v[1] = (float)((a * (double)v.v[1] - (double)v.v[0] * c) / det);
    return r;
}
MM_DEV mm_tup<3> mm_solve_linear_3(mm_tup<9> m, mm_tup<3> v) {
    double a[9], b[3];
    for (int i = 0; i < 9; ++i) a[i] = m.v[i];
    for (int i = 0; i < 3; ++i) b[i] = v.v[i];
    double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    mm_tup<3> r;
    if (det == 0.0) { r.v[0] = r.v[1] = r.v[2] = 0.f; return r; }
    double d0 = b[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (b[1] * a[8] - a[5] * b[2]) + a[2] * (b[1] * a[7] - a[4] * b[2]);
    double d1 = a[0] * (b[1] * a[8] - a[5] * b[2]) - b[0] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * b[2] - b[1] * a[6]);
    double d2 = a[0] * (a[4] * b[2] - b[1] * a[7]) - a[1] * (a[3] * b[2] - b[1] * a[6]) + b[0] * (a[3] * a[7] - a[4] * a[6]);
    r.v[0] = (float)(d0 / det); r.v[1] = (float)(d1 / det); r.v[2] = (float)(d2 / det);
    return r;
}

// ---------------------------------------------------------------------- complex
// float _Complex as float2.  The inverse functions are evaluated in double and narrowed; the direct ones follow
// glibc's float formulas (below).
struct mm_cd { double re, im; };
MM_DEV mm_cd mm_cd_of(float2 z) { mm_cd r; r.re = z.x; r.im = z.y; return r; }
MM_DEV float2 mm_c_narrow(mm_cd z) { return make_float2((float)z.re, (float)z.im); }
MM_DEV mm_cd mm_cd_mul(mm_cd a, mm_cd b) { mm_cd r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }
MM_DEV mm_cd mm_cd_div(mm_cd a, mm_cd b) {
    double d = b.re * b.re + b.im * b.im;
    mm_cd r; r.re = (a.re * b.re + a.im * b.im) / d; r.im = (a.im * b.re - a.re * b.im) / d; return r;
}
MM_DEV mm_cd mm_cd_exp(mm_cd z) { double e = exp(z.re), s, c; sincos(z.im, &s, &c); mm_cd r; r.re = e * c; r.im = e * s; return r; }
MM_DEV mm_cd mm_cd_log(mm_cd z) { mm_cd r; r.re = 0.5 * log(z.re * z.re + z.im * z.im); r.im = atan2(z.im, z.re); return r; }
MM_DEV mm_cd mm_cd_sqrt(mm_cd z) {
    mm_cd r;
    if (z.re == 0.0 && z.im == 0.0) { r.re = 0.0; r.im = z.im; return r; }
    double m = hypot(z.re, z.im);
    if (z.re >= 0.0) { double t = sqrt(0.5 * (m + z.re)); r.re = t; r.im = z.im / (2.0 * t); }
    else { double t = sqrt(0.5 * (m - z.re)); r.re = fabs(z.im) / (2.0 * t); r.im = copysign(t, z.im); }
    return r;
}
MM_DEV mm_cd mm_cd_make(double re, double im) { mm_cd r; r.re = re; r.im = im; return r; }

// The complex elementary functions are the largest bodies of the runtime (double exp, sincos with range reduction, the log1p
// branches of clogf ...), and a filter like Droste names them at a dozen sites.  MM_COMPLEX_CALLS = 1 makes them real
// calls: compiled once per module instead of once per site (Droste: NVRTC 3.6 -> 2.0 s); the NVRTC driver sets it for
// modules with many such sites (nvrtc_module.cpp).  Results are the same bits either way.
#ifndef MM_COMPLEX_CALLS
#define MM_COMPLEX_CALLS 0
#endif
#if MM_COMPLEX_CALLS
#define MM_CFN __device__ __noinline__
#else
#define MM_CFN MM_DEV
#endif
MM_DEV float2 mm_complex(float re, float im) { return make_float2(re, im); }
MM_DEV float mm_creal(float2 z) { return z.x; }
MM_DEV float mm_cimag(float2 z) { return z.y; }
MM_DEV float2 mm_cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
MM_DEV float2 mm_csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
MM_DEV float2 mm_cneg(float2 a) { return make_float2(-a.x, -a.y); }
MM_DEV float2 mm_cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// The reference calls glibc's FLOAT complex functions (ops.lisp:196-213: csqrtf, csinf, ..., cpowf).  Those work in
// float arithmetic on top of the float libm primitives, so the functions below follow glibc's formulas
// (math/s_c*_template.c) step by step in float, with each primitive (sinf, coshf, logf, hypotf, atan2f ...) evaluated in
// double and narrowed, i.e. correctly rounded, except sinhf, coshf, atan2f and log1pf, which glibc does not round
// correctly and which are therefore glibc's own algorithms restated bit for bit (mm_glibc_float.h).  What is left to
// differ are glibc's sinf/cosf (98.7 % correctly rounded), expf and logf (99.9 %).
// Arguments beyond glibc's overflow guards (|x| > 44 or 88) and non-finite ones take the plain double formula.
#define MM_FLT_MIN 1.17549435e-38f
#define MM_FLT_EPSILON 1.19209290e-07f
MM_DEV bool mm_c_finite(float2 z) { return isfinite(z.x) && isfinite(z.y); }
MM_DEV float mm_cr_sinh(float x) { return mm_g_sinhf(x); }  // glibc's sinhf / coshf bit for bit (mm_glibc_float.h)
MM_DEV float mm_cr_cosh(float x) { return mm_g_coshf(x); }
MM_DEV void mm_cr_sincos(float x, float &s, float &c) {  // glibc: sincosf unless |x| <= FLT_MIN
    if (fabsf(x) > MM_FLT_MIN) mm_sincos(x, s, c);
    else { s = x; c = 1.0f; }
}
MM_CFN float2 mm_csqrt(float2 z) {
    // s_csqrt_template.c; the range-scaling branches (|z| > FLT_MAX/4, subnormal parts) take the double formula
    const float ax = fabsf(z.x), ay = fabsf(z.y);
    if (!mm_c_finite(z) || ax > 8.5e37f || ay > 8.5e37f || (ax < 2.0f * MM_FLT_MIN && ax != 0.0f) || (ay < 2.0f * MM_FLT_MIN && ay != 0.0f))
        return mm_c_narrow(mm_cd_sqrt(mm_cd_of(z)));
    if (z.y == 0.0f) {
        if (z.x < 0.0f) return make_float2(0.0f, copysignf(__fsqrt_rn(-z.x), z.y));
        return make_float2(fabsf(__fsqrt_rn(z.x)), copysignf(0.0f, z.y));
    }
    if (z.x == 0.0f) {
        const float r = __fsqrt_rn(__fmul_rn(0.5f, ay));
        return make_float2(r, copysignf(r, z.y));
    }
    const float d = mm_hypot(z.x, z.y);
    float r, sgn;
    if (z.x > 0.0f) {
        r = __fsqrt_rn(__fmul_rn(0.5f, __fadd_rn(d, z.x)));
        sgn = __fmul_rn(0.5f, __fdiv_rn(z.y, r));
    } else {
        sgn = __fsqrt_rn(__fmul_rn(0.5f, __fsub_rn(d, z.x)));
        r = fabsf(__fmul_rn(0.5f, __fdiv_rn(z.y, sgn)));
    }
    return make_float2(r, copysignf(sgn, z.y));
}
MM_CFN float2 mm_cexp(float2 z) {
    // glibc cexpf: expf(re) * (cosf(im), sinf(im)), each factor rounded to float first
    float e = mm_exp(z.x), s, c;
    mm_sincos(z.y, s, c);  // each equals the separately evaluated sinf / cosf
    return make_float2(e * c, e * s);
}
// logf, correctly rounded, in both math modes (the argument is positive and finite here)
MM_DEV float mm_log_cr(float a) { return (a > 0.0f && a < __int_as_float(0x7f800000)) ? mm_d_log_core(a) : (float)log((double)a); }
MM_CFN float2 mm_clog(float2 z) {
    // s_clog_template.c without its scaling of huge / subnormal arguments
    float ax = fabsf(z.x), ay = fabsf(z.y);
    if (!mm_c_finite(z) || (ax == 0.0f && ay == 0.0f) || ax > 1.7e38f || ay > 1.7e38f || (ax < MM_FLT_MIN && ay < MM_FLT_MIN))
        return mm_c_narrow(mm_cd_log(mm_cd_of(z)));
    if (ax < ay) { const float t = ax; ax = ay; ay = t; }
    float re;
    if (ax == 1.0f) re = __fmul_rn(mm_g_log1pf(__fmul_rn(ay, ay)), 0.5f);
    else if (ax > 1.0f && ax < 2.0f && ay < 1.0f) {
        float d2m1 = __fmul_rn(__fsub_rn(ax, 1.0f), __fadd_rn(ax, 1.0f));
        if (ay >= MM_FLT_EPSILON) d2m1 = __fadd_rn(d2m1, __fmul_rn(ay, ay));
        re = __fmul_rn(mm_g_log1pf(d2m1), 0.5f);
    } else if (ax < 1.0f && ax >= 0.5f && ay < MM_FLT_EPSILON / 2.0f) {
        const float d2m1 = __fmul_rn(__fsub_rn(ax, 1.0f), __fadd_rn(ax, 1.0f));
        re = __fmul_rn(mm_g_log1pf(d2m1), 0.5f);
    } else if (ax < 1.0f && ax >= 0.5f && __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)) >= 0.5f) {
        // __x2y2m1f: x^2 + y^2 - 1 evaluated exactly, rounded once (exact in double for float inputs)
        const double d = __dsub_rn(__dadd_rn(__dmul_rn((double)ax, (double)ax), __dmul_rn((double)ay, (double)ay)), 1.0);
        re = __fmul_rn(mm_g_log1pf((float)d), 0.5f);
    } else
        re = mm_log_cr(mm_hypot(ax, ay));
    return make_float2(re, mm_g_atan2f(z.y, z.x));
}
MM_DEV float mm_carg(float2 z) { return mm_g_atan2f(z.y, z.x); }  // cargf = atan2f, ops.lisp:205
MM_CFN float2 mm_cpow(float2 a, float2 b) {
    // glibc cpowf(x, c) = cexpf(c * clogf(x)) in float complex arithmetic
    float2 l = mm_clog(a);
    return mm_cexp(mm_cmul(b, l));
}
// cosh / sinh pair of glibc's ccoshf and csinhf for a finite argument with |re| <= 88
MM_DEV float2 mm_ccosh_core(float re, float im) {
    float s, c;
    mm_cr_sincos(im, s, c);
    return make_float2(__fmul_rn(mm_cr_cosh(re), c), __fmul_rn(mm_cr_sinh(re), s));
}
MM_DEV float2 mm_csinh_core(float re, float im) {
    float s, c;
    mm_cr_sincos(im, s, c);
    return make_float2(__fmul_rn(mm_cr_sinh(re), c), __fmul_rn(mm_cr_cosh(re), s));
}
MM_CFN float2 mm_csin(float2 z) {
    if (mm_c_finite(z) && fabsf(z.y) <= 88.0f) {  // s_csin_template.c
        float s, c;
        mm_cr_sincos(fabsf(z.x), s, c);
        if (signbit(z.x)) s = -s;
        return make_float2(__fmul_rn(mm_cr_cosh(z.y), s), __fmul_rn(mm_cr_sinh(z.y), c));
    }
    double s, c; sincos((double)z.x, &s, &c);
    return make_float2((float)(s * cosh((double)z.y)), (float)(c * sinh((double)z.y)));
}
MM_CFN float2 mm_ccos(float2 z) {
    if (mm_c_finite(z) && fabsf(z.y) <= 88.0f) return mm_ccosh_core(-z.y, z.x);  // ccosf(z) = ccoshf(i z), s_ccos_template.c
    double s, c; sincos((double)z.x, &s, &c);
    return make_float2((float)(c * cosh((double)z.y)), (float)(-s * sinh((double)z.y)));
}
// tail of glibc's ctanf / ctanhf for |v| > t = 44, where cosh(v)^2 overflows: the "hyperbolic" component is +-1 and
// the other one 4 sin cos / exp(2|v|), divided in two steps to avoid intermediate underflow (s_ctan_template.c)
MM_DEV float mm_ctan_small_part(float sinu, float cosu, float v) {
    const float t = 44.0f, exp_2t = (float)exp(2.0 * 44.0);
    float r = __fmul_rn(__fmul_rn(4.0f, sinu), cosu);
    v = __fsub_rn(fabsf(v), t);
    r = __fdiv_rn(r, exp_2t);
    if (v > t) r = __fdiv_rn(r, exp_2t);
    else r = __fdiv_rn(r, (float)exp((double)__fmul_rn(2.0f, v)));
    return r;
}
MM_CFN float2 mm_ctan(float2 z) {
    if (mm_c_finite(z)) {  // s_ctan_template.c
        float sinrx, cosrx, sinhix, coshix, den;
        mm_cr_sincos(z.x, sinrx, cosrx);
        if (fabsf(z.y) > 44.0f) return make_float2(mm_ctan_small_part(sinrx, cosrx, z.y), copysignf(1.0f, z.y));
        if (fabsf(z.y) > MM_FLT_MIN) { sinhix = mm_cr_sinh(z.y); coshix = mm_cr_cosh(z.y); }
        else { sinhix = z.y; coshix = 1.0f; }
        if (fabsf(sinhix) > __fmul_rn(fabsf(cosrx), MM_FLT_EPSILON)) den = __fadd_rn(__fmul_rn(cosrx, cosrx), __fmul_rn(sinhix, sinhix));
        else den = __fmul_rn(cosrx, cosrx);
        return make_float2(__fdiv_rn(__fmul_rn(sinrx, cosrx), den), __fdiv_rn(__fmul_rn(sinhix, coshix), den));
    }
    double s, c; sincos(2.0 * (double)z.x, &s, &c);
    double d = c + cosh(2.0 * (double)z.y);
    return make_float2((float)(s / d), (float)(sinh(2.0 * (double)z.y) / d));
}
MM_CFN float2 mm_csinh(float2 z) {
    if (mm_c_finite(z) && fabsf(z.x) <= 88.0f) return mm_csinh_core(z.x, z.y);  // s_csinh_template.c
    double s, c; sincos((double)z.y, &s, &c);
    return make_float2((float)(sinh((double)z.x) * c), (float)(cosh((double)z.x) * s));
}
MM_CFN float2 mm_ccosh(float2 z) {
    if (mm_c_finite(z) && fabsf(z.x) <= 88.0f) return mm_ccosh_core(z.x, z.y);  // s_ccosh_template.c
    double s, c; sincos((double)z.y, &s, &c);
    return make_float2((float)(cosh((double)z.x) * c), (float)(sinh((double)z.x) * s));
}
MM_CFN float2 mm_ctanh(float2 z) {
    if (mm_c_finite(z)) {  // s_ctanh_template.c
        float sinix, cosix, sinhrx, coshrx, den;
        mm_cr_sincos(z.y, sinix, cosix);
        if (fabsf(z.x) > 44.0f) return make_float2(copysignf(1.0f, z.x), mm_ctan_small_part(sinix, cosix, z.x));
        if (fabsf(z.x) > MM_FLT_MIN) { sinhrx = mm_cr_sinh(z.x); coshrx = mm_cr_cosh(z.x); }
        else { sinhrx = z.x; coshrx = 1.0f; }
        if (fabsf(sinhrx) > __fmul_rn(fabsf(cosix), MM_FLT_EPSILON)) den = __fadd_rn(__fmul_rn(sinhrx, sinhrx), __fmul_rn(cosix, cosix));
        else den = __fmul_rn(cosix, cosix);
        return make_float2(__fdiv_rn(__fmul_rn(sinhrx, coshrx), den), __fdiv_rn(__fmul_rn(sinix, cosix), den));
    }
    double s, c; sincos(2.0 * (double)z.y, &s, &c);
    double d = cosh(2.0 * (double)z.x) + c;
    return make_float2((float)(sinh(2.0 * (double)z.x) / d), (float)(s / d));
}
MM_DEV mm_cd mm_cd_asinh(mm_cd z) {
    mm_cd z2 = mm_cd_mul(z, z);
    z2.re += 1.0;
    mm_cd s = mm_cd_sqrt(z2);
    return mm_cd_log(mm_cd_make(z.re + s.re, z.im + s.im));
}
MM_DEV float2 mm_casinh(float2 z) { return mm_c_narrow(mm_cd_asinh(mm_cd_of(z))); }
MM_DEV float2 mm_casin(float2 z) {  // -i * asinh(i z)
    mm_cd w = mm_cd_asinh(mm_cd_make(-(double)z.y, (double)z.x));
    return mm_c_narrow(mm_cd_make(w.im, -w.re));
}
MM_DEV float2 mm_cacos(float2 z) {
    mm_cd w = mm_cd_asinh(mm_cd_make(-(double)z.y, (double)z.x));
    return mm_c_narrow(mm_cd_make(1.5707963267948966 - w.im, w.re));
}
MM_DEV float2 mm_cacosh(float2 z) {
    mm_cd a = mm_cd_sqrt(mm_cd_make((double)z.x + 1.0, (double)z.y)), b = mm_cd_sqrt(mm_cd_make((double)z.x - 1.0, (double)z.y));
    mm_cd p = mm_cd_mul(a, b);
    return mm_c_narrow(mm_cd_log(mm_cd_make((double)z.x + p.re, (double)z.y + p.im)));
}
MM_DEV mm_cd mm_cd_atanh(mm_cd z) {
    mm_cd a = mm_cd_log(mm_cd_make(1.0 + z.re, z.im)), b = mm_cd_log(mm_cd_make(1.0 - z.re, -z.im));
    return mm_cd_make(0.5 * (a.re - b.re), 0.5 * (a.im - b.im));
}
MM_DEV float2 mm_catanh(float2 z) { return mm_c_narrow(mm_cd_atanh(mm_cd_of(z))); }
MM_DEV float2 mm_catan(float2 z) {  // -i * atanh(i z)
    mm_cd w = mm_cd_atanh(mm_cd_make(-(double)z.y, (double)z.x));
    return mm_c_narrow(mm_cd_make(w.im, -w.re));
}
// complex gamma: Luke's 7-term approximation in complex double with reflection
// for Re z < 0 (reference builtins/spec_func.c:35-63)
__device__ __noinline__ float2 mm_cgamma(float2 zf) {
    const double coeff[7] = {41.624436916439068, -51.224241022374774, 11.338755813488977, -0.747732687772388,
                             0.008782877493061,  -1.899030264e-6,     1.946335e-9};
    mm_cd z = mm_cd_of(zf), denom = mm_cd_make(1.0, 0.0);
    if (z.re < 0.0) {
        int flr = mm_d2i(-floor(z.re));
        for (int n = 0; n < flr; ++n) denom = mm_cd_mul(denom, mm_cd_make(z.re + n, z.im));
        // the reference recurses on the float-narrowed z + flr
        float2 zr = mm_c_narrow(mm_cd_make(z.re + flr, z.im));
        z = mm_cd_of(zr);
    }
    mm_cd w = mm_cd_make(z.re - 1.0, z.im), s = mm_cd_make(coeff[0], 0.0), H = mm_cd_make(1.0, 0.0);
    for (int n = 1; n < 7; n++) {
        H = mm_cd_mul(H, mm_cd_div(mm_cd_make(w.re + 1 - n, w.im), mm_cd_make(w.re + n, w.im)));
        s.re += coeff[n] * H.re;
        s.im += coeff[n] * H.im;
    }
    mm_cd e = mm_cd_exp(mm_cd_make(-w.re - 5.5, -w.im));
    mm_cd base = mm_cd_make(w.re + 5.5, w.im), ex = mm_cd_make(w.re + 0.5, w.im);
    mm_cd pw = mm_cd_exp(mm_cd_mul(ex, mm_cd_log(base)));
    mm_cd r = mm_cd_mul(mm_cd_mul(mm_cd_make(2.506628274631 * e.re, 2.506628274631 * e.im), pw), s);
    float2 g = mm_c_narrow(r);
    if (zf.x < 0.0f) g = mm_c_narrow(mm_cd_div(mm_cd_of(g), denom));
    return g;
}

// ---------------------------------------------------------------------- colours
// packing R<<24|G<<16|B<<8|A (color.h:36-43); k/255.0 narrowed to float is
// computed exactly for every k in 0..255 as fma(k, hi, k*lo) with hi+lo = 1/255
MM_DEV float mm_unit_from_byte(unsigned k) {
    float f = (float)k;
    return __fmaf_rn(f, 0.003921568859368563f, __fmul_rn(f, -2.319175823606301e-10f));
}
MM_DEV float mm_red(mm_color c) { return mm_unit_from_byte(c >> 24); }
MM_DEV float mm_green(mm_color c) { return mm_unit_from_byte((c >> 16) & 0xff); }
MM_DEV float mm_blue(mm_color c) { return mm_unit_from_byte((c >> 8) & 0xff); }
MM_DEV float mm_alpha(mm_color c) { return mm_unit_from_byte(c & 0xff); }
MM_DEV mm_tup<4> mm_tuple_from_color(mm_color c) {
    mm_tup<4> t;
    t.v[0] = mm_red(c); t.v[1] = mm_green(c); t.v[2] = mm_blue(c); t.v[3] = mm_alpha(c);
    return t;
}
// MAKE_COLOR (opmacros.h:154): float product, truncated, masked
MM_DEV mm_color mm_make_color(float r, float g, float b, float a) {
    unsigned R = (unsigned)mm_f2i(__fmul_rn(mm_clamp01(r), 255.0f)) & 0xff, G = (unsigned)mm_f2i(__fmul_rn(mm_clamp01(g), 255.0f)) & 0xff;
    unsigned B = (unsigned)mm_f2i(__fmul_rn(mm_clamp01(b), 255.0f)) & 0xff, A = (unsigned)mm_f2i(__fmul_rn(mm_clamp01(a), 255.0f)) & 0xff;
    return (R << 24) | (G << 16) | (B << 8) | A;
}
MM_DEV float mm_apply_curve(const float *curve, float p) { return curve[mm_f2i(__fmul_rn(mm_clamp01(p), (float)(MM_CURVE_POINTS - 1)))]; }
MM_DEV mm_tup<4> mm_apply_gradient(const mm_color *grad, float p) {
    return mm_tuple_from_color(grad[mm_f2i(__fmul_rn(mm_clamp01(p), (float)(MM_CURVE_POINTS - 1)))]);
}

// --------------------------------------------------------------------- sampling
// builtins/builtins.c:41-119 (C remainder semantics == CUDA's).  A NaN or out-of-range coordinate arrives here as INT_MIN
// (the x86 conversion, mm_f2i), and the reference then negates it or subtracts it from width - 1: signed overflow, undefined
// in C and in CUDA C++ alike -- gcc -O2 and NVVM both conclude "the result of -x % width lies in [0, width)", drop the
// range check that follows, and the reference reads a texel row 32 rows in front of the image.  Here the negation and the
// subtraction wrap (unsigned arithmetic), INT_MIN stays negative or becomes 0, and the range check holds.
MM_DEV int mm_wrap_neg(int v) { return (int)(0u - (unsigned)v); }
MM_DEV int mm_wrap_sub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
template <int MODE_X, int MODE_Y> MM_DEV void mm_apply_edge_behaviour(int &x, int &y, int width, int height) {
    if (MODE_X == 1) { if (x < 0) x = x % width + width; else if (x >= width) x %= width; }
    else if (MODE_X == 2) { if (x < 0) x = mm_wrap_neg(x) % width; else if (x >= width) x = (width - 1) - (x % width); }
    else if (MODE_X == 3) {
        if (x < 0) { x = mm_wrap_neg(x) % width; y = mm_wrap_sub(height - 1, y); }
        else if (x >= width) { x = (width - 1) - (x % width); y = mm_wrap_sub(height - 1, y); }
    }
    if (MODE_Y == 1) { if (y < 0) y = y % height + height; else if (y >= height) y %= height; }
    else if (MODE_Y == 2) { if (y < 0) y = mm_wrap_neg(y) % height; else if (y >= height) y = (height - 1) - (y % height); }
    else if (MODE_Y == 3) {
        if (y < 0) { x = mm_wrap_sub(width - 1, x); y = mm_wrap_neg(y) % height; }
        else if (y >= height) { x = mm_wrap_sub(width - 1, x); y = (height - 1) - (y % height); }
    }
}

// one texel as packed colour; out of image -> edge colour (x tested first), frame out of range -> white
MM_DEV mm_color mm_get_pixel(const mm_params &P, const mm_image &img, int x, int y, int frame) {
    mm_apply_edge_behaviour<MM_EDGE_X, MM_EDGE_Y>(x, y, img.w, img.h);
    if (x < 0 || x >= img.w) return P.edge_color_x;
    if (y < 0 || y >= img.h) return P.edge_color_y;
    if (frame < 0 || frame >= img.num_frames) return 0xffffffffu;
    // RGBA8 little-endian word: R in the low byte; swap to R-high packing
    unsigned w = __ldg((const unsigned *)img.data + ((size_t)y * (size_t)img.w + (size_t)x));
    return __byte_perm(w, 0, 0x0123);
}

MM_DEV mm_color mm_sample_nearest(const mm_params &P, const mm_image &img, float x, float y, int frame) {
    x = __fmul_rn(__fadd_rn(x, img.mx), img.sx);
    y = -__fmul_rn(__fsub_rn(y, img.my), img.sy);
#if !MM_SUPERSAMPLING
    x = __fadd_rn(x, 0.5f);  // x += 0.5 in double then narrowed == float add
    y = __fadd_rn(y, 0.5f);
#endif
    return mm_get_pixel(P, img, mm_f2i(floorf(x)), mm_f2i(floorf(y)), frame);
}

// bilinear, exactly as builtins.c:165-245 + color.h:48-54: float weights, per
// channel ((c1*p1 + c2*p2) + c3*p3) + c4*p4 with every op rounded, then rintf to 8 bits
MM_DEV mm_color mm_sample_bilinear(const mm_params &P, const mm_image &img, float x, float y, int frame) {
    x = __fmul_rn(__fadd_rn(x, img.mx), img.sx);
    y = -__fmul_rn(__fsub_rn(y, img.my), img.sy);
    int x1 = mm_f2i(floorf(x)), y1 = mm_f2i(floorf(y));
    int x2 = x1 + 1, y2 = y1 + 1;
    float x2f = __fsub_rn(x, (float)x1), y2f = __fsub_rn(y, (float)y1);
    float x1f = __fsub_rn(1.0f, x2f), y1f = __fsub_rn(1.0f, y2f);
    float p1 = __fmul_rn(x1f, y1f), p2 = __fmul_rn(x1f, y2f), p3 = __fmul_rn(x2f, y1f), p4 = __fmul_rn(x2f, y2f);
    mm_color c1 = mm_get_pixel(P, img, x1, y1, frame), c2 = mm_get_pixel(P, img, x1, y2, frame);
    mm_color c3 = mm_get_pixel(P, img, x2, y1, frame), c4 = mm_get_pixel(P, img, x2, y2, frame);
    mm_color result = 0;
#pragma unroll
    for (int sh = 24; sh >= 0; sh -= 8) {
        float a = (float)((c1 >> sh) & 0xff), b = (float)((c2 >> sh) & 0xff), c = (float)((c3 >> sh) & 0xff), d = (float)((c4 >> sh) & 0xff);
        float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, p1), __fmul_rn(b, p2)), __fmul_rn(c, p3)), __fmul_rn(d, p4));
        // (color_t)rintf(s) & 0xff: round half to even, then the x86 cast semantics
        unsigned q = (unsigned)mm_f2i(rintf(s)) & 0xff;
        result |= q << sh;
    }
    return result;
}

// ---- interior fast paths -------------------------------------------------------
// When every texel a sample touches lies inside the image (the common case) the
// generic path's per-texel edge handling, 64-bit addressing and chains of
// int<->float conversions (FRND, F2I, I2F on the narrow XU pipe) are replaced by:
//   * one float range test per axis (NaN and out-of-range fall to the generic path),
//   * floor and the texel index from one round-down add of 2^23 (mm_floor_biased),
//   * byte -> float through PRMT into the mantissa of 2^23 (exact) for one channel and through I2F.U8 for the other
//     three, which balances issue slots against the XU pipe (mm_texel_channel),
//   * rintf + (int) through the 1.5 * 2^23 magic constant (round half to even, exact for 0..255),
//   * byte/255 from the rounded float directly -- or, when the sample is the pixel (WORD), the rounded bytes themselves.
// Results are bit-identical to the generic path.
#define MM_MAGIC_ROUND 12582912.0f  /* 1.5 * 2^23 */
// `magic` is 0x4B000000 read from mm_params: as a literal the compiler would keep it as PRMT's immediate and spend a
// MOV per selector instead; as data it stays in one register and the selector is the immediate.
MM_DEV float mm_byte_as_float(unsigned word, unsigned magic, unsigned sel /* 0x7650 | byte index */) {
    return __fsub_rn(__int_as_float(__byte_perm(word, magic, sel)), 8388608.0f);
}
// Channel k of a texel as a float.  The PRMT + FADD pair costs two issue slots on the FP32/ALU pipes, the conversion
// instruction (I2F.U8 with a byte selector) one slot but several cycles of the narrow XU pipe: the kernels are bound by
// issue slots, so MM_XU_CHANNELS of the four channels go through the XU pipe.  Measured on Ident 8192^2 (bilinear) for
// 0 / 1 / 2 / 3 / 4 channels: 0.313 / 0.309 / 0.299 / 0.295 / 0.303 ms (with all four the XU pipe starts to limit).
#ifndef MM_XU_CHANNELS
#define MM_XU_CHANNELS 3
#endif
MM_DEV float mm_texel_channel(unsigned word, unsigned magic, int k) {
    if (k >= 4 - MM_XU_CHANNELS) return (float)((word >> (8 * k)) & 0xffu);
    return mm_byte_as_float(word, magic, 0x7650u | (unsigned)k);
}
MM_DEV float mm_unit_from_rounded(float qf) {  // qf holds an integer 0..255 exactly
    return __fmaf_rn(qf, 0.003921568859368563f, __fmul_rn(qf, -2.319175823606301e-10f));
}
// 2^23 + floor(v) for 0 <= v < 2^22, in one add rounded towards minus infinity: the low mantissa bits are floor(v) as
// an integer, and subtracting 2^23 again (exact) gives floor(v) as a float -- no conversion instructions
MM_DEV float mm_floor_biased(float v) { return __fadd_rd(v, 8388608.0f); }

// ---- packed FP32 arithmetic (Blackwell FFMA2: two floats per lane and issue slot) -----------------------------------
// The blend of a bilinear sample is 7 multiply / add operations per channel, each rounded on its own.  On two channels
// at once it takes half the issue slots, which is what bounds the sampling kernels.  ptxas contracts mul.f32x2 +
// add.f32x2 into FFMA2 whatever --fmad says (and rewrites fma(a, b, -0) / fma(s, 1, m) to get there), which would round
// once instead of twice; so the product is written fma(a, b, NZ) and the sum fma(s, ONE, m) with NZ = (-0, -0) and
// ONE = (1, 1) arriving as kernel DATA (mm_params): a*b + (-0) is RN(a*b) for every a, b including the sign of a zero
// product, s*1 + m is RN(s + m), and two FMAs cannot be fused or simplified when their constants are not known.
MM_DEV unsigned long long mm_pk(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
MM_DEV void mm_unpk(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
MM_DEV unsigned long long mm_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// ((a*p1 + b*p2) + c*p3) + d*p4 + 1.5*2^23 on a pair of channels, every product and sum rounded to float
MM_DEV unsigned long long mm_blend2(const mm_params &P, unsigned long long a, unsigned long long b, unsigned long long c, unsigned long long d,
                                    unsigned long long p1, unsigned long long p2, unsigned long long p3, unsigned long long p4) {
    const unsigned long long nz = P.pk_neg_zero, one = P.pk_one;
    unsigned long long s = mm_fma2(mm_fma2(a, p1, nz), one, mm_fma2(b, p2, nz));
    s = mm_fma2(s, one, mm_fma2(c, p3, nz));
    s = mm_fma2(s, one, mm_fma2(d, p4, nz));
    return mm_fma2(s, one, P.pk_magic_round);
}

MM_DEV void mm_texel_floats(unsigned word, unsigned magic, float (&c)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k] = mm_texel_channel(word, magic, k);
}
// One pixel from its four texels' channel floats (a = (x1,y1), b = (x1,y2), c = (x2,y1), d = (x2,y2)) and weights:
// per channel ((a*p1 + b*p2) + c*p3) + d*p4, then rintf through the 1.5 * 2^23 constant (round half to even, exact for
// 0..255: the low mantissa byte of the biased sum is (int)rintf(s) & 0xff), on two channels per instruction.
template <bool WORD>
MM_DEV void mm_blend_texels(const mm_params &P, const float (&a)[4], const float (&b)[4], const float (&c)[4], const float (&d)[4], float p1, float p2,
                            float p3, float p4, mm_tup<4> &out, unsigned &word) {
    const unsigned long long w1 = mm_pk(p1, p1), w2 = mm_pk(p2, p2), w3 = mm_pk(p3, p3), w4 = mm_pk(p4, p4);
    float biased[4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
        mm_unpk(mm_blend2(P, mm_pk(a[2 * h], a[2 * h + 1]), mm_pk(b[2 * h], b[2 * h + 1]), mm_pk(c[2 * h], c[2 * h + 1]), mm_pk(d[2 * h], d[2 * h + 1]), w1, w2,
                          w3, w4),
                biased[2 * h], biased[2 * h + 1]);
    if (WORD)
        word = __byte_perm(__byte_perm((unsigned)__float_as_int(biased[0]), (unsigned)__float_as_int(biased[1]), 0x0040),
                           __byte_perm((unsigned)__float_as_int(biased[2]), (unsigned)__float_as_int(biased[3]), 0x0040), 0x5410);
    else {
#pragma unroll
        for (int k = 0; k < 4; ++k) out.v[k] = mm_unit_from_rounded(__fsub_rn(biased[k], MM_MAGIC_ROUND));
    }
}

// WORD: instead of the four channel floats k/255, return the rounded bytes k themselves as one RGBA8 word in memory
// order (the direct-output case, see mm_orig_val_out).
template <bool WORD> MM_DEV bool mm_bilinear_interior(const mm_params &P, const mm_image &img, float x, float y, float t, mm_tup<4> &out, unsigned &word) {
#if MM_EDGE_X == 0 && MM_EDGE_Y == 0
    const float px = __fmul_rn(__fadd_rn(x, img.mx), img.sx);
    const float py = -__fmul_rn(__fsub_rn(y, img.my), img.sy);
    // x1 = floor(px) in [0, w-2] and y1 in [0, h-2]  <=>  0 <= px < w-1 and 0 <= py < h-1 (w, h < 2^22: the host
    // sets the fast_* bounds to -1 otherwise).  The frame (int)t, truncated, is valid  <=>  -1 < t < num_frames.
    if (!(px >= 0.0f && px < img.fast_wm1 && py >= 0.0f && py < img.fast_hm1 && t > -1.0f && t < img.fast_nf)) return false;
    const float bx = mm_floor_biased(px), by = mm_floor_biased(py);
    const float fx = __fsub_rn(bx, 8388608.0f), fy = __fsub_rn(by, 8388608.0f);
    const float x2f = __fsub_rn(px, fx), y2f = __fsub_rn(py, fy);
    const float x1f = __fsub_rn(1.0f, x2f), y1f = __fsub_rn(1.0f, y2f);
    const float p1 = __fmul_rn(x1f, y1f), p2 = __fmul_rn(x1f, y2f), p3 = __fmul_rn(x2f, y1f), p4 = __fmul_rn(x2f, y2f);
    // texel index y1 * w + x1, plus the 0x4B000000 of bx's exponent bits that fast_base already subtracts
    const unsigned *row0 = img.fast_base + (((unsigned)__float_as_int(by) & 0x7fffffu) * (unsigned)img.w + (unsigned)__float_as_int(bx));
    const unsigned *row1 = row0 + img.w;
    const unsigned t1 = __ldg(row0), t3 = __ldg(row0 + 1), t2 = __ldg(row1), t4 = __ldg(row1 + 1);  // pixel1..4 of builtins.c:221-224
    const unsigned magic = P.magic23;
    float a[4], b[4], c[4], d[4];  // memory byte k is R, G, B, A
    mm_texel_floats(t1, magic, a);
    mm_texel_floats(t2, magic, b);
    mm_texel_floats(t3, magic, c);
    mm_texel_floats(t4, magic, d);
    mm_blend_texels<WORD>(P, a, b, c, d, p1, p2, p3, p4, out, word);
    return true;
#else
    return false;
#endif
}

template <bool WORD> MM_DEV bool mm_nearest_interior(const mm_params &P, const mm_image &img, float x, float y, float t, mm_tup<4> &out, unsigned &word) {
#if MM_EDGE_X == 0 && MM_EDGE_Y == 0
    float px = __fmul_rn(__fadd_rn(x, img.mx), img.sx);
    float py = -__fmul_rn(__fsub_rn(y, img.my), img.sy);
#if !MM_SUPERSAMPLING
    px = __fadd_rn(px, 0.5f);
    py = __fadd_rn(py, 0.5f);
#endif
    if (!(px >= 0.0f && px < img.fast_w && py >= 0.0f && py < img.fast_h && t > -1.0f && t < img.fast_nf)) return false;
    const unsigned xb = (unsigned)__float_as_int(mm_floor_biased(px)), yb = (unsigned)__float_as_int(mm_floor_biased(py));
    const unsigned texel = __ldg(img.fast_base + ((yb & 0x7fffffu) * (unsigned)img.w + xb));
    if (WORD) { word = texel; return true; }
    const unsigned magic = P.magic23;
#pragma unroll
    for (int k = 0; k < 4; ++k) out.v[k] = mm_unit_from_rounded(mm_texel_channel(texel, magic, k));
    return true;
#else
    return false;
#endif
}

// ---- exterior fast paths ---------------------------------------------------------
// A sample whose texels ALL lie outside the image in x gets the x edge colour from every texel (x is tested first,
// builtins.c:121-130), one whose columns are all inside but whose rows are all outside the y edge colour.  For the
// bilinear sampler the blend of four equal bytes c is c again: the float weights sum to 1 within 4 * 2^-24 and the
// seven roundings add less than 2^-21 relative, so |s - c| < 3e-4 and rintf(s) = c (tests/test_cabi.py checks the bound).  Filters that recurse until a
// tap lands inside the picture (Droste) take this path for most of their taps.  NaN coordinates fail every test.
MM_DEV bool mm_bilinear_exterior(const mm_params &P, const mm_image &img, float x, float y, mm_color &edge) {
#if MM_EDGE_X == 0 && MM_EDGE_Y == 0
    if (!(img.fast_w > 0.0f)) return false;
    const float px = __fmul_rn(__fadd_rn(x, img.mx), img.sx);
    const float py = -__fmul_rn(__fsub_rn(y, img.my), img.sy);
    // beyond +-2^31 the reference's (int) casts saturate and its weights px - (float)x1 are no longer in [0, 1]: general path
    if (!(fabsf(px) < 2147483648.0f && fabsf(py) < 2147483648.0f)) return false;
    // x1 = floor(px) and x1 + 1 both outside [0, w)  <=>  px < -1 or px >= w
    if (px < -1.0f || px >= img.fast_w) { edge = P.edge_color_x; return true; }
    if (px >= 0.0f && px < img.fast_wm1 && (py < -1.0f || py >= img.fast_h)) { edge = P.edge_color_y; return true; }
#endif
    return false;
}
MM_DEV bool mm_nearest_exterior(const mm_params &P, const mm_image &img, float x, float y, mm_color &edge) {
#if MM_EDGE_X == 0 && MM_EDGE_Y == 0
    if (!(img.fast_w > 0.0f)) return false;
    float px = __fmul_rn(__fadd_rn(x, img.mx), img.sx);
    float py = -__fmul_rn(__fsub_rn(y, img.my), img.sy);
#if !MM_SUPERSAMPLING
    px = __fadd_rn(px, 0.5f);
    py = __fadd_rn(py, 0.5f);
#endif
    if (px < 0.0f || px >= img.fast_w) { edge = P.edge_color_x; return true; }
    if (py < 0.0f || py >= img.fast_h) { edge = P.edge_color_y; return true; }
#endif
    return false;
}

// (int)lrintf(v) as the reference's x86-64 build computes it (builtins.c:257-258): cvtss2si into a 64-bit long -- NaN and
// |v| >= 2^63 give 0x8000000000000000 -- of which the cast keeps the low 32 bits.  So a NaN coordinate reads column / row 0
// (not "out of range" as the 32-bit conversions of the drawable samplers make it), and so do 2^32, 2^33 ...
MM_DEV int mm_lrintf_to_int(float v) {
    if (fabsf(v) < 2147483648.0f) return __float2int_rn(v);  // the usual case: the 64-bit result fits its low word
    return fabsf(v) < 9223372036854775808.0f ? (int)__float2ll_rn(v) : 0;
}
// get_floatmap_pixel, builtins.c:249-265: nearest via lrintf (round half even)
MM_DEV mm_tup<4> mm_floatmap_pixel(const mm_image &img, float x, float y) {
    mm_tup<4> t;
    float fx = __fadd_rn(__fmul_rn(img.ax, x), img.bx), fy = __fadd_rn(__fmul_rn(img.ay, y), img.by);
    int ix = mm_lrintf_to_int(fx), iy = mm_lrintf_to_int(fy);
    if (ix < 0 || ix >= img.w || iy < 0 || iy >= img.h) { t.v[0] = t.v[1] = t.v[2] = t.v[3] = 0.0f; return t; }
    float4 v = __ldg((const float4 *)img.data + ((size_t)iy * (size_t)img.w + (size_t)ix));
    t.v[0] = v.x; t.v[1] = v.y; t.v[2] = v.z; t.v[3] = v.w;
    return t;
}

// defined at the end of every generated module (backend/cuda_emit.cpp)
__device__ mm_tup<4> mm_closure_dispatch(const mm_params &P, const mm_image &img, float x, float y, float t);

// ORIG_VAL (opmacros.h:199-216); closures known at compile time are inlined or called directly, the others dispatch
MM_DEV mm_tup<4> mm_orig_val_full(const mm_params &P, int image, float x, float y, float t) {
    const mm_image &img = P.images[image];
    x = __fmul_rn(x, img.xf);  // 1.0 unless a RESIZE wrapper survived to run time
    y = __fmul_rn(y, img.yf);
    if (img.kind == MM_IMAGE_FLOATMAP) return mm_floatmap_pixel(img, x, y);
    if (img.kind == MM_IMAGE_CLOSURE) return mm_closure_dispatch(P, img, x, y, t);  // img->v.closure.func, opmacros.h:208
    mm_tup<4> r;
#if MM_AA
    unsigned unused;
    if (mm_bilinear_interior<false>(P, img, x, y, t, r, unused)) return r;
    mm_color edge;
    if (mm_bilinear_exterior(P, img, x, y, edge)) return mm_tuple_from_color(edge);
    return mm_tuple_from_color(mm_sample_bilinear(P, img, x, y, mm_f2i(t)));
#else
    unsigned unused;
    if (mm_nearest_interior<false>(P, img, x, y, t, r, unused)) return r;
    mm_color edge;
    if (mm_nearest_exterior(P, img, x, y, edge)) return mm_tuple_from_color(edge);
    return mm_tuple_from_color(mm_sample_nearest(P, img, x, y, mm_f2i(t)));
#endif
}
// the nearest sampler regardless of MM_AA: render_image of a drawable (builtins.c:306)
MM_DEV mm_tup<4> mm_orig_val_nearest(const mm_params &P, const mm_image &img, float x, float y) {
    return mm_tuple_from_color(mm_sample_nearest(P, img, x, y, 0));
}

// The general ORIG_VAL as a real call: keeps everything but the interior fast path out of the pixel loop of a
// direct-output kernel (inlined, its branches cost the hot path about ten instructions of convergence bookkeeping).
__device__ __noinline__ mm_tup<4> mm_orig_val_call(const mm_params &P, int image, float x, float y, float t) { return mm_orig_val_full(P, image, x, y, t); }
// What generated code calls.  With MM_COMPLEX_CALLS (mmb_set_fast_compile) only a drawable's interior fast path is inlined at
// a sample site and the rest (floatmaps, closures, border and exterior texels, edge modes) is the one call above.
MM_DEV mm_tup<4> mm_orig_val(const mm_params &P, int image, float x, float y, float t) {
#if MM_COMPLEX_CALLS
    const mm_image &img = P.images[image];
    if (img.kind == MM_IMAGE_DRAWABLE) {
        mm_tup<4> r;
        unsigned unused;
        const float xs = __fmul_rn(x, img.xf), ys = __fmul_rn(y, img.yf);
#if MM_AA
        if (mm_bilinear_interior<false>(P, img, xs, ys, t, r, unused)) return r;
#else
        if (mm_nearest_interior<false>(P, img, xs, ys, t, r, unused)) return r;
#endif
    }
    return mm_orig_val_call(P, image, x, y, t);
#else
    return mm_orig_val_full(P, image, x, y, t);
#endif
}

// ORIG_VAL whose result is the pixel itself (cuda_emit.cpp: find_direct_output).  For RGBA8 output of a drawable's
// interior the sample's rounded bytes ARE the output bytes: the reference turns byte k into (float)(k / 255.0), clamps,
// multiplies by 255.0 in double and truncates, which gives k back for every k in 0..255 (tests/test_cabi.py checks the
// 256 cases), so the conversion to floats and back is skipped.  Everything else takes the general path.
MM_DEV mm_tup<4> mm_orig_val_out(const mm_params &P, int image, float x, float y, float t, unsigned &word, bool &have_word) {
    const mm_image &img = P.images[image];
    if (P.out_mode == 0 && img.kind == MM_IMAGE_DRAWABLE) {
        mm_tup<4> r = mm_tup<4>{};
        const float xs = __fmul_rn(x, img.xf), ys = __fmul_rn(y, img.yf);
#if MM_AA
        have_word = mm_bilinear_interior<true>(P, img, xs, ys, t, r, word);
#else
        have_word = mm_nearest_interior<true>(P, img, xs, ys, t, r, word);
#endif
        if (have_word) return r;
    }
    have_word = false;
    return mm_orig_val_call(P, image, x, y, t);
}

// ---- quad kernels: four horizontally adjacent pixels of one row per thread ---------------------------------------
// The "local access" variant of the pixel kernel (cuda_emit.cpp: quad mode), for filters whose samples are separable:
// the sample's x depends on the column only and its y on the row only (in(xy), translations, scalings, zooms, flips,
// stencils with constant offsets, everything that then only works on the colour).  The reference computes what
// depends on the column once per slice (init_slice, new_template.c.in:339-373); here a thread renders a 4 x 1 strip, so
// that per sample
//   * the row half of the work (coordinate transform, floor, weights, row addresses) is done once for four pixels,
//   * neighbouring pixels whose texel columns are consecutive (x1(p+1) == x1(p) + 1: the common case at unit scale)
//     share texels: 10 loads and byte -> float conversions of 40 channels instead of 16 and 64,
//   * the four RGBA8 results leave as one 128-bit store.
// Results are bit-identical to the one-pixel samplers: the same float operations per pixel, in the same order.
// Quad kernels spend their time waiting for texels (up to 16 loads per strip in flight, long-scoreboard stalls lead the
// profile), so they are compiled for more resident blocks than the compiler would choose: 4 with the bilinear sampler
// (64 registers; 80 and 3 blocks unconstrained), 5 with the nearest one (48 registers).
#ifndef MM_QUAD_BLOCKS
#define MM_QUAD_BLOCKS (MM_AA ? 4 : 5)
#endif
MM_DEV void mm_pixel_coords_quad(int &col, int &row, int rows) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    col = (blockIdx.x * MM_BLOCK_W + lane) * 4;
    row = blockIdx.y * (MM_BLOCK_H * rows) + warp;
}
// Returns 0xf when all four samples lie in the image's interior and were taken here, else 0 (the caller then takes the
// pixels one by one through the general sampler: border strips only).
template <bool WORD>
MM_DEV unsigned mm_bilinear_quad(const mm_params &P, const mm_image &img, const float (&x)[4], float y, float t, mm_tup<4> (&out)[4], unsigned (&word)[4]) {
#if MM_EDGE_X == 0 && MM_EDGE_Y == 0
    const float py = -__fmul_rn(__fsub_rn(y, img.my), img.sy);
    if (!(py >= 0.0f && py < img.fast_hm1 && t > -1.0f && t < img.fast_nf)) return 0u;
    float px[4];
    bool inside = true;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        px[p] = __fmul_rn(__fadd_rn(x[p], img.mx), img.sx);
        inside = inside && px[p] >= 0.0f && px[p] < img.fast_wm1;
    }
    if (!inside) return 0u;
    const float by = mm_floor_biased(py), fy = __fsub_rn(by, 8388608.0f);
    const float y2f = __fsub_rn(py, fy), y1f = __fsub_rn(1.0f, y2f);
    const unsigned *row0 = img.fast_base + ((unsigned)__float_as_int(by) & 0x7fffffu) * (unsigned)img.w;
    const unsigned *row1 = row0 + img.w;
    const unsigned magic = P.magic23;
    float bx[4], x2f[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        bx[p] = mm_floor_biased(px[p]);
        x2f[p] = __fsub_rn(px[p], __fsub_rn(bx[p], 8388608.0f));
    }
    const unsigned i0 = (unsigned)__float_as_int(bx[0]);
    // Decided per warp: where the lanes disagree (at unit scale floor(px) is the column or the one before it, as the float
    // rounding of px falls) a warp would run both variants.
    const bool consecutive = (unsigned)__float_as_int(bx[1]) == i0 + 1 && (unsigned)__float_as_int(bx[2]) == i0 + 2 && (unsigned)__float_as_int(bx[3]) == i0 + 3;
    if (__all_sync(__activemask(), consecutive)) {
        // consecutive texel columns: pixel p blends columns p and p + 1 of five
        unsigned ta[5], tb[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) { ta[j] = __ldg(row0 + i0 + j); tb[j] = __ldg(row1 + i0 + j); }
        float la[4], lb[4];
        mm_texel_floats(ta[0], magic, la);
        mm_texel_floats(tb[0], magic, lb);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float ra[4], rb[4];
            mm_texel_floats(ta[p + 1], magic, ra);
            mm_texel_floats(tb[p + 1], magic, rb);
            const float x1f = __fsub_rn(1.0f, x2f[p]);
            mm_blend_texels<WORD>(P, la, lb, ra, rb, __fmul_rn(x1f, y1f), __fmul_rn(x1f, y2f), __fmul_rn(x2f[p], y1f), __fmul_rn(x2f[p], y2f), out[p], word[p]);
#pragma unroll
            for (int k = 0; k < 4; ++k) { la[k] = ra[k]; lb[k] = rb[k]; }
        }
    } else {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const unsigned ix = (unsigned)__float_as_int(bx[p]);
            float a[4], b[4], c[4], d[4];
            mm_texel_floats(__ldg(row0 + ix), magic, a);
            mm_texel_floats(__ldg(row1 + ix), magic, b);
            mm_texel_floats(__ldg(row0 + ix + 1), magic, c);
            mm_texel_floats(__ldg(row1 + ix + 1), magic, d);
            const float x1f = __fsub_rn(1.0f, x2f[p]);
            mm_blend_texels<WORD>(P, a, b, c, d, __fmul_rn(x1f, y1f), __fmul_rn(x1f, y2f), __fmul_rn(x2f[p], y1f), __fmul_rn(x2f[p], y2f), out[p], word[p]);
        }
    }
    return 0xfu;
#else
    return 0u;
#endif
}
template <bool WORD>
MM_DEV unsigned mm_nearest_quad(const mm_params &P, const mm_image &img, const float (&x)[4], float y, float t, mm_tup<4> (&out)[4], unsigned (&word)[4]) {
#if MM_EDGE_X == 0 && MM_EDGE_Y == 0
    float py = -__fmul_rn(__fsub_rn(y, img.my), img.sy);
#if !MM_SUPERSAMPLING
    py = __fadd_rn(py, 0.5f);
#endif
    if (!(py >= 0.0f && py < img.fast_h && t > -1.0f && t < img.fast_nf)) return 0u;
    float px[4];
    bool inside = true;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        px[p] = __fmul_rn(__fadd_rn(x[p], img.mx), img.sx);
#if !MM_SUPERSAMPLING
        px[p] = __fadd_rn(px[p], 0.5f);
#endif
        inside = inside && px[p] >= 0.0f && px[p] < img.fast_w;
    }
    if (!inside) return 0u;
    const unsigned *row = img.fast_base + ((unsigned)__float_as_int(mm_floor_biased(py)) & 0x7fffffu) * (unsigned)img.w;
    // The block's next tile is eight rows down: at unit scale its texel row is eight rows down too, requested into L2 now
    // (a guess that costs one instruction; inside the image by the test).  Measured on Invert 8192^2: 0.145 -> 0.137 ms;
    // the bilinear sampler, which waits on L1 / L2 hits rather than on DRAM, gained nothing from the same hint.
    if (py + 8.0f < img.fast_h) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 8u * (unsigned)img.w + (unsigned)__float_as_int(mm_floor_biased(px[0]))));
    const unsigned magic = P.magic23;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const unsigned texel = __ldg(row + (unsigned)__float_as_int(mm_floor_biased(px[p])));
        if (WORD) word[p] = texel;
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) out[p].v[k] = mm_unit_from_rounded(mm_texel_channel(texel, magic, k));
        }
    }
    return 0xfu;
#else
    return 0u;
#endif
}
// ORIG_VAL of four pixels of a row whose sample rows coincide (y, t shared)
MM_DEV void mm_orig_val_quad(const mm_params &P, int image, const float (&x)[4], float y, float t, mm_tup<4> (&r)[4]) {
    const mm_image &img = P.images[image];
    unsigned done = 0u;
    if (img.kind == MM_IMAGE_DRAWABLE) {
        float xs[4];
        unsigned unused[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) xs[p] = __fmul_rn(x[p], img.xf);
#if MM_AA
        done = mm_bilinear_quad<false>(P, img, xs, __fmul_rn(y, img.yf), t, r, unused);
#else
        done = mm_nearest_quad<false>(P, img, xs, __fmul_rn(y, img.yf), t, r, unused);
#endif
    }
    if (done != 0xfu) {  // unrolled: a runtime index would put the arrays into local memory
#pragma unroll
        for (int p = 0; p < 4; ++p) r[p] = mm_orig_val_call(P, image, x[p], y, t);
    }
}
// ... whose results are the pixels themselves (see mm_orig_val_out): `have` gets a bit per pixel whose RGBA8 word is in `word`
MM_DEV void mm_orig_val_out_quad(const mm_params &P, int image, const float (&x)[4], float y, float t, mm_tup<4> (&r)[4], unsigned (&word)[4], unsigned &have) {
    const mm_image &img = P.images[image];
    have = 0u;
#pragma unroll
    for (int p = 0; p < 4; ++p) r[p] = mm_tup<4>{};
    if (P.out_mode == 0 && img.kind == MM_IMAGE_DRAWABLE) {
        float xs[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) xs[p] = __fmul_rn(x[p], img.xf);
#if MM_AA
        have = mm_bilinear_quad<true>(P, img, xs, __fmul_rn(y, img.yf), t, r, word);
#else
        have = mm_nearest_quad<true>(P, img, xs, __fmul_rn(y, img.yf), t, r, word);
#endif
    }
    if (have != 0xfu) {
#pragma unroll
        for (int p = 0; p < 4; ++p) r[p] = mm_orig_val_call(P, image, x[p], y, t);
    }
}

// local (compact) row of this launch -> absolute image row
MM_DEV int mm_actual_row(const mm_params &P, int row) {
    if (P.row_interleave <= 1) return row + P.first_row;
    return P.first_row + ((row >> 3) * P.row_interleave + P.row_phase) * 8 + (row & 7);
}

// ----------------------------------------------------------------------- output
// new_template.c.in:272-293: clamp, times 255.0 in double, truncate
MM_DEV unsigned mm_quant(float v) { return __float2uint_rz(__fmul_rz(mm_clamp01(v), 255.0f)); }
// The same byte in the low mantissa bits of a float: fma(v, 255, 2^23) rounded towards zero is 2^23 + floor(v * 255) with the
// product exact inside the fma, i.e. the truncated double product of the reference -- one FP32 instruction, no conversion
MM_DEV unsigned mm_quant_bits(float v) { return (unsigned)__float_as_int(__fmaf_rz(mm_clamp01(v), 255.0f, 8388608.0f)); }

// rowp: the output row of this pixel (P.out + compact row * P.out_stride)
MM_DEV void mm_store_word(char *rowp, int col, unsigned word) { ((unsigned *)rowp)[col] = word; }  // out_mode 0 only
MM_DEV void mm_store_pixel(const mm_params &P, char *rowp, int col, const mm_tup<4> &t) {
    if (P.out_mode == 0) {  // RGBA8
        const unsigned rg = __byte_perm(mm_quant_bits(t.v[0]), mm_quant_bits(t.v[1]), 0x0040);
        const unsigned ba = __byte_perm(mm_quant_bits(t.v[2]), mm_quant_bits(t.v[3]), 0x0040);
        ((unsigned *)rowp)[col] = __byte_perm(rg, ba, 0x5410);
        return;
    }
    if (P.out_mode == 1) {
        ((float4 *)rowp)[col] = make_float4(t.v[0], t.v[1], t.v[2], t.v[3]);
        return;
    }
    unsigned char *p = (unsigned char *)rowp + (size_t)col * P.bpp;
    if (P.bpp == 3) {
        p[0] = mm_quant(t.v[0]); p[1] = mm_quant(t.v[1]); p[2] = mm_quant(t.v[2]);
    } else {
        double l = ((double)mm_clamp01(t.v[0]) * 0.299 + (double)mm_clamp01(t.v[1]) * 0.587 + (double)mm_clamp01(t.v[2]) * 0.114) * 255.0;
        p[0] = (unsigned char)mm_d2i(l);
        if (P.bpp == 2) p[1] = mm_quant(t.v[3]);
    }
}

// the RGBA8 word of a pixel (the out_mode 0 branch of mm_store_pixel)
MM_DEV unsigned mm_pack_pixel(const mm_tup<4> &t) {
    const unsigned rg = __byte_perm(mm_quant_bits(t.v[0]), mm_quant_bits(t.v[1]), 0x0040);
    const unsigned ba = __byte_perm(mm_quant_bits(t.v[2]), mm_quant_bits(t.v[3]), 0x0040);
    return __byte_perm(rg, ba, 0x5410);
}
// np of the four pixels from column col on lie inside the region; `have`: pixels whose word is already in `word`
MM_DEV void mm_store_quad(const mm_params &P, char *rowp, int col, int np, const mm_tup<4> (&t)[4], const unsigned (&word)[4], unsigned have) {
    if (P.out_mode == 0) {
        unsigned w[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) w[p] = ((have >> p) & 1u) ? word[p] : mm_pack_pixel(t[p]);
        if (np == 4 && P.vec_store) { *(uint4 *)(rowp + (size_t)col * 4) = make_uint4(w[0], w[1], w[2], w[3]); return; }
#pragma unroll
        for (int p = 0; p < 4; ++p)
            if (p < np) ((unsigned *)rowp)[col + p] = w[p];
        return;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
        if (p < np) mm_store_pixel(P, rowp, col + p, t[p]);
}

#include "mm_noise.cuh"
