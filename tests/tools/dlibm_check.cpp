// Test infrastructure: compares mathmap_b200/csrc/runtime/mm_dlibm.h (the header the device runtime includes, compiled
// here for the host) with the host's double libm narrowed to float, which is what the reference's ops compute.
//   g++ -O2 -ffp-contract=off -o check dlibm_check.cpp -lm && ./check [STRIDE]
// Prints, per function, the number of float arguments whose narrowed result differs and the largest difference in
// float ulps.  STRIDE 1 walks all 2^32 float bit patterns.
#include <cstdio>
#include <cstdlib>

#include "../../mathmap_b200/csrc/runtime/mm_dlibm.h"

static unsigned bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static float from_bits(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static long ulps(float a, float b) {
    if (a != a && b != b) return 0;
    if (a != a || b != b) return 1 << 30;
    auto key = [](float f) { int i = (int)bits(f); return (long)(i < 0 ? (int)0x80000000 - i : i); };
    long d = key(a) - key(b);
    if (d == 0 && bits(a) != bits(b)) return 1;  // +0 vs -0
    return d < 0 ? -d : d;
}

int main(int argc, char **argv) {
    const unsigned long long stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1021;
    long bad[8] = {0, 0, 0, 0, 0, 0, 0, 0}, worst[8] = {0, 0, 0, 0, 0, 0, 0, 0}, n_trig = 0, n_inv = 0, n_exp = 0, n_log = 0, n_atan = 0, n_atan2 = 0;
    for (unsigned long long u = 0; u < 0x100000000ull; u += stride) {
        const float x = from_bits((unsigned)u);
        const float ax = fabsf(x);
        if (ax >= 0x1p-27f && ax < 0x1p31f) {
            float s, c;
            mm_d_sincos_core(x, s, c);
            long ds = ulps(s, (float)sin((double)x)), dc = ulps(c, (float)cos((double)x));
            // the single-function variants must give the very same floats
            if (bits(mm_d_sin_core(x)) != bits(s)) ds = 1 << 30;
            if (bits(mm_d_cos_core(x)) != bits(c)) dc = 1 << 30;
            bad[0] += ds != 0; bad[1] += dc != 0;
            if (ds > worst[0]) worst[0] = ds;
            if (dc > worst[1]) worst[1] = dc;
            ++n_trig;
        }
        if (ax <= 1.0f) {
            const long da = ulps(mm_d_acos_core(x), (float)acos((double)x)), di = ulps(mm_d_asin_core(x), (float)asin((double)x));
            bad[2] += da != 0; bad[3] += di != 0;
            if (da > worst[2]) worst[2] = da;
            if (di > worst[3]) worst[3] = di;
            ++n_inv;
        }
        if (x >= -104.0f && x <= 89.0f) {
            const long de = ulps(mm_d_exp_core(x), (float)exp((double)x));
            bad[4] += de != 0;
            if (de > worst[4]) worst[4] = de;
            ++n_exp;
        }
        if (x == x) {
            const long dt = ulps(mm_d_atan_core(x), (float)atan((double)x));
            bad[6] += dt != 0;
            if (dt > worst[6]) worst[6] = dt;
            ++n_atan;
        }
        if (x > 0.0f && x < INFINITY) {
            const long dl = ulps(mm_d_log_core(x), (float)log((double)x));
            bad[5] += dl != 0;
            if (dl > worst[5]) worst[5] = dl;
            ++n_log;
        }
    }
    // atan2: pairs of arbitrary finite non-zero floats and of moderate values
    unsigned long long rs = 88172645463325252ull;
    auto next = [&]() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; };
    for (unsigned long long i = 0; i < 0x100000000ull / stride * 2; ++i) {
        const unsigned long long r = next();
        float y, x;
        if (i & 1) { y = from_bits((unsigned)r); x = from_bits((unsigned)(r >> 32)); }
        else { y = ((int)(r % 2000001) - 1000000) / 65536.0f; x = ((int)((r >> 32) % 2000001) - 1000000) / 65536.0f; }
        if (!(fabsf(x) < INFINITY) || !(fabsf(y) < INFINITY) || x == 0.0f || y == 0.0f) continue;
        const long d = ulps(mm_d_atan2_core(y, x), (float)atan2((double)y, (double)x));
        bad[7] += d != 0;
        if (d > worst[7]) worst[7] = d;
        ++n_atan2;
    }
    printf("arguments trig %ld inverse %ld exp %ld log %ld atan %ld atan2 %ld mismatches sin %ld cos %ld acos %ld asin %ld exp %ld log %ld atan %ld atan2 %ld worst ulps %ld %ld %ld %ld %ld %ld %ld %ld\n",
           n_trig, n_inv, n_exp, n_log, n_atan, n_atan2, bad[0], bad[1], bad[2], bad[3], bad[4], bad[5], bad[6], bad[7], worst[0], worst[1], worst[2], worst[3],
           worst[4], worst[5], worst[6], worst[7]);
    return 0;
}
