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
    long bad[4] = {0, 0, 0, 0}, worst[4] = {0, 0, 0, 0}, n_trig = 0, n_inv = 0;
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
    }
    printf("arguments trig %ld inverse %ld mismatches sin %ld cos %ld acos %ld asin %ld worst ulps %ld %ld %ld %ld\n", n_trig, n_inv, bad[0], bad[1], bad[2],
           bad[3], worst[0], worst[1], worst[2], worst[3]);
    return 0;
}
