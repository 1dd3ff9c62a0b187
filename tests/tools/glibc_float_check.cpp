// Test infrastructure: compares the restated fdlibm float routines (mathmap_b200/csrc/runtime/mm_glibc_float.h, the
// header the device runtime includes) with the host's libm, bit for bit.
//   g++ -O2 -ffp-contract=off -o check glibc_float_check.cpp -lm && ./check [STRIDE]
// STRIDE 1 walks all 2^32 float bit patterns (a few minutes); the test suite uses a prime stride.
#include <cstdio>
#include <cstdlib>

#include "../../mathmap_b200/csrc/runtime/mm_glibc_float.h"

static bool same(float a, float b) { return mm_g_bits_host(a) == mm_g_bits_host(b) || (a != a && b != b); }

int main(int argc, char **argv) {
    const unsigned long long stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1021;
    long bad[6] = {0, 0, 0, 0, 0, 0}, total = 0;
    for (unsigned long long u = 0; u < 0x100000000ull; u += stride) {
        const float x = mm_g_float_host((unsigned)u);
        ++total;
        bad[0] += !same(atanf(x), mm_g_atanf(x));
        bad[1] += !same(expm1f(x), mm_g_expm1f(x));
        bad[2] += !same(sinhf(x), mm_g_sinhf(x));
        bad[3] += !same(coshf(x), mm_g_coshf(x));
        bad[4] += !same(log1pf(x), mm_g_log1pf(x));
    }
    // atan2f: pairs of arbitrary bit patterns and of moderate values
    unsigned long long s = 88172645463325252ull;
    auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    long pairs = 0;
    for (unsigned long long i = 0; i < 0x100000000ull / stride * 4; ++i) {
        const unsigned long long r = next();
        float y, x;
        if (i & 1) { y = mm_g_float_host((unsigned)r); x = mm_g_float_host((unsigned)(r >> 32)); }
        else { y = ((int)(r % 2000001) - 1000000) / 65536.0f; x = ((int)((r >> 32) % 2000001) - 1000000) / 65536.0f; }
        ++pairs;
        bad[5] += !same(atan2f(y, x), mm_g_atan2f(y, x));
    }
    printf("arguments %ld pairs %ld mismatches atanf %ld expm1f %ld sinhf %ld coshf %ld log1pf %ld atan2f %ld\n", total, pairs, bad[0], bad[1], bad[2],
           bad[3], bad[4], bad[5]);
    return (bad[0] | bad[1] | bad[2] | bad[3] | bad[4] | bad[5]) != 0;
}
