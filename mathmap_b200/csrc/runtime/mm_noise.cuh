// Coherent noise device functions: the published libnoise 1.0.0 algorithms
// (third-party; the reference vendors libnoisesrc-1.0.0.zip, patches in the
// 7th-order "QUALITY_BESTEST" blend via libnoise-bestest.diff, and calls
// Perlin / Billow / RidgedMulti / Voronoi modules from builtins/libnoise.cpp:31-88
// with frequency 1, seed 0).  Integer lattice hash + double arithmetic; with
// --fmad=false the results are bit-identical to the host's.
//
// The 256-entry gradient table lives in constant memory; lookups within a warp
// hit at most a few distinct entries per corner for smooth coordinates, and the
// control flow is warp-uniform (fixed octave counts, fixed 5x5x5 Voronoi search).
#pragma once

__constant__ double mm_noise_vectors[256 * 3] = {
#include "mm_noise_table.inc"
};

#define MM_X_NOISE_GEN 1619u
#define MM_Y_NOISE_GEN 31337u
#define MM_Z_NOISE_GEN 6971u
#define MM_SEED_NOISE_GEN 1013u

MM_DEV double mm_scurve7(double a) {
    double a2 = a * a, a4 = a2 * a2, a5 = a4 * a, a6 = a4 * a2, a7 = a5 * a2;
    return -20.0 * a7 + 70.0 * a6 - 84.0 * a5 + 35.0 * a4;
}
MM_DEV double mm_lerp(double n0, double n1, double a) { return ((1.0 - a) * n0) + (a * n1); }

MM_DEV double mm_gradient_noise(double fx, double fy, double fz, int ix, int iy, int iz, int seed) {
    int vi = (int)(MM_X_NOISE_GEN * (unsigned)ix + MM_Y_NOISE_GEN * (unsigned)iy + MM_Z_NOISE_GEN * (unsigned)iz + MM_SEED_NOISE_GEN * (unsigned)seed);
    vi ^= (vi >> 8);
    vi &= 0xff;
    double xg = mm_noise_vectors[vi * 3], yg = mm_noise_vectors[vi * 3 + 1], zg = mm_noise_vectors[vi * 3 + 2];
    double xp = fx - (double)ix, yp = fy - (double)iy, zp = fz - (double)iz;
    return ((xg * xp) + (yg * yp) + (zg * zp)) * 2.12;
}

MM_DEV int mm_noise_floor(double v) { return v > 0.0 ? mm_d2i(v) : mm_d2i(v) - 1; }

__device__ __noinline__ double mm_gradient_coherent_noise(double x, double y, double z, int seed) {
    int x0 = mm_noise_floor(x), x1 = x0 + 1, y0 = mm_noise_floor(y), y1 = y0 + 1, z0 = mm_noise_floor(z), z1 = z0 + 1;
    double xs = mm_scurve7(x - (double)x0), ys = mm_scurve7(y - (double)y0), zs = mm_scurve7(z - (double)z0);
    double n0, n1, ix0, ix1, iy0, iy1;
    n0 = mm_gradient_noise(x, y, z, x0, y0, z0, seed);
    n1 = mm_gradient_noise(x, y, z, x1, y0, z0, seed);
    ix0 = mm_lerp(n0, n1, xs);
    n0 = mm_gradient_noise(x, y, z, x0, y1, z0, seed);
    n1 = mm_gradient_noise(x, y, z, x1, y1, z0, seed);
    ix1 = mm_lerp(n0, n1, xs);
    iy0 = mm_lerp(ix0, ix1, ys);
    n0 = mm_gradient_noise(x, y, z, x0, y0, z1, seed);
    n1 = mm_gradient_noise(x, y, z, x1, y0, z1, seed);
    ix0 = mm_lerp(n0, n1, xs);
    n0 = mm_gradient_noise(x, y, z, x0, y1, z1, seed);
    n1 = mm_gradient_noise(x, y, z, x1, y1, z1, seed);
    ix1 = mm_lerp(n0, n1, xs);
    iy1 = mm_lerp(ix0, ix1, ys);
    return mm_lerp(iy0, iy1, zs);
}

MM_DEV double mm_value_noise(int x, int y, int z, int seed) {
    unsigned n = (MM_X_NOISE_GEN * (unsigned)x + MM_Y_NOISE_GEN * (unsigned)y + MM_Z_NOISE_GEN * (unsigned)z + MM_SEED_NOISE_GEN * (unsigned)seed) & 0x7fffffffu;
    n = (n >> 13) ^ n;
    int v = (int)((n * (n * n * 60493u + 19990303u) + 1376312589u) & 0x7fffffffu);
    return 1.0 - ((double)v / 1073741824.0);
}

MM_DEV double mm_make_int32_range(double n) {
    if (n >= 1073741824.0) return (2.0 * fmod(n, 1073741824.0)) - 1073741824.0;
    else if (n <= -1073741824.0) return (2.0 * fmod(n, 1073741824.0)) + 1073741824.0;
    return n;
}

__device__ __noinline__ float mm_libnoise_perlin(int octaves, float persistence_f, float lacunarity_f, float xf, float yf, float zf) {
    double x = xf, y = yf, z = zf, lacunarity = lacunarity_f, persistence = persistence_f, value = 0.0, cur = 1.0;
    for (int o = 0; o < octaves; o++) {
        double signal = mm_gradient_coherent_noise(mm_make_int32_range(x), mm_make_int32_range(y), mm_make_int32_range(z), o);
        value += signal * cur;
        x *= lacunarity; y *= lacunarity; z *= lacunarity;
        cur *= persistence;
    }
    return (float)value;
}

__device__ __noinline__ float mm_libnoise_billow(int octaves, float persistence_f, float lacunarity_f, float xf, float yf, float zf) {
    double x = xf, y = yf, z = zf, lacunarity = lacunarity_f, persistence = persistence_f, value = 0.0, cur = 1.0;
    for (int o = 0; o < octaves; o++) {
        double signal = mm_gradient_coherent_noise(mm_make_int32_range(x), mm_make_int32_range(y), mm_make_int32_range(z), o);
        signal = 2.0 * fabs(signal) - 1.0;
        value += signal * cur;
        x *= lacunarity; y *= lacunarity; z *= lacunarity;
        cur *= persistence;
    }
    value += 0.5;
    return (float)value;
}

__device__ __noinline__ float mm_libnoise_ridged_multi(int octaves, float lacunarity_f, float xf, float yf, float zf) {
    double x = xf, y = yf, z = zf, lacunarity = lacunarity_f;
    double value = 0.0, weight = 1.0, frequency = 1.0;
    if (octaves > 30) octaves = 30;
    for (int o = 0; o < octaves; o++) {
        // spectral weight pow(frequency, -1.0) with frequency = lacunarity^o built by repeated multiplication
        double sw = pow(frequency, -1.0);
        double signal = mm_gradient_coherent_noise(mm_make_int32_range(x), mm_make_int32_range(y), mm_make_int32_range(z), o & 0x7fffffff);
        signal = fabs(signal);
        signal = 1.0 - signal;
        signal *= signal;
        signal *= weight;
        weight = signal * 2.0;
        if (weight > 1.0) weight = 1.0;
        if (weight < 0.0) weight = 0.0;
        value += (signal * sw);
        x *= lacunarity; y *= lacunarity; z *= lacunarity;
        frequency *= lacunarity;
    }
    return (float)((value * 1.25) - 1.0);
}

__device__ __noinline__ float mm_libnoise_voronoi(float displacement_f, float xf, float yf, float zf) {
    double x = xf, y = yf, z = zf, displacement = displacement_f;
    int xInt = mm_noise_floor(x), yInt = mm_noise_floor(y), zInt = mm_noise_floor(z);
    double minDist = 2147483647.0, xc = 0, yc = 0, zc = 0;
    for (int zCur = zInt - 2; zCur <= zInt + 2; zCur++)
        for (int yCur = yInt - 2; yCur <= yInt + 2; yCur++)
            for (int xCur = xInt - 2; xCur <= xInt + 2; xCur++) {
                double xPos = xCur + mm_value_noise(xCur, yCur, zCur, 0);
                double yPos = yCur + mm_value_noise(xCur, yCur, zCur, 1);
                double zPos = zCur + mm_value_noise(xCur, yCur, zCur, 2);
                double xDist = xPos - x, yDist = yPos - y, zDist = zPos - z;
                double dist = xDist * xDist + yDist * yDist + zDist * zDist;
                if (dist < minDist) { minDist = dist; xc = xPos; yc = yPos; zc = zPos; }
            }
    return (float)(0.0 + (displacement * mm_value_noise(mm_d2i(floor(xc)), mm_d2i(floor(yc)), mm_d2i(floor(zc)), 0)));
}
