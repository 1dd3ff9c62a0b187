/* ORACLE (test infrastructure): coherent noise, restating the published
 * algorithms of libnoise 1.0.0 (third-party, vendored by the reference as
 * libnoisesrc-1.0.0.zip and patched by libnoise-bestest.diff) as the reference
 * calls them from builtins/libnoise.cpp:31-88:
 *   gradient / value lattice noise, integer hash .... noise/src/noisegen.cpp
 *   7th-order blend curve "QUALITY_BESTEST" ......... libnoise-bestest.diff (SCurve7)
 *   Perlin / Billow / RidgedMulti / Voronoi sums ..... noise/src/module/{perlin,billow,ridgedmulti,voronoi}.cpp
 * Module defaults used by the wrappers: frequency 1, seed 0, ridged offset 1 /
 * gain 2 / spectral exponent 1, Voronoi distance disabled.
 * All arithmetic is double on float arguments, narrowed to float on return.
 */
#include "mmo_runtime.h"

static const double k_vectors[256 * 3] = {
#include "noise_table.inc"
};

enum { X_NOISE_GEN = 1619, Y_NOISE_GEN = 31337, Z_NOISE_GEN = 6971, SEED_NOISE_GEN = 1013, SHIFT_NOISE_GEN = 8 };

static double scurve7(double a) {
    double a2 = a * a;
    double a4 = a2 * a2;
    double a5 = a4 * a;
    double a6 = a4 * a2;
    double a7 = a5 * a2;
    return -20.0 * a7 + 70.0 * a6 - 84.0 * a5 + 35.0 * a4;
}
static double lerp(double n0, double n1, double a) { return ((1.0 - a) * n0) + (a * n1); }

static double gradient_noise(double fx, double fy, double fz, int ix, int iy, int iz, int seed) {
    /* unsigned arithmetic == the reference's wrapping int arithmetic */
    unsigned int vi = (unsigned)X_NOISE_GEN * (unsigned)ix + (unsigned)Y_NOISE_GEN * (unsigned)iy + (unsigned)Z_NOISE_GEN * (unsigned)iz +
                      (unsigned)SEED_NOISE_GEN * (unsigned)seed;
    int vectorIndex = (int)vi;
    vectorIndex ^= (vectorIndex >> SHIFT_NOISE_GEN);
    vectorIndex &= 0xff;
    {
        double xg = k_vectors[vectorIndex * 3], yg = k_vectors[vectorIndex * 3 + 1], zg = k_vectors[vectorIndex * 3 + 2];
        double xp = (fx - (double)ix), yp = (fy - (double)iy), zp = (fz - (double)iz);
        return ((xg * xp) + (yg * yp) + (zg * zp)) * 2.12;
    }
}

static double gradient_coherent_noise(double x, double y, double z, int seed) {
    int x0 = (x > 0.0 ? (int)x : (int)x - 1), x1 = x0 + 1;
    int y0 = (y > 0.0 ? (int)y : (int)y - 1), y1 = y0 + 1;
    int z0 = (z > 0.0 ? (int)z : (int)z - 1), z1 = z0 + 1;
    double xs = scurve7(x - (double)x0), ys = scurve7(y - (double)y0), zs = scurve7(z - (double)z0);
    double n0, n1, ix0, ix1, iy0, iy1;
    n0 = gradient_noise(x, y, z, x0, y0, z0, seed);
    n1 = gradient_noise(x, y, z, x1, y0, z0, seed);
    ix0 = lerp(n0, n1, xs);
    n0 = gradient_noise(x, y, z, x0, y1, z0, seed);
    n1 = gradient_noise(x, y, z, x1, y1, z0, seed);
    ix1 = lerp(n0, n1, xs);
    iy0 = lerp(ix0, ix1, ys);
    n0 = gradient_noise(x, y, z, x0, y0, z1, seed);
    n1 = gradient_noise(x, y, z, x1, y0, z1, seed);
    ix0 = lerp(n0, n1, xs);
    n0 = gradient_noise(x, y, z, x0, y1, z1, seed);
    n1 = gradient_noise(x, y, z, x1, y1, z1, seed);
    ix1 = lerp(n0, n1, xs);
    iy1 = lerp(ix0, ix1, ys);
    return lerp(iy0, iy1, zs);
}

static int int_value_noise(int x, int y, int z, int seed) {
    unsigned int n = ((unsigned)X_NOISE_GEN * (unsigned)x + (unsigned)Y_NOISE_GEN * (unsigned)y + (unsigned)Z_NOISE_GEN * (unsigned)z +
                      (unsigned)SEED_NOISE_GEN * (unsigned)seed) & 0x7fffffffu;
    n = (n >> 13) ^ n;
    return (int)((n * (n * n * 60493u + 19990303u) + 1376312589u) & 0x7fffffffu);
}
static double value_noise(int x, int y, int z, int seed) { return 1.0 - ((double)int_value_noise(x, y, z, seed) / 1073741824.0); }

static double make_int32_range(double n) {
    if (n >= 1073741824.0) return (2.0 * fmod(n, 1073741824.0)) - 1073741824.0;
    else if (n <= -1073741824.0) return (2.0 * fmod(n, 1073741824.0)) + 1073741824.0;
    else return n;
}

float libnoise_perlin(int octaves, float persistence_f, float lacunarity_f, float xf, float yf, float zf) {
    double x = xf, y = yf, z = zf, lacunarity = lacunarity_f, persistence = persistence_f;
    double value = 0.0, signal, cur = 1.0;
    int o;
    for (o = 0; o < octaves; o++) {
        double nx = make_int32_range(x), ny = make_int32_range(y), nz = make_int32_range(z);
        signal = gradient_coherent_noise(nx, ny, nz, o);
        value += signal * cur;
        x *= lacunarity;
        y *= lacunarity;
        z *= lacunarity;
        cur *= persistence;
    }
    return value;
}

float libnoise_billow(int octaves, float persistence_f, float lacunarity_f, float xf, float yf, float zf) {
    double x = xf, y = yf, z = zf, lacunarity = lacunarity_f, persistence = persistence_f;
    double value = 0.0, signal, cur = 1.0;
    int o;
    for (o = 0; o < octaves; o++) {
        double nx = make_int32_range(x), ny = make_int32_range(y), nz = make_int32_range(z);
        signal = gradient_coherent_noise(nx, ny, nz, o);
        signal = 2.0 * fabs(signal) - 1.0;
        value += signal * cur;
        x *= lacunarity;
        y *= lacunarity;
        z *= lacunarity;
        cur *= persistence;
    }
    value += 0.5;
    return value;
}

#define RIDGED_MAX_OCTAVE 30
float libnoise_ridged_multi(int octaves, float lacunarity_f, float xf, float yf, float zf) {
    double x = xf, y = yf, z = zf, lacunarity = lacunarity_f;
    double weights[RIDGED_MAX_OCTAVE], frequency = 1.0;
    double signal, value = 0.0, weight = 1.0, offset = 1.0, gain = 2.0;
    int i, o;
    for (i = 0; i < RIDGED_MAX_OCTAVE; i++) {
        weights[i] = pow(frequency, -1.0);
        frequency *= lacunarity;
    }
    if (octaves > RIDGED_MAX_OCTAVE) octaves = RIDGED_MAX_OCTAVE; /* the reference throws; never reached by the tests */
    for (o = 0; o < octaves; o++) {
        double nx = make_int32_range(x), ny = make_int32_range(y), nz = make_int32_range(z);
        signal = gradient_coherent_noise(nx, ny, nz, o & 0x7fffffff);
        signal = fabs(signal);
        signal = offset - signal;
        signal *= signal;
        signal *= weight;
        weight = signal * gain;
        if (weight > 1.0) weight = 1.0;
        if (weight < 0.0) weight = 0.0;
        value += (signal * weights[o]);
        x *= lacunarity;
        y *= lacunarity;
        z *= lacunarity;
    }
    return (value * 1.25) - 1.0;
}

float libnoise_voronoi(float displacement_f, float xf, float yf, float zf) {
    double x = xf, y = yf, z = zf, displacement = displacement_f;
    int xInt = (x > 0.0 ? (int)x : (int)x - 1), yInt = (y > 0.0 ? (int)y : (int)y - 1), zInt = (z > 0.0 ? (int)z : (int)z - 1);
    double minDist = 2147483647.0, xc = 0, yc = 0, zc = 0;
    int xCur, yCur, zCur;
    for (zCur = zInt - 2; zCur <= zInt + 2; zCur++)
        for (yCur = yInt - 2; yCur <= yInt + 2; yCur++)
            for (xCur = xInt - 2; xCur <= xInt + 2; xCur++) {
                double xPos = xCur + value_noise(xCur, yCur, zCur, 0);
                double yPos = yCur + value_noise(xCur, yCur, zCur, 1);
                double zPos = zCur + value_noise(xCur, yCur, zCur, 2);
                double xDist = xPos - x, yDist = yPos - y, zDist = zPos - z;
                double dist = xDist * xDist + yDist * yDist + zDist * zDist;
                if (dist < minDist) {
                    minDist = dist;
                    xc = xPos;
                    yc = yPos;
                    zc = zPos;
                }
            }
    return 0.0 + (displacement * (double)value_noise((int)(floor(xc)), (int)(floor(yc)), (int)(floor(zc)), 0));
}
