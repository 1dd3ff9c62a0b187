// Developer harness for the Gaussian IIR passes (not part of the library or the tests): runs the shipped kernel
// (mmbackend::launch_gauss_iir) and the round-1 full-scratch kernel (tools/gauss_r01.cuh) on the same inputs, compares
// raw float bits (NaN == NaN) and times both with CUDA events.
//   tools/build_gauss_dev.sh && tools/_bin/gauss_dev [size]
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../mathmap_b200/csrc/backend/kernels.h"
#include "gauss_r01.cuh"

using namespace mmbackend;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

static unsigned rng_state = 12345u;
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static long compare(const std::vector<float> &a, const std::vector<float> &b) {
    long bad = 0;
    for (size_t i = 0; i < a.size(); ++i) {
        unsigned x, y;
        memcpy(&x, &a[i], 4);
        memcpy(&y, &b[i], 4);
        if (x != y && !(std::isnan(a[i]) && std::isnan(b[i]))) {
            if (bad < 5) printf("    mismatch at %zu: %.9g (%08x) vs %.9g (%08x)\n", i, a[i], x, b[i], y);
            ++bad;
        }
    }
    return bad;
}

// mode 0: random floats with zero runs; 1: bytes; 2: floats with inf / NaN / -0 / negative values sprinkled in
static long check(int w, int h, float sh, float sv, int mode) {
    const size_t px = (size_t)w * h;
    std::vector<float> f(px * 4);
    std::vector<unsigned char> b(px * 4);
    for (size_t i = 0; i < px * 4; ++i) {
        b[i] = (unsigned char)(rnd() & 0xff);
        f[i] = (float)(rnd() & 0xffff) / 65535.0f;
        if ((rnd() & 7) == 0) { f[i] = 0.0f; b[i] = 0; }
        if (mode == 2) {
            unsigned r = rnd() % 997;
            if (r == 0) f[i] = INFINITY;
            else if (r == 1) f[i] = NAN;
            else if (r == 2) f[i] = -INFINITY;
            else if (r < 40) f[i] = -0.0f;
            else if (r < 80) f[i] = -f[i];
            else if (r < 90) f[i] = 1e-41f;
        }
    }
    if (mode == 2 && px > 64) {  // a clean region too: most lines of small pictures would be all NaN otherwise
        for (size_t i = 0; i < px * 4; ++i)
            if ((i / 4) % (size_t)w < (size_t)w / 2 && !std::isfinite(f[i])) f[i] = 0.25f;
    }
    void *din;
    float *o1, *o2;
    double *s1, *s2;
    const size_t in_bytes = mode == 1 ? px * 4 : px * 16;
    CK(cudaMalloc(&din, in_bytes));
    CK(cudaMalloc(&o1, px * 16));
    CK(cudaMalloc(&o2, px * 16));
    CK(cudaMalloc(&s1, gauss_iir_scratch_bytes(w, h)));
    CK(cudaMalloc(&s2, gauss_iir_scratch_bytes_r01(w, h)));
    CK(cudaMemcpy(din, mode == 1 ? (void *)b.data() : (void *)f.data(), in_bytes, cudaMemcpyHostToDevice));
    CK(cudaMemset(o1, 0xff, px * 16));
    CK(cudaMemset(o2, 0xee, px * 16));
    launch_gauss_iir(din, mode == 1, o1, s1, w, h, sh, sv, 0);
    launch_gauss_iir_r01(din, mode == 1, o2, s2, w, h, sh, sv, 0);
    CK(cudaDeviceSynchronize());
    std::vector<float> r1(px * 4), r2(px * 4);
    CK(cudaMemcpy(r1.data(), o1, px * 16, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r2.data(), o2, px * 16, cudaMemcpyDeviceToHost));
    long bad = compare(r1, r2);
    long bad_inplace = 0;
    if (mode != 1) {  // in place
        launch_gauss_iir(din, false, (float *)din, s1, w, h, sh, sv, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(r1.data(), din, px * 16, cudaMemcpyDeviceToHost));
        bad_inplace = compare(r1, r2);
    }
    printf("  %5d x %5d  sigma %6.2f %6.2f  mode %d: %ld mismatches, in place %ld\n", w, h, sh, sv, mode, bad, bad_inplace);
    cudaFree(din); cudaFree(o1); cudaFree(o2); cudaFree(s1); cudaFree(s2);
    return bad + bad_inplace;
}

static void timing(int w, int h, float sigma, bool bytes) {
    const size_t px = (size_t)w * h;
    void *din;
    float *o;
    double *s1, *s2;
    CK(cudaMalloc(&din, bytes ? px * 4 : px * 16));
    CK(cudaMalloc(&o, px * 16));
    CK(cudaMalloc(&s1, gauss_iir_scratch_bytes(w, h)));
    CK(cudaMalloc(&s2, gauss_iir_scratch_bytes_r01(w, h)));
    {
        std::vector<unsigned char> b(bytes ? px * 4 : px * 16);
        if (bytes) for (auto &v : b) v = (unsigned char)(rnd() & 0xff);
        else { float *f = (float *)b.data(); for (size_t i = 0; i < px * 4; ++i) f[i] = (float)(rnd() & 0xffff) / 65535.0f; }
        CK(cudaMemcpy(din, b.data(), b.size(), cudaMemcpyHostToDevice));
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int which = 0; which < 2; ++which) {
        float best = 1e30f, sum = 0;
        const int reps = 10;
        for (int r = -3; r < reps; ++r) {
            cudaEventRecord(e0);
            if (which == 0) launch_gauss_iir(din, bytes, o, s1, w, h, sigma, sigma, 0);
            else launch_gauss_iir_r01(din, bytes, o, s2, w, h, sigma, sigma, 0);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 0) { sum += ms; if (ms < best) best = ms; }
        }
        printf("  %s %d x %d, %s input, both passes: mean %.3f ms, best %.3f ms\n", which == 0 ? "new" : "r01", w, h, bytes ? "RGBA8" : "float", sum / reps, best);
    }
    cudaFree(din); cudaFree(o); cudaFree(s1); cudaFree(s2);
}

int main(int argc, char **argv) {
    const int big = argc > 1 ? atoi(argv[1]) : 8192;
    long bad = 0;
    const int shapes[][2] = {{1, 1}, {2, 3}, {7, 5}, {9, 1}, {1, 9}, {33, 17}, {130, 40}, {40, 130}, {1000, 37}, {37, 1000}, {517, 1031}, {64, 64}, {257, 255}, {31, 33}, {8, 16}, {16, 8}, {35, 36}, {36, 35}, {63, 65}, {66, 67}, {44, 21}, {20, 50}, {52, 19}, {12, 300}, {100, 9}};
    if (argc <= 2)
    for (auto &s : shapes)
        for (int mode = 0; mode < 3; ++mode) bad += check(s[0], s[1], 3.0f + (s[0] % 7), 0.7f + (s[1] % 5) * 2.5f, mode);
    if (argc <= 2) { bad += check(2048, 1536, 32.0f, 32.0f, 0); bad += check(2048, 1536, 32.0f, 32.0f, 1); }
    printf("total mismatches: %ld\n", bad);
    if (big > 0) {
        timing(big, big, 32.0f, true);
        timing(big, big, 32.0f, false);
    }
    return bad ? 1 : 0;
}
