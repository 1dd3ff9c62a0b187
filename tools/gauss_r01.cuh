// The round-1 Gaussian IIR kernel (full double scratch image), kept OUT of the library for A/B runs of tools/gauss_dev.cu:
// it is the implementation the round-1 parity tests validated against the oracle, bit for bit.  Not built into
// libmathmap_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstring>

#define MM_AA 0
#define MM_SUPERSAMPLING 0
#include "../mathmap_b200/csrc/runtime/mm_runtime.cuh"
#include "../mathmap_b200/csrc/backend/kernels.h"

namespace mmbackend {

// ---------------------------------------------------------------- Gaussian IIR
// Two threads own one line of one channel (a column in the vertical pass, a row
// in the horizontal pass) and run the two sweeps of the 4th-order recursion in
// double exactly as gauss.c:175-196, one sweep each, exchanging their per-sample
// state through a double scratch line; the sum of the two is narrowed to float.
// Sample k of line l, channel c is at data[l * line_stride + k * elem_stride + c]
// (floats).  In the vertical pass consecutive threads are consecutive
// (column, channel) floats, so every step of the sweep is a coalesced row access.
struct GaussCoeffs {
    double n_p[5], n_m[5], d_p[5], d_m[5], bd_p[5], bd_m[5];
};

struct IirState {
    double s1, s2, s3, s4, v1, v2, v3, v4;  // previous four samples and outputs in sweep direction
};

// The coefficients of a sweep direction (M: anticausal) are read straight from the kernel parameter: with the
// direction a template parameter they are constant-bank operands of the DMULs and cost no registers, which
// leaves room for a deeper prefetch.
template <bool M> __device__ __forceinline__ double cf_n(const GaussCoeffs &C, int i) { return M ? C.n_m[i] : C.n_p[i]; }
template <bool M> __device__ __forceinline__ double cf_d(const GaussCoeffs &C, int i) { return M ? C.d_m[i] : C.d_p[i]; }
// n[j] - bd[j], the factor of `initial` in the boundary terms
template <bool M> __device__ __forceinline__ double cf_nb(const GaussCoeffs &C, int i) {
    return __dsub_rn(M ? C.n_m[i] : C.n_p[i], M ? C.bd_m[i] : C.bd_p[i]);
}

// One recursion step in the steady state (step index >= 4).  Order of operations
// as in gauss.c:182-190: acc = 0; acc += n[i]*s[i] - d[i]*v[i] for i = 0..4, where
// the i = 0 term reads acc itself for v[0] (still 0).  d[0] is 0.0 in both
// directions, so "n[0]*s0 - d[0]*0.0" is n[0]*s0 bit for bit (x - 0 == x, also for
// x == -0); the leading "0.0 +" stays because it turns -0 into +0.
template <bool M> __device__ __forceinline__ double iir_step_steady(const GaussCoeffs &C, double s0, const IirState &st) {
    double acc = __dadd_rn(0.0, __dmul_rn(cf_n<M>(C, 0), s0));
    acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cf_n<M>(C, 1), st.s1), __dmul_rn(cf_d<M>(C, 1), st.v1)));
    acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cf_n<M>(C, 2), st.s2), __dmul_rn(cf_d<M>(C, 2), st.v2)));
    acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cf_n<M>(C, 3), st.s3), __dmul_rn(cf_d<M>(C, 3), st.v3)));
    acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cf_n<M>(C, 4), st.s4), __dmul_rn(cf_d<M>(C, 4), st.v4)));
    return acc;
}

// One of the first four steps of a sweep (terms = step index < 4): the recursion
// terms that exist, then the boundary terms (n[j] - bd[j]) * initial for j = terms+1..4.
template <bool M> __device__ __forceinline__ double iir_step_boundary(const GaussCoeffs &C, double s0, const IirState &st, int terms, double initial) {
    double acc = __dadd_rn(0.0, __dmul_rn(cf_n<M>(C, 0), s0));
    if (terms >= 1) acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cf_n<M>(C, 1), st.s1), __dmul_rn(cf_d<M>(C, 1), st.v1)));
    if (terms >= 2) acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cf_n<M>(C, 2), st.s2), __dmul_rn(cf_d<M>(C, 2), st.v2)));
    if (terms >= 3) acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cf_n<M>(C, 3), st.s3), __dmul_rn(cf_d<M>(C, 3), st.v3)));
    if (terms < 1) acc = __dadd_rn(acc, __dmul_rn(cf_nb<M>(C, 1), initial));
    if (terms < 2) acc = __dadd_rn(acc, __dmul_rn(cf_nb<M>(C, 2), initial));
    if (terms < 3) acc = __dadd_rn(acc, __dmul_rn(cf_nb<M>(C, 3), initial));
    acc = __dadd_rn(acc, __dmul_rn(cf_nb<M>(C, 4), initial));
    return acc;
}

__device__ __forceinline__ void iir_shift(IirState &st, double s0, double acc) {
    st.s4 = st.s3; st.s3 = st.s2; st.s2 = st.s1; st.s1 = s0;
    st.v4 = st.v3; st.v3 = st.v2; st.v2 = st.v1; st.v1 = acc;
}

// Both sweeps of a line run CONCURRENTLY, one thread each, and meet in the
// middle.  Neither sweep depends on the other's state, only the final sum does,
// so (with h = ceil(n/2)):
//   phase 1  causal thread:      k = 0 .. h-1,     stores its state vp[k] in the scratch line
//            anticausal thread:  k = n-1 .. h,     stores its state vm[k]
//   phase 2  causal thread:      k = h .. n-1,     out[k] = (float)(vp[k] + vm[k]) with vm[k] from the scratch line
//            anticausal thread:  k = h-1 .. 0,     out[k] = (float)(vm[k] + vp[k]) with vp[k] from the scratch line
// Every value is computed by the same double operations in the same order as
// the sequential reference; only WHEN each output is written changes.  Compared
// with one thread running both sweeps back to back this doubles the number of
// independent recursions in flight and halves the length of each, at the same
// memory traffic.  The pass may run in place: in phase 2 each thread reads
// samples only from the half it then overwrites, and the four-sample history
// that reaches across the middle is carried in registers.
//
// Tiles of IIR_T steps, prefetched ahead of the dependent chain of double
// operations (two tiles ahead in phase 1, one in phase 2): with 3.5 warps per
// scheduler there is little else to hide memory latency behind.  Samples stay as loaded
// (bytes in the column pass straight from an RGBA8 picture) and are converted
// on use -- a conversion next to the load would wait for the data.
#define IIR_T 8

// A line's samples are floats, or (column pass straight from an RGBA8 picture) bytes converted like render_image does
__device__ __forceinline__ float iir_value(float raw) { return raw; }
__device__ __forceinline__ float iir_value(unsigned char raw) { return mm_unit_from_byte(raw); }

// one step without prefetch: the boundary steps and the tail that does not fill a tile
template <bool M, bool COMBINE, class S>
__device__ __forceinline__ void iir_single(IirState &st, const GaussCoeffs &C, const S *pp, float *oo, double *ss, int t, double initial) {
    const double s0 = (double)iir_value(*pp);
    const double acc = t < 4 ? iir_step_boundary<M>(C, s0, st, t, initial) : iir_step_steady<M>(C, s0, st);
    if (COMBINE) *oo = (float)__dadd_rn(acc, *ss);
    else *ss = acc;
    iir_shift(st, s0, acc);
}

template <bool COMBINE, class S>
__device__ __forceinline__ void iir_tile_load(S (&a)[IIR_T], double (&v)[IIR_T], const S *pp, const double *ss, long long de, long long ds) {
#pragma unroll
    for (int i = 0; i < IIR_T; ++i) {
        a[i] = pp[i * de];
        if (COMBINE) v[i] = ss[i * ds];
    }
}
template <bool M, bool COMBINE, class S>
__device__ __forceinline__ void iir_tile_run(IirState &st, const GaussCoeffs &C, const S (&a)[IIR_T], const double (&v)[IIR_T], float *oo, double *ss,
                                             long long de, long long ds) {
#pragma unroll
    for (int i = 0; i < IIR_T; ++i) {
        const double s0 = (double)iir_value(a[i]);
        const double acc = iir_step_steady<M>(C, s0, st);
        if (COMBINE) oo[i * de] = (float)__dadd_rn(acc, v[i]);
        else ss[i * ds] = acc;
        iir_shift(st, s0, acc);
    }
}

// steps t0 .. t1-1 of one sweep; pp/oo/ss point at the sample of step t0, de/ds
// are the signed strides (floats / doubles) from one step to the next
template <bool M, bool COMBINE, class S>
__device__ __forceinline__ void iir_run(IirState &st, const GaussCoeffs &C, const S *pp, float *oo, double *ss, long long de, long long ds, int t0, int t1,
                                        double initial) {
    int t = t0;
#pragma unroll 1
    for (; t < t1 && t < 4; ++t, pp += de, oo += de, ss += ds) iir_single<M, COMBINE>(st, C, pp, oo, ss, t, initial);
    const int ntiles = t1 > t ? (t1 - t) / IIR_T : 0;
    const long long te = IIR_T * de, ts = IIR_T * ds;
    if (ntiles > 0 && !COMBINE) {
        // phase 1 (no scratch to read): samples two tiles ahead through a ring of three register tiles
        S a0[IIR_T], a1[IIR_T], a2[IIR_T];
        double v0[IIR_T], v1[IIR_T], v2[IIR_T];  // unused here
        iir_tile_load<COMBINE>(a0, v0, pp, ss, de, ds);
        if (ntiles > 1) iir_tile_load<COMBINE>(a1, v1, pp + te, ss + ts, de, ds);
#pragma unroll 1
        for (int tile = 0; tile < ntiles; tile += 3) {  // at the top of each round the ring holds tiles `tile` (a0) and `tile + 1` (a1)
            if (tile + 2 < ntiles) iir_tile_load<COMBINE>(a2, v2, pp + 2 * te, ss + 2 * ts, de, ds);
            iir_tile_run<M, COMBINE>(st, C, a0, v0, oo, ss, de, ds);
            pp += te; oo += te; ss += ts;
            if (tile + 1 >= ntiles) break;
            if (tile + 3 < ntiles) iir_tile_load<COMBINE>(a0, v0, pp + 2 * te, ss + 2 * ts, de, ds);
            iir_tile_run<M, COMBINE>(st, C, a1, v1, oo, ss, de, ds);
            pp += te; oo += te; ss += ts;
            if (tile + 2 >= ntiles) break;
            if (tile + 4 < ntiles) iir_tile_load<COMBINE>(a1, v1, pp + 2 * te, ss + 2 * ts, de, ds);
            iir_tile_run<M, COMBINE>(st, C, a2, v2, oo, ss, de, ds);
            pp += te; oo += te; ss += ts;
        }
        t += ntiles * IIR_T;
    } else if (ntiles > 0) {
        // phase 2: samples and the other sweep's states one tile ahead (two would not fit the 128 registers that
        // seven blocks per SM leave each thread)
        S a0[IIR_T], a1[IIR_T];
        double v0[IIR_T], v1[IIR_T];
        iir_tile_load<COMBINE>(a0, v0, pp, ss, de, ds);
#pragma unroll 1
        for (int tile = 0; tile < ntiles; tile += 2) {
            if (tile + 1 < ntiles) iir_tile_load<COMBINE>(a1, v1, pp + te, ss + ts, de, ds);
            iir_tile_run<M, COMBINE>(st, C, a0, v0, oo, ss, de, ds);
            pp += te; oo += te; ss += ts;
            if (tile + 1 >= ntiles) break;
            if (tile + 2 < ntiles) iir_tile_load<COMBINE>(a0, v0, pp + te, ss + ts, de, ds);
            iir_tile_run<M, COMBINE>(st, C, a1, v1, oo, ss, de, ds);
            pp += te; oo += te; ss += ts;
        }
        t += ntiles * IIR_T;
    }
#pragma unroll 1
    for (; t < t1; ++t, pp += de, oo += de, ss += ds) iir_single<M, COMBINE>(st, C, pp, oo, ss, t, initial);
}

// One thread of a pair; ANTI = false runs the causal sweep (positions ascending), true the anticausal one.
template <bool ANTI, class S>
__device__ __forceinline__ void iir_thread(const GaussCoeffs &C, const S *in, float *out, double *scratch, int pair, int nn, long long line_stride,
                                           long long elem_stride, long long scratch_line_stride, long long scratch_elem_stride) {
    const int line = pair >> 2, ch = pair & 3;
    const int h = nn - nn / 2;                      // causal phase 1 covers [0, h), anticausal [h, n)
    const int len1 = ANTI ? nn - h : h;
    const int k0 = ANTI ? nn - 1 : 0;
    const long long de = ANTI ? -elem_stride : elem_stride, ds = ANTI ? -scratch_elem_stride : scratch_elem_stride;
    const S *p = in + (size_t)line * line_stride + ch + (long long)k0 * elem_stride;
    float *o = out + (size_t)line * line_stride + ch + (long long)k0 * elem_stride;
    double *sc = scratch + (size_t)line * scratch_line_stride + ch + (long long)k0 * scratch_elem_stride;
    IirState st = {0, 0, 0, 0, 0, 0, 0, 0};
    double initial = 0.0;
    if (nn > 0) initial = (double)iir_value(*p);
    iir_run<ANTI, false>(st, C, p, o, sc, de, ds, 0, len1, initial);
    __syncthreads();
    iir_run<ANTI, true>(st, C, p + len1 * de, o + len1 * de, sc + len1 * ds, de, ds, len1, nn, initial);
}

// Block = IIR_PAIRS causal threads followed by IIR_PAIRS anticausal threads for
// the same IIR_PAIRS (line, channel) recursions; the sweep direction is uniform
// per warp, the hand-over between the phases is one __syncthreads().
#define IIR_PAIRS 32
template <class S>
__global__ void __launch_bounds__(2 * IIR_PAIRS, 7) gauss_iir_lines_kernel(const S *in, float *out, double *scratch, int nlines, int n, long long line_stride,
                                                                         long long elem_stride, long long scratch_line_stride,
                                                                         long long scratch_elem_stride, const __grid_constant__ GaussCoeffs C) {
    const bool anti = threadIdx.x >= IIR_PAIRS;
    const int pair = blockIdx.x * IIR_PAIRS + (threadIdx.x - (anti ? IIR_PAIRS : 0));
    const bool valid = pair < nlines * 4;
    const int nn = valid ? n : 0;
    if (anti) iir_thread<true>(C, in, out, scratch, valid ? pair : 0, nn, line_stride, elem_stride, scratch_line_stride, scratch_elem_stride);
    else iir_thread<false>(C, in, out, scratch, valid ? pair : 0, nn, line_stride, elem_stride, scratch_line_stride, scratch_elem_stride);
}


size_t gauss_iir_scratch_bytes_r01(int width, int height) { return sizeof(double) * 4 * (size_t)width * height; }

// in -> out (float4 [height][width]; may alias); scratch: gauss_iir_scratch_bytes().  With in_is_rgba8 the input is
// uchar4 [height][width] whose bytes stand for k/255 (never aliases out).
void launch_gauss_iir_r01(const void *in, bool in_is_rgba8, float *out, double *scratch, int width, int height, float sigma_h, float sigma_v,
                      cudaStream_t stream) {
    GaussCoeffs c;
    // vertical pass: lines are columns (in and out have the same element strides: 4 channels per pixel)
    { double raw[30]; gauss_iir_constants_host(sigma_v, raw); memcpy(&c, raw, sizeof c); }
    {
        int threads = width * 4;
        if (in_is_rgba8)
            gauss_iir_lines_kernel<unsigned char><<<(threads + IIR_PAIRS - 1) / IIR_PAIRS, 2 * IIR_PAIRS, 0, stream>>>(
                (const unsigned char *)in, out, scratch, width, height, 4, (long long)width * 4, 4, (long long)width * 4, c);
        else
            gauss_iir_lines_kernel<float><<<(threads + IIR_PAIRS - 1) / IIR_PAIRS, 2 * IIR_PAIRS, 0, stream>>>(
                (const float *)in, out, scratch, width, height, 4, (long long)width * 4, 4, (long long)width * 4, c);
    }
    // horizontal pass: lines are rows, in place on `out`
    { double raw[30]; gauss_iir_constants_host(sigma_h, raw); memcpy(&c, raw, sizeof c); }
    {
        int threads = height * 4;
        gauss_iir_lines_kernel<float><<<(threads + IIR_PAIRS - 1) / IIR_PAIRS, 2 * IIR_PAIRS, 0, stream>>>(out, out, scratch, height, width, (long long)width * 4, 4,
                                                                                                        (long long)width * 4, 4, c);
    }
}


}  // namespace mmbackend
