// Statically compiled sm_100a kernels (everything that is not filter-specific):
//   * drawable -> floatmap resample (reference render_image, builtins/builtins.c:303-342)
//   * the native Gaussian blur: recursive IIR, columns then rows
//     (reference native-filters/gauss.c:127-262), and the sigma < 0.5 px FIR (:265-639)
//   * supersampling combine (reference mathmap_common.c:880-927)
// Compiled with -fmad=false: the recursions must round exactly like the host's
// separate double multiply / subtract / add.
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#define MM_AA 0
#define MM_SUPERSAMPLING 0
#include "../runtime/mm_runtime.cuh"
#include "kernels.h"

namespace mmbackend {

// ------------------------------------------------------------------ render_image
// out[y][x] = tuple(nearest sample of drawable at fx = (x - bx)/ax, fy = (y - by)/ay)
template <int EX, int EY>
__global__ void __launch_bounds__(256) drawable_to_floatmap_kernel(mm_image img, float4 *out, int width, int height, float ax, float bx, float ay,
                                                                   float by, mm_color edge_x, mm_color edge_y, int supersampling) {
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= width || y >= height) return;
    float fx = __fdiv_rn(__fsub_rn((float)x, bx), ax), fy = __fdiv_rn(__fsub_rn((float)y, by), ay);
    fx = __fmul_rn(fx, img.xf);
    fy = __fmul_rn(fy, img.yf);
    float px = __fmul_rn(__fadd_rn(fx, img.mx), img.sx), py = -__fmul_rn(__fsub_rn(fy, img.my), img.sy);
    if (!supersampling) { px = __fadd_rn(px, 0.5f); py = __fadd_rn(py, 0.5f); }
    int ix = mm_f2i(floorf(px)), iy = mm_f2i(floorf(py));
    mm_apply_edge_behaviour<EX, EY>(ix, iy, img.w, img.h);
    mm_color c;
    if (ix < 0 || ix >= img.w) c = edge_x;
    else if (iy < 0 || iy >= img.h) c = edge_y;
    else c = __byte_perm(__ldg((const unsigned *)img.data + ((size_t)iy * img.w + ix)), 0, 0x0123);
    out[(size_t)y * width + x] = make_float4(mm_red(c), mm_green(c), mm_blue(c), mm_alpha(c));
}

// The same lookup, but the result stays RGBA8 in memory order (R first): every channel of render_image's floatmap is
// k/255 of such a byte, so a consumer that converts on load (the blur's column pass) reads a quarter of the bytes.
template <int EX, int EY>
__global__ void __launch_bounds__(256) drawable_to_bytes_kernel(mm_image img, unsigned *out, int width, int height, float ax, float bx, float ay, float by,
                                                                mm_color edge_x, mm_color edge_y, int supersampling) {
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= width || y >= height) return;
    float fx = __fdiv_rn(__fsub_rn((float)x, bx), ax), fy = __fdiv_rn(__fsub_rn((float)y, by), ay);
    fx = __fmul_rn(fx, img.xf);
    fy = __fmul_rn(fy, img.yf);
    float px = __fmul_rn(__fadd_rn(fx, img.mx), img.sx), py = -__fmul_rn(__fsub_rn(fy, img.my), img.sy);
    if (!supersampling) { px = __fadd_rn(px, 0.5f); py = __fadd_rn(py, 0.5f); }
    int ix = mm_f2i(floorf(px)), iy = mm_f2i(floorf(py));
    mm_apply_edge_behaviour<EX, EY>(ix, iy, img.w, img.h);
    unsigned word;  // memory order
    if (ix < 0 || ix >= img.w) word = __byte_perm(edge_x, 0, 0x0123);
    else if (iy < 0 || iy >= img.h) word = __byte_perm(edge_y, 0, 0x0123);
    else word = __ldg((const unsigned *)img.data + ((size_t)iy * img.w + ix));
    out[(size_t)y * width + x] = word;
}

void launch_drawable_to_bytes(const mm_image &img, void *out, int width, int height, float ax, float bx, float ay, float by, int edge_x_mode,
                              int edge_y_mode, unsigned edge_x, unsigned edge_y, int supersampling, cudaStream_t stream) {
    dim3 grid((width + 31) / 32, (height + 7) / 8);
#define MM_CASE(EX, EY) \
    if (edge_x_mode == EX && edge_y_mode == EY) drawable_to_bytes_kernel<EX, EY><<<grid, 256, 0, stream>>>(img, (unsigned *)out, width, height, ax, bx, ay, by, edge_x, edge_y, supersampling);
    MM_CASE(0, 0) MM_CASE(0, 1) MM_CASE(0, 2) MM_CASE(0, 3)
    MM_CASE(1, 0) MM_CASE(1, 1) MM_CASE(1, 2) MM_CASE(1, 3)
    MM_CASE(2, 0) MM_CASE(2, 1) MM_CASE(2, 2) MM_CASE(2, 3)
    MM_CASE(3, 0) MM_CASE(3, 1) MM_CASE(3, 2) MM_CASE(3, 3)
#undef MM_CASE
}

// Is render_image of this drawable at width x height the identity on texels (output pixel (x, y) = texel (x, y), no
// edge pixel)?  Decided with the kernel's own float arithmetic, evaluated on the host for every column and row.
bool drawable_render_is_identity(const mm_image &img, int width, int height, float ax, float bx, float ay, float by, int supersampling) {
    if (img.w != width || img.h != height) return false;
    auto f2i = [](float f) { return (!(f > -2147483904.0f && f < 2147483648.0f)) ? (int)0x80000000 : (int)f; };
    for (int x = 0; x < width; ++x) {
        volatile float fx = ((float)x - bx) / ax;
        fx = fx * img.xf;
        volatile float px = (fx + img.mx) * img.sx;
        if (!supersampling) px = px + 0.5f;
        if (f2i(floorf(px)) != x) return false;
    }
    for (int y = 0; y < height; ++y) {
        volatile float fy = ((float)y - by) / ay;
        fy = fy * img.yf;
        volatile float t = (fy - img.my) * img.sy;
        volatile float py = -t;
        if (!supersampling) py = py + 0.5f;
        if (f2i(floorf(py)) != y) return false;
    }
    return true;
}

void launch_drawable_to_floatmap(const mm_image &img, float *out, int width, int height, float ax, float bx, float ay, float by, int edge_x_mode,
                                 int edge_y_mode, unsigned edge_x, unsigned edge_y, int supersampling, cudaStream_t stream) {
    dim3 grid((width + 31) / 32, (height + 7) / 8);
#define MM_CASE(EX, EY) \
    if (edge_x_mode == EX && edge_y_mode == EY) drawable_to_floatmap_kernel<EX, EY><<<grid, 256, 0, stream>>>(img, (float4 *)out, width, height, ax, bx, ay, by, edge_x, edge_y, supersampling);
    MM_CASE(0, 0) MM_CASE(0, 1) MM_CASE(0, 2) MM_CASE(0, 3)
    MM_CASE(1, 0) MM_CASE(1, 1) MM_CASE(1, 2) MM_CASE(1, 3)
    MM_CASE(2, 0) MM_CASE(2, 1) MM_CASE(2, 2) MM_CASE(2, 3)
    MM_CASE(3, 0) MM_CASE(3, 1) MM_CASE(3, 2) MM_CASE(3, 3)
#undef MM_CASE
}

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

// gauss.c:39-115, evaluated on the host in double exactly like the reference
static void find_iir_constants(GaussCoeffs &c, float std_dev) {
    const double div = sqrt(2 * M_PI) * std_dev;
    const double x0 = -1.783 / std_dev, x1 = -1.723 / std_dev, x2 = 0.6318 / std_dev, x3 = 1.997 / std_dev;
    const double x4 = 1.6803 / div, x5 = 3.735 / div, x6 = -0.6803 / div, x7 = -0.2598 / div;
    double *n_p = c.n_p, *n_m = c.n_m, *d_p = c.d_p, *d_m = c.d_m;
    n_p[0] = x4 + x6;
    n_p[1] = (exp(x1) * (x7 * sin(x3) - (x6 + 2 * x4) * cos(x3)) + exp(x0) * (x5 * sin(x2) - (2 * x6 + x4) * cos(x2)));
    n_p[2] = (2 * exp(x0 + x1) * ((x4 + x6) * cos(x3) * cos(x2) - x5 * cos(x3) * sin(x2) - x7 * cos(x2) * sin(x3)) + x6 * exp(2 * x0) +
              x4 * exp(2 * x1));
    n_p[3] = (exp(x1 + 2 * x0) * (x7 * sin(x3) - x6 * cos(x3)) + exp(x0 + 2 * x1) * (x5 * sin(x2) - x4 * cos(x2)));
    n_p[4] = 0.0;
    d_p[0] = 0.0;
    d_p[1] = -2 * exp(x1) * cos(x3) - 2 * exp(x0) * cos(x2);
    d_p[2] = 4 * cos(x3) * cos(x2) * exp(x0 + x1) + exp(2 * x1) + exp(2 * x0);
    d_p[3] = -2 * cos(x2) * exp(x0 + 2 * x1) - 2 * cos(x3) * exp(x1 + 2 * x0);
    d_p[4] = exp(2 * x0 + 2 * x1);
    for (int i = 0; i <= 4; i++) d_m[i] = d_p[i];
    n_m[0] = 0.0;
    for (int i = 1; i <= 4; i++) n_m[i] = n_p[i] - d_p[i] * n_p[0];
    double sum_n_p = 0.0, sum_n_m = 0.0, sum_d = 0.0;
    for (int i = 0; i <= 4; i++) { sum_n_p += n_p[i]; sum_n_m += n_m[i]; sum_d += d_p[i]; }
    const double a = sum_n_p / (1.0 + sum_d), b = sum_n_m / (1.0 + sum_d);
    for (int i = 0; i <= 4; i++) { c.bd_p[i] = d_p[i] * a; c.bd_m[i] = d_m[i] * b; }
}

void gauss_iir_constants_host(float std_dev, double *out30) {
    GaussCoeffs c;
    find_iir_constants(c, std_dev);
    memcpy(out30, &c, sizeof(double) * 30);
}

size_t gauss_iir_scratch_bytes_r01(int width, int height) { return sizeof(double) * 4 * (size_t)width * height; }

// in -> out (float4 [height][width]; may alias); scratch: gauss_iir_scratch_bytes().  With in_is_rgba8 the input is
// uchar4 [height][width] whose bytes stand for k/255 (never aliases out).
void launch_gauss_iir_r01(const void *in, bool in_is_rgba8, float *out, double *scratch, int width, int height, float sigma_h, float sigma_v,
                      cudaStream_t stream) {
    GaussCoeffs c;
    // vertical pass: lines are columns (in and out have the same element strides: 4 channels per pixel)
    find_iir_constants(c, sigma_v);
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
    find_iir_constants(c, sigma_h);
    {
        int threads = height * 4;
        gauss_iir_lines_kernel<float><<<(threads + IIR_PAIRS - 1) / IIR_PAIRS, 2 * IIR_PAIRS, 0, stream>>>(out, out, scratch, height, width, (long long)width * 4, 4,
                                                                                                        (long long)width * 4, 4, c);
    }
}

// ------------------------------------------------- sigma < 0.5 px: truncated FIR
// gauss.c:265-639.  `curve` holds exp(-i*i/(2 sigma^2)) for i = 0..length as floats.
// The reference picks between a run-length variant (when more than 3/4 of a
// line repeats its predecessor; note its int-truncated total and partial sums,
// gauss.c:383-404) and the direct sum per line and channel; both are evaluated
// here per line by one thread, in the reference's order of float operations.
struct RleCurve {
    int length;
    float total;
    const float *curve;  // device: curve[0..length]
    const float *sum;    // device: sum[-length..length] stored at [i + length]
};

__global__ void __launch_bounds__(128) gauss_rle_lines_kernel(const float *in, float *outp, int nlines, int n, long long line_stride, long long elem_stride,
                                                              RleCurve K) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nlines * 4) return;
    int line = tid >> 2, ch = tid & 3;
    const float *p = in + (size_t)line * line_stride + ch;
    float *o = outp + (size_t)line * line_stride + ch;
    const int L = K.length;
    auto pix = [&](int i) { i = i < 0 ? 0 : (i >= n ? n - 1 : i); return p[(size_t)i * elem_stride]; };
    // run_length_encode's `same`: samples equal to their right neighbour (scanned from the end)
    int same = 0;
    {
        float last = pix(n - 1);
        for (int i = n - 1; i >= 0; --i) {
            float c = pix(i);
            if (c == last) same++; else last = c;
        }
    }
    const bool encoded = same > (3 * n) / 4;
    const int ctotal_int = (int)K.total;
    for (int col = 0; col < n; ++col) {
        float val = 0.0f;
        if (!encoded) {
            val = __fadd_rn(val, __fmul_rn(pix(col), K.curve[0]));
            for (int i = 1; i <= L; ++i) val = __fadd_rn(val, __fmul_rn(__fadd_rn(pix(col + i), pix(col - i)), K.curve[i]));
            val = __fdiv_rn(val, K.total);
        } else {
            // walk runs of equal samples from col-L to col+L; rle counts runs looking rightwards
            int i = -L;
            float s1 = K.sum[0];
            while (true) {
                // length of the run starting at col+i (bounded by the padded line end)
                float v = pix(col + i);
                int nb = 1;
                while (col + i + nb < n + L && pix(col + i + nb) == v) ++nb;
                if (i + nb > L) { val = __fadd_rn(val, __fmul_rn(v, __fsub_rn(K.sum[2 * L], s1))); break; }
                int s2 = (int)K.sum[i + nb + L];
                val = __fadd_rn(val, __fmul_rn(v, __fsub_rn((float)s2, s1)));
                s1 = (float)s2;
                i += nb;
            }
            val = __fdiv_rn(val, (float)ctotal_int);
        }
        o[(size_t)col * elem_stride] = val;
    }
}

// make_rle_curve, gauss.c:265-300: the curve reaches out to where it drops below 1/255, whatever that length is
static int rle_curve_length(double sigma) {
    const double sigma2 = 2 * sigma * sigma;
    const double l = sqrt(-sigma2 * log(1.0 / 255.0));
    int n = (int)(ceil(l) * 2);
    if ((n % 2) == 0) n += 1;
    return n / 2;
}
// floats of device memory one curve takes: curve[0..length] and sum[0..2 length]
static size_t rle_curve_floats(double sigma) { return 3 * (size_t)rle_curve_length(sigma) + 2; }
size_t gauss_rle_curve_bytes(float sigma_h, float sigma_v) {
    return sizeof(float) * ((sigma_h > 0.0f ? rle_curve_floats(sigma_h) : 0) + (sigma_v > 0.0f ? rle_curve_floats(sigma_v) : 0)) + 16;
}
static void make_rle_curve(double sigma, RleCurve &K, float *dev, cudaStream_t stream) {
    const double sigma2 = 2 * sigma * sigma;
    const int length = rle_curve_length(sigma);
    K.length = length;
    std::vector<float> curve(2 * (size_t)length + 1), host(rle_curve_floats(sigma));
    curve[length] = 1.0f;
    for (int i = 1; i <= length; i++) {
        float temp = (float)exp(-(i * i) / sigma2);
        curve[length - i] = temp;
        curve[length + i] = temp;
    }
    float *hc = host.data(), *hs = host.data() + length + 1;
    for (int i = 0; i <= length; ++i) hc[i] = curve[length + i];
    hs[0] = 0;
    for (int i = 1; i <= length * 2; i++) hs[i] = curve[i - 1] + hs[i - 1];
    K.total = hs[2 * length] - hs[0];
    K.curve = dev;
    K.sum = dev + length + 1;
    // pageable source: the call returns once the bytes are staged, so `host` may go out of scope
    cudaMemcpyAsync(dev, host.data(), sizeof(float) * host.size(), cudaMemcpyHostToDevice, stream);
}

// in -> out (float4 [height][width], may alias); tmp is a third buffer of that size; curve_mem: gauss_rle_curve_bytes()
void launch_gauss_rle(const float *in, float *tmp, float *out, int width, int height, float sigma_h, float sigma_v, void *curve_mem, cudaStream_t stream) {
    RleCurve K;
    const float *src = in;
    float *dev = (float *)curve_mem;
    size_t bytes = sizeof(float) * 4 * (size_t)width * height;
    if (sigma_v > 0.0f) {
        make_rle_curve(sigma_v, K, dev, stream);
        dev += rle_curve_floats(sigma_v);
        int threads = width * 4;
        gauss_rle_lines_kernel<<<(threads + 127) / 128, 128, 0, stream>>>(src, tmp, width, height, 4, (long long)width * 4, K);
        src = tmp;
    }
    if (sigma_h > 0.0f) {
        make_rle_curve(sigma_h, K, dev, stream);
        int threads = height * 4;
        // a pass never runs in place (every output sample reads its neighbours): when the horizontal pass is the only
        // one and the caller's buffers alias, it goes through tmp
        float *dst = src == out ? tmp : out;
        gauss_rle_lines_kernel<<<(threads + 127) / 128, 128, 0, stream>>>(src, dst, height, width, (long long)width * 4, 4, K);
        if (dst != out) cudaMemcpyAsync(out, dst, bytes, cudaMemcpyDeviceToDevice, stream);
    } else
        cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToDevice, stream);
}

// -------------------------------------------------------- supersampling combine
// out = (l1[c] + l1[c+1] + 2*l2[c] + l3[c] + l3[c+1]) / 6 on bytes, where the
// "long" image has one more column and row r of it was rendered with offsets
// (-0.5, -0.5); its last row repeats the one before (see oracle/runtime/driver.c).
__global__ void __launch_bounds__(256) supersample_combine_kernel(const unsigned char *shortimg, const unsigned char *longimg, unsigned char *out, int width,
                                                                  int height, int long_rows, int bpp) {
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= width || y >= height) return;
    const size_t lw = (size_t)(width + 1) * bpp;
    const unsigned char *l1 = longimg + (size_t)y * lw, *l3 = longimg + (size_t)(y + 1 < long_rows ? y + 1 : y) * lw;
    const unsigned char *l2 = shortimg + (size_t)y * width * bpp;
    for (int i = 0; i < bpp; ++i) {
        int v = l1[x * bpp + i] + l1[(x + 1) * bpp + i] + 2 * l2[x * bpp + i] + l3[x * bpp + i] + l3[(x + 1) * bpp + i];
        out[((size_t)y * width + x) * bpp + i] = (unsigned char)(v / 6);
    }
}

void launch_supersample_combine(const unsigned char *shortimg, const unsigned char *longimg, unsigned char *out, int width, int height, int long_rows,
                                int bpp, cudaStream_t stream) {
    dim3 grid((width + 31) / 32, (height + 7) / 8);
    supersample_combine_kernel<<<grid, 256, 0, stream>>>(shortimg, longimg, out, width, height, long_rows, bpp);
}

}  // namespace mmbackend
