// The native Gaussian blur's recursive (IIR) passes for sm_100a: reference native-filters/gauss.c:127-262
// (gauss_iir: a causal and an anticausal 4th-order recursion in double per line and channel, summed and narrowed
// to float; columns first, then rows).  Compiled with -fmad=false.
//
// What bounds this path (8192^2 RGBA: 65 536 independent recursions of 8192 steps, twice):
//   * FP64 issue: B200 has 64 FP64 lanes per SM, one warp instruction every two cycles per scheduler;
//   * HBM: the two sweeps of a line need each other's value at every sample, and they arrive from opposite ends.
//     Handing a sweep's values over through memory costs 16 B per sample and channel (8 written, 8 read): 8.6 GB per
//     blur against 2.7 GB of samples and results -- the round-1 kernel ran at 0.66-0.75 of HBM and could not get
//     faster.
// This kernel hands over CHECKPOINTS instead (the four previous outputs of a sweep every CK_S steps, 2 B per sample)
// and re-runs the other sweep's recursion block by block where its values are needed -- the same double operations in
// the same order, hence the same bits -- so each line costs three sweeps of FP64 work instead of two, and the work per
// step is cut from 19 to 15 FP64 instructions by dropping the operations that cannot change a bit (see step_fast).
//
// Work split: two threads per (line, channel), one per sweep direction, that meet in the middle:
//   phase 1   each thread runs its own sweep over its half of the line, keeping only checkpoints;
//   phase 2   each thread continues its sweep over the other half; for every block of CK_S samples it first re-runs the
//             OTHER sweep from that sweep's checkpoint into shared memory (one block ahead), then adds.  The two
//             recursions of a thread are independent dependent chains, which hides FP64 latency.
// Samples reach the threads through a shared-memory ring filled by cp.async two tiles ahead (no registers held by
// loads in flight); checkpoints of the block after next are loaded into registers one block ahead.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "kernels.h"

namespace mmbackend {

namespace {

struct Coeffs {
    double n_p[5], n_m[5], d_p[5], d_m[5], bd_p[5], bd_m[5];
};

constexpr int CK_S = 16;      // steps per block = checkpoint spacing
constexpr int CK_R = 4;       // sample tiles in the ring
constexpr int CK_PAIRS = 32;  // (line, channel) recursions per thread block
constexpr int CK_T = 2 * CK_PAIRS;

#define CK_DEV __device__ __forceinline__

// coefficients of a sweep direction (M: anticausal) as constant-bank operands of the kernel parameter
template <bool M> CK_DEV double cn(const Coeffs &C, int i) { return M ? C.n_m[i] : C.n_p[i]; }
template <bool M> CK_DEV double cd(const Coeffs &C, int i) { return M ? C.d_m[i] : C.d_p[i]; }
template <bool M> CK_DEV double cnb(const Coeffs &C, int i) { return __dsub_rn(M ? C.n_m[i] : C.n_p[i], M ? C.bd_m[i] : C.bd_p[i]); }

// A line's samples are floats, or (column pass straight from an RGBA8 picture) bytes that stand for k/255 narrowed to
// float like render_image makes them; the byte -> double conversion is a 256-entry table in shared memory.
// k/255.0 narrowed to float, exactly (mm_runtime.cuh: mm_unit_from_byte)
CK_DEV float unit_from_byte(unsigned k) {
    const float f = (float)k;
    return __fmaf_rn(f, 0.003921568859368563f, __fmul_rn(f, -2.319175823606301e-10f));
}
CK_DEV double conv(float e, const double *) { return (double)e; }
CK_DEV double conv(unsigned char e, const double *lut) { return lut[e]; }
// raw + |0 * other|: raw itself (with -0 turned into +0) unless `other` is infinite or NaN, then NaN.  Restores, on
// the FP32 pipe, the NaN that the reference's 0.0 * sample terms produce (see step_fast).  Bytes are always finite.
CK_DEV float poisoned(float raw, float other) { return __fadd_rn(raw, fabsf(__fmul_rn(0.0f, other))); }
CK_DEV unsigned char poisoned(unsigned char raw, unsigned char) { return raw; }

template <class S> struct Chain {
    double s1, s2, s3, s4, v1, v2, v3, v4;  // previous four samples and outputs in sweep direction
    S r1, r2, r3, r4;                       // the previous four samples as loaded
};
template <class S> CK_DEV void chain_clear(Chain<S> &c) {
    c.s1 = c.s2 = c.s3 = c.s4 = c.v1 = c.v2 = c.v3 = c.v4 = 0.0;
    c.r1 = c.r2 = c.r3 = c.r4 = S(0);
}

// One of the first four steps of a sweep (t = step index < 4), operation by operation as gauss.c:175-196: the terms
// that exist, then the boundary terms (n[j] - bd[j]) * initial for j = t+1..4.  d[0] is 0.0 in both directions, so
// "n[0]*s0 - d[0]*0.0" is n[0]*s0 bit for bit.
template <bool M, class S> CK_DEV double step_boundary(const Coeffs &C, Chain<S> &st, S raw, int t, double initial, const double *lut) {
    const double s0 = conv(raw, lut);
    double acc = __dadd_rn(0.0, __dmul_rn(cn<M>(C, 0), s0));
    if (t >= 1) acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cn<M>(C, 1), st.s1), __dmul_rn(cd<M>(C, 1), st.v1)));
    if (t >= 2) acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cn<M>(C, 2), st.s2), __dmul_rn(cd<M>(C, 2), st.v2)));
    if (t >= 3) acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cn<M>(C, 3), st.s3), __dmul_rn(cd<M>(C, 3), st.v3)));
    if (t < 1) acc = __dadd_rn(acc, __dmul_rn(cnb<M>(C, 1), initial));
    if (t < 2) acc = __dadd_rn(acc, __dmul_rn(cnb<M>(C, 2), initial));
    if (t < 3) acc = __dadd_rn(acc, __dmul_rn(cnb<M>(C, 3), initial));
    acc = __dadd_rn(acc, __dmul_rn(cnb<M>(C, 4), initial));
    st.s4 = st.s3; st.s3 = st.s2; st.s2 = st.s1; st.s1 = s0;
    st.r4 = st.r3; st.r3 = st.r2; st.r2 = st.r1; st.r1 = raw;
    st.v4 = st.v3; st.v3 = st.v2; st.v2 = st.v1; st.v1 = acc;
    return acc;
}

// A steady-state step (t >= 4).  The reference evaluates, in this order,
//     acc = 0.0 + n0*s0;  acc += n1*s1 - d1*v1;  acc += n2*s2 - d2*v2;  acc += n3*s3 - d3*v3;  acc += n4*s4 - d4*v4
// (19 FP64 operations).  find_iir_constants makes n_p[4] and n_m[0] exactly 0.0, which leaves operations that cannot
// change a bit of the blur's output:
//   causal      "n4*s4 - d4*v4" is -(d4*v4) when s4 is finite, so "acc + (...)" is acc - d4*v4 (same rounding; and the
//               accumulator is never -0, so a zero term of either sign leaves it alone);  "0.0 + n0*s0" is n0*s0 unless
//               s0 is -0 (n0 > 0), which the sample's "+ |0 * ...|" below rules out.  The values are the reference's
//               bit for bit and never -0.
//   anticausal  "0.0 + 0.0*s0" is +0 for a finite s0, and "+0 + x" is x unless x is -0: the sweep's values are the
//               reference's except that a zero may carry the other sign, which neither the following steps (a zero
//               term or factor changes nothing but zero signs) nor the final sum notice: the causal value it is added
//               to is never -0, and +0 + (+-0) = +0.
// An infinite or NaN sample makes the dropped products NaN in the reference (0 * inf): four steps later in the causal
// sweep, in the same step in the anticausal one, and NaN from then on.  That is restored on the FP32 pipe: the causal
// sweep adds |0 * s[t-4]| to the sample it converts in step t, the anticausal sweep converts s[t-1] in step t (it has
// no n0 term) and adds |0 * s[t]|.  15 FP64 operations per step in both directions.
// The sample a steady-state step feeds to its recursion, as a double (see above): causal e[t] = s[t] + |0 * s[t-4]|,
// anticausal e[t] = s[t-1] + |0 * s[t]|.  Independent of the recursion, so it is prepared ahead of it.
template <bool M, class S> CK_DEV double prep(S raw, S back1, S back4, const double *lut) {
    return M ? conv(poisoned(back1, raw), lut) : conv(poisoned(raw, back4), lut);
}
// the 15 FP64 operations of a steady-state step on the prepared sample e
template <bool M, class S> CK_DEV double recur(const Coeffs &C, Chain<S> &st, double e) {
    double acc;
    if (!M) {
        acc = __dmul_rn(cn<M>(C, 0), e);
        acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cn<M>(C, 1), st.s1), __dmul_rn(cd<M>(C, 1), st.v1)));
        acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cn<M>(C, 2), st.s2), __dmul_rn(cd<M>(C, 2), st.v2)));
        acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cn<M>(C, 3), st.s3), __dmul_rn(cd<M>(C, 3), st.v3)));
        acc = __dsub_rn(acc, __dmul_rn(cd<M>(C, 4), st.v4));
        st.s4 = st.s3; st.s3 = st.s2; st.s2 = st.s1; st.s1 = e;
    } else {
        acc = __dsub_rn(__dmul_rn(cn<M>(C, 1), e), __dmul_rn(cd<M>(C, 1), st.v1));
        acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cn<M>(C, 2), st.s2), __dmul_rn(cd<M>(C, 2), st.v2)));
        acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cn<M>(C, 3), st.s3), __dmul_rn(cd<M>(C, 3), st.v3)));
        acc = __dadd_rn(acc, __dsub_rn(__dmul_rn(cn<M>(C, 4), st.s4), __dmul_rn(cd<M>(C, 4), st.v4)));
        st.s4 = st.s3; st.s3 = st.s2; st.s2 = e;
    }
    st.v4 = st.v3; st.v3 = st.v2; st.v2 = st.v1; st.v1 = acc;
    return acc;
}
template <bool M, class S> CK_DEV double step_fast(const Coeffs &C, Chain<S> &st, S raw, const double *lut) {
    const double acc = recur<M>(C, st, prep<M>(raw, st.r1, st.r4, lut));
    st.r4 = st.r3; st.r3 = st.r2; st.r2 = st.r1; st.r1 = raw;
    return acc;
}

// The state of a sweep before its step u0 > 0: outputs from its checkpoint, samples from the line (h[k] = sample of
// step u0 - 1 - k as loaded).  History samples need no "+ |0 * ...|": where that would matter the checkpointed outputs
// are already NaN.
template <class S> CK_DEV void chain_restore(Chain<S> &r, const double (&v)[4], const S (&h)[4], const double *lut) {
    r.v1 = v[0]; r.v2 = v[1]; r.v3 = v[2]; r.v4 = v[3];
    r.r1 = h[0]; r.r2 = h[1]; r.r3 = h[2]; r.r4 = h[3];
    r.s1 = conv(h[0], lut); r.s2 = conv(h[1], lut); r.s3 = conv(h[2], lut); r.s4 = conv(h[3], lut);
}

CK_DEV void cp_async4(unsigned smem_addr, const void *g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(g) : "memory");
}
CK_DEV void cp_async16(unsigned smem_addr, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
CK_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> CK_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// How a warp fills a sample tile (CK_S steps of its 32 recursions; element i of thread tid at slot[i * CK_T + tid]).
// The 32 recursions of a warp are 8 pixels x 4 channels, and a pixel's four channels are adjacent in memory in both
// passes, so the lanes copy 16-byte chunks for one another:
//   floats      a chunk = one pixel of one step (4 lanes' samples); lane l copies the pixel (l % 8) of steps l / 8 + 4m,
//               m = 0..3: 4 requests per lane and tile instead of 16 four-byte ones, which with their 64-bit address
//               arithmetic were a quarter of the kernel's instructions;
//   bytes       (column pass, rows of 32 bytes per step) a chunk = 16 recursions of one step; lane l copies half (l % 2) of
//               step l / 2: 1 request per lane and tile.  Needs rows that are multiples of 16 bytes; otherwise lanes 0-7
//               copy the 4 bytes of one pixel each, step by step (`bytes4`).
// The loader of a thread: which chunk column it copies (pointer of that column at the sweep's first / last sample) and where
// it lands.
template <class S> struct TileLoader {
    const char *p0, *q0;  // the chunk column at own step 0 / at the other sweep's step 0
    unsigned off;         // byte offset of the chunk column inside a tile row
    int i_first;          // first step this lane copies
    bool bytes4;          // bytes only: the 4-byte fallback (lanes 0-7)
};
template <class S> CK_DEV void tile_issue(S *slot, const TileLoader<S> &L, const char *col, long long deb, int cnt) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(slot) + L.off;
    constexpr unsigned row = CK_T * (unsigned)sizeof(S);
    if (sizeof(S) == 4) {
        const char *p = col + L.i_first * deb;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            if (L.i_first + 4 * m < cnt) cp_async16(sa + (L.i_first + 4 * m) * row, p);
            p += 4 * deb;
        }
    } else if (!L.bytes4) {
        if (L.i_first < cnt) cp_async16(sa + L.i_first * row, col + L.i_first * deb);
    } else if (L.i_first == 0) {  // lanes 0-7
        const char *p = col;
#pragma unroll 4
        for (int i = 0; i < cnt; ++i) {
            cp_async4(sa + i * row, p);
            p += deb;
        }
    }
    cp_async_commit();
}

// Shared memory of a block, as separate arrays so that the compiler knows that a store of a re-run value can never
// alias a sample or table read (with one dynamic buffer every such store fenced the loads behind it and the two chains
// of a thread ran one after the other).
template <class S> struct CkShared {
    double (*w)[CK_T];            // [CK_S][CK_T] values of the other sweep for the block being consumed / re-run (see ck_thread)
    S (*ring)[CK_S][CK_T];        // [CK_R][CK_S][CK_T] sample tiles
    const double *lut;            // [256], bytes only
};

// Where a pass's results go.  OUT_FLOAT: a float image laid out like the input (its own line stride).  OUT_RGBA8 (the
// row pass only, when the filter's pixel IS the blurred sample, see invocation.cpp): every thread quantises its channel like
// the pixel kernel's store (mm_runtime.cuh: mm_clamp01 with the template's MIN/MAX, times 255.0 truncated,
// new_template.c.in:272-293) and stores its byte; the four channel lanes of a pixel write one word between them (the
// LSU merges them), which measured faster than gathering the word with two shuffles per step.
enum { OUT_FLOAT = 0, OUT_RGBA8 = 1 };
CK_DEV unsigned quant_byte(float v) {
    const float m = (1.0f < v) ? 1.0f : v;   // MIN(1, v): NaN stays
    const float c = (0.0f < m) ? m : 0.0f;   // MAX(0, m): NaN -> 0
    return (unsigned)__float_as_int(__fmaf_rz(c, 255.0f, 8388608.0f)) & 0xffu;  // 2^23 + floor(c * 255), exact product
}
template <int OUT> struct OutCursor;
template <> struct OutCursor<OUT_FLOAT> {
    float *p;
    long long de;
    CK_DEV void put(float v) { *p = v; p += de; }
    CK_DEV void put4(float a, float b, float c, float d) { p[0] = a; p[de] = b; p[2 * de] = c; p[3 * de] = d; p += 4 * de; }
};
template <> struct OutCursor<OUT_RGBA8> {
    unsigned char *p;  // this thread's channel byte of the pixel of the next step in its row
    int dir;           // +4 / -4 bytes per step
    CK_DEV void put(float v) { *p = (unsigned char)quant_byte(v); p += dir; }
    CK_DEV void put4(float a, float b, float c, float d) {
        p[0] = (unsigned char)quant_byte(a); p[dir] = (unsigned char)quant_byte(b);
        p[2 * dir] = (unsigned char)quant_byte(c); p[3 * dir] = (unsigned char)quant_byte(d);
        p += 4 * dir;
    }
};

// The loops below are unrolled by CK_U = 4 steps, not by a whole block: the recursion looks four steps back, so after
// four steps every state variable is back in its register (no moves), and the bodies stay small enough for the
// instruction cache -- fully unrolled blocks (17 KB per direction and parity) stalled on instruction fetch for half of
// all cycles.
constexpr int CK_U = 4;
template <class S> CK_DEV void load_group(S (&raw)[CK_U], const S *p, int stride) {
#pragma unroll
    for (int k = 0; k < CK_U; ++k) raw[k] = p[k * stride];
}
// In all three loops a group's samples are loaded before its first step, and what happens to a step's value (kept in
// w, or added to the other sweep's value and stored) is done after the group, when the values are the state's v4..v1:
// a store or sum placed right behind its step would hold up the in-order issue of the next step until the dependent
// chain of the step has finished.
// CK_S steady-state steps of one sweep over a full tile, nothing kept but the state (phase 1)
template <bool M, class S> CK_DEV void plain_round(const Coeffs &C, Chain<S> &st, const S *tile, const double *lut) {
#pragma unroll 1
    for (int g = 0; g < CK_S / CK_U; ++g) {
        S raw[CK_U];
        load_group(raw, tile, CK_T);
#pragma unroll
        for (int k = 0; k < CK_U; ++k) step_fast<M>(C, st, raw[k], lut);
        tile += CK_U * CK_T;
    }
}
// A phase-2 round: the own sweep over its block (own_round), then the re-run of the other sweep over the next one
template <bool M, class S> CK_DEV void rerun_round(const Coeffs &C, Chain<S> &r, const S *tile, double *w, const double *lut) {
#pragma unroll 1
    for (int g = 0; g < CK_S / CK_U; ++g) {
        S raw[CK_U];
        load_group(raw, tile, CK_T);
#pragma unroll
        for (int k = 0; k < CK_U; ++k) step_fast<M>(C, r, raw[k], lut);
        w[0] = r.v4; w[CK_T] = r.v3; w[2 * CK_T] = r.v2; w[3 * CK_T] = r.v1;
        tile += CK_U * CK_T;
        w += CK_U * CK_T;
    }
}
template <bool M, class S, class O>
CK_DEV void own_round(const Coeffs &C, Chain<S> &st, const S *tile, const double *w, O &out, const double *lut) {
    tile += (CK_S - 1) * CK_T;
    w += (CK_S - 1) * CK_T;
#pragma unroll 1
    for (int g = 0; g < CK_S / CK_U; ++g) {
        S raw[CK_U];
        double wv[CK_U];
        load_group(raw, tile, -CK_T);
        load_group(wv, w, -CK_T);
#pragma unroll
        for (int k = 0; k < CK_U; ++k) step_fast<M>(C, st, raw[k], lut);
        out.put4((float)__dadd_rn(st.v4, wv[0]), (float)__dadd_rn(st.v3, wv[1]), (float)__dadd_rn(st.v2, wv[2]), (float)__dadd_rn(st.v1, wv[3]));
        tile -= CK_U * CK_T;
        w -= CK_U * CK_T;
    }
}
// One thread of a pair; ANTI = false runs the causal sweep (positions ascending), true the anticausal one.
// Own step t is position ANTI ? n-1-t : t; the other sweep's step u = n-1-t is the same position.
// Checkpoint j >= 1 of a sweep = its v1..v4 before its step j * CK_S, at ckpt[((dir * J + j) * 4 + i) * npairs + pair].
// Output: OUT_FLOAT: out is a float image, line l at out + l * out_line_stride (floats), samples elem_stride apart like the
// input's; OUT_RGBA8: out is bytes, line l at out + l * out_line_stride (bytes), sample k at its word k.
template <bool ANTI, int OUT, class S>
CK_DEV void ck_thread(const Coeffs &C, const S *in, void *out, long long out_line_stride, double *ckpt, int J, int pair, int npairs, int n,
                      long long line_stride, long long elem_stride, bool bytes16, const CkShared<S> &sm) {
    const int tid = threadIdx.x, lane = tid & 31;
    // Threads beyond the last recursion (the last block of a picture whose line count is not a multiple of 8) run one
    // of the last line's again, in the same warp and in lockstep with the thread that owns it: they read what it reads and
    // store what it stores.  That keeps the loops free of predicates -- with the stores under a condition the compiler
    // sank the loads of the other sweep's values into the branch, right in front of their use.
    // (the same channel of the last line -- or, where the warp copies bytes in 16-byte chunks, the recursion 16 before it: what the
    // repeated chunk puts into this thread's slot of the ring)
    if (pair >= npairs) pair = (sizeof(S) == 1 && bytes16) ? pair - 16 : npairs - 4 + (pair & 3);
    const int line = pair >> 2, ch = pair & 3;
    const int h = n - n / 2;
    const int len1 = ANTI ? n - h : h, len2 = n - len1;  // own / other phase-1 lengths
    const int J1 = (len1 + CK_S - 1) / CK_S, J2 = (len2 + CK_S - 1) / CK_S;
    const long long de = ANTI ? -elem_stride : elem_stride;
    const long long first = (long long)(ANTI ? n - 1 : 0) * elem_stride, last = (long long)(ANTI ? 0 : n - 1) * elem_stride;
    const S *base_in = in + (size_t)line * line_stride + ch;
    const S *P0 = base_in + first;   // own step 0
    const S *Q0 = base_in + last;    // the other sweep's step 0
    const double *lut = sm.lut;
    // what this thread copies into the ring for its warp (see TileLoader); out-of-range chunks repeat the last pixel's
    TileLoader<S> L;
    {
        const int wpair0 = blockIdx.x * CK_PAIRS;  // the warp's first recursion (both warps of a block serve the same 32)
        int gpair;
        if (sizeof(S) == 4) {
            gpair = wpair0 + 4 * (lane & 7);
            if (gpair >= npairs) gpair = npairs - 4;
            L.off = (unsigned)((tid - lane + 4 * (lane & 7)) * sizeof(S));
            L.i_first = lane >> 3;
            L.bytes4 = false;
        } else if (bytes16) {
            gpair = wpair0 + 16 * (lane & 1);
            if (gpair >= npairs) gpair = npairs - 16;
            L.off = (unsigned)(tid - lane + 16 * (lane & 1));
            L.i_first = lane >> 1;
            L.bytes4 = false;
        } else {
            gpair = wpair0 + 4 * (lane & 7);
            if (gpair >= npairs) gpair = npairs - 4;
            L.off = (unsigned)(tid - lane + 4 * (lane & 7));
            L.i_first = lane < 8 ? 0 : 1;  // lanes 8-31 copy nothing
            L.bytes4 = true;
        }
        const S *gb = in + (size_t)(gpair >> 2) * line_stride + (gpair & 3);
        L.p0 = (const char *)(gb + first);
        L.q0 = (const char *)(gb + last);
    }
    const long long deb = de * (long long)sizeof(S);  // byte stride of a step
    double *ck_own = ckpt + (size_t)(ANTI ? 1 : 0) * J * 4 * npairs + pair;
    const double *ck_other = ckpt + (size_t)(ANTI ? 0 : 1) * J * 4 * npairs + pair;

    Chain<S> st;
    chain_clear(st);
    double initial = 0.0, initial_other = 0.0;
    if (n > 0) { initial = conv(*P0, lut); initial_other = conv(*Q0, lut); }

    // ---- phase 1: own steps [0, len1), checkpoints only; tile j = own steps [j*CK_S, ...) in ring slot j % CK_R
#pragma unroll
    for (int j = 0; j < CK_R - 1; ++j) {
        const int c = len1 - j * CK_S;
        tile_issue(&sm.ring[j % CK_R][0][0], L, L.p0 + (long long)j * CK_S * deb, deb, c < CK_S ? (c < 0 ? 0 : c) : CK_S);
    }
    int slot = 0;  // j % CK_R
#pragma unroll 1
    for (int j = 0; j < J1; ++j) {
        __syncwarp();  // the slot refilled below was read by other lanes in the previous round
        {
            const int jn = j + CK_R - 1, c = len1 - jn * CK_S;
            tile_issue(&sm.ring[slot == 0 ? CK_R - 1 : slot - 1][0][0], L, L.p0 + (long long)jn * CK_S * deb, deb, c < CK_S ? (c < 0 ? 0 : c) : CK_S);
        }
        cp_async_wait<CK_R - 1>();
        __syncwarp();
        const S *tile = &sm.ring[slot][0][tid];
        const int cnt = len1 - j * CK_S < CK_S ? len1 - j * CK_S : CK_S;
        if (j == 0) {
            int i = 0;
#pragma unroll 1
            for (; i < cnt && i < 4; ++i) step_boundary<ANTI>(C, st, tile[i * CK_T], i, initial, lut);
#pragma unroll 1
            for (; i < cnt; ++i) step_fast<ANTI>(C, st, tile[i * CK_T], lut);
        } else {
            double *c = ck_own + (size_t)j * 4 * npairs;
            c[0] = st.v1; c[(size_t)npairs] = st.v2; c[(size_t)2 * npairs] = st.v3; c[(size_t)3 * npairs] = st.v4;
            if (cnt == CK_S) plain_round<ANTI>(C, st, tile, lut);
            else {
#pragma unroll 1
                for (int i = 0; i < cnt; ++i) step_fast<ANTI>(C, st, tile[i * CK_T], lut);
            }
        }
        slot = slot == CK_R - 1 ? 0 : slot + 1;
    }
    cp_async_wait<0>();
    __syncthreads();  // the pair's checkpoints are in memory; every ring slot is free

    // ---- phase 2: the other sweep's blocks jb = J2-1 .. 0 (its steps [jb*CK_S, ...)); tile jb holds the samples of those
    // steps in ITS order (element i = its step jb*CK_S + i), which the re-run reads ascending and the own sweep descending.
    // w holds the other sweep's values of one block (element i in slot i): a round runs the own sweep over block jb,
    // adding w from the last slot down, then re-runs the other sweep over block jb - 1 into w.  The two loops are kept
    // apart on purpose: interleaved in one loop (two dependent chains per thread) they measured the same or slower --
    // the kernel is bound by instruction issue (two cycles per warp instruction on average), not by FP64 latency.
    if (len2 <= 0) return;
    double *w = &sm.w[0][tid];
    auto load_checkpoint = [&](int jb, double (&v)[4], S (&hs)[4]) {  // state of the other sweep before its step jb * CK_S (jb >= 1)
        const double *c = ck_other + (size_t)jb * 4 * npairs;
        v[0] = c[0]; v[1] = c[(size_t)npairs]; v[2] = c[(size_t)2 * npairs]; v[3] = c[(size_t)3 * npairs];
        const S *q = Q0 - (long long)(jb * CK_S - 1) * de;  // its step jb*CK_S - 1
        hs[0] = q[0]; hs[1] = q[de]; hs[2] = q[2 * de]; hs[3] = q[3 * de];
    };
    // other's block jb (cnt steps from its step jb*CK_S) re-run from state r
    auto rerun_block = [&](Chain<S> &r, const S *tile, int jb, int cnt) {
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) {
            const int u = jb * CK_S + i;
            w[i * CK_T] = u < 4 ? step_boundary<!ANTI>(C, r, tile[i * CK_T], u, initial_other, lut) : step_fast<!ANTI>(C, r, tile[i * CK_T], lut);
        }
    };
    // tile jb lives in ring slot jb % CK_R
    int oslot = (J2 - 1) % CK_R;  // slot of the own block's tile
    auto slot_below = [](int s, int k) { s -= k; return s < 0 ? s + CK_R : s; };
#pragma unroll
    for (int k = 0; k < CK_R - 1; ++k) {
        const int jb = J2 - 1 - k;
        const int c = jb >= 0 ? (len2 - jb * CK_S < CK_S ? len2 - jb * CK_S : CK_S) : 0;
        tile_issue(&sm.ring[slot_below(oslot, k)][0][0], L, L.q0 - (long long)jb * CK_S * deb, -deb, c);
    }
    double ckv[4] = {0, 0, 0, 0};
    S hs[4] = {S(0), S(0), S(0), S(0)};
    {
        Chain<S> r;
        chain_clear(r);
        if (J2 - 1 >= 1) {
            load_checkpoint(J2 - 1, ckv, hs);
            chain_restore(r, ckv, hs, lut);
        }
        cp_async_wait<CK_R - 2>();
        __syncwarp();
        rerun_block(r, &sm.ring[oslot][0][tid], J2 - 1, len2 - (J2 - 1) * CK_S);
        if (J2 - 2 >= 1) load_checkpoint(J2 - 2, ckv, hs);
    }
#pragma unroll 1
    for (int jb = J2 - 1; jb >= 0; --jb) {
        const int cnt = len2 - jb * CK_S < CK_S ? len2 - jb * CK_S : CK_S;
        const int t0 = n - (jb * CK_S + cnt);  // own steps t0 .. t0+cnt-1 cover the other sweep's block jb
        __syncwarp();
        {
            const int jn = jb - (CK_R - 1);
            tile_issue(&sm.ring[slot_below(oslot, CK_R - 1)][0][0], L, L.q0 - (long long)jn * CK_S * deb, -deb, jn >= 0 ? CK_S : 0);
        }
        double nckv[4] = {0, 0, 0, 0};
        S nhs[4] = {S(0), S(0), S(0), S(0)};
        if (jb - 2 >= 1) load_checkpoint(jb - 2, nckv, nhs);  // used in the next round
        cp_async_wait<CK_R - 2>();  // all but the latest CK_R - 2 tiles have landed: tiles jb and jb - 1
        __syncwarp();
        const S *otile = &sm.ring[oslot][0][tid];
        const S *rtile = &sm.ring[slot_below(oslot, 1)][0][tid];
        OutCursor<OUT> oc;
        if constexpr (OUT == OUT_FLOAT) {
            oc.p = (float *)out + (size_t)line * out_line_stride + ch + first + (long long)t0 * de;
            oc.de = de;
        } else {
            oc.p = (unsigned char *)out + (size_t)line * out_line_stride + (size_t)(ANTI ? n - 1 - t0 : t0) * 4 + ch;
            oc.dir = ANTI ? -4 : 4;
        }
        Chain<S> r;
        chain_clear(r);
        if (jb - 1 >= 1) chain_restore(r, ckv, hs, lut);
        if (cnt == CK_S && t0 >= 4) own_round<ANTI>(C, st, otile, w, oc, lut);
        else {
#pragma unroll 1
            for (int q = 0; q < cnt; ++q) {
                const int t = t0 + q, i = cnt - 1 - q;
                const double a = t < 4 ? step_boundary<ANTI>(C, st, otile[i * CK_T], t, initial, lut) : step_fast<ANTI>(C, st, otile[i * CK_T], lut);
                oc.put((float)__dadd_rn(a, w[i * CK_T]));
            }
        }
        if (jb - 1 >= 1) rerun_round<!ANTI>(C, r, rtile, w, lut);
        else if (jb >= 1) rerun_block(r, rtile, 0, CK_S);
#pragma unroll
        for (int k = 0; k < 4; ++k) { ckv[k] = nckv[k]; hs[k] = nhs[k]; }
        oslot = slot_below(oslot, 1);
    }
    cp_async_wait<0>();
}

template <class S, int OUT>
__global__ void __launch_bounds__(CK_T, 7) gauss_iir_ckpt_kernel(const S *in, void *out, long long out_line_stride, double *ckpt, int nlines, int n,
                                                                  long long line_stride, long long elem_stride, int J, int bytes16, const __grid_constant__ Coeffs C) {
    __shared__ double sh_w[CK_S][CK_T];
    __shared__ S sh_ring[CK_R][CK_S][CK_T];
    __shared__ double sh_lut[sizeof(S) == 1 ? 256 : 1];
    if (sizeof(S) == 1) {
        for (int k = threadIdx.x; k < 256; k += CK_T) sh_lut[k] = (double)unit_from_byte((unsigned)k);
        __syncthreads();
    }
    const CkShared<S> sm = {sh_w, sh_ring, sh_lut};
    const bool anti = threadIdx.x >= CK_PAIRS;
    const int npairs = nlines * 4;
    const int pair = blockIdx.x * CK_PAIRS + (threadIdx.x - (anti ? CK_PAIRS : 0));
    if (anti) ck_thread<true, OUT>(C, in, out, out_line_stride, ckpt, J, pair, npairs, n, line_stride, elem_stride, bytes16 != 0, sm);
    else ck_thread<false, OUT>(C, in, out, out_line_stride, ckpt, J, pair, npairs, n, line_stride, elem_stride, bytes16 != 0, sm);
}

int ckpt_blocks(int n) { return ((n - n / 2) + CK_S - 1) / CK_S + 1; }

template <class S, int OUT>
void launch_pass(const S *in, void *out, long long out_line_stride, double *ckpt, int nlines, int n, long long line_stride, long long elem_stride,
                 float sigma, cudaStream_t stream) {
    if (nlines <= 0 || n <= 0) return;
    Coeffs c;
    double raw[30];
    gauss_iir_constants_host(sigma, raw);
    memcpy(&c, raw, sizeof c);
    const int npairs = nlines * 4;
    // bytes: 16-byte chunks need rows of the picture that start and end on 16-byte boundaries
    const int bytes16 = sizeof(S) == 1 && npairs % 16 == 0 && ((uintptr_t)in & 15) == 0 && (elem_stride & 15) == 0;
    gauss_iir_ckpt_kernel<S, OUT><<<(npairs + CK_PAIRS - 1) / CK_PAIRS, CK_T, 0, stream>>>(in, out, out_line_stride, ckpt, nlines, n, line_stride,
                                                                                          elem_stride, ckpt_blocks(n), bytes16, c);
}

}  // namespace

// doubles of checkpoint memory for a width x height blur (the larger of the two passes)
size_t gauss_iir_scratch_bytes(int width, int height) {
    const size_t col = (size_t)2 * ckpt_blocks(height) * 4 * (size_t)width * 4, row = (size_t)2 * ckpt_blocks(width) * 4 * (size_t)height * 4;
    return sizeof(double) * (col > row ? col : row) + 16;
}

// The vertical pass: in -> mid (float4 [height][width]; may alias a float input).  With in_is_rgba8 the input is
// uchar4 [height][width] whose bytes stand for k/255.  Lines are columns: consecutive recursions are consecutive floats /
// bytes of a row.
void launch_gauss_iir_columns(const void *in, bool in_is_rgba8, float *mid, double *scratch, int width, int height, float sigma_v, cudaStream_t stream) {
    if (in_is_rgba8) launch_pass<unsigned char, OUT_FLOAT>((const unsigned char *)in, mid, 4, scratch, width, height, 4, (long long)width * 4, sigma_v, stream);
    else launch_pass<float, OUT_FLOAT>((const float *)in, mid, 4, scratch, width, height, 4, (long long)width * 4, sigma_v, stream);
}
// The horizontal pass over rows [first_row, first_row + nrows) of mid (float4 [..][width]): row r goes to
// out + (r - first_row) * out_pitch bytes, as float4 pixels or (rgba8) as the quantised RGBA8 pixels the pixel kernel would
// store.  out may be mid's own rows (in place, float).
void launch_gauss_iir_rows(const float *mid, void *out, bool rgba8, long long out_pitch, double *scratch, int width, int first_row, int nrows, float sigma_h,
                           cudaStream_t stream) {
    const float *in = mid + (size_t)first_row * width * 4;
    if (rgba8) launch_pass<float, OUT_RGBA8>(in, out, out_pitch, scratch, nrows, width, (long long)width * 4, 4, sigma_h, stream);
    else launch_pass<float, OUT_FLOAT>(in, out, out_pitch / (long long)sizeof(float), scratch, nrows, width, (long long)width * 4, 4, sigma_h, stream);
}
// both passes, in -> out (float4 [height][width]; may alias); scratch: gauss_iir_scratch_bytes()
void launch_gauss_iir(const void *in, bool in_is_rgba8, float *out, double *scratch, int width, int height, float sigma_h, float sigma_v,
                      cudaStream_t stream) {
    launch_gauss_iir_columns(in, in_is_rgba8, out, scratch, width, height, sigma_v, stream);
    launch_gauss_iir_rows(out, out, false, (long long)width * 16, scratch, width, 0, height, sigma_h, stream);
}

}  // namespace mmbackend
