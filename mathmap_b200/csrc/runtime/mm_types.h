// Plain structs shared by host code and device code (NVRTC and nvcc).
#pragma once

#define MM_MAX_IMAGES 32
#define MM_CURVE_POINTS 1024

typedef unsigned int mm_color;

enum { MM_IMAGE_DRAWABLE = 0, MM_IMAGE_FLOATMAP = 1, MM_IMAGE_CLOSURE = 2 };

// One input or intermediate image.  Drawables are RGBA8 (R first); floatmaps are
// float4 per pixel (reference drawable.h:92-125, floatmap.c:30-47).
struct mm_image {
    const void *data;
    int kind;
    int w, h;
    int num_frames;
    float sx, sy, mx, my;  // drawable: scale and middle (userval.c:263-279)
    float ax, bx, ay, by;  // floatmap: pixel = a * coord + b
    float xf, yf;          // resize factors applied to sampling coordinates (opmacros.h:203-207)
    // Bounds of the samplers' interior fast paths as floats, filled by the host: w, h, w-1, h-1 and num_frames
    // when w, h < 2^22 and the image is a drawable, else all -1 (the fast-path tests then never hold).
    float fast_w, fast_h, fast_wm1, fast_hm1, fast_nf;
    // data - 0x4B000000 texels: the fast paths index texels with the raw bits of 2^23 + x1 (a float whose mantissa holds
    // the column), so the offset of the exponent bits is taken out of the base once, by the host
    const unsigned *fast_base;
    int closure_filter;    // MM_IMAGE_CLOSURE: which filter; `data` then points at that filter's packed uniforms (device memory)
};

#ifndef MM_MAX_CALL_DEPTH
#define MM_MAX_CALL_DEPTH 64 /* nesting levels of filter calls on the device, see mm_runtime.cuh */
#endif

// Per-launch parameters (the invocation / frame / slice fields calc_lines reads,
// reference mathmap.h:161-227, new_template.c.in:210-234).
struct mm_params {
    void *out;                // first row of the band
    long long out_stride;     // bytes per output row
    const float *xs;          // virtual x per absolute column (CALC_VIRTUAL_X evaluated on the host in double)
    const float *ys;          // virtual y per absolute row
    int region_x, region_y, region_w;
    int first_row, num_rows;  // band, in absolute rows (num_rows counts rendered rows)
    int row_limit;            // absolute row bound (exclusive)
    int row_interleave, row_phase;  // > 1: render only 8-row blocks b with b % row_interleave == row_phase, stored compactly
    int rows;                 // 32x8 tiles a block renders in sequence (grid.y = ceil(num_rows / (8 * rows)))
    int img_w, img_h, render_w, render_h;
    int frame;
    float t;
    float R;
    int bpp;
    int floatmap;
    int out_mode;             // 0: RGBA8 (bpp 4), 1: floatmap, 2: bpp 1..3 -- one test in the store instead of two
    unsigned magic23;         // 0x4B000000 (2^23 as float bits); passed as data so that byte -> float PRMTs keep it in a register
    mm_color edge_color_x, edge_color_y;
    // (-0, -0), (1, 1) and (1.5 * 2^23, 1.5 * 2^23) as packed float pairs, passed as data: see mm_fma2 in mm_runtime.cuh
    unsigned long long pk_neg_zero, pk_one, pk_magic_round;
    int vec_store;            // quad kernels: `out` and `out_stride` are multiples of 16 bytes, four RGBA8 pixels go out as one 128-bit store
    mm_image images[MM_MAX_IMAGES];
};

