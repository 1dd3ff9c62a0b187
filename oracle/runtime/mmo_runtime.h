/* mathmap_b200 ORACLE host runtime — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the CPU runtime the reference's generated code links
 * against.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build or call anything under oracle/.
 *
 * Pinned against the reference's golden PNGs by tests/test_oracle_goldens.py.
 *
 * What each part follows (paths into the reference tree):
 *   op semantics ...................... opmacros.h:30-216
 *   MIN/MAX/colour packing ............ new_template.c.in:51-84
 *   coordinate mapping ................ opmacros.h:156-157
 *   input sampling, edge modes ........ builtins/builtins.c:41-265, color.h:36-54
 *   drawable scale/middle ............. userval.c:263-279
 *   floatmaps ......................... floatmap.c:30-47, builtins/builtins.c:249-265
 *   render_image ...................... builtins/builtins.c:269-345
 *   row/column loops, quantisation .... new_template.c.in:208-312
 *   supersampling ..................... mathmap_common.c:880-927
 */
#ifndef MMO_RUNTIME_H
#define MMO_RUNTIME_H

#include <complex.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#define MAX(a, b) (((a) < (b)) ? (b) : (a))

typedef unsigned int color_t;
#define MAKE_RGBA_COLOR_UNSAFE(r, g, b, a) ((((color_t)(r)) << 24) | (((color_t)(g)) << 16) | (((color_t)(b)) << 8) | ((color_t)(a)))
#define MAKE_RGBA_COLOR(r, g, b, a) MAKE_RGBA_COLOR_UNSAFE((color_t)(r)&0xff, (color_t)(g)&0xff, (color_t)(b)&0xff, (color_t)(a)&0xff)
#define RED(c) ((c) >> 24)
#define GREEN(c) (((c) >> 16) & 0xff)
#define BLUE(c) (((c) >> 8) & 0xff)
#define ALPHA(c) ((c)&0xff)
#define RED_FLOAT(c) (RED(c) / 255.0)
#define GREEN_FLOAT(c) (GREEN(c) / 255.0)
#define BLUE_FLOAT(c) (BLUE(c) / 255.0)
#define ALPHA_FLOAT(c) (ALPHA(c) / 255.0)

#define USER_CURVE_POINTS 1024
#define NUM_FLOATMAP_CHANNELS 4

enum { MMO_IMAGE_DRAWABLE = 0, MMO_IMAGE_FLOATMAP = 1, MMO_IMAGE_CLOSURE = 2, MMO_IMAGE_RESIZE = 3 };
enum { EDGE_BEHAVIOUR_COLOR = 0, EDGE_BEHAVIOUR_WRAP = 1, EDGE_BEHAVIOUR_REFLECT = 2, EDGE_BEHAVIOUR_ROTATE = 3 };

struct mmo_image;
struct mmo_invocation;
struct mmo_pools;

typedef struct mmo_userval {
    union {
        int int_const;
        float float_const;
        int bool_const;
        color_t color;
        const float *curve;       /* USER_CURVE_POINTS samples */
        const color_t *gradient;  /* USER_CURVE_POINTS samples */
        struct mmo_image *image;
    } v;
} mmo_userval;

typedef float *(*mmo_filter_func)(struct mmo_invocation *, struct mmo_image *closure, float x, float y, float t, struct mmo_pools *pools);
typedef void (*mmo_calc_lines_func)(struct mmo_invocation *, struct mmo_image *closure, int frame, float t, int frame_w, int frame_h,
                                    int region_x, int region_y, int region_w, int region_h, float off_x, float off_y,
                                    int first_row, int last_row, void *q, int floatmap);

#define MMO_MAX_ARGS 96 /* arguments of one filter or closure (a composition of Droste + two more nodes has 37) */
typedef struct mmo_image {
    int type;
    int pixel_width, pixel_height;
    /* drawable: RGBA8 rows, R first */
    const unsigned char *data;
    float scale_x, scale_y, middle_x, middle_y;
    int num_frames;
    /* floatmap */
    float ax, bx, ay, by;
    float *fdata;
    /* closure */
    mmo_filter_func func;
    mmo_calc_lines_func calc_lines;
    void *xy_vars;
    int num_args;
    mmo_userval args[MMO_MAX_ARGS];
    /* resize */
    struct mmo_image *original;
    float x_factor, y_factor;
} mmo_image;

typedef struct mmo_invocation {
    int img_width, img_height;
    int render_width, render_height;
    float image_R;
    int antialiasing;
    int supersampling;
    int edge_behaviour_x, edge_behaviour_y;
    color_t edge_color_x, edge_color_y;
    int output_bpp;
    int row_stride;
    /* native filter cache (native-filters/cache.c): one entry list, single-threaded init_frame */
    struct mmo_cache_entry *cache;
    long taps; /* statistics: number of ORIG_VAL samples taken on drawables (for bench accounting) */
} mmo_invocation;

/* ---- pools: bump allocator reset per pixel (reference mmpools.c / lispreader/pools.c) ---- */
typedef struct mmo_chunk {
    struct mmo_chunk *next;
    size_t used, size;
    char *mem;
} mmo_chunk;
typedef struct mmo_pools {
    mmo_chunk *first, *cur;
} mmo_pools;

static inline void mmo_pools_init(mmo_pools *p) { p->first = p->cur = NULL; }
static inline void *mmo_pools_alloc(mmo_pools *p, size_t n) {
    n = (n + 15) & ~(size_t)15;
    while (1) {
        if (p->cur && p->cur->used + n <= p->cur->size) {
            void *r = p->cur->mem + p->cur->used;
            p->cur->used += n;
            memset(r, 0, n);
            return r;
        }
        if (p->cur && p->cur->next) {
            p->cur = p->cur->next;
            p->cur->used = 0;
            continue;
        }
        {
            size_t sz = n > 65536 ? n : 65536;
            mmo_chunk *c = (mmo_chunk *)malloc(sizeof(mmo_chunk));
            c->mem = (char *)malloc(sz);
            c->size = sz;
            c->used = 0;
            c->next = NULL;
            if (p->cur) p->cur->next = c; else p->first = c;
            p->cur = c;
        }
    }
}
static inline void mmo_pools_reset(mmo_pools *p) {
    p->cur = p->first;
    if (p->cur) p->cur->used = 0;
}
static inline void mmo_pools_free(mmo_pools *p) {
    mmo_chunk *c = p->first;
    while (c) { mmo_chunk *n = c->next; free(c->mem); free(c); c = n; }
    p->first = p->cur = NULL;
}

/* ---- op macros (opmacros.h:30-47, 127-128, 154-157) ---- */
#define NOP() (0.0)
#define INT2FLOAT(x) ((float)(x))
#define FLOAT2INT(x) ((int)(x))
#define INT2COMPLEX(x) ((float _Complex)(x))
#define FLOAT2COMPLEX(x) ((float _Complex)(x))
#define ADD(a, b) ((a) + (b))
#define SUB(a, b) ((a) - (b))
#define NEG(a) (-(a))
#define MUL(a, b) ((a) * (b))
#define DIV(a, b) ((float)(a) / (float)(b))
#define MOD(a, b) (fmod((a), (b)))
#define EQ(a, b) ((a) == (b))
#define LESS(a, b) ((a) < (b))
#define LEQ(a, b) ((a) <= (b))
#define NOT(a) (!(a))
#define PRINT_FLOAT(a) (0)
#define NEWLINE() (0)
#define COMPLEX(r, i) ((r) + (i)*I)
#define CLAMP01(x) (MAX(0, MIN(1, (x))))
#define MAKE_COLOR(r, g, b, a) (MAKE_RGBA_COLOR(CLAMP01((r)) * 255, CLAMP01((g)) * 255, CLAMP01((b)) * 255, CLAMP01((a)) * 255))
#define CALC_VIRTUAL_X(pxl, size, sampl_off) (((pxl) - ((size)-1) / 2.0 + (sampl_off)) / (((size)-1) / 2.0))
#define CALC_VIRTUAL_Y(pxl, size, sampl_off) ((-(pxl) + ((size)-1) / 2.0 - (sampl_off)) / (((size)-1) / 2.0))
#define IMAGE_PIXEL_WIDTH(i) ((i)->pixel_width)
#define IMAGE_PIXEL_HEIGHT(i) ((i)->pixel_height)
#define ALLOC_TUPLE(n) ((float *)mmo_pools_alloc(pools, sizeof(float) * (n)))
#define TUPLE_NTH(t, n) ((t)[(n)])
#define TUPLE_RED(t) CLAMP01(TUPLE_NTH((t), 0))
#define TUPLE_GREEN(t) CLAMP01(TUPLE_NTH((t), 1))
#define TUPLE_BLUE(t) CLAMP01(TUPLE_NTH((t), 2))
#define TUPLE_ALPHA(t) CLAMP01(TUPLE_NTH((t), 3))
#define APPLY_CURVE(c, p) ((c)[(int)(CLAMP01((p)) * (USER_CURVE_POINTS - 1))])
#define STRIP_RESIZE(i) ((i)->type == MMO_IMAGE_RESIZE ? (i)->original : (i))
#define UNINITED_IMAGE ((mmo_image *)0)

/* special functions: GSL is not available here, so these are declared by the
 * oracle's own restatements (spec_funcs.c); parity for them is UNPINNED
 * (no reference test exercises them). */
double mmo_gamma(double x);
double mmo_beta(double a, double b);
float _Complex mmo_cgamma(float _Complex z);
#define GAMMA(a) (((a) > 171.0) ? 0.0 : mmo_gamma((a)))
#define gsl_sf_beta(a, b) mmo_beta((a), (b))
/* elliptics (opmacros.h:101-125), see elliptic.c */
double mmo_ellint_Kcomp(double k);
double mmo_ellint_Ecomp(double k);
double mmo_ellint_F(double phi, double k);
double mmo_ellint_E(double phi, double k);
double mmo_ellint_P(double phi, double k, double n);
double mmo_ellint_D(double phi, double k);
double mmo_ellint_RC(double x, double y);
double mmo_ellint_RD(double x, double y, double z);
double mmo_ellint_RF(double x, double y, double z);
double mmo_ellint_RJ(double x, double y, double z, double p);
void mmo_elljac(double u, double m, double *sn, double *cn, double *dn);
float *mmo_ell_jac_tuple(float u, float m, struct mmo_pools *pools);
#define ELL_INT_K_COMP(k) mmo_ellint_Kcomp((k))
#define ELL_INT_E_COMP(k) mmo_ellint_Ecomp((k))
#define ELL_INT_F(phi, k) mmo_ellint_F((phi), (k))
#define ELL_INT_E(phi, k) mmo_ellint_E((phi), (k))
#define ELL_INT_P(phi, k, n) mmo_ellint_P((phi), (k), (n))
#define ELL_INT_D(phi, k, n) mmo_ellint_D((phi), (k))
#define ELL_INT_RC(x, y) mmo_ellint_RC((x), (y))
#define ELL_INT_RD(x, y, z) mmo_ellint_RD((x), (y), (z))
#define ELL_INT_RF(x, y, z) mmo_ellint_RF((x), (y), (z))
#define ELL_INT_RJ(x, y, z, p) mmo_ellint_RJ((x), (y), (z), (p))
#define ELL_JAC(u, m) mmo_ell_jac_tuple((u), (m), pools)
#define cgamma(z) mmo_cgamma((z))

/* GSL's gsl_linalg_HH_solve is absent: Cramer's rule in double (spec_funcs.c), parity unpinned */
float *mmo_solve_linear_2(const float *m, const float *v, struct mmo_pools *pools);
float *mmo_solve_linear_3(const float *m, const float *v, struct mmo_pools *pools);

/* libnoise (noise.c) */
float libnoise_perlin(int octaves, float persistence, float lacunarity, float x, float y, float z);
float libnoise_billow(int octaves, float persistence, float lacunarity, float x, float y, float z);
float libnoise_ridged_multi(int octaves, float lacunarity, float x, float y, float z);
float libnoise_voronoi(float displacement, float x, float y, float z);

/* native filters (gauss.c) */
mmo_image *native_filter_gaussian_blur(mmo_invocation *invocation, mmo_userval *args, mmo_pools *pools);

/* FFT natives (convolve.c) */
mmo_image *native_filter_convolve(mmo_invocation *invocation, mmo_userval *args, mmo_pools *pools);
mmo_image *native_filter_half_convolve(mmo_invocation *invocation, mmo_userval *args, mmo_pools *pools);
mmo_image *native_filter_visualize_fft(mmo_invocation *invocation, mmo_userval *args, mmo_pools *pools);

/* images.c */
mmo_image *mmo_floatmap_alloc(int width, int height, mmo_pools *pools);
mmo_image *mmo_floatmap_copy(mmo_image *src, mmo_pools *pools);
mmo_image *mmo_make_resize_image(mmo_image *image, float x_factor, float y_factor, mmo_pools *pools);
mmo_image *mmo_render_image(mmo_invocation *invocation, mmo_image *image, int width, int height, mmo_pools *pools, int force);
color_t mmo_get_orig_val_pixel(mmo_invocation *invocation, float x, float y, mmo_image *image, int frame);
color_t mmo_get_orig_val_intersample_pixel(mmo_invocation *invocation, float x, float y, mmo_image *image, int frame);
float *mmo_get_floatmap_pixel(mmo_invocation *invocation, mmo_image *image, float x, float y, float frame);
float *mmo_orig_val(mmo_invocation *invocation, float x, float y, mmo_image *img, float frame, mmo_pools *pools, int force_nearest);

#define RESIZE_IMAGE(i, xf, yf) (mmo_make_resize_image((i), (xf), (yf), pools))
#define ORIG_VAL(ix, iy, i, f) (mmo_orig_val(invocation, (ix), (iy), (i), (f), pools, 0))
#define RENDER(i, w, h) (mmo_render_image(invocation, (i), (w), (h), pools, 0))

static inline float *mmo_tuple_from_color(color_t c, mmo_pools *pools) {
    float *tuple = ALLOC_TUPLE(4);
    tuple[0] = RED_FLOAT(c);
    tuple[1] = GREEN_FLOAT(c);
    tuple[2] = BLUE_FLOAT(c);
    tuple[3] = ALPHA_FLOAT(c);
    return tuple;
}
#define APPLY_GRADIENT(g, p) (mmo_tuple_from_color((g)[(int)(CLAMP01((p)) * (USER_CURVE_POINTS - 1))], pools))

/* quantisation of one pixel (new_template.c.in:272-293) */
static inline void mmo_store_pixel(unsigned char *p, const float *return_tuple, int output_bpp) {
    int is_bw = output_bpp == 1 || output_bpp == 2;
    int need_alpha = output_bpp == 2 || output_bpp == 4;
    int alpha_index = output_bpp - 1;
    if (is_bw)
        p[0] = (TUPLE_RED(return_tuple) * 0.299 + TUPLE_GREEN(return_tuple) * 0.587 + TUPLE_BLUE(return_tuple) * 0.114) * 255.0;
    else {
        p[0] = TUPLE_RED(return_tuple) * 255.0;
        p[1] = TUPLE_GREEN(return_tuple) * 255.0;
        p[2] = TUPLE_BLUE(return_tuple) * 255.0;
    }
    if (need_alpha) p[alpha_index] = TUPLE_ALPHA(return_tuple) * 255.0;
}

#endif
