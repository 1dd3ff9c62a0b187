/* mathmap_b200 — C ABI of the B200-native evaluation backend for MathMap's
 * per-pixel render path.
 *
 * This header is the drop-in boundary: plain C, pointers and sizes only.  Each
 * entry point names the reference interface it replaces (paths into the
 * reference tree).  INTEGRATION.md shows the backends/cuda.c a maintainer adds
 * to the reference to bind these.
 *
 * Everything that renders requires a CUDA device and fails loudly (returns
 * nonzero / NULL and sets mmb_last_error) without one.  There is no CPU path.
 */
#ifndef MATHMAP_B200_H
#define MATHMAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmb_module mmb_module;         /* a compiled filter module; reference: mathmap_t after compile_mathmap */
typedef struct mmb_invocation mmb_invocation; /* reference: mathmap_invocation_t */

/* userval kinds, reference userval.h USERVAL_* */
enum { MMB_USERVAL_INT = 0, MMB_USERVAL_FLOAT = 1, MMB_USERVAL_BOOL = 2, MMB_USERVAL_COLOR = 3,
       MMB_USERVAL_CURVE = 4, MMB_USERVAL_GRADIENT = 5, MMB_USERVAL_IMAGE = 6 };
/* edge behaviour, reference mathmap.h EDGE_BEHAVIOUR_* */
enum { MMB_EDGE_COLOR = 0, MMB_EDGE_WRAP = 1, MMB_EDGE_REFLECT = 2, MMB_EDGE_ROTATE = 3 };
#define MMB_CURVE_POINTS 1024 /* reference userval.h USER_CURVE_POINTS */

/* ---- compile ---------------------------------------------------------------
 * mmb_compile: MathMap source -> optimised IR -> CUDA C.  Replaces
 *   compile_mathmap (mathmap_common.c:504-580) = parse_mathmap +
 *   compiler_compile_filters + gen_and_load_c_code.
 * mmb_load_ir: the same from IR text ("mmir 1", see csrc/ir/ir_text.cpp) that a
 *   reference-side backends/cuda.c prints from filter_code_t**.  Replaces
 *   gen_and_load_c_code (backends/cc.c:634-758, declared compiler.h:79-82).
 * Both run the loop-carried value pass (csrc/ir/passes.cpp; environment MMB_LOOP_CARRY=0 switches it off), so the
 *   reference compiler's IR ends in the same kernels as mmb_compile's.
 * Both return NULL on error with the message in mmb_last_error().  NVRTC
 * compilation for sm_100a happens lazily at the first render of a given
 * sampler/edge/output configuration.
 */
mmb_module *mmb_compile(const char *source);
mmb_module *mmb_load_ir(const char *ir_text);
/* Compositions (".mmc" designs, the composer's save format, designer/loadsave.c:30 designer_load_design): the MathMap
 * source the reference generates for a design (designer_filter.c:278 make_filter_source_from_design); node types are
 * looked up by the name of the main filter of the .mm / .mmc files found under filter_search_path
 * (expression_db.c:155 read_expression_db).  The string is malloc'd: release it with mmb_free_string. */
char *mmb_design_to_source(const char *design_text, const char *filter_search_path);
void mmb_free_string(char *s);
mmb_module *mmb_compile_design(const char *design_text, const char *filter_search_path);
void mmb_module_free(mmb_module *m); /* unload_c_code, backends/cc.c:761 */
const char *mmb_module_ir(const mmb_module *m);          /* optimised IR as text */
const char *mmb_module_cuda_source(mmb_module *m);       /* generated CUDA C for the default configuration */
const char *mmb_module_main_filter_name(const mmb_module *m);
int mmb_module_num_uservals(const mmb_module *m);        /* uservals of the main filter, userval_info_t list */
int mmb_module_userval_info(const mmb_module *m, int index, char *name, size_t name_len, int *type, float *min_value,
                            float *max_value, float *default_value);
int mmb_module_userval_index(const mmb_module *m, const char *name);

/* ---- invoke ----------------------------------------------------------------
 * mmb_invoke replaces invoke_mathmap (mathmap_common.c:747-790): output 4 bytes
 * per pixel, edge behaviour COLOR with colour 0, nearest sampling, no
 * supersampling, uservals at their defaults, R = sqrt(2).  `device` is the CUDA
 * device ordinal.
 */
mmb_invocation *mmb_invoke(mmb_module *m, int img_width, int img_height, int device);
void mmb_invocation_free(mmb_invocation *inv);
int mmb_set_antialiasing(mmb_invocation *inv, int enabled);   /* invocation_set_antialiasing, mathmap_common.c:737 (-i flag) */
int mmb_set_supersampling(mmb_invocation *inv, int enabled);  /* invocation->supersampling (-o flag) */
int mmb_set_edge_behaviour(mmb_invocation *inv, int mode_x, int mode_y, uint32_t color_x, uint32_t color_y);
int mmb_set_output_bpp(mmb_invocation *inv, int bpp);         /* invocation->output_bpp: 1, 2, 3 or 4 */
/* invocation->render_width / render_height (mathmap.h:180-181) when they differ from the image size: the GIMP preview
 * renders a scaled frame (mathmap.c:2191-2223).  __renderPixelW/H read them and native filters render their
 * intermediates at this size (native-filters/gauss.c:657, convolve.c:88).  Default: the image size. */
int mmb_set_render_size(mmb_invocation *inv, int render_width, int render_height);
int mmb_set_warp_shape(mmb_invocation *inv, int warp_width); /* pixels per warp row in the 32x8 tile: 32 (default), 16 or 8 */
int mmb_set_rows_per_thread(mmb_invocation *inv, int rows); /* 32x8 tiles one block renders in sequence (1, 2, 4 or 8; 0 = automatic, the default: up to 8 for straight-line pixel code on large grids, 1 for per-pixel loops, which ignore the setting) */
/* 1 (default): the pixel kernel of a frame is compiled for the values of the frame-constant conditions it branches on
 * (`if (userval)`: the reference's init_frame values, new_template.c.in:314-337, as compile-time constants), one NVRTC
 * compile per combination actually used; 0: one kernel that tests them per pixel */
int mmb_set_specialize(mmb_invocation *inv, int enabled);
/* 0 (default): every device function inlined into the pixel kernel; 1: the complex elementary functions (cexpf, clogf,
 * cpowf, csinf ...: the largest bodies of the device runtime) are real calls, compiled once per module instead of once per
 * call site -- Map/Droste: NVRTC 3.7 -> 2.1 s, kernel 4-5 % slower, same bits.  For callers that show the first frame
 * of a newly edited filter (the reference's only published timings are compile times: TODO:715-720); kernels of both kinds
 * are cached side by side */
int mmb_set_fast_compile(mmb_invocation *inv, int enabled);
int mmb_set_precise_math(mmb_invocation *inv, int enabled);   /* 1 (default): libm calls evaluated in double and narrowed, like the host; 0: CUDA float libm (<= 2 ulp, faster) */

/* userval bindings, reference userval.h userval_t / mathmap_cmdline.c:756-796 (-D name=value) */
int mmb_set_userval_int(mmb_invocation *inv, int index, int value);
int mmb_set_userval_float(mmb_invocation *inv, int index, float value);
int mmb_set_userval_bool(mmb_invocation *inv, int index, int value);
int mmb_set_userval_color(mmb_invocation *inv, int index, float r, float g, float b, float a);
int mmb_set_userval_color_packed(mmb_invocation *inv, int index, uint32_t rgba_packed); /* userval_t.v.color.value as it is: R in the high byte (color.h:36-43) */
int mmb_set_userval_curve(mmb_invocation *inv, int index, const float *values /* MMB_CURVE_POINTS */);
int mmb_set_userval_gradient(mmb_invocation *inv, int index, const uint32_t *rgba_packed /* MMB_CURVE_POINTS, R in the high byte */);
/* input drawables are RGBA8, rows top to bottom, R first (color.h:36-43 packing is applied on load).
 * _host copies to the device; _device adopts a device pointer owned by the caller (e.g. an NCCL-broadcast buffer). */
int mmb_set_userval_image_host(mmb_invocation *inv, int index, const uint8_t *rgba, int width, int height);
int mmb_set_userval_image_device(mmb_invocation *inv, int index, const void *device_rgba, int width, int height);

/* ---- render ------------------------------------------------------------------
 * mathfuncs_t {init_frame, init_slice, calc_lines} (compiler.h:50-66, drawable.h:57-60).
 * mmb_init_frame: frame-constant values are computed once on the host and
 *   native filters / render() run as kernels; reference init_frame_<f>
 *   (new_template.c.in:314-337) via invocation_new_frame (mathmap_common.c:798).
 * mmb_calc_lines: rows [first_row, last_row) of the region, one launch for the
 *   band; q is a HOST buffer laid out like the reference's (row stride
 *   width*bpp, or float[4] per pixel when floatmap != 0); the device result is
 *   copied into it.  Reference calc_lines_<f> (new_template.c.in:208-312).
 * mmb_calc_lines_device: same with a DEVICE buffer and an optional CUstream /
 *   cudaStream_t (as void*, 0 = default stream); no host copy, asynchronous.
 */
int mmb_init_frame(mmb_invocation *inv, int frame, float t);
int mmb_calc_lines(mmb_invocation *inv, int first_row, int last_row, void *q, int floatmap);
int mmb_calc_lines_device(mmb_invocation *inv, int first_row, int last_row, void *device_q, int floatmap, void *stream);

/* mmb_calc_lines_slice IS mathfuncs_t.calc_lines (compiler.h:50-56): `void calc_lines(mathmap_slice_t *slice, image_t
 * *closure, int first_row, int last_row, void *q, int floatmap)` with every parameter the reference's generated
 * calc_lines_<f> reads from its slice, frame and invocation (new_template.c.in:208-312):
 *   frame_render_width/height ... mathmap_frame_t (mathmap.h:207-219): the frame the virtual coordinates refer to
 *   region_x/y/width/height ..... mathmap_slice_t (mathmap.h:221-230): the rectangle of that frame this call covers;
 *                                 GIMP hands over tile-sized regions with region_x != 0 (mathmap.c:1160-1175)
 *   sampling_offset_x/y ......... mathmap_slice_t: 0, or -0.5 for the one-column-wider slice of supersampling
 *   row_stride .................. invocation->row_stride in bytes (mathmap.c:1168 sets the GIMP region's rowstride);
 *                                 floatmap output advances by frame_render_width float[4] pixels instead
 * Rows [max(0, first_row), min(last_row, region_y + region_height)) are rendered (absolute frame rows); q points at the
 * first rendered row, pixel (region_x + c) of a row at q + c * output_bpp.  No supersampling combine happens here: the
 * reference's caller does it with three calls per row (call_invocation, mathmap_common.c:880-927); mmb_set_supersampling
 * only selects the nearest sampler's missing +0.5 (builtins.c:155-159) like invocation->supersampling.
 * mmb_calc_lines above is the whole call_invocation for a full-width band, supersampling included. */
typedef struct mmb_slice {
    int frame_render_width, frame_render_height;
    int region_x, region_y, region_width, region_height;
    float sampling_offset_x, sampling_offset_y;
    int row_stride;
} mmb_slice;
int mmb_calc_lines_slice(mmb_invocation *inv, const mmb_slice *slice, int first_row, int last_row, void *q, int floatmap);
/* the same into device memory, asynchronously on `stream` (0 = the library's stream, the legacy default stream) */
int mmb_calc_lines_slice_device(mmb_invocation *inv, const mmb_slice *slice, int first_row, int last_row, void *device_q, int floatmap,
                                void *stream);
/* Row-band sharding across GPUs with load balance (the reference splits one frame into contiguous bands per
 * thread, mathmap_common.c:991-1003; escape-time filters make contiguous bands unequal): renders the 8-row
 * blocks b with b % count == phase into device_q, compactly (this rank's k-th block at rows [8k, 8k+8)). */
int mmb_calc_lines_interleaved_device(mmb_invocation *inv, int phase, int count, void *device_q, void *stream);
/* Batched entry for frame sharding: renders n frames (frame numbers
 * and t values given) into consecutive W*H*bpp device buffers starting at device_q. */
int mmb_render_frames_device(mmb_invocation *inv, int n, const int *frames, const float *ts, void *device_q, void *stream);
int mmb_synchronize(mmb_invocation *inv);
/* number of kernels launched by this invocation so far (bench accounting) */
long mmb_launch_count(const mmb_invocation *inv);
/* name of the most recently launched pixel kernel (for profiler filters) */
const char *mmb_kernel_name(const mmb_invocation *inv);

/* ---- native filters, callable directly (native-filters/native-filters.h) -----
 * Gaussian blur of a device float4 image, sigma in pixels (gauss.c:643-670 chooses
 * IIR for sigma >= 0.5 in both axes, the run-length FIR otherwise). */
int mmb_gaussian_blur_device(int device, const float *device_in, float *device_out, int width, int height, float sigma_h_px,
                             float sigma_v_px, void *stream);

/* NVRTC-compiles the module for sm_100a without needing a GPU (build check); returns cubin bytes or -1 */
long mmb_module_compile_check(mmb_module *m, int antialiasing, int precise_math);
long mmb_module_compile_check_fast(mmb_module *m, int antialiasing, int precise_math); /* the same with mmb_set_fast_compile(1) */
/* Optional persistent cubin cache, process-wide: with a directory set, compiled kernels are stored there under a key of the
 * NVRTC version, options, device runtime and generated source, and loaded instead of recompiled by later processes.  The
 * reference has no equivalent (gcc runs on every load of a filter, backends/cc.c:634-758).  NULL or "" turns it off (default). */
int mmb_set_cubin_cache_dir(const char *dir);
/* the IIR coefficients the blur computes on the host: 30 doubles n_p n_m d_p d_m bd_p bd_m (gauss.c:39-115) */
void mmb_gauss_iir_constants(float std_dev, double *out30);
const char *mmb_last_error(void);
const char *mmb_version(void);

#ifdef __cplusplus
}
#endif
#endif
