// Launchers of the statically compiled kernels (kernels.cu).
#pragma once
#include <cuda_runtime_api.h>

#include <cstddef>
#include <string>

#include "../runtime/mm_types.h"

namespace mmbackend {

void launch_drawable_to_floatmap(const mm_image &img, float *out, int width, int height, float ax, float bx, float ay, float by, int edge_x_mode,
                                 int edge_y_mode, unsigned edge_x, unsigned edge_y, int supersampling, cudaStream_t stream);
size_t gauss_iir_scratch_bytes(int width, int height);
void launch_drawable_to_bytes(const mm_image &img, void *out, int width, int height, float ax, float bx, float ay, float by, int edge_x_mode,
                              int edge_y_mode, unsigned edge_x, unsigned edge_y, int supersampling, cudaStream_t stream);
bool drawable_render_is_identity(const mm_image &img, int width, int height, float ax, float bx, float ay, float by, int supersampling);
void launch_gauss_iir(const void *in, bool in_is_rgba8, float *out, double *scratch, int width, int height, float sigma_h, float sigma_v,
                      cudaStream_t stream);
void launch_gauss_iir_columns(const void *in, bool in_is_rgba8, float *mid, double *scratch, int width, int height, float sigma_v, cudaStream_t stream);
void launch_gauss_iir_rows(const float *mid, void *out, bool rgba8, long long out_pitch, double *scratch, int width, int first_row, int nrows, float sigma_h,
                           cudaStream_t stream);
size_t gauss_rle_curve_bytes(float sigma_h, float sigma_v);
void launch_gauss_rle(const float *in, float *tmp, float *out, int width, int height, float sigma_h, float sigma_v, void *curve_mem, cudaStream_t stream);
void gauss_iir_constants_host(float std_dev, double *out30);
void launch_supersample_combine(const unsigned char *shortimg, const unsigned char *longimg, unsigned char *out, int width, int height, int long_rows,
                                int bpp, cudaStream_t stream);

// fft_natives.cu
void launch_floatmap_resample(const float *src, int sw, int sh, float sax, float sbx, float say, float sby, float xf, float yf, float *out, int width,
                              int height, float ax, float bx, float ay, float by, cudaStream_t stream);
bool fft_convolve(const float *in, const float *filt, float *out, int w, int h, int normalize, int copy_alpha, cudaStream_t stream, long *launches,
                  std::string &err);
bool fft_half_convolve(const float *in, const float *mask, float *out, int w, int h, int copy_alpha, cudaStream_t stream, long *launches, std::string &err);
bool fft_visualize(const float *in, float *out, int w, int h, int ignore_alpha, cudaStream_t stream, long *launches, std::string &err);

}  // namespace mmbackend
