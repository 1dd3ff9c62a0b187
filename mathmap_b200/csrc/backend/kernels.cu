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

// ------------------------------------------------------ Gaussian IIR: the constants
// The recursive passes themselves are in gauss_iir.cu.
struct GaussCoeffs {
    double n_p[5], n_m[5], d_p[5], d_m[5], bd_p[5], bd_m[5];
};

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
