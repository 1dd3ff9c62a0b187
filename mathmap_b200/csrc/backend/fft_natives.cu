// The FFT-based native filters on the device (reference native-filters/convolve.c:69-357):
// convolve, half_convolve, visualize_fft.  The reference uses FFTW3 double r2c/c2r 2-D
// transforms per channel; here cuFFT D2Z/Z2D (a library call, like the reference's) plus
// small hand-written kernels for the channel (de)interleave, the centred-kernel shift,
// normalisation, spectrum products and the mirrored magnitude layout.
#include <cuda_runtime.h>
#include <cufft.h>

#include <map>
#include <mutex>
#include <string>

#include "kernels.h"

namespace mmbackend {

namespace {

struct PlanKey {
    int device, h, w, type;
    bool operator<(const PlanKey &o) const { return std::tie(device, h, w, type) < std::tie(o.device, o.h, o.w, o.type); }
};
// Plans are shared by the invocations of a process; a plan carries its stream, so one lock covers a native filter
// from cufftSetStream to the completion of its last transform (fft_convolve etc. hold it for their whole run).
std::map<PlanKey, cufftHandle> g_plans;
std::mutex g_plans_mu;

bool get_plan(int h, int w, cufftType type, cudaStream_t stream, cufftHandle *out, std::string &err) {
    int device = 0;
    cudaGetDevice(&device);
    PlanKey key{device, h, w, (int)type};
    auto it = g_plans.find(key);
    if (it == g_plans.end()) {
        cufftHandle p;
        cufftResult r = cufftPlan2d(&p, h, w, type);
        if (r != CUFFT_SUCCESS) { err = "cufftPlan2d failed (" + std::to_string((int)r) + ")"; return false; }
        it = g_plans.emplace(key, p).first;
    }
    if (cufftSetStream(it->second, stream) != CUFFT_SUCCESS) { err = "cufftSetStream failed"; return false; }
    *out = it->second;
    return true;
}

// plane[i] = img[(i + shift) % n].channel
__global__ void __launch_bounds__(256) extract_channel_kernel(const float4 *img, double *plane, long long n, int channel, long long shift) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    long long j = i + shift;
    if (j >= n) j -= n;
    const float *p = (const float *)(img + j);
    plane[i] = (double)p[channel];
}

// Sum of a plane in a fixed order (the reference adds sequentially, convolve.c:117-122; any fixed order is as
// good, a run-to-run varying one is not): SUM_BLOCKS partial sums, then one block adds those.
#define SUM_BLOCKS 1024
__global__ void __launch_bounds__(256) sum_kernel(const double *plane, long long n, double *partials) {
    __shared__ double sh[256];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) s += plane[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) sum_partials_kernel(const double *partials, double *result) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < SUM_BLOCKS; i += 256) s += partials[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) *result = sh[0];
}

__global__ void __launch_bounds__(256) scale_by_inverse_kernel(double *plane, long long n, const double *sum) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double factor = 1.0 / *sum;
    plane[i] *= factor;
}

__global__ void __launch_bounds__(256) complex_multiply_kernel(cufftDoubleComplex *a, const cufftDoubleComplex *b, long long cn) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= cn) return;
    cufftDoubleComplex x = a[i], y = b[i], r;
    r.x = x.x * y.x - x.y * y.y;
    r.y = x.x * y.y + x.y * y.x;
    a[i] = r;
}

// half_convolve: spectrum[x + y*cw] *= mask[(x + y*w + nhalf) mod n].channel
__global__ void __launch_bounds__(256) mask_multiply_kernel(cufftDoubleComplex *a, const float4 *mask, int w, int h, int cw, long long nhalf, int channel) {
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= cw || y >= h) return;
    long long n = (long long)w * h, idx = (long long)x + (long long)y * w + nhalf;
    if (idx >= n) idx -= n;
    double m = (double)((const float *)(mask + idx))[channel];
    cufftDoubleComplex v = a[(size_t)y * cw + x];
    v.x *= m;
    v.y *= m;
    a[(size_t)y * cw + x] = v;
}

// out[i].channel = plane[i] / n
__global__ void __launch_bounds__(256) store_channel_kernel(const double *plane, float4 *out, long long n, int channel) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    ((float *)(out + i))[channel] = (float)(plane[i] / (double)n);
}

__global__ void __launch_bounds__(256) copy_alpha_kernel(const float4 *in, float4 *out, long long n, int set_one) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    ((float *)(out + i))[3] = set_one ? 1.0f : ((const float *)(in + i))[3];
}

// visualize_fft: |z| / sqrt(n) mirrored around the centre column, rows shifted by h/2.
// The reference scatters spectrum column x to out_x1 = cw-1-x and out_x2 = x+w-cw for x ascending
// (convolve.c:325-338); for even widths two columns are written twice and the larger x wins.  Written
// here as a gather per output pixel so the result is deterministic and identical.
__global__ void __launch_bounds__(256) magnitude_kernel(const cufftDoubleComplex *a, float4 *out, int w, int h, int cw, double sqrtn, int channel) {
    int c = blockIdx.x * 32 + (threadIdx.x & 31), out_y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (c >= w || out_y >= h) return;
    int y = out_y - h / 2;
    if (y < 0) y += h;
    int xa = cw - 1 - c, xb = c - (w - cw);
    int x = -1;
    if (xa >= 0 && xa < cw) x = xa;
    if (xb >= 0 && xb < cw && xb > x) x = xb;
    if (x < 0) return;
    cufftDoubleComplex z = a[(size_t)y * cw + x];
    ((float *)(out + ((size_t)c + (size_t)out_y * w)))[channel] = (float)(hypot(z.x, z.y) / sqrtn);
}

// nearest floatmap -> floatmap resample: render_image(force) of a floatmap (builtins.c:303-342 + :249-265)
__global__ void __launch_bounds__(256) floatmap_resample_kernel(const float4 *src, int sw, int sh, float sax, float sbx, float say, float sby, float xf, float yf,
                                                                float4 *out, int width, int height, float ax, float bx, float ay, float by) {
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= width || y >= height) return;
    float fx = __fmul_rn(__fdiv_rn(__fsub_rn((float)x, bx), ax), xf), fy = __fmul_rn(__fdiv_rn(__fsub_rn((float)y, by), ay), yf);
    float px = __fadd_rn(__fmul_rn(sax, fx), sbx), py = __fadd_rn(__fmul_rn(say, fy), sby);
    // (int)lrintf(), x86-64: NaN and |v| >= 2^63 become 0, otherwise the low 32 bits of the rounded value (mm_runtime.cuh: mm_lrintf_to_int)
    const int rx = fabsf(px) < 9223372036854775808.0f ? (int)__float2ll_rn(px) : 0, ry = fabsf(py) < 9223372036854775808.0f ? (int)__float2ll_rn(py) : 0;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rx >= 0 && rx < sw && ry >= 0 && ry < sh) v = src[(size_t)ry * sw + rx];
    out[(size_t)y * width + x] = v;
}

inline unsigned blocks(long long n) { return (unsigned)((n + 255) / 256); }

struct Work {
    double *plane = nullptr, *sum = nullptr;
    cufftDoubleComplex *spec_a = nullptr, *spec_b = nullptr;
    ~Work() {
        if (plane) cudaFree(plane);
        if (sum) cudaFree(sum);
        if (spec_a) cudaFree(spec_a);
        if (spec_b) cudaFree(spec_b);
    }
    bool alloc(long long n, long long cn, bool two, std::string &err) {
        if (cudaMalloc(&plane, sizeof(double) * n) != cudaSuccess || cudaMalloc(&sum, sizeof(double) * (SUM_BLOCKS + 1)) != cudaSuccess ||
            cudaMalloc(&spec_a, sizeof(cufftDoubleComplex) * cn) != cudaSuccess ||
            (two && cudaMalloc(&spec_b, sizeof(cufftDoubleComplex) * cn) != cudaSuccess)) {
            err = "cudaMalloc failed in FFT native filter";
            return false;
        }
        return true;
    }
};

}  // namespace

void launch_floatmap_resample(const float *src, int sw, int sh, float sax, float sbx, float say, float sby, float xf, float yf, float *out, int width,
                              int height, float ax, float bx, float ay, float by, cudaStream_t stream) {
    dim3 grid((width + 31) / 32, (height + 7) / 8);
    floatmap_resample_kernel<<<grid, 256, 0, stream>>>((const float4 *)src, sw, sh, sax, sbx, say, sby, xf, yf, (float4 *)out, width, height, ax, bx, ay, by);
}

bool fft_convolve(const float *in, const float *filt, float *out, int w, int h, int normalize, int copy_alpha, cudaStream_t stream, long *launches,
                  std::string &err) {
    std::lock_guard<std::mutex> lock(g_plans_mu);
    long long n = (long long)w * h, cn = (long long)h * (w / 2 + 1), nhalf = (long long)w * (h / 2) + w / 2;
    cufftHandle fwd, inv;
    if (!get_plan(h, w, CUFFT_D2Z, stream, &fwd, err) || !get_plan(h, w, CUFFT_Z2D, stream, &inv, err)) return false;
    Work wk;
    if (!wk.alloc(n, cn, true, err)) return false;
    int channels = copy_alpha ? 3 : 4;
    for (int c = 0; c < channels; ++c) {
        extract_channel_kernel<<<blocks(n), 256, 0, stream>>>((const float4 *)in, wk.plane, n, c, 0);
        if (cufftExecD2Z(fwd, wk.plane, wk.spec_a) != CUFFT_SUCCESS) { err = "cufftExecD2Z failed"; return false; }
        extract_channel_kernel<<<blocks(n), 256, 0, stream>>>((const float4 *)filt, wk.plane, n, c, n - nhalf);
        if (normalize) {
            sum_kernel<<<SUM_BLOCKS, 256, 0, stream>>>(wk.plane, n, wk.sum + 1);
            sum_partials_kernel<<<1, 256, 0, stream>>>(wk.sum + 1, wk.sum);
            scale_by_inverse_kernel<<<blocks(n), 256, 0, stream>>>(wk.plane, n, wk.sum);
            *launches += 3;
        }
        if (cufftExecD2Z(fwd, wk.plane, wk.spec_b) != CUFFT_SUCCESS) { err = "cufftExecD2Z failed"; return false; }
        complex_multiply_kernel<<<blocks(cn), 256, 0, stream>>>(wk.spec_a, wk.spec_b, cn);
        if (cufftExecZ2D(inv, wk.spec_a, wk.plane) != CUFFT_SUCCESS) { err = "cufftExecZ2D failed"; return false; }
        store_channel_kernel<<<blocks(n), 256, 0, stream>>>(wk.plane, (float4 *)out, n, c);
        *launches += 4;
    }
    if (copy_alpha) { copy_alpha_kernel<<<blocks(n), 256, 0, stream>>>((const float4 *)in, (float4 *)out, n, 0); ++*launches; }
    cudaStreamSynchronize(stream);  // the work buffers are freed on return
    return cudaGetLastError() == cudaSuccess;
}

bool fft_half_convolve(const float *in, const float *mask, float *out, int w, int h, int copy_alpha, cudaStream_t stream, long *launches, std::string &err) {
    std::lock_guard<std::mutex> lock(g_plans_mu);
    long long n = (long long)w * h, nhalf = (long long)w * (h / 2) + w / 2;
    int cw = w / 2 + 1;
    long long cn = (long long)h * cw;
    cufftHandle fwd, inv;
    if (!get_plan(h, w, CUFFT_D2Z, stream, &fwd, err) || !get_plan(h, w, CUFFT_Z2D, stream, &inv, err)) return false;
    Work wk;
    if (!wk.alloc(n, cn, false, err)) return false;
    int channels = copy_alpha ? 3 : 4;
    dim3 grid((cw + 31) / 32, (h + 7) / 8);
    for (int c = 0; c < channels; ++c) {
        extract_channel_kernel<<<blocks(n), 256, 0, stream>>>((const float4 *)in, wk.plane, n, c, 0);
        if (cufftExecD2Z(fwd, wk.plane, wk.spec_a) != CUFFT_SUCCESS) { err = "cufftExecD2Z failed"; return false; }
        mask_multiply_kernel<<<grid, 256, 0, stream>>>(wk.spec_a, (const float4 *)mask, w, h, cw, nhalf, c);
        if (cufftExecZ2D(inv, wk.spec_a, wk.plane) != CUFFT_SUCCESS) { err = "cufftExecZ2D failed"; return false; }
        store_channel_kernel<<<blocks(n), 256, 0, stream>>>(wk.plane, (float4 *)out, n, c);
        *launches += 3;
    }
    if (copy_alpha) { copy_alpha_kernel<<<blocks(n), 256, 0, stream>>>((const float4 *)in, (float4 *)out, n, 0); ++*launches; }
    cudaStreamSynchronize(stream);
    return cudaGetLastError() == cudaSuccess;
}

bool fft_visualize(const float *in, float *out, int w, int h, int ignore_alpha, cudaStream_t stream, long *launches, std::string &err) {
    std::lock_guard<std::mutex> lock(g_plans_mu);
    long long n = (long long)w * h;
    int cw = w / 2 + 1;
    long long cn = (long long)h * cw;
    cufftHandle fwd;
    if (!get_plan(h, w, CUFFT_D2Z, stream, &fwd, err)) return false;
    Work wk;
    if (!wk.alloc(n, cn, false, err)) return false;
    cudaMemsetAsync(out, 0, sizeof(float) * 4 * n, stream);
    int channels = ignore_alpha ? 3 : 4;
    dim3 grid((w + 31) / 32, (h + 7) / 8);
    double sqrtn = sqrt((double)n);
    for (int c = 0; c < channels; ++c) {
        extract_channel_kernel<<<blocks(n), 256, 0, stream>>>((const float4 *)in, wk.plane, n, c, 0);
        if (cufftExecD2Z(fwd, wk.plane, wk.spec_a) != CUFFT_SUCCESS) { err = "cufftExecD2Z failed"; return false; }
        magnitude_kernel<<<grid, 256, 0, stream>>>(wk.spec_a, (float4 *)out, w, h, cw, sqrtn, c);
        *launches += 2;
    }
    if (ignore_alpha) { copy_alpha_kernel<<<blocks(n), 256, 0, stream>>>((const float4 *)in, (float4 *)out, n, 1); ++*launches; }
    cudaStreamSynchronize(stream);
    return cudaGetLastError() == cudaSuccess;
}

}  // namespace mmbackend
