/* ORACLE (test infrastructure): the FFT-based native filters, restated from the
 * reference native-filters/convolve.c:
 *   native_filter_convolve ......... :69-175  (circular convolution, kernel image centred, optional normalisation)
 *   native_filter_half_convolve .... :178-269 (multiply the spectrum by a centred real mask)
 *   native_filter_visualize_fft .... :272-357 (|spectrum| / sqrt(n), centred and mirrored)
 * The reference calls FFTW3 (third-party, absent here: double r2c/c2r 2-D, unnormalised).  The transform itself is
 * restated as a plain mixed-radix DFT in double; it computes the same mathematical quantity, differences are at the
 * 1e-15 level before the result is narrowed to float.  Pinned by the golden utilities_visualize_fft.png.
 */
#include "mmo_runtime.h"

typedef double _Complex cplx;

/* out[k] = sum_j in[j*stride] * exp(sign * 2 pi i j k / n), recursive decimation by the smallest prime factor */
static void dft(const cplx *in, cplx *out, int n, int stride, int sign, cplx *scratch) {
    int p, r, k, m;
    if (n == 1) { out[0] = in[0]; return; }
    for (p = 2; p * p <= n; ++p) if (n % p == 0) break;
    if (p * p > n) p = n;
    if (p == n) {
        for (k = 0; k < n; ++k) {
            cplx s = 0;
            for (r = 0; r < n; ++r) s += in[(size_t)r * stride] * cexp(sign * 2.0 * M_PI * I * (double)(((long long)r * k) % n) / n);
            out[k] = s;
        }
        return;
    }
    m = n / p;
    for (r = 0; r < p; ++r) dft(in + (size_t)r * stride, scratch + (size_t)r * m, m, stride * p, sign, out);
    for (k = 0; k < m; ++k)
        for (r = 0; r < p; ++r) {  /* output index k + r*m */
            cplx s = 0;
            int q, idx = k + r * m;
            for (q = 0; q < p; ++q) s += scratch[(size_t)q * m + k] * cexp(sign * 2.0 * M_PI * I * (double)(((long long)q * idx) % n) / n);
            out[idx] = s;
        }
}

/* full complex 2-D transform of a real h x w plane (rows then columns) */
static void fft2(cplx *data, int h, int w, int sign) {
    int mx = h > w ? h : w, x, y;
    cplx *line = (cplx *)malloc(sizeof(cplx) * mx), *res = (cplx *)malloc(sizeof(cplx) * mx), *scr = (cplx *)malloc(sizeof(cplx) * mx);
    for (y = 0; y < h; ++y) {
        memcpy(line, data + (size_t)y * w, sizeof(cplx) * w);
        dft(line, res, w, 1, sign, scr);
        memcpy(data + (size_t)y * w, res, sizeof(cplx) * w);
    }
    for (x = 0; x < w; ++x) {
        for (y = 0; y < h; ++y) line[y] = data[(size_t)y * w + x];
        dft(line, res, h, 1, sign, scr);
        for (y = 0; y < h; ++y) data[(size_t)y * w + x] = res[y];
    }
    free(line);
    free(res);
    free(scr);
}

static mmo_image *as_floatmap(mmo_invocation *inv, mmo_image *img, int w, int h, mmo_pools *pools) {
    return mmo_render_image(inv, img, w, h, pools, 1);
}

mmo_image *native_filter_convolve(mmo_invocation *invocation, mmo_userval *args, mmo_pools *pools) {
    mmo_image *in_image = args[0].v.image, *filter_image = args[1].v.image, *out_image;
    int normalize = args[2].v.bool_const != 0, copy_alpha = args[3].v.bool_const != 0;
    int i, n, nhalf, channel, num_channels, w, h;
    cplx *a, *b;
    if (in_image->type != MMO_IMAGE_FLOATMAP) in_image = as_floatmap(invocation, in_image, invocation->render_width, invocation->render_height, pools);
    if (filter_image->type != MMO_IMAGE_FLOATMAP || filter_image->pixel_width != in_image->pixel_width || filter_image->pixel_height != in_image->pixel_height)
        filter_image = as_floatmap(invocation, filter_image, in_image->pixel_width, in_image->pixel_height, pools);
    w = in_image->pixel_width;
    h = in_image->pixel_height;
    out_image = mmo_floatmap_alloc(w, h, pools);
    n = h * w;
    nhalf = w * (h / 2) + w / 2;
    a = (cplx *)malloc(sizeof(cplx) * n);
    b = (cplx *)malloc(sizeof(cplx) * n);
    num_channels = copy_alpha ? 3 : 4;
    for (channel = 0; channel < num_channels; ++channel) {
        for (i = 0; i < n; ++i) a[i] = in_image->fdata[(size_t)i * 4 + channel];
        fft2(a, h, w, -1);
        for (i = 0; i < n; ++i) b[i] = filter_image->fdata[(size_t)((i + n - nhalf) % n) * 4 + channel];
        if (normalize) {
            double sum = 0.0, factor;
            for (i = 0; i < n; ++i) sum += creal(b[i]);
            factor = 1.0 / sum;
            for (i = 0; i < n; ++i) b[i] *= factor;
        }
        fft2(b, h, w, -1);
        for (i = 0; i < n; ++i) a[i] *= b[i];
        fft2(a, h, w, +1);
        for (i = 0; i < n; ++i) out_image->fdata[(size_t)i * 4 + channel] = creal(a[i]) / n;
    }
    if (copy_alpha)
        for (i = 0; i < n; ++i) out_image->fdata[(size_t)i * 4 + 3] = in_image->fdata[(size_t)i * 4 + 3];
    free(a);
    free(b);
    return out_image;
}

mmo_image *native_filter_half_convolve(mmo_invocation *invocation, mmo_userval *args, mmo_pools *pools) {
    mmo_image *in_image = args[0].v.image, *filter_image = args[1].v.image, *out_image;
    int copy_alpha = args[2].v.bool_const != 0;
    int i, n, nhalf, channel, num_channels, w, h, x, y;
    cplx *a;
    if (in_image->type != MMO_IMAGE_FLOATMAP) in_image = as_floatmap(invocation, in_image, invocation->render_width, invocation->render_height, pools);
    if (filter_image->type != MMO_IMAGE_FLOATMAP || filter_image->pixel_width != in_image->pixel_width || filter_image->pixel_height != in_image->pixel_height)
        filter_image = as_floatmap(invocation, filter_image, in_image->pixel_width, in_image->pixel_height, pools);
    w = in_image->pixel_width;
    h = in_image->pixel_height;
    out_image = mmo_floatmap_alloc(w, h, pools);
    n = h * w;
    nhalf = w * (h / 2) + w / 2;
    a = (cplx *)malloc(sizeof(cplx) * n);
    num_channels = copy_alpha ? 3 : 4;
    for (channel = 0; channel < num_channels; ++channel) {
        int cw = w / 2 + 1;
        for (i = 0; i < n; ++i) a[i] = in_image->fdata[(size_t)i * 4 + channel];
        fft2(a, h, w, -1);
        /* the reference multiplies the stored half spectrum (x < cw); the c2r inverse implies the conjugate half */
        for (y = 0; y < h; ++y)
            for (x = 0; x < cw; ++x) {
                int out_idx = x + y * w, in_idx = out_idx + nhalf;
                double m;
                if (in_idx >= n) in_idx -= n;
                m = filter_image->fdata[(size_t)in_idx * 4 + channel];
                a[(size_t)y * w + x] *= m;
            }
        for (y = 0; y < h; ++y)
            for (x = cw; x < w; ++x) a[(size_t)y * w + x] = conj(a[(size_t)((h - y) % h) * w + (w - x)]);
        fft2(a, h, w, +1);
        for (i = 0; i < n; ++i) out_image->fdata[(size_t)i * 4 + channel] = creal(a[i]) / n;
    }
    if (copy_alpha)
        for (i = 0; i < n; ++i) out_image->fdata[(size_t)i * 4 + 3] = in_image->fdata[(size_t)i * 4 + 3];
    free(a);
    return out_image;
}

mmo_image *native_filter_visualize_fft(mmo_invocation *invocation, mmo_userval *args, mmo_pools *pools) {
    mmo_image *in_image = args[0].v.image, *out_image;
    int ignore_alpha = args[1].v.bool_const != 0;
    int i, n, channel, num_channels, w, h, x, y, cw;
    double sqrtn;
    cplx *a;
    if (in_image->type != MMO_IMAGE_FLOATMAP) in_image = as_floatmap(invocation, in_image, invocation->render_width, invocation->render_height, pools);
    w = in_image->pixel_width;
    h = in_image->pixel_height;
    out_image = mmo_floatmap_alloc(w, h, pools);
    n = h * w;
    sqrtn = sqrt(n);
    cw = w / 2 + 1;
    a = (cplx *)malloc(sizeof(cplx) * n);
    memset(out_image->fdata, 0, sizeof(float) * 4 * (size_t)n);
    num_channels = ignore_alpha ? 3 : 4;
    for (channel = 0; channel < num_channels; ++channel) {
        for (i = 0; i < n; ++i) a[i] = in_image->fdata[(size_t)i * 4 + channel];
        fft2(a, h, w, -1);
        for (y = 0; y < h; ++y) {
            int out_y = y + h / 2;
            if (out_y >= h) out_y -= h;
            for (x = 0; x < cw; ++x) {
                int out_x1 = cw - 1 - x, out_x2 = x + w - cw;
                double val = cabs(a[(size_t)y * w + x]) / sqrtn;
                out_image->fdata[((size_t)out_x1 + (size_t)out_y * w) * 4 + channel] = val;
                out_image->fdata[((size_t)out_x2 + (size_t)out_y * w) * 4 + channel] = val;
            }
        }
    }
    if (ignore_alpha)
        for (i = 0; i < n; ++i) out_image->fdata[(size_t)i * 4 + 3] = 1.0;
    free(a);
    return out_image;
}
