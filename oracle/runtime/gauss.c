/* ORACLE (test infrastructure): the native Gaussian blur, restated from the
 * reference native-filters/gauss.c:
 *   IIR coefficients (4th-order causal/anticausal recursive Gaussian) .. :39-115
 *   gauss_iir: columns then rows, per channel, double accumulators ...... :127-262
 *   truncated-Gaussian run-length FIR for sigma < 0.5 px ................ :265-639
 *   entry point, sigma in pixels = |sigma * ax|, path selection ......... :643-670
 * and the per-invocation result cache native-filters/cache.c:111-156 (keyed by
 * filter + scalar arguments + image identity).
 */
#include "mmo_runtime.h"

static void find_iir_constants(double *n_p, double *n_m, double *d_p, double *d_m, double *bd_p, double *bd_m, float std_dev) {
    int i;
    double div = sqrt(2 * M_PI) * std_dev;
    double x0 = -1.783 / std_dev, x1 = -1.723 / std_dev, x2 = 0.6318 / std_dev, x3 = 1.997 / std_dev;
    double x4 = 1.6803 / div, x5 = 3.735 / div, x6 = -0.6803 / div, x7 = -0.2598 / div;
    double sum_n_p = 0.0, sum_n_m = 0.0, sum_d = 0.0, a, b;

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

    for (i = 0; i <= 4; i++) d_m[i] = d_p[i];
    n_m[0] = 0.0;
    for (i = 1; i <= 4; i++) n_m[i] = n_p[i] - d_p[i] * n_p[0];

    for (i = 0; i <= 4; i++) {
        sum_n_p += n_p[i];
        sum_n_m += n_m[i];
        sum_d += d_p[i];
    }
    a = sum_n_p / (1.0 + sum_d);
    b = sum_n_m / (1.0 + sum_d);
    for (i = 0; i <= 4; i++) {
        bd_p[i] = d_p[i] * a;
        bd_m[i] = d_m[i] * b;
    }
}

/* exported so tests can compare the device's host-computed coefficients */
void mmo_gauss_iir_constants(float std_dev, double *out30) {
    find_iir_constants(out30, out30 + 5, out30 + 10, out30 + 15, out30 + 20, out30 + 25, std_dev);
}

/* one line of length n, read with stride `stride` floats, written back in place */
static void iir_line(float *line, int stride, int n, const double *n_p, const double *n_m, const double *d_p, const double *d_m,
                     const double *bd_p, const double *bd_m, double *val_p, double *val_m, float *src) {
    int k, i, j, terms;
    float *sp_p, *sp_m, initial_p, initial_m;
    double *vp, *vm;
    memset(val_p, 0, n * sizeof(double));
    memset(val_m, 0, n * sizeof(double));
    for (k = 0; k < n; ++k) src[k] = line[(size_t)k * stride];
    sp_p = src;
    sp_m = src + (n - 1);
    vp = val_p;
    vm = val_m + (n - 1);
    initial_p = sp_p[0];
    initial_m = sp_m[0];
    for (k = 0; k < n; k++) {
        double *vpptr = vp, *vmptr = vm;
        terms = (k < 4) ? k : 4;
        for (i = 0; i <= terms; i++) {
            *vpptr += n_p[i] * sp_p[-i] - d_p[i] * vp[-i];
            *vmptr += n_m[i] * sp_m[i] - d_m[i] * vm[i];
        }
        for (j = i; j <= 4; j++) {
            *vpptr += (n_p[j] - bd_p[j]) * initial_p;
            *vmptr += (n_m[j] - bd_m[j]) * initial_m;
        }
        sp_p++;
        sp_m--;
        vp++;
        vm--;
    }
    for (k = 0; k < n; ++k) line[(size_t)k * stride] = val_p[k] + val_m[k];
}

static mmo_image *gauss_iir(mmo_image *floatmap, float horizontal_std_dev, float vertical_std_dev, mmo_pools *pools) {
    mmo_image *out = mmo_floatmap_copy(floatmap, pools);
    int width = floatmap->pixel_width, height = floatmap->pixel_height;
    int mx = MAX(width, height), channel, row, col;
    double n_p[5], n_m[5], d_p[5], d_m[5], bd_p[5], bd_m[5];
    double *val_p = (double *)malloc(mx * sizeof(double)), *val_m = (double *)malloc(mx * sizeof(double));
    float *src = (float *)malloc(mx * sizeof(float));

    find_iir_constants(n_p, n_m, d_p, d_m, bd_p, bd_m, vertical_std_dev);
    for (channel = 0; channel < NUM_FLOATMAP_CHANNELS; ++channel)
        for (col = 0; col < width; col++)
            iir_line(out->fdata + (size_t)col * 4 + channel, width * 4, height, n_p, n_m, d_p, d_m, bd_p, bd_m, val_p, val_m, src);

    find_iir_constants(n_p, n_m, d_p, d_m, bd_p, bd_m, horizontal_std_dev);
    for (channel = 0; channel < NUM_FLOATMAP_CHANNELS; ++channel)
        for (row = 0; row < height; row++)
            iir_line(out->fdata + (size_t)row * width * 4 + channel, 4, width, n_p, n_m, d_p, d_m, bd_p, bd_m, val_p, val_m, src);

    free(val_p);
    free(val_m);
    free(src);
    return out;
}

/* ---- sigma < 0.5 px: run-length / direct truncated-Gaussian FIR ---- */
static void make_rle_curve(double sigma, float **p_curve, int *p_length, float **p_sum, float *p_total) {
    const double sigma2 = 2 * sigma * sigma;
    const double l = sqrt(-sigma2 * log(1.0 / 255.0));
    int i, n, length;
    float *sum, *curve;
    n = ceil(l) * 2;
    if ((n % 2) == 0) n += 1;
    curve = (float *)malloc(sizeof(float) * n);
    length = n / 2;
    curve += length;
    curve[0] = 1.0;
    for (i = 1; i <= length; i++) {
        float temp = exp(-(i * i) / sigma2);
        curve[-i] = temp;
        curve[i] = temp;
    }
    sum = (float *)malloc(sizeof(float) * (2 * length + 1));
    sum[0] = 0;
    for (i = 1; i <= length * 2; i++) sum[i] = curve[i - length - 1] + sum[i - 1];
    sum += length;
    *p_total = sum[length] - sum[-length];
    *p_curve = curve;
    *p_sum = sum;
    *p_length = length;
}

static int run_length_encode(const float *src, int *rle, float *pix, int dist, int width, int border) {
    float last;
    int count = 0, i, same = 0;
    src += dist * (width - 1);
    rle += width + border - 1;
    pix += width + border - 1;
    last = *src;
    for (i = 0; i < border; i++) {
        count++;
        *pix-- = last;
        *rle-- = count;
    }
    for (i = 0; i < width; i++) {
        float c = *src;
        src -= dist;
        if (c == last) {
            count++;
            *pix-- = last;
            *rle-- = count;
            same++;
        } else {
            count = 1;
            last = c;
            *pix-- = last;
            *rle-- = count;
        }
    }
    for (i = 0; i < border; i++) {
        count++;
        *pix-- = last;
        *rle-- = count;
    }
    return same;
}

/* note the int-typed ctotal and s2: they truncate, as in the reference (:383-384, :404) */
static void do_encoded_lre(const int *enc, const float *src, float *dest, int width, int length, int dist, int ctotal, const float *csum) {
    int col;
    for (col = 0; col < width; col++, dest += dist) {
        const int *rpt;
        const float *pix;
        int nb, i, start = -length;
        float s1, val = 0.0;
        rpt = &enc[col + start];
        pix = &src[col + start];
        s1 = csum[start];
        nb = rpt[0];
        i = start + nb;
        while (i <= length) {
            int s2 = csum[i];
            val += pix[0] * (s2 - s1);
            s1 = s2;
            rpt = &rpt[nb];
            pix = &pix[nb];
            nb = rpt[0];
            i += nb;
        }
        val += pix[0] * (csum[length] - s1);
        val = val / ctotal;
        *dest = val;
    }
}

static void do_full_lre(const float *src, float *dest, int width, int length, int dist, const float *curve, float ctotal) {
    int col;
    for (col = 0; col < width; col++, dest += dist) {
        const float *x1, *x2, *c = &curve[0];
        int i;
        float val = 0.0;
        x1 = x2 = &src[col];
        val += x1[0] * c[0];
        c += 1;
        x1 += 1;
        x2 -= 1;
        for (i = length; i >= 1; --i) {
            val += (x1[0] + x2[-0]) * c[0];
            c += 1;
            x1 += 1;
            x2 -= 1;
        }
        val = val / ctotal;
        *dest = val;
    }
}

static void rle_pass(float *data, int line_stride, int elem_stride, int nlines, int n, float std_dev) {
    float *curve, *sum, total;
    int length, line, b, k;
    int *rle;
    float *pix, *src, *dest;
    if (!(std_dev > 0.0)) return;
    make_rle_curve(std_dev, &curve, &length, &sum, &total);
    rle = (int *)malloc(sizeof(int) * (n + 2 * length)) ;
    pix = (float *)malloc(sizeof(float) * (n + 2 * length));
    src = (float *)malloc(sizeof(float) * n * 4);
    dest = (float *)malloc(sizeof(float) * n * 4);
    for (line = 0; line < nlines; ++line) {
        float *base = data + (size_t)line * line_stride;
        for (k = 0; k < n; ++k) memcpy(src + k * 4, base + (size_t)k * elem_stride, sizeof(float) * 4);
        for (b = 0; b < 4; b++) {
            int same = run_length_encode(src + b, rle + length, pix + length, 4, n, length);
            if (same > (3 * n) / 4)
                do_encoded_lre(rle + length, pix + length, dest + b, n, length, 4, total, sum);
            else
                do_full_lre(pix + length, dest + b, n, length, 4, curve, total);
        }
        for (k = 0; k < n; ++k) memcpy(base + (size_t)k * elem_stride, dest + k * 4, sizeof(float) * 4);
    }
    free(rle);
    free(pix);
    free(src);
    free(dest);
    free(sum - length);
    free(curve - length);
}

static mmo_image *gauss_rle(mmo_image *floatmap, float horizontal_std_dev, float vertical_std_dev, mmo_pools *pools) {
    mmo_image *out = mmo_floatmap_copy(floatmap, pools);
    int width = floatmap->pixel_width, height = floatmap->pixel_height;
    rle_pass(out->fdata, 4, width * 4, width, height, vertical_std_dev);    /* columns */
    rle_pass(out->fdata, width * 4, 4, height, width, horizontal_std_dev);  /* rows */
    return out;
}

typedef struct mmo_cache_entry {
    struct mmo_cache_entry *next;
    mmo_image *in;
    float h, v;
    mmo_image *result;
} mmo_cache_entry;

mmo_image *native_filter_gaussian_blur(mmo_invocation *invocation, mmo_userval *args, mmo_pools *pools) {
    mmo_image *floatmap = args[0].v.image, *result;
    float horizontal_std_dev = args[1].v.float_const, vertical_std_dev = args[2].v.float_const;
    mmo_cache_entry *e;
    for (e = invocation->cache; e; e = e->next)
        if (e->in == floatmap && e->h == horizontal_std_dev && e->v == vertical_std_dev) return e->result;
    if (floatmap->type != MMO_IMAGE_FLOATMAP)
        floatmap = mmo_render_image(invocation, floatmap, invocation->render_width, invocation->render_height, pools, 0);
    horizontal_std_dev = fabs(horizontal_std_dev * floatmap->ax);
    vertical_std_dev = fabs(vertical_std_dev * floatmap->ay);
    if (horizontal_std_dev < 0.5 || vertical_std_dev < 0.5)
        result = gauss_rle(floatmap, horizontal_std_dev, vertical_std_dev, pools);
    else
        result = gauss_iir(floatmap, horizontal_std_dev, vertical_std_dev, pools);
    e = (mmo_cache_entry *)mmo_pools_alloc(pools, sizeof(mmo_cache_entry));
    e->in = args[0].v.image;
    e->h = args[1].v.float_const;
    e->v = args[2].v.float_const;
    e->result = result;
    e->next = invocation->cache;
    invocation->cache = e;
    return result;
}

/* Float-level entry for the tests: blurs data[height][width][4] in place with pixel-unit sigmas, the same dispatch
 * (IIR for sigma >= 0.5 px on both axes, truncated FIR otherwise) as native_filter_gaussian_blur above. */
void mmo_gaussian_blur_floats(float *data, int width, int height, float sigma_h_px, float sigma_v_px) {
    mmo_pools pools;
    mmo_image *fm, *result;
    mmo_pools_init(&pools);
    fm = mmo_floatmap_alloc(width, height, &pools);
    memcpy(fm->fdata, data, sizeof(float) * 4 * (size_t)width * height);
    if (sigma_h_px < 0.5 || sigma_v_px < 0.5) result = gauss_rle(fm, sigma_h_px, sigma_v_px, &pools);
    else result = gauss_iir(fm, sigma_h_px, sigma_v_px, &pools);
    memcpy(data, result->fdata, sizeof(float) * 4 * (size_t)width * height);
    mmo_pools_free(&pools);
}
