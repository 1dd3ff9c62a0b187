/* ORACLE (test infrastructure): render driver.  Restates the reference's frame /
 * slice / thread fan-out logic:
 *   invoke_mathmap defaults ............. mathmap_common.c:747-790 (bpp 4, edge COLOR, R = sqrt(2), stride W*4)
 *   invocation_new_frame ................ mathmap_common.c:798-816
 *   call_invocation + supersampling ..... mathmap_common.c:874-936
 *   band split across threads ........... mathmap_common.c:973-1006
 *   calc_image_values ................... userval.c:263-279
 */
#include <pthread.h>
#include <stdio.h>

#include "mmo_runtime.h"

void *mmo_main_new_frame(mmo_invocation *invocation, mmo_image *closure, int frame, float t, mmo_pools *frame_pools);
void mmo_main_calc_lines(mmo_invocation *invocation, mmo_image *closure, void *xy_vars, int frame, float t, int frame_w, int frame_h,
                         int region_x, int region_y, int region_w, int region_h, float off_x, float off_y, int first_row, int last_row,
                         void *q, int floatmap);

typedef struct mmo_render_params {
    int img_width, img_height;
    int antialiasing, supersampling;
    int edge_behaviour_x, edge_behaviour_y;
    unsigned int edge_color_x, edge_color_y;
    int output_bpp;
    int frame;
    float t;
    int num_threads;
    int floatmap;
    int num_uservals;
    mmo_userval *uservals;
    void *output; /* u8 [H][W][bpp] or float [H][W][4] */
    long taps;    /* out: drawable samples taken */
    const int *sample_rows; /* optional: render only these rows (bounded CPU-baseline samples); output has num_sample_rows rows */
    int num_sample_rows;
} mmo_render_params;

mmo_image *mmo_make_drawable(const unsigned char *rgba, int width, int height) {
    mmo_image *img = (mmo_image *)calloc(1, sizeof(mmo_image));
    img->type = MMO_IMAGE_DRAWABLE;
    img->pixel_width = width;
    img->pixel_height = height;
    img->data = rgba;
    img->num_frames = 1;
    img->scale_x = (width - 1) / 2.0;
    img->scale_y = (height - 1) / 2.0;
    img->middle_x = 1.0;
    img->middle_y = 1.0;
    return img;
}
void mmo_free_drawable(mmo_image *img) { free(img); }

typedef struct {
    mmo_invocation *invocation;
    mmo_image *closure;
    void *xy_vars;
    int frame;
    float t;
    int region_x, region_y, region_width, region_height;
    unsigned char *q;
    int floatmap;
    const int *sample_rows; /* when set: rows sample_rows[region_y .. region_y+region_height) one by one */
    size_t row_bytes;
} band_t;

static void call_invocation(band_t *b) {
    mmo_invocation *inv = b->invocation;
    int W = inv->render_width, H = inv->render_height;
    if (b->sample_rows) {
        int i;
        for (i = 0; i < b->region_height; ++i) {
            int row = b->sample_rows[b->region_y + i];
            mmo_main_calc_lines(inv, b->closure, b->xy_vars, b->frame, b->t, W, H, 0, 0, W, H, 0.0f, 0.0f, row, row + 1,
                                b->q + (size_t)i * b->row_bytes, b->floatmap);
        }
        return;
    }
    if (inv->supersampling && !b->floatmap) {
        int bpp = inv->output_bpp, row, col, i;
        unsigned char *line1 = (unsigned char *)malloc((b->region_width + 1) * bpp);
        unsigned char *line2 = (unsigned char *)malloc(b->region_width * bpp);
        unsigned char *line3 = (unsigned char *)malloc((b->region_width + 1) * bpp);
        unsigned char *q = b->q;
        mmo_main_calc_lines(inv, b->closure, b->xy_vars, b->frame, b->t, W, H, b->region_x, b->region_y, b->region_width + 1,
                            b->region_height, -0.5f, -0.5f, b->region_y, b->region_y + 1, line1, 0);
        /* a region of ONE row never writes line3 in the reference (its only line3 call is clamped away, see below) and the
         * combine reads uninitialised heap memory there; the oracle and the CUDA path define it as line1, what line3 holds at
         * the last row of every taller region */
        memcpy(line3, line1, (b->region_width + 1) * bpp);
        for (row = b->region_y; row < b->region_y + b->region_height; ++row) {
            unsigned char *p = q;
            mmo_main_calc_lines(inv, b->closure, b->xy_vars, b->frame, b->t, W, H, b->region_x, b->region_y, b->region_width,
                                b->region_height, 0.0f, 0.0f, row, row + 1, line2, 0);
            /* the "long" slice has one more row than the region in the reference only by
             * virtue of last_row clamping; row + 1 == region end is clamped away there too,
             * leaving line3 unchanged from the previous iteration (mathmap_common.c:897-901,
             * new_template.c.in:238-239) */
            mmo_main_calc_lines(inv, b->closure, b->xy_vars, b->frame, b->t, W, H, b->region_x, b->region_y, b->region_width + 1,
                                b->region_height, -0.5f, -0.5f, row + 1, row + 2, line3, 0);
            for (col = 0; col < b->region_width; ++col) {
                for (i = 0; i < bpp; ++i)
                    p[i] = (line1[col * bpp + i] + line1[(col + 1) * bpp + i] + 2 * line2[col * bpp + i] + line3[col * bpp + i] +
                            line3[(col + 1) * bpp + i]) / 6;
                p += bpp;
            }
            memcpy(line1, line3, (b->region_width + 1) * bpp);
            q += inv->row_stride;
        }
        free(line1);
        free(line2);
        free(line3);
    } else
        mmo_main_calc_lines(inv, b->closure, b->xy_vars, b->frame, b->t, W, H, b->region_x, b->region_y, b->region_width,
                            b->region_height, 0.0f, 0.0f, b->region_y, b->region_y + b->region_height, b->q, b->floatmap);
}

static void *band_thread(void *arg) {
    call_invocation((band_t *)arg);
    return NULL;
}

int mmo_render(mmo_render_params *p) {
    mmo_invocation inv;
    mmo_image closure;
    mmo_pools frame_pools;
    void *xy_vars;
    int i, n = p->num_threads < 1 ? 1 : p->num_threads;
    int first_row = 0, last_row = p->sample_rows ? p->num_sample_rows : p->img_height;
    band_t *bands;
    pthread_t *threads;

    memset(&inv, 0, sizeof inv);
    inv.img_width = inv.render_width = p->img_width;
    inv.img_height = inv.render_height = p->img_height;
    inv.image_R = sqrt(2.0);
    inv.antialiasing = p->antialiasing;
    inv.supersampling = p->supersampling;
    inv.edge_behaviour_x = p->edge_behaviour_x;
    inv.edge_behaviour_y = p->edge_behaviour_y;
    inv.edge_color_x = p->edge_color_x;
    inv.edge_color_y = p->edge_color_y;
    inv.output_bpp = p->output_bpp;
    inv.row_stride = p->img_width * p->output_bpp;

    memset(&closure, 0, sizeof closure);
    closure.type = MMO_IMAGE_CLOSURE;
    closure.pixel_width = p->img_width;
    closure.pixel_height = p->img_height;
    closure.num_args = p->num_uservals;
    if (p->num_uservals > MMO_MAX_ARGS) return -1;
    for (i = 0; i < p->num_uservals; ++i) closure.args[i] = p->uservals[i];

    mmo_pools_init(&frame_pools);
    xy_vars = mmo_main_new_frame(&inv, &closure, p->frame, p->t, &frame_pools);

    bands = (band_t *)calloc(n, sizeof(band_t));
    threads = (pthread_t *)calloc(n, sizeof(pthread_t));
    for (i = 0; i < n; ++i) {
        band_t *b = &bands[i];
        size_t row_bytes = p->floatmap ? sizeof(float) * 4 * (size_t)p->img_width : (size_t)inv.row_stride;
        b->invocation = &inv;
        b->closure = &closure;
        b->xy_vars = xy_vars;
        b->frame = p->frame;
        b->t = p->t;
        b->region_x = 0;
        b->region_width = p->img_width;
        b->region_y = first_row + (last_row - first_row) * i / n;
        b->region_height = first_row + (last_row - first_row) * (i + 1) / n - b->region_y;
        b->q = (unsigned char *)p->output + (size_t)(b->region_y - first_row) * row_bytes;
        b->floatmap = p->floatmap;
        b->sample_rows = p->sample_rows;
        b->row_bytes = row_bytes;
    }
    if (n == 1)
        call_invocation(&bands[0]);
    else {
        for (i = 0; i < n; ++i) pthread_create(&threads[i], NULL, band_thread, &bands[i]);
        for (i = 0; i < n; ++i) pthread_join(threads[i], NULL);
    }
    p->taps = inv.taps;
    free(bands);
    free(threads);
    mmo_pools_free(&frame_pools);
    return 0;
}
