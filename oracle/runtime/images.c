/* ORACLE (test infrastructure): input sampling, edge behaviour, floatmaps and
 * render_image, restated from the reference:
 *   apply_edge_behaviour ............ builtins/builtins.c:41-119
 *   get_pixel ....................... builtins/builtins.c:122-130 + mathmap.c:1195-1212 (bounds -> edge colour)
 *                                     + mathmap_cmdline.c:143-144 (frame out of range -> white)
 *   get_image_drawable .............. builtins/builtins.c:133-147
 *   nearest / bilinear sampling ..... builtins/builtins.c:150-245, color.h:48-54
 *   get_floatmap_pixel .............. builtins/builtins.c:249-265
 *   render_image .................... builtins/builtins.c:269-345
 *   floatmap_alloc/copy ............. floatmap.c:30-60
 *   ORIG_VAL dispatch ............... opmacros.h:199-216
 */
#include "mmo_runtime.h"

mmo_image *mmo_floatmap_alloc(int width, int height, mmo_pools *pools) {
    mmo_image *img = (mmo_image *)mmo_pools_alloc(pools, sizeof(mmo_image));
    img->type = MMO_IMAGE_FLOATMAP;
    img->pixel_width = width;
    img->pixel_height = height;
    img->ax = img->bx = (float)(width - 1) / 2.0;
    img->ay = img->by = (float)(height - 1) / 2.0;
    img->ay *= -1.0;
    img->fdata = (float *)mmo_pools_alloc(pools, sizeof(float) * (size_t)width * height * NUM_FLOATMAP_CHANNELS);
    return img;
}

mmo_image *mmo_floatmap_copy(mmo_image *src, mmo_pools *pools) {
    mmo_image *copy = mmo_floatmap_alloc(src->pixel_width, src->pixel_height, pools);
    copy->ax = src->ax;
    copy->bx = src->bx;
    copy->ay = src->ay;
    copy->by = src->by;
    memcpy(copy->fdata, src->fdata, sizeof(float) * (size_t)src->pixel_width * src->pixel_height * NUM_FLOATMAP_CHANNELS);
    return copy;
}

mmo_image *mmo_make_resize_image(mmo_image *image, float x_factor, float y_factor, mmo_pools *pools) {
    mmo_image *r = (mmo_image *)mmo_pools_alloc(pools, sizeof(mmo_image));
    r->type = MMO_IMAGE_RESIZE;
    r->pixel_width = image->pixel_width;
    r->pixel_height = image->pixel_height;
    r->original = image;
    r->x_factor = x_factor;
    r->y_factor = y_factor;
    return r;
}

/* builtins/builtins.c:41-119.  One deliberate difference: a NaN or out-of-range coordinate arrives as INT_MIN (the x86
 * conversion), and the reference negates it / subtracts it from width - 1 -- signed overflow, undefined behaviour.  gcc -O2
 * concludes that -x % width lies in [0, width), drops the range check in get_pixel, and the reference reads a texel row in
 * front of the image (seen with Droste's NaN coordinates under the reflect / rotate modes).  No result can be "identical" to
 * that; the oracle and the CUDA path both let these two operations wrap, so INT_MIN stays out of range and yields the edge
 * colour (or texel 0 when width is a power of two). */
static int wrap_neg(int v) { return (int)(0u - (unsigned)v); }
static int wrap_sub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
static void apply_edge_behaviour(mmo_invocation *invocation, int *_x, int *_y, int width, int height) {
    int x = *_x, y = *_y;
    switch (invocation->edge_behaviour_x) {
    case EDGE_BEHAVIOUR_WRAP:
        if (x < 0) x = x % width + width;
        else if (x >= width) x %= width;
        break;
    case EDGE_BEHAVIOUR_REFLECT:
        if (x < 0) x = wrap_neg(x) % width;
        else if (x >= width) x = (width - 1) - (x % width);
        break;
    case EDGE_BEHAVIOUR_ROTATE:
        if (x < 0) { x = wrap_neg(x) % width; y = wrap_sub(height - 1, y); }
        else if (x >= width) { x = (width - 1) - (x % width); y = wrap_sub(height - 1, y); }
        break;
    default: break;
    }
    switch (invocation->edge_behaviour_y) {
    case EDGE_BEHAVIOUR_WRAP:
        if (y < 0) y = y % height + height;
        else if (y >= height) y %= height;
        break;
    case EDGE_BEHAVIOUR_REFLECT:
        if (y < 0) y = wrap_neg(y) % height;
        else if (y >= height) y = (height - 1) - (y % height);
        break;
    case EDGE_BEHAVIOUR_ROTATE:
        if (y < 0) { x = wrap_sub(width - 1, x); y = wrap_neg(y) % height; }
        else if (y >= height) { x = wrap_sub(width - 1, x); y = (height - 1) - (y % height); }
        break;
    default: break;
    }
    *_x = x;
    *_y = y;
}

static color_t get_pixel(mmo_invocation *invocation, int x, int y, mmo_image *drawable, int frame) {
    const unsigned char *p;
    if (drawable == NULL || drawable->data == NULL) return MAKE_RGBA_COLOR(255, 255, 255, 255);
    apply_edge_behaviour(invocation, &x, &y, drawable->pixel_width, drawable->pixel_height);
    if (x < 0 || x >= drawable->pixel_width) return invocation->edge_color_x;
    if (y < 0 || y >= drawable->pixel_height) return invocation->edge_color_y;
    if (frame < 0 || frame >= drawable->num_frames) return MAKE_RGBA_COLOR(255, 255, 255, 255);
    p = drawable->data + 4 * ((size_t)drawable->pixel_width * y + x);
    return MAKE_RGBA_COLOR(p[0], p[1], p[2], p[3]);
}

static mmo_image *get_image_drawable(mmo_image *image, float *x, float *y) {
    if (image == NULL || image->data == NULL) return NULL;
    *x = (*x + image->middle_x) * image->scale_x;
    *y = -((*y - image->middle_y) * image->scale_y);
    return image;
}

color_t mmo_get_orig_val_pixel(mmo_invocation *invocation, float x, float y, mmo_image *image, int frame) {
    mmo_image *drawable = get_image_drawable(image, &x, &y);
    if (!invocation->supersampling) {
        x += 0.5;
        y += 0.5;
    }
    return get_pixel(invocation, floor(x), floor(y), drawable, frame);
}

typedef struct { float red, green, blue, alpha; } float_color_t;
#define COLOR_MUL_FLOAT(c, f) ((float_color_t){RED((c)) * (f), GREEN((c)) * (f), BLUE((c)) * (f), ALPHA((c)) * (f)})
#define FLOAT_COLOR_ADD(a, b) ((float_color_t){(a).red + (b).red, (a).green + (b).green, (a).blue + (b).blue, (a).alpha + (b).alpha})
#define FLOAT_COLOR_TO_COLOR(fc) (MAKE_RGBA_COLOR(rintf((fc).red), rintf((fc).green), rintf((fc).blue), rintf((fc).alpha)))

color_t mmo_get_orig_val_intersample_pixel(mmo_invocation *invocation, float x, float y, mmo_image *image, int frame) {
    int x1, x2, y1, y2;
    float x2fact, y2fact, x1fact, y1fact, p1fact, p2fact, p3fact, p4fact;
    color_t pixel1, pixel2, pixel3, pixel4;
    float_color_t fpixel1, fpixel2, fpixel3, fpixel4, fresult;
    mmo_image *drawable = get_image_drawable(image, &x, &y);

    /* pixel_inc is 1 outside the GIMP fast preview (mathmap.c:1321-1328) */
    x1 = floor(x);
    x2 = x1 + 1;
    x2fact = x - x1;
    y1 = floor(y);
    y2 = y1 + 1;
    y2fact = y - y1;

    x1fact = 1.0 - x2fact;
    y1fact = 1.0 - y2fact;

    p1fact = x1fact * y1fact;
    p2fact = x1fact * y2fact;
    p3fact = x2fact * y1fact;
    p4fact = x2fact * y2fact;

    pixel1 = get_pixel(invocation, x1, y1, drawable, frame);
    pixel2 = get_pixel(invocation, x1, y2, drawable, frame);
    pixel3 = get_pixel(invocation, x2, y1, drawable, frame);
    pixel4 = get_pixel(invocation, x2, y2, drawable, frame);

    fpixel1 = COLOR_MUL_FLOAT(pixel1, p1fact);
    fpixel2 = COLOR_MUL_FLOAT(pixel2, p2fact);
    fpixel3 = COLOR_MUL_FLOAT(pixel3, p3fact);
    fpixel4 = COLOR_MUL_FLOAT(pixel4, p4fact);

    fresult = FLOAT_COLOR_ADD(fpixel1, fpixel2);
    fresult = FLOAT_COLOR_ADD(fresult, fpixel3);
    fresult = FLOAT_COLOR_ADD(fresult, fpixel4);

    return FLOAT_COLOR_TO_COLOR(fresult);
}

float *mmo_get_floatmap_pixel(mmo_invocation *invocation, mmo_image *image, float x, float y, float frame) {
    static float black[] = {0.0, 0.0, 0.0, 0.0};
    int ix, iy;
    (void)invocation;
    (void)frame;
    ix = (int)lrintf(image->ax * x + image->bx);
    iy = (int)lrintf(image->ay * y + image->by);
    if (ix < 0 || ix >= image->pixel_width || iy < 0 || iy >= image->pixel_height) return black;
    return image->fdata + ((size_t)iy * image->pixel_width + ix) * 4;
}

float *mmo_orig_val(mmo_invocation *invocation, float x, float y, mmo_image *img, float f, mmo_pools *pools, int force_nearest) {
    if (img->type == MMO_IMAGE_RESIZE) {
        x *= img->x_factor;
        y *= img->y_factor;
        img = img->original;
    }
    if (img->type == MMO_IMAGE_CLOSURE) return img->func(invocation, img, x, y, f, pools);
    if (img->type == MMO_IMAGE_FLOATMAP) return mmo_get_floatmap_pixel(invocation, img, x, y, f);
    {
        color_t color;
        invocation->taps++;
        /* the frame argument is a float converted to the int parameter of the sampler */
        if (invocation->antialiasing && !force_nearest) color = mmo_get_orig_val_intersample_pixel(invocation, x, y, img, (int)f);
        else color = mmo_get_orig_val_pixel(invocation, x, y, img, (int)f);
        return mmo_tuple_from_color(color, pools);
    }
}

mmo_image *mmo_render_image(mmo_invocation *invocation, mmo_image *image, int width, int height, mmo_pools *pools, int force) {
    mmo_image *new_image;
    if (!force && image->type == MMO_IMAGE_FLOATMAP) return image;
    new_image = mmo_floatmap_alloc(width, height, pools);
    if (image->type == MMO_IMAGE_CLOSURE) {
        /* frame 0, t 0.0, full region, no sampling offset (builtins.c:288-297) */
        image->calc_lines(invocation, image, 0, 0.0f, width, height, 0, 0, width, height, 0.0f, 0.0f, 0, height, new_image->fdata, 1);
    } else {
        float ax = new_image->ax, bx = new_image->bx, ay = new_image->ay, by = new_image->by;
        int x, y;
        float *p = new_image->fdata;
        mmo_pools filter_pools;
        mmo_pools_init(&filter_pools);
        for (y = 0; y < height; ++y) {
            float fy = ((float)y - by) / ay;
            for (x = 0; x < width; ++x) {
                float fx = ((float)x - bx) / ax;
                float *tuple;
                mmo_pools_reset(&filter_pools);
                /* the reference hard-wires the nearest sampler here (builtins.c:306) */
                tuple = mmo_orig_val(invocation, fx, fy, image, 0.0, &filter_pools, 1);
                memcpy(p, tuple, sizeof(float) * 4);
                p += 4;
            }
        }
        mmo_pools_free(&filter_pools);
    }
    return new_image;
}
