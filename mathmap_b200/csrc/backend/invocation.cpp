// The render half of the C ABI: invocations, the per-frame host replay of
// frame-constant IR (the reference's init_frame, new_template.c.in:314-337),
// render_image / native filters as kernel launches, and calc_lines as one
// pixel-grid launch per band (new_template.c.in:208-312).
#include "../runtime/mm_elliptic.h"
#include <cuda_runtime_api.h>

#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>

#include "../../../include/mathmap_b200.h"
#include "../ir/eval.h"
#include "../runtime/mm_types.h"
#include "kernels.h"
#include "nvrtc_module.h"

using namespace mm;
using namespace mmbackend;

namespace mmbackend {
long compile_only(mmb_module *m, const KernelConfig &cfg, std::string &err);
}

namespace {

struct Fail {
    std::string msg;
};
[[noreturn]] void fail(const std::string &m) { throw Fail{m}; }
void ck(cudaError_t e, const char *what) {
    if (e != cudaSuccess) fail(std::string(what) + ": " + cudaGetErrorString(e));
}

// Host-side value of a frame-constant IR value.
struct HVal {
    Type type = T_INT;
    int i = 0;
    float f = 0.f;
    std::complex<float> c;
    uint32_t color = 0;
    int image = -1;             // index into Invocation::images
    const void *ptr = nullptr;  // device pointer of a curve / gradient
    std::vector<float> tuple;
};

enum ImageKind { IMG_DRAWABLE = 0, IMG_FLOATMAP = 1, IMG_CLOSURE = 2 };

struct HostImage {
    int kind = IMG_DRAWABLE;
    void *data = nullptr;  // device
    bool owned = false;
    int w = 0, h = 0;
    int num_frames = 1;
    float sx = 0, sy = 0, mx = 1, my = 1;
    float ax = 0, bx = 0, ay = 0, by = 0;
    float xf = 1.f, yf = 1.f;
    bool resized = false;
    int original = -1;  // for resize wrappers: the wrapped image
    const Filter *filter = nullptr;
    std::vector<HVal> args;
    // A Gaussian blur (IIR) whose horizontal pass has not run yet: `data` holds the vertical pass's result.  The pass is
    // run in place by materialise() when anything reads the image, or -- when the frame's pixels ARE this image's
    // (pass-through, see render_slice) -- band by band straight into the caller's rows.
    bool pending_rows = false;
    float pending_sigma_h = 0.f;
};

struct Userval {
    int type = UV_INT;
    int i = 0;
    float f = 0;
    uint32_t color = 0;
    void *table = nullptr;  // device curve/gradient
    int image = -1;
};

struct FrameData {
    std::vector<unsigned char> uniforms;
    // image descriptor table of the launch: the kernel's own image uniforms sit at their compile-time slots, images that
    // only closures reach (their filters read the slot from their uniforms) are appended behind them
    mm_image slots[MM_MAX_IMAGES];
    int nslots = 0;
    int next_dynamic = 0;
    std::map<int, int> dynamic_slot_of;  // host image handle -> slot
};

}  // namespace

struct mmb_invocation {
    mmb_module *m = nullptr;
    std::shared_ptr<ModuleBackend> backend;
    int device = 0;
    int W = 0, H = 0;                // invocation->img_width / img_height
    int render_w = 0, render_h = 0;  // invocation->render_width / render_height (differ from W, H in a scaled preview)
    KernelConfig cfg;
    uint32_t edge_color_x = 0, edge_color_y = 0;
    int bpp = 4;
    cudaStream_t stream = nullptr;
    std::vector<Userval> uservals;
    std::vector<HostImage> images;
    size_t persistent_images = 0;  // images [0, persistent) survive across frames
    int frame = 0;
    float t = 0.f;
    bool frame_ready = false;
    FrameData main_frame;
    std::string main_spec;       // KernelConfig::spec of this frame's launches of the main filter
    int passthrough_image = -1;  // this frame's pixels are the floatmap images[passthrough_image], sampled at their own positions
    std::map<std::string, bool> identity_cache;
    long launches = 0;
    std::string kernel_name;
    std::map<std::string, float *> coord_cache;
    // Device allocation pool for per-frame temporaries.  Two lanes: frames rendered back to back on two alternating
    // streams (mmb_render_frames_device) each recycle only what their own stream used, so stream order alone makes the
    // reuse safe.  Everything else runs in lane 0.
    std::multimap<size_t, void *> free_blocks[2];
    std::vector<std::pair<size_t, void *>> frame_blocks[2];  // blocks in use by the lane's current frame
    int lane = 0;
    cudaStream_t lane_stream[2] = {nullptr, nullptr};
    cudaEvent_t lane_event[3] = {nullptr, nullptr, nullptr};
    std::map<std::string, int> native_cache;                // per frame: key -> image index
    void *staging = nullptr;
    size_t staging_bytes = 0;
    void *staging2 = nullptr;
    size_t staging2_bytes = 0;
    void *default_curve = nullptr, *default_gradient = nullptr;
    cudaStream_t copy_stream = nullptr;  // device->host copies overlap the next chunk's kernel
    cudaStream_t aux_stream = nullptr;   // odd chunks render here so one chunk's tail overlaps the next chunk's head
    cudaEvent_t order_event = nullptr;   // start of a chunked band (timed)
    // Ordering between the library's own stream (init_frame's renders, blurs and uploads) and a stream the caller
    // passes to the *_device entry points: the caller's stream waits for `init_event`, recorded on the library's
    // stream before the launch; what the library does next (recycling pool blocks, overwriting images) waits for
    // `user_event`, recorded on the caller's stream after the launch.
    cudaEvent_t init_event = nullptr, user_event = nullptr;
    bool user_event_pending = false;
    // cost of each chunk of the last chunked band (ms between consecutive completions) and the band it belongs to
    std::vector<float> chunk_cost;
    int cost_fr = -1, cost_lr = -1, cost_chunk_rows = -1, cost_floatmap = -1, cost_width = -1;
    std::vector<cudaEvent_t> chunk_events;

    void *alloc(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        auto &fb = free_blocks[lane];
        auto it = fb.lower_bound(bytes);
        if (it != fb.end() && it->first <= bytes + bytes / 4) {
            void *p = it->second;
            size_t sz = it->first;
            fb.erase(it);
            frame_blocks[lane].push_back({sz, p});
            return p;
        }
        void *p = nullptr;
        ck(cudaMalloc(&p, bytes), "cudaMalloc");
        frame_blocks[lane].push_back({bytes, p});
        return p;
    }
    void release_frame_blocks() {
        for (auto &b : frame_blocks[lane]) free_blocks[lane].insert(b);
        frame_blocks[lane].clear();
    }
    void *ensure_staging(void *&buf, size_t &cap, size_t bytes) {
        if (cap < bytes) {
            if (buf) cudaFree(buf);
            buf = nullptr;
            cap = 0;
            ck(cudaMalloc(&buf, bytes), "cudaMalloc(staging)");
            cap = bytes;
        }
        return buf;
    }
};

namespace {

// ------------------------------------------------------------ coordinate arrays
// CALC_VIRTUAL_X / CALC_VIRTUAL_Y (opmacros.h:156-157) in double on the host,
// narrowed to float, once per (size, offset); kernels read them instead of doing
// double divisions per pixel.
const float *coords(mmb_invocation *inv, bool is_y, int size, float offset, int count) {
    char key[96];
    snprintf(key, sizeof key, "%c%d_%a_%d", is_y ? 'y' : 'x', size, (double)offset, count);
    auto it = inv->coord_cache.find(key);
    if (it != inv->coord_cache.end()) return it->second;
    std::vector<float> v(count);
    for (int p = 0; p < count; ++p) {
        if (is_y) v[p] = (float)((-(p) + ((size)-1) / 2.0 - (offset)) / (((size)-1) / 2.0));
        else v[p] = (float)(((p) - ((size)-1) / 2.0 + (offset)) / (((size)-1) / 2.0));
    }
    float *d = nullptr;
    ck(cudaMalloc((void **)&d, sizeof(float) * count), "cudaMalloc(coords)");
    ck(cudaMemcpyAsync(d, v.data(), sizeof(float) * count, cudaMemcpyHostToDevice, inv->stream), "cudaMemcpy(coords)");
    ck(cudaStreamSynchronize(inv->stream), "sync(coords)");
    inv->coord_cache[key] = d;
    return d;
}

mm_image to_device_desc(const HostImage &h) {
    mm_image d;
    memset(&d, 0, sizeof d);
    d.data = h.data;
    d.kind = h.kind == IMG_FLOATMAP ? MM_IMAGE_FLOATMAP : MM_IMAGE_DRAWABLE;
    d.w = h.w; d.h = h.h; d.num_frames = h.num_frames;
    d.sx = h.sx; d.sy = h.sy; d.mx = h.mx; d.my = h.my;
    d.ax = h.ax; d.bx = h.bx; d.ay = h.ay; d.by = h.by;
    d.xf = h.xf; d.yf = h.yf;
    // the fast paths address texel y1 * w + x1 + 0x4B000000 in 32 bits relative to fast_base
    const bool fast = d.kind == MM_IMAGE_DRAWABLE && h.w > 0 && h.h > 0 && h.w < (1 << 22) && h.h < (1 << 22) &&
                      (unsigned long long)h.w * (unsigned long long)h.h < 3000000000ull;
    d.fast_base = (const unsigned *)((const char *)h.data - 4ll * 0x4B000000ll);
    d.fast_w = fast ? (float)h.w : -1.0f;
    d.fast_h = fast ? (float)h.h : -1.0f;
    d.fast_wm1 = fast ? (float)(h.w - 1) : -1.0f;
    d.fast_hm1 = fast ? (float)(h.h - 1) : -1.0f;
    d.fast_nf = fast ? (float)h.num_frames : -1.0f;
    return d;
}

// Runs the pending horizontal pass of a deferred Gaussian blur in place (see HostImage::pending_rows).
void materialise(mmb_invocation *inv, int image) {
    if (image < 0 || image >= (int)inv->images.size()) return;
    HostImage &im = inv->images[image];
    if (!im.pending_rows) return;
    void *scratch = inv->alloc(gauss_iir_scratch_bytes(im.w, im.h));
    launch_gauss_iir_rows((const float *)im.data, im.data, false, (long long)im.w * 16, (double *)scratch, im.w, 0, im.h, im.pending_sigma_h, inv->stream);
    ck(cudaGetLastError(), "gaussian blur (rows) launch");
    inv->launches++;
    im.pending_rows = false;
}

// Does get_floatmap_pixel (builtins.c:249-265) at the virtual coordinates of pixel (x, y) of a w x h frame read texel
// (x, y) of this floatmap, for every pixel?  Decided with the device's own float operations (mm_floatmap_pixel:
// rintf(ax * X + bx)) on the host's copy of the coordinate arrays (CALC_VIRTUAL_X/Y, see coords()).
bool floatmap_lookup_is_identity(mmb_invocation *inv, const HostImage &im, int w, int h) {
    if (im.kind != IMG_FLOATMAP || im.resized || im.xf != 1.f || im.yf != 1.f || im.w != w || im.h != h || w < 2 || h < 2) return false;
    char key[160];
    snprintf(key, sizeof key, "%d_%d_%a_%a_%a_%a", w, h, (double)im.ax, (double)im.bx, (double)im.ay, (double)im.by);
    auto it = inv->identity_cache.find(key);
    if (it != inv->identity_cache.end()) return it->second;
    bool ok = true;
    for (int x = 0; x < w && ok; ++x) {
        volatile float X = (float)(((x) - ((w)-1) / 2.0 + 0.0) / (((w)-1) / 2.0));
        volatile float p = im.ax * X;
        volatile float f = p + im.bx;
        ok = f > -2147483904.0f && f < 2147483648.0f && (int)rintf(f) == x;
    }
    for (int y = 0; y < h && ok; ++y) {
        volatile float Y = (float)((-(y) + ((h)-1) / 2.0 - 0.0) / (((h)-1) / 2.0));
        volatile float p = im.ay * Y;
        volatile float f = p + im.by;
        ok = f > -2147483904.0f && f < 2147483648.0f && (int)rintf(f) == y;
    }
    inv->identity_cache[key] = ok;
    return ok;
}

// -------------------------------------------------------------- host replay
struct Replay {
    mmb_invocation *inv;
    const Filter *filter;
    const FilterCode *code;
    const std::vector<HVal> &uservals;  // of this filter instance
    int frame;
    float t;
    int depth;
    std::map<const Value *, HVal> env;

    Replay(mmb_invocation *i, const Filter *f, const std::vector<HVal> &uv, int fr, float tt, int d)
        : inv(i), filter(f), code(i->m->code_for(f)), uservals(uv), frame(fr), t(tt), depth(d) {}

    static HVal from_const(const Const &c) {
        HVal v;
        v.type = c.type;
        v.i = c.i; v.f = c.f; v.c = c.c; v.color = c.color;
        return v;
    }
    static Const to_const(const HVal &v) {
        Const c;
        c.type = v.type;
        c.i = v.i; c.f = v.f; c.c = v.c; c.color = v.color;
        return c;
    }
    HVal prim(const Primary &p) {
        if (p.is_const) return from_const(p.c);
        if (p.value->index < 0) { HVal z; z.type = p.value->cv->type == T_FLOAT ? T_FLOAT : T_INT; return z; }
        auto it = env.find(p.value);
        if (it == env.end()) fail("internal error: frame-constant value used before definition");
        return it->second;
    }
    static float as_float(const HVal &v) { return v.type == T_INT ? (float)v.i : v.f; }
    static int as_int(const HVal &v) { return v.type == T_INT ? v.i : (int)v.f; }
    static bool truth(const HVal &v) { return v.type == T_INT ? v.i != 0 : v.f != 0.0f; }
    // C assignment into a variable of the compvar's type
    static HVal coerce(HVal v, Type t) {
        if (t == T_FLOAT && v.type == T_INT) { v.type = T_FLOAT; v.f = (float)v.i; }
        else if (t == T_COMPLEX && v.type == T_INT) { v.type = T_COMPLEX; v.c = {(float)v.i, 0.f}; }
        else if (t == T_COMPLEX && v.type == T_FLOAT) { v.type = T_COMPLEX; v.c = {v.f, 0.f}; }
        return v;
    }

    int add_image(const HostImage &img) {
        inv->images.push_back(img);
        return (int)inv->images.size() - 1;
    }

    // render_image (builtins/builtins.c:269-345)
    int render_image(int idx, int width, int height, bool force = false);
    int gaussian_blur(const std::vector<HVal> &args);
    int fft_native(const std::string &name, const std::vector<HVal> &args);

    HVal eval(const Rhs *r, const CompVar *dest) {
        switch (r->kind) {
        case RHS_PRIMARY: return prim(r->prim);
        case RHS_INTERNAL: {
            HVal v;
            const std::string &n = r->internal;
            if (n == "t") { v.type = T_FLOAT; v.f = t; }
            else if (n == "R") { v.type = T_FLOAT; v.f = (float)sqrt(2.0); }
            else if (n == "frame") { v.type = T_INT; v.i = frame; }
            else if (n == "__canvasPixelW") { v.type = T_INT; v.i = inv->W; }
            else if (n == "__canvasPixelH") { v.type = T_INT; v.i = inv->H; }
            else if (n == "__renderPixelW") { v.type = T_INT; v.i = inv->render_w; }
            else if (n == "__renderPixelH") { v.type = T_INT; v.i = inv->render_h; }
            else fail("internal " + n + " is not frame-constant");
            return v;
        }
        case RHS_TUPLE: {
            HVal v;
            v.type = T_TUPLE;
            for (auto &a : r->args) v.tuple.push_back(as_float(prim(a)));
            return v;
        }
        case RHS_CLOSURE: {
            std::vector<HVal> args;
            for (auto &a : r->args) args.push_back(prim(a));
            HVal v;
            v.type = T_IMAGE;
            if (r->filter->kind == FILTER_NATIVE) {
                if (r->filter->name == "gaussian_blur") v.image = gaussian_blur(args);
                else v.image = fft_native(r->filter->name, args);
                return v;
            }
            HostImage img;
            img.kind = IMG_CLOSURE;
            img.filter = r->filter;
            // callee uservals are typed by the callee's declaration
            for (size_t k = 0; k < args.size(); ++k)
                if (r->filter->uservals[k].type == UV_FLOAT) args[k] = coerce(args[k], T_FLOAT);
            img.args = args;
            img.w = inv->W;
            img.h = inv->H;
            v.image = add_image(img);
            return v;
        }
        case RHS_OP: break;
        default: fail("internal error: rhs kind not frame-constant");
        }
        const OpInfo *op = r->op;
        std::vector<HVal> a;
        for (auto &p : r->args) a.push_back(prim(p));
        HVal v;
        switch (op->id) {
        case OP_USERVAL_INT: case OP_USERVAL_BOOL: case OP_USERVAL_FLOAT: case OP_USERVAL_COLOR:
        case OP_USERVAL_CURVE: case OP_USERVAL_GRADIENT: case OP_USERVAL_IMAGE: {
            int idx = a[0].i;
            if (idx < 0 || idx >= (int)uservals.size()) fail("userval index out of range");
            return uservals[idx];
        }
        case OP_IMAGE_PIXEL_WIDTH: v.type = T_INT; v.i = inv->images.at(a[0].image).w; return v;
        case OP_IMAGE_PIXEL_HEIGHT: v.type = T_INT; v.i = inv->images.at(a[0].image).h; return v;
        case OP_RESIZE_IMAGE: {
            HostImage img = inv->images.at(a[0].image);
            img.owned = false;
            img.resized = true;
            img.original = a[0].image;
            img.xf = as_float(a[1]);
            img.yf = as_float(a[2]);
            v.type = T_IMAGE;
            v.image = add_image(img);
            return v;
        }
        case OP_STRIP_RESIZE: {
            const HostImage &img = inv->images.at(a[0].image);
            v.type = T_IMAGE;
            v.image = img.resized ? img.original : a[0].image;
            return v;
        }
        case OP_RENDER:
            v.type = T_IMAGE;
            v.image = render_image(a[0].image, as_int(a[1]), as_int(a[2]));
            return v;
        case OP_MAKE_RGBA_COLOR: {
            auto q = [](float x) {
                float cl = (0 < ((1 < x) ? 1 : x)) ? ((1 < x) ? 1 : x) : 0;
                return (uint32_t)(int)(cl * 255) & 0xffu;
            };
            v.type = T_COLOR;
            v.color = (q(as_float(a[0])) << 24) | (q(as_float(a[1])) << 16) | (q(as_float(a[2])) << 8) | q(as_float(a[3]));
            return v;
        }
        case OP_RED: v.type = T_FLOAT; v.f = (float)((a[0].color >> 24) / 255.0); return v;
        case OP_GREEN: v.type = T_FLOAT; v.f = (float)(((a[0].color >> 16) & 0xff) / 255.0); return v;
        case OP_BLUE: v.type = T_FLOAT; v.f = (float)(((a[0].color >> 8) & 0xff) / 255.0); return v;
        case OP_ALPHA: v.type = T_FLOAT; v.f = (float)((a[0].color & 0xff) / 255.0); return v;
        case OP_ELL_JAC: {  // opmacros.h:119-125
            double sn, cn, dn;
            mm_elljac((double)as_float(a[0]), (double)as_float(a[1]), &sn, &cn, &dn);
            v.type = T_TUPLE;
            v.tuple = {(float)sn, (float)cn, (float)dn};
            return v;
        }
        case OP_TUPLE_NTH: {
            int n = a[1].i;
            v.type = T_FLOAT;
            v.f = (n >= 0 && n < (int)a[0].tuple.size()) ? a[0].tuple[n] : 0.f;
            return v;
        }
        default: break;
        }
        Const args[6], out;
        for (int k = 0; k < op->nargs; ++k) args[k] = to_const(a[k]);
        if (!eval_op(op, args, &out)) fail(std::string("op ") + op->name + " cannot be evaluated on the host");
        (void)dest;
        return from_const(out);
    }

    void set(const Value *lhs, HVal v) { env[lhs] = coerce(std::move(v), lhs->cv->type); }

    void phis(const Stmt *p, int branch) {
        std::vector<std::pair<const Value *, HVal>> vals;
        for (; p; p = p->next)
            if (p->kind == ST_PHI && p->lhs->level == 0) vals.push_back({p->lhs, eval(branch == 0 ? p->rhs : p->rhs2, p->lhs->cv)});
        for (auto &kv : vals) set(kv.first, kv.second);
    }
    static bool has_level0(const Stmt *s) {
        for (; s; s = s->next) {
            switch (s->kind) {
            case ST_ASSIGN: case ST_PHI: if (s->lhs->level == 0) return true; break;
            case ST_IF: if (has_level0(s->cons) || has_level0(s->alt) || has_level0(s->exit)) return true; break;
            case ST_WHILE: if (s->level == 0 || has_level0(s->body)) return true; break;
            default: break;
            }
        }
        return false;
    }
    void run(const Stmt *s) {
        for (; s; s = s->next) {
            switch (s->kind) {
            case ST_ASSIGN:
                if (s->lhs->level == 0) set(s->lhs, eval(s->rhs, s->lhs->cv));
                break;
            case ST_IF:
                if (!(has_level0(s->cons) || has_level0(s->alt) || has_level0(s->exit))) break;
                if (s->level == 0) {
                    bool c = truth(eval(s->cond, nullptr));
                    run(c ? s->cons : s->alt);
                    phis(s->exit, c ? 0 : 1);
                } else {  // condition is per-pixel: only speculated pure definitions live here
                    run(s->cons);
                    run(s->alt);
                }
                break;
            case ST_WHILE:
                if (s->level == 0) {
                    phis(s->entry, 0);
                    long guard = 0;
                    while (truth(eval(s->cond, nullptr))) {
                        run(s->body);
                        phis(s->entry, 1);
                        if (++guard > 100000000L) fail("frame-constant loop does not terminate");
                    }
                } else
                    run(s->body);
                break;
            default: break;
            }
        }
    }
};

void pack_uniform_bytes(mmb_invocation *inv, const FilterKernel &k, Replay &rp, std::vector<unsigned char> &bytes, FrameData &fd, bool fixed_slots,
                        int depth);

// Device descriptor of a host image.  A closure (filter + argument values) becomes MM_IMAGE_CLOSURE: its filter's frame
// constants are computed by a host replay with the closure's arguments, like the reference's lazily initialised
// closure->xy_vars (new_template.c.in:389-398), packed as that filter's uniforms and uploaded; ORIG_VAL on it calls
// mm_closure_<filter> through mm_closure_dispatch.
mm_image device_desc(mmb_invocation *inv, FrameData &fd, int image, float t, int depth) {
    if (image < 0 || image >= (int)inv->images.size()) fail("internal error: bad image handle");
    if (image != inv->passthrough_image) materialise(inv, image);  // the pass-through image keeps its horizontal pass for the band launches
    if (inv->images[image].kind != IMG_CLOSURE) return to_device_desc(inv->images[image]);
    if (depth > 16) fail("closure nesting too deep");
    const HostImage img = inv->images[image];  // copy: the replay below may add images
    const FilterKernel &fk = inv->backend->source.kernels.at(img.filter);
    if (!fk.closure_fn) fail("internal error: filter " + img.filter->name + " has no closure entry on the device");
    const std::vector<HVal> args = img.args;
    Replay crp(inv, img.filter, args, 0, t, depth + 1);
    crp.run(crp.code->first);
    std::vector<unsigned char> blob;
    pack_uniform_bytes(inv, fk, crp, blob, fd, false, depth + 1);
    void *d = inv->alloc(std::max<size_t>(blob.size(), 8));
    // pageable source: the call returns once the bytes are staged for the DMA, so the local blob may go out of scope --
    // no stream synchronisation per closure and frame
    ck(cudaMemcpyAsync(d, blob.data(), blob.size(), cudaMemcpyHostToDevice, inv->stream), "cudaMemcpyAsync(closure uniforms)");
    mm_image desc;
    memset(&desc, 0, sizeof desc);
    desc.kind = MM_IMAGE_CLOSURE;
    desc.data = d;
    desc.closure_filter = fk.filter_index;
    desc.w = img.w; desc.h = img.h; desc.num_frames = 1;
    desc.xf = img.xf; desc.yf = img.yf;
    desc.fast_w = desc.fast_h = desc.fast_wm1 = desc.fast_hm1 = desc.fast_nf = -1.0f;
    return desc;
}

int dynamic_slot(mmb_invocation *inv, FrameData &fd, int image, float t, int depth) {
    auto it = fd.dynamic_slot_of.find(image);
    if (it != fd.dynamic_slot_of.end()) return it->second;
    if (fd.next_dynamic >= MM_MAX_IMAGES) fail("too many images reachable from one filter (max " + std::to_string(MM_MAX_IMAGES) + ")");
    const int slot = fd.next_dynamic++;
    fd.dynamic_slot_of[image] = slot;
    fd.nslots = std::max(fd.nslots, slot + 1);
    const mm_image d = device_desc(inv, fd, image, t, depth);  // may allocate further slots
    fd.slots[slot] = d;
    return slot;
}

void pack_uniform_bytes(mmb_invocation *inv, const FilterKernel &k, Replay &rp, std::vector<unsigned char> &bytes, FrameData &fd, bool fixed_slots,
                        int depth) {
    bytes.assign(k.uniforms_size, 0);
    for (const UniformField &u : k.uniforms) {
        auto it = rp.env.find(u.value);
        // a frame-constant value defined in the branch of a frame-constant `if` the host did not take
        // is never read by the device either (same branch): leave it zero
        if (it == rp.env.end()) continue;
        const HVal &v = it->second;
        unsigned char *dst = bytes.data() + u.offset;
        switch (u.type) {
        case T_INT: case T_NIL: { int x = Replay::as_int(v); memcpy(dst, &x, 4); break; }
        case T_FLOAT: { float x = Replay::as_float(v); memcpy(dst, &x, 4); break; }
        case T_COMPLEX: { float z[2] = {v.c.real(), v.c.imag()}; memcpy(dst, z, 8); break; }
        case T_COLOR: memcpy(dst, &v.color, 4); break;
        case T_CURVE: case T_GRADIENT: memcpy(dst, &v.ptr, 8); break;
        case T_IMAGE: {
            int slot;
            if (fixed_slots) {
                // every image-typed uniform of the launched kernel owns the slot the kernel was compiled with
                // (cuda_emit.cpp), so the generated code indexes P.images with a literal
                slot = u.image_slot;
                if (slot < 0 || slot >= MM_MAX_IMAGES) fail("internal error: image slot out of range");
                const mm_image d = device_desc(inv, fd, v.image, rp.t, depth);
                fd.slots[slot] = d;
                fd.nslots = std::max(fd.nslots, slot + 1);
            } else
                slot = dynamic_slot(inv, fd, v.image, rp.t, depth);
            memcpy(dst, &slot, 4);
            break;
        }
        case T_TUPLE: case T_TREE_VECTOR: {
            for (int i = 0; i < std::max(1, u.tuple_len); ++i) {
                float x = i < (int)v.tuple.size() ? v.tuple[i] : 0.f;
                memcpy(dst + 4 * i, &x, 4);
            }
            break;
        }
        default: fail("internal error: unsupported uniform type");
        }
    }
}

// Packs the uniforms a filter's kernel reads and fills the image descriptor table of its launches.
void pack_frame(mmb_invocation *inv, const FilterKernel &k, Replay &rp, FrameData &fd) {
    fd.nslots = 0;
    memset(fd.slots, 0, sizeof fd.slots);
    fd.dynamic_slot_of.clear();
    int fixed = 0;
    for (const UniformField &u : k.uniforms) fixed += u.type == T_IMAGE;
    fd.next_dynamic = fixed;
    fd.nslots = fixed;
    pack_uniform_bytes(inv, k, rp, fd.uniforms, fd, true, rp.depth);
}

struct LaunchGeom {
    int frame_w, frame_h;   // frame_render_width/height the coordinates refer to
    int region_x, region_w;
    int first_row, num_rows;
    float off_x, off_y;
    int xs_count, ys_count;
    int interleave = 1, phase = 0;  // 8-row block interleaving across ranks (see mm_actual_row)
    int row_limit = -1;
    const float *xs_dev = nullptr, *ys_dev = nullptr;  // explicit per-column / per-row coordinates instead of CALC_VIRTUAL_X/Y
};

void launch_filter(mmb_invocation *inv, const Filter *f, const FrameData &fd, const LaunchGeom &g, void *out, long long out_stride, int floatmap,
                   int frame, float t) {
    std::string err;
    KernelConfig cfg = inv->cfg;
    if (f == inv->m->main && cfg.specialize) cfg.spec = inv->main_spec;
    auto lm = inv->backend->get(cfg, inv->device, err);
    if (!lm) fail(err);
    DriverApi *api = driver_api(err);
    if (!api) fail(err);
    const FilterKernel &k = inv->backend->source.kernels.at(f);
    mm_params P;
    memset(&P, 0, sizeof P);
    P.out = out;
    P.out_stride = out_stride;
    P.xs = g.xs_dev ? g.xs_dev : coords(inv, false, g.frame_w, g.off_x, g.xs_count);
    P.ys = g.ys_dev ? g.ys_dev : coords(inv, true, g.frame_h, g.off_y, g.ys_count);
    P.region_x = g.region_x;
    P.region_y = 0;
    P.region_w = g.region_w;
    P.first_row = g.first_row;
    P.num_rows = g.num_rows;
    P.row_interleave = g.interleave;
    P.row_phase = g.phase;
    P.row_limit = g.row_limit >= 0 ? g.row_limit : g.first_row + g.num_rows;
    P.img_w = inv->W; P.img_h = inv->H; P.render_w = inv->render_w; P.render_h = inv->render_h;
    P.frame = frame;
    P.t = t;
    P.R = (float)sqrt(2.0);
    P.bpp = inv->bpp;
    P.floatmap = floatmap;
    P.out_mode = floatmap ? 1 : (inv->bpp == 4 ? 0 : 2);
    P.magic23 = 0x4B000000u;
    P.pk_neg_zero = 0x8000000080000000ull;
    P.pk_one = 0x3f8000003f800000ull;
    P.pk_magic_round = 0x4b4000004b400000ull;
    P.edge_color_x = inv->edge_color_x;
    P.edge_color_y = inv->edge_color_y;
    for (int i = 0; i < fd.nslots; ++i) P.images[i] = fd.slots[i];
    // row pre-kernel: row-constant values are computed once per row into 4-byte arrays the pixel kernel reads
    std::vector<void *> rowvals((size_t)std::max(1, k.row_slots), nullptr);
    void *params[3] = {&P, (void *)fd.uniforms.data(), (void *)rowvals.data()};
    const int tile_w = k.quad ? 128 : 32;
    unsigned gx = (unsigned)((g.region_w + tile_w - 1) / tile_w), gy;
    P.vec_store = ((uintptr_t)out & 15) == 0 && (out_stride & 15) == 0;
    {
        // Tiles per block: as many as the kernel takes (k.auto_rows) while the grid still has about six waves of
        // blocks (8 resident blocks per SM), so that the tail of a small frame stays short.  Measured at 8192^2 for
        // 4 -> 8 tiles: Ident 0.295 -> 0.283 ms, Invert 0.255 -> 0.235 ms; at 3840x2160 (Sea): 11.7 -> 12.6 ms.
        int rows = k.auto_rows > 1 ? inv->cfg.rows : 1;  // kernels with per-pixel loops are compiled for one tile per block
        if (rows == 0) {
            static int sm_count[64];  // per device ordinal, 0 = not asked yet
            int &sms = sm_count[inv->device & 63];
            if (sms == 0 && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, inv->device) != cudaSuccess) sms = 148;
            const long long want = 6ll * 8 * sms;
            rows = k.auto_rows;
            while (rows > 1 && (long long)gx * ((g.num_rows + 8 * rows - 1) / (8 * rows)) < want) rows /= 2;
        }
        P.rows = rows;
        gy = (unsigned)((g.num_rows + 8 * rows - 1) / (8 * rows));
    }
    if (gx == 0 || gy == 0) return;
    auto rit = lm->row_functions.find(f);
    if (rit != lm->row_functions.end()) {
        size_t per = ((size_t)g.num_rows * 4 + 255) & ~(size_t)255;
        char *buf = (char *)inv->alloc(per * (size_t)k.row_slots);
        for (int i = 0; i < k.row_slots; ++i) rowvals[i] = buf + per * (size_t)i;
        int rrc = api->launch_kernel(rit->second, (unsigned)((g.num_rows + 255) / 256), 1, 1, 256, 1, 1, 0, inv->stream, params, nullptr);
        if (rrc != 0) fail("cuLaunchKernel(" + k.row_kernel_name + ") failed: " + api->error_string(rrc));
        inv->launches++;
    }
    int rc = api->launch_kernel(lm->functions.at(f), gx, gy, 1, 256, 1, 1, 0, inv->stream, params, nullptr);
    if (rc != 0) fail("cuLaunchKernel(" + k.kernel_name + ") failed: " + api->error_string(rc));
    inv->launches++;
    inv->kernel_name = k.kernel_name;
    if (lm->call_overflow) {  // modules with filter calls (recursion) only: did a call nest deeper than the device stack allows?
        int flag = 0;
        cudaError_t e = cudaMemcpyAsync(&flag, lm->call_overflow, sizeof flag, cudaMemcpyDeviceToHost, inv->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(inv->stream);
        if (e != cudaSuccess) fail(std::string("filter kernel failed: ") + cudaGetErrorString(e));
        if (flag) {
            cudaMemsetAsync(lm->call_overflow, 0, sizeof flag, inv->stream);
            fail("filter " + f->name + ": calls nested deeper than " + std::to_string(MM_MAX_CALL_DEPTH) + " levels (the device stack holds that many frames)");
        }
    }
}

int Replay::render_image(int idx, int width, int height, bool force) {
    materialise(inv, idx);
    HostImage src = inv->images.at(idx);
    // a RESIZE wrapper is not a floatmap to the reference (its type is IMAGE_RESIZE, builtins.c:275): it is resampled
    // through ORIG_VAL with its factors, and the result carries none
    if (src.kind == IMG_FLOATMAP && !src.resized && !force) return idx;
    HostImage out;
    out.kind = IMG_FLOATMAP;
    out.w = width;
    out.h = height;
    // floatmap_alloc, floatmap.c:30-47
    out.ax = out.bx = (float)((float)(width - 1) / 2.0);
    out.ay = out.by = (float)((float)(height - 1) / 2.0);
    out.ay = (float)(out.ay * -1.0);
    out.data = inv->alloc(sizeof(float) * 4 * (size_t)width * height);
    if (src.kind == IMG_CLOSURE) {
        if (depth > 16) fail("render() nesting too deep");
        // a new frame for the closure: frame 0, t = 0.0 (builtins.c:288)
        Replay sub(inv, src.filter, src.args, 0, 0.0f, depth + 1);
        sub.run(sub.code->first);
        const FilterKernel &k = inv->backend->source.kernels.at(src.filter);
        FrameData fd;
        pack_frame(inv, k, sub, fd);
        LaunchGeom g{width, height, 0, width, 0, height, 0.f, 0.f, width, height};
        if (src.resized) {
            // A resize wrapper around the closure (the caller's pixel-size factors): the reference then does not run the
            // closure's calc_lines but samples it pixel by pixel through ORIG_VAL at fx = ((float)x - bx) / ax scaled by
            // the factors (builtins.c:303-342, opmacros.h:203-207).  Same kernel, those coordinates instead of
            // CALC_VIRTUAL_X/Y.
            std::vector<float> c((size_t)width + height);
            for (int x = 0; x < width; ++x) c[x] = (((float)x - out.bx) / out.ax) * src.xf;
            for (int y = 0; y < height; ++y) c[(size_t)width + y] = (((float)y - out.by) / out.ay) * src.yf;
            float *d = (float *)inv->alloc(sizeof(float) * c.size());
            ck(cudaMemcpyAsync(d, c.data(), sizeof(float) * c.size(), cudaMemcpyHostToDevice, inv->stream), "cudaMemcpyAsync(coords)");  // pageable: staged on return
            g.xs_dev = d;
            g.ys_dev = d + width;
        }
        launch_filter(inv, src.filter, fd, g, out.data, (long long)sizeof(float) * 4 * width, 1, 0, 0.0f);
    } else if (src.kind == IMG_FLOATMAP) {
        // forced re-render of a floatmap: nearest lookup through get_floatmap_pixel (builtins.c:303-342)
        launch_floatmap_resample((const float *)src.data, src.w, src.h, src.ax, src.bx, src.ay, src.by, src.xf, src.yf, (float *)out.data, width, height,
                                 out.ax, out.bx, out.ay, out.by, inv->stream);
        inv->launches++;
    } else {
        mm_image d = to_device_desc(src);
        launch_drawable_to_floatmap(d, (float *)out.data, width, height, out.ax, out.bx, out.ay, out.by, inv->cfg.edge_x, inv->cfg.edge_y,
                                    inv->edge_color_x, inv->edge_color_y, inv->cfg.supersampling, inv->stream);
        inv->launches++;
    }
    return add_image(out);
}

// native_filter_convolve / half_convolve / visualize_fft, native-filters/convolve.c:69-357
int Replay::fft_native(const std::string &name, const std::vector<HVal> &args) {
    std::string key = name;
    for (auto &a : args) {
        char b[64];
        snprintf(b, sizeof b, ":%d:%d:%a", a.image, a.i, (double)a.f);
        key += b;
    }
    auto it = inv->native_cache.find(key);
    if (it != inv->native_cache.end()) return it->second;
    int in = args[0].image;
    materialise(inv, in);
    if (args.size() > 1 && args[1].image >= 0) materialise(inv, args[1].image);
    {
        const HostImage &i0 = inv->images.at(in);  // convolve.c:88 always re-renders; skipped where that is the identity
        if (i0.kind != IMG_FLOATMAP || i0.resized || i0.w != inv->render_w || i0.h != inv->render_h) in = render_image(in, inv->render_w, inv->render_h, true);
    }
    HostImage src = inv->images.at(in);
    HostImage out;
    out.kind = IMG_FLOATMAP;
    out.w = src.w;
    out.h = src.h;
    out.ax = out.bx = (float)((float)(src.w - 1) / 2.0);
    out.ay = out.by = (float)((float)(src.h - 1) / 2.0);
    out.ay = (float)(out.ay * -1.0);
    out.data = inv->alloc(sizeof(float) * 4 * (size_t)src.w * src.h);
    std::string err;
    bool ok;
    if (name == "visualize_fft") {
        ok = fft_visualize((const float *)src.data, (float *)out.data, src.w, src.h, as_int(args[1]) != 0, inv->stream, &inv->launches, err);
    } else if (name == "convolve" || name == "half_convolve") {
        int fi = args[1].image;
        const HostImage &f0 = inv->images.at(fi);
        if (f0.kind != IMG_FLOATMAP || f0.resized || f0.w != src.w || f0.h != src.h) fi = render_image(fi, src.w, src.h, true);
        HostImage filt = inv->images.at(fi);
        if (name == "convolve")
            ok = fft_convolve((const float *)src.data, (const float *)filt.data, (float *)out.data, src.w, src.h, as_int(args[2]) != 0, as_int(args[3]) != 0,
                              inv->stream, &inv->launches, err);
        else
            ok = fft_half_convolve((const float *)src.data, (const float *)filt.data, (float *)out.data, src.w, src.h, as_int(args[2]) != 0, inv->stream,
                                   &inv->launches, err);
    } else
        fail("native filter " + name + " is not implemented");
    if (!ok) fail(err.empty() ? "FFT native filter failed" : err);
    int idx = add_image(out);
    inv->native_cache[key] = idx;
    return idx;
}

// native_filter_gaussian_blur, native-filters/gauss.c:643-670
int Replay::gaussian_blur(const std::vector<HVal> &args) {
    int in = args[0].image;
    float h = as_float(args[1]), v = as_float(args[2]);
    char key[128];
    snprintf(key, sizeof key, "gauss:%d:%a:%a", in, (double)h, (double)v);
    auto it = inv->native_cache.find(key);
    if (it != inv->native_cache.end()) return it->second;
    materialise(inv, in);
    // A drawable input whose blur takes the IIR path is not rendered to a floatmap first: every channel of that
    // floatmap would be k/255 of a byte, so the column pass reads RGBA8 (the drawable itself when render_image is the
    // identity on texels, else a resampled RGBA8 copy) and converts on load; results are the same floats.
    const HostImage in_img = inv->images.at(in);
    if (in_img.kind == IMG_DRAWABLE) {
        const int w = inv->render_w, hgt = inv->render_h;  // gauss.c:657
        const float ax = (float)((float)(w - 1) / 2.0), ay = (float)((float)(hgt - 1) / 2.0) * -1.0f;  // floatmap_alloc, floatmap.c:30-47
        const float bx = ax, by = (float)((float)(hgt - 1) / 2.0);
        const float sh = (float)fabs((double)(h * ax)), sv = (float)fabs((double)(v * ay));
        if (!(sh < 0.5f || sv < 0.5f)) {
            HostImage out;
            out.kind = IMG_FLOATMAP;
            out.w = w;
            out.h = hgt;
            out.ax = ax; out.bx = bx; out.ay = ay; out.by = by;
            out.data = inv->alloc(sizeof(float) * 4 * (size_t)w * hgt);
            mm_image d = to_device_desc(in_img);
            const void *bytes_in = in_img.data;
            if (!drawable_render_is_identity(d, w, hgt, ax, bx, ay, by, inv->cfg.supersampling)) {
                void *resampled = inv->alloc((size_t)w * hgt * 4);
                launch_drawable_to_bytes(d, resampled, w, hgt, ax, bx, ay, by, inv->cfg.edge_x, inv->cfg.edge_y, inv->edge_color_x, inv->edge_color_y,
                                         inv->cfg.supersampling, inv->stream);
                inv->launches++;
                bytes_in = resampled;
            }
            // the vertical pass now, the horizontal one when the image is first read (materialise) or, band by band, as
            // the frame's own pixels (render_slice)
            void *scratch = inv->alloc(gauss_iir_scratch_bytes(w, hgt));
            launch_gauss_iir_columns(bytes_in, true, (float *)out.data, (double *)scratch, w, hgt, sv, inv->stream);
            inv->launches++;
            out.pending_rows = true;
            out.pending_sigma_h = sh;
            ck(cudaGetLastError(), "gaussian blur launch");
            int idx = add_image(out);
            inv->native_cache[key] = idx;
            return idx;
        }
    }
    int fm = in;
    if (inv->images.at(in).kind != IMG_FLOATMAP || inv->images.at(in).resized) fm = render_image(in, inv->render_w, inv->render_h);
    HostImage src = inv->images.at(fm);
    float sh = (float)fabs((double)(h * src.ax)), sv = (float)fabs((double)(v * src.ay));
    HostImage out = src;
    out.owned = false;
    out.xf = out.yf = 1.f;  // a native filter's result is a plain floatmap (gauss.c:147 floatmap_copy)
    out.resized = false;
    out.original = -1;
    out.pending_rows = false;
    size_t bytes = sizeof(float) * 4 * (size_t)src.w * src.h;
    out.data = inv->alloc(bytes);
    if (sh < 0.5f || sv < 0.5f) {
        void *tmp = inv->alloc(bytes);
        launch_gauss_rle((const float *)src.data, (float *)tmp, (float *)out.data, src.w, src.h, sh, sv, inv->alloc(gauss_rle_curve_bytes(sh, sv)), inv->stream);
        inv->launches += 2;
    } else {
        void *scratch = inv->alloc(gauss_iir_scratch_bytes(src.w, src.h));
        launch_gauss_iir_columns(src.data, false, (float *)out.data, (double *)scratch, src.w, src.h, sv, inv->stream);
        inv->launches++;
        out.pending_rows = true;
        out.pending_sigma_h = sh;
    }
    ck(cudaGetLastError(), "gaussian blur launch");
    int idx = add_image(out);
    inv->native_cache[key] = idx;
    return idx;
}

std::vector<HVal> main_uservals(mmb_invocation *inv) {
    std::vector<HVal> out;
    for (auto &u : inv->uservals) {
        HVal v;
        switch (u.type) {
        case UV_INT: case UV_BOOL: v.type = T_INT; v.i = u.i; break;
        case UV_FLOAT: v.type = T_FLOAT; v.f = u.f; break;
        case UV_COLOR: v.type = T_COLOR; v.color = u.color; break;
        case UV_CURVE: v.type = T_CURVE; v.ptr = u.table ? u.table : inv->default_curve; break;
        case UV_GRADIENT: v.type = T_GRADIENT; v.ptr = u.table ? u.table : inv->default_gradient; break;
        case UV_IMAGE: v.type = T_IMAGE; v.image = u.image; break;
        }
        out.push_back(v);
    }
    return out;
}

void set_device(mmb_invocation *inv) { ck(cudaSetDevice(inv->device), "cudaSetDevice"); }

template <class F> int guarded(F f) {
    try {
        f();
        return 0;
    } catch (Fail &e) {
        set_error(e.msg);
    } catch (mm::CompileError &e) {
        set_error(e.message);
    } catch (std::exception &e) {
        set_error(std::string("internal error: ") + e.what());
    }
    return -1;
}

// Makes the library's stream wait for work the caller's stream still has in flight on this invocation's buffers.
void wait_for_user_stream(mmb_invocation *inv) {
    if (!inv->user_event_pending) return;
    ck(cudaStreamWaitEvent(inv->stream, inv->user_event, 0), "cudaStreamWaitEvent");
    inv->user_event_pending = false;
}

// Launches of one entry point go to `user` (the caller's stream) instead of the library's own: the caller's stream
// first waits for everything init_frame queued, and the library's later work waits for these launches (done()).
struct StreamScope {
    mmb_invocation *inv;
    cudaStream_t saved;
    bool switched = false;
    StreamScope(mmb_invocation *i, cudaStream_t user) : inv(i), saved(i->stream) {
        if (!user || user == saved) return;
        if (!inv->init_event) ck(cudaEventCreateWithFlags(&inv->init_event, cudaEventDisableTiming), "cudaEventCreate");
        if (!inv->user_event) ck(cudaEventCreateWithFlags(&inv->user_event, cudaEventDisableTiming), "cudaEventCreate");
        ck(cudaEventRecord(inv->init_event, saved), "cudaEventRecord");
        ck(cudaStreamWaitEvent(user, inv->init_event, 0), "cudaStreamWaitEvent");
        inv->stream = user;
        switched = true;
    }
    void done() {
        if (!switched) return;
        ck(cudaEventRecord(inv->user_event, inv->stream), "cudaEventRecord");
        inv->user_event_pending = true;
    }
    ~StreamScope() { inv->stream = saved; }
};

// calc_lines_<f> for one slice (new_template.c.in:208-312): rows [max(0, first_row), min(last_row, region_y + region_height))
// and columns [region_x, region_x + region_width) of a frame_w x frame_h frame sampled at the slice's offsets; row r of
// the band starts at dev_out + (r - first_row) * out_stride, column region_x at its first byte.
struct SliceGeom {
    int frame_w, frame_h;
    int region_x, region_y, region_w, region_h;
    float off_x, off_y;
};
void render_slice(mmb_invocation *inv, const SliceGeom &sl, int first_row, int last_row, void *dev_out, long long out_stride, int floatmap) {
    if (!inv->frame_ready) fail("mmb_init_frame must be called before mmb_calc_lines");
    if (sl.frame_w <= 0 || sl.frame_h <= 0 || sl.region_w < 0 || sl.region_h < 0 || sl.region_x < 0 || sl.region_y < 0) fail("bad slice geometry");
    first_row = std::max(0, first_row);
    last_row = std::min(last_row, sl.region_y + sl.region_h);
    if (last_row <= first_row || sl.region_w == 0) return;
    if (inv->passthrough_image >= 0) {
        HostImage &im = inv->images[inv->passthrough_image];
        if (im.pending_rows && sl.region_x == 0 && sl.region_w == im.w && sl.frame_w == im.w && sl.frame_h == im.h && sl.off_x == 0.f && sl.off_y == 0.f &&
            last_row <= im.h && (floatmap || inv->bpp == 4)) {
            // whole rows of the frame at the frame's own sampling positions: the blur's horizontal pass over these rows
            // writes them (quantised like mm_store_pixel, or as floats); chunks may run on different streams, so each
            // launch has checkpoint memory of its own
            const int rows = last_row - first_row;
            void *scratch = inv->alloc(gauss_iir_scratch_bytes(im.w, rows));
            launch_gauss_iir_rows((const float *)im.data, dev_out, !floatmap, out_stride, (double *)scratch, im.w, first_row, rows, im.pending_sigma_h,
                                  inv->stream);
            ck(cudaGetLastError(), "gaussian blur (rows) launch");
            inv->launches++;
            inv->kernel_name = "gauss_iir_rows";
            return;
        }
        materialise(inv, inv->passthrough_image);  // anything else samples the finished floatmap through the pixel kernel
    }
    // the coordinate arrays cover every column / row index the launch reads (a slice may reach past the frame:
    // the supersampling slice is one column wider, mathmap_common.c:892)
    LaunchGeom g{sl.frame_w, sl.frame_h, sl.region_x, sl.region_w, first_row, last_row - first_row, sl.off_x, sl.off_y,
                 std::max(sl.frame_w, sl.region_x + sl.region_w) + 1, std::max(sl.frame_h, last_row) + 1};
    launch_filter(inv, inv->m->main, inv->main_frame, g, dev_out, out_stride, floatmap, inv->frame, inv->t);
}

// call_invocation (mathmap_common.c:874-936) for the band [first_row, last_row) of the whole frame: one slice, or with
// supersampling the short slice plus a one-column-wider slice sampled at (-0.5, -0.5), combined
// (l1[c] + l1[c+1] + 2 l2[c] + l3[c] + l3[c+1]) / 6.  The long slice has the band's region_height, so the row below
// the band's last one is clamped away and l3 repeats l1 there -- per band, exactly like the reference's threads.
void render_band(mmb_invocation *inv, int first_row, int last_row, void *dev_out, int floatmap) {
    if (!inv->frame_ready) fail("mmb_init_frame must be called before mmb_calc_lines");
    first_row = std::max(0, first_row);
    last_row = std::min(last_row, inv->render_h);
    if (last_row <= first_row) return;
    const int W = inv->render_w, H = inv->render_h;
    int rows = last_row - first_row;
    if (inv->cfg.supersampling && !floatmap) {
        int bpp = inv->bpp;
        void *s = inv->ensure_staging(inv->staging2, inv->staging2_bytes, (size_t)rows * W * bpp + (size_t)rows * (W + 1) * bpp + 512);
        unsigned char *shortimg = (unsigned char *)s;
        unsigned char *longimg = shortimg + (((size_t)rows * W * bpp + 255) & ~(size_t)255);
        render_slice(inv, SliceGeom{W, H, 0, first_row, W, rows, 0.f, 0.f}, first_row, last_row, shortimg, (long long)W * bpp, 0);
        render_slice(inv, SliceGeom{W, H, 0, first_row, W + 1, rows, -0.5f, -0.5f}, first_row, last_row, longimg, (long long)(W + 1) * bpp, 0);
        launch_supersample_combine(shortimg, longimg, (unsigned char *)dev_out, W, rows, rows, bpp, inv->stream);
        inv->launches++;
        ck(cudaGetLastError(), "supersample combine");
        return;
    }
    long long stride = floatmap ? (long long)sizeof(float) * 4 * W : (long long)W * inv->bpp;
    render_slice(inv, SliceGeom{W, H, 0, first_row, W, rows, 0.f, 0.f}, first_row, last_row, dev_out, stride, floatmap);
}

}  // namespace

// ------------------------------------------------------------------ C ABI
extern "C" {

const char *mmb_module_cuda_source(mmb_module *m) {
    if (!m) return nullptr;
    int rc = guarded([&] { get_module_backend(m); });
    return rc == 0 ? m->cuda_source.c_str() : nullptr;
}

// Build check without a GPU: NVRTC-compiles the module for sm_100a; returns cubin bytes or -1.
int mmb_set_cubin_cache_dir(const char *dir) {
    set_cubin_cache_dir(dir);
    return 0;
}

long mmb_module_compile_check(mmb_module *m, int antialiasing, int precise_math) {
    if (!m) return -1;
    KernelConfig cfg;
    cfg.aa = antialiasing;
    cfg.precise = precise_math;
    std::string err;
    long n = compile_only(m, cfg, err);
    if (n < 0) set_error(err);
    return n;
}

long mmb_module_compile_check_fast(mmb_module *m, int antialiasing, int precise_math) {
    if (!m) return -1;
    KernelConfig cfg;
    cfg.aa = antialiasing;
    cfg.precise = precise_math;
    cfg.fast_compile = 1;
    std::string err;
    long n = compile_only(m, cfg, err);
    if (n < 0) set_error(err);
    return n;
}

mmb_invocation *mmb_invoke(mmb_module *m, int img_width, int img_height, int device) {
    if (!m || img_width <= 0 || img_height <= 0) { set_error("mmb_invoke: bad arguments"); return nullptr; }
    auto inv = new mmb_invocation();
    int rc = guarded([&] {
        inv->m = m;
        inv->device = device;
        inv->W = inv->render_w = img_width;
        inv->H = inv->render_h = img_height;
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            fail(std::string("no CUDA device available (") + cudaGetErrorString(e) + "); mathmap_b200 has no CPU path");
        if (device < 0 || device >= count) fail("CUDA device ordinal out of range");
        set_device(inv);
        ck(cudaFree(nullptr), "CUDA context creation");
        inv->backend = get_module_backend(m);
        for (auto &u : m->main->uservals) {
            Userval v;
            v.type = u.type;
            switch (u.type) {
            case UV_INT: v.i = u.int_default; break;
            case UV_FLOAT: v.f = u.float_default; break;
            case UV_BOOL: v.i = u.bool_default; break;
            case UV_COLOR: v.color = 0x000000ffu; break;  // opaque black
            default: break;
            }
            inv->uservals.push_back(v);
        }
        // default curve: identity; default gradient: gray ramp (userval.c:281-309, mathmap.c:356-361)
        std::vector<float> curve(MMB_CURVE_POINTS);
        std::vector<uint32_t> grad(MMB_CURVE_POINTS);
        for (int i = 0; i < MMB_CURVE_POINTS; ++i) {
            curve[i] = (float)i / (float)(MMB_CURVE_POINTS - 1);
            uint32_t g = (uint32_t)(int)((double)i / (MMB_CURVE_POINTS - 1) * 255.0);
            grad[i] = (g << 24) | (g << 16) | (g << 8) | 255u;
        }
        ck(cudaMalloc(&inv->default_curve, sizeof(float) * MMB_CURVE_POINTS), "cudaMalloc");
        ck(cudaMalloc(&inv->default_gradient, sizeof(uint32_t) * MMB_CURVE_POINTS), "cudaMalloc");
        ck(cudaMemcpy(inv->default_curve, curve.data(), sizeof(float) * MMB_CURVE_POINTS, cudaMemcpyHostToDevice), "cudaMemcpy");
        ck(cudaMemcpy(inv->default_gradient, grad.data(), sizeof(uint32_t) * MMB_CURVE_POINTS, cudaMemcpyHostToDevice), "cudaMemcpy");
    });
    if (rc != 0) { delete inv; return nullptr; }
    return inv;
}

void mmb_invocation_free(mmb_invocation *inv) {
    if (!inv) return;
    cudaSetDevice(inv->device);
    if (inv->user_event_pending) cudaEventSynchronize(inv->user_event);
    cudaStreamSynchronize(inv->stream);
    for (int l = 0; l < 2; ++l) {
        if (inv->lane_stream[l]) { cudaStreamSynchronize(inv->lane_stream[l]); cudaStreamDestroy(inv->lane_stream[l]); }
        inv->lane = l;
        inv->release_frame_blocks();
        for (auto &b : inv->free_blocks[l]) cudaFree(b.second);
    }
    for (auto e : inv->lane_event)
        if (e) cudaEventDestroy(e);
    for (auto &img : inv->images)
        if (img.owned && img.data) cudaFree(img.data);
    for (auto &u : inv->uservals)
        if (u.table) cudaFree(u.table);
    for (auto &c : inv->coord_cache) cudaFree(c.second);
    for (auto e : inv->chunk_events) cudaEventDestroy(e);
    if (inv->aux_stream) cudaStreamDestroy(inv->aux_stream);
    if (inv->order_event) cudaEventDestroy(inv->order_event);
    if (inv->init_event) cudaEventDestroy(inv->init_event);
    if (inv->user_event) cudaEventDestroy(inv->user_event);
    if (inv->copy_stream) cudaStreamDestroy(inv->copy_stream);
    if (inv->staging) cudaFree(inv->staging);
    if (inv->staging2) cudaFree(inv->staging2);
    if (inv->default_curve) cudaFree(inv->default_curve);
    if (inv->default_gradient) cudaFree(inv->default_gradient);
    delete inv;
}

static bool null_inv(const mmb_invocation *inv, const char *fn) {
    if (inv) return false;
    set_error(std::string(fn) + ": NULL invocation");
    return true;
}
int mmb_set_antialiasing(mmb_invocation *inv, int enabled) {
    if (null_inv(inv, "mmb_set_antialiasing")) return -1;
    inv->cfg.aa = enabled ? 1 : 0;
    return 0;
}
int mmb_set_supersampling(mmb_invocation *inv, int enabled) {
    if (null_inv(inv, "mmb_set_supersampling")) return -1;
    inv->cfg.supersampling = enabled ? 1 : 0;
    return 0;
}
int mmb_set_precise_math(mmb_invocation *inv, int enabled) {
    if (null_inv(inv, "mmb_set_precise_math")) return -1;
    inv->cfg.precise = enabled ? 1 : 0;
    return 0;
}
// invocation->render_width / render_height: the size frames are rendered at when it differs from the image size (the
// GIMP preview, mathmap.c:2191-2223); natives render their intermediates at this size (gauss.c:657, convolve.c:88)
int mmb_set_render_size(mmb_invocation *inv, int render_width, int render_height) {
    if (null_inv(inv, "mmb_set_render_size")) return -1;
    if (render_width <= 0 || render_height <= 0) { set_error("mmb_set_render_size: bad size"); return -1; }
    inv->render_w = render_width;
    inv->render_h = render_height;
    inv->frame_ready = false;
    return 0;
}
int mmb_set_specialize(mmb_invocation *inv, int enabled) {
    if (!inv) return -1;
    inv->cfg.specialize = enabled ? 1 : 0;
    return 0;
}
int mmb_set_fast_compile(mmb_invocation *inv, int enabled) {
    if (null_inv(inv, "mmb_set_fast_compile")) return -1;
    inv->cfg.fast_compile = enabled ? 1 : 0;
    return 0;
}
int mmb_set_warp_shape(mmb_invocation *inv, int warp_width) {
    if (null_inv(inv, "mmb_set_warp_shape")) return -1;
    if (warp_width != 32 && warp_width != 16 && warp_width != 8) { set_error("warp width must be 32, 16 or 8"); return -1; }
    inv->cfg.warp_w = warp_width;
    return 0;
}
int mmb_set_rows_per_thread(mmb_invocation *inv, int rows) {
    if (null_inv(inv, "mmb_set_rows_per_thread")) return -1;
    if (rows != 0 && rows != 1 && rows != 2 && rows != 4 && rows != 8) { set_error("rows per thread must be 0 (automatic), 1, 2, 4 or 8"); return -1; }
    inv->cfg.rows = rows;
    return 0;
}
int mmb_set_edge_behaviour(mmb_invocation *inv, int mode_x, int mode_y, uint32_t color_x, uint32_t color_y) {
    if (null_inv(inv, "mmb_set_edge_behaviour")) return -1;
    if (mode_x < 0 || mode_x > 3 || mode_y < 0 || mode_y > 3) { set_error("bad edge behaviour"); return -1; }
    inv->cfg.edge_x = mode_x;
    inv->cfg.edge_y = mode_y;
    inv->edge_color_x = color_x;
    inv->edge_color_y = color_y;
    return 0;
}
int mmb_set_output_bpp(mmb_invocation *inv, int bpp) {
    if (null_inv(inv, "mmb_set_output_bpp")) return -1;
    if (bpp < 1 || bpp > 4) { set_error("output bpp must be 1..4"); return -1; }
    inv->bpp = bpp;
    return 0;
}

static Userval *userval_slot(mmb_invocation *inv, int index, int type) {
    if (!inv || index < 0 || index >= (int)inv->uservals.size()) { set_error("userval index out of range"); return nullptr; }
    if (inv->uservals[index].type != type) {
        set_error(std::string("userval ") + inv->m->main->uservals[index].name + " is of type " + userval_type_name(inv->uservals[index].type));
        return nullptr;
    }
    inv->frame_ready = false;
    return &inv->uservals[index];
}
int mmb_set_userval_int(mmb_invocation *inv, int index, int value) {
    Userval *u = userval_slot(inv, index, UV_INT);
    if (!u) return -1;
    u->i = value;
    return 0;
}
int mmb_set_userval_float(mmb_invocation *inv, int index, float value) {
    Userval *u = userval_slot(inv, index, UV_FLOAT);
    if (!u) return -1;
    u->f = value;
    return 0;
}
int mmb_set_userval_bool(mmb_invocation *inv, int index, int value) {
    Userval *u = userval_slot(inv, index, UV_BOOL);
    if (!u) return -1;
    u->i = value ? 1 : 0;
    return 0;
}
int mmb_set_userval_color(mmb_invocation *inv, int index, float r, float g, float b, float a) {
    Userval *u = userval_slot(inv, index, UV_COLOR);
    if (!u) return -1;
    auto q = [](float x) { x = x < 0 ? 0 : (x > 1 ? 1 : x); return (uint32_t)(int)(x * 255.0) & 0xffu; };  // MAKE_RGBA_COLOR_FLOAT, color.h:40
    u->color = (q(r) << 24) | (q(g) << 16) | (q(b) << 8) | q(a);
    return 0;
}
int mmb_set_userval_color_packed(mmb_invocation *inv, int index, uint32_t rgba_packed) {
    Userval *u = userval_slot(inv, index, UV_COLOR);
    if (!u) return -1;
    u->color = rgba_packed;
    return 0;
}
static int set_table(mmb_invocation *inv, int index, int type, const void *host, size_t bytes) {
    Userval *u = userval_slot(inv, index, type);
    if (!u) return -1;
    return guarded([&] {
        set_device(inv);
        if (!u->table) ck(cudaMalloc(&u->table, bytes), "cudaMalloc");
        ck(cudaMemcpy(u->table, host, bytes, cudaMemcpyHostToDevice), "cudaMemcpy");
    });
}
int mmb_set_userval_curve(mmb_invocation *inv, int index, const float *values) {
    return set_table(inv, index, UV_CURVE, values, sizeof(float) * MMB_CURVE_POINTS);
}
int mmb_set_userval_gradient(mmb_invocation *inv, int index, const uint32_t *rgba_packed) {
    return set_table(inv, index, UV_GRADIENT, rgba_packed, sizeof(uint32_t) * MMB_CURVE_POINTS);
}

static int set_image(mmb_invocation *inv, int index, const void *data, bool host, int width, int height) {
    Userval *u = userval_slot(inv, index, UV_IMAGE);
    if (!u) return -1;
    if (width <= 0 || height <= 0 || !data) { set_error("bad image"); return -1; }
    return guarded([&] {
        set_device(inv);
        wait_for_user_stream(inv);  // a caller's stream may still sample the image this call replaces
        if (inv->images.size() != inv->persistent_images) {  // drop per-frame images before touching the persistent prefix
            inv->images.resize(inv->persistent_images);
            inv->native_cache.clear();
        }
        HostImage img;
        img.kind = IMG_DRAWABLE;
        img.w = width;
        img.h = height;
        // calc_image_values, userval.c:263-279
        img.sx = (float)((width - 1) / 2.0);
        img.sy = (float)((height - 1) / 2.0);
        img.mx = 1.0f;
        img.my = 1.0f;
        size_t bytes = (size_t)width * height * 4;
        int slot = u->image;
        if (host) {
            void *d = nullptr;
            if (slot >= 0 && inv->images[slot].owned && (size_t)inv->images[slot].w * inv->images[slot].h * 4 == bytes) d = inv->images[slot].data;
            else {
                if (slot >= 0 && inv->images[slot].owned) cudaFree(inv->images[slot].data);
                ck(cudaMalloc(&d, bytes), "cudaMalloc(image)");
            }
            ck(cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, inv->stream), "cudaMemcpyAsync(image)");
            img.data = d;
            img.owned = true;
        } else {
            if (slot >= 0 && inv->images[slot].owned) cudaFree(inv->images[slot].data);
            img.data = const_cast<void *>(data);
            img.owned = false;
        }
        if (slot >= 0) inv->images[slot] = img;
        else {
            inv->images.push_back(img);
            u->image = (int)inv->images.size() - 1;
            inv->persistent_images = inv->images.size();
        }
    });
}
int mmb_set_userval_image_host(mmb_invocation *inv, int index, const uint8_t *rgba, int width, int height) {
    return set_image(inv, index, rgba, true, width, height);
}
int mmb_set_userval_image_device(mmb_invocation *inv, int index, const void *device_rgba, int width, int height) {
    return set_image(inv, index, device_rgba, false, width, height);
}

int mmb_init_frame(mmb_invocation *inv, int frame, float t) {
    if (!inv) return -1;
    return guarded([&] {
        set_device(inv);
        for (size_t i = 0; i < inv->uservals.size(); ++i)
            if (inv->uservals[i].type == UV_IMAGE && inv->uservals[i].image < 0) {
                // the reference samples white from a missing drawable (builtins.c:125-126); we require the binding
                fail("image argument `" + inv->m->main->uservals[i].name + "' has no drawable bound");
            }
        // previous frame's temporaries become reusable: in stream order on the library's stream, and after whatever a
        // caller's stream still runs on them
        wait_for_user_stream(inv);
        inv->release_frame_blocks();
        inv->images.resize(inv->persistent_images);
        inv->native_cache.clear();
        inv->frame = frame;
        inv->t = t;
        std::vector<HVal> uv = main_uservals(inv);
        inv->passthrough_image = -1;
        Replay rp(inv, inv->m->main, uv, frame, t, 0);
        rp.run(rp.code->first);
        const FilterKernel &mk = inv->backend->source.kernels.at(inv->m->main);
        // Pass-through: the filter's pixel is `img(xy)` (cuda_emit.cpp) and img is a blur whose horizontal pass is still
        // pending and whose texel (x, y) is what pixel (x, y) looks up.  That pass then writes the frame's rows itself
        // (render_slice) instead of a floatmap the pixel kernel would only copy and quantise.
        if (mk.passthrough_image) {
            auto it = rp.env.find(mk.passthrough_image);
            if (it != rp.env.end() && it->second.image >= 0 && it->second.image < (int)inv->images.size()) {
                const HostImage &im = inv->images[it->second.image];
                if (im.pending_rows && floatmap_lookup_is_identity(inv, im, inv->render_w, inv->render_h)) inv->passthrough_image = it->second.image;
            }
        }
        // the frame-constant branch conditions of the pixel kernel, as compile-time constants of this frame's kernel
        inv->main_spec.clear();
        for (size_t i = 0; i < mk.spec_conds.size(); ++i) {
            auto it = rp.env.find(mk.spec_conds[i]);
            if (it != rp.env.end()) inv->main_spec += "#define " + mk.spec_prefix + std::to_string(i) + (Replay::truth(it->second) ? " 1\n" : " 0\n");
        }
        pack_frame(inv, mk, rp, inv->main_frame);
        inv->frame_ready = true;
    });
}

int mmb_calc_lines_device(mmb_invocation *inv, int first_row, int last_row, void *device_q, int floatmap, void *stream) {
    if (!inv || !device_q) { set_error("mmb_calc_lines_device: bad arguments"); return -1; }
    return guarded([&] {
        set_device(inv);
        StreamScope scope(inv, (cudaStream_t)stream);
        render_band(inv, first_row, last_row, device_q, floatmap);
        scope.done();
    });
}

// Row-band sharding with load balance: renders the 8-row blocks b of the frame with
// b % count == phase into device_q, stored compactly (block k of this rank at rows [8k, 8k+8)).
// Contiguous bands (the reference's split, mathmap_common.c:997-998) are mmb_calc_lines_device.
int mmb_calc_lines_interleaved_device(mmb_invocation *inv, int phase, int count, void *device_q, void *stream) {
    if (!inv || !device_q || count < 1 || phase < 0 || phase >= count) { set_error("mmb_calc_lines_interleaved_device: bad arguments"); return -1; }
    return guarded([&] {
        set_device(inv);
        if (!inv->frame_ready) fail("mmb_init_frame must be called before rendering");
        if (inv->cfg.supersampling) fail("interleaved bands do not support supersampling");
        int blocks = (inv->render_h + 7) / 8;
        int mine = blocks > phase ? (blocks - phase + count - 1) / count : 0;
        if (mine == 0) return;
        materialise(inv, inv->passthrough_image);
        StreamScope scope(inv, (cudaStream_t)stream);
        LaunchGeom g{inv->render_w, inv->render_h, 0, inv->render_w, 0, mine * 8, 0.f, 0.f, inv->render_w + 1, inv->render_h + 1};
        g.interleave = count;
        g.phase = phase;
        g.row_limit = inv->render_h;
        launch_filter(inv, inv->m->main, inv->main_frame, g, device_q, (long long)inv->render_w * inv->bpp, 0, inv->frame, inv->t);
        scope.done();
    });
}

// The host-buffer flavour of calc_lines: the band is rendered on the device in chunks and copied into q (row pitch
// q_pitch bytes).  `sl` = the reference's slice (mmb_calc_lines_slice), or NULL for the whole-frame band of
// mmb_calc_lines, which also does the supersampling combine of call_invocation.
static void calc_lines_host(mmb_invocation *inv, const SliceGeom *sl, int first_row, int last_row, void *q, size_t q_pitch, int floatmap) {
    set_device(inv);
    const int width = sl ? sl->region_w : inv->render_w;
    const int row_end = sl ? sl->region_y + sl->region_h : inv->render_h;
    int fr = std::max(0, first_row), lr = std::min(last_row, row_end);
    if (lr <= fr || width <= 0) return;
    size_t row_bytes = floatmap ? sizeof(float) * 4 * (size_t)width : (size_t)width * inv->bpp;
    size_t bytes = row_bytes * (size_t)(lr - fr);
    void *d = inv->ensure_staging(inv->staging, inv->staging_bytes, bytes);
    // The band is rendered in chunks; the device->host copy of a chunk runs on a copy stream while the
    // kernels of the following chunks run (q pinned: true overlap; pageable q: still correct).  Chunks
    // alternate between two compute streams, so the last blocks of one chunk (rows of a filter differ in
    // cost) do not leave the device idle before the next chunk starts.
    //
    // Order: kernel then copy is a two-machine flow shop with equal copy times, so the makespan is smallest
    // when the cheapest chunks are rendered first (Johnson's rule): the copy engine starts early and the
    // expensive chunks at the end hide the copy backlog.  Chunk costs are measured (time between consecutive
    // completions) and reused by the next call on the same band, e.g. the next frame of an animation; the
    // first call renders from the outside in (top, bottom, second from top, ...).
    int rows = lr - fr;
    int chunks = 1;
    if (!(inv->cfg.supersampling && !sl) && bytes >= ((size_t)8 << 20)) chunks = (int)std::min<size_t>(32, std::max<size_t>(2, bytes >> 25));
    int chunk_rows = ((rows + chunks - 1) / chunks + 7) & ~7;
    chunks = (rows + chunk_rows - 1) / chunk_rows;
    if (!inv->copy_stream) ck(cudaStreamCreateWithFlags(&inv->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    if (!inv->aux_stream) ck(cudaStreamCreateWithFlags(&inv->aux_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    if (!inv->order_event) ck(cudaEventCreate(&inv->order_event), "cudaEventCreate");
    while ((int)inv->chunk_events.size() < chunks) {
        cudaEvent_t e;
        ck(cudaEventCreate(&e), "cudaEventCreate");
        inv->chunk_events.push_back(e);
    }
    std::vector<int> order(chunks);
    const bool have_costs = inv->cost_fr == fr && inv->cost_lr == lr && inv->cost_chunk_rows == chunk_rows && inv->cost_floatmap == floatmap &&
                            inv->cost_width == width && (int)inv->chunk_cost.size() == chunks;
    if (have_costs) {
        for (int i = 0; i < chunks; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return inv->chunk_cost[a] < inv->chunk_cost[b]; });
    } else {
        for (int i = 0, lo = 0, hi = chunks - 1; i < chunks; ++i) order[i] = (i & 1) ? hi-- : lo++;
    }
    cudaStream_t main_stream = inv->stream;
    ck(cudaEventRecord(inv->order_event, main_stream), "cudaEventRecord");
    // everything queued so far (init_frame's renders and blurs) precedes the chunks on both streams
    if (chunks > 1) ck(cudaStreamWaitEvent(inv->aux_stream, inv->order_event, 0), "cudaStreamWaitEvent");
    try {
        for (int k = 0; k < chunks; ++k) {
            const int ci = order[k];
            const int r0 = fr + ci * chunk_rows, r1 = std::min(lr, r0 + chunk_rows);
            char *dchunk = (char *)d + (size_t)(r0 - fr) * row_bytes;
            inv->stream = (k & 1) ? inv->aux_stream : main_stream;
            if (sl) render_slice(inv, *sl, r0, r1, dchunk, (long long)row_bytes, floatmap);
            else render_band(inv, r0, r1, dchunk, floatmap);
            ck(cudaEventRecord(inv->chunk_events[ci], inv->stream), "cudaEventRecord");
            ck(cudaStreamWaitEvent(inv->copy_stream, inv->chunk_events[ci], 0), "cudaStreamWaitEvent");
            char *qchunk = (char *)q + (size_t)(r0 - fr) * q_pitch;
            if (q_pitch == row_bytes || r1 - r0 == 1)
                ck(cudaMemcpyAsync(qchunk, dchunk, (size_t)(r1 - r0) * row_bytes, cudaMemcpyDeviceToHost, inv->copy_stream), "cudaMemcpyAsync(D2H)");
            else
                ck(cudaMemcpy2DAsync(qchunk, q_pitch, dchunk, row_bytes, row_bytes, (size_t)(r1 - r0), cudaMemcpyDeviceToHost, inv->copy_stream),
                   "cudaMemcpy2DAsync(D2H)");
        }
    } catch (...) {
        inv->stream = main_stream;
        throw;
    }
    inv->stream = main_stream;
    ck(cudaStreamSynchronize(inv->aux_stream), "cudaStreamSynchronize");
    ck(cudaStreamSynchronize(inv->copy_stream), "cudaStreamSynchronize");
    ck(cudaStreamSynchronize(inv->stream), "cudaStreamSynchronize");
    if (chunks > 1) {
        inv->chunk_cost.assign(chunks, 0.0f);
        cudaEvent_t prev = inv->order_event;
        for (int k = 0; k < chunks; ++k) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, prev, inv->chunk_events[order[k]]) != cudaSuccess) ms = 0.0f;
            inv->chunk_cost[order[k]] = ms > 0.0f ? ms : 0.0f;  // a chunk that finished before its predecessor: no cost of its own
            if (ms > 0.0f) prev = inv->chunk_events[order[k]];
        }
        inv->cost_fr = fr; inv->cost_lr = lr; inv->cost_chunk_rows = chunk_rows; inv->cost_floatmap = floatmap; inv->cost_width = width;
    }
}

int mmb_calc_lines(mmb_invocation *inv, int first_row, int last_row, void *q, int floatmap) {
    if (!inv || !q) { set_error("mmb_calc_lines: bad arguments"); return -1; }
    return guarded([&] {
        size_t row_bytes = floatmap ? sizeof(float) * 4 * (size_t)inv->render_w : (size_t)inv->render_w * inv->bpp;
        calc_lines_host(inv, nullptr, first_row, last_row, q, row_bytes, floatmap);
    });
}

// mathfuncs_t.calc_lines (new_template.c.in:208-312) with the reference's own parameters
static bool slice_geom(const mmb_slice *s, SliceGeom &g) {
    if (!s) return false;
    g = SliceGeom{s->frame_render_width, s->frame_render_height, s->region_x, s->region_y, s->region_width, s->region_height,
                  s->sampling_offset_x, s->sampling_offset_y};
    return true;
}
int mmb_calc_lines_slice(mmb_invocation *inv, const mmb_slice *slice, int first_row, int last_row, void *q, int floatmap) {
    SliceGeom g;
    if (!inv || !q || !slice_geom(slice, g)) { set_error("mmb_calc_lines_slice: bad arguments"); return -1; }
    return guarded([&] {
        // q advances by invocation->row_stride per row, or by frame_render_width float[4] pixels for floatmap output
        // (new_template.c.in:299-302)
        size_t pitch = floatmap ? sizeof(float) * 4 * (size_t)g.frame_w : (size_t)slice->row_stride;
        size_t row_bytes = floatmap ? sizeof(float) * 4 * (size_t)g.region_w : (size_t)g.region_w * inv->bpp;
        // one row may be wider than the stride: the reference's supersampling loop renders its region_width + 1 lines one
        // at a time into line buffers of their own (mathmap_common.c:897-905) with the frame's row_stride
        const int rows = std::min(last_row, g.region_y + g.region_h) - std::max(0, first_row);
        if (pitch < row_bytes && rows > 1) fail("mmb_calc_lines_slice: row stride smaller than a row of the region");
        calc_lines_host(inv, &g, first_row, last_row, q, pitch, floatmap);
    });
}
int mmb_calc_lines_slice_device(mmb_invocation *inv, const mmb_slice *slice, int first_row, int last_row, void *device_q, int floatmap, void *stream) {
    SliceGeom g;
    if (!inv || !device_q || !slice_geom(slice, g)) { set_error("mmb_calc_lines_slice_device: bad arguments"); return -1; }
    return guarded([&] {
        set_device(inv);
        long long pitch = floatmap ? (long long)sizeof(float) * 4 * g.frame_w : (long long)slice->row_stride;
        // RGBA8 pixels are stored as one 32-bit word, float pixels as one 128-bit word
        if (floatmap ? ((uintptr_t)device_q & 15) != 0 : (inv->bpp == 4 && (((uintptr_t)device_q | (uintptr_t)pitch) & 3) != 0))
            fail("mmb_calc_lines_slice_device: buffer and row stride must be aligned to the pixel size");
        StreamScope scope(inv, (cudaStream_t)stream);
        render_slice(inv, g, first_row, last_row, device_q, pitch, floatmap);
        scope.done();
    });
}

int mmb_render_frames_device(mmb_invocation *inv, int n, const int *frames, const float *ts, void *device_q, void *stream) {
    if (!inv || !device_q || n < 0 || (n > 0 && !ts)) { set_error("mmb_render_frames_device: bad arguments"); return -1; }
    size_t frame_bytes = (size_t)inv->render_w * inv->render_h * inv->bpp;
    if (n < 4) {
        for (int i = 0; i < n; ++i) {
            if (mmb_init_frame(inv, frames ? frames[i] : i, ts[i]) != 0) return -1;
            if (mmb_calc_lines_device(inv, 0, inv->render_h, (char *)device_q + frame_bytes * i, 0, stream) != 0) return -1;
        }
        return 0;
    }
    // Consecutive frames go to two alternating streams ("lanes"), each with its own pool of per-frame temporaries, so that
    // the ramp-up of one frame's kernels overlaps the tail of the previous frame's (a 3840x2160 frame is a 40 us kernel:
    // back to back on one stream the SMs idle at both ends of each).  The lanes start after what the caller's (or the
    // library's) stream has queued so far, and that stream waits for both lanes at the end.
    int rc = guarded([&] {
        set_device(inv);
        for (int l = 0; l < 2; ++l)
            if (!inv->lane_stream[l]) ck(cudaStreamCreateWithFlags(&inv->lane_stream[l], cudaStreamNonBlocking), "cudaStreamCreate");
        for (auto &e : inv->lane_event)
            if (!e) ck(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
        cudaStream_t base = stream ? (cudaStream_t)stream : inv->stream;
        wait_for_user_stream(inv);
        ck(cudaEventRecord(inv->lane_event[2], base), "cudaEventRecord");
        for (int l = 0; l < 2; ++l) ck(cudaStreamWaitEvent(inv->lane_stream[l], inv->lane_event[2], 0), "cudaStreamWaitEvent");
    });
    if (rc != 0) return rc;
    cudaStream_t saved = inv->stream;
    cudaStream_t base = stream ? (cudaStream_t)stream : saved;
    for (int i = 0; i < n && rc == 0; ++i) {
        inv->lane = i & 1;
        inv->stream = inv->lane_stream[inv->lane];
        rc = mmb_init_frame(inv, frames ? frames[i] : i, ts[i]);
        if (rc == 0) rc = mmb_calc_lines_device(inv, 0, inv->render_h, (char *)device_q + frame_bytes * i, 0, nullptr);
    }
    inv->lane = 0;
    inv->stream = saved;
    for (int l = 0; l < 2; ++l) {
        if (cudaEventRecord(inv->lane_event[l], inv->lane_stream[l]) != cudaSuccess || cudaStreamWaitEvent(base, inv->lane_event[l], 0) != cudaSuccess) {
            if (rc == 0) set_error("mmb_render_frames_device: stream ordering failed");
            rc = -1;
        }
    }
    // the frame state left behind belongs to lane 1's stream; the next init_frame on the library's stream must come after it
    if (base != saved && cudaStreamWaitEvent(saved, inv->lane_event[1], 0) != cudaSuccess) rc = -1;
    if (cudaStreamWaitEvent(saved, inv->lane_event[0], 0) != cudaSuccess) rc = -1;
    return rc;
}

int mmb_synchronize(mmb_invocation *inv) {
    if (null_inv(inv, "mmb_synchronize")) return -1;
    return guarded([&] {
        set_device(inv);
        if (inv->user_event_pending) ck(cudaEventSynchronize(inv->user_event), "cudaEventSynchronize");
        ck(cudaStreamSynchronize(inv->stream), "cudaStreamSynchronize");
        ck(cudaGetLastError(), "kernel execution");
    });
}

long mmb_launch_count(const mmb_invocation *inv) { return inv ? inv->launches : -1; }
const char *mmb_kernel_name(const mmb_invocation *inv) { return inv ? inv->kernel_name.c_str() : nullptr; }

// Work memory of mmb_gaussian_blur_device, kept per device between calls (grow-only): no cudaMalloc / cudaFree per call.
// Calls on one device are serialised by the mutex, and a call waits for its own work before it returns, so the buffers
// are free again when the next call takes them.
namespace {
struct BlurScratch {
    void *buf[3] = {nullptr, nullptr, nullptr};
    size_t cap[3] = {0, 0, 0};
    void *get(int i, size_t bytes) {
        if (cap[i] < bytes) {
            if (buf[i]) cudaFree(buf[i]);
            buf[i] = nullptr;
            cap[i] = 0;
            ck(cudaMalloc(&buf[i], bytes), "cudaMalloc(blur work memory)");
            cap[i] = bytes;
        }
        return buf[i];
    }
};
std::mutex g_blur_mu;
std::map<int, BlurScratch> g_blur_scratch;
}  // namespace

int mmb_gaussian_blur_device(int device, const float *device_in, float *device_out, int width, int height, float sigma_h_px, float sigma_v_px,
                             void *stream) {
    return guarded([&] {
        ck(cudaSetDevice(device), "cudaSetDevice");
        if ((((uintptr_t)device_in | (uintptr_t)device_out) & 15) != 0) fail("mmb_gaussian_blur_device: float4 pixels must be 16-byte aligned");
        if (width <= 0 || height <= 0) return;
        cudaStream_t s = (cudaStream_t)stream;
        size_t bytes = sizeof(float) * 4 * (size_t)width * height;
        std::lock_guard<std::mutex> lock(g_blur_mu);
        BlurScratch &w = g_blur_scratch[device];
        if (sigma_h_px < 0.5f || sigma_v_px < 0.5f)
            launch_gauss_rle(device_in, (float *)w.get(0, bytes), device_out, width, height, sigma_h_px, sigma_v_px,
                             w.get(1, gauss_rle_curve_bytes(sigma_h_px, sigma_v_px)), s);
        else
            launch_gauss_iir(device_in, false, device_out, (double *)w.get(2, gauss_iir_scratch_bytes(width, height)), width, height, sigma_h_px, sigma_v_px, s);
        ck(cudaGetLastError(), "gaussian blur launch");
        ck(cudaStreamSynchronize(s), "gaussian blur");
    });
}

// test hook: the IIR coefficients the host computes (30 doubles: n_p n_m d_p d_m bd_p bd_m)
void mmb_gauss_iir_constants(float std_dev, double *out30) { gauss_iir_constants_host(std_dev, out30); }

}  // extern "C"
