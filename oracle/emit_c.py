"""ORACLE (test infrastructure, not product code): IR text -> host C.

Restates what the reference's cc backend does with the optimised IR
(reference backends/cc.c:73-448: one C variable per SSA value typed by its
compvar, one C statement per IR statement, phis lowered to copies at the end
of each branch / loop entry and back-edge, frame-constant values kept in an
`xy_vars` struct filled by init_frame, row-constant values computed once per
row) inside the per-filter skeleton of new_template.c.in:208-422
(calc_lines / init_frame / filter_<name>).

Input is the "mmir 1" text produced by the product front end
(mathmap_b200/csrc/ir/ir_text.cpp) or by a reference-side backends/cuda.c stub.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
import re
import struct

C_TYPES = {
    "int": "int", "float": "float", "complex": "float _Complex", "color": "color_t",
    "curve": "const float *", "gradient": "const color_t *", "image": "mmo_image *",
    "tuple": "float *", "nil": "int", "tree_vector": "float *",
}
USERVAL_FIELD = {"int": "int_const", "float": "float_const", "bool": "bool_const", "color": "color",
                 "curve": "curve", "gradient": "gradient", "image": "image"}


# ---------------------------------------------------------------- s-expressions
def parse_sexpr(text):
    tokens = re.findall(r'"(?:\\.|[^"\\])*"|[()]|[^\s()]+', text)
    pos = 0

    def rd():
        nonlocal pos
        tok = tokens[pos]
        pos += 1
        if tok == "(":
            lst = []
            while tokens[pos] != ")":
                lst.append(rd())
            pos += 1
            return lst
        if tok.startswith('"'):
            return ("str", tok[1:-1].replace('\\"', '"').replace("\\\\", "\\"))
        return tok

    return rd()


class Filter:
    def __init__(self, sx):
        assert sx[0] == "filter"
        self.name = sx[1]
        self.flags = sx[2][1:]
        self.uservals = []  # (type, name, rest...)
        for u in sx[3][1:]:
            self.uservals.append((u[0], u[1][1], u[2:]))
        self.vartypes = {}
        self.varlengths = {}  # tuple / tree_vector compvars: element count
        for v in sx[4][1:]:
            self.vartypes[int(v[0])] = v[1]
            if len(v) > 2:
                self.varlengths[int(v[0])] = int(v[2])
        self.code = sx[5][1:]
        self.cname = re.sub(r"[^A-Za-z0-9_]", "_", self.name)


class Module:
    def __init__(self, text):
        sx = parse_sexpr(text)
        assert sx[0] == "mmir" and sx[1] == "1", "not an mmir 1 module"
        self.filters = {}
        self.order = []
        self.main = None
        for item in sx[2:]:
            if item[0] == "filter":
                f = Filter(item)
                self.filters[f.name] = f
                self.order.append(f)
            elif item[0] == "main":
                self.main = item[1]
        if self.main is None:
            self.main = self.order[-1].name
        # native filters known to the runtime: name -> (C function, [userval field per arg])
        self.natives = {
            "gaussian_blur": ("native_filter_gaussian_blur", ["image", "float_const", "float_const"]),
            "convolve": ("native_filter_convolve", ["image", "image", "bool_const", "bool_const"]),
            "half_convolve": ("native_filter_half_convolve", ["image", "image", "bool_const"]),
            "visualize_fft": ("native_filter_visualize_fft", ["image", "bool_const"]),
        }


def float_literal(tok):
    if tok == "nan":
        return "(0.0/0.0)"
    if tok == "inf":
        return "(1.0/0.0)"
    if tok == "-inf":
        return "(-1.0/0.0)"
    # the float32 value printed exactly as a double literal (cc.c prints %.17g of the float)
    f32 = struct.unpack("f", struct.pack("f", float(tok)))[0]
    s = repr(float(f32))
    if "e" not in s and "." not in s and "n" not in s:
        s += ".0"
    return s


class Emitter:
    def __init__(self, module, filt):
        self.m = module
        self.f = filt
        self.levels = {}   # value name -> level
        self.types = {}    # value name -> type name
        self.collect(self.f.code)

    # -- bookkeeping --------------------------------------------------------
    @staticmethod
    def vname(tok):
        cv, idx = tok[1:].split(".")
        return "v%s_%s" % (cv, idx)

    def collect(self, stmts):
        for s in stmts:
            if s[0] == "assign":
                self.note(s[1], int(s[3]))
            elif s[0] == "if":
                self.collect(s[3])
                self.collect(s[4])
                for p in s[5][1:]:
                    self.note(p[1], int(p[3]))
            elif s[0] == "while":
                for p in s[1][1:]:
                    self.note(p[1], int(p[3]))
                self.collect(s[4])

    def note(self, tok, level):
        n = self.vname(tok)
        self.levels[n] = level
        self.types[n] = self.f.vartypes[int(tok[1:].split(".")[0])]

    # -- expressions --------------------------------------------------------
    def prim(self, tok):
        if tok.startswith("%"):
            cv, idx = tok[1:].split(".")
            if idx == "u":
                return "UNINITED_IMAGE" if self.f.vartypes.get(int(cv)) == "image" else "0 /* uninitialized */"
            n = self.vname(tok)
            return ("xy_vars->" + n) if self.levels.get(n) == 0 else n
        kind, val = tok.split(":", 1)
        if kind == "i":
            return "(%s)" % val if val.startswith("-") else val
        if kind == "f":
            lit = float_literal(val)
            return "(%s)" % lit if lit.startswith("-") else lit
        if kind == "c":
            re_, im = val.split(",")
            return "(%s + %s * I)" % (float_literal(re_), float_literal(im))
        if kind == "k":
            return "%su" % val
        raise ValueError(tok)

    def closure_expr(self, fname, args):
        callee = self.m.filters.get(fname)
        if callee is None:
            cfunc, fields = self.m.natives.get(fname, (None, None))
            if cfunc is None:
                raise NotImplementedError("oracle: native filter %s is not supported" % fname)
            parts = ["({ mmo_userval args[%d]; " % max(1, len(args))]
            for i, a in enumerate(args):
                parts.append("args[%d].v.%s = %s; " % (i, fields[i], self.prim(a)))
            parts.append("%s(invocation, args, pools); })" % cfunc)
            return "".join(parts)
        parts = ["({ mmo_image *image = (mmo_image *)mmo_pools_alloc(pools, sizeof(mmo_image)); image->type = MMO_IMAGE_CLOSURE; "
                 "image->func = filter_%s; image->calc_lines = render_closure_%s; image->xy_vars = 0; image->num_args = %d; "
                 % (callee.cname, callee.cname, len(args))]
        for i, a in enumerate(args):
            parts.append("image->args[%d].v.%s = %s; " % (i, USERVAL_FIELD[callee.uservals[i][0]], self.prim(a)))
        parts.append("image->pixel_width = __canvasPixelW; image->pixel_height = __canvasPixelH; ")
        return "".join(parts)

    def rhs(self, r):
        if isinstance(r, str):
            return self.prim(r)
        head = r[0]
        if head == "internal":
            return r[1]
        if head == "op":
            name, args = r[1], [self.prim(a) for a in r[2:]]
            m = re.match(r"USERVAL_(\w+)_ACCESS", name)
            if m:
                return "(arguments[%s].v.%s)" % (args[0], USERVAL_FIELD[m.group(1).lower()])
            if name == "OUTPUT_TUPLE":
                return "((return_tuple = (%s)), 0)" % args[0]
            if name in ("SOLVE_LINEAR_2", "SOLVE_LINEAR_3"):
                return "mmo_solve_linear_%s(%s, pools)" % (name[-1], ",".join(args))
            if name in ("TREE_VECTOR_NTH", "SET_TREE_VECTOR_NTH"):
                # tree_vectors.c:83-146 keeps a persistent tree; as a value it is an array whose length travels with
                # the compvar type: get clamps the index, set copies and replaces one element
                tv = r[3]
                length = self.f.varlengths[int(tv[1:].split(".")[0])]
                if name == "TREE_VECTOR_NTH":
                    return "mmo_tree_vector_get(%s, %d, %s)" % (args[1], length, args[0])
                return "mmo_tree_vector_set(pools, %s, %d, %s, %s)" % (args[1], length, args[0], args[2])
            if name.startswith("SOLVE_"):
                return "mmo_unsupported_op(\"%s\")" % name
            if name == "RAND":
                return "mmo_rand(%s)" % ",".join(args)
            return "%s(%s)" % (name, ",".join(args))
        if head in ("tuple", "tree-vector"):
            n = len(r) - 1
            body = "".join("tuple[%d] = %s; " % (i, self.prim(a)) for i, a in enumerate(r[1:]))
            return "({ float *tuple = ALLOC_TUPLE(%d); %stuple; })" % (n, body)
        if head == "closure":
            e = self.closure_expr(r[1], r[2:])
            return e if r[1] not in self.m.filters else e + "image; })"
        if head == "filter":
            callee = self.m.filters[r[1]]
            nuv = len(callee.uservals)
            e = self.closure_expr(r[1], r[2:2 + nuv])
            x, y, t = (self.prim(a) for a in r[2 + nuv:5 + nuv])
            return e + "filter_%s(invocation, image, %s, %s, %s, pools); })" % (callee.cname, x, y, t)
        raise ValueError(head)

    # -- statements ---------------------------------------------------------
    def contains_level(self, stmts, pred):
        for s in stmts:
            if s[0] == "assign":
                if pred(int(s[3])):
                    return True
            elif s[0] == "if":
                if self.contains_level(s[3], pred) or self.contains_level(s[4], pred):
                    return True
                if any(pred(int(p[3])) for p in s[5][1:]):
                    return True
            elif s[0] == "while":
                if any(pred(int(p[3])) for p in s[1][1:]) or self.contains_level(s[4], pred):
                    return True
        return False

    def emit_phis(self, out, phis, branch, pred, ind):
        for p in phis:
            if not pred(int(p[3])):
                continue
            src = p[4 + branch]
            if isinstance(src, str) and src == p[1]:
                continue
            out.append("%s%s = %s;" % (ind, self.prim(p[1]), self.rhs(src)))

    def emit(self, out, stmts, pred, ind="    "):
        """Emit the statements whose level satisfies pred (a slice of the code)."""
        lo = min(l for l in (0, 1, 3) if pred(l))
        for s in stmts:
            if s[0] == "assign":
                if pred(int(s[3])):
                    out.append("%s%s = %s;" % (ind, self.prim(s[1]), self.rhs(s[4])))
            elif s[0] == "if":
                cond, level, cons, alt, phis = s[1], int(s[2]), s[3], s[4], s[5][1:]
                if not self.contains_level([s], pred):
                    continue
                if level <= lo or pred(level):
                    out.append("%sif (%s)\n%s{" % (ind, self.rhs(cond), ind))
                    self.emit(out, cons, pred, ind + "    ")
                    self.emit_phis(out, phis, 0, pred, ind + "    ")
                    out.append("%s}\n%selse\n%s{" % (ind, ind, ind))
                    self.emit(out, alt, pred, ind + "    ")
                    self.emit_phis(out, phis, 1, pred, ind + "    ")
                    out.append("%s}" % ind)
                else:  # condition not available at this level: pure definitions are speculated
                    self.emit(out, cons, pred, ind)
                    self.emit(out, alt, pred, ind)
            elif s[0] == "while":
                phis, cond, level, body = s[1][1:], s[2], int(s[3]), s[4]
                if pred(level):
                    self.emit_phis(out, phis, 0, pred, ind)
                    out.append("%swhile (%s)\n%s{" % (ind, self.rhs(cond), ind))
                    self.emit(out, body, pred, ind + "    ")
                    self.emit_phis(out, phis, 1, pred, ind + "    ")
                    out.append("%s}" % ind)
                elif level > lo:
                    self.emit(out, body, pred, ind)

    def decls(self, pred):
        lines = []
        for n in sorted(self.levels, key=lambda s: [int(x) for x in s[1:].split("_")]):
            if pred(self.levels[n]):
                lines.append("    %s %s;" % (C_TYPES[self.types[n]], n))
        return lines


PROLOGUE = """/* generated by oracle/emit_c.py -- ORACLE, test infrastructure only */
#include "mmo_runtime.h"
#include <stdio.h>
static int mmo_unsupported_op(const char *n) { fprintf(stderr, "oracle: op %s is not supported\\n", n); abort(); return 0; }
static mmo_image *mmo_unsupported_native(const char *n) { fprintf(stderr, "oracle: native filter %s is not supported\\n", n); abort(); return 0; }
/* RAND: the reference draws from glib's global generator (opmacros.h:127), which no two runs reproduce.  The oracle uses
 * the product's counter-based generator (seeded per pixel from column, row and frame; PCG output function) so that filters
 * calling rand() can still be compared pixel by pixel.  PARITY UNPINNED against the reference by nature. */
static unsigned mmo_rng_seed(int a, int b, int c) {
    unsigned h = (unsigned)a * 0x9E3779B1u ^ ((unsigned)b * 0x85EBCA77u + 0x7F4A7C15u) ^ ((unsigned)c * 0xC2B2AE3Du);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}
static double mmo_rand_next(unsigned *state, double a, double b) {
    unsigned w;
    *state = *state * 747796405u + 2891336453u;
    w = ((*state >> ((*state >> 28u) + 4u)) ^ *state) * 277803737u;
    w = (w >> 22u) ^ w;
    return a + (b - a) * ((double)w * (1.0 / 4294967296.0));
}
static int mmo_float_bits(float f) { int i; memcpy(&i, &f, 4); return i; }
static float mmo_tree_vector_get(const float *tv, int length, int index) {
    if (tv == 0) return 0.0f; /* never assigned */
    if (index < 0) index = 0; else if (index >= length) index = length - 1;
    return tv[index];
}
static float *mmo_tree_vector_set(mmo_pools *pools, const float *tv, int length, int index, float value) {
    float *copy = ALLOC_TUPLE(length);
    if (tv) memcpy(copy, tv, sizeof(float) * length); else memset(copy, 0, sizeof(float) * length);
    if (index < 0) index = 0; else if (index >= length) index = length - 1;
    copy[index] = value;
    return copy;
}
#define mmo_rand(a, b) mmo_rand_next(&mmo_rng, (a), (b))
"""

FILTER_TEMPLATE = """
typedef struct {
@XY_DECLS@
    int dummy;
} xy_vars_@N@;

static void init_frame_@N@(mmo_invocation *invocation, mmo_image *closure, int frame, float t, xy_vars_@N@ *xy_vars, mmo_pools *pools)
{
    int __canvasPixelW = invocation->img_width, __canvasPixelH = invocation->img_height;
    int __renderPixelW = invocation->render_width, __renderPixelH = invocation->render_height;
    float R = invocation->image_R;
    mmo_userval *arguments = closure->args;
    float *return_tuple = 0;
    unsigned mmo_rng = 0;
    (void)__canvasPixelW; (void)__canvasPixelH; (void)__renderPixelW; (void)__renderPixelH; (void)R; (void)arguments; (void)return_tuple; (void)frame; (void)t;
    (void)mmo_rng;
@XY_CODE@
}

void calc_lines_@N@(mmo_invocation *invocation, mmo_image *closure, void *_xy_vars, int frame, float t, int frame_render_width, int frame_render_height,
                    int region_x, int region_y, int region_width, int region_height, float sampling_offset_x, float sampling_offset_y,
                    int first_row, int last_row, void *q, int floatmap)
{
    xy_vars_@N@ *xy_vars = (xy_vars_@N@ *)_xy_vars;
    int __canvasPixelW = invocation->img_width, __canvasPixelH = invocation->img_height;
    int __renderPixelW = invocation->render_width, __renderPixelH = invocation->render_height;
    float R = invocation->image_R;
    mmo_userval *arguments = closure->args;
    int output_bpp = invocation->output_bpp;
    mmo_pools slice_pools, pixel_pools, *pools;
    int row, col;
    unsigned mmo_rng = 0;
@LOCAL_DECLS@
    (void)__canvasPixelW; (void)__canvasPixelH; (void)__renderPixelW; (void)__renderPixelH; (void)R; (void)arguments; (void)frame; (void)t;
    mmo_pools_init(&slice_pools);
    mmo_pools_init(&pixel_pools);
    first_row = MAX(0, first_row);
    last_row = MIN(last_row, region_y + region_height);
    for (row = first_row - region_y; row < last_row - region_y; ++row)
    {
        float y = CALC_VIRTUAL_Y(row + region_y, frame_render_height, sampling_offset_y);
        unsigned char *p = (unsigned char *)q;
        float *fp = (float *)q;
        float *return_tuple = 0;
        pools = &slice_pools;
        (void)y;
@ROW_CODE@
        pools = &pixel_pools;
        for (col = 0; col < region_width; ++col)
        {
            float x = CALC_VIRTUAL_X(col + region_x, frame_render_width, sampling_offset_x);
            (void)x;
            mmo_pools_reset(pools);
            mmo_rng = mmo_rng_seed(col + region_x, row + region_y, frame);
@PIXEL_CODE@
            if (floatmap)
            {
                int i;
                for (i = 0; i < NUM_FLOATMAP_CHANNELS; ++i) fp[i] = return_tuple[i];
            }
            else
                mmo_store_pixel(p, return_tuple, output_bpp);
            p += output_bpp;
            fp += NUM_FLOATMAP_CHANNELS;
        }
        if (floatmap) q = (float *)q + frame_render_width * NUM_FLOATMAP_CHANNELS;
        else q = (unsigned char *)q + invocation->row_stride;
    }
    mmo_pools_free(&pixel_pools);
    mmo_pools_free(&slice_pools);
}

void *new_frame_@N@(mmo_invocation *invocation, mmo_image *closure, int frame, float t, mmo_pools *frame_pools)
{
    xy_vars_@N@ *xy_vars = (xy_vars_@N@ *)mmo_pools_alloc(frame_pools, sizeof(xy_vars_@N@));
    init_frame_@N@(invocation, closure, frame, t, xy_vars, frame_pools);
    return xy_vars;
}

/* render_image() of a closure: new frame (frame 0, t as given), whole region, floatmap output */
static void render_closure_@N@(mmo_invocation *invocation, mmo_image *closure, int frame, float t, int frame_w, int frame_h,
                               int region_x, int region_y, int region_w, int region_h, float off_x, float off_y,
                               int first_row, int last_row, void *q, int floatmap)
{
    mmo_pools frame_pools;
    void *xy;
    mmo_pools_init(&frame_pools);
    xy = new_frame_@N@(invocation, closure, frame, t, &frame_pools);
    calc_lines_@N@(invocation, closure, xy, frame, t, frame_w, frame_h, region_x, region_y, region_w, region_h, off_x, off_y, first_row, last_row, q, floatmap);
    mmo_pools_free(&frame_pools);
}

/* single-pixel evaluation for calls that were not inlined (new_template.c.in:375-422) */
static float *filter_@N@(mmo_invocation *invocation, mmo_image *closure, float x, float y, float t, mmo_pools *pools)
{
    int frame = 0;
    int __canvasPixelW = invocation->img_width, __canvasPixelH = invocation->img_height;
    int __renderPixelW = invocation->render_width, __renderPixelH = invocation->render_height;
    float R = invocation->image_R;
    mmo_userval *arguments = closure->args;
    float *return_tuple = 0;
    xy_vars_@N@ *xy_vars;
    unsigned mmo_rng = mmo_rng_seed(mmo_float_bits(x), mmo_float_bits(y), mmo_float_bits(t));
@LOCAL_DECLS@
    (void)__canvasPixelW; (void)__canvasPixelH; (void)__renderPixelW; (void)__renderPixelH; (void)R; (void)arguments; (void)frame;
    if (closure->xy_vars == 0)
    {
        /* lives as long as the closure's own allocation pool in the reference; here: leaked per closure instance */
        static __thread mmo_pools closure_pools;
        xy_vars = (xy_vars_@N@ *)mmo_pools_alloc(&closure_pools, sizeof(xy_vars_@N@));
        init_frame_@N@(invocation, closure, frame, t, xy_vars, &closure_pools);
        closure->xy_vars = xy_vars;
    }
    else
        xy_vars = (xy_vars_@N@ *)closure->xy_vars;
@ROW_CODE@
@PIXEL_CODE@
    return return_tuple;
}
"""


def emit_module(ir_text):
    """Returns the C source for a whole module (all filters + main entry points)."""
    m = Module(ir_text)
    parts = [PROLOGUE]
    for f in m.order:
        parts.append("static float *filter_%s(mmo_invocation *, mmo_image *, float, float, float, mmo_pools *);" % f.cname)
        parts.append("static void render_closure_%s(mmo_invocation *, mmo_image *, int, float, int, int, int, int, int, int, float, float, int, int, void *, int);" % f.cname)
    for f in m.order:
        e = Emitter(m, f)
        xy_code, row_code, pix_code = [], [], []
        e.emit(xy_code, f.code, lambda l: l == 0, "    ")
        e.emit(row_code, f.code, lambda l: l == 1, "        ")
        e.emit(pix_code, f.code, lambda l: l == 3, "            ")
        src = FILTER_TEMPLATE
        src = src.replace("@XY_DECLS@", "\n".join(e.decls(lambda l: l == 0)))
        src = src.replace("@LOCAL_DECLS@", "\n".join(e.decls(lambda l: l != 0)))
        src = src.replace("@XY_CODE@", "\n".join(xy_code))
        src = src.replace("@ROW_CODE@", "\n".join(row_code))
        src = src.replace("@PIXEL_CODE@", "\n".join(pix_code))
        src = src.replace("@N@", f.cname)
        parts.append(src)
    main = m.filters[m.main]
    parts.append("""
void *mmo_main_new_frame(mmo_invocation *invocation, mmo_image *closure, int frame, float t, mmo_pools *frame_pools)
{ return new_frame_%(n)s(invocation, closure, frame, t, frame_pools); }
void mmo_main_calc_lines(mmo_invocation *invocation, mmo_image *closure, void *xy_vars, int frame, float t, int frame_w, int frame_h,
                         int region_x, int region_y, int region_w, int region_h, float off_x, float off_y, int first_row, int last_row, void *q, int floatmap)
{ calc_lines_%(n)s(invocation, closure, xy_vars, frame, t, frame_w, frame_h, region_x, region_y, region_w, region_h, off_x, off_y, first_row, last_row, q, floatmap); }
""" % {"n": main.cname})
    return "\n".join(parts), m
