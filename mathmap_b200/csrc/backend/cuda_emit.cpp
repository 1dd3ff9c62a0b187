#include "cuda_emit.h"
#include "../runtime/mm_types.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <set>
#include <sstream>

namespace mmbackend {

using namespace mm;

namespace {

[[noreturn]] void unsupported(const std::string &what) {
    CompileError e;
    e.message = "CUDA backend: " + what;
    throw e;
}

std::string sanitize(const std::string &s) {
    std::string r;
    for (char c : s) r += (isalnum((unsigned char)c) || c == '_') ? c : '_';
    return r;
}

std::string float_literal(float f) {
    if (std::isnan(f)) return "__int_as_float(0x7fc00000)";
    if (std::isinf(f)) return f > 0 ? "__int_as_float(0x7f800000)" : "__int_as_float(0xff800000)";
    char buf[64];
    snprintf(buf, sizeof buf, "%.9g", (double)f);
    std::string s = buf;
    if (!strpbrk(buf, ".en")) s += ".0";
    return s + "f";
}

// ---- exact strength reduction: LESS/LEQ(sqrt(x), c) with x provably >= 0 or NaN ----
// sqrt is correctly rounded, hence monotonic: sqrt_rn(x) < c  <=>  x < t where t is the smallest
// float with sqrt_rn(t) >= c (found by stepping floats around c*c on the host); similarly for <=.
// The result is bit-identical and removes the IEEE square root from escape-time loops (abs(z) < 2).
bool nonneg_or_nan(const Primary &p, int depth = 0) {
    if (p.is_const) return (p.c.type == T_INT && p.c.i >= 0) || (p.c.type == T_FLOAT && p.c.f >= 0.0f);
    const Value *v = p.value;
    if (v->index < 0 || !v->def || depth > 12) return false;
    const Stmt *d = v->def;
    if (d->kind != ST_ASSIGN) return false;
    const Rhs *r = d->rhs;
    if (r->kind == RHS_PRIMARY) return nonneg_or_nan(r->prim, depth + 1);
    if (r->kind != RHS_OP) return false;
    if (v->cv->type != T_FLOAT) return false;
    switch (r->op->id) {
    case OP_MUL: {  // x * x
        const Primary &a = r->args[0], &b = r->args[1];
        if (!a.is_const && !b.is_const && a.value == b.value) return true;
        return nonneg_or_nan(a, depth + 1) && nonneg_or_nan(b, depth + 1);
    }
    case OP_ADD: return nonneg_or_nan(r->args[0], depth + 1) && nonneg_or_nan(r->args[1], depth + 1);
    case OP_ABS: case OP_SQRT: case OP_HYPOT: return true;
    default: return false;
    }
}
// smallest float t with sqrtf(t) >= c (strict) / largest float t with sqrtf(t) <= c (non-strict); c > 0 finite
bool sqrt_threshold(float c, bool strict, float *out) {
    if (!(c > 0.0f) || !std::isfinite(c) || c > 1e18f) return false;
    float t = c * c;
    if (!std::isfinite(t) || t <= 0.0f) return false;
    if (strict) {
        while (sqrtf(t) >= c) t = nextafterf(t, 0.0f);     // now sqrt(t) < c
        while (sqrtf(t) < c) t = nextafterf(t, INFINITY);  // smallest with sqrt(t) >= c
    } else {
        while (sqrtf(t) <= c) t = nextafterf(t, INFINITY);  // now sqrt(t) > c
        while (sqrtf(t) > c) t = nextafterf(t, 0.0f);       // largest with sqrt(t) <= c
    }
    *out = t;
    return true;
}

struct Emitter {
    const mmb_module &mod;
    const FilterCode &code;
    // what this emitter instance prints:
    //   PIXEL_ALL  every device statement (levels 1 and 3) in the pixel kernel
    //   PIXEL_ONLY level 3 only; row-constant (level 1) values are read from the row pre-kernel's arrays
    //   ROW        level 1 only: the row pre-kernel (the reference's per-row "x-const" code, new_template.c.in:251-259)
    //   CALL       a __device__ function evaluating everything per call (non-inlined filter calls)
    enum Mode { PIXEL_ALL, PIXEL_ONLY, ROW, CALL };
    Mode mode;
    bool call_flavour;
    std::ostringstream out;
    std::set<const Value *> own_uniform_set, *uniform_set_p = &own_uniform_set;
    std::vector<const Value *> own_uniform_order, *uniform_order_p = &own_uniform_order;
    std::set<const Value *> row_export_set;
    std::vector<const Value *> row_exports;  // level-1 values the pixel kernel reads
    std::set<const Filter *> &called;  // filters reached through RHS_FILTER
    std::set<const Stmt *> fused;       // statements already emitted as part of a fused pair
    std::string slot_prefix;            // kernels: image uniforms are referenced as the enum constant <slot_prefix><value>
    // Quad mode (mm_runtime.cuh: quad kernels): a thread renders four horizontally adjacent pixels.  Per-pixel values
    // (level 3) are arrays of four, indexed by `qp` while a statement is printed for one of the pixels; qp < 0 at the
    // top level, where a statement is printed once (row-level values, quad samplers) or four times.
    bool quad = false;
    int qp = -1;
    // Column statements: top-level assignments whose value depends on the column only (the reference's "y-const" values,
    // computed once per slice by init_slice, new_template.c.in:339-373).  The pixel kernel prints them once per thread,
    // before its loop over the block's tiles (col_phase 1), and leaves them out of the loop body (col_phase 2).
    std::set<const Stmt *> col_stmts;
    int col_phase = 0;
    void find_column_statements(const Stmt *first) {
        std::set<const Value *> col_values;
        for (const Stmt *s = first; s; s = s->next) {
            if (s->kind != ST_ASSIGN || s->lhs->level < 3 || !(s->lhs->const_bits & CONST_Y) || s == direct_sample || s == direct_output) continue;
            const Rhs *r = s->rhs;
            bool ok = r->kind == RHS_PRIMARY || r->kind == RHS_TUPLE || (r->kind == RHS_INTERNAL && r->internal == "x") ||
                      (r->kind == RHS_OP && r->op->pure && r->op->id != OP_ORIG_VAL && r->op->id != OP_RAND && r->op->id != OP_OUTPUT_TUPLE);
            auto arg_ok = [&](const Primary &p) { return p.is_const || p.value->index < 0 || p.value->level == 0 || col_values.count(p.value) != 0; };
            if (ok && r->kind == RHS_PRIMARY) ok = arg_ok(r->prim);
            if (ok && (r->kind == RHS_OP || r->kind == RHS_TUPLE))
                for (auto &a : r->args) ok = ok && arg_ok(a);
            if (!ok) continue;
            col_stmts.insert(s);
            col_values.insert(s->lhs);
        }
    }
    std::vector<const Value *> *spec_conds = nullptr;  // the pixel kernel only: frame-constant branch conditions (FilterKernel::spec_conds)
    std::string cond_prefix;                           // "MM_COND_<filter>_"

    Emitter(const mmb_module &m, const FilterCode &c, Mode md, std::set<const Filter *> &cl) : mod(m), code(c), mode(md), call_flavour(md == CALL), called(cl) {}

    bool on_device(int level) const {
        switch (mode) {
        case CALL: return true;
        case PIXEL_ALL: return level >= 1;
        case PIXEL_ONLY: return level >= 3;
        case ROW: return level == 1;
        }
        return false;
    }

    static std::string vname(const Value *v) { return "v" + std::to_string(v->cv->id) + "_" + std::to_string(v->index); }
    bool is_vec(const Value *v) const { return quad && v->index >= 0 && v->level >= 3; }
    // the name of a value of this kernel as an lvalue / rvalue for the pixel being printed
    std::string ref(const Value *v) const { return is_vec(v) ? vname(v) + "[" + std::to_string(qp < 0 ? 0 : qp) + "]" : vname(v); }
    std::string pix(const char *name) const { return quad ? std::string(name) + "[" + std::to_string(qp < 0 ? 0 : qp) + "]" : std::string(name); }
    static std::string ctype(const CompVar *cv) {
        switch (cv->type) {
        case T_INT: case T_NIL: return "int";
        case T_FLOAT: return "float";
        case T_COMPLEX: return "float2";
        case T_COLOR: return "mm_color";
        case T_CURVE: return "const float *";
        case T_GRADIENT: return "const mm_color *";
        case T_IMAGE: return "int";
        case T_TUPLE:
        case T_TREE_VECTOR:  // the reference's persistent tree (tree_vectors.c) is a value type here: a register / local array
            return "mm_tup<" + std::to_string(std::max(1, cv->tuple_len)) + ">";
        default: unsupported("internal error: value of unknown type");
        }
    }

    std::string prim(const Primary &p) {
        if (p.is_const) {
            switch (p.c.type) {
            case T_INT: return p.c.i < 0 ? "(" + std::to_string(p.c.i) + ")" : std::to_string(p.c.i);
            case T_FLOAT: { std::string s = float_literal(p.c.f); return p.c.f < 0 ? "(" + s + ")" : s; }
            case T_COMPLEX: return "make_float2(" + float_literal(p.c.c.real()) + ", " + float_literal(p.c.c.imag()) + ")";
            case T_COLOR: return std::to_string(p.c.color) + "u";
            default: unsupported("constant of non-scalar type");
            }
        }
        const Value *v = p.value;
        if (v->index < 0) return (v->cv->type == T_TUPLE || v->cv->type == T_TREE_VECTOR) ? ctype(v->cv) + "{}" : (v->cv->type == T_COMPLEX ? "make_float2(0.f, 0.f)" : "0");
        if (!on_device(v->level)) {
            if (v->level == 0) {
                if (uniform_set_p->insert(v).second) uniform_order_p->push_back(v);
                if (v->cv->type == T_IMAGE && !slot_prefix.empty()) return slot_prefix + vname(v);
                return "U." + vname(v);
            }
            if (mode == PIXEL_ONLY && v->level == 1) {  // loaded from the row arrays at kernel entry
                if (row_export_set.insert(v).second) row_exports.push_back(v);
                return vname(v);
            }
            unsupported("internal error: value of level " + std::to_string(v->level) + " is not visible here");
        }
        return ref(v);
    }
    Type ptype(const Primary &p) const { return primary_type(p); }
    std::string as_float(const Primary &p) {
        std::string s = prim(p);
        return ptype(p) == T_FLOAT ? s : "(float)" + s;
    }
    std::string as_complex(const Primary &p) {
        if (ptype(p) == T_COMPLEX) return prim(p);
        return "make_float2(" + as_float(p) + ", 0.f)";
    }

    std::string internal(const std::string &n) {
        if (n == "x") return quad ? pix("mm_x") : n;
        if (n == "y" || n == "t" || n == "frame") return n;
        if (n == "R") return "P.R";
        if (n == "__canvasPixelW") return "P.img_w";
        if (n == "__canvasPixelH") return "P.img_h";
        if (n == "__renderPixelW") return "P.render_w";
        if (n == "__renderPixelH") return "P.render_h";
        unsupported("internal " + n + " is not available on the device");
    }

    std::string userval_access(int op, const Primary &idx) {
        if (!call_flavour) unsupported("userval access outside the frame-constant slice");
        (void)op;
        return "a" + std::to_string(idx.c.i);
    }

    std::string op_expr(const Rhs *r, Type result_type) {
        const OpInfo *op = r->op;
        auto A = [&](int i) { return prim(r->args[i]); };
        auto F = [&](int i) { return as_float(r->args[i]); };
        // an int parameter fed a float value: C's implicit conversion (x86 truncation semantics)
        auto I = [&](int i) { return ptype(r->args[i]) == T_FLOAT ? "mm_f2i(" + prim(r->args[i]) + ")" : prim(r->args[i]); };
        auto Z = [&](int i) { return as_complex(r->args[i]); };
        Type mx = T_INT;
        for (auto &a : r->args) mx = std::max(mx, ptype(a));
        auto arith = [&](const char *sym, const char *cfn) {
            if (mx == T_COMPLEX) return std::string(cfn) + "(" + Z(0) + ", " + Z(1) + ")";
            return "(" + A(0) + " " + sym + " " + A(1) + ")";
        };
        auto fn1 = [&](const char *name) { return std::string(name) + "(" + F(0) + ")"; };
        auto fn2 = [&](const char *name) { return std::string(name) + "(" + F(0) + ", " + F(1) + ")"; };
        auto cfn1 = [&](const char *name) { return std::string(name) + "(" + Z(0) + ")"; };
        switch (op->id) {
        case OP_NOP: return "0";
        case OP_INT2FLOAT: return "(float)" + A(0);
        case OP_FLOAT2INT: return "mm_f2i(" + F(0) + ")";
        case OP_INT2COMPLEX: case OP_FLOAT2COMPLEX: return Z(0);
        case OP_ADD: return arith("+", "mm_cadd");
        case OP_SUB: return arith("-", "mm_csub");
        case OP_MUL: return arith("*", "mm_cmul");
        case OP_NEG: return mx == T_COMPLEX ? "mm_cneg(" + Z(0) + ")" : "(-" + A(0) + ")";
        case OP_DIV: return fn2("mm_div");
        case OP_MOD: return fn2("mm_mod");
        case OP_ABS: return mx == T_INT ? "mm_abs(" + A(0) + ")" : "mm_abs(" + F(0) + ")";
        case OP_MIN: return mx == T_INT ? "mm_min(" + A(0) + ", " + A(1) + ")" : fn2("mm_min");
        case OP_MAX: return mx == T_INT ? "mm_max(" + A(0) + ", " + A(1) + ")" : fn2("mm_max");
        case OP_SQRT: return fn1("mm_sqrt");
        case OP_HYPOT: return fn2("mm_hypot");
        case OP_SIN: return fn1("mm_sin");
        case OP_COS: return fn1("mm_cos");
        case OP_TAN: return fn1("mm_tan");
        case OP_ASIN: return fn1("mm_asin");
        case OP_ACOS: return fn1("mm_acos");
        case OP_ATAN: return fn1("mm_atan");
        case OP_ATAN2: return fn2("mm_atan2");
        case OP_POW: {
            // pow(x, 2): x * x is exact in double for a float x and the host's pow returns exactly representable results
            // exactly (its error bound is below one ulp), so the narrowed value is the correctly rounded float product
            const Primary &ex = r->args[1];
            if (ex.is_const && ((ex.c.type == T_INT && ex.c.i == 2) || (ex.c.type == T_FLOAT && ex.c.f == 2.0f))) return "mm_sqr(" + F(0) + ")";
            return fn2("mm_pow");
        }
        case OP_EXP: return fn1("mm_exp");
        case OP_LOG: return fn1("mm_log");
        case OP_SINH: return fn1("mm_sinh");
        case OP_COSH: return fn1("mm_cosh");
        case OP_TANH: return fn1("mm_tanh");
        case OP_ASINH: return fn1("mm_asinh");
        case OP_ACOSH: return fn1("mm_acosh");
        case OP_ATANH: return fn1("mm_atanh");
        case OP_GAMMA: return fn1("mm_gamma");
        case OP_BETA: return fn2("mm_beta");
        case OP_FLOOR: return fn1("mm_floor");
        case OP_CEIL: return fn1("mm_ceil");
        case OP_EQ: return "(" + A(0) + " == " + A(1) + ")";
        case OP_LESS:
        case OP_LEQ: {
            const char *sym = op->id == OP_LESS ? " < " : " <= ";
            const Primary &l = r->args[0], &c = r->args[1];
            if (!l.is_const && l.value->index >= 0 && l.value->def && l.value->def->kind == ST_ASSIGN && l.value->def->rhs->kind == RHS_OP &&
                l.value->def->rhs->op->id == OP_SQRT && l.value->cv->type == T_FLOAT && c.is_const && (c.c.type == T_INT || c.c.type == T_FLOAT) &&
                on_device(l.value->level)) {
                const Primary &x = l.value->def->rhs->args[0];
                float cf = c.c.type == T_INT ? (float)c.c.i : c.c.f, thr;
                if (ptype(x) == T_FLOAT && nonneg_or_nan(x) && sqrt_threshold(cf, op->id == OP_LESS, &thr))
                    return "(" + prim(x) + sym + float_literal(thr) + ") /* sqrt(x)" + sym + float_literal(cf) + ", exact */";
            }
            return "(" + A(0) + sym + A(1) + ")";
        }
        case OP_NOT: return "(!" + A(0) + ")";
        case OP_PRINT: case OP_NEWLINE: case OP_START_DEBUG_TUPLE: case OP_SET_DEBUG_TUPLE_DATA: return "0";
        case OP_APPLY_CURVE: return "mm_apply_curve(" + A(0) + ", " + F(1) + ")";
        case OP_APPLY_GRADIENT: return "mm_apply_gradient(" + A(0) + ", " + F(1) + ")";
        case OP_ORIG_VAL: return "mm_orig_val(P, " + A(2) + ", " + F(0) + ", " + F(1) + ", " + F(3) + ")";
        case OP_IMAGE_PIXEL_WIDTH: return "P.images[" + A(0) + "].w";
        case OP_IMAGE_PIXEL_HEIGHT: return "P.images[" + A(0) + "].h";
        case OP_MAKE_RGBA_COLOR: return "mm_make_color(" + F(0) + ", " + F(1) + ", " + F(2) + ", " + F(3) + ")";
        case OP_RED: return "mm_red(" + A(0) + ")";
        case OP_GREEN: return "mm_green(" + A(0) + ")";
        case OP_BLUE: return "mm_blue(" + A(0) + ")";
        case OP_ALPHA: return "mm_alpha(" + A(0) + ")";
        case OP_TUPLE_NTH: return A(0) + ".v[" + A(1) + "]";
        case OP_TREE_VECTOR_NTH: return "mm_tree_vector_nth(" + I(0) + ", " + A(1) + ")";  // opmacros.h:189
        case OP_SET_TREE_VECTOR_NTH: return "mm_set_tree_vector_nth(" + I(0) + ", " + A(1) + ", " + F(2) + ")";  // opmacros.h:190
        case OP_COMPLEX: return "mm_complex(" + F(0) + ", " + F(1) + ")";
        case OP_C_REAL: return Z(0) + ".x";
        case OP_C_IMAG: return Z(0) + ".y";
        case OP_C_SQRT: return cfn1("mm_csqrt");
        case OP_C_SIN: return cfn1("mm_csin");
        case OP_C_COS: return cfn1("mm_ccos");
        case OP_C_TAN: return cfn1("mm_ctan");
        case OP_C_ASIN: return cfn1("mm_casin");
        case OP_C_ACOS: return cfn1("mm_cacos");
        case OP_C_ATAN: return cfn1("mm_catan");
        case OP_C_POW: return "mm_cpow(" + Z(0) + ", " + Z(1) + ")";
        case OP_C_EXP: return cfn1("mm_cexp");
        case OP_C_LOG: return cfn1("mm_clog");
        case OP_C_ARG: return cfn1("mm_carg");
        case OP_C_SINH: return cfn1("mm_csinh");
        case OP_C_COSH: return cfn1("mm_ccosh");
        case OP_C_TANH: return cfn1("mm_ctanh");
        case OP_C_ASINH: return cfn1("mm_casinh");
        case OP_C_ACOSH: return cfn1("mm_cacosh");
        case OP_C_ATANH: return cfn1("mm_catanh");
        case OP_C_GAMMA: return cfn1("mm_cgamma");
        case OP_LIBNOISE_PERLIN: return "mm_libnoise_perlin(" + A(0) + ", " + F(1) + ", " + F(2) + ", " + F(3) + ", " + F(4) + ", " + F(5) + ")";
        case OP_LIBNOISE_BILLOW: return "mm_libnoise_billow(" + A(0) + ", " + F(1) + ", " + F(2) + ", " + F(3) + ", " + F(4) + ", " + F(5) + ")";
        case OP_LIBNOISE_RIDGED_MULTI: return "mm_libnoise_ridged_multi(" + A(0) + ", " + F(1) + ", " + F(2) + ", " + F(3) + ", " + F(4) + ")";
        case OP_LIBNOISE_VORONOI: return "mm_libnoise_voronoi(" + F(0) + ", " + F(1) + ", " + F(2) + ", " + F(3) + ")";
        case OP_USERVAL_INT: case OP_USERVAL_FLOAT: case OP_USERVAL_BOOL: case OP_USERVAL_COLOR:
        case OP_USERVAL_CURVE: case OP_USERVAL_GRADIENT: case OP_USERVAL_IMAGE:
            return userval_access(op->id, r->args[0]);
        case OP_ELL_INT_K_COMP: return fn1("mm_ell_int_k_comp");
        case OP_ELL_INT_E_COMP: return fn1("mm_ell_int_e_comp");
        case OP_ELL_INT_F: return fn2("mm_ell_int_f");
        case OP_ELL_INT_E: return fn2("mm_ell_int_e");
        case OP_ELL_INT_P: return "mm_ell_int_p(" + F(0) + ", " + F(1) + ", " + F(2) + ")";
        case OP_ELL_INT_D: return "mm_ell_int_d(" + F(0) + ", " + F(1) + ")";  // the third argument is unused by GSL
        case OP_ELL_INT_RC: return fn2("mm_ell_int_rc");
        case OP_ELL_INT_RD: return "mm_ell_int_rd(" + F(0) + ", " + F(1) + ", " + F(2) + ")";
        case OP_ELL_INT_RF: return "mm_ell_int_rf(" + F(0) + ", " + F(1) + ", " + F(2) + ")";
        case OP_ELL_INT_RJ: return "mm_ell_int_rj(" + F(0) + ", " + F(1) + ", " + F(2) + ", " + F(3) + ")";
        case OP_ELL_JAC: return fn2("mm_ell_jac");
        case OP_RAND: return "mm_rand(" + pix("mm_rng") + ", " + F(0) + ", " + F(1) + ")";
        case OP_SOLVE_LINEAR_2: return "mm_solve_linear_2(" + A(0) + ", " + A(1) + ")";
        case OP_SOLVE_LINEAR_3: return "mm_solve_linear_3(" + A(0) + ", " + A(1) + ")";
        case OP_OUTPUT_TUPLE: return "(" + pix("mm_ret") + " = " + A(0) + ", 0)";
        case OP_STRIP_RESIZE: return A(0);  // device image handles never carry a resize wrapper of their own
        default:
            (void)result_type;
            unsupported(std::string("op ") + op->name + " cannot run per pixel on the device (it needs the host or an absent third-party library)");
        }
    }

    std::string rhs_expr(const Rhs *r, const CompVar *dest) {
        switch (r->kind) {
        case RHS_PRIMARY: return prim(r->prim);
        case RHS_INTERNAL: return internal(r->internal);
        case RHS_OP: return op_expr(r, dest ? dest->type : T_INT);
        case RHS_TUPLE:
        case RHS_TREE_VECTOR: {
            std::string s = "mm_tup<" + std::to_string(dest ? std::max(1, dest->tuple_len) : (int)r->args.size()) + ">{{";
            for (size_t i = 0; i < r->args.size(); ++i) s += (i ? ", " : "") + as_float(r->args[i]);
            return s + "}}";
        }
        case RHS_FILTER: {
            const Filter *callee = r->filter;
            called.insert(callee);
            size_t nuv = callee->uservals.size();
            std::string s = "mm_call_" + sanitize(callee->name) + "(P";
            for (size_t i = 0; i < nuv; ++i) {
                int t = callee->uservals[i].type;
                s += ", " + ((t == UV_FLOAT) ? as_float(r->args[i]) : prim(r->args[i]));
            }
            // the nesting level travels with the call: recursion depth is the filter's data (IFS Functional's `depth`), the
            // device stack is not the host's -- see MM_MAX_CALL_DEPTH in mm_runtime.cuh
            s += ", " + as_float(r->args[nuv]) + ", " + as_float(r->args[nuv + 1]) + ", " + as_float(r->args[nuv + 2]) + (mode == CALL ? ", mm_depth + 1)" : ", 0)");
            return s;
        }
        case RHS_CLOSURE: unsupported("an image closure that is not frame-constant");
        default: unsupported("internal error: unknown rhs kind");
        }
    }

    // does this statement list contain anything evaluated on the device?
    bool has_device(const Stmt *s) const {
        for (; s; s = s->next) {
            switch (s->kind) {
            case ST_ASSIGN: case ST_PHI: if (on_device(s->lhs->level)) return true; break;
            case ST_IF: if (has_device(s->cons) || has_device(s->alt) || has_device(s->exit)) return true; break;
            case ST_WHILE: if (on_device(s->level) && s->level >= 1) return true; if (has_device(s->body)) return true; break;
            default: break;
            }
        }
        return false;
    }

    // which: 0 every phi, 1 per-pixel (array) phis only, 2 the others only
    void emit_phis_one(const Stmt *phis, int branch, const std::string &ind, int which) {
        std::vector<const Stmt *> live;
        for (const Stmt *p = phis; p; p = p->next)
            if (p->kind == ST_PHI && on_device(p->lhs->level)) {
                if ((which == 1 && !is_vec(p->lhs)) || (which == 2 && is_vec(p->lhs))) continue;
                const Rhs *src = branch == 0 ? p->rhs : p->rhs2;
                if (src->kind == RHS_PRIMARY && !src->prim.is_const && src->prim.value == p->lhs) continue;
                live.push_back(p);
            }
        // parallel-copy hazard: a source that is the destination of another copy in this set
        bool hazard = false;
        for (const Stmt *p : live) {
            const Rhs *src = branch == 0 ? p->rhs : p->rhs2;
            if (src->kind == RHS_PRIMARY && !src->prim.is_const)
                for (const Stmt *q : live)
                    if (q != p && q->lhs == src->prim.value) hazard = true;
        }
        if (!hazard) {
            for (const Stmt *p : live) out << ind << ref(p->lhs) << " = " << rhs_expr(branch == 0 ? p->rhs : p->rhs2, p->lhs->cv) << ";\n";
            return;
        }
        out << ind << "{\n";
        int k = 0;
        for (const Stmt *p : live) out << ind << "    " << ctype(p->lhs->cv) << " mm_pc" << k++ << " = " << rhs_expr(branch == 0 ? p->rhs : p->rhs2, p->lhs->cv) << ";\n";
        k = 0;
        for (const Stmt *p : live) out << ind << "    " << ref(p->lhs) << " = mm_pc" << k++ << ";\n";
        out << ind << "}\n";
    }
    void emit_phis(const Stmt *phis, int branch, const std::string &ind) {
        if (!quad || qp >= 0) { emit_phis_one(phis, branch, ind, 0); return; }
        emit_phis_one(phis, branch, ind, 2);
        for (qp = 0; qp < 4; ++qp) emit_phis_one(phis, branch, ind, 1);
        qp = -1;
    }

    // "Direct output": the filter's last two statements are  v = ORIG_VAL(...); OUTPUT_TUPLE(v)  with no other use
    // of v, i.e. the pixel IS a sample (every distortion filter).  The sampler's interior fast path may then hand
    // over the RGBA8 word it has already rounded instead of four floats that the store would quantise again
    // (k/255 narrowed to float, times 255, truncated, is k for every byte k).
    const Stmt *direct_sample = nullptr, *direct_output = nullptr;
    void find_direct_output(const Stmt *first) {
        const Stmt *prev = nullptr, *last = nullptr;
        for (const Stmt *s = first; s; s = s->next) { prev = last; last = s; }
        if (!prev || prev->kind != ST_ASSIGN || last->kind != ST_ASSIGN) return;
        if (last->rhs->kind != RHS_OP || last->rhs->op->id != OP_OUTPUT_TUPLE || prev->rhs->kind != RHS_OP || prev->rhs->op->id != OP_ORIG_VAL) return;
        const Primary &a = last->rhs->args[0];
        if (a.is_const || a.value != prev->lhs || prev->lhs->uses.size() != 1) return;
        if (!on_device(prev->lhs->level) || !on_device(last->lhs->level)) return;
        direct_sample = prev;
        direct_output = last;
    }

    // Can this ORIG_VAL go through a quad sampler?  Printed at the top level of a quad kernel, sample row and frame
    // shared by the four pixels (the x argument may differ per pixel or not).
    bool quad_sample(const Stmt *s) const {
        if (!quad || qp >= 0 || s->kind != ST_ASSIGN || s->rhs->kind != RHS_OP || s->rhs->op->id != OP_ORIG_VAL || !is_vec(s->lhs)) return false;
        auto scalar = [&](const Primary &p) { return p.is_const || !is_vec(p.value); };
        const Rhs *r = s->rhs;
        return scalar(r->args[1]) && scalar(r->args[2]) && scalar(r->args[3]);
    }
    // the x arguments of the four pixels as a float[4] initialiser
    std::string quad_xs(const Primary &x) {
        std::string l = "{";
        for (qp = 0; qp < 4; ++qp) l += (qp ? ", " : "") + as_float(x);
        qp = -1;
        return l + "}";
    }

    // one assignment, for the pixel `qp` in quad mode
    void emit_assign(const Stmt *s, const std::string &ind) {
        if (s == direct_sample) {
            const Rhs *r = s->rhs;
            if (quad)
                out << ind << "{ bool mm_h; " << ref(s->lhs) << " = mm_orig_val_out(P, " << prim(r->args[2]) << ", " << as_float(r->args[0]) << ", "
                    << as_float(r->args[1]) << ", " << as_float(r->args[3]) << ", " << pix("mm_word") << ", mm_h); mm_have |= (unsigned)mm_h << " << qp
                    << "; }\n";
            else
                out << ind << vname(s->lhs) << " = mm_orig_val_out(P, " << prim(r->args[2]) << ", " << as_float(r->args[0]) << ", "
                    << as_float(r->args[1]) << ", " << as_float(r->args[3]) << ", mm_word, mm_have_word);\n";
            return;
        }
        // sin(v) and cos(v) of the same value in one block share one range reduction
        if (s->rhs->kind == RHS_OP && (s->rhs->op->id == OP_SIN || s->rhs->op->id == OP_COS) && s->lhs->cv->type == T_FLOAT) {
            const int other = s->rhs->op->id == OP_SIN ? OP_COS : OP_SIN;
            const Stmt *mate = nullptr;
            for (const Stmt *q = s->next; q; q = q->next)
                if (q->kind == ST_ASSIGN && q->rhs->kind == RHS_OP && q->rhs->op->id == other && on_device(q->lhs->level) &&
                    q->lhs->cv->type == T_FLOAT && !q->rhs->args[0].is_const && !s->rhs->args[0].is_const &&
                    q->rhs->args[0].value == s->rhs->args[0].value && is_vec(q->lhs) == is_vec(s->lhs) &&
                    (col_stmts.count(q) != 0) == (col_stmts.count(s) != 0)) { mate = q; break; }
            if (mate) {
                fused.insert(mate);
                const Stmt *sn = s->rhs->op->id == OP_SIN ? s : mate, *cs = s->rhs->op->id == OP_SIN ? mate : s;
                out << ind << "mm_sincos(" << as_float(s->rhs->args[0]) << ", " << ref(sn->lhs) << ", " << ref(cs->lhs) << ");\n";
                return;
            }
        }
        out << ind << ref(s->lhs) << " = " << rhs_expr(s->rhs, s->lhs->cv) << ";\n";
    }

    void emit_stmts(const Stmt *s, const std::string &ind) {
        for (; s; s = s->next) {
            switch (s->kind) {
            case ST_ASSIGN:
                if (!on_device(s->lhs->level) || fused.count(s)) break;
                if ((col_phase == 1) != (col_stmts.count(s) != 0) && col_phase != 0) break;
                if (quad_sample(s)) {
                    const Rhs *r = s->rhs;
                    out << ind << "{ const float mm_qx[4] = " << quad_xs(r->args[0]) << "; ";
                    if (s == direct_sample)
                        out << "mm_orig_val_out_quad(P, " << prim(r->args[2]) << ", mm_qx, " << as_float(r->args[1]) << ", " << as_float(r->args[3]) << ", "
                            << vname(s->lhs) << ", mm_word, mm_have); }\n";
                    else
                        out << "mm_orig_val_quad(P, " << prim(r->args[2]) << ", mm_qx, " << as_float(r->args[1]) << ", " << as_float(r->args[3]) << ", "
                            << vname(s->lhs) << "); }\n";
                    break;
                }
                if (quad && qp < 0 && is_vec(s->lhs)) {
                    for (qp = 0; qp < 4; ++qp) emit_assign(s, ind);
                    qp = -1;
                } else
                    emit_assign(s, ind);
                break;
            case ST_IF:
                if (col_phase == 1) break;
                if (!(has_device(s->cons) || has_device(s->alt) || has_device(s->exit))) break;
                if (mode == ROW && s->level > 1) {  // per-pixel condition: only speculated pure row-level definitions live here
                    emit_stmts(s->cons, ind);
                    emit_stmts(s->alt, ind);
                    break;
                }
                if (quad && qp < 0 && s->level >= 3) {  // a per-pixel condition: the whole construct once per pixel
                    for (qp = 0; qp < 4; ++qp) emit_if(s, ind);
                    qp = -1;
                } else
                    emit_if(s, ind);
                break;
            case ST_WHILE:
                if (col_phase == 1) break;
                if (!on_device(s->level)) {
                    // a loop of another level: at most speculated definitions of this level inside
                    if (mode == ROW && s->level > 1) emit_stmts(s->body, ind);
                    break;
                }
                emit_phis(s->entry, 0, ind);
                out << ind << "while (" << rhs_expr(s->cond, nullptr) << ") {\n";
                emit_stmts(s->body, ind + "    ");
                emit_phis(s->entry, 1, ind + "    ");
                out << ind << "}\n";
                break;
            default: break;
            }
        }
    }
    void emit_if(const Stmt *s, const std::string &ind) {
        std::string cond = rhs_expr(s->cond, nullptr);
        const Rhs *c = s->cond;
        if (spec_conds && c->kind == RHS_PRIMARY && !c->prim.is_const && c->prim.value->index >= 0 && c->prim.value->level == 0 &&
            c->prim.value->cv->type == T_INT) {
            size_t n = 0;
            while (n < spec_conds->size() && (*spec_conds)[n] != c->prim.value) ++n;
            if (n == spec_conds->size()) spec_conds->push_back(c->prim.value);
            cond = cond_prefix + std::to_string(n) + "(" + cond + ")";
        }
        out << ind << "if (" << cond << ") {\n";
        emit_stmts(s->cons, ind + "    ");
        emit_phis(s->exit, 0, ind + "    ");
        out << ind << "} else {\n";
        emit_stmts(s->alt, ind + "    ");
        emit_phis(s->exit, 1, ind + "    ");
        out << ind << "}\n";
    }

    void collect_decls(const Stmt *s, std::vector<const Value *> &vals) const {
        for (; s; s = s->next) {
            switch (s->kind) {
            case ST_ASSIGN: case ST_PHI: if (on_device(s->lhs->level)) vals.push_back(s->lhs); break;
            case ST_IF: collect_decls(s->cons, vals); collect_decls(s->alt, vals); collect_decls(s->exit, vals); break;
            case ST_WHILE: collect_decls(s->entry, vals); collect_decls(s->body, vals); break;
            default: break;
            }
        }
    }
};

size_t field_size(Type t, int tuple_len) {
    switch (t) {
    case T_COMPLEX: return 8;
    case T_CURVE: case T_GRADIENT: return 8;
    case T_TUPLE: case T_TREE_VECTOR: return 4 * (size_t)std::max(1, tuple_len);
    default: return 4;
    }
}

}  // namespace

CudaModuleSource emit_cuda_module(const mmb_module &m) {
    CudaModuleSource src;
    std::ostringstream text;
    std::set<const Filter *> called, emitted_calls;
    std::ostringstream kernels_text;

    // Filters that survive optimisation as closure VALUES (RHS_CLOSURE): a closure that is sampled through ORIG_VAL in
    // code that could not inline it (another closure's argument, a rendered closure's input) is called on the device
    // through mm_closure_dispatch, the analogue of the reference's img->v.closure.func pointer (opmacros.h:208-209).
    std::set<const Filter *> closure_targets;
    {
        std::function<void(const Stmt *)> scan = [&](const Stmt *st) {
            for (; st; st = st->next) {
                if (st->kind == ST_ASSIGN && st->rhs->kind == RHS_CLOSURE && st->rhs->filter->kind == FILTER_MATHMAP) closure_targets.insert(st->rhs->filter);
                else if (st->kind == ST_IF) { scan(st->cons); scan(st->alt); }
                else if (st->kind == ST_WHILE) scan(st->body);
            }
        };
        for (auto &fp : m.mod->filters)
            if (fp->kind == FILTER_MATHMAP) scan(m.code_for(fp.get())->first);
    }
    std::ostringstream dispatch;
    int filter_index = -1;

    // forward declarations of callable filters are emitted once we know which are called
    std::vector<std::string> bodies;
    for (auto &fp : m.mod->filters) {
        const Filter *f = fp.get();
        ++filter_index;
        if (f->kind != FILTER_MATHMAP) continue;
        // Kernels only for what can be launched: the main filter and filters that survive as closure values (rendered to a
        // floatmap or called through mm_closure_dispatch).  The other filters of a module (a composition's node types, whose
        // bodies the front end has inlined into the main filter) would only cost compile time: the reference's
        // benchmark composition "Gaussian Blur -> Spin Zoom -> Droste" compiles in 4.1 s instead of 8.3 s.
        if (f != m.main && closure_targets.count(f) == 0) continue;
        const FilterCode *code = m.code_for(f);
        std::string name = sanitize(f->name);
        // Is a row pre-kernel worth it?  Only when the row-constant slice holds real work (libm, division,
        // sampling, noise); a couple of multiplies are cheaper to redo per pixel than to store and reload.
        int row_cost = 0;
        {
            std::function<void(const Stmt *)> scan = [&](const Stmt *st) {
                for (; st; st = st->next) {
                    if (st->kind == ST_ASSIGN && st->lhs->level == 1 && st->rhs->kind == RHS_OP) {
                        int id = st->rhs->op->id;
                        if ((id >= OP_SQRT && id <= OP_BETA) || (id >= OP_C_SQRT && id <= OP_C_GAMMA) || id == OP_ORIG_VAL ||
                            (id >= OP_LIBNOISE_PERLIN && id <= OP_LIBNOISE_VORONOI) || id == OP_MOD)
                            row_cost += 20;
                        else if (id == OP_DIV) row_cost += 4;
                        else row_cost += 1;
                    } else if (st->kind == ST_IF) { scan(st->cons); scan(st->alt); }
                    else if (st->kind == ST_WHILE) { if (st->level == 1) row_cost += 20; scan(st->body); }
                }
            };
            scan(code->first);
        }
        const bool use_rows = row_cost >= 20;
        // Quad kernel (four pixels of a row per thread, mm_runtime.cuh): for straight-line pixel code with at least one
        // sample whose row and frame do not depend on the column -- in(xy), translations, scalings, stencils with
        // constant offsets ... (the reference's "y-const" values, compiler.c:2867-3234, new_template.c.in:339-373)
        bool quad = false;
        {
            std::function<bool(const Stmt *)> has_pixel_loop = [&](const Stmt *st) {
                for (; st; st = st->next) {
                    if (st->kind == ST_WHILE && st->level >= 1) return true;
                    if (st->kind == ST_IF && (has_pixel_loop(st->cons) || has_pixel_loop(st->alt))) return true;
                }
                return false;
            };
            std::function<bool(const Stmt *)> has_row_sample = [&](const Stmt *st) {
                for (; st; st = st->next) {
                    if (st->kind == ST_ASSIGN && st->rhs->kind == RHS_OP && st->rhs->op->id == OP_ORIG_VAL && st->lhs->level >= 3) {
                        auto below = [](const Primary &q, int l) { return q.is_const || q.value->index < 0 || q.value->level < l; };
                        if (below(st->rhs->args[1], 3) && below(st->rhs->args[2], 1) && below(st->rhs->args[3], 3)) return true;
                    } else if (st->kind == ST_IF && st->level < 3 && (has_row_sample(st->cons) || has_row_sample(st->alt)))
                        return true;
                }
                return false;
            };
            quad = !has_pixel_loop(code->first) && has_row_sample(code->first);
        }
        Emitter e(m, *code, use_rows ? Emitter::PIXEL_ONLY : Emitter::PIXEL_ALL, called);
        e.quad = quad;
        e.slot_prefix = "mm_slot_" + name + "_";
        std::vector<const Value *> spec_conds;
        e.spec_conds = &spec_conds;
        e.cond_prefix = "MM_COND_" + name + "_";
        // body first (discovers uniforms and row exports)
        std::vector<const Value *> decls;
        e.collect_decls(code->first, decls);
        e.find_direct_output(code->first);
        e.find_column_statements(code->first);
        e.col_phase = 1;
        e.emit_stmts(code->first, "    ");
        const std::string col_body = e.out.str();
        e.out.str("");
        e.col_phase = 2;
        e.emit_stmts(code->first, "    ");
        std::string body = e.out.str();
        Emitter er(m, *code, Emitter::ROW, called);
        er.slot_prefix = e.slot_prefix;
        er.uniform_set_p = &e.own_uniform_set;
        er.uniform_order_p = &e.own_uniform_order;
        std::vector<const Value *> row_decls;
        std::string row_body;
        if (use_rows) {
            er.collect_decls(code->first, row_decls);
            er.emit_stmts(code->first, "    ");
            row_body = er.out.str();
        }

        // the closure flavour: the pixel code as a device function of (x, y, t); frame constants come from the same
        // uniforms struct (filled by a host replay with the closure's arguments), image slots from its fields
        const bool closure_fn = closure_targets.count(f) != 0;
        Emitter ec(m, *code, Emitter::PIXEL_ALL, called);
        std::vector<const Value *> closure_decls;
        if (closure_fn) {
            ec.uniform_set_p = &e.own_uniform_set;
            ec.uniform_order_p = &e.own_uniform_order;
            ec.collect_decls(code->first, closure_decls);
            ec.emit_stmts(code->first, "    ");
        }

        FilterKernel k;
        k.filter = f;
        k.filter_index = filter_index;
        k.closure_fn = closure_fn;
        k.kernel_name = "mm_kernel_" + name;
        // layout: 8-byte fields first, then 4-byte ones
        std::vector<const Value *> order = e.own_uniform_order;
        std::stable_sort(order.begin(), order.end(), [](const Value *a, const Value *b) {
            auto big = [](const Value *v) { return v->cv->type == T_CURVE || v->cv->type == T_GRADIENT || v->cv->type == T_COMPLEX; };
            return big(a) && !big(b);
        });
        size_t off = 0;
        int image_slots = 0;
        std::ostringstream st, slots;
        st << "struct mm_uniforms_" << name << " {\n";
        for (const Value *v : order) {
            UniformField uf;
            uf.value = v;
            uf.type = v->cv->type;
            uf.tuple_len = v->cv->tuple_len;
            uf.size = field_size(uf.type, uf.tuple_len);
            uf.offset = off;
            off += uf.size;
            if (uf.type == T_IMAGE) {
                if (image_slots >= MM_MAX_IMAGES) unsupported("more than " + std::to_string(MM_MAX_IMAGES) + " image values in one filter");
                uf.image_slot = image_slots++;
                slots << "enum { " << e.slot_prefix << Emitter::vname(v) << " = " << uf.image_slot << " };\n";
            }
            k.uniforms.push_back(uf);
            st << "    " << Emitter::ctype(v->cv) << " " << Emitter::vname(v) << ";\n";
        }
        if (off == 0) { st << "    int mm_unused;\n"; off = 4; }
        off = (off + 7) & ~(size_t)7;
        st << "};\n" << slots.str();
        k.uniforms_size = off;

        // row arrays: one 4-byte array per scalar component of every exported row-constant value
        auto ncomp = [](const Value *v) {
            switch (v->cv->type) {
            case T_COMPLEX: return 2;
            case T_TUPLE: case T_TREE_VECTOR: return std::max(1, v->cv->tuple_len);
            case T_CURVE: case T_GRADIENT: return -1;
            default: return 1;
            }
        };
        std::ostringstream rv, rv_load, rv_store;
        rv << "struct mm_rowvals_" << name << " {\n";
        if (use_rows) {
            for (const Value *v : e.row_exports) {
                int nc = ncomp(v);
                if (nc < 0) unsupported("row-constant curve/gradient values");
                k.row_slots += nc;
                std::string vn = Emitter::vname(v);
                const char *elt = (v->cv->type == T_INT || v->cv->type == T_IMAGE || v->cv->type == T_NIL) ? "int" : (v->cv->type == T_COLOR ? "unsigned" : "float");
                for (int c2 = 0; c2 < nc; ++c2) rv << "    " << elt << " *" << vn << "_" << c2 << ";\n";
                rv_load << "    " << Emitter::ctype(v->cv) << " " << vn << ";\n";
                if (v->cv->type == T_TUPLE || v->cv->type == T_TREE_VECTOR) {
                    for (int c2 = 0; c2 < nc; ++c2) {
                        rv_load << "    " << vn << ".v[" << c2 << "] = __ldg(RV." << vn << "_" << c2 << " + row);\n";
                        rv_store << "    RV." << vn << "_" << c2 << "[row] = " << vn << ".v[" << c2 << "];\n";
                    }
                } else if (v->cv->type == T_COMPLEX) {
                    rv_load << "    " << vn << ".x = __ldg(RV." << vn << "_0 + row); " << vn << ".y = __ldg(RV." << vn << "_1 + row);\n";
                    rv_store << "    RV." << vn << "_0[row] = " << vn << ".x; RV." << vn << "_1[row] = " << vn << ".y;\n";
                } else {
                    rv_load << "    " << vn << " = __ldg(RV." << vn << "_0 + row);\n";
                    rv_store << "    RV." << vn << "_0[row] = " << vn << ";\n";
                }
            }
        }
        if (k.row_slots == 0) rv << "    float *mm_unused;\n";
        rv << "};\n";
        const bool have_rows = use_rows && k.row_slots > 0;
        if (have_rows) k.row_kernel_name = "mm_rows_" + name;
        // Straight-line pixel code gains from rendering several tiles per block (the per-column work and the constant
        // loads are shared); loops (escape-time iteration, Droste levels) make tiles uneven and measured slower that way.
        // auto_rows is the most tiles a block of this kernel takes; the launch picks fewer for small grids.
        {
            std::function<bool(const Stmt *)> has_loop = [&](const Stmt *st) {
                for (; st; st = st->next) {
                    if (st->kind == ST_WHILE && st->level >= 1) return true;
                    if (st->kind == ST_IF && (has_loop(st->cons) || has_loop(st->alt))) return true;
                }
                return false;
            };
            k.auto_rows = has_loop(code->first) ? 1 : 8;
        }

        std::ostringstream fn;
        fn << st.str() << rv.str();
        k.spec_conds = spec_conds;
        k.spec_prefix = "MM_SPEC_" + name + "_";
        for (size_t i = 0; i < spec_conds.size(); ++i)
            fn << "#ifdef " << k.spec_prefix << i << "\n#define " << e.cond_prefix << i << "(x) (" << k.spec_prefix << i << ")\n#else\n#define " << e.cond_prefix
               << i << "(x) (x)\n#endif\n";
        if (have_rows) {
            fn << "extern \"C\" __global__ void __launch_bounds__(256) " << k.row_kernel_name << "(const __grid_constant__ mm_params P, const __grid_constant__ mm_uniforms_"
               << name << " U, const __grid_constant__ mm_rowvals_" << name << " RV) {\n"
               << "    const int row = blockIdx.x * 256 + threadIdx.x;\n"
               << "    if (row >= P.num_rows) return;\n"
               << "    const int arow = mm_actual_row(P, row);\n"
               << "    if (arow >= P.row_limit) return;\n"
               << "    const float y = __ldg(P.ys + arow);\n"
               << "    const float t = P.t; const int frame = P.frame; (void)t; (void)frame; (void)y;\n";
            for (const Value *v : row_decls) fn << "    " << Emitter::ctype(v->cv) << " " << Emitter::vname(v) << ";\n";
            fn << row_body << rv_store.str() << "}\n";
        }
        k.quad = quad;
        if (e.direct_sample) {
            // img(xy) with x, y the pixel's own coordinates and a frame-constant image and frame (pass-through, invocation.cpp)
            const Rhs *r = e.direct_sample->rhs;
            auto internal_is = [](const Primary &p, const char *iname) {
                return !p.is_const && p.value->index >= 0 && p.value->def && p.value->def->kind == ST_ASSIGN && p.value->def->rhs->kind == RHS_INTERNAL &&
                       p.value->def->rhs->internal == iname;
            };
            const Primary &im = r->args[2], &tt = r->args[3];
            if (internal_is(r->args[0], "x") && internal_is(r->args[1], "y") && !im.is_const && im.value->index >= 0 && im.value->level == 0 &&
                (tt.is_const || tt.value->index < 0 || tt.value->level == 0))
                k.passthrough_image = im.value;
        }
        // quad kernels wait on their texel loads: MM_QUAD_BLOCKS blocks per SM (mm_runtime.cuh) -- unless the module calls filters or
        // closures on the device: a register cap makes their frames spill, and the recursion through mm_closure_dispatch then
        // overruns the device stack (IFS Functional)
        fn << "extern \"C\" __global__ void __launch_bounds__(MM_BLOCK_W * MM_BLOCK_H" << (quad && closure_targets.empty() && called.empty() ? ", MM_QUAD_BLOCKS" : "") << ") " << k.kernel_name
           << "(const __grid_constant__ mm_params P, const __grid_constant__ mm_uniforms_"
           << name << " U, const __grid_constant__ mm_rowvals_" << name << " RV) {\n"
           << "    int col, mm_row0;\n"
           << (k.auto_rows > 1 ? "    const int mm_rows = P.rows;\n" : "    constexpr int mm_rows = 1;  // per-pixel loops: one tile per block\n");
        if (quad)
            fn << "    mm_pixel_coords_quad(col, mm_row0, mm_rows);  // columns col .. col+3\n"
               << "    if (col >= P.region_w) return;\n"
               << "    const int mm_np = P.region_w - col < 4 ? P.region_w - col : 4;  // pixels of the strip inside the region\n"
               << "    float mm_x[4];\n"
               << "#pragma unroll\n"
               << "    for (int mm_p = 0; mm_p < 4; ++mm_p) mm_x[mm_p] = __ldg(P.xs + ((mm_p < mm_np ? col + mm_p : col + mm_np - 1) + P.region_x));\n"
               << "    const float t = P.t; const int frame = P.frame; (void)t; (void)frame; (void)mm_x;\n";
        else
            fn << "    mm_pixel_coords(col, mm_row0, mm_rows);\n"
               << "    if (col >= P.region_w) return;\n"
               << "    const float x = __ldg(P.xs + (col + P.region_x));\n"
               << "    const float t = P.t; const int frame = P.frame; (void)t; (void)frame; (void)x;\n";
        // every value of the pixel code is declared here, once; what depends on the column only is computed here too
        for (const Value *v : decls) fn << "    " << Emitter::ctype(v->cv) << " " << Emitter::vname(v) << (e.is_vec(v) ? "[4]" : "") << ";\n";
        if (!col_body.empty()) fn << "    // per column, once for all tiles of the block (init_slice, new_template.c.in:339-373)\n" << col_body;
        fn << "    // tile k of this block: compact row mm_row0 + 8 k, absolute row mm_arow0 + k * mm_astep (8-row blocks may be interleaved over\n"
           << "    // ranks), output row pointer advancing by 8 rows -- all loop-invariant work is done here, once\n"
           << "    const int mm_arow0 = mm_actual_row(P, mm_row0), mm_astep = MM_BLOCK_H * (P.row_interleave > 1 ? P.row_interleave : 1);\n"
           << "    char *mm_outp = (char *)P.out + (size_t)mm_row0 * (size_t)P.out_stride;\n"
           << "    const size_t mm_ostep = (size_t)MM_BLOCK_H * (size_t)P.out_stride;\n"
           << "#pragma unroll 1\n"
           << "    for (int mm_rep = 0; mm_rep < mm_rows; ++mm_rep, mm_outp += mm_ostep) {\n"
           << "    const int row = mm_row0 + mm_rep * MM_BLOCK_H;\n"
           << "    if (row >= P.num_rows) return;\n"
           << "    const int arow = mm_arow0 + mm_rep * mm_astep;\n"
           << "    if (arow >= P.row_limit) return;\n"
           << "    const float y = __ldg(P.ys + arow); (void)y;\n";
        if (quad)
            fn << "    mm_tup<4> mm_ret[4] = {};\n"
               << "    unsigned mm_rng[4];\n"
               << "#pragma unroll\n"
               << "    for (int mm_p = 0; mm_p < 4; ++mm_p) mm_rng[mm_p] = mm_rng_seed(col + mm_p + P.region_x, arow, P.frame);\n"
               << "    (void)mm_rng;\n";
        else
            fn << "    mm_tup<4> mm_ret = mm_tup<4>{};\n"
               << "    unsigned mm_rng = mm_rng_seed(col + P.region_x, arow, P.frame); (void)mm_rng;\n";
        if (have_rows) fn << rv_load.str();
        if (quad) fn << "    unsigned mm_word[4] = {0, 0, 0, 0}, mm_have = 0; (void)mm_word;\n";
        else if (e.direct_sample) fn << "    unsigned mm_word = 0; bool mm_have_word = false;\n";
        fn << body;
        if (quad) fn << "    mm_store_quad(P, mm_outp, col, mm_np, mm_ret, mm_word, mm_have);\n";
        else
            fn << (e.direct_sample ? "    if (mm_have_word) mm_store_word(mm_outp, col, mm_word); else\n" : "") << "    mm_store_pixel(P, mm_outp, col, mm_ret);\n";
        fn << "    }\n}\n";
        if (closure_fn) {  // new_template.c.in:375-422 filter_$name, with the frame constants precomputed
            fn << "__device__ mm_tup<4> mm_closure_" << name << "(const mm_params &P, const mm_uniforms_" << name << " &U, float x, float y, float t) {\n"
               << "    const int frame = 0;\n    (void)frame; (void)x; (void)y; (void)t;\n    mm_tup<4> mm_ret = mm_tup<4>{};\n"
               << "    unsigned mm_rng = mm_rng_seed(__float_as_int(x), __float_as_int(y), __float_as_int(t)); (void)mm_rng;\n";
            for (const Value *v : closure_decls) fn << "    " << Emitter::ctype(v->cv) << " " << Emitter::vname(v) << ";\n";
            fn << ec.out.str() << "    return mm_ret;\n}\n";
            dispatch << "    case " << filter_index << ": return mm_closure_" << name << "(P, *(const mm_uniforms_" << name << " *)img.data, x, y, t);\n";
        }
        bodies.push_back(fn.str());
        src.kernels[f] = k;
    }

    // device functions for filters that are called rather than inlined (recursion, colour arguments)
    std::ostringstream calls, protos;
    std::set<const Filter *> pending = called;
    while (!pending.empty()) {
        const Filter *f = *pending.begin();
        pending.erase(pending.begin());
        if (!emitted_calls.insert(f).second) continue;
        const FilterCode *code = m.code_for(f);
        std::string name = sanitize(f->name);
        std::set<const Filter *> more;
        Emitter e(m, *code, Emitter::CALL, more);
        std::vector<const Value *> decls;
        e.collect_decls(code->first, decls);
        e.emit_stmts(code->first, "    ");
        std::ostringstream sig;
        sig << "__device__ mm_tup<4> mm_call_" << name << "(const mm_params &P";
        for (auto &u : f->uservals) {
            const char *ct = "int";
            switch (u.type) {
            case UV_FLOAT: ct = "float"; break;
            case UV_COLOR: ct = "mm_color"; break;
            case UV_CURVE: ct = "const float *"; break;
            case UV_GRADIENT: ct = "const mm_color *"; break;
            default: break;
            }
            sig << ", " << ct << " a" << u.index;
        }
        sig << ", float x, float y, float t, int mm_depth)";
        protos << sig.str() << ";\n";
        calls << sig.str() << " {\n    if (mm_depth >= MM_MAX_CALL_DEPTH) { mm_call_overflow = 1; return mm_tup<4>{}; }\n"
              << "    const int frame = 0;\n    (void)frame;\n    mm_tup<4> mm_ret = mm_tup<4>{};\n"
              << "    unsigned mm_rng = mm_rng_seed(__float_as_int(x), __float_as_int(y), __float_as_int(t)); (void)mm_rng;\n";
        for (const Value *v : decls) calls << "    " << Emitter::ctype(v->cv) << " " << Emitter::vname(v) << ";\n";
        calls << e.out.str() << "    return mm_ret;\n}\n";
        for (const Filter *g : more)
            if (!emitted_calls.count(g)) pending.insert(g);
    }

    text << "// generated by mathmap_b200 backend/cuda_emit.cpp\n";
    src.has_calls = !emitted_calls.empty();
    if (src.has_calls) text << "__device__ int mm_call_overflow;  // set when a filter call nests deeper than MM_MAX_CALL_DEPTH; read by the host after the launch\n";
    text << protos.str();
    for (auto &b : bodies) text << b;
    text << calls.str();
    // declared in mm_runtime.cuh; img.data points at the closure filter's uniforms, packed by the host
    text << "__device__ mm_tup<4> mm_closure_dispatch(const mm_params &P, const mm_image &img, float x, float y, float t) {\n"
         << "    switch (img.closure_filter) {\n" << dispatch.str() << "    default: break;\n    }\n"
         << "    (void)P; (void)x; (void)y; (void)t;\n    return mm_tup<4>{};\n}\n";
    src.text = text.str();
    return src;
}

}  // namespace mmbackend
