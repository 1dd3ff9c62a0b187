// Operator table and IR helpers.  The table restates reference ops.lisp:112-253
// (arity, type propagation rule, result type, purity, foldability, argument
// types); names are the reference's op->name strings so that IR text written by
// a reference-side backends/cuda.c stub resolves without a mapping table.
#include "ir.h"

#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <unordered_map>

namespace mm {

static const char *k_type_names[] = {"nil", "int", "float", "complex", "color", "curve", "gradient", "image", "tuple", "tree_vector"};
const char *type_name(Type t) { return k_type_names[(int)t]; }
Type type_from_name(const std::string &s) {
    for (int i = 0; i < 10; ++i)
        if (s == k_type_names[i]) return (Type)i;
    return T_NIL;
}
const char *userval_type_name(int t) {
    static const char *n[] = {"int", "float", "bool", "color", "curve", "gradient", "image"};
    return n[t];
}

#define F T_FLOAT
#define I T_INT
#define C T_COMPLEX
#define IMG T_IMAGE
#define TUP T_TUPLE
#define COL T_COLOR
#define TV T_TREE_VECTOR
static OpInfo k_ops[NUM_OPS] = {
    // id, name, nargs, prop, type, pure, foldable, argtypes
    {OP_NOP, "NOP", 0, TP_CONST, I, true, true, {}},
    {OP_INT2FLOAT, "INT2FLOAT", 1, TP_CONST, F, true, true, {I}},
    {OP_FLOAT2INT, "FLOAT2INT", 1, TP_CONST, I, true, true, {F}},
    {OP_INT2COMPLEX, "INT2COMPLEX", 1, TP_CONST, C, true, true, {I}},
    {OP_FLOAT2COMPLEX, "FLOAT2COMPLEX", 1, TP_CONST, C, true, true, {F}},
    {OP_ADD, "ADD", 2, TP_MAX, T_NIL, true, true, {}},
    {OP_SUB, "SUB", 2, TP_MAX, T_NIL, true, true, {}},
    {OP_NEG, "NEG", 1, TP_MAX, T_NIL, true, true, {}},
    {OP_MUL, "MUL", 2, TP_MAX, T_NIL, true, true, {}},
    {OP_DIV, "DIV", 2, TP_CONST, F, true, true, {F, F}},
    {OP_MOD, "MOD", 2, TP_CONST, F, true, true, {F, F}},
    {OP_ABS, "fabs", 1, TP_MAX_FLOAT, T_NIL, true, true, {}},
    {OP_MIN, "MIN", 2, TP_MAX_FLOAT, T_NIL, true, true, {}},
    {OP_MAX, "MAX", 2, TP_MAX_FLOAT, T_NIL, true, true, {}},
    {OP_SQRT, "sqrt", 1, TP_CONST, F, true, true, {F}},
    {OP_HYPOT, "hypot", 2, TP_CONST, F, true, true, {F, F}},
    {OP_SIN, "sin", 1, TP_CONST, F, true, true, {F}},
    {OP_COS, "cos", 1, TP_CONST, F, true, true, {F}},
    {OP_TAN, "tan", 1, TP_CONST, F, true, true, {F}},
    {OP_ASIN, "asin", 1, TP_CONST, F, true, true, {F}},
    {OP_ACOS, "acos", 1, TP_CONST, F, true, true, {F}},
    {OP_ATAN, "atan", 1, TP_CONST, F, true, true, {F}},
    {OP_ATAN2, "atan2", 2, TP_CONST, F, true, true, {F, F}},
    {OP_POW, "pow", 2, TP_CONST, F, true, true, {F, F}},
    {OP_EXP, "exp", 1, TP_CONST, F, true, true, {F}},
    {OP_LOG, "log", 1, TP_CONST, F, true, true, {F}},
    {OP_SINH, "sinh", 1, TP_CONST, F, true, true, {F}},
    {OP_COSH, "cosh", 1, TP_CONST, F, true, true, {F}},
    {OP_TANH, "tanh", 1, TP_CONST, F, true, true, {F}},
    {OP_ASINH, "asinh", 1, TP_CONST, F, true, true, {F}},
    {OP_ACOSH, "acosh", 1, TP_CONST, F, true, true, {F}},
    {OP_ATANH, "atanh", 1, TP_CONST, F, true, true, {F}},
    {OP_GAMMA, "GAMMA", 1, TP_CONST, F, true, false, {F}},
    {OP_BETA, "gsl_sf_beta", 2, TP_CONST, F, true, false, {F, F}},
    {OP_FLOOR, "floor", 1, TP_CONST, I, true, true, {F}},
    {OP_CEIL, "ceil", 1, TP_CONST, I, true, true, {F}},
    {OP_EQ, "EQ", 2, TP_CONST, I, true, true, {F, F}},
    {OP_LESS, "LESS", 2, TP_CONST, I, true, true, {F, F}},
    {OP_LEQ, "LEQ", 2, TP_CONST, I, true, true, {F, F}},
    {OP_NOT, "NOT", 1, TP_CONST, I, true, true, {I}},
    {OP_PRINT, "PRINT_FLOAT", 1, TP_CONST, I, false, false, {F}},
    {OP_NEWLINE, "NEWLINE", 0, TP_CONST, I, false, false, {}},
    {OP_START_DEBUG_TUPLE, "START_DEBUG_TUPLE", 1, TP_CONST, I, false, false, {I}},
    {OP_SET_DEBUG_TUPLE_DATA, "SET_DEBUG_TUPLE_DATA", 2, TP_CONST, I, false, false, {I, F}},
    {OP_APPLY_CURVE, "APPLY_CURVE", 2, TP_CONST, F, true, false, {T_CURVE, F}},
    {OP_APPLY_GRADIENT, "APPLY_GRADIENT", 2, TP_CONST, TUP, true, false, {T_GRADIENT, F}},
    {OP_ORIG_VAL, "ORIG_VAL", 4, TP_CONST, TUP, true, false, {F, F, IMG, F}},
    {OP_RESIZE_IMAGE, "RESIZE_IMAGE", 3, TP_CONST, IMG, true, false, {IMG, F, F}},
    {OP_STRIP_RESIZE, "STRIP_RESIZE", 1, TP_CONST, IMG, true, false, {IMG}},
    {OP_RENDER, "RENDER", 3, TP_CONST, IMG, true, false, {IMG, I, I}},
    {OP_IMAGE_PIXEL_WIDTH, "IMAGE_PIXEL_WIDTH", 1, TP_CONST, I, true, false, {IMG}},
    {OP_IMAGE_PIXEL_HEIGHT, "IMAGE_PIXEL_HEIGHT", 1, TP_CONST, I, true, false, {IMG}},
    {OP_MAKE_RGBA_COLOR, "MAKE_COLOR", 4, TP_CONST, COL, true, false, {F, F, F, F}},
    {OP_RED, "RED_FLOAT", 1, TP_CONST, F, true, false, {COL}},
    {OP_GREEN, "GREEN_FLOAT", 1, TP_CONST, F, true, false, {COL}},
    {OP_BLUE, "BLUE_FLOAT", 1, TP_CONST, F, true, false, {COL}},
    {OP_ALPHA, "ALPHA_FLOAT", 1, TP_CONST, F, true, false, {COL}},
    {OP_TUPLE_NTH, "TUPLE_NTH", 2, TP_CONST, F, true, false, {TUP, I}},
    {OP_TREE_VECTOR_NTH, "TREE_VECTOR_NTH", 2, TP_CONST, F, true, false, {I, TV}},
    {OP_SET_TREE_VECTOR_NTH, "SET_TREE_VECTOR_NTH", 3, TP_CONST, TV, true, false, {I, TV, F}},
    {OP_COMPLEX, "COMPLEX", 2, TP_CONST, C, true, true, {F, F}},
    {OP_C_REAL, "crealf", 1, TP_CONST, F, true, true, {C}},
    {OP_C_IMAG, "cimagf", 1, TP_CONST, F, true, true, {C}},
    {OP_C_SQRT, "csqrtf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_SIN, "csinf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_COS, "ccosf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_TAN, "ctanf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_ASIN, "casinf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_ACOS, "cacosf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_ATAN, "catanf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_POW, "cpowf", 2, TP_CONST, C, true, true, {C, C}},
    {OP_C_EXP, "cexpf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_LOG, "clogf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_ARG, "cargf", 1, TP_CONST, F, true, true, {C}},
    {OP_C_SINH, "csinhf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_COSH, "ccoshf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_TANH, "ctanhf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_ASINH, "casinhf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_ACOSH, "cacoshf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_ATANH, "catanhf", 1, TP_CONST, C, true, true, {C}},
    {OP_C_GAMMA, "cgamma", 1, TP_CONST, C, true, false, {C}},
    {OP_ELL_INT_K_COMP, "ELL_INT_K_COMP", 1, TP_CONST, F, true, true, {F}},
    {OP_ELL_INT_E_COMP, "ELL_INT_E_COMP", 1, TP_CONST, F, true, true, {F}},
    {OP_ELL_INT_F, "ELL_INT_F", 2, TP_CONST, F, true, true, {F, F}},
    {OP_ELL_INT_E, "ELL_INT_E", 2, TP_CONST, F, true, true, {F, F}},
    {OP_ELL_INT_P, "ELL_INT_P", 3, TP_CONST, F, true, true, {F, F, F}},
    {OP_ELL_INT_D, "ELL_INT_D", 3, TP_CONST, F, true, true, {F, F, F}},
    {OP_ELL_INT_RC, "ELL_INT_RC", 2, TP_CONST, F, true, true, {F, F}},
    {OP_ELL_INT_RD, "ELL_INT_RD", 3, TP_CONST, F, true, true, {F, F, F}},
    {OP_ELL_INT_RF, "ELL_INT_RF", 3, TP_CONST, F, true, true, {F, F, F}},
    {OP_ELL_INT_RJ, "ELL_INT_RJ", 4, TP_CONST, F, true, true, {F, F, F, F}},
    {OP_ELL_JAC, "ELL_JAC", 2, TP_CONST, TUP, true, false, {F, F}},
    {OP_SOLVE_LINEAR_2, "SOLVE_LINEAR_2", 2, TP_CONST, TUP, true, false, {TUP, TUP}},
    {OP_SOLVE_LINEAR_3, "SOLVE_LINEAR_3", 2, TP_CONST, TUP, true, false, {TUP, TUP}},
    {OP_SOLVE_POLY_2, "SOLVE_POLY_2", 3, TP_CONST, TUP, true, false, {F, F, F}},
    {OP_SOLVE_POLY_3, "SOLVE_POLY_3", 4, TP_CONST, TUP, true, false, {F, F, F, F}},
    {OP_RAND, "RAND", 2, TP_CONST, F, false, false, {F, F}},
    {OP_LIBNOISE_PERLIN, "libnoise_perlin", 6, TP_CONST, F, true, false, {I, F, F, F, F, F}},
    {OP_LIBNOISE_BILLOW, "libnoise_billow", 6, TP_CONST, F, true, false, {I, F, F, F, F, F}},
    {OP_LIBNOISE_RIDGED_MULTI, "libnoise_ridged_multi", 5, TP_CONST, F, true, false, {I, F, F, F, F}},
    {OP_LIBNOISE_VORONOI, "libnoise_voronoi", 4, TP_CONST, F, true, false, {F, F, F, F}},
    {OP_USERVAL_INT, "USERVAL_INT_ACCESS", 1, TP_CONST, I, true, false, {I}},
    {OP_USERVAL_FLOAT, "USERVAL_FLOAT_ACCESS", 1, TP_CONST, F, true, false, {I}},
    {OP_USERVAL_BOOL, "USERVAL_BOOL_ACCESS", 1, TP_CONST, I, true, false, {I}},
    {OP_USERVAL_COLOR, "USERVAL_COLOR_ACCESS", 1, TP_CONST, COL, true, false, {I}},
    {OP_USERVAL_CURVE, "USERVAL_CURVE_ACCESS", 1, TP_CONST, T_CURVE, true, false, {I}},
    {OP_USERVAL_GRADIENT, "USERVAL_GRADIENT_ACCESS", 1, TP_CONST, T_GRADIENT, true, false, {I}},
    {OP_USERVAL_IMAGE, "USERVAL_IMAGE_ACCESS", 1, TP_CONST, IMG, true, false, {I}},
    {OP_OUTPUT_TUPLE, "OUTPUT_TUPLE", 1, TP_CONST, I, false, false, {TUP}},
};
#undef F
#undef I
#undef C
#undef IMG
#undef TUP
#undef COL
#undef TV

const OpInfo *op_info(int op) {
    assert(op >= 0 && op < NUM_OPS && k_ops[op].id == op);
    return &k_ops[op];
}
const OpInfo *op_by_name(const std::string &name) {
    static std::unordered_map<std::string, const OpInfo *> m;
    if (m.empty())
        for (int i = 0; i < NUM_OPS; ++i) m[k_ops[i].name] = &k_ops[i];
    auto it = m.find(name);
    return it == m.end() ? nullptr : it->second;
}

Internal *Filter::lookup_internal(const std::string &n, bool touch) {
    for (auto &i : internals)
        if (i.name == n) {
            if (touch) i.used = true;
            return &i;
        }
    return nullptr;
}
bool Filter::uses_ra() const {
    for (auto &i : internals)
        if ((i.name == "r" || i.name == "a") && i.used) return true;
    return false;
}
bool Filter::uses_t() const {
    for (auto &i : internals)
        if (i.name == "t" && i.used) return true;
    return false;
}

void add_use(Value *v, Stmt *s) { v->uses.push_back(s); }
void remove_use(Value *v, Stmt *s) {
    for (size_t i = 0; i < v->uses.size(); ++i)
        if (v->uses[i] == s) {
            v->uses.erase(v->uses.begin() + i);
            return;
        }
    assert(!"use not found");
}
Stmt *last_stmt(Stmt *s) {
    if (!s) return nullptr;
    while (s->next) s = s->next;
    return s;
}

Type primary_type(const Primary &p) { return p.is_const ? p.c.type : p.value->cv->type; }

// reference compiler.c:2752-2800 (rhs_type)
Type rhs_type(const Rhs *rhs) {
    switch (rhs->kind) {
    case RHS_PRIMARY: return primary_type(rhs->prim);
    case RHS_INTERNAL: return T_FLOAT;
    case RHS_OP:
        if (rhs->op->prop == TP_CONST) return rhs->op->type;
        {
            int mx = T_INT;
            for (auto &a : rhs->args) mx = std::max(mx, (int)primary_type(a));
            return (Type)mx;
        }
    case RHS_FILTER:
    case RHS_TUPLE: return T_TUPLE;
    case RHS_CLOSURE: return T_IMAGE;
    case RHS_TREE_VECTOR: return T_TREE_VECTOR;
    }
    return T_NIL;
}

int tuple_length_of_rhs(const Rhs *rhs) {
    switch (rhs->kind) {
    case RHS_TUPLE:
    case RHS_TREE_VECTOR: return (int)rhs->args.size();
    case RHS_FILTER: return 4;
    case RHS_OP:
        switch (rhs->op->id) {
        case OP_ORIG_VAL:
        case OP_APPLY_GRADIENT: return 4;
        case OP_ELL_JAC: return 3;
        case OP_SOLVE_LINEAR_2:
        case OP_SOLVE_POLY_2: return 2;
        case OP_SOLVE_LINEAR_3:
        case OP_SOLVE_POLY_3: return 3;
        default: return 0;
        }
    default: return 0;
    }
}

std::string format_float(float f) {
    if (std::isnan(f)) return "nan";
    if (std::isinf(f)) return f > 0 ? "inf" : "-inf";
    char buf[64];
    snprintf(buf, sizeof buf, "%.9g", (double)f);
    // make sure it reads back as a float literal, not an int
    if (!strpbrk(buf, ".en")) strcat(buf, ".0");
    return buf;
}

std::string primary_to_string(const Primary &p) {
    char buf[128];
    if (!p.is_const) {
        if (p.value->index < 0)
            snprintf(buf, sizeof buf, "%%%d.u", p.value->cv->id);
        else
            snprintf(buf, sizeof buf, "%%%d.%d", p.value->cv->id, p.value->index);
        return buf;
    }
    switch (p.c.type) {
    case T_INT: snprintf(buf, sizeof buf, "i:%d", p.c.i); return buf;
    case T_FLOAT: return "f:" + format_float(p.c.f);
    case T_COMPLEX: return "c:" + format_float(p.c.c.real()) + "," + format_float(p.c.c.imag());
    case T_COLOR: snprintf(buf, sizeof buf, "k:%u", p.c.color); return buf;
    default: return "?";
    }
}

}  // namespace mm
