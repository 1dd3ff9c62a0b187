// mathmap_b200 — intermediate representation of a compiled filter.
//
// The shape mirrors what the reference backend boundary receives
// (filter_code_t / statement_t / rhs_t / primary_t / value_t / compvar_t,
// reference compiler-internals.h:60-235): structured SSA with explicit phi
// lists on if-exits and while-entries.  It is a fresh C++ data model, not a
// translation; the text form (ir_text.cpp) is what a reference-side
// backends/cuda.c serialises over the C ABI.
#pragma once
#include <complex>
#include <cstdint>
#include <deque>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace mm {

// Runtime types, in promotion order (reference ops.lisp:37-66).
enum Type : int {
    T_NIL = 0,
    T_INT = 1,
    T_FLOAT = 2,
    T_COMPLEX = 3,
    T_COLOR = 4,
    T_CURVE = 5,
    T_GRADIENT = 6,
    T_IMAGE = 7,
    T_TUPLE = 8,
    T_TREE_VECTOR = 9,
};
const char *type_name(Type t);
Type type_from_name(const std::string &s);

enum TypeProp { TP_CONST, TP_MAX, TP_MAX_FLOAT };

// Bits of "independent of …" (reference internals.h:31-34).
enum : int { CONST_NONE = 0, CONST_X = 1, CONST_Y = 2, CONST_T = 4, CONST_ALL = 7 };

struct OpInfo {
    int id;
    const char *name;  // interchange name == the reference's op->name
    int nargs;
    TypeProp prop;
    Type type;  // result type when prop == TP_CONST
    bool pure;
    bool foldable;
    Type argtypes[6];
};

// Operator ids.  Order is ours; the interchange key is the name.
enum Op : int {
    OP_NOP, OP_INT2FLOAT, OP_FLOAT2INT, OP_INT2COMPLEX, OP_FLOAT2COMPLEX,
    OP_ADD, OP_SUB, OP_NEG, OP_MUL, OP_DIV, OP_MOD,
    OP_ABS, OP_MIN, OP_MAX,
    OP_SQRT, OP_HYPOT, OP_SIN, OP_COS, OP_TAN, OP_ASIN, OP_ACOS, OP_ATAN, OP_ATAN2,
    OP_POW, OP_EXP, OP_LOG, OP_SINH, OP_COSH, OP_TANH, OP_ASINH, OP_ACOSH, OP_ATANH,
    OP_GAMMA, OP_BETA,
    OP_FLOOR, OP_CEIL, OP_EQ, OP_LESS, OP_LEQ, OP_NOT,
    OP_PRINT, OP_NEWLINE, OP_START_DEBUG_TUPLE, OP_SET_DEBUG_TUPLE_DATA,
    OP_APPLY_CURVE, OP_APPLY_GRADIENT, OP_ORIG_VAL,
    OP_RESIZE_IMAGE, OP_STRIP_RESIZE, OP_RENDER,
    OP_IMAGE_PIXEL_WIDTH, OP_IMAGE_PIXEL_HEIGHT,
    OP_MAKE_RGBA_COLOR, OP_RED, OP_GREEN, OP_BLUE, OP_ALPHA,
    OP_TUPLE_NTH, OP_TREE_VECTOR_NTH, OP_SET_TREE_VECTOR_NTH,
    OP_COMPLEX, OP_C_REAL, OP_C_IMAG, OP_C_SQRT, OP_C_SIN, OP_C_COS, OP_C_TAN,
    OP_C_ASIN, OP_C_ACOS, OP_C_ATAN, OP_C_POW, OP_C_EXP, OP_C_LOG, OP_C_ARG,
    OP_C_SINH, OP_C_COSH, OP_C_TANH, OP_C_ASINH, OP_C_ACOSH, OP_C_ATANH, OP_C_GAMMA,
    OP_ELL_INT_K_COMP, OP_ELL_INT_E_COMP, OP_ELL_INT_F, OP_ELL_INT_E, OP_ELL_INT_P,
    OP_ELL_INT_D, OP_ELL_INT_RC, OP_ELL_INT_RD, OP_ELL_INT_RF, OP_ELL_INT_RJ, OP_ELL_JAC,
    OP_SOLVE_LINEAR_2, OP_SOLVE_LINEAR_3, OP_SOLVE_POLY_2, OP_SOLVE_POLY_3,
    OP_RAND,
    OP_LIBNOISE_PERLIN, OP_LIBNOISE_BILLOW, OP_LIBNOISE_RIDGED_MULTI, OP_LIBNOISE_VORONOI,
    OP_USERVAL_INT, OP_USERVAL_FLOAT, OP_USERVAL_BOOL, OP_USERVAL_COLOR,
    OP_USERVAL_CURVE, OP_USERVAL_GRADIENT, OP_USERVAL_IMAGE,
    OP_OUTPUT_TUPLE,
    NUM_OPS
};
const OpInfo *op_info(int op);
const OpInfo *op_by_name(const std::string &name);

struct Stmt;
struct Value;
struct Filter;

// A "compvar": one scalar slot of a user variable or a compiler temporary.
// The runtime type is a property of the compvar, shared by all of its SSA
// values (reference compiler.c:2752-2865).
struct CompVar {
    int id = 0;
    Type type = T_INT;
    std::string label;  // "name[i]" for user variables, "" for temporaries
    std::vector<Value *> values;
    Value *current = nullptr;
    int tuple_len = 0;  // for T_TUPLE compvars, resolved before emission
};

struct Value {
    CompVar *cv = nullptr;
    int index = -1;  // -1: never assigned (reads as 0)
    Stmt *def = nullptr;
    std::vector<Stmt *> uses;
    int const_bits = CONST_NONE;  // filled by analyze_constants
    bool hoisted = false;         // level == 0: evaluated once per frame on the host
    int level = 3;                // 0 frame, 1 row, 3 pixel (see analyze_constants)
};

struct Const {
    Type type = T_INT;
    int i = 0;
    float f = 0.f;
    std::complex<float> c;
    uint32_t color = 0;
};

struct Primary {
    bool is_const = false;
    Value *value = nullptr;
    Const c;
    static Primary of(Value *v) { Primary p; p.value = v; return p; }
    static Primary ic(int i) { Primary p; p.is_const = true; p.c.type = T_INT; p.c.i = i; return p; }
    static Primary fc(float f) { Primary p; p.is_const = true; p.c.type = T_FLOAT; p.c.f = f; return p; }
    static Primary cc(std::complex<float> z) { Primary p; p.is_const = true; p.c.type = T_COMPLEX; p.c.c = z; return p; }
};

enum RhsKind { RHS_PRIMARY = 1, RHS_INTERNAL, RHS_OP, RHS_FILTER, RHS_CLOSURE, RHS_TUPLE, RHS_TREE_VECTOR };

struct InlineHistory {
    Filter *filter;
    std::shared_ptr<InlineHistory> next;
};

struct Rhs {
    RhsKind kind = RHS_PRIMARY;
    Primary prim;            // RHS_PRIMARY
    std::string internal;    // RHS_INTERNAL
    const OpInfo *op = nullptr;
    std::vector<Primary> args;  // op / filter / closure / tuple args
    Filter *filter = nullptr;   // RHS_FILTER / RHS_CLOSURE
    std::shared_ptr<InlineHistory> history;
};

enum StmtKind { ST_NIL = 0, ST_ASSIGN, ST_PHI, ST_IF, ST_WHILE };

struct Stmt {
    StmtKind kind = ST_NIL;
    // assign / phi
    Value *lhs = nullptr;
    Rhs *rhs = nullptr;
    Rhs *rhs2 = nullptr;
    Value *old_value = nullptr;
    // if
    Rhs *cond = nullptr;  // also the while invariant
    Stmt *cons = nullptr, *alt = nullptr, *exit = nullptr;
    // while
    Stmt *entry = nullptr, *body = nullptr;
    Stmt *parent = nullptr;
    Stmt *next = nullptr;
    int level = 3;  // if: level of the condition; while: level the whole loop runs at
};

// ---- filters (reference mathmap.h filter_t, userval.h userval_info_t) ----

enum UservalType { UV_INT = 0, UV_FLOAT, UV_BOOL, UV_COLOR, UV_CURVE, UV_GRADIENT, UV_IMAGE };
const char *userval_type_name(int t);

enum : unsigned { IMAGE_FLAG_UNIT = 1, IMAGE_FLAG_SQUARE = 2 };

struct UservalInfo {
    std::string name;
    int type = UV_INT;
    int index = 0;
    int int_min = 0, int_max = 0, int_default = 0;
    float float_min = 0, float_max = 0, float_default = 0;
    int bool_default = 0;
    unsigned image_flags = 0;
    std::string doc;
};

struct Expr;
enum FilterKind { FILTER_MATHMAP, FILTER_NATIVE };

struct Internal {
    std::string name;
    int const_bits;
    bool used = false;
};

struct Variable {
    std::string name;
    int tag = 0, length = 0;
    std::vector<CompVar *> compvars;
    // subscripted somewhere by a computed index: held as ONE tree-vector compvar (compvars[0]) instead of one
    // compvar per element (compiler.c:2521-2570 find_all_vector_variables)
    bool is_vector = false;
};

struct Filter {
    FilterKind kind = FILTER_MATHMAP;
    std::string name;
    std::string doc;
    unsigned flags = IMAGE_FLAG_UNIT | IMAGE_FLAG_SQUARE;
    std::vector<UservalInfo> uservals;
    // mathmap filters
    std::vector<Internal> internals;
    std::vector<std::unique_ptr<Variable>> variables;
    Expr *body = nullptr;
    // native filters
    std::string native_name;
    int index = 0;  // position in the module's filter list
    Internal *lookup_internal(const std::string &n, bool touch);
    bool uses_ra() const;
    bool uses_t() const;
};

// Everything produced by compiling one filter: the statement list plus the
// arenas that own it.
struct FilterCode {
    Filter *filter = nullptr;
    Stmt *first = nullptr;
    std::deque<Stmt> stmts;
    std::deque<Rhs> rhss;
    std::deque<Value> values;
    std::deque<CompVar> compvars;
    Stmt *new_stmt(StmtKind k) { stmts.emplace_back(); stmts.back().kind = k; return &stmts.back(); }
    Rhs *new_rhs(RhsKind k) { rhss.emplace_back(); rhss.back().kind = k; return &rhss.back(); }
    Value *new_value(CompVar *cv) { values.emplace_back(); Value *v = &values.back(); v->cv = cv; cv->values.push_back(v); return v; }
    CompVar *new_compvar(Type t, const std::string &label = "") {
        compvars.emplace_back(); CompVar *c = &compvars.back(); c->id = (int)compvars.size(); c->type = t; c->label = label; return c;
    }
};

struct CompileError {
    std::string message;
    int line = -1, column = -1;
};

// helpers over rhs / statements
template <class F> void for_each_value_in_rhs(Rhs *rhs, F f) {
    if (!rhs) return;
    switch (rhs->kind) {
    case RHS_PRIMARY: if (!rhs->prim.is_const) f(rhs->prim.value); break;
    case RHS_INTERNAL: break;
    default: for (auto &a : rhs->args) if (!a.is_const) f(a.value); break;
    }
}
template <class F> void for_each_primary_in_rhs(Rhs *rhs, F f) {
    if (!rhs) return;
    switch (rhs->kind) {
    case RHS_PRIMARY: f(rhs->prim); break;
    case RHS_INTERNAL: break;
    default: for (auto &a : rhs->args) f(a); break;
    }
}
void add_use(Value *v, Stmt *s);
void remove_use(Value *v, Stmt *s);
Stmt *last_stmt(Stmt *s);
Type primary_type(const Primary &p);
Type rhs_type(const Rhs *rhs);
int tuple_length_of_rhs(const Rhs *rhs);  // 0 if not a tuple producer

std::string format_float(float f);  // round-trippable text of a float32
std::string dump_ir(const FilterCode &code);  // one (filter ...) form
std::string dump_module_ir(const std::vector<const FilterCode *> &codes, const std::string &main_name);
std::string primary_to_string(const Primary &p);

}  // namespace mm
