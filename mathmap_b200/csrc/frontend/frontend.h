// MathMap language front end: tokens, typed expression trees, overload table.
//
// Behavioural contract (what must stay unchanged for .mm filters to mean the
// same thing): reference scanner.c:219-392 (tokens), parser.y:62-267 (grammar
// and precedences), exprtree.c (typing rules of every node), overload.c:212-279
// (first-match overload resolution with tag/length unification),
// macros.c:96-195 (xy ra XY WH I pi e and the __origVal macros).
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../ir/ir.h"

namespace mm {

struct TupleInfo {
    int tag = 0;
    int length = 0;
};

enum ExprKind {
    EX_INT_CONST = 1, EX_FLOAT_CONST, EX_FUNC, EX_INTERNAL, EX_SEQUENCE, EX_ASSIGNMENT,
    EX_VARIABLE, EX_IF_THEN, EX_IF_THEN_ELSE, EX_WHILE, EX_DO_WHILE, EX_TUPLE, EX_SELECT,
    EX_CAST, EX_SUB_ASSIGNMENT, EX_USERVAL, EX_FILTER_CLOSURE
};

struct OverloadEntry;

struct Expr {
    ExprKind kind;
    TupleInfo result;
    int line = -1, column = -1;
    int int_const = 0;
    float float_const = 0.f;
    Internal *internal = nullptr;
    Variable *var = nullptr;
    const UservalInfo *userval = nullptr;
    const OverloadEntry *entry = nullptr;
    Filter *filter = nullptr;          // EX_FILTER_CLOSURE
    std::vector<Expr *> args;          // func args / tuple elems / closure args / subscripts
    Expr *a = nullptr, *b = nullptr, *c = nullptr;  // generic children
    int tagnum = 0;                    // cast
    bool vector_select = false;        // EX_SELECT with a computed subscript (compiler.c:2543-2553)
};

class Gen;  // IR generation context (irgen.h)
struct GenArgs;
typedef void (*BuiltinGen)(Gen &g, GenArgs &a);

struct OverloadPatternElem {
    // tag: >0 constant tag number, 0 wildcard, <0 variable id (-1..-26)
    int tag = 0;
    // length: >0 constant, 0 wildcard, <0 variable id
    int length = 0;
};

struct OverloadEntry {
    std::string name;
    OverloadPatternElem result;
    std::vector<OverloadPatternElem> args;
    BuiltinGen gen = nullptr;                              // builtin
    std::function<Expr *(std::vector<Expr *> &)> macro;    // macro
    std::string impl_name;
};

// One compilation unit: the filters of one .mm source plus the registries the
// reference keeps in globals (tags.c, overload.c, macros.c).
struct Module {
    std::vector<std::unique_ptr<Filter>> filters;
    Filter *main_filter = nullptr;
    std::vector<std::string> tags;  // tags[n-1] is the name of tag number n
    std::vector<OverloadEntry> overloads;
    std::deque<Expr> exprs;
    int gensym_counter = 0;
    int nil_tag, xy_tag, ra_tag, rgba_tag, ri_tag, image_tag, curve_tag, gradient_tag;

    Module();
    int tag_number(const std::string &name);
    const std::string &tag_name(int n) const { return tags[n - 1]; }
    Filter *lookup_filter(const std::string &name);
    Expr *new_expr(ExprKind k) { exprs.emplace_back(); exprs.back().kind = k; return &exprs.back(); }
    // overload registration: spec is "RESULT <- ARG, ARG" with each item TAG:LEN
    void register_builtin(const char *name, const char *impl, const char *spec, BuiltinGen gen);
    void register_macro(const char *name, const char *spec, std::function<Expr *(std::vector<Expr *> &)> fn);
    const OverloadEntry *resolve(const std::string &name, const std::vector<TupleInfo> &args, TupleInfo *result) const;
};

void register_all_builtins(Module &m);  // builtins.cpp

void register_native_filters(Module &m);
// Parses `source` into `m` (filters, main filter).  Throws CompileError.
void parse_module(Module &m, const std::string &source);
// A composition (".mmc" design text) -> MathMap source; node types are looked up by main-filter name among the
// .mm / .mmc files under `filter_search_path` (design.cpp).  Throws CompileError.
std::string design_to_source(const std::string &design_text, const std::string &filter_search_path);

}  // namespace mm
