// IR generation: structured-SSA builder, tree -> IR lowering, filter prologue
// bindings, and the helper vocabulary builtins.cpp is written in.
//
// Behaviour restated from reference compiler.c: SSA builder with lazily created
// phis (:1013-1378), gen_code (:1876-2226), coordinate/limit/userval bindings
// (:2229-2511), gen_filter_code (:2611-2664).
#pragma once
#include <functional>

#include "frontend.h"

namespace mm {

typedef Primary P;

struct GenArgs {
    std::vector<std::vector<CompVar *>> args;
    std::vector<int> lengths, tags;
    std::vector<CompVar *> result;  // the temporaries the body assigns
};

class Gen {
   public:
    Gen(Module &m, FilterCode &c) : mod(m), code(c) {}
    Module &mod;
    FilterCode &code;
    Filter *filter = nullptr;  // filter whose body is being lowered (changes while inlining)

    // ---- SSA builder ----
    CompVar *temp(Type t = T_INT) { CompVar *cv = code.new_compvar(t); cv->current = code.new_value(cv); return cv; }
    P cur(CompVar *cv) { return P::of(cv->current); }
    Rhs *rhs_prim(P p);
    Rhs *rhs_internal(const std::string &name);
    Rhs *rhs_op(int op, std::vector<P> args);
    Rhs *rhs_tuple(std::vector<P> args);
    Rhs *rhs_closure(Filter *f, std::vector<P> args);
    Rhs *rhs_filter(Filter *f, std::vector<P> args);
    void assign(CompVar *cv, Rhs *rhs);   // new SSA value of cv
    void assign_value(Value *v, Rhs *rhs);
    void start_if(Rhs *cond);
    void switch_branch();
    void end_if();
    void start_while(CompVar *invariant);
    void end_while();
    void emit_nil();
    Stmt **emit_loc = nullptr;
    std::vector<Stmt *> stack;
    std::shared_ptr<InlineHistory> history;
    // insertion used by passes
    Stmt **emit_before(Stmt *stmt, Stmt **loc, Stmt *parent);
    Stmt *make_assign_stmt(CompVar *cv, Rhs *rhs);
    void commit_assign(Stmt *stmt);

    // ---- builtin vocabulary ----
    static P ic(int i) { return P::ic(i); }
    static P fc(float f) { return P::fc(f); }
    P op(int o, std::vector<P> args, Type t = T_INT) { CompVar *cv = temp(t); assign(cv, rhs_op(o, std::move(args))); return cur(cv); }
    P op1(int o, P a) { return op(o, {a}); }
    P op2(int o, P a, P b) { return op(o, {a, b}); }
    P add(P a, P b) { return op2(OP_ADD, a, b); }
    P sub(P a, P b) { return op2(OP_SUB, a, b); }
    P mul(P a, P b) { return op2(OP_MUL, a, b); }
    P div(P a, P b) { return op2(OP_DIV, a, b); }
    P neg(P a) { return op1(OP_NEG, a); }
    // left fold, as the reference's n-ary + and * expand (builtins.lisp:116-118)
    P sum(std::vector<P> v) { P r = v[0]; for (size_t i = 1; i < v.size(); ++i) r = add(r, v[i]); return r; }
    P prod(std::vector<P> v) { P r = v[0]; for (size_t i = 1; i < v.size(); ++i) r = mul(r, v[i]); return r; }
    CompVar *let(P p) { CompVar *cv = temp(); assign(cv, rhs_prim(p)); return cv; }
    void set(CompVar *cv, P p) { assign(cv, rhs_prim(p)); }
    P internal(const std::string &name);  // honours prologue bindings

    typedef std::function<void(CompVar *)> Cond;
    Cond c_op(int o, P a, P b) { return [=](CompVar *cv) { assign(cv, rhs_op(o, {a, b})); }; }
    Cond c_eq(P a, P b) { return c_op(OP_EQ, a, b); }
    Cond c_less(P a, P b) { return c_op(OP_LESS, a, b); }
    Cond c_leq(P a, P b) { return c_op(OP_LEQ, a, b); }
    Cond c_and(Cond a, Cond b) {
        return [=](CompVar *cv) { a(cv); start_if(rhs_prim(cur(cv))); b(cv); switch_branch(); end_if(); };
    }
    Cond c_or(Cond a, Cond b) {
        return [=](CompVar *cv) { a(cv); start_if(rhs_prim(cur(cv))); switch_branch(); b(cv); end_if(); };
    }
    Cond c_not(Cond a) {
        return [=](CompVar *cv) { a(cv); assign(cv, rhs_op(OP_NOT, {cur(cv)})); };
    }
    void if_(Cond c, std::function<void()> then_fn, std::function<void()> else_fn) {
        CompVar *cv = temp();
        c(cv);
        start_if(rhs_prim(cur(cv)));
        then_fn();
        switch_branch();
        else_fn();
        end_if();
    }

    // ---- tree lowering ----
    struct Binding {
        int kind;  // 0 internal, 1 userval
        const void *key;
        Value *value;
    };
    std::vector<Binding> bindings;
    Value *lookup_binding(int kind, const void *key);
    void gen_code(Expr *tree, CompVar **dest, bool alloced);
    CompVar *gen_tree_vector(Expr *tree, CompVar **dest, bool alloced);
    Stmt *gen_filter_code(Filter *f, CompVar *tuple, const std::vector<P> *args, Rhs **tuple_rhs,
                          std::shared_ptr<InlineHistory> hist);

   private:
    Value *new_lhs(CompVar *cv) { return code.new_value(cv); }
    void emit_stmt(Stmt *s);
    void record_def_uses(Stmt *s);
    void make_current(Value *v);
    Value *internal_value(const std::string &name, bool allow_bindings);
    Value *resize_image_if_necessary(P image, unsigned flags);
    void bind_limits();
    void bind_xy(Value *x, Value *y);
    void bind_from_uservals();
    void bind_from_args(const std::vector<P> &args);
    void bind_ra();
    void bind_internal(const std::string &name, Rhs *rhs, Type t = T_INT);
    void gen_args(std::vector<Expr *> &trees, GenArgs &ga);
};

}  // namespace mm
