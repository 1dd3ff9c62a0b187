// Structured-SSA construction and tree -> IR lowering.  See irgen.h for the
// reference locations each part restates.
#include "irgen.h"

#include <cassert>
#include <cmath>

namespace mm {

[[noreturn]] static void gen_fail(const std::string &msg, const Expr *at = nullptr) {
    CompileError e;
    e.message = msg;
    if (at) { e.line = at->line; e.column = at->column; }
    throw e;
}

// ---------------------------------------------------------------- rhs makers
Rhs *Gen::rhs_prim(P p) { Rhs *r = code.new_rhs(RHS_PRIMARY); r->prim = p; return r; }
Rhs *Gen::rhs_internal(const std::string &name) { Rhs *r = code.new_rhs(RHS_INTERNAL); r->internal = name; return r; }
Rhs *Gen::rhs_op(int op, std::vector<P> args) {
    Rhs *r = code.new_rhs(RHS_OP);
    r->op = op_info(op);
    assert((int)args.size() == r->op->nargs);
    r->args = std::move(args);
    return r;
}
Rhs *Gen::rhs_tuple(std::vector<P> args) { Rhs *r = code.new_rhs(RHS_TUPLE); r->args = std::move(args); return r; }
Rhs *Gen::rhs_closure(Filter *f, std::vector<P> args) {
    Rhs *r = code.new_rhs(RHS_CLOSURE);
    r->filter = f;
    r->args = std::move(args);
    r->history = history;
    return r;
}
Rhs *Gen::rhs_filter(Filter *f, std::vector<P> args) {
    Rhs *r = code.new_rhs(RHS_FILTER);
    r->filter = f;
    r->args = std::move(args);
    r->history = history;
    return r;
}

// --------------------------------------------------------------- SSA builder
void Gen::record_def_uses(Stmt *s) {
    switch (s->kind) {
    case ST_NIL: break;
    case ST_ASSIGN:
        for_each_value_in_rhs(s->rhs, [&](Value *v) { add_use(v, s); });
        s->lhs->def = s;
        break;
    case ST_PHI:
        for_each_value_in_rhs(s->rhs, [&](Value *v) { add_use(v, s); });
        for_each_value_in_rhs(s->rhs2, [&](Value *v) { add_use(v, s); });
        s->lhs->def = s;
        break;
    case ST_IF:
    case ST_WHILE:
        for_each_value_in_rhs(s->cond, [&](Value *v) { add_use(v, s); });
        break;
    }
}

void Gen::emit_stmt(Stmt *s) {
    assert(emit_loc);
    s->parent = stack.empty() ? nullptr : stack.back();
    s->next = *emit_loc;
    *emit_loc = s;
    emit_loc = &s->next;
    record_def_uses(s);
}

void Gen::emit_nil() { emit_stmt(code.new_stmt(ST_NIL)); }

void Gen::make_current(Value *v) {
    CompVar *cv = v->cv;
    int mx = 0;
    for (Value *o : cv->values) mx = std::max(mx, o->index);
    if (v->index < 0) v->index = mx + 1;
    cv->current = v;
}

static Stmt *find_phi(Stmt *list, CompVar *cv) {
    for (; list; list = list->next)
        if (list->kind == ST_PHI && list->lhs->cv == cv) return list;
    return nullptr;
}

static bool within_limit(Stmt *s, Stmt *limit) {
    if (!limit) return true;
    do s = s->parent; while (s && s != limit);
    return s != nullptr;
}

static void rewrite_primary_in_rhs(Rhs *rhs, Value *old, P nw, Stmt *stmt) {
    auto fix = [&](P &p) {
        if (!p.is_const && p.value == old) {
            remove_use(old, stmt);
            p = nw;
            if (!nw.is_const) add_use(nw.value, stmt);
        }
    };
    if (!rhs) return;
    if (rhs->kind == RHS_PRIMARY) fix(rhs->prim);
    else if (rhs->kind != RHS_INTERNAL)
        for (auto &a : rhs->args) fix(a);
}

void rewrite_use(Stmt *stmt, Value *old, P nw) {
    switch (stmt->kind) {
    case ST_ASSIGN: rewrite_primary_in_rhs(stmt->rhs, old, nw, stmt); break;
    case ST_PHI:
        rewrite_primary_in_rhs(stmt->rhs, old, nw, stmt);
        rewrite_primary_in_rhs(stmt->rhs2, old, nw, stmt);
        break;
    case ST_IF:
    case ST_WHILE: rewrite_primary_in_rhs(stmt->cond, old, nw, stmt); break;
    default: break;
    }
}

static void rewrite_uses_within(Value *old, Value *nw, Stmt *limit) {
    if (old == nw) return;
    std::vector<Stmt *> users = old->uses;  // snapshot; rewriting edits the list
    for (Stmt *s : users) {
        bool still = false;
        for (Stmt *u : old->uses) if (u == s) { still = true; break; }
        if (!still) continue;
        // phis on the entry of the loop being built keep their pre-loop operand
        if (within_limit(s, limit) && !(s->kind == ST_PHI && s->parent == limit)) rewrite_use(s, old, P::of(nw));
    }
}

void Gen::commit_assign(Stmt *stmt) {
    if (!stack.empty()) {
        Stmt *tos = stack.back();
        CompVar *cv = stmt->lhs->cv;
        if (tos->kind == ST_IF) {
            Stmt *phi = find_phi(tos->exit, cv);
            if (!phi) {
                phi = code.new_stmt(ST_PHI);
                phi->lhs = new_lhs(cv);
                phi->rhs = rhs_prim(cur(cv));
                add_use(cv->current, phi);
                phi->rhs2 = rhs_prim(cur(cv));
                add_use(cv->current, phi);
                phi->old_value = cv->current;
                phi->lhs->def = phi;
                phi->parent = tos;
                phi->next = tos->exit;
                tos->exit = phi;
            }
            Rhs *&slot = tos->alt == nullptr ? phi->rhs : phi->rhs2;
            assert(slot->kind == RHS_PRIMARY && !slot->prim.is_const);
            remove_use(slot->prim.value, phi);
            slot = rhs_prim(P::of(stmt->lhs));
            add_use(stmt->lhs, phi);
        } else {
            assert(tos->kind == ST_WHILE);
            Stmt *phi = find_phi(tos->entry, cv);
            if (!phi) {
                phi = code.new_stmt(ST_PHI);
                phi->lhs = new_lhs(cv);
                phi->rhs = rhs_prim(cur(cv));
                add_use(cv->current, phi);
                phi->lhs->def = phi;
                phi->parent = tos;
                phi->next = tos->entry;
                tos->entry = phi;
                phi->rhs2 = rhs_prim(P::of(stmt->lhs));
                add_use(stmt->lhs, phi);
                phi->old_value = cv->current;
                // reads of the pre-loop value inside the loop now see the phi
                Value *pre = cv->current;
                // give the phi its index now so dumps are readable
                rewrite_uses_within(pre, phi->lhs, tos);
            } else {
                assert(phi->rhs2->kind == RHS_PRIMARY && !phi->rhs2->prim.is_const);
                remove_use(phi->rhs2->prim.value, phi);
                phi->rhs2 = rhs_prim(P::of(stmt->lhs));
                add_use(stmt->lhs, phi);
            }
        }
    }
    make_current(stmt->lhs);
}

Stmt *Gen::make_assign_stmt(CompVar *cv, Rhs *rhs) {
    Stmt *s = code.new_stmt(ST_ASSIGN);
    s->lhs = new_lhs(cv);
    s->rhs = rhs;
    return s;
}

void Gen::assign(CompVar *cv, Rhs *rhs) {
    Stmt *s = make_assign_stmt(cv, rhs);
    emit_stmt(s);
    commit_assign(s);
}

void Gen::assign_value(Value *v, Rhs *rhs) {
    Stmt *s = code.new_stmt(ST_ASSIGN);
    s->lhs = v;
    s->rhs = rhs;
    emit_stmt(s);
    commit_assign(s);
}

Stmt **Gen::emit_before(Stmt *stmt, Stmt **loc, Stmt *parent) {
    stmt->parent = parent;
    stmt->next = *loc;
    *loc = stmt;
    record_def_uses(stmt);
    if (stmt->kind == ST_ASSIGN) make_current(stmt->lhs);
    return &stmt->next;
}

void Gen::start_if(Rhs *cond) {
    Stmt *s = code.new_stmt(ST_IF);
    s->cond = cond;
    emit_stmt(s);
    stack.push_back(s);
    emit_loc = &s->cons;
}

static void reset_values_for_phis(Stmt *phi, bool del) {
    for (; phi; phi = phi->next) {
        if (phi->kind != ST_PHI) continue;
        if (phi->old_value) phi->lhs->cv->current = phi->old_value;
        if (del) phi->old_value = nullptr;
    }
}

void Gen::switch_branch() {
    Stmt *s = stack.back();
    assert(s->kind == ST_IF && s->alt == nullptr);
    if (!s->cons) emit_nil();
    reset_values_for_phis(s->exit, false);
    emit_loc = &s->alt;
}

void Gen::end_if() {
    Stmt *s = stack.back();
    assert(s->kind == ST_IF && s->cons);
    if (!s->alt) emit_nil();
    if (!s->exit) {
        Stmt *nil = code.new_stmt(ST_NIL);
        nil->parent = s;
        s->exit = nil;
    }
    stack.pop_back();
    reset_values_for_phis(s->exit, true);
    for (Stmt *phi = s->exit; phi; phi = phi->next)
        if (phi->kind == ST_PHI) commit_assign(phi);
    emit_loc = &s->next;
}

void Gen::start_while(CompVar *inv) {
    Stmt *s = code.new_stmt(ST_WHILE);
    Stmt *phi = code.new_stmt(ST_PHI);
    phi->lhs = new_lhs(inv);
    phi->rhs = rhs_prim(cur(inv));
    phi->rhs2 = rhs_prim(cur(inv));
    add_use(inv->current, phi);
    add_use(inv->current, phi);
    phi->lhs->def = phi;
    phi->old_value = inv->current;
    make_current(phi->lhs);
    s->cond = rhs_prim(cur(inv));
    emit_stmt(s);
    stack.push_back(s);
    phi->parent = s;
    phi->next = s->entry;
    s->entry = phi;
    emit_loc = &s->body;
}

void Gen::end_while() {
    Stmt *s = stack.back();
    stack.pop_back();
    assert(s->kind == ST_WHILE);
    if (!s->body) {
        stack.push_back(s);
        emit_nil();
        stack.pop_back();
    }
    reset_values_for_phis(s->entry, true);
    for (Stmt *phi = s->entry; phi; phi = phi->next)
        if (phi->kind == ST_PHI) commit_assign(phi);
    emit_loc = &s->next;
}

// ------------------------------------------------------------------ bindings
Value *Gen::lookup_binding(int kind, const void *key) {
    for (auto it = bindings.rbegin(); it != bindings.rend(); ++it)
        if (it->kind == kind && it->key == key) return it->value;
    return nullptr;
}

Value *Gen::internal_value(const std::string &name, bool allow_bindings) {
    Internal *in = filter->lookup_internal(name, false);
    assert(in);
    if (allow_bindings)
        if (Value *v = lookup_binding(0, in)) return v;
    CompVar *t = temp();
    assign(t, rhs_internal(name));
    return t->current;
}

P Gen::internal(const std::string &name) { return P::of(internal_value(name, true)); }

void Gen::bind_internal(const std::string &name, Rhs *rhs, Type t) {
    Internal *in = filter->lookup_internal(name, false);
    assert(in);
    CompVar *cv = temp(t);
    assign(cv, rhs);
    bindings.push_back({0, in, cv->current});
}

static bool needs_xy_scaling(unsigned flags) { return (flags & (IMAGE_FLAG_UNIT | IMAGE_FLAG_SQUARE)) != IMAGE_FLAG_UNIT; }

// compiler.c:1716-1773
Value *Gen::resize_image_if_necessary(P image, unsigned flags) {
    CompVar *resized = temp(T_IMAGE);
    if (!needs_xy_scaling(flags)) {
        assign(resized, rhs_prim(image));
        return resized->current;
    }
    P pw = op1(OP_IMAGE_PIXEL_WIDTH, image), ph = op1(OP_IMAGE_PIXEL_HEIGHT, image);
    P xf, yf;
    if (flags == 0) {
        xf = div(ic(2), pw);
        yf = div(ic(2), ph);
    } else {
        assert(flags == (IMAGE_FLAG_UNIT | IMAGE_FLAG_SQUARE));
        P mx = op2(OP_MAX, pw, ph);
        xf = div(mx, pw);
        yf = div(mx, ph);
    }
    assign(resized, rhs_op(OP_STRIP_RESIZE, {image}));
    assign(resized, rhs_op(OP_RESIZE_IMAGE, {cur(resized), xf, yf}));
    return resized->current;
}

// compiler.c:2339-2401
void Gen::bind_limits() {
    switch (filter->flags & (IMAGE_FLAG_UNIT | IMAGE_FLAG_SQUARE)) {
    case 0: {
        bind_internal("W", rhs_prim(P::of(internal_value("__canvasPixelW", false))));
        bind_internal("H", rhs_prim(P::of(internal_value("__canvasPixelH", false))));
        P wm1 = sub(P::of(internal_value("__canvasPixelW", true)), ic(1));
        bind_internal("X", rhs_op(OP_DIV, {wm1, ic(2)}));
        P hm1 = sub(P::of(internal_value("__canvasPixelH", true)), ic(1));
        bind_internal("Y", rhs_op(OP_DIV, {hm1, ic(2)}));
        break;
    }
    case IMAGE_FLAG_UNIT:
        bind_internal("W", rhs_prim(ic(2)));
        bind_internal("H", rhs_prim(ic(2)));
        bind_internal("X", rhs_prim(ic(1)));
        bind_internal("Y", rhs_prim(ic(1)));
        break;
    default: {
        P mx = op2(OP_MAX, P::of(internal_value("__canvasPixelW", false)), P::of(internal_value("__canvasPixelH", false)));
        bind_internal("X", rhs_op(OP_DIV, {P::of(internal_value("__canvasPixelW", false)), mx}));
        P xv = P::of(bindings.back().value);
        bind_internal("Y", rhs_op(OP_DIV, {P::of(internal_value("__canvasPixelH", false)), mx}));
        P yv = P::of(bindings.back().value);
        bind_internal("W", rhs_op(OP_MUL, {xv, ic(2)}));
        bind_internal("H", rhs_op(OP_MUL, {yv, ic(2)}));
        break;
    }
    }
}

// compiler.c:2403-2420
void Gen::bind_xy(Value *x, Value *y) {
    bind_internal("x", rhs_op(OP_MUL, {P::of(x), P::of(internal_value("X", true))}));
    bind_internal("y", rhs_op(OP_MUL, {P::of(y), P::of(internal_value("Y", true))}));
}

static int userval_getter_op(int type) {
    switch (type) {
    case UV_INT: return OP_USERVAL_INT;
    case UV_FLOAT: return OP_USERVAL_FLOAT;
    case UV_BOOL: return OP_USERVAL_BOOL;
    case UV_CURVE: return OP_USERVAL_CURVE;
    case UV_GRADIENT: return OP_USERVAL_GRADIENT;
    case UV_IMAGE: return OP_USERVAL_IMAGE;
    default: return -1;  // colours have no single-variable representation
    }
}
static Type userval_var_type(int type) {
    switch (type) {
    case UV_INT: case UV_BOOL: return T_INT;
    case UV_FLOAT: return T_FLOAT;
    case UV_CURVE: return T_CURVE;
    case UV_GRADIENT: return T_GRADIENT;
    case UV_IMAGE: return T_IMAGE;
    default: return T_INT;
    }
}

// compiler.c:2248-2281
void Gen::bind_from_uservals() {
    for (const UservalInfo &u : filter->uservals) {
        int getter = userval_getter_op(u.type);
        if (getter < 0) continue;
        CompVar *cv = temp(userval_var_type(u.type));
        if (u.type == UV_IMAGE) {
            CompVar *image = temp(T_IMAGE);
            assign(image, rhs_op(getter, {ic(u.index)}));
            Value *resized = resize_image_if_necessary(cur(image), u.image_flags);
            assign(cv, rhs_prim(P::of(resized)));
        } else
            assign(cv, rhs_op(getter, {ic(u.index)}));
        bindings.push_back({1, &u, cv->current});
    }
}

// compiler.c:2422-2466
void Gen::bind_from_args(const std::vector<P> &args) {
    int n = (int)args.size();
    assert(n == (int)filter->uservals.size() + 3);
    for (int i = 0; i < n - 3; ++i) {
        const UservalInfo &u = filter->uservals[i];
        if (userval_getter_op(u.type) < 0) gen_fail("cannot inline filter " + filter->name + ": colour arguments");
        CompVar *cv = temp(userval_var_type(u.type));
        if (u.type == UV_IMAGE)
            assign(cv, rhs_prim(P::of(resize_image_if_necessary(args[i], u.image_flags))));
        else
            assign(cv, rhs_prim(args[i]));
        bindings.push_back({1, &u, cv->current});
    }
    CompVar *xt = temp();
    assign(xt, rhs_prim(args[n - 3]));
    CompVar *yt = temp();
    assign(yt, rhs_prim(args[n - 2]));
    bind_xy(xt->current, yt->current);
    bind_internal("t", rhs_prim(args[n - 1]));
}

// compiler.c:2468-2511
void Gen::bind_ra() {
    CompVar *r = temp(T_FLOAT), *a = temp(T_FLOAT), *x_over_r = temp(T_FLOAT);
    Value *x = internal_value("x", true), *y = internal_value("y", true);
    assign(r, rhs_op(OP_HYPOT, {P::of(x), P::of(y)}));
    start_if(rhs_op(OP_EQ, {cur(r), fc(0.0f)}));
    assign(a, rhs_prim(fc(0.0f)));
    switch_branch();
    assign(x_over_r, rhs_op(OP_DIV, {P::of(x), cur(r)}));
    assign(a, rhs_op(OP_ACOS, {cur(x_over_r)}));
    end_if();
    start_if(rhs_op(OP_LESS, {P::of(y), fc(0.0f)}));
    assign(a, rhs_op(OP_SUB, {fc((float)(2 * M_PI)), cur(a)}));
    switch_branch();
    end_if();
    Internal *ri = filter->lookup_internal("r", false), *ai = filter->lookup_internal("a", false);
    CompVar *rb = temp(T_FLOAT);
    assign(rb, rhs_prim(cur(r)));
    bindings.push_back({0, ri, rb->current});
    CompVar *ab = temp(T_FLOAT);
    assign(ab, rhs_prim(cur(a)));
    bindings.push_back({0, ai, ab->current});
}

// -------------------------------------------------------------- tree -> IR
void Gen::gen_args(std::vector<Expr *> &trees, GenArgs &ga) {
    for (Expr *t : trees) {
        std::vector<CompVar *> dst(t->result.length, nullptr);
        ga.lengths.push_back(t->result.length);
        ga.tags.push_back(t->result.tag);
        gen_code(t, dst.data(), false);
        ga.args.push_back(std::move(dst));
    }
}

static bool single_const(const Expr *e, int *iv) {
    if (e->kind == EX_INT_CONST) { *iv = e->int_const; return true; }
    if (e->kind == EX_FLOAT_CONST) { *iv = (int)e->float_const; return true; }
    return false;
}

// compiler.c:2521-2570: a variable read or written through a subscript that is not a literal becomes a "vector
// variable"; a select with such a subscript is marked so that its tuple operand is materialised as a tree vector
static void find_vector_variables(Expr *t) {
    if (!t) return;
    switch (t->kind) {
    case EX_SELECT:
        find_vector_variables(t->a);
        for (Expr *sub : t->args) {
            int dummy;
            find_vector_variables(sub);
            if (!single_const(sub, &dummy)) {
                t->vector_select = true;
                if (t->a->kind == EX_VARIABLE) t->a->var->is_vector = true;
            }
        }
        return;
    case EX_SUB_ASSIGNMENT:
        find_vector_variables(t->a);
        for (Expr *sub : t->args) {
            int dummy;
            find_vector_variables(sub);
            if (!single_const(sub, &dummy)) t->var->is_vector = true;
        }
        return;
    default:
        for (Expr *e : t->args) find_vector_variables(e);
        find_vector_variables(t->a);
        find_vector_variables(t->b);
        find_vector_variables(t->c);
    }
}

static CompVar *new_tree_vector_compvar(Gen &g, int length, const std::string &name) {
    CompVar *cv = g.code.new_compvar(T_TREE_VECTOR, name);
    cv->tuple_len = length;
    cv->current = g.code.new_value(cv);
    return cv;
}

static void alloc_var_compvars(Gen &g, Variable *v) {
    if (v->is_vector) {  // compiler.c:1776-1785
        if (v->compvars.size() != 1) v->compvars.assign(1, nullptr);
        if (!v->compvars[0]) v->compvars[0] = new_tree_vector_compvar(g, v->length, v->name);
        return;
    }
    if ((int)v->compvars.size() != v->length) v->compvars.assign(v->length, nullptr);
    for (int i = 0; i < v->length; ++i)
        if (!v->compvars[i]) {
            Type t = (v->tag == g.mod.image_tag && v->length == 1) ? T_IMAGE : T_INT;
            CompVar *cv = g.code.new_compvar(t, v->name + "[" + std::to_string(i) + "]");
            cv->current = g.code.new_value(cv);
            v->compvars[i] = cv;
        }
}

// compiler.c:1838-1873
CompVar *Gen::gen_tree_vector(Expr *tree, CompVar **dest, bool alloced) {
    if (tree->kind == EX_VARIABLE && tree->var->is_vector) {
        alloc_var_compvars(*this, tree->var);
        CompVar *tv = tree->var->compvars[0];
        for (int i = 0; i < tree->result.length; ++i) {
            if (!alloced) dest[i] = temp(T_FLOAT);
            assign(dest[i], rhs_op(OP_TREE_VECTOR_NTH, {ic(i), cur(tv)}));
        }
        return tv;
    }
    gen_code(tree, dest, alloced);
    std::vector<P> args;
    for (int i = 0; i < tree->result.length; ++i) args.push_back(cur(dest[i]));
    CompVar *tv = new_tree_vector_compvar(*this, tree->result.length, "");
    Rhs *r = code.new_rhs(RHS_TREE_VECTOR);
    r->args = std::move(args);
    assign(tv, r);
    return tv;
}

void Gen::gen_code(Expr *tree, CompVar **dest, bool alloced) {
    auto result_type = [&](const TupleInfo &ti) { return (ti.tag == mod.image_tag && ti.length == 1) ? T_IMAGE : T_INT; };
    switch (tree->kind) {
    case EX_INT_CONST:
        if (!alloced) dest[0] = temp(T_INT);
        assign(dest[0], rhs_prim(ic(tree->int_const)));
        break;
    case EX_FLOAT_CONST:
        if (!alloced) dest[0] = temp(T_FLOAT);
        assign(dest[0], rhs_prim(fc(tree->float_const)));
        break;
    case EX_TUPLE:
        for (size_t i = 0; i < tree->args.size(); ++i) gen_code(tree->args[i], dest + i, alloced);
        break;
    case EX_SELECT: {
        std::vector<CompVar *> temps(tree->a->result.length, nullptr);
        CompVar *tree_vector = nullptr;
        if (tree->vector_select) tree_vector = gen_tree_vector(tree->a, temps.data(), false);
        else gen_code(tree->a, temps.data(), false);
        for (size_t i = 0; i < tree->args.size(); ++i) {
            int sub;
            if (single_const(tree->args[i], &sub)) {
                sub = std::max(0, std::min(sub, tree->a->result.length - 1));
                if (!alloced) dest[i] = temps[sub];
                else assign(dest[i], rhs_prim(cur(temps[sub])));
            } else {  // compiler.c:1941-1955
                CompVar *subscript = nullptr;
                if (!alloced) dest[i] = temp(T_INT);
                gen_code(tree->args[i], &subscript, false);
                assign(dest[i], rhs_op(OP_TREE_VECTOR_NTH, {cur(subscript), cur(tree_vector)}));
            }
        }
        break;
    }
    case EX_VARIABLE:
        alloc_var_compvars(*this, tree->var);
        if (tree->var->is_vector) {  // compiler.c:1962-1971
            for (int i = 0; i < tree->var->length; ++i) {
                if (!alloced) dest[i] = temp(T_INT);
                assign(dest[i], rhs_op(OP_TREE_VECTOR_NTH, {ic(i), cur(tree->var->compvars[0])}));
            }
            break;
        }
        for (int i = 0; i < tree->var->length; ++i)
            if (!alloced) dest[i] = tree->var->compvars[i];
            else assign(dest[i], rhs_prim(cur(tree->var->compvars[i])));
        break;
    case EX_INTERNAL: {
        Value *bv = lookup_binding(0, tree->internal);
        if (!alloced) dest[0] = temp(T_INT);
        if (bv) assign(dest[0], rhs_prim(P::of(bv)));
        else assign(dest[0], rhs_internal(tree->internal->name));
        break;
    }
    case EX_ASSIGNMENT:
        alloc_var_compvars(*this, tree->var);
        if (tree->var->is_vector) {  // compiler.c:1995-1999
            CompVar *tv = gen_tree_vector(tree->a, dest, alloced);
            assign(tree->var->compvars[0], rhs_prim(cur(tv)));
            break;
        }
        gen_code(tree->a, tree->var->compvars.data(), true);
        for (int i = 0; i < tree->result.length; ++i)
            if (alloced) assign(dest[i], rhs_prim(cur(tree->var->compvars[i])));
            else dest[i] = tree->var->compvars[i];
        break;
    case EX_SUB_ASSIGNMENT: {
        std::vector<CompVar *> temps(tree->a->result.length, nullptr);
        alloc_var_compvars(*this, tree->var);
        gen_code(tree->a, temps.data(), false);
        for (size_t i = 0; i < tree->args.size(); ++i) {
            int sub;
            if (tree->var->is_vector) {  // compiler.c:2027-2038
                CompVar *tv = tree->var->compvars[0], *subscript = nullptr;
                gen_code(tree->args[i], &subscript, false);
                assign(tv, rhs_op(OP_SET_TREE_VECTOR_NTH, {cur(subscript), cur(tv), cur(temps[i])}));
            } else if (single_const(tree->args[i], &sub)) {
                sub = std::max(0, std::min(sub, tree->var->length - 1));
                assign(tree->var->compvars[sub], rhs_prim(cur(temps[i])));
            } else
                gen_fail("internal error: computed subscript on a variable that is not a vector variable", tree);
            if (alloced) assign(dest[i], rhs_prim(cur(temps[i])));
            else dest[i] = temps[i];
        }
        break;
    }
    case EX_CAST: gen_code(tree->a, dest, alloced); break;
    case EX_FUNC: {
        GenArgs ga;
        gen_args(tree->args, ga);
        if (!alloced)
            for (int i = 0; i < tree->result.length; ++i) dest[i] = temp(result_type(tree->result));
        // the body assigns fresh temporaries which are then copied to the
        // destination, like the reference's generated gen_* functions
        for (int i = 0; i < tree->result.length; ++i) ga.result.push_back(temp(dest[i]->type));
        tree->entry->gen(*this, ga);
        for (int i = 0; i < tree->result.length; ++i) assign(dest[i], rhs_prim(cur(ga.result[i])));
        break;
    }
    case EX_SEQUENCE: {
        std::vector<CompVar *> left(tree->a->result.length, nullptr);
        gen_code(tree->a, left.data(), false);
        gen_code(tree->b, dest, alloced);
        break;
    }
    case EX_IF_THEN:
    case EX_IF_THEN_ELSE: {
        std::vector<CompVar *> result(tree->result.length);
        for (auto &r : result) r = temp(result_type(tree->result));
        CompVar *condition = nullptr;
        gen_code(tree->a, &condition, false);
        start_if(rhs_prim(cur(condition)));
        gen_code(tree->b, result.data(), true);
        switch_branch();
        if (tree->kind == EX_IF_THEN_ELSE) gen_code(tree->c, result.data(), true);
        end_if();
        for (int i = 0; i < tree->result.length; ++i)
            if (alloced) assign(dest[i], rhs_prim(cur(result[i])));
            else dest[i] = result[i];
        break;
    }
    case EX_DO_WHILE:
    case EX_WHILE: {
        CompVar *invariant = temp(T_INT);
        std::vector<CompVar *> body_result(tree->b->result.length, nullptr);
        if (tree->kind == EX_DO_WHILE) gen_code(tree->b, body_result.data(), false);
        gen_code(tree->a, &invariant, true);
        start_while(invariant);
        std::fill(body_result.begin(), body_result.end(), nullptr);
        gen_code(tree->b, body_result.data(), false);
        gen_code(tree->a, &invariant, true);
        end_while();
        if (!alloced) dest[0] = temp(T_INT);
        assign(dest[0], rhs_prim(ic(0)));
        break;
    }
    case EX_USERVAL:
        if (tree->userval->type == UV_COLOR) {
            CompVar *t = temp(T_INT);
            assign(t, rhs_op(OP_USERVAL_COLOR, {ic(tree->userval->index)}));
            static const int color_ops[4] = {OP_RED, OP_GREEN, OP_BLUE, OP_ALPHA};
            for (int i = 0; i < 4; ++i) {
                if (!alloced) dest[i] = temp(T_FLOAT);
                assign(dest[i], rhs_op(color_ops[i], {cur(t)}));
            }
        } else {
            Value *bv = lookup_binding(1, tree->userval);
            assert(bv);
            if (!alloced) dest[0] = temp(userval_var_type(tree->userval->type));
            assign(dest[0], rhs_prim(P::of(bv)));
        }
        break;
    case EX_FILTER_CLOSURE: {
        GenArgs ga;
        gen_args(tree->args, ga);
        Filter *callee = tree->filter;
        std::vector<P> prims;
        for (size_t i = 0; i < callee->uservals.size(); ++i) {
            const UservalInfo &u = callee->uservals[i];
            if (u.type == UV_COLOR) {
                CompVar *c = temp(T_COLOR);
                assign(c, rhs_op(OP_MAKE_RGBA_COLOR, {cur(ga.args[i][0]), cur(ga.args[i][1]), cur(ga.args[i][2]), cur(ga.args[i][3])}));
                prims.push_back(cur(c));
            } else if (u.type == UV_IMAGE) {
                CompVar *c = temp(T_IMAGE);
                assign(c, rhs_op(OP_STRIP_RESIZE, {cur(ga.args[i][0])}));
                prims.push_back(cur(c));
            } else
                prims.push_back(cur(ga.args[i][0]));
        }
        CompVar *image = temp(T_IMAGE);
        if (!alloced) dest[0] = temp(T_IMAGE);
        assign(image, rhs_closure(callee, prims));
        Value *resized = resize_image_if_necessary(cur(image), filter->flags);
        assign(dest[0], rhs_prim(P::of(resized)));
        break;
    }
    }
}

// compiler.c:2611-2664
Stmt *Gen::gen_filter_code(Filter *f, CompVar *tuple, const std::vector<P> *args, Rhs **tuple_rhs,
                           std::shared_ptr<InlineHistory> hist) {
    Filter *filter_save = filter;
    Stmt **emit_loc_save = emit_loc;
    auto history_save = history;
    auto bindings_save = bindings;
    std::vector<Stmt *> stack_save;
    stack_save.swap(stack);

    filter = f;
    for (auto &v : f->variables) v->compvars.clear();
    find_vector_variables(f->body);
    history = std::make_shared<InlineHistory>(InlineHistory{f, hist});
    bindings.clear();

    Stmt *first = nullptr;
    emit_loc = &first;
    bind_limits();
    if (args)
        bind_from_args(*args);
    else {
        bind_from_uservals();
        if (needs_xy_scaling(f->flags)) bind_xy(internal_value("x", false), internal_value("y", false));
    }
    if (f->uses_ra()) bind_ra();

    std::vector<CompVar *> result(f->body->result.length, nullptr);
    gen_code(f->body, result.data(), false);
    Rhs *rhs = rhs_tuple({cur(result[0]), cur(result[1]), cur(result[2]), cur(result[3])});
    if (tuple_rhs) *tuple_rhs = rhs;
    if (tuple) assign(tuple, rhs);

    filter = filter_save;
    emit_loc = emit_loc_save;
    history = history_save;
    bindings = bindings_save;
    stack.swap(stack_save);
    return first;
}

}  // namespace mm
