// IR optimisation passes, type propagation, constness analysis and the
// per-filter compile driver.
//
// Pass list and order follow the reference driver (compiler.c:4703-4764):
// closure application, inlining, copy propagation, tuple-nth, make-tuple, CSE,
// copy propagation, constant folding, algebraic unit/zero simplification,
// ORIG_VAL/RESIZE_IMAGE lowering, STRIP_RESIZE, closure pixel-size, dead
// assignments, dead branches, dead controls — iterated to a fixpoint with no
// wall-clock timeout (the reference's 2 s budget makes its IR host-speed
// dependent; `--bench-no-compile-time-limit` is the deterministic behaviour).
// Then type propagation (compiler.c:2802-2865) and constness analysis.
#include "passes.h"

#include <cassert>
#include <cstdlib>
#include <functional>
#include <map>
#include <set>
#include <unordered_map>
#include <unordered_set>

#include "../frontend/irgen.h"
#include "eval.h"

namespace mm {

void rewrite_use(Stmt *stmt, Value *old, P nw);  // irgen.cpp

namespace {

typedef std::function<void(Stmt *)> StmtFn;

// Visits every statement (including phi lists) in program order.
void walk(Stmt *s, const StmtFn &f) {
    for (; s; s = s->next) {
        switch (s->kind) {
        case ST_IF:
            f(s);
            walk(s->cons, f);
            walk(s->alt, f);
            walk(s->exit, f);
            break;
        case ST_WHILE:
            walk(s->entry, f);
            f(s);
            walk(s->body, f);
            break;
        default: f(s); break;
        }
    }
}

bool rhs_is_pure(const Rhs *r) {
    if (r->kind == RHS_OP) return r->op->pure;
    if (r->kind == RHS_FILTER) return false;
    return true;
}

void remove_uses_in_rhs(Rhs *r, Stmt *s) {
    for_each_value_in_rhs(r, [&](Value *v) { remove_use(v, s); });
}
void add_uses_in_rhs(Rhs *r, Stmt *s) {
    for_each_value_in_rhs(r, [&](Value *v) { add_use(v, s); });
}
void replace_rhs(Rhs **slot, Rhs *nw, Stmt *s) {
    remove_uses_in_rhs(*slot, s);
    *slot = nw;
    add_uses_in_rhs(nw, s);
}
bool is_assign_with_op(const Stmt *s, int op) {
    return s && s->kind == ST_ASSIGN && s->rhs->kind == RHS_OP && s->rhs->op->id == op;
}

struct Optimizer {
    Module &mod;
    Gen &g;
    FilterCode &code;
    Filter *filter;
    Stmt *&first;

    // -------------------------------------------------- closure application
    // ORIG_VAL(x, y, closure-of-mathmap-filter, t)  ->  FILTER(args..., x, y, t)
    void closure_application() {
        walk(first, [&](Stmt *s) {
            if (!is_assign_with_op(s, OP_ORIG_VAL) || s->rhs->args[2].is_const) return;
            Stmt *def = s->rhs->args[2].value->def;
            if (!def || def->kind != ST_ASSIGN || def->rhs->kind != RHS_CLOSURE || def->rhs->filter->kind != FILTER_MATHMAP) return;
            std::vector<P> args = def->rhs->args;
            args.push_back(s->rhs->args[0]);
            args.push_back(s->rhs->args[1]);
            args.push_back(s->rhs->args[3]);
            Rhs *nw = code.new_rhs(RHS_FILTER);
            nw->filter = def->rhs->filter;
            nw->args = args;
            nw->history = def->rhs->history;
            replace_rhs(&s->rhs, nw, s);
        });
    }

    // ------------------------------------------------------------- inlining
    static bool can_inline(Filter *f, const std::shared_ptr<InlineHistory> &h) {
        if (f->kind != FILTER_MATHMAP) return false;
        for (auto p = h; p; p = p->next)
            if (p->filter == f) return false;
        for (auto &u : f->uservals)
            if (u.type == UV_COLOR) return false;
        return true;
    }
    bool inline_list(Stmt **loc) {
        bool changed = false;
        while (*loc) {
            Stmt *s = *loc;
            if (s->kind == ST_ASSIGN && s->rhs->kind == RHS_FILTER && can_inline(s->rhs->filter, s->rhs->history)) {
                Rhs *tuple_rhs = nullptr;
                std::vector<P> args = s->rhs->args;
                Stmt *stmts = g.gen_filter_code(s->rhs->filter, nullptr, &args, &tuple_rhs, s->rhs->history);
                replace_rhs(&s->rhs, tuple_rhs, s);
                if (stmts) {
                    for (Stmt *it = stmts; it; it = it->next) it->parent = s->parent;
                    last_stmt(stmts)->next = s;
                    *loc = stmts;
                }
                changed = true;
            } else if (s->kind == ST_IF) {
                changed |= inline_list(&s->cons);
                changed |= inline_list(&s->alt);
            } else if (s->kind == ST_WHILE)
                changed |= inline_list(&s->body);
            // advance to the statement after s (s may have moved behind inserted code)
            loc = &s->next;
        }
        return changed;
    }

    // ----------------------------------------------------- copy propagation
    bool copy_propagation() {
        bool changed = false;
        std::unordered_map<Value *, P> copies;
        auto rewrite_rhs = [&](Rhs *r, Stmt *s) {
            std::vector<Value *> vals;
            for_each_value_in_rhs(r, [&](Value *v) { vals.push_back(v); });
            for (Value *v : vals) {
                auto it = copies.find(v);
                if (it != copies.end()) {
                    rewrite_use(s, v, it->second);
                    changed = true;
                }
            }
        };
        walk(first, [&](Stmt *s) {
            switch (s->kind) {
            case ST_PHI:
                rewrite_rhs(s->rhs, s);
                rewrite_rhs(s->rhs2, s);
                break;
            case ST_ASSIGN:
                rewrite_rhs(s->rhs, s);
                if (s->rhs->kind == RHS_PRIMARY) copies[s->lhs] = s->rhs->prim;
                break;
            case ST_IF:
            case ST_WHILE: rewrite_rhs(s->cond, s); break;
            default: break;
            }
        });
        return changed;
    }

    // --------------------------------------------------------------- tuples
    bool tuple_nth() {
        bool changed = false;
        walk(first, [&](Stmt *s) {
            if (!is_assign_with_op(s, OP_TUPLE_NTH) || s->rhs->args[0].is_const) return;
            Stmt *def = s->rhs->args[0].value->def;
            if (!def || def->kind != ST_ASSIGN || def->rhs->kind != RHS_TUPLE) return;
            assert(s->rhs->args[1].is_const && s->rhs->args[1].c.type == T_INT);
            int n = s->rhs->args[1].c.i;
            assert(n < (int)def->rhs->args.size());
            replace_rhs(&s->rhs, g.rhs_prim(def->rhs->args[n]), s);
            changed = true;
        });
        return changed;
    }
    bool make_tuple() {
        bool changed = false;
        walk(first, [&](Stmt *s) {
            if (s->kind != ST_ASSIGN || s->rhs->kind != RHS_TUPLE) return;
            Value *tuple = nullptr;
            size_t i = 0;
            for (; i < s->rhs->args.size(); ++i) {
                P &a = s->rhs->args[i];
                if (a.is_const) break;
                Stmt *def = a.value->def;
                if (!is_assign_with_op(def, OP_TUPLE_NTH)) break;
                if (def->rhs->args[1].c.i != (int)i || def->rhs->args[0].is_const) break;
                Value *t = def->rhs->args[0].value;
                if (!tuple) tuple = t;
                else if (tuple != t) break;
            }
            if (i == s->rhs->args.size() && tuple) {
                // only when the source tuple has exactly this length
                Stmt *tdef = tuple->def;
                int tl = tdef && (tdef->kind == ST_ASSIGN) ? tuple_length_of_rhs(tdef->rhs) : 0;
                if (tl != 0 && tl != (int)s->rhs->args.size()) return;
                replace_rhs(&s->rhs, g.rhs_prim(P::of(tuple)), s);
                changed = true;
            }
        });
        return changed;
    }

    // ------------------------------------------------------------------ CSE
    static std::string rhs_key(const Rhs *r) {
        std::string k;
        switch (r->kind) {
        case RHS_INTERNAL: return "I:" + r->internal;
        case RHS_OP: k = std::string("O:") + r->op->name; break;
        case RHS_TUPLE: k = "T:"; break;
        default: return "";
        }
        for (auto &a : r->args) k += "," + primary_to_string(a);
        return k;
    }
    bool cse_list(Stmt *s, std::map<std::string, Value *> avail) {
        bool changed = false;
        for (; s; s = s->next) {
            if (s->kind == ST_ASSIGN) {
                Rhs *r = s->rhs;
                bool ok = (r->kind == RHS_OP && r->op->pure) || r->kind == RHS_INTERNAL;
                if (!ok) continue;
                std::string k = rhs_key(r);
                auto it = avail.find(k);
                if (it != avail.end()) {
                    replace_rhs(&s->rhs, g.rhs_prim(P::of(it->second)), s);
                    changed = true;
                } else
                    avail[k] = s->lhs;
            } else if (s->kind == ST_IF) {
                changed |= cse_list(s->cons, avail);
                changed |= cse_list(s->alt, avail);
            } else if (s->kind == ST_WHILE)
                changed |= cse_list(s->body, avail);
        }
        return changed;
    }

    // ------------------------------------------------ loop-carried values
    // irgen lays `while c do b end` out rotated, as the reference does (compiler.c:2119-2137): the condition's statements
    // once in front of the loop and again at the end of the body, for the next iteration's test.  An escape-time loop
    // (`while abs(c) < 2 ... c = c*c + p`) then computes the squares of the NEW c for the test at the end of iteration k-1
    // and the squares of the SAME numbers, now read through the loop's phis, for the product at the top of iteration k.
    // For a pure top-level body statement S = op(args) whose arguments are phis of this loop, constants and values defined
    // outside it: when op(entry values) is available in front of the loop (E) and op(back-edge values) is a top-level body
    // statement (B, executed in every iteration), then S in iteration k is B of iteration k-1 -- the same pure op on the
    // same operand values -- and E in the first one.  A new phi(E, B) carries the value and S becomes a copy of it.
    // Bit-exact by construction (nothing is re-associated, only not computed twice); Render/Mandelbrot.mm loses 4 of its
    // 14 products per iteration (27.16 -> 24.21 ms at 16384 x 16384 on a B200, profiles/r02_loop_carry_ab.json).
    bool loop_carry_enabled = loop_carry_default();
    static bool loop_carry_default() {
        const char *e = getenv("MMB_LOOP_CARRY");  // 0 switches the pass off (what a comparison against the plain IR needs)
        return e ? atoi(e) != 0 : true;
    }
    static bool defined_inside(const Value *v, const Stmt *loop) {
        for (const Stmt *s = v->def; s; s = s->parent)
            if (s == loop) return true;
        return false;
    }
    // the back-edge operand of a loop phi is a copy (`x.2 = tmp`) that copy propagation, visiting the phi first, leaves alone
    static P through_copies(P p) {
        for (int n = 0; n < 64 && !p.is_const && p.value->def && p.value->def->kind == ST_ASSIGN && p.value->def->rhs->kind == RHS_PRIMARY; ++n)
            p = p.value->def->rhs->prim;
        return p;
    }
    // scalar arithmetic only: conversions, + - * / %, the math functions, comparisons, and the complex functions -- whose
    // operands and results are ints, floats or complex numbers (no tuples, tree vectors, images or colours)
    static bool scalar_result(const Rhs *r) {
        if (r->kind != RHS_OP || !r->op->pure) return false;
        const int id = r->op->id;
        return (id >= OP_INT2FLOAT && id <= OP_NOT) || (id >= OP_COMPLEX && id <= OP_C_GAMMA);
    }
    bool loop_carried_values(Stmt *s, std::map<std::string, Value *> avail) {
        bool changed = false;
        for (; s; s = s->next) {
            if (s->kind == ST_ASSIGN) {
                if (s->rhs->kind == RHS_OP && s->rhs->op->pure) avail.emplace(rhs_key(s->rhs), s->lhs);
            } else if (s->kind == ST_IF) {
                changed |= loop_carried_values(s->cons, avail);
                changed |= loop_carried_values(s->alt, avail);
            } else if (s->kind == ST_WHILE) {
                std::unordered_map<const Value *, std::pair<P, P>> phis;  // phi -> (value on entry, value on the back edge)
                for (Stmt *p = s->entry; p; p = p->next)
                    if (p->kind == ST_PHI && p->rhs->kind == RHS_PRIMARY && p->rhs2->kind == RHS_PRIMARY)
                        phis[p->lhs] = {through_copies(p->rhs->prim), through_copies(p->rhs2->prim)};
                std::map<std::string, Value *> in_body;
                for (Stmt *b = s->body; b; b = b->next)
                    if (b->kind == ST_ASSIGN && scalar_result(b->rhs)) in_body.emplace(rhs_key(b->rhs), b->lhs);
                for (Stmt *c = s->body; c; c = c->next) {
                    if (c->kind != ST_ASSIGN || !scalar_result(c->rhs)) continue;
                    std::string on_entry = std::string("O:") + c->rhs->op->name, on_back = on_entry;
                    bool reads_phi = false, ok = true;
                    for (auto &a : c->rhs->args) {
                        auto it = a.is_const ? phis.end() : phis.find(a.value);
                        if (it != phis.end()) {
                            reads_phi = true;
                            on_entry += "," + primary_to_string(it->second.first);
                            on_back += "," + primary_to_string(it->second.second);
                        } else if (!a.is_const && defined_inside(a.value, s)) {
                            ok = false;
                            break;
                        } else {
                            on_entry += "," + primary_to_string(a);
                            on_back += "," + primary_to_string(a);
                        }
                    }
                    if (!ok || !reads_phi) continue;
                    // S may only be read inside the loop: behind it, S is the last EXECUTED iteration's value while the
                    // carrying phi already holds the next one's
                    for (const Stmt *u : c->lhs->uses) {
                        const Stmt *up = u;
                        while (up && up != s) up = up->parent;
                        if (!up) { ok = false; break; }
                    }
                    if (!ok) continue;
                    auto e = avail.find(on_entry);
                    auto b = in_body.find(on_back);
                    if (e == avail.end() || b == in_body.end() || b->second == c->lhs) continue;
                    int top_id = 0;
                    for (auto &existing : code.compvars) top_id = std::max(top_id, existing.id);
                    CompVar *cv = g.temp(T_INT);  // propagate_types gives it the operands' type
                    cv->id = top_id + 1;          // loaded IR keeps the ids of its text: stay clear of them
                    Value *carried = code.new_value(cv);
                    carried->index = 1;
                    Stmt *phi = code.new_stmt(ST_PHI);
                    phi->lhs = carried;
                    carried->def = phi;
                    phi->rhs = g.rhs_prim(P::of(e->second));
                    phi->rhs2 = g.rhs_prim(P::of(b->second));
                    add_use(e->second, phi);
                    add_use(b->second, phi);
                    phi->parent = s;
                    phi->next = s->entry;
                    s->entry = phi;
                    replace_rhs(&c->rhs, g.rhs_prim(P::of(carried)), c);
                    changed = true;
                }
                changed |= loop_carried_values(s->body, avail);
            }
        }
        return changed;
    }

    // ------------------------------------------------- folding / simplifying
    bool fold_rhs(Rhs **slot, Stmt *s) {
        Rhs *r = *slot;
        if (!r || r->kind != RHS_OP || !r->op->foldable) return false;
        Const args[6];
        for (int i = 0; i < r->op->nargs; ++i) {
            if (!r->args[i].is_const) return false;
            args[i] = r->args[i].c;
        }
        Const out;
        if (!eval_op(r->op, args, &out)) return false;
        P p;
        p.is_const = true;
        p.c = out;
        replace_rhs(slot, g.rhs_prim(p), s);
        return true;
    }
    bool constant_folding() {
        bool changed = false;
        walk(first, [&](Stmt *s) {
            switch (s->kind) {
            case ST_PHI: changed |= fold_rhs(&s->rhs2, s);  // fallthrough
            case ST_ASSIGN: changed |= fold_rhs(&s->rhs, s); break;
            case ST_IF:
            case ST_WHILE: changed |= fold_rhs(&s->cond, s); break;
            default: break;
            }
        });
        return changed;
    }
    // compiler.c:3462-3541: x+0, 0+x, x-0, x*1, 1*x, x*0, 0*x, x/1, x^1, x^0
    bool simplify_rhs(Rhs **slot, Stmt *s) {
        Rhs *r = *slot;
        if (!r || r->kind != RHS_OP) return false;
        auto cval = [&](int i) { return const_as_float(r->args[i].c); };
        auto unit = [&](float u, bool left, bool right) {
            if (left && r->args[0].is_const && cval(0) == u) { replace_rhs(slot, g.rhs_prim(r->args[1]), s); return true; }
            if (right && r->args[1].is_const && cval(1) == u) { replace_rhs(slot, g.rhs_prim(r->args[0]), s); return true; }
            return false;
        };
        auto zero = [&](float z, int result, bool left, bool right) {
            if ((left && r->args[0].is_const && cval(0) == z) || (right && r->args[1].is_const && cval(1) == z)) {
                // int result only when the first argument is an int constant (compiler.c:3490-3493)
                P p = (r->args[0].is_const && r->args[0].c.type == T_INT) ? P::ic(result) : P::fc((float)result);
                replace_rhs(slot, g.rhs_prim(p), s);
                return true;
            }
            return false;
        };
        switch (r->op->id) {
        case OP_ADD: return unit(0.f, true, true);
        case OP_SUB: return unit(0.f, false, true);
        case OP_MUL: return unit(1.f, true, true) || zero(0.f, 0, true, true);
        case OP_DIV: return unit(1.f, false, true);
        case OP_POW: return unit(1.f, false, true) || zero(0.f, 1, false, true);
        default: return false;
        }
    }
    bool simplify_ops() {
        bool changed = false;
        walk(first, [&](Stmt *s) {
            switch (s->kind) {
            case ST_PHI: changed |= simplify_rhs(&s->rhs2, s);  // fallthrough
            case ST_ASSIGN: changed |= simplify_rhs(&s->rhs, s); break;
            case ST_IF:
            case ST_WHILE: changed |= simplify_rhs(&s->cond, s); break;
            default: break;
            }
        });
        return changed;
    }

    // ------------------------------------------------------- resize lowering
    // compopt/resize.c:30-101: ORIG_VAL(x, y, RESIZE_IMAGE(img, xf, yf), t) -> ORIG_VAL(x*xf, y*yf, img, t)
    bool orig_val_resize(Stmt **loc) {
        bool changed = false;
        while (*loc) {
            Stmt *s = *loc;
            if (is_assign_with_op(s, OP_ORIG_VAL) && !s->rhs->args[2].is_const) {
                Stmt *def = s->rhs->args[2].value->def;
                if (is_assign_with_op(def, OP_RESIZE_IMAGE)) {
                    P ox = s->rhs->args[0], oy = s->rhs->args[1];
                    P image = def->rhs->args[0], xf = def->rhs->args[1], yf = def->rhs->args[2];
                    CompVar *nx = g.temp(T_INT), *ny = g.temp(T_INT);
                    loc = g.emit_before(g.make_assign_stmt(nx, g.rhs_op(OP_MUL, {ox, xf})), loc, s->parent);
                    loc = g.emit_before(g.make_assign_stmt(ny, g.rhs_op(OP_MUL, {oy, yf})), loc, s->parent);
                    assert(*loc == s);
                    Rhs *nw = g.rhs_op(OP_ORIG_VAL, {g.cur(nx), g.cur(ny), image, s->rhs->args[3]});
                    replace_rhs(&s->rhs, nw, s);
                    changed = true;
                }
            } else if (s->kind == ST_IF) {
                changed |= orig_val_resize(&s->cons);
                changed |= orig_val_resize(&s->alt);
            } else if (s->kind == ST_WHILE)
                changed |= orig_val_resize(&s->body);
            loc = &s->next;
        }
        return changed;
    }
    // compopt/resize.c:109-169
    bool strip_resize() {
        bool changed = false;
        walk(first, [&](Stmt *s) {
            if (!is_assign_with_op(s, OP_STRIP_RESIZE) || s->rhs->args[0].is_const) return;
            Stmt *def = s->rhs->args[0].value->def;
            if (!def) return;
            if (is_assign_with_op(def, OP_RESIZE_IMAGE)) {
                replace_rhs(&s->rhs, g.rhs_prim(def->rhs->args[0]), s);
                changed = true;
            } else if (def->kind == ST_ASSIGN && (def->rhs->kind == RHS_CLOSURE || is_assign_with_op(def, OP_USERVAL_IMAGE))) {
                replace_rhs(&s->rhs, g.rhs_prim(P::of(def->lhs)), s);
                changed = true;
            }
        });
        return changed;
    }
    // compopt/simplify.c:28-62 + simplify.lisp:33-39
    bool closure_pixel_size() {
        bool changed = false;
        walk(first, [&](Stmt *s) {
            if (!(is_assign_with_op(s, OP_IMAGE_PIXEL_WIDTH) || is_assign_with_op(s, OP_IMAGE_PIXEL_HEIGHT))) return;
            if (s->rhs->args[0].is_const) return;
            Stmt *def = s->rhs->args[0].value->def;
            if (!def || def->kind != ST_ASSIGN || def->rhs->kind != RHS_CLOSURE) return;
            Filter *cf = def->rhs->filter;
            for (size_t i = 0; i < cf->uservals.size(); ++i)
                if (cf->uservals[i].type == UV_IMAGE) {
                    Rhs *nw = g.rhs_op(s->rhs->op->id, {def->rhs->args[i]});
                    replace_rhs(&s->rhs, nw, s);
                    changed = true;
                    return;
                }
            bool w = s->rhs->op->id == OP_IMAGE_PIXEL_WIDTH;
            replace_rhs(&s->rhs, g.rhs_internal(w ? "__canvasPixelW" : "__canvasPixelH"), s);
            changed = true;
        });
        return changed;
    }

    // ------------------------------------------------------------ dead code
    bool dead_assignments() {
        std::vector<Value *> work;
        std::unordered_set<Value *> live;
        walk(first, [&](Stmt *s) {
            switch (s->kind) {
            case ST_PHI:
                if (!rhs_is_pure(s->rhs2) || !rhs_is_pure(s->rhs)) work.push_back(s->lhs);
                break;
            case ST_ASSIGN:
                if (!rhs_is_pure(s->rhs)) work.push_back(s->lhs);
                break;
            case ST_IF:
            case ST_WHILE: for_each_value_in_rhs(s->cond, [&](Value *v) { work.push_back(v); }); break;
            default: break;
            }
        });
        while (!work.empty()) {
            Value *v = work.back();
            work.pop_back();
            if (!live.insert(v).second) continue;
            Stmt *d = v->def;
            if (!d) continue;
            if (d->kind == ST_PHI) for_each_value_in_rhs(d->rhs2, [&](Value *x) { work.push_back(x); });
            if (d->kind == ST_PHI || d->kind == ST_ASSIGN) for_each_value_in_rhs(d->rhs, [&](Value *x) { work.push_back(x); });
        }
        bool changed = false;
        walk(first, [&](Stmt *s) {
            if ((s->kind == ST_ASSIGN || s->kind == ST_PHI) && !live.count(s->lhs)) {
                remove_uses_in_rhs(s->rhs, s);
                if (s->kind == ST_PHI) remove_uses_in_rhs(s->rhs2, s);
                s->kind = ST_NIL;
                changed = true;
            }
        });
        return changed;
    }
    // compiler.c:3640-3735
    bool dead_branches(Stmt *s) {
        bool changed = false;
        for (; s; s = s->next) {
            if (s->kind == ST_IF) {
                if (s->cond->kind == RHS_PRIMARY && s->cond->prim.is_const) {
                    bool truth = const_is_true(s->cond->prim.c);
                    Stmt *branch = truth ? s->cons : s->alt;
                    Stmt *insertion = s;
                    while (branch) {
                        Stmt *next = branch->next;
                        if (branch->kind != ST_NIL) {
                            branch->parent = s->parent;
                            branch->next = insertion->next;
                            insertion->next = branch;
                            insertion = branch;
                        }
                        branch = next;
                    }
                    // the branch not taken disappears: drop the uses it holds
                    walk(truth ? s->alt : s->cons, [&](Stmt *d) {
                        if (d->kind == ST_ASSIGN) remove_uses_in_rhs(d->rhs, d);
                        else if (d->kind == ST_PHI) { remove_uses_in_rhs(d->rhs, d); remove_uses_in_rhs(d->rhs2, d); }
                        else if (d->kind == ST_IF || d->kind == ST_WHILE) remove_uses_in_rhs(d->cond, d);
                        d->kind = ST_NIL;
                    });
                    Stmt *phi = s->exit;
                    while (phi) {
                        Stmt *next = phi->next;
                        if (phi->kind == ST_PHI) {
                            remove_uses_in_rhs(truth ? phi->rhs2 : phi->rhs, phi);
                            phi->kind = ST_ASSIGN;
                            if (!truth) phi->rhs = phi->rhs2;
                            phi->rhs2 = nullptr;
                            phi->parent = s->parent;
                            phi->next = insertion->next;
                            insertion->next = phi;
                            insertion = phi;
                        }
                        phi = next;
                    }
                    s->kind = ST_NIL;
                    s->cons = s->alt = s->exit = nullptr;
                    s->cond = nullptr;
                    changed = true;
                } else {
                    changed |= dead_branches(s->cons);
                    changed |= dead_branches(s->alt);
                }
            } else if (s->kind == ST_WHILE)
                changed |= dead_branches(s->body);
        }
        return changed;
    }
    static bool stmts_empty(Stmt *s) {
        while (s && s->kind == ST_NIL) s = s->next;
        return s == nullptr;
    }
    bool dead_controls(Stmt *s) {
        bool changed = false;
        for (; s; s = s->next) {
            if (s->kind == ST_IF) {
                if (stmts_empty(s->cons) && stmts_empty(s->alt) && stmts_empty(s->exit) && rhs_is_pure(s->cond)) {
                    remove_uses_in_rhs(s->cond, s);
                    s->kind = ST_NIL;
                    changed = true;
                } else {
                    changed |= dead_controls(s->cons);
                    changed |= dead_controls(s->alt);
                }
            } else if (s->kind == ST_WHILE)
                changed |= dead_controls(s->body);
        }
        return changed;
    }

    // re-link parents after structural edits (cheap and keeps invariants simple)
    static void fix_parents(Stmt *s, Stmt *parent) {
        for (; s; s = s->next) {
            s->parent = parent;
            if (s->kind == ST_IF) { fix_parents(s->cons, s); fix_parents(s->alt, s); fix_parents(s->exit, s); }
            else if (s->kind == ST_WHILE) { fix_parents(s->entry, s); fix_parents(s->body, s); }
        }
    }

    void run() {
        bool changed = true;
        int iter = 0;
        while (changed && iter++ < 200) {
            closure_application();
            changed = false;
            changed |= inline_list(&first);
            changed |= copy_propagation();
            changed |= tuple_nth();
            changed |= make_tuple();
            changed |= cse_list(first, {});
            changed |= copy_propagation();
            if (loop_carry_enabled) {
                fix_parents(first, nullptr);
                changed |= loop_carried_values(first, {});
                changed |= copy_propagation();
            }
            changed |= constant_folding();
            changed |= simplify_ops();
            changed |= orig_val_resize(&first);
            changed |= strip_resize();
            changed |= closure_pixel_size();
            changed |= dead_assignments();
            changed |= dead_branches(first);
            changed |= dead_controls(first);
            fix_parents(first, nullptr);
        }
    }
};

// Removes ST_NIL statements so emitters see a clean list.
Stmt *strip_nils(Stmt *s) {
    Stmt *head = nullptr, **tail = &head;
    for (; s; s = s->next) {
        if (s->kind == ST_NIL) continue;
        if (s->kind == ST_IF) { s->cons = strip_nils(s->cons); s->alt = strip_nils(s->alt); s->exit = strip_nils(s->exit); }
        if (s->kind == ST_WHILE) { s->entry = strip_nils(s->entry); s->body = strip_nils(s->body); }
        *tail = s;
        tail = &s->next;
    }
    *tail = nullptr;
    return head;
}

}  // namespace

// The loop-carried value pass alone, for IR that did not come through this front end (mmb_load_ir: the reference's own
// compiler hands its optimised IR over the boundary and must end up with the same kernels).  A fixpoint on IR that has
// already been through it.
void carry_loop_values(Module &mod, FilterCode &code) {
    Gen g(mod, code);
    Optimizer opt{mod, g, code, code.filter, code.first};
    if (!opt.loop_carry_enabled) return;
    Optimizer::fix_parents(code.first, nullptr);
    bool changed = false;
    for (int iter = 0; iter < 16 && opt.loop_carried_values(code.first, {}); ++iter) {
        changed = true;
        opt.copy_propagation();
        opt.dead_assignments();
    }
    if (!changed) return;
    code.first = strip_nils(code.first);
    Optimizer::fix_parents(code.first, nullptr);
}

// -------------------------------------------------------------------- types
void propagate_types(FilterCode &code) {
    bool changed = true;
    while (changed) {
        changed = false;
        walk(code.first, [&](Stmt *s) {
            if (s->kind != ST_ASSIGN && s->kind != ST_PHI) return;
            Type t = rhs_type(s->rhs);
            if (s->kind == ST_PHI) {
                Type t2 = rhs_type(s->rhs2);
                if (t2 > t) t = t2;
            }
            if (t > s->lhs->cv->type) {
                s->lhs->cv->type = t;
                changed = true;
            }
        });
    }
    // tuple lengths (our addition: device code keeps tuples in registers and
    // needs a static length per tuple compvar)
    changed = true;
    while (changed) {
        changed = false;
        walk(code.first, [&](Stmt *s) {
            if (s->kind != ST_ASSIGN && s->kind != ST_PHI) return;
            if (s->lhs->cv->type != T_TUPLE && s->lhs->cv->type != T_TREE_VECTOR) return;
            auto len_of = [&](Rhs *r) {
                if (!r) return 0;
                if (r->kind == RHS_PRIMARY) return r->prim.is_const ? 0 : r->prim.value->cv->tuple_len;
                if (r->kind == RHS_OP && r->op->id == OP_SET_TREE_VECTOR_NTH) return r->args[1].is_const ? 0 : r->args[1].value->cv->tuple_len;
                return tuple_length_of_rhs(r);
            };
            int l = std::max(len_of(s->rhs), s->kind == ST_PHI ? len_of(s->rhs2) : 0);
            if (l > s->lhs->cv->tuple_len) {
                s->lhs->cv->tuple_len = l;
                changed = true;
            }
        });
    }
}

// ---------------------------------------------------------------- constness
// A value's const bits say which of x, y, t it does NOT depend on.  An op
// result is the AND of its arguments'; values assigned under control flow are
// additionally limited by the controlling condition; phis by both inputs and
// the condition.  Values with CONST_X|CONST_Y whose defining op the host can
// evaluate are "hoisted": computed once per frame by the host replay
// (backend/host_eval.cpp), like the reference's init_frame slice
// (new_template.c.in:314-337, compiler.c:2867-3234).
static bool host_can_eval(const Rhs *r) {
    switch (r->kind) {
    case RHS_PRIMARY:
    case RHS_INTERNAL: return true;
    case RHS_TUPLE: return true;
    case RHS_CLOSURE: return true;  // closure construction / native filter call
    case RHS_FILTER: return false;
    case RHS_OP:
        if (!r->op->pure) return false;
        switch (r->op->id) {
        case OP_ORIG_VAL: case OP_APPLY_CURVE: case OP_APPLY_GRADIENT:
        case OP_GAMMA: case OP_BETA: case OP_C_GAMMA:
        case OP_SOLVE_LINEAR_2: case OP_SOLVE_LINEAR_3: case OP_SOLVE_POLY_2: case OP_SOLVE_POLY_3:
        case OP_LIBNOISE_PERLIN: case OP_LIBNOISE_BILLOW: case OP_LIBNOISE_RIDGED_MULTI: case OP_LIBNOISE_VORONOI:
        case OP_TREE_VECTOR_NTH: case OP_SET_TREE_VECTOR_NTH:
            return false;
        default: return true;
        }
    default: return false;
    }
}

void analyze_constants(FilterCode &code) {
    Filter *f = code.filter;
    for (auto &v : code.values) { v.const_bits = CONST_ALL; v.hoisted = false; }
    auto prim_bits = [&](const P &p) { return p.is_const ? (int)CONST_ALL : (p.value->index < 0 ? (int)CONST_ALL : p.value->const_bits); };
    auto rhs_bits = [&](const Rhs *r) {
        int b = CONST_ALL;
        switch (r->kind) {
        case RHS_INTERNAL: {
            Internal *in = f->lookup_internal(r->internal, false);
            b = in ? in->const_bits : CONST_NONE;
            break;
        }
        case RHS_PRIMARY: b = prim_bits(r->prim); break;
        case RHS_OP:
            if (!r->op->pure) return (int)CONST_NONE;
            // fallthrough
        default:
            for (auto &a : r->args) b &= prim_bits(a);
            if (r->kind == RHS_FILTER) b = CONST_NONE;
            break;
        }
        return b;
    };
    bool changed = true;
    std::function<void(Stmt *, int)> go = [&](Stmt *s, int ctrl) {
        for (; s; s = s->next) {
            switch (s->kind) {
            case ST_ASSIGN: {
                (void)ctrl;  // pure definitions may be speculated; only phis depend on control
                int nb = s->lhs->const_bits & rhs_bits(s->rhs);
                if (nb != s->lhs->const_bits) { s->lhs->const_bits = nb; changed = true; }
                break;
            }
            case ST_IF: {
                int cb = rhs_bits(s->cond) & ctrl;
                go(s->cons, cb);
                go(s->alt, cb);
                // A phi of an `if` is a pure function of the condition and its two inputs: the control AROUND the `if` decides
                // whether the value is needed, not what it is.  (The reference limits it by the enclosing conditions too,
                // compiler.c analyze_phis_constants' const_max, and then builds e.g. the closure of an inlined image
                // argument per pixel inside a loop; here such a value is a frame constant, computed by the host replay,
                // which already speculates pure definitions under per-pixel control.)
                const int own = rhs_bits(s->cond);
                for (Stmt *p = s->exit; p; p = p->next) {
                    if (p->kind != ST_PHI) continue;
                    int b = rhs_bits(p->rhs) & rhs_bits(p->rhs2) & own;
                    if ((p->lhs->const_bits & b) != p->lhs->const_bits) { p->lhs->const_bits &= b; changed = true; }
                }
                break;
            }
            case ST_WHILE: {
                // loop-carried: iterate locally until stable
                for (int it = 0; it < 64; ++it) {
                    bool before = changed;
                    changed = false;
                    int cb = rhs_bits(s->cond) & ctrl;
                    for (Stmt *p = s->entry; p; p = p->next) {
                        if (p->kind != ST_PHI) continue;
                        int b = rhs_bits(p->rhs) & rhs_bits(p->rhs2) & cb;
                        if ((p->lhs->const_bits & b) != p->lhs->const_bits) { p->lhs->const_bits &= b; changed = true; }
                    }
                    cb = rhs_bits(s->cond) & ctrl;
                    go(s->body, cb);
                    bool again = changed;
                    changed = before || again;
                    if (!again) break;
                }
                break;
            }
            default: break;
            }
        }
    };
    for (int it = 0; it < 64 && changed; ++it) {
        changed = false;
        go(code.first, CONST_ALL);
    }

    // ---- evaluation levels ------------------------------------------------
    // 0: once per frame, replayed on the host (reference: init_frame / xy_vars)
    // 1: once per output row (reference: the "x-const" row locals)
    // 3: every pixel
    // A definition's level is the max of its own base level (from const bits and
    // whether the host can evaluate it) and its arguments' levels.  Pure
    // definitions may sit at a lower level than the control structure around
    // them (they are speculated); phis are bound to their controlling condition;
    // a loop runs at the max level of everything inside it.
    for (auto &v : code.values) v.level = 0;
    auto prim_level = [&](const P &p) { return (p.is_const || p.value->index < 0) ? 0 : p.value->level; };
    auto rhs_level = [&](const Rhs *r) {
        int l = 0;
        if (r->kind == RHS_PRIMARY) l = prim_level(r->prim);
        else if (r->kind != RHS_INTERNAL)
            for (auto &a : r->args) l = std::max(l, prim_level(a));
        return l;
    };
    auto base_level = [&](const Value *v, const Rhs *r) {
        int l = 3;
        if ((v->const_bits & (CONST_X | CONST_Y)) == (CONST_X | CONST_Y)) l = 0;
        else if (v->const_bits & CONST_X) l = 1;
        if (l == 0 && !host_can_eval(r)) l = 1;
        if (r->kind == RHS_OP && !r->op->pure) l = 3;
        if (r->kind == RHS_FILTER) l = 3;
        return l;
    };
    bool ch = true;
    auto raise = [&](Value *v, int l) { if (l > v->level) { v->level = l; ch = true; } };
    std::function<int(Stmt *)> lv = [&](Stmt *s) {
        int mx = 0;  // max level of anything in this list
        for (; s; s = s->next) {
            switch (s->kind) {
            case ST_ASSIGN:
                raise(s->lhs, std::max(base_level(s->lhs, s->rhs), rhs_level(s->rhs)));
                mx = std::max(mx, s->lhs->level);
                break;
            case ST_IF: {
                int cl = rhs_level(s->cond);
                if (!host_can_eval(s->cond)) cl = std::max(cl, 1);
                s->level = cl;
                mx = std::max(mx, std::max(lv(s->cons), lv(s->alt)));
                for (Stmt *p = s->exit; p; p = p->next) {
                    if (p->kind != ST_PHI) continue;
                    raise(p->lhs, std::max(cl, std::max(rhs_level(p->rhs), rhs_level(p->rhs2))));
                    mx = std::max(mx, p->lhs->level);
                }
                break;
            }
            case ST_WHILE: {
                int ll = std::max(s->level, rhs_level(s->cond));
                if (!host_can_eval(s->cond)) ll = std::max(ll, 1);
                for (Stmt *p = s->entry; p; p = p->next)
                    if (p->kind == ST_PHI) ll = std::max(ll, std::max(p->lhs->level, std::max(rhs_level(p->rhs), rhs_level(p->rhs2))));
                ll = std::max(ll, lv(s->body));
                for (Stmt *p = s->entry; p; p = p->next)
                    if (p->kind == ST_PHI) raise(p->lhs, ll);
                if (ll != s->level) { s->level = ll; ch = true; }
                mx = std::max(mx, ll);
                break;
            }
            default: break;
            }
        }
        return mx;
    };
    std::function<void(Stmt *)> reset = [&](Stmt *s) {
        for (; s; s = s->next) {
            s->level = 0;
            if (s->kind == ST_IF) { reset(s->cons); reset(s->alt); }
            if (s->kind == ST_WHILE) reset(s->body);
        }
    };
    reset(code.first);
    while (ch) {
        ch = false;
        lv(code.first);
    }
    for (auto &v : code.values) v.hoisted = v.level == 0;
}

std::unique_ptr<FilterCode> compile_filter(Module &mod, Filter *filter, bool optimize) {
    auto code = std::make_unique<FilterCode>();
    code->filter = filter;
    Gen g(mod, *code);
    CompVar *tuple_tmp = g.temp(T_TUPLE);
    code->first = g.gen_filter_code(filter, tuple_tmp, nullptr, nullptr, nullptr);
    g.filter = filter;
    Stmt *last = last_stmt(code->first);
    g.emit_loc = last ? &last->next : &code->first;
    CompVar *dummy = g.temp(T_INT);
    g.assign(dummy, g.rhs_op(OP_OUTPUT_TUPLE, {g.cur(tuple_tmp)}));
    g.emit_loc = nullptr;
    if (optimize) {
        Optimizer opt{mod, g, *code, filter, code->first};
        opt.run();
    }
    code->first = strip_nils(code->first);
    Optimizer::fix_parents(code->first, nullptr);
    propagate_types(*code);
    analyze_constants(*code);
    return code;
}

}  // namespace mm
