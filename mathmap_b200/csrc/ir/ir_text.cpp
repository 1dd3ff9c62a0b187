// Text form of the IR ("mmir 1").  This is the interchange format of the C ABI:
// a reference-side backends/cuda.c walks filter_code_t (reference
// compiler-internals.h:182-235) and prints this; dump_ir() prints the same form
// from our own front end, and the oracle's C emitter (oracle/emit_c.py) reads it.
//
//   (mmir 1
//    (filter NAME (flags unit square)
//     (uservals (int "name" MIN MAX DEFAULT) (image "in" FLAGS) ...)
//     (vars (ID TYPE [TUPLELEN]) ...)
//     (code STMT ...)))
//   STMT := (assign %cv.idx CONSTBITS LEVEL RHS)
//         | (if RHS LEVEL (STMT ...) (STMT ...) (phis (phi %lhs CONSTBITS LEVEL RHS RHS2) ...))
//         | (while (phis ...) RHS LEVEL (STMT ...))
//   LEVEL: 0 once per frame (host), 1 once per row, 3 per pixel
//   RHS  := PRIM | (internal NAME) | (op NAME PRIM ...) | (tuple PRIM ...)
//         | (closure FILTERNAME PRIM ...) | (filter FILTERNAME PRIM ...)
//   PRIM := %cv.idx | %cv.u (never assigned: reads 0) | i:INT | f:FLOAT | c:RE,IM | k:UINT
#include <set>
#include <sstream>

#include "ir.h"

namespace mm {

static void dump_rhs(std::ostringstream &o, const Rhs *r) {
    switch (r->kind) {
    case RHS_PRIMARY: o << primary_to_string(r->prim); return;
    case RHS_INTERNAL: o << "(internal " << r->internal << ")"; return;
    case RHS_OP: o << "(op " << r->op->name; break;
    case RHS_TUPLE: o << "(tuple"; break;
    case RHS_TREE_VECTOR: o << "(tree-vector"; break;
    case RHS_CLOSURE: o << "(closure " << r->filter->name; break;
    case RHS_FILTER: o << "(filter " << r->filter->name; break;
    }
    for (auto &a : r->args) o << " " << primary_to_string(a);
    o << ")";
}

static void dump_phis(std::ostringstream &o, const Stmt *p, int ind) {
    o << "(phis";
    for (; p; p = p->next) {
        if (p->kind != ST_PHI) continue;
        o << "\n" << std::string(ind + 1, ' ') << "(phi " << primary_to_string(Primary::of(p->lhs)) << " " << p->lhs->const_bits << " "
          << p->lhs->level << " ";
        dump_rhs(o, p->rhs);
        o << " ";
        dump_rhs(o, p->rhs2);
        o << ")";
    }
    o << ")";
}

static void dump_stmts(std::ostringstream &o, const Stmt *s, int ind) {
    std::string pad(ind, ' ');
    for (; s; s = s->next) {
        switch (s->kind) {
        case ST_NIL: break;
        case ST_ASSIGN:
            o << pad << "(assign " << primary_to_string(Primary::of(s->lhs)) << " " << s->lhs->const_bits << " " << s->lhs->level << " ";
            dump_rhs(o, s->rhs);
            o << ")\n";
            break;
        case ST_PHI: break;
        case ST_IF:
            o << pad << "(if ";
            dump_rhs(o, s->cond);
            o << " " << s->level << "\n" << pad << " (\n";
            dump_stmts(o, s->cons, ind + 2);
            o << pad << " )\n" << pad << " (\n";
            dump_stmts(o, s->alt, ind + 2);
            o << pad << " )\n" << pad << " ";
            dump_phis(o, s->exit, ind + 1);
            o << ")\n";
            break;
        case ST_WHILE:
            o << pad << "(while ";
            dump_phis(o, s->entry, ind + 1);
            o << "\n" << pad << " ";
            dump_rhs(o, s->cond);
            o << " " << s->level << "\n" << pad << " (\n";
            dump_stmts(o, s->body, ind + 2);
            o << pad << " ))\n";
            break;
        }
    }
}

static void collect_compvars(const Stmt *s, std::set<const CompVar *> &out) {
    auto rhs = [&](const Rhs *r) {
        if (r) for_each_value_in_rhs(const_cast<Rhs *>(r), [&](Value *v) { out.insert(v->cv); });
    };
    for (; s; s = s->next) {
        switch (s->kind) {
        case ST_PHI: rhs(s->rhs2);  // fallthrough
        case ST_ASSIGN: rhs(s->rhs); out.insert(s->lhs->cv); break;
        case ST_IF: rhs(s->cond); collect_compvars(s->cons, out); collect_compvars(s->alt, out); collect_compvars(s->exit, out); break;
        case ST_WHILE: rhs(s->cond); collect_compvars(s->entry, out); collect_compvars(s->body, out); break;
        default: break;
        }
    }
}

static std::string quote(const std::string &s) {
    std::string r = "\"";
    for (char c : s) {
        if (c == '"' || c == '\\') r += '\\';
        if (c == '\n') { r += "\\n"; continue; }
        r += c;
    }
    return r + "\"";
}

std::string dump_ir(const FilterCode &code) {
    std::ostringstream o;
    const Filter *f = code.filter;
    o << " (filter " << f->name << " (flags";
    if (f->flags & IMAGE_FLAG_UNIT) o << " unit";
    if (f->flags & IMAGE_FLAG_SQUARE) o << " square";
    o << ")\n  (uservals";
    for (auto &u : f->uservals) {
        o << "\n   (" << userval_type_name(u.type) << " " << quote(u.name);
        switch (u.type) {
        case UV_INT: o << " " << u.int_min << " " << u.int_max << " " << u.int_default; break;
        case UV_FLOAT: o << " " << format_float(u.float_min) << " " << format_float(u.float_max) << " " << format_float(u.float_default); break;
        case UV_BOOL: o << " " << u.bool_default; break;
        case UV_IMAGE: o << " " << u.image_flags; break;
        default: break;
        }
        o << ")";
    }
    o << ")\n  (vars";
    std::set<const CompVar *> live;
    collect_compvars(code.first, live);
    for (auto &cv : code.compvars) {
        if (!live.count(&cv)) continue;
        o << "\n   (" << cv.id << " " << type_name(cv.type);
        if (cv.type == T_TUPLE || cv.type == T_TREE_VECTOR) o << " " << cv.tuple_len;
        o << ")";
    }
    o << ")\n  (code\n";
    dump_stmts(o, code.first, 3);
    o << "  ))\n";
    return o.str();
}

std::string dump_module_ir(const std::vector<const FilterCode *> &codes, const std::string &main_name) {
    std::string s = "(mmir 1\n";
    for (const FilterCode *c : codes)
        if (c) s += dump_ir(*c);
    s += " (main " + main_name + "))\n";
    return s;
}

}  // namespace mm
