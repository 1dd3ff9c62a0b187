// Reader for the "mmir 1" text (see ir_text.cpp for the grammar).  This is the
// receiving end of the drop-in boundary: the reference's own compiler produces
// filter_code_t** (compiler.c:4798-4832), a backends/cuda.c stub prints it as
// text, and mmb_load_ir() rebuilds the same in-memory IR our front end produces.
// Levels/const bits in the text are ignored and recomputed here so that a
// producer need not implement our analysis.
#include <cstdlib>
#include <cstring>
#include <map>

#include "../cabi_module.h"
#include "passes.h"

namespace mmbackend {

using namespace mm;

namespace {

struct Sx {
    bool is_list = false;
    bool is_string = false;
    std::string atom;
    std::vector<Sx> items;
    const Sx &operator[](size_t i) const { return items.at(i); }
    size_t size() const { return items.size(); }
};

[[noreturn]] void bad(const std::string &msg) {
    CompileError e;
    e.message = "mmir: " + msg;
    throw e;
}

struct Reader {
    const std::string &s;
    size_t pos = 0;
    explicit Reader(const std::string &t) : s(t) {}
    void skip() {
        while (pos < s.size() && (isspace((unsigned char)s[pos]) || s[pos] == ';')) {
            if (s[pos] == ';') while (pos < s.size() && s[pos] != '\n') ++pos;
            else ++pos;
        }
    }
    Sx read() {
        skip();
        if (pos >= s.size()) bad("unexpected end of text");
        Sx x;
        if (s[pos] == '(') {
            ++pos;
            x.is_list = true;
            for (;;) {
                skip();
                if (pos >= s.size()) bad("unbalanced parenthesis");
                if (s[pos] == ')') { ++pos; break; }
                x.items.push_back(read());
            }
            return x;
        }
        if (s[pos] == '"') {
            ++pos;
            x.is_string = true;
            while (pos < s.size() && s[pos] != '"') {
                if (s[pos] == '\\' && pos + 1 < s.size()) ++pos;
                x.atom += s[pos++];
            }
            ++pos;
            return x;
        }
        while (pos < s.size() && !isspace((unsigned char)s[pos]) && s[pos] != '(' && s[pos] != ')') x.atom += s[pos++];
        return x;
    }
};

float parse_float(const std::string &t) {
    if (t == "nan") return NAN;
    if (t == "inf") return INFINITY;
    if (t == "-inf") return -INFINITY;
    return strtof(t.c_str(), nullptr);
}

struct Builder {
    Module &mod;
    FilterCode &code;
    std::map<int, CompVar *> cvs;
    std::map<std::pair<int, int>, Value *> vals;

    CompVar *cv(int id) {
        auto it = cvs.find(id);
        if (it == cvs.end()) bad("value of undeclared variable %" + std::to_string(id));
        return it->second;
    }
    Value *value(const std::string &tok) {
        size_t dot = tok.find('.');
        if (tok[0] != '%' || dot == std::string::npos) bad("bad value token " + tok);
        int id = atoi(tok.c_str() + 1);
        CompVar *c = cv(id);
        std::string idx = tok.substr(dot + 1);
        int index = idx == "u" ? -1 : atoi(idx.c_str());
        auto key = std::make_pair(id, index);
        auto it = vals.find(key);
        if (it != vals.end()) return it->second;
        Value *v = code.new_value(c);
        v->index = index;
        vals[key] = v;
        return v;
    }
    Primary prim(const Sx &x) {
        if (x.is_list) bad("expected a primary");
        const std::string &t = x.atom;
        if (t[0] == '%') return Primary::of(value(t));
        if (t.size() < 3 || t[1] != ':') bad("bad primary " + t);
        std::string v = t.substr(2);
        switch (t[0]) {
        case 'i': return Primary::ic(atoi(v.c_str()));
        case 'f': return Primary::fc(parse_float(v));
        case 'c': {
            size_t comma = v.find(',');
            return Primary::cc({parse_float(v.substr(0, comma)), parse_float(v.substr(comma + 1))});
        }
        case 'k': {
            Primary p;
            p.is_const = true;
            p.c.type = T_COLOR;
            p.c.color = (uint32_t)strtoul(v.c_str(), nullptr, 10);
            return p;
        }
        default: bad("bad primary " + t);
        }
    }
    Rhs *rhs(const Sx &x) {
        if (!x.is_list) {
            Rhs *r = code.new_rhs(RHS_PRIMARY);
            r->prim = prim(x);
            return r;
        }
        const std::string &head = x[0].atom;
        Rhs *r = nullptr;
        size_t first_arg = 1;
        if (head == "internal") {
            r = code.new_rhs(RHS_INTERNAL);
            r->internal = x[1].atom;
            return r;
        } else if (head == "op") {
            r = code.new_rhs(RHS_OP);
            r->op = op_by_name(x[1].atom);
            if (!r->op) bad("unknown op " + x[1].atom);
            first_arg = 2;
            if ((int)x.size() - 2 != r->op->nargs) bad("wrong number of arguments for op " + x[1].atom);
        } else if (head == "tuple") {
            r = code.new_rhs(RHS_TUPLE);
        } else if (head == "tree-vector") {
            r = code.new_rhs(RHS_TREE_VECTOR);
        } else if (head == "closure" || head == "filter") {
            r = code.new_rhs(head == "closure" ? RHS_CLOSURE : RHS_FILTER);
            r->filter = mod.lookup_filter(x[1].atom);
            if (!r->filter) bad("unknown filter " + x[1].atom);
            first_arg = 2;
        } else
            bad("unknown rhs form " + head);
        for (size_t i = first_arg; i < x.size(); ++i) r->args.push_back(prim(x[i]));
        return r;
    }
    void add_uses(Rhs *r, Stmt *s) { for_each_value_in_rhs(r, [&](Value *v) { add_use(v, s); }); }
    Stmt *phis(const Sx &x, Stmt *parent) {
        if (!x.is_list || x.size() < 1 || x[0].atom != "phis") bad("expected (phis ...)");
        Stmt *head = nullptr, **tail = &head;
        for (size_t i = 1; i < x.size(); ++i) {
            const Sx &p = x[i];
            if (p.size() != 6 || p[0].atom != "phi") bad("malformed phi");
            Stmt *s = code.new_stmt(ST_PHI);
            s->lhs = value(p[1].atom);
            s->lhs->def = s;
            s->rhs = rhs(p[4]);
            s->rhs2 = rhs(p[5]);
            add_uses(s->rhs, s);
            add_uses(s->rhs2, s);
            s->parent = parent;
            *tail = s;
            tail = &s->next;
        }
        return head;
    }
    Stmt *stmts(const Sx &list, size_t from, Stmt *parent) {
        Stmt *head = nullptr, **tail = &head;
        for (size_t i = from; i < list.size(); ++i) {
            const Sx &x = list[i];
            if (!x.is_list || x.size() < 1) bad("malformed statement");
            const std::string &k = x[0].atom;
            Stmt *s = nullptr;
            if (k == "assign") {
                if (x.size() != 5) bad("malformed assign");
                s = code.new_stmt(ST_ASSIGN);
                s->lhs = value(x[1].atom);
                s->lhs->def = s;
                s->rhs = rhs(x[4]);
                add_uses(s->rhs, s);
            } else if (k == "if") {
                if (x.size() != 6) bad("malformed if");
                s = code.new_stmt(ST_IF);
                s->cond = rhs(x[1]);
                add_uses(s->cond, s);
                s->cons = stmts(x[3], 0, s);
                s->alt = stmts(x[4], 0, s);
                s->exit = phis(x[5], s);
            } else if (k == "while") {
                if (x.size() != 5) bad("malformed while");
                s = code.new_stmt(ST_WHILE);
                s->entry = phis(x[1], s);
                s->cond = rhs(x[2]);
                add_uses(s->cond, s);
                s->body = stmts(x[4], 0, s);
            } else
                bad("unknown statement " + k);
            s->parent = parent;
            *tail = s;
            tail = &s->next;
        }
        return head;
    }
};

}  // namespace

std::unique_ptr<mmb_module> load_ir_text(const std::string &text) {
    Reader rd(text);
    Sx top = rd.read();
    if (!top.is_list || top.size() < 3 || top[0].atom != "mmir" || top[1].atom != "1") bad("not an (mmir 1 ...) module");
    auto m = std::make_unique<mmb_module>();
    m->mod = std::make_unique<Module>();
    Module &mod = *m->mod;
    register_native_filters(mod);
    std::string main_name;
    // first pass: declare filters so calls can be resolved in any order
    std::vector<const Sx *> filter_forms;
    for (size_t i = 2; i < top.size(); ++i) {
        const Sx &f = top[i];
        if (!f.is_list || f.size() < 1) bad("malformed module item");
        if (f[0].atom == "main") { main_name = f[1].atom; continue; }
        if (f[0].atom != "filter" || f.size() != 6) bad("malformed filter form");
        auto fp = std::make_unique<Filter>();
        fp->kind = FILTER_MATHMAP;
        fp->name = f[1].atom;
        fp->flags = 0;
        for (size_t k = 1; k < f[2].size(); ++k) {
            if (f[2][k].atom == "unit") fp->flags |= IMAGE_FLAG_UNIT;
            if (f[2][k].atom == "square") fp->flags |= IMAGE_FLAG_SQUARE;
        }
        for (size_t k = 1; k < f[3].size(); ++k) {
            const Sx &u = f[3][k];
            UservalInfo ui;
            const std::string &t = u[0].atom;
            static const char *names[] = {"int", "float", "bool", "color", "curve", "gradient", "image"};
            ui.type = -1;
            for (int q = 0; q < 7; ++q) if (t == names[q]) ui.type = q;
            if (ui.type < 0) bad("unknown userval type " + t);
            ui.name = u[1].atom;
            ui.index = (int)k - 1;
            switch (ui.type) {
            case UV_INT: ui.int_min = atoi(u[2].atom.c_str()); ui.int_max = atoi(u[3].atom.c_str()); ui.int_default = atoi(u[4].atom.c_str()); break;
            case UV_FLOAT: ui.float_min = parse_float(u[2].atom); ui.float_max = parse_float(u[3].atom); ui.float_default = parse_float(u[4].atom); break;
            case UV_BOOL: ui.bool_default = atoi(u[2].atom.c_str()); break;
            case UV_IMAGE: ui.image_flags = (unsigned)atoi(u[2].atom.c_str()); break;
            default: break;
            }
            fp->uservals.push_back(ui);
        }
        static const struct { const char *n; int c; } ints[] = {
            {"x", CONST_Y | CONST_T}, {"y", CONST_X | CONST_T}, {"r", CONST_T}, {"a", CONST_T},
            {"t", CONST_X | CONST_Y}, {"R", CONST_ALL}, {"__canvasPixelW", CONST_ALL}, {"__canvasPixelH", CONST_ALL},
            {"__renderPixelW", CONST_ALL}, {"__renderPixelH", CONST_ALL}, {"frame", CONST_X | CONST_Y},
            {"X", CONST_ALL}, {"Y", CONST_ALL}, {"W", CONST_ALL}, {"H", CONST_ALL}};
        for (auto &in : ints) fp->internals.push_back(Internal{in.n, in.c, false});
        fp->index = (int)mod.filters.size();
        mod.filters.push_back(std::move(fp));
        filter_forms.push_back(&f);
    }
    if (filter_forms.empty()) bad("module has no filters");
    m->codes.resize(mod.filters.size());
    size_t k = 0;
    for (auto &fp : mod.filters) {
        if (fp->kind != FILTER_MATHMAP) continue;
        const Sx &f = *filter_forms[k++];
        auto code = std::make_unique<FilterCode>();
        code->filter = fp.get();
        Builder b{mod, *code};
        for (size_t v = 1; v < f[4].size(); ++v) {
            const Sx &d = f[4][v];
            int id = atoi(d[0].atom.c_str());
            CompVar *c = code->new_compvar(type_from_name(d[1].atom));
            c->id = id;
            if (d.size() > 2) c->tuple_len = atoi(d[2].atom.c_str());
            b.cvs[id] = c;
        }
        code->first = b.stmts(f[5], 1, nullptr);
        carry_loop_values(mod, *code);
        propagate_types(*code);
        analyze_constants(*code);
        m->codes[fp->index] = std::move(code);
    }
    mod.main_filter = main_name.empty() ? mod.filters.back().get() : mod.lookup_filter(main_name);
    if (!mod.main_filter || mod.main_filter->kind != FILTER_MATHMAP) bad("main filter not found");
    m->main = mod.main_filter;
    std::vector<const FilterCode *> ptrs;
    for (auto &c : m->codes) if (c) ptrs.push_back(c.get());
    m->ir_text = dump_module_ir(ptrs, m->main->name);
    return m;
}

}  // namespace mmbackend
