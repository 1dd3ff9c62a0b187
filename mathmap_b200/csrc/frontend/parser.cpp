// Scanner + recursive-descent parser + typed tree construction for the MathMap
// language.  Follows the reference's token rules (scanner.c:219-392), grammar
// and operator precedences (parser.y:48-267: ';' < '=' < ||,&&,xor < comparisons
// < +,- < *,/,% < ^(right) < unary < cast ':' ) and the typing performed by the
// tree constructors (exprtree.c:468-1385).  Written as precedence climbing
// instead of an LALR table; semantic actions run in the same order as the
// bison actions so that variables are registered at the same points.
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "frontend.h"

namespace mm {

namespace {

enum Tok {
    TK_EOF = 0, TK_IDENT = 256, TK_STRING, TK_INT, TK_FLOAT, TK_RANGE, TK_FILTER,
    TK_FLOAT_TYPE, TK_INT_TYPE, TK_BOOL_TYPE, TK_COLOR_TYPE, TK_GRADIENT_TYPE, TK_CURVE_TYPE, TK_IMAGE_TYPE,
    TK_IF, TK_THEN, TK_ELSE, TK_END, TK_WHILE, TK_DO, TK_FOR, TK_XOR,
    TK_EQUAL, TK_LE, TK_GE, TK_NE, TK_OR, TK_AND, TK_CONVERT
};

struct Token {
    int kind = TK_EOF;
    std::string text;
    int line = 0, column = 0;
};

[[noreturn]] void fail(const std::string &msg, int line, int column) {
    CompileError e;
    e.message = msg;
    e.line = line;
    e.column = column;
    throw e;
}

class Lexer {
   public:
    explicit Lexer(const std::string &s) : src(s) {}
    Token next() {
        skip_ws_and_comments();
        Token t;
        t.line = line;
        t.column = col;
        if (pos >= src.size()) return t;
        char c = src[pos];
        if (c == '_' || isalpha((unsigned char)c)) {
            size_t b = pos;
            while (pos < src.size() && (src[pos] == '_' || isalnum((unsigned char)src[pos]))) adv();
            t.text = src.substr(b, pos - b);
            t.kind = keyword(t.text, t);
            return t;
        }
        if (c == '"') {
            adv();
            size_t b = pos;
            while (pos < src.size() && src[pos] != '"') adv();
            if (pos >= src.size()) fail("String not terminated", t.line, t.column);
            t.text = src.substr(b, pos - b);
            adv();
            t.kind = TK_STRING;
            return t;
        }
        if (c == '.' || isdigit((unsigned char)c)) {
            // digits with at most one '.', ".." is the range token (scanner.c:287-330)
            size_t b = pos;
            if (c == '.' && pos + 1 < src.size() && src[pos + 1] == '.') {
                adv(); adv();
                t.kind = TK_RANGE;
                t.text = "..";
                return t;
            }
            bool have_dot = false, have_digits = false;
            while (pos < src.size()) {
                char d = src[pos];
                if (isdigit((unsigned char)d)) { have_digits = true; adv(); }
                else if (d == '.') {
                    if (pos + 1 < src.size() && src[pos + 1] == '.' && !have_dot) break;  // "1..5"
                    if (have_dot) break;
                    have_dot = true;
                    adv();
                } else break;
            }
            if (!have_digits) fail("Misplaced decimal point", t.line, t.column);
            t.text = src.substr(b, pos - b);
            t.kind = have_dot ? TK_FLOAT : TK_INT;
            return t;
        }
        static const char *singles = "-<>!,()+*/%=;^:[]";
        char d = pos + 1 < src.size() ? src[pos + 1] : 0;
        struct { const char *s; int k; } twos[] = {{"==", TK_EQUAL}, {"<=", TK_LE}, {">=", TK_GE}, {"!=", TK_NE},
                                                  {"||", TK_OR}, {"&&", TK_AND}, {"::", TK_CONVERT}};
        for (auto &tw : twos)
            if (c == tw.s[0] && d == tw.s[1]) {
                adv(); adv();
                t.kind = tw.k;
                t.text = tw.s;
                return t;
            }
        if (strchr(singles, c)) {
            adv();
            t.kind = c;
            t.text = std::string(1, c);
            return t;
        }
        fail("Illegal character", t.line, t.column);
    }

   private:
    const std::string &src;
    size_t pos = 0;
    int line = 0, col = 0;
    void adv() {
        if (src[pos] == '\n') { ++line; col = 0; } else ++col;
        ++pos;
    }
    void skip_ws_and_comments() {
        for (;;) {
            while (pos < src.size() && isspace((unsigned char)src[pos])) adv();
            if (pos < src.size() && src[pos] == '#') {
                while (pos < src.size() && src[pos] != '\n') adv();
                continue;
            }
            break;
        }
    }
    int keyword(const std::string &s, const Token &t) {
        static const struct { const char *n; int k; } kws[] = {
            {"filter", TK_FILTER}, {"if", TK_IF}, {"then", TK_THEN}, {"else", TK_ELSE}, {"end", TK_END},
            {"while", TK_WHILE}, {"do", TK_DO}, {"for", TK_FOR}, {"xor", TK_XOR},
            {"int", TK_INT_TYPE}, {"float", TK_FLOAT_TYPE}, {"bool", TK_BOOL_TYPE}, {"color", TK_COLOR_TYPE},
            {"curve", TK_CURVE_TYPE}, {"gradient", TK_GRADIENT_TYPE}, {"image", TK_IMAGE_TYPE}};
        for (auto &k : kws)
            if (s == k.n) return k.k;
        if (s == "function" || s == "lambda") fail("The identifier `" + s + "' is reserved.", t.line, t.column);
        return TK_IDENT;
    }
};

// Precedence levels, parser.y:48-59
enum { P_SEQ = 1, P_ASSIGN = 2, P_LOGIC = 3, P_CMP = 4, P_ADD = 5, P_MUL = 6, P_POW = 7, P_UNARY = 8 };

struct BinOp { int prec; bool right; const char *func; };

const BinOp *binop_for(int tok) {
    static const BinOp add{P_ADD, false, "__add"}, sub{P_ADD, false, "__sub"}, mul{P_MUL, false, "__mul"},
        div{P_MUL, false, "__div"}, mod{P_MUL, false, "__mod"}, pw{P_POW, true, "__pow"},
        eq{P_CMP, false, "__equal"}, lt{P_CMP, false, "__less"}, gt{P_CMP, false, "__greater"},
        le{P_CMP, false, "__lessequal"}, ge{P_CMP, false, "__greaterequal"}, ne{P_CMP, false, "__notequal"},
        lor{P_LOGIC, false, "__or"}, land{P_LOGIC, false, "__and"}, lxor{P_LOGIC, false, "__xor"};
    switch (tok) {
    case '+': return &add;
    case '-': return &sub;
    case '*': return &mul;
    case '/': return &div;
    case '%': return &mod;
    case '^': return &pw;
    case TK_EQUAL: return &eq;
    case '<': return &lt;
    case '>': return &gt;
    case TK_LE: return &le;
    case TK_GE: return &ge;
    case TK_NE: return &ne;
    case TK_OR: return &lor;
    case TK_AND: return &land;
    case TK_XOR: return &lxor;
    default: return nullptr;
    }
}

const char *op_text_for_func(const std::string &f) {
    static const struct { const char *f, *o; } t[] = {
        {"__add", "+"}, {"__sub", "-"}, {"__mul", "*"}, {"__div", "/"}, {"__mod", "%"}, {"__pow", "^"},
        {"__equal", "=="}, {"__less", "<"}, {"__greater", ">"}, {"__lessequal", "<="}, {"__greaterequal", ">="},
        {"__notequal", "!="}, {"__or", "||"}, {"__and", "&&"}, {"__xor", "xor"}, {"__neg", "-"}, {"__not", "!"}};
    for (auto &e : t)
        if (f == e.f) return e.o;
    return nullptr;
}

class Parser {
   public:
    Parser(Module &m, const std::string &src) : mod(m), lex(src) { advance(); advance(); }

    void parse_filters() {
        while (cur.kind != TK_EOF) parse_filter();
        if (mod.filters.empty() || !mod.main_filter) fail("At least one filter must be defined.", 0, 0);
    }

   private:
    Module &mod;
    Lexer lex;
    Token cur, peek;
    Filter *filt = nullptr;

    void advance() {
        cur = peek;
        peek = lex.next();
    }
    [[noreturn]] void parse_error() { fail("Parse error.", cur.line, cur.column); }
    void expect(int k) {
        if (cur.kind != k) parse_error();
        advance();
    }
    bool accept(int k) {
        if (cur.kind == k) { advance(); return true; }
        return false;
    }

    // ---- declarations -------------------------------------------------
    std::vector<std::string> parse_options() {
        std::vector<std::string> opts;
        while (cur.kind == TK_IDENT) {
            opts.push_back(cur.text);
            advance();
            if (accept('(')) {  // sub-options are parsed and ignored (mathmap_common.c:59-72 only reads top level)
                parse_options();
                expect(')');
            }
        }
        return opts;
    }
    static unsigned flags_from_options(const std::vector<std::string> &opts) {
        bool pixel = false, stretched = false;
        for (auto &o : opts) { if (o == "pixel") pixel = true; if (o == "stretched") stretched = true; }
        unsigned f = 0;
        if (!pixel) { f |= IMAGE_FLAG_UNIT; if (!stretched) f |= IMAGE_FLAG_SQUARE; }
        return f;
    }

    struct Num { bool is_int; int i; float f; int line, col; };
    Num parse_signed_number() {
        bool neg = accept('-');
        Num n{};
        n.line = cur.line; n.col = cur.column;
        if (cur.kind == TK_INT) { n.is_int = true; n.i = atoi(cur.text.c_str()); if (neg) n.i = -n.i; }
        else if (cur.kind == TK_FLOAT) { n.is_int = false; n.f = (float)strtod(cur.text.c_str(), nullptr); if (neg) n.f = -n.f; }
        else parse_error();
        advance();
        return n;
    }

    void parse_arg_decl(Filter *f) {
        std::vector<std::string> opts = parse_options();
        UservalInfo u;
        int line = cur.line, col = cur.column;
        switch (cur.kind) {
        case TK_FLOAT_TYPE: u.type = UV_FLOAT; break;
        case TK_INT_TYPE: u.type = UV_INT; break;
        case TK_BOOL_TYPE: u.type = UV_BOOL; break;
        case TK_COLOR_TYPE: u.type = UV_COLOR; break;
        case TK_GRADIENT_TYPE: u.type = UV_GRADIENT; break;
        case TK_CURVE_TYPE: u.type = UV_CURVE; break;
        case TK_IMAGE_TYPE: u.type = UV_IMAGE; break;
        case TK_FILTER: fail("Filter-typed arguments are not supported.", line, col);
        default: parse_error();
        }
        advance();
        if (cur.kind != TK_IDENT) parse_error();
        u.name = cur.text;
        advance();
        bool have_limits = false;
        // defaults when no limits are given: mathmap_common.c:99-115
        u.int_min = -100000; u.int_max = 100000; u.int_default = 0;
        u.float_min = -1.f; u.float_max = 1.f; u.float_default = 0.f;
        int lim_imin = 0, lim_imax = 0;
        float lim_fmin = 0, lim_fmax = 0;
        if (accept(':')) {
            Num lo = parse_signed_number();
            expect('-');
            Num hi = parse_signed_number();
            have_limits = true;
            bool int_limits = lo.is_int && hi.is_int;
            if (int_limits) {
                if (lo.i >= hi.i) fail("Lower limit must be less than upper limit", lo.line, lo.col);
            } else {
                float a = lo.is_int ? (float)lo.i : lo.f, b = hi.is_int ? (float)hi.i : hi.f;
                if (a >= b) fail("Lower limit must be less than upper limit", lo.line, lo.col);
            }
            if (u.type == UV_INT) {
                if (!int_limits) fail("Only integers can be limits for an int argument", lo.line, lo.col);
                lim_imin = lo.i; lim_imax = hi.i;
                u.int_min = lo.i; u.int_max = hi.i; u.int_default = lo.i;
            } else if (u.type == UV_FLOAT) {
                lim_fmin = lo.is_int ? (float)lo.i : lo.f;
                lim_fmax = hi.is_int ? (float)hi.i : hi.f;
                u.float_min = lim_fmin; u.float_max = lim_fmax; u.float_default = lim_fmin;
            } else
                fail("Limits applied to wrongly typed argument", lo.line, lo.col);
        }
        if (cur.kind == '(') {
            advance();
            Num d = parse_signed_number();
            expect(')');
            switch (u.type) {
            case UV_INT:
                if (!d.is_int) fail("Only integers can be defaults for an int argument", d.line, d.col);
                if (d.i < lim_imin || d.i > lim_imax) fail("Default value outside of bounds", d.line, d.col);
                if (have_limits) u.int_default = d.i;
                break;
            case UV_FLOAT: {
                float v = d.is_int ? (float)d.i : d.f;
                if (v < lim_fmin || v > lim_fmax) fail("Default value outside of bounds", d.line, d.col);
                if (have_limits) u.float_default = v;
                break;
            }
            case UV_BOOL:
                if (!d.is_int) fail("Only integers can be defaults for a bool argument", d.line, d.col);
                u.bool_default = d.i ? 1 : 0;
                break;
            default: fail("Default applied to wrongly typed argument", d.line, d.col);
            }
        }
        if (cur.kind == TK_STRING) { u.doc = cur.text; advance(); }
        if (u.type == UV_IMAGE) u.image_flags = flags_from_options(opts);
        for (auto &o : f->uservals)
            if (o.name == u.name) fail("The argument `" + u.name + "' is declared more than once.", line, col);
        u.index = (int)f->uservals.size();
        f->uservals.push_back(u);
    }

    void parse_filter() {
        std::vector<std::string> opts = parse_options();
        expect(TK_FILTER);
        if (cur.kind != TK_IDENT) parse_error();
        auto fp = std::make_unique<Filter>();
        Filter *f = fp.get();
        f->kind = FILTER_MATHMAP;
        f->name = cur.text;
        int line = cur.line, col = cur.column;
        advance();
        expect('(');
        if (cur.kind != ')') {
            parse_arg_decl(f);
            while (accept(',')) parse_arg_decl(f);
        }
        expect(')');
        if (cur.kind == TK_STRING) { f->doc = cur.text; advance(); }
        f->flags = flags_from_options(opts);
        if (mod.lookup_filter(f->name)) fail("The filter `" + f->name + "' is defined more than once.", line, col);
        // internals and their constness: mathmap_common.c:193-215
        static const struct { const char *n; int c; } ints[] = {
            {"x", CONST_Y | CONST_T}, {"y", CONST_X | CONST_T}, {"r", CONST_T}, {"a", CONST_T},
            {"t", CONST_X | CONST_Y}, {"R", CONST_ALL}, {"__canvasPixelW", CONST_ALL}, {"__canvasPixelH", CONST_ALL},
            {"__renderPixelW", CONST_ALL}, {"__renderPixelH", CONST_ALL}, {"frame", CONST_X | CONST_Y},
            {"X", CONST_ALL}, {"Y", CONST_ALL}, {"W", CONST_ALL}, {"H", CONST_ALL}};
        for (auto &i : ints) f->internals.push_back(Internal{i.n, i.c, false});
        f->index = (int)mod.filters.size();
        mod.filters.push_back(std::move(fp));
        filt = f;
        Expr *body = parse_expr(P_SEQ);
        accept(';');
        expect(TK_END);
        if (body->result.tag != mod.rgba_tag || body->result.length != 4)
            fail("The filter `" + f->name + "' must have the result type rgba:4.", body->line, body->column);
        f->body = body;
        mod.main_filter = f;
        filt = nullptr;
    }

    // ---- tree constructors (exprtree.c) -------------------------------
    Expr *mk(ExprKind k, const Token &at) {
        Expr *e = mod.new_expr(k);
        e->line = at.line;
        e->column = at.column;
        return e;
    }
    Expr *make_int(int v, const Token &at) {
        Expr *e = mk(EX_INT_CONST, at);
        e->int_const = v;
        e->result = {mod.nil_tag, 1};
        return e;
    }
    Expr *make_float(float v, const Token &at) {
        Expr *e = mk(EX_FLOAT_CONST, at);
        e->float_const = v;
        e->result = {mod.nil_tag, 1};
        return e;
    }
    Variable *lookup_variable(const std::string &n) {
        for (auto &v : filt->variables)
            if (v->name == n) return v.get();
        return nullptr;
    }
    Variable *register_variable(const std::string &n, TupleInfo ti) {
        auto v = std::make_unique<Variable>();
        v->name = n;
        v->tag = ti.tag;
        v->length = ti.length;
        filt->variables.push_back(std::move(v));
        return filt->variables.back().get();
    }
    const UservalInfo *lookup_userval(const std::string &n) {
        for (auto &u : filt->uservals)
            if (u.name == n) return &u;
        return nullptr;
    }
    Expr *make_var_expr(Variable *v, const Token &at) {
        Expr *e = mk(EX_VARIABLE, at);
        e->var = v;
        e->result = {v->tag, v->length};
        return e;
    }
    Expr *make_tuple(std::vector<Expr *> elems, const Token &at) {
        for (Expr *el : elems)
            if (el->result.length != 1) fail("Tuples cannot contain tuples of length other than 1.", el->line, el->column);
        Expr *e = mk(EX_TUPLE, at);
        e->args = std::move(elems);
        e->result = {mod.nil_tag, (int)e->args.size()};
        return e;
    }
    Expr *make_cast(const std::string &tag, Expr *tuple, const Token &at) {
        Expr *e = mk(EX_CAST, at);
        e->tagnum = mod.tag_number(tag);
        e->a = tuple;
        e->result = {e->tagnum, tuple->result.length};
        return e;
    }
    Expr *make_internal_or_null(const std::string &name, const Token &at) {
        Internal *in = filt->lookup_internal(name, true);
        if (!in) return nullptr;
        Expr *e = mk(EX_INTERNAL, at);
        e->internal = in;
        e->result = {mod.nil_tag, 1};
        return e;
    }
    Expr *make_var(const std::string &name, const Token &at) {
        if (Expr *e = make_internal_or_null(name, at)) return e;
        // variable macros: macros.c:181-189
        if (name == "xy") return make_cast("xy", make_tuple({make_var("x", at), make_var("y", at)}, at), at);
        if (name == "ra") return make_cast("ra", make_tuple({make_var("r", at), make_var("a", at)}, at), at);
        if (name == "XY") return make_cast("xy", make_tuple({make_var("X", at), make_var("Y", at)}, at), at);
        if (name == "WH") return make_cast("xy", make_tuple({make_var("W", at), make_var("H", at)}, at), at);
        if (name == "I") return make_cast("ri", make_tuple({make_int(0, at), make_int(1, at)}, at), at);
        if (name == "pi") return make_float((float)M_PI, at);
        if (name == "e") return make_float((float)M_E, at);
        if (const UservalInfo *u = lookup_userval(name)) return make_userval(u, {}, at);
        if (Variable *v = lookup_variable(name)) return make_var_expr(v, at);
        fail("Undefined variable " + name + ".", at.line, at.column);
    }
    static bool is_variable_macro(const std::string &n) {
        return n == "xy" || n == "ra" || n == "XY" || n == "WH" || n == "I" || n == "pi" || n == "e";
    }
    Expr *make_userval(const UservalInfo *u, std::vector<Expr *> args, const Token &at) {
        Expr *e = mk(EX_USERVAL, at);
        e->userval = u;
        switch (u->type) {
        case UV_INT: case UV_FLOAT: case UV_BOOL: case UV_COLOR:
            if (!args.empty()) fail("Number, bool and color inputs take no arguments.", at.line, at.column);
            e->result = u->type == UV_COLOR ? TupleInfo{mod.rgba_tag, 4} : TupleInfo{mod.nil_tag, 1};
            return e;
        case UV_CURVE:
            e->result = {mod.curve_tag, 1};
            if (args.size() == 1) { args.insert(args.begin(), e); return make_function("__applyCurve", args, at); }
            if (!args.empty()) fail("A curve takes one argument.", at.line, at.column);
            return e;
        case UV_GRADIENT:
            e->result = {mod.gradient_tag, 1};
            if (args.size() == 1) { args.insert(args.begin(), e); return make_function("__applyGradient", args, at); }
            if (!args.empty()) fail("A gradient takes one argument.", at.line, at.column);
            return e;
        case UV_IMAGE:
            e->result = {mod.image_tag, 1};
            if (args.size() == 1 || args.size() == 2) { args.push_back(e); return make_function("__origVal", args, at); }
            if (!args.empty()) fail("An image takes one or two arguments.", at.line, at.column);
            return e;
        }
        return e;
    }
    Expr *make_filter_call(Filter *f, std::vector<Expr *> args, const Token &at) {
        int nuv = (int)f->uservals.size();
        int n = (int)args.size();
        if (n < nuv || n >= nuv + 3)
            fail("Filter " + f->name + " takes " + std::to_string(nuv) + " to " + std::to_string(nuv + 2) +
                     " arguments but is called with " + std::to_string(n) + ".", at.line, at.column);
        bool is_closure = n == nuv;
        for (int i = 0; i < nuv; ++i) {
            const UservalInfo &u = f->uservals[i];
            if (u.type == UV_COLOR) {
                if (args[i]->result.tag != mod.rgba_tag || args[i]->result.length != 4)
                    fail("Can only pass tuples of type rgba:4 as colors.", args[i]->line, args[i]->column);
            } else if (args[i]->result.length != 1)
                fail("Can only pass tuples of length 1 as numbers, booleans, curves, gradients, or images.", args[i]->line, args[i]->column);
        }
        Expr *closure = mk(EX_FILTER_CLOSURE, at);
        closure->filter = f;
        closure->args.assign(args.begin(), args.begin() + nuv);
        closure->result = {mod.image_tag, 1};
        if (is_closure) return closure;
        Expr *coord = args[nuv];
        if (coord->result.length != 2 || (coord->result.tag != mod.xy_tag && coord->result.tag != mod.ra_tag))
            fail("The coordinate argument to a filter must be a tuple of type xy:2 or ra:2.", coord->line, coord->column);
        if (coord->result.tag == mod.ra_tag) coord = make_function("toXY", {coord}, at);
        Expr *time;
        if (n == nuv + 2) {
            time = args[nuv + 1];
            if (time->result.length != 1) fail("The time argument to a filter must be a tuple of length 1.", time->line, time->column);
        } else
            time = make_var("t", at);
        return make_function("__origVal", {coord, time, closure}, at);
    }
    Expr *make_image_call(Expr *image, std::vector<Expr *> args, const Token &at) {
        if (args.size() != 1 && args.size() != 2) fail("An image must be invoked with one or two arguments.", at.line, at.column);
        Expr *coord = args[0];
        if (coord->result.length != 2 || (coord->result.tag != mod.xy_tag && coord->result.tag != mod.ra_tag))
            fail("The coordinate argument to an image must be of type xy:2 or ra:2.", at.line, at.column);
        if (coord->result.tag == mod.ra_tag) args[0] = make_function("toXY", {coord}, at);
        if (args.size() == 2 && args[1]->result.length != 1) fail("The time argument to an image have length 1.", at.line, at.column);
        args.push_back(image);
        return make_function("__origVal", args, at);
    }

   public:
    Expr *make_function(const std::string &name, std::vector<Expr *> args, const Token &at) {
        if (const UservalInfo *u = lookup_userval(name)) return make_userval(u, args, at);
        if (Filter *f = mod.lookup_filter(name)) return make_filter_call(f, args, at);
        if (args.empty()) fail("Unable to resolve invocation of function `" + name + "'.", at.line, at.column);
        std::vector<TupleInfo> infos;
        for (Expr *a : args) infos.push_back(a->result);
        TupleInfo result;
        if (const OverloadEntry *en = mod.resolve(name, infos, &result)) {
            if (en->gen) {
                Expr *e = mk(EX_FUNC, at);
                e->entry = en;
                e->args = args;
                e->result = result;
                return e;
            }
            cur_macro_token = &at;
            Expr *e = en->macro(args);
            e->line = at.line;
            e->column = at.column;
            return e;
        }
        if (Variable *v = lookup_variable(name)) {
            if (v->tag != mod.image_tag || v->length != 1)
                fail("Variable " + name + " is not an image and cannot be invoked.", at.line, at.column);
            return make_image_call(make_var_expr(v, at), args, at);
        }
        if (const char *op = op_text_for_func(name))
            fail(std::string("Unable to resolve invocation of operator `") + op + "'.", at.line, at.column);
        fail("Unable to resolve invocation of function `" + name + "'.", at.line, at.column);
    }
    const Token *cur_macro_token = nullptr;

    // macros.c:120-166: bind the coordinate to a temporary, convert to xy, supply t
    Expr *macro_orig_val_image(std::vector<Expr *> &args, bool with_frame) {
        const Token &at = *cur_macro_token;
        char buf[64];
        snprintf(buf, sizeof buf, "__tmp%d", ++mod.gensym_counter);
        Variable *tmp = register_variable(buf, args[0]->result);
        Expr *assign = make_assignment_to(tmp, args[0], at);
        Expr *xy = make_function("toXY", {make_var_expr(tmp, at)}, at);
        Expr *call;
        if (with_frame)
            call = make_function("__origVal", {xy, args[1], args[2]}, at);
        else
            call = make_function("__origVal", {xy, make_var("t", at), args[1]}, at);
        return make_sequence(assign, call, at);
    }

   private:
    Expr *make_sequence(Expr *l, Expr *r, const Token &at) {
        Expr *e = mk(EX_SEQUENCE, at);
        e->a = l;
        e->b = r;
        e->result = r->result;
        return e;
    }
    Expr *make_assignment_to(Variable *v, Expr *value, const Token &at) {
        if (v->tag != value->result.tag || v->length != value->result.length)
            fail("Variable " + v->name + " is being assigned two different types.", at.line, at.column);
        Expr *e = mk(EX_ASSIGNMENT, at);
        e->var = v;
        e->a = value;
        e->result = value->result;
        return e;
    }
    Expr *make_assignment(const std::string &name, Expr *value, const Token &at) {
        Variable *v = lookup_variable(name);
        if (!v) {
            if (filt->lookup_internal(name, false) || is_variable_macro(name))
                fail("Cannot assign to internal variable `" + name + "'.", at.line, at.column);
            if (lookup_userval(name)) fail("Cannot assign to filter argument `" + name + "'.", at.line, at.column);
            v = register_variable(name, value->result);
        }
        return make_assignment_to(v, value, at);
    }
    Expr *make_sub_assignment(const std::string &name, std::vector<Expr *> subs, Expr *value, const Token &at) {
        Variable *v = lookup_variable(name);
        if (!v) fail("Undefined variable " + name + ".", at.line, at.column);
        if ((int)subs.size() != value->result.length) fail("Lhs does not match rhs in sub assignment.", at.line, at.column);
        Expr *e = mk(EX_SUB_ASSIGNMENT, at);
        e->var = v;
        e->args = std::move(subs);
        e->a = value;
        e->result = value->result;
        return e;
    }
    Expr *make_select(Expr *tuple, std::vector<Expr *> subs, const Token &at) {
        Expr *e = mk(EX_SELECT, at);
        e->a = tuple;
        int n = (int)subs.size();
        for (Expr *s : subs)
            if (s->result.length != 1) fail("Tuples cannot contain tuples of length other than 1.", s->line, s->column);
        e->args = std::move(subs);
        e->result = n == 1 ? TupleInfo{mod.nil_tag, 1} : TupleInfo{tuple->result.tag, n};
        return e;
    }
    Expr *make_if(Expr *c, Expr *t, Expr *f, const Token &at) {
        if (c->result.length != 1) fail("Condition to if statement must have length 1.", c->line, c->column);
        if (f && (t->result.tag != f->result.tag || t->result.length != f->result.length))
            fail("Consequent and alternative must have the same type in if statement.", t->line, t->column);
        Expr *e = mk(f ? EX_IF_THEN_ELSE : EX_IF_THEN, at);
        e->a = c;
        e->b = t;
        e->c = f;
        e->result = t->result;
        return e;
    }
    Expr *make_while(Expr *inv, Expr *body, bool do_while, const Token &at) {
        if (inv->result.length != 1) fail("Invariant of while loop must have length 1.", inv->line, inv->column);
        Expr *e = mk(do_while ? EX_DO_WHILE : EX_WHILE, at);
        e->a = inv;
        e->b = body;
        e->result = {mod.nil_tag, 1};
        return e;
    }

    // ---- expressions --------------------------------------------------
    std::vector<Expr *> parse_exprlist(int closer) {
        std::vector<Expr *> v;
        if (cur.kind == closer) return v;
        v.push_back(parse_expr(P_SEQ));
        while (accept(',')) v.push_back(parse_expr(P_SEQ));
        return v;
    }
    std::vector<Expr *> parse_subscripts() {
        std::vector<Expr *> v;
        do {
            if (cur.kind == TK_INT && peek.kind == TK_RANGE) {
                Token at = cur;
                int first = atoi(cur.text.c_str());
                advance();
                advance();
                if (cur.kind != TK_INT) parse_error();
                int last = atoi(cur.text.c_str());
                advance();
                if (first > last) fail("Invalid range " + std::to_string(first) + ".." + std::to_string(last) + ".", at.line, at.column);
                for (int i = first; i <= last; ++i) v.push_back(make_int(i, at));
            } else
                v.push_back(parse_expr(P_SEQ));
        } while (accept(','));
        return v;
    }
    void parse_end() {
        accept(';');
        expect(TK_END);
    }

    Expr *parse_primary() {
        Token at = cur;
        switch (cur.kind) {
        case TK_INT: advance(); return make_int(atoi(at.text.c_str()), at);
        case TK_FLOAT: advance(); return make_float((float)strtod(at.text.c_str(), nullptr), at);
        case '[': {
            advance();
            std::vector<Expr *> el = parse_exprlist(']');
            expect(']');
            if (el.empty()) parse_error();
            return make_tuple(el, at);
        }
        case '(': {
            advance();
            Expr *e = parse_expr(P_SEQ);
            expect(')');
            if (cur.kind == '[') {
                advance();
                std::vector<Expr *> subs = parse_subscripts();
                expect(']');
                return make_select(e, subs, at);
            }
            return e;
        }
        case TK_IF: {
            advance();
            Expr *c = parse_expr(P_SEQ);
            expect(TK_THEN);
            Expr *t = parse_expr(P_SEQ);
            accept(';');
            if (accept(TK_ELSE)) {
                Expr *f = parse_expr(P_SEQ);
                parse_end();
                return make_if(c, t, f, at);
            }
            expect(TK_END);
            return make_if(c, t, nullptr, at);
        }
        case TK_WHILE: {
            advance();
            Expr *c = parse_expr(P_SEQ);
            expect(TK_DO);
            Expr *b = parse_expr(P_SEQ);
            parse_end();
            return make_while(c, b, false, at);
        }
        case TK_DO: {
            advance();
            Expr *b = parse_expr(P_SEQ);
            expect(TK_WHILE);
            Expr *c = parse_expr(P_SEQ);
            parse_end();
            return make_while(c, b, true, at);
        }
        case TK_FOR: {
            advance();
            if (cur.kind != TK_IDENT) parse_error();
            std::string counter = cur.text;
            Token cat = cur;
            advance();
            expect('=');
            Expr *start = parse_expr(P_LOGIC);
            if (start->result.length != 1)
                fail("The start and end of a for loop interval must be tuples of length 1.", start->line, start->column);
            Expr *counter_init = make_assignment(counter, start, cat);
            expect(TK_RANGE);
            Expr *end = parse_expr(P_SEQ);
            expect(TK_DO);
            Expr *body = parse_expr(P_SEQ);
            parse_end();
            if (end->result.length != 1 || start->result.tag != end->result.tag)
                fail("The start and end of a for loop interval must be tuples of the same tag and length 1.", cat.line, cat.column);
            // desugaring: exprtree.c:1355-1385
            char buf[64];
            snprintf(buf, sizeof buf, "__gensym%d", ++mod.gensym_counter);
            Expr *end_init = make_assignment(buf, end, cat);
            Expr *init = make_sequence(counter_init, end_init, cat);
            Expr *inc = make_assignment(counter, make_function("__add", {make_var(counter, cat), make_int(1, cat)}, cat), cat);
            Expr *inv = make_function("__lessequal", {make_var(counter, cat), make_var(buf, cat)}, cat);
            return make_sequence(init, make_while(inv, make_sequence(body, inc, cat), false, cat), cat);
        }
        case TK_IDENT: {
            std::string name = cur.text;
            advance();
            if (cur.kind == '(') {
                advance();
                std::vector<Expr *> args = parse_exprlist(')');
                expect(')');
                return make_function(name, args, at);
            }
            if (cur.kind == '[') {
                advance();
                std::vector<Expr *> subs = parse_subscripts();
                expect(']');
                if (cur.kind == '=') {
                    advance();
                    Expr *v = parse_expr(P_LOGIC);
                    return make_sub_assignment(name, subs, v, at);
                }
                return make_select(make_var(name, at), subs, at);
            }
            if (cur.kind == ':') {
                advance();
                Expr *v = parse_unary();
                return make_cast(name, v, at);
            }
            if (cur.kind == TK_CONVERT) fail("The `::' conversion operator is not supported.", at.line, at.column);
            if (cur.kind == '=') {
                advance();
                Expr *v = parse_expr(P_LOGIC);
                return make_assignment(name, v, at);
            }
            return make_var(name, at);
        }
        default: parse_error();
        }
    }

    Expr *parse_unary() {
        if (cur.kind == '-' || cur.kind == '!') {
            Token at = cur;
            advance();
            Expr *v = parse_unary();
            return make_function(at.kind == '-' ? "__neg" : "__not", {v}, at);
        }
        return parse_primary();
    }

    Expr *parse_expr(int min_prec) {
        Expr *lhs = parse_unary();
        for (;;) {
            if (cur.kind == ';') {
                if (min_prec > P_SEQ) break;
                if (peek.kind == TK_END || peek.kind == TK_ELSE) break;
                Token at = cur;
                advance();
                Expr *rhs = parse_expr(P_SEQ + 1);
                lhs = make_sequence(lhs, rhs, at);
                continue;
            }
            const BinOp *op = binop_for(cur.kind);
            if (!op || op->prec < min_prec) break;
            Token at = cur;
            advance();
            Expr *rhs = parse_expr(op->right ? op->prec : op->prec + 1);
            lhs = make_function(op->func, {lhs, rhs}, at);
        }
        return lhs;
    }

    friend void mm::parse_module(Module &, const std::string &);
};

}  // namespace

void register_native_filters(Module &m) {
    // native filters are registered before the user's filters: mathmap_common.c:346-377
    if (!m.filters.empty()) return;
    auto add_native = [&](const char *name, const char *func, std::vector<UservalInfo> uvs) {
        auto f = std::make_unique<Filter>();
        f->kind = FILTER_NATIVE;
        f->name = name;
        f->native_name = func;
        for (size_t i = 0; i < uvs.size(); ++i) uvs[i].index = (int)i;
        f->uservals = std::move(uvs);
        f->index = (int)m.filters.size();
        m.filters.push_back(std::move(f));
    };
    auto img = [](const char *n) { UservalInfo u; u.name = n; u.type = UV_IMAGE; u.image_flags = 0; return u; };
    auto flt = [](const char *n, float lo, float hi, float d) {
        UservalInfo u; u.name = n; u.type = UV_FLOAT; u.float_min = lo; u.float_max = hi; u.float_default = d; return u;
    };
    auto bl = [](const char *n, int d) { UservalInfo u; u.name = n; u.type = UV_BOOL; u.bool_default = d; return u; };
    add_native("gaussian_blur", "native_filter_gaussian_blur",
               {img("in"), flt("horizontal_std_dev", 0.f, 2.f, 0.01f), flt("vertical_std_dev", 0.f, 2.f, 0.01f)});
    add_native("convolve", "native_filter_convolve", {img("in"), img("kernel"), bl("normalize", 1), bl("copy_alpha", 1)});
    add_native("half_convolve", "native_filter_half_convolve", {img("in"), img("mask"), bl("copy_alpha", 1)});
    add_native("visualize_fft", "native_filter_visualize_fft", {img("in"), bl("ignore_alpha", 1)});
}

void parse_module(Module &m, const std::string &source) {
    register_native_filters(m);
    Parser p(m, source);
    // the __origVal macros (macros.c:191-194) need the parser to build trees
    m.register_macro("__origVal", "rgba:4 <- xy:2, image:1", [&p](std::vector<Expr *> &a) { return p.macro_orig_val_image(a, false); });
    m.register_macro("__origVal", "rgba:4 <- ra:2, image:1", [&p](std::vector<Expr *> &a) { return p.macro_orig_val_image(a, false); });
    m.register_macro("__origVal", "rgba:4 <- ra:2, nil:1, image:1", [&p](std::vector<Expr *> &a) { return p.macro_orig_val_image(a, true); });
    p.parse_filters();
    // the macro closures reference the parser on this stack frame; drop them
    for (auto &o : m.overloads) o.macro = nullptr;
}

}  // namespace mm
