// The MathMap builtin function library as IR generators.
//
// WHAT each overload computes — signature, registration order (which decides
// overload resolution), the exact association order of every sum and product,
// integer-vs-float literals, the guards (x/0 -> 0, log(x<=0) -> 0, ...) — is the
// contract taken from reference builtins.lisp:432-1395.  HOW it is written is
// ours: plain C++ lambdas over the Gen vocabulary instead of a Lisp DSL that
// prints C.  Integer literals below are deliberate: they stay TYPE_INT in the
// IR exactly where the reference's literals are integers.
#include <cmath>

#include "irgen.h"

namespace mm {

namespace {

struct Ctx {
    Gen &g;
    GenArgs &a;
    P arg(int k, int i) const { return g.cur(a.args[k][i]); }
    int len(int k) const { return a.lengths[k]; }
    void res(int i, P p) const { g.assign(a.result[i], g.rhs_prim(p)); }
    int nres() const { return (int)a.result.size(); }
    // sum of a list the way the reference's (sum expr) expands: t0 + t1, then + t2 ...
    P sum(const std::vector<P> &v) const { return g.sum(v); }
    P cplx(P re, P im) const { return g.op2(OP_COMPLEX, re, im); }
    void res_complex(P c) const {
        res(0, g.op1(OP_C_REAL, c));
        res(1, g.op1(OP_C_IMAG, c));
    }
};

#define BUILTIN(name, impl, spec, ...)                                   \
    m.register_builtin(name, impl, spec, [](Gen &g, GenArgs &ga) {       \
        Ctx c{g, ga};                                                    \
        (void)c;                                                         \
        __VA_ARGS__                                                      \
    })

// element-wise binary op with a scalar/elementwise right operand chooser
template <class F> void each(const Ctx &c, F f) {
    for (int i = 0; i < c.nres(); ++i) c.res(i, f(i));
}

}  // namespace

// Helpers to register families that differ only by the IR op.
template <int OP> static void gen_complex_unary(Gen &g, GenArgs &ga) {
    Ctx c{g, ga};
    c.res_complex(g.op1(OP, c.cplx(c.arg(0, 0), c.arg(0, 1))));
}
template <int OP> static void gen_real_unary(Gen &g, GenArgs &ga) {
    Ctx c{g, ga};
    c.res(0, g.op1(OP, c.arg(0, 0)));
}
template <int OP, int N> static void gen_real_nary(Gen &g, GenArgs &ga) {
    Ctx c{g, ga};
    std::vector<P> args;
    for (int i = 0; i < N; ++i) args.push_back(c.arg(i, 0));
    c.res(0, g.op(OP, args));
}
template <int OP> static void gen_elementwise_nn(Gen &g, GenArgs &ga) {
    Ctx c{g, ga};
    each(c, [&](int i) { return g.op2(OP, c.arg(0, i), c.arg(1, i)); });
}
template <int OP> static void gen_elementwise_ns(Gen &g, GenArgs &ga) {
    Ctx c{g, ga};
    each(c, [&](int i) { return g.op2(OP, c.arg(0, i), c.arg(1, 0)); });
}
// x OP y with "y == 0 -> 0" guard, three shapes (builtins.lisp:635-669)
template <int OP> static void gen_guarded_1(Gen &g, GenArgs &ga) {
    Ctx c{g, ga};
    g.if_(g.c_eq(c.arg(1, 0), g.ic(0)), [&] { c.res(0, g.ic(0)); }, [&] { c.res(0, g.op2(OP, c.arg(0, 0), c.arg(1, 0))); });
}
template <int OP> static void gen_guarded_s(Gen &g, GenArgs &ga) {
    Ctx c{g, ga};
    g.if_(g.c_eq(c.arg(1, 0), g.ic(0)), [&] { each(c, [&](int) { return g.ic(0); }); },
          [&] { each(c, [&](int i) { return g.op2(OP, c.arg(0, i), c.arg(1, 0)); }); });
}
template <int OP> static void gen_guarded_n(Gen &g, GenArgs &ga) {
    Ctx c{g, ga};
    for (int i = 0; i < c.len(1); ++i)
        g.if_(g.c_eq(c.arg(1, i), g.ic(0)), [&] { c.res(i, g.ic(0)); }, [&] { c.res(i, g.op2(OP, c.arg(0, i), c.arg(1, i))); });
}
// 4-component algebras: result[k] = sum over 4 signed products a[i]*b[j]
struct Term { int i, j, sign; };
static void gen_algebra4(Gen &g, GenArgs &ga, const Term t[4][4]) {
    Ctx c{g, ga};
    for (int k = 0; k < 4; ++k) {
        std::vector<P> terms;
        for (int n = 0; n < 4; ++n) {
            P p = g.mul(c.arg(0, t[k][n].i), c.arg(1, t[k][n].j));
            terms.push_back(t[k][n].sign < 0 ? g.neg(p) : p);
        }
        c.res(k, c.sum(terms));
    }
}
static void gen_abs_euclid(Gen &g, GenArgs &ga) {
    Ctx c{g, ga};
    std::vector<P> sq;
    for (int i = 0; i < c.len(0); ++i) sq.push_back(g.mul(c.arg(0, i), c.arg(0, i)));
    c.res(0, g.op1(OP_SQRT, c.sum(sq)));
}
static void gen_noise_full(Gen &g, GenArgs &ga, int op) {
    Ctx c{g, ga};
    c.res(0, g.op(op, {c.arg(0, 0), c.arg(1, 0), c.arg(2, 0), c.arg(3, 0), c.arg(3, 1), c.arg(3, 2)}));
}
// complex Jacobi elliptic functions (builtins.lisp:1013-1036)
static void gen_ell_jac_ri(Gen &g, GenArgs &ga, int which) {
    Ctx c{g, ga};
    P m = c.arg(1, 0);
    P v = g.op2(OP_ELL_JAC, c.arg(0, 0), m);
    P v1 = g.op2(OP_ELL_JAC, c.arg(0, 1), g.sub(g.ic(1), m));
    auto nth = [&](P t, int i) { return g.op2(OP_TUPLE_NTH, t, g.ic(i)); };
    P s = nth(v, 0), cc = nth(v, 1), d = nth(v, 2), s1 = nth(v1, 0), c1 = nth(v1, 1), d1 = nth(v1, 2);
    P denom = g.add(g.mul(c1, c1), g.mul(m, g.mul(g.mul(s, s), g.mul(s1, s1))));
    P rn, in;
    if (which == 0) { rn = g.mul(s, d1); in = g.mul(g.mul(cc, d), g.mul(s1, c1)); }
    else if (which == 1) { rn = g.mul(cc, c1); in = g.neg(g.mul(g.mul(s, d), g.mul(s1, d1))); }
    else { rn = g.mul(c1, g.mul(d, d1)); in = g.sub(g.mul(s, s1), g.mul(m, cc)); }
    c.res(0, g.div(rn, denom));
    c.res(1, g.div(in, denom));
}

void register_all_builtins(Module &m) {
    BUILTIN("print", "print", "nil:1 <- _:_", {
        for (int i = 0; i < c.len(0); ++i) g.op1(OP_PRINT, c.arg(0, i));
        g.op(OP_NEWLINE, {});
        c.res(0, g.ic(0));
    });

    // ---- addition / subtraction ------------------------------------------
    m.register_builtin("__add", "add_ri", "ri:2 <- ri:2, ri:2", gen_elementwise_nn<OP_ADD>);
    BUILTIN("__add", "add_ri_1", "ri:2 <- ri:2, _:1", {
        c.res(0, g.add(c.arg(0, 0), c.arg(1, 0)));
        c.res(1, g.add(c.arg(0, 1), g.ic(0)));
    });
    BUILTIN("__add", "add_1_ri", "ri:2 <- _:1, ri:2", {
        c.res(0, g.add(c.arg(1, 0), c.arg(0, 0)));
        c.res(1, g.add(c.arg(1, 1), g.ic(0)));
    });
    m.register_builtin("__add", "add_1", "T:1 <- T:1, T:1", gen_elementwise_nn<OP_ADD>);
    m.register_builtin("__add", "add_s", "T:L <- T:L, _:1", gen_elementwise_ns<OP_ADD>);
    m.register_builtin("__add", "add_n", "T:L <- T:L, T:L", gen_elementwise_nn<OP_ADD>);

    m.register_builtin("__sub", "sub_ri", "ri:2 <- ri:2, ri:2", gen_elementwise_nn<OP_SUB>);
    BUILTIN("__sub", "sub_ri_1", "ri:2 <- ri:2, _:1", {
        c.res(0, g.sub(c.arg(0, 0), c.arg(1, 0)));
        c.res(1, g.sub(c.arg(0, 1), g.ic(0)));
    });
    BUILTIN("__sub", "sub_1_ri", "ri:2 <- _:1, ri:2", {
        c.res(0, g.sub(c.arg(0, 0), c.arg(1, 0)));
        c.res(1, g.sub(g.ic(0), c.arg(1, 1)));
    });
    m.register_builtin("__sub", "sub_1", "T:1 <- T:1, T:1", gen_elementwise_nn<OP_SUB>);
    m.register_builtin("__sub", "sub_s", "T:L <- T:L, _:1", gen_elementwise_ns<OP_SUB>);
    m.register_builtin("__sub", "sub_n", "T:L <- T:L, T:L", gen_elementwise_nn<OP_SUB>);

    BUILTIN("__neg", "neg", "T:L <- T:L", { each(c, [&](int i) { return g.neg(c.arg(0, i)); }); });

    // ---- multiplication --------------------------------------------------
    BUILTIN("__mul", "mul_ri", "ri:2 <- ri:2, ri:2", {
        c.res(0, g.sub(g.mul(c.arg(0, 0), c.arg(1, 0)), g.mul(c.arg(0, 1), c.arg(1, 1))));
        c.res(1, g.add(g.mul(c.arg(0, 0), c.arg(1, 1)), g.mul(c.arg(1, 0), c.arg(0, 1))));
    });
    BUILTIN("__mul", "mul_1_ri", "ri:2 <- _:1, ri:2", { each(c, [&](int i) { return g.mul(c.arg(0, 0), c.arg(1, i)); }); });
    BUILTIN("__mul", "mul_m2x2", "m2x2:4 <- m2x2:4, m2x2:4", {
        const int n = 2;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
            std::vector<P> t;
            for (int k = 0; k < n; ++k) t.push_back(g.mul(c.arg(0, i * n + k), c.arg(1, k * n + j)));
            c.res(i * n + j, c.sum(t));
        }
    });
    BUILTIN("__mul", "mul_m3x3", "m3x3:9 <- m3x3:9, m3x3:9", {
        const int n = 3;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
            std::vector<P> t;
            for (int k = 0; k < n; ++k) t.push_back(g.mul(c.arg(0, i * n + k), c.arg(1, k * n + j)));
            c.res(i * n + j, c.sum(t));
        }
    });
    BUILTIN("__mul", "mul_v2m2x2", "v2:2 <- v2:2, m2x2:4", {
        const int n = 2;
        for (int i = 0; i < n; ++i) {
            std::vector<P> t;
            for (int j = 0; j < n; ++j) t.push_back(g.mul(c.arg(0, j), c.arg(1, i + n * j)));
            c.res(i, c.sum(t));
        }
    });
    BUILTIN("__mul", "mul_v3m3x3", "v3:3 <- v3:3, m3x3:9", {
        const int n = 3;
        for (int i = 0; i < n; ++i) {
            std::vector<P> t;
            for (int j = 0; j < n; ++j) t.push_back(g.mul(c.arg(0, j), c.arg(1, i + n * j)));
            c.res(i, c.sum(t));
        }
    });
    BUILTIN("__mul", "mul_m2x2v2", "v2:2 <- m2x2:4, v2:2", {
        const int n = 2;
        for (int i = 0; i < n; ++i) {
            std::vector<P> t;
            for (int j = 0; j < n; ++j) t.push_back(g.mul(c.arg(0, j + n * i), c.arg(1, j)));
            c.res(i, c.sum(t));
        }
    });
    BUILTIN("__mul", "mul_m3x3v3", "v3:3 <- m3x3:9, v3:3", {
        const int n = 3;
        for (int i = 0; i < n; ++i) {
            std::vector<P> t;
            for (int j = 0; j < n; ++j) t.push_back(g.mul(c.arg(0, j + n * i), c.arg(1, j)));
            c.res(i, c.sum(t));
        }
    });
    // quaternions, "cquat" and hypercomplex products: sign tables per component
    BUILTIN("__mul", "mul_quat", "quat:4 <- quat:4, quat:4", {
        static const Term t[4][4] = {{{0, 0, 1}, {1, 1, -1}, {2, 2, -1}, {3, 3, -1}},
                                     {{0, 1, 1}, {1, 0, 1}, {2, 3, 1}, {3, 2, -1}},
                                     {{0, 2, 1}, {2, 0, 1}, {1, 3, -1}, {3, 1, 1}},
                                     {{0, 3, 1}, {3, 0, 1}, {1, 2, 1}, {2, 1, -1}}};
        gen_algebra4(g, ga, t);
    });
    BUILTIN("__mul", "mul_cquat", "cquat:4 <- cquat:4, cquat:4", {
        static const Term t[4][4] = {{{0, 0, 1}, {1, 1, -1}, {2, 2, 1}, {3, 3, 1}},
                                     {{0, 1, 1}, {1, 0, 1}, {2, 3, 1}, {3, 2, 1}},
                                     {{0, 2, 1}, {2, 0, 1}, {1, 3, -1}, {3, 1, -1}},
                                     {{0, 3, 1}, {3, 0, 1}, {1, 2, -1}, {2, 1, -1}}};
        gen_algebra4(g, ga, t);
    });
    BUILTIN("__mul", "mul_hyper", "hyper:4 <- hyper:4, hyper:4", {
        static const Term t[4][4] = {{{0, 0, 1}, {1, 1, -1}, {2, 2, -1}, {3, 3, 1}},
                                     {{0, 1, 1}, {1, 0, 1}, {2, 3, -1}, {3, 2, -1}},
                                     {{0, 2, 1}, {2, 0, 1}, {1, 3, -1}, {3, 1, -1}},
                                     {{0, 3, 1}, {3, 0, 1}, {1, 2, 1}, {2, 1, 1}}};
        gen_algebra4(g, ga, t);
    });
    m.register_builtin("__mul", "mul_1", "T:1 <- T:1, T:1", gen_elementwise_nn<OP_MUL>);
    m.register_builtin("__mul", "mul_s", "T:L <- T:L, _:1", gen_elementwise_ns<OP_MUL>);
    m.register_builtin("__mul", "mul_n", "T:L <- T:L, T:L", gen_elementwise_nn<OP_MUL>);

    // ---- division / remainder ------------------------------------------
    BUILTIN("__div", "div_ri", "ri:2 <- ri:2, ri:2", {
        g.if_(g.c_and(g.c_eq(c.arg(1, 0), g.ic(0)), g.c_eq(c.arg(1, 1), g.ic(0))),
              [&] { c.res(0, g.ic(0)); c.res(1, g.ic(0)); },
              [&] {
                  P cc = g.add(g.mul(c.arg(1, 0), c.arg(1, 0)), g.mul(c.arg(1, 1), c.arg(1, 1)));
                  c.res(0, g.div(g.add(g.mul(c.arg(0, 0), c.arg(1, 0)), g.mul(c.arg(0, 1), c.arg(1, 1))), cc));
                  c.res(1, g.div(g.add(g.mul(g.neg(c.arg(0, 0)), c.arg(1, 1)), g.mul(c.arg(1, 0), c.arg(0, 1))), cc));
              });
    });
    BUILTIN("__div", "div_1_ri", "ri:2 <- T:1, ri:2", {
        P tmp = g.add(g.mul(c.arg(1, 0), c.arg(1, 0)), g.mul(c.arg(1, 1), c.arg(1, 1)));
        g.if_(g.c_eq(tmp, g.ic(0)), [&] { c.res(0, g.ic(0)); c.res(1, g.ic(0)); },
              [&] {
                  c.res(0, g.div(g.mul(c.arg(0, 0), c.arg(1, 0)), tmp));
                  c.res(1, g.neg(g.div(g.mul(c.arg(0, 0), c.arg(1, 1)), tmp)));
              });
    });
    BUILTIN("__div", "div_v2m2x2", "v2:2 <- _:2, m2x2:4", {
        CompVar *mt = g.temp(T_TUPLE);
        g.assign(mt, g.rhs_tuple({c.arg(1, 0), c.arg(1, 1), c.arg(1, 2), c.arg(1, 3)}));
        CompVar *vt = g.temp(T_TUPLE);
        g.assign(vt, g.rhs_tuple({c.arg(0, 0), c.arg(0, 1)}));
        P r = g.op2(OP_SOLVE_LINEAR_2, g.cur(mt), g.cur(vt));
        for (int i = 0; i < 2; ++i) c.res(i, g.op2(OP_TUPLE_NTH, r, g.ic(i)));
    });
    BUILTIN("__div", "div_v3m3x3", "v3:3 <- _:3, m3x3:9", {
        std::vector<P> mm;
        for (int i = 0; i < 9; ++i) mm.push_back(c.arg(1, i));
        CompVar *mt = g.temp(T_TUPLE);
        g.assign(mt, g.rhs_tuple(mm));
        CompVar *vt = g.temp(T_TUPLE);
        g.assign(vt, g.rhs_tuple({c.arg(0, 0), c.arg(0, 1), c.arg(0, 2)}));
        P r = g.op2(OP_SOLVE_LINEAR_3, g.cur(mt), g.cur(vt));
        for (int i = 0; i < 3; ++i) c.res(i, g.op2(OP_TUPLE_NTH, r, g.ic(i)));
    });
    m.register_builtin("__div", "div_1", "T:1 <- T:1, T:1", gen_guarded_1<OP_DIV>);
    m.register_builtin("__div", "div_s", "T:L <- T:L, _:1", gen_guarded_s<OP_DIV>);
    m.register_builtin("__div", "div_n", "T:L <- T:L, T:L", gen_guarded_n<OP_DIV>);
    m.register_builtin("__mod", "mod_1", "T:1 <- T:1, T:1", gen_guarded_1<OP_MOD>);
    m.register_builtin("__mod", "mod_s", "T:L <- T:L, _:1", gen_guarded_s<OP_MOD>);
    m.register_builtin("__mod", "mod_n", "T:L <- T:L, T:L", gen_guarded_n<OP_MOD>);
    BUILTIN("pmod", "pmod", "T:1 <- T:1, T:1", {
        P md = g.op2(OP_MOD, c.arg(0, 0), c.arg(1, 0));
        g.if_(g.c_less(c.arg(0, 0), g.ic(0)), [&] { c.res(0, g.add(md, c.arg(1, 0))); }, [&] { c.res(0, md); });
    });

    m.register_builtin("sqrt", "sqrt_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_SQRT>);
    m.register_builtin("sqrt", "sqrt_1", "T:1 <- T:1", gen_real_unary<OP_SQRT>);
    BUILTIN("sum", "sum", "nil:1 <- T:L", {
        std::vector<P> v;
        for (int i = 0; i < c.len(0); ++i) v.push_back(c.arg(0, i));
        c.res(0, c.sum(v));
    });

    // ---- vectors ---------------------------------------------------------
    BUILTIN("dotp", "dotp", "nil:1 <- T:L, T:L", {
        std::vector<P> v;
        for (int i = 0; i < c.len(0); ++i) v.push_back(g.mul(c.arg(0, i), c.arg(1, i)));
        c.res(0, c.sum(v));
    });
    BUILTIN("crossp", "crossp", "T:3 <- T:3, T:3", {
        c.res(0, g.sub(g.mul(c.arg(0, 1), c.arg(1, 2)), g.mul(c.arg(0, 2), c.arg(1, 1))));
        c.res(1, g.sub(g.mul(c.arg(0, 2), c.arg(1, 0)), g.mul(c.arg(0, 0), c.arg(1, 2))));
        c.res(2, g.sub(g.mul(c.arg(0, 0), c.arg(1, 1)), g.mul(c.arg(0, 1), c.arg(1, 0))));
    });
    BUILTIN("det", "det_m2x2", "nil:1 <- m2x2:4", {
        c.res(0, g.sub(g.mul(c.arg(0, 0), c.arg(0, 3)), g.mul(c.arg(0, 1), c.arg(0, 2))));
    });
    BUILTIN("det", "det_m3x3", "nil:1 <- m3x3:9", {
        auto p3 = [&](int x, int y, int z) { return g.prod({c.arg(0, x), c.arg(0, y), c.arg(0, z)}); };
        P pos = c.sum({p3(0, 4, 8), p3(1, 5, 6), p3(2, 3, 7)});
        P ng = c.sum({p3(2, 4, 6), p3(0, 5, 7), p3(1, 3, 8)});
        c.res(0, g.sub(pos, ng));
    });
    BUILTIN("normalize", "normalize", "T:L <- T:L", {
        std::vector<P> sq;
        for (int i = 0; i < c.len(0); ++i) sq.push_back(g.mul(c.arg(0, i), c.arg(0, i)));
        P l = c.sum(sq);
        g.if_(g.c_eq(l, g.ic(0)), [&] { each(c, [&](int) { return g.ic(0); }); },
              [&] { each(c, [&](int i) { return g.div(c.arg(0, i), g.op1(OP_SQRT, l)); }); });
    });
    BUILTIN("abs", "abs_ri", "nil:1 <- ri:2", { c.res(0, g.op2(OP_HYPOT, c.arg(0, 0), c.arg(0, 1))); });
    m.register_builtin("abs", "abs_quat", "nil:1 <- quat:4", gen_abs_euclid);
    m.register_builtin("abs", "abs_cquat", "nil:1 <- cquat:4", gen_abs_euclid);
    m.register_builtin("abs", "abs_hyper", "nil:1 <- hyper:4", gen_abs_euclid);
    m.register_builtin("abs", "abs_v2", "nil:1 <- v2:2", gen_abs_euclid);
    m.register_builtin("abs", "abs_v3", "nil:1 <- v3:3", gen_abs_euclid);
    BUILTIN("abs", "abs_1", "T:1 <- T:1", { c.res(0, g.op1(OP_ABS, c.arg(0, 0))); });
    BUILTIN("abs", "abs_n", "T:L <- T:L", { each(c, [&](int i) { return g.op1(OP_ABS, c.arg(0, i)); }); });

    // ---- trigonometry ----------------------------------------------------
    BUILTIN("deg2rad", "deg2rad", "nil:1 <- _:1", { c.res(0, g.mul(c.arg(0, 0), g.fc(0.017453292519943295722f))); });
    BUILTIN("rad2deg", "rad2deg", "deg:1 <- _:1", { c.res(0, g.mul(c.arg(0, 0), g.fc(57.2957795130823208768f))); });
    m.register_builtin("sin", "sin_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_SIN>);
    m.register_builtin("sin", "sin", "T:1 <- T:1", gen_real_unary<OP_SIN>);
    m.register_builtin("cos", "cos_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_COS>);
    m.register_builtin("cos", "cos", "T:1 <- T:1", gen_real_unary<OP_COS>);
    m.register_builtin("tan", "tan_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_TAN>);
    m.register_builtin("tan", "tan", "T:1 <- T:1", gen_real_unary<OP_TAN>);
    m.register_builtin("asin", "asin_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_ASIN>);
    BUILTIN("asin", "asin", "T:1 <- T:1", {
        g.if_(g.c_or(g.c_less(c.arg(0, 0), g.ic(-1)), g.c_less(g.ic(1), c.arg(0, 0))), [&] { c.res(0, g.ic(0)); },
              [&] { c.res(0, g.op1(OP_ASIN, c.arg(0, 0))); });
    });
    m.register_builtin("acos", "acos_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_ACOS>);
    BUILTIN("acos", "acos", "T:1 <- T:1", {
        g.if_(g.c_or(g.c_less(c.arg(0, 0), g.ic(-1)), g.c_less(g.ic(1), c.arg(0, 0))), [&] { c.res(0, g.ic(0)); },
              [&] { c.res(0, g.op1(OP_ACOS, c.arg(0, 0))); });
    });
    m.register_builtin("atan", "atan_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_ATAN>);
    m.register_builtin("atan", "atan", "T:1 <- T:1", gen_real_unary<OP_ATAN>);
    m.register_builtin("atan", "atan2", "T:1 <- T:1, T:1", gen_real_nary<OP_ATAN2, 2>);

    // ---- powers, exponentials -------------------------------------------
    BUILTIN("__pow", "pow_ri_1", "ri:2 <- ri:2, T:1", {
        c.res_complex(g.op2(OP_C_POW, c.cplx(c.arg(0, 0), c.arg(0, 1)), c.cplx(c.arg(1, 0), g.fc(0.0f))));
    });
    BUILTIN("__pow", "pow_ri", "ri:2 <- ri:2, ri:2", {
        c.res_complex(g.op2(OP_C_POW, c.cplx(c.arg(0, 0), c.arg(0, 1)), c.cplx(c.arg(1, 0), c.arg(1, 1))));
    });
    BUILTIN("__pow", "pow_1_ri", "ri:2 <- T:1, ri:2", {
        c.res_complex(g.op2(OP_C_POW, c.cplx(c.arg(0, 0), g.fc(0.0f)), c.cplx(c.arg(1, 0), c.arg(1, 1))));
    });
    BUILTIN("__pow", "pow_1", "T:1 <- T:1, T:1", {
        g.if_(g.c_and(g.c_leq(c.arg(1, 0), g.ic(0)), g.c_eq(c.arg(0, 0), g.ic(0))), [&] { c.res(0, g.ic(0)); },
              [&] { c.res(0, g.op2(OP_POW, c.arg(0, 0), c.arg(1, 0))); });
    });
    BUILTIN("__pow", "pow_s", "T:L <- T:L, _:1", {
        for (int i = 0; i < c.len(0); ++i)
            g.if_(g.c_and(g.c_leq(c.arg(1, 0), g.ic(0)), g.c_eq(c.arg(0, i), g.ic(0))), [&] { c.res(i, g.ic(0)); },
                  [&] { c.res(i, g.op2(OP_POW, c.arg(0, i), c.arg(1, 0))); });
    });
    m.register_builtin("exp", "exp_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_EXP>);
    m.register_builtin("exp", "exp_1", "T:1 <- T:1", gen_real_unary<OP_EXP>);
    m.register_builtin("log", "log_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_LOG>);
    BUILTIN("log", "log_1", "T:1 <- T:1", {
        g.if_(g.c_leq(c.arg(0, 0), g.ic(0)), [&] { c.res(0, g.ic(0)); }, [&] { c.res(0, g.op1(OP_LOG, c.arg(0, 0))); });
    });
    BUILTIN("arg", "arg_ri", "nil:1 <- ri:2", { c.res(0, g.op1(OP_C_ARG, c.cplx(c.arg(0, 0), c.arg(0, 1)))); });
    BUILTIN("conj", "conj_ri", "ri:2 <- ri:2", {
        c.res(0, c.arg(0, 0));
        c.res(1, g.neg(c.arg(0, 1)));
    });
    m.register_builtin("sinh", "sinh_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_SINH>);
    m.register_builtin("sinh", "sinh_1", "T:1 <- T:1", gen_real_unary<OP_SINH>);
    m.register_builtin("cosh", "cosh_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_COSH>);
    m.register_builtin("cosh", "cosh_1", "T:1 <- T:1", gen_real_unary<OP_COSH>);
    m.register_builtin("tanh", "tanh_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_TANH>);
    m.register_builtin("tanh", "tanh_1", "T:1 <- T:1", gen_real_unary<OP_TANH>);
    m.register_builtin("asinh", "asinh_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_ASINH>);
    m.register_builtin("asinh", "asinh_1", "T:1 <- T:1", gen_real_unary<OP_ASINH>);
    m.register_builtin("acosh", "acosh_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_ACOSH>);
    m.register_builtin("acosh", "acosh_1", "T:1 <- T:1", gen_real_unary<OP_ACOSH>);
    m.register_builtin("atanh", "atanh_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_ATANH>);
    m.register_builtin("atanh", "atanh_1", "T:1 <- T:1", gen_real_unary<OP_ATANH>);
    m.register_builtin("gamma", "gamma_ri", "ri:2 <- ri:2", gen_complex_unary<OP_C_GAMMA>);
    BUILTIN("gamma", "gamma_1", "T:1 <- T:1", {
        g.if_(g.c_less(c.arg(0, 0), g.ic(0)), [&] { c.res(0, g.ic(0)); }, [&] { c.res(0, g.op1(OP_GAMMA, c.arg(0, 0))); });
    });
    BUILTIN("beta", "beta_1", "T:1 <- T:1, T:1", {
        g.if_(g.c_or(g.c_less(c.arg(0, 0), g.ic(0)), g.c_less(c.arg(1, 0), g.ic(0))), [&] { c.res(0, g.ic(0)); },
              [&] { c.res(0, g.op2(OP_BETA, c.arg(0, 0), c.arg(1, 0))); });
    });

    // ---- elliptic integrals / functions ----------------------------------
    m.register_builtin("ell_int_Kcomp", "ell_int_Kcomp", "T:1 <- T:1", gen_real_nary<OP_ELL_INT_K_COMP, 1>);
    m.register_builtin("ell_int_Ecomp", "ell_int_Ecomp", "T:1 <- T:1", gen_real_nary<OP_ELL_INT_E_COMP, 1>);
    m.register_builtin("ell_int_F", "ell_int_F", "T:1 <- T:1, T:1", gen_real_nary<OP_ELL_INT_F, 2>);
    m.register_builtin("ell_int_E", "ell_int_E", "T:1 <- T:1, T:1", gen_real_nary<OP_ELL_INT_E, 2>);
    m.register_builtin("ell_int_P", "ell_int_P", "T:1 <- T:1, T:1, T:1", gen_real_nary<OP_ELL_INT_P, 3>);
    m.register_builtin("ell_int_D", "ell_int_D", "T:1 <- T:1, T:1, T:1", gen_real_nary<OP_ELL_INT_D, 3>);
    m.register_builtin("ell_int_RC", "ell_int_RC", "T:1 <- T:1, T:1", gen_real_nary<OP_ELL_INT_RC, 2>);
    m.register_builtin("ell_int_RD", "ell_int_RD", "T:1 <- T:1, T:1, T:1", gen_real_nary<OP_ELL_INT_RD, 3>);
    m.register_builtin("ell_int_RF", "ell_int_RF", "T:1 <- T:1, T:1, T:1", gen_real_nary<OP_ELL_INT_RF, 3>);
    m.register_builtin("ell_int_RJ", "ell_int_RJ", "T:1 <- T:1, T:1, T:1, T:1", gen_real_nary<OP_ELL_INT_RJ, 4>);
    BUILTIN("ell_jac_sn", "ell_jac_sn_1", "T:1 <- T:1, T:1", {
        c.res(0, g.op2(OP_TUPLE_NTH, g.op2(OP_ELL_JAC, c.arg(0, 0), c.arg(1, 0)), g.ic(0)));
    });
    BUILTIN("ell_jac_cn", "ell_jac_cn_1", "T:1 <- T:1, T:1", {
        c.res(0, g.op2(OP_TUPLE_NTH, g.op2(OP_ELL_JAC, c.arg(0, 0), c.arg(1, 0)), g.ic(1)));
    });
    BUILTIN("ell_jac_dn", "ell_jac_dn_1", "T:1 <- T:1, T:1", {
        c.res(0, g.op2(OP_TUPLE_NTH, g.op2(OP_ELL_JAC, c.arg(0, 0), c.arg(1, 0)), g.ic(2)));
    });
    BUILTIN("ell_jac_sn", "ell_jac_sn_ri", "ri:2 <- ri:2, _:1", { gen_ell_jac_ri(g, ga, 0); });
    BUILTIN("ell_jac_cn", "ell_jac_cn_ri", "ri:2 <- ri:2, _:1", { gen_ell_jac_ri(g, ga, 1); });
    BUILTIN("ell_jac_dn", "ell_jac_dn_ri", "ri:2 <- ri:2, _:1", { gen_ell_jac_ri(g, ga, 2); });

    // ---- rounding, selection, interpolation --------------------------------
    m.register_builtin("floor", "floor", "T:1 <- T:1", gen_real_unary<OP_FLOOR>);
    m.register_builtin("ceil", "ceil", "T:1 <- T:1", gen_real_unary<OP_CEIL>);
    BUILTIN("sign", "sign_n", "T:L <- T:L", {
        for (int i = 0; i < c.len(0); ++i)
            g.if_(g.c_less(c.arg(0, i), g.ic(0)), [&] { c.res(i, g.ic(-1)); },
                  [&] { g.if_(g.c_less(g.ic(0), c.arg(0, i)), [&] { c.res(i, g.ic(1)); }, [&] { c.res(i, g.ic(0)); }); });
    });
    BUILTIN("min", "min_n", "T:L <- T:L, T:L", { each(c, [&](int i) { return g.op2(OP_MIN, c.arg(0, i), c.arg(1, i)); }); });
    BUILTIN("max", "max_n", "T:L <- T:L, T:L", { each(c, [&](int i) { return g.op2(OP_MAX, c.arg(0, i), c.arg(1, i)); }); });
    BUILTIN("clamp", "clamp", "T:L <- T:L, T:L, T:L", {
        for (int i = 0; i < c.len(0); ++i)
            g.if_(g.c_less(c.arg(0, i), c.arg(1, i)), [&] { c.res(i, c.arg(1, i)); }, [&] {
                g.if_(g.c_less(c.arg(2, i), c.arg(0, i)), [&] { c.res(i, c.arg(2, i)); }, [&] { c.res(i, c.arg(0, i)); });
            });
    });
    BUILTIN("lerp", "lerp_1", "T:L <- _:1, T:L, T:L", {
        P l = g.sub(g.ic(1), c.arg(0, 0));
        for (int i = 0; i < c.len(1); ++i) c.res(i, g.add(g.mul(l, c.arg(1, i)), g.mul(c.arg(0, 0), c.arg(2, i))));
    });
    BUILTIN("lerp", "lerp_n", "T:L <- T:L, T:L, T:L", {
        for (int i = 0; i < c.len(1); ++i)
            c.res(i, g.add(g.mul(g.sub(g.ic(1), c.arg(0, i)), c.arg(1, i)), g.mul(c.arg(0, i), c.arg(2, i))));
    });
    BUILTIN("scale", "scale", "T:L <- T:L, T:L, T:L, T:L, T:L", {
        for (int i = 0; i < c.len(0); ++i) {
            P dv = g.sub(c.arg(2, i), c.arg(1, i));
            g.if_(g.c_eq(dv, g.ic(0)), [&] { c.res(i, g.ic(0)); }, [&] {
                P q = g.div(g.sub(c.arg(0, i), c.arg(1, i)), dv);
                c.res(i, g.add(g.mul(q, g.sub(c.arg(4, i), c.arg(3, i))), c.arg(3, i)));
            });
        }
    });

    // ---- logic (eager at the language level; guards short-circuit in the IR)
    BUILTIN("__not", "not", "T:1 <- T:1", {
        g.if_(g.c_eq(c.arg(0, 0), g.ic(0)), [&] { c.res(0, g.ic(1)); }, [&] { c.res(0, g.ic(0)); });
    });
    BUILTIN("__or", "or", "T:1 <- T:1, T:1", {
        g.if_(g.c_and(g.c_eq(c.arg(0, 0), g.ic(0)), g.c_eq(c.arg(1, 0), g.ic(0))), [&] { c.res(0, g.ic(0)); }, [&] { c.res(0, g.ic(1)); });
    });
    BUILTIN("__and", "and", "T:1 <- T:1, T:1", {
        g.if_(g.c_or(g.c_eq(c.arg(0, 0), g.ic(0)), g.c_eq(c.arg(1, 0), g.ic(0))), [&] { c.res(0, g.ic(0)); }, [&] { c.res(0, g.ic(1)); });
    });
    BUILTIN("__xor", "xor", "T:1 <- T:1, T:1", {
        g.if_(g.c_or(g.c_and(g.c_not(g.c_eq(c.arg(0, 0), g.ic(0))), g.c_eq(c.arg(1, 0), g.ic(0))),
                     g.c_and(g.c_not(g.c_eq(c.arg(1, 0), g.ic(0))), g.c_eq(c.arg(0, 0), g.ic(0)))),
              [&] { c.res(0, g.ic(1)); }, [&] { c.res(0, g.ic(0)); });
    });

    // ---- comparison ------------------------------------------------------
    BUILTIN("__equal", "equal_ri", "nil:1 <- ri:2, ri:2", {
        g.if_(g.c_and(g.c_eq(c.arg(0, 0), c.arg(1, 0)), g.c_eq(c.arg(0, 1), c.arg(1, 1))), [&] { c.res(0, g.ic(1)); }, [&] { c.res(0, g.ic(0)); });
    });
    BUILTIN("__equal", "equal_ri_1", "nil:1 <- ri:2, _:1", {
        g.if_(g.c_and(g.c_eq(c.arg(0, 0), c.arg(1, 0)), g.c_eq(c.arg(0, 1), g.ic(0))), [&] { c.res(0, g.ic(1)); }, [&] { c.res(0, g.ic(0)); });
    });
    BUILTIN("__equal", "equal_1_ri", "nil:1 <- _:1, ri:2", {
        g.if_(g.c_and(g.c_eq(c.arg(1, 0), c.arg(0, 0)), g.c_eq(c.arg(1, 1), g.ic(0))), [&] { c.res(0, g.ic(1)); }, [&] { c.res(0, g.ic(0)); });
    });
    BUILTIN("__equal", "equal", "nil:1 <- T:1, T:1", { c.res(0, g.op2(OP_EQ, c.arg(0, 0), c.arg(1, 0))); });
    BUILTIN("__less", "less", "nil:1 <- T:1, T:1", { c.res(0, g.op2(OP_LESS, c.arg(0, 0), c.arg(1, 0))); });
    BUILTIN("__greater", "greater", "nil:1 <- T:1, T:1", { c.res(0, g.op2(OP_LESS, c.arg(1, 0), c.arg(0, 0))); });
    BUILTIN("__lessequal", "lessequal", "nil:1 <- T:1, T:1", { c.res(0, g.op2(OP_LEQ, c.arg(0, 0), c.arg(1, 0))); });
    BUILTIN("__greaterequal", "greaterequal", "nil:1 <- T:1, T:1", { c.res(0, g.op2(OP_LEQ, c.arg(1, 0), c.arg(0, 0))); });
    BUILTIN("__notequal", "notequal", "nil:1 <- T:1, T:1", { c.res(0, g.op1(OP_NOT, g.op2(OP_EQ, c.arg(0, 0), c.arg(1, 0)))); });
    BUILTIN("inintv", "inintv", "nil:1 <- T:1, T:1, T:1", {
        g.if_(g.c_and(g.c_leq(c.arg(1, 0), c.arg(0, 0)), g.c_leq(c.arg(0, 0), c.arg(2, 0))), [&] { c.res(0, g.ic(1)); }, [&] { c.res(0, g.ic(0)); });
    });

    // ---- application -----------------------------------------------------
    BUILTIN("__applyCurve", "apply_curve", "nil:1 <- curve:1, _:1", { c.res(0, g.op2(OP_APPLY_CURVE, c.arg(0, 0), c.arg(1, 0))); });
    BUILTIN("__applyGradient", "apply_gradient", "rgba:4 <- gradient:1, _:1", {
        P t = g.op2(OP_APPLY_GRADIENT, c.arg(0, 0), c.arg(1, 0));
        for (int i = 0; i < 4; ++i) c.res(i, g.op2(OP_TUPLE_NTH, t, g.ic(i)));
    });
    BUILTIN("__origVal", "origValXY", "rgba:4 <- xy:2, nil:1, image:1", {
        P t = g.op(OP_ORIG_VAL, {c.arg(0, 0), c.arg(0, 1), c.arg(2, 0), c.arg(1, 0)});
        for (int i = 0; i < 4; ++i) c.res(i, g.op2(OP_TUPLE_NTH, t, g.ic(i)));
    });
    BUILTIN("render", "render", "image:1 <- image:1", {
        c.res(0, g.op(OP_RENDER, {c.arg(0, 0), g.internal("__renderPixelW"), g.internal("__renderPixelH")}, T_IMAGE));
    });
    BUILTIN("pixelSize", "pixelSize", "xy:2 <- image:1", {
        c.res(0, g.op1(OP_IMAGE_PIXEL_WIDTH, c.arg(0, 0)));
        c.res(1, g.op1(OP_IMAGE_PIXEL_HEIGHT, c.arg(0, 0)));
    });

    // ---- colours ---------------------------------------------------------
    BUILTIN("red", "red", "nil:1 <- rgba:4", { c.res(0, c.arg(0, 0)); });
    BUILTIN("green", "green", "nil:1 <- rgba:4", { c.res(0, c.arg(0, 1)); });
    BUILTIN("blue", "blue", "nil:1 <- rgba:4", { c.res(0, c.arg(0, 2)); });
    BUILTIN("alpha", "alpha", "nil:1 <- rgba:4", { c.res(0, c.arg(0, 3)); });
    BUILTIN("gray", "gray", "nil:1 <- rgba:4", {
        c.res(0, c.sum({g.mul(g.fc(0.299f), c.arg(0, 0)), g.mul(g.fc(0.587f), c.arg(0, 1)), g.mul(g.fc(0.114f), c.arg(0, 2))}));
    });
    BUILTIN("rgbColor", "rgbColor", "rgba:4 <- T:1, T:1, T:1", {
        c.res(0, c.arg(0, 0)); c.res(1, c.arg(1, 0)); c.res(2, c.arg(2, 0)); c.res(3, g.ic(1));
    });
    BUILTIN("rgbaColor", "rgbaColor", "rgba:4 <- T:1, T:1, T:1, T:1", {
        c.res(0, c.arg(0, 0)); c.res(1, c.arg(1, 0)); c.res(2, c.arg(2, 0)); c.res(3, c.arg(3, 0));
    });
    BUILTIN("grayColor", "grayColor", "rgba:4 <- T:1", {
        c.res(0, c.arg(0, 0)); c.res(1, c.arg(0, 0)); c.res(2, c.arg(0, 0)); c.res(3, g.ic(1));
    });
    BUILTIN("grayaColor", "grayaColor", "rgba:4 <- T:1, T:1", {
        c.res(0, c.arg(0, 0)); c.res(1, c.arg(0, 0)); c.res(2, c.arg(0, 0)); c.res(3, c.arg(1, 0));
    });
    BUILTIN("toHSVA", "toHSVA", "hsva:4 <- rgba:4", {
        auto clamp01 = [&](P p) { return g.op2(OP_MAX, g.ic(0), g.op2(OP_MIN, g.ic(1), p)); };
        P r = clamp01(c.arg(0, 0)), gg = clamp01(c.arg(0, 1)), b = clamp01(c.arg(0, 2));
        c.res(3, clamp01(c.arg(0, 3)));
        P mx = g.op2(OP_MAX, r, g.op2(OP_MAX, gg, b));
        P mn = g.op2(OP_MIN, r, g.op2(OP_MIN, gg, b));
        c.res(2, mx);
        g.if_(g.c_eq(mx, g.ic(0)), [&] { c.res(0, g.ic(0)); c.res(1, g.ic(0)); }, [&] {
            P delta = g.sub(mx, mn);
            CompVar *h = g.let(g.ic(0));
            c.res(1, g.div(delta, mx));
            g.if_(g.c_eq(r, mx), [&] { g.set(h, g.div(g.sub(gg, b), delta)); }, [&] {
                g.if_(g.c_eq(gg, mx), [&] { g.set(h, g.add(g.ic(2), g.div(g.sub(b, r), delta))); },
                      [&] { g.set(h, g.add(g.ic(4), g.div(g.sub(r, gg), delta))); });
            });
            g.set(h, g.div(g.cur(h), g.fc(6.0f)));
            g.if_(g.c_less(g.cur(h), g.ic(0)), [&] { c.res(0, g.add(g.cur(h), g.ic(1))); }, [&] { c.res(0, g.cur(h)); });
        });
    });
    BUILTIN("toRGBA", "toRGBA", "rgba:4 <- hsva:4", {
        auto clamp01 = [&](P p) { return g.op2(OP_MAX, g.ic(0), g.op2(OP_MIN, g.ic(1), p)); };
        P s = clamp01(c.arg(0, 1)), v = clamp01(c.arg(0, 2));
        c.res(3, clamp01(c.arg(0, 3)));
        auto set3 = [&](P x, P y, P z) { c.res(0, x); c.res(1, y); c.res(2, z); };
        g.if_(g.c_eq(s, g.ic(0)), [&] { set3(v, v, v); }, [&] {
            CompVar *h = g.temp();
            g.assign(h, g.rhs_op(OP_MAX, {g.ic(0), c.arg(0, 0)}));
            g.if_(g.c_leq(g.ic(1), g.cur(h)), [&] { g.set(h, g.ic(0)); }, [&] { g.set(h, g.mul(g.cur(h), g.ic(6))); });
            P hh = g.cur(h);
            P i = g.op1(OP_FLOOR, hh);
            P f = g.sub(hh, i);
            P p = g.mul(v, g.sub(g.ic(1), s));
            P q = g.mul(v, g.sub(g.ic(1), g.mul(s, f)));
            P t = g.mul(v, g.sub(g.ic(1), g.mul(s, g.sub(g.ic(1), f))));
            g.if_(g.c_eq(i, g.ic(0)), [&] { set3(v, t, p); }, [&] {
                g.if_(g.c_eq(i, g.ic(1)), [&] { set3(q, v, p); }, [&] {
                    g.if_(g.c_eq(i, g.ic(2)), [&] { set3(p, v, t); }, [&] {
                        g.if_(g.c_eq(i, g.ic(3)), [&] { set3(p, q, v); }, [&] {
                            g.if_(g.c_eq(i, g.ic(4)), [&] { set3(t, p, v); }, [&] { set3(v, p, q); });
                        });
                    });
                });
            });
        });
    });

    // ---- coordinates -----------------------------------------------------
    BUILTIN("toXY", "toXY", "xy:2 <- ra:2", {
        c.res(0, g.mul(g.op1(OP_COS, c.arg(0, 1)), c.arg(0, 0)));
        c.res(1, g.mul(g.op1(OP_SIN, c.arg(0, 1)), c.arg(0, 0)));
    });
    BUILTIN("toXY", "toXY_trivial", "xy:2 <- xy:2", { c.res(0, c.arg(0, 0)); c.res(1, c.arg(0, 1)); });
    BUILTIN("toRA", "toRA", "ra:2 <- xy:2", {
        P r = g.op2(OP_HYPOT, c.arg(0, 0), c.arg(0, 1));
        g.if_(g.c_eq(r, g.ic(0)), [&] { c.res(0, g.ic(0)); c.res(1, g.ic(0)); }, [&] {
            P a = g.op1(OP_ACOS, g.div(c.arg(0, 0), r));
            c.res(0, r);
            g.if_(g.c_less(c.arg(0, 1), g.ic(0)), [&] { c.res(1, g.sub(g.mul(g.ic(2), g.fc((float)M_PI)), a)); }, [&] { c.res(1, a); });
        });
    });
    BUILTIN("toRA", "toRA_trivial", "ra:2 <- ra:2", { c.res(0, c.arg(0, 0)); c.res(1, c.arg(0, 1)); });

    // ---- random / noise ----------------------------------------------------
    BUILTIN("rand", "rand", "T:1 <- T:1, T:1", { c.res(0, g.op2(OP_RAND, c.arg(0, 0), c.arg(1, 0))); });
    BUILTIN("noise", "noise-perlin-simple", "nil:1 <- _:3", {
        c.res(0, g.op(OP_LIBNOISE_PERLIN, {g.ic(1), g.ic(0), g.ic(0), c.arg(0, 0), c.arg(0, 1), c.arg(0, 2)}));
    });
    BUILTIN("noise", "noise-perlin-full", "nil:1 <- _:1, _:1, _:1, _:3", { gen_noise_full(g, ga, OP_LIBNOISE_PERLIN); });
    BUILTIN("noiseBillow", "noise-billow", "nil:1 <- _:1, _:1, _:1, _:3", { gen_noise_full(g, ga, OP_LIBNOISE_BILLOW); });
    BUILTIN("noiseRidgedMulti", "noise-ridged-multi", "nil:1 <- _:1, _:1, _:3", {
        c.res(0, g.op(OP_LIBNOISE_RIDGED_MULTI, {c.arg(0, 0), c.arg(1, 0), c.arg(2, 0), c.arg(2, 1), c.arg(2, 2)}));
    });
    BUILTIN("voronoiCells", "noise-voronoi", "nil:1 <- _:3", {
        c.res(0, g.op(OP_LIBNOISE_VORONOI, {g.ic(1), c.arg(0, 0), c.arg(0, 1), c.arg(0, 2)}));
    });
}

}  // namespace mm
