// Host evaluation of pure scalar ops on constants.  Used by constant folding and
// by the per-frame host replay of frame-constant ("xy-const") statements.
//
// Semantics restated from reference opmacros.h:30-47 and the folder generator
// ops.lisp:319-333: TP_CONST ops convert each constant to the op's declared
// argument type, call the C expression (libm in double on float arguments) and
// narrow to the op's result type; TP_MAX / TP_MAX_FLOAT ops first cast all
// arguments to the widest argument type.
#include "../runtime/mm_elliptic.h"
#include <cmath>
#include <complex>

#include "eval.h"

namespace mm {

typedef __complex__ float cfloat;
static inline cfloat to_c(std::complex<float> z) { cfloat r; __real__ r = z.real(); __imag__ r = z.imag(); return r; }
static inline std::complex<float> from_c(cfloat z) { return {__real__ z, __imag__ z}; }

float const_as_float(const Const &c) {
    switch (c.type) {
    case T_INT: return (float)c.i;
    case T_FLOAT: return c.f;
    case T_COMPLEX: return c.c.real();
    default: return 0.f;
    }
}
int const_as_int(const Const &c) {
    switch (c.type) {
    case T_INT: return c.i;
    case T_FLOAT: return (int)c.f;
    default: return 0;
    }
}
std::complex<float> const_as_complex(const Const &c) {
    switch (c.type) {
    case T_INT: return {(float)c.i, 0.f};
    case T_FLOAT: return {c.f, 0.f};
    case T_COMPLEX: return c.c;
    default: return {0.f, 0.f};
    }
}
static Const mk_i(int v) { Const c; c.type = T_INT; c.i = v; return c; }
static Const mk_f(float v) { Const c; c.type = T_FLOAT; c.f = v; return c; }
static Const mk_c(std::complex<float> v) { Const c; c.type = T_COMPLEX; c.c = v; return c; }

bool const_is_true(const Const &c) {
    if (c.type == T_INT) return c.i != 0;
    if (c.type == T_FLOAT) return c.f != 0.0f;
    return false;
}

bool eval_op(const OpInfo *op, const Const *a, Const *out) {
    for (int i = 0; i < op->nargs; ++i)
        if (a[i].type != T_INT && a[i].type != T_FLOAT && a[i].type != T_COMPLEX) return false;
    if (op->prop != TP_CONST) {
        int mx = T_INT;
        for (int i = 0; i < op->nargs; ++i) mx = std::max(mx, (int)a[i].type);
        if (op->prop == TP_MAX_FLOAT && mx > T_FLOAT) return false;
        if (mx == T_INT) {
            int x = const_as_int(a[0]), y = op->nargs > 1 ? const_as_int(a[1]) : 0;
            switch (op->id) {
            case OP_ADD: *out = mk_i((int)((unsigned)x + (unsigned)y)); return true;
            case OP_SUB: *out = mk_i((int)((unsigned)x - (unsigned)y)); return true;
            case OP_NEG: *out = mk_i((int)(0u - (unsigned)x)); return true;
            case OP_MUL: *out = mk_i((int)((unsigned)x * (unsigned)y)); return true;
            case OP_ABS: *out = mk_i((int)fabs((double)x)); return true;
            case OP_MIN: *out = mk_i(x < y ? x : y); return true;
            case OP_MAX: *out = mk_i(x < y ? y : x); return true;
            default: return false;
            }
        }
        if (mx == T_FLOAT) {
            float x = const_as_float(a[0]), y = op->nargs > 1 ? const_as_float(a[1]) : 0.f;
            switch (op->id) {
            case OP_ADD: *out = mk_f(x + y); return true;
            case OP_SUB: *out = mk_f(x - y); return true;
            case OP_NEG: *out = mk_f(-x); return true;
            case OP_MUL: *out = mk_f(x * y); return true;
            case OP_ABS: *out = mk_f((float)fabs((double)x)); return true;
            case OP_MIN: *out = mk_f(x < y ? x : y); return true;
            case OP_MAX: *out = mk_f(x < y ? y : x); return true;
            default: return false;
            }
        }
        cfloat x = to_c(const_as_complex(a[0])), y = to_c(op->nargs > 1 ? const_as_complex(a[1]) : std::complex<float>());
        switch (op->id) {
        case OP_ADD: *out = mk_c(from_c(x + y)); return true;
        case OP_SUB: *out = mk_c(from_c(x - y)); return true;
        case OP_NEG: *out = mk_c(from_c(-x)); return true;
        case OP_MUL: *out = mk_c(from_c(x * y)); return true;
        default: return false;
        }
    }
    auto F = [&](int i) { return const_as_float(a[i]); };
    auto D = [&](int i) { return (double)const_as_float(a[i]); };
    auto Z = [&](int i) { return to_c(const_as_complex(a[i])); };
    switch (op->id) {
    case OP_NOP: *out = mk_i(0); return true;
    case OP_INT2FLOAT: *out = mk_f((float)const_as_int(a[0])); return true;
    case OP_FLOAT2INT: *out = mk_i((int)F(0)); return true;
    case OP_INT2COMPLEX: *out = mk_c({(float)const_as_int(a[0]), 0.f}); return true;
    case OP_FLOAT2COMPLEX: *out = mk_c({F(0), 0.f}); return true;
    case OP_DIV: *out = mk_f(F(0) / F(1)); return true;
    case OP_MOD: *out = mk_f((float)fmod(D(0), D(1))); return true;
    case OP_SQRT: *out = mk_f((float)sqrt(D(0))); return true;
    case OP_HYPOT: *out = mk_f((float)hypot(D(0), D(1))); return true;
    case OP_SIN: *out = mk_f((float)sin(D(0))); return true;
    case OP_COS: *out = mk_f((float)cos(D(0))); return true;
    case OP_TAN: *out = mk_f((float)tan(D(0))); return true;
    case OP_ASIN: *out = mk_f((float)asin(D(0))); return true;
    case OP_ACOS: *out = mk_f((float)acos(D(0))); return true;
    case OP_ATAN: *out = mk_f((float)atan(D(0))); return true;
    case OP_ATAN2: *out = mk_f((float)atan2(D(0), D(1))); return true;
    case OP_POW: *out = mk_f((float)pow(D(0), D(1))); return true;
    case OP_EXP: *out = mk_f((float)exp(D(0))); return true;
    case OP_LOG: *out = mk_f((float)log(D(0))); return true;
    case OP_SINH: *out = mk_f((float)sinh(D(0))); return true;
    case OP_COSH: *out = mk_f((float)cosh(D(0))); return true;
    case OP_TANH: *out = mk_f((float)tanh(D(0))); return true;
    case OP_ASINH: *out = mk_f((float)asinh(D(0))); return true;
    case OP_ACOSH: *out = mk_f((float)acosh(D(0))); return true;
    case OP_ATANH: *out = mk_f((float)atanh(D(0))); return true;
    case OP_FLOOR: *out = mk_i((int)floor(D(0))); return true;
    case OP_CEIL: *out = mk_i((int)ceil(D(0))); return true;
    case OP_EQ: *out = mk_i(F(0) == F(1)); return true;
    case OP_LESS: *out = mk_i(F(0) < F(1)); return true;
    case OP_LEQ: *out = mk_i(F(0) <= F(1)); return true;
    case OP_NOT: *out = mk_i(!const_as_int(a[0])); return true;
    case OP_COMPLEX: *out = mk_c({F(0), F(1)}); return true;
    case OP_C_REAL: *out = mk_f(__real__ Z(0)); return true;
    case OP_C_IMAG: *out = mk_f(__imag__ Z(0)); return true;
    case OP_C_SQRT: *out = mk_c(from_c(__builtin_csqrtf(Z(0)))); return true;
    case OP_C_SIN: *out = mk_c(from_c(__builtin_csinf(Z(0)))); return true;
    case OP_C_COS: *out = mk_c(from_c(__builtin_ccosf(Z(0)))); return true;
    case OP_C_TAN: *out = mk_c(from_c(__builtin_ctanf(Z(0)))); return true;
    case OP_C_ASIN: *out = mk_c(from_c(__builtin_casinf(Z(0)))); return true;
    case OP_C_ACOS: *out = mk_c(from_c(__builtin_cacosf(Z(0)))); return true;
    case OP_C_ATAN: *out = mk_c(from_c(__builtin_catanf(Z(0)))); return true;
    case OP_C_POW: *out = mk_c(from_c(__builtin_cpowf(Z(0), Z(1)))); return true;
    case OP_C_EXP: *out = mk_c(from_c(__builtin_cexpf(Z(0)))); return true;
    case OP_C_LOG: *out = mk_c(from_c(__builtin_clogf(Z(0)))); return true;
    case OP_C_ARG: *out = mk_f(__builtin_cargf(Z(0))); return true;
    case OP_C_SINH: *out = mk_c(from_c(__builtin_csinhf(Z(0)))); return true;
    case OP_C_COSH: *out = mk_c(from_c(__builtin_ccoshf(Z(0)))); return true;
    case OP_C_TANH: *out = mk_c(from_c(__builtin_ctanhf(Z(0)))); return true;
    case OP_C_ASINH: *out = mk_c(from_c(__builtin_casinhf(Z(0)))); return true;
    case OP_C_ACOSH: *out = mk_c(from_c(__builtin_cacoshf(Z(0)))); return true;
    case OP_C_ATANH: *out = mk_c(from_c(__builtin_catanhf(Z(0)))); return true;
    // GSL elliptic integrals (opmacros.h:101-117), restated in ../runtime/mm_elliptic.h
    case OP_ELL_INT_K_COMP: *out = mk_f((float)mm_ellint_kcomp(D(0))); return true;
    case OP_ELL_INT_E_COMP: *out = mk_f((float)mm_ellint_ecomp(D(0))); return true;
    case OP_ELL_INT_F: *out = mk_f((float)mm_ellint_f(D(0), D(1))); return true;
    case OP_ELL_INT_E: *out = mk_f((float)mm_ellint_e(D(0), D(1))); return true;
    case OP_ELL_INT_P: *out = mk_f((float)mm_ellint_p(D(0), D(1), D(2))); return true;
    case OP_ELL_INT_D: *out = mk_f((float)mm_ellint_d(D(0), D(1))); return true;
    case OP_ELL_INT_RC: *out = mk_f((float)mm_ellint_rc(D(0), D(1))); return true;
    case OP_ELL_INT_RD: *out = mk_f((float)mm_ellint_rd(D(0), D(1), D(2))); return true;
    case OP_ELL_INT_RF: *out = mk_f((float)mm_ellint_rf(D(0), D(1), D(2))); return true;
    case OP_ELL_INT_RJ: *out = mk_f((float)mm_ellint_rj(D(0), D(1), D(2), D(3))); return true;
    default: return false;
    }
}

}  // namespace mm
