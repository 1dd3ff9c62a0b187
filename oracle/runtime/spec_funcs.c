/* ORACLE (test infrastructure): special functions.
 *
 * mmo_cgamma restates reference builtins/spec_func.c:35-63 (Luke's 7-term
 * approximation, complex double arithmetic, reflection for Re z < 0).
 *
 * mmo_gamma / mmo_beta stand in for GSL's gsl_sf_gamma / gsl_sf_beta (GSL is a
 * third-party dependency absent from this environment; opmacros.h:43, ops.lisp:148-149).
 * They use libm's tgamma/lgamma.  PARITY UNPINNED: no reference test exercises
 * gamma/beta (tests/run_tests.sh has the Gamma Correction lines commented out).
 */
#include "mmo_runtime.h"

float _Complex mmo_cgamma(float _Complex z) {
    static const double coeff[7] = {41.624436916439068, -51.224241022374774, 11.338755813488977, -0.747732687772388,
                                    0.008782877493061,  -1.899030264e-6,     1.946335e-9};
    double _Complex s, H, w;
    int n;
    if (creal(z) < 0.0) {
        double _Complex denom = 1.0;
        int flr = -floor(creal(z));
        for (n = 0; n < flr; ++n) denom = denom * (z + n);
        return mmo_cgamma(z + flr) / denom;
    }
    w = z - 1.0;
    s = coeff[0];
    H = 1.0;
    for (n = 1; n < 7; n++) {
        H *= (w + 1 - n) / (w + n);
        s += coeff[n] * H;
    }
    return (2.506628274631 * cexp(-w - 5.5) * cpow(w + 5.5, w + 0.5) * s);
}

double mmo_gamma(double x) { return tgamma(x); }
double mmo_beta(double a, double b) { return exp(lgamma(a) + lgamma(b) - lgamma(a + b)); }

/* matrix "division": stands in for gsl_linalg_HH_solve (opmacros.h:66-86); Cramer's rule in double, 0 when singular.  PARITY UNPINNED. */
float *mmo_solve_linear_2(const float *m, const float *v, mmo_pools *pools) {
    float *r = ALLOC_TUPLE(2);
    double a = m[0], b = m[1], c = m[2], d = m[3], det = a * d - b * c;
    if (det == 0.0) { r[0] = r[1] = 0.f; return r; }
    r[0] = (float)(((double)v[0] * d - b * (double)v[1]) / det);
    r[1] = (float)((a * (double)v[1] - (double)v[0] * c) / det);
    return r;
}
float *mmo_solve_linear_3(const float *m, const float *v, mmo_pools *pools) {
    float *r = ALLOC_TUPLE(3);
    double a[9], b[3], det, d0, d1, d2;
    int i;
    for (i = 0; i < 9; ++i) a[i] = m[i];
    for (i = 0; i < 3; ++i) b[i] = v[i];
    det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    if (det == 0.0) { r[0] = r[1] = r[2] = 0.f; return r; }
    d0 = b[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (b[1] * a[8] - a[5] * b[2]) + a[2] * (b[1] * a[7] - a[4] * b[2]);
    d1 = a[0] * (b[1] * a[8] - a[5] * b[2]) - b[0] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * b[2] - b[1] * a[6]);
    d2 = a[0] * (a[4] * b[2] - b[1] * a[7]) - a[1] * (a[3] * b[2] - b[1] * a[6]) + b[0] * (a[3] * a[7] - a[4] * a[6]);
    r[0] = (float)(d0 / det); r[1] = (float)(d1 / det); r[2] = (float)(d2 / det);
    return r;
}
