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
