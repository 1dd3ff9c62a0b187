/* ORACLE (test infrastructure): elliptic integrals and Jacobi elliptic functions.
 *
 * The reference evaluates these through GSL (opmacros.h:101-125: gsl_sf_ellint_*(..., GSL_PREC_SINGLE) and
 * gsl_sf_elljac_e), a third-party dependency that is not in the reference tree, and holds no golden vector for
 * them: PARITY UNPINNED.  This file computes the same mathematical functions in double, on purpose by other
 * routes than the product's restatement of GSL (mathmap_b200/csrc/runtime/mm_elliptic.h), so that the parity
 * tests compare two independent evaluations:
 *   K, E complete ........ arithmetic-geometric mean (Abramowitz+Stegun 17.6)
 *   RC, RD, RF, RJ ....... Carlson duplication run to a 1e-4 relative spread, fifth-order series
 *   F, E, P, D ........... Legendre forms reduced to the Carlson forms, phi folded into (-pi/2, pi/2]
 *   sn, cn, dn ........... AGM scale followed by the backward phi recursion (Abramowitz+Stegun 16.4)
 * Both were checked against scipy.special to better than 1e-9.  Domain errors give NaN.
 */
#include "mmo_runtime.h"

#define TOL 1e-4

static double max3(double a, double b, double c) { double m = a > b ? a : b; return m > c ? m : c; }

double mmo_ellint_RC(double x, double y) {
    double mu, s;
    int n;
    if (x < 0.0 || y < 0.0 || x + y < 5.0 * DBL_MIN) return NAN;
    for (n = 0; n < 10000; ++n) {
        double lam;
        mu = (x + y + y) / 3.0;
        s = (y + mu) / mu - 2.0;
        if (fabs(s) < TOL) break;
        lam = 2.0 * sqrt(x) * sqrt(y) + y;
        x = (x + lam) / 4.0;
        y = (y + lam) / 4.0;
    }
    return (1.0 + s * s * (0.3 + s * (1.0 / 7.0 + s * (0.375 + s * 9.0 / 22.0)))) / sqrt(mu);
}

double mmo_ellint_RF(double x, double y, double z) {
    double mu, dx, dy, dz, e2, e3;
    int n;
    if (x < 0.0 || y < 0.0 || z < 0.0 || x + y < 5.0 * DBL_MIN || x + z < 5.0 * DBL_MIN || y + z < 5.0 * DBL_MIN) return NAN;
    for (n = 0; n < 10000; ++n) {
        double rx, ry, rz, lam;
        mu = (x + y + z) / 3.0;
        dx = 2.0 - (mu + x) / mu; dy = 2.0 - (mu + y) / mu; dz = 2.0 - (mu + z) / mu;
        if (max3(fabs(dx), fabs(dy), fabs(dz)) < TOL) break;
        rx = sqrt(x); ry = sqrt(y); rz = sqrt(z);
        lam = rx * (ry + rz) + ry * rz;
        x = (x + lam) / 4.0; y = (y + lam) / 4.0; z = (z + lam) / 4.0;
    }
    e2 = dx * dy - dz * dz;
    e3 = dx * dy * dz;
    return (1.0 + (e2 / 24.0 - 0.1 - 3.0 * e3 / 44.0) * e2 + e3 / 14.0) / sqrt(mu);
}

double mmo_ellint_RD(double x, double y, double z) {
    double mu, dx, dy, dz, sigma = 0.0, p4 = 1.0, ea, eb, ec, ed, ef, s1, s2;
    int n;
    if (x < 0.0 || y < 0.0 || x + y < 1e-200 || z < 1e-200) return NAN;
    for (n = 0; n < 10000; ++n) {
        double rx, ry, rz, lam;
        mu = (x + y + 3.0 * z) / 5.0;
        dx = (mu - x) / mu; dy = (mu - y) / mu; dz = (mu - z) / mu;
        if (max3(fabs(dx), fabs(dy), fabs(dz)) < TOL) break;
        rx = sqrt(x); ry = sqrt(y); rz = sqrt(z);
        lam = rx * (ry + rz) + ry * rz;
        sigma += p4 / (rz * (z + lam));
        p4 /= 4.0;
        x = (x + lam) / 4.0; y = (y + lam) / 4.0; z = (z + lam) / 4.0;
    }
    ea = dx * dy; eb = dz * dz; ec = ea - eb; ed = ea - 6.0 * eb; ef = ed + ec + ec;
    s1 = ed * (-3.0 / 14.0 + 0.25 * (9.0 / 22.0) * ed - 1.5 * (3.0 / 26.0) * dz * ef);
    s2 = dz * (ef / 6.0 + dz * (-(9.0 / 22.0) * ec + dz * (3.0 / 26.0) * ea));
    return 3.0 * sigma + p4 * (1.0 + s1 + s2) / (mu * sqrt(mu));
}

double mmo_ellint_RJ(double x, double y, double z, double p) {
    const double lo = cbrt(5.0 * DBL_MIN);
    double mu, dx, dy, dz, dp, sigma = 0.0, p4 = 1.0, ea, eb, ec, e2, e3, s1, s2, s3;
    int n;
    if (x < 0.0 || y < 0.0 || z < 0.0 || x + y < lo || x + z < lo || y + z < lo || p < lo) return NAN;
    for (n = 0; n < 10000; ++n) {
        double rx, ry, rz, lam, alfa, beta;
        mu = (x + y + z + p + p) / 5.0;
        dx = (mu - x) / mu; dy = (mu - y) / mu; dz = (mu - z) / mu; dp = (mu - p) / mu;
        if (fmax(max3(fabs(dx), fabs(dy), fabs(dz)), fabs(dp)) < TOL) break;
        rx = sqrt(x); ry = sqrt(y); rz = sqrt(z);
        lam = rx * (ry + rz) + ry * rz;
        alfa = p * (rx + ry + rz) + rx * ry * rz;
        alfa *= alfa;
        beta = p * (p + lam) * (p + lam);
        sigma += p4 * mmo_ellint_RC(alfa, beta);
        p4 /= 4.0;
        x = (x + lam) / 4.0; y = (y + lam) / 4.0; z = (z + lam) / 4.0; p = (p + lam) / 4.0;
    }
    ea = dx * (dy + dz) + dy * dz; eb = dx * dy * dz; ec = dp * dp;
    e2 = ea - 3.0 * ec; e3 = eb + 2.0 * dp * (ea - ec);
    s1 = 1.0 + e2 * (-3.0 / 14.0 + 0.75 * (3.0 / 22.0) * e2 - 1.5 * (3.0 / 26.0) * e3);
    s2 = eb * (0.5 / 3.0 + dp * (-2.0 * (3.0 / 22.0) + dp * (3.0 / 26.0)));
    s3 = dp * ea * (1.0 / 3.0 - dp * (3.0 / 22.0)) - dp * ec / 3.0;
    return 3.0 * sigma + p4 * (s1 + s2 + s3) / (mu * sqrt(mu));
}

/* K and E by the AGM: K = pi / (2 a_N), E = K (1 - sum 2^(n-1) c_n^2) */
static void agm_KE(double k, double *K, double *E) {
    double a = 1.0, b = sqrt(1.0 - k * k), c = fabs(k), sum = 0.5 * c * c, pw = 0.5;
    int n;
    for (n = 0; n < 64 && fabs(c) > 1e-17 * a; ++n) {
        double an = 0.5 * (a + b);
        c = 0.5 * (a - b);
        b = sqrt(a * b);
        a = an;
        pw *= 2.0;
        sum += pw * c * c;
    }
    *K = M_PI / (2.0 * a);
    *E = *K * (1.0 - sum);
}

double mmo_ellint_Kcomp(double k) {
    double K, E;
    if (k * k >= 1.0) return NAN;
    agm_KE(k, &K, &E);
    return K;
}
double mmo_ellint_Ecomp(double k) {
    double K, E;
    if (k * k >= 1.0) return NAN;
    agm_KE(k, &K, &E);
    return E;
}

static double fold(double *phi) {
    double nc = floor(*phi / M_PI + 0.5);
    *phi -= nc * M_PI;
    return nc;
}

double mmo_ellint_F(double phi, double k) {
    double nc = fold(&phi), s = sin(phi), r = s * mmo_ellint_RF(1.0 - s * s, 1.0 - k * k * s * s, 1.0);
    return nc != 0.0 ? r + 2.0 * nc * mmo_ellint_Kcomp(k) : r;
}
double mmo_ellint_E(double phi, double k) {
    double nc = fold(&phi), s = sin(phi), x = 1.0 - s * s, y = 1.0 - k * k * s * s, r;
    if (x < DBL_EPSILON) return (2.0 * nc + (s >= 0.0 ? 1.0 : -1.0)) * mmo_ellint_Ecomp(k);
    r = s * mmo_ellint_RF(x, y, 1.0) - k * k / 3.0 * s * s * s * mmo_ellint_RD(x, y, 1.0);
    return nc != 0.0 ? r + 2.0 * nc * mmo_ellint_Ecomp(k) : r;
}
double mmo_ellint_P(double phi, double k, double n) {
    double nc = fold(&phi), s = sin(phi), x = 1.0 - s * s, y = 1.0 - k * k * s * s;
    double r = s * mmo_ellint_RF(x, y, 1.0) - n / 3.0 * s * s * s * mmo_ellint_RJ(x, y, 1.0, 1.0 + n * s * s);
    if (nc != 0.0) {
        if (k * k >= 1.0) return NAN;
        r += 2.0 * nc * (mmo_ellint_RF(0.0, 1.0 - k * k, 1.0) - n / 3.0 * mmo_ellint_RJ(0.0, 1.0 - k * k, 1.0, 1.0 + n));
    }
    return r;
}
double mmo_ellint_D(double phi, double k) {
    double nc = fold(&phi), s = sin(phi), r = s * s * s / 3.0 * mmo_ellint_RD(1.0 - s * s, 1.0 - k * k * s * s, 1.0);
    if (nc != 0.0) {
        if (k * k >= 1.0) return NAN;
        r += 2.0 * nc * mmo_ellint_RD(0.0, 1.0 - k * k, 1.0) / 3.0;
    }
    return r;
}

/* Abramowitz+Stegun 16.4: a_n, c_n by the AGM, phi_N = 2^N a_N u, then backwards
 * sin(2 phi_(n-1) - phi_n) = (c_n / a_n) sin(phi_n); sn = sin phi_0, cn = cos phi_0, dn = cos phi_0 / cos(phi_1 - phi_0) */
void mmo_elljac(double u, double m, double *sn, double *cn, double *dn) {
    double a[32], c[32], b, phi, prev = 0.0, twon = 1.0;
    int n = 0, i;
    if (fabs(m) > 1.0) { *sn = *cn = *dn = 0.0; return; }
    if (fabs(m) < 2.0 * DBL_EPSILON) { *sn = sin(u); *cn = cos(u); *dn = 1.0; return; }
    if (fabs(m - 1.0) < 2.0 * DBL_EPSILON) { *sn = tanh(u); *cn = 1.0 / cosh(u); *dn = *cn; return; }
    if (m < 0.0) {  /* negative parameter: A+S 16.10 */
        double mu1 = 1.0 / (1.0 - m), mu = -m * mu1, k1 = sqrt(1.0 - m), s, cc, d;
        mmo_elljac(u * k1, mu, &s, &cc, &d);
        *sn = s / (d * k1);
        *cn = cc / d;
        *dn = 1.0 / d;
        return;
    }
    a[0] = 1.0;
    b = sqrt(1.0 - m);
    c[0] = sqrt(m);
    while (fabs(c[n] / a[n]) > DBL_EPSILON && n < 30) {
        a[n + 1] = 0.5 * (a[n] + b);
        c[n + 1] = 0.5 * (a[n] - b);
        b = sqrt(a[n] * b);
        twon *= 2.0;
        ++n;
    }
    phi = twon * a[n] * u;
    for (i = n; i > 0; --i) {
        prev = phi;
        phi = 0.5 * (asin(c[i] * sin(phi) / a[i]) + phi);
    }
    *sn = sin(phi);
    *cn = cos(phi);
    {
        double t = cos(prev - phi);
        *dn = fabs(t) < 0.1 ? sqrt(1.0 - m * (*sn) * (*sn)) : *cn / t;
    }
}

float *mmo_ell_jac_tuple(float u, float m, mmo_pools *pools) {
    double sn, cn, dn;
    float *r = ALLOC_TUPLE(3);
    mmo_elljac(u, m, &sn, &cn, &dn);
    r[0] = sn; r[1] = cn; r[2] = dn;
    return r;
}
