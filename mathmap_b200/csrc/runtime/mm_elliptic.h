// Elliptic integrals and Jacobi elliptic functions, shared by the host (constant
// folding and the frame-constant replay) and the device runtime.
//
// The reference calls GSL for these (opmacros.h:101-125: gsl_sf_ellint_* with
// GSL_PREC_SINGLE, gsl_sf_elljac_e); GSL is a third-party dependency that is not
// part of the reference tree (debian/control: libgsl0-dev).  What follows
// restates GSL's published algorithms: Carlson's duplication theorems for
// RC/RD/RF/RJ with the single-precision tolerance 0.03, Legendre forms reduced to
// them, Abramowitz+Stegun 17.3.33/17.3.36 near k^2 = 1, and the descending Landen
// (arithmetic-geometric mean) recursion for sn/cn/dn.  Everything is evaluated in
// double and narrowed to float by the caller, like the reference's C.  A domain
// error yields NaN (the reference turns GSL's error handler off, mathmap.c:353).
// PARITY UNPINNED: the reference holds no golden vector for these ops.
#pragma once

#ifdef __CUDACC__
#define MM_ELL_FN __host__ __device__ inline
#else
#include <cmath>
#define MM_ELL_FN inline
#endif

#define MM_ELL_DBL_EPSILON 2.2204460492503131e-16
#define MM_ELL_SQRT_DBL_EPSILON 1.4901161193847656e-08
#define MM_ELL_DBL_MIN 2.2250738585072014e-308
#define MM_ELL_DBL_MAX 1.7976931348623157e+308
#define MM_ELL_ERRTOL 0.03  // GSL_PREC_SINGLE
#define MM_ELL_NMAX 10000
#define MM_ELL_PI 3.14159265358979323846

MM_ELL_FN double mm_ell_nan() { return nan(""); }
MM_ELL_FN double mm_ell_max2(double a, double b) { return a > b ? a : b; }
MM_ELL_FN double mm_ell_max3(double a, double b, double c) { return mm_ell_max2(mm_ell_max2(a, b), c); }
MM_ELL_FN double mm_ell_min2(double a, double b) { return a < b ? a : b; }

MM_ELL_FN double mm_ellint_rc(double x, double y) {
    const double lolim = 5.0 * MM_ELL_DBL_MIN, uplim = 0.2 * MM_ELL_DBL_MAX;
    if (x < 0.0 || y < 0.0 || x + y < lolim) return mm_ell_nan();
    if (!(mm_ell_max2(x, y) < uplim)) return mm_ell_nan();
    const double c1 = 1.0 / 7.0, c2 = 9.0 / 22.0;
    double xn = x, yn = y, mu, sn;
    for (int n = 0;; ++n) {
        mu = (xn + yn + yn) / 3.0;
        sn = (yn + mu) / mu - 2.0;
        if (fabs(sn) < MM_ELL_ERRTOL) break;
        const double lamda = 2.0 * sqrt(xn) * sqrt(yn) + yn;
        xn = (xn + lamda) * 0.25;
        yn = (yn + lamda) * 0.25;
        if (n + 1 == MM_ELL_NMAX) return mm_ell_nan();
    }
    const double s = sn * sn * (0.3 + sn * (c1 + sn * (0.375 + sn * c2)));
    return (1.0 + s) / sqrt(mu);
}

MM_ELL_FN double mm_ellint_rd(double x, double y, double z) {
    const double lolim = 2.0 / pow(MM_ELL_DBL_MAX, 2.0 / 3.0), uplim = pow(0.1 * MM_ELL_ERRTOL / MM_ELL_DBL_MIN, 2.0 / 3.0);
    if (mm_ell_min2(x, y) < 0.0 || mm_ell_min2(x + y, z) < lolim) return mm_ell_nan();
    if (!(mm_ell_max3(x, y, z) < uplim)) return mm_ell_nan();
    const double c1 = 3.0 / 14.0, c2 = 1.0 / 6.0, c3 = 9.0 / 22.0, c4 = 3.0 / 26.0;
    double xn = x, yn = y, zn = z, sigma = 0.0, power4 = 1.0, mu, xndev, yndev, zndev;
    for (int n = 0;; ++n) {
        mu = (xn + yn + 3.0 * zn) * 0.2;
        xndev = (mu - xn) / mu;
        yndev = (mu - yn) / mu;
        zndev = (mu - zn) / mu;
        const double epslon = mm_ell_max3(fabs(xndev), fabs(yndev), fabs(zndev));
        if (epslon < MM_ELL_ERRTOL) break;
        const double xnroot = sqrt(xn), ynroot = sqrt(yn), znroot = sqrt(zn);
        const double lamda = xnroot * (ynroot + znroot) + ynroot * znroot;
        sigma += power4 / (znroot * (zn + lamda));
        power4 *= 0.25;
        xn = (xn + lamda) * 0.25;
        yn = (yn + lamda) * 0.25;
        zn = (zn + lamda) * 0.25;
        if (n + 1 == MM_ELL_NMAX) return mm_ell_nan();
    }
    const double ea = xndev * yndev, eb = zndev * zndev, ec = ea - eb, ed = ea - 6.0 * eb, ef = ed + ec + ec;
    const double s1 = ed * (-c1 + 0.25 * c3 * ed - 1.5 * c4 * zndev * ef);
    const double s2 = zndev * (c2 * ef + zndev * (-c3 * ec + zndev * c4 * ea));
    return 3.0 * sigma + power4 * (1.0 + s1 + s2) / (mu * sqrt(mu));
}

MM_ELL_FN double mm_ellint_rf(double x, double y, double z) {
    const double lolim = 5.0 * MM_ELL_DBL_MIN, uplim = 0.2 * MM_ELL_DBL_MAX;
    if (x < 0.0 || y < 0.0 || z < 0.0) return mm_ell_nan();
    if (x + y < lolim || x + z < lolim || y + z < lolim) return mm_ell_nan();
    if (!(mm_ell_max3(x, y, z) < uplim)) return mm_ell_nan();
    const double c1 = 1.0 / 24.0, c2 = 3.0 / 44.0, c3 = 1.0 / 14.0;
    double xn = x, yn = y, zn = z, mu, xndev, yndev, zndev;
    for (int n = 0;; ++n) {
        mu = (xn + yn + zn) / 3.0;
        xndev = 2.0 - (mu + xn) / mu;
        yndev = 2.0 - (mu + yn) / mu;
        zndev = 2.0 - (mu + zn) / mu;
        const double epslon = mm_ell_max3(fabs(xndev), fabs(yndev), fabs(zndev));
        if (epslon < MM_ELL_ERRTOL) break;
        const double xnroot = sqrt(xn), ynroot = sqrt(yn), znroot = sqrt(zn);
        const double lamda = xnroot * (ynroot + znroot) + ynroot * znroot;
        xn = (xn + lamda) * 0.25;
        yn = (yn + lamda) * 0.25;
        zn = (zn + lamda) * 0.25;
        if (n + 1 == MM_ELL_NMAX) return mm_ell_nan();
    }
    const double e2 = xndev * yndev - zndev * zndev, e3 = xndev * yndev * zndev;
    const double s = 1.0 + (c1 * e2 - 0.1 - c2 * e3) * e2 + c3 * e3;
    return s / sqrt(mu);
}

MM_ELL_FN double mm_ellint_rj(double x, double y, double z, double p) {
    const double lolim = pow(5.0 * MM_ELL_DBL_MIN, 1.0 / 3.0), uplim = 0.3 * pow(0.2 * MM_ELL_DBL_MAX, 1.0 / 3.0);
    if (x < 0.0 || y < 0.0 || z < 0.0) return mm_ell_nan();
    if (x + y < lolim || x + z < lolim || y + z < lolim || p < lolim) return mm_ell_nan();
    if (!(mm_ell_max2(mm_ell_max3(x, y, z), p) < uplim)) return mm_ell_nan();
    const double c1 = 3.0 / 14.0, c2 = 1.0 / 3.0, c3 = 3.0 / 22.0, c4 = 3.0 / 26.0;
    double xn = x, yn = y, zn = z, pn = p, sigma = 0.0, power4 = 1.0, mu, xndev, yndev, zndev, pndev;
    for (int n = 0;; ++n) {
        mu = (xn + yn + zn + pn + pn) * 0.2;
        xndev = (mu - xn) / mu;
        yndev = (mu - yn) / mu;
        zndev = (mu - zn) / mu;
        pndev = (mu - pn) / mu;
        const double epslon = mm_ell_max2(mm_ell_max3(fabs(xndev), fabs(yndev), fabs(zndev)), fabs(pndev));
        if (epslon < MM_ELL_ERRTOL) break;
        const double xnroot = sqrt(xn), ynroot = sqrt(yn), znroot = sqrt(zn);
        const double lamda = xnroot * (ynroot + znroot) + ynroot * znroot;
        double alfa = pn * (xnroot + ynroot + znroot) + xnroot * ynroot * znroot;
        alfa = alfa * alfa;
        const double beta = pn * (pn + lamda) * (pn + lamda);
        sigma += power4 * mm_ellint_rc(alfa, beta);
        power4 *= 0.25;
        xn = (xn + lamda) * 0.25;
        yn = (yn + lamda) * 0.25;
        zn = (zn + lamda) * 0.25;
        pn = (pn + lamda) * 0.25;
        if (n + 1 == MM_ELL_NMAX) return mm_ell_nan();
    }
    const double ea = xndev * (yndev + zndev) + yndev * zndev, eb = xndev * yndev * zndev, ec = pndev * pndev;
    const double e2 = ea - 3.0 * ec, e3 = eb + 2.0 * pndev * (ea - ec);
    const double s1 = 1.0 + e2 * (-c1 + 0.75 * c3 * e2 - 1.5 * c4 * e3);
    const double s2 = eb * (0.5 * c2 + pndev * (-c3 - c3 + pndev * c4));
    const double s3 = pndev * ea * (c2 - pndev * c3) - c2 * pndev * ec;
    return 3.0 * sigma + power4 * (s1 + s2 + s3) / (mu * sqrt(mu));
}

// complete integrals
MM_ELL_FN double mm_ellint_kcomp(double k) {
    if (k * k >= 1.0) return mm_ell_nan();
    const double y = 1.0 - k * k;
    if (k * k >= 1.0 - MM_ELL_SQRT_DBL_EPSILON) {  // Abramowitz+Stegun 17.3.33
        const double ta = 1.38629436112 + y * (0.09666344259 + y * 0.03590092383);
        const double tb = -log(y) * (0.5 + y * (0.12498593597 + y * 0.06880248576));
        return ta + tb;
    }
    return mm_ellint_rf(0.0, y, 1.0);
}

MM_ELL_FN double mm_ellint_ecomp(double k) {
    if (k * k >= 1.0) return mm_ell_nan();
    const double y = 1.0 - k * k;
    if (k * k >= 1.0 - MM_ELL_SQRT_DBL_EPSILON) {  // Abramowitz+Stegun 17.3.36
        const double ta = 1.0 + y * (0.44325141463 + y * (0.06260601220 + 0.04757383546 * y));
        const double tb = -y * log(y) * (0.24998368310 + y * (0.09200180037 + 0.04069697526 * y));
        return ta + tb;
    }
    return mm_ellint_rf(0.0, y, 1.0) - k * k / 3.0 * mm_ellint_rd(0.0, y, 1.0);
}

MM_ELL_FN double mm_ellint_pcomp(double k, double n) {
    if (k * k >= 1.0) return mm_ell_nan();
    const double y = 1.0 - k * k;
    return mm_ellint_rf(0.0, y, 1.0) - (n / 3.0) * mm_ellint_rj(0.0, y, 1.0, 1.0 + n);
}

MM_ELL_FN double mm_ellint_dcomp(double k) {
    if (k * k >= 1.0) return mm_ell_nan();
    return mm_ellint_rd(0.0, 1.0 - k * k, 1.0) / 3.0;
}

// incomplete (Legendre) integrals: phi is reduced to (-pi/2, pi/2] by whole periods nc
MM_ELL_FN double mm_ellint_f(double phi, double k) {
    const double nc = floor(phi / MM_ELL_PI + 0.5);
    phi = phi - nc * MM_ELL_PI;
    const double sin_phi = sin(phi), sin2_phi = sin_phi * sin_phi;
    const double x = 1.0 - sin2_phi, y = 1.0 - k * k * sin2_phi;
    double r = sin_phi * mm_ellint_rf(x, y, 1.0);
    if (nc != 0.0) r += 2.0 * nc * mm_ellint_kcomp(k);
    return r;
}

MM_ELL_FN double mm_ellint_e(double phi, double k) {
    const double nc = floor(phi / MM_ELL_PI + 0.5);
    phi = phi - nc * MM_ELL_PI;
    const double sin_phi = sin(phi), sin2_phi = sin_phi * sin_phi;
    const double x = 1.0 - sin2_phi, y = 1.0 - k * k * sin2_phi;
    if (x < MM_ELL_DBL_EPSILON) {
        const double re = mm_ellint_ecomp(k);
        return 2.0 * nc * re + (sin_phi >= 0.0 ? 1.0 : -1.0) * re;
    }
    const double sin3_phi = sin2_phi * sin_phi;
    double r = sin_phi * mm_ellint_rf(x, y, 1.0) - k * k / 3.0 * sin3_phi * mm_ellint_rd(x, y, 1.0);
    if (nc != 0.0) r += 2.0 * nc * mm_ellint_ecomp(k);
    return r;
}

MM_ELL_FN double mm_ellint_p(double phi, double k, double n) {
    const double nc = floor(phi / MM_ELL_PI + 0.5);
    phi = phi - nc * MM_ELL_PI;
    const double sin_phi = sin(phi), sin2_phi = sin_phi * sin_phi, sin3_phi = sin2_phi * sin_phi;
    const double x = 1.0 - sin2_phi, y = 1.0 - k * k * sin2_phi;
    double r = sin_phi * mm_ellint_rf(x, y, 1.0) - n / 3.0 * sin3_phi * mm_ellint_rj(x, y, 1.0, 1.0 + n * sin2_phi);
    if (nc != 0.0) r += 2.0 * nc * mm_ellint_pcomp(k, n);
    return r;
}

// the third argument of the reference's ell_int_D is ignored by GSL (and absent from GSL >= 2)
MM_ELL_FN double mm_ellint_d(double phi, double k) {
    const double nc = floor(phi / MM_ELL_PI + 0.5);
    phi = phi - nc * MM_ELL_PI;
    const double sin_phi = sin(phi), sin2_phi = sin_phi * sin_phi, sin3_phi = sin2_phi * sin_phi;
    const double x = 1.0 - sin2_phi, y = 1.0 - k * k * sin2_phi;
    double r = sin3_phi / 3.0 * mm_ellint_rd(x, y, 1.0);
    if (nc != 0.0) r += 2.0 * nc * mm_ellint_dcomp(k);
    return r;
}

// sn, cn, dn of (u | m) by the descending Landen transformation
MM_ELL_FN void mm_elljac(double u, double m, double *sn, double *cn, double *dn) {
    if (fabs(m) > 1.0) { *sn = *cn = *dn = 0.0; return; }
    if (fabs(m) < 2.0 * MM_ELL_DBL_EPSILON) { *sn = sin(u); *cn = cos(u); *dn = 1.0; return; }
    if (fabs(m - 1.0) < 2.0 * MM_ELL_DBL_EPSILON) { *sn = tanh(u); *cn = 1.0 / cosh(u); *dn = *cn; return; }
    const int N = 16;
    double mu[16], nu[16], c[16], d[16];
    int n = 0;
    mu[0] = 1.0;
    nu[0] = sqrt(1.0 - m);
    while (fabs(mu[n] - nu[n]) > 4.0 * MM_ELL_DBL_EPSILON * fabs(mu[n] + nu[n])) {
        mu[n + 1] = 0.5 * (mu[n] + nu[n]);
        nu[n + 1] = sqrt(mu[n] * nu[n]);
        ++n;
        if (n >= N - 1) break;
    }
    const double sin_umu = sin(u * mu[n]), cos_umu = cos(u * mu[n]);
    if (fabs(sin_umu) < fabs(cos_umu)) {
        const double t = sin_umu / cos_umu;
        c[n] = mu[n] * t;
        d[n] = 1.0;
        while (n > 0) {
            --n;
            c[n] = d[n + 1] * c[n + 1];
            const double r = (c[n + 1] * c[n + 1]) / mu[n + 1];
            d[n] = (r + nu[n]) / (r + mu[n]);
        }
        *dn = sqrt(1.0 - m) / d[n];
        *cn = (*dn) * (cos_umu >= 0.0 ? 1.0 : -1.0) / hypot(1.0, c[n]);
        *sn = (*cn) * c[n] / sqrt(1.0 - m);
    } else {
        const double t = cos_umu / sin_umu;
        c[n] = mu[n] * t;
        d[n] = 1.0;
        while (n > 0) {
            --n;
            c[n] = d[n + 1] * c[n + 1];
            const double r = (c[n + 1] * c[n + 1]) / mu[n + 1];
            d[n] = (r + nu[n]) / (r + mu[n]);
        }
        *dn = d[n];
        *sn = (sin_umu >= 0.0 ? 1.0 : -1.0) / hypot(1.0, c[n]);
        *cn = c[n] * (*sn);
    }
}
