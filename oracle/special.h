/* oracle/special.h -- TEST INFRASTRUCTURE ONLY (the CPU oracle).
 *
 * Special functions that the reference obtains from third-party code that is
 * NOT present under /root/reference:
 *
 *   - `special-fun` (git dep https://github.com/pkgw/special-fun, "^0.1", rev
 *     unpinned, Cargo.toml:17; a binding to Cephes): `besseli`, `besselj`,
 *     `bessely` with real order.  Call sites: src/heyvaerts.rs:331-332,
 *     335-336, 359-360, 363, 437-438, 440-441.
 *   - GSL `gsl_sf_hyperg_2F1(1/2, -k/2; 3/2; 1)` (src/gsl.rs:261-263, called
 *     from src/pitchy_pl.rs:98 and src/pitchy_kappa.rs:93).
 *   - GSL QAGIU for the Juettner normalisation (src/thermal_juettner.rs:56-64),
 *     replaced here by the closed form T K_2(1/T) that the integral equals.
 *
 * Cephes aims at full double accuracy for these functions, so the oracle
 * restates the *mathematical* functions with textbook algorithms (ascending
 * series; Temme's series for Y_mu) and tests/ pins them against scipy.special.
 * Parity for the g >= 10 J/Y branch of Heyvaerts is "unpinned" in the sense of
 * no reference test reaches that branch.
 */
#ifndef RIMPHONY_ORACLE_SPECIAL_H
#define RIMPHONY_ORACLE_SPECIAL_H

#include <float.h>
#include <math.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846264338327950288
#endif

/* Taylor coefficients of 1/Gamma(1+mu) about mu = 0 (generated with mpmath). */
static const double orc_rgamma1p_coef[28] = {
    1.0,
    0.57721566490153286061,
    -0.65587807152025388108,
    -0.042002635034095235529,
    0.1665386113822914895,
    -0.042197734555544336748,
    -0.0096219715278769735621,
    0.0072189432466630995424,
    -0.0011651675918590651121,
    -0.00021524167411495097282,
    0.00012805028238811618615,
    -0.000020134854780788238656,
    -1.2504934821426706573e-6,
    1.1330272319816958824e-6,
    -2.0563384169776071035e-7,
    6.1160951044814158179e-9,
    5.0020076444692229301e-9,
    -1.1812745704870201446e-9,
    1.0434267116911005105e-10,
    7.782263439905071254e-12,
    -3.6968056186422057082e-12,
    5.100370287454475979e-13,
    -2.0583260535665067832e-14,
    -5.3481225394230179824e-15,
    1.2267786282382607902e-15,
    -1.1812593016974587695e-16,
    1.1866922547516003326e-18,
    1.4123806553180317816e-18
};

/* For |mu| <= 1/2:
 *   gam2  = (1/Gamma(1-mu) + 1/Gamma(1+mu)) / 2
 *   gam1  = (1/Gamma(1-mu) - 1/Gamma(1+mu)) / (2 mu)
 *   gampl = 1/Gamma(1+mu),  gammi = 1/Gamma(1-mu)
 */
static inline void orc_gamma_pair(double mu, double *gam1, double *gam2,
                                  double *gampl, double *gammi)
{
    const double m2 = mu * mu;
    double even = 0.0, odd = 0.0;
    int j;
    for (j = 26; j >= 0; j -= 2)
        even = even * m2 + orc_rgamma1p_coef[j];
    for (j = 27; j >= 1; j -= 2)
        odd = odd * m2 + orc_rgamma1p_coef[j];
    *gam2 = even;
    *gam1 = -odd;
    *gampl = even + mu * odd;
    *gammi = even - mu * odd;
}

/* 1/Gamma(1 + nu) for nu in (-1, ~20), through the |mu| <= 1/2 series and the
 * recurrence Gamma(z+1) = z Gamma(z). */
static inline double orc_rgamma1p(double nu)
{
    double m = floor(nu + 0.5);
    double mu = nu - m;
    double g1, g2, gp, gm, r;
    int i, im = (int)m;
    orc_gamma_pair(mu, &g1, &g2, &gp, &gm);
    r = gp; /* 1/Gamma(1+mu) */
    if (im >= 0) {
        for (i = 1; i <= im; i++)
            r /= (mu + i);
    } else {
        /* im == -1: 1/Gamma(mu) = mu / Gamma(1+mu) */
        r *= mu;
    }
    return r;
}

/* J_nu(x) by its ascending series; intended for nu > -1 and 0 < x <~ 6. */
static inline double orc_bessel_j_series(double nu, double x)
{
    const double q = -0.25 * x * x;
    double term = 1.0, sum = 1.0;
    int k;
    for (k = 1; k < 200; k++) {
        term *= q / (k * (k + nu));
        sum += term;
        if (fabs(term) < 1e-17 * fabs(sum))
            break;
    }
    return pow(0.5 * x, nu) * orc_rgamma1p(nu) * sum;
}

/* Temme's series: Y_mu(x) and Y_{mu+1}(x) for |mu| <= 1/2, 0 < x <~ 4. */
static inline void orc_bessel_y_temme(double mu, double x, double *y_mu, double *y_mu1)
{
    const double eps = DBL_EPSILON;
    double gam1, gam2, gampl, gammi;
    const double x2 = 0.5 * x;
    const double pimu = M_PI * mu;
    const double fact = (fabs(pimu) < eps) ? 1.0 : pimu / sin(pimu);
    double d = -log(x2);
    double e = mu * d;
    const double fact2 = (fabs(e) < eps) ? 1.0 : sinh(e) / e;
    double ff, p, q, r, c, sum, sum1, pimu2, fact3;
    int i;

    orc_gamma_pair(mu, &gam1, &gam2, &gampl, &gammi);
    ff = 2.0 / M_PI * fact * (gam1 * cosh(e) + gam2 * fact2 * d);
    e = exp(e);
    p = e / (gampl * M_PI);
    q = 1.0 / (e * M_PI * gammi);
    pimu2 = 0.5 * pimu;
    fact3 = (fabs(pimu2) < eps) ? 1.0 : sin(pimu2) / pimu2;
    r = M_PI * pimu2 * fact3 * fact3;
    c = 1.0;
    d = -x2 * x2;
    sum = ff + r * q;
    sum1 = p;
    for (i = 1; i < 500; i++) {
        double del, del1;
        ff = (i * ff + p + q) / (i * (double)i - mu * mu);
        c *= d / i;
        p /= (i - mu);
        q /= (i + mu);
        del = c * (ff + r * q);
        sum += del;
        del1 = c * p - i * del;
        sum1 += del1;
        if (fabs(del) < (1.0 + fabs(sum)) * eps * 0.1)
            break;
    }
    *y_mu = -sum;
    *y_mu1 = -sum1 * (2.0 / x);
}

/* J_nu(x) and Y_nu(x) for real order nu in (-1, ~8) and 0 < x <~ 4: the only
 * region in which heyvaerts.rs reaches its besselj/bessely branch (sigma < 3,
 * x < sigma). */
static inline void orc_bessel_jy(double nu, double x, double *j, double *y)
{
    if (!(x > 0.0) || !(nu > -1.0)) {
        *j = NAN;
        *y = NAN;
        return;
    }

    if (nu >= -0.5) {
        double m = floor(nu + 0.5);
        double mu = nu - m;
        double ya, yb;
        int i, im = (int)m;
        orc_bessel_y_temme(mu, x, &ya, &yb);
        /* upward recurrence (stable for Y) */
        for (i = 1; i <= im; i++) {
            double yn = 2.0 * (mu + i) / x * yb - ya;
            ya = yb;
            yb = yn;
        }
        *y = ya;
        *j = orc_bessel_j_series(nu, x);
    } else {
        /* reflection: a = -nu in (1/2, 1) */
        const double a = -nu;
        double ya, yb, ja;
        orc_bessel_y_temme(a - 1.0, x, &yb, &ya); /* Y_{a-1}, Y_a */
        (void)yb;
        ja = orc_bessel_j_series(a, x);
        *y = sin(a * M_PI) * ja + cos(a * M_PI) * ya;
        *j = orc_bessel_j_series(nu, x);
    }
}

/* Modified Bessel I_nu(x), real (possibly negative non-integer) order, by the
 * ascending series; the reference calls it for nu = +-1/3, +-2/3 and
 * 0 < x < 10 (heyvaerts.rs:330-333, 358-361, 434-438). */
static inline double orc_bessel_i_series(double nu, double x)
{
    const double q = 0.25 * x * x;
    double term = 1.0, sum = 1.0;
    int k;
    for (k = 1; k < 500; k++) {
        term *= q / (k * (k + nu));
        sum += term;
        if (term < 1e-17 * sum)
            break;
    }
    return pow(0.5 * x, nu) * orc_rgamma1p(nu) * sum;
}

/* K_2(z) from int_0^inf exp(-z cosh t) cosh(2t) dt with the trapezoid rule
 * (exponentially convergent). */
static inline double orc_bessel_k2(double z)
{
    const double h = 0.0625;
    double sum = 0.5 * exp(-z);
    int i;
    for (i = 1; i < 4000; i++) {
        double t = i * h;
        double arg = z * cosh(t);
        double v;
        if (arg > 745.0)
            break;
        v = exp(-arg) * cosh(2.0 * t);
        sum += v;
    }
    return sum * h;
}

/* 2F1(1/2, -k/2; 3/2; 1) by Gauss's theorem. */
static inline double orc_pitch_angle_integral(double k)
{
    return 0.5 * sqrt(M_PI) * exp(lgamma(1.0 + 0.5 * k) - lgamma(1.5 + 0.5 * k));
}

#endif
