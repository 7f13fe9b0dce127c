// rb_bessel.cuh -- the Leung, Gammie & Noble (2011, Appendix B) fast Bessel
// evaluator J_n(x), J_n'(x) as a device function.
//
// Replaces the reference's one native component, leung-bessel/src/bessel.c
// (pkgw_bessel_j :318-376, pkgw_bessel_dj :379-405, BesselJ_Meissel_First
// :94-149, BesselJ_Meissel_Second :57-88, BesselJ_Debye_Eps_Exp :159-213,
// exp_factor :22-51).  Parity with the reference means reproducing its
// approximations, including the (n+1)*Gamma(n) substitution of bessel.c:123-124
// that biases Meissel's first expansion by -1/(n+1); this is NOT a
// general-purpose accurate Bessel function.
//
// B200-first restructuring: in the Symphony gamma integral the 31 lanes of a
// warp evaluate J_n and J_{n+1} at 31 different arguments but the SAME order,
// so everything that depends on the order alone (ln n, the region thresholds
// as bounds on (n-x)/n instead of a log10 per call, lgamma(n) by Stirling's
// series, the 1/n series) is hoisted into a warp-uniform `LeungOrder` that is
// prepared once per gamma integral.
#pragma once

#include "rb_core.cuh"

namespace rb {

constexpr double kLn10 = 2.302585092994045684017991454684364;
constexpr double kLog10e = 0.434294481903251827651128918916605;

// bessel.c:313-316
constexpr double kMinusEtaA = 0.174857;
constexpr double kMinusEtaB = 0.295966;
constexpr double kPlusEtaA = 0.151550;
constexpr double kPlusEtaB = 0.438914;
constexpr double kEtaSlope = -0.6666666;
constexpr double kNJn = 30.0;

// 10^intercept, so that thresholds on eta = log10(eps) become thresholds on eps
constexpr double kTenMinusA = 1.4957430721771774;  // 10^0.174857
constexpr double kTenMinusB = 1.9768148733821513;  // 10^0.295966
constexpr double kTenPlusA = 1.4175879078875209;   // 10^0.151550
constexpr double kTenPlusB = 2.747350062605397;   // 10^0.438914

// bessel.c:171-174: At[m] = sin(pi (m+1)/3) 6^((m+1)/3) Gamma((m+1)/3),
// evaluated in double exactly as the reference's lazy initialiser does.
constexpr double kAt0 = 4.215772156113926;
constexpr double kAt1 = 3.8721718521798665;
constexpr double kAt3 = -8.431544312227844;
constexpr double kAt4 = -15.488687408719468;
constexpr double kAt6 = 67.45235449782282;
constexpr double kAt7 = 154.88687408719463;
constexpr double kAt9 = -944.3329629695189;
constexpr double kAt10 = -2478.189985395116;
constexpr double kAt12 = 18886.65925939032;
constexpr double kAt13 = 54520.17967869256;
constexpr double kAt15 = -491053.1407441488;

enum { kOrderInteger = 0, kOrderLeung = 1, kOrderInvalid = 2 };

// Everything about J_n(.) that depends on the order only.
struct LeungOrder {
    double n;
    double ninv;
    double lo_minus, hi_minus; // Debye / blend / Meissel-1 boundaries on (n-x)/n
    double lo_plus, hi_plus;   // same for (x-n)/x, x > n
    double eta_lo_minus, eta_lo_plus;
    double c_std;      // -Vsum2 - lgamma(n)                       (bessel.c:145)
    double c_stirling; // 0.5 ln(n/2pi) + loggamma_exp - Vsum2     (bessel.c:135)
    int kind;
    int nint;
};

// ln Gamma(n) for n >= 30 by Stirling's series (absolute error < 1e-16 + 1 ulp).
RB_FN double lgamma_stirling(double n, double ln_n, double ninv)
{
    const double t = ninv * ninv;
    const double series =
        ninv * (1.0 / 12.0 +
                t * (-1.0 / 360.0 + t * (1.0 / 1260.0 + t * (-1.0 / 1680.0 + t * (1.0 / 1188.0)))));
    return (n - 0.5) * ln_n - n + 0.918938533204672741780329736405618 + series;
}

RB_FN_NOINLINE void leung_prepare(double n, LeungOrder &o)
{
    o.n = n;
    if (!(n >= 0.0)) {
        o.kind = kOrderInvalid;
        return;
    }
    if (n < kNJn) {
        const int ni = (int)n;
        o.nint = ni;
        o.kind = ((double)ni == n) ? kOrderInteger : kOrderInvalid;
        return;
    }
    o.kind = kOrderLeung;
    o.nint = 0;
    const double ninv = 1.0 / n;
    const double ln_n = rb_log(n);
    o.ninv = ninv;

    // eta thresholds -0.6666666 log10 n + intercept, as bounds on eps itself
    const double pw = rb_exp(kEtaSlope * ln_n); // n^-0.6666666
    o.lo_minus = pw * kTenMinusA;
    o.hi_minus = pw * kTenMinusB;
    o.lo_plus = pw * kTenPlusA;
    o.hi_plus = pw * kTenPlusB;
    const double logn = ln_n * kLog10e;
    o.eta_lo_minus = kEtaSlope * logn + kMinusEtaA;
    o.eta_lo_plus = kEtaSlope * logn + kPlusEtaA;

    const double t2 = ninv * ninv;
    const double vsum2 = -(ninv * (420.0 + (-14.0 + (-4.0 + 3.0 * t2) * t2) * t2)) / 5040.0;
    const double t3 = t2 * t2;
    const double loggamma_exp = (ninv * (-420.0 + 14.0 * t2 - 4.0 * t3 + 3.0 * t3 * t2)) / 5040.0;
    o.c_std = -vsum2 - lgamma_stirling(n, ln_n, ninv);
    o.c_stirling = 0.5 * (ln_n - 1.837877066409345483560659472811235) + loggamma_exp - vsum2;
}

// Polynomial coefficients of the expansions below live in the constant bank and are used as
// c[bank][offset] operands of DFMA/DMUL.  As literals the compiler materialises each 64-bit value
// with two UMOVs per use: ~10 % of all instructions the product kernels executed, and code size
// in a fetch-bound loop (profiles/).  (Not `const` on the device, so that they are not folded
// back into immediates; the values and the order of operations are unchanged.)
#ifdef RB_DEVICE_BUILD
#define RB_CTABLE __constant__
#else
#define RB_CTABLE static const
#endif
RB_CTABLE double kMeisselTab[47] = {
    1.0 / 10321920.0, 860160.0, 1290240.0, 2580480.0,
    645120.0, 28672.0, 2709504.0, 6547968.0,
    672000.0, 23224320.0, 18708480.0, 1048320.0,
    8192.0, 2519040.0, 60518400.0, 151828480.0,
    61254720.0, 2163168.0, 138700800.0, 800163840.0,
    940423680.0, 228049920.0, 5537280.0, 6144.0,
    2644992.0, 299351808.0, 3405435264.0, 8653594320.0,
    5897669400.0, 954875250.0, 16907985.0, 625766400.0,
    12841758720.0, 60631119360.0, 86387857920.0, 38435160960.0,
    4450158720.0, 59968440.0, 0.984023040e9, 0.442810368e9,
    0.303114240e9, 0.233192960e9, 0.190139040e9, 0.160692840e9,
    0.139204065e9, 0.1476034560e10, 1.0e-3,
};
RB_CTABLE double kDebyeTab[53] = {
    1.0 / 10321920.0, 810485676000000.0 * kAt4, 19451656224000000.0 * kAt0, 19451656224000000.0 * kAt1,
    3241942704000000.0 * kAt3, 27016189200000.0 * kAt6, 53603550000.0 * kAt9, 40608750.0 * kAt12,
    14875.0 * kAt15, 5403237840000.0 * kAt6, 69470200800000.0 * kAt4, 401283384.0 * kAt15,
    4707059994.0 * kAt13, 3012121710.0 * kAt12, 36011689560.0 * kAt10, 8027667648000.0 * kAt9,
    8027667648000.0 * kAt7, 1296777081600000.0 * kAt3, 41423013450.0 * kAt12, 484040056500.0 * kAt10,
    67540473000000.0 * kAt6, 1748257220.0 * kAt15, 19964735910.0 * kAt13, 2594411820000.0 * kAt9,
    29331862560000.0 * kAt7, 78248884350.0 * kAt12, 860873013000.0 * kAt10, 94556662200000.0 * kAt6,
    1938419560.0 * kAt15, 20997160275.0 * kAt13, 2283511230000.0 * kAt9, 47153256150.0 * kAt12,
    459918459000.0 * kAt10, 849093050.0 * kAt15, 8397889500.0 * kAt13, 643242600000.0 * kAt9,
    11448186750.0 * kAt12, 173573400.0 * kAt15, 1474097625.0 * kAt13, 1161410250.0 * kAt12,
    17481100.0 * kAt15, 833000.0 * kAt15, kAt7, kAt10,
    kAt13, 1.0e55, 21612951360000.0, 3859455600000.0,
    88445857500.0, 5360355000.0, 113704500.0, 3123750.0,
    0.58354968672000000e17,
};
RB_CTABLE double kExpFacTab[9] = {
    1.0 / 10321920.0, 40320.0, 20160.0, 6720.0,
    1680.0, 336.0, 56.0, 8.0,
    690.0,
};

// bessel.c:22-51
RB_FN double exp_factor(double f_factor, double f_exp)
{
    if (f_factor == 0.0)
        return 0.0;
    const double fabs_exp = fabs(f_exp);
    if (fabs_exp < 1e-3) {
        const double x = f_exp;
        return f_factor *
               (1.0 + ((kExpFacTab[1] + (kExpFacTab[2] + (kExpFacTab[3] + (kExpFacTab[4] + (kExpFacTab[5] + (kExpFacTab[6] + (kExpFacTab[7] + x) * x) * x) * x) * x) * x) * x) * x / kExpFacTab[1]));
    }
    // (one exp call site for the three cases of bessel.c:40-50)
    double mult = f_factor, arg = f_exp;
    if (fabs_exp > kExpFacTab[8]) {
        const double log_f = rb_log(fabs(f_factor));
        if (log_f * f_exp < 0.0) {
            mult = (f_factor < 0.0) ? -1.0 : 1.0;
            arg = log_f + f_exp;
        }
    }
    return mult * rb_exp(arg);
}

// Meissel's first expansion, x < n (bessel.c:94-149).
RB_FN double leung_meissel_first(const LeungOrder &o, double x)
{
    const double n = o.n;
#ifdef RB_LEAN_MATH
    const double z = x * o.ninv; // (the product path; the faithful kernels divide, as bessel.c does)
    const double eps = (n - x) * o.ninv;
#else
    const double z = x / n;
    const double eps = (n - x) / n;
#endif
    const double Z = rb_sqrt(eps * (1.0 + z));
    const double U = rb_rcp(n * Z * Z * Z);
    const double t1 = z * z;
    const double D = kMeisselTab[0];

    // V_n sum, a polynomial in U whose coefficients are polynomials in z^2
    const double p0 = (kMeisselTab[1] + kMeisselTab[2] * t1) * D;
    const double p1 = ((-kMeisselTab[3] - kMeisselTab[4] * t1) * t1) * D;
    const double p2 = (-kMeisselTab[5] + (kMeisselTab[6] + (kMeisselTab[7] + kMeisselTab[8] * t1) * t1) * t1) * D;
    const double p3 = ((-kMeisselTab[3] + (-kMeisselTab[9] + (-kMeisselTab[10] - kMeisselTab[11] * t1) * t1) * t1) * t1) * D;
    const double p4 =
        (-kMeisselTab[12] + (-kMeisselTab[13] + (-kMeisselTab[14] + (-kMeisselTab[15] + (-kMeisselTab[16] - kMeisselTab[17] * t1) * t1) * t1) * t1) * t1) * D;
    const double p5 =
        ((kMeisselTab[3] + (kMeisselTab[18] + (kMeisselTab[19] + (kMeisselTab[20] + (kMeisselTab[21] + kMeisselTab[22] * t1) * t1) * t1) * t1) * t1) * t1) * D;
    const double p6 = (kMeisselTab[23] + (-kMeisselTab[24] + (-kMeisselTab[25] + (-kMeisselTab[26] + (-kMeisselTab[27] + (-kMeisselTab[28] +
                       (-kMeisselTab[29] - kMeisselTab[30] * t1) * t1) * t1) * t1) * t1) * t1) * t1) * D;
    const double p7 = ((kMeisselTab[3] + (kMeisselTab[31] + (kMeisselTab[32] + (kMeisselTab[33] + (kMeisselTab[34] + (kMeisselTab[35] +
                       (kMeisselTab[36] + kMeisselTab[37] * t1) * t1) * t1) * t1) * t1) * t1) * t1) * t1) * D;
    const double vsum1 = U * (p0 + U * (p1 + U * (p2 + U * (p3 + U * (p4 + U * (p5 + U * (p6 + U * p7)))))));

    // "I substitute Gamma(n+1) with (n+1)*Gamma(n) in the denominator" (bessel.c:123)
    const double factor = rb_rcp((n + 1.0) * rb_sqrt(Z));

    double exp_val;
    if (eps < 1e-4 && n > 1e3) {
        const double exp2 = -n * rb_sqrt(2.0 * eps) * eps *
                            (kMeisselTab[38] + (kMeisselTab[39] + (kMeisselTab[40] + (kMeisselTab[41] + (kMeisselTab[42] +
                             (kMeisselTab[43] + kMeisselTab[44] * eps) * eps) * eps) * eps) * eps) * eps) * rb_rcp(kMeisselTab[45]);
        exp_val = o.c_stirling + exp2 - vsum1;
    } else {
        double inv_zp1;
        if (Z < kMeisselTab[46])
            inv_zp1 = 1.0 + (-1.0 + (1.0 + (-1.0 + (1.0 + (-1.0 + (1.0 - Z) * Z) * Z) * Z) * Z) * Z) * Z;
        else
            inv_zp1 = rb_rcp(1.0 + Z);
        exp_val = n * (rb_log(x * inv_zp1) - (1.0 - Z)) - vsum1 + o.c_std;
    }
    return exp_factor(factor, exp_val);
}

// Meissel's second expansion, x > n (bessel.c:57-88).  The reference uses
// 80-bit acosl/cosl; the device has no long double.  The Symphony integrand
// never gets here (z < n identically), so this is kept out of line.
RB_FN_NOINLINE double leung_meissel_second(double n, double x)
{
    const double z = x / n;
    const double eps = (x - n) / n;
    const double Z = sqrt(eps * (1.0 + z));
    const double U = 1.0 / (n * Z * Z * Z);
    const double t1 = z * z;
    const double t2 = U * U;

    const double exp_val =
        (t1 * t2 * (-3072.0 - 768.0 * t1 + (3072.0 + (27648.0 + (22272.0 + 1248.0 * t1) * t1) * t1 +
        (-3072.0 + (-165120.0 + (-952576.0 + (-1119552.0 + (-271488.0 - 6592.0 * t1) * t1) * t1) * t1) * t1 +
        (3072.0 + (744960.0 + (15287808.0 + (72179904.0 + (102842688.0 + (45756144.0 + (5297808.0 +
        71391.0 * t1) * t1) * t1) * t1) * t1) * t1) * t1) * t2) * t2) * t2)) / 0.12288e5;

    const double Qt = n * (Z - acos(n / x));

    const double Qsum =
        -(U * (860160.0 + 1290240.0 * t1 + (28672.0 + (-2709504.0 + (-6547968.0 - 672000.0 * t1) * t1) * t1 +
        (8192.0 + (2519040.0 + (60518400.0 + (151828480.0 + (61254720.0 + 2163168.0 * t1) * t1) * t1) * t1) * t1 +
        (-6144.0 + (2644992.0 + (299351808.0 + (3405435264.0 + (8653594320.0 + (5897669400.0 +
        (954875250.0 + 16907985.0 * t1) * t1) * t1) * t1) * t1) * t1) * t1) * t2) * t2) * t2)) / 0.10321920e8;

    const double factor = sqrt(2.0 / (kPi * n * Z)) * cos(Qsum + Qt - 0.25 * kPi);
    return exp_factor(factor, exp_val);
}

// The Debye "epsilon" expansion about x = n (bessel.c:159-213): a degree-14
// polynomial in ez = x - n whose coefficients are polynomials in x^(1/3).
RB_FN double leung_debye_eps(double n, double x)
{
    if (x > kDebyeTab[45])
        return NAN;

    const double ez = x - n;
    const double z = rb_cbrt(x);
    const double t3 = z * z;
    const double t4 = x * z;
    const double t10 = t4 * t4;
    const double t38 = kDebyeTab[1] * t3;
    const double t70 = kDebyeTab[42] * t3;
    const double t93 = kDebyeTab[43] * t3;
    const double t107 = kDebyeTab[44] * t3;
    const double t114 = ez * ez;
    const double t146 = t10 * t10;

    const double lead = (-kDebyeTab[9] + (kDebyeTab[10] + kDebyeTab[2] * t4) * t3) * t10 * z;

    const double e0 = -kDebyeTab[11] + (kDebyeTab[12] + (kDebyeTab[13] + (-kDebyeTab[14] +
                      (kDebyeTab[15] + (-kDebyeTab[16] + (-kDebyeTab[17] +
                      kDebyeTab[3] * t3) * t4) * t3) * z) * t3) * z) * t3;
    const double e1 = (-kDebyeTab[18] + (kDebyeTab[19] + (kDebyeTab[20] - t38) * t4) * t3) * x;
    const double e2 = kDebyeTab[21] + (-kDebyeTab[22] + (-kDebyeTab[23] +
                      (kDebyeTab[24] + kDebyeTab[4] * t4) * t3) * t4) * t3;
    const double e3 = (kDebyeTab[25] + (-kDebyeTab[26] + (-kDebyeTab[27] + t38) * t4) * t3) * x;
    const double e4 = -kDebyeTab[28] + (kDebyeTab[29] + (kDebyeTab[30] - kDebyeTab[46] * t70) * t4) * t3;
    const double e5 = (-kDebyeTab[31] + (kDebyeTab[32] + kDebyeTab[5] * t4) * t3) * x;
    const double e6 = kDebyeTab[33] + (-kDebyeTab[34] + (-kDebyeTab[35] + kDebyeTab[47] * t70) * t4) * t3;
    const double e7 = (kDebyeTab[36] - kDebyeTab[48] * t93) * x;
    const double e8 = -kDebyeTab[37] + (kDebyeTab[38] + kDebyeTab[6] * t4) * t3;
    const double e9 = (-kDebyeTab[39] + kDebyeTab[49] * t93) * x;
    const double e10 = -kDebyeTab[50] * t107 + kDebyeTab[40];
    const double e11 = kDebyeTab[7] * x;
    const double e12 = -kDebyeTab[41] + kDebyeTab[51] * t107 + kDebyeTab[8] * t114;

    const double poly = (e0 + (e1 + (e2 + (e3 + (e4 + (e5 + (e6 + (e7 + (e8 + (e9 + (e10 + (e11 + e12 * ez) * ez) * ez) * ez) *
                        ez) * ez) * ez) * ez) * ez) * ez) * ez) * ez) * ez;

    return rb_div(lead + poly, kPi * t146 * kDebyeTab[52]);
}

// The same expansion for the orders n and n + 1 at one x (the Symphony integrand needs J_n and J_{n+1} at
// every node): the coefficients e_0 ... e_11 and the lead term depend on x only, so they are computed once and
// feed two Horner chains in ez = x - n and x - n - 1.  Same values as two calls of leung_debye_eps().
RB_FN void leung_debye_eps_pair(double n0, double n1, double x, double &d0, double &d1)
{
    if (x > kDebyeTab[45]) {
        d0 = d1 = NAN;
        return;
    }
    const double ez0 = x - n0, ez1 = x - n1;
    const double z = rb_cbrt(x);
    const double t3 = z * z;
    const double t4 = x * z;
    const double t10 = t4 * t4;
    const double t38 = kDebyeTab[1] * t3;
    const double t70 = kDebyeTab[42] * t3;
    const double t93 = kDebyeTab[43] * t3;
    const double t107 = kDebyeTab[44] * t3;
    const double t146 = t10 * t10;
    const double e12a = -kDebyeTab[41] + kDebyeTab[51] * t107;
    double p0 = e12a + kDebyeTab[8] * (ez0 * ez0), p1 = e12a + kDebyeTab[8] * (ez1 * ez1);
#define RB_DEBYE_STEP(expr)        \
    {                              \
        const double ek_ = (expr); \
        p0 = ek_ + p0 * ez0;       \
        p1 = ek_ + p1 * ez1;       \
    }
    RB_DEBYE_STEP(kDebyeTab[7] * x)
    RB_DEBYE_STEP(-kDebyeTab[50] * t107 + kDebyeTab[40])
    RB_DEBYE_STEP((-kDebyeTab[39] + kDebyeTab[49] * t93) * x)
    RB_DEBYE_STEP(-kDebyeTab[37] + (kDebyeTab[38] + kDebyeTab[6] * t4) * t3)
    RB_DEBYE_STEP((kDebyeTab[36] - kDebyeTab[48] * t93) * x)
    RB_DEBYE_STEP(kDebyeTab[33] + (-kDebyeTab[34] + (-kDebyeTab[35] + kDebyeTab[47] * t70) * t4) * t3)
    RB_DEBYE_STEP((-kDebyeTab[31] + (kDebyeTab[32] + kDebyeTab[5] * t4) * t3) * x)
    RB_DEBYE_STEP(-kDebyeTab[28] + (kDebyeTab[29] + (kDebyeTab[30] - kDebyeTab[46] * t70) * t4) * t3)
    RB_DEBYE_STEP((kDebyeTab[25] + (-kDebyeTab[26] + (-kDebyeTab[27] + t38) * t4) * t3) * x)
    RB_DEBYE_STEP(kDebyeTab[21] + (-kDebyeTab[22] + (-kDebyeTab[23] + (kDebyeTab[24] + kDebyeTab[4] * t4) * t3) * t4) * t3)
    RB_DEBYE_STEP((-kDebyeTab[18] + (kDebyeTab[19] + (kDebyeTab[20] - t38) * t4) * t3) * x)
    RB_DEBYE_STEP(-kDebyeTab[11] + (kDebyeTab[12] + (kDebyeTab[13] + (-kDebyeTab[14] + (kDebyeTab[15] + (-kDebyeTab[16] +
                  (-kDebyeTab[17] + kDebyeTab[3] * t3) * t4) * t3) * z) * t3) * z) * t3)
#undef RB_DEBYE_STEP
    const double lead = (-kDebyeTab[9] + (kDebyeTab[10] + kDebyeTab[2] * t4) * t3) * t10 * z;
    const double inv_den = rb_rcp(kPi * t146 * kDebyeTab[52]);
    d0 = (lead + p0 * ez0) * inv_den;
    d1 = (lead + p1 * ez1) * inv_den;
}

// Integer order 0 <= n < 30 (the reference calls gsl_sf_bessel_Jn,
// bessel.c:327-334).  Miller's backward recurrence normalised by
// J_0 + 2 sum J_2k = 1 where the recurrence is stable (x < n + 1, the only
// region the Symphony integrand reaches); upward recurrence from j0/j1 beyond.
RB_FN_NOINLINE double bessel_jn_small_int(int n, double x)
{
    if (x == 0.0)
        return n == 0 ? 1.0 : 0.0;

    if (x >= (double)n + 1.0) {
        double jm = j0(x);
        if (n == 0)
            return jm;
        double jc = j1(x);
        const double tox = 2.0 / x;
        for (int k = 1; k < n; k++) {
            const double jn_ = k * tox * jc - jm;
            jm = jc;
            jc = jn_;
        }
        return jc;
    }

    const double tox = 2.0 / x;
    const int m = 2 * ((n + 46) / 2);
    double bjp = 0.0, bj = 1.0, sum = 0.0, ans = 0.0;
    for (int k = m; k > 0; k--) {
        const double bjm = k * tox * bj - bjp;
        bjp = bj;
        bj = bjm;
        if (fabs(bj) > 1e150) {
            bj *= 1e-150;
            bjp *= 1e-150;
            ans *= 1e-150;
            sum *= 1e-150;
        }
        // bj now holds the unnormalised J_{k-1}
        if (((k - 1) & 1) == 0 && k - 1 > 0)
            sum += bj;
        if (k - 1 == n)
            ans = bj;
    }
    sum = 2.0 * sum + bj; // bj = J_0
    return ans / sum;
}

// J_n(x) and J_{n+1}(x) for integer 0 <= n, n + 1 < 30 and 0 < x < n + 1 from ONE backward recurrence (the
// Symphony integrand needs both at every node, and the recurrence is the whole cost of a discrete harmonic).
RB_FN_NOINLINE void bessel_jn_pair_small_int(int n, double x, double &jn, double &jn1)
{
    const double tox = 2.0 * rb_rcp(x);
    int k = 2 * ((n + 47) / 2);
    double kd = (double)k;
    double bjp = 0.0, bj = 1.0, sum = 0.0, a0 = 0.0, a1 = 0.0;
    // state at the top of a step: bj = J_k, bjp = J_{k+1} (unnormalised); a step goes from k to k - 1.
    // Two loops with the same body: down to k = n + 1, where J_{n+1} and then J_n are taken, and on to k = 0.
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
        const int k_end = phase ? 0 : n + 1;
#pragma unroll 1
        while (k > k_end) {
            const double bjm = kd * tox * bj - bjp;
            bjp = bj;
            bj = bjm;
            kd -= 1.0;
            k--;
            if (fabs(bj) > 1e150) {
                bj *= 1e-150;
                bjp *= 1e-150;
                a0 *= 1e-150;
                a1 *= 1e-150;
                sum *= 1e-150;
            }
            if ((k & 1) == 0 && k > 0)
                sum += bj;
            if (k == n)
                a0 = bj; // (taken in the first step of the second phase)
        }
        if (phase == 0)
            a1 = bj;
    }
    const double inv = rb_rcp(2.0 * sum + bj); // bj = J_0
    jn = a0 * inv;
    jn1 = a1 * inv;
}

// pkgw_bessel_j (bessel.c:318-376).
RB_FN double leung_j(const LeungOrder &o, double x)
{
    if (o.kind == kOrderInvalid || !(x >= 0.0))
        return NAN;
    if (o.kind == kOrderInteger)
        return bessel_jn_small_int(o.nint, x);

    const double n = o.n;
    if (x == n)
        return leung_debye_eps(n, x);

    if (x < n) {
        const double eps = (n - x) / n;
        if (eps < o.lo_minus)
            return leung_debye_eps(n, x);
        if (eps > o.hi_minus)
            return leung_meissel_first(o, x);
        const double debye = leung_debye_eps(n, x);
        const double meissel1 = leung_meissel_first(o, x);
        const double eta = rb_log(eps) * kLog10e;
        const double pos = (eta - o.eta_lo_minus) / (kMinusEtaB - kMinusEtaA);
        return debye * (1.0 - pos) + meissel1 * pos;
    } else {
        const double eps = (x - n) / x;
        if (eps < o.lo_plus)
            return leung_debye_eps(n, x);
        if (eps > o.hi_plus)
            return leung_meissel_second(n, x);
        const double debye = leung_debye_eps(n, x);
        const double meissel2 = leung_meissel_second(n, x);
        const double eta = rb_log(eps) * kLog10e;
        const double pos = (eta - o.eta_lo_plus) / (kPlusEtaB - kPlusEtaA);
        return debye * (1.0 - pos) + meissel2 * pos;
    }
}

// J_n(x) and J_n'(x) = n J_n(x)/x - J_{n+1}(x) (pkgw_bessel_dj, bessel.c:379-405),
// given the prepared orders n and n + 1.
RB_FN void leung_j_and_dj(const LeungOrder &on, const LeungOrder &on1, double x, double &jn, double &djn)
{
    jn = leung_j(on, x);
    if (on.n >= 1e15) {
        djn = NAN;
        return;
    }
    const double jnp1 = leung_j(on1, x);
    if (x == 0.0) {
        if (on.n >= 2.0)
            djn = 0.0;
        else if (on.n == 0.0)
            djn = -jnp1;
        else
            djn = on.n * jn / DBL_MIN - jnp1;
        return;
    }
    djn = on.n * jn / x - jnp1;
}

} // namespace rb
