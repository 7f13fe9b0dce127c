// rb_special.cuh -- special functions needed by the Heyvaerts Faraday
// integrals on the device.
//
// Replaces the reference's calls into `special-fun` (Cephes) at
// src/heyvaerts.rs:331-332, 335-336, 359-360, 363, 437-438, 440-441:
//   besseli(+-1/3), besseli(+-2/3)  at 0 < g < 10
//   besselj / bessely of real order sigma, sigma - 1 at 0 < x < sigma < 3
// (the J/Y branch is only reachable for sigma < 3: it needs g >= 10, i.e. a
// point of the quasi-resonant domain far below the turning point, and for
// sigma >= 3^{3/2}... the domain is capped at pomega_phys where g <= ~1; see
// DESIGN.md section 5).  Algorithms: ascending series for I_nu and J_nu,
// Temme's series for Y_mu, |mu| <= 1/2, followed by upward recurrence.
#pragma once

#include "rb_core.cuh"

namespace rb {

// Taylor coefficients of 1/Gamma(1 + mu) about mu = 0.
RB_TABLE double RGAMMA1P_COEF[28] = {
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
    1.4123806553180317816e-18};

// For |mu| <= 1/2: gampl = 1/Gamma(1+mu), gammi = 1/Gamma(1-mu),
// gam1 = (gammi - gampl)/(2 mu), gam2 = (gammi + gampl)/2.
RB_FN void gamma_pair(double mu, double &gam1, double &gam2, double &gampl, double &gammi)
{
    const double m2 = mu * mu;
    double even = 0.0, odd = 0.0;
    for (int j = 26; j >= 0; j -= 2)
        even = even * m2 + RGAMMA1P_COEF[j];
    for (int j = 27; j >= 1; j -= 2)
        odd = odd * m2 + RGAMMA1P_COEF[j];
    gam2 = even;
    gam1 = -odd;
    gampl = even + mu * odd;
    gammi = even - mu * odd;
}

// 1/Gamma(1 + nu), nu in (-1, ~8).
RB_FN double rgamma1p(double nu)
{
    const double m = floor(nu + 0.5);
    const double mu = nu - m;
    double g1, g2, gp, gm;
    gamma_pair(mu, g1, g2, gp, gm);
    double r = gp;
    const int im = (int)m;
    if (im >= 0) {
        for (int i = 1; i <= im; i++)
            r *= rb_rcp(mu + i);
    } else {
        r *= mu; // 1/Gamma(mu) = mu/Gamma(1 + mu)
    }
    return r;
}

// J_nu(x), nu > -1, 0 < x <~ 6, ascending series.  (Out of line, like bessel_y_temme: the pair
// function calls them five times and sits in the inner loop of the s sin(theta) < 3 points.)
RB_FN_NOINLINE double bessel_j_series(double nu, double x)
{
    const double q = -0.25 * x * x;
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 200; k++) {
        term *= q * rb_rcp(k * (k + nu));
        sum += term;
        if (fabs(term) < 1e-17 * fabs(sum))
            break;
    }
    return rb_exp(nu * rb_log(0.5 * x)) * rgamma1p(nu) * sum;
}

// Temme's series: Y_mu(x), Y_{mu+1}(x), |mu| <= 1/2, 0 < x <~ 4.
RB_FN_NOINLINE void bessel_y_temme(double mu, double x, double &y_mu, double &y_mu1)
{
    const double eps = DBL_EPSILON;
    double gam1, gam2, gampl, gammi;
    gamma_pair(mu, gam1, gam2, gampl, gammi);
    const double x2 = 0.5 * x;
    const double pimu = kPi * mu;
    const double fact = (fabs(pimu) < eps) ? 1.0 : rb_div(pimu, sin(pimu));
    double d = -rb_log(x2);
    double e = mu * d;
    const double fact2 = (fabs(e) < eps) ? 1.0 : rb_div(sinh(e), e);
    double ff = 2.0 / kPi * fact * (gam1 * cosh(e) + gam2 * fact2 * d);
    e = rb_exp(e);
    double p = rb_div(e, gampl * kPi);
    double q = rb_rcp(e * kPi * gammi);
    const double pimu2 = 0.5 * pimu;
    const double fact3 = (fabs(pimu2) < eps) ? 1.0 : rb_div(sin(pimu2), pimu2);
    const double r = kPi * pimu2 * fact3 * fact3;
    double c = 1.0;
    d = -x2 * x2;
    double sum = ff + r * q;
    double sum1 = p;
    for (int i = 1; i < 500; i++) {
        // 1 / (i - mu), 1 / (i + mu) and 1 / (i^2 - mu^2) from one reciprocal
        const double im = i - mu, ip = i + mu;
        const double rr = rb_rcp(im * ip);
        ff = (i * ff + p + q) * rr;
        c *= d * rb_rcp((double)i);
        p *= ip * rr;
        q *= im * rr;
        const double del = c * (ff + r * q);
        sum += del;
        const double del1 = c * p - i * del;
        sum1 += del1;
        if (fabs(del) < (1.0 + fabs(sum)) * eps * 0.1)
            break;
    }
    y_mu = -sum;
    y_mu1 = -sum1 * (2.0 * rb_rcp(x));
}

// J and Y of orders sigma and sigma - 1 at x, for 0 < sigma <~ 8, 0 < x <~ 4.
RB_FN_NOINLINE void bessel_jy_pair(double sigma, double x, double &j_s, double &j_sm1, double &y_s, double &y_sm1)
{
    if (!(x > 0.0) || !(sigma > 0.0)) {
        j_s = j_sm1 = y_s = y_sm1 = NAN;
        return;
    }
    j_s = bessel_j_series(sigma, x);
    j_sm1 = bessel_j_series(sigma - 1.0, x);

    const double lo = sigma - 1.0;
    if (lo >= -0.5) {
        const double m = floor(lo + 0.5);
        const double mu = lo - m;
        double ya, yb;
        bessel_y_temme(mu, x, ya, yb); // Y_mu, Y_{mu+1}
        const int im = (int)m;
        for (int i = 1; i <= im; i++) {
            const double yn = 2.0 * (mu + i) * rb_rcp(x) * yb - ya;
            ya = yb;
            yb = yn;
        }
        // ya = Y_{sigma-1}, yb = Y_sigma
        y_sm1 = ya;
        y_s = yb;
    } else {
        // sigma < 1/2: Y_sigma directly, Y_{sigma-1} = Y_{-a}, a = 1 - sigma in (1/2, 1), by reflection
        double ys, tmp;
        bessel_y_temme(sigma, x, ys, tmp);
        y_s = ys;
        const double a = 1.0 - sigma;
        double ym, ya;
        bessel_y_temme(-sigma, x, ym, ya); // Y_{-sigma}, Y_{1-sigma} = Y_a
        const double ja = bessel_j_series(a, x);
        y_sm1 = sin(a * kPi) * ja + cos(a * kPi) * ya;
    }
}

constexpr int kISeriesMax = 48;
// 1/(k (k + nu)), rows nu = 1/3, -1/3, 2/3, -2/3; column k (generated with mpmath)
RB_TABLE double ISERIES_RECIP[4][48] = {
    {0.0, 0.75, 0.21428571428571428571, 0.1, 0.057692307692307692308, 0.0375, 0.026315789473684210526, 0.019480519480519480519, 0.015, 0.011904761904761904762, 0.0096774193548387096774, 0.0080213903743315508021, 0.0067567567567567567568, 0.0057692307692307692308, 0.0049833887043189368771, 0.0043478260869565217391, 0.0038265306122448979592, 0.003393665158371040724, 0.003030303030303030303, 0.002722323049001814882, 0.0024590163934426229508, 0.0022321428571428571429, 0.0020352781546811397558, 0.0018633540372670807453, 0.0017123287671232876712, 0.0015789473684210526316, 0.0014605647517039922103, 0.001355013550135501355, 0.0012605042016806722689, 0.00117554858934169279, 0.0010989010989010989011, 0.001029512697323266987, 0.00096649484536082474227, 0.00090909090909090909091, 0.00085665334094802969732, 0.00080862533692722371968, 0.00076452599388379204893, 0.00072393822393822393822, 0.0006864988558352402746, 0.00065189048239895697523, 0.00061983471074380165289, 0.00059008654602675059009, 0.0005624296962879640045, 0.00053667262969588550984, 0.00051264524948735475051, 0.00049019607843137254902, 0.00046918986549890522365, 0.00044950554390170812107},
    {0.0, 1.5, 0.3, 0.125, 0.068181818181818181818, 0.042857142857142857143, 0.029411764705882352941, 0.021428571428571428571, 0.016304347826086956522, 0.012820512820512820513, 0.010344827586206896552, 0.0085227272727272727273, 0.0071428571428571428571, 0.0060728744939271255061, 0.0052264808362369337979, 0.0045454545454545454545, 0.0039893617021276595745, 0.0035294117647058823529, 0.0031446540880503144654, 0.0028195488721804511278, 0.0025423728813559322034, 0.0023041474654377880184, 0.0020979020979020979021, 0.0019181585677749360614, 0.0017605633802816901408, 0.0016216216216216216216, 0.0014985014985014985015, 0.0013888888888888888889, 0.0012908777969018932874, 0.0012028869286287089014, 0.0011235955056179775281, 0.0010518934081346423562, 0.00098684210526315789474, 0.0009276437847866419295, 0.00087361677344205008736, 0.00082417582417582417582, 0.0007788161993769470405, 0.00073710073710073710074, 0.00069864927806241266884, 0.00066312997347480106101, 0.00063025210084033613445, 0.00059976009596161535386, 0.00057142857142857142857, 0.00054505813953488372093, 0.00052047189451769604441, 0.00049751243781094527363, 0.00047603935258648048239, 0.0004559270516717325228},
    {0.0, 0.6, 0.1875, 0.090909090909090909091, 0.053571428571428571429, 0.035294117647058823529, 0.025, 0.018633540372670807453, 0.014423076923076923077, 0.011494252873563218391, 0.009375, 0.0077922077922077922078, 0.0065789473684210526316, 0.0056285178236397748593, 0.0048701298701298701299, 0.0042553191489361702128, 0.00375, 0.0033296337402885682575, 0.0029761904761904761905, 0.0026761819803746654773, 0.0024193548387096774194, 0.0021978021978021978022, 0.0020053475935828877005, 0.0018371096142069810165, 0.0016891891891891891892, 0.0015584415584415584416, 0.0014423076923076923077, 0.0013386880856760374833, 0.0012458471760797342193, 0.0011623401782254939946, 0.0010869565217391304348, 0.0010186757215619694397, 0.0009566326530612244898, 0.00090009000900090009001, 0.000848416289592760181, 0.00080106809078771695594, 0.00075757575757575757576, 0.00071753169098301841665, 0.00068058076225045372051, 0.00064641241111829347123, 0.0006147540983606557377, 0.00058536585365853658537, 0.00055803571428571428571, 0.00053257589206461920824, 0.00050881953867028493894, 0.00048661800486618004866, 0.00046583850931677018634, 0.00044636214848980806428},
    {0.0, 3.0, 0.375, 0.14285714285714285714, 0.075, 0.046153846153846153846, 0.03125, 0.022556390977443609023, 0.017045454545454545455, 0.013333333333333333333, 0.010714285714285714286, 0.0087976539589442815249, 0.0073529411764705882353, 0.0062370062370062370062, 0.0053571428571428571429, 0.0046511627906976744186, 0.0040760869565217391304, 0.0036014405762304921969, 0.0032051282051282051282, 0.0028708133971291866029, 0.0025862068965517241379, 0.0023419203747072599532, 0.0021306818181818181818, 0.0019467878001297858533, 0.0017857142857142857143, 0.0016438356164383561644, 0.0015182186234817813765, 0.001406469760900140647, 0.0013066202090592334495, 0.0012170385395537525355, 0.0011363636363636363636, 0.0010634526763559021624, 0.00099734042553191489362, 0.00093720712277413308341, 0.00088235294117647058824, 0.00083217753120665742025, 0.00078616352201257861635, 0.00074386312918423010166, 0.00070488721804511278195, 0.00066889632107023411371, 0.00063559322033898305085, 0.00060471679096956258819, 0.00057603686635944700461, 0.00054934993590917414393, 0.00052447552447552447552, 0.0005012531328320802005, 0.00047953964194373401535, 0.00045920710240318383591},
};

// I_{1/3}, I_{-1/3}, I_{2/3}, I_{-2/3} at 0 < g < ~12 by the ascending series;
// the four share (g/2)^2 and one cube root.  (Summing backwards from a tabulated number of terms, one multiply and
// one fused multiply-add per term and no convergence test, was measured 10 % slower for the whole Heyvaerts
// kernel: the table has to be conservative and most nodes need three or four terms.)
RB_FN void bessel_i_thirds(double g, double &ip13, double &im13, double &ip23, double &im23)
{
    const double q = 0.25 * g * g;
    double t0 = 1.0, t1 = 1.0, t2 = 1.0, t3 = 1.0;
    double s0 = 1.0, s1 = 1.0, s2 = 1.0, s3 = 1.0;
    for (int k = 1; k < kISeriesMax; k++) {
        t0 *= q * ISERIES_RECIP[0][k];
        t1 *= q * ISERIES_RECIP[1][k];
        t2 *= q * ISERIES_RECIP[2][k];
        t3 *= q * ISERIES_RECIP[3][k];
        s0 += t0;
        s1 += t1;
        s2 += t2;
        s3 += t3;
#ifdef RB_ISERIES_ONE_TEST
        if (t3 < 1e-17 * s3) // nu = -2/3 has the largest terms of the four
            break;
#else
        if (t3 < 1e-17 * s3 && t1 < 1e-17 * s1)
            break;
#endif
    }
    const double c = rb_cbrt(0.5 * g); // (g/2)^(1/3)
    const double ci = rb_rcp(c);
    ip13 = c * 1.119846521722185685 * s0;
    im13 = ci * 0.73848811162164831294 * s1;
    ip23 = c * c * 1.1077321674324724694 * s2;
    im23 = ci * ci * 0.37328217390739522833 * s3;
}

} // namespace rb
