/* oracle/gsl_restated.h -- TEST INFRASTRUCTURE ONLY (the CPU oracle).
 *
 * Restatement of the three GSL entry points that rimphony's hot path calls
 * (reference: src/gsl.rs:169-180 `gsl_integration_qag` with key=3,
 * src/gsl.rs:246 `gsl_deriv_central`).  GSL itself is a third-party system
 * dependency of the reference (gsl-sys/build.rs:11-14, version ">= 1.0",
 * otherwise unpinned) and is absent from /root/reference and from this image,
 * so the algorithm is restated from the published QUADPACK / GSL description
 * (SURVEY.md Appendix A).  The 31-point Gauss-Kronrod constants were derived
 * from first principles with mpmath (roots of P_15 and of the Stieltjes
 * polynomial E_16, weights from the Legendre moment equations) and agree with
 * the QUADPACK qk31 table to 1e-25.
 *
 * Nothing outside tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may call this code.
 */
#ifndef RIMPHONY_ORACLE_GSL_RESTATED_H
#define RIMPHONY_ORACLE_GSL_RESTATED_H

#include <float.h>
#include <math.h>
#include <stddef.h>

typedef double (*orc_fn)(double x, void *ctx);

/* GSL_MAX_DBL: plain ternary, so a NaN second operand wins (unlike fmax). */
#define ORC_GSL_MAX(a, b) ((a) > (b) ? (a) : (b))

/* Status codes, numbered like GSL's gsl_errno.h so that messages in tests read
 * the same as the reference's `GslError`. */
enum {
    ORC_SUCCESS = 0,
    ORC_EFAILED = 5,
    ORC_EMAXITER = 11,
    ORC_EBADTOL = 13,
    ORC_EROUND = 18,
    ORC_ESING = 21
};

#define ORC_QAG_MAX_INTERVALS 5000

typedef struct {
    size_t limit;
    size_t size;
    double alist[ORC_QAG_MAX_INTERVALS];
    double blist[ORC_QAG_MAX_INTERVALS];
    double rlist[ORC_QAG_MAX_INTERVALS];
    double elist[ORC_QAG_MAX_INTERVALS];
    /* statistics for SURVEY/BASELINE style probes */
    size_t max_size_seen;
    unsigned long n_rule_applications;
    int trace; /* debug: print every bisection (parity studies) */
} orc_workspace;

static const double orc_xgk31[16] = {
    0.998002298693397060285172840152271,
    0.987992518020485428489565718586613,
    0.967739075679139134257347978784337,
    0.937273392400705904307758947710209,
    0.897264532344081900882509656454496,
    0.848206583410427216200648320774217,
    0.790418501442465932967649294817947,
    0.724417731360170047416186054613938,
    0.650996741297416970533735895313275,
    0.570972172608538847537226737253911,
    0.485081863640239680693655740232351,
    0.394151347077563369897207370981045,
    0.299180007153168812166780024266389,
    0.201194093997434522300628303394596,
    0.101142066918717499027074231447392,
    0.000000000000000000000000000000000
};

static const double orc_wgk31[16] = {
    0.005377479872923348987792051430128,
    0.015007947329316122538374763075807,
    0.025460847326715320186874001019653,
    0.035346360791375846222037948478360,
    0.044589751324764876608227299373280,
    0.053481524690928087265343147239430,
    0.062009567800670640285139230960803,
    0.069854121318728258709520077099147,
    0.076849680757720378894432777482659,
    0.083080502823133021038289247286104,
    0.088564443056211770647275443693774,
    0.093126598170825321225486872747346,
    0.096642726983623678505179907627589,
    0.099173598721791959332393173484603,
    0.100769845523875595044946662617570,
    0.101330007014791549017374792767493
};

/* weights of the embedded 15-point Gauss rule (7 symmetric pairs + centre) */
static const double orc_wg15[8] = {
    0.030753241996117268354628393577204,
    0.070366047488108124709267416450667,
    0.107159220467171935011869546685869,
    0.139570677926154314447804794511028,
    0.166269205816993933553200860481209,
    0.186161000015562211026800561866423,
    0.198431485327111576456118326443839,
    0.202578241925561272880620199967519
};

static inline void orc_workspace_init(orc_workspace *w, size_t limit)
{
    w->limit = limit > ORC_QAG_MAX_INTERVALS ? ORC_QAG_MAX_INTERVALS : limit;
    w->size = 0;
    w->max_size_seen = 0;
    w->n_rule_applications = 0;
    w->trace = 0;
}

static inline double orc_rescale_error(double err, double result_abs, double result_asc)
{
    err = fabs(err);

    if (result_asc != 0 && err != 0) {
        double scale = pow((200 * err / result_asc), 1.5);
        if (scale < 1)
            err = result_asc * scale;
        else
            err = result_asc;
    }

    if (result_abs > DBL_MIN / (50 * DBL_EPSILON)) {
        double min_err = 50 * DBL_EPSILON * result_abs;
        if (min_err > err)
            err = min_err;
    }

    return err;
}

/* One application of the (15, 31) Gauss-Kronrod pair on [a, b]. */
static inline void orc_qk31(orc_fn f, void *ctx, double a, double b,
                            double *result, double *abserr, double *resabs, double *resasc)
{
    const int n = 16;
    double fv1[16], fv2[16];
    const double center = 0.5 * (a + b);
    const double half_length = 0.5 * (b - a);
    const double abs_half_length = fabs(half_length);
    const double f_center = f(center, ctx);

    double result_gauss = f_center * orc_wg15[n / 2 - 1];
    double result_kronrod = f_center * orc_wgk31[n - 1];
    double result_abs = fabs(result_kronrod);
    double result_asc, mean, err;
    int j;

    for (j = 0; j < (n - 1) / 2; j++) {
        const int jtw = j * 2 + 1;
        const double abscissa = half_length * orc_xgk31[jtw];
        const double fval1 = f(center - abscissa, ctx);
        const double fval2 = f(center + abscissa, ctx);
        const double fsum = fval1 + fval2;
        fv1[jtw] = fval1;
        fv2[jtw] = fval2;
        result_gauss += orc_wg15[j] * fsum;
        result_kronrod += orc_wgk31[jtw] * fsum;
        result_abs += orc_wgk31[jtw] * (fabs(fval1) + fabs(fval2));
    }

    for (j = 0; j < n / 2; j++) {
        const int jtwm1 = j * 2;
        const double abscissa = half_length * orc_xgk31[jtwm1];
        const double fval1 = f(center - abscissa, ctx);
        const double fval2 = f(center + abscissa, ctx);
        fv1[jtwm1] = fval1;
        fv2[jtwm1] = fval2;
        result_kronrod += orc_wgk31[jtwm1] * (fval1 + fval2);
        result_abs += orc_wgk31[jtwm1] * (fabs(fval1) + fabs(fval2));
    }

    mean = result_kronrod * 0.5;
    result_asc = orc_wgk31[n - 1] * fabs(f_center - mean);

    for (j = 0; j < n - 1; j++)
        result_asc += orc_wgk31[j] * (fabs(fv1[j] - mean) + fabs(fv2[j] - mean));

    err = (result_kronrod - result_gauss) * half_length;
    result_kronrod *= half_length;
    result_abs *= abs_half_length;
    result_asc *= abs_half_length;

    *result = result_kronrod;
    *resabs = result_abs;
    *resasc = result_asc;
    *abserr = orc_rescale_error(err, result_abs, result_asc);
}

static inline int orc_subinterval_too_small(double a1, double a2, double b2)
{
    const double e = DBL_EPSILON;
    const double u = DBL_MIN;
    double tmp = (1 + 100 * e) * (fabs(a2) + 1000 * u);
    return fabs(a1) <= tmp && fabs(b2) <= tmp;
}

/* Index of the interval with the largest error estimate; among equals the
 * earliest-stored one (QUADPACK's qpsrt keeps a descending order list and QAG
 * always takes its head, nrmax = 0). */
static inline size_t orc_ws_argmax(const orc_workspace *w)
{
    size_t i, best = 0;
    for (i = 1; i < w->size; i++)
        if (w->elist[i] > w->elist[best])
            best = i;
    return best;
}

/* Debug aid for the parity studies (tests/golden/make_stability.py and friends): when set, every
 * bisection of a QAG call whose workspace has `trace` on is printed to stderr. */
#include <stdio.h>

/* gsl_integration_qag with key = GSL_INTEG_GAUSS31. */
static inline int orc_qag31(orc_fn f, void *ctx, double a, double b,
                            double epsabs, double epsrel, orc_workspace *w,
                            double *result, double *abserr)
{
    double area, errsum, result0, abserr0, resabs0, resasc0, tolerance, round_off;
    size_t iteration = 0, i;
    int roundoff_type1 = 0, roundoff_type2 = 0, error_type = 0;
    const size_t limit = w->limit;

    w->size = 0;
    *result = 0;
    *abserr = 0;

    if (epsabs <= 0 && (epsrel < 50 * DBL_EPSILON || epsrel < 0.5e-28))
        return ORC_EBADTOL;

    orc_qk31(f, ctx, a, b, &result0, &abserr0, &resabs0, &resasc0);
    w->n_rule_applications++;

    w->alist[0] = a;
    w->blist[0] = b;
    w->rlist[0] = result0;
    w->elist[0] = abserr0;
    w->size = 1;
    if (w->max_size_seen < 1)
        w->max_size_seen = 1;

    tolerance = ORC_GSL_MAX(epsabs, epsrel * fabs(result0));
    round_off = 50 * DBL_EPSILON * resabs0;

    if (abserr0 <= round_off && abserr0 > tolerance) {
        *result = result0;
        *abserr = abserr0;
        return ORC_EROUND;
    } else if ((abserr0 <= tolerance && abserr0 != resasc0) || abserr0 == 0.0) {
        *result = result0;
        *abserr = abserr0;
        return ORC_SUCCESS;
    } else if (limit == 1) {
        *result = result0;
        *abserr = abserr0;
        return ORC_EMAXITER;
    }

    area = result0;
    errsum = abserr0;
    iteration = 1;

    do {
        double a1, b1, a2, b2, a_i, b_i, r_i, e_i;
        double area1 = 0, area2 = 0, area12 = 0;
        double error1 = 0, error2 = 0, error12 = 0;
        double resasc1, resasc2, resabs1, resabs2;
        size_t i_max = orc_ws_argmax(w);

        a_i = w->alist[i_max];
        b_i = w->blist[i_max];
        r_i = w->rlist[i_max];
        e_i = w->elist[i_max];

        a1 = a_i;
        b1 = 0.5 * (a_i + b_i);
        a2 = b1;
        b2 = b_i;

        orc_qk31(f, ctx, a1, b1, &area1, &error1, &resabs1, &resasc1);
        orc_qk31(f, ctx, a2, b2, &area2, &error2, &resabs2, &resasc2);
        w->n_rule_applications += 2;

        area12 = area1 + area2;
        error12 = error1 + error2;

        errsum += (error12 - e_i);
        area += area12 - r_i;

        if (resasc1 != error1 && resasc2 != error2) {
            double delta = r_i - area12;

            if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i)
                roundoff_type1++;
            if (iteration >= 10 && error12 > e_i)
                roundoff_type2++;
        }

        tolerance = ORC_GSL_MAX(epsabs, epsrel * fabs(area));
        if (w->trace)
            fprintf(stderr, "  qag it %zu bisect [%.15g, %.15g] r_old %.6g e_old %.3g -> r (%.6g, %.6g) e (%.3g, %.3g) | area %.8g errsum %.3g tol %.3g\n",
                    iteration, a_i, b_i, r_i, e_i, area1, area2, error1, error2, area, errsum, tolerance);

        if (errsum > tolerance) {
            if (roundoff_type1 >= 6 || roundoff_type2 >= 20)
                error_type = 2; /* round off error */

            /* set error flag in the case of bad integrand behaviour at
               a point of the integration range */
            if (orc_subinterval_too_small(a1, a2, b2))
                error_type = 3;
        }

        /* replace the bisected interval by its halves (larger error first, as
           QUADPACK's update does) */
        if (error2 > error1) {
            w->alist[i_max] = a2;
            w->blist[i_max] = b2;
            w->rlist[i_max] = area2;
            w->elist[i_max] = error2;
            w->alist[w->size] = a1;
            w->blist[w->size] = b1;
            w->rlist[w->size] = area1;
            w->elist[w->size] = error1;
        } else {
            w->alist[i_max] = a1;
            w->blist[i_max] = b1;
            w->rlist[i_max] = area1;
            w->elist[i_max] = error1;
            w->alist[w->size] = a2;
            w->blist[w->size] = b2;
            w->rlist[w->size] = area2;
            w->elist[w->size] = error2;
        }
        w->size++;
        if (w->size > w->max_size_seen)
            w->max_size_seen = w->size;

        iteration++;
    } while (iteration < limit && !error_type && errsum > tolerance);

    {
        double sum = 0;
        for (i = 0; i < w->size; i++)
            sum += w->rlist[i];
        *result = sum;
    }
    *abserr = errsum;

    if (errsum <= tolerance)
        return ORC_SUCCESS;
    else if (error_type == 2)
        return ORC_EROUND;
    else if (error_type == 3)
        return ORC_ESING;
    else if (iteration == limit)
        return ORC_EMAXITER;
    else
        return ORC_EFAILED;
}

static inline void orc_central_deriv(orc_fn f, void *ctx, double x, double h,
                                     double *result, double *abserr_round, double *abserr_trunc)
{
    /* 5-point rule (x-h, x-h/2, x, x+h/2, x+h); the central point is not used. */
    double fm1 = f(x - h, ctx);
    double fp1 = f(x + h, ctx);
    double fmh = f(x - h / 2, ctx);
    double fph = f(x + h / 2, ctx);

    double r3 = 0.5 * (fp1 - fm1);
    double r5 = (4.0 / 3.0) * (fph - fmh) - (1.0 / 3.0) * r3;

    double e3 = (fabs(fp1) + fabs(fm1)) * DBL_EPSILON;
    double e5 = 2.0 * (fabs(fph) + fabs(fmh)) * DBL_EPSILON + e3;

    double dy = ORC_GSL_MAX(fabs(r3 / h), fabs(r5 / h)) * (fabs(x) / h) * DBL_EPSILON;

    *result = r5 / h;
    *abserr_trunc = fabs((r5 - r3) / h);
    *abserr_round = fabs(e5 / h) + dy;
}

/* gsl_deriv_central: always succeeds. */
static inline int orc_deriv_central(orc_fn f, void *ctx, double x, double h,
                                    double *result, double *abserr)
{
    double r_0, round, trunc, error;
    orc_central_deriv(f, ctx, x, h, &r_0, &round, &trunc);
    error = round + trunc;

    if (round < trunc && (round > 0 && trunc > 0)) {
        double r_opt, round_opt, trunc_opt, error_opt;
        double h_opt = h * pow(round / (2.0 * trunc), 1.0 / 3.0);
        orc_central_deriv(f, ctx, x, h_opt, &r_opt, &round_opt, &trunc_opt);
        error_opt = round_opt + trunc_opt;

        if (error_opt < error && fabs(r_opt - r_0) < 4.0 * error) {
            r_0 = r_opt;
            error = error_opt;
        }
    }

    *result = r_0;
    *abserr = error;
    return ORC_SUCCESS;
}

#endif
