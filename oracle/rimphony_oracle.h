/* oracle/rimphony_oracle.h -- TEST INFRASTRUCTURE ONLY (see rimphony_oracle.c). */
#ifndef RIMPHONY_ORACLE_H
#define RIMPHONY_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same numbering as include/rimphony_b200.h */
enum { ORC_DIST_POWER_LAW = 0, ORC_DIST_THERMAL_JUETTNER = 1, ORC_DIST_PITCHY_PL = 2, ORC_DIST_PITCHY_KAPPA = 3 };
enum { ORC_COEFF_EMISSION = 0, ORC_COEFF_ABSORPTION = 1, ORC_COEFF_FARADAY = 2 };
enum { ORC_STOKES_I = 0, ORC_STOKES_Q = 1, ORC_STOKES_V = 2 };

typedef struct {
    int kind;
    double p, k;
    double gamma_min, gamma_max, inv_gamma_cutoff;
    double kappa, width, inv_kappa_width;
    double neg_inverse_t;
    double norm;
} orc_dist;

typedef struct {
    uint64_t n_symphony_integrand;
    uint64_t n_gamma_qag;
    uint64_t n_gamma_qag_failed;
    uint64_t n_chunks;
    uint64_t max_gamma_intervals;
    uint64_t max_n_intervals;
    uint64_t n_heyvaerts_element;
    uint64_t n_heyvaerts_qag;
    uint64_t n_heyvaerts_jy;
    double hey_nr_val, hey_qr_val;
    /* where a Heyvaerts calculation failed (diagnostics for the parity study; 0 = it did not):
     * stage 1 NR centre, 2 NR right march, 3 NR left march, 4 QR march; the outer QAG's status;
     * the status, outer variable and kind (0 NR, 1 QR) of the FIRST inner QAG that failed */
    int hey_fail_stage, hey_fail_outer_status, hey_fail_inner_status, hey_fail_inner_kind;
    double hey_fail_lo, hey_fail_hi, hey_fail_inner_var;
    /* the first non-finite element value: kind (0 none, 1 NR, 2 QR), its coordinates and value */
    int hey_nan_kind;
    double hey_nan_sigma, hey_nan_pomega, hey_nan_x, hey_nan_gamma, hey_nan_mu, hey_nan_value;
} orc_stats;

int orc_dist_init(orc_dist *d, int kind, const double *params, int n_params);
double orc_calc_f(const orc_dist *d, double gamma, double cos_xi);
void orc_calc_f_derivatives(const orc_dist *d, double gamma, double cos_xi, double *dfdg, double *dfdcx);

double orc_symphony(const orc_dist *d, int coeff, int stokes, double s, double theta, orc_stats *stats);
double orc_symphony_lobes(const orc_dist *d, int coeff, int stokes, double s, double theta,
                          orc_stats *stats, double lobes[2]);
double orc_symphony_diagnostic(const orc_dist *d, int coeff, int stokes, double s, double theta,
                               int what, double a, double b);
double orc_heyvaerts(const orc_dist *d, int stokes, double s, double theta, orc_stats *stats);

double orc_compute_dimensionless(const orc_dist *d, int coeff, int stokes, double s, double theta,
                                 orc_stats *stats);
double orc_compute_cgs(const orc_dist *d, int coeff, int stokes, double nu, double b, double n_e,
                       double theta, orc_stats *stats);
void orc_compute_all_dimensionless(const orc_dist *d, double s, double theta, double out[8],
                                   double lobes[4], orc_stats *stats);
int orc_batch_compute_all_dimensionless(int kind, int64_t n_points, const double *s,
                                        const double *theta, const double *const *params,
                                        int n_params, unsigned coeff_mask, double *out8,
                                        double *lobes4, int n_threads);
int orc_num_threads(void);
/* stability studies only (tests/golden/make_stability.py); 0 restores the reference's 1e-3 */
void orc_set_epsrel(double symphony_epsrel, double heyvaerts_epsrel);
/* debug: print the bisections of the Heyvaerts outer QAG calls to stderr */
void orc_set_trace(int heyvaerts_outer);

double orc_test_hey_outer_integrand(const orc_dist *d, int stokes, double s, double theta, int which,
                                    double outer_var, orc_stats *stats);
double orc_ref_bessel_j(double n, double x);
double orc_ref_bessel_dj(double n, double x);
double orc_test_bessel_i(double nu, double x);
void orc_test_bessel_jy(double nu, double x, double *j, double *y);
double orc_test_bessel_k2(double z);
double orc_test_pitch_angle_integral(double k);
int orc_test_qag(int which, double a_param, double lo, double hi, double epsrel, double *result,
                 double *abserr, int *n_intervals);
double orc_test_deriv(int which, double a_param, double x, double h);
void orc_gk31_tables(double xgk[16], double wgk[16], double wg[8]);

#ifdef __cplusplus
}
#endif
#endif
