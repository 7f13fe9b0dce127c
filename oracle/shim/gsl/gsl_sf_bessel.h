/* oracle/shim/gsl/gsl_sf_bessel.h -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference's leung-bessel/src/bessel.c includes <gsl/gsl_sf_bessel.h>
 * only for gsl_sf_bessel_Jn(int n, double x) (bessel.c:19, 249, 333): exact
 * integer-order J_n for n < 30.  GSL is not in this image; glibc's jn() is the
 * same mathematical function to double precision, so the shim maps one onto
 * the other and lets bessel.c compile unmodified, in place.
 */
#ifndef ORACLE_SHIM_GSL_SF_BESSEL_H
#define ORACLE_SHIM_GSL_SF_BESSEL_H
#ifndef _DEFAULT_SOURCE
#define _DEFAULT_SOURCE 1
#endif
#include <math.h>
extern double jn(int, double);
static inline double gsl_sf_bessel_Jn(int n, double x) { return jn(n, x); }
#endif
