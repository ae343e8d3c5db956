/* Stub of <gsl/gsl_sf_pow_int.h> for compiling the reference's suzerain/diffwave.c without GSL
 * (test infrastructure only).  gsl_sf_pow_int is restated from its published algorithm
 * (GSL 2.8 specfunc/pow_int.c: binary powering, x -> 1/x for negative n). */
#ifndef ORACLE_SHIM_GSL_SF_POW_INT_H
#define ORACLE_SHIM_GSL_SF_POW_INT_H
static inline double gsl_sf_pow_int(double x, int n)
{
    double value = 1.0;
    if (n < 0) { n = -n; x = 1.0 / x; }
    do {
        if (n & 1) value *= x;
        n >>= 1;
        x *= x;
    } while (n);
    return value;
}
#endif
