/* suzerain_b200_dropin.h -- the reference's own per-pencil entry points, name for name and
 * argument for argument, served by the B200 kernels (libsuzerain_b200_dropin.so).
 *
 * These are the six symbols apps/perfect/operator_hybrid_isothermal.cpp and
 * tests/test_rholut_imexop{,00}.cpp bind (suzerain/rholut_imexop.h:186-207, 209-238, 323-345,
 * 347-368, 425-445, 447-469): a caller built against the reference headers links against this
 * library instead of suzerain/rholut_imexop.c without touching its source.
 *
 *   - complex_double arguments are passed BY VALUE exactly as in the reference (C99
 *     `double _Complex`; in C++ a struct of two doubles has the same SysV x86-64 / AAPCS64
 *     calling convention);
 *   - the five scalar-ordering integers are accepted; the assembled operator is produced for the
 *     ordering the application uses, rho_E=0, rho_u=1, rho_v=2, rho_w=3, rho=4
 *     (operator_hybrid_isothermal.cpp:644-653) -- anything else aborts through the reference's
 *     error convention (message + abort, suzerain/error.c:40-56);
 *   - `buf` is accepted and ignored (the device kernels need no host scratch);
 *   - the B-spline workspace is the reference's struct (suzerain/bsplineop.h:125-180), read in
 *     place: k, n, nderiv, kl[], ku[], max_kl, max_ku, ld and the band storage D_T[0..2].
 *
 * Define SZB_DROPIN_USE_REFERENCE_TYPES after including the reference's own
 * <suzerain/rholut_imexop.h> to have the prototypes below re-declare the reference's functions
 * with the reference's types: the compiler then rejects any mismatch (tests/test_dropin_abi.py).
 */
#ifndef SUZERAIN_B200_DROPIN_H
#define SUZERAIN_B200_DROPIN_H

#include "suzerain_b200.h"

#ifdef SZB_DROPIN_USE_REFERENCE_TYPES
typedef complex_double                       szb_dropin_complex;
typedef suzerain_rholut_imexop_scenario      szb_dropin_scenario;
typedef suzerain_rholut_imexop_ref           szb_dropin_ref;
typedef suzerain_rholut_imexop_refld         szb_dropin_refld;
typedef suzerain_bsplineop_workspace         szb_dropin_bsplineop_workspace;
typedef suzerain_bsmbsm                      szb_dropin_bsmbsm;
#else
#if defined(__cplusplus)
typedef szb_complex                          szb_dropin_complex;      /* two doubles: same by-value ABI as double _Complex */
#else
typedef double _Complex                      szb_dropin_complex;
#endif
typedef szb_rholut_imexop_scenario           szb_dropin_scenario;
typedef szb_rholut_imexop_ref                szb_dropin_ref;
typedef szb_rholut_imexop_refld              szb_dropin_refld;
typedef szb_bsmbsm                           szb_dropin_bsmbsm;
/* suzerain_bsplineop_workspace, field for field (suzerain/bsplineop.h:125-180) */
typedef struct szb_dropin_bsplineop_workspace {
    int      method;        /* enum suzerain_bsplineop_method */
    int      k, n, nderiv;
    int     *kl, *ku;
    int      max_kl, max_ku, ld;
    double **D_T;
} szb_dropin_bsplineop_workspace;
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* suzerain/rholut_imexop.h:186-207.  The header names the inputs E,w,v,u,rho; positions are
 * E,u,v,w,rho (rholut_imexop.c:52-62). */
void suzerain_rholut_imexop_accumulate(
        const szb_dropin_complex phi, const double km, const double kn,
        const szb_dropin_scenario *s, const szb_dropin_ref *r, const szb_dropin_refld *ld,
        const szb_dropin_bsplineop_workspace *w,
        const szb_dropin_complex *in_rho_E, const szb_dropin_complex *in_rho_w,
        const szb_dropin_complex *in_rho_v, const szb_dropin_complex *in_rho_u,
        const szb_dropin_complex *in_rho, const szb_dropin_complex beta,
        szb_dropin_complex *out_rho_E, szb_dropin_complex *out_rho_u, szb_dropin_complex *out_rho_v,
        szb_dropin_complex *out_rho_w, szb_dropin_complex *out_rho,
        const double *a, const double *b, const double *c);

/* suzerain/rholut_imexop.h:209-238 */
void suzerain_rholut_imexop_accumulate00(
        const szb_dropin_complex phi,
        const szb_dropin_scenario *s, const szb_dropin_ref *r, const szb_dropin_refld *ld,
        const szb_dropin_bsplineop_workspace *w,
        const szb_dropin_complex *in_rho_E, const szb_dropin_complex *in_rho_w,
        const szb_dropin_complex *in_rho_v, const szb_dropin_complex *in_rho_u,
        const szb_dropin_complex *in_rho, const szb_dropin_complex beta,
        szb_dropin_complex *out_rho_E, szb_dropin_complex *out_rho_u, szb_dropin_complex *out_rho_v,
        szb_dropin_complex *out_rho_w, szb_dropin_complex *out_rho,
        const double *c);

/* suzerain/rholut_imexop.h:323-345 */
void suzerain_rholut_imexop_packc(
        const szb_dropin_complex phi, const double km, const double kn,
        const szb_dropin_scenario *s, const szb_dropin_ref *r, const szb_dropin_refld *ld,
        const szb_dropin_bsplineop_workspace *w,
        const int rho_E, const int rho_u, const int rho_v, const int rho_w, const int rho,
        szb_dropin_complex *buf, szb_dropin_bsmbsm *A_T, szb_dropin_complex *patpt,
        const double *a, const double *b, const double *c);

/* suzerain/rholut_imexop.h:347-368 */
void suzerain_rholut_imexop_packc00(
        const szb_dropin_complex phi,
        const szb_dropin_scenario *s, const szb_dropin_ref *r, const szb_dropin_refld *ld,
        const szb_dropin_bsplineop_workspace *w,
        const int rho_E, const int rho_u, const int rho_v, const int rho_w, const int rho,
        szb_dropin_complex *buf, szb_dropin_bsmbsm *A_T, szb_dropin_complex *patpt,
        const double *c);

/* suzerain/rholut_imexop.h:425-445 */
void suzerain_rholut_imexop_packf(
        const szb_dropin_complex phi, const double km, const double kn,
        const szb_dropin_scenario *s, const szb_dropin_ref *r, const szb_dropin_refld *ld,
        const szb_dropin_bsplineop_workspace *w,
        const int rho_E, const int rho_u, const int rho_v, const int rho_w, const int rho,
        szb_dropin_complex *buf, szb_dropin_bsmbsm *A_T, szb_dropin_complex *patpt,
        const double *a, const double *b, const double *c);

/* suzerain/rholut_imexop.h:447-469 */
void suzerain_rholut_imexop_packf00(
        const szb_dropin_complex phi,
        const szb_dropin_scenario *s, const szb_dropin_ref *r, const szb_dropin_refld *ld,
        const szb_dropin_bsplineop_workspace *w,
        const int rho_E, const int rho_u, const int rho_v, const int rho_w, const int rho,
        szb_dropin_complex *buf, szb_dropin_bsmbsm *A_T, szb_dropin_complex *patpt,
        const double *c);

#ifdef __cplusplus
}
#endif

#endif
