/* Stub for the oracle build: GSL is absent.  suzerain/bsplineop.h only needs
 * the workspace type name in prototypes; bsplineop.c itself is NOT compiled
 * (operator construction is redone in oracle/bspline.py and validated
 * against the reference's golden collocation matrices). */
#ifndef GSL_BSPLINE_H_STUB
#define GSL_BSPLINE_H_STUB
#include <stddef.h>
typedef struct gsl_bspline_workspace_stub gsl_bspline_workspace;
typedef struct gsl_bspline_deriv_workspace_stub gsl_bspline_deriv_workspace;
typedef struct gsl_matrix_stub gsl_matrix;
typedef struct gsl_vector_stub gsl_vector;
#endif
