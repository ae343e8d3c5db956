/* Stub for the oracle build: GSL is absent.  Only the type and allocator
 * used by suzerain/bsmbsm.c (suzerain_bsmbsm_permutation) are provided. */
#ifndef GSL_PERMUTATION_H_STUB
#define GSL_PERMUTATION_H_STUB
#include <stdlib.h>
typedef struct { size_t size; size_t *data; } gsl_permutation;
static inline gsl_permutation *gsl_permutation_alloc(size_t n)
{
    gsl_permutation *p = (gsl_permutation *) malloc(sizeof(*p));
    if (!p) return NULL;
    p->size = n;
    p->data = (size_t *) malloc((n ? n : 1) * sizeof(size_t));
    if (!p->data) { free(p); return NULL; }
    return p;
}
static inline void gsl_permutation_free(gsl_permutation *p)
{
    if (p) { free(p->data); free(p); }
}
#endif
