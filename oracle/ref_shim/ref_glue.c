/*
 * ref_glue.c -- TEST INFRASTRUCTURE ONLY (oracle/_ref build).
 *
 * Thin C harness linked together with the *unmodified* reference sources
 * (compiled from where they lie under /root/reference by oracle/Makefile)
 * into oracle/_ref/libsuzerain_ref.so.  It exists so that tests and the
 * bench.py cpu_baseline / --impl reference legs can drive the reference's own
 * per-pencil loop body without the C++ application layer (Boost, Eigen, MPI,
 * ESIO, log4cxx are absent here).
 *
 * What is restated here rather than compiled from the reference:
 *   - the loop body of operator_hybrid_isothermal::invert_mass_plus_scaled_operator
 *     (apps/perfect/operator_hybrid_isothermal.cpp:617-686) and of
 *     accumulate_mass_plus_scaled_operator (:306-334);
 *   - IsothermalPATPTEnforcer::{op,rhs} (same file, :396-526), a C++ class
 *     template private to that translation unit;
 *   - bsmbsm_solver::{supply_B,demand_X} (suzerain/bsmbsm_solver.hpp:150-156,
 *     :274-280) which are one call each to suzerain_bsmbsm_zaPxpby;
 *   - bsmbsm_solver_zgbsv::solve_hook / bsmbsm_solver_zcgbsvx::solve_hook
 *     (suzerain/bsmbsm_solver.cpp:155-182, :377-414).
 * Everything arithmetic (assembly, pack, permutation, banded LU, refinement,
 * banded mat-vecs) is the reference's own object code or LAPACK (OpenBLAS
 * inside SciPy; MKL is unavailable in this image).
 *
 * Nothing under suzerain_b200/ may link or load this library.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <suzerain/common.h>
#include <suzerain/blas_et_al.h>
#include <suzerain/bsmbsm.h>
#include <suzerain/gbmatrix.h>
#include <suzerain/rholut_imexop.h>

/* ---- MKL-convention adapters named by the stub mkl_blas.h ---- */
extern float  _Complex scipy_cdotc_(const int *, const void *, const int *, const void *, const int *);
extern double _Complex scipy_zdotc_(const int *, const void *, const int *, const void *, const int *);
extern void scipy_openblas_set_num_threads(int);

void ref_shim_cdotc(void *r, const int *n, const void *x, const int *incx,
                    const void *y, const int *incy)
{ *(float _Complex *) r = scipy_cdotc_(n, x, incx, y, incy); }

void ref_shim_zdotc(void *r, const int *n, const void *x, const int *incx,
                    const void *y, const int *incy)
{ *(double _Complex *) r = scipy_zdotc_(n, x, incx, y, incy); }

static int ref_shim_last_xerbla = 0;
void ref_shim_xerbla(const char *srname, const int *info, const int len)
{
    ref_shim_last_xerbla = *info;
    fprintf(stderr, "ref_shim xerbla: %.*s info=%d\n", len, srname ? srname : "", *info);
}
int ref_last_xerbla(void) { int t = ref_shim_last_xerbla; ref_shim_last_xerbla = 0; return t; }

void ref_set_blas_threads(int n) { scipy_openblas_set_num_threads(n); }

/* ---- Wall boundary-condition description (flat mirror of the data the
 * enforcer's constructor derives; operator_hybrid_isothermal.cpp:419-461) ---- */
typedef struct ref_bc {
    int    enforce_lower;   /* always 1 in the app (:599-601)          */
    int    enforce_upper;   /* grid.two_sided()                        */
    double E_factor[2];     /* T_w/(g(g-1)) + Ma^2/2 |u_w|^2  (:448-455) */
    double vel_factor[2][3];/* wall u,v,w                     (:456-461) */
} ref_bc;

/* Rewrites up to 8 columns of PA^TP^T; follows :470-510.  ndx order is
 * e=0,mx=1,my=2,mz=3,rho=4 (apps/perfect's ndx::). */
static void enforcer_op(const suzerain_bsmbsm *A, const ref_bc *bc,
                        complex_double *patpt, int ld)
{
    const int wall_pt[2] = { 0, A->n - 1 };
    const int wb = bc->enforce_lower ? 0 : 1, we = bc->enforce_upper ? 2 : 1;
    for (int wall = wb; wall < we; ++wall) {
        const int irho = suzerain_bsmbsm_qinv(A->S, A->n, 4*A->n + wall_pt[wall]);
        for (int eqn = 0; eqn < 4; ++eqn) {
            const int ieq = suzerain_bsmbsm_qinv(A->S, A->n, eqn*A->n + wall_pt[wall]);
            const double factor = eqn == 0 ? bc->E_factor[wall]
                                           : bc->vel_factor[wall][eqn - 1];
            int begin, end;
            complex_double * const col = (complex_double *) suzerain_gbmatrix_col(
                    A->N, A->N, A->KL, A->KU, (void *) patpt, ld,
                    sizeof(complex_double), ieq, &begin, &end);
            if (col[ieq] == 0) col[ieq] = 1;
            const complex_double scaling = col[ieq];
            for (int i = begin; i < end; ++i) {
                col[i] = i == ieq  ? +scaling
                       : i == irho ? -scaling*factor
                       :             0;
            }
        }
    }
}

static void enforcer_rhs(const suzerain_bsmbsm *A, const ref_bc *bc,
                         complex_double *b)
{
    const int wall_pt[2] = { 0, A->n - 1 };
    const int wb = bc->enforce_lower ? 0 : 1, we = bc->enforce_upper ? 2 : 1;
    for (int wall = wb; wall < we; ++wall)
        for (int eqn = 0; eqn < 4; ++eqn)
            b[suzerain_bsmbsm_qinv(A->S, A->n, eqn*A->n + wall_pt[wall])] = 0;
}

/* Exposed for unit tests of the restated enforcer. */
void ref_enforcer_op(int S, int n, int kl, int ku, const ref_bc *bc,
                     complex_double *patpt, int ld)
{
    suzerain_bsmbsm A = suzerain_bsmbsm_construct(S, n, kl, ku);
    enforcer_op(&A, bc, patpt, ld);
}

/* ---- assembled operator for one wavenumber: packc (+ optional BCs) ---- */
int ref_assemble(const double phi[2], double km, double kn,
                 const suzerain_rholut_imexop_scenario *s,
                 const suzerain_rholut_imexop_ref *r,
                 const suzerain_rholut_imexop_refld *ld,
                 const suzerain_bsplineop_workspace *w,
                 const ref_bc *bc, /* NULL: skip BCs */
                 const double *a, const double *b, const double *c,
                 int packf,        /* 0: LD rows; 1: LD+KL rows, offset KL */
                 complex_double *out)
{
    suzerain_bsmbsm A = suzerain_bsmbsm_construct(5, w->n, w->max_kl, w->max_ku);
    const int nbuf = A.ld*A.n > 75 ? A.ld*A.n : 75;
    complex_double *buf = (complex_double *) malloc(nbuf*sizeof(*buf));
    if (!buf) return -1;
    const complex_double cphi = phi[0] + _Complex_I*phi[1];
    if (packf) {
        suzerain_rholut_imexop_packf(cphi, km, kn, s, r, ld, w, 0, 1, 2, 3, 4,
                                     buf, &A, out, a, b, c);
        if (bc) enforcer_op(&A, bc, out + A.KL, A.LD + A.KL);
    } else {
        suzerain_rholut_imexop_packc(cphi, km, kn, s, r, ld, w, 0, 1, 2, 3, 4,
                                     buf, &A, out, a, b, c);
        if (bc) enforcer_op(&A, bc, out, A.LD);
    }
    free(buf);
    return 0;
}

/* ---- invert loop body over a batch of active pencils ----
 * state: npencil contiguous pencils of 5*n complex (interleaved-state pencil,
 *        field stride n), solved in place.
 * solver: 0 = zgbsv (in-place zgbtrf + zgbtrs), 1 = zcgbsvx with the
 *        reference default spec (reuse=false, aiter=1, siter=-1, diter=5,
 *        tolsc=0; specification_zgbsv.cpp:46-54).
 * ipiv_out (may be NULL): npencil*N LAPACK 1-based pivots.
 * iters_out (may be NULL): npencil ints, diter reported by zcgbsvx.
 * nextra: additional right hand sides solved per pencil against the same
 *        factorisation (the integral-constraint columns, :676-685);
 *        extra is npencil*nextra*N complex, in place.
 * Returns the first nonzero info (or 0); first_bad gets its pencil index. */
int ref_invert_batch(int solver, const double phi[2],
                     const suzerain_rholut_imexop_scenario *s,
                     const suzerain_rholut_imexop_ref *r,
                     const suzerain_rholut_imexop_refld *ld,
                     const suzerain_bsplineop_workspace *w,
                     const ref_bc *bc,
                     const double *a, const double *b, const double *c,
                     int npencil, const double *km, const double *kn,
                     complex_double *state,
                     int nextra, complex_double *extra,
                     int *ipiv_out, int *iters_out,
                     int nthreads, int *first_bad)
{
    const suzerain_bsmbsm A0 = suzerain_bsmbsm_construct(5, w->n, w->max_kl, w->max_ku);
    const complex_double cphi = phi[0] + _Complex_I*phi[1];
    int rc = 0, bad = -1;
    if (nthreads < 1) nthreads = 1;

#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        suzerain_bsmbsm A = A0;
        const int N = A.N, ldlu = A.KL + A.LD;
        const int nbuf = A.ld*A.n > 75 ? A.ld*A.n : 75;
        complex_double *buf  = (complex_double *) malloc(nbuf*sizeof(*buf));
        complex_double *LU   = (complex_double *) malloc((size_t) ldlu*N*sizeof(*LU));
        complex_double *PAPT = (complex_double *) malloc((size_t) A.LD*N*sizeof(*PAPT));
        complex_double *PB   = (complex_double *) malloc(N*sizeof(*PB));
        complex_double *PX   = (complex_double *) malloc(N*sizeof(*PX));
        complex_double *R    = (complex_double *) malloc(N*sizeof(*R));
        int            *ipiv = (int *) malloc(N*sizeof(int));

#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int p = 0; p < npencil; ++p) {
            complex_double * const x = state + (size_t) p*N;
            int info = 0;
            char fact = 'N';
            int apprx = 0;
            double afrob = -1;
            if (solver == 0) {
                suzerain_rholut_imexop_packf(cphi, km[p], kn[p], s, r, ld, w,
                        0, 1, 2, 3, 4, buf, &A, LU, a, b, c);
                enforcer_op(&A, bc, LU + A.KL, ldlu);
            } else {
                suzerain_rholut_imexop_packc(cphi, km[p], kn[p], s, r, ld, w,
                        0, 1, 2, 3, 4, buf, &A, PAPT, a, b, c);
                enforcer_op(&A, bc, PAPT, A.LD);
            }
            for (int rhs = 0; rhs <= nextra && !info; ++rhs) {
                complex_double * const v = rhs == 0 ? x
                    : extra + ((size_t) p*nextra + (rhs - 1))*N;
                suzerain_bsmbsm_zaPxpby('N', A.S, A.n, 1, v, 1, 0, PB, 1);
                enforcer_rhs(&A, bc, PB);
                if (solver == 0) {
                    if (fact == 'N') {
                        info = suzerain_lapack_zgbtrf(N, N, A.KL, A.KU, LU, ldlu, ipiv);
                        fact = 'F';
                    }
                    if (!info)
                        info = suzerain_lapack_zgbtrs('T', N, A.KL, A.KU, 1,
                                                      LU, ldlu, ipiv, PB, N);
                    if (!info)
                        suzerain_bsmbsm_zaPxpby('T', A.S, A.n, 1, PB, 1, 0, v, 1);
                } else {
                    int siter = -1, diter = 5;
                    double tolsc = 0, res = 0;
                    info = suzerain_lapackext_zcgbsvx(&fact, &apprx, 1, 'T',
                            N, A.KL, A.KU, PAPT, &afrob, LU, ipiv, PB, PX,
                            &siter, &diter, &tolsc, R, &res);
                    if (!info)
                        suzerain_bsmbsm_zaPxpby('T', A.S, A.n, 1, PX, 1, 0, v, 1);
                    if (iters_out && rhs == 0) iters_out[p] = diter;
                }
            }
            if (ipiv_out) memcpy(ipiv_out + (size_t) p*N, ipiv, N*sizeof(int));
            if (info) {
#ifdef _OPENMP
#pragma omp critical
#endif
                if (bad < 0 || p < bad) { bad = p; rc = info; }
            }
        }
        free(buf); free(LU); free(PAPT); free(PB); free(PX); free(R); free(ipiv);
    }
    if (first_bad) *first_bad = bad;
    return rc;
}

/* ---- invert loop body for every solver specification of
 * apps/perfect/test_implicit_solvers.sh:24-30 ----
 * method: 0 zgbsv, 1 zcgbsvx, 2 zgbsvx (specification_zgbsv.cpp).  The pencils are taken
 * in rows of rowlen consecutive ones (one kz row of operator_hybrid_isothermal.cpp:613-688):
 * inside a row they are solved sequentially with the solver state (fact_, apprx_) carried
 * from one to the next exactly as bsmbsm_solver::{supplied_PAPT,apprx,solve} do
 * (bsmbsm_solver.cpp:80-102), so that reuse=true exercises the reference's
 * approximate-factorisation path; rows are independent (apprx(false) at :622).
 * iters_out: zcgbsvx diter / siter pairs (2 ints per pencil); stats_out (may be NULL):
 * per pencil {refactored?, berr or res}. */
int ref_invert_spec_batch(int method, int equil, int reuse, int aiter, int siter0, int diter0,
                          double tolsc0, const double phi[2],
                          const suzerain_rholut_imexop_scenario *s,
                          const suzerain_rholut_imexop_ref *r,
                          const suzerain_rholut_imexop_refld *ld,
                          const suzerain_bsplineop_workspace *w,
                          const ref_bc *bc,
                          const double *a, const double *b, const double *c,
                          int npencil, int rowlen, const double *km, const double *kn,
                          complex_double *state, int *iters_out, double *stats_out,
                          int nthreads, int *first_bad)
{
    const suzerain_bsmbsm A0 = suzerain_bsmbsm_construct(5, w->n, w->max_kl, w->max_ku);
    const complex_double cphi = phi[0] + _Complex_I*phi[1];
    int rc = 0, bad = -1;
    if (nthreads < 1) nthreads = 1;
    if (rowlen < 1) rowlen = 1;
    const int nrows = (npencil + rowlen - 1)/rowlen;
    const char default_fact = (method == 2 && equil) ? 'E' : 'N';

#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        suzerain_bsmbsm A = A0;
        const int N = A.N, ldlu = A.KL + A.LD;
        const int nbuf = A.ld*A.n > 75 ? A.ld*A.n : 75;
        complex_double *buf  = (complex_double *) malloc(nbuf*sizeof(*buf));
        complex_double *LU   = (complex_double *) malloc((size_t) ldlu*N*sizeof(*LU));
        complex_double *PAPT = (complex_double *) malloc((size_t) A.LD*N*sizeof(*PAPT));
        complex_double *PB   = (complex_double *) malloc(N*sizeof(*PB));
        complex_double *PX   = (complex_double *) malloc(N*sizeof(*PX));
        complex_double *R    = (complex_double *) malloc(2*(size_t) N*sizeof(*R));
        double         *rcw  = (double *) malloc(3*(size_t) N*sizeof(double));
        int            *ipiv = (int *) malloc(N*sizeof(int));

#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int row = 0; row < nrows; ++row) {
            char fact = default_fact, equed = 'N';
            int apprx = 0;
            double afrob = -1;
            /* "Factorization reuse will not aid us across large jumps in km" (:622) */
            for (int p = row*rowlen; p < npencil && p < (row + 1)*rowlen; ++p) {
                complex_double * const x = state + (size_t) p*N;
                int info = 0;
                if (method == 0) {
                    suzerain_rholut_imexop_packf(cphi, km[p], kn[p], s, r, ld, w,
                            0, 1, 2, 3, 4, buf, &A, LU, a, b, c);
                    enforcer_op(&A, bc, LU + A.KL, ldlu);
                } else {
                    suzerain_rholut_imexop_packc(cphi, km[p], kn[p], s, r, ld, w,
                            0, 1, 2, 3, 4, buf, &A, PAPT, a, b, c);
                    enforcer_op(&A, bc, PAPT, A.LD);
                }
                /* supplied_PAPT() */
                if (method == 1) afrob = -1;
                if (reuse) apprx = fact != default_fact; else fact = default_fact;
                suzerain_bsmbsm_zaPxpby('N', A.S, A.n, 1, x, 1, 0, PB, 1);
                enforcer_rhs(&A, bc, PB);
                if (method == 0) {
                    info = suzerain_lapack_zgbtrf(N, N, A.KL, A.KU, LU, ldlu, ipiv);
                    if (!info) info = suzerain_lapack_zgbtrs('T', N, A.KL, A.KU, 1, LU, ldlu, ipiv, PB, N);
                    if (!info) suzerain_bsmbsm_zaPxpby('T', A.S, A.n, 1, PB, 1, 0, x, 1);
                } else if (method == 2) {
                    double rcond = 0, ferr = 0, berr = 0;
                    info = suzerain_lapack_zgbsvx(fact, 'T', N, A.KL, A.KU, 1, PAPT, A.LD, LU, ldlu,
                            ipiv, &equed, rcw, rcw + N, PB, N, PX, N, &rcond, &ferr, &berr, R, rcw + 2*N);
                    fact = 'F';
                    if (!info) suzerain_bsmbsm_zaPxpby('T', A.S, A.n, 1, PX, 1, 0, x, 1);
                    if (stats_out) {
                        stats_out[2*p] = (equed == 'R' || equed == 'B') + 2*(equed == 'C' || equed == 'B');
                        stats_out[2*p + 1] = berr;
                    }
                } else {
                    int siter = siter0, diter = diter0;
                    double tolsc = tolsc0, res = 0;
                    info = suzerain_lapackext_zcgbsvx(&fact, &apprx, aiter, 'T',
                            N, A.KL, A.KU, PAPT, &afrob, LU, ipiv, PB, PX,
                            &siter, &diter, &tolsc, R, &res);
                    if (!info) suzerain_bsmbsm_zaPxpby('T', A.S, A.n, 1, PX, 1, 0, x, 1);
                    if (iters_out) { iters_out[2*p] = diter; iters_out[2*p + 1] = siter; }
                    if (stats_out) { stats_out[2*p] = apprx; stats_out[2*p + 1] = res; }
                }
                if (info) {
#ifdef _OPENMP
#pragma omp critical
#endif
                    if (bad < 0 || p < bad) { bad = p; rc = info; }
                }
            }
        }
        free(buf); free(LU); free(PAPT); free(PB); free(PX); free(R); free(rcw); free(ipiv);
    }
    if (first_bad) *first_bad = bad;
    return rc;
}

/* ---- accumulate loop body over a batch of active pencils ----
 * in/out: npencil pencils of 5*n complex each, field stride n. */
void ref_accumulate_batch(const double phi[2],
                          const suzerain_rholut_imexop_scenario *s,
                          const suzerain_rholut_imexop_ref *r,
                          const suzerain_rholut_imexop_refld *ld,
                          const suzerain_bsplineop_workspace *w,
                          const double *a, const double *b, const double *c,
                          int npencil, const double *km, const double *kn,
                          const complex_double *in, const double beta[2],
                          complex_double *out, int nthreads)
{
    const complex_double cphi  = phi[0]  + _Complex_I*phi[1];
    const complex_double cbeta = beta[0] + _Complex_I*beta[1];
    const int n = w->n;
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (int p = 0; p < npencil; ++p) {
        const complex_double *i0 = in  + (size_t) p*5*n;
        complex_double       *o0 = out + (size_t) p*5*n;
        suzerain_rholut_imexop_accumulate(cphi, km[p], kn[p], s, r, ld, w,
                i0, i0 + n, i0 + 2*n, i0 + 3*n, i0 + 4*n, cbeta,
                o0, o0 + n, o0 + 2*n, o0 + 3*n, o0 + 4*n, a, b, c);
    }
}

/* ---- plain pre-assembled banded solve (bsmbsm_solver_zgbsv protocol) ---- */
int ref_zgbsv_T(int N, int KL, int KU, complex_double *LU, int ldlu,
                int *ipiv, complex_double *B, int nrhs)
{
    int info = suzerain_lapack_zgbtrf(N, N, KL, KU, LU, ldlu, ipiv);
    if (!info) info = suzerain_lapack_zgbtrs('T', N, KL, KU, nrhs, LU, ldlu, ipiv, B, N);
    return info;
}

/* suzerain_bsplineop_accumulate_complex (suzerain/bsplineop.c:260-297) restated around the
 * reference's own suzerain_blas_zgbmv_d_z: bsplineop.c itself needs GSL (absent) for the
 * operator construction it also contains. */
void ref_bsplineop_accumulate_complex(int nderiv, int nrhs, const double alpha[2],
        const complex_double *x, int incx, int ldx, const double beta[2],
        complex_double *y, int incy, int ldy, const suzerain_bsplineop_workspace *w)
{
    const complex_double a = alpha[0] + _Complex_I * alpha[1], b = beta[0] + _Complex_I * beta[1];
    for (int j = 0; j < nrhs; ++j)
        suzerain_blas_zgbmv_d_z('T', w->n, w->n, w->kl[nderiv], w->ku[nderiv], a, w->D_T[nderiv], w->ld,
                                x + (size_t) j * ldx, incx, b, y + (size_t) j * ldy, incy);
}

/* suzerain_bsplineop_accumulate (suzerain/bsplineop.c:222-258), suzerain_bsplineop_apply (:299-337) and
 * suzerain_bsplineop_apply_complex (:339-381) restated the same way around the reference's own
 * suzerain_blas_dgbmv / suzerain_blas_dcopy: real coefficients, and the in-place forms (a scratch copy of each
 * vector; real and imaginary parts separately with stride 2 for the complex one). */
void ref_bsplineop_accumulate(int nderiv, int nrhs, double alpha, const double *x, int incx, int ldx,
        double beta, double *y, int incy, int ldy, const suzerain_bsplineop_workspace *w)
{
    for (int j = 0; j < nrhs; ++j)
        suzerain_blas_dgbmv('T', w->n, w->n, w->kl[nderiv], w->ku[nderiv], alpha, w->D_T[nderiv], w->ld,
                            x + (size_t) j * ldx, incx, beta, y + (size_t) j * ldy, incy);
}
void ref_bsplineop_apply(int nderiv, int nrhs, double alpha, double *x, int incx, int ldx,
        const suzerain_bsplineop_workspace *w)
{
    double *scratch = suzerain_blas_malloc(w->n * sizeof(double));
    for (int j = 0; j < nrhs; ++j) {
        double *x_j = x + (size_t) j * ldx;
        suzerain_blas_dcopy(w->n, x_j, incx, scratch, 1);
        suzerain_blas_dgbmv('T', w->n, w->n, w->kl[nderiv], w->ku[nderiv], alpha, w->D_T[nderiv], w->ld,
                            scratch, 1, 0.0, x_j, incx);
    }
    suzerain_blas_free(scratch);
}
void ref_bsplineop_apply_complex(int nderiv, int nrhs, double alpha, complex_double *x, int incx, int ldx,
        const suzerain_bsplineop_workspace *w)
{
    double *scratch = suzerain_blas_malloc(w->n * sizeof(double));
    for (int j = 0; j < nrhs; ++j) {
        double *xreal_j = (double *) (x + (size_t) j * ldx);
        for (int c = 0; c < 2; ++c) {
            suzerain_blas_dcopy(w->n, xreal_j + c, 2 * incx, scratch, 1);
            suzerain_blas_dgbmv('T', w->n, w->n, w->kl[nderiv], w->ku[nderiv], alpha, w->D_T[nderiv], w->ld,
                                scratch, 1, 0.0, xreal_j + c, 2 * incx);
        }
    }
    suzerain_blas_free(scratch);
}

/* by-pointer wrappers: ctypes cannot pass C99 complex by value */
void suzerain_diffwave_apply(int, int, complex_double, complex_double *, double, double, int,
                             int, int, int, int, int, int, int, int);
void suzerain_diffwave_accumulate(int, int, complex_double, const complex_double *, complex_double,
                                  complex_double *, double, double, int, int, int, int, int, int, int, int, int);
void ref_diffwave_apply(int dxcnt, int dzcnt, const double alpha[2], complex_double *x,
        double Lx, double Lz, int Ny, int Nx, int dNx, int dkbx, int dkex,
        int Nz, int dNz, int dkbz, int dkez)
{
    suzerain_diffwave_apply(dxcnt, dzcnt, alpha[0] + _Complex_I * alpha[1], x, Lx, Lz, Ny,
                            Nx, dNx, dkbx, dkex, Nz, dNz, dkbz, dkez);
}
void ref_diffwave_accumulate(int dxcnt, int dzcnt, const double alpha[2], const complex_double *x,
        const double beta[2], complex_double *y, double Lx, double Lz, int Ny,
        int Nx, int dNx, int dkbx, int dkex, int Nz, int dNz, int dkbz, int dkez)
{
    suzerain_diffwave_accumulate(dxcnt, dzcnt, alpha[0] + _Complex_I * alpha[1], x,
                                 beta[0] + _Complex_I * beta[1], y, Lx, Lz, Ny,
                                 Nx, dNx, dkbx, dkex, Nz, dNz, dkbz, dkez);
}

/* ---- linearize::rhome_y (apps/perfect/operator_hybrid_isothermal.cpp:691-761): ONE wavenumber-
 * independent operator from suzerain_rholut_imexop_packf00, enforcer, one zgbtrf; then supply_B /
 * rhs BC / zgbtrs('T') / demand_X per pencil.  state: npencil contiguous pencils of 5*n complex,
 * solved in place; ipiv_out (may be NULL): N pivots.  Returns zgbtrf's info. ---- */
int ref_invert00_batch(const double phi[2],
                       const suzerain_rholut_imexop_scenario *s,
                       const suzerain_rholut_imexop_ref *r,
                       const suzerain_rholut_imexop_refld *ld,
                       const suzerain_bsplineop_workspace *w,
                       const ref_bc *bc, const double *c,
                       int npencil, complex_double *state, int *ipiv_out)
{
    suzerain_bsmbsm A = suzerain_bsmbsm_construct(5, w->n, w->max_kl, w->max_ku);
    const complex_double cphi = phi[0] + _Complex_I*phi[1];
    const int N = A.N, ldlu = A.KL + A.LD;
    const int nbuf = A.ld*A.n > 75 ? A.ld*A.n : 75;
    complex_double *buf = (complex_double *) malloc(nbuf*sizeof(*buf));
    complex_double *LU  = (complex_double *) malloc((size_t) ldlu*N*sizeof(*LU));
    complex_double *PB  = (complex_double *) malloc(N*sizeof(*PB));
    int *ipiv = (int *) malloc(N*sizeof(int));
    suzerain_rholut_imexop_packf00(cphi, s, r, ld, w, 0, 1, 2, 3, 4, buf, &A, LU, c);
    enforcer_op(&A, bc, LU + A.KL, ldlu);
    int info = suzerain_lapack_zgbtrf(N, N, A.KL, A.KU, LU, ldlu, ipiv);
    for (int p = 0; p < npencil && !info; ++p) {
        complex_double * const x = state + (size_t) p*N;
        suzerain_bsmbsm_zaPxpby('N', A.S, A.n, 1, x, 1, 0, PB, 1);
        enforcer_rhs(&A, bc, PB);
        info = suzerain_lapack_zgbtrs('T', N, A.KL, A.KU, 1, LU, ldlu, ipiv, PB, N);
        if (!info) suzerain_bsmbsm_zaPxpby('T', A.S, A.n, 1, PB, 1, 0, x, 1);
    }
    if (ipiv_out) memcpy(ipiv_out, ipiv, N*sizeof(int));
    free(buf); free(LU); free(PB); free(ipiv);
    return info;
}

/* suzerain_rholut_imexop_accumulate00 over a batch (operator_hybrid_isothermal.cpp rhome_y branch
 * of apply / accumulate) */
void ref_accumulate00_batch(const double phi[2],
                            const suzerain_rholut_imexop_scenario *s,
                            const suzerain_rholut_imexop_ref *r,
                            const suzerain_rholut_imexop_refld *ld,
                            const suzerain_bsplineop_workspace *w, const double *c,
                            int npencil, const complex_double *in, const double beta[2],
                            complex_double *out)
{
    const complex_double cphi  = phi[0]  + _Complex_I*phi[1];
    const complex_double cbeta = beta[0] + _Complex_I*beta[1];
    const int n = w->n;
    for (int p = 0; p < npencil; ++p) {
        const complex_double *i0 = in  + (size_t) p*5*n;
        complex_double       *o0 = out + (size_t) p*5*n;
        suzerain_rholut_imexop_accumulate00(cphi, s, r, ld, w, i0, i0 + n, i0 + 2*n, i0 + 3*n, i0 + 4*n,
                                            cbeta, o0, o0 + n, o0 + 2*n, o0 + 3*n, o0 + 4*n, c);
    }
}

/* suzerain_rholut_imexop_pack{c,f}00 into out (as ref_assemble) */
int ref_assemble00(const double phi[2],
                   const suzerain_rholut_imexop_scenario *s,
                   const suzerain_rholut_imexop_ref *r,
                   const suzerain_rholut_imexop_refld *ld,
                   const suzerain_bsplineop_workspace *w,
                   const ref_bc *bc, const double *c, int packf, complex_double *out)
{
    suzerain_bsmbsm A = suzerain_bsmbsm_construct(5, w->n, w->max_kl, w->max_ku);
    const int nbuf = A.ld*A.n > 75 ? A.ld*A.n : 75;
    complex_double *buf = (complex_double *) malloc(nbuf*sizeof(*buf));
    if (!buf) return -1;
    const complex_double cphi = phi[0] + _Complex_I*phi[1];
    if (packf) {
        suzerain_rholut_imexop_packf00(cphi, s, r, ld, w, 0, 1, 2, 3, 4, buf, &A, out, c);
        if (bc) enforcer_op(&A, bc, out + A.KL, A.LD + A.KL);
    } else {
        suzerain_rholut_imexop_packc00(cphi, s, r, ld, w, 0, 1, 2, 3, 4, buf, &A, out, c);
        if (bc) enforcer_op(&A, bc, out, A.LD);
    }
    free(buf);
    return 0;
}
