/*
 * suzerain_b200.h -- C ABI of the B200-native implicit wall-normal operator path.
 *
 * Drop-in boundary for Suzerain's SMR91 hybrid implicit/explicit substep
 * (SURVEY.md section 8b).  Plain C: pointers, ints, doubles; complex numbers are
 * pairs of doubles (re, im), binary compatible with C99 `double _Complex`,
 * C++ `std::complex<double>` and the reference's `complex_double`.
 *
 * Every entry point names the reference interface it replaces (file:line
 * relative to the reference tree).  Three layers are exported:
 *
 *   1. per-pencil HOST-pointer functions with the reference's own signatures
 *      (thin wrappers: copy in, launch, copy out) -- what the reference's unit
 *      tests call;
 *   2. batched DEVICE-pointer functions (one launch for all local (kx,kz)
 *      pencils) -- what a GPU-resident time stepper calls;
 *   3. whole-field HOST-pointer functions mirroring the three virtuals of
 *      operator_hybrid_isothermal -- what apps/perfect calls through
 *      lowstorage::linear_operator; these include H2D/D2H copies.
 *
 * All functions return 0 on success, <0 for an invalid argument (-k = k-th
 * argument, LAPACK convention, suzerain/blas_et_al/blas.c:68-74), >0 for a
 * numerical failure (singular pivot, 1-based row, as zgbtrf), and
 * SZB_ECUDA (-1000 - cudaError) for CUDA runtime failures.  There is no CPU
 * fallback: without a usable CUDA device every compute entry point fails.
 */
#ifndef SUZERAIN_B200_H
#define SUZERAIN_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SZB_ECUDA_BASE (-1000)

typedef struct szb_complex { double re, im; } szb_complex;

/* ------------------------------------------------------------------------ *
 * Band-storage index algebra.  Replaces suzerain/gbmatrix.h:51-68.
 * ------------------------------------------------------------------------ */
int szb_gbmatrix_offset(int ld, int kl, int ku, int i, int j);
int szb_gbmatrix_in_band(int kl, int ku, int i, int j);

/* ------------------------------------------------------------------------ *
 * BSMBSM structure and permutation.  Replaces suzerain/bsmbsm.h:104-186.
 * Field-for-field identical to `suzerain_bsmbsm`.
 * ------------------------------------------------------------------------ */
typedef struct szb_bsmbsm {
    int S, n, kl, ku, ld, N, KL, KU, LD;
} szb_bsmbsm;

szb_bsmbsm szb_bsmbsm_construct(int S, int n, int kl, int ku); /* bsmbsm.h:130-149 */
int szb_bsmbsm_q   (int S, int n, int i);                      /* bsmbsm.h:162-169 */
int szb_bsmbsm_qinv(int S, int n, int i);                      /* bsmbsm.h:182-186 */

/* y <- alpha P x + beta y ('N') or alpha P^T x + beta y ('T') on nbatch
 * contiguous length-S*n device vectors.  Replaces suzerain_bsmbsm_zaPxpby
 * (suzerain/bsmbsm_aPxpby_complex.def:37-336); x must not alias y. */
int szb_bsmbsm_zaPxpby_batch(char trans, int S, int n,
                             const double alpha[2], const szb_complex *d_x,
                             const double beta[2],        szb_complex *d_y,
                             int nbatch, void *stream);

/* ------------------------------------------------------------------------ *
 * B-spline collocation operators.  Replaces suzerain_bsplineop_workspace /
 * suzerain_bsplineop_alloc (suzerain/bsplineop.h:125-180,
 * suzerain/bsplineop.c:101-201,403-652) for
 * SUZERAIN_BSPLINEOP_COLLOCATION_GREVILLE, together with the breakpoint ->
 * knot/Greville bookkeeping the reference obtains from GSL.
 * ------------------------------------------------------------------------ */
typedef struct szb_bsplineop szb_bsplineop;

int  szb_bsplineop_alloc(int k, int nbreak, const double *breakpoints,
                         int nderiv, szb_bsplineop **out);
void szb_bsplineop_free(szb_bsplineop *w);
int  szb_bsplineop_k     (const szb_bsplineop *w);
int  szb_bsplineop_n     (const szb_bsplineop *w);
int  szb_bsplineop_nderiv(const szb_bsplineop *w);
int  szb_bsplineop_kl    (const szb_bsplineop *w, int d);
int  szb_bsplineop_ku    (const szb_bsplineop *w, int d);
int  szb_bsplineop_max_kl(const szb_bsplineop *w);
int  szb_bsplineop_max_ku(const szb_bsplineop *w);
int  szb_bsplineop_ld    (const szb_bsplineop *w);
/* Host pointer equivalent to the reference's w->D_T[d] (bsplineop.h:163-187):
 * transposed operator, general band storage, leading dimension ld, already
 * stepped past unused super-diagonals. */
const double *szb_bsplineop_D_T(const szb_bsplineop *w, int d);
/* Greville abscissae (collocation points), n doubles. */
int  szb_bsplineop_greville(const szb_bsplineop *w, double *xi);
/* Build from caller-supplied operator storage ((nderiv+1) blocks of ld*n
 * doubles laid out as the reference's single calloc'd block,
 * bsplineop.c:163-187) instead of evaluating the basis. */
int  szb_bsplineop_from_storage(int k, int n, int nderiv, const int *kl,
                                const int *ku, const double *storage,
                                szb_bsplineop **out);

/* Grid stretching used to place breakpoints (suzerain/htstretch.c:40-53,
 * 112-125; suzerain/support/support.cpp:288-300). */
double szb_htstretch1(double delta, double L, double x);
double szb_htstretch2(double delta, double L, double x);

/* y <- alpha D^(d) x + beta y for nrhs contiguous real or complex pencils on
 * the device.  Replaces suzerain_bsplineop_accumulate{,_complex}
 * (suzerain/bsplineop.c:222-297) as batched by operator_tools.hpp:77-116. */
int szb_bsplineop_accumulate_complex_batch(const szb_bsplineop *w, int d, int nrhs,
        const double alpha[2], const szb_complex *d_x, size_t ldx,
        const double beta[2],        szb_complex *d_y, size_t ldy, void *stream);

/* The other three members of the family (suzerain/bsplineop.h:246-295, 360-367; bsplineop.c:222-258, 299-381), same
 * batching: x <- alpha D^(d) x in place for complex and for real pencils (the reference's scratch copy is the
 * kernel's staging buffer) and y <- alpha D^(d) x + beta y for real pencils.  alpha and beta are real here, as in
 * the reference.  Return values follow the argument positions of the complex call. */
int szb_bsplineop_apply_complex_batch(const szb_bsplineop *w, int d, int nrhs, double alpha,
        szb_complex *d_x, size_t ldx, void *stream);
int szb_bsplineop_accumulate_batch(const szb_bsplineop *w, int d, int nrhs, double alpha,
        const double *d_x, size_t ldx, double beta, double *d_y, size_t ldy, void *stream);
int szb_bsplineop_apply_batch(const szb_bsplineop *w, int d, int nrhs, double alpha,
        double *d_x, size_t ldx, void *stream);

/* ------------------------------------------------------------------------ *
 * Linearised perfect-gas operator (M + phi L).  Replaces
 * suzerain/rholut_imexop.h:66-469.  Struct layouts are field-for-field those
 * of suzerain_rholut_imexop_{scenario,ref,refld}.
 * ------------------------------------------------------------------------ */
typedef struct szb_rholut_imexop_scenario {
    double Re, Pr, Ma, alpha, gamma;
} szb_rholut_imexop_scenario;

#define SZB_NREF 26
typedef struct szb_rholut_imexop_ref {
    double *ux, *uy, *uz, *u2, *uxux, *uxuy, *uxuz, *uyuy, *uyuz, *uzuz,
           *nu, *nuux, *nuuy, *nuuz, *nuu2, *nuuxux, *nuuxuy, *nuuxuz,
           *nuuyuy, *nuuyuz, *nuuzuz, *ex_gradrho, *ey_gradrho, *ez_gradrho,
           *e_divm, *e_deltarho;
} szb_rholut_imexop_ref;

typedef struct szb_rholut_imexop_refld {
    int ux, uy, uz, u2, uxux, uxuy, uxuz, uyuy, uyuz, uzuz,
        nu, nuux, nuuy, nuuz, nuu2, nuuxux, nuuxuy, nuuxuz,
        nuuyuy, nuuyuz, nuuzuz, ex_gradrho, ey_gradrho, ez_gradrho,
        e_divm, e_deltarho;
} szb_rholut_imexop_refld;

/* Isothermal wall data consumed by IsothermalPATPTEnforcer
 * (apps/perfect/operator_hybrid_isothermal.cpp:396-526): wall temperature and
 * velocities from specification_isothermal, and which walls are enforced. */
typedef struct szb_isothermal {
    int    enforce_lower, enforce_upper;
    double lower_T, lower_u, lower_v, lower_w;
    double upper_T, upper_u, upper_v, upper_w;
} szb_isothermal;

/* Solver specification.  Replaces specification_zgbsv
 * (suzerain/specification_zgbsv.cpp:46-124). */
enum { SZB_SOLVER_ZGBSV = 0, SZB_SOLVER_ZCGBSVX = 1, SZB_SOLVER_ZGBSVX = 2 };
typedef struct szb_zgbsv_spec {
    int    method;   /* SZB_SOLVER_*                                   */
    int    aiter;    /* zcgbsvx: iterations before stagnation test (1)  */
    int    diter;    /* zcgbsvx: max double-precision refinements  (5)  */
    double tolsc;    /* zcgbsvx: 0 => absolute eps tolerance       (0)  */
    int    equil;    /* zgbsvx: equilibrate (FACT = 'E')           (0)  */
    int    reuse;    /* zcgbsvx: a HINT here -- every pencil is factored
                      * afresh on the device (the reference reuses the
                      * previous wavenumber's factors as a preconditioner
                      * and refines to the same criterion)          (0)  */
    int    siter;    /* zcgbsvx: a HINT here -- no single-precision
                      * factorisation is attempted                  (-1) */
} szb_zgbsv_spec;
szb_zgbsv_spec szb_zgbsv_spec_default(void);        /* zcgbsvx defaults */

/* Device-resident operator context: operators D_T[0..2], reference
 * profiles, scenario, wall data, optional NRBC matrices. */
typedef struct szb_imexop szb_imexop;

int  szb_imexop_create(const szb_bsplineop *w, szb_imexop **out);
void szb_imexop_destroy(szb_imexop *op);
int  szb_imexop_set_scenario(szb_imexop *op, const szb_rholut_imexop_scenario *s);
/* Gathers the 26 strided host profiles (references.cpp:50-108 exports stride
 * 42) into a dense device table. */
int  szb_imexop_set_refs(szb_imexop *op, const szb_rholut_imexop_ref *r,
                         const szb_rholut_imexop_refld *ld);
/* The same from a DEVICE copy of the reference's `references` block (42 rows x Ny, column-major,
 * leading dimension ld >= 31: apps/perfect/references.hpp:82-125), e.g. the buffer a sharded stepper
 * has just summed over ranks (MPI_Allreduce of apps/perfect/perfect.cpp:1397 -> ncclAllReduce): rows
 * q::u .. q::e_deltarho are the 26 profiles, in the order of references::rholut_imexop.  Asynchronous
 * on `stream`; no host copy. */
#define SZB_REFERENCES_ROWS  42
#define SZB_REFERENCES_FIRST 5
int  szb_imexop_set_refs_device(szb_imexop *op, const double *d_references, int ld, void *stream);
/* The producer of that block: collect_references (apps/perfect/perfect.cpp:1266-1400).  Sums the 42 quantities
 * of apps/perfect/references.hpp:83-128 over the (z, x) points of the `ny` local wall-normal planes [y0, y0 + ny)
 * of the physical-space state d_sphys (fields e, mx, my, mz, rho in ndx order, each [ny][nzx] doubles, field
 * stride `field_stride` doubles: the physical_view of suzerain/physical_view.hpp:44-87), multiplies by `scale`
 * and writes d_refs (42 x Ny, column-major, ld = 42); the columns of planes this rank does not own are zeroed
 * (:1275-1277), so that an all-reduce(SUM) over ranks completes the profile.  Single rank: scale = chi =
 * 1 / (dNx dNz) (suzerain/pencil_grid.hpp:159-163); several ranks: scale = 1, all-reduce, then scale by chi
 * (:1396-1399).  top_is_inviscid: mu = lambda = 0 on the global plane Ny - 1 (one-sided grids, :1287-1290).
 * beta is the viscosity exponent of definition_scenario; Re and Pr of `scenario` are not used.
 * d_workspace: `*workspace_needed` bytes of device scratch (query with d_sphys = d_refs = NULL).
 * Deterministic (no atomics); asynchronous on `stream`. */
int  szb_collect_references_device(const szb_rholut_imexop_scenario *scenario, double beta, int Ny, int y0, int ny,
        size_t nzx, const double *d_sphys, size_t field_stride, int top_is_inviscid, double scale,
        double *d_refs, void *d_workspace, size_t workspace_bytes, size_t *workspace_needed, void *stream);
int  szb_imexop_set_isothermal(szb_imexop *op, const szb_isothermal *iso);
/* 5x5 column-major Giles matrices (upper_nrbc_{a,b,c},
 * operator_hybrid_isothermal.cpp:771-777); any may be NULL. */
int  szb_imexop_set_nrbc(szb_imexop *op, const double *a, const double *b,
                         const double *c);
szb_bsmbsm szb_imexop_bsmbsm(const szb_imexop *op);
/* Which linearisation the implicit operator uses (apps/perfect/linearize_type.hpp, read at
 * operator_hybrid_isothermal.cpp:136-241, 276-374, 608-761): rhome_xyz (default) or rhome_y, the
 * wavenumber-independent operator of suzerain_rholut_imexop_{accumulate,packc,packf}00
 * (suzerain/rholut_imexop.h:209-238, 330-368, 447-469): accumulate / apply then use km = kn = 0
 * for every pencil and invert factors ONE operator and solves every pencil with it. */
enum { SZB_LINEARIZE_RHOME_XYZ = 0, SZB_LINEARIZE_RHOME_Y = 1 };
int  szb_imexop_set_linearization(szb_imexop *op, int linearization);

/* Batched, device pointers.  Batch entry p works on the pencil whose field f
 * starts at
 *   base + slot(p)*pencil_stride + f*field_stride  (complex elements, y stride 1)
 * with slot(p) = d_index ? d_index[p] : p, so that the active (non-dealiased)
 * pencils of a state can be processed without compaction.
 * d_km/d_kn: npencil wavenumbers on the device, indexed by batch entry. */

/* out <- (M + phi L) in + beta out.  Replaces the loop body of
 * operator_hybrid_isothermal::accumulate_mass_plus_scaled_operator
 * (operator_hybrid_isothermal.cpp:306-334) over suzerain_rholut_imexop_accumulate
 * (rholut_imexop.c:43-547).  in == out is allowed only with beta == 0 and
 * identical strides (apply_mass_plus_scaled_operator, :103-241). */
int szb_imexop_accumulate_batch(const szb_imexop *op, const double phi[2],
        int npencil, const double *d_km, const double *d_kn, const int *d_index,
        const szb_complex *d_in, size_t in_field_stride, size_t in_pencil_stride,
        const double beta[2],
        szb_complex *d_out, size_t out_field_stride, size_t out_pencil_stride,
        void *stream);

/* Assemble P (M + phi L)^T P^T in LAPACK band storage for every pencil.
 * Replaces suzerain_rholut_imexop_pack{c,f} (rholut_imexop.def:41-597) and,
 * when with_bc != 0, IsothermalPATPTEnforcer::op.  packf != 0: LD+KL rows per
 * column with the matrix offset by KL rows (LU-ready); else LD rows.
 * Pencil p's matrix starts at d_patpt + p*N*rows. */
int szb_imexop_pack_batch(const szb_imexop *op, const double phi[2],
        int npencil, const double *d_km, const double *d_kn,
        int packf, int with_bc, szb_complex *d_patpt, void *stream);

/* state <- P^T (P (M + phi L) P^T with wall BCs)^{-1} P state, in place, for
 * every pencil: assemble + BCs + permute + factor + solve + permute back.
 * Replaces the hot loop of invert_mass_plus_scaled_operator
 * (operator_hybrid_isothermal.cpp:617-686).  nextra additional right hand
 * sides per pencil (integral constraints, :676-685) are solved against the
 * same factorisation: pencil p's c-th extra RHS is d_extra + (p*nextra+c)*N.
 * d_ipiv (optional, npencil*N ints) receives LAPACK 1-based pivots.
 * d_info (npencil ints) receives per-pencil zgbtrf-style info.
 * d_iters (optional, npencil ints) receives the zcgbsvx diter counter. */
int szb_imexop_invert_batch(const szb_imexop *op, const szb_zgbsv_spec *spec,
        const double phi[2], int npencil, const double *d_km, const double *d_kn,
        const int *d_index,
        szb_complex *d_state, size_t field_stride, size_t pencil_stride,
        int nextra, szb_complex *d_extra,
        int *d_ipiv, int *d_info, int *d_iters, void *stream);

/* Bytes of device scratch the batched invert holds (for capacity planning). */
size_t szb_imexop_workspace_bytes(const szb_imexop *op);

/* a <-> b for the listed pencils (null: all) of two device states in different layouts: the
 * exchange of an interleaved_state with a contiguous_state that lowstorage::step performs after
 * every accumulate (suzerain/lowstorage.hpp:1511, suzerain/state.hpp:486-520,607-630). */
int szb_state_exchange(int npencil, const int *d_index, int S, int n,
                       szb_complex *d_a, size_t a_field_stride, size_t a_pencil_stride,
                       szb_complex *d_b, size_t b_field_stride, size_t b_pencil_stride, void *stream);

/* Zero-fill the listed pencils (dealiased / Nyquist modes,
 * operator_hybrid_isothermal.cpp:632-637). */
int szb_zero_pencils(int npencil, const int *d_index, int S, int n,
        szb_complex *d_state, size_t field_stride, size_t pencil_stride,
        void *stream);

/* ------------------------------------------------------------------------ *
 * Batched banded LU and solve on pre-assembled systems: the bsmbsm_solver
 * protocol (suzerain/bsmbsm_solver.hpp:70-330).  LAPACK band storage,
 * system b at d_ab + b*stride.  Pivot rule: first maximum of |re|+|im|
 * (izamax), row interchanges as zgbtf2, ipiv 1-based.
 * ------------------------------------------------------------------------ */
/* Replaces suzerain_lapack_zgbtrf (suzerain/blas_et_al/lapack.c:185-197). */
int szb_zgbtrf_batch(int n, int kl, int ku, szb_complex *d_ab, int ldab,
                     size_t stride, int *d_ipiv, int *d_info, int nbatch,
                     void *stream);
/* Replaces suzerain_lapack_zgbtrs (lapack.c:261-277); trans in {'N','T'}. */
int szb_zgbtrs_batch(char trans, int n, int kl, int ku, int nrhs,
                     const szb_complex *d_ab, int ldab, size_t stride,
                     const int *d_ipiv, szb_complex *d_b, int ldb,
                     size_t strideb, int nbatch, void *stream);
/* Replaces suzerain_lapackext_zcgbsvx for fact='N', siter<0
 * (suzerain/blas_et_al/dsgbsvx.def:71-318): d_ab is the unfactored matrix
 * (ldab = kl+1+ku), d_afb receives the factors (2kl+1+ku rows). */
int szb_zcgbsvx_batch(char trans, int n, int kl, int ku, int aiter, int diter,
                      double tolsc,
                      const szb_complex *d_ab, size_t stride_ab,
                      szb_complex *d_afb, size_t stride_afb, int *d_ipiv,
                      const szb_complex *d_b, szb_complex *d_x,
                      int *d_iters, double *d_res, int *d_info, int nbatch,
                      void *stream);

/* bsmbsm_solver::solve for ONE system on HOST storage laid out as the reference's solver object keeps it
 * (suzerain/bsmbsm_solver.hpp:70-330, .cpp:58-78): lu is (LD+KL) x N column-major; zgbsv works in place (the
 * matrix sits in rows KL.. of lu, pb is overwritten by the solution, papt / px unused); zcgbsvx reads the
 * unfactored papt (LD x N), writes the factors to lu and the solution to px, one refined solve per right hand
 * side (bsmbsm_solver.cpp:396-401).  Returns zgbtrf's info.  include/suzerain_b200_solver.hpp wraps it in the
 * reference's supply_B -> supplied_PAPT -> solve -> demand_X protocol. */
int szb_bsmbsm_solver_solve(const szb_bsmbsm *A, const szb_zgbsv_spec *spec, char trans, int nrhs,
                            szb_complex *lu, const szb_complex *papt, int *ipiv,
                            szb_complex *pb, szb_complex *px, int *iters, double *res);

/* ------------------------------------------------------------------------ *
 * Per-pencil HOST-pointer wrappers with the reference's signatures.
 * ------------------------------------------------------------------------ */
/* suzerain_rholut_imexop_accumulate (rholut_imexop.h:186-207; positions are
 * E,u,v,w,rho as rholut_imexop.c:52-62). */
int szb_rholut_imexop_accumulate(const double phi[2], double km, double kn,
        const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
        const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
        const szb_complex *in_rho_E, const szb_complex *in_rho_u,
        const szb_complex *in_rho_v, const szb_complex *in_rho_w,
        const szb_complex *in_rho, const double beta[2],
        szb_complex *out_rho_E, szb_complex *out_rho_u, szb_complex *out_rho_v,
        szb_complex *out_rho_w, szb_complex *out_rho,
        const double *a, const double *b, const double *c);
/* suzerain_rholut_imexop_pack{c,f} (rholut_imexop.h:323-469) with the
 * ordering rho_E=0, rho_u=1, rho_v=2, rho_w=3, rho=4 the application uses
 * (operator_hybrid_isothermal.cpp:644-653); `buf` is not needed. */
int szb_rholut_imexop_packc(const double phi[2], double km, double kn,
        const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
        const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
        szb_bsmbsm *A_T, szb_complex *patpt,
        const double *a, const double *b, const double *c);
int szb_rholut_imexop_packf(const double phi[2], double km, double kn,
        const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
        const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
        szb_bsmbsm *A_T, szb_complex *patpt,
        const double *a, const double *b, const double *c);

/* suzerain_rholut_imexop_{accumulate,packc,packf}00 (rholut_imexop.h:209-238, 330-368,
 * 447-469): the km = kn = 0 special cases used by linearize::rhome_y. */
int szb_rholut_imexop_accumulate00(const double phi[2],
        const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
        const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
        const szb_complex *in_rho_E, const szb_complex *in_rho_u,
        const szb_complex *in_rho_v, const szb_complex *in_rho_w,
        const szb_complex *in_rho, const double beta[2],
        szb_complex *out_rho_E, szb_complex *out_rho_u, szb_complex *out_rho_v,
        szb_complex *out_rho_w, szb_complex *out_rho, const double *c);
int szb_rholut_imexop_packc00(const double phi[2],
        const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
        const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
        szb_bsmbsm *A_T, szb_complex *patpt, const double *c);
int szb_rholut_imexop_packf00(const double phi[2],
        const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
        const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
        szb_bsmbsm *A_T, szb_complex *patpt, const double *c);

/* ------------------------------------------------------------------------ *
 * Whole-field HOST-pointer entry points: the three virtuals of
 * operator_hybrid_isothermal (apps/perfect/operator_hybrid_isothermal.hpp:
 * 112-140) as called through lowstorage::linear_operator
 * (suzerain/lowstorage.hpp:631-681).
 * ------------------------------------------------------------------------ */
/* Local wave-space extents, as read from specification_grid / pencil_grid at
 * operator_hybrid_isothermal.cpp:120-134. */
typedef struct szb_wavegrid {
    int    Nx, dNx, dkbx, dkex;   /* grid.N.x, grid.dN.x, local_wave_{start,end}.x */
    int    Nz, dNz, dkbz, dkez;
    double Lx, Lz;                /* grid.L.x, grid.L.z */
} szb_wavegrid;

/* Number of local pencils / of those that are active (not dealiased). */
int szb_wavegrid_npencils(const szb_wavegrid *g);
int szb_wavegrid_nactive (const szb_wavegrid *g);
/* Fills km, kn and an active flag per local pencil in state order (x fastest,
 * then z); km/kn follow operator_hybrid_isothermal.cpp:53-65,617-630. */
int szb_wavegrid_wavenumbers(const szb_wavegrid *g, double *km, double *kn,
                             int *active);

/* Wave-space differentiation of Ny-long pencils stored [kz][kx][y] on the device:
 *   apply:      x <- alpha (i kx)^dxcnt (i kz)^dzcnt x
 *   accumulate: y <- alpha (i kx)^dxcnt (i kz)^dzcnt x + beta y
 * with dealiased and Nyquist modes zeroed (scaled by beta).  Replace
 * suzerain_diffwave_apply / suzerain_diffwave_accumulate (suzerain/diffwave.c:65-129,
 * 131-198; the wave-space half of the nonlinear operator, navier_stokes.hpp:227-331). */
int szb_diffwave_apply_batch(int dxcnt, int dzcnt, const double alpha[2],
        szb_complex *d_x, const szb_wavegrid *g, int Ny, void *stream);
int szb_diffwave_accumulate_batch(int dxcnt, int dzcnt, const double alpha[2],
        const szb_complex *d_x, const double beta[2], szb_complex *d_y,
        const szb_wavegrid *g, int Ny, void *stream);

/* `state` is the interleaved state [5][Ny][Nx_loc][Nz_loc] with strides
 * (Ny, 1, 5*Ny, 5*Ny*Nx_loc) (suzerain/storage.hpp:235-236). */
int szb_operator_apply_mass_plus_scaled_operator(const szb_imexop *op,
        const szb_wavegrid *g, const double phi[2], szb_complex *state);
/* `output` is contiguous state: field stride out_field_stride, pencil (m,n)
 * at out + f*out_field_stride + m*Ny + n*Ny*Nx_loc (storage.hpp:267-268). */
int szb_operator_accumulate_mass_plus_scaled_operator(const szb_imexop *op,
        const szb_wavegrid *g, const double phi[2], const szb_complex *input,
        const double beta[2], szb_complex *output, size_t out_field_stride);
/* ic0: optional nconstraints right hand sides of 5*Ny complex each, solved on
 * the (0,0) pencil's factorisation when it is local (:676-685). */
int szb_operator_invert_mass_plus_scaled_operator(const szb_imexop *op,
        const szb_zgbsv_spec *spec, const szb_wavegrid *g, const double phi[2],
        szb_complex *state, int nconstraints, szb_complex *ic0,
        int *first_bad_pencil);

/* Library/device probes. */
int         szb_device_count(void);
const char *szb_version(void);
/* Number of kernel launches issued by this library since process start. */
unsigned long long szb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SUZERAIN_B200_H */
