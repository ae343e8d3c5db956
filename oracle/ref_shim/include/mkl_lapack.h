/* Stub for the oracle build: Intel MKL is absent from this image.  Maps the
 * Fortran-ABI names the reference's lapack.c calls onto the LP64 OpenBLAS that
 * ships inside SciPy (symbols scipy_<name>_).  Declarations are deliberately
 * unprototyped (K&R) because every call site passes by reference.
 * Test infrastructure only; generated once by hand-run script in the
 * builder's session and committed. */
#ifndef MKL_LAPACK_STUB
#define MKL_LAPACK_STUB
extern void   scipy_cgbcon_();
#define cgbcon scipy_cgbcon_
extern void   scipy_cgbrfs_();
#define cgbrfs scipy_cgbrfs_
extern void   scipy_cgbsv_();
#define cgbsv scipy_cgbsv_
extern void   scipy_cgbsvx_();
#define cgbsvx scipy_cgbsvx_
extern void   scipy_cgbtrf_();
#define cgbtrf scipy_cgbtrf_
extern void   scipy_cgbtrs_();
#define cgbtrs scipy_cgbtrs_
extern void   scipy_clacpy_();
#define clacpy scipy_clacpy_
extern void   scipy_dgbcon_();
#define dgbcon scipy_dgbcon_
extern void   scipy_dgbrfs_();
#define dgbrfs scipy_dgbrfs_
extern void   scipy_dgbsv_();
#define dgbsv scipy_dgbsv_
extern void   scipy_dgbsvx_();
#define dgbsvx scipy_dgbsvx_
extern void   scipy_dgbtrf_();
#define dgbtrf scipy_dgbtrf_
extern void   scipy_dgbtrs_();
#define dgbtrs scipy_dgbtrs_
extern void   scipy_dlacpy_();
#define dlacpy scipy_dlacpy_
extern void   scipy_sgbcon_();
#define sgbcon scipy_sgbcon_
extern void   scipy_sgbrfs_();
#define sgbrfs scipy_sgbrfs_
extern void   scipy_sgbsv_();
#define sgbsv scipy_sgbsv_
extern void   scipy_sgbsvx_();
#define sgbsvx scipy_sgbsvx_
extern void   scipy_sgbtrf_();
#define sgbtrf scipy_sgbtrf_
extern void   scipy_sgbtrs_();
#define sgbtrs scipy_sgbtrs_
extern void   scipy_slacpy_();
#define slacpy scipy_slacpy_
extern void   scipy_zgbcon_();
#define zgbcon scipy_zgbcon_
extern void   scipy_zgbrfs_();
#define zgbrfs scipy_zgbrfs_
extern void   scipy_zgbsv_();
#define zgbsv scipy_zgbsv_
extern void   scipy_zgbsvx_();
#define zgbsvx scipy_zgbsvx_
extern void   scipy_zgbtrf_();
#define zgbtrf scipy_zgbtrf_
extern void   scipy_zgbtrs_();
#define zgbtrs scipy_zgbtrs_
extern void   scipy_zlacpy_();
#define zlacpy scipy_zlacpy_
extern double scipy_dlamch_();
#define dlamch scipy_dlamch_
extern double scipy_dlangb_();
#define dlangb scipy_dlangb_
extern double scipy_zlangb_();
#define zlangb scipy_zlangb_
extern float  scipy_slamch_();
#define slamch scipy_slamch_
extern float  scipy_slangb_();
#define slangb scipy_slangb_
extern float  scipy_clangb_();
#define clangb scipy_clangb_
#endif
