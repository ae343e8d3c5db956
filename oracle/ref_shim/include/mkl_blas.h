/* Stub for the oracle build: Intel MKL is absent from this image.  Maps the
 * Fortran-ABI names the reference's blas.c calls onto the LP64 OpenBLAS that
 * ships inside SciPy (symbols scipy_<name>_).  Declarations are deliberately
 * unprototyped (K&R) because every call site passes by reference.
 * Test infrastructure only; generated once by hand-run script in the
 * builder's session and committed. */
#ifndef MKL_BLAS_STUB
#define MKL_BLAS_STUB
extern void   scipy_caxpby_();
#define caxpby scipy_caxpby_
extern void   scipy_caxpy_();
#define caxpy scipy_caxpy_
extern void   scipy_ccopy_();
#define ccopy scipy_ccopy_
extern void   scipy_cgbmv_();
#define cgbmv scipy_cgbmv_
extern void   scipy_cscal_();
#define cscal scipy_cscal_
extern void   scipy_cswap_();
#define cswap scipy_cswap_
extern void   scipy_daxpby_();
#define daxpby scipy_daxpby_
extern void   scipy_daxpy_();
#define daxpy scipy_daxpy_
extern void   scipy_dcopy_();
#define dcopy scipy_dcopy_
extern void   scipy_dgbmv_();
#define dgbmv scipy_dgbmv_
extern void   scipy_dsbmv_();
#define dsbmv scipy_dsbmv_
extern void   scipy_dscal_();
#define dscal scipy_dscal_
extern void   scipy_dswap_();
#define dswap scipy_dswap_
extern void   scipy_saxpby_();
#define saxpby scipy_saxpby_
extern void   scipy_saxpy_();
#define saxpy scipy_saxpy_
extern void   scipy_scopy_();
#define scopy scipy_scopy_
extern void   scipy_sgbmv_();
#define sgbmv scipy_sgbmv_
extern void   scipy_ssbmv_();
#define ssbmv scipy_ssbmv_
extern void   scipy_sscal_();
#define sscal scipy_sscal_
extern void   scipy_sswap_();
#define sswap scipy_sswap_
extern void   scipy_zaxpby_();
#define zaxpby scipy_zaxpby_
extern void   scipy_zaxpy_();
#define zaxpy scipy_zaxpy_
extern void   scipy_zcopy_();
#define zcopy scipy_zcopy_
extern void   scipy_zgbmv_();
#define zgbmv scipy_zgbmv_
extern void   scipy_zscal_();
#define zscal scipy_zscal_
extern void   scipy_zswap_();
#define zswap scipy_zswap_
extern double scipy_dasum_();
#define dasum scipy_dasum_
extern double scipy_ddot_();
#define ddot scipy_ddot_
extern double scipy_dnrm2_();
#define dnrm2 scipy_dnrm2_
extern double scipy_dzasum_();
#define dzasum scipy_dzasum_
extern double scipy_dznrm2_();
#define dznrm2 scipy_dznrm2_
extern float  scipy_sasum_();
#define sasum scipy_sasum_
extern float  scipy_sdot_();
#define sdot scipy_sdot_
extern float  scipy_snrm2_();
#define snrm2 scipy_snrm2_
extern float  scipy_scasum_();
#define scasum scipy_scasum_
extern float  scipy_scnrm2_();
#define scnrm2 scipy_scnrm2_
/* index-of-extremum helpers used directly (not through BLAS_FUNC) by blas.c:378-470 */
extern int    scipy_isamax_();
#define isamax scipy_isamax_
extern int    scipy_idamax_();
#define idamax scipy_idamax_
extern int    scipy_icamax_();
#define icamax scipy_icamax_
extern int    scipy_izamax_();
#define izamax scipy_izamax_
extern int    scipy_isamin_();
#define isamin scipy_isamin_
extern int    scipy_idamin_();
#define idamin scipy_idamin_
extern int    scipy_icamin_();
#define icamin scipy_icamin_
extern int    scipy_izamin_();
#define izamin scipy_izamin_
/* MKL returns complex dot products through a leading pointer argument;
 * gfortran-built OpenBLAS returns them by value.  Adapt in ref_glue.c. */
extern void ref_shim_cdotc(void *r, const int *n, const void *x, const int *incx, const void *y, const int *incy);
extern void ref_shim_zdotc(void *r, const int *n, const void *x, const int *incx, const void *y, const int *incy);
#define cdotc ref_shim_cdotc
#define zdotc ref_shim_zdotc
/* xerbla: keep errors inside the process rather than OpenBLAS's exit(). */
extern void ref_shim_xerbla(const char *srname, const int *info, const int len);
#define xerbla ref_shim_xerbla
#endif
