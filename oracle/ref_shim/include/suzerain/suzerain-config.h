/* Stub written for the oracle build (oracle/Makefile): stands in for the
 * autoconf-generated suzerain-config.h, which cannot be generated here
 * (no autotools).  Test infrastructure only. */
#ifndef SUZERAIN_CONFIG_H_STUB
#define SUZERAIN_CONFIG_H_STUB
#define SUZERAIN_HAVE_MKL 1
#define SUZERAIN_BLAS_ALIGNMENT 64
#endif
