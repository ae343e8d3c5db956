/* Stub for the oracle build (MKL absent). */
#ifndef MKL_TYPES_STUB
#define MKL_TYPES_STUB
#include <complex.h>
typedef int MKL_INT;
typedef float _Complex MKL_Complex8;
typedef double _Complex MKL_Complex16;
#endif
