/* Stub for the oracle build: nothing from gsl_permute.h is used. */
