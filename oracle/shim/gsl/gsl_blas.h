/* TEST INFRASTRUCTURE ONLY (oracle shim) -- gslcblas reference-loop semantics, see gsl_shim.c. */
#ifndef ORACLE_GSL_BLAS_H
#define ORACLE_GSL_BLAS_H
#include "gsl_matrix.h"
#ifdef __cplusplus
extern "C" {
#endif
int gsl_blas_dgemm(CBLAS_TRANSPOSE_t TransA, CBLAS_TRANSPOSE_t TransB, double alpha,
                   const gsl_matrix* A, const gsl_matrix* B, double beta, gsl_matrix* C);
int gsl_blas_dgemv(CBLAS_TRANSPOSE_t TransA, double alpha, const gsl_matrix* A,
                   const gsl_vector* X, double beta, gsl_vector* Y);
int gsl_blas_ddot(const gsl_vector* X, const gsl_vector* Y, double* result);
double gsl_blas_dnrm2(const gsl_vector* X);
#ifdef __cplusplus
}
#endif
#endif
