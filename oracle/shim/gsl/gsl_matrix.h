/* TEST INFRASTRUCTURE ONLY (oracle shim) -- GSL matrix API subset used by the reference. */
#ifndef ORACLE_GSL_MATRIX_H
#define ORACLE_GSL_MATRIX_H
#include "gsl_shim_types.h"
#include "gsl_vector.h"
#ifdef __cplusplus
extern "C" {
#endif
gsl_matrix* gsl_matrix_alloc(size_t n1, size_t n2);
void gsl_matrix_free(gsl_matrix* m);
int gsl_matrix_memcpy(gsl_matrix* dst, const gsl_matrix* src);
_gsl_matrix_view gsl_matrix_submatrix(gsl_matrix* m, size_t i, size_t j, size_t n1, size_t n2);
int gsl_matrix_sub(gsl_matrix* a, const gsl_matrix* b);
int gsl_matrix_add(gsl_matrix* a, const gsl_matrix* b);
int gsl_matrix_add_constant(gsl_matrix* a, double x);
_gsl_vector_view gsl_matrix_column(gsl_matrix* m, size_t j);
_gsl_vector_view gsl_matrix_row(gsl_matrix* m, size_t i);
double* gsl_matrix_ptr(gsl_matrix* m, size_t i, size_t j);
double gsl_matrix_get(const gsl_matrix* m, size_t i, size_t j);
void gsl_matrix_set(gsl_matrix* m, size_t i, size_t j, double x);
_gsl_matrix_const_view gsl_matrix_const_view_array(const double* base, size_t n1, size_t n2);
_gsl_matrix_view gsl_matrix_view_array(double* base, size_t n1, size_t n2);
void gsl_matrix_set_identity(gsl_matrix* m);
void gsl_matrix_set_zero(gsl_matrix* m);
int gsl_matrix_transpose(gsl_matrix* m);
int gsl_matrix_transpose_memcpy(gsl_matrix* dst, const gsl_matrix* src);
#ifdef __cplusplus
}
#endif
#endif
