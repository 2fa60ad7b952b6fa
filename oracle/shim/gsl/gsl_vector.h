/* TEST INFRASTRUCTURE ONLY (oracle shim) -- GSL vector API subset used by the reference. */
#ifndef ORACLE_GSL_VECTOR_H
#define ORACLE_GSL_VECTOR_H
#include "gsl_shim_types.h"
#ifdef __cplusplus
extern "C" {
#endif
gsl_vector* gsl_vector_alloc(size_t n);
void gsl_vector_free(gsl_vector* v);
int gsl_vector_memcpy(gsl_vector* dst, const gsl_vector* src);
double* gsl_vector_ptr(gsl_vector* v, size_t i);
double gsl_vector_get(const gsl_vector* v, size_t i);
void gsl_vector_set(gsl_vector* v, size_t i, double x);
void gsl_vector_set_zero(gsl_vector* v);
int gsl_vector_add_constant(gsl_vector* v, double x);
double gsl_vector_max(const gsl_vector* v);
double gsl_vector_min(const gsl_vector* v);
_gsl_vector_view gsl_vector_view_array(double* base, size_t n);
_gsl_vector_const_view gsl_vector_const_view_array(const double* base, size_t n);
#ifdef __cplusplus
}
#endif
#endif
