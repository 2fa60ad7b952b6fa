/* TEST INFRASTRUCTURE ONLY (oracle shim). */
#ifndef ORACLE_GSL_PERMUTATION_H
#define ORACLE_GSL_PERMUTATION_H
#include "gsl_shim_types.h"
#ifdef __cplusplus
extern "C" {
#endif
gsl_permutation* gsl_permutation_alloc(size_t n);
void gsl_permutation_free(gsl_permutation* p);
#ifdef __cplusplus
}
#endif
#endif
