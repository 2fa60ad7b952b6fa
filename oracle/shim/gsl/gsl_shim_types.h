/* TEST INFRASTRUCTURE ONLY -- part of oracle/ (see oracle/README.md).
 *
 * Minimal re-statement of the GSL 2.x public types that the reference's
 * obvious::Matrix / obvious::Vector wrapper touches directly
 * (reference: src/obcore/math/linalg/gsl/Matrix.h:154-160 `gsl_matrix* _M`,
 *  src/obcore/math/linalg/gsl/Matrix.cpp:21-22 `_M->size1/_M->size2`,
 *  :248 `col.vector.data`).  Field names follow GSL's documented structs.
 * GSL itself is NOT installed in this image and NOT vendored by the reference
 * (CMakeLists.txt:91-92 links system `gsl gslcblas`, version unpinned).
 */
#ifndef ORACLE_GSL_SHIM_TYPES_H
#define ORACLE_GSL_SHIM_TYPES_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { size_t size; double* data; } gsl_block;

typedef struct { size_t size; size_t stride; double* data; gsl_block* block; int owner; } gsl_vector;
typedef struct { gsl_vector vector; } _gsl_vector_view;
typedef _gsl_vector_view gsl_vector_view;
typedef struct { gsl_vector vector; } _gsl_vector_const_view;
typedef const _gsl_vector_const_view gsl_vector_const_view;

typedef struct { size_t size1; size_t size2; size_t tda; double* data; gsl_block* block; int owner; } gsl_matrix;
typedef struct { gsl_matrix matrix; } _gsl_matrix_view;
typedef _gsl_matrix_view gsl_matrix_view;
typedef struct { gsl_matrix matrix; } _gsl_matrix_const_view;
typedef const _gsl_matrix_const_view gsl_matrix_const_view;

typedef struct { size_t size; size_t* data; } gsl_permutation;

enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
typedef enum CBLAS_TRANSPOSE CBLAS_TRANSPOSE_t;

#define GSL_SUCCESS 0
#define GSL_DBL_EPSILON 2.2204460492503131e-16
#define GSL_DBL_MIN 2.2250738585072014e-308

#ifdef __cplusplus
}
#endif
#endif
