/* TEST INFRASTRUCTURE ONLY -- part of oracle/ (see oracle/README.md).
 *
 * Functional re-statement of the subset of GSL 2.x + gslcblas that the
 * reference's obvious::Matrix/Vector wrapper calls
 * (reference: src/obcore/math/linalg/gsl/Matrix.cpp, Vector.cpp).
 * GSL is a third-party dependency that is absent from /root/reference and from
 * this image (CMakeLists.txt:91-92 `gsl gslcblas`, docker/Dockerfile:21-22
 * `libgsl-dev` on Ubuntu 22.04 => most likely GSL 2.7.1; version UNPINNED by
 * the reference).  What is restated here is GSL's *published* algorithm:
 *
 *  - dgemm / dgemv / ddot / dnrm2: the gslcblas reference loops
 *    (cblas/source_gemm_r.h, source_gemv_r.h, source_dot_r.h, source_nrm2_r.h):
 *    loop order and the `if (temp != 0.0)` skip in the NoTrans x NoTrans and
 *    Trans x NoTrans branches are part of the observable rounding behaviour
 *    (SURVEY.md App. A.2).
 *  - LU_decomp: partial pivoting, first-maximum pivot, scaling of the
 *    sub-column by the reciprocal 1/Ajj as GSL >= 2.6 does (linalg/lu.c,
 *    LU_decomp_L2); LU_invert: column-by-column solve of L U x = P e_j
 *    (GSL <= 2.5 `gsl_linalg_LU_invert`).  The 3x3 pose inverse is the only
 *    hot-path use; the parity harness computes it ONCE with this routine and
 *    hands the same bits to the reference path and to the CUDA path
 *    (SURVEY.md 8c mitigation 3), so LU is not on the parity surface.
 *  - SV_decomp_jacobi: one-sided Jacobi orthogonalisation (linalg/svd.c),
 *    used only by Matrix::pcaAnalysis (matcher normals, host side).
 *  - gsl_stats_mean: running mean in long double (statistics/mean_source.c).
 *
 * "parity unpinned" at this boundary: the reference holds no test that pins
 * GSL results.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gsl/gsl_blas.h"
#include "gsl/gsl_linalg.h"
#include "gsl/gsl_statistics_double.h"

static void shim_abort(const char* what)
{
  fprintf(stderr, "oracle gsl shim: %s is not implemented (unused on the hot path)\n", what);
  abort();
}

/* ------------------------------------------------------------------ block/vector */

static gsl_block* block_alloc(size_t n)
{
  gsl_block* b = (gsl_block*)malloc(sizeof(gsl_block));
  b->size = n;
  b->data = (double*)malloc((n ? n : 1) * sizeof(double));
  return b;
}

gsl_vector* gsl_vector_alloc(size_t n)
{
  gsl_vector* v = (gsl_vector*)malloc(sizeof(gsl_vector));
  v->block = block_alloc(n);
  v->data = v->block->data;
  v->size = n;
  v->stride = 1;
  v->owner = 1;
  return v;
}

void gsl_vector_free(gsl_vector* v)
{
  if(!v) return;
  if(v->owner) { free(v->block->data); free(v->block); }
  free(v);
}

int gsl_vector_memcpy(gsl_vector* dst, const gsl_vector* src)
{
  for(size_t i = 0; i < src->size; i++) dst->data[i * dst->stride] = src->data[i * src->stride];
  return GSL_SUCCESS;
}

double* gsl_vector_ptr(gsl_vector* v, size_t i) { return v->data + i * v->stride; }
double gsl_vector_get(const gsl_vector* v, size_t i) { return v->data[i * v->stride]; }
void gsl_vector_set(gsl_vector* v, size_t i, double x) { v->data[i * v->stride] = x; }
void gsl_vector_set_zero(gsl_vector* v) { for(size_t i = 0; i < v->size; i++) v->data[i * v->stride] = 0.0; }

int gsl_vector_add_constant(gsl_vector* v, double x)
{
  for(size_t i = 0; i < v->size; i++) v->data[i * v->stride] += x;
  return GSL_SUCCESS;
}

double gsl_vector_max(const gsl_vector* v)
{
  double max = v->data[0];
  for(size_t i = 0; i < v->size; i++)
  {
    double x = v->data[i * v->stride];
    if(x > max) max = x;
    if(isnan(x)) return x;
  }
  return max;
}

double gsl_vector_min(const gsl_vector* v)
{
  double min = v->data[0];
  for(size_t i = 0; i < v->size; i++)
  {
    double x = v->data[i * v->stride];
    if(x < min) min = x;
    if(isnan(x)) return x;
  }
  return min;
}

_gsl_vector_view gsl_vector_view_array(double* base, size_t n)
{
  _gsl_vector_view view;
  view.vector.size = n; view.vector.stride = 1; view.vector.data = base;
  view.vector.block = NULL; view.vector.owner = 0;
  return view;
}

_gsl_vector_const_view gsl_vector_const_view_array(const double* base, size_t n)
{
  _gsl_vector_const_view view;
  view.vector.size = n; view.vector.stride = 1; view.vector.data = (double*)base;
  view.vector.block = NULL; view.vector.owner = 0;
  return view;
}

/* ------------------------------------------------------------------ matrix */

gsl_matrix* gsl_matrix_alloc(size_t n1, size_t n2)
{
  gsl_matrix* m = (gsl_matrix*)malloc(sizeof(gsl_matrix));
  m->block = block_alloc(n1 * n2);
  m->data = m->block->data;
  m->size1 = n1; m->size2 = n2; m->tda = n2; m->owner = 1;
  return m;
}

void gsl_matrix_free(gsl_matrix* m)
{
  if(!m) return;
  if(m->owner) { free(m->block->data); free(m->block); }
  free(m);
}

int gsl_matrix_memcpy(gsl_matrix* dst, const gsl_matrix* src)
{
  if(dst->size1 != src->size1 || dst->size2 != src->size2)
  {
    fprintf(stderr, "oracle gsl shim: gsl_matrix_memcpy size mismatch (%zux%zu <- %zux%zu)\n",
            dst->size1, dst->size2, src->size1, src->size2);
    abort();
  }
  for(size_t i = 0; i < src->size1; i++)
    for(size_t j = 0; j < src->size2; j++)
      dst->data[i * dst->tda + j] = src->data[i * src->tda + j];
  return GSL_SUCCESS;
}

_gsl_matrix_view gsl_matrix_submatrix(gsl_matrix* m, size_t i, size_t j, size_t n1, size_t n2)
{
  _gsl_matrix_view view;
  view.matrix.data = m->data + (i * m->tda + j);
  view.matrix.size1 = n1; view.matrix.size2 = n2; view.matrix.tda = m->tda;
  view.matrix.block = m->block; view.matrix.owner = 0;
  return view;
}

int gsl_matrix_sub(gsl_matrix* a, const gsl_matrix* b)
{
  for(size_t i = 0; i < a->size1; i++)
    for(size_t j = 0; j < a->size2; j++) a->data[i * a->tda + j] -= b->data[i * b->tda + j];
  return GSL_SUCCESS;
}

int gsl_matrix_add(gsl_matrix* a, const gsl_matrix* b)
{
  for(size_t i = 0; i < a->size1; i++)
    for(size_t j = 0; j < a->size2; j++) a->data[i * a->tda + j] += b->data[i * b->tda + j];
  return GSL_SUCCESS;
}

int gsl_matrix_add_constant(gsl_matrix* a, double x)
{
  for(size_t i = 0; i < a->size1; i++)
    for(size_t j = 0; j < a->size2; j++) a->data[i * a->tda + j] += x;
  return GSL_SUCCESS;
}

_gsl_vector_view gsl_matrix_column(gsl_matrix* m, size_t j)
{
  _gsl_vector_view view;
  view.vector.data = m->data + j; view.vector.size = m->size1; view.vector.stride = m->tda;
  view.vector.block = m->block; view.vector.owner = 0;
  return view;
}

_gsl_vector_view gsl_matrix_row(gsl_matrix* m, size_t i)
{
  _gsl_vector_view view;
  view.vector.data = m->data + i * m->tda; view.vector.size = m->size2; view.vector.stride = 1;
  view.vector.block = m->block; view.vector.owner = 0;
  return view;
}

double* gsl_matrix_ptr(gsl_matrix* m, size_t i, size_t j) { return m->data + (i * m->tda + j); }
double gsl_matrix_get(const gsl_matrix* m, size_t i, size_t j) { return m->data[i * m->tda + j]; }
void gsl_matrix_set(gsl_matrix* m, size_t i, size_t j, double x) { m->data[i * m->tda + j] = x; }

_gsl_matrix_const_view gsl_matrix_const_view_array(const double* base, size_t n1, size_t n2)
{
  _gsl_matrix_const_view view;
  view.matrix.data = (double*)base; view.matrix.size1 = n1; view.matrix.size2 = n2; view.matrix.tda = n2;
  view.matrix.block = NULL; view.matrix.owner = 0;
  return view;
}

_gsl_matrix_view gsl_matrix_view_array(double* base, size_t n1, size_t n2)
{
  _gsl_matrix_view view;
  view.matrix.data = base; view.matrix.size1 = n1; view.matrix.size2 = n2; view.matrix.tda = n2;
  view.matrix.block = NULL; view.matrix.owner = 0;
  return view;
}

void gsl_matrix_set_identity(gsl_matrix* m)
{
  for(size_t i = 0; i < m->size1; i++)
    for(size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] = (i == j) ? 1.0 : 0.0;
}

void gsl_matrix_set_zero(gsl_matrix* m)
{
  for(size_t i = 0; i < m->size1; i++)
    for(size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] = 0.0;
}

int gsl_matrix_transpose(gsl_matrix* m)
{
  for(size_t i = 0; i < m->size1; i++)
    for(size_t j = i + 1; j < m->size2; j++)
    {
      double tmp = m->data[i * m->tda + j];
      m->data[i * m->tda + j] = m->data[j * m->tda + i];
      m->data[j * m->tda + i] = tmp;
    }
  return GSL_SUCCESS;
}

int gsl_matrix_transpose_memcpy(gsl_matrix* dst, const gsl_matrix* src)
{
  for(size_t i = 0; i < dst->size1; i++)
    for(size_t j = 0; j < dst->size2; j++) dst->data[i * dst->tda + j] = src->data[j * src->tda + i];
  return GSL_SUCCESS;
}

gsl_permutation* gsl_permutation_alloc(size_t n)
{
  gsl_permutation* p = (gsl_permutation*)malloc(sizeof(gsl_permutation));
  p->size = n;
  p->data = (size_t*)malloc((n ? n : 1) * sizeof(size_t));
  return p;
}

void gsl_permutation_free(gsl_permutation* p)
{
  if(!p) return;
  free(p->data);
  free(p);
}

/* ------------------------------------------------------------------ blas */

/* cblas_dgemm, row major (gslcblas source_gemm_r.h). */
int gsl_blas_dgemm(CBLAS_TRANSPOSE_t TransA, CBLAS_TRANSPOSE_t TransB, double alpha,
                   const gsl_matrix* A, const gsl_matrix* B, double beta, gsl_matrix* C)
{
  const size_t n1 = C->size1;
  const size_t n2 = C->size2;
  const size_t K = (TransA == CblasNoTrans) ? A->size2 : A->size1;
  const double* F = A->data; const size_t ldf = A->tda;
  const double* G = B->data; const size_t ldg = B->tda;
  double* Cd = C->data; const size_t ldc = C->tda;
  size_t i, j, k;

  {
    const size_t MA = (TransA == CblasNoTrans) ? A->size1 : A->size2;
    const size_t NB = (TransB == CblasNoTrans) ? B->size2 : B->size1;
    const size_t KB = (TransB == CblasNoTrans) ? B->size1 : B->size2;
    if(MA != n1 || NB != n2 || KB != K)
    {
      fprintf(stderr, "oracle gsl shim: dgemm invalid length\n");
      abort();
    }
  }

  if(alpha == 0.0 && beta == 1.0) return GSL_SUCCESS;

  if(beta == 0.0)
  {
    for(i = 0; i < n1; i++) for(j = 0; j < n2; j++) Cd[ldc * i + j] = 0.0;
  }
  else if(beta != 1.0)
  {
    for(i = 0; i < n1; i++) for(j = 0; j < n2; j++) Cd[ldc * i + j] *= beta;
  }

  if(alpha == 0.0) return GSL_SUCCESS;

  if(TransA == CblasNoTrans && TransB == CblasNoTrans)
  {
    for(k = 0; k < K; k++)
      for(i = 0; i < n1; i++)
      {
        const double temp = alpha * F[ldf * i + k];
        if(temp != 0.0)
          for(j = 0; j < n2; j++) Cd[ldc * i + j] += temp * G[ldg * k + j];
      }
  }
  else if(TransA == CblasNoTrans && TransB == CblasTrans)
  {
    for(i = 0; i < n1; i++)
      for(j = 0; j < n2; j++)
      {
        double temp = 0.0;
        for(k = 0; k < K; k++) temp += F[ldf * i + k] * G[ldg * j + k];
        Cd[ldc * i + j] += alpha * temp;
      }
  }
  else if(TransA == CblasTrans && TransB == CblasNoTrans)
  {
    for(k = 0; k < K; k++)
      for(i = 0; i < n1; i++)
      {
        const double temp = alpha * F[ldf * k + i];
        if(temp != 0.0)
          for(j = 0; j < n2; j++) Cd[ldc * i + j] += temp * G[ldg * k + j];
      }
  }
  else
  {
    for(i = 0; i < n1; i++)
      for(j = 0; j < n2; j++)
      {
        double temp = 0.0;
        for(k = 0; k < K; k++) temp += F[ldf * k + i] * G[ldg * j + k];
        Cd[ldc * i + j] += alpha * temp;
      }
  }
  return GSL_SUCCESS;
}

/* cblas_dgemv, row major (gslcblas source_gemv_r.h). */
int gsl_blas_dgemv(CBLAS_TRANSPOSE_t TransA, double alpha, const gsl_matrix* A,
                   const gsl_vector* X, double beta, gsl_vector* Y)
{
  const size_t M = A->size1, N = A->size2;
  const size_t lenY = (TransA == CblasNoTrans) ? M : N;
  size_t i, j;
  if(alpha == 0.0 && beta == 1.0) return GSL_SUCCESS;
  if(beta == 0.0) { for(i = 0; i < lenY; i++) Y->data[i * Y->stride] = 0.0; }
  else if(beta != 1.0) { for(i = 0; i < lenY; i++) Y->data[i * Y->stride] *= beta; }
  if(alpha == 0.0) return GSL_SUCCESS;
  if(TransA == CblasNoTrans)
  {
    for(i = 0; i < M; i++)
    {
      double temp = 0.0;
      for(j = 0; j < N; j++) temp += X->data[j * X->stride] * A->data[A->tda * i + j];
      Y->data[i * Y->stride] += alpha * temp;
    }
  }
  else
  {
    for(j = 0; j < M; j++)
    {
      const double temp = alpha * X->data[j * X->stride];
      if(temp != 0.0)
        for(i = 0; i < N; i++) Y->data[i * Y->stride] += temp * A->data[A->tda * j + i];
    }
  }
  return GSL_SUCCESS;
}

int gsl_blas_ddot(const gsl_vector* X, const gsl_vector* Y, double* result)
{
  double r = 0.0;
  for(size_t i = 0; i < X->size; i++) r += X->data[i * X->stride] * Y->data[i * Y->stride];
  *result = r;
  return GSL_SUCCESS;
}

/* gslcblas source_nrm2_r.h: scaled sum of squares. */
double gsl_blas_dnrm2(const gsl_vector* X)
{
  double scale = 0.0, ssq = 1.0;
  const size_t N = X->size;
  if(N == 0) return 0.0;
  if(N == 1) return fabs(X->data[0]);
  for(size_t i = 0; i < N; i++)
  {
    const double x = X->data[i * X->stride];
    if(x != 0.0)
    {
      const double ax = fabs(x);
      if(scale < ax) { ssq = 1.0 + ssq * (scale / ax) * (scale / ax); scale = ax; }
      else { ssq += (ax / scale) * (ax / scale); }
    }
  }
  return scale * sqrt(ssq);
}

/* ------------------------------------------------------------------ linalg */

int gsl_linalg_LU_decomp(gsl_matrix* A, gsl_permutation* p, int* signum)
{
  const size_t N = A->size1;
  size_t i, j, k;
  *signum = 1;
  for(i = 0; i < N; i++) p->data[i] = i;

  for(j = 0; j < N; j++)
  {
    /* pivot: first maximum of |A(i,j)|, i >= j (idamax) */
    double max = fabs(A->data[j * A->tda + j]);
    size_t i_pivot = j;
    for(i = j + 1; i < N; i++)
    {
      double aij = fabs(A->data[i * A->tda + j]);
      if(aij > max) { max = aij; i_pivot = i; }
    }
    if(i_pivot != j)
    {
      for(k = 0; k < N; k++)
      {
        double tmp = A->data[j * A->tda + k];
        A->data[j * A->tda + k] = A->data[i_pivot * A->tda + k];
        A->data[i_pivot * A->tda + k] = tmp;
      }
      size_t t = p->data[j]; p->data[j] = p->data[i_pivot]; p->data[i_pivot] = t;
      *signum = -(*signum);
    }
    {
      const double ajj = A->data[j * A->tda + j];
      if(fabs(ajj) >= GSL_DBL_MIN)
      {
        const double inv = 1.0 / ajj;
        for(i = j + 1; i < N; i++) A->data[i * A->tda + j] *= inv;
      }
      else
      {
        for(i = j + 1; i < N; i++) A->data[i * A->tda + j] /= ajj;
      }
      /* rank-1 update of the trailing block (dger with alpha = -1) */
      for(i = j + 1; i < N; i++)
      {
        const double tmp = -1.0 * A->data[i * A->tda + j];
        for(k = j + 1; k < N; k++) A->data[i * A->tda + k] += A->data[j * A->tda + k] * tmp;
      }
    }
  }
  return GSL_SUCCESS;
}

static void lu_svx(const gsl_matrix* LU, const gsl_permutation* p, double* x /* stride 1, size N */)
{
  const size_t N = LU->size1;
  size_t i, j;
  /* apply permutation: x <- P x */
  double* tmp = (double*)malloc(N * sizeof(double));
  for(i = 0; i < N; i++) tmp[i] = x[p->data[i]];
  for(i = 0; i < N; i++) x[i] = tmp[i];
  free(tmp);
  /* solve L y = Pb (unit lower), dtrsv NoTrans Lower Unit: forward substitution */
  for(i = 1; i < N; i++)
  {
    double t = x[i];
    for(j = 0; j < i; j++) t -= LU->data[i * LU->tda + j] * x[j];
    x[i] = t;
  }
  /* solve U x = y, dtrsv NoTrans Upper NonUnit: backsubstitution */
  if(N > 0)
  {
    x[N - 1] = x[N - 1] / LU->data[(N - 1) * LU->tda + (N - 1)];
    for(i = N - 1; i > 0 && i--;)
    {
      double t = x[i];
      for(j = i + 1; j < N; j++) t -= LU->data[i * LU->tda + j] * x[j];
      x[i] = t / LU->data[i * LU->tda + i];
    }
  }
}

int gsl_linalg_LU_invert(const gsl_matrix* LU, const gsl_permutation* p, gsl_matrix* inverse)
{
  const size_t N = LU->size1;
  double* col = (double*)malloc(N * sizeof(double));
  for(size_t j = 0; j < N; j++)
  {
    for(size_t i = 0; i < N; i++) col[i] = (i == j) ? 1.0 : 0.0;
    lu_svx(LU, p, col);
    for(size_t i = 0; i < N; i++) inverse->data[i * inverse->tda + j] = col[i];
  }
  free(col);
  return GSL_SUCCESS;
}

int gsl_linalg_LU_solve(const gsl_matrix* LU, const gsl_permutation* p, const gsl_vector* b, gsl_vector* x)
{
  const size_t N = LU->size1;
  double* col = (double*)malloc(N * sizeof(double));
  for(size_t i = 0; i < N; i++) col[i] = b->data[i * b->stride];
  lu_svx(LU, p, col);
  for(size_t i = 0; i < N; i++) x->data[i * x->stride] = col[i];
  free(col);
  return GSL_SUCCESS;
}

int gsl_linalg_SV_decomp(gsl_matrix* A, gsl_matrix* V, gsl_vector* S, gsl_vector* work)
{
  (void)A; (void)V; (void)S; (void)work;
  shim_abort("gsl_linalg_SV_decomp");
  return -1;
}

int gsl_linalg_QR_decomp(gsl_matrix* A, gsl_vector* tau)
{
  (void)A; (void)tau;
  shim_abort("gsl_linalg_QR_decomp");
  return -1;
}

int gsl_linalg_QR_lssolve(const gsl_matrix* QR, const gsl_vector* tau, const gsl_vector* b, gsl_vector* x, gsl_vector* residual)
{
  (void)QR; (void)tau; (void)b; (void)x; (void)residual;
  shim_abort("gsl_linalg_QR_lssolve");
  return -1;
}

/* One-sided Jacobi SVD (GSL linalg/svd.c, gsl_linalg_SV_decomp_jacobi). */
int gsl_linalg_SV_decomp_jacobi(gsl_matrix* A, gsl_matrix* Q, gsl_vector* S)
{
  const size_t M = A->size1;
  const size_t N = A->size2;
  size_t i, j, k;
  int count = 1;
  int sweep = 0;
  int sweepmax = 5 * (int)N;
  double tolerance = 10 * M * GSL_DBL_EPSILON;

  if(sweepmax < 12) sweepmax = 12;

  gsl_matrix_set_identity(Q);

  for(j = 0; j < N; j++)
  {
    gsl_vector_view cj = gsl_matrix_column(A, j);
    double sj = gsl_blas_dnrm2(&cj.vector);
    gsl_vector_set(S, j, GSL_DBL_EPSILON * sj);
  }

  while(count > 0 && sweep <= sweepmax)
  {
    count = (int)(N * (N - 1) / 2);
    for(j = 0; j + 1 < N; j++)
    {
      for(k = j + 1; k < N; k++)
      {
        double a = 0.0, b = 0.0, p = 0.0, q = 0.0;
        double cosine, sine, v, abserr_a, abserr_b;
        int sorted, orthog, noisya, noisyb;

        gsl_vector_view cj = gsl_matrix_column(A, j);
        gsl_vector_view ck = gsl_matrix_column(A, k);

        gsl_blas_ddot(&cj.vector, &ck.vector, &p);
        p *= 2.0;

        a = gsl_blas_dnrm2(&cj.vector);
        b = gsl_blas_dnrm2(&ck.vector);

        q = a * a - b * b;
        v = hypot(p, q);

        abserr_a = gsl_vector_get(S, j);
        abserr_b = gsl_vector_get(S, k);

        sorted = (a >= b);
        orthog = (fabs(p) <= tolerance * (a * b));
        noisya = (a < abserr_a);
        noisyb = (b < abserr_b);

        if(sorted && (orthog || noisya || noisyb))
        {
          count--;
          continue;
        }

        if(v == 0 || !sorted)
        {
          cosine = 0.0;
          sine = 1.0;
        }
        else
        {
          cosine = sqrt((v + q) / (2.0 * v));
          sine = p / (2.0 * v * cosine);
        }

        for(i = 0; i < M; i++)
        {
          const double Aik = gsl_matrix_get(A, i, k);
          const double Aij = gsl_matrix_get(A, i, j);
          gsl_matrix_set(A, i, j, Aij * cosine + Aik * sine);
          gsl_matrix_set(A, i, k, -Aij * sine + Aik * cosine);
        }

        gsl_vector_set(S, j, fabs(cosine) * abserr_a + fabs(sine) * abserr_b);
        gsl_vector_set(S, k, fabs(sine) * abserr_a + fabs(cosine) * abserr_b);

        for(i = 0; i < N; i++)
        {
          const double Qij = gsl_matrix_get(Q, i, j);
          const double Qik = gsl_matrix_get(Q, i, k);
          gsl_matrix_set(Q, i, j, Qij * cosine + Qik * sine);
          gsl_matrix_set(Q, i, k, -Qij * sine + Qik * cosine);
        }
      }
    }
    sweep++;
  }

  {
    double prev_norm = -1.0;
    for(j = 0; j < N; j++)
    {
      gsl_vector_view column = gsl_matrix_column(A, j);
      double norm = gsl_blas_dnrm2(&column.vector);
      if(norm == 0.0 || prev_norm == 0.0 || (j > 0 && norm <= tolerance * prev_norm))
      {
        gsl_vector_set(S, j, 0.0);
        for(i = 0; i < M; i++) gsl_matrix_set(A, i, j, 0.0);
        prev_norm = 0.0;
      }
      else
      {
        gsl_vector_set(S, j, norm);
        for(i = 0; i < M; i++) gsl_matrix_set(A, i, j, gsl_matrix_get(A, i, j) * (1.0 / norm));
        prev_norm = norm;
      }
    }
  }
  return GSL_SUCCESS;
}

/* ------------------------------------------------------------------ statistics */

double gsl_stats_mean(const double data[], size_t stride, size_t n)
{
  long double mean = 0;
  for(size_t i = 0; i < n; i++) mean += (data[i * stride] - mean) / (i + 1);
  return (double)mean;
}
