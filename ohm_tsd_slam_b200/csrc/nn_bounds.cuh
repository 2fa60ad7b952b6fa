// Single-precision lower bound of the squared distance from a query to anything inside an axis-aligned box, as the
// RNM scorer's pruning uses it (match.cu, k_score_rnm).  Host/device code so that tests/cpp/nnbound_check.cpp can run
// exactly these functions on the CPU against the double-precision distances.
//
// The box is kept in single precision, rounded OUTWARD (tsd_box_make); the query is rounded to nearest; `e` bounds, per
// axis, |single-precision difference - true difference|: two conversions (half an ulp each of a magnitude below
// |x| + |m|) and one subtraction (half an ulp of the result, covered by the factor 1 - 2e-7); the product and the sum
// are covered by 1 - 1e-6.  Whatever is inside the box in double precision is therefore at a squared distance >= the
// value returned; a group is skipped only if that value exceeds the best squared distance found so far, rounded UP to
// single precision.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define TSD_NB_HD __host__ __device__ __forceinline__
#else
#define TSD_NB_HD static inline
struct float4 { float x, y, z, w; };
#endif

// per-axis error bound for a query (xf, yf) against a model whose largest |coordinate| is mabs
TSD_NB_HD float tsd_nb_err(float xf, float yf, float mabs) { return 1.3e-7f * (fmaxf(fabsf(xf), fabsf(yf)) + mabs) + 1e-30f; }

// b = {x0, x1, y0, y1}, rounded outward
TSD_NB_HD float tsd_nb_box_lb(float4 b, float xf, float yf, float e)
{
  const float ex = fmaxf(fmaxf(b.x - xf, xf - b.y) * (1.f - 2e-7f) - e, 0.f);
  const float ey = fmaxf(fmaxf(b.z - yf, yf - b.w) * (1.f - 2e-7f) - e, 0.f);
  return (ex * ex + ey * ey) * (1.f - 1e-6f);
}
