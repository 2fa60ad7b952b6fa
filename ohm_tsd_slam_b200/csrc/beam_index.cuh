// Beam index of a point in the sensor frame: the device-side replacement of
// SensorPolar2D::backProject (reference src/obvision/reconstruct/grid/SensorPolar2D.cpp:117-135).
//
// Reference semantics (exact):   phi = atan2(y, x)
//                                phi <= phiLower -> -2 ;  phi >= phiUpper -> -1
//                                else (int)round((phi - phiMin) * (1.0 / angularRes))
//
// A double-precision atan2 costs ~150 FP64-pipe instructions, more than the whole rest of the TSD cell
// update, and the FP64 pipe (64 lanes/clk/SM on B200) is what co-limits TsdGrid::push with HBM.  So:
//   1. a candidate beam k comes from a single-precision polynomial atan2 (|error| < 2e-6 rad, ~20 FP32
//      instructions: one reciprocal, six fused multiply-adds, three quadrant fix-ups);
//   2. it is CONFIRMED in double precision with two cross products against the directions of the two
//      half-beam boundaries  B_k = phiMin + (k - 1/2) res  and  B_{k+1}  (table `dirs`, in shared memory):
//         cross(dir(B_k), p)   >  +margin * (|x|+|y|)       (p strictly left of the lower boundary)
//         cross(dir(B_k+1), p) <  -margin * (|x|+|y|)       (p strictly right of the upper boundary)
//      margin = 1e-11 dwarfs every rounding error involved (table entries, products, the reference's own
//      atan2: all < 1e-15), so a confirmed candidate IS the reference's answer;
//   3. anything not confirmed (wrong candidate: ~3e-4 of cells; inside the margin band or the FOV edges:
//      ~1e-8 of cells) takes the slow path, which evaluates the reference formula verbatim with the
//      double-precision atan2.
// The slow path can differ from glibc only where the two libms' atan2 differ in the last ulp AND that ulp
// flips the rounding, i.e. with probability ~1e-13 per slow-path cell.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define TSD_HD __host__ __device__ __forceinline__
#else
#define TSD_HD static inline
#endif

struct BeamModel
{
  double phi_min;
  double res_inv;    // 1.0 / angularRes, as SensorPolar2D.cpp:127
  double phi_lower;
  double phi_upper;
  float phi_min_f;
  float res_inv_f;
  float phi_lower_f;
  float phi_upper_f;
  int n;             // number of beams
  int fast_ok;       // 0: field of view leaves [-pi, pi] or the model is degenerate -> always the slow path
};

#define TSD_BEAM_MARGIN 1e-11

#if defined(__CUDA_ARCH__)
#define TSD_FMAF(a, b, c) __fmaf_rn((a), (b), (c))
#define TSD_FDIV(a, b) __fdividef((a), (b))
#else
#define TSD_FMAF(a, b, c) fmaf((a), (b), (c))
#define TSD_FDIV(a, b) ((a) / (b))
#endif

// atan2 in single precision, |error| < 2e-6 rad: odd minimax polynomial of degree 11 on [0,1] plus octant
// reduction.  Only ever proposes a candidate beam; every candidate is confirmed in double precision below.
TSD_HD float fast_atan2f(float y, float x)
{
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float t = TSD_FDIV(mn, mx);
  const float s = t * t;
  float p = -0.011719098314642906f;
  p = TSD_FMAF(p, s, 0.05264726281166077f);
  p = TSD_FMAF(p, s, -0.11642640829086304f);
  p = TSD_FMAF(p, s, 0.19354034960269928f);
  p = TSD_FMAF(p, s, -0.33262282609939575f);
  p = TSD_FMAF(p, s, 0.9999772310256958f);
  float r = p * t;
  if(ay > ax) r = 1.57079632679489662f - r;
  if(x < 0.0f) r = 3.14159265358979324f - r;
  return copysignf(r, y);
}

// reference formula, verbatim (slow path)
TSD_HD int beam_index_exact(const BeamModel& bm, double x, double y)
{
  const double phi = atan2(y, x);
  if(phi <= bm.phi_lower) return -2;
  if(phi >= bm.phi_upper) return -1;
  return (int)round((phi - bm.phi_min) * bm.res_inv);
}

// dirs: (n + 1) entries {cos B_k, sin B_k}, k = 0..n.  *slow is set when the slow path was taken.
TSD_HD int beam_index(const BeamModel& bm, const double2* __restrict__ dirs, double x, double y, bool* slow)
{
  if(bm.fast_ok)
  {
    const float phif = fast_atan2f((float)y, (float)x);
    const float vf = (phif - bm.phi_min_f) * bm.res_inv_f;
    const int k = (int)rintf(vf);
    if(k >= 0 && k < bm.n)
    {
#if defined(__CUDA_ARCH__)
      const double2 lo = __ldg(dirs + k);
      const double2 hi = __ldg(dirs + k + 1);
#else
      const double2 lo = dirs[k];
      const double2 hi = dirs[k + 1];
#endif
      const double m = TSD_BEAM_MARGIN * (fabs(x) + fabs(y));
      const double s_lo = lo.x * y - lo.y * x;
      const double s_hi = hi.x * y - hi.y * x;
      if(s_lo > m && s_hi < -m)
      {
        *slow = false;
        return k;
      }
    }
    else
    {
      // Clearly outside the field of view: the candidate angle is good to ~2e-6 rad, 1e-3 rad of slack decides the
      // reference's two comparisons (lower bound first) without the double-precision atan2.
      if(phif < bm.phi_lower_f - 1e-3f) { *slow = false; return -2; }
      if(phif > bm.phi_upper_f + 1e-3f) { *slow = false; return -1; }
    }
  }
  *slow = true;
  return beam_index_exact(bm, x, y);
}
