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
#include <string.h>

#include <cmath>

#if defined(__CUDACC__)
#define TSD_HD __host__ __device__ __forceinline__
#else
#define TSD_HD static inline
struct float2 { float x, y; };   // (plain C++ build of the host-side checks: tests/cpp/fastpath_check.cpp)
struct double2 { double x, y; };
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


// ------------------------------------------------------------------------------------------------------------------
// Single-precision front end of the cell update (k_update, grid.cu).  Host/device code so that tests/cpp/
// fastpath_check.cpp can run exactly these functions on the CPU against the reference's double-precision expressions.
//
// What a cell of an active partition needs decided (TsdGrid.cpp:250-274 + addTsd, TsdGridPartition.h:170-212):
//   (a) its beam  round((atan2(y', x') - phiMin) / res)  -- an integer;
//   (b) whether sd = r - dist is >= -maxTruncation (the cell is rewritten at all) and whether sd / maxTruncation
//       clips to 1.0 (free space in front of the surface) -- two comparisons.
// Both are decided in single precision whenever the answer is certain, i.e. separated from the alternative by more
// than the error bound of the single-precision evaluation; the (few) other cells take the double-precision
// expressions of the reference.  Certain answers are by construction the reference's answers.
//   beam:  candidate angle from an odd degree-15 minimax polynomial (|error| <= 1.5e-7 rad over every float in [0,1],
//          checked exhaustively) with octant reduction; s = fma(phi, 1/res, -phiMin/res); certain iff s is further
//          than `margin` (~1.2e-3 beams for a 1081-beam 270-degree scanner) from a half-integer.
//   gates: squared cell distance against a per-beam pair of thresholds (tsd_gate_entry).
// ------------------------------------------------------------------------------------------------------------------
struct FastModel
{
  double txp, typ;   // the point the pose inverse maps to the origin (the sensor position as Pi implies it)
  float rinv_f;      // 1 / angularRes
  float off_f;       // -phiMin / angularRes
  float half_m;      // 0.5 - margin; < 0: the front end is off, every cell takes the exact route
  int n;
};

#if defined(__CUDA_ARCH__)
#define TSD_RCPF(x) tsd_rcp_approx(x)
__device__ __forceinline__ float tsd_rcp_approx(float x)
{
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#define TSD_F2I_BITS(x) __float_as_int(x)
#define TSD_D2F_RU(x) __double2float_ru(x)
#define TSD_D2F_RD(x) __double2float_rd(x)
#define TSD_LDG(p) __ldg(p)
#define TSD_FADD(a, b) __fadd_rn((a), (b))
#else
#ifndef TSD_RCPF
#define TSD_RCPF(x) (1.0f / (x))
#endif
static inline int tsd_f2i_bits(float x) { int i; memcpy(&i, &x, 4); return i; }
#define TSD_F2I_BITS(x) tsd_f2i_bits(x)
static inline float tsd_fadd(float a, float b) { volatile float r = a + b; return r; }  // (never folded or widened)
#define TSD_FADD(a, b) tsd_fadd((a), (b))
static inline float tsd_d2f_ru(double x) { float f = (float)x; return ((double)f < x) ? nextafterf(f, INFINITY) : f; }
static inline float tsd_d2f_rd(double x) { float f = (float)x; return ((double)f > x) ? nextafterf(f, -INFINITY) : f; }
#define TSD_D2F_RU(x) tsd_d2f_ru(x)
#define TSD_D2F_RD(x) tsd_d2f_rd(x)
#define TSD_LDG(p) (*(p))
#endif

// Host: the front end's parameters for one scan (Pi = pose inverse, row-major 3x3).  The candidate angle comes from
// sensor-frame coordinates formed as Pi00*(X - tx') + Pi01*(Y - ty') in single precision; that is accurate relative
// to the cell's distance only when the 2x2 part of Pi is a (scaled) rotation.  The cut of atan2 at +-pi must lie
// outside the field of view by a margin: next to it the SIGN of a single-precision y' is not reliable.
static inline void tsd_fast_model(const double Pi[9], double phi_min, double angular_res, double phi_lower, double phi_upper,
                                  int n, FastModel* o)
{
  o->txp = o->typ = 0.0;
  o->rinv_f = o->off_f = 0.0f;
  o->half_m = -1.0f;
  o->n = n;
  const double pi = 3.14159265358979323846;
  if(!(angular_res > 1e-9 && phi_lower >= -pi + 0.01 && phi_upper <= pi - 0.01 && phi_upper > phi_lower && n >= 1)) return;
  const double res_inv = 1.0 / angular_res;
  const double a = Pi[0], b = Pi[1], c = Pi[3], d = Pi[4];
  const double det = a * d - b * c, n0 = a * a + b * b, n1 = c * c + d * d;
  const bool rot = std::isfinite(det) && n0 > 0.0 && std::fabs(det) >= 0.999 * n0 && std::fabs(n0 - n1) <= 1e-3 * n0 &&
                   std::fabs(a * c + b * d) <= 1e-3 * n0;
  // error budget of the candidate angle (rad): polynomial 1.5e-7 + reciprocal and octant fix-ups 5.6e-7 + inputs
  // 3.6e-7 < 1.1e-6; budgeted 2.4e-6.  The scaled angle adds the roundings of fma(phi, rinv_f, off_f):
  // < 2.5e-7 * (|off| + n + 2) beams.  Half again on top of the sum.
  const double off = -phi_min * res_inv;
  const double margin = 1.5 * (2.4e-6 * res_inv + 2.5e-7 * (std::fabs(off) + (double)n + 2.0));
  const bool small = std::fabs(res_inv) * 3.2 + std::fabs(off) < 2.0e6;  // the magic-number rounding needs |s| < 2^22
  if(!(rot && small && margin < 0.2 && std::isfinite(Pi[2]) && std::isfinite(Pi[5]))) return;
  o->txp = (-Pi[2] * d + b * Pi[5]) / det;
  o->typ = (-a * Pi[5] + Pi[2] * c) / det;
  if(!(std::isfinite(o->txp) && std::isfinite(o->typ))) return;
  o->rinv_f = (float)res_inv;
  o->off_f = (float)off;
  o->half_m = (float)(0.5 - margin);
}

// Per beam, the two squared cell distances that decide a cell's fate without the double-precision square root:
// d^2 > hi2: the signed distance is below -maxTruncation (or the beam is masked / the cell is beyond the
// low-reflectivity range of a beam without return): the cell is not rewritten; d^2 < lo2: the signed distance is above
// +maxTruncation, the new tsd is exactly 1.0.  In between (and for the cells near a sensor whose beam has no return)
// the exact route decides.  The slack (4e-6 relative + 1e-5 m, outward rounding) dwarfs the error of the
// single-precision d^2 (< 2e-7 relative).
TSD_HD void tsd_gate_entry(double r, bool m, double max_trunc, double low_refl, float* lo2o, float* hi2o)
{
  float lo2 = -1.0f, hi2 = -INFINITY;
  if(m && !isnan(r))
  {
    const double hi = isinf(r) ? low_refl : r + max_trunc;
    const double h = hi + 4e-6 * fabs(hi) + 1e-5;
    if(h >= 0.0) hi2 = TSD_D2F_RU(h * h * (1.0 + 1e-12));
    if(!isinf(r))
    {
      const double lo = r - max_trunc;
      const double l = lo - 4e-6 * fabs(lo) - 1e-5;
      if(l > 0.0) lo2 = TSD_D2F_RD(l * l * (1.0 - 1e-12));
    }
    if(isnan(hi2)) { hi2 = INFINITY; lo2 = -1.0f; }  // NaN sensor parameters: everything takes the exact route
  }
  *lo2o = lo2;
  *hi2o = hi2;
}

// classes: 0 not rewritten, 1 rewritten with tsd_new = 1.0, 2 beam known (kOut) / distance needs the exact route,
//          3 beam needs the exact route.
TSD_HD int tsd_classify_cell(float rinv_f, float off_f, float half_m, int n, const float2* __restrict__ gate, float xf, float yf,
                             float d2f, int& kOut)
{
  const float ax = fabsf(xf), ay = fabsf(yf);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float rc = TSD_RCPF(mx);
  const float t = mn * rc;
  const float q = t * t;
  float p = -0.0040545654483139515f;
  p = TSD_FMAF(p, q, 0.021862952038645744f);
  p = TSD_FMAF(p, q, -0.0559123195707798f);
  p = TSD_FMAF(p, q, 0.0964219719171524f);
  p = TSD_FMAF(p, q, -0.1390862911939621f);
  p = TSD_FMAF(p, q, 0.19946566224098206f);
  p = TSD_FMAF(p, q, -0.33329859375953674f);
  p = TSD_FMAF(p, q, 0.9999993443489075f);
  float r = p * t;
  r = (ay > ax) ? 1.57079632679489662f - r : r;
  r = (xf < 0.0f) ? 3.14159265358979324f - r : r;
  const float phi = copysignf(r, yf);
  const float sc = TSD_FMAF(phi, rinv_f, off_f);
  // round to nearest without a conversion: adding 1.5 * 2^23 leaves the integer in the low mantissa bits
  const float u = TSD_FADD(sc, 12582912.0f);
  const float kf = TSD_FADD(u, -12582912.0f);
  const int k = TSD_F2I_BITS(u) - 0x4b400000;
  // certain: away from the half-integers, and not a cell (almost) on top of the sensor, whose angle means nothing
  const bool certain = (fabsf(sc - kf) < half_m) && (mx >= 1e-3f);
  const bool inside = (unsigned)k < (unsigned)n;
  const float2 g = TSD_LDG(gate + (inside ? k : 0));
  kOut = k;
  int cls = (d2f > g.y) ? 0 : ((d2f < g.x) ? 1 : 2);
  cls = inside ? cls : 0;  // certain and outside the field of view: backProject gives -1 / -2, no update
  return certain ? cls : 3;
}
