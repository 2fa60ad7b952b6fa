// The serial sum `p += r`, 32 times, in closed form -- where that is exact.  Host/device code so that
// tests/cpp/closedform_check.cpp can run exactly this function on the CPU against the serial additions.
//
// While p stays inside one binade [2^e, 2^(e+1)) (same sign) its values are multiples of u = ulp(p), and fl(p + r) = p + d
// with d = RN_u(r), the multiple of u nearest to r, at every step -- unless r lies exactly half-way between two multiples
// (round-to-even would then look at p's parity).  d is read off one real addition (d = fl(p + r) - p: both are multiples
// of u of the same binade, the difference is exact); p + k d is exact for k <= 32 (|d| < 2^(e-5) if 32 steps stay inside
// the binade, so k d has fewer than 53 bits, and the sum is a multiple of u inside the binade).  The sequence is monotonic,
// so it stays inside the binade iff its last element does.
//   returns true:  after k steps (1 <= k <= 32) the serial sum is exactly p + k * (*d); *end = p + 32 * (*d)
//   returns false: a binade boundary or zero is crossed, p is zero / subnormal-ish, or r is a rounding tie: replay serially
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define TSD_CF_HD __host__ __device__ __forceinline__
#else
#define TSD_CF_HD static inline
#endif

TSD_CF_HD uint32_t tsd_cf_hi(double v)
{
#ifdef __CUDA_ARCH__
  return (uint32_t)__double2hiint(v);
#else
  uint64_t b;
  memcpy(&b, &v, 8);
  return (uint32_t)(b >> 32);
#endif
}
TSD_CF_HD double tsd_cf_from_hi(uint32_t hi)
{
#ifdef __CUDA_ARCH__
  return __hiloint2double((int)hi, 0);
#else
  const uint64_t b = (uint64_t)hi << 32;
  double v;
  memcpy(&v, &b, 8);
  return v;
#endif
}

TSD_CF_HD bool tsd_closed_form_pass(double p, double r, double* d, double* end)
{
  const double dd = (p + r) - p;
  const double e = p + 32.0 * dd;
  const uint32_t h = tsd_cf_hi(p) >> 20;  // sign and exponent
  const uint32_t x = h & 0x7ffu;
  // half an ulp of the binade (exponent - 53); 0 if that would be subnormal: then the strict test below fails
  const double hu = tsd_cf_from_hi(x > 53u ? ((x - 53u) << 20) : 0u);
  *d = dd;
  *end = e;
  return (h == (tsd_cf_hi(e) >> 20)) && (fabs(r - dd) < hu);
}
