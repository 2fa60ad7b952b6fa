// The single-precision box bound of the RNM scorer's pruning (ohm_tsd_slam_b200/csrc/nn_bounds.cuh, the very functions
// k_score_rnm calls) against double-precision distances on the CPU: for random boxes (built from random points, rounded
// outward as the kernel does), random and adversarial queries (far away, next to the box, inside it, with coordinates up
// to the largest map), the bound must never exceed the squared distance (0 + dx*dx) + dy*dy to ANY of the box's points
// rounded UP to single precision -- the quantity the kernel compares it with.
//   nnbound_check      prints a summary; exit code 0 = no violation
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

#include "../../ohm_tsd_slam_b200/csrc/nn_bounds.cuh"

static float f_rd(double v) { float f = (float)v; if((double)f > v) f = std::nextafterf(f, -INFINITY); return f; }
static float f_ru(double v) { float f = (float)v; if((double)f < v) f = std::nextafterf(f, INFINITY); return f; }

int main()
{
  std::mt19937_64 rng(777);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  uint64_t cases = 0, bad = 0, pruned = 0;
  for(int it = 0; it < 300000; it++)
  {
    // a group of up to 16 points along a short stretch of contour somewhere in a map of up to 1.6 km
    const double scale = std::ldexp(1.0, (int)(U(rng) * 11));
    const double cx = (U(rng) - 0.5) * scale, cy = (U(rng) - 0.5) * scale, ext = 0.01 + U(rng) * 2.0;
    const int n = 1 + (int)(U(rng) * 16);
    std::vector<double> px(n), py(n);
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300, mabs_d = 0.0;
    for(int k = 0; k < n; k++)
    {
      px[k] = cx + (U(rng) - 0.5) * ext;
      py[k] = cy + (U(rng) - 0.5) * ext * U(rng);
      x0 = std::fmin(x0, px[k]); x1 = std::fmax(x1, px[k]);
      y0 = std::fmin(y0, py[k]); y1 = std::fmax(y1, py[k]);
    }
    float4 b;
    b.x = f_rd(x0); b.y = f_ru(x1); b.z = f_rd(y0); b.w = f_ru(y1);
    // the kernel's mabs is the largest |box coordinate| of the whole model: at least this box's
    const float mabs = std::fmax(std::fmax(std::fabs(b.x), std::fabs(b.y)), std::fmax(std::fabs(b.z), std::fabs(b.w)));
    (void)mabs_d;
    for(int q = 0; q < 8; q++)
    {
      double x, y;
      const int kind = q % 4;
      if(kind == 0) { x = cx + (U(rng) - 0.5) * scale; y = cy + (U(rng) - 0.5) * scale; }                  // anywhere
      else if(kind == 1) { x = x1 + U(rng) * 1e-3 * (q & 4 ? 1e-3 : 1.0); y = y0 + U(rng) * (y1 - y0); }  // just outside an edge
      else if(kind == 2) { x = x0 + U(rng) * (x1 - x0); y = y0 + U(rng) * (y1 - y0); }                     // inside
      else { x = x0 - U(rng) * ext; y = y1 + U(rng) * ext; }                                               // off a corner
      const float xf = (float)x, yf = (float)y;
      const float lb = tsd_nb_box_lb(b, xf, yf, tsd_nb_err(xf, yf, mabs));
      cases++;
      for(int k = 0; k < n; k++)
      {
        const double d0 = x - px[k], d1 = y - py[k];
        double d = 0.0;
        d += d0 * d0;
        d += d1 * d1;
        if(lb > f_ru(d)) { if(bad < 5) fprintf(stderr, "violation: lb %.9g > d %.17g (box %g %g %g %g, query %.17g %.17g)\n", lb, d, b.x, b.y, b.z, b.w, x, y); bad++; }
      }
      // how sharp the bound is: the share of queries outside the box by more than 1 mm for which it is positive
      if(lb > 0.f) pruned++;
    }
  }
  printf("%llu queries, bound positive for %.1f %%, %llu violations\n", (unsigned long long)cases, 100.0 * (double)pruned / (double)cases,
         (unsigned long long)bad);
  return bad ? 1 : 0;
}
