// The closed form of the ray caster's position sums (ohm_tsd_slam_b200/csrc/closed_form.cuh, the very function k_raycast
// calls) against the serial additions `position += ray` of the reference (RayCastPolar2D.cpp:243-270), on the CPU:
// whenever the function says "closed", all 32 partial sums must equal p + k d BIT FOR BIT.  Sweeps: positions over many
// binades and next to binade boundaries, steps of every magnitude a ray direction of length cellSize can have, both
// signs, steps that are exact multiples / exact ties of the position's ulp, tiny and zero steps.
//   closedform_check      prints a summary; exit code 0 = no mismatch
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>

#include "../../ohm_tsd_slam_b200/csrc/closed_form.cuh"

static uint64_t bits(double v) { uint64_t b; memcpy(&b, &v, 8); return b; }

int main()
{
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  uint64_t cases = 0, closed = 0, bad = 0, ties = 0;
  auto check = [&](double p, double r)
  {
    double d, e;
    cases++;
    if(!tsd_closed_form_pass(p, r, &d, &e)) return;
    closed++;
    volatile double s = p;  // (volatile: no excess precision, no reassociation)
    for(int k = 1; k <= 32; k++)
    {
      s = s + r;
      const double c = p + (double)k * d;
      if(bits(c) != bits((double)s)) { if(bad < 5) fprintf(stderr, "mismatch p=%a r=%a k=%d serial=%a closed=%a\n", p, r, k, (double)s, c); bad++; return; }
    }
    if(bits(e) != bits((double)s)) bad++;
  };
  const double cell = 0.025;
  // random positions in the map (0 .. 1700 m), random directions of length cellSize
  for(int i = 0; i < 4000000; i++)
  {
    const double p = U(rng) * std::ldexp(1.0, (int)(U(rng) * 12) - 1);
    const double a = U(rng) * 2.0 * M_PI;
    check(p, cell * std::cos(a));
    check(-p, cell * std::sin(a));
  }
  // next to binade boundaries, from both sides, steps towards and away from them
  for(int e = -4; e <= 11; e++)
    for(int i = 0; i < 40000; i++)
    {
      const double b = std::ldexp(1.0, e);
      const double p = b + (U(rng) - 0.5) * 64.0 * cell * U(rng);
      const double r = cell * (U(rng) * 2.0 - 1.0);
      if(p > 0) { check(p, r); check(p, -r); }
    }
  // steps that are exact multiples of the position's ulp, exact ties, and one ulp of the step next to a tie
  for(int i = 0; i < 400000; i++)
  {
    const double p = 1.0 + U(rng) * 500.0;
    int ex;
    std::frexp(p, &ex);
    const double u = std::ldexp(1.0, ex - 53);
    const double q = std::floor(U(rng) * 1e9) * u;
    check(p, q);
    const double t = q + 0.5 * u;
    double dd, ee;
    if(!tsd_closed_form_pass(p, t, &dd, &ee)) ties++;
    check(p, t);
    check(p, std::nextafter(t, 0.0));
    check(p, std::nextafter(t, 1.0));
    check(p, -t);
  }
  // tiny and zero steps, tiny and zero positions
  for(int i = 0; i < 200000; i++)
  {
    const double p = U(rng) * 400.0;
    check(p, 0.0);
    check(p, std::ldexp(U(rng), -60));
    check(std::ldexp(U(rng), -40 - (int)(U(rng) * 1000)), cell * U(rng));
    check(0.0, cell * U(rng));
  }
  printf("%llu cases, %llu closed (%.1f %%), %llu exact ties refused, %llu mismatches\n", (unsigned long long)cases,
         (unsigned long long)closed, 100.0 * (double)closed / (double)cases, (unsigned long long)ties, (unsigned long long)bad);
  return bad ? 1 : 0;
}
