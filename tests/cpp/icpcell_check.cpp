// The cell arithmetic of k_icp's nearest-neighbour search (ohm_tsd_slam_b200/csrc/icp_cells.cuh, the very functions the
// kernel calls) on the CPU: for random queries and model points -- anywhere in the 3x3 block, on and next to cell
// boundaries, with the grid origin up to kilometres away, for cell edges from 6 mm to 1 m -- a model point that lies in
// the neighbour (dx, dy) of the query's cell is never nearer than the edge gaps say, so skipping a cell whose gap sum
// exceeds the best squared distance so far cannot lose the nearest neighbour (or a tie).
//   icpcell_check      prints a summary; exit code 0 = no violation
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>

#include "../../ohm_tsd_slam_b200/csrc/icp_cells.cuh"

int main()
{
  std::mt19937_64 rng(4242);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  uint64_t cases = 0, bad = 0, informative = 0;
  const double hs[] = {0.1, 0.00625, 1.0, 0.1 * (1.0 + 1e-9), 0.025};
  for(int it = 0; it < 3000000; it++)
  {
    const double h = hs[it % 5], invh = 1.0 / h;
    const double bx0 = (U(rng) - 0.5) * std::ldexp(1.0, (int)(U(rng) * 12)), by0 = (U(rng) - 0.5) * std::ldexp(1.0, (int)(U(rng) * 12));
    // a query somewhere within 3000 cells of the origin of the hash; every tenth one on or next to a cell boundary
    double x = bx0 + (U(rng) - 0.5) * 6000.0 * h, y = by0 + (U(rng) - 0.5) * 6000.0 * h;
    if(it % 10 == 0) x = bx0 + std::floor((x - bx0) * invh) * h + (U(rng) - 0.5) * 1e-12 * (it % 20 ? 1.0 : 0.0);
    const int qx = tsd_icp_cell_of(x, bx0, invh), qy = tsd_icp_cell_of(y, by0, invh);
    double l2, r2, d2, u2;
    tsd_icp_edge_gaps2(x, y, bx0, by0, h, invh, qx, qy, &l2, &r2, &d2, &u2);
    for(int k = 0; k < 6; k++)
    {
      // a model point in the 3x3 block around the query (or a little beyond), sometimes on a boundary
      double mx = x + (U(rng) - 0.5) * 3.0 * h, my = y + (U(rng) - 0.5) * 3.0 * h;
      if(k == 0) mx = bx0 + (double)(qx + (it & 1 ? 1 : 0)) * h;  // the cell's own left / right edge
      if(k == 1) my = by0 + (double)(qy + (it & 2 ? 1 : 0)) * h;
      const int dx = tsd_icp_cell_of(mx, bx0, invh) - qx, dy = tsd_icp_cell_of(my, by0, invh) - qy;
      if(dx < -1 || dx > 1 || dy < -1 || dy > 1) continue;
      const double g2 = (dx < 0 ? l2 : (dx > 0 ? r2 : 0.0)) + (dy < 0 ? d2 : (dy > 0 ? u2 : 0.0));
      const double d0 = x - mx, d1 = y - my;
      double d = 0.0;
      d += d0 * d0;
      d += d1 * d1;
      cases++;
      if(g2 > 0.0) informative++;
      if(g2 > d)
      {
        if(bad < 5) fprintf(stderr, "violation: gap %.17g > d %.17g (h %g, q %.17g %.17g, m %.17g %.17g, dx %d dy %d)\n", g2, d, h, x, y, mx, my, dx, dy);
        bad++;
      }
    }
  }
  printf("%llu pairs, gap bound positive for %.1f %%, %llu violations\n", (unsigned long long)cases,
         100.0 * (double)informative / (double)cases, (unsigned long long)bad);
  return bad ? 1 : 0;
}
