// Cell arithmetic of k_icp's nearest-neighbour search (icp.cu): the cell of a point in the spatial hash and the squared
// distances from a query to the edges of its own cell, which decide whether a neighbouring cell has to be opened.
// Host/device code so that tests/cpp/icpcell_check.cpp can run exactly these functions on the CPU: whatever model point
// lies (by tsd_icp_cell_of) in the neighbour (dx, dy) of the query's cell is at a squared distance, computed as the
// search computes it, NOT BELOW gap2(dx) + gap2(dy) -- so a cell skipped because that sum exceeds the best distance so
// far cannot hold a point as near.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define TSD_IC_HD __host__ __device__ __forceinline__
#else
#define TSD_IC_HD static inline
#endif

// cell coordinate of a point; clamped so that the hash input stays small and rings never overflow
TSD_IC_HD int tsd_icp_cell_of(double v, double v0, double invh)
{
  const double t = floor((v - v0) * invh);
  return (int)fmin(fmax(t, -1.0e6), 1.0e6);
}

// squared distances of (x, y) to the left / right / lower / upper edge of its own cell (qx, qy), shaved so that the
// rounding of the cell computation never claims more distance than there is
TSD_IC_HD void tsd_icp_edge_gaps2(double x, double y, double bx0, double by0, double h, double invh, int qx, int qy, double* l2,
                                  double* r2, double* d2, double* u2)
{
  const double fx = (x - bx0) * invh - (double)qx, fy = (y - by0) * invh - (double)qy;
  const double shave = 1e-9 * h;
  const double gl = fmax(fx * h - shave, 0.0), gr = fmax((1.0 - fx) * h - shave, 0.0);
  const double gd = fmax(fy * h - shave, 0.0), gu = fmax((1.0 - fy) * h - shave, 0.0);
  *l2 = gl * gl; *r2 = gr * gr; *d2 = gd * gd; *u2 = gu * gu;
}
