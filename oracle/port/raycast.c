/* TEST INFRASTRUCTURE ONLY -- part of oracle/ (see oracle/port/port.h).
 *
 * Restatement of obvious::RayCastPolar2D (reference src/obvision/reconstruct/grid/RayCastPolar2D.cpp).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "port.h"

int port_interpolate_bilinear_one(const port_grid_t* g, const double coord[2], double* tsd);
int port_interpolate_normal_one(const port_grid_t* g, const double coord[2], double normal[2]);

/* mathbase.h:39-53 */
static inline double ob_max(double a, double b) { return (a >= b) ? a : b; }
static inline double ob_min(double a, double b) { return (a <= b) ? a : b; }

static uint64_t g_fine_steps, g_coarse_steps;

void port_raycast_steps(uint64_t* fine_steps, uint64_t* coarse_steps)
{
  if(fine_steps) *fine_steps = g_fine_steps;
  if(coarse_steps) *coarse_steps = g_coarse_steps;
}

typedef struct
{
  double xmin, ymin, xmax, ymax; /* RayCastPolar2D.cpp:128-146 */
  double idxMin, idxMax;         /* :148-149 */
} rc_state_t;

/* RayCastPolar2D.cpp:194-281.  *key receives 2*step + abort for the first event (UINT64_MAX: none), where
 * step counts iterations of the fine loop from 0. */
static int ray_cast_from_current_view(const port_grid_t* grid, const rc_state_t* rc, const double tr[2],
                                      const double ray[2], double coordinates[2], double normal[2], uint64_t* key)
{
  int32_t cellsX, cellsY, dim;
  double cellSize;
  port_grid_get_geometry(grid, &cellsX, &cellsY, &dim, &cellSize, NULL, NULL, NULL, NULL, NULL);
  const int xDim = cellsX;
  const int yDim = cellsY;

  double position[2];
  double interp = 0.0;
  *key = UINT64_MAX;

  double xmin = rc->xmin;
  double ymin = rc->ymin;
  if(fabs(ray[0]) > 10e-6) xmin = ((double)(ray[0] > 0.0 ? 0 : (xDim - 1) * cellSize) - tr[0]) / ray[0];
  if(fabs(ray[1]) > 10e-6) ymin = ((double)(ray[1] > 0.0 ? 0 : (yDim - 1) * cellSize) - tr[1]) / ray[1];
  double idxMin = ob_max(xmin, ymin);
  idxMin = ob_max(idxMin, 0.0);

  double xmax = rc->xmax;
  double ymax = rc->ymax;
  if(fabs(ray[0]) > 10e-6) xmax = ((double)(ray[0] > 0.0 ? (xDim - 1) * cellSize : 0) - tr[0]) / ray[0];
  if(fabs(ray[1]) > 10e-6) ymax = ((double)(ray[1] > 0.0 ? (yDim - 1) * cellSize : 0) - tr[1]) / ray[1];
  double idxMax = ob_min(xmax, ymax);

  idxMin = ob_max(idxMin, rc->idxMin);
  idxMax = ob_min(idxMax, rc->idxMax);

  if(idxMin >= idxMax) return 0;

  /* :223-235 traverse partitions roughly to clip minimum index */
  double partitionSize = dim;
  for(double i = idxMin; i < idxMax; i += partitionSize)
  {
    double tsd_tmp;
    position[0] = tr[0] + i * ray[0];
    position[1] = tr[1] + i * ray[1];
    g_coarse_steps++;
    int retval = port_interpolate_bilinear_one(grid, position, &tsd_tmp);
    if(retval != TSD_INTERPOLATE_EMPTYPARTITION && retval != TSD_INTERPOLATE_INVALIDINDEX) break;
    else idxMin = i;
  }

  double tsd_prev;
  position[0] = tr[0] + idxMin * ray[0];
  position[1] = tr[1] + idxMin * ray[1];
  if(port_interpolate_bilinear_one(grid, position, &tsd_prev) != TSD_INTERPOLATE_SUCCESS) tsd_prev = NAN;

  int found = 0;
  uint64_t step = 0;
  for(double i = idxMin; i <= idxMax; i += 1.0, step++)
  {
    position[0] += ray[0];
    position[1] += ray[1];
    g_fine_steps++;

    double tsd = NAN;
    if(port_interpolate_bilinear_one(grid, position, &tsd) != TSD_INTERPOLATE_SUCCESS)
    {
      tsd_prev = tsd;
      continue;
    }

    if(tsd_prev > 0 && tsd < 0)
    {
      interp = tsd_prev / (tsd_prev - tsd);
      found = 1;
      *key = 2 * step;
      break;
    }
    else if(tsd_prev < 0 && tsd > 0)
    {
      found = 0;
      *key = 2 * step + 1;
      break;
    }
    tsd_prev = tsd;
  }

  if(!found) return 0;

  coordinates[0] = position[0] + ray[0] * (interp - 1.0);
  coordinates[1] = position[1] + ray[1] * (interp - 1.0);

  return port_interpolate_normal_one(grid, coordinates, normal);
}

/* `M = T * M` with T 3x3 and M 3x1: gslcblas dgemm NoTrans x NoTrans:
 * for k: for i: temp = 1.0*A[i,k]; if(temp != 0.0) C[i] += temp*B[k].  Only rows 0,1 are read back. */
static void mat3_times_col(const double T[9], const double v[3], double out[3])
{
  out[0] = out[1] = out[2] = 0.0;
  for(int k = 0; k < 3; k++)
    for(int i = 0; i < 3; i++)
    {
      const double temp = 1.0 * T[3 * i + k];
      if(temp != 0.0) out[i] += temp * v[k];
    }
}

static void rc_setup(const port_grid_t* g, const tsd_scan_t* s, rc_state_t* rc)
{
  double cellSize, minX, maxX, minY, maxY;
  port_grid_get_geometry(g, NULL, NULL, NULL, &cellSize, &minX, &maxX, &minY, &maxY, NULL);
  const double tr[2] = {s->pose[2], s->pose[5]};
  /* TsdGrid.h:342-347 isInsideGrid; RayCastPolar2D.cpp:128-146 */
  if(tr[0] > minX && tr[0] < maxX && tr[1] > minY && tr[1] < maxY)
  {
    rc->xmin = -10e9; rc->ymin = -10e9; rc->xmax = 10e9; rc->ymax = 10e9;
  }
  else
  {
    rc->xmin = 10e9; rc->ymin = 10e9; rc->xmax = -10e9; rc->ymax = -10e9;
  }
  rc->idxMin = s->min_range / cellSize;
  rc->idxMax = s->max_range / cellSize;
}

/* RayCastPolar2D.cpp:113-192 */
int port_raycast_mask(port_grid_t* g, const tsd_scan_t* s, const double* rays_world, double* coords,
                      double* normals, uint8_t* mask, uint32_t* count)
{
  unsigned int cnt = 0;
  const double* T = s->pose_inv; /* Matrix T = sensor->getTransformation(); T.invert(); */
  const double tr[2] = {s->pose[2], s->pose[5]};
  rc_state_t rc;
  rc_setup(g, s, &rc);
  g_fine_steps = g_coarse_steps = 0;

  double M[3], N[3], c[2], n[2];
  M[2] = 1.0;
  N[2] = 0.0;
  for(int beam = 0; beam < s->n; beam++)
  {
    double ray[2] = {rays_world[beam], rays_world[s->n + beam]};
    uint64_t key;
    if(ray_cast_from_current_view(g, &rc, tr, ray, c, n, &key))
    {
      double Mo[3], No[3];
      M[0] = c[0]; M[1] = c[1];
      N[0] = n[0]; N[1] = n[1];
      mat3_times_col(T, M, Mo);
      mat3_times_col(T, N, No);
      coords[2 * beam] = Mo[0];
      coords[2 * beam + 1] = Mo[1];
      normals[2 * beam] = No[0];
      normals[2 * beam + 1] = No[1];
      mask[beam] = 1;
      cnt++;
    }
    else
    {
      mask[beam] = 0;
    }
  }
  if(count) *count = cnt;
  return TSD_OK;
}

void port_raycast_keys(port_grid_t* g, const tsd_scan_t* s, const double* rays_world, uint64_t* keys)
{
  const double tr[2] = {s->pose[2], s->pose[5]};
  rc_state_t rc;
  rc_setup(g, s, &rc);
  double c[2], n[2];
  for(int beam = 0; beam < s->n; beam++)
  {
    double ray[2] = {rays_world[beam], rays_world[s->n + beam]};
    ray_cast_from_current_view(g, &rc, tr, ray, c, n, &keys[beam]);
  }
}
