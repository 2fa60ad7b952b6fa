/* TEST INFRASTRUCTURE ONLY -- part of oracle/ (see oracle/port/port.h).
 *
 * Restatement of obvious::TsdGrid / TsdGridPartition / TsdGridComponent::isInRange /
 * SensorPolar2D::backProject.  Citations are file:line in /root/reference/src/obvision/reconstruct/grid
 * unless a longer path is given.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "port.h"

#define MAXWEIGHT 32.0 /* reconstruct_defs.h:4 TSDGRIDMAXWEIGHT */
#define TSDINC 1.0     /* reconstruct_defs.h:6 */

typedef struct
{
  double tsd;
  double weight;
} cell_t; /* TsdGridPartition.h:16-20 */

typedef struct
{
  cell_t* grid;       /* (dim+1) x (dim+1), NULL until init (TsdGridPartition.cpp:97) */
  int initialized;    /* TsdGridPartition.h:160 */
  double init_weight; /* TsdGridPartition.h:158 */
  unsigned int x, y;  /* first cell index (TsdGridPartition.h:154-156) */
  double edge[4][2];  /* TsdGridPartition.cpp:48-62 */
  double centroid[2]; /* :65-66 */
  double circumradius;/* :68-70 */
  double max_truncation, inv_max_truncation, eps; /* :93-95 */
} part_t;

struct port_grid
{
  int cells_x, cells_y, dim, parts_x, parts_y;
  double cell_size, inv_cell_size, max_truncation;
  double min_x, max_x, min_y, max_y;
  part_t* parts;
  tsd_push_stats_t stats;
};

/* mathbase.h:39-53 -- note the NaN asymmetry: min(a,b) returns b when a is NaN */
static inline double ob_min(double a, double b) { return (a <= b) ? a : b; }

/* TsdGrid.cpp:112-169 init + TsdGridPartition.cpp:15-74 ctor */
port_grid_t* port_grid_create(double cell_size, int layout_partition, int layout_grid)
{
  port_grid_t* g = (port_grid_t*)calloc(1, sizeof(*g));
  g->cell_size = cell_size;
  g->inv_cell_size = 1.0 / cell_size;
  g->cells_x = 1 << layout_grid;
  g->cells_y = g->cells_x;
  g->dim = 1 << layout_partition;
  if(g->dim > g->cells_x) { free(g); return NULL; }
  g->parts_x = g->cells_x / g->dim;
  g->parts_y = g->cells_y / g->dim;
  g->max_truncation = 2.0 * cell_size;
  g->min_x = 0.0;
  g->max_x = ((double)g->cells_x + 0.5) * cell_size;
  g->min_y = 0.0;
  g->max_y = ((double)g->cells_y + 0.5) * cell_size;
  g->parts = (part_t*)calloc((size_t)g->parts_x * g->parts_y, sizeof(part_t));
  for(int py = 0; py < g->parts_y; py++)
    for(int px = 0; px < g->parts_x; px++)
    {
      part_t* p = &g->parts[py * g->parts_x + px];
      const unsigned int x = px * g->dim, y = py * g->dim;
      const unsigned int cx = g->dim, cy = g->dim;
      p->x = x;
      p->y = y;
      p->edge[0][0] = ((double)x + 0.5) * cell_size;
      p->edge[0][1] = ((double)y + 0.5) * cell_size;
      p->edge[1][0] = ((double)(x + cx) + 0.5) * cell_size;
      p->edge[1][1] = ((double)y + 0.5) * cell_size;
      p->edge[2][0] = ((double)x + 0.5) * cell_size;
      p->edge[2][1] = ((double)(y + cy) + 0.5) * cell_size;
      p->edge[3][0] = ((double)(x + cx) + 0.5) * cell_size;
      p->edge[3][1] = ((double)(y + cy) + 0.5) * cell_size;
      p->centroid[0] = (p->edge[0][0] + p->edge[1][0] + p->edge[2][0] + p->edge[3][0]) / 4.0;
      p->centroid[1] = (p->edge[0][1] + p->edge[1][1] + p->edge[2][1] + p->edge[3][1]) / 4.0;
      const double dx = p->edge[3][0] - p->edge[0][0];
      const double dy = p->edge[3][1] - p->edge[0][1];
      p->circumradius = sqrt(dx * dx + dy * dy) * 0.5;
      p->init_weight = 0.0;
    }
  return g;
}

void port_grid_destroy(port_grid_t* g)
{
  if(!g) return;
  for(int i = 0; i < g->parts_x * g->parts_y; i++) free(g->parts[i].grid);
  free(g->parts);
  free(g);
}

/* TsdGrid.cpp:206-215 */
void port_grid_set_max_truncation(port_grid_t* g, double val)
{
  if(val < 2 * g->cell_size) val = 2 * g->cell_size;
  g->max_truncation = val;
}

void port_grid_get_geometry(const port_grid_t* g, int32_t* cells_x, int32_t* cells_y, int32_t* partition_size,
                            double* cell_size, double* min_x, double* max_x, double* min_y, double* max_y,
                            double* max_truncation)
{
  if(cells_x) *cells_x = g->cells_x;
  if(cells_y) *cells_y = g->cells_y;
  if(partition_size) *partition_size = g->dim;
  if(cell_size) *cell_size = g->cell_size;
  if(min_x) *min_x = g->min_x;
  if(max_x) *max_x = g->max_x;
  if(min_y) *min_y = g->min_y;
  if(max_y) *max_y = g->max_y;
  if(max_truncation) *max_truncation = g->max_truncation;
}

/* TsdGridPartition.cpp:88-134 (the homogeneous cell coordinates are recomputed on the fly in push) */
static void part_init(port_grid_t* g, part_t* p, double max_truncation)
{
  if(p->initialized) return;
  p->max_truncation = max_truncation;
  p->inv_max_truncation = 1.0 / max_truncation;
  p->eps = -g->cell_size / 2.0;
  const int n = (g->dim + 1) * (g->dim + 1);
  p->grid = (cell_t*)malloc(n * sizeof(cell_t));
  if(p->init_weight > 0.0)
  {
    for(int i = 0; i < n; i++) { p->grid[i].tsd = 1.0; p->grid[i].weight = p->init_weight; }
  }
  else
  {
    for(int i = 0; i < n; i++) { p->grid[i].tsd = NAN; p->grid[i].weight = p->init_weight; }
  }
  p->initialized = 1;
  g->stats.newly_initialized++;
}

/* TsdGridPartition.cpp:136-164 */
static void part_increase_emptiness(port_grid_t* g, part_t* p)
{
  if(p->initialized)
  {
    const int n = (g->dim + 1) * (g->dim + 1);
    for(int i = 0; i < n; i++)
    {
      cell_t* cell = &p->grid[i];
      if(isnan(cell->tsd))
      {
        cell->weight += 1.0;
        cell->tsd = 1.0;
      }
      else
      {
        cell->weight = ob_min(cell->weight + 1, MAXWEIGHT);
        cell->tsd = (cell->tsd * (cell->weight - 1.0) + 1.0) / cell->weight;
      }
    }
    g->stats.cell_updates += (uint64_t)n;
  }
  else
  {
    p->init_weight += 1.0;
    p->init_weight = ob_min(p->init_weight, MAXWEIGHT);
  }
  g->stats.emptied_tiles++;
}

/* TsdGridPartition.h:170-212 */
static void part_add_tsd(port_grid_t* g, part_t* p, unsigned int x, unsigned int y, double sd, double weight)
{
  if(sd >= -p->max_truncation)
  {
    cell_t* cell = &p->grid[y * (g->dim + 1) + x];
    double tsd = ob_min(sd * p->inv_max_truncation, TSDINC);
    double w = 0.01;
    if(fabs(sd) < p->eps) w = 1.0; /* eps is negative: never true */
    w *= weight;
    if(isnan(cell->tsd))
    {
      cell->tsd = tsd;
      cell->weight += w;
    }
    else
    {
      cell->tsd = (cell->tsd * cell->weight + tsd * w) / (cell->weight + w);
      cell->weight = ob_min(cell->weight + w, MAXWEIGHT);
    }
    g->stats.cell_updates++;
  }
}

/* SensorPolar2D.cpp:117-135.  coords2D = PoseInv * M^T through gslcblas dgemm NoTrans x Trans:
 * temp = 0; temp += A[i,k]*B[j,k] (k = 0,1,2); C = 0 + 1.0*temp  (SURVEY.md App. A.2). */
static int back_project_one(const tsd_scan_t* s, double X, double Y)
{
  const double* P = s->pose_inv;
  double tx = 0.0;
  tx += P[0] * X;
  tx += P[1] * Y;
  tx += P[2] * 1.0;
  double ty = 0.0;
  ty += P[3] * X;
  ty += P[4] * Y;
  ty += P[5] * 1.0;
  const double cx = 0.0 + 1.0 * tx;
  const double cy = 0.0 + 1.0 * ty;
  const double angular_res_inv = 1.0 / s->angular_res;
  const double phi = atan2(cy, cx);
  if(phi <= s->phi_lower) return -2;
  if(phi >= s->phi_upper) return -1;
  return (int)round((phi - s->phi_min) * angular_res_inv);
}

void port_back_project(const tsd_scan_t* scan, int32_t n, const double* xy, int32_t* idx)
{
  for(int i = 0; i < n; i++) idx[i] = back_project_one(scan, xy[2 * i], xy[2 * i + 1]);
}

/* TsdGridComponent.cpp:43-124 (leaf branch) */
static int part_is_in_range(port_grid_t* g, part_t* p, const double pos[2], const tsd_scan_t* s, double max_truncation)
{
  /* mathbase.h:369-378 euklideanDistance */
  double sqr = 0.0;
  for(int i = 0; i < 2; i++)
  {
    double tmp = pos[i] - p->centroid[i];
    sqr += tmp * tmp;
  }
  const double distance = sqrt(sqr);

  double closestVoxelDist = distance - p->circumradius - max_truncation;
  if(closestVoxelDist > s->max_range) return 0;
  double farthestVoxelDist = distance + p->circumradius + max_truncation;
  if(farthestVoxelDist < s->min_range) return 0;

  const double* data = s->ranges;
  const uint8_t* mask = s->mask;
  const int measurements = s->n;

  int idxEdge[4];
  for(int i = 0; i < 4; i++) idxEdge[i] = back_project_one(s, p->edge[i][0], p->edge[i][1]);

  int isAnyEdgeVisible = 0;
  int areAllEdgesVisible = 1;
  for(int i = 0; i < 4; i++)
  {
    if(idxEdge[i] == -1) { idxEdge[i] = measurements - 1; areAllEdgesVisible = 0; }
    else if(idxEdge[i] == -2) { idxEdge[i] = 0; areAllEdgesVisible = 0; }
    else isAnyEdgeVisible = 1;
  }
  if(!isAnyEdgeVisible) return 0;

  /* mathbase.h:55-64 minmaxArray */
  int minIdx = idxEdge[0], maxIdx = idxEdge[0];
  for(int i = 1; i < 4; i++)
  {
    if(minIdx > idxEdge[i]) minIdx = idxEdge[i];
    else if(maxIdx < idxEdge[i]) maxIdx = idxEdge[i];
  }

  int isVisible = 0;
  for(int j = minIdx; j <= maxIdx; j++) isVisible = isVisible || ((data[j] > closestVoxelDist) && mask[j]);
  if(!isVisible) return 0;

  if(areAllEdgesVisible)
  {
    int isEmpty = 1;
    for(int j = minIdx; j <= maxIdx; j++)
    {
      if(isinf(data[j])) isEmpty = isEmpty && (distance < s->low_reflectivity_range);
      else isEmpty = isEmpty && (data[j] > farthestVoxelDist) && mask[j];
    }
    if(isEmpty)
    {
      part_increase_emptiness(g, p);
      return 0;
    }
  }
  return 1;
}

/* TsdGrid.cpp:372-427 */
static void propagate_borders(port_grid_t* g)
{
  const int width = g->dim, height = g->dim, pitch = g->dim + 1;
  for(int py = 0; py < g->parts_y; py++)
    for(int px = 0; px < g->parts_x; px++)
    {
      part_t* cur = &g->parts[py * g->parts_x + px];
      if(!cur->initialized) continue;
      if(px < g->parts_x - 1)
      {
        part_t* right = &g->parts[py * g->parts_x + px + 1];
        if(right->initialized)
          for(int i = 0; i < height; i++) cur->grid[i * pitch + width] = right->grid[i * pitch + 0];
      }
      if(py < g->parts_y - 1)
      {
        part_t* up = &g->parts[(py + 1) * g->parts_x + px];
        if(up->initialized)
          for(int i = 0; i < width; i++) cur->grid[height * pitch + i] = up->grid[0 * pitch + i];
      }
      if(px < g->parts_x - 1 && py < g->parts_y - 1)
      {
        part_t* upRight = &g->parts[(py + 1) * g->parts_x + px + 1];
        if(upRight->initialized) cur->grid[height * pitch + width] = upRight->grid[0];
      }
    }
}

/* TsdGrid.cpp:217-284 */
void port_grid_push(port_grid_t* g, const tsd_scan_t* s)
{
  const double* data = s->ranges;
  const uint8_t* mask = s->mask;
  double tr[2] = {s->pose[2], s->pose[5]}; /* Sensor.cpp:114-118 getPosition */
  const unsigned int partSize = g->dim * g->dim;
  memset(&g->stats, 0, sizeof(g->stats));

  for(unsigned int i = 0; i < (unsigned int)(g->parts_x * g->parts_y); i++)
  {
    part_t* part = &g->parts[i];
    if(!part_is_in_range(g, part, tr, s, g->max_truncation)) continue;

    part_init(g, part, g->max_truncation);
    g->stats.active_tiles++;
    g->stats.cell_visits += partSize;

    const double* partCentroid = part->centroid;
    double distCentroid = sqrt((partCentroid[0] - tr[0]) * (partCentroid[0] - tr[0]) +
                               (partCentroid[1] - tr[1]) * (partCentroid[1] - tr[1]));
    if(distCentroid > s->max_range) distCentroid = s->max_range;
    double partWeight = (s->max_range - distCentroid) / s->max_range;
    partWeight *= partWeight;

    const double lowReflectivityRange = s->low_reflectivity_range;

    unsigned int c = 0;
    for(unsigned int iy = part->y; iy < part->y + g->dim; iy++)
      for(unsigned int ix = part->x; ix < part->x + g->dim; ix++, c++)
      {
        /* TsdGridPartition.cpp:127-128 */
        const double X = ((double)ix + 0.5) * g->cell_size;
        const double Y = ((double)iy + 0.5) * g->cell_size;
        const int index = back_project_one(s, X, Y);
        if(index >= 0)
        {
          if(mask[index])
          {
            if(!isinf(data[index]))
            {
              const double sd = data[index] - sqrt((X - tr[0]) * (X - tr[0]) + (Y - tr[1]) * (Y - tr[1]));
              part_add_tsd(g, part, ix - part->x, iy - part->y, sd, partWeight);
            }
            else
            {
              const double dist = sqrt((X - tr[0]) * (X - tr[0]) + (Y - tr[1]) * (Y - tr[1]));
              if(dist < lowReflectivityRange)
                part_add_tsd(g, part, ix - part->x, iy - part->y, g->max_truncation, partWeight);
            }
          }
        }
      }
  }
  propagate_borders(g);
}

void port_grid_last_push_stats(port_grid_t* g, tsd_push_stats_t* out) { *out = g->stats; }

/* TsdGrid.cpp:609-638 */
int port_grid_free_footprint(port_grid_t* g, double cx, double cy, double width, double height)
{
  unsigned int minX = (unsigned int)((cx - width * 0.5) / g->cell_size + 0.5);
  unsigned int maxX = (unsigned int)((cx + width * 0.5) / g->cell_size + 0.5);
  unsigned int minY = (unsigned int)((cy - height * 0.5) / g->cell_size + 0.5);
  unsigned int maxY = (unsigned int)((cy + height * 0.5) / g->cell_size + 0.5);
  if((minX > (unsigned int)g->cells_x) || (maxX > (unsigned int)g->cells_x) || (minY > (unsigned int)g->cells_y) ||
     (maxY > (unsigned int)g->cells_y))
    return TSD_E_RANGE;
  const unsigned int dim = (unsigned int)g->dim;
  for(unsigned int rows = minY; rows < maxY; rows++)
    for(unsigned int cols = minX; cols < maxX; cols++)
    {
      unsigned int py = rows / dim, px = cols / dim;
      part_t* p = &g->parts[py * g->parts_x + px];
      if(!p->initialized) part_init(g, p, g->max_truncation);
      unsigned int cyy = rows % dim, cxx = cols % dim;
      p->grid[cyy * (dim + 1) + cxx].tsd = TSDINC;
    }
  return TSD_OK;
}

/* TsdGrid.h:306-340 */
static int coord2cell(const port_grid_t* g, const double coord[2], int* p, int* x, int* y, double* dx, double* dy)
{
  const double dCoordX = coord[0] * g->inv_cell_size;
  const double dCoordY = coord[1] * g->inv_cell_size;
  int xIdx = (int)floor(dCoordX);
  int yIdx = (int)floor(dCoordY);
  *dx = ((double)xIdx + 0.5) * g->cell_size;
  *dy = ((double)yIdx + 0.5) * g->cell_size;
  if(coord[0] < *dx) { xIdx--; (*dx) -= g->cell_size; }
  if(coord[1] < *dy) { yIdx--; (*dy) -= g->cell_size; }
  if((xIdx >= g->cells_x) || (xIdx < 0) || (yIdx >= g->cells_y) || (yIdx < 0)) return 0;
  *p = yIdx / g->dim * g->parts_x + xIdx / g->dim;
  *x = xIdx % g->dim;
  *y = yIdx % g->dim;
  return 1;
}

/* TsdGrid.h:284-304 + TsdGridPartition.h:214-221 */
int port_interpolate_bilinear_one(const port_grid_t* g, const double coord[2], double* tsd)
{
  int p, x, y;
  double dx, dy;
  if(!coord2cell(g, coord, &p, &x, &y, &dx, &dy)) return TSD_INTERPOLATE_INVALIDINDEX;
  const part_t* part = &g->parts[p];
  if(!part->initialized) return TSD_INTERPOLATE_EMPTYPARTITION;
  const double wx = fabs((coord[0] - dx) * g->inv_cell_size);
  const double wy = fabs((coord[1] - dy) * g->inv_cell_size);
  const int pitch = g->dim + 1;
  const cell_t* c = part->grid;
  *tsd = c[y * pitch + x].tsd * (1. - wy) * (1. - wx) + c[(y + 1) * pitch + (x + 0)].tsd * wy * (1. - wx) +
         c[(y + 0) * pitch + (x + 1)].tsd * (1. - wy) * wx + c[(y + 1) * pitch + (x + 1)].tsd * wy * wx;
  if(isnan(*tsd)) return TSD_INTERPOLATE_ISNAN;
  return TSD_INTERPOLATE_SUCCESS;
}

/* TsdGrid.cpp:517-546 + mathbase.h:212-218 norm2 */
int port_interpolate_normal_one(const port_grid_t* g, const double coord[2], double normal[2])
{
  double neighbor[2];
  double depthInc = 0, depthDec = 0;
  neighbor[0] = coord[0] + g->cell_size;
  neighbor[1] = coord[1];
  if(port_interpolate_bilinear_one(g, neighbor, &depthInc) != TSD_INTERPOLATE_SUCCESS) return 0;
  neighbor[0] = coord[0] - g->cell_size;
  if(port_interpolate_bilinear_one(g, neighbor, &depthDec) != TSD_INTERPOLATE_SUCCESS) return 0;
  normal[0] = depthInc - depthDec;
  neighbor[0] = coord[0];
  neighbor[1] = coord[1] + g->cell_size;
  if(port_interpolate_bilinear_one(g, neighbor, &depthInc) != TSD_INTERPOLATE_SUCCESS) return 0;
  neighbor[1] = coord[1] - g->cell_size;
  if(port_interpolate_bilinear_one(g, neighbor, &depthDec) != TSD_INTERPOLATE_SUCCESS) return 0;
  normal[1] = depthInc - depthDec;
  double len = sqrt(normal[0] * normal[0] + normal[1] * normal[1]);
  if(fabs(len) <= 10e-6) return 1;
  normal[0] /= len;
  normal[1] /= len;
  return 1;
}

void port_grid_interpolate_bilinear(port_grid_t* g, int32_t n, const double* xy, double* tsd, int32_t* status)
{
  for(int i = 0; i < n; i++)
  {
    double v = NAN;
    status[i] = port_interpolate_bilinear_one(g, &xy[2 * i], &v);
    tsd[i] = v;
  }
}

void port_grid_interpolate_normal(port_grid_t* g, int32_t n, const double* xy, double* normals, int32_t* ok)
{
  for(int i = 0; i < n; i++)
  {
    double nn[2] = {NAN, NAN};
    ok[i] = port_interpolate_normal_one(g, &xy[2 * i], nn);
    normals[2 * i] = nn[0];
    normals[2 * i + 1] = nn[1];
  }
}

int32_t port_grid_num_partitions(const port_grid_t* g) { return g->parts_x * g->parts_y; }

/* TsdGridPartition.h:66,72 isInitialized / isEmpty */
void port_grid_partition_states(port_grid_t* g, int32_t* state, double* init_weight)
{
  for(int i = 0; i < g->parts_x * g->parts_y; i++)
  {
    const part_t* p = &g->parts[i];
    state[i] = p->initialized ? TSD_PARTITION_CONTENT
                              : (p->init_weight > 0.0 ? TSD_PARTITION_EMPTY : TSD_PARTITION_UNINITIALIZED);
    if(init_weight) init_weight[i] = p->init_weight;
  }
}

int port_grid_download_partition(port_grid_t* g, int32_t p, double* tsd, double* weight)
{
  const part_t* part = &g->parts[p];
  if(!part->initialized) return TSD_E_INVALID;
  const int n = (g->dim + 1) * (g->dim + 1);
  for(int i = 0; i < n; i++) { tsd[i] = part->grid[i].tsd; weight[i] = part->grid[i].weight; }
  return TSD_OK;
}

int port_grid_upload_partition(port_grid_t* g, int32_t p, const double* tsd, const double* weight)
{
  part_t* part = &g->parts[p];
  if(!part->initialized) { part_init(g, part, g->max_truncation); }
  const int n = (g->dim + 1) * (g->dim + 1);
  for(int i = 0; i < n; i++) { part->grid[i].tsd = tsd[i]; part->grid[i].weight = weight[i]; }
  return TSD_OK;
}

void port_grid_fill(port_grid_t* g, double tsd, double weight, int only_uninitialized)
{
  const int n = (g->dim + 1) * (g->dim + 1);
  for(int i = 0; i < g->parts_x * g->parts_y; i++)
  {
    part_t* part = &g->parts[i];
    if(part->initialized && only_uninitialized) continue;
    if(!part->initialized) part_init(g, part, g->max_truncation);
    for(int k = 0; k < n; k++) { part->grid[k].tsd = tsd; part->grid[k].weight = weight; }
  }
}

/* gsl/Matrix.cpp:168-179 through the shim's LU (oracle/shim/gsl_shim.c): partial pivoting, reciprocal
 * scaling, column-wise solves.  Same statement order as tsd_invert3x3 in the product's host code. */
void port_invert3x3(const double in[9], double out[9])
{
  double A[3][3];
  int perm[3] = {0, 1, 2};
  for(int i = 0; i < 3; i++) for(int j = 0; j < 3; j++) A[i][j] = in[3 * i + j];
  for(int j = 0; j < 3; j++)
  {
    double max = fabs(A[j][j]);
    int ip = j;
    for(int i = j + 1; i < 3; i++)
    {
      double a = fabs(A[i][j]);
      if(a > max) { max = a; ip = i; }
    }
    if(ip != j)
    {
      for(int k = 0; k < 3; k++) { double t = A[j][k]; A[j][k] = A[ip][k]; A[ip][k] = t; }
      int t = perm[j]; perm[j] = perm[ip]; perm[ip] = t;
    }
    const double ajj = A[j][j];
    if(fabs(ajj) >= 2.2250738585072014e-308)
    {
      const double inv = 1.0 / ajj;
      for(int i = j + 1; i < 3; i++) A[i][j] *= inv;
    }
    else
    {
      for(int i = j + 1; i < 3; i++) A[i][j] /= ajj;
    }
    for(int i = j + 1; i < 3; i++)
    {
      const double tmp = -1.0 * A[i][j];
      for(int k = j + 1; k < 3; k++) A[i][k] += A[j][k] * tmp;
    }
  }
  for(int c = 0; c < 3; c++)
  {
    double e[3], x[3];
    for(int i = 0; i < 3; i++) e[i] = (i == c) ? 1.0 : 0.0;
    for(int i = 0; i < 3; i++) x[i] = e[perm[i]];
    for(int i = 1; i < 3; i++)
    {
      double t = x[i];
      for(int j = 0; j < i; j++) t -= A[i][j] * x[j];
      x[i] = t;
    }
    x[2] = x[2] / A[2][2];
    for(int i = 1; i >= 0; i--)
    {
      double t = x[i];
      for(int j = i + 1; j < 3; j++) t -= A[i][j] * x[j];
      x[i] = t / A[i][i];
    }
    for(int i = 0; i < 3; i++) out[3 * i + c] = x[i];
  }
}

/* ---------------------------------------------------------------- map publication (ThreadGrid.cpp:84,125) */

/* RayCastAxisAligned2D::calcCoords (RayCastAxisAligned2D.cpp:13-105): zero crossings of the TSD along the cell
 * rows and columns of every allocated inner partition (borders included), plus an occupancy grid.  coords gets
 * 2 doubles per crossing, *cnt counts doubles (as the reference does).  normals / occupied may be NULL.
 * Reference quirk kept: the normal is always evaluated at coords[0..1] -- the FIRST crossing -- because the
 * call passes the array base (RayCastAxisAligned2D.cpp:52,73), and it is left untouched when that fails. */
void port_axis_map(port_grid_t* g, double* coords, double* normals, uint32_t* cnt, int8_t* occupied)
{
  const unsigned int partitionsInX = (unsigned)(g->cells_x / g->dim);
  const unsigned int partitionsInY = partitionsInX;
  const double cellSize = g->cell_size;
  const unsigned int dim = (unsigned)g->dim;
  const unsigned int cellsPPart = dim * dim;
  const unsigned int cellsPPX = dim;
  const unsigned int pitch = dim + 1;
  unsigned int gridOffset = 0;
  *cnt = 0;
  for(unsigned int y = 1; y < partitionsInY - 1; y++)
  {
    for(unsigned int x = 1; x < partitionsInX - 1; x++)
    {
      const part_t* p = &g->parts[y * partitionsInX + x];
      if(p->initialized)
      {
        /* isEmpty() is false for an initialised partition (TsdGridPartition.h:72) */
        if(occupied) gridOffset = y * cellsPPart * partitionsInX + x * cellsPPX;
        for(unsigned int py = 0; py < dim + 1; py++)
        {
          double tsd_prev = p->grid[py * pitch + 0].tsd;
          double interp = 0.0;
          if(occupied) occupied[gridOffset + py * (unsigned)g->cells_x] = ((tsd_prev > 0.0) ? 0 : -1);
          for(unsigned int px = 1; px < dim + 1; px++)
          {
            const double tsd = p->grid[py * pitch + px].tsd;
            if(occupied) occupied[gridOffset + py * (unsigned)g->cells_x + px] = ((tsd > 0.0) ? 0 : -1);
            if((tsd_prev > 0 && tsd < 0) || (tsd_prev < 0 && tsd > 0))
            {
              interp = tsd_prev / (tsd_prev - tsd);
              coords[(*cnt)] = px * cellSize + cellSize * (interp - 1.0) + (x * dim) * cellSize;
              coords[(*cnt) + 1] = py * cellSize + (y * dim) * cellSize;
              if(normals) port_interpolate_normal_one(g, coords, &normals[*cnt]);
              (*cnt) += 2;
            }
            tsd_prev = tsd;
          }
        }
        for(unsigned int px = 0; px < dim + 1; px++)
        {
          double tsd_prev = p->grid[0 * pitch + px].tsd;
          double interp = 0.0;
          for(unsigned int py = 1; py < dim + 1; py++)
          {
            const double tsd = p->grid[py * pitch + px].tsd;
            if((tsd_prev > 0 && tsd < 0) || (tsd_prev < 0 && tsd > 0))
            {
              interp = tsd_prev / (tsd_prev - tsd);
              coords[(*cnt)] = px * cellSize + (x * dim) * cellSize;
              coords[(*cnt) + 1] = py * cellSize + cellSize * (interp - 1.0) + (y * dim) * cellSize;
              if(normals) port_interpolate_normal_one(g, coords, &normals[*cnt]);
              (*cnt) += 2;
            }
            tsd_prev = tsd;
          }
        }
      }
      else if(p->init_weight > 0.0) /* isEmpty(): seen as free space, never allocated */
      {
        if(occupied)
        {
          gridOffset = y * cellsPPart * partitionsInX + x * cellsPPX;
          for(unsigned int py = 0; py < dim; py++)
            for(unsigned int px = 0; px < dim; px++) occupied[gridOffset + py * (unsigned)g->cells_x + px] = 0;
        }
      }
    }
  }
}

/* TsdGrid::grid2ColorImage (TsdGrid.cpp:429-488); px / py are running sums, as in the reference */
void port_color_image(port_grid_t* g, uint8_t* image, uint32_t width, uint32_t height)
{
  unsigned char rgb[3];
  const double stepW = g->max_x / (double)width;
  const double stepH = g->max_y / (double)height;
  double py = 0.0;
  unsigned int i = 0;
  for(unsigned int h = 0; h < height; h++)
  {
    double px = 0.0;
    for(unsigned int w = 0; w < width; w++, i++)
    {
      double coord[2] = {px, py};
      int p, x, y;
      double dx, dy;
      double tsd = NAN;
      int isEmpty = 0;
      if(coord2cell(g, coord, &p, &x, &y, &dx, &dy))
      {
        const part_t* part = &g->parts[p];
        if(part->initialized) tsd = part->grid[y * (g->dim + 1) + x].tsd;
        isEmpty = (!part->initialized && part->init_weight > 0.0);
      }
      if(tsd > 0.0)
      {
        rgb[0] = (unsigned char)(tsd * 255.0);
        rgb[1] = 255;
        rgb[2] = (unsigned char)(tsd * 255.0);
      }
      else if(tsd < 0.0)
      {
        rgb[0] = (unsigned char)((1.0 + tsd) * 255.0);
        rgb[1] = 0;
        rgb[2] = 0;
      }
      else if(isEmpty) { rgb[0] = 255; rgb[1] = 255; rgb[2] = 255; }
      else { rgb[0] = 0; rgb[1] = 0; rgb[2] = 0; }
      memcpy(&image[3 * i], rgb, 3);
      px += stepW;
    }
    py += stepH;
  }
}

/* ---------------------------------------------------------------- checkpoint format (TsdGrid.cpp:25-110, :548-607) */

/* TsdGrid::storeGrid: text, one value per line, the stream's default formatting (6 significant digits, "%g") */
int port_grid_store(port_grid_t* g, const char* path)
{
  if(!path || !path[0]) return 0;
  FILE* f = fopen(path, "w");
  if(!f) return 0;
  int layout_partition = 0, layout_grid = 0;
  while((1 << layout_partition) < g->dim) layout_partition++;
  while((1 << layout_grid) < g->cells_x) layout_grid++;
  fprintf(f, "%g\n%d\n%d\n%g\n", g->cell_size, layout_partition, layout_grid, g->max_truncation);
  const int pitch = g->dim + 1;
  for(int y = 0; y < g->parts_y; y++)
    for(int x = 0; x < g->parts_x; x++)
    {
      const part_t* p = &g->parts[y * g->parts_x + x];
      if(p->initialized)
      {
        fprintf(f, "2\n"); /* CONTENT */
        for(int py = 0; py < g->dim; py++)
          for(int px = 0; px < g->dim; px++) fprintf(f, "%g\n%g\n", p->grid[py * pitch + px].tsd, p->grid[py * pitch + px].weight);
      }
      else if(p->init_weight > 0.0) fprintf(f, "1\n%g\n", p->init_weight); /* EMPTY */
      else fprintf(f, "0\n");                                               /* UNINITIALIZED */
    }
  fclose(f);
  return 1;
}

static double get_double_line(FILE* f) /* tools.cpp:190-200 */
{
  char line[1024];
  if(!fgets(line, sizeof(line), f)) return NAN;
  if(line[0] == '\n' || line[0] == 0) return NAN;
  return strtod(line, NULL);
}

static int get_int_line(FILE* f) /* tools.cpp:207-215 */
{
  char line[1024];
  if(!fgets(line, sizeof(line), f)) return 0;
  if(line[0] == '\n' || line[0] == 0) return 0;
  return atoi(line);
}

/* TsdGrid::TsdGrid(const std::string&, FILE_SOURCE) */
port_grid_t* port_grid_load(const char* path)
{
  FILE* f = fopen(path, "r");
  if(!f) return NULL;
  const double cellSize = get_double_line(f);
  const int layoutPartition = get_int_line(f);
  const int layoutGrid = get_int_line(f);
  if(layoutGrid < 0 || layoutPartition < 0 || layoutGrid > 15 || layoutPartition > 15) { fclose(f); return NULL; }
  const double maxTruncation = get_double_line(f);
  port_grid_t* g = port_grid_create(cellSize, layoutPartition, layoutGrid);
  if(!g) { fclose(f); return NULL; }
  port_grid_set_max_truncation(g, maxTruncation);
  const int pitch = g->dim + 1;
  for(int y = 0; y < g->parts_y; y++)
    for(int x = 0; x < g->parts_x; x++)
    {
      const int id = get_int_line(f);
      part_t* p = &g->parts[y * g->parts_x + x];
      if(id == 0) continue;
      else if(id == 1)
      {
        p->init_weight = get_double_line(f);
        p->init_weight = (MAXWEIGHT < p->init_weight) ? MAXWEIGHT : p->init_weight; /* std::min(a, b): b < a ? b : a */
      }
      else if(id == 2)
      {
        part_init(g, p, g->max_truncation);
        for(int py = 0; py < g->dim; py++)
          for(int px = 0; px < g->dim; px++)
          {
            p->grid[py * pitch + px].tsd = get_double_line(f);
            p->grid[py * pitch + px].weight = get_double_line(f);
          }
      }
      else { fclose(f); port_grid_destroy(g); return NULL; }
    }
  fclose(f);
  return g;
}
