// Map publication on the device: what ThreadGrid (src/ThreadGrid.cpp:84,125) asks of the
// map every two seconds -- the zero crossings + occupancy grid of RayCastAxisAligned2D::calcCoords
// (src/obvision/reconstruct/grid/RayCastAxisAligned2D.cpp:13-105) and TsdGrid::grid2ColorImage
// (TsdGrid.cpp:429-488) -- without copying the cell state to the host.
//
// calcCoords walks the inner partitions in row-major order and, inside an allocated one, the 33 cell rows
// (borders included) and then the 33 cell columns; the ORDER of the emitted points is part of its result.  Here:
//   k_axis_count   one warp per inner partition: the 33x33 cells go to shared memory, crossings are counted
//   cub scan       exclusive prefix over the partitions in the reference's order
//   k_axis_emit    the same warp pass again, writing every crossing at its rank (ballot + popc inside a row)
//   k_occupancy    one thread per grid cell: the value the LAST writer of the reference's sequential loop leaves
//                  (a partition also writes its border column / row into the first cells of its +x / +y / +xy
//                  neighbours, and later partitions overwrite earlier ones)
//   k_color_image  one thread per pixel; the pixel coordinates are the reference's running sums (host tables)
// Unsharded grids only (the publisher runs next to the mapper).
#include <string.h>

#include <vector>

#include <cub/device/device_scan.cuh>

#include "common.cuh"

using namespace tsd;

namespace
{

struct AxisParams
{
  GridView g;
  const double* tsd;
  int inner;         // partitions per axis that calcCoords visits: parts - 2
  double cell_size;
  unsigned* counts;  // per inner partition: crossings (rows + columns)
  unsigned* offsets; // exclusive prefix of counts
  double* coords;    // 2 doubles per crossing
  unsigned cap;      // capacity of coords in points
};

// (*p)(py, px) of TsdGridPartition.h:83 for py, px in 0..32: the 33x33 view with the replicated border
__device__ __forceinline__ int cell33(int py, int px)
{
  if(px < 32) return (py < 32) ? py * 32 + px : TSD_BORDER_OFF + 32 + px;
  return (py < 32) ? TSD_BORDER_OFF + py : TSD_BORDER_OFF + 64;
}

#define AXIS_WARPS 4

// mode 0: count, mode 1: emit
template <int MODE>
__global__ void __launch_bounds__(AXIS_WARPS * 32) k_axis(AxisParams A)
{
  __shared__ double s_tile[AXIS_WARPS][33 * 33 + 3];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int pi = blockIdx.x * AXIS_WARPS + w;  // inner partition, in the reference's order
  if(pi >= A.inner * A.inner) return;
  const int x = 1 + pi % A.inner, y = 1 + pi / A.inner;
  const int p = y * A.g.parts_x + x;
  if(!A.g.flags[p])
  {
    if(MODE == 0 && lane == 0) A.counts[pi] = 0u;
    return;
  }
  const double* t = A.tsd + (size_t)p * TSD_TILE_STRIDE;
  double* tile = s_tile[w];
  for(int i = lane; i < 33 * 33; i += 32) tile[i] = __ldg(t + cell33(i / 33, i % 33));
  __syncwarp();
  unsigned n = 0;
  const unsigned base = (MODE == 1) ? A.offsets[pi] : 0u;
  const double cs = A.cell_size;
  // rows: lane l looks at px = l + 1 against px = l (RayCastAxisAligned2D.cpp:36-59)
  for(int py = 0; py < 33; py++)
  {
    const double prev = tile[py * 33 + lane], v = tile[py * 33 + lane + 1];
    const bool hit = (prev > 0 && v < 0) || (prev < 0 && v > 0);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if(MODE == 1 && hit)
    {
      const unsigned k = base + n + __popc(m & ((1u << lane) - 1u));
      if(k < A.cap)
      {
        const unsigned px = (unsigned)lane + 1u;
        const double interp = prev / (prev - v);
        A.coords[2 * (size_t)k] = px * cs + cs * (interp - 1.0) + (double)((unsigned)x * 32u) * cs;
        A.coords[2 * (size_t)k + 1] = (unsigned)py * cs + (double)((unsigned)y * 32u) * cs;
      }
    }
    n += __popc(m);
  }
  // columns: lane l looks at py = l + 1 against py = l (:61-80)
  for(int px = 0; px < 33; px++)
  {
    const double prev = tile[lane * 33 + px], v = tile[(lane + 1) * 33 + px];
    const bool hit = (prev > 0 && v < 0) || (prev < 0 && v > 0);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if(MODE == 1 && hit)
    {
      const unsigned k = base + n + __popc(m & ((1u << lane) - 1u));
      if(k < A.cap)
      {
        const unsigned py = (unsigned)lane + 1u;
        const double interp = prev / (prev - v);
        A.coords[2 * (size_t)k] = (unsigned)px * cs + (double)((unsigned)x * 32u) * cs;
        A.coords[2 * (size_t)k + 1] = py * cs + cs * (interp - 1.0) + (double)((unsigned)y * 32u) * cs;
      }
    }
    n += __popc(m);
  }
  if(MODE == 0 && lane == 0) A.counts[pi] = n;
}

// What calcCoords leaves in occupiedGrid[gy * cellsX + gx].  Writers of a cell, in the order of the reference's
// loops: the partitions (X-1,Y-1), (X,Y-1), (X-1,Y) through their border corner / row / column, then the cell's
// own partition (X,Y).  Only inner partitions write; an allocated one writes (tsd > 0 ? 0 : -1) for its 33x33
// cells, an unallocated one that was seen empty writes 0 for its own 32x32 cells.  The last writer wins; a cell
// nobody writes keeps the caller's value.
__global__ void k_occupancy(GridView g, const double* tsd, const double* initw, int8_t* occ)
{
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= (size_t)g.cells_x * g.cells_y) return;
  const int gx = (int)(idx % g.cells_x), gy = (int)(idx / g.cells_x);
  const int X = gx >> 5, Y = gy >> 5, lx = gx & 31, ly = gy & 31;
  const int P = g.parts_x;
  auto inner = [&](int px, int py) { return px >= 1 && px <= P - 2 && py >= 1 && py <= P - 2; };
  auto value = [&](int px, int py, int cy, int cx) -> int8_t
  {
    const double v = __ldg(tsd + (size_t)(py * P + px) * TSD_TILE_STRIDE + cell33(cy, cx));
    return (v > 0.0) ? 0 : -1;
  };
  if(inner(X, Y))
  {
    const int p = Y * P + X;
    if(g.flags[p]) { occ[idx] = value(X, Y, ly, lx); return; }
    if(initw[p] > 0.0) { occ[idx] = 0; return; }
  }
  if(lx == 0 && inner(X - 1, Y) && g.flags[Y * P + X - 1]) { occ[idx] = value(X - 1, Y, ly, 32); return; }
  if(ly == 0 && inner(X, Y - 1) && g.flags[(Y - 1) * P + X]) { occ[idx] = value(X, Y - 1, 32, lx); return; }
  if(lx == 0 && ly == 0 && inner(X - 1, Y - 1) && g.flags[(Y - 1) * P + X - 1]) { occ[idx] = value(X - 1, Y - 1, 32, 32); return; }
}

// TsdGrid::grid2ColorImage (TsdGrid.cpp:429-488); xs / ys: the running sums px += stepW, py += stepH
__global__ void k_color_image(GridView g, const double* tsd, const double* initw, const double* xs, const double* ys,
                              unsigned width, unsigned height, uint8_t* image)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= (size_t)width * height) return;
  const unsigned w = (unsigned)(i % width), h = (unsigned)(i / width);
  const double cx = xs[w], cy = ys[h];
  // coord2Cell (TsdGrid.h:306-340)
  int xIdx = __double2int_rd(cx * g.inv_cell_size);
  int yIdx = __double2int_rd(cy * g.inv_cell_size);
  const double dx = ((double)xIdx + 0.5) * g.cell_size;
  const double dy = ((double)yIdx + 0.5) * g.cell_size;
  if(cx < dx) xIdx--;
  if(cy < dy) yIdx--;
  double v = __longlong_as_double(0x7ff8000000000000LL);
  bool isEmpty = false;
  if(!((xIdx >= g.cells_x) || (xIdx < 0) || (yIdx >= g.cells_y) || (yIdx < 0)))
  {
    const int p = (yIdx >> 5) * g.parts_x + (xIdx >> 5);
    if(g.flags[p]) v = __ldg(tsd + (size_t)p * TSD_TILE_STRIDE + (yIdx & 31) * 32 + (xIdx & 31));
    else isEmpty = initw[p] > 0.0;
  }
  uint8_t r, gg, b;
  if(v > 0.0)
  {
    r = (uint8_t)(v * 255.0);
    gg = 255;
    b = (uint8_t)(v * 255.0);
  }
  else if(v < 0.0)
  {
    r = (uint8_t)((1.0 + v) * 255.0);
    gg = 0;
    b = 0;
  }
  else if(isEmpty) { r = gg = b = 255; }
  else { r = gg = b = 0; }
  image[3 * i] = r;
  image[3 * i + 1] = gg;
  image[3 * i + 2] = b;
}

// TsdGrid::interpolateNormal (TsdGrid.cpp:517-546) at one point, with the reference's partial writes: the x
// component is stored before the y lookups are tried.  out = {nx, ny, stage}: stage 0 nothing written, 1 only nx
// (raw difference), 2 both (normalised unless the length is <= 1e-5).
__global__ void k_normal_partial(GridView g, const double* xy, double* out)
{
  const double cx = xy[0], cy = xy[1];
  double inc = 0, dec = 0;
  out[2] = 0.0;
  if(sample_bilinear(g, cx + g.cell_size, cy, &inc) != TSD_INTERPOLATE_SUCCESS) return;
  if(sample_bilinear(g, cx - g.cell_size, cy, &dec) != TSD_INTERPOLATE_SUCCESS) return;
  double nx = inc - dec;
  out[0] = nx;
  out[2] = 1.0;
  if(sample_bilinear(g, cx, cy + g.cell_size, &inc) != TSD_INTERPOLATE_SUCCESS) return;
  if(sample_bilinear(g, cx, cy - g.cell_size, &dec) != TSD_INTERPOLATE_SUCCESS) return;
  double ny = inc - dec;
  const double len = sqrt(nx * nx + ny * ny);
  if(!(fabs(len) <= 10e-6))
  {
    nx /= len;
    ny /= len;
  }
  out[0] = nx;
  out[1] = ny;
  out[2] = 2.0;
}

}  // namespace

extern "C" {

int tsdg_axis_aligned_map(tsd_grid_t* g, double* coords, uint32_t cap_points, double* normals, uint32_t* count,
                          int8_t* occupied)
{
  TSD_LOCK(g);
  if(!g || !count || (cap_points > 0 && !coords)) return TSD_E_INVALID;
  if(g->band) { set_error("map publication works on an unsharded grid"); return TSD_E_INVALID; }
  TSD_CUDA(cudaSetDevice(g->device));
  *count = 0;
  const int inner = g->parts_x - 2;
  const size_t cells = (size_t)g->cells_x * g->cells_y;
  GridView gv = grid_view(g);
  unsigned total = 0;
  if(inner > 0)
  {
    const int n = inner * inner;
    // scratch: counts | offsets | scan temp | coords
    size_t tempBytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tempBytes, (unsigned*)nullptr, (unsigned*)nullptr, n, g->stream);
    const size_t offCounts = 0, offOffsets = sizeof(unsigned) * (size_t)n, offTemp = ((2 * sizeof(unsigned) * (size_t)n + 255) / 256) * 256;
    const size_t offCoords = ((offTemp + tempBytes + 255) / 256) * 256;
    int rc = grid_ensure_scratch(g, offCoords + sizeof(double) * 2 * (size_t)cap_points);
    if(rc) return rc;
    unsigned char* d = (unsigned char*)g->d_scratch;
    AxisParams A;
    A.g = gv;
    A.tsd = g->d_tsd;
    A.inner = inner;
    A.cell_size = g->cell_size;
    A.counts = (unsigned*)(d + offCounts);
    A.offsets = (unsigned*)(d + offOffsets);
    A.coords = (double*)(d + offCoords);
    A.cap = cap_points;
    const int ctas = (n + AXIS_WARPS - 1) / AXIS_WARPS;
    k_axis<0><<<ctas, AXIS_WARPS * 32, 0, g->stream>>>(A);
    TSD_LAUNCHED();
    TSD_CUDA(cub::DeviceScan::ExclusiveSum(d + offTemp, tempBytes, A.counts, A.offsets, n, g->stream));
    k_axis<1><<<ctas, AXIS_WARPS * 32, 0, g->stream>>>(A);
    TSD_LAUNCHED();
    unsigned last[2];
    TSD_CUDA(cudaMemcpyAsync(&last[0], A.counts + (n - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, g->stream));
    TSD_CUDA(cudaMemcpyAsync(&last[1], A.offsets + (n - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, g->stream));
    TSD_CUDA(cudaStreamSynchronize(g->stream));
    total = last[0] + last[1];
    const unsigned got = total < cap_points ? total : cap_points;
    if(got) TSD_CUDA(cudaMemcpy(coords, A.coords, sizeof(double) * 2 * (size_t)got, cudaMemcpyDeviceToHost));
    if(total > cap_points)
    {
      set_error("tsdg_axis_aligned_map: %u crossings, room for %u", total, cap_points);
      *count = 2 * got;
      return TSD_E_RANGE;
    }
  }
  *count = 2 * total;  // the reference counts doubles
  if(normals && total > 0)
  {
    // RayCastAxisAligned2D.cpp:52,73 pass the array BASE to interpolateNormal: every normal is the one at the
    // first crossing; a lookup that fails half-way leaves what it wrote before (k_normal_partial)
    int rc = grid_ensure_scratch(g, sizeof(double) * 8);
    if(rc) return rc;
    double* d = (double*)g->d_scratch;
    double n0[3] = {0, 0, 0};
    TSD_CUDA(cudaMemcpyAsync(d, coords, sizeof(double) * 2, cudaMemcpyHostToDevice, g->stream));
    k_normal_partial<<<1, 1, 0, g->stream>>>(gv, d, d + 2);
    TSD_LAUNCHED();
    TSD_CUDA(cudaMemcpyAsync(n0, d + 2, sizeof(double) * 3, cudaMemcpyDeviceToHost, g->stream));
    TSD_CUDA(cudaStreamSynchronize(g->stream));
    if(n0[2] >= 1.0)
      for(unsigned k = 0; k < total; k++)
      {
        normals[2 * k] = n0[0];
        if(n0[2] >= 2.0) normals[2 * k + 1] = n0[1];
      }
  }
  if(occupied)
  {
    int rc = grid_ensure_scratch(g, cells);
    if(rc) return rc;
    int8_t* d_occ = (int8_t*)g->d_scratch;
    TSD_CUDA(cudaMemcpyAsync(d_occ, occupied, cells, cudaMemcpyHostToDevice, g->stream));  // cells nobody writes keep their value
    k_occupancy<<<(unsigned)((cells + 255) / 256), 256, 0, g->stream>>>(gv, g->d_tsd, g->d_initw, d_occ);
    TSD_LAUNCHED();
    TSD_CUDA(cudaMemcpyAsync(occupied, d_occ, cells, cudaMemcpyDeviceToHost, g->stream));
    TSD_CUDA(cudaStreamSynchronize(g->stream));
  }
  return TSD_OK;
}

int tsdg_color_image(tsd_grid_t* g, uint8_t* image, uint32_t width, uint32_t height)
{
  TSD_LOCK(g);
  if(!g || !image || width == 0 || height == 0) return TSD_E_INVALID;
  if(g->band) { set_error("map publication works on an unsharded grid"); return TSD_E_INVALID; }
  TSD_CUDA(cudaSetDevice(g->device));
  const size_t pixels = (size_t)width * height;
  const size_t tabBytes = sizeof(double) * ((size_t)width + height);
  int rc = grid_ensure_scratch(g, tabBytes + 3 * pixels);
  if(rc) return rc;
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  // TsdGrid.cpp:433-440,484-486: running sums, not w * stepW
  double* tab = (double*)g->h_scratch;
  const double stepW = g->max_x / (double)width, stepH = g->max_y / (double)height;
  double acc = 0.0;
  for(uint32_t w = 0; w < width; w++) { tab[w] = acc; acc += stepW; }
  acc = 0.0;
  for(uint32_t h = 0; h < height; h++) { tab[width + h] = acc; acc += stepH; }
  double* d_tab = (double*)g->d_scratch;
  uint8_t* d_img = (uint8_t*)g->d_scratch + tabBytes;
  TSD_CUDA(cudaMemcpyAsync(d_tab, tab, tabBytes, cudaMemcpyHostToDevice, g->stream));
  k_color_image<<<(unsigned)((pixels + 255) / 256), 256, 0, g->stream>>>(grid_view(g), g->d_tsd, g->d_initw, d_tab, d_tab + width,
                                                                            width, height, d_img);
  TSD_LAUNCHED();
  TSD_CUDA(cudaMemcpyAsync(image, d_img, 3 * pixels, cudaMemcpyDeviceToHost, g->stream));
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  return TSD_OK;
}

}  // extern "C"
