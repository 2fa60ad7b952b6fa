// RayCastPolar2D on the device (K5): one warp per beam.
//
// Reference: src/obvision/reconstruct/grid/RayCastPolar2D.cpp:113-192 (calcCoordsFromCurrentViewMask),
// :27-111 (calcCoordsFromCurrentView), :194-281 (rayCastFromCurrentView).
//
// The reference marches a beam serially: `position += ray` once per step and `i += 1.0` for the loop bound
// (`i += 32.0` in the coarse partition-skipping loop).  Those running sums are NOT tr + k*ray in floating
// point, so every lane of the warp replays the same serial additions (uniform work, 3 DADD per step) and
// keeps the value of its own step; what is distributed over the lanes is the expensive part, the bilinear
// sample (4 dependent-latency loads out of L2 + ~40 FP64 ops).  The first +/- sign change (hit) or -/+
// sign change (abort) of a 32-step chunk is found with two ballots.
#include <string.h>

#include "common.cuh"

using namespace tsd;

struct RayParams
{
  GridView g;
  ScanDev scan;
  const double* rays;  // 2 x n, world frame, length = cellSize
  double* out;         // 4 x n : cx cy nx ny (sensor frame)
  unsigned long long* keys;   // per beam: 4*step + code of the first event (code 0: hit, 1: hit whose normal
                              // failed, 2: abort), ~0 = none.  A sharded grid min-reduces these over the bands.
  unsigned long long* steps;  // [0] fine [1] coarse
  double xmin, ymin, xmax, ymax;  // RayCastPolar2D.cpp:128-146
  double idxMin, idxMax;          // :148-149
  int band;                       // 1: sharded grid, emit band-local first events (tsdg_raycast_band_keys)
};

#define RC_WARPS 4
#define NO_EVENT 0x7fffffffffffffffULL  // INT64_MAX: the largest key under a signed or unsigned min-reduction

__global__ void __launch_bounds__(RC_WARPS * 32) k_raycast(RayParams rp)
{
  const int lane = threadIdx.x & 31;
  const int beam = blockIdx.x * RC_WARPS + (threadIdx.x >> 5);
  if(beam >= rp.scan.n) return;
  const GridView& g = rp.g;
  const double ray0 = rp.rays[beam];
  const double ray1 = rp.rays[rp.scan.n + beam];
  const double tr0 = rp.scan.P[2], tr1 = rp.scan.P[5];
  const int xDim = g.cells_x, yDim = g.cells_y;
  const double cellSize = g.cell_size;

  unsigned long long key = NO_EVENT;
  unsigned long long nFine = 0, nCoarse = 0;
  bool found = false;
  double cx = 0, cy = 0, nx = 0, ny = 0;

  // RayCastPolar2D.cpp:205-221
  double xmin = rp.xmin, ymin = rp.ymin;
  if(fabs(ray0) > 10e-6) xmin = ((double)(ray0 > 0.0 ? 0 : (xDim - 1) * cellSize) - tr0) / ray0;
  if(fabs(ray1) > 10e-6) ymin = ((double)(ray1 > 0.0 ? 0 : (yDim - 1) * cellSize) - tr1) / ray1;
  double idxMin = ob_max(xmin, ymin);
  idxMin = ob_max(idxMin, 0.0);
  double xmax = rp.xmax, ymax = rp.ymax;
  if(fabs(ray0) > 10e-6) xmax = ((double)(ray0 > 0.0 ? (xDim - 1) * cellSize : 0) - tr0) / ray0;
  if(fabs(ray1) > 10e-6) ymax = ((double)(ray1 > 0.0 ? (yDim - 1) * cellSize : 0) - tr1) / ray1;
  double idxMax = ob_min(xmax, ymax);
  idxMin = ob_max(idxMin, rp.idxMin);
  idxMax = ob_min(idxMax, rp.idxMax);

  if(!(idxMin >= idxMax))
  {
    // :223-235 coarse loop, 32 iterations per pass
    {
      double i = idxMin;
      bool done = false;
      while(!done)
      {
        double mine = 0.0, ii = i;
#pragma unroll 8
        for(int k = 0; k < 32; k++)
        {
          if(k == lane) mine = ii;
          ii += 32.0;
        }
        const bool valid = mine < idxMax;
        bool stop = false;
        if(valid)
        {
          double tmp;
          const int rv = sample_bilinear(g, tr0 + mine * ray0, tr1 + mine * ray1, &tmp);
          stop = (rv != TSD_INTERPOLATE_EMPTYPARTITION && rv != TSD_INTERPOLATE_INVALIDINDEX);
        }
        const unsigned mStop = __ballot_sync(0xffffffffu, stop);
        const unsigned mInval = __ballot_sync(0xffffffffu, !valid);
        const int fStop = mStop ? (__ffs(mStop) - 1) : 32;
        const int fInval = mInval ? (__ffs(mInval) - 1) : 32;
        const int last = min(fStop, fInval);  // iterations [0, last) failed and moved idxMin
        if(last > 0) idxMin = __shfl_sync(0xffffffffu, mine, last - 1);
        nCoarse += (unsigned)min(fStop + 1, fInval);
        if(last < 32) done = true;
        i = ii;
      }
    }

    // :237-241
    double pos0 = tr0 + idxMin * ray0;
    double pos1 = tr1 + idxMin * ray1;
    double carry;
    {
      double v;
      carry = (sample_bilinear(g, pos0, pos1, &v) == TSD_INTERPOLATE_SUCCESS) ? v : __longlong_as_double(0x7ff8000000000000LL);
    }

    // :243-270 fine loop, 32 steps per pass.
    // The position chain of a pass is computed by all lanes (uniform DADDs), every lane keeps the position of its
    // own step; the chain of pass c+1 is issued between the loads and the use of pass c's samples, so the L2
    // latency of the samples hides behind it.
    // The loop counter `i += 1.0` of the reference only decides when the loop ends: it is replayed exactly
    // (serially) only for passes that come within 2 steps of idxMax; elsewhere idxMin + k decides safely.
    // (every lane runs the whole chain and keeps the position of its own step in registers: parking the positions
    //  in shared memory made each addition wait for the previous store to read its operands)
    double nmx = 0.0, nmy = 0.0;
    double iExact = idxMin;          // i of iteration kExact (exact serial value)
    unsigned long long kExact = 0;
    unsigned long long base = 0;
#pragma unroll 8
    for(int k = 0; k < 32; k++)
    {
      pos0 += ray0;
      pos1 += ray1;
      if(k == lane) { nmx = pos0; nmy = pos1; }
    }
    while(true)
    {
      const double mx = nmx, my = nmy;
      // validity of this lane's iteration: i_k <= idxMax
      const double est = idxMin + (double)(base + (unsigned)lane);
      bool valid;
      const bool nearEnd = !(idxMin + (double)(base + 31u) + 2.0 < idxMax);
      if(!nearEnd) valid = true;
      else
      {
        double mi = 0.0;
        // replay the reference's counter up to this pass (uniform), keep this lane's value
        while(kExact < base) { iExact += 1.0; kExact++; }
        double ii = iExact;
#pragma unroll 8
        for(int k = 0; k < 32; k++)
        {
          if(k == lane) mi = ii;
          ii += 1.0;
        }
        valid = mi <= idxMax;
        (void)est;
      }
      double v = __longlong_as_double(0x7ff8000000000000LL);
      double t = 0.0;
      const SampleLoads sl = sample_issue(g, mx, my);  // all lanes: addresses are clamped, results masked below
      // next pass's positions while the loads above are in flight
      {

#pragma unroll 8
        for(int k = 0; k < 32; k++)
        {
          pos0 += ray0;
          pos1 += ray1;
          if(k == lane) { nmx = pos0; nmy = pos1; }
        }
      }
      const int rv = sample_finish(sl, &t);
      if(valid && rv == TSD_INTERPOLATE_SUCCESS) v = t;
      double prev = __shfl_up_sync(0xffffffffu, v, 1);
      if(lane == 0) prev = carry;
      // a step belongs to the band that owns its sample's partition (everything, for an unsharded grid)
      const bool mine = valid && (sl.py >= g.row_begin) && (sl.py < g.row_end);
      const bool hit = mine && (prev > 0) && (v < 0);
      const bool abortEv = mine && (prev < 0) && (v > 0);
      const unsigned mHit = __ballot_sync(0xffffffffu, hit);
      const unsigned mEv = mHit | __ballot_sync(0xffffffffu, abortEv);
      const unsigned mInval = __ballot_sync(0xffffffffu, !valid);
      if(mEv)
      {
        const int f = __ffs(mEv) - 1;
        nFine += (unsigned)(f + 1);
        const bool isHit = (mHit >> f) & 1u;
        key = 4ULL * (base + (unsigned)f) + (isHit ? 1ULL : 2ULL);
        if(isHit)
        {
          // :259, :277-280 on the lane that owns the step
          int ok = 0;
          if(lane == f)
          {
            const double interp = prev / (prev - v);
            cx = mx + ray0 * (interp - 1.0);
            cy = my + ray1 * (interp - 1.0);
            ok = sample_normal(g, cx, cy, &nx, &ny) ? 1 : 0;
          }
          ok = __shfl_sync(0xffffffffu, ok, f);
          cx = __shfl_sync(0xffffffffu, cx, f);
          cy = __shfl_sync(0xffffffffu, cy, f);
          nx = __shfl_sync(0xffffffffu, nx, f);
          ny = __shfl_sync(0xffffffffu, ny, f);
          found = ok != 0;
          if(found) key &= ~3ULL;
        }
        break;
      }
      if(mInval)
      {
        nFine += (unsigned)(__ffs(mInval) - 1);
        break;
      }
      nFine += 32;
      base += 32;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
  }

  if(lane == 0)
  {
    if(found)
    {
      // :168-177  M = T * [c;1], N = T * [n;0]  (sensor frame)
      double m0, m1, n0, n1;
      mat3_vec_nn(rp.scan.Pi, cx, cy, 1.0, &m0, &m1);
      mat3_vec_nn(rp.scan.Pi, nx, ny, 0.0, &n0, &n1);
      rp.out[4 * beam + 0] = m0;
      rp.out[4 * beam + 1] = m1;
      rp.out[4 * beam + 2] = n0;
      rp.out[4 * beam + 3] = n1;
    }
    else
    {
      rp.out[4 * beam + 0] = 0.0;
      rp.out[4 * beam + 1] = 0.0;
      rp.out[4 * beam + 2] = 0.0;
      rp.out[4 * beam + 3] = 0.0;
    }
    rp.keys[beam] = key;
    atomicAdd(&rp.steps[0], nFine);
    atomicAdd(&rp.steps[1], nCoarse);
  }
}

#define RC_IS_HIT(k) ((k) != NO_EVENT && ((k) & 3ULL) == 0ULL)

static int raycast_launch(tsd_grid_t* g, const tsd_scan_t* scan, const double* rays_world, bool download = true)
{
  TSD_LOCK(g);
  if(!g || !scan || !rays_world) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  RayParams rp;
  memset(&rp, 0, sizeof(rp));
  int rc = grid_stage_scan(g, scan, &rp.scan, rays_world);
  if(rc) return rc;
  const int n = scan->n;
  g->rc_steps_prev[0] = g->h_rc_steps[0];  // the device counters run on; a call's steps are the difference
  g->rc_steps_prev[1] = g->h_rc_steps[1];
  rp.g = grid_view(g);
  rp.rays = g->d_rays;
  rp.out = g->d_rc_out;
  rp.keys = g->d_rc_keys;
  rp.steps = g->d_rc_steps;
  const double trx = scan->pose[2], try_ = scan->pose[5];
  // TsdGrid::isInsideGrid (TsdGrid.h:342-347), RayCastPolar2D.cpp:128-146
  if(trx > g->min_x && trx < g->max_x && try_ > g->min_y && try_ < g->max_y)
  {
    rp.xmin = -10e9; rp.ymin = -10e9; rp.xmax = 10e9; rp.ymax = 10e9;
  }
  else
  {
    rp.xmin = 10e9; rp.ymin = 10e9; rp.xmax = -10e9; rp.ymax = -10e9;
  }
  rp.idxMin = scan->min_range / g->cell_size;
  rp.idxMax = scan->max_range / g->cell_size;
  k_raycast<<<(n + RC_WARPS - 1) / RC_WARPS, RC_WARPS * 32, 0, g->stream>>>(rp);
  TSD_LAUNCHED();
  if(download) TSD_CUDA(cudaMemcpyAsync(g->h_rc, g->d_rc, g->rc_bytes, cudaMemcpyDeviceToHost, g->stream));
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  return TSD_OK;
}

extern "C" {

int tsdg_raycast_mask(tsd_grid_t* g, const tsd_scan_t* scan, const double* rays_world, double* coords,
                      double* normals, uint8_t* mask, uint32_t* count)
{
  TSD_LOCK(g);
  if(!coords || !normals || !mask) return TSD_E_INVALID;
  int rc = raycast_launch(g, scan, rays_world);
  if(rc) return rc;
  uint32_t cnt = 0;
  for(int b = 0; b < scan->n; b++)
  {
    if(RC_IS_HIT(g->h_rc_keys[b]))
    {
      coords[2 * b] = g->h_rc_out[4 * b];
      coords[2 * b + 1] = g->h_rc_out[4 * b + 1];
      normals[2 * b] = g->h_rc_out[4 * b + 2];
      normals[2 * b + 1] = g->h_rc_out[4 * b + 3];
      mask[b] = 1;
      cnt++;
    }
    else
      mask[b] = 0;  // coords/normals of a miss are left untouched, as RayCastPolar2D.cpp:181-184 does
  }
  if(count) *count = cnt;
  return TSD_OK;
}

int tsdg_raycast(tsd_grid_t* g, const tsd_scan_t* scan, const double* rays_world, double* coords, double* normals,
                 uint32_t* count)
{
  TSD_LOCK(g);
  if(!coords || !normals || !count) return TSD_E_INVALID;
  int rc = raycast_launch(g, scan, rays_world);
  if(rc) return rc;
  // beam order = the reference's single-thread order
  uint32_t cnt = 0;
  for(int b = 0; b < scan->n; b++)
  {
    if(RC_IS_HIT(g->h_rc_keys[b]))
    {
      coords[cnt] = g->h_rc_out[4 * b];
      normals[cnt++] = g->h_rc_out[4 * b + 2];
      coords[cnt] = g->h_rc_out[4 * b + 1];
      normals[cnt++] = g->h_rc_out[4 * b + 3];
    }
  }
  *count = cnt;
  return TSD_OK;
}

int tsdg_last_raycast_steps(tsd_grid_t* g, uint64_t* fine_steps, uint64_t* coarse_steps)
{
  TSD_LOCK(g);
  if(!g) return TSD_E_INVALID;
  if(!g->h_rc_steps) return TSD_E_INVALID;
  if(fine_steps) *fine_steps = g->h_rc_steps[0] - g->rc_steps_prev[0];
  if(coarse_steps) *coarse_steps = g->h_rc_steps[1] - g->rc_steps_prev[1];
  return TSD_OK;
}

int tsdg_raycast_band_keys(tsd_grid_t* g, const tsd_scan_t* scan, const double* rays_world, uint64_t** dev_keys,
                           double** dev_payload)
{
  TSD_LOCK(g);
  if(!dev_keys || !dev_payload) return TSD_E_INVALID;
  int rc = raycast_launch(g, scan, rays_world, false);
  if(rc) return rc;
  *dev_keys = reinterpret_cast<uint64_t*>(g->d_rc_keys);
  *dev_payload = g->d_rc_out;
  return TSD_OK;
}

}  // extern "C"
