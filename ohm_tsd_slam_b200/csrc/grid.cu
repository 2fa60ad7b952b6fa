// TsdGrid on the device: dense SoA cell state in HBM, scan integration (push), border refresh,
// bilinear sampling and partition accessors.
//
// Reference: src/obvision/reconstruct/grid/TsdGrid.{h,cpp}, TsdGridPartition.{h,cpp},
// TsdGridComponent.cpp:43-124 (isInRange), SensorPolar2D.cpp:117-135 (backProject).
//
// Data layout (DESIGN.md "HBM layout"): two arrays `tsd` and `weight`, one block of TSD_TILE_STRIDE
// doubles per partition: 32x32 interior cells row-major (8192 B, 128-B aligned) followed by the 65
// replicated border cells.  One byte `flags` (initialised) and one double `initw` (TsdGridPartition::
// _initWeight) per partition.  The homogeneous cell-coordinate matrices the reference stores per partition
// (24.6 KB each) are not stored: they are recomputed from the cell index.
//
// One push (of one scan, or of two scans of the same sensor model: tsdg_push_batch) = 2 launches on the handle's stream:
//   k_classify  4 lanes per partition of the scan's range box: TsdGridComponent::isInRange incl. the emptiness
//               side effect, one work-list entry per touched partition (K1); also the per-column / per-row tables
//   k_update    persistent CTAs over the work list: addTsd per cell / increaseEmptiness, border strips mirrored
//               into the neighbours, and -- in the last CTA to finish -- the borders of newly allocated partitions
//               and the per-push bookkeeping (K2, K3, most of K4)
// plus k_borders after a fill / upload (full border refresh) and k_halo_sync / k_borders on sharded grids.
#include <stdarg.h>
#include <string.h>
#include <cmath>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace tsd
{

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_error = buf;
}

int fill_scan_dev(const tsd_scan_t* s, ScanDev* o)
{
  o->n = s->n;
  for(int i = 0; i < 9; i++) { o->P[i] = s->pose[i]; o->Pi[i] = s->pose_inv[i]; }
  o->max_range = s->max_range;
  o->min_range = s->min_range;
  o->low_refl = s->low_reflectivity_range;
  BeamModel& bm = o->bm;
  bm.phi_min = s->phi_min;
  bm.res_inv = 1.0 / s->angular_res;  // SensorPolar2D.cpp:127
  bm.phi_lower = s->phi_lower;
  bm.phi_upper = s->phi_upper;
  bm.phi_min_f = (float)s->phi_min;
  bm.res_inv_f = (float)bm.res_inv;
  bm.phi_lower_f = (float)s->phi_lower;
  bm.phi_upper_f = (float)s->phi_upper;
  bm.n = s->n;
  const double pi = 3.14159265358979323846;
  bm.fast_ok = (s->angular_res > 1e-9 && s->phi_lower >= -pi - 1e-9 && s->phi_upper <= pi + 1e-9 &&
                s->phi_upper > s->phi_lower && s->n >= 1)
                   ? 1
                   : 0;
  FastModel fm;
  tsd_fast_model(o->Pi, s->phi_min, s->angular_res, s->phi_lower, s->phi_upper, s->n, &fm);  // beam_index.cuh
  o->txp = fm.txp;
  o->typ = fm.typ;
  o->rinv_f = fm.rinv_f;
  o->off_f = fm.off_f;
  o->half_m = fm.half_m;
  return TSD_OK;
}

GridView grid_view(const tsd_grid* g)
{
  GridView v;
  v.tsd = g->d_tsd;
  v.flags = g->d_flags;
  v.cells_x = g->cells_x;
  v.cells_y = g->cells_y;
  v.parts_x = g->parts_x;
  v.parts_y = g->parts_y;
  v.row_begin = g->row_begin;
  v.row_end = g->row_end;
  v.alloc_begin = g->alloc_begin;
  v.alloc_end = g->alloc_end;
  v.cell_size = g->cell_size;
  v.inv_cell_size = g->inv_cell_size;
  return v;
}

}  // namespace tsd

using namespace tsd;

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

#define PUSH_MAX_SCANS 4  // scans one push launch integrates (tsdg_push_batch); launches are built for 1, 2 and 4

struct PushParams
{
  ScanDev scans[PUSH_MAX_SCANS];
  int nscan;
  double cell_size;
  double max_trunc;
  double inv_max_trunc;  // 1.0 / maxTruncation (TsdGridPartition.cpp:94)
  int cells_x, cells_y, parts_x, parts_y, n_parts;
  int parts_shift;  // parts_x == 1 << parts_shift
  int scan_cap;                     // padded size of a staged scan (ensure_scan_capacity)
  int cl_px0, cl_py0, cl_w, cl_h;  // partitions k_classify looks at: the scan's range box (scan_partition_box)
  int row_begin, row_end;
  int alloc_begin, alloc_end;
  int band;
  int fused_tail;  // 1: k_update's last CTA runs the pull pass and the push tail (no k_borders launch)
  double* tsd;
  double* weight;
  uint8_t* flags;
  double* initw;
  uint32_t* active;    // work list: partition | bit 31 = allocated before this push
  uint32_t* kinds;     // per list entry: 2 bits per scan (0 untouched, 1 increaseEmptiness, 2 active)
  double* active_w;    // per list entry and scan: partition weight (capacity n_owned per scan)
  int list_cap;        // n_owned
  uint32_t* newly;    // partitions allocated by this push (count: counters[4], owned ones only in the list)
  uint32_t* pending;
  uint32_t* counters;
  unsigned long long* stats64;
  const double* coltab;  // per scan: 3 * cells_x
  const double* rowtab;  // per scan: 3 * cells_y
  const double2* dirs;
  float2* colxy;         // per scan: cells_x   (single-precision tables of the fast path, fill_tables)
  float* cold;           // per scan: cells_x
  float2* rowxy;         // per scan: cells_y
  float* rowd;           // per scan: cells_y
  float2* gate;          // per scan: scan_cap  (fill_gate)
  ScanDev* scans_dev;    // copy of scans[] in global memory (written by k_classify) for out-of-line code
  unsigned update_filter;  // measurement aid (tsdg_set_update_filter): bit 0 = skip K2 work, bit 1 = skip K3 work
  unsigned long long* prof;  // UPDATE_PROFILE builds: 8 cycle counters per CTA of k_update
};

// Per grid column X = ((double)ix + 0.5) * cellSize (TsdGridPartition.cpp:127): the two products of
// SensorPolar2D.cpp:125 that depend on X only, with gslcblas' accumulation order (temp = 0; temp += a*b ...),
// and the squared offset of TsdGrid.cpp:262.  Same per grid row.  Runs inside k_classify (first threads).
__device__ __forceinline__ void fill_tables(const PushParams& pp, const ScanDev& sc, double* coltab, double* rowtab, float2* colxy,
                                            float* cold, float2* rowxy, float* rowd, int i0)
{
  const double* Pi = sc.Pi;
  if(i0 < pp.cl_w * TSD_TILE)
  {
    const int i = pp.cl_px0 * TSD_TILE + i0;
    const double X = ((double)i + 0.5) * pp.cell_size;
    double a = 0.0;
    a += Pi[0] * X;
    double b = 0.0;
    b += Pi[3] * X;
    const double d = X - sc.P[2];
    coltab[i] = a;
    coltab[pp.cells_x + i] = b;
    coltab[2 * pp.cells_x + i] = d * d;
    const double dp = X - sc.txp;
    colxy[i] = make_float2((float)(Pi[0] * dp), (float)(Pi[3] * dp));
    cold[i] = (float)(d * d);
  }
  if(i0 < pp.cl_h * TSD_TILE)
  {
    const int i = pp.cl_py0 * TSD_TILE + i0;
    const double Y = ((double)i + 0.5) * pp.cell_size;
    const double d = Y - sc.P[5];
    rowtab[i] = Pi[1] * Y;
    rowtab[pp.cells_y + i] = Pi[4] * Y;
    rowtab[2 * pp.cells_y + i] = d * d;
    const double dp = Y - sc.typ;
    rowxy[i] = make_float2((float)(Pi[1] * dp), (float)(Pi[4] * dp));
    rowd[i] = (float)(d * d);
  }
}

// Per beam: the squared-distance gates of the single-precision front end (tsd_gate_entry, beam_index.cuh).
__device__ __forceinline__ void fill_gate(const PushParams& pp, const ScanDev& sc, float2* gate, int i)
{
  if(i >= sc.n) return;
  float lo2, hi2;
  tsd_gate_entry(sc.ranges[i], sc.mask[i] != 0, pp.max_trunc, sc.low_refl, &lo2, &hi2);
  gate[i] = make_float2(lo2, hi2);
}

// SensorPolar2D::backProject for one homogeneous point, complete (sign of zero included), for the 4 edge
// points of a partition.
__device__ __forceinline__ int back_project_edge(const ScanDev& s, const double2* dirs, double X, double Y)
{
  const double* P = s.Pi;
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
  bool slow;
  const int k = beam_index(s.bm, dirs, cx, cy, &slow);
  if(k < 0 && !slow)
  {
    // the fast path decides "outside" but not which side when y' is a signed zero: take the exact route
    if(cy == 0.0) return beam_index_exact(s.bm, cx, cy);
  }
  return k;
}

#define CLASSIFY_THREADS 256
#define CLASSIFY_MAX_BEAMS 1152  // scans up to this (padded) size are copied to shared memory by every classifier CTA

// K1: TsdGridComponent::isInRange (TsdGridComponent.cpp:43-124) for every partition of the range box.  Four lanes
// per partition (one per edge point), eight partitions per warp; the beam interval [minIdx, maxIdx] is scanned by
// the four lanes.  The first threads also fill the column / row tables of this push.
// A push launch can integrate up to PUSH_MAX_SCANS scans (tsdg_push_batch): every partition is classified for scan
// 0, then for scan 1, with the state scan 0 left (allocation flag, emptiness weight) -- the order TsdGrid::push
// calls would produce -- and gets ONE work-list entry with a 2-bit outcome per scan.
template <int NS>
__global__ void __launch_bounds__(CLASSIFY_THREADS) k_classify(PushParams pp, double* coltab, double* rowtab)
{
  const int gtid = blockIdx.x * CLASSIFY_THREADS + threadIdx.x;
  if(gtid < NS) pp.scans_dev[gtid] = pp.scans[gtid];
#pragma unroll
  for(int si = 0; si < NS; si++)
  {
    fill_tables(pp, pp.scans[si], coltab + (size_t)si * 3 * pp.cells_x, rowtab + (size_t)si * 3 * pp.cells_y,
                pp.colxy + (size_t)si * pp.cells_x, pp.cold + (size_t)si * pp.cells_x, pp.rowxy + (size_t)si * pp.cells_y,
                pp.rowd + (size_t)si * pp.cells_y, gtid);
    fill_gate(pp, pp.scans[si], pp.gate + (size_t)si * pp.scan_cap, gtid);
  }
  // The beam loops below walk ranges / mask with dependent loads; from global memory about half of them missed L1
  // (a CTA touches the scan only briefly), so every CTA first copies the scans into shared memory (9.7 KB each).
  __shared__ __align__(16) double s_ranges[NS][CLASSIFY_MAX_BEAMS];
  __shared__ __align__(16) uint8_t s_mask[NS][CLASSIFY_MAX_BEAMS];
  const bool stagedScan = pp.scan_cap <= CLASSIFY_MAX_BEAMS;
  if(stagedScan)
  {
    // 16-byte copies of the padded staging slots (scan_cap is a multiple of 64 and the slots are 64-byte aligned)
#pragma unroll
    for(int si = 0; si < NS; si++)
    {
      const double2* r2 = reinterpret_cast<const double2*>(pp.scans[si].ranges);
      const uint4* m4 = reinterpret_cast<const uint4*>(pp.scans[si].mask);
      for(int i = threadIdx.x; i < pp.scan_cap / 2; i += CLASSIFY_THREADS) reinterpret_cast<double2*>(s_ranges[si])[i] = r2[i];
      for(int i = threadIdx.x; i < pp.scan_cap / 16; i += CLASSIFY_THREADS) reinterpret_cast<uint4*>(s_mask[si])[i] = m4[i];
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane & 3;
  const int gshift = lane & ~3;
  const unsigned gmask = 0xfu << gshift;
  const int q = gtid >> 2;  // index inside the range box; everything outside fails the range cull below anyway
  // no lane leaves before the warp-wide parts below: `alive` carries the reference's early returns
  const bool exists = q < pp.cl_w * pp.cl_h;
  const int px = exists ? pp.cl_px0 + q % pp.cl_w : 0, py = exists ? pp.cl_py0 + q / pp.cl_w : 0;
  const int p = py * pp.parts_x + px;
  const unsigned int x0 = px * TSD_TILE, y0 = py * TSD_TILE;
  const double cs = pp.cell_size;

  // TsdGridPartition.cpp:48-70 edge coordinates, centroid, circumradius
  const double e0x = ((double)x0 + 0.5) * cs;
  const double e0y = ((double)y0 + 0.5) * cs;
  const double e1x = ((double)(x0 + TSD_TILE) + 0.5) * cs;
  const double e2y = ((double)(y0 + TSD_TILE) + 0.5) * cs;
  const double cenx = (e0x + e1x + e0x + e1x) / 4.0;
  const double ceny = (e0y + e0y + e2y + e2y) / 4.0;
  const double ddx = e1x - e0x, ddy = e2y - e0y;
  const double circumradius = sqrt(ddx * ddx + ddy * ddy) * 0.5;
  const bool owned = (py >= pp.row_begin && py < pp.row_end);

  // state of the partition as the scans of this launch see it one after the other (first lane of the group)
  bool stateKnown = false, wasInit = false, wasInitBefore = false, allocatedHere = false;
  unsigned kinds = 0;
  double wItem[NS];
  unsigned nActive = 0, nEmptied = 0;

#pragma unroll
  for(int si = 0; si < NS; si++)
  {
    const ScanDev& s = pp.scans[si];
    const double* ranges = stagedScan ? s_ranges[si] : s.ranges;
    const uint8_t* mask = stagedScan ? s_mask[si] : s.mask;
    wItem[si] = 0.0;
    bool alive = exists;
    const double trx = s.P[2], try_ = s.P[5];
    // euklideanDistance(pos, centroid) (mathbase.h:369-378)
    double sqr = 0.0;
    {
      const double t0 = trx - cenx;
      sqr += t0 * t0;
      const double t1 = try_ - ceny;
      sqr += t1 * t1;
    }
    const double distance = sqrt(sqr);
    const double closest = distance - circumradius - pp.max_trunc;
    if(closest > s.max_range) alive = false;
    const double farthest = distance + circumradius + pp.max_trunc;
    if(farthest < s.min_range) alive = false;

    // one edge point per lane
    int idxEdge = 0;
    if(alive)
    {
      const double X = (sub & 1) ? e1x : e0x;
      const double Y = (sub & 2) ? e2y : e0y;
      idxEdge = back_project_edge(s, pp.dirs, X, Y);
    }
    bool visibleEdge = true;
    if(idxEdge == -1) { idxEdge = s.n - 1; visibleEdge = false; }
    else if(idxEdge == -2) { idxEdge = 0; visibleEdge = false; }
    if(idxEdge > s.n - 1) idxEdge = s.n - 1;  // the reference would read past the scan here
    const unsigned vis4 = (__ballot_sync(0xffffffffu, visibleEdge) >> gshift) & 0xfu;
    if(vis4 == 0u) alive = false;  // !isAnyEdgeVisible
    const bool allVisible = (vis4 == 0xfu);
    int minIdx = idxEdge, maxIdx = idxEdge;
#pragma unroll
    for(int o = 1; o < 4; o <<= 1)
    {
      const int a = __shfl_xor_sync(0xffffffffu, minIdx, o);
      const int b = __shfl_xor_sync(0xffffffffu, maxIdx, o);
      minIdx = min(minIdx, a);
      maxIdx = max(maxIdx, b);
    }

    // TsdGridComponent.cpp:96-118: the beams between the outermost edge beams.  Narrow intervals are scanned by
    // the partition's four lanes, wide ones (partitions next to the sensor) by the whole warp.
    bool vis = false, empty = true;
    const bool wide = alive && (maxIdx - minIdx > 96);
    if(alive && !wide)
    {
      for(int j = minIdx + sub; j <= maxIdx; j += 4)
      {
        const double d = ranges[j];
        const bool m = mask[j] != 0;
        vis = vis || ((d > closest) && m);
        if(isinf(d)) empty = empty && (distance < s.low_refl);
        else empty = empty && (d > farthest) && m;
      }
    }
    unsigned wideLeaders = __ballot_sync(0xffffffffu, wide && sub == 0);
    while(wideLeaders)
    {
      const int gl = __ffs(wideLeaders) - 1;
      wideLeaders &= wideLeaders - 1;
      const int lo = __shfl_sync(0xffffffffu, minIdx, gl), hi = __shfl_sync(0xffffffffu, maxIdx, gl);
      const double cl = __shfl_sync(0xffffffffu, closest, gl), fa = __shfl_sync(0xffffffffu, farthest, gl);
      const double di = __shfl_sync(0xffffffffu, distance, gl);
      bool v = false, e = true;
      for(int j = lo + lane; j <= hi; j += 32)
      {
        const double d = ranges[j];
        const bool m = mask[j] != 0;
        v = v || ((d > cl) && m);
        if(isinf(d)) e = e && (di < s.low_refl);
        else e = e && (d > fa) && m;
      }
      const bool anyV = __any_sync(0xffffffffu, v);
      const bool allE = __all_sync(0xffffffffu, e);
      if((lane >> 2) == (gl >> 2)) { vis = anyV; empty = allE; }
    }
    const unsigned mVis = __ballot_sync(0xffffffffu, vis);
    const unsigned mEmpty = __ballot_sync(0xffffffffu, empty);
    if(!alive || (mVis & gmask) == 0u || sub != 0) continue;
    const bool allEmpty = (mEmpty & gmask) == gmask;
    if(!stateKnown)
    {
      wasInit = wasInitBefore = pp.flags[p] != 0;
      stateKnown = true;
    }
    if(allVisible && allEmpty)
    {
      // increaseEmptiness (TsdGridPartition.cpp:136-164)
      if(wasInit) kinds |= 1u << (2 * si);
      else
      {
        double w = pp.initw[p];
        w += 1.0;
        w = ob_min(w, TSD_MAXWEIGHT);
        pp.initw[p] = w;
      }
      nEmptied++;
    }
    else
    {
      // active: TsdGrid.cpp:237-243
      double distCentroid = sqrt((cenx - trx) * (cenx - trx) + (ceny - try_) * (ceny - try_));
      if(distCentroid > s.max_range) distCentroid = s.max_range;
      double partWeight = (s.max_range - distCentroid) / s.max_range;
      partWeight *= partWeight;
      double w = 0.01;  // TsdGridPartition.h:194-196 (the fabs(sd) < _eps branch is dead: _eps < 0)
      w *= partWeight;
      wItem[si] = w;
      kinds |= 2u << (2 * si);
      if(!wasInit)
      {
        pp.flags[p] = 1;
        wasInit = true;
        allocatedHere = true;
      }
      nActive++;
    }
  }
  if(sub != 0 || !owned) return;  // lists and statistics are about a band's own partitions
  if(kinds)
  {
    const uint32_t slot = atomicAdd(&pp.counters[0], 1u);
    pp.active[slot] = (uint32_t)p | (wasInitBefore ? 0x80000000u : 0u);
    pp.kinds[slot] = kinds;
#pragma unroll
    for(int si = 0; si < NS; si++) pp.active_w[(size_t)si * pp.list_cap + slot] = wItem[si];
  }
  if(allocatedHere)
  {
    atomicAdd(&pp.counters[4], 1u);
    pp.newly[atomicAdd(&pp.counters[18], 1u)] = (uint32_t)p;
  }
  if(nEmptied) atomicAdd(&pp.counters[6], nEmptied);
  if(nActive) atomicAdd(&pp.counters[7], nActive);
}

#define UPDATE_CONSUMERS 256                        // threads that update cells: 8 warps, 4 cells each per partition
#define UPDATE_THREADS (UPDATE_CONSUMERS + 32)      // + one producer warp (work list, metadata, bulk copies)
#ifndef UPDATE_STAGES
#define UPDATE_STAGES 2   // measured on C2: 2 stages 76.8 us, 3 stages 85.2 us (the shared memory a third stage takes is L1 the
                          // kernel's table look-ups and register spills live in); the copies saturate HBM with 2 x 3 CTAs/SM
#endif
#ifndef UPDATE_CTAS_PER_SM
#define UPDATE_CTAS_PER_SM 3
#endif
#define TILE_BYTES (TSD_TILE_STRIDE * 8)            // one partition of one array: 8832 B = 69 x 128 B, borders included
#define STAGE_BYTES (2 * TILE_BYTES)                // tsd + weight
#define BEAM_UNDECIDED (-3)

// Beam index of the slow path (the reference formula verbatim), kept out of line: ~0.1 % of the cells.
__device__ __noinline__ int beam_index_slow(uint32_t* counter, double phi_min, double res_inv, double phi_lower, double phi_upper,
                                            double x, double y)
{
  atomicAdd(counter, 1u);
  const double phi = atan2(y, x);  // SensorPolar2D.cpp:126-134
  if(phi <= phi_lower) return -2;
  if(phi >= phi_upper) return -1;
  return (int)round((phi - phi_min) * res_inv);
}

// Branch-free front half of beam_index() (beam_index.cuh): candidate + double-precision confirmation.
// Returns the beam, -2 / -1 for points clearly outside the field of view, BEAM_UNDECIDED otherwise.
__device__ __forceinline__ int beam_index_fast(const BeamModel& bm, const double2* __restrict__ dirs, double x, double y)
{
  const float xf = (float)x, yf = (float)y;
  const float ax = fabsf(xf), ay = fabsf(yf);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  float rc;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(mx));  // mx == 0 -> NaN candidate -> undecided
  const float t = mn * rc;
  const float s = t * t;
  float p = -0.011719098314642906f;
  p = __fmaf_rn(p, s, 0.05264726281166077f);
  p = __fmaf_rn(p, s, -0.11642640829086304f);
  p = __fmaf_rn(p, s, 0.19354034960269928f);
  p = __fmaf_rn(p, s, -0.33262282609939575f);
  p = __fmaf_rn(p, s, 0.9999772310256958f);
  float r = p * t;
  r = (ay > ax) ? 1.57079632679489662f - r : r;
  r = (xf < 0.0f) ? 3.14159265358979324f - r : r;
  const float phif = copysignf(r, yf);
  const int k = __float2int_rn((phif - bm.phi_min_f) * bm.res_inv_f);
  const bool inside = (unsigned)k < (unsigned)bm.n;
  const int kc = inside ? k : 0;
  const double2 lo = __ldg(dirs + kc);
  const double2 hi = __ldg(dirs + kc + 1);
  const double m = TSD_BEAM_MARGIN * (fabs(x) + fabs(y));
  const double s_lo = lo.x * y - lo.y * x;
  const double s_hi = hi.x * y - hi.y * x;
  const bool confirmed = inside && (s_lo > m) && (s_hi < -m);
  // clearly outside the field of view: 1e-3 rad of slack on a candidate good to 2e-6 rad (lower bound first)
  const bool below = !inside && (phif < bm.phi_lower_f - 1e-3f);
  const bool above = !inside && (phif > bm.phi_upper_f + 1e-3f);
  return confirmed ? k : (below ? -2 : (above ? -1 : BEAM_UNDECIDED));
}

// ---- K2, single-precision front end: tsd_classify_cell / tsd_gate_entry / tsd_fast_model in beam_index.cuh -------------
__device__ __forceinline__ int classify_cell(const ScanDev& s, const float2* __restrict__ gate, float xf, float yf, float d2f,
                                             int& kOut)
{
  return tsd_classify_cell(s.rinv_f, s.off_f, s.half_m, s.n, gate, xf, yf, d2f, kOut);
}

// The reference's own expressions for one cell (classes 2 and 3).  Out of line (~0.3 % of the cells), and reading
// the sensor model from the copy k_classify left in global memory: a reference to the kernel's parameter block would
// make the compiler copy all of it to the stack of every thread.
__device__ __noinline__ int beam_of_point_exact(const ScanDev* __restrict__ sg, const double2* __restrict__ dirs, uint32_t* counter,
                                                double x, double y)
{
  const BeamModel bm = sg->bm;
  int i = bm.fast_ok ? beam_index_fast(bm, dirs, x, y) : BEAM_UNDECIDED;
  if(i == BEAM_UNDECIDED) i = beam_index_slow(counter, bm.phi_min, bm.res_inv, bm.phi_lower, bm.phi_upper, x, y);
  return i;
}

__device__ __forceinline__ void exact_cell(const PushParams& pp, const ScanDev& s, int si, const double* ct, const double* rt, int ix,
                                           int iy, double rD, int cls, int k, bool& ok, double& n)
{
  int idx = k;
  if(cls == 3)
  {
    const double x = (ct[ix] + rt[iy]) + s.Pi[2] * 1.0;  // SensorPolar2D.cpp:125 with gslcblas' accumulation order
    const double y = (ct[pp.cells_x + ix] + rt[pp.cells_y + iy]) + s.Pi[5] * 1.0;
    idx = beam_of_point_exact(pp.scans_dev + si, pp.dirs, pp.counters + 5, x, y);
  }
  ok = false;
  if(idx >= 0)
  {
    idx = min(idx, s.n - 1);
    const unsigned m = __ldg(s.mask + idx);
    const double r = __ldg(s.ranges + idx);
    const double dist = sqrt(ct[2 * pp.cells_x + ix] + rD);  // TsdGrid.cpp:262
    const bool inf = isinf(r);
    const double sd = inf ? pp.max_trunc : r - dist;
    ok = (m != 0) && (!inf || dist < s.low_refl) && (sd >= -pp.max_trunc);
    n = fmin(sd * pp.inv_max_trunc, 1.0);  // == obvious::min(a, 1.0): the constant is never NaN
  }
}

// Two horizontally adjacent cells of TsdGrid.cpp:250-274 + TsdGridPartition::addTsd (TsdGridPartition.h:170-212).
// cxy / cd / rxy / rd: the single-precision table entries of the two columns and of the row; returns the number of
// cells rewritten (0..2).
__device__ __forceinline__ unsigned update_pair(const PushParams& pp, const ScanDev& s, int si, int cls0, int cls1, int k0, int k1,
                                                int gx, int gy, double wTile, double2& tv, double2& wv)
{
  const double den0 = wv.x + wTile, den1 = wv.y + wTile;
  // The common pair: each cell is either not rewritten (class 0) or free space seen again -- class 1 (tsd_new == 1.0)
  // on a cell whose tsd is 1.0.  Then (1.0 * w + 1.0 * wTile) / (w + wTile) has exact products, numerator and
  // denominator are the same rounded sum, and x / x == 1.0 for every finite non-zero x: tsd stays 1.0 without a
  // division, weight = min(w + wTile, 32).  Decided with integer tests on the high words (den a positive normal
  // number: then also min(den, 32) is a comparison of high words).
  const int dh0 = __double2hiint(den0), dh1 = __double2hiint(den1);
  const bool unit0 = (__double2hiint(tv.x) == 0x3ff00000) && (__double2loint(tv.x) == 0);
  const bool unit1 = (__double2hiint(tv.y) == 0x3ff00000) && (__double2loint(tv.y) == 0);
  const bool easy0 = (cls0 == 0) || (cls0 == 1 && unit0 && (unsigned)(dh0 - 0x00100000) < 0x7fe00000u);
  const bool easy1 = (cls1 == 0) || (cls1 == 1 && unit1 && (unsigned)(dh1 - 0x00100000) < 0x7fe00000u);
  if(easy0 && easy1)
  {
    if(cls0 == 1) wv.x = (dh0 >= 0x40400000) ? TSD_MAXWEIGHT : den0;
    if(cls1 == 1) wv.y = (dh1 >= 0x40400000) ? TSD_MAXWEIGHT : den1;
    return (unsigned)(cls0 + cls1);
  }
  bool ok0 = (cls0 == 1), ok1 = (cls1 == 1);
  double n0 = 1.0, n1 = 1.0;
  if((cls0 | cls1) & 2)
  {
    const double* ct = pp.coltab + (size_t)si * 3 * pp.cells_x;
    const double* rt = pp.rowtab + (size_t)si * 3 * pp.cells_y;
    const double rD = rt[2 * pp.cells_y + gy];
    if(cls0 & 2) exact_cell(pp, s, si, ct, rt, gx, gy, rD, cls0, k0, ok0, n0);
    if(cls1 & 2) exact_cell(pp, s, si, ct, rt, gx + 1, gy, rD, cls1, k1, ok1, n1);
  }
  if(ok0 || ok1)
  {
    const bool f0 = isnan(tv.x), f1 = isnan(tv.y);  // first measurement of the cell
    const bool one0 = (tv.x == 1.0) && (n0 == 1.0) && (den0 != 0.0) && (fabs(den0) < __longlong_as_double(0x7ff0000000000000LL));
    const bool one1 = (tv.y == 1.0) && (n1 == 1.0) && (den1 != 0.0) && (fabs(den1) < __longlong_as_double(0x7ff0000000000000LL));
    double q0 = 1.0, q1 = 1.0;
    if((ok0 && !f0 && !one0) || (ok1 && !f1 && !one1))
    {
      q0 = (tv.x * wv.x + n0 * wTile) / den0;
      q1 = (tv.y * wv.y + n1 * wTile) / den1;
    }
    if(ok0)
    {
      tv.x = f0 ? n0 : q0;
      wv.x = f0 ? den0 : fmin(den0, TSD_MAXWEIGHT);
    }
    if(ok1)
    {
      tv.y = f1 ? n1 : q1;
      wv.y = f1 ? den1 : fmin(den1, TSD_MAXWEIGHT);
    }
  }
  return (ok0 ? 1u : 0u) + (ok1 ? 1u : 0u);
}

// increaseEmptiness for one cell (TsdGridPartition.cpp:140-157), branch-free
__device__ __forceinline__ void empty_cell(double& tsd, double& weight)
{
  const bool first = isnan(tsd);
  const double w1 = weight + 1.0;
  const double wn = fmin(w1, TSD_MAXWEIGHT);
  const double q = (tsd * (wn - 1.0) + 1.0) / wn;
  tsd = first ? 1.0 : q;
  weight = first ? w1 : wn;
}

// The same for two cells.  Free space that has been seen for a while sits at the weight cap: the divisor is then
// the power of two TSD_MAXWEIGHT = 32 and the IEEE quotient equals the product with 1/32 exactly (no subnormal
// results are reachable: |tsd * 31 + 1| is 0 or >= 2^-53), which spares the division in the steady state.
__device__ __forceinline__ void empty_pair(double2& tv, double2& wv)
{
  const bool f0 = isnan(tv.x), f1 = isnan(tv.y);
  const double a0 = wv.x + 1.0, a1 = wv.y + 1.0;
  const double w0 = fmin(a0, TSD_MAXWEIGHT), w1 = fmin(a1, TSD_MAXWEIGHT);
  const double n0 = tv.x * (w0 - 1.0) + 1.0, n1 = tv.y * (w1 - 1.0) + 1.0;
  double q0, q1;
  if(w0 == TSD_MAXWEIGHT && w1 == TSD_MAXWEIGHT)
  {
    q0 = n0 * (1.0 / TSD_MAXWEIGHT);
    q1 = n1 * (1.0 / TSD_MAXWEIGHT);
  }
  else
  {
    q0 = n0 / w0;
    q1 = n1 / w1;
  }
  tv.x = f0 ? 1.0 : q0;
  tv.y = f1 ? 1.0 : q1;
  wv.x = f0 ? a0 : w0;
  wv.y = f1 ? a1 : w1;
}

// Border bookkeeping inside k_update (replaces most of the reference's propagateBorders pass, TsdGrid.cpp:372-427).
// Invariant after every push: a border strip whose source neighbour (+x, +y, +xy) is initialised equals that
// neighbour's first column / row / cell.  So (a) a partition never needs to update such a strip itself
// (increaseEmptiness, init) -- whoever owns the source keeps it current -- and (b) every partition that k_update
// rewrites stores its own first column / row / cell into the strips of its -x, -y, -xy neighbours that mirror
// them.  What is left for k_borders is to *pull* the strips of partitions that became initialised outside this
// mechanism (allocated by this push, by freeFootprint or by an upload).
// Which neighbours exist and are allocated comes as a bit mask per item (the producer warp of k_update): bits 0-2 the
// mirror targets -x / -y / -xy (inside this grid or band), bits 3-5 the border sources +x / +y / +xy.
// thread (xp, y) holds the final values of cells (y, xp) and (y, xp + 1) of the partition at `base`
__device__ __forceinline__ void mirror_to_neighbours(const PushParams& pp, unsigned nb, size_t base, int xp, int y,
                                                     const double2& tv, const double2& wv)
{
  const size_t rowStride = (size_t)pp.parts_x * TSD_TILE_STRIDE;
  if(xp == 0 && (nb & 1u))
  {
    // first column -> right border of the -x neighbour
    pp.tsd[base - TSD_TILE_STRIDE + TSD_BORDER_OFF + y] = tv.x;
    pp.weight[base - TSD_TILE_STRIDE + TSD_BORDER_OFF + y] = wv.x;
  }
  if(y == 0 && (nb & 2u))
  {
    // first row -> top border of the -y neighbour
    *reinterpret_cast<double2*>(pp.tsd + base - rowStride + TSD_BORDER_OFF + 32 + xp) = tv;
    *reinterpret_cast<double2*>(pp.weight + base - rowStride + TSD_BORDER_OFF + 32 + xp) = wv;
  }
  if(y == 0 && xp == 0 && (nb & 4u))
  {
    pp.tsd[base - rowStride - TSD_TILE_STRIDE + TSD_BORDER_OFF + 64] = tv.x;
    pp.weight[base - rowStride - TSD_TILE_STRIDE + TSD_BORDER_OFF + 64] = wv.x;
  }
}

__device__ __forceinline__ void pull_pass(const PushParams& pp, int warp, int nwarps, int lane);
__device__ __forceinline__ void push_tail(const PushParams& pp);

// ---- mbarrier / bulk-copy (TMA) primitives of the k_update pipeline --------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (or the hint elapses)
// instead of returning after its short default limit -- a waiting warp issues almost nothing
__device__ __forceinline__ bool mbar_try_wait_parked(uint32_t bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}
// A warp that finds its phase incomplete sleeps between looks.  (mbarrier.try_wait with a suspend-time hint came back
// every ~50 ns on this part -- measured: 140 rounds of 4 instructions per wait on `full`, 14% of everything k_update
// issued -- and each round takes issue slots from the warps that have work.)
#ifndef UPDATE_WAIT_NS
#define UPDATE_WAIT_NS 200
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  if(mbar_try_wait(bar, parity)) return;
#if UPDATE_WAIT_NS > 0
  do { __nanosleep(UPDATE_WAIT_NS); } while(!mbar_try_wait(bar, parity));
#else
  while(!mbar_try_wait_parked(bar, parity)) {}
#endif
}
// global -> shared bulk copy (TMA unit, no register staging); completion is counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// Per item in flight: what the consumers need to know about the partition besides its cells.  The producer also
// stages the item's slices of the single-precision per-column / per-row tables here (32 entries each per scan), so
// that a consumer's chain starts with shared-memory loads instead of L2 round trips.
template <int NS>
struct UpdMeta
{
  uint32_t entry;  // work-list entry: partition | bit 31 = its cells existed before this push (and are in the stage);
                   // 0xffffffff = no more work
  uint32_t nbm;    // bits 0-5 neighbour masks, bits 8.. the 2-bit outcomes of the scans
  int gx, gy;      // grid coordinates of the partition's first cell
  size_t base;     // offset of the partition's cells in the tsd / weight arrays
  double initw;    // TsdGridPartition::_initWeight (only read for a partition this push allocates)
  double wt[NS];
  alignas(16) float2 cxy[NS][TSD_TILE];  // fill_tables: Pi00*(X-tx'), Pi10*(X-tx') of the partition's 32 columns (read in pairs)
  float2 rxy[NS][TSD_TILE];  //              Pi01*(Y-ty'), Pi11*(Y-ty') of its 32 rows
  float cd[NS][TSD_TILE];    //              (X-tx)^2
  float rd[NS][TSD_TILE];    //              (Y-ty)^2
};
// The metadata ring has one slot more than the stage ring: a consumer warp releases a stage as soon as it holds the
// cells in registers but keeps reading the item's metadata while it computes.  The producer writes slot i % (S + 1) for
// item i after every consumer warp released the stage of item i - S, which a warp does after it finished item
// i - S - 1 -- the previous user of that slot -- entirely.
#define UPDATE_META_SLOTS (UPDATE_STAGES + 1)
template <int NS>
constexpr size_t update_smem_bytes()
{
  return UPDATE_STAGES * STAGE_BYTES + UPDATE_META_SLOTS * sizeof(UpdMeta<NS>) + 2 * UPDATE_STAGES * 8;
}

// K2 + K3.  Persistent CTAs over the work list (one entry per touched partition), UPDATE_CTAS_PER_SM per SM.
// Warp 8 is the PRODUCER.  It draws chunks of consecutive list entries from a global ticket counter -- guided
// self-scheduling: a chunk is 1/(4 #CTAs) of what is left, so that the CTAs finish together although partitions cost
// between a few hundred (free space) and several thousand cycles (a surface in it, two lasers) -- one entry per lane: the
// entry, the scans' outcomes and partition weights, the allocation flags of the six neighbours that the border logic needs;
// the next chunk is fetched while the current one is fed to the pipeline.  Feeding one partition: its table slices are
// loaded (a lane per column / row), the stage is waited for (mbarrier `empty`), metadata and tables are stored, and
// the TMA unit copies the partition's 8832 B of tsd and 8832 B of weight (interior + border strips, contiguous) into
// the stage with two cp.async.bulk whose bytes complete the stage's `full` mbarrier.  No thread computes an address
// per cell, no register stages the data, and the copies of UPDATE_STAGES - 1 partitions are in flight while one is
// computed.
// Warps 0-7 are the CONSUMERS: thread t of 256 owns the cell pair x = 2*(t%16), 2*(t%16)+1 in rows t/16 and
// t/16 + 16 (16-byte shared loads, a warp reads two adjacent 256-B rows).  They wait on `full`, take their
// cells (and the border cell they look after) into registers, release the stage (one arrive per warp), apply
// the scans of the launch one after the other (a partition seen by both lasers of a robot is read and written
// once), and store what changed with 16-byte stores.
template <int NS>
__global__ void __launch_bounds__(UPDATE_THREADS, UPDATE_CTAS_PER_SM) k_update(PushParams pp)
{
  extern __shared__ __align__(128) unsigned char smem[];
  UpdMeta<NS>* s_meta = reinterpret_cast<UpdMeta<NS>*>(smem + UPDATE_STAGES * STAGE_BYTES);
  const uint32_t bar0 = smem_u32(smem + UPDATE_STAGES * STAGE_BYTES + UPDATE_META_SLOTS * sizeof(UpdMeta<NS>));  // full[s], then empty[s]
  const uint32_t nItems = pp.counters[0];
  const uint32_t G = gridDim.x;
  const int t = threadIdx.x;
  const int lane = t & 31;
  if(t == 0)
  {
    for(int s = 0; s < UPDATE_STAGES; s++)
    {
      mbar_init(bar0 + 8 * s, 1);                                         // the producer's arrive (+ the copies' bytes)
      mbar_init(bar0 + 8 * (UPDATE_STAGES + s), UPDATE_CONSUMERS / 32);   // one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  unsigned updates = 0, emptied = 0;

  if(t >= UPDATE_CONSUMERS)
  {
    // ------------------------------------------------------------------------------------------ producer warp
    int s = 0, ms = 0;
    uint32_t round = 0;
    // one chunk of the work list in registers, a lane per entry
    struct Chunk
    {
      uint32_t start, cnt, e, m;
      double wt[NS];
    };
    // The first chunk of every CTA is static (entries [blockIdx * first, +first)): no ticket, no contention of all CTAs
    // on one counter at the start of the launch; the tickets hand out what follows.
    const uint32_t first = min(32u, nItems / (4 * G));
    auto grab = [&](uint32_t seen, bool initial) -> Chunk  // seen: entries known to be handed out already
    {
      Chunk c;
      uint32_t start = 0, want = 0;
      if(initial && first > 0)
      {
        start = blockIdx.x * first;
        want = first;
      }
      else
      {
        if(lane == 0)
        {
          const uint32_t left = nItems > seen ? nItems - seen : 0u;
          want = min(32u, max(1u, left / (6 * G)));
          start = atomicAdd(&pp.counters[19], want) + G * first;
        }
        start = __shfl_sync(0xffffffffu, start, 0);
        want = __shfl_sync(0xffffffffu, want, 0);
      }
      c.start = start;
      c.cnt = start < nItems ? min(want, nItems - start) : 0u;
      c.e = 0xffffffffu;
      c.m = 0;
#pragma unroll
      for(int si = 0; si < NS; si++) c.wt[si] = 0.0;
      if((uint32_t)lane < c.cnt)
      {
        const uint32_t it = start + (uint32_t)lane;
        c.e = pp.active[it];
#pragma unroll
        for(int si = 0; si < NS; si++) c.wt[si] = pp.active_w[(size_t)si * pp.list_cap + it];
        const int p = (int)(c.e & 0x7fffffffu);
        const int px = p & (pp.parts_x - 1), py = p >> pp.parts_shift;
        const bool hasL = px > 0, hasD = py > pp.row_begin;
        const bool hasR = px < pp.parts_x - 1, hasU = py < pp.parts_y - 1;
        unsigned m = pp.kinds[it] << 8;
        if(hasL && pp.flags[p - 1]) m |= 1u;
        if(hasD && pp.flags[p - pp.parts_x]) m |= 2u;
        if(hasL && hasD && pp.flags[p - pp.parts_x - 1]) m |= 4u;
        if(hasR && pp.flags[p + 1]) m |= 8u;
        if(hasU && pp.flags[p + pp.parts_x]) m |= 16u;
        if(hasR && hasU && pp.flags[p + pp.parts_x + 1]) m |= 32u;
        c.m = m;
      }
      return c;
    };
#ifdef UPDATE_PROFILE
    long long pr_t0 = clock64(), pr_wait = 0, pr_grab = 0, pr_items = 0;
#endif
    Chunk cur = grab(G * first, true);
#ifdef UPDATE_PROFILE
    pr_grab += clock64() - pr_t0;
#endif
    while(true)
    {
      // the next chunk's ticket and metadata are on their way while this one is fed (cur.cnt == 0: the sentinel)
      Chunk nxt;
      nxt.cnt = 0;
#ifdef UPDATE_PROFILE
      long long g0 = clock64();
#endif
      if(cur.cnt) nxt = grab(max(cur.start + cur.cnt, G * first), false);
#ifdef UPDATE_PROFILE
      pr_grab += clock64() - g0;
      pr_items += cur.cnt;
#endif
      const int cnt = cur.cnt ? (int)cur.cnt : 1;
      for(int j = 0; j < cnt; j++)
      {
        const uint32_t ej = __shfl_sync(0xffffffffu, cur.e, j);
        const uint32_t mj = __shfl_sync(0xffffffffu, cur.m, j);
        double wj[NS];
#pragma unroll
        for(int si = 0; si < NS; si++) wj[si] = __shfl_sync(0xffffffffu, cur.wt[si], j);
        const uint32_t full = bar0 + 8 * s, empty = bar0 + 8 * (UPDATE_STAGES + s);
#ifdef UPDATE_PROFILE
        long long w0 = clock64();
#endif
        if(lane == 0) mbar_wait(empty, (round & 1u) ^ 1u);  // (passes at once the first time round)
#ifdef UPDATE_PROFILE
        pr_wait += clock64() - w0;
#endif
        UpdMeta<NS>& meta = s_meta[ms];
        const bool item = ej != 0xffffffffu;
        const bool cells = item && (ej & 0x80000000u);
        const uint32_t p = ej & 0x7fffffffu;
        const int gx = (int)(p & (uint32_t)(pp.parts_x - 1)) * TSD_TILE;
        const int gy = (int)(p >> pp.parts_shift) * TSD_TILE;
        const size_t nb = (size_t)(p - pp.alloc_begin * pp.parts_x) * TSD_TILE_STRIDE;
        if(lane == 0)
        {
          meta.entry = ej;
          meta.nbm = mj;
          if(item)
          {
            meta.gx = gx;
            meta.gy = gy;
            meta.base = nb;
            meta.initw = cells ? 0.0 : pp.initw[p];
          }
#pragma unroll
          for(int si = 0; si < NS; si++) meta.wt[si] = wj[si];
          // one arrival for the metadata just written + the bytes the copies below will deliver
          if(item) mbar_arrive_expect_tx(full, (cells ? (uint32_t)STAGE_BYTES : 0u) + (uint32_t)(NS * 6 * 128));
          else mbar_arrive(full);  // the end of the list
        }
        __syncwarp();
        // Everything else comes by bulk copy, one copy per lane: the partition's tsd and weight (lanes 0, 1) and, per
        // scan, its 32-entry slices of the four single-precision tables (256 + 256 + 128 + 128 B, 16-byte aligned:
        // gx and gy are multiples of 32).  Nothing passes through this warp's registers, nothing is waited for here.
        if(item)
        {
          const void* src = nullptr;
          uint32_t dst = 0, bytes = 0;
          if(lane < 2)
          {
            if(cells)
            {
              src = (lane == 0 ? pp.tsd : pp.weight) + nb;
              dst = smem_u32(smem + (size_t)s * STAGE_BYTES) + (uint32_t)lane * TILE_BYTES;
              bytes = TILE_BYTES;
            }
          }
          else if(lane < 2 + 4 * NS)
          {
            const int si = (lane - 2) >> 2, which = (lane - 2) & 3;
            if(which == 0) { src = pp.colxy + (size_t)si * pp.cells_x + gx; dst = smem_u32(&meta.cxy[si][0]); bytes = 256; }
            else if(which == 1) { src = pp.rowxy + (size_t)si * pp.cells_y + gy; dst = smem_u32(&meta.rxy[si][0]); bytes = 256; }
            else if(which == 2) { src = pp.cold + (size_t)si * pp.cells_x + gx; dst = smem_u32(&meta.cd[si][0]); bytes = 128; }
            else { src = pp.rowd + (size_t)si * pp.cells_y + gy; dst = smem_u32(&meta.rd[si][0]); bytes = 128; }
          }
          if(bytes) bulk_g2s(dst, src, bytes, full);
        }
        if(++s == UPDATE_STAGES) { s = 0; round++; }
        if(++ms == UPDATE_META_SLOTS) ms = 0;
      }
      if(cur.cnt == 0) break;
      cur = nxt;
    }
#ifdef UPDATE_PROFILE
    if(lane == 0 && pp.prof)
    {
      pp.prof[blockIdx.x * 8 + 0] = clock64() - pr_t0;
      pp.prof[blockIdx.x * 8 + 1] = pr_wait;
      pp.prof[blockIdx.x * 8 + 2] = pr_grab;
      pp.prof[blockIdx.x * 8 + 3] = pr_items;
    }
#endif
  }
  else
  {
    // ------------------------------------------------------------------------------------------ consumer warps
    const int xp = (t & 15) * 2;
    const int yb = t >> 4;
    const bool edge = (xp == 0) || (yb == 0);
    const int ci0 = yb * TSD_TILE + xp, ci1 = ci0 + 16 * TSD_TILE;  // this thread's cell pairs: rows yb and yb + 16
    int s = 0, ms = 0;
    uint32_t par = 0;
#ifdef UPDATE_PROFILE
    long long co_t0 = clock64(), co_wait = 0, co_first = 0;
#endif
    while(true)
    {
      const uint32_t full = bar0 + 8 * s, empty = bar0 + 8 * (UPDATE_STAGES + s);
#ifdef UPDATE_PROFILE
      long long w0 = clock64();
#endif
      mbar_wait(full, par);
#ifdef UPDATE_PROFILE
      { long long w1 = clock64(); if(co_first == 0) co_first = w1 - co_t0; else co_wait += w1 - w0; }
#endif
      const double* st = reinterpret_cast<const double*>(smem + (size_t)s * STAGE_BYTES);
      const UpdMeta<NS>& meta = s_meta[ms];
      const uint32_t eCur = meta.entry;
      if(eCur == 0xffffffffu) break;
      const unsigned nbm = meta.nbm;
      if(++s == UPDATE_STAGES) { s = 0; par ^= 1u; }
      if(++ms == UPDATE_META_SLOTS) ms = 0;
      const unsigned kinds = nbm >> 8;
      const bool wasInit = (eCur & 0x80000000u) != 0;
      // the border cell this thread looks after (t < 65): only strips WITHOUT an allocated source neighbour are
      // the partition's own business (the others mirror the neighbour, which keeps them current)
      const bool myStrip = t < 65 && !((nbm >> (t < 32 ? 3 : (t < 64 ? 4 : 5))) & 1u);
      // a partition this push allocates starts from TsdGridPartition::init (TsdGridPartition.cpp:98-119): in the list
      // because some scan found it active, and no scan before that one did anything to it
      const size_t base = meta.base;
      double* T = pp.tsd + base;
      double* W = pp.weight + base;
      double2 tv0, wv0, tv1, wv1;
      if(wasInit)
      {
        tv0 = *reinterpret_cast<const double2*>(st + ci0);
        wv0 = *reinterpret_cast<const double2*>(st + TSD_TILE_STRIDE + ci0);
        tv1 = *reinterpret_cast<const double2*>(st + ci1);
        wv1 = *reinterpret_cast<const double2*>(st + TSD_TILE_STRIDE + ci1);
      }
      else
      {
        const double initW = meta.initw;
        const double initT = (initW > 0.0) ? 1.0 : __longlong_as_double(0x7ff8000000000000LL);
        tv0 = tv1 = make_double2(initT, initT);
        wv0 = wv1 = make_double2(initW, initW);
      }
      // The border cell this thread looks after is finished first (it only sees the partition's initialisation and
      // the increaseEmptiness scans, never addTsd), so that it does not occupy registers during the cell update.
      if(myStrip && (!wasInit || (kinds & 0x55u)))
      {
        double bt, bw;
        if(wasInit)
        {
          bt = st[TSD_BORDER_OFF + t];
          bw = st[TSD_TILE_STRIDE + TSD_BORDER_OFF + t];
        }
        else
        {
          bw = meta.initw;
          bt = (bw > 0.0) ? 1.0 : __longlong_as_double(0x7ff8000000000000LL);
        }
        if(!(pp.update_filter & 2u))
          for(unsigned k = kinds; k; k >>= 2)
            if((k & 3u) == 1u) empty_cell(bt, bw);
        T[TSD_BORDER_OFF + t] = bt;
        W[TSD_BORDER_OFF + t] = bw;
      }
      __syncwarp();
      if(lane == 0) mbar_arrive(empty);  // this warp has taken everything it needs out of the stage
      bool dirty0 = !wasInit, dirty1 = !wasInit;
      // The scans one after the other, rolled (the body unrolled NS times is beyond the instruction cache: measured,
      // 39 % of all stall samples "no instruction"); the two cell pairs of a thread side by side.
#pragma unroll 1
      for(int si = 0; si < NS; si++)
      {
        const unsigned kind = (kinds >> (2 * si)) & 3u;
        if(kind == 2u && !(pp.update_filter & 1u))
        {
          const float4 cxy = *reinterpret_cast<const float4*>(&meta.cxy[si][xp]);
          const float2 cd = *reinterpret_cast<const float2*>(&meta.cd[si][xp]);
          const float2 rxy0 = meta.rxy[si][yb], rxy1 = meta.rxy[si][yb + 16];
          const float rd0 = meta.rd[si][yb], rd1 = meta.rd[si][yb + 16];
          const double wTile = meta.wt[si];
          const float2* gate = pp.gate + (size_t)si * pp.scan_cap;
          const ScanDev& sc = pp.scans[si];
          // the four cells are classified side by side (four independent dependency chains in straight-line code),
          // then the two pairs are updated
          int k00, k01, k10, k11;
          const int c00 = classify_cell(sc, gate, cxy.x + rxy0.x, cxy.y + rxy0.y, cd.x + rd0, k00);
          const int c01 = classify_cell(sc, gate, cxy.z + rxy0.x, cxy.w + rxy0.y, cd.y + rd0, k01);
          const int c10 = classify_cell(sc, gate, cxy.x + rxy1.x, cxy.y + rxy1.y, cd.x + rd1, k10);
          const int c11 = classify_cell(sc, gate, cxy.z + rxy1.x, cxy.w + rxy1.y, cd.y + rd1, k11);
          const unsigned u0 = update_pair(pp, sc, si, c00, c01, k00, k01, meta.gx + xp, meta.gy + yb, wTile, tv0, wv0);
          const unsigned u1 = update_pair(pp, sc, si, c10, c11, k10, k11, meta.gx + xp, meta.gy + yb + 16, wTile, tv1, wv1);
          updates += u0 + u1;
          dirty0 = dirty0 || (u0 != 0u);
          dirty1 = dirty1 || (u1 != 0u);
        }
        else if(kind == 1u && !(pp.update_filter & 2u))
        {
          // K3: increaseEmptiness on an allocated partition, all 33x33 cells
          empty_pair(tv0, wv0);
          empty_pair(tv1, wv1);
          dirty0 = dirty1 = true;
          if(t == 0) emptied++;
        }
      }
      if(dirty0)
      {
        *reinterpret_cast<double2*>(T + ci0) = tv0;
        *reinterpret_cast<double2*>(W + ci0) = wv0;
      }
      if(dirty1)
      {
        *reinterpret_cast<double2*>(T + ci1) = tv1;
        *reinterpret_cast<double2*>(W + ci1) = wv1;
      }
      if(edge && (nbm & 7u))
      {
        mirror_to_neighbours(pp, nbm & 7u, base, xp, yb, tv0, wv0);
        mirror_to_neighbours(pp, nbm & 7u, base, xp, yb + 16, tv1, wv1);
      }
    }
#ifdef UPDATE_PROFILE
    if(t == 0 && pp.prof)
    {
      pp.prof[blockIdx.x * 8 + 4] = clock64() - co_t0;
      pp.prof[blockIdx.x * 8 + 5] = co_wait;
      pp.prof[blockIdx.x * 8 + 6] = co_first;
      unsigned smid;
      asm("mov.u32 %0, %%smid;" : "=r"(smid));
      pp.prof[blockIdx.x * 8 + 7] = smid;
    }
#endif
  }

  // one atomic per CTA
  __shared__ unsigned long long s_upd[UPDATE_THREADS / 32];
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) updates += __shfl_xor_sync(0xffffffffu, updates, o);
  if(lane == 0) s_upd[t >> 5] = (unsigned long long)updates + (unsigned long long)emptied * (33 * 33);
  __syncthreads();
  if(t == 0)
  {
    unsigned long long u = 0;
    for(int i = 0; i < UPDATE_THREADS / 32; i++) u += s_upd[i];
    if(u) atomicAdd(&pp.stats64[0], u);
  }
  if(pp.fused_tail)
  {
    // the last CTA to get here has seen every partition of this push written: it pulls the borders of the
    // partitions this push allocated (usually none) and closes the push
    __shared__ unsigned s_last;
    if(t == 0)
    {
      __threadfence();
      s_last = (atomicAdd(&pp.counters[17], 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if(s_last)
    {
      __threadfence();
      pull_pass(pp, t >> 5, UPDATE_THREADS / 32, lane);
      __syncthreads();
      if(t == 0)
      {
        __threadfence();
        push_tail(pp);
        pp.counters[17] = 0;
      }
    }
  }
}


// K4: TsdGrid::propagateBorders (TsdGrid.cpp:372-427) restricted to the partitions whose cells changed
// in this push or since the last one (pending): a touched partition refreshes its own
// right/top/corner border from its +x/+y/+xy neighbours, and the border of its -x/-y/-xy neighbours that
// mirrors its first column / row / cell.  Untouched pairs keep the values of the previous push, which is
// what the reference's full pass would rewrite them with.
__device__ __forceinline__ void refresh_borders_of(const PushParams& pp, int px, int py, int lane)
{
  // cur = (px,py) must be initialised and owned
  const int p = py * pp.parts_x + px;
  const size_t base = (size_t)(p - pp.alloc_begin * pp.parts_x) * TSD_TILE_STRIDE;
  double* T = pp.tsd + base;
  double* W = pp.weight + base;
  if(px < pp.parts_x - 1 && pp.flags[p + 1])
  {
    const size_t nb = base + TSD_TILE_STRIDE;
    T[TSD_BORDER_OFF + lane] = pp.tsd[nb + lane * TSD_TILE];
    W[TSD_BORDER_OFF + lane] = pp.weight[nb + lane * TSD_TILE];
  }
  if(py < pp.parts_y - 1 && py + 1 < pp.alloc_end && pp.flags[p + pp.parts_x])
  {
    const size_t nb = base + (size_t)pp.parts_x * TSD_TILE_STRIDE;
    T[TSD_BORDER_OFF + 32 + lane] = pp.tsd[nb + lane];
    W[TSD_BORDER_OFF + 32 + lane] = pp.weight[nb + lane];
  }
  if(lane == 0 && px < pp.parts_x - 1 && py < pp.parts_y - 1 && py + 1 < pp.alloc_end && pp.flags[p + pp.parts_x + 1])
  {
    const size_t nb = base + (size_t)(pp.parts_x + 1) * TSD_TILE_STRIDE;
    T[TSD_BORDER_OFF + 64] = pp.tsd[nb];
    W[TSD_BORDER_OFF + 64] = pp.weight[nb];
  }
}

// The six border strips that depend on one touched partition T: T's own right / top / corner (from its +x, +y,
// +xy neighbours) and the strips of its -x, -y, -xy neighbours that mirror T's first column / row / cell.  All
// loads are issued before the first store: one memory round trip per touched partition.
__device__ __forceinline__ void refresh_strips_around(const PushParams& pp, int px, int py, int lane)
{
  const int p = py * pp.parts_x + px;
  const size_t base = (size_t)(p - pp.alloc_begin * pp.parts_x) * TSD_TILE_STRIDE;
  const size_t rowStride = (size_t)pp.parts_x * TSD_TILE_STRIDE;
  const bool hasR = px < pp.parts_x - 1 && pp.flags[p + 1];
  const bool hasU = py < pp.parts_y - 1 && py + 1 < pp.alloc_end && pp.flags[p + pp.parts_x];
  const bool hasUR = px < pp.parts_x - 1 && py < pp.parts_y - 1 && py + 1 < pp.alloc_end && pp.flags[p + pp.parts_x + 1];
  const bool hasL = px > 0 && pp.flags[p - 1];
  const bool hasD = py > pp.row_begin && pp.flags[p - pp.parts_x];
  const bool hasDL = px > 0 && py > pp.row_begin && pp.flags[p - pp.parts_x - 1];
  double tR = 0, wR = 0, tU = 0, wU = 0, tC = 0, wC = 0, tc0 = 0, wc0 = 0, tr0 = 0, wr0 = 0;
  if(hasR) { tR = pp.tsd[base + TSD_TILE_STRIDE + lane * TSD_TILE]; wR = pp.weight[base + TSD_TILE_STRIDE + lane * TSD_TILE]; }
  if(hasU) { tU = pp.tsd[base + rowStride + lane]; wU = pp.weight[base + rowStride + lane]; }
  if(hasUR && lane == 0) { tC = pp.tsd[base + rowStride + TSD_TILE_STRIDE]; wC = pp.weight[base + rowStride + TSD_TILE_STRIDE]; }
  if(hasL) { tc0 = pp.tsd[base + lane * TSD_TILE]; wc0 = pp.weight[base + lane * TSD_TILE]; }
  if(hasD || hasDL) { tr0 = pp.tsd[base + lane]; wr0 = pp.weight[base + lane]; }
  if(hasR) { pp.tsd[base + TSD_BORDER_OFF + lane] = tR; pp.weight[base + TSD_BORDER_OFF + lane] = wR; }
  if(hasU) { pp.tsd[base + TSD_BORDER_OFF + 32 + lane] = tU; pp.weight[base + TSD_BORDER_OFF + 32 + lane] = wU; }
  if(hasUR && lane == 0) { pp.tsd[base + TSD_BORDER_OFF + 64] = tC; pp.weight[base + TSD_BORDER_OFF + 64] = wC; }
  if(hasL)
  {
    pp.tsd[base - TSD_TILE_STRIDE + TSD_BORDER_OFF + lane] = tc0;
    pp.weight[base - TSD_TILE_STRIDE + TSD_BORDER_OFF + lane] = wc0;
  }
  if(hasD)
  {
    pp.tsd[base - rowStride + TSD_BORDER_OFF + 32 + lane] = tr0;
    pp.weight[base - rowStride + TSD_BORDER_OFF + 32 + lane] = wr0;
  }
  if(hasDL && lane == 0)
  {
    pp.tsd[base - rowStride - TSD_TILE_STRIDE + TSD_BORDER_OFF + 64] = tr0;
    pp.weight[base - rowStride - TSD_TILE_STRIDE + TSD_BORDER_OFF + 64] = wr0;
  }
}

// pull pass: partitions allocated by this push (their strips were not maintained before) and partitions
// allocated / modified outside push since the last one
__device__ __forceinline__ void pull_pass(const PushParams& pp, int warp, int nwarps, int lane)
{
  const uint32_t nN = pp.counters[18], nP = pp.counters[2];
  for(uint32_t item = warp; item < nN + nP; item += nwarps)
  {
    const uint32_t p = (item < nN) ? pp.newly[item] : pp.pending[item - nN];
    if(!pp.flags[p]) continue;
    refresh_strips_around(pp, p % pp.parts_x, p / pp.parts_x, lane);
  }
}

// snapshot the per-push statistics (so that a later push does not clobber them before they are read) and zero
// the per-push counters, the pending list and the refresh-all flag
__device__ __forceinline__ void push_tail(const PushParams& pp)
{
  for(int i = 0; i < 8; i++) pp.counters[8 + i] = pp.counters[i];
  pp.stats64[1] = pp.stats64[0];
  pp.counters[20] = (uint32_t)(pp.stats64[0] & 0xffffffffULL);  // the same figure next to the counters: one D2H copy
  pp.counters[21] = (uint32_t)(pp.stats64[0] >> 32);
  for(int i = 0; i < 8; i++) pp.counters[i] = 0;
  pp.counters[18] = 0;
  pp.counters[19] = 0;  // k_update's work ticket
  pp.stats64[0] = 0;
}

// Sharded grid, after boundary rows were exchanged: (a) the band's top row takes its top / corner strips from the
// halo row above (those partitions belong to another GPU and may have changed); (b) the halo row BELOW the band
// arrived with the top / corner strips its owner had before it saw this band's new first row -- they mirror this
// band's own cells, so they are completed here with exactly the values the owner computes under (a).  With (b) one
// exchange per synchronisation suffices.  Columns [lo0, hi0] / [lo1, hi1]: what changed at the lower / upper
// boundary (a corner strip looks one partition to the right, hence the -1).
__device__ __forceinline__ void band_boundary_refresh(const PushParams& pp, int lo0, int hi0, int lo1, int hi1, int warp,
                                                      int nwarps, int lane)
{
  if(pp.row_end < pp.parts_y && hi1 >= lo1)
    for(int px = max(lo1 - 1, 0) + warp; px <= hi1; px += nwarps)
      if(pp.flags[(pp.row_end - 1) * pp.parts_x + px]) refresh_borders_of(pp, px, pp.row_end - 1, lane);
  if(pp.row_begin > 0 && hi0 >= lo0)
    for(int px = max(lo0 - 1, 0) + warp; px <= hi0; px += nwarps)
      if(pp.flags[(pp.row_begin - 1) * pp.parts_x + px]) refresh_borders_of(pp, px, pp.row_begin - 1, lane);
}

// Halo synchronisation of a sharded grid in ONE kernel over peer memory (NVLink / NVSwitch P2P stores; the
// neighbours' arrays are mapped with CUDA IPC, or are plain device pointers when both bands live in one process):
//   A  tell both neighbours "I am done reading the halo rows you gave me last time" (stream order guarantees it),
//   B  wait for the same message from them,
//   C  store this band's lowest / highest partition row (the dirty columns) straight into the neighbours' halo rows,
//   D  fence, then the last CTA raises "data ready" at both neighbours,
//   E  wait for the neighbours' "data ready",
//   F  band_boundary_refresh.
// Signals are sequence numbers (release / acquire at system scope); both sides of a boundary count alike.
struct HaloParams
{
  PushParams pp;
  int active[2];          // boundary below / above takes part in this synchronisation
  int px0[2], px1[2];     // dirty partition columns per boundary
  const double* src_t[2]; // this band's lowest / highest owned row
  const double* src_w[2];
  double* dst_t[2];       // the neighbour's halo row above / below ITS band
  double* dst_w[2];
  uint32_t* peer_sig[2];
  uint32_t* my_sig;
  uint32_t seq[2];
};

__global__ void __launch_bounds__(256) k_halo_sync(HaloParams hp)
{
  const int t = threadIdx.x;
  // A: my signal slots at the neighbour: [1]/[3] at the band below (I am its upper neighbour), [0]/[2] above
  if(blockIdx.x == 0 && t < 2 && hp.active[t]) st_release_sys(hp.peer_sig[t] + (t == 0 ? 3 : 2), hp.seq[t]);
  // B
  if(t < 2 && hp.active[t]) wait_seq(hp.my_sig + 2 + t, hp.seq[t], hp.my_sig + 5);
  __syncthreads();
  // C: 16-byte stores; a partition is 8832 B = 552 x 16 B, so every column offset is aligned
  for(int b = 0; b < 2; b++)
  {
    if(!hp.active[b]) continue;
    const size_t off = (size_t)hp.px0[b] * TSD_TILE_STRIDE;
    const size_t n2 = (size_t)(hp.px1[b] - hp.px0[b] + 1) * (TSD_TILE_STRIDE / 2);
    const double2* st = reinterpret_cast<const double2*>(hp.src_t[b] + off);
    const double2* sw = reinterpret_cast<const double2*>(hp.src_w[b] + off);
    double2* dt = reinterpret_cast<double2*>(hp.dst_t[b] + off);
    double2* dw = reinterpret_cast<double2*>(hp.dst_w[b] + off);
    for(size_t i = (size_t)blockIdx.x * blockDim.x + t; i < n2; i += (size_t)gridDim.x * blockDim.x)
    {
      dt[i] = st[i];
      dw[i] = sw[i];
    }
  }
  // D
  __threadfence_system();
  __syncthreads();
  if(t == 0)
  {
    if(atomicAdd(hp.my_sig + 4, 1u) == gridDim.x - 1)
    {
      hp.my_sig[4] = 0;
      __threadfence_system();
      for(int b = 0; b < 2; b++)
        if(hp.active[b]) st_release_sys(hp.peer_sig[b] + (b == 0 ? 1 : 0), hp.seq[b]);
    }
  }
  // E
  if(t < 2 && hp.active[t]) wait_seq(hp.my_sig + t, hp.seq[t], hp.my_sig + 5);
  __syncthreads();
  // F
  const int warp = (blockIdx.x * blockDim.x + t) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  band_boundary_refresh(hp.pp, hp.active[0] ? hp.px0[0] : 0, hp.active[0] ? hp.px1[0] : -1, hp.active[1] ? hp.px0[1] : 0,
                        hp.active[1] ? hp.px1[1] : -1, warp, nwarps, t & 31);
}

// K4 as a kernel of its own.  mode 0: end of a push that was asked to refresh every border (after a fill / upload /
// free-footprint), with the push tail; mode 1: refresh everything, no push involved; mode 2: sharded grid, after
// the halo exchange: the band's top row takes its top / corner strips from the halo row above (the partitions
// there belong to another GPU and may have changed).
__global__ void __launch_bounds__(256) k_borders(PushParams pp, int mode)
{
  const int all = (mode == 1);
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const bool refreshAll = (mode != 2) && (all || pp.counters[3]);
  if(refreshAll)
  {
    for(int p = pp.row_begin * pp.parts_x + warp; p < pp.row_end * pp.parts_x; p += nwarps)
      if(pp.flags[p]) refresh_borders_of(pp, p % pp.parts_x, p / pp.parts_x, lane);
  }
  if(mode == 2)
  {
    band_boundary_refresh(pp, 0, pp.parts_x - 1, 0, pp.parts_x - 1, warp, nwarps, lane);
    return;
  }
  if(!refreshAll) pull_pass(pp, warp, nwarps, lane);
  __syncthreads();
  if(threadIdx.x == 0)
  {
    __threadfence();
    const unsigned ticket = atomicAdd(&pp.counters[16], 1u);
    if(ticket == gridDim.x - 1)
    {
      push_tail(pp);
      pp.counters[16] = 0;
    }
  }
}

// TsdGrid::freeFootprint (TsdGrid.cpp:609-638), phase 1: initialise the partitions under the footprint
__global__ void k_footprint_init(PushParams pp, int pxMin, int pxMax, int pyMin, int pyMax)
{
  const int w = pxMax - pxMin + 1;
  const int tile = blockIdx.x;
  const int px = pxMin + tile % w, py = pyMin + tile / w;
  if(py > pyMax) return;
  const int p = py * pp.parts_x + px;
  if(pp.flags[p]) return;
  if(py >= pp.row_begin && py < pp.row_end)
  {
    const double initW = pp.initw[p];
    const double initT = (initW > 0.0) ? 1.0 : __longlong_as_double(0x7ff8000000000000LL);
    const size_t base = (size_t)(p - pp.alloc_begin * pp.parts_x) * TSD_TILE_STRIDE;
    for(int i = threadIdx.x; i < TSD_BORDER_OFF + 65; i += blockDim.x)
    {
      pp.tsd[base + i] = initT;
      pp.weight[base + i] = initW;
    }
    if(threadIdx.x == 0) pp.pending[atomicAdd(&pp.counters[2], 1u)] = (uint32_t)p;
  }
  __syncthreads();
  if(threadIdx.x == 0) pp.flags[p] = 1;
}

// phase 2: tsd = TSDINC for every cell of the footprint rectangle
__global__ void k_footprint_set(PushParams pp, unsigned minX, unsigned maxX, unsigned minY, unsigned maxY)
{
  const unsigned w = maxX - minX;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if(w == 0 || i >= w * (maxY - minY)) return;
  const unsigned cols = minX + i % w, rows = minY + i / w;
  const int py = rows >> 5, px = cols >> 5;
  if(py < pp.row_begin || py >= pp.row_end) return;
  const int p = py * pp.parts_x + px;
  const size_t base = (size_t)(p - pp.alloc_begin * pp.parts_x) * TSD_TILE_STRIDE;
  pp.tsd[base + (rows & 31) * TSD_TILE + (cols & 31)] = 1.0;
}

// one CTA-iteration per partition; only_uninit: leave allocated partitions alone
__global__ void k_fill(PushParams pp, double tsd, double weight, int only_uninit)
{
  for(int p = blockIdx.x; p < pp.n_parts; p += gridDim.x)
  {
    const bool skip = only_uninit && pp.flags[p];
    __syncthreads();
    if(skip) continue;
    const int py = p / pp.parts_x;
    if(py >= pp.row_begin && py < pp.row_end)
    {
      const size_t base = (size_t)(p - pp.alloc_begin * pp.parts_x) * TSD_TILE_STRIDE;
      for(int i = threadIdx.x; i < TSD_BORDER_OFF + 65; i += blockDim.x)
      {
        pp.tsd[base + i] = tsd;
        pp.weight[base + i] = weight;
      }
    }
    if(threadIdx.x == 0) pp.flags[p] = 1;
  }
}

__global__ void k_interpolate(GridView g, int n, const double* xy, double* tsd, int* status)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  double v = __longlong_as_double(0x7ff8000000000000LL);
  status[i] = sample_bilinear(g, xy[2 * i], xy[2 * i + 1], &v);
  tsd[i] = v;
}

__global__ void k_interpolate_normal(GridView g, int n, const double* xy, double* normals, int* ok)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  double nx = __longlong_as_double(0x7ff8000000000000LL), ny = nx;
  ok[i] = sample_normal(g, xy[2 * i], xy[2 * i + 1], &nx, &ny) ? 1 : 0;
  normals[2 * i] = nx;
  normals[2 * i + 1] = ny;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

static PushParams make_params(const tsd_grid* g)
{
  PushParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.cell_size = g->cell_size;
  pp.max_trunc = g->max_truncation;
  pp.inv_max_trunc = 1.0 / g->max_truncation;
  pp.cells_x = g->cells_x;
  pp.cells_y = g->cells_y;
  pp.parts_x = g->parts_x;
  pp.parts_shift = 0;
  while((1 << pp.parts_shift) < g->parts_x) pp.parts_shift++;
  pp.parts_y = g->parts_y;
  pp.n_parts = g->n_parts;
  pp.row_begin = g->row_begin;
  pp.row_end = g->row_end;
  pp.alloc_begin = g->alloc_begin;
  pp.alloc_end = g->alloc_end;
  pp.band = g->band ? 1 : 0;
  pp.tsd = g->d_tsd;
  pp.weight = g->d_weight;
  pp.flags = g->d_flags;
  pp.initw = g->d_initw;
  pp.active = g->d_active;
  pp.active_w = g->d_active_w;
  pp.kinds = g->d_kinds;
  pp.list_cap = g->n_owned;
  pp.newly = g->d_newly;
  pp.pending = g->d_pending;
  pp.counters = g->d_counters;
  pp.stats64 = g->d_stats64;
  pp.coltab = g->d_coltab;
  pp.rowtab = g->d_rowtab;
  pp.dirs = g->d_dirs;
  pp.colxy = reinterpret_cast<float2*>(g->d_col4);  // one allocation: cells_x float2 per scan, then cells_x float per scan
  pp.cold = reinterpret_cast<float*>(g->d_col4) + (size_t)2 * g->cells_x * PUSH_MAX_SCANS;
  pp.rowxy = reinterpret_cast<float2*>(g->d_row4);
  pp.rowd = reinterpret_cast<float*>(g->d_row4) + (size_t)2 * g->cells_y * PUSH_MAX_SCANS;
  pp.gate = g->d_gate;
  pp.scan_cap = g->scan_cap;
  pp.update_filter = g->update_filter;
  pp.scans_dev = g->d_scans;
  pp.prof = g->d_prof;
  return pp;
}

namespace tsd
{

int grid_ensure_scratch(tsd_grid* g, size_t bytes)
{
  if(bytes <= g->scratch_cap) return TSD_OK;
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  if(g->d_scratch) cudaFree(g->d_scratch);
  if(g->h_scratch) cudaFreeHost(g->h_scratch);
  g->d_scratch = g->h_scratch = nullptr;
  g->scratch_cap = 0;
  size_t cap = 1 << 16;
  while(cap < bytes) cap <<= 1;
  TSD_CUDA(cudaMalloc(&g->d_scratch, cap));
  TSD_CUDA(cudaMallocHost(&g->h_scratch, cap));
  g->scratch_cap = cap;
  return TSD_OK;
}

// One device block and one pinned mirror hold everything a scan brings with it, so that staging is a single
// H2D copy:  [ ranges: cap doubles | mask: cap bytes | rays: 2*cap doubles | (ranges | mask) of scans 1..3 of a batch ]
// (a push of one scan or a ray cast copies a prefix; a batch copies through its last scan).  The ray-caster's results come
// back in one D2H copy of  [ out: 4*cap doubles | keys: cap u64 | steps: 2 u64 ].
static int ensure_scan_capacity(tsd_grid* g, int n)
{
  if(n <= g->scan_cap) return TSD_OK;
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  cudaFree(g->d_in); cudaFree(g->d_dirs); cudaFree(g->d_rc); cudaFree(g->d_gate);
  cudaFreeHost(g->h_in2[0]); cudaFreeHost(g->h_in2[1]); cudaFreeHost(g->h_rc);
  g->d_in = g->d_rc = nullptr; g->d_dirs = nullptr; g->d_gate = nullptr; g->h_in = g->h_in2[0] = g->h_in2[1] = g->h_rc = nullptr;
  g->ev_in_used[0] = g->ev_in_used[1] = false;
  g->scan_cap = 0;
  const int cap = ((n + 63) / 64) * 64 + 64;
  g->in_bytes = PUSH_MAX_SCANS * (sizeof(double) * cap + cap) + sizeof(double) * 2 * cap;
  g->rc_bytes = sizeof(double) * 4 * cap + sizeof(unsigned long long) * cap + sizeof(unsigned long long) * 2;
  TSD_CUDA(cudaMalloc(&g->d_in, g->in_bytes));
  TSD_CUDA(cudaMalloc(&g->d_rc, g->rc_bytes));
  TSD_CUDA(cudaMalloc(&g->d_dirs, sizeof(double2) * (cap + 1)));
  TSD_CUDA(cudaMalloc(&g->d_gate, sizeof(float2) * (size_t)cap * PUSH_MAX_SCANS));
  // (the padding behind a scan's last beam is copied into shared memory with the scan, 16 bytes at a time: defined bytes)
  TSD_CUDA(cudaMemsetAsync(g->d_in, 0, g->in_bytes, g->stream));
  for(int i = 0; i < 2; i++)
  {
    TSD_CUDA(cudaMallocHost(&g->h_in2[i], g->in_bytes));
    memset(g->h_in2[i], 0, g->in_bytes);
    if(!g->ev_in[i]) TSD_CUDA(cudaEventCreateWithFlags(&g->ev_in[i], cudaEventDisableTiming));
  }
  g->h_in = g->h_in2[0];
  g->in_next = 0;
  TSD_CUDA(cudaMallocHost(&g->h_rc, g->rc_bytes));
  TSD_CUDA(cudaMemsetAsync(g->d_rc, 0, g->rc_bytes, g->stream));
  memset(g->h_rc, 0, g->rc_bytes);
  g->d_ranges = reinterpret_cast<double*>(g->d_in);
  g->d_mask = g->d_in + sizeof(double) * cap;
  g->d_rays = reinterpret_cast<double*>(g->d_in + sizeof(double) * cap + cap);
  g->h_ranges = reinterpret_cast<double*>(g->h_in);
  g->h_mask = g->h_in + sizeof(double) * cap;
  g->h_rays = reinterpret_cast<double*>(g->h_in + sizeof(double) * cap + cap);
  g->d_rc_out = reinterpret_cast<double*>(g->d_rc);
  g->d_rc_keys = reinterpret_cast<unsigned long long*>(g->d_rc + sizeof(double) * 4 * cap);
  g->d_rc_steps = g->d_rc_keys + cap;
  g->h_rc_out = reinterpret_cast<double*>(g->h_rc);
  g->h_rc_keys = reinterpret_cast<unsigned long long*>(g->h_rc + sizeof(double) * 4 * cap);
  g->h_rc_steps = g->h_rc_keys + cap;
  g->rc_steps_prev[0] = g->rc_steps_prev[1] = 0;
  g->scan_cap = cap;
  g->dirs_n = -1;
  return TSD_OK;
}

int grid_stage_scan(tsd_grid* g, const tsd_scan_t* scan, ScanDev* sd, const double* rays_world)
{
  return grid_stage_scans(g, scan, 1, sd, rays_world);
}

// Stages 1 or 2 scans (same sensor model) with ONE H2D copy; sd[i] gets scan i.
int grid_stage_scans(tsd_grid* g, const tsd_scan_t* scans, int n, ScanDev* sd, const double* rays_world)
{
  if(!scans || n < 1 || n > PUSH_MAX_SCANS) { set_error("invalid scan batch"); return TSD_E_INVALID; }
  for(int i = 0; i < n; i++)
    if(scans[i].n < 1 || !scans[i].ranges || !scans[i].mask) { set_error("invalid scan"); return TSD_E_INVALID; }
  const tsd_scan_t* scan = &scans[0];
  for(int i = 1; i < n; i++)
    if(scans[i].n != scan->n || scans[i].phi_min != scan->phi_min || scans[i].angular_res != scan->angular_res)
    {
      set_error("the scans of a batch must come from the same sensor model");
      return TSD_E_INVALID;
    }
  if(n == 3) { set_error("a launch takes 1, 2 or 4 scans"); return TSD_E_INVALID; }
  int rc = ensure_scan_capacity(g, scan->n);
  if(rc) return rc;
  // Two pinned blocks take turns: only the copy out of THIS block (two staging calls ago) must have drained, not the
  // kernels of the previous push -- tsdg_push_async returns while the previous push is still running.
  const int blk = g->in_next;
  g->in_next ^= 1;
  if(g->ev_in_used[blk]) TSD_CUDA(cudaEventSynchronize(g->ev_in[blk]));
  g->h_in = g->h_in2[blk];
  g->h_ranges = reinterpret_cast<double*>(g->h_in);
  g->h_mask = g->h_in + sizeof(double) * g->scan_cap;
  g->h_rays = reinterpret_cast<double*>(g->h_in + sizeof(double) * g->scan_cap + g->scan_cap);
  const size_t slotBytes = sizeof(double) * g->scan_cap + g->scan_cap;
  const size_t raysBytes = sizeof(double) * 2 * g->scan_cap;
  size_t bytes = 0;
  for(int i = 0; i < n; i++)
  {
    const size_t off = (i == 0) ? 0 : slotBytes + raysBytes + (size_t)(i - 1) * slotBytes;
    fill_scan_dev(&scans[i], &sd[i]);
    memcpy(g->h_in + off, scans[i].ranges, sizeof(double) * scans[i].n);
    memcpy(g->h_in + off + sizeof(double) * g->scan_cap, scans[i].mask, scans[i].n);
    sd[i].ranges = reinterpret_cast<double*>(g->d_in + off);
    sd[i].mask = g->d_in + off + sizeof(double) * g->scan_cap;
    bytes = off + sizeof(double) * g->scan_cap + scans[i].n;
  }
  if(rays_world)
  {
    memcpy(g->h_rays, rays_world, sizeof(double) * 2 * scan->n);
    const size_t withRays = slotBytes + sizeof(double) * 2 * scan->n;
    bytes = bytes > withRays ? bytes : withRays;
  }
  TSD_CUDA(cudaMemcpyAsync(g->d_in, g->h_in, bytes, cudaMemcpyHostToDevice, g->stream));
  TSD_CUDA(cudaEventRecord(g->ev_in[blk], g->stream));
  g->ev_in_used[blk] = true;
  if(g->dirs_n != scan->n || g->dirs_phi_min != scan->phi_min || g->dirs_res != scan->angular_res)
  {
    // directions of the half-beam boundaries B_k = phiMin + (k - 1/2) res, k = 0..n (beam_index.cuh)
    std::vector<double2> dirs(scan->n + 1);
    for(int k = 0; k <= scan->n; k++)
    {
      const double b = scan->phi_min + ((double)k - 0.5) * scan->angular_res;
      dirs[k].x = cos(b);
      dirs[k].y = sin(b);
    }
    TSD_CUDA(cudaMemcpyAsync(g->d_dirs, dirs.data(), sizeof(double2) * (scan->n + 1), cudaMemcpyHostToDevice, g->stream));
    TSD_CUDA(cudaStreamSynchronize(g->stream));
    g->dirs_n = scan->n;
    g->dirs_phi_min = scan->phi_min;
    g->dirs_res = scan->angular_res;
  }
  return TSD_OK;
}

}  // namespace tsd

extern "C" {

const char* tsd_last_error(void) { return tsd::t_error.c_str(); }

int tsd_device_count(void)
{
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

uint64_t tsd_kernel_launches(void) { return tsd::g_launches.load(); }

// Same statement order as the GSL-shim LU the oracle runs (oracle/shim/gsl_shim.c, oracle/port/grid.c):
// partial pivoting with first maximum, reciprocal scaling, rank-1 update, column-wise forward/back solves.
int tsd_invert3x3(const double in[9], double out[9])
{
  if(!in || !out) return TSD_E_INVALID;
  double A[3][3];
  int perm[3] = {0, 1, 2};
  for(int i = 0; i < 3; i++) for(int j = 0; j < 3; j++) A[i][j] = in[3 * i + j];
  for(int j = 0; j < 3; j++)
  {
    double max = fabs(A[j][j]);
    int ip = j;
    for(int i = j + 1; i < 3; i++)
    {
      const double a = fabs(A[i][j]);
      if(a > max) { max = a; ip = i; }
    }
    if(ip != j)
    {
      for(int k = 0; k < 3; k++) { const double t = A[j][k]; A[j][k] = A[ip][k]; A[ip][k] = t; }
      const int t = perm[j]; perm[j] = perm[ip]; perm[ip] = t;
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
    double x[3];
    for(int i = 0; i < 3; i++) x[i] = (perm[i] == c) ? 1.0 : 0.0;
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
  return TSD_OK;
}

int tsdg_create_band(double cell_size, int layout_partition, int layout_grid, int device, int part_row_begin,
                     int part_row_end, tsd_grid_t** out)
{
  if(!out) return TSD_E_INVALID;
  *out = nullptr;
  if(layout_partition != 5) { set_error("only LAYOUT_32x32 partitions are supported (the node's layout)"); return TSD_E_INVALID; }
  if(layout_grid < 5 || layout_grid > 16 || !(cell_size > 0.0)) { set_error("invalid grid layout"); return TSD_E_INVALID; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if(e != cudaSuccess || ndev == 0)
  {
    cudaGetLastError();
    set_error("no CUDA device: libtsdslam_b200 has no CPU path");
    return TSD_E_NO_DEVICE;
  }
  if(device < 0 || device >= ndev) { set_error("invalid device ordinal %d", device); return TSD_E_INVALID; }
  TSD_CUDA(cudaSetDevice(device));
  tsd_grid* g = new tsd_grid();
  memset(g, 0, sizeof(*g));
  g->mtx = new std::recursive_mutex();
  g->device = device;
  g->layout_grid = layout_grid;
  g->cell_size = cell_size;
  g->inv_cell_size = 1.0 / cell_size;   // TsdGrid.cpp:117
  g->cells_x = 1 << layout_grid;
  g->cells_y = g->cells_x;
  g->parts_x = g->cells_x / TSD_TILE;
  g->parts_y = g->cells_y / TSD_TILE;
  g->n_parts = g->parts_x * g->parts_y;
  if(part_row_begin < 0) part_row_begin = 0;
  if(part_row_end < 0 || part_row_end > g->parts_y) part_row_end = g->parts_y;
  if(part_row_begin >= part_row_end) { delete g; set_error("empty band"); return TSD_E_INVALID; }
  g->row_begin = part_row_begin;
  g->row_end = part_row_end;
  g->n_owned = (part_row_end - part_row_begin) * g->parts_x;
  g->band = (part_row_begin > 0 || part_row_end < g->parts_y);
  g->alloc_begin = g->band && part_row_begin > 0 ? part_row_begin - 1 : part_row_begin;
  g->alloc_end = g->band && part_row_end < g->parts_y ? part_row_end + 1 : part_row_end;
  g->n_alloc = (g->alloc_end - g->alloc_begin) * g->parts_x;
  g->max_truncation = 2.0 * cell_size;  // TsdGrid.cpp:136
  g->min_x = 0.0;
  g->max_x = ((double)g->cells_x + 0.5) * cell_size;  // TsdGrid.cpp:141-144
  g->min_y = 0.0;
  g->max_y = ((double)g->cells_y + 0.5) * cell_size;
  g->dirs_n = -1;
  cudaDeviceProp prop;
  TSD_CUDA(cudaGetDeviceProperties(&prop, device));
  g->sm_count = prop.multiProcessorCount;
  TSD_CUDA(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
  const size_t cellBytes = sizeof(double) * (size_t)g->n_alloc * TSD_TILE_STRIDE;
  e = cudaMalloc(&g->d_tsd, cellBytes);
  if(e == cudaSuccess) e = cudaMalloc(&g->d_weight, cellBytes);
  if(e != cudaSuccess)
  {
    cudaGetLastError();
    set_error("cudaMalloc of %zu bytes of cell state failed: %s", 2 * cellBytes, cudaGetErrorString(e));
    tsdg_destroy(g);
    return TSD_E_NOMEM;
  }
  TSD_CUDA(cudaMalloc(&g->d_flags, g->n_parts));
  TSD_CUDA(cudaMalloc(&g->d_initw, sizeof(double) * g->n_parts));
  TSD_CUDA(cudaMalloc(&g->d_active, sizeof(uint32_t) * g->n_owned));
  TSD_CUDA(cudaMalloc(&g->d_active_w, sizeof(double) * g->n_owned * PUSH_MAX_SCANS));
  TSD_CUDA(cudaMalloc(&g->d_newly, sizeof(uint32_t) * g->n_owned));
  TSD_CUDA(cudaMalloc(&g->d_signal, sizeof(uint32_t) * 8));
  TSD_CUDA(cudaMemset(g->d_signal, 0, sizeof(uint32_t) * 8));
  TSD_CUDA(cudaMalloc(&g->d_pending, sizeof(uint32_t) * g->n_owned));
  TSD_CUDA(cudaMalloc(&g->d_counters, sizeof(uint32_t) * 32));
  TSD_CUDA(cudaMalloc(&g->d_stats64, sizeof(unsigned long long) * 4));
  TSD_CUDA(cudaMalloc(&g->d_coltab, sizeof(double) * 3 * g->cells_x * PUSH_MAX_SCANS));
  TSD_CUDA(cudaMalloc(&g->d_rowtab, sizeof(double) * 3 * g->cells_y * PUSH_MAX_SCANS));
  TSD_CUDA(cudaMalloc(&g->d_kinds, sizeof(uint32_t) * g->n_owned));
  TSD_CUDA(cudaMalloc(&g->d_scans, sizeof(tsd::ScanDev) * PUSH_MAX_SCANS));
#ifdef UPDATE_PROFILE
  TSD_CUDA(cudaMalloc(&g->d_prof, sizeof(unsigned long long) * 8 * g->sm_count * UPDATE_CTAS_PER_SM));
  TSD_CUDA(cudaMemset(g->d_prof, 0, sizeof(unsigned long long) * 8 * g->sm_count * UPDATE_CTAS_PER_SM));
#endif
  TSD_CUDA(cudaMalloc(&g->d_col4, sizeof(float4) * g->cells_x * PUSH_MAX_SCANS));
  TSD_CUDA(cudaMalloc(&g->d_row4, sizeof(float4) * g->cells_y * PUSH_MAX_SCANS));
  TSD_CUDA(cudaMallocHost(&g->h_counters, sizeof(uint32_t) * 32));
  TSD_CUDA(cudaMallocHost(&g->h_stats64, sizeof(unsigned long long) * 4));
  TSD_CUDA(cudaMemsetAsync(g->d_flags, 0, g->n_parts, g->stream));
  TSD_CUDA(cudaMemsetAsync(g->d_initw, 0, sizeof(double) * g->n_parts, g->stream));
  TSD_CUDA(cudaMemsetAsync(g->d_counters, 0, sizeof(uint32_t) * 32, g->stream));
  TSD_CUDA(cudaMemsetAsync(g->d_stats64, 0, sizeof(unsigned long long) * 4, g->stream));
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  *out = g;
  return TSD_OK;
}

int tsdg_create(double cell_size, int layout_partition, int layout_grid, int device, tsd_grid_t** out)
{
  return tsdg_create_band(cell_size, layout_partition, layout_grid, device, 0, -1, out);
}

int tsdg_destroy(tsd_grid_t* g)
{
  if(!g) return TSD_OK;
  cudaSetDevice(g->device);
  if(g->stream) cudaStreamSynchronize(g->stream);
  cudaFree(g->d_tsd); cudaFree(g->d_weight); cudaFree(g->d_flags); cudaFree(g->d_initw); cudaFree(g->d_active);
  cudaFree(g->d_active_w); cudaFree(g->d_newly); cudaFree(g->d_signal);
  for(int b = 0; b < 2; b++)
    if(g->peer[b].connected && g->peer[b].ipc)
    {
      cudaIpcCloseMemHandle(g->peer[b].tsd); cudaIpcCloseMemHandle(g->peer[b].weight); cudaIpcCloseMemHandle(g->peer[b].signal);
    }
  for(int r = 0; r < 16; r++)
    if(g->peer_rcx[r] && g->peer_rcx_ipc[r]) cudaIpcCloseMemHandle(g->peer_rcx[r]);
  cudaFree(g->d_rcx);
  cudaFree(g->d_pending); cudaFree(g->d_counters);
  cudaFree(g->d_stats64); cudaFree(g->d_coltab); cudaFree(g->d_rowtab); cudaFree(g->d_col4); cudaFree(g->d_row4); cudaFree(g->d_scans); cudaFree(g->d_prof); cudaFree(g->d_gate); cudaFree(g->d_kinds); cudaFree(g->d_dirs); cudaFree(g->d_in);
  cudaFree(g->d_rc); cudaFree(g->d_scratch);
  cudaFreeHost(g->h_in2[0]); cudaFreeHost(g->h_in2[1]); cudaFreeHost(g->h_rc); cudaFreeHost(g->h_scratch); cudaFreeHost(g->h_counters);
  for(int i = 0; i < 2; i++) if(g->ev_in[i]) cudaEventDestroy(g->ev_in[i]);
  cudaFreeHost(g->h_stats64);
  for(int i = 0; i < 4; i++) if(g->ev[i]) cudaEventDestroy(g->ev[i]);
  if(g->ev_order) cudaEventDestroy(g->ev_order);
  if(g->stream) cudaStreamDestroy(g->stream);
  cudaGetLastError();
  delete g->mtx;
  delete g;
  return TSD_OK;
}

int tsdg_set_max_truncation(tsd_grid_t* g, double val)
{
  TSD_LOCK(g);
  if(!g) return TSD_E_INVALID;
  if(val < 2 * g->cell_size) val = 2 * g->cell_size;  // TsdGrid.cpp:208-212
  g->max_truncation = val;
  return TSD_OK;
}

int tsdg_get_geometry(const tsd_grid_t* g, int32_t* cells_x, int32_t* cells_y, int32_t* partition_size,
                      double* cell_size, double* min_x, double* max_x, double* min_y, double* max_y,
                      double* max_truncation)
{
  TSD_LOCK(g);
  if(!g) return TSD_E_INVALID;
  if(cells_x) *cells_x = g->cells_x;
  if(cells_y) *cells_y = g->cells_y;
  if(partition_size) *partition_size = TSD_TILE;
  if(cell_size) *cell_size = g->cell_size;
  if(min_x) *min_x = g->min_x;
  if(max_x) *max_x = g->max_x;
  if(min_y) *min_y = g->min_y;
  if(max_y) *max_y = g->max_y;
  if(max_truncation) *max_truncation = g->max_truncation;
  return TSD_OK;
}

int tsdg_free_footprint(tsd_grid_t* g, double cx, double cy, double width, double height)
{
  TSD_LOCK(g);
  if(!g) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  // TsdGrid.cpp:611-622
  const unsigned minX = static_cast<unsigned int>((cx - width * 0.5) / g->cell_size + 0.5);
  const unsigned maxX = static_cast<unsigned int>((cx + width * 0.5) / g->cell_size + 0.5);
  const unsigned minY = static_cast<unsigned int>((cy - height * 0.5) / g->cell_size + 0.5);
  const unsigned maxY = static_cast<unsigned int>((cy + height * 0.5) / g->cell_size + 0.5);
  if((minX > (unsigned)g->cells_x) || (maxX > (unsigned)g->cells_x) || (minY > (unsigned)g->cells_y) ||
     (maxY > (unsigned)g->cells_y))
  {
    set_error("freeFootprint: indices out of bounds");
    return TSD_E_RANGE;
  }
  if(maxX <= minX || maxY <= minY) return TSD_OK;
  PushParams pp = make_params(g);
  const int pxMin = minX >> 5, pxMax = (maxX - 1) >> 5, pyMin = minY >> 5, pyMax = (maxY - 1) >> 5;
  const int tiles = (pxMax - pxMin + 1) * (pyMax - pyMin + 1);
  k_footprint_init<<<tiles, 256, 0, g->stream>>>(pp, pxMin, pxMax, pyMin, pyMax);
  TSD_LAUNCHED();
  const unsigned cells = (maxX - minX) * (maxY - minY);
  k_footprint_set<<<(cells + 255) / 256, 256, 0, g->stream>>>(pp, minX, maxX, minY, maxY);
  TSD_LAUNCHED();
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  return TSD_OK;
}

static int push_finish(tsd_grid* g, const PushParams& pp);

// Partitions a scan taken at (tx, ty) can touch: everything else fails the range cull of isInRange
// (TsdGridComponent.cpp:50-58: centroid distance - circumradius - maxTruncation > maxRange).  One partition of
// slack on every side covers the half-cell offsets and the rounding of the cull itself.  box = {px0, py0, px1, py1},
// inclusive, clipped to the grid; a non-finite pose or range selects the whole grid.
static void scan_partition_box(const tsd_grid* g, double tx, double ty, double max_range, int box[4])
{
  box[0] = 0; box[1] = 0; box[2] = g->parts_x - 1; box[3] = g->parts_y - 1;
  const double part = TSD_TILE * g->cell_size;
  const double reach = max_range + g->max_truncation + part;  // circumradius < one partition edge
  if(!(std::isfinite(tx) && std::isfinite(ty) && std::isfinite(reach)) || !(reach >= 0.0)) return;
  const double lo[2] = {std::floor((tx - reach) / part) - 1.0, std::floor((ty - reach) / part) - 1.0};
  const double hi[2] = {std::floor((tx + reach) / part) + 1.0, std::floor((ty + reach) / part) + 1.0};
  const int n[2] = {g->parts_x, g->parts_y};
  for(int a = 0; a < 2; a++)
  {
    // a box entirely outside the grid degenerates to one row / column of partitions, all of which are culled
    const double l = lo[a] < 0.0 ? 0.0 : (lo[a] > n[a] - 1 ? n[a] - 1 : lo[a]);
    const double h = hi[a] < 0.0 ? 0.0 : (hi[a] > n[a] - 1 ? n[a] - 1 : hi[a]);
    box[a] = (int)l;
    box[2 + a] = (int)h;
  }
}

int tsdg_stage_scan(tsd_grid_t* g, const tsd_scan_t* scan) { return tsdg_stage_batch(g, scan, 1); }

int tsdg_stage_batch(tsd_grid_t* g, const tsd_scan_t* scans, int32_t n)
{
  TSD_LOCK(g);
  if(!g) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  int rc = grid_stage_scans(g, scans, n, g->staged, nullptr);
  if(rc) return rc;
  g->has_staged = true;
  g->staged_n = n;
  return TSD_OK;
}

int tsdg_push_staged(tsd_grid_t* g)
{
  TSD_LOCK(g);
  if(!g || !g->has_staged) { set_error("no staged scan"); return TSD_E_INVALID; }
  TSD_CUDA(cudaSetDevice(g->device));
  PushParams pp = make_params(g);
  const int ns = g->staged_n;
  pp.nscan = ns;
  pp.scan_cap = g->scan_cap;
  for(int i = 0; i < ns; i++) pp.scans[i] = g->staged[i];
  pp.dirs = g->d_dirs;
  g->stats_fresh = false;
  // counters [0] work-list entries, [4..7] statistics are per push; [2] pending and [3] refresh-all persist until
  // consumed by the push tail, which also zeroes the per-push ones
  int box[4];
  scan_partition_box(g, pp.scans[0].P[2], pp.scans[0].P[5], pp.scans[0].max_range, box);
  for(int i = 1; i < ns; i++)
  {
    int b2[4];
    scan_partition_box(g, pp.scans[i].P[2], pp.scans[i].P[5], pp.scans[i].max_range, b2);
    box[0] = b2[0] < box[0] ? b2[0] : box[0];
    box[1] = b2[1] < box[1] ? b2[1] : box[1];
    box[2] = b2[2] > box[2] ? b2[2] : box[2];
    box[3] = b2[3] > box[3] ? b2[3] : box[3];
  }
  if(g->band)
  {
    // a band keeps allocation flags / emptiness weights of its own rows and of the two rows next to them
    if(box[1] < g->row_begin - 1) box[1] = g->row_begin - 1;
    if(box[3] > g->row_end) box[3] = g->row_end;
    if(box[1] < 0) box[1] = 0;
    if(box[3] > g->parts_y - 1) box[3] = g->parts_y - 1;
    if(box[3] < box[1]) return TSD_OK;  // the scan cannot reach this band
  }
  pp.cl_px0 = box[0];
  pp.cl_py0 = box[1];
  pp.cl_w = box[2] - box[0] + 1;
  pp.cl_h = box[3] - box[1] + 1;
  int nmax = TSD_TILE * (pp.cl_w > pp.cl_h ? pp.cl_w : pp.cl_h);  // the first threads also fill the per-push tables
  if(nmax < pp.scans[0].n) nmax = pp.scans[0].n;
  const int nthreads = (4 * pp.cl_w * pp.cl_h > nmax) ? 4 * pp.cl_w * pp.cl_h : nmax;
  if(g->timing) TSD_CUDA(cudaEventRecord(g->ev[0], g->stream));
  const int cctas = (nthreads + CLASSIFY_THREADS - 1) / CLASSIFY_THREADS;
  if(ns == 4) k_classify<4><<<cctas, CLASSIFY_THREADS, 0, g->stream>>>(pp, g->d_coltab, g->d_rowtab);
  else if(ns == 2) k_classify<2><<<cctas, CLASSIFY_THREADS, 0, g->stream>>>(pp, g->d_coltab, g->d_rowtab);
  else k_classify<1><<<cctas, CLASSIFY_THREADS, 0, g->stream>>>(pp, g->d_coltab, g->d_rowtab);
  TSD_LAUNCHED();
  if(g->timing) TSD_CUDA(cudaEventRecord(g->ev[1], g->stream));
  const size_t smem = ns == 4 ? update_smem_bytes<4>() : (ns == 2 ? update_smem_bytes<2>() : update_smem_bytes<1>());
  {
    // the pipeline stages live in dynamic shared memory beyond the 48 KB default: opt in once per device
    static std::mutex mtx;
    static bool done[64] = {false};
    std::lock_guard<std::mutex> lk(mtx);
    if(!done[g->device & 63])
    {
      TSD_CUDA(cudaFuncSetAttribute(k_update<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_smem_bytes<1>()));
      TSD_CUDA(cudaFuncSetAttribute(k_update<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_smem_bytes<2>()));
      TSD_CUDA(cudaFuncSetAttribute(k_update<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_smem_bytes<4>()));
      done[g->device & 63] = true;
    }
  }
  int ctas = g->sm_count * UPDATE_CTAS_PER_SM;
  if(ctas > g->n_owned) ctas = g->n_owned;
  // nothing asked for a full border refresh: K4's remainder runs in k_update's last CTA.  (Sharded grids too: the
  // strips that depend on the halo row above the band are refreshed after the exchange, tsdg_band_push_finish.)
  pp.fused_tail = g->refresh_all_pending ? 0 : 1;
  if(g->band) g->band_push_open = true;
  if(ns == 4)
  {
    k_update<4><<<ctas, UPDATE_THREADS, smem, g->stream>>>(pp);
  }
  else if(ns == 2)
  {
    k_update<2><<<ctas, UPDATE_THREADS, smem, g->stream>>>(pp);
  }
  else
  {
    k_update<1><<<ctas, UPDATE_THREADS, smem, g->stream>>>(pp);
  }
  TSD_LAUNCHED();
  if(g->timing) TSD_CUDA(cudaEventRecord(g->ev[2], g->stream));
  if(pp.fused_tail)
  {
    if(g->timing) TSD_CUDA(cudaEventRecord(g->ev[3], g->stream));
    g->pushed_once = true;
    return TSD_OK;
  }
  return push_finish(g, pp);
}

static int push_finish(tsd_grid* g, const PushParams& pp)
{
  g->refresh_all_pending = false;
  int bctas = g->sm_count;
  k_borders<<<bctas, 256, 0, g->stream>>>(pp, 0);
  TSD_LAUNCHED();
  if(g->timing) TSD_CUDA(cudaEventRecord(g->ev[3], g->stream));
  // (pending list, refresh-all flag and per-push counters are consumed by k_borders' last CTA; statistics
  //  are fetched on demand)
  g->pushed_once = true;
  return TSD_OK;
}

int tsdg_push_async(tsd_grid_t* g, const tsd_scan_t* scan)
{
  TSD_LOCK(g);
  int rc = tsdg_stage_scan(g, scan);
  if(rc) return rc;
  return tsdg_push_staged(g);
}

// Integrates n scans in the order given, as n TsdGrid::push calls would (ThreadMapping::eventLoop drains its queue of
// sensors one push after the other, ThreadMapping.cpp:43-62).  Two scans of the same sensor model at a time share
// one classify + one update launch: partitions both scans touch are read and written once.
int tsdg_push_batch_async(tsd_grid_t* g, const tsd_scan_t* scans, int32_t n)
{
  TSD_LOCK(g);
  if(!g || !scans || n < 1) return TSD_E_INVALID;
  int i = 0;
  while(i < n)
  {
    auto same = [&](int a, int b)
    {
      return scans[a].n == scans[b].n && scans[a].phi_min == scans[b].phi_min && scans[a].angular_res == scans[b].angular_res;
    };
    int take = 1;
    if(i + 3 < n && same(i, i + 1) && same(i, i + 2) && same(i, i + 3)) take = 4;
    else if(i + 1 < n && same(i, i + 1)) take = 2;
    int rc = tsdg_stage_batch(g, scans + i, take);
    if(rc) return rc;
    rc = tsdg_push_staged(g);
    if(rc) return rc;
    i += take;
  }
  return TSD_OK;
}

int tsdg_push_batch(tsd_grid_t* g, const tsd_scan_t* scans, int32_t n)
{
  TSD_LOCK(g);
  int rc = tsdg_push_batch_async(g, scans, n);
  if(rc) return rc;
  // statistics of the LAST launch of the batch (both scans of a pair together) come back with the synchronisation
  TSD_CUDA(cudaMemcpyAsync(g->h_counters, g->d_counters, sizeof(uint32_t) * 24, cudaMemcpyDeviceToHost, g->stream));
  rc = tsdg_sync(g);
  g->stats_fresh = (rc == TSD_OK);
  return rc;
}

void* tsdg_stream(tsd_grid_t* g) { return g ? (void*)g->stream : nullptr; }

// Stream-ordered hand-over between the handle's stream and a caller's stream (the one its collectives run on),
// without blocking the host: direction 0 makes `other` wait for everything queued on the handle's stream,
// direction 1 makes the handle's stream wait for everything queued on `other`.
int tsdg_stream_order(tsd_grid_t* g, void* other, int direction)
{
  TSD_LOCK(g);
  if(!g) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  if(!g->ev_order) TSD_CUDA(cudaEventCreateWithFlags(&g->ev_order, cudaEventDisableTiming));
  cudaStream_t o = (cudaStream_t)other;
  if(direction == 0)
  {
    TSD_CUDA(cudaEventRecord(g->ev_order, g->stream));
    TSD_CUDA(cudaStreamWaitEvent(o, g->ev_order, 0));
  }
  else
  {
    TSD_CUDA(cudaEventRecord(g->ev_order, o));
    TSD_CUDA(cudaStreamWaitEvent(g->stream, g->ev_order, 0));
  }
  return TSD_OK;
}

int tsdg_band_push_finish(tsd_grid_t* g)
{
  TSD_LOCK(g);
  if(!g || !g->band) { set_error("not a sharded grid"); return TSD_E_INVALID; }
  TSD_CUDA(cudaSetDevice(g->device));
  PushParams pp = make_params(g);
  g->band_push_open = false;
  if(g->row_end < g->parts_y || g->row_begin > 0)
  {
    k_borders<<<g->sm_count, 256, 0, g->stream>>>(pp, 2);
    TSD_LAUNCHED();
  }
  return TSD_OK;
}

// --- halo synchronisation over peer memory ---------------------------------------------------------------
struct BandExport
{
  cudaIpcMemHandle_t tsd, weight, signal;
  int32_t alloc_begin, row_begin, row_end, parts_x;
};
static_assert(sizeof(BandExport) <= TSD_BAND_EXPORT_BYTES, "tsd_band_export_t too small");

int tsdg_band_export(tsd_grid_t* g, void* blob)
{
  TSD_LOCK(g);
  if(!g || !blob || !g->band) { set_error("not a sharded grid"); return TSD_E_INVALID; }
  TSD_CUDA(cudaSetDevice(g->device));
  BandExport e;
  memset(&e, 0, sizeof(e));
  TSD_CUDA(cudaIpcGetMemHandle(&e.tsd, g->d_tsd));
  TSD_CUDA(cudaIpcGetMemHandle(&e.weight, g->d_weight));
  TSD_CUDA(cudaIpcGetMemHandle(&e.signal, g->d_signal));
  e.alloc_begin = g->alloc_begin;
  e.row_begin = g->row_begin;
  e.row_end = g->row_end;
  e.parts_x = g->parts_x;
  memset(blob, 0, TSD_BAND_EXPORT_BYTES);
  memcpy(blob, &e, sizeof(e));
  return TSD_OK;
}

static int check_neighbour(const tsd_grid* g, int side, int nb_row_begin, int nb_row_end, int nb_parts_x)
{
  if(nb_parts_x != g->parts_x) { set_error("neighbouring band has another grid geometry"); return TSD_E_INVALID; }
  if(side == 0 && nb_row_end != g->row_begin) { set_error("band below does not end where this band begins"); return TSD_E_INVALID; }
  if(side == 1 && nb_row_begin != g->row_end) { set_error("band above does not begin where this band ends"); return TSD_E_INVALID; }
  return TSD_OK;
}

int tsdg_band_connect(tsd_grid_t* g, int side, const void* blob)
{
  TSD_LOCK(g);
  if(!g || !blob || !g->band || side < 0 || side > 1) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  BandExport e;
  memcpy(&e, blob, sizeof(e));
  int rc = check_neighbour(g, side, e.row_begin, e.row_end, e.parts_x);
  if(rc) return rc;
  tsd_grid::Peer& pr = g->peer[side];
  if(pr.connected) { set_error("neighbour already connected"); return TSD_E_INVALID; }
  void *pt = nullptr, *pw = nullptr, *ps = nullptr;
  TSD_CUDA(cudaIpcOpenMemHandle(&pt, e.tsd, cudaIpcMemLazyEnablePeerAccess));
  TSD_CUDA(cudaIpcOpenMemHandle(&pw, e.weight, cudaIpcMemLazyEnablePeerAccess));
  TSD_CUDA(cudaIpcOpenMemHandle(&ps, e.signal, cudaIpcMemLazyEnablePeerAccess));
  pr.tsd = static_cast<double*>(pt);
  pr.weight = static_cast<double*>(pw);
  pr.signal = static_cast<uint32_t*>(ps);
  pr.alloc_begin = e.alloc_begin;
  pr.ipc = true;
  pr.connected = true;
  return TSD_OK;
}

int tsdg_band_connect_local(tsd_grid_t* g, int side, tsd_grid_t* nb)
{
  TSD_LOCK(g);
  if(!g || !nb || !g->band || !nb->band || side < 0 || side > 1) return TSD_E_INVALID;
  int rc = check_neighbour(g, side, nb->row_begin, nb->row_end, nb->parts_x);
  if(rc) return rc;
  tsd_grid::Peer& pr = g->peer[side];
  if(pr.connected) { set_error("neighbour already connected"); return TSD_E_INVALID; }
  if(nb->device != g->device)
  {
    TSD_CUDA(cudaSetDevice(g->device));
    int can = 0;
    TSD_CUDA(cudaDeviceCanAccessPeer(&can, g->device, nb->device));
    if(!can) { set_error("no peer access between devices %d and %d", g->device, nb->device); return TSD_E_INVALID; }
    cudaError_t pe = cudaDeviceEnablePeerAccess(nb->device, 0);
    if(pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) TSD_CUDA(pe);
    cudaGetLastError();
  }
  pr.tsd = nb->d_tsd;
  pr.weight = nb->d_weight;
  pr.signal = nb->d_signal;
  pr.alloc_begin = nb->alloc_begin;
  pr.ipc = false;
  pr.connected = true;
  return TSD_OK;
}

int tsdg_band_halo_sync(tsd_grid_t* g, int lo_px0, int lo_px1, int hi_px0, int hi_px1)
{
  TSD_LOCK(g);
  if(!g || !g->band) { set_error("not a sharded grid"); return TSD_E_INVALID; }
  TSD_CUDA(cudaSetDevice(g->device));
  HaloParams hp;
  memset(&hp, 0, sizeof(hp));
  hp.pp = make_params(g);
  const int px0[2] = {lo_px0, hi_px0}, px1[2] = {lo_px1, hi_px1};
  bool any = false;
  for(int b = 0; b < 2; b++)
  {
    const bool has = (b == 0) ? g->row_begin > 0 : g->row_end < g->parts_y;
    if(!has || px1[b] < px0[b]) continue;
    if(px0[b] < 0 || px1[b] >= g->parts_x) { set_error("halo columns out of range"); return TSD_E_RANGE; }
    if(!g->peer[b].connected) { set_error("neighbouring band not connected (tsdg_band_connect)"); return TSD_E_INVALID; }
    any = true;
    hp.active[b] = 1;
    hp.px0[b] = px0[b];
    hp.px1[b] = px1[b];
    const int src_row = (b == 0) ? g->row_begin : g->row_end - 1;
    const size_t src_off = (size_t)(src_row - g->alloc_begin) * g->parts_x * TSD_TILE_STRIDE;
    hp.src_t[b] = g->d_tsd + src_off;
    hp.src_w[b] = g->d_weight + src_off;
    // the same row in the neighbour's allocation: its halo above (b == 0) / below (b == 1) its band
    const size_t dst_off = (size_t)(src_row - g->peer[b].alloc_begin) * g->parts_x * TSD_TILE_STRIDE;
    hp.dst_t[b] = g->peer[b].tsd + dst_off;
    hp.dst_w[b] = g->peer[b].weight + dst_off;
    hp.peer_sig[b] = g->peer[b].signal;
    hp.seq[b] = ++g->halo_seq[b];
  }
  g->band_push_open = false;
  if(!any) return TSD_OK;
  hp.my_sig = g->d_signal;
  // all CTAs must be co-resident (they wait on remote flags): at most one per SM
  int ctas = g->sm_count / 2;
  if(ctas < 1) ctas = 1;
  k_halo_sync<<<ctas, 256, 0, g->stream>>>(hp);
  TSD_LAUNCHED();
  g->halo_used = true;
  return TSD_OK;
}

int tsdg_band_flags(tsd_grid_t* g, uint8_t** flags, uint64_t* count)
{
  TSD_LOCK(g);
  if(!g || !flags || !count) return TSD_E_INVALID;
  *flags = g->d_flags;
  *count = (uint64_t)g->n_parts;
  return TSD_OK;
}

int tsdg_scan_box(const tsd_grid_t* g, const tsd_scan_t* scan, int32_t box[4])
{
  TSD_LOCK(g);
  if(!g || !scan || !box) return TSD_E_INVALID;
  int b[4];
  scan_partition_box(g, scan->pose[2], scan->pose[5], scan->max_range, b);
  for(int i = 0; i < 4; i++) box[i] = b[i];
  return TSD_OK;
}

// which: 0 = my lowest partition row (the band below wants it), 1 = my highest row (the band above wants it),
//        2 = halo slot below my band (filled from the lower neighbour's highest row),
//        3 = halo slot above my band (filled from the upper neighbour's lowest row).
int tsdg_band_row(tsd_grid_t* g, int which, double** tsd, double** weight, uint64_t* count)
{
  TSD_LOCK(g);
  if(!g || !tsd || !weight || !count || which < 0 || which > 3) return TSD_E_INVALID;
  int row;
  if(which == 0) row = g->row_begin;
  else if(which == 1) row = g->row_end - 1;
  else if(which == 2) row = g->row_begin - 1;
  else row = g->row_end;
  *tsd = *weight = nullptr;
  *count = 0;
  if(row < g->alloc_begin || row >= g->alloc_end) return TSD_OK;  // no such neighbour
  const size_t off = (size_t)(row - g->alloc_begin) * g->parts_x * TSD_TILE_STRIDE;
  *tsd = g->d_tsd + off;
  *weight = g->d_weight + off;
  *count = (uint64_t)g->parts_x * TSD_TILE_STRIDE;
  return TSD_OK;
}

// UPDATE_PROFILE builds: cycle counters of the last k_update launch, 8 per CTA: producer {total, waiting for a free
// stage, fetching work, items}, consumer warp 0 {total, waiting for data after the first item, until the first item, SM}
int tsdg_debug_profile(tsd_grid_t* g, unsigned long long* out, int max_ctas)
{
  TSD_LOCK(g);
  if(!g || !out) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  if(!g->d_prof) return TSD_E_INVALID;
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  int n = g->sm_count * UPDATE_CTAS_PER_SM;
  if(n > max_ctas) n = max_ctas;
  TSD_CUDA(cudaMemcpy(out, g->d_prof, sizeof(unsigned long long) * 8 * n, cudaMemcpyDeviceToHost));
  return n;
}

int tsdg_set_update_filter(tsd_grid_t* g, unsigned mask)
{
  TSD_LOCK(g);
  if(!g) return TSD_E_INVALID;
  g->update_filter = mask & 3u;
  return TSD_OK;
}

int tsdg_set_timing(tsd_grid_t* g, int enable)
{
  TSD_LOCK(g);
  if(!g) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  if(enable && !g->ev[0])
    for(int i = 0; i < 4; i++) TSD_CUDA(cudaEventCreate(&g->ev[i]));
  g->timing = enable != 0;
  return TSD_OK;
}

int tsdg_last_push_kernel_ms(tsd_grid_t* g, float ms[4])
{
  TSD_LOCK(g);
  if(!g || !ms || !g->timing || !g->pushed_once) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  TSD_CUDA(cudaEventElapsedTime(&ms[0], g->ev[0], g->ev[1]));
  TSD_CUDA(cudaEventElapsedTime(&ms[1], g->ev[1], g->ev[2]));
  TSD_CUDA(cudaEventElapsedTime(&ms[2], g->ev[2], g->ev[3]));
  TSD_CUDA(cudaEventElapsedTime(&ms[3], g->ev[0], g->ev[3]));
  return TSD_OK;
}

int tsdg_sync(tsd_grid_t* g)
{
  TSD_LOCK(g);
  if(!g) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  if(g->band && g->halo_used)
  {
    // a halo synchronisation whose neighbour never showed up leaves a flag instead of trapping (k_halo_sync)
    uint32_t err = 0;
    TSD_CUDA(cudaMemcpy(&err, g->d_signal + 5, sizeof(err), cudaMemcpyDeviceToHost));
    if(err)
    {
      const uint32_t zero = 0;
      cudaMemcpy(g->d_signal + 5, &zero, sizeof(zero), cudaMemcpyHostToDevice);
      set_error("halo synchronisation timed out: a neighbouring band did not take part (both sides of a boundary must call "
                "tsdg_band_halo_sync equally often)");
      return TSD_E_CUDA;
    }
  }
  return TSD_OK;
}

int tsdg_push(tsd_grid_t* g, const tsd_scan_t* scan)
{
  TSD_LOCK(g);
  int rc = tsdg_push_async(g, scan);
  if(rc) return rc;
  // the blocking call brings the push statistics back with its one synchronisation
  TSD_CUDA(cudaMemcpyAsync(g->h_counters, g->d_counters, sizeof(uint32_t) * 24, cudaMemcpyDeviceToHost, g->stream));
  rc = tsdg_sync(g);
  g->stats_fresh = (rc == TSD_OK);
  return rc;
}

int tsdg_last_push_stats(tsd_grid_t* g, tsd_push_stats_t* out)
{
  TSD_LOCK(g);
  if(!g || !out) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  if(!g->stats_fresh)
  {
    TSD_CUDA(cudaMemcpyAsync(g->h_counters, g->d_counters, sizeof(uint32_t) * 24, cudaMemcpyDeviceToHost, g->stream));
    TSD_CUDA(cudaStreamSynchronize(g->stream));
  }
  out->cell_updates = (uint64_t)g->h_counters[20] | ((uint64_t)g->h_counters[21] << 32);
  out->active_tiles = g->h_counters[8 + 7];
  out->cell_visits = (uint64_t)g->h_counters[8 + 7] * TSD_TILE_CELLS;
  out->emptied_tiles = g->h_counters[8 + 6];
  out->newly_initialized = g->h_counters[8 + 4];
  out->fallback_cells = g->h_counters[8 + 5];
  return TSD_OK;
}

int tsdg_interpolate_bilinear(tsd_grid_t* g, int32_t n, const double* xy, double* tsd, int32_t* status)
{
  TSD_LOCK(g);
  if(!g || n < 0 || (n > 0 && (!xy || !tsd || !status))) return TSD_E_INVALID;
  if(n == 0) return TSD_OK;
  TSD_CUDA(cudaSetDevice(g->device));
  const size_t bytes = (size_t)n * (sizeof(double) * 3 + sizeof(int));
  int rc = grid_ensure_scratch(g, bytes);
  if(rc) return rc;
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  double* h = (double*)g->h_scratch;
  double* d = (double*)g->d_scratch;
  memcpy(h, xy, sizeof(double) * 2 * n);
  TSD_CUDA(cudaMemcpyAsync(d, h, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, g->stream));
  double* d_tsd = d + 2 * (size_t)n;
  int* d_st = (int*)(d_tsd + n);
  k_interpolate<<<(n + 127) / 128, 128, 0, g->stream>>>(grid_view(g), n, d, d_tsd, d_st);
  TSD_LAUNCHED();
  TSD_CUDA(cudaMemcpyAsync(h + 2 * (size_t)n, d_tsd, sizeof(double) * n + sizeof(int) * n, cudaMemcpyDeviceToHost, g->stream));
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  memcpy(tsd, h + 2 * (size_t)n, sizeof(double) * n);
  memcpy(status, h + 3 * (size_t)n, sizeof(int) * n);
  return TSD_OK;
}

int tsdg_interpolate_normal(tsd_grid_t* g, int32_t n, const double* xy, double* normals, int32_t* ok)
{
  TSD_LOCK(g);
  if(!g || n < 0 || (n > 0 && (!xy || !normals || !ok))) return TSD_E_INVALID;
  if(n == 0) return TSD_OK;
  TSD_CUDA(cudaSetDevice(g->device));
  const size_t bytes = (size_t)n * (sizeof(double) * 4 + sizeof(int));
  int rc = grid_ensure_scratch(g, bytes);
  if(rc) return rc;
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  double* h = (double*)g->h_scratch;
  double* d = (double*)g->d_scratch;
  memcpy(h, xy, sizeof(double) * 2 * n);
  TSD_CUDA(cudaMemcpyAsync(d, h, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, g->stream));
  double* d_n = d + 2 * (size_t)n;
  int* d_ok = (int*)(d_n + 2 * (size_t)n);
  k_interpolate_normal<<<(n + 127) / 128, 128, 0, g->stream>>>(grid_view(g), n, d, d_n, d_ok);
  TSD_LAUNCHED();
  TSD_CUDA(cudaMemcpyAsync(h + 2 * (size_t)n, d_n, sizeof(double) * 2 * n + sizeof(int) * n, cudaMemcpyDeviceToHost, g->stream));
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  memcpy(normals, h + 2 * (size_t)n, sizeof(double) * 2 * n);
  memcpy(ok, h + 4 * (size_t)n, sizeof(int) * n);
  return TSD_OK;
}

int tsdg_num_partitions(const tsd_grid_t* g, int32_t* n)
{
  TSD_LOCK(g);
  if(!g || !n) return TSD_E_INVALID;
  *n = g->n_parts;
  return TSD_OK;
}

int tsdg_partition_states(tsd_grid_t* g, int32_t* state, double* init_weight)
{
  TSD_LOCK(g);
  if(!g || !state) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  std::vector<uint8_t> flags(g->n_parts);
  std::vector<double> iw(g->n_parts);
  TSD_CUDA(cudaMemcpy(flags.data(), g->d_flags, g->n_parts, cudaMemcpyDeviceToHost));
  TSD_CUDA(cudaMemcpy(iw.data(), g->d_initw, sizeof(double) * g->n_parts, cudaMemcpyDeviceToHost));
  for(int p = 0; p < g->n_parts; p++)
  {
    // TsdGridPartition.h:66,72 isInitialized / isEmpty
    state[p] = flags[p] ? TSD_PARTITION_CONTENT : (iw[p] > 0.0 ? TSD_PARTITION_EMPTY : TSD_PARTITION_UNINITIALIZED);
    if(init_weight) init_weight[p] = iw[p];
  }
  return TSD_OK;
}

static void tile_to_33(const double* tile, double* out33)
{
  for(int y = 0; y < 32; y++)
  {
    for(int x = 0; x < 32; x++) out33[y * 33 + x] = tile[y * 32 + x];
    out33[y * 33 + 32] = tile[TSD_BORDER_OFF + y];
  }
  for(int x = 0; x < 32; x++) out33[32 * 33 + x] = tile[TSD_BORDER_OFF + 32 + x];
  out33[32 * 33 + 32] = tile[TSD_BORDER_OFF + 64];
}

static void tile_from_33(const double* in33, double* tile)
{
  for(int y = 0; y < 32; y++)
  {
    for(int x = 0; x < 32; x++) tile[y * 32 + x] = in33[y * 33 + x];
    tile[TSD_BORDER_OFF + y] = in33[y * 33 + 32];
  }
  for(int x = 0; x < 32; x++) tile[TSD_BORDER_OFF + 32 + x] = in33[32 * 33 + x];
  tile[TSD_BORDER_OFF + 64] = in33[32 * 33 + 32];
  for(int i = TSD_BORDER_OFF + 65; i < TSD_TILE_STRIDE; i++) tile[i] = 0.0;
}

int tsdg_download_partition(tsd_grid_t* g, int32_t p, double* tsd, double* weight)
{
  TSD_LOCK(g);
  if(!g || p < 0 || p >= g->n_parts || !tsd || !weight) return TSD_E_INVALID;
  const int py = p / g->parts_x;
  if(py < g->row_begin || py >= g->row_end) { set_error("partition %d is not owned by this band", p); return TSD_E_INVALID; }
  TSD_CUDA(cudaSetDevice(g->device));
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  uint8_t flag = 0;
  TSD_CUDA(cudaMemcpy(&flag, g->d_flags + p, 1, cudaMemcpyDeviceToHost));
  if(!flag) return TSD_E_INVALID;
  double tile[TSD_TILE_STRIDE];
  const size_t base = (size_t)(p - g->alloc_begin * g->parts_x) * TSD_TILE_STRIDE;
  TSD_CUDA(cudaMemcpy(tile, g->d_tsd + base, sizeof(tile), cudaMemcpyDeviceToHost));
  tile_to_33(tile, tsd);
  TSD_CUDA(cudaMemcpy(tile, g->d_weight + base, sizeof(tile), cudaMemcpyDeviceToHost));
  tile_to_33(tile, weight);
  return TSD_OK;
}

int tsdg_upload_partition(tsd_grid_t* g, int32_t p, const double* tsd, const double* weight)
{
  TSD_LOCK(g);
  if(!g || p < 0 || p >= g->n_parts || !tsd || !weight) return TSD_E_INVALID;
  const int py = p / g->parts_x;
  if(py < g->row_begin || py >= g->row_end) { set_error("partition %d is not owned by this band", p); return TSD_E_INVALID; }
  TSD_CUDA(cudaSetDevice(g->device));
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  double tile[TSD_TILE_STRIDE];
  const size_t base = (size_t)(p - g->alloc_begin * g->parts_x) * TSD_TILE_STRIDE;
  tile_from_33(tsd, tile);
  TSD_CUDA(cudaMemcpy(g->d_tsd + base, tile, sizeof(tile), cudaMemcpyHostToDevice));
  tile_from_33(weight, tile);
  TSD_CUDA(cudaMemcpy(g->d_weight + base, tile, sizeof(tile), cudaMemcpyHostToDevice));
  const uint8_t one = 1;
  TSD_CUDA(cudaMemcpy(g->d_flags + p, &one, 1, cudaMemcpyHostToDevice));
  const uint32_t all = 1;  // the next push refreshes every border, like the reference's full propagateBorders
  TSD_CUDA(cudaMemcpy(g->d_counters + 3, &all, sizeof(all), cudaMemcpyHostToDevice));
  g->refresh_all_pending = true;
  return TSD_OK;
}

int tsdg_fill(tsd_grid_t* g, double tsd, double weight, int only_uninitialized)
{
  TSD_LOCK(g);
  if(!g) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  PushParams pp = make_params(g);
  k_fill<<<g->sm_count * 8, 256, 0, g->stream>>>(pp, tsd, weight, only_uninitialized);
  TSD_LAUNCHED();
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  const uint32_t all = 1;
  TSD_CUDA(cudaMemcpy(g->d_counters + 3, &all, sizeof(all), cudaMemcpyHostToDevice));
  g->refresh_all_pending = true;
  return TSD_OK;
}

}  // extern "C"
