// RayCastPolar2D on the device (K5): one warp per beam.
//
// Reference: src/obvision/reconstruct/grid/RayCastPolar2D.cpp:113-192 (calcCoordsFromCurrentViewMask),
// :27-111 (calcCoordsFromCurrentView), :194-281 (rayCastFromCurrentView).
//
// The reference marches a beam serially: `position += ray` once per step and `i += 1.0` for the loop bound
// (`i += 32.0` in the coarse partition-skipping loop).  Those running sums are NOT tr + k*ray in floating
// point.  The positions of a 32-step pass are obtained in closed form where that is exact (same binade, no
// rounding tie: see `advance` below) and by replaying the serial additions on every lane where it is not; what
// is distributed over the lanes is the expensive part, the bilinear sample (5 loads + ~40 FP64 ops), with the
// samples of RC_DEPTH passes in flight.  The first +/- sign change (hit) or -/+ sign change (abort) of a pass is
// found with two ballots.  On a sharded grid the kernel's epilogue is the all-gather of the per-beam first
// events (stores into every band's exchange block over peer memory, k_raycast_merge picks the earliest).
#include <string.h>

#include "common.cuh"
#include "closed_form.cuh"

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
  // Sharded grid, exchange over peer memory (tsdg_raycast_mask_sharded): rcx_world > 0 makes every beam's first event
  // go into slot rcx_rank of EVERY band's exchange block (NVLink P2P stores from the marching kernel itself: the
  // all-gather is part of the kernel that computes the data), followed by a release signal to every band.
  int rcx_world, rcx_rank;
  uint32_t rcx_seq;
  unsigned long long* rcx_keys[TSD_RCX_MAX];  // band r's key array of the current parity: [slot][TSD_RCX_CAP]
  double* rcx_out[TSD_RCX_MAX];               // band r's payload array: [slot][TSD_RCX_CAP][4]
  uint32_t* rcx_sig[TSD_RCX_MAX];             // band r's signal words: [slot]
  uint32_t* rcx_ticket;                       // this band's CTA ticket
};

#define RC_WARPS 4
#ifndef RC_PREFETCH_PASSES
#define RC_PREFETCH_PASSES 6
#endif
#ifndef RC_DEPTH
#define RC_DEPTH 2   // PAIRS of passes of a ray whose samples are in flight at once (single passes, C3, ms per call: 1 -> 0.41,
                     // 2 -> 0.30, 3 and 4 -> 0.29)
#endif

// L2 prefetch of the two cell rows a bilinear sample at (cx, cy) reads (clamped like sample_issue: always a valid address)
__device__ __forceinline__ void sample_prefetch(const GridView& g, double cx, double cy)
{
  int xIdx = __double2int_rd(cx * g.inv_cell_size);
  int yIdx = __double2int_rd(cy * g.inv_cell_size);
  xIdx = min(max(xIdx, 0), g.cells_x - 1);
  yIdx = min(max(yIdx, g.alloc_begin * 32), g.alloc_end * 32 - 1);
  const int py = yIdx >> 5, px = xIdx >> 5;
  const double* t = g.tsd + (size_t)((py - g.alloc_begin) * g.parts_x + px) * TSD_TILE_STRIDE + (yIdx & 31) * 32 + (xIdx & 31);
  asm volatile("prefetch.global.L2 [%0];" ::"l"(t));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(t + 32));
}
#define NO_EVENT 0x7fffffffffffffffULL  // INT64_MAX: the largest key under a signed or unsigned min-reduction

__device__ __forceinline__ void raycast_beam(const RayParams& rp, const int beam, const int lane)
{
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
    // own step.  The samples of TWO passes are in flight: while pass c is evaluated, the loads of pass c + 1 are on
    // their way and the chain of pass c + 2 is being added up, so that a pass costs its arithmetic, not a trip to L2 or
    // HBM (a ray of the 250 m sensor is 300 passes one after the other: the longest ray is the kernel's duration).
    // Loads past the end of a ray are harmless (addresses are clamped to valid memory, results masked by `valid`).
    // The loop counter `i += 1.0` of the reference only decides when the loop ends: it is replayed exactly
    // (serially) only for passes that come within 2 steps of idxMax; elsewhere idxMin + k decides safely.
    // (every lane runs the whole chain and keeps the position of its own step in registers: parking the positions
    //  in shared memory made each addition wait for the previous store to read its operands)
    double iExact = idxMin;          // i of iteration kExact (exact serial value)
    unsigned long long kExact = 0;
    unsigned long long base = 0;
    // The serial sum `position += ray` of a pass in closed form.  While a coordinate p stays inside one binade
    // [2^e, 2^(e+1)) its values are multiples of u = ulp(p), and fl(p + r) = p + d with d = RN_u(r) the same multiple of u
    // at every step -- unless r lies exactly half-way between two multiples (then round-to-even decides by p's parity).
    // d is read off one real addition (d = fl(p + r) - p, exact), and p + k d is exact for k <= 32 (k d has fewer than
    // 53 bits; the sum is a multiple of u inside the binade).  So if start and end of the pass share sign and exponent
    // (the sequence is monotonic) and there is no tie, lane k's position is p + (k + 1) d: 2 operations instead of a
    // chain of 32 dependent additions that every lane replays.  Otherwise (a coordinate crosses a power of two: about ten
    // passes per ray) the pass falls back to the serial chain.
    auto advance = [&](double* ox, double* oy)
    {
      double d0, d1, e0, e1;
      const bool closed = tsd_closed_form_pass(pos0, ray0, &d0, &e0) & tsd_closed_form_pass(pos1, ray1, &d1, &e1);
      if(closed)
      {
        const double k = (double)(lane + 1);
        *ox = pos0 + k * d0;
        *oy = pos1 + k * d1;
        pos0 = e0;
        pos1 = e1;
      }
      else
      {
        double ax = 0.0, ay = 0.0;
#pragma unroll 8
        for(int k = 0; k < 32; k++)
        {
          pos0 += ray0;
          pos1 += ray1;
          if(k == lane) { ax = pos0; ay = pos1; }
        }
        *ox = ax;
        *oy = ay;
      }
    };
    // validity of this lane's iteration in the pass that starts at step b: i_k <= idxMax
    auto valid_of = [&](unsigned long long b) -> bool
    {
      const bool nearEnd = !(idxMin + (double)(b + 31u) + 2.0 < idxMax);
      if(!nearEnd) return true;
      double mi = 0.0;
      // replay the reference's counter `i += 1.0` up to this pass (uniform), keep this lane's value.  Whole passes in
      // closed form where that is exact (closed_form.cuh with r = 1: a ray that reaches the end of its range replayed
      // 10 000 serial additions here, 40 us of the longest rays' 220), the rest one by one.
      while(kExact + 32u <= b)
      {
        double dI, eI;
        if(tsd_closed_form_pass(iExact, 1.0, &dI, &eI)) iExact = eI;
        else
          for(int k = 0; k < 32; k++) iExact += 1.0;
        kExact += 32u;
      }
      while(kExact < b) { iExact += 1.0; kExact++; }
      double ii = iExact;
#pragma unroll 8
      for(int k = 0; k < 32; k++)
      {
        if(k == lane) mi = ii;
        ii += 1.0;
      }
      return mi <= idxMax;
    };
    // the first event of a pass, if any; true when the ray is finished
    auto events = [&](const SampleLoads& cur, double cmx, double cmy, bool valid, double v, double prev) -> bool
    {
      // a step belongs to the band that owns its sample's partition (everything, for an unsharded grid)
      const bool mine = valid && (cur.py >= g.row_begin) && (cur.py < g.row_end);
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
            cx = cmx + ray0 * (interp - 1.0);
            cy = cmy + ray1 * (interp - 1.0);
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
        return true;
      }
      if(mInval)
      {
        nFine += (unsigned)(__ffs(mInval) - 1);
        return true;
      }
      nFine += 32;
      base += 32;
      return false;
    };
    // TWO consecutive passes at a time (64 steps): their samples are finished side by side -- two independent dependency
    // chains per lane, where one pass alone leaves the warp waiting on its own previous instruction most of the time --,
    // then their events are looked at in order, then both sets of registers are refilled with the samples of the pair
    // RC_DEPTH pairs ahead.  Returns true when the ray is finished.  The sets are only ever refilled in place, never moved
    // (a move of a loaded value waits for its load).
    auto pass2 = [&](SampleLoads& A, double& ax, double& ay, SampleLoads& B, double& bx, double& by) -> bool
    {
      const bool validA = valid_of(base), validB = valid_of(base + 32u);
      // positions of the pair RC_DEPTH pairs ahead while the loads of the pairs in between are in flight
      double nax, nay, nbx, nby;
      advance(&nax, &nay);
      advance(&nbx, &nby);
      const double nanv = __longlong_as_double(0x7ff8000000000000LL);
      double tA = 0.0, tB = 0.0;
      const int rvA = sample_finish(A, &tA);
      const int rvB = sample_finish(B, &tB);
      const double vA = (validA && rvA == TSD_INTERPOLATE_SUCCESS) ? tA : nanv;
      const double vB = (validB && rvB == TSD_INTERPOLATE_SUCCESS) ? tB : nanv;
      double prevA = __shfl_up_sync(0xffffffffu, vA, 1);
      double prevB = __shfl_up_sync(0xffffffffu, vB, 1);
      const double lastA = __shfl_sync(0xffffffffu, vA, 31);
      if(lane == 0) { prevA = carry; prevB = lastA; }
      if(events(A, ax, ay, validA, vA, prevA)) return true;
      if(events(B, bx, by, validB, vB, prevB)) return true;
      carry = __shfl_sync(0xffffffffu, vB, 31);
      ax = nax; ay = nay;
      A = sample_issue(g, nax, nay);
      bx = nbx; by = nby;
      B = sample_issue(g, nbx, nby);
      // ... and the cells the ray reaches RC_PREFETCH_PASSES passes later are called into L2 (approximate positions are
      // good enough for that): a ray walks through memory it has never touched -- a new partition every 32 steps, a new
      // 2 MB page every partition row.
      sample_prefetch(g, nax + (32.0 * RC_PREFETCH_PASSES) * ray0, nay + (32.0 * RC_PREFETCH_PASSES) * ray1);
      sample_prefetch(g, nbx + (32.0 * RC_PREFETCH_PASSES) * ray0, nby + (32.0 * RC_PREFETCH_PASSES) * ray1);
      return false;
    };
    // RC_DEPTH pairs of passes in flight, each with its own registers
    double pmx[2 * RC_DEPTH], pmy[2 * RC_DEPTH];
    SampleLoads psl[2 * RC_DEPTH];
#pragma unroll
    for(int d = 0; d < 2 * RC_DEPTH; d++)
    {
      advance(&pmx[d], &pmy[d]);
      psl[d] = sample_issue(g, pmx[d], pmy[d]);
    }
    bool done = false;
    while(!done)
    {
#pragma unroll
      for(int d = 0; d < RC_DEPTH; d++)
        if(!done) done = pass2(psl[2 * d], pmx[2 * d], pmy[2 * d], psl[2 * d + 1], pmx[2 * d + 1], pmy[2 * d + 1]);
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
    if(rp.rcx_world > 0)
    {
      const double o0 = rp.out[4 * beam + 0], o1 = rp.out[4 * beam + 1], o2 = rp.out[4 * beam + 2], o3 = rp.out[4 * beam + 3];
      const size_t slot = (size_t)rp.rcx_rank * TSD_RCX_CAP + (size_t)beam;
      for(int r = 0; r < rp.rcx_world; r++)
      {
        rp.rcx_keys[r][slot] = key;
        *reinterpret_cast<double2*>(rp.rcx_out[r] + 4 * slot) = make_double2(o0, o1);
        *reinterpret_cast<double2*>(rp.rcx_out[r] + 4 * slot + 2) = make_double2(o2, o3);
      }
    }
  }
}

__global__ void __launch_bounds__(RC_WARPS * 32) k_raycast(RayParams rp)
{
  const int lane = threadIdx.x & 31;
  const int beam = blockIdx.x * RC_WARPS + (threadIdx.x >> 5);
  if(beam < rp.scan.n) raycast_beam(rp, beam, lane);
  if(rp.rcx_world > 0)
  {
    // the last CTA to finish tells every band (itself included) that this band's slot is complete
    __threadfence_system();
    __syncthreads();
    if(threadIdx.x == 0)
    {
      if(atomicAdd(rp.rcx_ticket, 1u) == gridDim.x - 1)
      {
        *rp.rcx_ticket = 0;
        __threadfence_system();
        for(int r = 0; r < rp.rcx_world; r++) st_release_sys(rp.rcx_sig[r] + rp.rcx_rank, rp.rcx_seq);
      }
    }
  }
}

// Second half of the sharded ray cast: wait until every band's slot of this band's exchange block is complete, then
// every beam takes the earliest event among the bands (keys are 4 * step + code: the reference's serial march stops at
// the first event) and the payload of the band that saw it.  One launch; the result lands where the unsharded ray
// caster leaves its own (keys + {cx cy nx ny}), so the host side is shared.
struct MergeParams
{
  int n, world;
  uint32_t seq;
  const uint32_t* sig;   // this band's signal words [slot]
  uint32_t* err;
  const unsigned long long* keys_in;  // [slot][TSD_RCX_CAP]
  const double* out_in;               // [slot][TSD_RCX_CAP][4]
  unsigned long long* keys;
  double* out;
};

__global__ void __launch_bounds__(256) k_raycast_merge(MergeParams mp)
{
  if(threadIdx.x < mp.world) wait_seq(mp.sig + threadIdx.x, mp.seq, mp.err);
  __syncthreads();
  const int beam = blockIdx.x * blockDim.x + threadIdx.x;
  if(beam >= mp.n) return;
  unsigned long long best = NO_EVENT;
  int br = 0;
  for(int r = 0; r < mp.world; r++)
  {
    const unsigned long long k = mp.keys_in[(size_t)r * TSD_RCX_CAP + beam];
    if(k < best) { best = k; br = r; }
  }
  const double* src = mp.out_in + 4 * ((size_t)br * TSD_RCX_CAP + beam);
  const bool any = best != NO_EVENT;
  mp.keys[beam] = best;
  mp.out[4 * beam + 0] = any ? src[0] : 0.0;
  mp.out[4 * beam + 1] = any ? src[1] : 0.0;
  mp.out[4 * beam + 2] = any ? src[2] : 0.0;
  mp.out[4 * beam + 3] = any ? src[3] : 0.0;
}

#define RC_IS_HIT(k) ((k) != NO_EVENT && ((k) & 3ULL) == 0ULL)

// layout of a band's exchange block
#define RCX_KEYS_BYTES ((size_t)2 * TSD_RCX_MAX * TSD_RCX_CAP * sizeof(unsigned long long))
#define RCX_OUT_BYTES ((size_t)2 * TSD_RCX_MAX * TSD_RCX_CAP * 4 * sizeof(double))
#define RCX_SIG_OFF (RCX_KEYS_BYTES + RCX_OUT_BYTES)   // uint32: [TSD_RCX_MAX] signals, [TSD_RCX_MAX] ticket, [+1] timeout flag
#define RCX_BYTES (RCX_SIG_OFF + (TSD_RCX_MAX + 8) * sizeof(uint32_t))

static int raycast_exchange_finish(tsd_grid_t* g)
{
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  uint32_t err = 0;
  TSD_CUDA(cudaMemcpy(&err, g->d_rcx + RCX_SIG_OFF + (TSD_RCX_MAX + 1) * sizeof(uint32_t), sizeof(err), cudaMemcpyDeviceToHost));
  if(err)
  {
    cudaMemset(g->d_rcx + RCX_SIG_OFF + (TSD_RCX_MAX + 1) * sizeof(uint32_t), 0, sizeof(uint32_t));
    set_error("sharded ray cast timed out: not every band took part (the sharded ray cast is a collective)");
    return TSD_E_CUDA;
  }
  return TSD_OK;
}

static int raycast_launch(tsd_grid_t* g, const tsd_scan_t* scan, const double* rays_world, bool download = true, bool exchange = false,
                          bool sync = true)
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
  if(exchange)
  {
    if(g->rcx_world < 1 || !g->d_rcx) { set_error("ray-cast exchange not connected (tsdg_band_rcx_connect)"); return TSD_E_INVALID; }
    if(n > TSD_RCX_CAP) { set_error("scan too large for the ray-cast exchange block (%d beams)", TSD_RCX_CAP); return TSD_E_INVALID; }
    const uint32_t seq = ++g->rcx_seq;
    const size_t par = seq & 1u;
    rp.rcx_world = g->rcx_world;
    rp.rcx_rank = g->rcx_rank;
    rp.rcx_seq = seq;
    for(int r = 0; r < g->rcx_world; r++)
    {
      unsigned char* blk = (r == g->rcx_rank) ? g->d_rcx : g->peer_rcx[r];
      rp.rcx_keys[r] = reinterpret_cast<unsigned long long*>(blk) + par * TSD_RCX_MAX * TSD_RCX_CAP;
      rp.rcx_out[r] = reinterpret_cast<double*>(blk + RCX_KEYS_BYTES) + par * TSD_RCX_MAX * TSD_RCX_CAP * 4;
      rp.rcx_sig[r] = reinterpret_cast<uint32_t*>(blk + RCX_SIG_OFF);
    }
    rp.rcx_ticket = reinterpret_cast<uint32_t*>(g->d_rcx + RCX_SIG_OFF) + TSD_RCX_MAX;
  }
  k_raycast<<<(n + RC_WARPS - 1) / RC_WARPS, RC_WARPS * 32, 0, g->stream>>>(rp);
  TSD_LAUNCHED();
  if(exchange)
  {
    MergeParams mp;
    mp.n = n;
    mp.world = g->rcx_world;
    mp.seq = g->rcx_seq;
    mp.sig = reinterpret_cast<const uint32_t*>(g->d_rcx + RCX_SIG_OFF);
    mp.err = reinterpret_cast<uint32_t*>(g->d_rcx + RCX_SIG_OFF) + TSD_RCX_MAX + 1;
    const size_t par = g->rcx_seq & 1u;
    mp.keys_in = reinterpret_cast<const unsigned long long*>(g->d_rcx) + par * TSD_RCX_MAX * TSD_RCX_CAP;
    mp.out_in = reinterpret_cast<const double*>(g->d_rcx + RCX_KEYS_BYTES) + par * TSD_RCX_MAX * TSD_RCX_CAP * 4;
    mp.keys = g->d_rc_keys;
    mp.out = g->d_rc_out;
    k_raycast_merge<<<(n + 255) / 256, 256, 0, g->stream>>>(mp);
    TSD_LAUNCHED();
  }
  if(download) TSD_CUDA(cudaMemcpyAsync(g->h_rc, g->d_rc, g->rc_bytes, cudaMemcpyDeviceToHost, g->stream));
  if(!sync) return TSD_OK;
  if(exchange) return raycast_exchange_finish(g);
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  return TSD_OK;
}

// the ray cast enqueued on the grid's stream, results left on the device (tsdg_localize, icp.cu)
int tsd_raycast_enqueue(tsd_grid_t* g, const tsd_scan_t* scan, const double* rays_world)
{
  return raycast_launch(g, scan, rays_world, false, false, false);
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

// --- sharded ray cast over peer memory --------------------------------------------------------------------------
struct RcxExport
{
  cudaIpcMemHandle_t block;
  int32_t row_begin, row_end, parts_x;
};
static_assert(sizeof(RcxExport) <= TSD_BAND_EXPORT_BYTES, "tsd_band_export_t too small");

static int rcx_ensure(tsd_grid_t* g)
{
  if(g->d_rcx) return TSD_OK;
  TSD_CUDA(cudaSetDevice(g->device));
  TSD_CUDA(cudaMalloc(&g->d_rcx, RCX_BYTES));
  TSD_CUDA(cudaMemset(g->d_rcx, 0, RCX_BYTES));
  return TSD_OK;
}

int tsdg_band_rcx_export(tsd_grid_t* g, void* blob)
{
  TSD_LOCK(g);
  if(!g || !blob || !g->band) { set_error("not a sharded grid"); return TSD_E_INVALID; }
  int rc = rcx_ensure(g);
  if(rc) return rc;
  RcxExport e;
  memset(&e, 0, sizeof(e));
  TSD_CUDA(cudaIpcGetMemHandle(&e.block, g->d_rcx));
  e.row_begin = g->row_begin;
  e.row_end = g->row_end;
  e.parts_x = g->parts_x;
  memset(blob, 0, TSD_BAND_EXPORT_BYTES);
  memcpy(blob, &e, sizeof(e));
  return TSD_OK;
}

int tsdg_band_rcx_connect(tsd_grid_t* g, int rank, int world, const void* blobs)
{
  TSD_LOCK(g);
  if(!g || !g->band || !blobs || world < 1 || world > TSD_RCX_MAX || rank < 0 || rank >= world) return TSD_E_INVALID;
  int rc = rcx_ensure(g);
  if(rc) return rc;
  TSD_CUDA(cudaSetDevice(g->device));
  for(int r = 0; r < world; r++)
  {
    if(r == rank) continue;
    RcxExport e;
    memcpy(&e, (const unsigned char*)blobs + (size_t)r * TSD_BAND_EXPORT_BYTES, sizeof(e));
    if(e.parts_x != g->parts_x) { set_error("band %d has another grid geometry", r); return TSD_E_INVALID; }
    void* p = nullptr;
    TSD_CUDA(cudaIpcOpenMemHandle(&p, e.block, cudaIpcMemLazyEnablePeerAccess));
    g->peer_rcx[r] = static_cast<unsigned char*>(p);
    g->peer_rcx_ipc[r] = true;
  }
  g->rcx_rank = rank;
  g->rcx_world = world;
  return TSD_OK;
}

int tsdg_band_rcx_connect_local(tsd_grid_t* g, int rank, int world, tsd_grid_t** bands)
{
  TSD_LOCK(g);
  if(!g || !g->band || !bands || world < 1 || world > TSD_RCX_MAX || rank < 0 || rank >= world) return TSD_E_INVALID;
  int rc = rcx_ensure(g);
  if(rc) return rc;
  for(int r = 0; r < world; r++)
  {
    if(r == rank) continue;
    if(!bands[r] || !bands[r]->band) return TSD_E_INVALID;
    rc = rcx_ensure(bands[r]);
    if(rc) return rc;
    if(bands[r]->device != g->device)
    {
      TSD_CUDA(cudaSetDevice(g->device));
      int can = 0;
      TSD_CUDA(cudaDeviceCanAccessPeer(&can, g->device, bands[r]->device));
      if(!can) { set_error("no peer access between devices %d and %d", g->device, bands[r]->device); return TSD_E_INVALID; }
      cudaError_t pe = cudaDeviceEnablePeerAccess(bands[r]->device, 0);
      if(pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) TSD_CUDA(pe);
      cudaGetLastError();
    }
    g->peer_rcx[r] = bands[r]->d_rcx;
    g->peer_rcx_ipc[r] = false;
  }
  g->rcx_rank = rank;
  g->rcx_world = world;
  return TSD_OK;
}

static int raycast_copy_out(tsd_grid_t* g, int n, double* coords, double* normals, uint8_t* mask, uint32_t* count)
{
  uint32_t cnt = 0;
  for(int b = 0; b < n; b++)
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
      mask[b] = 0;
  }
  if(count) *count = cnt;
  return TSD_OK;
}

int tsdg_raycast_mask_sharded(tsd_grid_t* g, const tsd_scan_t* scan, const double* rays_world, double* coords, double* normals,
                              uint8_t* mask, uint32_t* count)
{
  TSD_LOCK(g);
  if(!coords || !normals || !mask) return TSD_E_INVALID;
  int rc = raycast_launch(g, scan, rays_world, true, true);
  if(rc) return rc;
  return raycast_copy_out(g, scan->n, coords, normals, mask, count);
}

/* the two halves of tsdg_raycast_mask_sharded, for bands that live in ONE process: launch on every band, then collect */
int tsdg_raycast_sharded_launch(tsd_grid_t* g, const tsd_scan_t* scan, const double* rays_world)
{
  TSD_LOCK(g);
  return raycast_launch(g, scan, rays_world, true, true, false);
}

int tsdg_raycast_sharded_collect(tsd_grid_t* g, int32_t n, double* coords, double* normals, uint8_t* mask, uint32_t* count)
{
  TSD_LOCK(g);
  if(!g || !coords || !normals || !mask || n < 1 || n > g->scan_cap) return TSD_E_INVALID;
  int rc = raycast_exchange_finish(g);
  if(rc) return rc;
  return raycast_copy_out(g, n, coords, normals, mask, count);
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
