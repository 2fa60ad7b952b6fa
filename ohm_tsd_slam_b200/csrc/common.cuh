// Shared host/device definitions of libtsdslam_b200 (sm_100a only; built with -fmad=false so that no
// a*b+c is contracted: the reference is built for baseline x86-64, every product and sum rounds on its own,
// SURVEY.md App. A.1).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>

#include "../../include/tsdslam_b200.h"
#include "beam_index.cuh"

#define TSD_TILE 32               // cells per partition edge (LAYOUT_32x32, SlamNode.cpp:77)
#define TSD_TILE_CELLS 1024
#define TSD_TILE_STRIDE 1104      // doubles per partition and array: 1024 interior + 65 border + 15 pad (8832 B = 69 * 128)
#define TSD_BORDER_OFF 1024       // [0,32): column x=32, rows y=0..31 ; [32,64): row y=32, columns 0..31 ; 64: corner
#define TSD_MAXWEIGHT 32.0        // reconstruct_defs.h:4
#define TSD_NOT_OWNED 4           // sample status: partition belongs to another band (sharded grid only)
#define TSD_RCX_MAX 16            // bands of a sharded grid that can exchange ray-cast events over peer memory
#define TSD_RCX_CAP 2048          // beams per scan in that exchange

namespace tsd
{

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define TSD_CUDA(call)                                                                                   \
  do                                                                                                     \
  {                                                                                                      \
    cudaError_t e__ = (call);                                                                            \
    if(e__ != cudaSuccess)                                                                               \
    {                                                                                                    \
      tsd::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__));      \
      return (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver) ? TSD_E_NO_DEVICE : TSD_E_CUDA; \
    }                                                                                                    \
  } while(0)

#define TSD_LAUNCHED()                                                                                   \
  do                                                                                                     \
  {                                                                                                      \
    tsd::g_launches.fetch_add(1, std::memory_order_relaxed);                                             \
    TSD_CUDA(cudaGetLastError());                                                                        \
  } while(0)

// mathbase.h:39-53 -- note the NaN asymmetry: ob_min(a,b) returns b when a is NaN
__host__ __device__ __forceinline__ double ob_min(double a, double b) { return (a <= b) ? a : b; }
__host__ __device__ __forceinline__ double ob_max(double a, double b) { return (a >= b) ? a : b; }

// Read-only view of the cell state that the samplers need (raycast, interpolate, TSD matcher).
struct GridView
{
  const double* tsd;     // owned partitions, TSD_TILE_STRIDE doubles each
  const uint8_t* flags;  // all partitions of the grid: 1 = initialised
  int cells_x, cells_y;
  int parts_x, parts_y;
  int row_begin, row_end;      // owned partition rows
  int alloc_begin, alloc_end;  // readable partition rows: owned + one halo row on each side of a band
  double cell_size, inv_cell_size;
};

// The POD part of tsd_scan_t plus device pointers to the staged measurement.
struct ScanDev
{
  const double* ranges;
  const uint8_t* mask;
  int n;
  double P[9];     // pose
  double Pi[9];    // pose inverse
  double max_range, min_range, low_refl;
  BeamModel bm;
  // single-precision front end of k_update (grid.cu "fast path"): sensor position as the pose INVERSE implies it
  // (Pi * [txp typ 1]' = [0 0 1]'), and the scaled-angle parameters s = fma(phi, rinv_f, off_f), a cell's beam is
  // round(s) for certain iff |s - round(s)| < half_m (half_m < 0: never, every cell takes the exact route)
  double txp, typ;
  float rinv_f, off_f, half_m;
};

#ifdef __CUDACC__

// TsdGrid::coord2Cell + interpolateBilinear (TsdGrid.h:284-340) + TsdGridPartition::interpolateBilinear
// (TsdGridPartition.h:214-221).  The replicated border of a partition lives behind its 1024 interior cells.
__device__ __forceinline__ int sample_bilinear(const GridView& g, double cx, double cy, double* out)
{
  const double dCoordX = cx * g.inv_cell_size;
  const double dCoordY = cy * g.inv_cell_size;
  int xIdx = __double2int_rd(dCoordX);
  int yIdx = __double2int_rd(dCoordY);
  double dx = ((double)xIdx + 0.5) * g.cell_size;
  double dy = ((double)yIdx + 0.5) * g.cell_size;
  if(cx < dx) { xIdx--; dx -= g.cell_size; }
  if(cy < dy) { yIdx--; dy -= g.cell_size; }
  if((xIdx >= g.cells_x) || (xIdx < 0) || (yIdx >= g.cells_y) || (yIdx < 0)) return TSD_INTERPOLATE_INVALIDINDEX;
  const int py = yIdx >> 5, px = xIdx >> 5;
  const int p = py * g.parts_x + px;
  const int x = xIdx & 31, y = yIdx & 31;
  if(!g.flags[p]) return TSD_INTERPOLATE_EMPTYPARTITION;
  if(py < g.alloc_begin || py >= g.alloc_end) return TSD_NOT_OWNED;
  const double wx = fabs((cx - dx) * g.inv_cell_size);
  const double wy = fabs((cy - dy) * g.inv_cell_size);
  const double* t = g.tsd + (size_t)(p - g.alloc_begin * g.parts_x) * TSD_TILE_STRIDE;
  const int i00 = y * 32 + x;
  const int i10 = (y == 31) ? (TSD_BORDER_OFF + 32 + x) : (i00 + 32);                 // [y+1][x]
  const int i01 = (x == 31) ? (TSD_BORDER_OFF + y) : (i00 + 1);                       // [y][x+1]
  const int i11 = (x == 31) ? ((y == 31) ? (TSD_BORDER_OFF + 64) : (TSD_BORDER_OFF + y + 1))
                            : ((y == 31) ? (TSD_BORDER_OFF + 32 + x + 1) : (i00 + 33)); // [y+1][x+1]
  const double g00 = __ldg(t + i00), g10 = __ldg(t + i10), g01 = __ldg(t + i01), g11 = __ldg(t + i11);
  const double v = g00 * (1. - wy) * (1. - wx) + g10 * wy * (1. - wx) + g01 * (1. - wy) * wx + g11 * wy * wx;
  *out = v;
  if(isnan(v)) return TSD_INTERPOLATE_ISNAN;
  return TSD_INTERPOLATE_SUCCESS;
}

// sample_bilinear split in two so that a caller can put independent work between the loads and their first
// use: sample_issue() computes the cell, issues the flag + 4 cell loads from clamped, always-valid addresses;
// sample_finish() forms the same sum and status as sample_bilinear.
struct SampleLoads
{
  double g00, g10, g01, g11, wx, wy;
  int pre;  // 0, or the status already known from the geometry (INVALIDINDEX / NOT_OWNED)
  int py;   // partition row of the sample (decides which band owns the step)
  unsigned char flag;
};

__device__ __forceinline__ SampleLoads sample_issue(const GridView& g, double cx, double cy)
{
  SampleLoads L;
  const double dCoordX = cx * g.inv_cell_size;
  const double dCoordY = cy * g.inv_cell_size;
  int xIdx = __double2int_rd(dCoordX);
  int yIdx = __double2int_rd(dCoordY);
  double dx = ((double)xIdx + 0.5) * g.cell_size;
  double dy = ((double)yIdx + 0.5) * g.cell_size;
  if(cx < dx) { xIdx--; dx -= g.cell_size; }
  if(cy < dy) { yIdx--; dy -= g.cell_size; }
  const bool inb = !((xIdx >= g.cells_x) || (xIdx < 0) || (yIdx >= g.cells_y) || (yIdx < 0));
  const int xc = inb ? xIdx : 0, yc = inb ? yIdx : g.row_begin * 32;
  const int py = yc >> 5, px = xc >> 5;
  const int p = py * g.parts_x + px;
  const int x = xc & 31, y = yc & 31;
  L.py = inb ? py : -1;
  L.flag = __ldg(g.flags + p);
  const bool owned = !(py < g.alloc_begin || py >= g.alloc_end);  // readable here
  L.wx = fabs((cx - dx) * g.inv_cell_size);
  L.wy = fabs((cy - dy) * g.inv_cell_size);
  const double* t = g.tsd + (size_t)(owned ? (p - g.alloc_begin * g.parts_x) : 0) * TSD_TILE_STRIDE;
  const int i00 = y * 32 + x;
  const int i10 = (y == 31) ? (TSD_BORDER_OFF + 32 + x) : (i00 + 32);
  const int i01 = (x == 31) ? (TSD_BORDER_OFF + y) : (i00 + 1);
  const int i11 = (x == 31) ? ((y == 31) ? (TSD_BORDER_OFF + 64) : (TSD_BORDER_OFF + y + 1))
                            : ((y == 31) ? (TSD_BORDER_OFF + 32 + x + 1) : (i00 + 33));
  L.g00 = __ldg(t + i00);
  L.g10 = __ldg(t + i10);
  L.g01 = __ldg(t + i01);
  L.g11 = __ldg(t + i11);
  L.pre = !inb ? TSD_INTERPOLATE_INVALIDINDEX : (!owned ? TSD_NOT_OWNED : 0);
  return L;
}

__device__ __forceinline__ int sample_finish(const SampleLoads& L, double* out)
{
  const double v = L.g00 * (1. - L.wy) * (1. - L.wx) + L.g10 * L.wy * (1. - L.wx) + L.g01 * (1. - L.wy) * L.wx +
                   L.g11 * L.wy * L.wx;
  *out = v;
  if(L.pre == TSD_INTERPOLATE_INVALIDINDEX) return TSD_INTERPOLATE_INVALIDINDEX;
  if(!L.flag) return TSD_INTERPOLATE_EMPTYPARTITION;
  if(L.pre == TSD_NOT_OWNED) return TSD_NOT_OWNED;
  if(isnan(v)) return TSD_INTERPOLATE_ISNAN;
  return TSD_INTERPOLATE_SUCCESS;
}

// TsdGrid::interpolateNormal (TsdGrid.cpp:517-546) + norm2 (mathbase.h:212-218)
__device__ __forceinline__ bool sample_normal(const GridView& g, double cx, double cy, double* nx, double* ny)
{
  double inc = 0, dec = 0;
  if(sample_bilinear(g, cx + g.cell_size, cy, &inc) != TSD_INTERPOLATE_SUCCESS) return false;
  if(sample_bilinear(g, cx - g.cell_size, cy, &dec) != TSD_INTERPOLATE_SUCCESS) return false;
  double n0 = inc - dec;
  if(sample_bilinear(g, cx, cy + g.cell_size, &inc) != TSD_INTERPOLATE_SUCCESS) return false;
  if(sample_bilinear(g, cx, cy - g.cell_size, &dec) != TSD_INTERPOLATE_SUCCESS) return false;
  double n1 = inc - dec;
  const double len = sqrt(n0 * n0 + n1 * n1);
  if(!(fabs(len) <= 10e-6))
  {
    n0 /= len;
    n1 /= len;
  }
  *nx = n0;
  *ny = n1;
  return true;
}

// `M = T * v` for a 3x3 T and a 3-vector v the way gslcblas dgemm NoTrans x NoTrans does it (k outer,
// zero coefficients skipped).  Rows 0 and 1 only.
__device__ __forceinline__ void mat3_vec_nn(const double* T, double v0, double v1, double v2, double* o0, double* o1)
{
  double r0 = 0.0, r1 = 0.0;
  if(T[0] != 0.0) r0 += T[0] * v0;
  if(T[3] != 0.0) r1 += T[3] * v0;
  if(T[1] != 0.0) r0 += T[1] * v1;
  if(T[4] != 0.0) r1 += T[4] * v1;
  if(T[2] != 0.0) r0 += T[2] * v2;
  if(T[5] != 0.0) r1 += T[5] * v2;
  *o0 = r0;
  *o1 = r1;
}

// ---- release / acquire signals at system scope (peer memory over NVLink, or another process on the same GPU) ----
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// sequence numbers wrap: compare as a signed distance.  A neighbour that never calls (the C ABI asks both sides of a
// boundary to call equally often) must not hang the GPU, and must not take the CUDA context -- and the map -- down
// either: after ~10 s the wait gives up and raises an error flag that the host reports (TSD_E_CUDA) at the next
// synchronisation; the kernel finishes with whatever arrived.
__device__ __forceinline__ void wait_seq(const uint32_t* p, uint32_t seq, uint32_t* err)
{
  for(unsigned spins = 0; (int32_t)(ld_acquire_sys(p) - seq) < 0; spins++)
  {
    __nanosleep(256);
    if(spins > (1u << 25))
    {
      atomicExch(err, 1u);
      return;
    }
  }
}


#endif  // __CUDACC__

int fill_scan_dev(const tsd_scan_t* scan, ScanDev* out);  // scalars only (pointers are set by the caller)

}  // namespace tsd

// the grid handle is shared between grid.cu, raycast.cu and match.cu
struct tsd_grid
{
  // One handle may be used from several host threads (the node's mapper pushes while its localisers ray-cast and
  // its grid thread publishes, ThreadMapping.cpp:46-61, ThreadLocalize.cpp:353, ThreadGrid.cpp:84): every entry point
  // holds this lock from staging through launch, read-back and synchronisation.  (Recursive: entry points call each other.)
  std::recursive_mutex* mtx;
  int device;
  cudaStream_t stream;
  int layout_grid;
  int cells_x, cells_y, parts_x, parts_y, n_parts;
  int row_begin, row_end, n_owned;  // owned partition rows / partitions
  int alloc_begin, alloc_end, n_alloc;  // allocated rows: owned + halo rows (bands only)
  bool band;                        // sharded grid: push runs in two phases around the halo exchange
  bool band_push_open;
  bool refresh_all_pending;        // an upload / fill asked for a full border refresh at the next push
  double cell_size, inv_cell_size, max_truncation;
  double min_x, max_x, min_y, max_y;
  bool pushed_once;
  bool stats_fresh;                // h_counters holds the statistics of the last push (blocking tsdg_push)

  double* d_tsd;
  double* d_weight;
  uint8_t* d_flags;    // n_parts
  double* d_initw;     // n_parts
  // per-push work lists (capacity n_owned each) and their per-item data
  uint32_t* d_active;  // bit 31: partition was initialised before this push
  uint32_t* d_kinds;   // per work-list entry: 2-bit outcome per scan of the launch
  double* d_active_w;  // 0.01 * partWeight per active item
  uint32_t* d_newly;   // partitions allocated by the current push
  uint32_t* d_pending; // partitions initialised/modified outside push: borders refreshed by the next push
  uint32_t* d_counters;   // [0] work-list entries [2] pending [3] refresh-all flag [4] newly allocated [5] slow-path cells
                          // [6] emptied tiles [7] active tiles [8..15] snapshot of the last push [16..21] tickets, list sizes, update count
  unsigned long long* d_stats64;  // [0] cell updates
  double* d_coltab;    // 3 * cells_x : A (0.0 + Pi00*X), B (0.0 + Pi10*X), D ((X-tx)^2)
  double* d_rowtab;    // 3 * cells_y : A (Pi01*Y), B (Pi11*Y), D ((Y-ty)^2)
  unsigned long long* d_prof;  // UPDATE_PROFILE builds only
  tsd::ScanDev* d_scans;  // the scans of the current push launch, for out-of-line device code
  float4* d_col4;      // per scan cells_x x {Pi00*(X-tx'), Pi10*(X-tx')} as float2, then cells_x x (X-tx)^2 as float (k_update fast path)
  float4* d_row4;      // cells_y per scan: {Pi01*(Y-ty'), Pi11*(Y-ty'), (Y-ty)^2, 0}
  float2* d_gate;      // scan_cap per scan: per beam the squared distances below / above which a cell is free space for
                       // certain / not rewritten for certain (grid.cu fill_gate)
  // sensor model tables
  double2* d_dirs;
  int dirs_n;
  double dirs_phi_min, dirs_res;
  // staged scan + rays: one device block and one pinned mirror (grid.cu ensure_scan_capacity)
  int scan_cap;
  size_t in_bytes, rc_bytes;
  unsigned char* d_in;
  unsigned char* h_in;   // pinned; = h_in2[in_next ^ 1] after a staging call
  unsigned char* h_in2[2];  // two pinned blocks take turns, so that staging a scan does not wait for the previous push
  cudaEvent_t ev_in[2];     // recorded after the H2D copy out of block i
  bool ev_in_used[2];
  int in_next;
  unsigned char* d_rc;
  unsigned char* h_rc;   // pinned
  unsigned long long rc_steps_prev[2];
  double* d_ranges;
  uint8_t* d_mask;
  double* h_ranges;
  uint8_t* h_mask;
  // raycast buffers (sized with the scan)
  double* d_rays;
  double* d_rc_out;        // 4 doubles per beam: cx cy nx ny
  unsigned long long* d_rc_keys;
  unsigned long long* d_rc_steps;  // [0] fine [1] coarse
  double* h_rc_out;        // pinned
  unsigned long long* h_rc_keys;  // pinned
  unsigned long long* h_rc_steps; // pinned
  double* h_rays;          // pinned
  // generic pinned / device scratch for batched queries
  size_t scratch_cap;
  void* d_scratch;
  void* h_scratch;
  // stats of the last push (pinned)
  uint32_t* h_counters;
  unsigned long long* h_stats64;
  tsd_push_stats_t last_stats;
  int sm_count;
  tsd::ScanDev staged[4];  // scans staged by tsdg_stage_scan / tsdg_stage_batch (device pointers + scalars)
  int staged_n;
  // halo synchronisation over peer memory (bands only): [0] = the band below, [1] = the band above
  uint32_t* d_signal;        // [0]/[1] data-ready from below/above, [2]/[3] ack from below/above, [4] CTA ticket, [5] timeout flag
  struct Peer
  {
    bool connected, ipc;
    double* tsd;             // the neighbour's cell arrays (its allocation base) as seen from this process
    double* weight;
    uint32_t* signal;
    int alloc_begin;         // the neighbour's first allocated partition row
  } peer[2];
  uint32_t halo_seq[2];      // synchronisations done per boundary (both sides count alike)
  // ray-cast exchange over peer memory (raycast.cu): every band stores its per-beam first events into its slot of EVERY
  // band's block; [2 parities][TSD_RCX_MAX slots][TSD_RCX_CAP beams] keys, the same of 4-double payloads, then signals
  unsigned char* d_rcx;
  unsigned char* peer_rcx[16];
  bool peer_rcx_ipc[16];
  int rcx_rank, rcx_world;
  uint32_t rcx_seq;
  bool halo_used;            // k_halo_sync ran: tsdg_sync looks at its timeout flag (d_signal[5])
  bool has_staged;
  unsigned update_filter;  // measurement aid: k_update skips K2 (bit 0) / K3 (bit 1) work (tsdg_set_update_filter)
  bool timing;          // record CUDA events around the push kernels (bench.py's live roofline)
  cudaEvent_t ev[4];
  cudaEvent_t ev_order;
};

int tsd_raycast_enqueue(tsd_grid_t* g, const tsd_scan_t* scan, const double* rays_world);  // raycast.cu

#define TSD_LOCK(g)                                      \
  std::unique_lock<std::recursive_mutex> tsd_lock__;     \
  if(g) tsd_lock__ = std::unique_lock<std::recursive_mutex>(*(g)->mtx)

namespace tsd
{
int grid_stage_scan(tsd_grid* g, const tsd_scan_t* scan, ScanDev* sd, const double* rays_world);  // one H2D copy
int grid_stage_scans(tsd_grid* g, const tsd_scan_t* scans, int n, ScanDev* sd, const double* rays_world);
int grid_ensure_scratch(tsd_grid* g, size_t bytes);
GridView grid_view(const tsd_grid* g);
}
