// Icp::iterate on the device (K6-K8): the whole registration loop -- pre-filter, exact nearest-neighbour
// pairing, distance filter, reciprocal filter, closed-form estimate, transform update -- is ONE kernel
// launch of one thread-block cluster (ICP_CLUSTER CTAs, one SM each); model, scene and the search structure
// live in shared memory for all iterations, so an ICP run costs one H2D copy, one launch and one D2H copy.
//
// Reference: src/obvision/registration/icp/Icp.cpp:464-512 (iterate), :410-462 (step), :371-408
// (applyTransformation); assign/PairAssignment.cpp:38-84; assign/FlannPairAssignment.cpp:64-92;
// assign/filter/OutOfBoundsFilter2D.cpp:27-37, DistanceFilter.cpp:32-64, ReciprocalFilter.cpp:32-78;
// ClosedFormEstimator2D.cpp:36-109.  Wiring: src/ThreadLocalize.cpp:210-225, :571-581.
//
// Pairing replaces FLANN's kd-tree by a spatial hash over the model points (cells of edge h = dist_max / 4,
// hashed into ICP_SLOTS buckets), built once per run (the model does not move during ICP).  The search is
// EXACT: square rings of cells around the query are visited until the best squared distance is strictly
// below the squared distance to everything unvisited, or until everything unvisited is beyond the distance
// filter's current threshold (such a pair is dropped by DistanceFilter.cpp:38 whatever its model index).
// Hash collisions only add candidates, never remove any.  Distances are computed as FLANN's L2 functor does
// ((0 + dx*dx) + dy*dy); ties go to the lowest model index, the rule the oracle's FLANN stand-in uses.
//
// Work split and communication inside the cluster: see k_icp.
//
// Sums of the estimator are block reductions with a fixed tree (over the model points in hash-slot order), taken
// about a provisional centre and corrected exactly, and the rotation's cos / sin come from the sums directly instead of
// through atan2: results are deterministic -- the same bits on every run and in every CTA -- but not bit-identical to
// the reference's sequential sums.  Pair lists are compared exactly, poses to 1e-9 (tests/test_gpu_parity.py).
#include <string.h>

#include <vector>

#include <cooperative_groups.h>

#include "common.cuh"
#include "icp_cells.cuh"

using namespace tsd;

#ifndef ICP_THREADS
#define ICP_THREADS 256   // per CTA.  An iteration is a dozen short phases that every warp walks through, and most of what a CTA
                          // issues is that walk, not work: measured per CTA and iteration, 1024 threads issue 21 k warp
                          // instructions, 256 threads 6 k, for the same pairs (at most 256 queries per CTA either way)
#endif
#define ICP_MAX_POINTS 2048
#define ICP_SLOTS 4096  // hash buckets
#define ICP_SUM_THREADS 256  // threads that walk the model points for the estimator's sums
#ifndef ICP_CLUSTER
#define ICP_CLUSTER 8   // CTAs (SMs) per registration: the portable cluster size
#endif

struct IcpParams
{
  int nM, nS;
  const int* nM_dev;    // if set: the number of model points is read on the device (tsdg_localize: the ray caster's hits)
  int max_iterations;
  unsigned conv_cnt;
  double max_rms;
  double max_dist_sqr, min_dist_sqr, multiplier;
  double x_min, x_max, y_min, y_max;
  double pose[9];
  double t_init[16];
  int has_init;
  double hash_h;        // cell edge of the spatial hash over the model
  double coarse_h;      // cell edge of the coarse occupancy bitmap (>= the largest distance threshold)
  int max_rings;        // rings after which everything unvisited is beyond the distance filter
  const double* model;  // nM x 2
  const double* scene;  // nS x 2
  // outputs
  double* result;       // [0..8] T 3x3, [9] mse, [10] pairs, [11] iterations, [12] state
  // trace
  int trace;            // 0: skip the per-iteration pair lists (icp_set_trace)
  int cap;
  unsigned* tr_model;
  unsigned* tr_scene;
  int* tr_count;
  double* tr_mse;
  double* tr_T;
};

struct tsd_icp
{
  std::recursive_mutex* mtx;  // one registration at a time per handle (each localiser thread owns one in the node)
  int device;
  cudaStream_t stream;
  IcpParams p;
  int cap;
  double* d_model;
  double* d_scene;
  double* d_result;
  unsigned* d_tr_model;
  unsigned* d_tr_scene;
  int* d_tr_count;
  double* d_tr_mse;
  double* d_tr_T;
  double* h_stage;   // pinned: model + scene
  double* h_result;  // pinned
  int last_nM, last_nS;
  int* d_nM;         // tsdg_localize: the model's size, counted on the device
  int trace;
  int trace_cap_it;
};

// cell coordinate of a point (icp_cells.cuh: host/device, checked on the CPU by tests/cpp/icpcell_check.cpp)
__device__ __forceinline__ int cell_of(double v, double v0, double invh) { return tsd_icp_cell_of(v, v0, invh); }

__device__ __forceinline__ unsigned slot_of(int cx, int cy)
{
  const unsigned hsh = (unsigned)cx * 73856093u ^ (unsigned)cy * 19349663u;
  return (hsh ^ (hsh >> 15)) & (ICP_SLOTS - 1);
}

// sums NV values per thread over the block; results land in s_red[64 + k]
template <int NV>
__device__ __forceinline__ void block_sum_n(double* v, double* s_red, int tid)
{
#pragma unroll
  for(int k = 0; k < NV; k++)
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  __syncthreads();
  if((tid & 31) == 0)
#pragma unroll
    for(int k = 0; k < NV; k++) s_red[k * 32 + (tid >> 5)] = v[k];
  __syncthreads();
  if(tid < 32 * NV)
  {
    double r = s_red[tid];  // warp k holds the 32 partials of value k
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if((tid & 31) == 0) s_red[192 + (tid >> 5)] = r;
  }
  __syncthreads();
#pragma unroll
  for(int k = 0; k < NV; k++) v[k] = s_red[192 + k];
}

// Distributed shared memory is only ever written with plain stores (st.shared::cluster through a mapa address), each
// word by exactly one remote thread between two cluster barriers.  Remote 64-bit min atomics are not usable: for a shared::cluster address that is not the CTA's own window, the
// compiler's atomicMin(unsigned long long) expands to a plain load / compare / store (seen in the SASS, and as
// lost updates in the pair lists), so every atomic below stays inside the CTA that owns the word.
__device__ __forceinline__ uint32_t dsmem_addr(const void* own_smem, unsigned rank)
{
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(own_smem);
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void dsmem_st_v2u64(uint32_t addr, unsigned long long a, unsigned long long b)
{
  asm volatile("st.shared::cluster.v2.u64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void dsmem_st_u32(uint32_t addr, unsigned v)
{
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
#ifdef ICP_PROFILE
#define ICP_STAMP(k) { const long long now__ = clock64(); pf[k] += now__ - pf_t; pf_t = now__; }
#else
#define ICP_STAMP(k)
#endif
__device__ __forceinline__ void st_release_cluster(uint32_t addr, unsigned v)
{
  asm volatile("st.release.cluster.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_cluster(const unsigned* own_smem)
{
  unsigned v;
  asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(own_smem)) : "memory");
  return v;
}

// One thread-block CLUSTER of ICP_CLUSTER CTAs (one SM each) runs the whole loop.  Every CTA keeps the model, its
// search structure, the scene and the transformation in its own shared memory and evolves them identically (same
// code, same data, same reduction trees), so no state is ever broadcast.  What is split is the nearest-neighbour
// search, the one part of an iteration that is instructions rather than latency (measured: one SM alone issues
// 17 k cycles of it per iteration on the C3 input, eight SMs 4 k): query i belongs to CTA i % ICP_CLUSTER, one lane
// per query.
//   * search: the previous iteration's neighbour seeds the bound; a lane first finds out which cells it has to open at
//     all (its own, then the neighbours that are occupied and not farther than the bound -- a cell is skipped only if
//     it is STRICTLY farther, so the result and the lowest-index tie rule are those of the full search), then walks
//     its own short list: a warp runs as many rounds as its busiest lane has cells, not nine.  The few queries without
//     a model point within one cell edge are finished by their whole warp (window of the distance filter's radius).
//   * exchange: every lane stores its result {distance, model index} into EVERY CTA's copy of the result table
//     (16-byte st.shared::cluster), then one thread per peer releases a flag in that peer's shared memory and every
//     CTA waits for its ICP_CLUSTER flags -- no cluster-wide barrier of 8192 threads, no second round.  Tables are
//     double-buffered by iteration parity: a CTA can only get two iterations ahead of the slowest reader.
//   * reciprocal filter, estimator, pose update and scene transform are replicated.  The estimator is ONE reduction:
//     centroids and cross sums in one pass, the cross sums taken about a provisional centre c' (last iteration's
//     centroids) and corrected exactly: sum (a - c)(b - d) = sum (a - c')(b - d') - n (c - c')(d - d'); only
//     ICP_SUM_THREADS threads walk the model points so that few warps pay for the 64-bit shuffle trees; cos and sin
//     of atan2(n0, n1) are n1 / |n| and n0 / |n|.
// Per iteration: one flag exchange and five block barriers.  (History, cycles per iteration on C3, 869 model / 686
// scene points: two-level reciprocal filter with two cluster barriers and a two-pass estimator 25 k; everything on
// one CTA 25 k, of which 17 k search.)
struct IcpSmem
{
  double *mx, *my, *sx, *sy, *lb, *red, *T, *part, *pbd;
  unsigned short *pend, *pos;
  ulonglong2* nnd;
  unsigned long long* best;
  unsigned *win, *scan, *coarse, *occ, *flag;
  int* prev;
  unsigned short *bstart, *bcnt, *bidx;
};
__device__ __forceinline__ IcpSmem icp_carve(unsigned char* smem, int nM, int nS)
{
  IcpSmem S;
  S.mx = reinterpret_cast<double*>(smem);
  S.my = S.mx + nM;
  S.sx = S.my + nM;
  S.sy = S.sx + nS;
  S.nnd = reinterpret_cast<ulonglong2*>(S.sy + nS);               // 2 x nS: {distance bits, model index} of every query, by parity
  S.lb = reinterpret_cast<double*>(S.nnd + 2 * (size_t)nS);      // nQ: own queries
  S.pbd = S.lb + (nS + ICP_CLUSTER - 1) / ICP_CLUSTER;            // nQ: bound of a far query before its window scan
  S.red = S.pbd + (nS + ICP_CLUSTER - 1) / ICP_CLUSTER;                                              // 320 partials + 16 results
  S.part = S.red + 336;                                           // 7 x ICP_SUM_THREADS partial sums
  S.T = S.part + 7 * ICP_SUM_THREADS;                                              // Tfinal 16, Tlast 16
  S.best = reinterpret_cast<unsigned long long*>(S.T + 32);       // nM
  S.win = reinterpret_cast<unsigned*>(S.best + nM);               // nM
  S.prev = reinterpret_cast<int*>(S.win + nM);                    // nQ: own queries
  S.flag = reinterpret_cast<unsigned*>(S.prev + (nS + ICP_CLUSTER - 1) / ICP_CLUSTER);  // ICP_CLUSTER
  S.scan = S.flag + ICP_CLUSTER;                                  // 40 + 32
  S.coarse = S.scan + 72;                                         // 128
  S.occ = S.coarse + 128;                                         // 128
  S.bstart = reinterpret_cast<unsigned short*>(S.occ + 128);      // ICP_SLOTS + 2
  S.bcnt = S.bstart + (ICP_SLOTS + 2);                            // ICP_SLOTS
  S.pend = S.bcnt + ICP_SLOTS;                                    // nQ (+ pad): far queries of this iteration
  S.bidx = S.pend + (((nS + ICP_CLUSTER - 1) / ICP_CLUSTER + 2) & ~1);
  S.pos = S.bidx + ((nM + 2) & ~1);                                    // nM (+ pad)
  return S;
}
static size_t icp_smem_bytes(int nM, int nS)
{
  const size_t nQ = ((size_t)nS + ICP_CLUSTER - 1) / ICP_CLUSTER;
  size_t b = sizeof(double) * (2 * (size_t)nM + 2 * (size_t)nS + 2 * nQ + 336 + 7 * ICP_SUM_THREADS + 32) + 32 * (size_t)nS;  // mx my sx sy lb red T, nnd
  b += 8 * (size_t)nM + 4 * (size_t)nM + 4 * nQ + 4 * (ICP_CLUSTER + 72 + 256);                    // best win prev flag scan coarse occ
  b += 2 * ((size_t)ICP_SLOTS + 2 + ICP_SLOTS + 2 * ((size_t)nM + 2) + nQ + 2);
  return b + 64;
}

__global__ void __cluster_dims__(ICP_CLUSTER, 1, 1) __launch_bounds__(ICP_THREADS, 1) k_icp(IcpParams P)
{
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int nM = P.nM_dev ? *P.nM_dev : P.nM, nS = P.nS;
  if(nM <= 0)
  {
    // Icp.cpp:467-471: no model, no registration (the host checks this itself when it knows the count)
    if(rank == 0 && tid == 0)
    {
      for(int i = 0; i < 9; i++) P.result[i] = (i % 4 == 0) ? 1.0 : 0.0;
      P.result[9] = 0.0; P.result[10] = 0.0; P.result[11] = 0.0;
      P.result[12] = (double)TSD_ICP_NOTMATCHABLE;
      P.result[13] = 0.0;
    }
    return;  // (every CTA of the cluster, before any of them waits for another)
  }
  const int nQ = (nS + ICP_CLUSTER - 1) / ICP_CLUSTER;  // queries of this CTA (upper bound)
  const IcpSmem S = icp_carve(smem, nM, nS);
  double* const s_mx = S.mx; double* const s_my = S.my; double* const s_sx = S.sx; double* const s_sy = S.sy;
  const unsigned long long INF64 = 0xffffffffffffffffULL;
  for(int i = tid; i < nM; i += ICP_THREADS) { S.best[i] = INF64; S.win[i] = 0xffffffffu; }
  for(int i = tid; i < nS; i += ICP_THREADS)
  {
    s_sx[i] = P.scene[2 * i]; s_sy[i] = P.scene[2 * i + 1];
  }
  for(int i = tid; i < nQ; i += ICP_THREADS) { S.lb[i] = 0.0; S.prev[i] = -1; }
  if(tid < ICP_CLUSTER) S.flag[tid] = 0u;
  if(tid == 0) S.scan[64] = 0u;
  for(int i = tid; i < ICP_SLOTS; i += ICP_THREADS) S.bcnt[i] = 0;
  if(tid < 128) { S.coarse[tid] = 0u; S.occ[tid] = 0u; }
  if(tid < 16) { S.T[tid] = (tid % 5 == 0) ? 1.0 : 0.0; S.T[16 + tid] = S.T[tid]; }
  __syncthreads();

  // ---- spatial hash of the model: counting sort of the points by hash slot ----
  const double h = P.hash_h, invh = 1.0 / P.hash_h;
  const double invhc = 1.0 / P.coarse_h;
  // The model is kept SORTED by hash slot (ties: by index), and from here on a model point is known by its position in
  // that order: the points of a cell are consecutive in memory, so the search reads coordinates without the detour over
  // an index list.  S.bidx[p] is the index the caller knows the point by (the tie rule and the pair lists use it).
  const double bx0 = P.model[0], by0 = P.model[1];
  for(int i = tid; i < nM; i += ICP_THREADS)
  {
    const double mxi = P.model[2 * i], myi = P.model[2 * i + 1];
    const unsigned b = slot_of(cell_of(mxi, bx0, invh), cell_of(myi, by0, invh));
    atomicAdd(reinterpret_cast<unsigned*>(S.bcnt) + (b >> 1), (b & 1) ? 0x10000u : 1u);
    const unsigned c = slot_of(cell_of(mxi, bx0, invhc), cell_of(myi, by0, invhc));
    atomicOr(&S.coarse[c >> 5], 1u << (c & 31));
    atomicOr(&S.occ[b >> 5], 1u << (b & 31));
  }
  __syncthreads();
  {
    // exclusive prefix over the ICP_SLOTS counters, ICP_SLOTS / ICP_THREADS consecutive ones per thread
    constexpr int PER = ICP_SLOTS / ICP_THREADS;
    unsigned mine = 0;
#pragma unroll 1
    for(int k = 0; k < PER; k++) mine += S.bcnt[PER * tid + k];
    unsigned incl = mine;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if(lane >= o) incl += t;
    }
    if(lane == 31) S.scan[tid >> 5] = incl;
    __syncthreads();
    unsigned off = incl - mine;
    for(int w = 0; w < (tid >> 5); w++) off += S.scan[w];
#pragma unroll 1
    for(int k = 0; k < PER; k++)
    {
      const unsigned c = S.bcnt[PER * tid + k];
      S.bstart[PER * tid + k] = (unsigned short)off;
      S.bcnt[PER * tid + k] = 0;
      off += c;
    }
    if(tid == ICP_THREADS - 1) S.bstart[ICP_SLOTS] = (unsigned short)off;
    __syncthreads();
  }
  for(int i = tid; i < nM; i += ICP_THREADS)
  {
    const unsigned b = slot_of(cell_of(P.model[2 * i], bx0, invh), cell_of(P.model[2 * i + 1], by0, invh));
    const unsigned old = atomicAdd(reinterpret_cast<unsigned*>(S.bcnt) + (b >> 1), (b & 1) ? 0x10000u : 1u);
    const unsigned within = (b & 1) ? (old >> 16) : (old & 0xffffu);
    S.bidx[S.bstart[b] + within] = (unsigned short)i;
  }
  __syncthreads();
  // (the atomics above hand out places within a slot in any order: put them in index order, so that every CTA of the
  //  cluster -- and every run -- numbers the points the same way)
  for(int b = tid; b < ICP_SLOTS; b += ICP_THREADS)
  {
    const int k0 = S.bstart[b], k1 = S.bstart[b + 1];
    for(int k = k0 + 1; k < k1; k++)
    {
      const unsigned short v = S.bidx[k];
      int j = k - 1;
      while(j >= k0 && S.bidx[j] > v) { S.bidx[j + 1] = S.bidx[j]; j--; }
      S.bidx[j + 1] = v;
    }
  }
  __syncthreads();
  for(int k = tid; k < nM; k += ICP_THREADS)
  {
    const int i = S.bidx[k];
    s_mx[k] = P.model[2 * i];
    s_my[k] = P.model[2 * i + 1];
    S.pos[i] = (unsigned short)k;
  }
  // ---- Icp::iterate (Icp.cpp:480-487): initial transformation ----
  if(P.has_init)
  {
    const double r00 = P.t_init[0], r01 = P.t_init[1], r10 = P.t_init[4], r11 = P.t_init[5];
    const double t0 = P.t_init[3], t1 = P.t_init[7];
    for(int i = tid; i < nS; i += ICP_THREADS)
    {
      const double x = s_sx[i], y = s_sy[i];
      double a = 0.0; a += x * r00; a += y * r01; a = 0.0 + 1.0 * a;
      double b = 0.0; b += x * r10; b += y * r11; b = 0.0 + 1.0 * b;
      s_sx[i] = a + t0;
      s_sy[i] = b + t1;
    }
    if(tid == 0)
    {
      double out[16];
      for(int i = 0; i < 16; i++) out[i] = 0.0;
      for(int k = 0; k < 4; k++)
        for(int i = 0; i < 4; i++)
        {
          const double temp = 1.0 * P.t_init[4 * i + k];
          if(temp != 0.0)
            for(int j = 0; j < 4; j++) out[4 * i + j] += temp * S.T[4 * k + j];
        }
      for(int i = 0; i < 16; i++) S.T[i] = out[i];
    }
  }
  cluster.sync();  // every CTA is resident and initialised before anybody addresses its shared memory

  int eRetval = TSD_ICP_PROCESSING;
  unsigned iter = 0;
  double rms_prev = 10e12;
  unsigned conv_cnt = 0;
  double rms = 0.0;
  unsigned pairs = 0;
  double distSqr = P.max_dist_sqr;
  // provisional centres of the estimator's cross sums (any point near the data; afterwards the last centroids)
  double pm0 = bx0, pm1 = by0, ps0 = s_sx[0], ps1 = s_sy[0];
  const double INF = __longlong_as_double(0x7ff0000000000000LL);

#ifdef ICP_PROFILE
  long long pf[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pf_t = clock64();
#endif
  while(eRetval == TSD_ICP_PROCESSING)
  {
    ICP_STAMP(7)
    // ---- A: pre-filter + exact 1-NN + distance filter, one lane per query ----
    ulonglong2* const nndCur = S.nnd + (size_t)(iter & 1u) * nS;
    // what a finished search leaves behind: the lower bound and the seed for the next iteration, and the result in
    // every CTA's table (its own included)
    auto finish = [&](int ql, int q, int best, double bestD, double lbNew)
    {
      S.lb[ql] = lbNew;
      const bool keep = (best >= 0) && (bestD <= distSqr);  // DistanceFilter.cpp:38
      S.prev[ql] = best >= 0 ? best : S.prev[ql];
      const unsigned long long db = (unsigned long long)__double_as_longlong(bestD);
      const unsigned long long mi = keep ? (unsigned long long)best : 0xffffffffffffffffULL;
      const uint32_t own = (uint32_t)__cvta_generic_to_shared(nndCur + q);
#pragma unroll 1
      for(unsigned r = 0; r < ICP_CLUSTER; r++)
      {
        uint32_t ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(own), "r"(r));
        dsmem_st_v2u64(ra, db, mi);
      }
    };
    // (What an iteration costs is the instructions its 32 warps issue, so warps without queries skip the search altogether.)
    for(int l0 = 0; l0 < nQ; l0 += ICP_THREADS)
    {
      // (queries are dealt to the warps like cards: a warp's time is its busiest lane's, and eight warps with eleven
      //  queries each finish sooner than three warps with thirty-two)
      if(l0 + (tid >> 5) >= nQ) continue;           // (warp-uniform: the warp's first query)
      const int ql = l0 + (tid & 31) * (ICP_THREADS / 32) + (tid >> 5);  // index among this CTA's queries
      const int q = ql * ICP_CLUSTER + (int)rank;   // scene point
      const bool exists = ql < nQ && q < nS;
      const double x = exists ? s_sx[q] : 0.0, y = exists ? s_sy[q] : 0.0;
      // OutOfBoundsFilter2D.cpp:27-37: S.transform(pose) = S * R^T + t
      double tx = 0.0; tx += x * P.pose[0]; tx += y * P.pose[1]; tx = 0.0 + 1.0 * tx; tx += P.pose[2];
      double ty = 0.0; ty += x * P.pose[3]; ty += y * P.pose[4]; ty = 0.0 + 1.0 * ty; ty += P.pose[5];
      bool search = exists && !(tx < P.x_min || tx > P.x_max || ty < P.y_min || ty > P.y_max);
      int best = -1;
      double bestD = INF;
      // A point whose nearest model point is provably farther than the distance filter's threshold cannot yield a
      // pair (DistanceFilter.cpp:38): its search is skipped.  lb is a lower bound of that distance, carried over
      // from the last search and reduced by how far the point moved since.
      double lbNew = exists ? S.lb[ql] : 0.0;
      if(search && lbNew * lbNew > distSqr) search = false;
      else if(search) lbNew = 0.0;
      // (the loops below are deliberately NOT unrolled: unrolled, the kernel is 100 KB of code, and an iteration that walks
      //  through it once is bound by instruction fetch -- measured: 3 k cycles for a phase that issues 300 instructions)
      if(search)
      {
        const int cqx = cell_of(x, bx0, invhc), cqy = cell_of(y, by0, invhc);
        unsigned any = 0;
#pragma unroll
        for(int c = 0; c < 9; c++)
        {
          const unsigned cs = slot_of(cqx + c % 3 - 1, cqy + c / 3 - 1);
          any |= (S.coarse[cs >> 5] >> (cs & 31)) & 1u;
        }
        search = any != 0;
        if(!search) lbNew = P.coarse_h * (1.0 - 1e-6);  // nothing within one coarse cell
      }
      ICP_STAMP(8)
      const int qx = cell_of(x, bx0, invh), qy = cell_of(y, by0, invh);
      bool pending = false;
      // One lane per query.  The lane opens its own cell first, then finds out which neighbours it has to open at all
      // (occupied, and not farther than the bound), then walks that short list: a warp runs as many rounds as its
      // busiest lane has cells, not nine.  Cell c of the 3x3 block is (c % 3 - 1, c / 3 - 1); bit 4 is the query's own.
      unsigned open = 0;
      double l2 = 0.0, r2 = 0.0, d2 = 0.0, u2 = 0.0;  // squared distances to the edges of the own cell
      if(search)
      {
        const int pm = S.prev[ql];
        if(pm >= 0)
        {
          const double d0 = x - s_mx[pm], d1 = y - s_my[pm];
          double d = 0.0;
          d += d0 * d0;
          d += d1 * d1;
          bestD = d;
          best = pm;
        }
        // shaved so that rounding never skips a cell wrongly
        tsd_icp_edge_gaps2(x, y, bx0, by0, h, invh, qx, qy, &l2, &r2, &d2, &u2);
        open = 1u << 4;
      }
      bool first = true;
      ICP_STAMP(12)
#pragma unroll 1
      while(__any_sync(0xffffffffu, open != 0))
      {
#ifdef ICP_PROFILE
        pf[13] += 1;
#endif
        if(open)
        {
          const int c = first ? 4 : (__ffs(open) - 1);
          open &= ~(1u << c);
          const int dx = c % 3 - 1, dy = c / 3 - 1;
          const double g2 = (dx < 0 ? l2 : (dx > 0 ? r2 : 0.0)) + (dy < 0 ? d2 : (dy > 0 ? u2 : 0.0));
          const unsigned b = slot_of(qx + dx, qy + dy);
          if(!(g2 > bestD) && ((S.occ[b >> 5] >> (b & 31)) & 1u))  // (the bound may have shrunk since the list was made)
          {
            const int k1 = S.bstart[b + 1];
#pragma unroll 4
            for(int k = S.bstart[b]; k < k1; k++)
            {
              const double d0 = x - s_mx[k];
              const double d1 = y - s_my[k];
              double d = 0.0;
              d += d0 * d0;
              d += d1 * d1;
              if(d < bestD || (d == bestD && S.bidx[k] < S.bidx[best])) { bestD = d; best = k; }
            }
          }
          if(first)
          {
            // the neighbours worth opening, given what the own cell and last iteration's neighbour yielded
            first = false;
            // (straight-line code: the eight tests overlap; with three to five warps per CTA in this phase, latency counts)
#pragma unroll
            for(int c2 = 0; c2 < 9; c2++)
            {
              if(c2 == 4) continue;
              const int ex = c2 % 3 - 1, ey = c2 / 3 - 1;
              const double e2 = (ex < 0 ? l2 : (ex > 0 ? r2 : 0.0)) + (ey < 0 ? d2 : (ey > 0 ? u2 : 0.0));
              const unsigned bb = slot_of(qx + ex, qy + ey);
              const unsigned occ = (S.occ[bb >> 5] >> (bb & 31)) & 1u;
              if(!(e2 > bestD)) open |= occ << c2;  // (a cell farther than the bound holds nothing as close)
            }
          }
        }
      }
      ICP_STAMP(9)
      ICP_STAMP(10)
      if(search)
      {
        // everything unvisited lies in cells at Chebyshev distance > 1, i.e. farther than h
        const double lb = h * (1.0 - 1e-9);
        const double lb2 = lb * lb;
        if(bestD < lb2 || distSqr < lb2) lbNew = fmin(sqrt(bestD), lb) * (1.0 - 1e-9);  // shaved against rounding
        else pending = true;
      }
      ICP_STAMP(11)
      // the few queries without a model point within one cell edge go to a list that all warps of the CTA work off below
      if(pending)
      {
        const unsigned at = atomicAdd(S.scan + 64, 1u);
        S.pend[at] = (unsigned short)ql;
        S.pbd[ql] = bestD;
        S.prev[ql] = best;  // (seed or nothing yet)
      }
      else if(exists) finish(ql, q, best, bestD, lbNew);
    }
    __syncthreads();
    {
      const unsigned nPend = S.scan[64];
      if(nPend)
      {
        // smallest R with (R h)^2 > distSqr: the window then holds every point the distance filter can keep
        int R = 2;
        while(R < P.max_rings && !(distSqr < ((double)R * h * (1.0 - 1e-9)) * ((double)R * h * (1.0 - 1e-9)))) R++;
#pragma unroll 1
        for(unsigned at = (unsigned)(tid >> 5); at < nPend; at += ICP_THREADS / 32)
        {
          // one warp per query: the 32 lanes scan the rest of the window
          const int ql = S.pend[at];
          const int q = ql * ICP_CLUSTER + (int)rank;
          const double px = s_sx[q], py = s_sy[q];
          const int pqx = cell_of(px, bx0, invh), pqy = cell_of(py, by0, invh);
          double bestD = S.pbd[ql];
          int best = lane == 0 ? S.prev[ql] : -1;  // (only lane 0 merges and writes the result)
          // a cell in ring k is farther than (k - 1) h: the bound found so far (last iteration's neighbour) limits the rings
          int Rq = R;
          if(bestD < INF) Rq = min(R, (int)(sqrt(bestD) * invh * (1.0 + 1e-9)) + 1);
          const int Wq = 2 * Rq + 1;
          double wd = INF;
          int wb = -1;
#pragma unroll 1
          for(int c = lane; c < Wq * Wq; c += 32)
          {
            const int dx = c % Wq - Rq, dy = c / Wq - Rq;
            if(dx >= -1 && dx <= 1 && dy >= -1 && dy <= 1) continue;  // done above
            const unsigned b = slot_of(pqx + dx, pqy + dy);
            if(!((S.occ[b >> 5] >> (b & 31)) & 1u)) continue;
            const int k1 = S.bstart[b + 1];
            for(int k = S.bstart[b]; k < k1; k++)
            {
              const double d0 = px - s_mx[k];
              const double d1 = py - s_my[k];
              double d = 0.0;
              d += d0 * d0;
              d += d1 * d1;
              if(d < wd || (d == wd && S.bidx[k] < S.bidx[wb])) { wd = d; wb = k; }
            }
          }
#pragma unroll
          for(int o = 16; o > 0; o >>= 1)
          {
            const double od = __shfl_xor_sync(0xffffffffu, wd, o);
            const int ob = __shfl_xor_sync(0xffffffffu, wb, o);
            if(ob >= 0 && (wb < 0 || od < wd || (od == wd && S.bidx[ob] < S.bidx[wb]))) { wd = od; wb = ob; }
          }
          if(lane == 0)
          {
            if(wb >= 0 && (best < 0 || wd < bestD || (wd == bestD && S.bidx[wb] < S.bidx[best]))) { bestD = wd; best = wb; }
            finish(ql, q, best, bestD, fmin(sqrt(bestD), (double)R * h * (1.0 - 1e-9)) * (1.0 - 1e-9));
          }
        }
      }
    }
    __syncthreads();
    ICP_STAMP(0)
    // flag exchange: this CTA's results are in everybody's table / everybody's results are in mine
    if(tid < ICP_CLUSTER)
    {
      asm volatile("fence.acq_rel.cluster;" ::: "memory");
      st_release_cluster(dsmem_addr(S.flag + rank, (unsigned)tid), iter + 1u);
      while(ld_acquire_cluster(S.flag + tid) < iter + 1u) {}
    }
    __syncthreads();
    ICP_STAMP(5)
    // ---- B: ReciprocalFilter.cpp:32-78: closest scene point per model point (lowest scene index on ties) ----
    if(tid == 0) S.scan[64] = 0u;  // the list of far queries is empty again
    // The smallest distance per model point is found in two rounds of native 32-bit shared-memory atomics (high word,
    // then low word among those that share it): a 64-bit atomicMin on shared memory is a compare-and-swap loop, which
    // was 16% of what the active warps did.  (Distances are non-negative, so their bit patterns order like the values.)
    unsigned* const bhi = reinterpret_cast<unsigned*>(S.best);
    unsigned* const blo = bhi + nM;
    for(int q = tid; q < nS; q += ICP_THREADS)
    {
      const ulonglong2 e = nndCur[q];
      if(e.y != 0xffffffffffffffffULL) atomicMin(&bhi[e.y], (unsigned)(e.x >> 32));
    }
    __syncthreads();
    for(int q = tid; q < nS; q += ICP_THREADS)
    {
      const ulonglong2 e = nndCur[q];
      if(e.y != 0xffffffffffffffffULL && (unsigned)(e.x >> 32) == bhi[e.y]) atomicMin(&blo[e.y], (unsigned)e.x);
    }
    __syncthreads();
    for(int q = tid; q < nS; q += ICP_THREADS)
    {
      const ulonglong2 e = nndCur[q];
      if(e.y != 0xffffffffffffffffULL && (unsigned)(e.x >> 32) == bhi[e.y] && (unsigned)e.x == blo[e.y])
        atomicMin(&S.win[e.y], (unsigned)q);
    }
    __syncthreads();
    ICP_STAMP(1)
    // DistanceFilter.cpp:62-63
    distSqr *= P.multiplier;
    if(distSqr < P.min_dist_sqr) distSqr = P.min_dist_sqr;

    // ---- C: ClosedFormEstimator2D::setPairs + estimateTransformation (ClosedFormEstimator2D.cpp:36-109) ----
    if(P.trace && rank == 0)
    {
      // the pair list in model order
      unsigned baseCount = 0;
      for(int j = 0; j * ICP_THREADS < nM; j++)
      {
        const int m = tid + j * ICP_THREADS;  // the caller's index
        const unsigned w = (m < nM) ? S.win[S.pos[m]] : 0xffffffffu;
        const bool has = w != 0xffffffffu;
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        if(lane == 0) S.scan[tid >> 5] = __popc(bal);
        __syncthreads();
        unsigned off = baseCount;
        unsigned total = 0;
        for(int w2 = 0; w2 < ICP_THREADS / 32; w2++)
        {
          const unsigned c = S.scan[w2];
          if(w2 < (tid >> 5)) off += c;
          total += c;
        }
        if(has)
        {
          const unsigned pos = off + __popc(bal & ((1u << lane) - 1u));
          if((int)iter < P.max_iterations && pos < (unsigned)P.cap)
          {
            P.tr_model[(size_t)iter * P.cap + pos] = (unsigned)m;
            P.tr_scene[(size_t)iter * P.cap + pos] = w;
          }
        }
        baseCount += total;
        __syncthreads();
      }
    }
    // The sums are latency, not work (a few hundred pairs): ICP_SUM_THREADS threads walk the model points.
    if(tid < ICP_SUM_THREADS)
    {
      double acc[7] = {0, 0, 0, 0, 0, 0, 0};  // sum m0 m1 s0 s1 | r | cross sums about (pm, ps): yx - xy, xx + yy
      unsigned cnt = 0;
      for(int m = tid; m < nM; m += ICP_SUM_THREADS)
      {
        const unsigned sidx = S.win[m];
        S.win[m] = 0xffffffffu;  // for the next iteration (nobody else looks at them before the barriers below)
        S.best[m] = INF64;       // (model points m and nM + m of the two 32-bit arrays the reciprocal filter sees: all ones)
        if(sidx != 0xffffffffu)
        {
          const double m0 = s_mx[m], m1 = s_my[m], s0 = s_sx[sidx], s1 = s_sy[sidx];
          acc[0] += m0; acc[1] += m1;
          acc[2] += s0; acc[3] += s1;
          const double dx = s0 - m0;
          const double dy = s1 - m1;
          acc[4] += dx * dx + dy * dy;
          cnt++;
          const double a0 = m0 - pm0, a1 = m1 - pm1, b0 = s0 - ps0, b1 = s1 - ps1;
          acc[5] += a1 * b0 - a0 * b1;
          acc[6] += a0 * b0 + a1 * b1;
        }
      }
      // the partial sums go through shared memory: a 64-bit shuffle tree over 7 values is 560 instructions per warp
#pragma unroll
      for(int k = 0; k < 7; k++) S.part[k * ICP_SUM_THREADS + tid] = acc[k];
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if(lane == 0) S.scan[32 + (tid >> 5)] = cnt;
    }
    __syncthreads();
    if(tid < 7 * 32)
    {
      // warp k adds up value k: every lane ICP_SUM_THREADS / 32 partials in a fixed order, then one shuffle tree
      const int k = tid >> 5;
      double v = 0.0;
#pragma unroll
      for(int j = 0; j < ICP_SUM_THREADS / 32; j++) v += S.part[k * ICP_SUM_THREADS + j * 32 + lane];
#pragma unroll
      for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if(lane == 0) S.red[k] = v;
    }
    __syncthreads();
    ICP_STAMP(2)
    if(tid < 32)
    {
      double v[7];
#pragma unroll
      for(int k = 0; k < 7; k++) v[k] = S.red[k];
      unsigned np = lane < ICP_SUM_THREADS / 32 ? S.scan[32 + lane] : 0u;
      np = __reduce_add_sync(0xffffffffu, np);
      double* R_ = S.red + 320;  // results: pairs, rms, the new provisional centres
      double* Tl = S.T + 16;
      if(lane == 0) R_[0] = (double)np;
      if(np > 2)
      {
        // (every lane computes the same few numbers; lanes 0-15 then each own one element of the matrices)
        const double sizeInv = 1.0 / (double)np;
        const double r = v[4] * sizeInv;
        const double cm0 = v[0] * sizeInv, cm1 = v[1] * sizeInv, cs0 = v[2] * sizeInv, cs1 = v[3] * sizeInv;
        // sum over the pairs of (m - cm)(s - cs)^T from the sums about the provisional centres
        const double em0 = cm0 - pm0, em1 = cm1 - pm1, es0 = cs0 - ps0, es1 = cs1 - ps1;
        const double n = (double)np;
        const double nominator = v[5] - n * (em1 * es0 - em0 * es1);
        const double denominator = v[6] - n * (em0 * es0 + em1 * es1);
        // cos / sin of deltaTheta = atan2(nominator, denominator) (ClosedFormEstimator2D.cpp:93-96)
        const double hyp = sqrt(nominator * nominator + denominator * denominator);
        double c = 1.0, sn = 0.0;
        if(hyp > 0.0) { c = denominator / hyp; sn = nominator / hyp; }
        const double deltaX = (cm0 - (c * cs0 - sn * cs1));
        const double deltaY = (cm1 - (c * cs1 + sn * cs0));
        if(lane == 0)
        {
          R_[1] = r;
          R_[2] = cm0; R_[3] = cm1; R_[4] = cs0; R_[5] = cs1;
        }
        if(lane < 16)
        {
          double tl = (lane % 5 == 0) ? 1.0 : 0.0;
          if(lane == 0 || lane == 5) tl = c;
          if(lane == 1) tl = -sn;
          if(lane == 4) tl = sn;
          if(lane == 3) tl = deltaX;
          if(lane == 7) tl = deltaY;
          Tl[lane] = tl;
        }
        __syncwarp();
        // Tfinal = Tlast * Tfinal (Icp.cpp:454): element (i, j) accumulates over k in order, zero factors skipped (dgemm)
        double o = 0.0;
        if(lane < 16)
        {
          const int i = lane >> 2, j = lane & 3;
#pragma unroll
          for(int k = 0; k < 4; k++)
          {
            const double temp = 1.0 * Tl[4 * i + k];
            if(temp != 0.0) o += temp * S.T[4 * k + j];
          }
        }
        __syncwarp();
        if(lane < 16) S.T[lane] = o;
      }
    }
    __syncthreads();
    ICP_STAMP(3)
    pairs = (unsigned)S.red[320];
    int retval = TSD_ICP_PROCESSING;
    if(pairs > 2)
    {
      rms = S.red[321];
      pm0 = S.red[322]; pm1 = S.red[323]; ps0 = S.red[324]; ps1 = S.red[325];
      // applyTransformation (Icp.cpp:371-408)
      const double* Tl = S.T + 16;
      const double r00 = Tl[0], r01 = Tl[1], r10 = Tl[4], r11 = Tl[5], t0 = Tl[3], t1 = Tl[7];
      for(int i = tid; i < nS; i += ICP_THREADS)
      {
        const double x = s_sx[i], y = s_sy[i];
        double a = 0.0; a += x * r00; a += y * r01; a = 0.0 + 1.0 * a;
        double b = 0.0; b += x * r10; b += y * r11; b = 0.0 + 1.0 * b;
        const double nx = a + t0, ny = b + t1;
        s_sx[i] = nx;
        s_sy[i] = ny;
        // the point moved by |(nx,ny) - (x,y)|: its nearest-neighbour distance shrank by at most that
        if((unsigned)(i % ICP_CLUSTER) == rank)
        {
          // (|dx| + |dy| bounds the distance moved from above; no square root)
          const double lb = S.lb[i / ICP_CLUSTER] - (fabs(nx - x) + fabs(ny - y)) * (1.0 + 1e-9) - 1e-12;
          S.lb[i / ICP_CLUSTER] = lb > 0.0 ? lb : 0.0;
        }
      }
      {
        // the provisional scene centre moves with the scene
        double a = 0.0; a += ps0 * r00; a += ps1 * r01;
        double b = 0.0; b += ps0 * r10; b += ps1 * r11;
        ps0 = a + t0; ps1 = b + t1;
      }
    }
    else
    {
      retval = TSD_ICP_NOTMATCHABLE;
    }
    if(rank == 0 && tid == 0 && (int)iter < P.max_iterations)
    {
      P.tr_count[iter] = (int)pairs;
      P.tr_mse[iter] = rms;
      for(int i = 0; i < 16; i++) P.tr_T[16 * iter + i] = S.T[i];
    }
    __syncthreads();  // the scene points and lower bounds just written are read by other threads in the next search
    ICP_STAMP(4)
    eRetval = retval;
    // Icp.cpp:496-507
    iter++;
    if(fabs(rms - rms_prev) < 10e-10) conv_cnt++;
    else conv_cnt = 0;
    if((rms <= P.max_rms || conv_cnt >= P.conv_cnt)) eRetval = TSD_ICP_SUCCESS;
    else if(iter >= (unsigned)P.max_iterations) eRetval = TSD_ICP_MAXITERATIONS;
    rms_prev = rms;
  }
  // iterations that did not run have no pair list
  if(rank == 0)
    for(int i = (int)iter + tid; i < P.max_iterations; i += ICP_THREADS) P.tr_count[i] = -1;
  if(rank == 0 && tid == 0)
  {
    // getFinalTransformation (Icp.cpp:528-546)
    P.result[0] = S.T[0]; P.result[1] = S.T[1]; P.result[2] = S.T[3];
    P.result[3] = S.T[4]; P.result[4] = S.T[5]; P.result[5] = S.T[7];
    P.result[6] = 0; P.result[7] = 0; P.result[8] = 1;
    P.result[9] = rms;
    P.result[10] = (double)pairs;
    P.result[11] = (double)iter;
    P.result[12] = (double)eRetval;
    P.result[13] = (double)nM;
#ifdef ICP_PROFILE
    printf("k_icp %u iterations, cycles per iteration: NN %lld | exchange %lld | recip %lld | gather+sums %lld | pose %lld | apply %lld | loop top %lld\n",
           iter, pf[0] / iter, pf[5] / iter, pf[1] / iter, pf[2] / iter, pf[3] / iter, pf[4] / iter, pf[7] / iter);
    printf("   cell walk: seed %lld cycles, %lld rounds per iteration\n", pf[12] / iter, pf[13] / iter);
    printf("   NN: prefilter+coarse %lld | prev+own cell+mask %lld | neighbour cells %lld | far queries %lld | tail+barrier %lld\n", pf[8] / iter,
           pf[9] / iter, pf[10] / iter, pf[11] / iter, pf[0] / iter);
#endif
  }
  cluster.sync();  // no CTA's shared memory goes away while a neighbour may still address it
}

extern "C" {

int icp_create(uint32_t max_iterations, double dist_max, double dist_min, uint32_t dist_iterations,
               const double bounds[4], int device, tsd_icp_t** out)
{
  if(!out || !bounds) return TSD_E_INVALID;
  *out = nullptr;
  int ndev = 0;
  if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
  {
    cudaGetLastError();
    set_error("no CUDA device: libtsdslam_b200 has no CPU path");
    return TSD_E_NO_DEVICE;
  }
  if(device < 0 || device >= ndev) { set_error("invalid device ordinal %d", device); return TSD_E_INVALID; }
  TSD_CUDA(cudaSetDevice(device));
  tsd_icp* h = new tsd_icp();
  memset(h, 0, sizeof(*h));
  h->mtx = new std::recursive_mutex();
  h->device = device;
  IcpParams& p = h->p;
  p.max_iterations = (int)max_iterations;  // ThreadLocalize.cpp:224
  p.conv_cnt = max_iterations;             // :225
  p.max_rms = 0.0;                         // :223
  // DistanceFilter.cpp:11-20
  p.max_dist_sqr = dist_max * dist_max;
  p.min_dist_sqr = dist_min * dist_min;
  double it = (double)(uint32_t)(dist_iterations - 1u);
  if(dist_iterations < 1) it = 1.0;
  p.multiplier = pow((dist_min / dist_max), 1.0 / it);
  p.x_min = bounds[0]; p.x_max = bounds[1]; p.y_min = bounds[2]; p.y_max = bounds[3];
  {
    double hh = fabs(dist_max) / 4.0;
    if(!(hh >= 1e-3)) hh = 1e-3;
    if(hh > 1e6) hh = 1e6;
    p.hash_h = hh;
    p.coarse_h = (fabs(dist_max) > hh ? fabs(dist_max) : hh) * (1.0 + 1e-9);
    double rings = ceil(fabs(dist_max) / hh) + 2.0;
    if(!(rings < 64.0)) rings = 64.0;
    p.max_rings = (int)rings;
  }
  h->cap = ICP_MAX_POINTS;
  TSD_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  TSD_CUDA(cudaMalloc(&h->d_model, sizeof(double) * 4 * h->cap));  // model, then scene
  h->d_scene = nullptr;
  TSD_CUDA(cudaMalloc(&h->d_result, sizeof(double) * 16));
  TSD_CUDA(cudaMalloc(&h->d_nM, sizeof(int)));
  const size_t mi = max_iterations > 0 ? max_iterations : 1;
  h->trace_cap_it = (int)mi;
  TSD_CUDA(cudaMalloc(&h->d_tr_model, sizeof(unsigned) * mi * h->cap));
  TSD_CUDA(cudaMalloc(&h->d_tr_scene, sizeof(unsigned) * mi * h->cap));
  TSD_CUDA(cudaMalloc(&h->d_tr_count, sizeof(int) * mi));
  TSD_CUDA(cudaMalloc(&h->d_tr_mse, sizeof(double) * mi));
  TSD_CUDA(cudaMalloc(&h->d_tr_T, sizeof(double) * 16 * mi));
  TSD_CUDA(cudaMallocHost(&h->h_stage, sizeof(double) * 4 * h->cap));
  TSD_CUDA(cudaMallocHost(&h->h_result, sizeof(double) * 16));
  TSD_CUDA(cudaFuncSetAttribute(k_icp, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)icp_smem_bytes(ICP_MAX_POINTS, ICP_MAX_POINTS)));
  *out = h;
  return TSD_OK;
}

int icp_destroy(tsd_icp_t* h)
{
  if(!h) return TSD_OK;
  cudaSetDevice(h->device);
  if(h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_model); cudaFree(h->d_scene); cudaFree(h->d_result); cudaFree(h->d_nM); cudaFree(h->d_tr_model); cudaFree(h->d_tr_scene);
  cudaFree(h->d_tr_count); cudaFree(h->d_tr_mse); cudaFree(h->d_tr_T);
  cudaFreeHost(h->h_stage); cudaFreeHost(h->h_result);
  if(h->stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h->mtx;
  delete h;
  return TSD_OK;
}

int icp_run(tsd_icp_t* h, const double* model, const double* normals, int32_t n_model, const double* scene,
            int32_t n_scene, const double pose[9], const double* t_init, double t_out[9], double* mse,
            uint32_t* pairs, uint32_t* iterations, int32_t* state)
{
  TSD_LOCK(h);
  (void)normals;  // ClosedFormEstimator2D ignores normals (ClosedFormEstimator2D.cpp:26-34)
  if(!h || !pose || !t_out || !mse || !pairs || !iterations || !state) return TSD_E_INVALID;
  for(int i = 0; i < 9; i++) t_out[i] = (i % 4 == 0) ? 1.0 : 0.0;
  *mse = 0.0; *pairs = 0; *iterations = 0;
  h->last_nM = h->last_nS = 0;
  // Icp.cpp:467-471
  if(n_model <= 0 || n_scene <= 0) { *state = TSD_ICP_NOTMATCHABLE; return TSD_OK; }
  if(!model || !scene) return TSD_E_INVALID;
  if(n_model > ICP_MAX_POINTS || n_scene > ICP_MAX_POINTS)
  {
    set_error("icp_run supports at most %d model and scene points", ICP_MAX_POINTS);
    return TSD_E_INVALID;
  }
  TSD_CUDA(cudaSetDevice(h->device));
  TSD_CUDA(cudaStreamSynchronize(h->stream));
  // model and scene travel in one copy
  memcpy(h->h_stage, model, sizeof(double) * 2 * n_model);
  memcpy(h->h_stage + 2 * n_model, scene, sizeof(double) * 2 * n_scene);
  TSD_CUDA(cudaMemcpyAsync(h->d_model, h->h_stage, sizeof(double) * 2 * ((size_t)n_model + n_scene), cudaMemcpyHostToDevice, h->stream));
  IcpParams p = h->p;
  p.nM = n_model;
  p.nM_dev = nullptr;
  p.nS = n_scene;
  for(int i = 0; i < 9; i++) p.pose[i] = pose[i];
  p.has_init = t_init ? 1 : 0;
  for(int i = 0; i < 16; i++) p.t_init[i] = t_init ? t_init[i] : ((i % 5 == 0) ? 1.0 : 0.0);
  p.model = h->d_model;
  p.scene = h->d_model + 2 * (size_t)n_model;
  p.result = h->d_result;
  p.cap = h->cap;
  p.trace = h->trace;
  p.tr_model = h->d_tr_model;
  p.tr_scene = h->d_tr_scene;
  p.tr_count = h->d_tr_count;
  p.tr_mse = h->d_tr_mse;
  p.tr_T = h->d_tr_T;
  k_icp<<<ICP_CLUSTER, ICP_THREADS, icp_smem_bytes(n_model, n_scene), h->stream>>>(p);  // one cluster (__cluster_dims__)
  TSD_LAUNCHED();
  TSD_CUDA(cudaMemcpyAsync(h->h_result, h->d_result, sizeof(double) * 13, cudaMemcpyDeviceToHost, h->stream));
  TSD_CUDA(cudaStreamSynchronize(h->stream));
  for(int i = 0; i < 9; i++) t_out[i] = h->h_result[i];
  *mse = h->h_result[9];
  *pairs = (uint32_t)h->h_result[10];
  *iterations = (uint32_t)h->h_result[11];
  *state = (int32_t)h->h_result[12];
  h->last_nM = n_model;
  h->last_nS = n_scene;
  return TSD_OK;
}

// ThreadLocalize::maskMatrix (ThreadLocalize.cpp:738-755) on the device: the ray caster's hits, in beam order, become the
// ICP model.  One block; n <= ICP_MAX_POINTS beams.
__global__ void __launch_bounds__(1024) k_icp_model_from_raycast(const double* __restrict__ out4, const unsigned long long* __restrict__ keys,
                                                                 int n, double* __restrict__ model, int* __restrict__ n_model)
{
  __shared__ unsigned s_cnt[32];
  const int tid = threadIdx.x, lane = tid & 31;
  unsigned base = 0;
  for(int b0 = 0; b0 < n; b0 += 1024)
  {
    const int b = b0 + tid;
    const unsigned long long k = b < n ? keys[b] : 0x7fffffffffffffffULL;
    const bool hit = k != 0x7fffffffffffffffULL && (k & 3ULL) == 0ULL;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if(lane == 0) s_cnt[tid >> 5] = __popc(bal);
    __syncthreads();
    unsigned off = base, total = 0;
    for(int w = 0; w < 32; w++)
    {
      const unsigned c = s_cnt[w];
      if(w < (tid >> 5)) off += c;
      total += c;
    }
    if(hit)
    {
      const unsigned pos = off + __popc(bal & ((1u << lane) - 1u));
      model[2 * pos] = out4[4 * b];
      model[2 * pos + 1] = out4[4 * b + 1];
    }
    base += total;
    __syncthreads();
  }
  if(tid == 0) *n_model = (int)base;
}

// One localisation step without the model ever visiting the host: RayCastPolar2D::calcCoordsFromCurrentViewMask, maskMatrix
// and Icp::iterate (ThreadLocalize.cpp:333-361, :571-581) as three launches on the grid's stream -- ray cast, compaction
// of the hits into the model, k_icp reading the model's size on the device -- with one upload (scan, ray directions,
// scene) before and one download (pose, mse, pairs, iterations, state, model size) after.
int tsdg_localize(tsd_grid_t* g, tsd_icp_t* h, const tsd_scan_t* scan, const double* rays_world, const double* scene,
                  int32_t n_scene, const double* t_init, double t_out[9], double* mse, uint32_t* pairs, uint32_t* iterations,
                  int32_t* state, uint32_t* n_model)
{
  TSD_LOCK(g);
  std::unique_lock<std::recursive_mutex> lk2;
  if(h) lk2 = std::unique_lock<std::recursive_mutex>(*h->mtx);
  if(!g || !h || !scan || !rays_world || !t_out || !mse || !pairs || !iterations || !state || !n_model) return TSD_E_INVALID;
  for(int i = 0; i < 9; i++) t_out[i] = (i % 4 == 0) ? 1.0 : 0.0;
  *mse = 0.0; *pairs = 0; *iterations = 0; *n_model = 0;
  if(g->device != h->device) { set_error("tsdg_localize: grid and icp handle live on different devices"); return TSD_E_INVALID; }
  if(scan->n > ICP_MAX_POINTS || n_scene > ICP_MAX_POINTS)
  {
    set_error("tsdg_localize supports at most %d beams and scene points", ICP_MAX_POINTS);
    return TSD_E_INVALID;
  }
  if(n_scene > 0 && !scene) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  TSD_CUDA(cudaStreamSynchronize(h->stream));  // (an icp_run of another thread's making still owns the staging buffer)
  int rc = tsd_raycast_enqueue(g, scan, rays_world);
  if(rc) return rc;
  if(n_scene <= 0)
  {
    // Icp.cpp:467-471; the model's size is still reported
    k_icp_model_from_raycast<<<1, 1024, 0, g->stream>>>(g->d_rc_out, g->d_rc_keys, scan->n, h->d_model, h->d_nM);
    TSD_LAUNCHED();
    int nm = 0;
    TSD_CUDA(cudaMemcpyAsync(&nm, h->d_nM, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    TSD_CUDA(cudaStreamSynchronize(g->stream));
    *n_model = (uint32_t)nm;
    *state = TSD_ICP_NOTMATCHABLE;
    return TSD_OK;
  }
  double* d_scene = h->d_model + 2 * (size_t)h->cap;
  memcpy(h->h_stage, scene, sizeof(double) * 2 * n_scene);
  TSD_CUDA(cudaMemcpyAsync(d_scene, h->h_stage, sizeof(double) * 2 * n_scene, cudaMemcpyHostToDevice, g->stream));
  k_icp_model_from_raycast<<<1, 1024, 0, g->stream>>>(g->d_rc_out, g->d_rc_keys, scan->n, h->d_model, h->d_nM);
  TSD_LAUNCHED();
  IcpParams p = h->p;
  p.nM = scan->n;  // upper bound (shared memory is sized for it)
  p.nM_dev = h->d_nM;
  p.nS = n_scene;
  for(int i = 0; i < 9; i++) p.pose[i] = scan->pose[i];
  p.has_init = t_init ? 1 : 0;
  for(int i = 0; i < 16; i++) p.t_init[i] = t_init ? t_init[i] : ((i % 5 == 0) ? 1.0 : 0.0);
  p.model = h->d_model;
  p.scene = d_scene;
  p.result = h->d_result;
  p.cap = h->cap;
  p.trace = h->trace;
  p.tr_model = h->d_tr_model;
  p.tr_scene = h->d_tr_scene;
  p.tr_count = h->d_tr_count;
  p.tr_mse = h->d_tr_mse;
  p.tr_T = h->d_tr_T;
  k_icp<<<ICP_CLUSTER, ICP_THREADS, icp_smem_bytes(scan->n, n_scene), g->stream>>>(p);
  TSD_LAUNCHED();
  TSD_CUDA(cudaMemcpyAsync(h->h_result, h->d_result, sizeof(double) * 14, cudaMemcpyDeviceToHost, g->stream));
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  for(int i = 0; i < 9; i++) t_out[i] = h->h_result[i];
  *mse = h->h_result[9];
  *pairs = (uint32_t)h->h_result[10];
  *iterations = (uint32_t)h->h_result[11];
  *state = (int32_t)h->h_result[12];
  *n_model = (uint32_t)h->h_result[13];
  h->last_nM = (int)*n_model;
  h->last_nS = n_scene;
  return TSD_OK;
}

// PairAssignment::determinePairs on its own (assign/PairAssignment.cpp:38-84 with the node's filters: OutOfBoundsFilter2D
// before, FlannPairAssignment::determinePairsSequential (FlannPairAssignment.cpp:64-92), DistanceFilter and
// ReciprocalFilter after): the pair list of ONE pass over the scene as given -- the first step of Icp::iterate with an
// identity initial guess and freshly reset filters.  Pairs come in model-index order (ReciprocalFilter.cpp:32-78).
int icp_pairs(tsd_icp_t* h, const double* model, int32_t n_model, const double* scene, int32_t n_scene, const double pose[9],
              uint32_t* pair_model, uint32_t* pair_scene, double* dist_sqr, uint32_t* n_pairs)
{
  TSD_LOCK(h);
  if(!h || !pair_model || !pair_scene || !n_pairs || !pose) return TSD_E_INVALID;
  *n_pairs = 0;
  if(h->trace_cap_it < 1) { set_error("icp_pairs needs a handle created for at least one iteration"); return TSD_E_INVALID; }
  const int keepIt = h->p.max_iterations, keepTrace = h->trace;
  h->p.max_iterations = 1;
  h->trace = 1;
  double T[9], mse;
  uint32_t pairs = 0, its = 0;
  int32_t state = 0;
  int rc = icp_run(h, model, nullptr, n_model, scene, n_scene, pose, nullptr, T, &mse, &pairs, &its, &state);
  h->p.max_iterations = keepIt;
  h->trace = keepTrace;
  if(rc) return rc;
  if(state == TSD_ICP_NOTMATCHABLE) return TSD_OK;  // empty model or scene: no pairs
  int cnt = -1;
  TSD_CUDA(cudaMemcpy(&cnt, h->d_tr_count, sizeof(int), cudaMemcpyDeviceToHost));
  if(cnt <= 0) return TSD_OK;
  TSD_CUDA(cudaMemcpy(pair_model, h->d_tr_model, sizeof(unsigned) * cnt, cudaMemcpyDeviceToHost));
  TSD_CUDA(cudaMemcpy(pair_scene, h->d_tr_scene, sizeof(unsigned) * cnt, cudaMemcpyDeviceToHost));
  if(dist_sqr)
    for(int i = 0; i < cnt; i++)
    {
      // flann::L2: result += diff * diff, dimension by dimension
      const double d0 = scene[2 * pair_scene[i]] - model[2 * pair_model[i]];
      const double d1 = scene[2 * pair_scene[i] + 1] - model[2 * pair_model[i] + 1];
      double d = 0.0;
      d += d0 * d0;
      d += d1 * d1;
      dist_sqr[i] = d;
    }
  *n_pairs = (uint32_t)cnt;
  return TSD_OK;
}

int icp_set_termination(tsd_icp_t* h, double max_rms, uint32_t convergence_counter)
{
  TSD_LOCK(h);
  if(!h) return TSD_E_INVALID;
  h->p.max_rms = max_rms;
  h->p.conv_cnt = convergence_counter;
  return TSD_OK;
}

int icp_set_max_iterations(tsd_icp_t* h, uint32_t max_iterations)
{
  TSD_LOCK(h);
  if(!h || max_iterations > (uint32_t)h->trace_cap_it) { set_error("icp_set_max_iterations: beyond the capacity given to icp_create"); return TSD_E_INVALID; }
  h->p.max_iterations = (int)max_iterations;
  return TSD_OK;
}

int icp_set_trace(tsd_icp_t* h, int enable)
{
  TSD_LOCK(h);
  if(!h) return TSD_E_INVALID;
  h->trace = enable != 0;
  return TSD_OK;
}

int icp_get_trace(tsd_icp_t* h, int32_t max_it, int32_t cap, uint32_t* pair_model, uint32_t* pair_scene,
                  int32_t* pair_count, double* mse, double* t_final16, int32_t* n_it)
{
  TSD_LOCK(h);
  if(!h || !pair_model || !pair_scene || !pair_count || !mse || !t_final16 || !n_it) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(h->device));
  TSD_CUDA(cudaStreamSynchronize(h->stream));
  const int mi = h->p.max_iterations < max_it ? h->p.max_iterations : max_it;
  std::vector<int> cnt(mi > 0 ? mi : 1);
  *n_it = 0;
  if(mi <= 0) return TSD_OK;
  TSD_CUDA(cudaMemcpy(cnt.data(), h->d_tr_count, sizeof(int) * mi, cudaMemcpyDeviceToHost));
  std::vector<unsigned> row(h->cap);
  int its = 0;
  for(int it = 0; it < mi; it++)
  {
    if(cnt[it] < 0) break;
    its++;
    pair_count[it] = cnt[it];
    const int n = cnt[it] < cap ? cnt[it] : cap;
    TSD_CUDA(cudaMemcpy(row.data(), h->d_tr_model + (size_t)it * h->cap, sizeof(unsigned) * n, cudaMemcpyDeviceToHost));
    memcpy(pair_model + (size_t)it * cap, row.data(), sizeof(unsigned) * n);
    TSD_CUDA(cudaMemcpy(row.data(), h->d_tr_scene + (size_t)it * h->cap, sizeof(unsigned) * n, cudaMemcpyDeviceToHost));
    memcpy(pair_scene + (size_t)it * cap, row.data(), sizeof(unsigned) * n);
  }
  TSD_CUDA(cudaMemcpy(mse, h->d_tr_mse, sizeof(double) * its, cudaMemcpyDeviceToHost));
  TSD_CUDA(cudaMemcpy(t_final16, h->d_tr_T, sizeof(double) * 16 * its, cudaMemcpyDeviceToHost));
  *n_it = its;
  return TSD_OK;
}

}  // extern "C"
