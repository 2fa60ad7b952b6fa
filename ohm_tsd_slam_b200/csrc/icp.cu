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
// Sums of the estimator are block reductions with a fixed tree, so results are deterministic but not
// bit-identical to the reference's sequential sums (and atan2/sin/cos differ from glibc in the last ulp
// anyway): pair lists are compared exactly, poses to 1e-9 (tests/test_icp_gpu.py).
#include <string.h>

#include <vector>

#include <cooperative_groups.h>

#include "common.cuh"

using namespace tsd;

#define ICP_THREADS 1024
#define ICP_MAX_POINTS 2048
#define ICP_SLOTS 4096  // hash buckets
#ifndef ICP_CLUSTER
#define ICP_CLUSTER 8   // CTAs (SMs) per registration: the portable cluster size
#endif

struct IcpParams
{
  int nM, nS;
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
  int trace;
  int trace_cap_it;
};

// cell coordinate of a point; clamped so that the hash input stays small and rings never overflow
__device__ __forceinline__ int cell_of(double v, double v0, double invh)
{
  const double t = floor((v - v0) * invh);
  return (int)fmin(fmax(t, -1.0e6), 1.0e6);
}

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
// One thread-block CLUSTER of ICP_CLUSTER CTAs (one SM each) runs the whole loop.  Every CTA keeps the model, its
// search structure, the scene and the transformation in its own shared memory and evolves them identically
// (same code, same data, same reduction trees), so no state is ever broadcast.  What is split is the expensive
// part, the nearest-neighbour search: query i belongs to CTA i % ICP_CLUSTER, and four lanes share a query (the
// cells of a ring are dealt round-robin to the lanes, then two shuffles pick the winner).  The reciprocal filter
// (closest scene point per model point) is reduced in two levels: every CTA filters its own queries with
// shared-memory atomics and stores each local winner into the inbox of model point m's host, CTA m % ICP_CLUSTER
// (distributed shared memory, one 16-byte store); the host keeps the best of its ICP_CLUSTER candidates and stores
// the final winner into every CTA's copy.  Two cluster barriers per iteration:
//      NN search, local reciprocal filter, local winners -> hosts' inboxes   | cluster.sync |
//      hosts pick (best distance, lowest scene index), winners -> every CTA   | cluster.sync |
//      estimator sums over the winners, pose update, scene transform -- replicated, identical in every CTA
__global__ void __cluster_dims__(ICP_CLUSTER, 1, 1) __launch_bounds__(ICP_THREADS, 1) k_icp(IcpParams P)
{
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x;
  const int nM = P.nM, nS = P.nS;
  const int nQ = (nS + ICP_CLUSTER - 1) / ICP_CLUSTER;  // queries per CTA (upper bound)
  const int nH = (nM + ICP_CLUSTER - 1) / ICP_CLUSTER;  // model points hosted per CTA (upper bound)
  // shared memory carve-up (identical in every CTA, so that map_shared_rank offsets agree)
  double* s_mx = reinterpret_cast<double*>(smem);
  double* s_my = s_mx + nM;
  double* s_sx = s_my + nM;
  double* s_sy = s_sx + nS;
  double* s_d2 = s_sy + nS;                                                        // nQ: own queries
  double* s_lb = s_d2 + nQ;                                                        // nQ: lower bound of the NN distance
  unsigned long long* s_best = reinterpret_cast<unsigned long long*>(s_lb + nQ);  // nM: over this CTA's queries
  double* s_red = reinterpret_cast<double*>(s_best + nM);                          // 208
  double* s_T = s_red + 208;                                                       // Tfinal 16, Tlast 16
  unsigned* s_win = reinterpret_cast<unsigned*>(s_T + 32);                         // nM: over this CTA's queries
  unsigned* s_fin = s_win + nM;                                                    // nM: final winner of every model point
  int* s_nn = reinterpret_cast<int*>(s_fin + nM);                                  // nQ
  unsigned* s_scan = reinterpret_cast<unsigned*>(s_nn + nQ);                       // 40
  unsigned short* s_bstart = reinterpret_cast<unsigned short*>(s_scan + 40);       // ICP_SLOTS + 2
  unsigned short* s_bcnt = s_bstart + (ICP_SLOTS + 2);                             // ICP_SLOTS
  unsigned short* s_bidx = s_bcnt + ICP_SLOTS;                                     // nM (+1 pad)
  unsigned* s_coarse = reinterpret_cast<unsigned*>(s_bidx + ((nM + 2) & ~1));      // 128 words: coarse occupancy bitmap
  unsigned* s_occ = s_coarse + 128;                                                // 128 words: non-empty hash slots
  // candidates sent to this host: ICP_CLUSTER x nH entries {distance bits, scene index}, one 16-byte store each
  ulonglong2* s_inbox = reinterpret_cast<ulonglong2*>((reinterpret_cast<uintptr_t>(s_occ + 128) + 15) & ~(uintptr_t)15);

  const unsigned long long INF64 = 0xffffffffffffffffULL;
  for(int i = tid; i < nM; i += ICP_THREADS) { s_mx[i] = P.model[2 * i]; s_my[i] = P.model[2 * i + 1]; }
  for(int i = tid; i < nS; i += ICP_THREADS) { s_sx[i] = P.scene[2 * i]; s_sy[i] = P.scene[2 * i + 1]; }
  for(int i = tid; i < nQ; i += ICP_THREADS) s_lb[i] = 0.0;
  for(int i = tid; i < ICP_CLUSTER * nH; i += ICP_THREADS) s_inbox[i] = make_ulonglong2(INF64, 0xffffffffULL);
  for(int i = tid; i < ICP_SLOTS; i += ICP_THREADS) s_bcnt[i] = 0;
  if(tid < 128) { s_coarse[tid] = 0u; s_occ[tid] = 0u; }
  if(tid < 16) { s_T[tid] = (tid % 5 == 0) ? 1.0 : 0.0; s_T[16 + tid] = s_T[tid]; }
  __syncthreads();

  // ---- spatial hash of the model: counting sort of the points by hash slot (built by every CTA for itself) ----
  const double h = P.hash_h, invh = 1.0 / P.hash_h;
  const double invhc = 1.0 / P.coarse_h;  // coarse cells: edge >= the distance filter's largest threshold
  const double bx0 = s_mx[0], by0 = s_my[0];
  for(int i = tid; i < nM; i += ICP_THREADS)
  {
    const unsigned b = slot_of(cell_of(s_mx[i], bx0, invh), cell_of(s_my[i], by0, invh));
    atomicAdd(reinterpret_cast<unsigned*>(s_bcnt) + (b >> 1), (b & 1) ? 0x10000u : 1u);  // u16 counters, nM <= 2048
    const unsigned c = slot_of(cell_of(s_mx[i], bx0, invhc), cell_of(s_my[i], by0, invhc));
    atomicOr(&s_coarse[c >> 5], 1u << (c & 31));
    atomicOr(&s_occ[b >> 5], 1u << (b & 31));
  }
  __syncthreads();
  {
    // exclusive prefix over ICP_SLOTS = 4 * ICP_THREADS counters
    const unsigned c0 = s_bcnt[4 * tid], c1 = s_bcnt[4 * tid + 1], c2 = s_bcnt[4 * tid + 2], c3 = s_bcnt[4 * tid + 3];
    const unsigned mine = c0 + c1 + c2 + c3;
    unsigned incl = mine;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if((tid & 31) >= o) incl += t;
    }
    if((tid & 31) == 31) s_scan[tid >> 5] = incl;
    __syncthreads();
    unsigned off = incl - mine;
    for(int w = 0; w < (tid >> 5); w++) off += s_scan[w];
    s_bstart[4 * tid] = (unsigned short)off;
    s_bstart[4 * tid + 1] = (unsigned short)(off + c0);
    s_bstart[4 * tid + 2] = (unsigned short)(off + c0 + c1);
    s_bstart[4 * tid + 3] = (unsigned short)(off + c0 + c1 + c2);
    if(tid == ICP_THREADS - 1) s_bstart[ICP_SLOTS] = (unsigned short)(off + mine);
    __syncthreads();
    s_bcnt[4 * tid] = 0; s_bcnt[4 * tid + 1] = 0; s_bcnt[4 * tid + 2] = 0; s_bcnt[4 * tid + 3] = 0;
    __syncthreads();
  }
  for(int i = tid; i < nM; i += ICP_THREADS)
  {
    const unsigned b = slot_of(cell_of(s_mx[i], bx0, invh), cell_of(s_my[i], by0, invh));
    const unsigned old = atomicAdd(reinterpret_cast<unsigned*>(s_bcnt) + (b >> 1), (b & 1) ? 0x10000u : 1u);
    const unsigned within = (b & 1) ? (old >> 16) : (old & 0xffffu);
    s_bidx[s_bstart[b] + within] = (unsigned short)i;
  }
  __syncthreads();

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
      // Tfinal = Tinit * Tfinal(identity), dgemm NoTrans x NoTrans with zero skipping
      double out[16];
      for(int i = 0; i < 16; i++) out[i] = 0.0;
      for(int k = 0; k < 4; k++)
        for(int i = 0; i < 4; i++)
        {
          const double temp = 1.0 * P.t_init[4 * i + k];
          if(temp != 0.0)
            for(int j = 0; j < 4; j++) out[4 * i + j] += temp * s_T[4 * k + j];
        }
      for(int i = 0; i < 16; i++) s_T[i] = out[i];
    }
    __syncthreads();
  }
  cluster.sync();  // every CTA is resident before anybody addresses its shared memory

  int eRetval = TSD_ICP_PROCESSING;
  unsigned iter = 0;
  double rms_prev = 10e12;
  unsigned conv_cnt = 0;
  double rms = 0.0;  // the caller passes *rms = 0.0 (ThreadLocalize.cpp:577)
  unsigned pairs = 0;
  double distSqr = P.max_dist_sqr;  // DistanceFilter::reset (DistanceFilter.cpp:27-30)

  // four lanes per query
  const int quad = tid >> 2, ql = tid & 3;
  const unsigned qmask = 0xfu << ((tid & 31) & ~3);

  while(eRetval == TSD_ICP_PROCESSING)
  {
    for(int m = tid; m < nM; m += ICP_THREADS) { s_best[m] = INF64; s_win[m] = 0xffffffffu; }
    __syncthreads();

    // ---- A: pre-filter + exact 1-NN + distance filter, for the queries of this CTA ----
    // Four lanes per query scan the 3x3 cells around it (one pass for all queries: nQ <= ICP_THREADS / 4).  That
    // settles every query with a model point within one cell edge.  The others are finished one after the other
    // by their whole warp: the 32 lanes scan the rest of the (2R+1)^2 window that covers the distance filter's
    // current radius, so that a few far-off points do not hold up the cluster.
    {
      const int q = quad;
      const int i = q * ICP_CLUSTER + (int)rank;
      const bool exists = q < nQ && i < nS;
      const int lane = tid & 31;
      const double x = exists ? s_sx[i] : 0.0, y = exists ? s_sy[i] : 0.0;
      // OutOfBoundsFilter2D.cpp:27-37: S.transform(pose) = S * R^T + t
      double tx = 0.0; tx += x * P.pose[0]; tx += y * P.pose[1]; tx = 0.0 + 1.0 * tx; tx += P.pose[2];
      double ty = 0.0; ty += x * P.pose[3]; ty += y * P.pose[4]; ty = 0.0 + 1.0 * ty; ty += P.pose[5];
      bool search = exists && !(tx < P.x_min || tx > P.x_max || ty < P.y_min || ty > P.y_max);
      int best = -1;
      double bestD = __longlong_as_double(0x7ff0000000000000LL);
      // A point whose nearest model point is provably farther than the distance filter's threshold cannot
      // yield a pair (DistanceFilter.cpp:38): its search is skipped.  s_lb[q] is a lower bound of that
      // distance, carried over from the last search and reduced by how far the point moved since.
      double lbNew = exists ? s_lb[q] : 0.0;
      if(search && lbNew * lbNew > distSqr) search = false;
      else if(search) lbNew = 0.0;
      if(search)
      {
        // a model point within the distance filter's radius lies in the 3x3 coarse cells around the query
        const int cqx = cell_of(x, bx0, invhc), cqy = cell_of(y, by0, invhc);
        unsigned any = 0;
#pragma unroll
        for(int dy = -1; dy <= 1; dy++)
#pragma unroll
          for(int dx = -1; dx <= 1; dx++)
          {
            const unsigned c = slot_of(cqx + dx, cqy + dy);
            any |= (s_coarse[c >> 5] >> (c & 31)) & 1u;
          }
        search = any != 0;
        if(!search) lbNew = P.coarse_h * (1.0 - 1e-6);  // nothing within one coarse cell
      }
      // (search, x, y are uniform over the four lanes of the query)
      const int qx = cell_of(x, bx0, invh), qy = cell_of(y, by0, invh);
      bool pending = false;
      if(search)
      {
        for(int c = ql; c < 9; c += 4)
        {
          const unsigned b = slot_of(qx - 1 + c % 3, qy - 1 + c / 3);
          if(!((s_occ[b >> 5] >> (b & 31)) & 1u)) continue;
          const int k1 = s_bstart[b + 1];
          for(int k = s_bstart[b]; k < k1; k++)
          {
            const int m = s_bidx[k];
            const double d0 = x - s_mx[m];
            const double d1 = y - s_my[m];
            double d = 0.0;
            d += d0 * d0;
            d += d1 * d1;
            if(d < bestD || (d == bestD && m < best)) { bestD = d; best = m; }
          }
        }
#pragma unroll
        for(int o = 1; o < 4; o <<= 1)
        {
          const double od = __shfl_xor_sync(qmask, bestD, o);
          const int ob = __shfl_xor_sync(qmask, best, o);
          if(ob >= 0 && (best < 0 || od < bestD || (od == bestD && ob < best))) { bestD = od; best = ob; }
        }
        // everything unvisited lies in cells at Chebyshev distance > 1, i.e. farther than h
        const double lb = h * (1.0 - 1e-9);
        const double lb2 = lb * lb;
        if(bestD < lb2 || distSqr < lb2) lbNew = fmin(sqrt(bestD), lb) * (1.0 - 1e-9);  // shaved against rounding
        else pending = true;
      }
      // smallest R with (R h)^2 > distSqr: the window then holds every point the distance filter can keep
      int R = 2;
      while(R < P.max_rings && !(distSqr < ((double)R * h * (1.0 - 1e-9)) * ((double)R * h * (1.0 - 1e-9)))) R++;
      const int W = 2 * R + 1;
      unsigned todo = __ballot_sync(0xffffffffu, pending && ql == 0);
      while(todo)
      {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const double px = __shfl_sync(0xffffffffu, x, src), py = __shfl_sync(0xffffffffu, y, src);
        const int pqx = __shfl_sync(0xffffffffu, qx, src), pqy = __shfl_sync(0xffffffffu, qy, src);
        double wd = __longlong_as_double(0x7ff0000000000000LL);
        int wb = -1;
        for(int c = lane; c < W * W; c += 32)
        {
          const int dx = c % W - R, dy = c / W - R;
          if(dx >= -1 && dx <= 1 && dy >= -1 && dy <= 1) continue;  // done above
          const unsigned b = slot_of(pqx + dx, pqy + dy);
          if(!((s_occ[b >> 5] >> (b & 31)) & 1u)) continue;
          const int k1 = s_bstart[b + 1];
          for(int k = s_bstart[b]; k < k1; k++)
          {
            const int m = s_bidx[k];
            const double d0 = px - s_mx[m];
            const double d1 = py - s_my[m];
            double d = 0.0;
            d += d0 * d0;
            d += d1 * d1;
            if(d < wd || (d == wd && m < wb)) { wd = d; wb = m; }
          }
        }
#pragma unroll
        for(int o = 16; o > 0; o >>= 1)
        {
          const double od = __shfl_xor_sync(0xffffffffu, wd, o);
          const int ob = __shfl_xor_sync(0xffffffffu, wb, o);
          if(ob >= 0 && (wb < 0 || od < wd || (od == wd && ob < wb))) { wd = od; wb = ob; }
        }
        if((lane & ~3) == src)
        {
          if(wb >= 0 && (best < 0 || wd < bestD || (wd == bestD && wb < best))) { bestD = wd; best = wb; }
          lbNew = fmin(sqrt(bestD), (double)R * h * (1.0 - 1e-9)) * (1.0 - 1e-9);
        }
      }
      if(exists && ql == 0)
      {
        s_lb[q] = lbNew;
        const bool keep = (best >= 0) && (bestD <= distSqr);  // DistanceFilter.cpp:38
        s_nn[q] = keep ? best : -1;
        s_d2[q] = bestD;
        if(keep) atomicMin(&s_best[best], (unsigned long long)__double_as_longlong(bestD));
      }
    }
    __syncthreads();
    // ---- B: ReciprocalFilter.cpp:32-78: closest scene point per model point (lowest scene index on ties) ----
    // level 1: among this CTA's queries
    for(int q = tid; q < nQ; q += ICP_THREADS)
    {
      const int i = q * ICP_CLUSTER + (int)rank;
      if(i >= nS) continue;
      const int m = s_nn[q];
      if(m >= 0 && (unsigned long long)__double_as_longlong(s_d2[q]) == s_best[m]) atomicMin(&s_win[m], (unsigned)i);
    }
    __syncthreads();
    // the local winner of model point m goes to m's host: slot [this CTA][m / ICP_CLUSTER] of its inbox
    for(int q = tid; q < nQ; q += ICP_THREADS)
    {
      const int i = q * ICP_CLUSTER + (int)rank;
      if(i >= nS) continue;
      const int m = s_nn[q];
      if(m >= 0 && s_win[m] == (unsigned)i)
      {
        const unsigned host = m % ICP_CLUSTER, slot = rank * nH + m / ICP_CLUSTER;
        dsmem_st_v2u64(dsmem_addr(s_inbox + slot, host), (unsigned long long)__double_as_longlong(s_d2[q]), (unsigned long long)i);
      }
    }
    cluster.sync();
    // level 2: the host of model point m = k * ICP_CLUSTER + rank picks the winner among the ICP_CLUSTER candidates
    // (lane l of a group looks at CTA l's), clears its inbox and tells every CTA (lane l tells CTA l)
    for(int k0 = 0; k0 < nH; k0 += ICP_THREADS / ICP_CLUSTER)
    {
      const int k = k0 + tid / ICP_CLUSTER;
      const unsigned peer = tid % ICP_CLUSTER;
      const int m = k * ICP_CLUSTER + (int)rank;
      const bool hosted = k < nH && m < nM;
      unsigned long long bd = INF64;
      unsigned wi = 0xffffffffu;
      if(hosted)
      {
        const ulonglong2 e = s_inbox[peer * nH + k];
        bd = e.x;
        wi = (unsigned)e.y;
        s_inbox[peer * nH + k] = make_ulonglong2(INF64, 0xffffffffULL);
      }
#pragma unroll
      for(int o = 1; o < ICP_CLUSTER; o <<= 1)
      {
        const unsigned long long ob = __shfl_xor_sync(0xffffffffu, bd, o);
        const unsigned ow = __shfl_xor_sync(0xffffffffu, wi, o);
        if(ob < bd || (ob == bd && ow < wi)) { bd = ob; wi = ow; }
      }
      if(hosted) dsmem_st_u32(dsmem_addr(s_fin + m, peer), (bd == INF64) ? 0xffffffffu : wi);
    }
    cluster.sync();
    // DistanceFilter.cpp:62-63
    distSqr *= P.multiplier;
    if(distSqr < P.min_dist_sqr) distSqr = P.min_dist_sqr;

    // ---- C: ClosedFormEstimator2D::setPairs (+ the pair list in model order when tracing) ----
    // (replicated: every CTA holds all winners)
    double acc[6] = {0, 0, 0, 0, 0, 0};  // cm0 cm1 cs0 cs1 r count
    unsigned winOf[(ICP_MAX_POINTS + ICP_THREADS - 1) / ICP_THREADS];
#pragma unroll
    for(int j = 0; j < (ICP_MAX_POINTS + ICP_THREADS - 1) / ICP_THREADS; j++)
    {
      const int m = tid + j * ICP_THREADS;
      winOf[j] = (m < nM) ? s_fin[m] : 0xffffffffu;
    }
    if(P.trace && rank == 0)
    {
      unsigned baseCount = 0;
#pragma unroll
      for(int j = 0; j < (ICP_MAX_POINTS + ICP_THREADS - 1) / ICP_THREADS; j++)
      {
        const int m = tid + j * ICP_THREADS;
        if(j * ICP_THREADS >= nM) break;
        const bool has = winOf[j] != 0xffffffffu;
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        if((tid & 31) == 0) s_scan[tid >> 5] = __popc(bal);
        __syncthreads();
        unsigned off = baseCount;
        unsigned total = 0;
        for(int w = 0; w < ICP_THREADS / 32; w++)
        {
          const unsigned c = s_scan[w];
          if(w < (tid >> 5)) off += c;
          total += c;
        }
        if(has)
        {
          const unsigned pos = off + __popc(bal & ((1u << (tid & 31)) - 1u));
          if((int)iter < P.max_iterations && pos < (unsigned)P.cap)
          {
            P.tr_model[(size_t)iter * P.cap + pos] = (unsigned)m;
            P.tr_scene[(size_t)iter * P.cap + pos] = winOf[j];
          }
        }
        baseCount += total;
        __syncthreads();
      }
    }
#pragma unroll
    for(int j = 0; j < (ICP_MAX_POINTS + ICP_THREADS - 1) / ICP_THREADS; j++)
    {
      const int m = tid + j * ICP_THREADS;
      const unsigned sidx = winOf[j];
      if(sidx != 0xffffffffu)
      {
        acc[0] += s_mx[m]; acc[1] += s_my[m];
        acc[2] += s_sx[sidx]; acc[3] += s_sy[sidx];
        const double dx = s_sx[sidx] - s_mx[m];
        const double dy = s_sy[sidx] - s_my[m];
        acc[4] += dx * dx + dy * dy;
        acc[5] += 1.0;
      }
    }
    block_sum_n<6>(acc, s_red, tid);
    pairs = (unsigned)acc[5];

    int retval = TSD_ICP_PROCESSING;
    if(pairs > 2)
    {
      const double sizeInv = 1.0 / (double)pairs;
      const double r = acc[4] * sizeInv;
      const double cm0 = acc[0] * sizeInv, cm1 = acc[1] * sizeInv, cs0 = acc[2] * sizeInv, cs1 = acc[3] * sizeInv;
      rms = r;
      // estimateTransformation (ClosedFormEstimator2D.cpp:74-109)
      double nd[2] = {0, 0};
#pragma unroll
      for(int j = 0; j < (ICP_MAX_POINTS + ICP_THREADS - 1) / ICP_THREADS; j++)
      {
        const int m = tid + j * ICP_THREADS;
        const unsigned sidx = winOf[j];
        if(sidx != 0xffffffffu)
        {
          const double xFCm = s_mx[m] - cm0, yFCm = s_my[m] - cm1;
          const double xSCs = s_sx[sidx] - cs0, ySCs = s_sy[sidx] - cs1;
          nd[0] += yFCm * xSCs - xFCm * ySCs;
          nd[1] += xFCm * xSCs + yFCm * ySCs;
        }
      }
      block_sum_n<2>(nd, s_red, tid);
      if(tid == 0)
      {
        const double deltaTheta = atan2(nd[0], nd[1]);
        const double c = cos(deltaTheta), s = sin(deltaTheta);
        const double deltaX = (cm0 - (c * cs0 - s * cs1));
        const double deltaY = (cm1 - (c * cs1 + s * cs0));
        double* Tl = s_T + 16;
        for(int i = 0; i < 16; i++) Tl[i] = (i % 5 == 0) ? 1.0 : 0.0;
        Tl[0] = c; Tl[1] = -s; Tl[3] = deltaX;
        Tl[4] = s; Tl[5] = c;  Tl[7] = deltaY;
        Tl[11] = 0;
        // Tfinal = Tlast * Tfinal (Icp.cpp:454)
        double out[16];
        for(int i = 0; i < 16; i++) out[i] = 0.0;
        for(int k = 0; k < 4; k++)
          for(int i = 0; i < 4; i++)
          {
            const double temp = 1.0 * Tl[4 * i + k];
            if(temp != 0.0)
              for(int j = 0; j < 4; j++) out[4 * i + j] += temp * s_T[4 * k + j];
          }
        for(int i = 0; i < 16; i++) s_T[i] = out[i];
      }
      __syncthreads();
      // applyTransformation (Icp.cpp:371-408)
      {
        const double* Tl = s_T + 16;
        const double r00 = Tl[0], r01 = Tl[1], r10 = Tl[4], r11 = Tl[5], t0 = Tl[3], t1 = Tl[7];
        for(int i = tid; i < nS; i += ICP_THREADS)
        {
          const double x = s_sx[i], y = s_sy[i];
          double a = 0.0; a += x * r00; a += y * r01; a = 0.0 + 1.0 * a;
          double b = 0.0; b += x * r10; b += y * r11; b = 0.0 + 1.0 * b;
          const double nx = a + t0, ny = b + t1;
          s_sx[i] = nx;
          s_sy[i] = ny;
          if((unsigned)(i % ICP_CLUSTER) == rank)
          {
            // the point moved by |(nx,ny) - (x,y)|: its nearest-neighbour distance shrank by at most that
            const double mvx = nx - x, mvy = ny - y;
            const double lb = s_lb[i / ICP_CLUSTER] - sqrt(mvx * mvx + mvy * mvy) * (1.0 + 1e-9) - 1e-12;
            s_lb[i / ICP_CLUSTER] = lb > 0.0 ? lb : 0.0;
          }
        }
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
      for(int i = 0; i < 16; i++) P.tr_T[16 * iter + i] = s_T[i];
    }
    __syncthreads();
    eRetval = retval;
    // Icp.cpp:496-507
    iter++;
    if(fabs(rms - rms_prev) < 10e-10) conv_cnt++;
    else conv_cnt = 0;
    if((rms <= P.max_rms || conv_cnt >= P.conv_cnt)) eRetval = TSD_ICP_SUCCESS;
    else if(iter >= (unsigned)P.max_iterations) eRetval = TSD_ICP_MAXITERATIONS;
    rms_prev = rms;
  }

  if(rank == 0 && tid == 0)
  {
    // getFinalTransformation (Icp.cpp:528-546)
    P.result[0] = s_T[0]; P.result[1] = s_T[1]; P.result[2] = s_T[3];
    P.result[3] = s_T[4]; P.result[4] = s_T[5]; P.result[5] = s_T[7];
    P.result[6] = 0; P.result[7] = 0; P.result[8] = 1;
    P.result[9] = rms;
    P.result[10] = (double)pairs;
    P.result[11] = (double)iter;
    P.result[12] = (double)eRetval;
  }
  cluster.sync();  // no CTA's shared memory goes away while a neighbour may still address it
}

static size_t icp_smem_bytes(int nM, int nS)
{
  const size_t nQ = ((size_t)nS + ICP_CLUSTER - 1) / ICP_CLUSTER, nH = ((size_t)nM + ICP_CLUSTER - 1) / ICP_CLUSTER;
  size_t b = 0;
  b += sizeof(double) * (2 * (size_t)nM + 2 * (size_t)nS + 2 * nQ);  // mx my sx sy d2 lb
  b += sizeof(unsigned long long) * nM;                            // best
  b += sizeof(double) * (208 + 32);                               // red + T
  b += sizeof(unsigned) * (2 * (size_t)nM) + sizeof(int) * nQ + sizeof(unsigned) * 40;
  b += 16 * ICP_CLUSTER * nH + 16;                                 // inbox
  b += sizeof(unsigned short) * (ICP_SLOTS + 2 + ICP_SLOTS + (size_t)nM + 2) + sizeof(unsigned) * 256;
  return b + 64;
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
  TSD_CUDA(cudaMalloc(&h->d_model, sizeof(double) * 2 * h->cap));
  TSD_CUDA(cudaMalloc(&h->d_scene, sizeof(double) * 2 * h->cap));
  TSD_CUDA(cudaMalloc(&h->d_result, sizeof(double) * 16));
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
  cudaFree(h->d_model); cudaFree(h->d_scene); cudaFree(h->d_result); cudaFree(h->d_tr_model); cudaFree(h->d_tr_scene);
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
  memcpy(h->h_stage, model, sizeof(double) * 2 * n_model);
  memcpy(h->h_stage + 2 * h->cap, scene, sizeof(double) * 2 * n_scene);
  TSD_CUDA(cudaMemcpyAsync(h->d_model, h->h_stage, sizeof(double) * 2 * n_model, cudaMemcpyHostToDevice, h->stream));
  TSD_CUDA(cudaMemcpyAsync(h->d_scene, h->h_stage + 2 * h->cap, sizeof(double) * 2 * n_scene, cudaMemcpyHostToDevice, h->stream));
  IcpParams p = h->p;
  p.nM = n_model;
  p.nS = n_scene;
  for(int i = 0; i < 9; i++) p.pose[i] = pose[i];
  p.has_init = t_init ? 1 : 0;
  for(int i = 0; i < 16; i++) p.t_init[i] = t_init ? t_init[i] : ((i % 5 == 0) ? 1.0 : 0.0);
  p.model = h->d_model;
  p.scene = h->d_scene;
  p.result = h->d_result;
  p.cap = h->cap;
  p.trace = h->trace;
  p.tr_model = h->d_tr_model;
  p.tr_scene = h->d_tr_scene;
  p.tr_count = h->d_tr_count;
  p.tr_mse = h->d_tr_mse;
  p.tr_T = h->d_tr_T;
  if(p.max_iterations > 0) TSD_CUDA(cudaMemsetAsync(h->d_tr_count, 0xff, sizeof(int) * p.max_iterations, h->stream));
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
