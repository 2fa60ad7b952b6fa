// Icp::iterate on the device (K6-K8): the whole registration loop -- pre-filter, exact nearest-neighbour
// pairing, distance filter, reciprocal filter, closed-form estimate, transform update -- is ONE kernel
// launch of one persistent CTA; model, scene and the search structure live in shared memory for all
// iterations, so an ICP run costs one H2D copy, one launch and one D2H copy.
//
// Reference: src/obvision/registration/icp/Icp.cpp:464-512 (iterate), :410-462 (step), :371-408
// (applyTransformation); assign/PairAssignment.cpp:38-84; assign/FlannPairAssignment.cpp:64-92;
// assign/filter/OutOfBoundsFilter2D.cpp:27-37, DistanceFilter.cpp:32-64, ReciprocalFilter.cpp:32-78;
// ClosedFormEstimator2D.cpp:36-109.  Wiring: src/ThreadLocalize.cpp:210-225, :571-581.
//
// Pairing replaces FLANN's kd-tree by a uniform bucket grid over the model points, built once per run
// (the model does not move during ICP).  The search is EXACT: rings of buckets are visited until the best
// squared distance is strictly below the squared distance to everything unvisited, or until everything
// unvisited is beyond the distance filter's current threshold (such a pair is dropped by
// DistanceFilter.cpp:38 whatever its model index).  Distances are computed as FLANN's L2 functor does
// ((0 + dx*dx) + dy*dy); ties go to the lowest model index, the rule the oracle's FLANN stand-in uses.
//
// Sums of the estimator are block reductions with a fixed tree, so results are deterministic but not
// bit-identical to the reference's sequential sums (and atan2/sin/cos differ from glibc in the last ulp
// anyway): pair lists are compared exactly, poses to 1e-9 (tests/test_icp_gpu.py).
#include <string.h>

#include <vector>

#include "common.cuh"

using namespace tsd;

#define ICP_THREADS 1024
#define ICP_MAX_POINTS 2048
#define ICP_G 64  // bucket grid is ICP_G x ICP_G

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
  const double* model;  // nM x 2
  const double* scene;  // nS x 2
  // outputs
  double* result;       // [0..8] T 3x3, [9] mse, [10] pairs, [11] iterations, [12] state
  // trace
  int cap;
  unsigned* tr_model;
  unsigned* tr_scene;
  int* tr_count;
  double* tr_mse;
  double* tr_T;
};

struct tsd_icp
{
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
};

__device__ __forceinline__ double block_sum(double v, double* s_red, int tid)
{
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if((tid & 31) == 0) s_red[tid >> 5] = v;
  __syncthreads();
  double r = (tid < ICP_THREADS / 32) ? s_red[tid] : 0.0;
  if(tid < 32)
  {
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if(tid == 0) s_red[32] = r;
  }
  __syncthreads();
  return s_red[32];
}

__device__ __forceinline__ int bucket_of(double v, double v0, double invh)
{
  int b = __double2int_rd((v - v0) * invh);
  return min(max(b, 0), ICP_G - 1);
}

__global__ void __launch_bounds__(ICP_THREADS, 1) k_icp(IcpParams P)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x;
  const int nM = P.nM, nS = P.nS;
  // shared memory carve-up
  double* s_mx = reinterpret_cast<double*>(smem);
  double* s_my = s_mx + nM;
  double* s_sx = s_my + nM;
  double* s_sy = s_sx + nS;
  double* s_d2 = s_sy + nS;                                          // nS
  unsigned long long* s_best = reinterpret_cast<unsigned long long*>(s_d2 + nS);  // nM
  double* s_red = reinterpret_cast<double*>(s_best + nM);           // 136 (block_sum: 33, bounding box: 128)
  double* s_T = s_red + 136;                                         // Tfinal 16, Tlast 16, misc 8
  unsigned* s_win = reinterpret_cast<unsigned*>(s_T + 40);           // nM
  int* s_nn = reinterpret_cast<int*>(s_win + nM);                    // nS
  unsigned* s_scan = reinterpret_cast<unsigned*>(s_nn + nS);         // 40
  unsigned short* s_bstart = reinterpret_cast<unsigned short*>(s_scan + 40);  // ICP_G*ICP_G + 1
  unsigned short* s_bidx = s_bstart + (ICP_G * ICP_G + 2);           // nM
  unsigned short* s_bcnt = s_bidx + ((nM + 1) & ~1);                 // ICP_G*ICP_G

  for(int i = tid; i < nM; i += ICP_THREADS) { s_mx[i] = P.model[2 * i]; s_my[i] = P.model[2 * i + 1]; }
  for(int i = tid; i < nS; i += ICP_THREADS) { s_sx[i] = P.scene[2 * i]; s_sy[i] = P.scene[2 * i + 1]; }
  for(int i = tid; i < ICP_G * ICP_G; i += ICP_THREADS) s_bcnt[i] = 0;
  if(tid < 16) { s_T[tid] = (tid % 5 == 0) ? 1.0 : 0.0; s_T[16 + tid] = s_T[tid]; }
  __syncthreads();

  // ---- bucket grid over the model (bounding box by block reduction with min/max) ----
  double bx0, by0, invh, h;
  {
    double lx = 1e300, ly = 1e300, hx = -1e300, hy = -1e300;
    for(int i = tid; i < nM; i += ICP_THREADS)
    {
      lx = fmin(lx, s_mx[i]); hx = fmax(hx, s_mx[i]);
      ly = fmin(ly, s_my[i]); hy = fmax(hy, s_my[i]);
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
    {
      lx = fmin(lx, __shfl_xor_sync(0xffffffffu, lx, o)); hx = fmax(hx, __shfl_xor_sync(0xffffffffu, hx, o));
      ly = fmin(ly, __shfl_xor_sync(0xffffffffu, ly, o)); hy = fmax(hy, __shfl_xor_sync(0xffffffffu, hy, o));
    }
    double* s_bb = s_red;  // scratch: 4 x 32
    if((tid & 31) == 0) { s_bb[tid >> 5] = lx; s_bb[32 + (tid >> 5)] = hx; s_bb[64 + (tid >> 5)] = ly; s_bb[96 + (tid >> 5)] = hy; }
    __syncthreads();
    if(tid == 0)
    {
      for(int w = 1; w < ICP_THREADS / 32; w++)
      {
        lx = fmin(lx, s_bb[w]); hx = fmax(hx, s_bb[32 + w]); ly = fmin(ly, s_bb[64 + w]); hy = fmax(hy, s_bb[96 + w]);
      }
      double ext = fmax(hx - lx, hy - ly);
      if(!(ext > 1e-6)) ext = 1e-6;
      const double hh = ext / ICP_G * (1.0 + 1e-9);
      s_T[32] = lx; s_T[33] = ly; s_T[34] = hh; s_T[35] = 1.0 / hh;
    }
    __syncthreads();
    bx0 = s_T[32]; by0 = s_T[33]; h = s_T[34]; invh = s_T[35];
  }
  for(int i = tid; i < nM; i += ICP_THREADS)
  {
    const int b = bucket_of(s_my[i], by0, invh) * ICP_G + bucket_of(s_mx[i], bx0, invh);
    atomicAdd(reinterpret_cast<unsigned*>(s_bcnt) + (b >> 1), (b & 1) ? 0x10000u : 1u);  // u16 counters, nM <= 2048
  }
  __syncthreads();
  if(tid == 0)
  {
    unsigned acc = 0;
    for(int b = 0; b < ICP_G * ICP_G; b++) { s_bstart[b] = (unsigned short)acc; acc += s_bcnt[b]; s_bcnt[b] = 0; }
    s_bstart[ICP_G * ICP_G] = (unsigned short)acc;
  }
  __syncthreads();
  for(int i = tid; i < nM; i += ICP_THREADS)
  {
    const int b = bucket_of(s_my[i], by0, invh) * ICP_G + bucket_of(s_mx[i], bx0, invh);
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

  int eRetval = TSD_ICP_PROCESSING;
  unsigned iter = 0;
  double rms_prev = 10e12;
  unsigned conv_cnt = 0;
  double rms = 0.0;  // the caller passes *rms = 0.0 (ThreadLocalize.cpp:577)
  unsigned pairs = 0;
  double distSqr = P.max_dist_sqr;  // DistanceFilter::reset (DistanceFilter.cpp:27-30)
  const unsigned long long INF64 = 0xffffffffffffffffULL;

  while(eRetval == TSD_ICP_PROCESSING)
  {
    for(int m = tid; m < nM; m += ICP_THREADS) { s_best[m] = INF64; s_win[m] = 0xffffffffu; }
    __syncthreads();

    // ---- A: pre-filter + exact 1-NN + distance filter ----
    for(int i = tid; i < nS; i += ICP_THREADS)
    {
      const double x = s_sx[i], y = s_sy[i];
      // OutOfBoundsFilter2D.cpp:27-37: S.transform(pose) = S * R^T + t
      double tx = 0.0; tx += x * P.pose[0]; tx += y * P.pose[1]; tx = 0.0 + 1.0 * tx; tx += P.pose[2];
      double ty = 0.0; ty += x * P.pose[3]; ty += y * P.pose[4]; ty = 0.0 + 1.0 * ty; ty += P.pose[5];
      const bool masked = (tx < P.x_min || tx > P.x_max || ty < P.y_min || ty > P.y_max);
      int best = -1;
      double bestD = __longlong_as_double(0x7ff0000000000000LL);
      if(!masked)
      {
        const int qx = bucket_of(x, bx0, invh), qy = bucket_of(y, by0, invh);
        for(int r = 0; r < ICP_G; r++)
        {
          const int x0 = qx - r, x1 = qx + r, y0 = qy - r, y1 = qy + r;
          for(int by = max(y0, 0); by <= min(y1, ICP_G - 1); by++)
          {
            const bool edgeRow = (by == y0 || by == y1);
            const int step = edgeRow ? 1 : max(x1 - x0, 1);
            for(int bx = x0; bx <= x1; bx += step)
            {
              if(bx < 0 || bx >= ICP_G) continue;
              const int b = by * ICP_G + bx;
              for(int k = s_bstart[b]; k < s_bstart[b + 1]; k++)
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
          }
          // everything unvisited is farther than r*h (conservatively)
          const double lb = (double)r * h * (1.0 - 1e-9);
          const double lb2 = lb * lb;
          if(bestD < lb2 || distSqr < lb2) break;
          if(x0 <= 0 && y0 <= 0 && x1 >= ICP_G - 1 && y1 >= ICP_G - 1) break;
        }
      }
      const bool keep = (best >= 0) && (bestD <= distSqr);  // DistanceFilter.cpp:38
      s_nn[i] = keep ? best : -1;
      s_d2[i] = bestD;
      if(keep) atomicMin(&s_best[best], (unsigned long long)__double_as_longlong(bestD));
    }
    __syncthreads();
    // ---- B: ReciprocalFilter.cpp:32-78: closest scene point per model point (lowest scene index on ties) ----
    for(int i = tid; i < nS; i += ICP_THREADS)
    {
      const int m = s_nn[i];
      if(m >= 0 && (unsigned long long)__double_as_longlong(s_d2[i]) == s_best[m]) atomicMin(&s_win[m], (unsigned)i);
    }
    __syncthreads();
    // DistanceFilter.cpp:62-63
    distSqr *= P.multiplier;
    if(distSqr < P.min_dist_sqr) distSqr = P.min_dist_sqr;

    // ---- C: pair list in model order (trace) + ClosedFormEstimator2D::setPairs ----
    double cm0 = 0, cm1 = 0, cs0 = 0, cs1 = 0, r = 0;
    unsigned baseCount = 0;
    for(int m0 = 0; m0 < nM; m0 += ICP_THREADS)
    {
      const int m = m0 + tid;
      const bool has = (m < nM) && (s_win[m] != 0xffffffffu);
      const unsigned bal = __ballot_sync(0xffffffffu, has);
      if((tid & 31) == 0) s_scan[tid >> 5] = __popc(bal);
      __syncthreads();
      unsigned off = baseCount;
      for(int w = 0; w < (tid >> 5); w++) off += s_scan[w];
      unsigned total = 0;
      for(int w = 0; w < ICP_THREADS / 32; w++) total += s_scan[w];
      if(has)
      {
        const unsigned pos = off + __popc(bal & ((1u << (tid & 31)) - 1u));
        const unsigned sidx = s_win[m];
        if((int)iter < P.max_iterations && pos < (unsigned)P.cap)
        {
          P.tr_model[(size_t)iter * P.cap + pos] = (unsigned)m;
          P.tr_scene[(size_t)iter * P.cap + pos] = sidx;
        }
        cm0 += s_mx[m]; cm1 += s_my[m];
        cs0 += s_sx[sidx]; cs1 += s_sy[sidx];
        const double dx = s_sx[sidx] - s_mx[m];
        const double dy = s_sy[sidx] - s_my[m];
        r += dx * dx + dy * dy;
      }
      baseCount += total;
      __syncthreads();
    }
    pairs = baseCount;

    int retval = TSD_ICP_PROCESSING;
    if(pairs > 2)
    {
      cm0 = block_sum(cm0, s_red, tid);
      cm1 = block_sum(cm1, s_red, tid);
      cs0 = block_sum(cs0, s_red, tid);
      cs1 = block_sum(cs1, s_red, tid);
      r = block_sum(r, s_red, tid);
      const double sizeInv = 1.0 / (double)pairs;
      r *= sizeInv; cm0 *= sizeInv; cm1 *= sizeInv; cs0 *= sizeInv; cs1 *= sizeInv;
      rms = r;
      // estimateTransformation (ClosedFormEstimator2D.cpp:74-109)
      double nom = 0, den = 0;
      for(int m = tid; m < nM; m += ICP_THREADS)
      {
        const unsigned sidx = s_win[m];
        if(sidx != 0xffffffffu)
        {
          const double xFCm = s_mx[m] - cm0, yFCm = s_my[m] - cm1;
          const double xSCs = s_sx[sidx] - cs0, ySCs = s_sy[sidx] - cs1;
          nom += yFCm * xSCs - xFCm * ySCs;
          den += xFCm * xSCs + yFCm * ySCs;
        }
      }
      nom = block_sum(nom, s_red, tid);
      den = block_sum(den, s_red, tid);
      if(tid == 0)
      {
        const double deltaTheta = atan2(nom, den);
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
          s_sx[i] = a + t0;
          s_sy[i] = b + t1;
        }
      }
    }
    else
    {
      retval = TSD_ICP_NOTMATCHABLE;
    }
    if(tid == 0 && (int)iter < P.max_iterations)
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

  if(tid == 0)
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
}

static size_t icp_smem_bytes(int nM, int nS)
{
  size_t b = 0;
  b += sizeof(double) * (2 * (size_t)nM + 2 * (size_t)nS + nS);  // mx my sx sy d2
  b += sizeof(unsigned long long) * nM;                           // best
  b += sizeof(double) * 176;                                      // red + T
  b += sizeof(unsigned) * nM + sizeof(int) * nS + sizeof(unsigned) * 40;
  b += sizeof(unsigned short) * (ICP_G * ICP_G + 2 + ((nM + 1) & ~1) + ICP_G * ICP_G);
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
  h->cap = ICP_MAX_POINTS;
  TSD_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  TSD_CUDA(cudaMalloc(&h->d_model, sizeof(double) * 2 * h->cap));
  TSD_CUDA(cudaMalloc(&h->d_scene, sizeof(double) * 2 * h->cap));
  TSD_CUDA(cudaMalloc(&h->d_result, sizeof(double) * 16));
  const size_t mi = max_iterations > 0 ? max_iterations : 1;
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
  delete h;
  return TSD_OK;
}

int icp_run(tsd_icp_t* h, const double* model, const double* normals, int32_t n_model, const double* scene,
            int32_t n_scene, const double pose[9], const double* t_init, double t_out[9], double* mse,
            uint32_t* pairs, uint32_t* iterations, int32_t* state)
{
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
  p.tr_model = h->d_tr_model;
  p.tr_scene = h->d_tr_scene;
  p.tr_count = h->d_tr_count;
  p.tr_mse = h->d_tr_mse;
  p.tr_T = h->d_tr_T;
  if(p.max_iterations > 0) TSD_CUDA(cudaMemsetAsync(h->d_tr_count, 0xff, sizeof(int) * p.max_iterations, h->stream));
  k_icp<<<1, ICP_THREADS, icp_smem_bytes(n_model, n_scene), h->stream>>>(p);
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

int icp_get_trace(tsd_icp_t* h, int32_t max_it, int32_t cap, uint32_t* pair_model, uint32_t* pair_scene,
                  int32_t* pair_count, double* mse, double* t_final16, int32_t* n_it)
{
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
