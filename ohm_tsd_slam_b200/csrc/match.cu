// Hypothesis scoring of the RANSAC pre-registration matchers (K9 TSD_PDF, K10 RandomNormal, K11 PDF):
// one warp per hypothesis, control set / model staged in shared memory, lanes stride over the control
// points.  The winner is picked on the device in the reference's single-thread order (first best).
//
// Reference: src/obvision/registration/ransacMatching/TSD_PDFMatching.cpp:206-260,
// RandomNormalMatching.cpp:251-359, PDFMatching.cpp:235-388 and :435-487.
//
// Per-hypothesis products / sums are combined lane-partial first, then across lanes, i.e. in a fixed but
// different association than the reference's left-to-right loop; cos/sin/atan2/exp come from the CUDA
// math library.  Scores therefore match the oracle to ~1e-12 relative, not bit for bit (tests use 1e-9).
#include <string.h>

#include <vector>

#include "common.cuh"

using namespace tsd;

struct tsd_matcher
{
  std::recursive_mutex* mtx;
  int device;
  cudaStream_t stream;
  size_t cap;
  void* d_buf;
  void* h_buf;  // pinned
};

struct HypCommon
{
  int n_hyp;
  const tsd_hypothesis_t* hyps;
  const double* model;  // n x 2
  const double* scene;  // n x 2
  const double* phi_m;
  const double* phi_s;
  double phi_max;
  int n_control;
  const double* control;  // 3 x n_control
};

// TSD_PDFMatching.cpp:206-220 (identical in the other two matchers).  false: skipped.
__device__ __forceinline__ bool hypothesis_transform(const HypCommon& hc, int h, double T[9], double* phi_out)
{
  const int idx = hc.hyps[h].idx_model, i = hc.hyps[h].idx_scene;
  const double pi = 3.14159265358979323846;
  double phi = hc.phi_m[idx] - hc.phi_s[i];
  if(phi > pi) phi -= 2.0 * pi;
  else if(phi < -pi) phi += 2.0 * pi;
  *phi_out = phi;
  if(!(fabs(phi) < hc.phi_max)) return false;
  const double c = cos(phi), s = sin(phi);
  T[0] = c; T[1] = -s; T[2] = 0.0;
  T[3] = s; T[4] = c;  T[5] = 0.0;
  T[6] = 0.0; T[7] = 0.0; T[8] = 1.0;
  const double sx = hc.scene[2 * i], sy = hc.scene[2 * i + 1];
  T[2] = hc.model[2 * idx] - (T[0] * sx + T[1] * sy);
  T[5] = hc.model[2 * idx + 1] - (T[3] * sx + T[4] * sy);
  return true;
}

// one column of `A * Control` (dgemm NoTrans x NoTrans, zero coefficients skipped), rows 0 and 1
__device__ __forceinline__ void transform_control(const double* A, double c0, double c1, double c2, double* x, double* y)
{
  mat3_vec_nn(A, c0, c1, c2, x, y);
}

#define MATCH_WARPS 8

// ---------------------------------------------------------------- K9: TSD_PDFMatching.cpp:222-251
__global__ void __launch_bounds__(MATCH_WARPS * 32) k_score_tsd(HypCommon hc, GridView g, const double* t_sensor,
                                                                double zrand, double* score)
{
  extern __shared__ double s_ctrl[];  // 3 x n_control
  for(int i = threadIdx.x; i < 3 * hc.n_control; i += blockDim.x) s_ctrl[i] = hc.control[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warpsTotal = gridDim.x * MATCH_WARPS;
  for(int h = blockIdx.x * MATCH_WARPS + (threadIdx.x >> 5); h < hc.n_hyp; h += warpsTotal)
  {
    double T[9], phi;
    if(!hypothesis_transform(hc, h, T, &phi))
    {
      if(lane == 0) score[h] = -1.0;
      continue;
    }
    // TMap = TSensor * T (dgemm NoTrans x NoTrans)
    double TMap[9];
#pragma unroll
    for(int i = 0; i < 9; i++) TMap[i] = 0.0;
#pragma unroll
    for(int k = 0; k < 3; k++)
#pragma unroll
      for(int i = 0; i < 3; i++)
      {
        const double temp = 1.0 * t_sensor[3 * i + k];
        if(temp != 0.0)
        {
#pragma unroll
          for(int j = 0; j < 3; j++) TMap[3 * i + j] += temp * T[3 * k + j];
        }
      }
    double prob = 1.0;
    for(int s = lane; s < hc.n_control; s += 32)
    {
      double x, y;
      transform_control(TMap, s_ctrl[s], s_ctrl[hc.n_control + s], s_ctrl[2 * hc.n_control + s], &x, &y);
      double tsd;
      if(sample_bilinear(g, x, y, &tsd) == TSD_INTERPOLATE_SUCCESS) prob *= (1.0 - (1.0 - zrand) * fabs(tsd));
      else prob *= zrand;
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) prob *= __shfl_xor_sync(0xffffffffu, prob, o);
    if(lane == 0) score[h] = prob;
  }
}

// ---------------------------------------------------------------- K10: RandomNormalMatching.cpp:265-342
struct RnmParams
{
  const double* phi_control;
  int n_valid;
  const double* model_valid;  // n_valid x 2
  const double* phi_valid;
  double theta_min, theta_max, scale_distance, scale_orientation;
  int* cnt_match;
  int* max_cnt_match;
  double* err_sum;
};

__global__ void __launch_bounds__(MATCH_WARPS * 32) k_score_rnm(HypCommon hc, RnmParams rp)
{
  extern __shared__ double s_buf[];
  double* s_ctrl = s_buf;                      // 3 x C
  double* s_phic = s_ctrl + 3 * hc.n_control;  // C
  double* s_mx = s_phic + hc.n_control;        // n_valid
  double* s_my = s_mx + rp.n_valid;
  double* s_mphi = s_my + rp.n_valid;
  for(int i = threadIdx.x; i < 3 * hc.n_control; i += blockDim.x) s_ctrl[i] = hc.control[i];
  for(int i = threadIdx.x; i < hc.n_control; i += blockDim.x) s_phic[i] = rp.phi_control[i];
  for(int i = threadIdx.x; i < rp.n_valid; i += blockDim.x)
  {
    s_mx[i] = rp.model_valid[2 * i];
    s_my[i] = rp.model_valid[2 * i + 1];
    s_mphi[i] = rp.phi_valid[i];
  }
  // Exact nearest neighbour without looking at every model point: the valid model points come in beam order, i.e.
  // along the scan contour, so 32 consecutive points are a compact group.  Every group gets a bounding box; a query
  // scans a group only if the box could hold a point at least as near as the best one found so far.  Lanes of a warp
  // hold consecutive control points (also along a contour), so they mostly want the same few groups and the warp
  // stays coherent.  Distances are formed exactly as in the brute-force scan, ties go to the lowest index: the
  // result is the brute-force result (the FLANN stand-in's rule), at ~1/7 of the distance evaluations.
  //
  // The box test runs in SINGLE precision on conservatively rounded numbers (boxes rounded outward, the distance bounded
  // from below by more than the conversion error, the running best rounded up): a double-precision instruction takes two
  // issue slots on this part, and the 32 box tests per control point were as many of them as the scans they save.
  // (Screening the POINTS of a group the same way was measured and lost: the lanes of a warp hold different control
  // points, some lane nearly always needs the exact distance, and the warp then pays for both.)
  const int nGroups = (rp.n_valid + 31) >> 5;
  float4* s_boxf = reinterpret_cast<float4*>((reinterpret_cast<uintptr_t>(s_mphi + rp.n_valid) + 15) & ~(uintptr_t)15);  // per group: x0 x1 y0 y1, rounded outward
  float* s_mabs = reinterpret_cast<float*>(s_boxf + nGroups);       // [0]: largest |coordinate| of the model
  __syncthreads();
  for(int g = threadIdx.x; g < nGroups; g += blockDim.x)
  {
    double x0 = s_mx[32 * g], x1 = x0, y0 = s_my[32 * g], y1 = y0;
    for(int k = 32 * g + 1; k < min(32 * g + 32, rp.n_valid); k++)
    {
      x0 = fmin(x0, s_mx[k]); x1 = fmax(x1, s_mx[k]);
      y0 = fmin(y0, s_my[k]); y1 = fmax(y1, s_my[k]);
    }
    s_boxf[g] = make_float4(__double2float_rd(x0), __double2float_ru(x1), __double2float_rd(y0), __double2float_ru(y1));
  }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    float a = 0.f;
    for(int g = 0; g < nGroups; g++)
      a = fmaxf(a, fmaxf(fmaxf(fabsf(s_boxf[g].x), fabsf(s_boxf[g].y)), fmaxf(fabsf(s_boxf[g].z), fabsf(s_boxf[g].w))));
    s_mabs[0] = a;
  }
  __syncthreads();
  const float mabs = s_mabs[0];
  const int lane = threadIdx.x & 31;
  const int warpsTotal = gridDim.x * MATCH_WARPS;
  for(int h = blockIdx.x * MATCH_WARPS + (threadIdx.x >> 5); h < hc.n_hyp; h += warpsTotal)
  {
    double T[9], phi;
    if(!hypothesis_transform(hc, h, T, &phi))
    {
      if(lane == 0) { rp.cnt_match[h] = -1; rp.max_cnt_match[h] = 0; rp.err_sum[h] = 0.0; }
      continue;
    }
    int maxCnt = 0, cnt = 0, prevBest = -1;
    double errSum = 0.0;
    for(int s = lane; s < hc.n_control; s += 32)
    {
      double x, y;
      transform_control(T, s_ctrl[s], s_ctrl[hc.n_control + s], s_ctrl[2 * hc.n_control + s], &x, &y);
      const double theta = atan2(y, x);
      if(theta > rp.theta_max || theta < rp.theta_min) continue;  // :274-277
      maxCnt++;
      // exact 1-NN among the valid model points ((0 + dx*dx) + dy*dy, lowest index on ties)
      int bi = -1;
      double bd = __longlong_as_double(0x7ff0000000000000LL);
      float bdf = __int_as_float(0x7f800000);  // bd rounded up
      // single-precision images of the query and a bound e on |image difference - true difference| per axis:
      // two conversions (half an ulp each of a magnitude below |x| + |m|) and one subtraction (half an ulp of the result)
      const float xf = (float)x, yf = (float)y;
      const float e = 1.3e-7f * (fmaxf(fabsf(xf), fabsf(yf)) + mabs) + 1e-30f;
      auto scan_group = [&](int g)
      {
        const int k1 = min(32 * g + 32, rp.n_valid);
        for(int k = 32 * g; k < k1; k++)
        {
          const double d0 = x - s_mx[k];
          const double d1 = y - s_my[k];
          double d = 0.0;
          d += d0 * d0;
          d += d1 * d1;
          if(d < bd || (d == bd && k < bi)) { bd = d; bi = k; bdf = __double2float_ru(d); }
        }
      };
      // seed with the group the previous control point of this lane ended in (usually the right one already)
      const int seed = (prevBest >= 0) ? (prevBest >> 5) : -1;
      if(seed >= 0) scan_group(seed);
      for(int g = 0; g < nGroups; g++)
      {
        if(g == seed) continue;
        const float4 b = s_boxf[g];
        const float ex = fmaxf(fmaxf(b.x - xf, xf - b.y) * (1.f - 2e-7f) - e, 0.f);
        const float ey = fmaxf(fmaxf(b.z - yf, yf - b.w) * (1.f - 2e-7f) - e, 0.f);
        // lower bound of the squared distance to anything in the box
        if((ex * ex + ey * ey) * (1.f - 1e-6f) <= bdf) scan_group(g);
      }
      prevBest = bi;
      if(bi < 0) continue;
      const double normalConsensus = (1.0 - cos(s_mphi[bi] - s_phic[s] - phi)) / 2.0;
      const double err = bd * rp.scale_distance + normalConsensus * rp.scale_orientation;
      errSum += err;
      if(err < 1.0) cnt++;
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
    {
      maxCnt += __shfl_xor_sync(0xffffffffu, maxCnt, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      errSum += __shfl_xor_sync(0xffffffffu, errSum, o);
    }
    if(lane == 0) { rp.cnt_match[h] = cnt; rp.max_cnt_match[h] = maxCnt; rp.err_sum[h] = errSum; }
  }
}

// ---------------------------------------------------------------- K11: PDFMatching.cpp:304-370, :435-487
struct PdfParams
{
  int n_valid;
  int sorted;  // model_angles strictly increasing (the usual case: model points come in beam order): binary search
  const double* model_angles;
  const double* model_dists;
  double p[12];
  double* prob;
  int* fov_count;
};

__device__ __forceinline__ double probability_of_two_single_scans(const double* p, double m, double s)
{
  const double zhit = p[0], zphi = p[1], zshort = p[2], zmax = p[3], zrand = p[4];
  const double rangemax = p[6], sigphi = p[7], sighit = p[8], lamshort = p[9];
  const double sigphit = 1.0 / (sqrt(2.0 * 3.14159265358979323846) * sighit);
  double phit = 0, pphi = 0, pshort = 0, pmax = 0, prand = 0;
  // pow(M_E, x) of the reference == exp(x) to ~1e-16 relative
  if(s < rangemax) phit = sigphit * exp((-0.5 * ((m - s) * (m - s))) / (sighit * sighit));
  if(zphi != 0.0) pphi = sigphi * exp((-0.5 * s * s) / (sigphi * sigphi));
  if(s < m)
  {
    const double n = 1.0 / (1.0 - exp(-lamshort * m));
    pshort = n * lamshort * exp(-lamshort * s);
  }
  if(s >= rangemax) pmax = 1.0;
  if(s < rangemax) prand = 1.0 / rangemax;
  return zhit * phit + zshort * pshort + zmax * pmax + zrand * prand + zphi * pphi;
}

__global__ void __launch_bounds__(MATCH_WARPS * 32) k_score_pdf(HypCommon hc, PdfParams pp)
{
  extern __shared__ double s_buf[];
  double* s_ctrl = s_buf;                     // 3 x C
  double* s_ang = s_ctrl + 3 * hc.n_control;  // n_valid
  double* s_dst = s_ang + pp.n_valid;
  for(int i = threadIdx.x; i < 3 * hc.n_control; i += blockDim.x) s_ctrl[i] = hc.control[i];
  for(int i = threadIdx.x; i < pp.n_valid; i += blockDim.x) { s_ang[i] = pp.model_angles[i]; s_dst[i] = pp.model_dists[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warpsTotal = gridDim.x * MATCH_WARPS;
  const double pi = 3.14159265358979323846;
  const double angleThresh = (pi / 180.0) * pp.p[10];
  for(int h = blockIdx.x * MATCH_WARPS + (threadIdx.x >> 5); h < hc.n_hyp; h += warpsTotal)
  {
    double T[9], phi;
    if(!hypothesis_transform(hc, h, T, &phi))
    {
      if(lane == 0) { pp.prob[h] = -1.0; pp.fov_count[h] = 0; }
      continue;
    }
    double prob = 1.0;
    int fov = 0;
    for(int s = lane; s < hc.n_control; s += 32)
    {
      double x, y;
      transform_control(T, s_ctrl[s], s_ctrl[hc.n_control + s], s_ctrl[2 * hc.n_control + s], &x, &y);
      const double angle = atan2(y, x);
      const double distance = sqrt(x * x + y * y);
      // arg-min of |angle - modelAngle[k]|, first index on ties (PDFMatching.cpp:321-333 is a linear scan)
      double minAngleDiff = 2 * pi;
      int idxMin = 0;
      if(pp.sorted)
      {
        // Sorted angles: the rounded differences are weakly V-shaped in k, so the scan's answer is the leftmost
        // local minimum next to the insertion point of `angle`.
        const int nv = pp.n_valid;
        int lo = 0, hi = nv;
        while(lo < hi)
        {
          const int mid = (lo + hi) >> 1;
          if(s_ang[mid] < angle) lo = mid + 1;
          else hi = mid;
        }
        int c = (lo < nv) ? lo : nv - 1;
        double dc = fabs(angle - s_ang[c]);
        while(c > 0)
        {
          const double d = fabs(angle - s_ang[c - 1]);
          if(d <= dc) { c--; dc = d; }
          else break;
        }
        while(c + 1 < nv)
        {
          const double d = fabs(angle - s_ang[c + 1]);
          if(d < dc) { c++; dc = d; }
          else break;
        }
        if(dc < minAngleDiff) { minAngleDiff = dc; idxMin = c; }
      }
      else
      {
        for(int k = 0; k < pp.n_valid; k++)
        {
          const double diff = fabs(angle - s_ang[k]);
          if(diff < minAngleDiff) { minAngleDiff = diff; idxMin = k; }
        }
      }
      if(minAngleDiff < angleThresh) fov++;
      prob *= probability_of_two_single_scans(pp.p, s_dst[idxMin], distance);
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
    {
      prob *= __shfl_xor_sync(0xffffffffu, prob, o);
      fov += __shfl_xor_sync(0xffffffffu, fov, o);
    }
    if(hc.n_control == 0) prob = 0.0;  // PDFMatching.cpp:359-363
    if(lane == 0) { pp.prob[h] = prob; pp.fov_count[h] = fov; }
  }
}

// first maximum of score[h] over h with score[h] > 0 (and, for PDF, fov[h] > fov_min): the reference's
// `if(prob > bestProb)` with bestProb = 0 evaluated in list order.
__global__ void __launch_bounds__(1024) k_first_max(int n, const double* score, const int* fov, double fov_min, int* best)
{
  __shared__ double s_v[32];
  __shared__ int s_i[32];
  double bv = 0.0;
  int bi = -1;
  for(int h = threadIdx.x; h < n; h += blockDim.x)
  {
    const double v = score[h];
    const bool ok = (v > 0.0) && (fov == nullptr || (double)fov[h] > fov_min);
    if(ok && (v > bv || bi < 0)) { bv = v; bi = h; }  // h ascending per thread: first max kept
  }
#pragma unroll
  for(int o = 16; o > 0; o >>= 1)
  {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if(oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
  }
  if((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    for(int w = 1; w < 32; w++)
    {
      const double ov = s_v[w];
      const int oi = s_i[w];
      if(oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    *best = bi;
  }
}

// ---------------------------------------------------------------- host side
static int ensure(tsd_matcher* m, size_t bytes)
{
  if(bytes <= m->cap) return TSD_OK;
  TSD_CUDA(cudaStreamSynchronize(m->stream));
  if(m->d_buf) cudaFree(m->d_buf);
  if(m->h_buf) cudaFreeHost(m->h_buf);
  m->d_buf = m->h_buf = nullptr;
  m->cap = 0;
  size_t cap = 1 << 20;
  while(cap < bytes) cap <<= 1;
  TSD_CUDA(cudaMalloc(&m->d_buf, cap));
  TSD_CUDA(cudaMallocHost(&m->h_buf, cap));
  m->cap = cap;
  return TSD_OK;
}

// bump allocator over the pinned/device buffer pair; everything 16-byte aligned
struct Arena
{
  unsigned char* h;
  unsigned char* d;
  size_t off;
  template <typename T>
  T* put(const T* src, size_t count, T** hostCopy = nullptr)
  {
    const size_t bytes = ((sizeof(T) * count + 15) / 16) * 16;
    if(src) memcpy(h + off, src, sizeof(T) * count);
    if(hostCopy) *hostCopy = reinterpret_cast<T*>(h + off);
    T* dp = reinterpret_cast<T*>(d + off);
    off += bytes;
    return dp;
  }
};

static size_t a16(size_t b) { return ((b + 15) / 16) * 16; }

static void set_identity3(double T[9])
{
  for(int i = 0; i < 9; i++) T[i] = (i % 4 == 0) ? 1.0 : 0.0;
}

// T of the winning hypothesis, recomputed on the host exactly as the kernels' hypothesis_transform does
// but with the host libm (this is the matrix handed to ICP as the initial guess).
static void best_transform(int best, const tsd_hypothesis_t* hyps, const double* model, const double* scene,
                           const double* phi_m, const double* phi_s, double T[9])
{
  set_identity3(T);
  if(best < 0) return;
  const int idx = hyps[best].idx_model, i = hyps[best].idx_scene;
  double phi = phi_m[idx] - phi_s[i];
  if(phi > M_PI) phi -= 2.0 * M_PI;
  else if(phi < -M_PI) phi += 2.0 * M_PI;
  const double c = cos(phi), s = sin(phi);
  T[0] = c; T[1] = -s; T[3] = s; T[4] = c;
  const double sx = scene[2 * i], sy = scene[2 * i + 1];
  T[2] = model[2 * idx] - (T[0] * sx + T[1] * sy);
  T[5] = model[2 * idx + 1] - (T[3] * sx + T[4] * sy);
}

static int grid_for(int n_hyp, int sm)
{
  int ctas = (n_hyp + MATCH_WARPS - 1) / MATCH_WARPS;
  const int cap = sm * 8;
  return ctas < cap ? (ctas > 0 ? ctas : 1) : cap;
}

extern "C" {

int match_create(int device, tsd_matcher_t** out)
{
  if(!out) return TSD_E_INVALID;
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
  tsd_matcher* m = new tsd_matcher();
  memset(m, 0, sizeof(*m));
  m->mtx = new std::recursive_mutex();
  m->device = device;
  TSD_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  *out = m;
  return TSD_OK;
}

int match_destroy(tsd_matcher_t* m)
{
  if(!m) return TSD_OK;
  cudaSetDevice(m->device);
  if(m->stream) cudaStreamSynchronize(m->stream);
  cudaFree(m->d_buf);
  cudaFreeHost(m->h_buf);
  if(m->stream) cudaStreamDestroy(m->stream);
  cudaGetLastError();
  delete m->mtx;
  delete m;
  return TSD_OK;
}

int match_score_tsd(tsd_matcher_t* m, tsd_grid_t* grid, int32_t n_hyp, const tsd_hypothesis_t* hyps, int32_t n,
                    const double* model, const double* scene, const double* phi_m, const double* phi_s,
                    double phi_max, int32_t n_control, const double* control, const double t_sensor[9],
                    double zrand, double* score, int32_t* best, double t_best[9])
{
  TSD_LOCK(m);
  std::unique_lock<std::recursive_mutex> grid_lock__;
  if(grid) grid_lock__ = std::unique_lock<std::recursive_mutex>(*grid->mtx);
  if(!m || !grid || n_hyp < 0 || n <= 0 || !hyps || !model || !scene || !phi_m || !phi_s || n_control < 0 ||
     (n_control > 0 && !control) || !t_sensor || !best || !t_best)
    return TSD_E_INVALID;
  set_identity3(t_best);
  *best = -1;
  if(n_hyp == 0) return TSD_OK;
  if(grid->device != m->device) { set_error("matcher and grid live on different devices"); return TSD_E_INVALID; }
  TSD_CUDA(cudaSetDevice(m->device));
  const size_t inBytes = a16(sizeof(tsd_hypothesis_t) * n_hyp) + 2 * a16(sizeof(double) * 2 * n) + 2 * a16(sizeof(double) * n) +
                         a16(sizeof(double) * 3 * (n_control + 1)) + a16(sizeof(double) * 9);
  const size_t outBytes = a16(sizeof(double) * n_hyp) + 16;
  int rc = ensure(m, inBytes + outBytes);
  if(rc) return rc;
  TSD_CUDA(cudaStreamSynchronize(m->stream));
  Arena a{(unsigned char*)m->h_buf, (unsigned char*)m->d_buf, 0};
  HypCommon hc;
  hc.n_hyp = n_hyp;
  hc.hyps = a.put(hyps, n_hyp);
  hc.model = a.put(model, 2 * (size_t)n);
  hc.scene = a.put(scene, 2 * (size_t)n);
  hc.phi_m = a.put(phi_m, n);
  hc.phi_s = a.put(phi_s, n);
  hc.phi_max = phi_max;
  hc.n_control = n_control;
  hc.control = a.put(control, 3 * (size_t)n_control + (n_control == 0 ? 1 : 0));
  const double* d_ts = a.put(t_sensor, 9);
  const size_t inEnd = a.off;
  double* h_score;
  double* d_score = a.put<double>(nullptr, n_hyp, &h_score);
  int* h_best;
  int* d_best = a.put<int>(nullptr, 4, &h_best);
  TSD_CUDA(cudaMemcpyAsync(m->d_buf, m->h_buf, inEnd, cudaMemcpyHostToDevice, m->stream));
  // the grid's stream may still be pushing: order after it
  TSD_CUDA(cudaStreamSynchronize(grid->stream));
  const size_t smem = sizeof(double) * 3 * (n_control + 1);
  if(smem > 48 * 1024) TSD_CUDA(cudaFuncSetAttribute(k_score_tsd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_score_tsd<<<grid_for(n_hyp, grid->sm_count), MATCH_WARPS * 32, smem, m->stream>>>(hc, grid_view(grid), d_ts, zrand, d_score);
  TSD_LAUNCHED();
  k_first_max<<<1, 1024, 0, m->stream>>>(n_hyp, d_score, nullptr, 0.0, d_best);
  TSD_LAUNCHED();
  if(score) TSD_CUDA(cudaMemcpyAsync(h_score, d_score, sizeof(double) * n_hyp, cudaMemcpyDeviceToHost, m->stream));
  TSD_CUDA(cudaMemcpyAsync(h_best, d_best, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  TSD_CUDA(cudaStreamSynchronize(m->stream));
  if(score) memcpy(score, h_score, sizeof(double) * n_hyp);
  *best = *h_best;
  best_transform(*best, hyps, model, scene, phi_m, phi_s, t_best);
  return TSD_OK;
}

int match_score_rnm(tsd_matcher_t* m, int32_t n_hyp, const tsd_hypothesis_t* hyps, int32_t n, const double* model,
                    const double* scene, const double* phi_m, const double* phi_s, double phi_max,
                    int32_t n_control, const double* control, const double* phi_control, int32_t n_valid,
                    const double* model_valid, const double* phi_valid, double theta_min, double theta_max,
                    double scale_distance, double scale_orientation, uint32_t cnt_match_thresh, int32_t* cnt_match,
                    int32_t* max_cnt_match, double* err_sum, int32_t* best, double t_best[9])
{
  TSD_LOCK(m);
  if(!m || n_hyp < 0 || n <= 0 || !hyps || !model || !scene || !phi_m || !phi_s || n_control < 0 ||
     (n_control > 0 && (!control || !phi_control)) || n_valid < 0 || (n_valid > 0 && (!model_valid || !phi_valid)) ||
     !best || !t_best)
    return TSD_E_INVALID;
  set_identity3(t_best);
  *best = -1;
  if(n_hyp == 0) return TSD_OK;
  TSD_CUDA(cudaSetDevice(m->device));
  const size_t inBytes = a16(sizeof(tsd_hypothesis_t) * n_hyp) + 2 * a16(sizeof(double) * 2 * n) + 2 * a16(sizeof(double) * n) +
                         a16(sizeof(double) * 3 * (n_control + 1)) + a16(sizeof(double) * (n_control + 1)) +
                         a16(sizeof(double) * 2 * (n_valid + 1)) + a16(sizeof(double) * (n_valid + 1));
  const size_t outBytes = 2 * a16(sizeof(int) * n_hyp) + a16(sizeof(double) * n_hyp);
  int rc = ensure(m, inBytes + outBytes);
  if(rc) return rc;
  TSD_CUDA(cudaStreamSynchronize(m->stream));
  Arena a{(unsigned char*)m->h_buf, (unsigned char*)m->d_buf, 0};
  HypCommon hc;
  hc.n_hyp = n_hyp;
  hc.hyps = a.put(hyps, n_hyp);
  hc.model = a.put(model, 2 * (size_t)n);
  hc.scene = a.put(scene, 2 * (size_t)n);
  hc.phi_m = a.put(phi_m, n);
  hc.phi_s = a.put(phi_s, n);
  hc.phi_max = phi_max;
  hc.n_control = n_control;
  hc.control = a.put(control, 3 * (size_t)n_control + (n_control == 0 ? 1 : 0));
  RnmParams rp;
  rp.phi_control = a.put(phi_control, (size_t)n_control + (n_control == 0 ? 1 : 0));
  rp.n_valid = n_valid;
  rp.model_valid = a.put(model_valid, 2 * (size_t)n_valid + (n_valid == 0 ? 1 : 0));
  rp.phi_valid = a.put(phi_valid, (size_t)n_valid + (n_valid == 0 ? 1 : 0));
  rp.theta_min = theta_min;
  rp.theta_max = theta_max;
  rp.scale_distance = scale_distance;
  rp.scale_orientation = scale_orientation;
  const size_t inEnd = a.off;
  int *h_cnt, *h_max;
  double* h_err;
  rp.cnt_match = a.put<int>(nullptr, n_hyp, &h_cnt);
  rp.max_cnt_match = a.put<int>(nullptr, n_hyp, &h_max);
  rp.err_sum = a.put<double>(nullptr, n_hyp, &h_err);
  TSD_CUDA(cudaMemcpyAsync(m->d_buf, m->h_buf, inEnd, cudaMemcpyHostToDevice, m->stream));
  const size_t smem = sizeof(double) * (4 * (size_t)n_control + 3 * (size_t)n_valid) + 16 * (((size_t)n_valid + 31) / 32) + 8 * (size_t)n_valid + 32;
  if(smem > 200 * 1024) { set_error("control set / model too large for shared memory"); return TSD_E_INVALID; }
  if(smem > 48 * 1024) TSD_CUDA(cudaFuncSetAttribute(k_score_rnm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sm = 148;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, m->device);
  k_score_rnm<<<grid_for(n_hyp, sm), MATCH_WARPS * 32, smem, m->stream>>>(hc, rp);
  TSD_LAUNCHED();
  TSD_CUDA(cudaMemcpyAsync(h_cnt, rp.cnt_match, a.off - inEnd, cudaMemcpyDeviceToHost, m->stream));
  TSD_CUDA(cudaStreamSynchronize(m->stream));
  // RandomNormalMatching.cpp:338-359: the reference's ordered, non-associative best-select, replayed in
  // list order over the per-hypothesis triples
  double bestRatio = 0.0;
  unsigned bestCnt = 0;
  double bestErr = 1e12;
  for(int h = 0; h < n_hyp; h++)
  {
    if(h_cnt[h] < 0) continue;
    const unsigned cntMatch = (unsigned)h_cnt[h];
    if(cntMatch <= cnt_match_thresh) continue;
    const double ratio = (double)cntMatch / (double)h_max[h];
    const double equalThres = 1e-5;
    const bool rateCondition = ((ratio - bestRatio) > equalThres) && (cntMatch > bestCnt);
    const bool similarityCondition = ((ratio - bestRatio) < equalThres) && (cntMatch == bestCnt) && h_err[h] < bestErr;
    if(rateCondition || similarityCondition)
    {
      bestRatio = ratio;
      bestCnt = cntMatch;
      bestErr = h_err[h];
      *best = h;
    }
  }
  if(cnt_match) memcpy(cnt_match, h_cnt, sizeof(int) * n_hyp);
  if(max_cnt_match) memcpy(max_cnt_match, h_max, sizeof(int) * n_hyp);
  if(err_sum) memcpy(err_sum, h_err, sizeof(double) * n_hyp);
  best_transform(*best, hyps, model, scene, phi_m, phi_s, t_best);
  return TSD_OK;
}

int match_score_pdf(tsd_matcher_t* m, int32_t n_hyp, const tsd_hypothesis_t* hyps, int32_t n, const double* model,
                    const double* scene, const double* phi_m, const double* phi_s, double phi_max,
                    int32_t n_control, const double* control, int32_t n_valid, const double* model_angles,
                    const double* model_dists, const double params[12], double* prob, int32_t* fov_count,
                    int32_t* best, double t_best[9])
{
  TSD_LOCK(m);
  if(!m || n_hyp < 0 || n <= 0 || !hyps || !model || !scene || !phi_m || !phi_s || n_control < 0 ||
     (n_control > 0 && !control) || n_valid <= 0 || !model_angles || !model_dists || !params || !best || !t_best)
    return TSD_E_INVALID;
  set_identity3(t_best);
  *best = -1;
  if(n_hyp == 0) return TSD_OK;
  TSD_CUDA(cudaSetDevice(m->device));
  const size_t inBytes = a16(sizeof(tsd_hypothesis_t) * n_hyp) + 2 * a16(sizeof(double) * 2 * n) + 2 * a16(sizeof(double) * n) +
                         a16(sizeof(double) * 3 * (n_control + 1)) + 2 * a16(sizeof(double) * n_valid);
  const size_t outBytes = a16(sizeof(double) * n_hyp) + a16(sizeof(int) * n_hyp) + 16;
  int rc = ensure(m, inBytes + outBytes);
  if(rc) return rc;
  TSD_CUDA(cudaStreamSynchronize(m->stream));
  Arena a{(unsigned char*)m->h_buf, (unsigned char*)m->d_buf, 0};
  HypCommon hc;
  hc.n_hyp = n_hyp;
  hc.hyps = a.put(hyps, n_hyp);
  hc.model = a.put(model, 2 * (size_t)n);
  hc.scene = a.put(scene, 2 * (size_t)n);
  hc.phi_m = a.put(phi_m, n);
  hc.phi_s = a.put(phi_s, n);
  hc.phi_max = phi_max;
  hc.n_control = n_control;
  hc.control = a.put(control, 3 * (size_t)n_control + (n_control == 0 ? 1 : 0));
  PdfParams pp;
  pp.n_valid = n_valid;
  pp.sorted = 1;
  for(int k = 1; k < n_valid; k++)
    if(!(model_angles[k - 1] < model_angles[k])) { pp.sorted = 0; break; }
  pp.model_angles = a.put(model_angles, n_valid);
  pp.model_dists = a.put(model_dists, n_valid);
  for(int i = 0; i < 12; i++) pp.p[i] = params[i];
  const size_t inEnd = a.off;
  double* h_prob;
  int* h_fov;
  int* h_best;
  pp.prob = a.put<double>(nullptr, n_hyp, &h_prob);
  pp.fov_count = a.put<int>(nullptr, n_hyp, &h_fov);
  int* d_best = a.put<int>(nullptr, 4, &h_best);
  TSD_CUDA(cudaMemcpyAsync(m->d_buf, m->h_buf, inEnd, cudaMemcpyHostToDevice, m->stream));
  const size_t smem = sizeof(double) * (3 * (size_t)n_control + 2 * (size_t)n_valid + 2);
  if(smem > 200 * 1024) { set_error("control set / model too large for shared memory"); return TSD_E_INVALID; }
  if(smem > 48 * 1024) TSD_CUDA(cudaFuncSetAttribute(k_score_pdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sm = 148;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, m->device);
  k_score_pdf<<<grid_for(n_hyp, sm), MATCH_WARPS * 32, smem, m->stream>>>(hc, pp);
  TSD_LAUNCHED();
  // accept: prob > bestProb && fieldOfViewCount > pointsInControl * percentagePointsInC (PDFMatching.cpp:373)
  k_first_max<<<1, 1024, 0, m->stream>>>(n_hyp, pp.prob, pp.fov_count, (double)(unsigned)n_control * params[5], d_best);
  TSD_LAUNCHED();
  TSD_CUDA(cudaMemcpyAsync(h_prob, pp.prob, a.off - inEnd, cudaMemcpyDeviceToHost, m->stream));
  TSD_CUDA(cudaStreamSynchronize(m->stream));
  if(prob) memcpy(prob, h_prob, sizeof(double) * n_hyp);
  if(fov_count) memcpy(fov_count, h_fov, sizeof(int) * n_hyp);
  *best = *h_best;
  best_transform(*best, hyps, model, scene, phi_m, phi_s, t_best);
  return TSD_OK;
}

}  // extern "C"
