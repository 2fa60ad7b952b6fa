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
#include "nn_bounds.cuh"

using namespace tsd;

struct tsd_matcher
{
  std::recursive_mutex* mtx;
  int device;
  cudaStream_t stream;
  size_t cap;
  void* d_buf;
  void* h_buf;  // pinned
  // match_prepare's result: one block on the device and its mirror in pinned host memory, same layout, so that a host
  // pointer into the mirror names its device twin (resident())
  size_t prep_cap, prep_bytes;
  unsigned char* d_prep;
  unsigned char* h_prep;
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
#define RNM_GROUP 16   // valid model points per bounding box
#define RNM_SUPER 8    // boxes per super-box

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
  // Two levels: groups of RNM_GROUP points, and super-groups of RNM_SUPER groups whose box is tested first (per control
  // point: ~9 + 2 x 8 box tests and 2-3 scans of 16 points instead of 34 box tests and 2-3 scans of 32: 6.2 -> 5.8 ms).
  // (Measured and dropped, profiles/r02_notes.md: opening a box for the whole warp if any lane wants it, with a vote that
  //  skips the exact distance of a point no lane can use: 8.9 ms -- the union of 32 lanes' boxes is most of the model.)
  const int nGroups = (rp.n_valid + RNM_GROUP - 1) / RNM_GROUP;
  const int nSuper = (nGroups + RNM_SUPER - 1) / RNM_SUPER;
  float4* s_boxf = reinterpret_cast<float4*>((reinterpret_cast<uintptr_t>(s_mphi + rp.n_valid) + 15) & ~(uintptr_t)15);  // per group: x0 x1 y0 y1, rounded outward
  float4* s_sboxf = s_boxf + nGroups;                               // per super-group
  float* s_mabs = reinterpret_cast<float*>(s_sboxf + nSuper);       // [0]: largest |coordinate| of the model
  __syncthreads();
  for(int g = threadIdx.x; g < nGroups; g += blockDim.x)
  {
    double x0 = s_mx[RNM_GROUP * g], x1 = x0, y0 = s_my[RNM_GROUP * g], y1 = y0;
    for(int k = RNM_GROUP * g + 1; k < min(RNM_GROUP * g + RNM_GROUP, rp.n_valid); k++)
    {
      x0 = fmin(x0, s_mx[k]); x1 = fmax(x1, s_mx[k]);
      y0 = fmin(y0, s_my[k]); y1 = fmax(y1, s_my[k]);
    }
    s_boxf[g] = make_float4(__double2float_rd(x0), __double2float_ru(x1), __double2float_rd(y0), __double2float_ru(y1));
  }
  __syncthreads();
  for(int sg = threadIdx.x; sg < nSuper; sg += blockDim.x)
  {
    float4 b = s_boxf[RNM_SUPER * sg];
    for(int g = RNM_SUPER * sg + 1; g < min(RNM_SUPER * sg + RNM_SUPER, nGroups); g++)
    {
      const float4 c = s_boxf[g];
      b.x = fminf(b.x, c.x); b.y = fmaxf(b.y, c.y); b.z = fminf(b.z, c.z); b.w = fmaxf(b.w, c.w);
    }
    s_sboxf[sg] = b;
  }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    float a = 0.f;
    for(int g = 0; g < nSuper; g++)
      a = fmaxf(a, fmaxf(fmaxf(fabsf(s_sboxf[g].x), fabsf(s_sboxf[g].y)), fmaxf(fabsf(s_sboxf[g].z), fabsf(s_sboxf[g].w))));
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
      const float e = tsd_nb_err(xf, yf, mabs);
      auto scan_group = [&](int g)
      {
        const int k1 = min(RNM_GROUP * g + RNM_GROUP, rp.n_valid);
        for(int k = RNM_GROUP * g; k < k1; k++)
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
      const int seed = (prevBest >= 0) ? (prevBest / RNM_GROUP) : -1;
      if(seed >= 0) scan_group(seed);
      // lower bound of the squared distance to anything in a box
      auto box_lb = [&](const float4 b) { return tsd_nb_box_lb(b, xf, yf, e); };
      for(int sg = 0; sg < nSuper; sg++)
      {
        if(!(box_lb(s_sboxf[sg]) <= bdf)) continue;
        const int g1 = min(RNM_SUPER * sg + RNM_SUPER, nGroups);
        for(int g = RNM_SUPER * sg; g < g1; g++)
          if(g != seed && box_lb(s_boxf[g]) <= bdf) scan_group(g);
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

// ---------------------------------------------------------------- matcher pre-processing on the device
// RandomMatching.cpp:41-183 (extractSamples, pickControlSet, calcNormals, calcPhi, subsampleMask) and the trial /
// hypothesis enumeration of the three matchers (TSD_PDFMatching.cpp:59-205 == RandomNormalMatching.cpp:94-247 ==
// PDFMatching.cpp:67-233), so that a relocalisation with 10^5 hypotheses needs no host work between the scan and the
// scores.  Two DELIBERATE departures from the reference, both confined to this entry point (the adapter's match() keeps
// the reference's host code and libc rand() so that its goldens can be replayed):
//   * random numbers come from a counter-based generator (SplitMix64 finaliser over (seed, stream, index)) instead of the
//     sequential rand(): "keep point i" is rng(seed, 0, i) % 1000 >= threshold as in subsampleMask; "draw K of the valid
//     indices without replacement, in random order" is "the K smallest of the keys rng(seed, stream, index), in key order"
//     -- the same distribution as erasing random elements of a shrinking vector, but a function of (seed, index) that
//     every thread can evaluate on its own;
//   * the centroid inside pcaAnalysis is a running mean in double precision (gsl_stats_mean uses long double, which the
//     device does not have): normals agree with the reference's to ~1e-15.
__host__ __device__ __forceinline__ uint64_t tsd_rng(uint64_t seed, uint32_t stream, uint32_t idx)
{
  uint64_t z = seed + 0x9E3779B97F4A7C15ULL * ((((uint64_t)stream << 32) | idx) + 1ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

#define PREP_MAX_WINDOW 32

__device__ double prep_nrm2(const double* x, int n, int stride)  // gslcblas source_nrm2_r.h
{
  double scale = 0.0, ssq = 1.0;
  if(n == 1) return fabs(x[0]);
  for(int i = 0; i < n; i++)
  {
    const double v = x[i * stride];
    if(v != 0.0)
    {
      const double ax = fabs(v);
      if(scale < ax) { ssq = 1.0 + ssq * (scale / ax) * (scale / ax); scale = ax; }
      else { ssq += (ax / scale) * (ax / scale); }
    }
  }
  return scale * sqrt(ssq);
}

// obvious::Matrix::pcaAnalysis for a rows x 2 matrix (gsl/Matrix.cpp:227-327: centroid, M'M, one-sided Jacobi SVD,
// projections, extents); only the short axis' direction and the two squared lengths are handed back
__device__ void prep_pca(const double* Ain, int rows, double* xShort, double* yShort, double* lenLongSqr, double* lenShortSqr)
{
  const double eps = 2.2204460492503131e-16;
  double M[2 * PREP_MAX_WINDOW];
  double cent[2];
  for(int c = 0; c < 2; c++)
  {
    double mean = 0.0;
    for(int i = 0; i < rows; i++) mean += (Ain[2 * i + c] - mean) / (double)(i + 1);
    cent[c] = mean;
  }
  for(int c = 0; c < 2; c++)
    for(int i = 0; i < rows; i++) M[2 * i + c] = Ain[2 * i + c] + -cent[c];
  double A[4] = {0.0, 0.0, 0.0, 0.0};
  for(int k = 0; k < rows; k++)
    for(int i = 0; i < 2; i++)
    {
      const double temp = 1.0 * M[2 * k + i];
      if(temp != 0.0)
        for(int j = 0; j < 2; j++) A[2 * i + j] += temp * M[2 * k + j];
    }
  double V[4] = {1.0, 0.0, 0.0, 1.0}, S[2];
  const double tolerance = 10 * 2 * eps;
  for(int j = 0; j < 2; j++) S[j] = eps * prep_nrm2(A + j, 2, 2);
  int count = 1, sweep = 0;
  while(count > 0 && sweep <= 12)
  {
    count = 1;
    double pp = 0.0;
    for(int i = 0; i < 2; i++) pp += A[2 * i] * A[2 * i + 1];
    pp *= 2.0;
    const double a = prep_nrm2(A, 2, 2), b = prep_nrm2(A + 1, 2, 2);
    const double q = a * a - b * b;
    const double v = hypot(pp, q);
    const double abserr_a = S[0], abserr_b = S[1];
    const bool sorted = (a >= b), orthog = (fabs(pp) <= tolerance * (a * b)), noisya = (a < abserr_a), noisyb = (b < abserr_b);
    if(sorted && (orthog || noisya || noisyb)) count--;
    else
    {
      double cosine, sine;
      if(v == 0 || !sorted) { cosine = 0.0; sine = 1.0; }
      else
      {
        cosine = sqrt((v + q) / (2.0 * v));
        sine = pp / (2.0 * v * cosine);
      }
      for(int i = 0; i < 2; i++)
      {
        const double Aik = A[2 * i + 1], Aij = A[2 * i];
        A[2 * i] = Aij * cosine + Aik * sine;
        A[2 * i + 1] = -Aij * sine + Aik * cosine;
      }
      S[0] = fabs(cosine) * abserr_a + fabs(sine) * abserr_b;
      S[1] = fabs(sine) * abserr_a + fabs(cosine) * abserr_b;
      for(int i = 0; i < 2; i++)
      {
        const double Qij = V[2 * i], Qik = V[2 * i + 1];
        V[2 * i] = Qij * cosine + Qik * sine;
        V[2 * i + 1] = -Qij * sine + Qik * cosine;
      }
    }
    sweep++;
  }
  // extents of the projections on the two axes; the axes' end points are centre -+ V(:, i) * ext / 2, so their
  // difference is V(:, i) * ext up to the rounding of (c + e) - (c - e), which is replayed
  double ext[2], mid[2];
  for(int i = 0; i < 2; i++)
  {
    double mx = 0.0, mn = 0.0;
    for(int j = 0; j < rows; j++)
    {
      double temp = 0.0;
      for(int k = 0; k < 2; k++) temp += V[2 * k + i] * M[2 * j + k];
      const double pj = 0.0 + 1.0 * temp;
      if(j == 0) { mx = pj; mn = pj; }
      else { if(pj > mx) mx = pj; if(pj < mn) mn = pj; }
    }
    ext[i] = mx - mn;
    mid[i] = (ext[i] > 1e-6) ? (mx + mn) / 2.0 : 0.0;
  }
  for(int i = 0; i < 2; i++)
    for(int j = 0; j < 2; j++) cent[j] += V[2 * j + i] * mid[i];
  double d[2][2];
  for(int i = 0; i < 2; i++)
    for(int j = 0; j < 2; j++)
    {
      const double e = V[2 * j + i] * ext[i] / 2.0;
      d[i][j] = (cent[j] + e) - (cent[j] - e);
    }
  *lenLongSqr = d[0][0] * d[0][0] + d[0][1] * d[0][1];
  *xShort = d[1][0];
  *yShort = d[1][1];
  *lenShortSqr = d[1][0] * d[1][0] + d[1][1] * d[1][1];
}

// RandomMatching::calcNormals (RandomMatching.cpp:77-146) + calcPhi (:148-169) for point i
__device__ void prep_normal(const double* P, const uint8_t* maskIn, uint8_t* maskOut, double* N, double* phi, int points, int r, int i)
{
  if(i < r || i >= points - r) { maskOut[i] = 0; phi[i] = -1e6; N[2 * i] = 0.0; N[2 * i + 1] = 0.0; return; }
  N[2 * i] = 0.0; N[2 * i + 1] = 0.0;
  if(maskIn[i])
  {
    double A[2 * PREP_MAX_WINDOW];
    int cnt = 0;
    for(int j = -r; j < r; j++)
      if(maskIn[i + j]) { A[2 * cnt] = P[2 * (i + j)]; A[2 * cnt + 1] = P[2 * (i + j) + 1]; cnt++; }
    if(cnt > 3)
    {
      double xs, ys, ll, ls;
      prep_pca(A, cnt, &xs, &ys, &ll, &ls);
      if(ls > 1e-6 && (ll / ls) < 4.0) maskOut[i] = 0;
      else
      {
        const double len = sqrt(ls);
        if((P[2 * i] * xs + P[2 * i + 1] * ys) < 0.0) { N[2 * i] = xs / len; N[2 * i + 1] = ys / len; }
        else { N[2 * i] = -xs / len; N[2 * i + 1] = -ys / len; }
      }
    }
    else maskOut[i] = 0;
  }
  phi[i] = maskOut[i] ? atan2(N[2 * i + 1], N[2 * i]) : -1e6;
}

struct PrepParams
{
  int n, r;
  unsigned size_control, trials;
  int span;
  uint64_t seed;
  const double* M;
  const double* S;
  const uint8_t* maskM;
  const uint8_t* maskS;
  uint8_t* maskMpca;
  uint8_t* maskSpca;
  double* NM;   // scratch n x 2
  double* NS;
  double* phiM;
  double* phiS;
  int* idxM;    // n
  int* idxS;
  int* idxControl;
  int* idxTrials;
  int* prefS;   // n + 1: number of valid (post-PCA) scene points before index i
  unsigned* hypOff;  // trials + 1
  double* control;   // 3 x C
  double* phiControl;
  double* modelValid;  // nValidM x 2
  double* phiValid;
  double* modelAngles;
  double* modelDists;
  unsigned long long* keys;  // scratch, n
  int* header;       // [0] nValidM [1] nValidS [2] nControl [3] nTrials [4] nHyp
  double* headerD;   // [0] thetaMin [1] thetaMax
};

// ordered compaction of the indices i in [r, n - r) with mask[i] (extractSamples); one block
__device__ int prep_compact(const uint8_t* mask, int n, int r, int* out, unsigned* s_scan)
{
  const int tid = threadIdx.x, lane = tid & 31, nw = blockDim.x >> 5;
  int base = 0;
  for(int i0 = 0; i0 < n; i0 += blockDim.x)
  {
    const int i = i0 + tid;
    const bool v = i >= r && i < n - r && mask[i];
    const unsigned bal = __ballot_sync(0xffffffffu, v);
    if(lane == 0) s_scan[tid >> 5] = __popc(bal);
    __syncthreads();
    int off = base, total = 0;
    for(int w = 0; w < nw; w++)
    {
      const int c = (int)s_scan[w];
      if(w < (tid >> 5)) off += c;
      total += c;
    }
    if(v) out[off + __popc(bal & ((1u << lane) - 1u))] = i;
    base += total;
    __syncthreads();
  }
  return base;
}

// the K entries of idx[0..nv) with the smallest (rng(seed, stream, idx), idx): in that order, or (ascending) in index order
__device__ void prep_pick(const int* idx, int nv, int K, uint64_t seed, uint32_t stream, int* out, bool ascending, int* s_tmp,
                          unsigned long long* keys, unsigned* s_scan)
{
  for(int a = threadIdx.x; a < nv; a += blockDim.x) keys[a] = tsd_rng(seed, stream, (uint32_t)idx[a]);
  __syncthreads();
  // rank of every entry among the keys; picked <=> rank < K
  for(int a = threadIdx.x; a < nv; a += blockDim.x)
  {
    const uint64_t ka = keys[a];
    int rank = 0;
    for(int b = 0; b < nv; b++)
    {
      const uint64_t kb = keys[b];
      rank += (kb < ka || (kb == ka && b < a)) ? 1 : 0;
    }
    if(!ascending) { if(rank < K) out[rank] = idx[a]; }
    else s_tmp[a] = rank < K ? 1 : 0;
  }
  if(!ascending) return;
  // idx is ascending: the place of a picked entry is the number of picked entries before it (ordered compaction)
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, nw = blockDim.x >> 5;
  int base = 0;
  for(int a0 = 0; a0 < nv; a0 += blockDim.x)
  {
    const int a = a0 + tid;
    const bool v = a < nv && s_tmp[a] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, v);
    if(lane == 0) s_scan[tid >> 5] = __popc(bal);
    __syncthreads();
    int off = base, total = 0;
    for(int w = 0; w < nw; w++)
    {
      const int c = (int)s_scan[w];
      if(w < (tid >> 5)) off += c;
      total += c;
    }
    if(v) out[off + __popc(bal & ((1u << lane) - 1u))] = idx[a];
    base += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) k_prepare(PrepParams pp)
{
  __shared__ unsigned s_scan[40];
  const int tid = threadIdx.x, n = pp.n, r = pp.r;
  // ----- model: normals, orientations, valid indices
  for(int i = tid; i < n; i += blockDim.x) pp.maskMpca[i] = pp.maskM[i];
  __syncthreads();
  for(int i = tid; i < n; i += blockDim.x) prep_normal(pp.M, pp.maskM, pp.maskMpca, pp.NM, pp.phiM, n, r, i);
  __syncthreads();
  const int nvm = prep_compact(pp.maskMpca, n, r, pp.idxM, s_scan);
  // ----- scene: subsampling (RandomNormalMatching.cpp:126-131, RandomMatching.cpp:171-183), then the same
  int valid = 0;
  for(int i = tid; i < n; i += blockDim.x) valid += pp.maskS[i] ? 1 : 0;
  valid = __reduce_add_sync(0xffffffffu, valid);
  if((tid & 31) == 0) s_scan[tid >> 5] = (unsigned)valid;
  __syncthreads();
  valid = 0;
  for(int w = 0; w < (int)(blockDim.x >> 5); w++) valid += (int)s_scan[w];
  __syncthreads();
  double probability = 180.0 / (double)valid;
  int thresh = 0;
  if(probability < 0.99)
  {
    if(probability < 0.0) probability = 0.0;
    thresh = (int)(1000.0 - probability * 1000.0 + 0.5);
  }
  for(int i = tid; i < n; i += blockDim.x)
  {
    uint8_t keep = pp.maskS[i];
    if(thresh > 0 && (int)(tsd_rng(pp.seed, 0u, (uint32_t)i) % 1000ULL) < thresh) keep = 0;
    pp.maskSpca[i] = keep;
  }
  __syncthreads();
  for(int i = tid; i < n; i += blockDim.x) prep_normal(pp.S, pp.maskS, pp.maskSpca, pp.NS, pp.phiS, n, r, i);
  __syncthreads();
  const int nvs = prep_compact(pp.maskSpca, n, r, pp.idxS, s_scan);
  // ----- control set (pickControlSet) and trial order
  const int C = min((int)pp.size_control, nvs);
  const int T = min((int)pp.trials, nvm);
  // The control set is handed out in scene (= contour) order: it is a set -- every scorer sums or multiplies over it -- and
  // lanes that hold neighbouring control points want the same parts of the model (k_score_rnm: 4.9 -> 2.2 ms on 38 k
  // hypotheses against the random order pickControlSet leaves behind).  The trials keep their random order: it is the
  // order of the hypothesis list, and the first best hypothesis wins.
  prep_pick(pp.idxS, nvs, C, pp.seed, 1u, pp.idxControl, true, pp.prefS, pp.keys, s_scan);
  __syncthreads();
  prep_pick(pp.idxM, nvm, T, pp.seed, 2u, pp.idxTrials, false, nullptr, pp.keys, s_scan);
  __syncthreads();
  // valid scene points before index i
  if(tid == 0)
  {
    int acc = 0;
    for(int i = 0; i < n; i++) { pp.prefS[i] = acc; acc += pp.maskSpca[i] ? 1 : 0; }
    pp.prefS[n] = acc;
  }
  __syncthreads();
  for(int c = tid; c < C; c += blockDim.x)
  {
    const int idx = pp.idxControl[c];
    pp.control[c] = pp.S[2 * idx];
    pp.control[C + c] = pp.S[2 * idx + 1];
    pp.control[2 * C + c] = 1.0;
    pp.phiControl[c] = atan2(pp.NS[2 * idx + 1], pp.NS[2 * idx]);
  }
  for(int k = tid; k < nvm; k += blockDim.x)
  {
    const int idx = pp.idxM[k];
    const double x = pp.M[2 * idx], y = pp.M[2 * idx + 1];
    pp.modelValid[2 * k] = x;
    pp.modelValid[2 * k + 1] = y;
    pp.phiValid[k] = pp.phiM[idx];
    pp.modelAngles[k] = atan2(y, x);
    pp.modelDists[k] = sqrt(x * x + y * y);
  }
  // ----- hypotheses per trial: the valid scene indices in [max(idx - span, r), min(idx + span, n - r)); their offsets in
  // the list are an exclusive prefix sum over the trials (block scan, 1024 trials at a time)
  {
    const int lane = tid & 31, nw = blockDim.x >> 5;
    unsigned base = 0;
    for(int t0 = 0; t0 < T; t0 += blockDim.x)
    {
      const int t = t0 + tid;
      unsigned c = 0;
      if(t < T)
      {
        const int idx = pp.idxTrials[t];
        const int iMin = max(idx - pp.span, r), iMax = min(idx + pp.span, n - r);
        if(iMax > iMin) c = (unsigned)(pp.prefS[iMax] - pp.prefS[iMin]);
      }
      unsigned incl = c;
#pragma unroll
      for(int o = 1; o < 32; o <<= 1)
      {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if(lane >= o) incl += v;
      }
      if(lane == 31) s_scan[tid >> 5] = incl;
      __syncthreads();
      unsigned off = base, total = 0;
      for(int w = 0; w < nw; w++)
      {
        const unsigned v = s_scan[w];
        if(w < (tid >> 5)) off += v;
        total += v;
      }
      if(t < T) pp.hypOff[t] = off + incl - c;
      base += total;
      __syncthreads();
    }
    if(tid == 0) pp.hypOff[T] = base;
    if(tid == 0) s_scan[39] = base;
    __syncthreads();
  }
  if(tid == 0)
  {
    const unsigned acc = s_scan[39];
    pp.header[0] = nvm; pp.header[1] = nvs; pp.header[2] = C; pp.header[3] = T; pp.header[4] = (int)acc;
    if(nvm > 0)
    {
      pp.headerD[0] = atan2(pp.M[2 * pp.idxM[0] + 1], pp.M[2 * pp.idxM[0]]);
      pp.headerD[1] = atan2(pp.M[2 * pp.idxM[nvm - 1] + 1], pp.M[2 * pp.idxM[nvm - 1]]);
    }
    else { pp.headerD[0] = 0.0; pp.headerD[1] = 0.0; }
  }
}

// block t writes the hypotheses of trial t, in scene order, at their place in the list
__global__ void __launch_bounds__(256) k_prepare_emit(PrepParams pp, tsd_hypothesis_t* hyps, unsigned cap)
{
  const int T = pp.header[3];
  for(int t = blockIdx.x; t < T; t += gridDim.x)
  {
    const int idx = pp.idxTrials[t];
    const int iMin = max(idx - pp.span, pp.r), iMax = min(idx + pp.span, pp.n - pp.r);
    const unsigned off = pp.hypOff[t];
    for(int i = iMin + (int)threadIdx.x; i < iMax; i += blockDim.x)
      if(pp.maskSpca[i])
      {
        const unsigned at = off + (unsigned)(pp.prefS[i] - pp.prefS[iMin]);
        if(at < cap) { hyps[at].idx_model = idx; hyps[at].idx_scene = i; }
      }
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

// An input array that lies inside the mirror of the prepared set (match_prepare) is already on the device: its twin sits
// at the same offset of the device block.  Everything else is staged and uploaded as before.
template <typename T>
static const T* put_in(tsd_matcher* m, Arena& a, const T* src, size_t count)
{
  const unsigned char* p = reinterpret_cast<const unsigned char*>(src);
  if(src && m->h_prep && p >= m->h_prep && p + sizeof(T) * count <= m->h_prep + m->prep_bytes)
    return reinterpret_cast<const T*>(m->d_prep + (p - m->h_prep));
  return a.put(src, count);
}

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
  cudaFree(m->d_prep);
  cudaFreeHost(m->h_prep);
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
  hc.hyps = put_in(m, a, hyps, n_hyp);
  hc.model = put_in(m, a, model, 2 * (size_t)n);
  hc.scene = put_in(m, a, scene, 2 * (size_t)n);
  hc.phi_m = put_in(m, a, phi_m, n);
  hc.phi_s = put_in(m, a, phi_s, n);
  hc.phi_max = phi_max;
  hc.n_control = n_control;
  hc.control = put_in(m, a, control, 3 * (size_t)n_control + (n_control == 0 ? 1 : 0));
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
  hc.hyps = put_in(m, a, hyps, n_hyp);
  hc.model = put_in(m, a, model, 2 * (size_t)n);
  hc.scene = put_in(m, a, scene, 2 * (size_t)n);
  hc.phi_m = put_in(m, a, phi_m, n);
  hc.phi_s = put_in(m, a, phi_s, n);
  hc.phi_max = phi_max;
  hc.n_control = n_control;
  hc.control = put_in(m, a, control, 3 * (size_t)n_control + (n_control == 0 ? 1 : 0));
  RnmParams rp;
  rp.phi_control = put_in(m, a, phi_control, (size_t)n_control + (n_control == 0 ? 1 : 0));
  rp.n_valid = n_valid;
  rp.model_valid = put_in(m, a, model_valid, 2 * (size_t)n_valid + (n_valid == 0 ? 1 : 0));
  rp.phi_valid = put_in(m, a, phi_valid, (size_t)n_valid + (n_valid == 0 ? 1 : 0));
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
  const size_t nGroupsH = ((size_t)n_valid + RNM_GROUP - 1) / RNM_GROUP;
  const size_t smem = sizeof(double) * (4 * (size_t)n_control + 3 * (size_t)n_valid) + 16 * (nGroupsH + (nGroupsH + RNM_SUPER - 1) / RNM_SUPER) + 64;
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
  hc.hyps = put_in(m, a, hyps, n_hyp);
  hc.model = put_in(m, a, model, 2 * (size_t)n);
  hc.scene = put_in(m, a, scene, 2 * (size_t)n);
  hc.phi_m = put_in(m, a, phi_m, n);
  hc.phi_s = put_in(m, a, phi_s, n);
  hc.phi_max = phi_max;
  hc.n_control = n_control;
  hc.control = put_in(m, a, control, 3 * (size_t)n_control + (n_control == 0 ? 1 : 0));
  PdfParams pp;
  pp.n_valid = n_valid;
  pp.sorted = 1;
  for(int k = 1; k < n_valid; k++)
    if(!(model_angles[k - 1] < model_angles[k])) { pp.sorted = 0; break; }
  pp.model_angles = put_in(m, a, model_angles, n_valid);
  pp.model_dists = put_in(m, a, model_dists, n_valid);
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

int match_prepare(tsd_matcher_t* m, int32_t n, const double* model, const uint8_t* mask_m, const double* scene,
                  const uint8_t* mask_s, int32_t pca_search_range, uint32_t size_control_set, uint32_t trials, double phi_max,
                  double resolution, uint64_t seed, tsd_match_prep_t* out)
{
  TSD_LOCK(m);
  if(!m || !out || n < 3 || !model || !mask_m || !scene || !mask_s) return TSD_E_INVALID;
  memset(out, 0, sizeof(*out));
  const int r = pca_search_range / 2;
  if(r < 1 || 2 * r > PREP_MAX_WINDOW) { set_error("match_prepare: pca_search_range must be 2 .. %d", PREP_MAX_WINDOW); return TSD_E_INVALID; }
  if(n > 4096) { set_error("match_prepare supports at most 4096 points"); return TSD_E_INVALID; }
  // RandomNormalMatching.cpp:184-197
  if(!(phi_max <= M_PI * 0.5)) phi_max = M_PI * 0.5;
  if(!(resolution > 1e-6)) { set_error("match_prepare: resolution not properly set (%g)", resolution); return TSD_E_INVALID; }
  int span = (int)floor(phi_max / resolution);
  if(span > n) span = n;
  if(span < 0) span = 0;
  const size_t tmax = std::min<size_t>(trials, (size_t)n), cmax = std::min<size_t>(size_control_set, (size_t)n);
  const size_t capHyp = tmax * std::min<size_t>(2 * (size_t)span, (size_t)n) + 1;
  TSD_CUDA(cudaSetDevice(m->device));
  TSD_CUDA(cudaStreamSynchronize(m->stream));
  // layout of the block (identical on both sides)
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t at = off; off += a16(bytes); return at; };
  const size_t oM = take(16 * (size_t)n), oS = take(16 * (size_t)n), oMaskM = take(n), oMaskS = take(n);
  const size_t inEnd = off;
  const size_t oMaskMp = take(n), oMaskSp = take(n), oNM = take(16 * (size_t)n), oNS = take(16 * (size_t)n);
  const size_t oPhiM = take(8 * (size_t)n), oPhiS = take(8 * (size_t)n), oIdxM = take(4 * (size_t)n), oIdxS = take(4 * (size_t)n);
  const size_t oIdxC = take(4 * (cmax + 1)), oIdxT = take(4 * (tmax + 1)), oPref = take(4 * ((size_t)n + 1)), oHypOff = take(4 * (tmax + 2));
  const size_t oCtrl = take(8 * 3 * (cmax + 1)), oPhiC = take(8 * (cmax + 1));
  const size_t oMV = take(16 * (size_t)n), oPV = take(8 * (size_t)n), oAng = take(8 * (size_t)n), oDst = take(8 * (size_t)n);
  const size_t oHdr = take(32), oHdrD = take(16), oKeys = take(8 * (size_t)n);
  const size_t fixedEnd = off;
  const size_t oHyps = take(sizeof(tsd_hypothesis_t) * capHyp);
  if(off > m->prep_cap)
  {
    if(m->d_prep) cudaFree(m->d_prep);
    if(m->h_prep) cudaFreeHost(m->h_prep);
    m->d_prep = m->h_prep = nullptr;
    m->prep_cap = m->prep_bytes = 0;
    size_t cap = 1 << 20;
    while(cap < off) cap <<= 1;
    TSD_CUDA(cudaMalloc(&m->d_prep, cap));
    TSD_CUDA(cudaMallocHost(&m->h_prep, cap));
    m->prep_cap = cap;
  }
  m->prep_bytes = 0;  // (nothing is resident until this call has succeeded)
  unsigned char *H = m->h_prep, *D = m->d_prep;
  memcpy(H + oM, model, 16 * (size_t)n);
  memcpy(H + oS, scene, 16 * (size_t)n);
  memcpy(H + oMaskM, mask_m, n);
  memcpy(H + oMaskS, mask_s, n);
  TSD_CUDA(cudaMemcpyAsync(D, H, inEnd, cudaMemcpyHostToDevice, m->stream));
  PrepParams pp;
  pp.n = n; pp.r = r; pp.size_control = (unsigned)cmax; pp.trials = (unsigned)tmax; pp.span = span; pp.seed = seed;
  pp.M = (const double*)(D + oM); pp.S = (const double*)(D + oS);
  pp.maskM = D + oMaskM; pp.maskS = D + oMaskS; pp.maskMpca = D + oMaskMp; pp.maskSpca = D + oMaskSp;
  pp.NM = (double*)(D + oNM); pp.NS = (double*)(D + oNS); pp.phiM = (double*)(D + oPhiM); pp.phiS = (double*)(D + oPhiS);
  pp.idxM = (int*)(D + oIdxM); pp.idxS = (int*)(D + oIdxS); pp.idxControl = (int*)(D + oIdxC); pp.idxTrials = (int*)(D + oIdxT);
  pp.prefS = (int*)(D + oPref); pp.hypOff = (unsigned*)(D + oHypOff);
  pp.control = (double*)(D + oCtrl); pp.phiControl = (double*)(D + oPhiC);
  pp.modelValid = (double*)(D + oMV); pp.phiValid = (double*)(D + oPV); pp.modelAngles = (double*)(D + oAng); pp.modelDists = (double*)(D + oDst);
  pp.header = (int*)(D + oHdr); pp.headerD = (double*)(D + oHdrD); pp.keys = (unsigned long long*)(D + oKeys);
  k_prepare<<<1, 1024, 0, m->stream>>>(pp);
  TSD_LAUNCHED();
  k_prepare_emit<<<(unsigned)std::max<size_t>(tmax, 1), 256, 0, m->stream>>>(pp, (tsd_hypothesis_t*)(D + oHyps), (unsigned)capHyp);
  TSD_LAUNCHED();
  TSD_CUDA(cudaMemcpyAsync(H + inEnd, D + inEnd, fixedEnd - inEnd, cudaMemcpyDeviceToHost, m->stream));
  TSD_CUDA(cudaStreamSynchronize(m->stream));
  const int* hdr = (const int*)(H + oHdr);
  const double* hdrD = (const double*)(H + oHdrD);
  const int nvm = hdr[0], nvs = hdr[1];
  int nHyp = hdr[4];
  if((size_t)nHyp > capHyp) { set_error("match_prepare: hypothesis list overflow (%d > %zu)", nHyp, capHyp); return TSD_E_INVALID; }
  // RandomNormalMatching.cpp:160-170: too few valid points -> no match
  if(nvs < 3 || nvm < 3) nHyp = 0;
  if(nHyp > 0)
  {
    TSD_CUDA(cudaMemcpyAsync(H + oHyps, D + oHyps, sizeof(tsd_hypothesis_t) * (size_t)nHyp, cudaMemcpyDeviceToHost, m->stream));
    TSD_CUDA(cudaStreamSynchronize(m->stream));
  }
  m->prep_bytes = oHyps + sizeof(tsd_hypothesis_t) * (size_t)nHyp;
  out->n = n;
  out->n_hyp = nHyp;
  out->n_control = hdr[2];
  out->n_trials = hdr[3];
  out->n_valid_m = nvm;
  out->n_valid_s = nvs;
  out->span = span;
  out->phi_max = phi_max;
  out->theta_min = hdrD[0];
  out->theta_max = hdrD[1];
  out->hyps = (const tsd_hypothesis_t*)(H + oHyps);
  out->model = (const double*)(H + oM);
  out->scene = (const double*)(H + oS);
  out->phi_m = (const double*)(H + oPhiM);
  out->phi_s = (const double*)(H + oPhiS);
  out->mask_m_pca = H + oMaskMp;
  out->mask_s_pca = H + oMaskSp;
  out->idx_m_valid = (const int32_t*)(H + oIdxM);
  out->idx_s_valid = (const int32_t*)(H + oIdxS);
  out->idx_control = (const int32_t*)(H + oIdxC);
  out->idx_trials = (const int32_t*)(H + oIdxT);
  out->control = (const double*)(H + oCtrl);
  out->phi_control = (const double*)(H + oPhiC);
  out->model_valid = (const double*)(H + oMV);
  out->phi_valid = (const double*)(H + oPV);
  out->model_angles = (const double*)(H + oAng);
  out->model_dists = (const double*)(H + oDst);
  return TSD_OK;
}

uint64_t match_rng(uint64_t seed, uint32_t stream, uint32_t index) { return tsd_rng(seed, stream, index); }

}  // extern "C"
