// Scan pre-processing on the device: what ThreadLocalize does to a LaserScan between the ROS
// callback and the first kernel of the hot path --
//   Sensor::setRealMeasurementData(vector<float>, scale)   reference src/obvision/reconstruct/Sensor.cpp:136-145
//   SensorPolar2D::setStandardMask                          reconstruct/grid/SensorPolar2D.cpp:59-98
//       = resetMask, maskZeroDepth, maskInvalidDepth (Sensor.cpp:246-272), maskDepthDiscontinuity(3 deg)
//   Sensor::dataToCartesianVectorMask                       Sensor.cpp:168-190  (scene points, sensor frame)
//   ThreadLocalize::maskMatrix                              src/ThreadLocalize.cpp:738-755 (compaction of the scene)
// -- one launch of one CTA: a scan is ~1000 beams, the work is latency, not throughput.  Expression order follows the
// reference (no FMA contraction: the library is built with -fmad=false); sin / cos of the angular resolution and the
// threshold come from the host so that the only library function evaluated here is asin.
#include "common.cuh"

using namespace tsd;

struct ScanPrepParams
{
  int n;
  const float* in;     // LaserScan ranges
  float scale;
  double max_range, cosphi, sinphi, thresh;
  const double* rays_local;  // 2 x n (row-major): SensorPolar2D::_raysLocal
  double* data;              // n
  uint8_t* mask;             // n
  double* scene;             // n x 2, written where scene_mask != 0
  uint8_t* scene_mask;       // n
  double* scene_valid;       // n_valid x 2, beam order
  unsigned* n_valid;
};

#define PREP_THREADS 1024

__global__ void __launch_bounds__(PREP_THREADS) k_scan_prepare(ScanPrepParams p)
{
  __shared__ unsigned s_warp[PREP_THREADS / 32];
  __shared__ unsigned s_base;
  const int t = threadIdx.x;
  // Sensor.cpp:143-144, :250-272
  for(int i = t; i < p.n; i += PREP_THREADS)
  {
    double d = (double)(p.in[i] * p.scale);
    bool m = true;
    m = m && (d != 0.0);
    if(d > p.max_range) d = __longlong_as_double(0x7ff0000000000000LL);
    if(isnan(d))
    {
      m = false;
      d = __longlong_as_double(0x7ff0000000000000LL);
    }
    p.data[i] = d;
    p.mask[i] = m ? 1 : 0;
  }
  __syncthreads();
  // SensorPolar2D.cpp:67-98 (reads the final data of the neighbours, clears mask bits only)
  for(int i = t; i < p.n; i += PREP_THREADS)
  {
    if(i < 1 || i >= p.n - 1) continue;
    double betamin = 3.14159265358979323846;
    const double a = p.data[i];
    if(isinf(a)) continue;
    for(int j = -1; j <= 1; j++)
    {
      const double b = p.data[i + j];
      if(isinf(b)) continue;
      const double c = sqrt(a * a + b * b - 2 * a * b * p.cosphi);
      if(a > b)
      {
        const double beta = asin(b / c * p.sinphi);
        if(beta < betamin) betamin = beta;
      }
    }
    if(betamin < p.thresh) p.mask[i] = 0;
  }
  __syncthreads();
  // Sensor.cpp:168-190 and the compaction of ThreadLocalize.cpp:738-755 (a block-wide exclusive scan, chunk by chunk)
  if(t == 0) s_base = 0;
  __syncthreads();
  for(int i0 = 0; i0 < p.n; i0 += PREP_THREADS)
  {
    const int i = i0 + t;
    bool v = false;
    double x = 0.0, y = 0.0;
    if(i < p.n)
    {
      const double d = p.data[i];
      v = !isinf(d) && p.mask[i];
      if(v)
      {
        x = p.rays_local[i] * d;
        y = p.rays_local[p.n + i] * d;
        p.scene[2 * i] = x;
        p.scene[2 * i + 1] = y;
      }
      p.scene_mask[i] = v ? 1 : 0;
    }
    const unsigned b = __ballot_sync(0xffffffffu, v);
    if((t & 31) == 0) s_warp[t >> 5] = __popc(b);
    __syncthreads();
    unsigned before = s_base;
    for(int w = 0; w < (t >> 5); w++) before += s_warp[w];
    if(v)
    {
      const unsigned k = before + __popc(b & ((1u << (t & 31)) - 1u));
      p.scene_valid[2 * k] = x;
      p.scene_valid[2 * k + 1] = y;
    }
    __syncthreads();
    if(t == 0)
    {
      unsigned tot = 0;
      for(int w = 0; w < PREP_THREADS / 32; w++) tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
  }
  if(t == 0) *p.n_valid = s_base;
}

extern "C" int tsds_prepare_scan(tsd_grid_t* g, int32_t n, const float* ranges, float scale, double max_range, double angular_res,
                                 const double* rays_local, double* data, uint8_t* mask, double* scene, uint8_t* scene_mask,
                                 double* scene_valid, uint32_t* n_valid)
{
  TSD_LOCK(g);
  if(!g || n < 1 || !ranges || !rays_local || !data || !mask) return TSD_E_INVALID;
  TSD_CUDA(cudaSetDevice(g->device));
  // one scratch block: [ in f32 n | rays 2n f64 | data n f64 | scene 2n | scene_valid 2n | mask n | scene_mask n | count ]
  const size_t N = (size_t)n, pad = (N + 7) & ~(size_t)7;
  const size_t oRays = sizeof(float) * pad, oData = oRays + sizeof(double) * 2 * N, oScene = oData + sizeof(double) * N;
  const size_t oValid = oScene + sizeof(double) * 2 * N, oMask = oValid + sizeof(double) * 2 * N, oSMask = oMask + pad;
  const size_t oCnt = oSMask + pad, bytes = oCnt + 16;
  int rc = grid_ensure_scratch(g, bytes);
  if(rc) return rc;
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  unsigned char* h = (unsigned char*)g->h_scratch;
  unsigned char* d = (unsigned char*)g->d_scratch;
  memcpy(h, ranges, sizeof(float) * N);
  memcpy(h + oRays, rays_local, sizeof(double) * 2 * N);
  TSD_CUDA(cudaMemcpyAsync(d, h, oData, cudaMemcpyHostToDevice, g->stream));
  ScanPrepParams p;
  p.n = n;
  p.in = (const float*)d;
  p.scale = scale;
  p.max_range = max_range;
  sincos(angular_res, &p.sinphi, &p.cosphi);                  // SensorPolar2D.cpp:71
  p.thresh = 3.0 * 3.14159265358979323846 / 180.0;            // deg2rad(3.0), mathbase.h
  p.rays_local = (const double*)(d + oRays);
  p.data = (double*)(d + oData);
  p.scene = (double*)(d + oScene);
  p.scene_valid = (double*)(d + oValid);
  p.mask = d + oMask;
  p.scene_mask = d + oSMask;
  p.n_valid = (unsigned*)(d + oCnt);
  k_scan_prepare<<<1, PREP_THREADS, 0, g->stream>>>(p);
  TSD_LAUNCHED();
  TSD_CUDA(cudaMemcpyAsync(h + oData, d + oData, bytes - oData, cudaMemcpyDeviceToHost, g->stream));
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  memcpy(data, h + oData, sizeof(double) * N);
  memcpy(mask, h + oMask, N);
  const unsigned cnt = *(const unsigned*)(h + oCnt);
  if(scene_mask) memcpy(scene_mask, h + oSMask, N);
  if(scene)  // misses keep the caller's values, as Sensor.cpp:183-187 skips them
  {
    const double* s = (const double*)(h + oScene);
    const uint8_t* sm = h + oSMask;
    for(size_t i = 0; i < N; i++)
      if(sm[i]) { scene[2 * i] = s[2 * i]; scene[2 * i + 1] = s[2 * i + 1]; }
  }
  if(scene_valid) memcpy(scene_valid, h + oValid, sizeof(double) * 2 * cnt);
  if(n_valid) *n_valid = cnt;
  return TSD_OK;
}
