// Host-side proof obligation of k_update's single-precision front end (ohm_tsd_slam_b200/csrc/beam_index.cuh:
// tsd_fast_model / tsd_gate_entry / tsd_classify_cell): whenever the front end calls a cell "certain", its answer
// must be the answer of the reference's double-precision expressions
//   SensorPolar2D::backProject  (reference src/obvision/reconstruct/grid/SensorPolar2D.cpp:117-135)
//   TsdGrid::push per-cell part (TsdGrid.cpp:250-274) + TsdGridPartition::addTsd gate (TsdGridPartition.h:174-191)
// Runs the SAME functions the kernel runs (the header is host/device code), with the hardware's approximate
// reciprocal replaced by 1/x perturbed by up to +-2 ulp, on random sensor models, poses, cells and ranges, incl.
// cells placed adversarially next to beam boundaries and next to the truncation band.
// Usage: fastpath_check [cells per model]     exit code 0 = no violation; prints statistics.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <random>
#include <vector>

static int g_rcp_mode = 0;
static inline float rcp_test(float x)
{
  float r = 1.0f / x;
  switch(g_rcp_mode)
  {
    case 1: r = nextafterf(r, INFINITY); break;
    case 2: r = nextafterf(r, -INFINITY); break;
    case 3: r = nextafterf(nextafterf(r, INFINITY), INFINITY); break;
    case 4: r = nextafterf(nextafterf(r, -INFINITY), -INFINITY); break;
    default: break;
  }
  return r;
}
#define TSD_RCPF(x) rcp_test(x)
#include "../../ohm_tsd_slam_b200/csrc/beam_index.cuh"

struct Model { int n; double res, phi_min; };

static void invert_pose(double x, double y, double th, double P[9], double Pi[9])
{
  const double c = cos(th), s = sin(th);
  P[0] = c; P[1] = -s; P[2] = x; P[3] = s; P[4] = c; P[5] = y; P[6] = 0; P[7] = 0; P[8] = 1;
  // closed form is good enough here: the kernel takes whatever inverse the caller hands over
  Pi[0] = c; Pi[1] = s; Pi[2] = -(c * x + s * y); Pi[3] = -s; Pi[4] = c; Pi[5] = -(-s * x + c * y); Pi[6] = 0; Pi[7] = 0; Pi[8] = 1;
}

int main(int argc, char** argv)
{
  const long per_model = argc > 1 ? atol(argv[1]) : 2000000;
  const Model models[] = {{1081, M_PI / 720.0, -135.0 * M_PI / 180.0}, {361, M_PI / 240.0, -135.0 * M_PI / 180.0},
                          {541, M_PI / 360.0, -0.5 * M_PI - 0.3}, {2000, 0.0005, -0.4}, {720, M_PI / 360.0, -M_PI}};
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  long total = 0, certain = 0, cls_n[4] = {0, 0, 0, 0}, bad = 0, off_models = 0;
  const double cell = 0.025;
  for(const Model& md : models)
  {
    const double phi_lower = -0.5 * md.res + md.phi_min;                 // SensorPolar2D.cpp:26
    const double phi_upper = md.phi_min + ((double)md.n - 0.5) * md.res; // SensorPolar2D.cpp:30
    const double res_inv = 1.0 / md.res;
    for(int poseIdx = 0; poseIdx < 8; poseIdx++)
    {
      const double grid = (poseIdx & 1) ? 1638.4 : 102.4;
      const double tx = grid * (0.1 + 0.8 * U(rng)), ty = grid * (0.1 + 0.8 * U(rng)), th = (U(rng) * 2 - 1) * M_PI;
      double P[9], Pi[9];
      invert_pose(tx, ty, th, P, Pi);
      FastModel fm;
      tsd_fast_model(Pi, md.phi_min, md.res, phi_lower, phi_upper, md.n, &fm);
      if(fm.half_m < 0.0f) { off_models++; continue; }
      const double T = cell * (2.0 + 3.0 * U(rng)), invT = 1.0 / T, L = 0.5 + 2.0 * U(rng);
      // a scan
      std::vector<double> ranges(md.n);
      std::vector<uint8_t> mask(md.n);
      std::vector<float2> gate(md.n);
      const double base = 0.3 + U(rng) * ((poseIdx & 1) ? 800.0 : 30.0);
      for(int k = 0; k < md.n; k++)
      {
        double r = base * (0.7 + 0.6 * U(rng));
        const double u = U(rng);
        if(u < 0.03) r = INFINITY;
        else if(u < 0.05) r = 0.0;
        else if(u < 0.06) r = NAN;
        r = (double)(float)r;
        ranges[k] = r;
        mask[k] = (U(rng) < 0.97) ? 1 : 0;
        tsd_gate_entry(r, mask[k] != 0, T, L, &gate[k].x, &gate[k].y);
      }
      for(long it = 0; it < per_model / 8; it++)
      {
        // a cell centre of the grid, chosen by kind: anywhere / next to a beam boundary / next to the truncation band
        const int kind = (int)(U(rng) * 4);
        double d = exp(log(1e-3) + U(rng) * (log(base * 1.6) - log(1e-3)));
        double a = md.phi_min + (U(rng) * (md.n + 40) - 20.0) * md.res;   // sensor-frame angle, beyond the FOV too
        if(kind == 1)  // adversarial angle: a half-beam boundary +- up to 3e-3 beams
          a = md.phi_min + ((double)(int)(U(rng) * (md.n + 1)) - 0.5 + (U(rng) * 2 - 1) * 3e-3) * md.res;
        if(kind == 2)  // adversarial distance: the beam's range +- T +- a few 1e-5
        {
          int k = (int)lround((a - md.phi_min) * res_inv);
          if(k >= 0 && k < md.n && std::isfinite(ranges[k]))
            d = ranges[k] + ((U(rng) < 0.5) ? T : -T) + (U(rng) * 2 - 1) * 1e-4 * (1.0 + ranges[k] * 0.1);
          if(!(d > 1e-4)) d = 1e-3;
        }
        if(kind == 3) a = (U(rng) < 0.5 ? M_PI : -M_PI) + (U(rng) * 2 - 1) * 1e-6;  // the cut of atan2
        const double wx = tx + d * cos(a + th), wy = ty + d * sin(a + th);
        const long ix = lround(wx / cell - 0.5), iy = lround(wy / cell - 0.5);
        const double X = ((double)ix + 0.5) * cell, Y = ((double)iy + 0.5) * cell;  // TsdGridPartition.cpp:127-128
        // --- the reference (gslcblas accumulation order)
        double xs = 0.0; xs += Pi[0] * X; xs += Pi[1] * Y; xs += Pi[2] * 1.0;
        double ys = 0.0; ys += Pi[3] * X; ys += Pi[4] * Y; ys += Pi[5] * 1.0;
        const double phi = atan2(ys, xs);
        int idx;
        if(phi <= phi_lower) idx = -2;
        else if(phi >= phi_upper) idx = -1;
        else idx = (int)round((phi - md.phi_min) * res_inv);
        const double dist = sqrt((X - P[2]) * (X - P[2]) + (Y - P[5]) * (Y - P[5]));
        bool upd = false;
        double nref = 0.0;
        if(idx >= 0 && idx < md.n && mask[idx])
        {
          const double r = ranges[idx];
          if(std::isinf(r)) { if(dist < L) { upd = (T >= -T); nref = fmin(T * invT, 1.0); } }
          else { const double sd = r - dist; upd = sd >= -T; nref = fmin(sd * invT, 1.0); }
        }
        // --- the front end, on the tables fill_tables builds
        const float cxf = (float)(Pi[0] * (X - fm.txp)), cyf = (float)(Pi[3] * (X - fm.txp));
        const float rxf = (float)(Pi[1] * (Y - fm.typ)), ryf = (float)(Pi[4] * (Y - fm.typ));
        const float cdf = (float)((X - P[2]) * (X - P[2])), rdf = (float)((Y - P[5]) * (Y - P[5]));
        for(g_rcp_mode = 0; g_rcp_mode < 5; g_rcp_mode++)
        {
          int k = 0;
          const int cls = tsd_classify_cell(fm.rinv_f, fm.off_f, fm.half_m, md.n, gate.data(), cxf + rxf, cyf + ryf, cdf + rdf, k);
          if(g_rcp_mode == 0) { total++; cls_n[cls]++; if(cls != 3) certain++; }
          bool okc = true;
          if(cls == 3) continue;
          const bool inside = (unsigned)k < (unsigned)md.n;
          if(!inside) okc = (idx < 0) || (idx >= md.n);          // outside the field of view -> no update
          else
          {
            okc = (idx == k);
            if(cls == 0) okc = okc && !upd;
            if(cls == 1) okc = okc && upd && nref == 1.0 && !std::isinf(ranges[k]);
          }
          if(!okc)
          {
            if(bad < 10)
              fprintf(stderr, "VIOLATION model n=%d cls=%d k=%d idx=%d upd=%d n=%.17g dist=%.17g r=%.17g rcp_mode=%d phi=%.17g\n", md.n,
                      cls, k, idx, (int)upd, nref, dist, inside ? ranges[k] : -1.0, g_rcp_mode, phi);
            bad++;
          }
        }
      }
    }
  }
  printf("cells %ld certain %.4f%% classes: skip %ld free %ld exact-distance %ld exact-beam %ld; models with the front end off: %ld; "
         "violations %ld\n",
         total, 100.0 * certain / (double)(total ? total : 1), cls_n[0], cls_n[1], cls_n[2], cls_n[3], off_models, bad);
  return bad ? 1 : 0;
}
