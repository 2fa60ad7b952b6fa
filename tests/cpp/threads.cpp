// One grid handle used the way the node uses obvious::TsdGrid: a mapper thread pushes (ThreadMapping::eventLoop,
// reference src/ThreadMapping.cpp:43-62) while two localiser threads ray-cast and sample the same grid
// (ThreadLocalize::eventLoop, src/ThreadLocalize.cpp:310-409; two of them in the double-laser configuration,
// src/SlamNode.cpp:104-121), with no synchronisation between the threads.
//
// To make the outcome independent of the interleaving, the mapper integrates scans into the RIGHT part of the map (short
// range) and the localisers look at the LEFT part: every ray cast and every sample must then equal, bit for bit, what
// the same calls return in a single-threaded run, and so must the final map.  What the test exercises is the host side
// of the C ABI (staging buffers, streams, scratch memory shared by the calls on one handle).
//
//   threads <iterations>      exit code 0 = identical; prints a one-line summary
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/tsdslam_b200.h"

static const int N = 361;
static const double RES = M_PI / 240.0, PHI_MIN = -135.0 * M_PI / 180.0;
static const double CELL = 0.025;
static const int LAYOUT = 9;  // 512 x 512 cells, 12.8 m
static const double LO = 0.5, HI = 12.3;  // the room

struct ScanData
{
  std::vector<double> ranges, rays;
  std::vector<uint8_t> mask;
  tsd_scan_t s;
};

static void make_scan(double x, double y, double th, double max_range, ScanData* d)
{
  d->ranges.assign(N, 0.0);
  d->mask.assign(N, 1);
  d->rays.assign(2 * N, 0.0);
  for(int i = 0; i < N; i++)
  {
    const double a = th + PHI_MIN + i * RES, c = cos(a), s = sin(a);
    double t = 1e30;
    if(c > 1e-12) t = fmin(t, (HI - x) / c);
    if(c < -1e-12) t = fmin(t, (LO - x) / c);
    if(s > 1e-12) t = fmin(t, (HI - y) / s);
    if(s < -1e-12) t = fmin(t, (LO - y) / s);
    d->ranges[i] = (t > max_range) ? INFINITY : (double)(float)t;
    if(i % 37 == 5) d->mask[i] = 0;
    d->rays[i] = c * CELL;       // world-frame beam directions of length cellSize (Sensor::getNormalizedRayMap)
    d->rays[N + i] = s * CELL;
  }
  tsd_scan_t& sc = d->s;
  memset(&sc, 0, sizeof(sc));
  sc.n = N;
  sc.ranges = d->ranges.data();
  sc.mask = d->mask.data();
  const double P[9] = {cos(th), -sin(th), x, sin(th), cos(th), y, 0, 0, 1};
  memcpy(sc.pose, P, sizeof(P));
  tsd_invert3x3(sc.pose, sc.pose_inv);
  sc.phi_min = PHI_MIN;
  sc.angular_res = RES;
  sc.phi_lower = -0.5 * RES + PHI_MIN;
  sc.phi_upper = PHI_MIN + ((double)N - 0.5) * RES;
  sc.max_range = max_range;
  sc.min_range = 0.001;
  sc.low_reflectivity_range = 2.0;
}

#define CK(call)                                                                   \
  do                                                                               \
  {                                                                                \
    int rc__ = (call);                                                             \
    if(rc__ != 0)                                                                  \
    {                                                                              \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc__, tsd_last_error());            \
      exit(2);                                                                     \
    }                                                                              \
  } while(0)

struct Results
{
  std::vector<double> coords, normals, tsd;  // per iteration and localiser
  std::vector<uint8_t> mask;
  std::vector<uint32_t> hits;
  std::vector<int32_t> status;
  std::vector<uint64_t> map;  // checksum per partition at the end
};

static tsd_grid_t* build_grid()
{
  tsd_grid_t* g = nullptr;
  CK(tsdg_create(CELL, 5, LAYOUT, 0, &g));
  CK(tsdg_set_max_truncation(g, 3 * CELL));
  ScanData d;
  for(int k = 0; k < 6; k++)  // something to look at on the left, something on the right
  {
    make_scan(2.0 + 0.05 * k, 3.0 + 0.4 * k, 0.3 * k, 4.0, &d);
    CK(tsdg_push(g, &d.s));
    make_scan(10.5, 3.0 + 1.2 * k, 3.0 + 0.2 * k, 1.7, &d);
    CK(tsdg_push(g, &d.s));
  }
  return g;
}

static void pusher(tsd_grid_t* g, int iters)
{
  ScanData d;
  for(int i = 0; i < iters; i++)
  {
    make_scan(10.4 + 0.001 * (i % 100), 2.5 + 0.015 * i, 2.6 + 0.01 * i, 1.7, &d);
    if(i % 3 == 0) CK(tsdg_push_async(g, &d.s));  // the enqueue-only variant now and then: its staging must not be reused early
    else CK(tsdg_push(g, &d.s));
  }
  CK(tsdg_sync(g));
}

static void localiser(tsd_grid_t* g, int which, int iters, Results* r)
{
  ScanData d;
  std::vector<double> xy(2 * 64);
  for(int i = 0; i < iters; i++)
  {
    make_scan(2.0 + 0.3 * which + 0.002 * i, 2.5 + 0.01 * i + which, 0.2 * which + 0.005 * i, 4.0, &d);
    const size_t o = ((size_t)i * 2 + which);
    CK(tsdg_raycast_mask(g, &d.s, d.rays.data(), &r->coords[o * 2 * N], &r->normals[o * 2 * N], &r->mask[o * N], &r->hits[o]));
    for(int k = 0; k < 64; k++)
    {
      xy[2 * k] = 1.0 + 0.05 * k + 0.001 * i;
      xy[2 * k + 1] = 2.0 + 0.07 * k + which;
    }
    CK(tsdg_interpolate_bilinear(g, 64, xy.data(), &r->tsd[o * 64], &r->status[o * 64]));
  }
}

static void checksum(tsd_grid_t* g, Results* r)
{
  int32_t np = 0;
  CK(tsdg_num_partitions(g, &np));
  std::vector<int32_t> st(np);
  CK(tsdg_partition_states(g, st.data(), nullptr));
  r->map.assign(np, 0);
  std::vector<double> t(33 * 33), w(33 * 33);
  for(int p = 0; p < np; p++)
  {
    if(st[p] != TSD_PARTITION_CONTENT) { r->map[p] = (uint64_t)st[p]; continue; }
    CK(tsdg_download_partition(g, p, t.data(), w.data()));
    uint64_t a = 1469598103934665603ull;
    for(int i = 0; i < 33 * 33; i++)
    {
      uint64_t u, v;
      memcpy(&u, &t[i], 8);
      memcpy(&v, &w[i], 8);
      if(t[i] != t[i]) u = 0x7ff8000000000000ull;
      a = (a ^ u) * 1099511628211ull;
      a = (a ^ v) * 1099511628211ull;
    }
    r->map[p] = a;
  }
}

static void alloc(Results* r, int iters)
{
  const size_t m = (size_t)iters * 2;
  r->coords.assign(m * 2 * N, 7.25);
  r->normals.assign(m * 2 * N, 7.25);
  r->mask.assign(m * N, 9);
  r->hits.assign(m, 0);
  r->tsd.assign(m * 64, 0.0);
  r->status.assign(m * 64, -1);
}

int main(int argc, char** argv)
{
  const int iters = argc > 1 ? atoi(argv[1]) : 500;
  if(tsd_device_count() == 0) { fprintf(stderr, "no CUDA device: libtsdslam_b200 has no CPU path\n"); return 3; }
  Results serial, threaded;
  alloc(&serial, iters);
  alloc(&threaded, iters);
  {
    tsd_grid_t* g = build_grid();
    pusher(g, iters);
    localiser(g, 0, iters, &serial);
    localiser(g, 1, iters, &serial);
    checksum(g, &serial);
    tsdg_destroy(g);
  }
  {
    tsd_grid_t* g = build_grid();
    std::thread a(pusher, g, iters), b(localiser, g, 0, iters, &threaded), c(localiser, g, 1, iters, &threaded);
    a.join();
    b.join();
    c.join();
    checksum(g, &threaded);
    tsdg_destroy(g);
  }
  auto same = [](const std::vector<double>& x, const std::vector<double>& y) { return memcmp(x.data(), y.data(), x.size() * 8) == 0; };
  uint64_t hits = 0;
  for(uint32_t h : serial.hits) hits += h;
  const bool ok = same(serial.coords, threaded.coords) && same(serial.normals, threaded.normals) && serial.mask == threaded.mask &&
                  serial.hits == threaded.hits && same(serial.tsd, threaded.tsd) && serial.status == threaded.status &&
                  serial.map == threaded.map && hits > (uint64_t)iters * 50;
  printf("%d iterations, 1 pusher + 2 localisers on one handle: %llu ray hits, %s\n", iters, (unsigned long long)hits,
         ok ? "identical to the single-threaded run" : "DIFFERENT from the single-threaded run");
  return ok ? 0 : 1;
}
