// A TsdGrid sharded in bands of partition rows INSIDE the library: one handle, several devices (or several bands on
// one device), driven by one host thread -- what obvious::TsdGrid needs to be sharded at all, since SlamNode constructs
// one grid object (reference src/SlamNode.cpp:77) and every thread calls into it.  (bench.py at N > 1 runs one process
// per GPU instead and uses the band entry points directly; the kernels are the same.)
//
//   push      every band a scan can reach (tsdg_scan_box) integrates it into its own rows; no communication
//   read      before anything reads across a band boundary, boundary rows go to the neighbours' halo rows by one kernel
//             per band over peer memory (tsdg_band_halo_sync), allocation flags of the rows pushed to since the last
//             read are copied band to band, and the ray cast is the collective of raycast.cu: every band's marching
//             kernel stores its per-beam first events into every band's exchange block, a merge kernel keeps the earliest
// Results equal the unsharded grid's bit for bit (tests/test_sharded_gpu.py).
#include <algorithm>
#include <vector>

#include "common.cuh"

using namespace tsd;

struct tsd_sharded
{
  std::recursive_mutex* mtx;
  std::vector<tsd_grid_t*> bands;
  std::vector<int> row_begin, row_end;
  int parts_x, parts_y;
  // what changed since the last synchronisation
  bool dirty;
  int col_lo, col_hi;   // partition columns pushed to (halo rows travel for these only); col_lo > col_hi: all
  int row_lo, row_hi;   // partition rows whose allocation flags may have changed; row_lo > row_hi with dirty_all: all
  bool dirty_all;
  std::vector<char> pushed;  // bands the last push reached
};

static bool band_reached(const int box[4], int b, int e) { return box[1] <= e && box[3] >= b - 1; }

static void note_box(tsd_sharded* s, const int box[4])
{
  if(!s->dirty) { s->col_lo = box[0]; s->col_hi = box[2]; s->row_lo = box[1]; s->row_hi = box[3]; }
  else
  {
    s->col_lo = std::min(s->col_lo, box[0]); s->col_hi = std::max(s->col_hi, box[2]);
    s->row_lo = std::min(s->row_lo, box[1]); s->row_hi = std::max(s->row_hi, box[3]);
  }
  s->dirty = true;
}

static int sync_all(tsd_sharded* s)
{
  for(tsd_grid_t* g : s->bands)
  {
    int rc = tsdg_sync(g);
    if(rc) return rc;
  }
  return TSD_OK;
}

// halos + flags, if anything was pushed since the last call
static int make_readable(tsd_sharded* s)
{
  if(!s->dirty) return TSD_OK;
  const int nb = (int)s->bands.size();
  const bool all = s->dirty_all;
  const int c0 = all ? 0 : std::max(s->col_lo, 0), c1 = all ? s->parts_x - 1 : std::min(s->col_hi, s->parts_x - 1);
  if(nb > 1)
  {
    // every band's kernel waits for its neighbours': all of them are enqueued before anything synchronises
    for(tsd_grid_t* g : s->bands)
    {
      int rc = tsdg_band_halo_sync(g, c0, c1, c0, c1);
      if(rc) return rc;
    }
    // allocation flags: the owner's rows to every other band (the partition-skipping loop of the ray caster walks rays
    // through other bands' rows, RayCastPolar2D.cpp:223-235); stream-ordered after the owner's pushes
    int rc = sync_all(s);
    if(rc) return rc;
    const int r0 = all ? 0 : std::max(s->row_lo - 1, 0), r1 = all ? s->parts_y - 1 : std::min(s->row_hi + 1, s->parts_y - 1);
    for(int o = 0; o < nb; o++)
    {
      const int a = std::max(r0, s->row_begin[o]), b = std::min(r1, s->row_end[o] - 1);
      if(a > b) continue;
      const size_t off = (size_t)a * s->parts_x, bytes = (size_t)(b - a + 1) * s->parts_x;
      for(int d = 0; d < nb; d++)
      {
        if(d == o) continue;
        TSD_CUDA(cudaMemcpyPeerAsync(s->bands[d]->d_flags + off, s->bands[d]->device, s->bands[o]->d_flags + off, s->bands[o]->device,
                                     bytes, s->bands[o]->stream));
      }
    }
    rc = sync_all(s);
    if(rc) return rc;
  }
  s->dirty = false;
  s->dirty_all = false;
  return TSD_OK;
}

extern "C" {

int tsdg_create_sharded(double cell_size, int layout_partition, int layout_grid, int n_bands, const int* devices, tsd_sharded_t** out)
{
  if(!out) return TSD_E_INVALID;
  *out = nullptr;
  if(layout_grid < 5 || layout_grid > 16) { set_error("invalid grid layout"); return TSD_E_INVALID; }
  const int parts = (1 << layout_grid) / TSD_TILE;
  if(n_bands < 1 || n_bands > TSD_RCX_MAX || n_bands > parts) { set_error("1 .. %d bands", TSD_RCX_MAX); return TSD_E_INVALID; }
  const int ndev = tsd_device_count();
  if(ndev == 0) { set_error("no CUDA device: libtsdslam_b200 has no CPU path"); return TSD_E_NO_DEVICE; }
  tsd_sharded* s = new tsd_sharded();
  s->mtx = new std::recursive_mutex();
  s->parts_x = s->parts_y = parts;
  s->dirty = false;
  s->dirty_all = false;
  s->col_lo = s->row_lo = 0;
  s->col_hi = s->row_hi = -1;
  const int base = parts / n_bands, rem = parts % n_bands;
  int b = 0, rc = TSD_OK;
  for(int i = 0; i < n_bands && rc == TSD_OK; i++)
  {
    const int e = b + base + (i < rem ? 1 : 0);
    tsd_grid_t* g = nullptr;
    const int dev = devices ? devices[i] : (i % ndev);
    // (one band is simply the whole grid; tsdg_create_band then makes an unsharded handle)
    rc = tsdg_create_band(cell_size, layout_partition, layout_grid, dev, b, e, &g);
    if(rc == TSD_OK)
    {
      s->bands.push_back(g);
      s->row_begin.push_back(b);
      s->row_end.push_back(e);
    }
    b = e;
  }
  for(int i = 0; i < (int)s->bands.size() && rc == TSD_OK && n_bands > 1; i++)
  {
    if(i > 0) rc = tsdg_band_connect_local(s->bands[i], 0, s->bands[i - 1]);
    if(rc == TSD_OK && i + 1 < n_bands) rc = tsdg_band_connect_local(s->bands[i], 1, s->bands[i + 1]);
    if(rc == TSD_OK) rc = tsdg_band_rcx_connect_local(s->bands[i], i, n_bands, s->bands.data());
  }
  if(rc != TSD_OK)
  {
    for(tsd_grid_t* g : s->bands) tsdg_destroy(g);
    delete s->mtx;
    delete s;
    return rc;
  }
  s->pushed.assign(n_bands, 0);
  *out = s;
  return TSD_OK;
}

int tsdg_sharded_destroy(tsd_sharded_t* s)
{
  if(!s) return TSD_OK;
  for(tsd_grid_t* g : s->bands) tsdg_destroy(g);
  delete s->mtx;
  delete s;
  return TSD_OK;
}

int tsdg_sharded_num_bands(const tsd_sharded_t* s) { return s ? (int)s->bands.size() : 0; }

tsd_grid_t* tsdg_sharded_band(tsd_sharded_t* s, int i) { return (s && i >= 0 && i < (int)s->bands.size()) ? s->bands[i] : nullptr; }

int tsdg_sharded_set_max_truncation(tsd_sharded_t* s, double val)
{
  TSD_LOCK(s);
  if(!s) return TSD_E_INVALID;
  for(tsd_grid_t* g : s->bands) tsdg_set_max_truncation(g, val);
  return TSD_OK;
}

int tsdg_sharded_free_footprint(tsd_sharded_t* s, double cx, double cy, double w, double h)
{
  TSD_LOCK(s);
  if(!s) return TSD_E_INVALID;
  int rc = TSD_OK;
  for(tsd_grid_t* g : s->bands)
  {
    const int r = tsdg_free_footprint(g, cx, cy, w, h);
    if(r) rc = r;
  }
  s->dirty = true;
  s->dirty_all = true;
  return rc;
}

int tsdg_sharded_push_batch(tsd_sharded_t* s, const tsd_scan_t* scans, int32_t n)
{
  TSD_LOCK(s);
  if(!s || !scans || n < 1) return TSD_E_INVALID;
  std::vector<int> boxes(4 * (size_t)n);
  for(int i = 0; i < n; i++)
  {
    int32_t bx[4];
    int rc = tsdg_scan_box(s->bands[0], &scans[i], bx);
    if(rc) return rc;
    for(int k = 0; k < 4; k++) boxes[4 * i + k] = bx[k];
  }
  for(size_t b = 0; b < s->bands.size(); b++)
  {
    bool mine = false;
    for(int i = 0; i < n; i++) mine = mine || band_reached(&boxes[4 * i], s->row_begin[b], s->row_end[b]);
    s->pushed[b] = mine ? 1 : 0;
    if(mine)
    {
      int rc = tsdg_push_batch_async(s->bands[b], scans, n);  // enqueue only: the bands integrate concurrently
      if(rc) return rc;
    }
  }
  for(int i = 0; i < n; i++) note_box(s, &boxes[4 * i]);
  return sync_all(s);
}

int tsdg_sharded_push(tsd_sharded_t* s, const tsd_scan_t* scan) { return tsdg_sharded_push_batch(s, scan, 1); }

int tsdg_sharded_sync(tsd_sharded_t* s)
{
  TSD_LOCK(s);
  if(!s) return TSD_E_INVALID;
  return make_readable(s);
}

int tsdg_sharded_last_push_stats(tsd_sharded_t* s, tsd_push_stats_t* out)
{
  TSD_LOCK(s);
  if(!s || !out) return TSD_E_INVALID;
  memset(out, 0, sizeof(*out));
  for(size_t b = 0; b < s->bands.size(); b++)
  {
    if(!s->pushed[b]) continue;
    tsd_push_stats_t st;
    int rc = tsdg_last_push_stats(s->bands[b], &st);
    if(rc) return rc;
    out->cell_updates += st.cell_updates;
    out->cell_visits += st.cell_visits;
    out->active_tiles += st.active_tiles;
    out->emptied_tiles += st.emptied_tiles;
    out->newly_initialized += st.newly_initialized;
    out->fallback_cells += st.fallback_cells;
  }
  return TSD_OK;
}

int tsdg_sharded_raycast_mask(tsd_sharded_t* s, const tsd_scan_t* scan, const double* rays_world, double* coords, double* normals,
                              uint8_t* mask, uint32_t* count)
{
  TSD_LOCK(s);
  if(!s || !scan || !rays_world || !coords || !normals || !mask) return TSD_E_INVALID;
  int rc = make_readable(s);
  if(rc) return rc;
  if(s->bands.size() == 1) return tsdg_raycast_mask(s->bands[0], scan, rays_world, coords, normals, mask, count);
  for(tsd_grid_t* g : s->bands)
  {
    rc = tsdg_raycast_sharded_launch(g, scan, rays_world);
    if(rc) return rc;
  }
  // every band ends up with the full result; band 0's is handed out, the others are drained
  std::vector<double> c2(2 * (size_t)scan->n), n2(2 * (size_t)scan->n);
  std::vector<uint8_t> m2(scan->n);
  for(size_t b = s->bands.size(); b-- > 1;)
  {
    uint32_t cnt = 0;
    rc = tsdg_raycast_sharded_collect(s->bands[b], scan->n, c2.data(), n2.data(), m2.data(), &cnt);
    if(rc) return rc;
  }
  return tsdg_raycast_sharded_collect(s->bands[0], scan->n, coords, normals, mask, count);
}

int tsdg_sharded_interpolate_bilinear(tsd_sharded_t* s, int32_t n, const double* xy, double* tsd, int32_t* status)
{
  TSD_LOCK(s);
  if(!s || n < 0 || (n > 0 && (!xy || !tsd || !status))) return TSD_E_INVALID;
  if(n == 0) return TSD_OK;
  int rc = make_readable(s);
  if(rc) return rc;
  // every band samples every point; the band that owns a point's partition answers (the others report TSD_NOT_OWNED = 4)
  std::vector<double> t(n);
  std::vector<int32_t> st(n);
  std::vector<char> have(n, 0);
  for(size_t b = 0; b < s->bands.size(); b++)
  {
    rc = tsdg_interpolate_bilinear(s->bands[b], n, xy, t.data(), st.data());
    if(rc) return rc;
    for(int i = 0; i < n; i++)
      if(!have[i] && (st[i] != TSD_NOT_OWNED || b + 1 == s->bands.size()))
      {
        // a point whose interpolation cell straddles a band boundary is answered by the band that owns its lower-left
        // cell (it reads the rest from its halo row); INVALIDINDEX / EMPTYPARTITION do not depend on the band
        tsd[i] = t[i];
        status[i] = st[i];
        have[i] = 1;
      }
  }
  return TSD_OK;
}

int tsdg_sharded_partition_states(tsd_sharded_t* s, int32_t* state, double* init_weight)
{
  TSD_LOCK(s);
  if(!s || !state) return TSD_E_INVALID;
  const size_t np = (size_t)s->parts_x * s->parts_y;
  std::vector<int32_t> st(np);
  std::vector<double> iw(np);
  for(size_t b = 0; b < s->bands.size(); b++)
  {
    int rc = tsdg_partition_states(s->bands[b], st.data(), iw.data());
    if(rc) return rc;
    for(size_t p = (size_t)s->row_begin[b] * s->parts_x; p < (size_t)s->row_end[b] * s->parts_x; p++)
    {
      state[p] = st[p];
      if(init_weight) init_weight[p] = iw[p];
    }
  }
  return TSD_OK;
}

int tsdg_sharded_download_partition(tsd_sharded_t* s, int32_t p, double* tsd, double* weight)
{
  TSD_LOCK(s);
  if(!s || p < 0 || p >= s->parts_x * s->parts_y) return TSD_E_INVALID;
  int rc = make_readable(s);  // the border strips of a band's top row come from the band above
  if(rc) return rc;
  const int py = p / s->parts_x;
  for(size_t b = 0; b < s->bands.size(); b++)
    if(py >= s->row_begin[b] && py < s->row_end[b]) return tsdg_download_partition(s->bands[b], p, tsd, weight);
  return TSD_E_INVALID;
}

}  // extern "C"
