// The reference's checkpoint format for the device grid: TsdGrid::storeGrid
// (src/obvision/reconstruct/grid/TsdGrid.cpp:548-607) and the file constructor (:25-110).  Text, one value per
// line in the stream's default formatting (6 significant digits), partitions row-major: 0 = never seen,
// 1 + init weight = seen empty, 2 + 1024 x (tsd, weight) = allocated.  Borders are not stored: a loaded partition
// has the borders TsdGridPartition::init gives it, and the next push refreshes them (propagateBorders).
// Host code around bulk D2H / H2D copies of whole partition rows; unsharded grids only.
#include <locale.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"

using namespace tsd;

namespace
{

// The format is the "C" locale's (decimal point); a host application may have set another LC_NUMERIC: pin the
// calling thread to "C" for the duration of a store / load.
struct CLocale
{
  locale_t c, old;
  CLocale() : c(newlocale(LC_ALL_MASK, "C", (locale_t)0)), old((locale_t)0) { if(c) old = uselocale(c); }
  ~CLocale() { if(c) { uselocale(old); freelocale(c); } }
};

// tools.cpp:190-215.  The reference reads a file that ends early as zeros / NaN without a word; here the end of the
// file is remembered so that tsdg_load can refuse a truncated checkpoint.
thread_local bool t_hit_eof = false;

double get_double_line(FILE* f)
{
  char line[1024];
  if(!fgets(line, sizeof(line), f)) { t_hit_eof = true; return NAN; }
  if(line[0] == '\n' || line[0] == 0) return NAN;
  return strtod(line, NULL);
}

int get_int_line(FILE* f)
{
  char line[1024];
  if(!fgets(line, sizeof(line), f)) { t_hit_eof = true; return 0; }
  if(line[0] == '\n' || line[0] == 0) return 0;
  return atoi(line);
}

}  // namespace

extern "C" {

int tsdg_store(tsd_grid_t* g, const char* path)
{
  TSD_LOCK(g);
  if(!g || !path || !path[0]) { set_error("tsdg_store: invalid path"); return TSD_E_INVALID; }
  if(g->band) { set_error("tsdg_store works on an unsharded grid"); return TSD_E_INVALID; }
  TSD_CUDA(cudaSetDevice(g->device));
  TSD_CUDA(cudaStreamSynchronize(g->stream));
  CLocale pinned;
  FILE* f = fopen(path, "w");
  if(!f) { set_error("tsdg_store: cannot open %s", path); return TSD_E_INVALID; }
  std::vector<uint8_t> flags(g->n_parts);
  std::vector<double> initw(g->n_parts);
  TSD_CUDA(cudaMemcpy(flags.data(), g->d_flags, g->n_parts, cudaMemcpyDeviceToHost));
  TSD_CUDA(cudaMemcpy(initw.data(), g->d_initw, sizeof(double) * g->n_parts, cudaMemcpyDeviceToHost));
  fprintf(f, "%g\n%d\n%d\n%g\n", g->cell_size, 5, g->layout_grid, g->max_truncation);
  const size_t rowDoubles = (size_t)g->parts_x * TSD_TILE_STRIDE;
  std::vector<double> t(rowDoubles), w(rowDoubles);
  std::string buf;
  char num[64];
  for(int y = 0; y < g->parts_y; y++)
  {
    bool any = false;
    for(int x = 0; x < g->parts_x; x++) any = any || flags[(size_t)y * g->parts_x + x];
    if(any)
    {
      TSD_CUDA(cudaMemcpy(t.data(), g->d_tsd + (size_t)y * rowDoubles, sizeof(double) * rowDoubles, cudaMemcpyDeviceToHost));
      TSD_CUDA(cudaMemcpy(w.data(), g->d_weight + (size_t)y * rowDoubles, sizeof(double) * rowDoubles, cudaMemcpyDeviceToHost));
    }
    buf.clear();
    for(int x = 0; x < g->parts_x; x++)
    {
      const size_t p = (size_t)y * g->parts_x + x;
      if(flags[p])
      {
        buf += "2\n";
        const double* pt = t.data() + (size_t)x * TSD_TILE_STRIDE;
        const double* pw = w.data() + (size_t)x * TSD_TILE_STRIDE;
        for(int i = 0; i < TSD_TILE_CELLS; i++)
        {
          const int n = snprintf(num, sizeof(num), "%g\n%g\n", pt[i], pw[i]);
          buf.append(num, (size_t)n);
        }
      }
      else if(initw[p] > 0.0)
      {
        const int n = snprintf(num, sizeof(num), "1\n%g\n", initw[p]);
        buf.append(num, (size_t)n);
      }
      else buf += "0\n";
    }
    if(fwrite(buf.data(), 1, buf.size(), f) != buf.size()) { fclose(f); set_error("tsdg_store: write failed"); return TSD_E_INVALID; }
  }
  fclose(f);
  return TSD_OK;
}

int tsdg_load(const char* path, int device, tsd_grid_t** out)
{
  if(!path || !out) return TSD_E_INVALID;
  *out = nullptr;
  CLocale pinned;
  FILE* f = fopen(path, "r");
  if(!f) { set_error("tsdg_load: cannot open %s", path); return TSD_E_INVALID; }
  t_hit_eof = false;
  const double cellSize = get_double_line(f);
  const int layoutPartition = get_int_line(f);
  const int layoutGrid = get_int_line(f);
  const double maxTruncation = get_double_line(f);
  tsd_grid_t* g = nullptr;
  int rc = tsdg_create(cellSize, layoutPartition, layoutGrid, device, &g);
  if(rc) { fclose(f); return rc; }
  rc = tsdg_set_max_truncation(g, maxTruncation);
  if(rc) { fclose(f); tsdg_destroy(g); return rc; }
  std::vector<uint8_t> flags(g->n_parts, 0);
  std::vector<double> initw(g->n_parts, 0.0);
  const size_t rowDoubles = (size_t)g->parts_x * TSD_TILE_STRIDE;
  std::vector<double> t(rowDoubles), w(rowDoubles);
  const double nan = __builtin_nan("");
  bool anyContent = false;
  for(int y = 0; y < g->parts_y; y++)
  {
    bool any = false;
    for(int x = 0; x < g->parts_x; x++)
    {
      const size_t p = (size_t)y * g->parts_x + x;
      const int id = get_int_line(f);
      if(id == 0) continue;
      if(id == 1)
      {
        const double v = get_double_line(f);
        initw[p] = (TSD_MAXWEIGHT < v) ? TSD_MAXWEIGHT : v;  // std::min(v, TSDGRIDMAXWEIGHT)
      }
      else if(id == 2)
      {
        if(!any)
        {
          // (the cells of partitions that are not content are never read: zeros will do)
          memset(t.data(), 0, sizeof(double) * rowDoubles);
          memset(w.data(), 0, sizeof(double) * rowDoubles);
          any = true;
        }
        double* pt = t.data() + (size_t)x * TSD_TILE_STRIDE;
        double* pw = w.data() + (size_t)x * TSD_TILE_STRIDE;
        for(int i = 0; i < TSD_TILE_CELLS; i++)
        {
          pt[i] = get_double_line(f);
          pw[i] = get_double_line(f);
        }
        // the border cells keep what TsdGridPartition::init wrote (init weight 0: tsd NaN, weight 0)
        for(int i = TSD_TILE_CELLS; i < TSD_TILE_STRIDE; i++) { pt[i] = nan; pw[i] = 0.0; }
        flags[p] = 1;
      }
      else
      {
        fclose(f);
        tsdg_destroy(g);
        set_error("tsdg_load: unknown partition identifier %d for partition (%d/%d)", id, x, y);
        return TSD_E_INVALID;
      }
    }
    if(t_hit_eof)
    {
      fclose(f);
      set_error("tsdg_load: %s ends in partition row %d of %d: truncated checkpoint", path, y, g->parts_y);
      tsdg_destroy(g);
      return TSD_E_INVALID;
    }
    if(any)
    {
      anyContent = true;
      cudaError_t ce = cudaMemcpy(g->d_tsd + (size_t)y * rowDoubles, t.data(), sizeof(double) * rowDoubles, cudaMemcpyHostToDevice);
      if(ce == cudaSuccess) ce = cudaMemcpy(g->d_weight + (size_t)y * rowDoubles, w.data(), sizeof(double) * rowDoubles, cudaMemcpyHostToDevice);
      if(ce != cudaSuccess)
      {
        fclose(f);
        set_error("tsdg_load: %s", cudaGetErrorString(ce));
        tsdg_destroy(g);
        return TSD_E_CUDA;
      }
    }
  }
  fclose(f);
  cudaMemcpy(g->d_flags, flags.data(), g->n_parts, cudaMemcpyHostToDevice);
  cudaMemcpy(g->d_initw, initw.data(), sizeof(double) * g->n_parts, cudaMemcpyHostToDevice);
  if(anyContent)
  {
    const uint32_t all = 1;  // the next push refreshes every border, like the reference's full propagateBorders
    cudaMemcpy(g->d_counters + 3, &all, sizeof(all), cudaMemcpyHostToDevice);
    g->refresh_all_pending = true;
  }
  cudaError_t e = cudaGetLastError();
  if(e != cudaSuccess) { set_error("tsdg_load: %s", cudaGetErrorString(e)); tsdg_destroy(g); return TSD_E_CUDA; }
  *out = g;
  return TSD_OK;
}

}  // extern "C"
