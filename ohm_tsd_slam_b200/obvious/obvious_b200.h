// Header-compatible host-side mirror of the reference's `namespace obvious` interface for the mapping /
// localisation hot path, over the C ABI of libtsdslam_b200 (include/tsdslam_b200.h).
//
// A caller written against the reference's classes -- ThreadMapping / ThreadLocalize
// (reference src/ThreadMapping.cpp:32-76, src/ThreadLocalize.cpp:177-225, :310-409, :513-591) -- compiles
// against these headers unchanged (same class names, method names, argument meaning and error behaviour);
// the forwarding headers next to this file reproduce the reference's include paths.  What stays on the host
// here stays on the host in the reference design too: the sensor's pose / ray
// map / masks and the matchers' pre-processing.  Everything that touches cells, beams, pairs or hypotheses
// goes to the device; there is no CPU fallback -- without a CUDA device the constructors throw.
//
// Each class cites the reference interface it replaces.
#pragma once
#define OBVIOUS_B200_H 1

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/tsdslam_b200.h"

namespace obvious
{

typedef double obfloat;  // obcore/base/types.h:28-31

#define OBVIOUS_B200_CHECK(call)                                                                       \
  do                                                                                                    \
  {                                                                                                     \
    int rc__ = (call);                                                                                  \
    if(rc__ != 0) throw std::runtime_error(std::string(#call) + " failed: " + tsd_last_error());        \
  } while(0)

// obcore/math/mathbase.h:175-190
static inline double deg2rad(const double deg) { return ((M_PI * deg) / 180.0); }
static inline double rad2deg(const double rad) { return ((rad * 180.0) / M_PI); }

// ---------------------------------------------------------------------------------------------------------
// obvious::Matrix (obcore/math/linalg/gsl/Matrix.h) -- the subset the node uses.  Dense, row-major.
// operator* follows gslcblas' dgemm NoTrans x NoTrans rounding order (k outer, zero coefficients skipped) so
// that poses accumulate with the same bits as in the reference (Sensor.cpp:50-60).
// ---------------------------------------------------------------------------------------------------------
class Matrix
{
public:
  Matrix(unsigned int rows, unsigned int cols, double* data = NULL) : _rows(rows), _cols(cols), _d((size_t)rows * cols)
  {
    if(data) setData(data);
  }
  Matrix(const Matrix& M) : _rows(M._rows), _cols(M._cols), _d(M._d) {}
  // copy constructor of a submatrix (gsl/Matrix.cpp:27-32)
  Matrix(Matrix M, unsigned int i, unsigned int j, unsigned int rows, unsigned int cols) : _rows(rows), _cols(cols), _d((size_t)rows * cols)
  {
    for(unsigned int r = 0; r < rows; r++)
      for(unsigned int c = 0; c < cols; c++) (*this)(r, c) = M(i + r, j + c);
  }
  Matrix& operator=(const Matrix& M)
  {
    _rows = M._rows; _cols = M._cols; _d = M._d;
    return *this;
  }
  double& operator()(unsigned int row, unsigned int col) { return _d[(size_t)row * _cols + col]; }
  double operator()(unsigned int row, unsigned int col) const { return _d[(size_t)row * _cols + col]; }
  unsigned int getRows() const { return _rows; }
  unsigned int getCols() const { return _cols; }
  void getData(double* array) const { memcpy(array, _d.data(), sizeof(double) * _d.size()); }
  void setData(double* array) { memcpy(_d.data(), array, sizeof(double) * _d.size()); }
  const double* data() const { return _d.data(); }
  void setIdentity()
  {
    for(unsigned int r = 0; r < _rows; r++)
      for(unsigned int c = 0; c < _cols; c++) (*this)(r, c) = (r == c) ? 1.0 : 0.0;
  }
  void setZero() { std::fill(_d.begin(), _d.end(), 0.0); }
  friend Matrix operator*(const Matrix& A, const Matrix& B)
  {
    Matrix C(A._rows, B._cols);
    C.setZero();
    for(unsigned int k = 0; k < A._cols; k++)
      for(unsigned int i = 0; i < A._rows; i++)
      {
        const double temp = 1.0 * A(i, k);
        if(temp != 0.0)
          for(unsigned int j = 0; j < B._cols; j++) C(i, j) += temp * B(k, j);
      }
    return C;
  }
  // gsl/Matrix.cpp:168-179 (LU with partial pivoting); 3x3 through the library's routine
  void invert()
  {
    if(_rows != _cols) throw std::runtime_error("Matrix::invert: not square");
    if(_rows == 3)
    {
      double out[9];
      OBVIOUS_B200_CHECK(tsd_invert3x3(_d.data(), out));
      setData(out);
      return;
    }
    const unsigned int n = _rows;
    std::vector<double> A(_d), inv((size_t)n * n);
    std::vector<unsigned int> perm(n);
    for(unsigned int i = 0; i < n; i++) perm[i] = i;
    for(unsigned int j = 0; j < n; j++)
    {
      unsigned int ip = j;
      double mx = fabs(A[j * n + j]);
      for(unsigned int i = j + 1; i < n; i++)
        if(fabs(A[i * n + j]) > mx) { mx = fabs(A[i * n + j]); ip = i; }
      if(ip != j)
      {
        for(unsigned int k = 0; k < n; k++) std::swap(A[j * n + k], A[ip * n + k]);
        std::swap(perm[j], perm[ip]);
      }
      const double invp = 1.0 / A[j * n + j];
      for(unsigned int i = j + 1; i < n; i++) A[i * n + j] *= invp;
      for(unsigned int i = j + 1; i < n; i++)
      {
        const double t = -1.0 * A[i * n + j];
        for(unsigned int k = j + 1; k < n; k++) A[i * n + k] += A[j * n + k] * t;
      }
    }
    for(unsigned int c = 0; c < n; c++)
    {
      std::vector<double> x(n);
      for(unsigned int i = 0; i < n; i++) x[i] = (perm[i] == c) ? 1.0 : 0.0;
      for(unsigned int i = 1; i < n; i++)
      {
        double t = x[i];
        for(unsigned int j = 0; j < i; j++) t -= A[i * n + j] * x[j];
        x[i] = t;
      }
      for(int i = (int)n - 1; i >= 0; i--)
      {
        double t = x[i];
        for(unsigned int j = i + 1; j < n; j++) t -= A[i * n + j] * x[j];
        x[i] = t / A[i * n + i];
      }
      for(unsigned int i = 0; i < n; i++) inv[i * n + c] = x[i];
    }
    _d = inv;
  }
  Matrix getInverse()
  {
    Matrix M = *this;
    M.invert();
    return M;
  }
  void print() const
  {
    for(unsigned int r = 0; r < _rows; r++)
    {
      for(unsigned int c = 0; c < _cols; c++) std::cout << (*this)(r, c) << " ";
      std::cout << std::endl;
    }
  }
  friend std::ostream& operator<<(std::ostream& os, Matrix& M)
  {
    for(unsigned int r = 0; r < M._rows; r++)
    {
      os << M(r, 0);
      for(unsigned int c = 1; c < M._cols; c++) os << " " << M(r, c);
      os << std::endl;
    }
    return os;
  }

private:
  unsigned int _rows, _cols;
  std::vector<double> _d;
};

// obcore/math/linalg/MatrixFactory.cpp:88-96
struct MatrixFactory
{
  static Matrix TransformationMatrix33(obfloat phi, obfloat tx, obfloat ty)
  {
    Matrix M(3, 3);
    M.setIdentity();
    M(0, 2) = tx;
    M(1, 2) = ty;
    const obfloat cphi = cos(phi), sphi = sin(phi);
    M(0, 0) = cphi; M(0, 1) = -sphi;
    M(1, 0) = sphi; M(1, 1) = cphi;
    return M;
  }
};

// ---------------------------------------------------------------------------------------------------------
// obvious::SensorPolar2D (reconstruct/grid/SensorPolar2D.h, reconstruct/Sensor.h).  Host class, as in the
// reference; the kernels read a POD snapshot of it (tsd_scan_t).
// ---------------------------------------------------------------------------------------------------------
class SensorPolar2D
{
public:
  // SensorPolar2D.cpp:11-48
  SensorPolar2D(unsigned int size, double angularRes, double phiMin, double maxRange = INFINITY, double minRange = 0.0,
                double lowReflectivityRange = INFINITY)
      : _size(size), _angularRes(angularRes), _phiMin(phiMin), _maxRange(maxRange), _minRange(minRange),
        _lowReflectivityRange(lowReflectivityRange), _rayNorm(1.0), _T(3, 3), _rays(2, size), _raysLocal(2, size)
  {
    _data.assign(size, 0.0);
    _mask = new bool[size];
    for(unsigned int i = 0; i < size; i++) _mask[i] = true;
    _phiLowerBound = -0.5 * _angularRes + _phiMin;
    _phiUpperBound = _phiMin + (((double)size) - 0.5) * _angularRes;
    for(unsigned int i = 0; i < size; i++)
    {
      const double phi = _phiMin + ((double)i) * _angularRes;
      _rays(0, i) = cos(phi);
      _rays(1, i) = sin(phi);
    }
    _raysLocal = _rays;
    _T.setIdentity();
  }
  ~SensorPolar2D() { delete[] _mask; }
  SensorPolar2D(const SensorPolar2D&) = delete;
  SensorPolar2D& operator=(const SensorPolar2D&) = delete;

  unsigned int getRealMeasurementSize() { return _size; }
  double getAngularResolution() const { return _angularRes; }
  double getPhiMin() const { return _phiMin; }
  double getPhiLowerBound() const { return _phiLowerBound; }
  double getPhiUpperBound() const { return _phiUpperBound; }
  double getMaximumRange() { return _maxRange; }
  double getMinimumRange() { return _minRange; }
  double getLowReflectivityRange() { return _lowReflectivityRange; }
  double* getRealMeasurementData() { return _data.data(); }
  bool* getRealMeasurementMask() { return _mask; }
  // Sensor.cpp:125-145
  void setRealMeasurementData(double* data, double scale = 1.0)
  {
    if(scale == 1.0) memcpy(_data.data(), data, _size * sizeof(double));
    else
      for(unsigned int i = 0; i < _size; i++) _data[i] = data[i] * scale;
  }
  void setRealMeasurementData(std::vector<float> data, float scale = 1.0)
  {
    for(unsigned int i = 0; i < data.size() && i < _size; i++) _data[i] = (double)(data[i] * scale);
  }
  void setRealMeasurementMask(bool* mask) { memcpy(_mask, mask, _size * sizeof(*mask)); }
  void resetMask()
  {
    for(unsigned int i = 0; i < _size; i++) _mask[i] = true;
  }
  // Sensor.cpp:252-272
  void maskZeroDepth()
  {
    for(unsigned int i = 0; i < _size; i++) _mask[i] = _mask[i] && (_data[i] != 0.0);
  }
  void maskInvalidDepth()
  {
    for(unsigned int i = 0; i < _size; i++)
    {
      if(_data[i] > _maxRange) _data[i] = INFINITY;
      if(std::isnan(_data[i]))
      {
        _mask[i] = false;
        _data[i] = INFINITY;
      }
    }
  }
  // SensorPolar2D.cpp:67-98
  void maskDepthDiscontinuity(double thresh)
  {
    const int radius = 1;
    double cosphi, sinphi;
    sincos(_angularRes, &sinphi, &cosphi);
    for(int i = radius; i < ((int)_size) - radius; i++)
    {
      double betamin = M_PI;
      const double a = _data[i];
      if(std::isinf(a)) continue;
      for(int j = -radius; j <= radius; j++)
      {
        const double b = _data[i + j];
        if(std::isinf(b)) continue;
        const double c = sqrt(a * a + b * b - 2 * a * b * cosphi);
        if(a > b)
        {
          const double beta = asin(b / c * sinphi);
          if(beta < betamin) betamin = beta;
        }
      }
      if(betamin < thresh) _mask[i] = false;
    }
  }
  // SensorPolar2D.cpp:59-65
  void setStandardMask()
  {
    resetMask();
    maskZeroDepth();
    maskInvalidDepth();
    maskDepthDiscontinuity(deg2rad(3.0));
  }
  // Sensor.cpp:36-60
  Matrix* getNormalizedRayMap(double norm)
  {
    if(norm != _rayNorm)
    {
      for(unsigned int i = 0; i < _size; i++)
        for(unsigned int j = 0; j < 2; j++) _rays(j, i) *= (norm / _rayNorm);
      _rayNorm = norm;
    }
    return &_rays;
  }
  void transform(Matrix* T)
  {
    Matrix R(*T, 0, 0, 2, 2);
    _rays = R * _rays;
    _T = _T * *T;
  }
  Matrix getTransformation() { return _T; }
  void setTransformation(Matrix T) { _T = T; }
  void resetTransformation() { _T.setIdentity(); }
  void getPosition(obfloat* tr)
  {
    tr[0] = _T(0, 2);
    tr[1] = _T(1, 2);
  }
  // Sensor.cpp:168-190
  unsigned int dataToCartesianVectorMask(double*& coords, bool*& validityMask)
  {
    unsigned int cnt = 0, validPoints = 0;
    for(unsigned int i = 0; i < _size; i++)
    {
      if(!std::isinf(_data[i]) && _mask[i])
      {
        for(unsigned int j = 0; j < 2; j++) coords[cnt++] = _raysLocal(j, i) * _data[i];
        validPoints++;
        validityMask[i] = true;
      }
      else
      {
        cnt += 2;
        validityMask[i] = false;
      }
    }
    return validPoints;
  }

  // POD snapshot for the C ABI; `mask8` must outlive the call it is passed to
  void snapshot(tsd_scan_t* s, std::vector<uint8_t>* mask8)
  {
    mask8->resize(_size);
    for(unsigned int i = 0; i < _size; i++) (*mask8)[i] = _mask[i] ? 1 : 0;
    memset(s, 0, sizeof(*s));
    s->n = (int32_t)_size;
    s->ranges = _data.data();
    s->mask = mask8->data();
    _T.getData(s->pose);
    OBVIOUS_B200_CHECK(tsd_invert3x3(s->pose, s->pose_inv));
    s->phi_min = _phiMin;
    s->angular_res = _angularRes;
    s->phi_lower = _phiLowerBound;
    s->phi_upper = _phiUpperBound;
    s->max_range = _maxRange;
    s->min_range = _minRange;
    s->low_reflectivity_range = _lowReflectivityRange;
  }

private:
  unsigned int _size;
  double _angularRes, _phiMin, _phiLowerBound, _phiUpperBound, _maxRange, _minRange, _lowReflectivityRange, _rayNorm;
  Matrix _T, _rays, _raysLocal;
  std::vector<double> _data;
  bool* _mask;
};

// ---------------------------------------------------------------------------------------------------------
// obvious::TsdGrid (reconstruct/grid/TsdGrid.h)
// ---------------------------------------------------------------------------------------------------------
enum EnumTsdGridLayout
{
  LAYOUT_1x1 = 0, LAYOUT_2x2 = 1, LAYOUT_4x4 = 2, LAYOUT_8x8 = 3, LAYOUT_16x16 = 4, LAYOUT_32x32 = 5, LAYOUT_64x64 = 6,
  LAYOUT_128x128 = 7, LAYOUT_256x256 = 8, LAYOUT_512x512 = 9, LAYOUT_1024x1024 = 10, LAYOUT_2048x2048 = 11,
  LAYOUT_4096x4096 = 12, LAYOUT_8192x8192 = 13, LAYOUT_16384x16384 = 14, LAYOUT_36768x36768 = 15,
  LAYOUT_65536x65536 = 16  // beyond the reference's enum (TsdGrid.h:11-26): needs the dense device grid
};

enum EnumTsdGridInterpolate
{
  INTERPOLATE_SUCCESS = 0, INTERPOLATE_INVALIDINDEX = 1, INTERPOLATE_EMPTYPARTITION = 2, INTERPOLATE_ISNAN = 3
};

enum EnumTsdGridPartitionIdentifier { UNINITIALIZED = 0, EMPTY = 1, CONTENT = 2 };  // TsdGrid.h:33-35
enum EnumTsdGridLoadSource { FILE_SOURCE = 0, STRING_SOURCE = 1 };                  // TsdGrid.h:37-39

class TsdGrid
{
public:
  // TsdGrid.cpp:20-23 (SlamNode.cpp:77).  `device` is the only addition: the CUDA ordinal, default 0.
  TsdGrid(const obfloat cellSize, const EnumTsdGridLayout layoutPartition, const EnumTsdGridLayout layoutGrid, int device = 0)
      : _h(NULL), _sh(NULL), _pushed(false)
  {
    OBVIOUS_B200_CHECK(tsdg_create(cellSize, (int)layoutPartition, (int)layoutGrid, device, &_h));
    refresh();
  }
  // Not in the reference: the same grid sharded over nBands bands of partition rows inside the library, band i on
  // devices[i] (NULL: device i mod #devices).  push / pushBatch / RayCastPolar2D / interpolateBilinear / freeFootprint
  // return what the unsharded grid returns, bit for bit; the publisher's calls (RayCastAxisAligned2D, grid2ColorImage),
  // interpolateNormal, storeGrid and TSD_PDFMatching want an unsharded grid and say so.
  TsdGrid(const obfloat cellSize, const EnumTsdGridLayout layoutPartition, const EnumTsdGridLayout layoutGrid, int nBands,
          const int* devices)
      : _h(NULL), _sh(NULL), _pushed(false)
  {
    OBVIOUS_B200_CHECK(tsdg_create_sharded(cellSize, (int)layoutPartition, (int)layoutGrid, nBands, devices, &_sh));
    _h = tsdg_sharded_band(_sh, 0);  // geometry getters
    refresh();
  }
  // TsdGrid.cpp:25-110: a grid from a file written by storeGrid (FILE_SOURCE only)
  TsdGrid(const std::string& data, const EnumTsdGridLoadSource source = FILE_SOURCE, int device = 0) : _h(NULL), _sh(NULL), _pushed(false)
  {
    if(source != FILE_SOURCE)
    {
      fprintf(stderr, "obvious_b200: TsdGrid(STRING_SOURCE) is not supported\n");
      std::exit(3);
    }
    if(tsdg_load(data.c_str(), device, &_h) != TSD_OK)
    {
      fprintf(stderr, "obvious_b200: %s\n", tsd_last_error());
      std::exit(2);  // the reference exits as well (TsdGrid.cpp:39-43)
    }
    _pushed = true;
    refresh();
  }
  // TsdGrid.cpp:548-607
  bool storeGrid(const std::string& path)
  {
    if(!path.size()) return false;
    unsharded("storeGrid");
    return tsdg_store(_h, path.c_str()) == TSD_OK;
  }
  virtual ~TsdGrid()
  {
    if(_sh) tsdg_sharded_destroy(_sh);
    else tsdg_destroy(_h);
  }
  TsdGrid(const TsdGrid&) = delete;
  TsdGrid& operator=(const TsdGrid&) = delete;

  unsigned int getCellsX() const { return (unsigned int)_cellsX; }
  unsigned int getCellsY() const { return (unsigned int)_cellsY; }
  obfloat getCellSize() const { return _cellSize; }
  obfloat getMinX() const { return _minX; }
  obfloat getMaxX() const { return _maxX; }
  obfloat getMinY() const { return _minY; }
  obfloat getMaxY() const { return _maxY; }
  unsigned int getPartitionSize() const { return (unsigned int)_partSize; }
  void getCentroid(double centroid[2])
  {
    centroid[0] = (_minX + _maxX) * 0.5;
    centroid[1] = (_minY + _maxY) * 0.5;
  }
  void setMaxTruncation(const double val)
  {
    if(_sh) OBVIOUS_B200_CHECK(tsdg_sharded_set_max_truncation(_sh, val));
    else OBVIOUS_B200_CHECK(tsdg_set_max_truncation(_h, val));
    refresh();
  }
  double getMaxTruncation() const { return _maxTruncation; }
  // TsdGrid.cpp:217-284
  void push(SensorPolar2D* sensor)
  {
    tsd_scan_t s;
    std::vector<uint8_t> m;
    sensor->snapshot(&s, &m);
    if(_sh) OBVIOUS_B200_CHECK(tsdg_sharded_push(_sh, &s));
    else OBVIOUS_B200_CHECK(tsdg_push(_h, &s));
    _pushed = true;
  }
  // Not in the reference: every queued sensor in one call, with the result of push() on each in turn.
  // ThreadMapping::eventLoop drains its queue exactly like that (ThreadMapping.cpp:43-62); two sensors of the same
  // model (the two lasers of a robot) then share one classification and one cell-update launch (tsdg_push_batch).
  void pushBatch(const std::vector<SensorPolar2D*>& sensors)
  {
    if(sensors.empty()) return;
    std::vector<tsd_scan_t> s(sensors.size());
    std::vector<std::vector<uint8_t> > m(sensors.size());
    for(size_t i = 0; i < sensors.size(); i++) sensors[i]->snapshot(&s[i], &m[i]);
    if(_sh) OBVIOUS_B200_CHECK(tsdg_sharded_push_batch(_sh, s.data(), (int32_t)s.size()));
    else OBVIOUS_B200_CHECK(tsdg_push_batch(_h, s.data(), (int32_t)s.size()));
    _pushed = true;
  }
  bool containsData() { return _pushed; }
  // TsdGrid.h:284-304
  EnumTsdGridInterpolate interpolateBilinear(obfloat coord[2], obfloat* tsd)
  {
    int32_t st = 0;
    double v = NAN;
    if(_sh) OBVIOUS_B200_CHECK(tsdg_sharded_interpolate_bilinear(_sh, 1, coord, &v, &st));
    else OBVIOUS_B200_CHECK(tsdg_interpolate_bilinear(_h, 1, coord, &v, &st));
    if(st == INTERPOLATE_SUCCESS || st == INTERPOLATE_ISNAN) *tsd = v;
    return (EnumTsdGridInterpolate)st;
  }
  // TsdGrid.cpp:517-546
  bool interpolateNormal(const obfloat coord[2], obfloat normal[2])
  {
    int32_t ok = 0;
    double n[2];
    unsharded("interpolateNormal");
    OBVIOUS_B200_CHECK(tsdg_interpolate_normal(_h, 1, coord, n, &ok));
    if(ok) { normal[0] = n[0]; normal[1] = n[1]; }
    return ok != 0;
  }
  // TsdGrid.h:342-347
  bool isInsideGrid(SensorPolar2D* sensor)
  {
    obfloat coord[2];
    sensor->getPosition(coord);
    return (coord[0] > _minX && coord[0] < _maxX && coord[1] > _minY && coord[1] < _maxY);
  }
  // TsdGrid.cpp:609-638
  bool freeFootprint(const obfloat centerCoords[2], const obfloat width, const obfloat height)
  {
    const int rc = _sh ? tsdg_sharded_free_footprint(_sh, centerCoords[0], centerCoords[1], width, height)
                       : tsdg_free_footprint(_h, centerCoords[0], centerCoords[1], width, height);
    if(rc == TSD_E_RANGE) return false;
    OBVIOUS_B200_CHECK(rc);
    return true;
  }
  // TsdGrid.cpp:429-488 (ThreadGrid.cpp:125)
  void grid2ColorImage(unsigned char* image, unsigned int width, unsigned int height)
  {
    unsharded("grid2ColorImage");
    OBVIOUS_B200_CHECK(tsdg_color_image(_h, image, width, height));
  }
  // the unsharded handle (the calls that take one do not work on a sharded grid)
  tsd_grid_t* handle() const
  {
    unsharded("this call");
    return _h;
  }
  tsd_sharded_t* shardedHandle() const { return _sh; }

private:
  void refresh()
  {
    OBVIOUS_B200_CHECK(tsdg_get_geometry(_h, &_cellsX, &_cellsY, &_partSize, &_cellSize, &_minX, &_maxX, &_minY, &_maxY,
                                         &_maxTruncation));
  }
  void unsharded(const char* what) const
  {
    if(!_sh) return;
    throw std::runtime_error(std::string("obvious_b200: ") + what + " needs an unsharded TsdGrid");
  }
  tsd_grid_t* _h;
  tsd_sharded_t* _sh;
  int32_t _cellsX, _cellsY, _partSize;
  double _cellSize, _minX, _maxX, _minY, _maxY, _maxTruncation;
  bool _pushed;
};

// ---------------------------------------------------------------------------------------------------------
// obvious::RayCastPolar2D (reconstruct/grid/RayCastPolar2D.h)
// ---------------------------------------------------------------------------------------------------------
class RayCastPolar2D
{
public:
  RayCastPolar2D() {}
  ~RayCastPolar2D() {}
  // RayCastPolar2D.cpp:113-192
  unsigned int calcCoordsFromCurrentViewMask(TsdGrid* grid, SensorPolar2D* sensor, double* coords, double* normals, bool* mask)
  {
    tsd_scan_t s;
    std::vector<uint8_t> m;
    sensor->snapshot(&s, &m);
    Matrix* R = sensor->getNormalizedRayMap(grid->getCellSize());
    std::vector<uint8_t> hit(s.n);
    uint32_t cnt = 0;
    if(grid->shardedHandle())
      OBVIOUS_B200_CHECK(tsdg_sharded_raycast_mask(grid->shardedHandle(), &s, R->data(), coords, normals, hit.data(), &cnt));
    else
      OBVIOUS_B200_CHECK(tsdg_raycast_mask(grid->handle(), &s, R->data(), coords, normals, hit.data(), &cnt));
    for(int i = 0; i < s.n; i++) mask[i] = hit[i] != 0;
    return cnt;
  }
  // RayCastPolar2D.cpp:27-111 (beam order, the reference's single-thread order)
  void calcCoordsFromCurrentView(TsdGrid* grid, SensorPolar2D* sensor, double* coords, double* normals, unsigned int* ctr)
  {
    tsd_scan_t s;
    std::vector<uint8_t> m;
    sensor->snapshot(&s, &m);
    Matrix* R = sensor->getNormalizedRayMap(grid->getCellSize());
    uint32_t cnt = 0;
    OBVIOUS_B200_CHECK(tsdg_raycast(grid->handle(), &s, R->data(), coords, normals, &cnt));
    *ctr = cnt;
  }
};

// ---------------------------------------------------------------------------------------------------------
// obvious::RayCastAxisAligned2D (reconstruct/grid/RayCastAxisAligned2D.h): the map publisher's ray caster
// (ThreadGrid.cpp:84).  coords must hold cellsX * cellsY doubles, as ThreadGrid allocates it (:21).
// ---------------------------------------------------------------------------------------------------------
class RayCastAxisAligned2D
{
public:
  RayCastAxisAligned2D() {}
  virtual ~RayCastAxisAligned2D() {}
  // RayCastAxisAligned2D.cpp:13-105
  void calcCoords(TsdGrid* grid, obfloat* coords, obfloat* normals, unsigned int* cnt, char* occupiedGrid = NULL)
  {
    uint32_t n = 0;
    const uint32_t cap = (uint32_t)(((size_t)grid->getCellsX() * grid->getCellsY()) / 2);
    OBVIOUS_B200_CHECK(tsdg_axis_aligned_map(grid->handle(), coords, cap, normals, &n, (int8_t*)occupiedGrid));
    *cnt = n;
  }
};

// ---------------------------------------------------------------------------------------------------------
// ICP: the strategy objects of the reference keep their names and constructors; they carry configuration,
// the work happens in one device kernel (icp_run).  Only the combination the node wires
// (ThreadLocalize.cpp:210-225) is supported: FlannPairAssignment + OutOfBoundsFilter2D + DistanceFilter +
// ReciprocalFilter + ClosedFormEstimator2D; anything else makes Icp::iterate return ICP_ERROR.
// ---------------------------------------------------------------------------------------------------------
enum EnumIcpState
{
  ICP_IDLE = 0, ICP_PROCESSING = 1, ICP_NOTMATCHABLE = 2, ICP_MAXITERATIONS = 3, ICP_TIMEELAPSED = 4, ICP_SUCCESS = 5,
  ICP_CONVERGED = 6, ICP_ERROR = 7
};

struct StrCartesianIndexPair  // registration/icp/assign/assignbase.h:38-44
{
  unsigned int indexFirst;
  unsigned int indexSecond;
};

class IPreAssignmentFilter  // assign/filter/IPreAssignmentFilter.h
{
public:
  virtual ~IPreAssignmentFilter() {}
};

class IPostAssignmentFilter  // assign/filter/IPostAssignmentFilter.h
{
public:
  IPostAssignmentFilter() { _active = true; }
  virtual ~IPostAssignmentFilter() {}
  virtual void activate() { _active = true; }
  virtual void deactivate() { _active = false; }
  virtual void reset() {}
  bool _active;
};

class OutOfBoundsFilter2D : public IPreAssignmentFilter  // assign/filter/OutOfBoundsFilter2D.h
{
public:
  OutOfBoundsFilter2D(double xMin, double xMax, double yMin, double yMax) : _T(3, 3)
  {
    _b[0] = xMin; _b[1] = xMax; _b[2] = yMin; _b[3] = yMax;
    _T.setIdentity();
  }
  void setPose(Matrix* T) { _T = *T; }
  double _b[4];
  Matrix _T;
};

class DistanceFilter : public IPostAssignmentFilter  // assign/filter/DistanceFilter.h
{
public:
  DistanceFilter(double maxdist, double mindist, unsigned int iterations) : _max(maxdist), _min(mindist), _it(iterations) {}
  double _max, _min;
  unsigned int _it;
};

class ReciprocalFilter : public IPostAssignmentFilter  // assign/filter/ReciprocalFilter.h
{
public:
  ReciprocalFilter() {}
};

class PairAssignment  // assign/PairAssignment.h
{
public:
  PairAssignment(int dimension = 2) : _dimension(dimension), _model(NULL), _sizeModel(0), _hp(NULL) {}
  virtual ~PairAssignment() { icp_destroy(_hp); }
  void addPreFilter(IPreAssignmentFilter* filter) { _vPrefilter.push_back(filter); }
  void addPostFilter(IPostAssignmentFilter* filter) { _vPostfilter.push_back(filter); }
  int getDimension() { return _dimension; }
  // PairAssignment.h:49-63.  The model pointer is borrowed, as in the reference (FlannPairAssignment.cpp:48-50).
  virtual void setModel(double** model, int size) { _model = model; _sizeModel = size; }
  // PairAssignment.cpp:38-84 used on its own (Icp::iterate does not come through here: it runs the whole loop in one
  // kernel): pre-filter, exact nearest neighbours, post-filters in one pass on the device (icp_pairs).  Supports the
  // filters the node wires (ThreadLocalize.cpp:212-221); any of them may be absent.
  virtual void determinePairs(double** scene, bool* mask, int size)
  {
    _pairs.clear();
    _distancesSqr.clear();
    _nonPairs.clear();
    if(!_model || _sizeModel < 1 || size < 1) return;
    OutOfBoundsFilter2D* fb = NULL;
    DistanceFilter* fd = NULL;
    bool reciprocal = false;
    for(size_t i = 0; i < _vPrefilter.size(); i++)
      if(!fb) fb = dynamic_cast<OutOfBoundsFilter2D*>(_vPrefilter[i]);
    for(size_t i = 0; i < _vPostfilter.size(); i++)
    {
      if(!fd) fd = dynamic_cast<DistanceFilter*>(_vPostfilter[i]);
      if(dynamic_cast<ReciprocalFilter*>(_vPostfilter[i])) reciprocal = true;
    }
    // masked scene points do not take part (PairAssignment.cpp:49-58): compact, remember the original indices
    std::vector<double> M(2 * (size_t)_sizeModel), S;
    std::vector<unsigned int> idx;
    for(int i = 0; i < _sizeModel; i++) { M[2 * i] = _model[i][0]; M[2 * i + 1] = _model[i][1]; }
    for(int i = 0; i < size; i++)
      if(!mask || mask[i]) { S.push_back(scene[i][0]); S.push_back(scene[i][1]); idx.push_back((unsigned int)i); }
    if(idx.empty()) return;
    const double inf = 1e150;
    double bounds[4] = {-inf, inf, -inf, inf}, pose[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if(fb) { memcpy(bounds, fb->_b, sizeof(bounds)); fb->_T.getData(pose); }
    const double dmax = fd ? fd->_max : 1e150, dmin = fd ? fd->_min : 1e150;
    if(!_hp || memcmp(_hpB, bounds, sizeof(bounds)) != 0 || _hpMax != dmax || _hpMin != dmin)
    {
      icp_destroy(_hp);
      _hp = NULL;
      OBVIOUS_B200_CHECK(icp_create(1, dmax, dmin, fd ? fd->_it : 2, bounds, 0, &_hp));
      memcpy(_hpB, bounds, sizeof(bounds)); _hpMax = dmax; _hpMin = dmin;
    }
    const size_t cap = std::min<size_t>((size_t)_sizeModel, idx.size());
    std::vector<uint32_t> pm(cap), ps(cap);
    std::vector<double> d(cap);
    uint32_t n = 0;
    OBVIOUS_B200_CHECK(icp_pairs(_hp, M.data(), _sizeModel, S.data(), (int32_t)idx.size(), pose, pm.data(), ps.data(), d.data(), &n));
    (void)reciprocal;  // (the device pass always applies the reciprocal filter: the node's wiring)
    for(uint32_t i = 0; i < n; i++)
    {
      StrCartesianIndexPair p;
      p.indexFirst = pm[i];
      p.indexSecond = idx[ps[i]];
      _pairs.push_back(p);
      _distancesSqr.push_back(d[i]);
    }
  }
  void determinePairs(double** scene, int size) { determinePairs(scene, NULL, size); }
  std::vector<StrCartesianIndexPair>* getPairs() { return &_pairs; }
  std::vector<double>* getDistancesSqr() { return &_distancesSqr; }
  std::vector<unsigned int>* getNonPairs() { return &_nonPairs; }
  void reset() { _pairs.clear(); _distancesSqr.clear(); _nonPairs.clear(); }
  std::vector<IPreAssignmentFilter*> _vPrefilter;
  std::vector<IPostAssignmentFilter*> _vPostfilter;

protected:
  int _dimension;
  double** _model;
  int _sizeModel;
  std::vector<StrCartesianIndexPair> _pairs;
  std::vector<double> _distancesSqr;
  std::vector<unsigned int> _nonPairs;
  tsd_icp_t* _hp;
  double _hpB[4] = {0, 0, 0, 0}, _hpMax = 0, _hpMin = 0;
};

class FlannPairAssignment : public PairAssignment  // assign/FlannPairAssignment.h
{
public:
  FlannPairAssignment(int dimension = 2, double eps = 0.0, bool parallelSearch = false) : PairAssignment(dimension)
  {
    (void)eps;
    (void)parallelSearch;
  }
};

class IRigidEstimator  // icp/IRigidEstimator.h
{
public:
  virtual ~IRigidEstimator() {}
};

class ClosedFormEstimator2D : public IRigidEstimator  // icp/ClosedFormEstimator2D.h
{
public:
  ClosedFormEstimator2D() {}
};

class Icp  // icp/Icp.h
{
public:
  Icp(PairAssignment* assigner, IRigidEstimator* estimator, int device = 0)
      : _assigner(assigner), _estimator(estimator), _h(NULL), _device(device), _maxIterations(3), _maxRMS(0.1), _convCnt(5),
        _Tfinal(3, 3)
  {
    _Tfinal.setIdentity();
  }
  ~Icp() { icp_destroy(_h); }
  void setMaxRMS(double rms) { _maxRMS = rms; }
  double getMaxRMS() { return _maxRMS; }
  void setMaxIterations(unsigned int iterations) { _maxIterations = iterations; }
  unsigned int getMaxIterations() { return _maxIterations; }
  void setConvergenceCounter(unsigned int convCnt) { _convCnt = convCnt; }
  unsigned int getConvergenceCounter() { return _convCnt; }
  PairAssignment* getPairAssigner() { return _assigner; }
  IRigidEstimator* getRigidEstimator() { return _estimator; }
  void reset() { _Tfinal.setIdentity(); }
  // Icp.cpp:150-205, :257-314 (probability 1.0 only: the node never subsamples)
  void setModel(Matrix* coords, Matrix* normals = NULL, double probability = 1.0)
  {
    (void)probability;
    _model.assign(coords->data(), coords->data() + (size_t)coords->getRows() * 2);
    _normals.clear();
    if(normals) _normals.assign(normals->data(), normals->data() + (size_t)normals->getRows() * 2);
  }
  void setScene(Matrix* coords, Matrix* normals = NULL, double probability = 1.0)
  {
    (void)normals;
    (void)probability;
    _scene.assign(coords->data(), coords->data() + (size_t)coords->getRows() * 2);
  }
  // Icp.cpp:464-512
  EnumIcpState iterate(double* rms, unsigned int* pairs, unsigned int* iterations, Matrix* Tinit = NULL)
  {
    OutOfBoundsFilter2D* fb = NULL;
    DistanceFilter* fd = NULL;
    ReciprocalFilter* fr = NULL;
    for(size_t i = 0; i < _assigner->_vPrefilter.size(); i++)
      if(!fb) fb = dynamic_cast<OutOfBoundsFilter2D*>(_assigner->_vPrefilter[i]);
    for(size_t i = 0; i < _assigner->_vPostfilter.size(); i++)
    {
      if(!fd) fd = dynamic_cast<DistanceFilter*>(_assigner->_vPostfilter[i]);
      if(!fr) fr = dynamic_cast<ReciprocalFilter*>(_assigner->_vPostfilter[i]);
    }
    if(!fb || !fd || !fr || !dynamic_cast<FlannPairAssignment*>(_assigner) || !dynamic_cast<ClosedFormEstimator2D*>(_estimator) ||
       _assigner->_vPrefilter.size() != 1 || _assigner->_vPostfilter.size() != 2)
    {
      fprintf(stderr, "obvious::Icp (b200): only the pipeline of ThreadLocalize.cpp:210-225 is supported\n");
      return ICP_ERROR;
    }
    if(!_h || _hIt != _maxIterations || _hMax != fd->_max || _hMin != fd->_min || _hDistIt != fd->_it ||
       memcmp(_hB, fb->_b, sizeof(_hB)) != 0)
    {
      icp_destroy(_h);
      _h = NULL;
      OBVIOUS_B200_CHECK(icp_create(_maxIterations, fd->_max, fd->_min, fd->_it, fb->_b, _device, &_h));
      _hIt = _maxIterations; _hMax = fd->_max; _hMin = fd->_min; _hDistIt = fd->_it;
      memcpy(_hB, fb->_b, sizeof(_hB));
    }
    OBVIOUS_B200_CHECK(icp_set_termination(_h, _maxRMS, _convCnt));
    double T9[9], pose[9], Ti[16];
    fb->_T.getData(pose);
    if(Tinit) Tinit->getData(Ti);
    int32_t state = ICP_ERROR;
    uint32_t p = 0, it = 0;
    OBVIOUS_B200_CHECK(icp_run(_h, _model.data(), _normals.empty() ? NULL : _normals.data(), (int32_t)(_model.size() / 2),
                               _scene.data(), (int32_t)(_scene.size() / 2), pose, Tinit ? Ti : NULL, T9, rms, &p, &it, &state));
    *pairs = p;
    *iterations = it;
    _Tfinal.setData(T9);
    return (EnumIcpState)state;
  }
  Matrix getFinalTransformation() { return _Tfinal; }  // Icp.cpp:528-546

private:
  PairAssignment* _assigner;
  IRigidEstimator* _estimator;
  tsd_icp_t* _h;
  int _device;
  unsigned int _maxIterations;
  double _maxRMS;
  unsigned int _convCnt;
  Matrix _Tfinal;
  std::vector<double> _model, _normals, _scene;
  unsigned int _hIt = 0, _hDistIt = 0;
  double _hMax = 0, _hMin = 0, _hB[4] = {0, 0, 0, 0};
};

// ---------------------------------------------------------------------------------------------------------
// RANSAC matchers (registration/ransacMatching/*).  Pre-processing on the host as in the reference
// (RandomMatching.cpp:41-183: PCA normals, validity, random scene subsampling, random control set, random
// trial order -- all with libc rand(), seeded with srand(time(NULL)) like TSD_PDFMatching.cpp:164); every
// hypothesis of the trial loops is scored on the device in one call.
// ---------------------------------------------------------------------------------------------------------
class RandomMatching
{
public:
  RandomMatching(unsigned int sizeControlSet, int device = 0)
      : _sizeControlSet(sizeControlSet), _m(NULL), _pcaSearchRange(10), _devPrep(false), _devSeed(1)
  {
    OBVIOUS_B200_CHECK(match_create(device, &_m));
  }
  virtual ~RandomMatching() { match_destroy(_m); }
  void activateTrace() {}
  void deactivateTrace() {}
  // Not in the reference: the pre-processing of match() (normals, subsampling, control set, trials, hypothesis list) on
  // the device with a counter-based random generator (match_prepare) instead of the host code with libc rand().  Same
  // distributions, different draws; every call of match() uses the next seed.
  void setDevicePreprocessing(bool on, uint64_t seed = 1)
  {
    _devPrep = on;
    _devSeed = seed;
  }

protected:
  struct Prep
  {
    bool ok;
    int n;
    std::vector<double> M, S, phiM, phiS, control, phiControl;
    std::vector<uint8_t> maskMpca, maskSpca;
    std::vector<unsigned int> idxMValid, idxSValid, idxControl;
    std::vector<tsd_hypothesis_t> hyps;
    double phiMax, thetaMin, thetaMax;
  };

  // obvious::Matrix::pcaAnalysis (obcore/math/linalg/gsl/Matrix.cpp:227-327) for an n x 2 matrix, statement by
  // statement, with the published algorithms of the GSL routines it calls: gsl_stats_mean (running mean in long
  // double), cblas_dgemm in the reference loop orders, gsl_linalg_SV_decomp_jacobi (one-sided Jacobi, linalg/svd.c).
  // axes: 2 x 4 row-major, {x0 x1 y0 y1} of the long and of the short principal axis.  Bit-identical with the reference
  // built against those routines (tests: the adapter's match() against tests/golden/matchers_tiny.npz).
  static double nrm2(const double* x, int n, int stride)  // gslcblas source_nrm2_r.h
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

  static void pcaAxes(const double* Ain, int rows, double axes[8])
  {
    const double eps = 2.2204460492503131e-16;  // GSL_DBL_EPSILON
    std::vector<double> M(Ain, Ain + 2 * (size_t)rows);
    double cent[2];
    for(int c = 0; c < 2; c++)
    {
      long double mean = 0;
      for(int i = 0; i < rows; i++) mean += (Ain[2 * i + c] - mean) / (i + 1);
      cent[c] = (double)mean;
    }
    for(int c = 0; c < 2; c++)
      for(int i = 0; i < rows; i++) M[2 * i + c] += -cent[c];
    // MtM = M' * M  (dgemm Trans, NoTrans: k outer, zero coefficients skipped)
    double A[4] = {0.0, 0.0, 0.0, 0.0};
    for(int k = 0; k < rows; k++)
      for(int i = 0; i < 2; i++)
      {
        const double temp = 1.0 * M[2 * k + i];
        if(temp != 0.0)
          for(int j = 0; j < 2; j++) A[2 * i + j] += temp * M[2 * k + j];
      }
    // one-sided Jacobi SVD of the 2 x 2 matrix A (columns j = 0, k = 1); V accumulates the rotations
    double V[4] = {1.0, 0.0, 0.0, 1.0}, S[2];
    const double tolerance = 10 * 2 * eps;
    for(int j = 0; j < 2; j++) S[j] = eps * nrm2(A + j, 2, 2);
    int count = 1, sweep = 0;
    const int sweepmax = 12;
    while(count > 0 && sweep <= sweepmax)
    {
      count = 1;
      {
        double pp = 0.0;
        for(int i = 0; i < 2; i++) pp += A[2 * i] * A[2 * i + 1];
        pp *= 2.0;
        const double a = nrm2(A, 2, 2), b = nrm2(A + 1, 2, 2);
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
      }
      sweep++;
    }
    // (the singular values and the normalisation of A's columns are not used by pcaAnalysis)
    // P = V' * M'  (dgemm Trans, Trans): P[i][j] = sum_k V[k][i] * M[j][k]
    std::vector<double> P(2 * (size_t)rows);
    for(int i = 0; i < 2; i++)
      for(int j = 0; j < rows; j++)
      {
        double temp = 0.0;
        for(int k = 0; k < 2; k++) temp += V[2 * k + i] * M[2 * j + k];
        P[(size_t)i * rows + j] = 0.0 + 1.0 * temp;
      }
    auto vmax = [&](int i) { double m = P[(size_t)i * rows]; for(int j = 0; j < rows; j++) if(P[(size_t)i * rows + j] > m) m = P[(size_t)i * rows + j]; return m; };
    auto vmin = [&](int i) { double m = P[(size_t)i * rows]; for(int j = 0; j < rows; j++) if(P[(size_t)i * rows + j] < m) m = P[(size_t)i * rows + j]; return m; };
    for(int i = 0; i < 2; i++)
    {
      const double max = vmax(i), min = vmin(i);
      const double ext = max - min;
      double align = 0.0;
      if(ext > 1e-6) align = (max + min) / 2.0;
      for(int j = 0; j < 2; j++)
      {
        const double e = V[2 * j + i] * align;
        cent[j] += e;
      }
    }
    for(int i = 0; i < 2; i++)
    {
      const double ext = vmax(i) - vmin(i);
      for(int j = 0; j < 2; j++)
      {
        const double e = V[2 * j + i] * ext / 2.0;
        axes[i * 4 + 2 * j] = cent[j] - e;
        axes[i * 4 + 2 * j + 1] = cent[j] + e;
      }
    }
  }

  // RandomMatching.cpp:77-146
  static void calcNormals(const std::vector<double>& M, std::vector<double>& N, int points, const bool* maskIn, uint8_t* maskOut,
                          int searchRadius)
  {
    for(int i = 0; i < searchRadius && i < points; i++) maskOut[i] = 0;
    for(int i = std::max(points - searchRadius, 0); i < points; i++) maskOut[i] = 0;
    std::vector<double> A(4 * searchRadius);
    for(int i = searchRadius; i < points - searchRadius; i++)
    {
      if(!maskIn[i]) continue;
      unsigned int cnt = 0;
      for(int j = -searchRadius; j < searchRadius; j++)
        if(maskIn[i + j]) cnt++;
      if(cnt > 3)
      {
        cnt = 0;
        for(int j = -searchRadius; j < searchRadius; j++)
          if(maskIn[i + j])
          {
            A[2 * cnt] = M[2 * (i + j)];
            A[2 * cnt + 1] = M[2 * (i + j) + 1];
            cnt++;
          }
        double Axes[8];
        pcaAxes(A.data(), (int)cnt, Axes);
        const double xLong = Axes[1] - Axes[0];
        const double yLong = Axes[3] - Axes[2];
        const double xShort = Axes[5] - Axes[4];
        const double yShort = Axes[7] - Axes[6];
        const double lenLongSqr = xLong * xLong + yLong * yLong;
        const double lenShortSqr = xShort * xShort + yShort * yShort;
        if(lenShortSqr > 1e-6 && (lenLongSqr / lenShortSqr) < 4.0)
        {
          maskOut[i] = 0;
          continue;
        }
        const double len = sqrt(lenShortSqr);
        if((M[2 * i] * xShort + M[2 * i + 1] * yShort) < 0.0)
        {
          N[2 * i] = xShort / len;
          N[2 * i + 1] = yShort / len;
        }
        else
        {
          N[2 * i] = -xShort / len;
          N[2 * i + 1] = -yShort / len;
        }
      }
      else
        maskOut[i] = 0;
    }
  }

  // TSD_PDFMatching.cpp:59-205 == RandomNormalMatching.cpp:94-247 == PDFMatching.cpp:67-233
  Prep prepare(Matrix* Mm, const bool* maskM, Matrix* Sm, const bool* maskS, unsigned int trialsIn, double phiMax, double resolution)
  {
    Prep p;
    p.ok = false;
    const int n = (int)Mm->getRows();
    p.n = n;
    if((int)Sm->getRows() != n || n < 3) return p;
    p.M.assign(Mm->data(), Mm->data() + 2 * (size_t)n);
    p.S.assign(Sm->data(), Sm->data() + 2 * (size_t)n);
    const int r = _pcaSearchRange / 2;
    std::vector<double> NM(2 * (size_t)n, 0.0), NS(2 * (size_t)n, 0.0);
    p.maskMpca.resize(n);
    p.maskSpca.resize(n);
    for(int i = 0; i < n; i++) { p.maskMpca[i] = maskM[i] ? 1 : 0; p.maskSpca[i] = maskS[i] ? 1 : 0; }
    calcNormals(p.M, NM, n, maskM, p.maskMpca.data(), r);
    p.phiM.resize(n);
    p.phiS.resize(n);
    for(int i = 0; i < n; i++) p.phiM[i] = p.maskMpca[i] ? atan2(NM[2 * i + 1], NM[2 * i]) : -1e6;  // calcPhi, RandomMatching.cpp:148-169
    for(int i = r; i < n - r; i++)
      if(p.maskMpca[i]) p.idxMValid.push_back(i);
    unsigned int validPoints = 0;
    for(int i = 0; i < n; i++)
      if(p.maskSpca[i]) validPoints++;
    const double probability = 180.0 / (double)validPoints;
    if(probability < 0.99)
    {
      const int thresh = (int)(1000.0 - std::min(std::max(probability, 0.0), 1.0) * 1000.0 + 0.5);
      for(int i = 0; i < n; i++)
        if((rand() % 1000) < thresh) p.maskSpca[i] = 0;
    }
    calcNormals(p.S, NS, n, maskS, p.maskSpca.data(), r);
    for(int i = 0; i < n; i++) p.phiS[i] = p.maskSpca[i] ? atan2(NS[2 * i + 1], NS[2 * i]) : -1e6;
    for(int i = r; i < n - r; i++)
      if(p.maskSpca[i]) p.idxSValid.push_back(i);
    // pickControlSet, RandomMatching.cpp:52-75
    unsigned int sizeControlSet = std::min<unsigned int>(_sizeControlSet, (unsigned int)p.idxSValid.size());
    std::vector<unsigned int> idxTemp = p.idxSValid;
    while(p.idxControl.size() < sizeControlSet)
    {
      const unsigned int rr = rand() % idxTemp.size();
      p.idxControl.push_back(idxTemp[rr]);
      idxTemp.erase(idxTemp.begin() + rr);
    }
    const size_t C = p.idxControl.size();
    p.control.resize(3 * std::max<size_t>(C, 1));
    p.phiControl.resize(std::max<size_t>(C, 1));
    for(size_t i = 0; i < C; i++)
    {
      const unsigned int idx = p.idxControl[i];
      p.control[i] = p.S[2 * idx];
      p.control[C + i] = p.S[2 * idx + 1];
      p.control[2 * C + i] = 1.0;
      p.phiControl[i] = atan2(NS[2 * idx + 1], NS[2 * idx]);
    }
    if(p.idxSValid.size() < 3 || p.idxMValid.size() < 3) return p;
    p.thetaMin = atan2(p.M[2 * p.idxMValid.front() + 1], p.M[2 * p.idxMValid.front()]);
    p.thetaMax = atan2(p.M[2 * p.idxMValid.back() + 1], p.M[2 * p.idxMValid.back()]);
    unsigned int trials = std::min<unsigned int>(trialsIn, (unsigned int)p.idxMValid.size());
    phiMax = std::min(phiMax, M_PI * 0.5);
    p.phiMax = phiMax;
    if(!(resolution > 1e-6)) return p;
    int span = (int)floor(phiMax / resolution);
    if(span > n) span = n;
    srand(time(NULL));
    std::vector<unsigned int> idxTrials = p.idxMValid;
    for(unsigned int trial = 0; trial < trials; trial++)
    {
      const int randIdx = rand() % (idxTrials.size());
      const int idx = (int)idxTrials[randIdx];
      idxTrials.erase(idxTrials.begin() + randIdx);
      const int iMin = std::max(idx - span, r);
      const int iMax = std::min(idx + span, n - r);
      for(int i = iMin; i < iMax; i++)
        if(p.maskSpca[i])
        {
          tsd_hypothesis_t h;
          h.idx_model = idx;
          h.idx_scene = i;
          p.hyps.push_back(h);
        }
    }
    p.ok = true;
    return p;
  }

  // match_prepare on M / S; false: nothing to score
  bool prepareOnDevice(Matrix* Mm, const bool* maskM, Matrix* Sm, const bool* maskS, unsigned int trials, double phiMax,
                       double resolution, tsd_match_prep_t* P)
  {
    const int n = (int)Mm->getRows();
    if((int)Sm->getRows() != n || n < 3 || !(resolution > 1e-6)) return false;
    std::vector<uint8_t> mm(n), ms(n);
    for(int i = 0; i < n; i++) { mm[i] = maskM[i] ? 1 : 0; ms[i] = maskS[i] ? 1 : 0; }
    OBVIOUS_B200_CHECK(match_prepare(_m, n, Mm->data(), mm.data(), Sm->data(), ms.data(), _pcaSearchRange, _sizeControlSet, trials,
                                     phiMax, resolution, _devSeed++, P));
    return P->n_hyp > 0;
  }

  static Matrix toMatrix(const double T[9])
  {
    Matrix M(3, 3);
    M.setData(const_cast<double*>(T));
    return M;
  }

  unsigned int _sizeControlSet;
  tsd_matcher_t* _m;
  int _pcaSearchRange;
  bool _devPrep;
  uint64_t _devSeed;
};

class TSD_PDFMatching : public RandomMatching  // ransacMatching/TSD_PDFMatching.h
{
public:
  TSD_PDFMatching(TsdGrid& grid, unsigned int trials = 30, double epsThresh = 0.15, unsigned int sizeControlSet = 360,
                  double zrand = 0.05)
      : RandomMatching(sizeControlSet), _grid(grid), _trials(trials), _zrand(zrand)
  {
    (void)epsThresh;
  }
  // TSD_PDFMatching.cpp:31-294
  Matrix match(Matrix TSensor, Matrix* M, const bool* maskM, Matrix* NM, Matrix* S, const bool* maskS,
               double phiMax = M_PI / 4.0, const double transMax = 1.5, const double resolution = 0.0)
  {
    (void)NM; (void)transMax;
    Matrix TBest(3, 3);
    TBest.setIdentity();
    double ts[9], T[9];
    TSensor.getData(ts);
    int32_t best = -1;
    if(_devPrep)
    {
      tsd_match_prep_t P;
      if(!prepareOnDevice(M, maskM, S, maskS, _trials, phiMax, resolution, &P)) return TBest;
      OBVIOUS_B200_CHECK(match_score_tsd(_m, _grid.handle(), P.n_hyp, P.hyps, P.n, P.model, P.scene, P.phi_m, P.phi_s, P.phi_max,
                                         P.n_control, P.control, ts, _zrand, NULL, &best, T));
      return toMatrix(T);
    }
    Prep p = prepare(M, maskM, S, maskS, _trials, phiMax, resolution);
    if(!p.ok || p.hyps.empty()) return TBest;
    OBVIOUS_B200_CHECK(match_score_tsd(_m, _grid.handle(), (int32_t)p.hyps.size(), p.hyps.data(), p.n, p.M.data(), p.S.data(),
                                       p.phiM.data(), p.phiS.data(), p.phiMax, (int32_t)p.idxControl.size(), p.control.data(), ts,
                                       _zrand, NULL, &best, T));
    return toMatrix(T);
  }

private:
  TsdGrid& _grid;
  unsigned int _trials;
  double _zrand;
};

class RandomNormalMatching : public RandomMatching  // ransacMatching/RandomNormalMatching.h
{
public:
  RandomNormalMatching(unsigned int trials = 50, double epsThresh = 0.15, unsigned int sizeControlSet = 180)
      : RandomMatching(sizeControlSet), _trials(trials), _scaleDistance(1.0 / (epsThresh * epsThresh)), _scaleOrientation(0.33)
  {
  }
  // RandomNormalMatching.cpp:67-395
  Matrix match(Matrix* M, const bool* maskM, Matrix* NM, Matrix* S, const bool* maskS, double phiMax = M_PI / 4.0,
               const double transMax = 1.5, const double resolution = 0.0)
  {
    (void)NM; (void)transMax;
    Matrix TBest(3, 3);
    TBest.setIdentity();
    if(_devPrep)
    {
      tsd_match_prep_t P;
      double Td[9];
      int32_t bestd = -1;
      if(!prepareOnDevice(M, maskM, S, maskS, _trials, phiMax, resolution, &P)) return TBest;
      OBVIOUS_B200_CHECK(match_score_rnm(_m, P.n_hyp, P.hyps, P.n, P.model, P.scene, P.phi_m, P.phi_s, P.phi_max, P.n_control,
                                         P.control, P.phi_control, P.n_valid_m, P.model_valid, P.phi_valid, P.theta_min, P.theta_max,
                                         _scaleDistance, _scaleOrientation, (uint32_t)(P.n_control / 3), NULL, NULL, NULL, &bestd, Td));
      return toMatrix(Td);
    }
    Prep p = prepare(M, maskM, S, maskS, _trials, phiMax, resolution);
    if(!p.ok || p.hyps.empty()) return TBest;
    std::vector<double> mv(2 * p.idxMValid.size()), pv(p.idxMValid.size());
    for(size_t k = 0; k < p.idxMValid.size(); k++)
    {
      mv[2 * k] = p.M[2 * p.idxMValid[k]];
      mv[2 * k + 1] = p.M[2 * p.idxMValid[k] + 1];
      pv[k] = p.phiM[p.idxMValid[k]];
    }
    double T[9];
    int32_t best = -1;
    OBVIOUS_B200_CHECK(match_score_rnm(_m, (int32_t)p.hyps.size(), p.hyps.data(), p.n, p.M.data(), p.S.data(), p.phiM.data(),
                                       p.phiS.data(), p.phiMax, (int32_t)p.idxControl.size(), p.control.data(),
                                       p.phiControl.data(), (int32_t)pv.size(), mv.data(), pv.data(), p.thetaMin, p.thetaMax,
                                       _scaleDistance, _scaleOrientation, (uint32_t)(p.idxControl.size() / 3), NULL, NULL, NULL,
                                       &best, T));
    return toMatrix(T);
  }

private:
  unsigned int _trials;
  double _scaleDistance, _scaleOrientation;
};

class PDFMatching : public RandomMatching  // ransacMatching/PDFMatching.h
{
public:
  PDFMatching(unsigned int trials = 100, double epsThresh = 0.15, unsigned int sizeControlSet = 140, double zhit = 0.45,
              double zphi = 0.0, double zshort = 0.25, double zmax = 0.05, double zrand = 0.25, double percentagePointsInC = 0.9,
              double rangemax = 20, double sigphi = M_PI / 180.0 * 3, double sighit = 0.2, double lamshort = 0.08,
              double maxAngleDiff = 3.0, double maxAnglePenalty = 0.5)
      : RandomMatching(sizeControlSet), _trials(trials)
  {
    (void)epsThresh;
    const double p[12] = {zhit, zphi, zshort, zmax, zrand, percentagePointsInC, rangemax, sigphi, sighit, lamshort, maxAngleDiff,
                          maxAnglePenalty};
    memcpy(_p, p, sizeof(p));
  }
  // PDFMatching.cpp:47-432
  Matrix match(Matrix* M, const bool* maskM, Matrix* NM, Matrix* S, const bool* maskS, double phiMax = M_PI / 4.0,
               const double transMax = 1.5, const double resolution = 0.0)
  {
    (void)NM; (void)transMax;
    Matrix TBest(3, 3);
    TBest.setIdentity();
    if(_devPrep)
    {
      tsd_match_prep_t P;
      double Td[9];
      int32_t bestd = -1;
      if(!prepareOnDevice(M, maskM, S, maskS, _trials, phiMax, resolution, &P)) return TBest;
      OBVIOUS_B200_CHECK(match_score_pdf(_m, P.n_hyp, P.hyps, P.n, P.model, P.scene, P.phi_m, P.phi_s, P.phi_max, P.n_control,
                                         P.control, P.n_valid_m, P.model_angles, P.model_dists, _p, NULL, NULL, &bestd, Td));
      return toMatrix(Td);
    }
    Prep p = prepare(M, maskM, S, maskS, _trials, phiMax, resolution);
    if(!p.ok || p.hyps.empty()) return TBest;
    std::vector<double> ang(p.idxMValid.size()), dst(p.idxMValid.size());
    for(size_t k = 0; k < p.idxMValid.size(); k++)
    {
      const double x = p.M[2 * p.idxMValid[k]], y = p.M[2 * p.idxMValid[k] + 1];
      ang[k] = atan2(y, x);
      dst[k] = sqrt(x * x + y * y);
    }
    double T[9];
    int32_t best = -1;
    OBVIOUS_B200_CHECK(match_score_pdf(_m, (int32_t)p.hyps.size(), p.hyps.data(), p.n, p.M.data(), p.S.data(), p.phiM.data(),
                                       p.phiS.data(), p.phiMax, (int32_t)p.idxControl.size(), p.control.data(), (int32_t)ang.size(),
                                       ang.data(), dst.data(), _p, NULL, NULL, &best, T));
    return toMatrix(T);
  }

private:
  unsigned int _trials;
  double _p[12];
};

}  // namespace obvious
