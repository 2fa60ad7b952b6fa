// TEST INFRASTRUCTURE ONLY -- part of oracle/ (see oracle/README.md).
//
// C API over the UNMODIFIED reference classes (namespace obvious), compiled from
// the sources where they lie under /root/reference/src by oracle/Makefile into
// oracle/_ref/libohm_ref.so.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load that library; the product
// (libtsdslam_b200.so) never does.
//
// Every entry point is a thin forwarder to the reference method named in its
// comment; no arithmetic of the hot path is re-implemented here.
//
// Determinism: the reference matchers call srand(time(NULL)) and rand()
// (RandomMatching.cpp:65,178; TSD_PDFMatching.cpp:164,190; Icp.cpp:90-93).  This
// library defines its own rand()/srand() (linked -Bsymbolic so the reference TUs
// bind to them): srand() is a no-op and rand() is a fixed LCG that only
// ref_seed() re-seeds, so that a parity run can replay the same draw sequence.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <omp.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

// Read-only access to TsdGridPartition::_initWeight / _grid for state dumps: the
// reference keeps them private (TsdGridPartition.h:140-166).  Object layout is
// unchanged by this; every standard header is included above already.
#define private public
#include "obvision/reconstruct/grid/TsdGridPartition.h"
#undef private

#include "obcore/base/Logger.h"
#include "obcore/math/linalg/linalg.h"
#include "obvision/reconstruct/grid/RayCastAxisAligned2D.h"
#include "obvision/reconstruct/grid/RayCastPolar2D.h"
#include "obvision/reconstruct/grid/SensorPolar2D.h"
#include "obvision/reconstruct/grid/TsdGrid.h"
#include "obvision/registration/icp/ClosedFormEstimator2D.h"
#include "obvision/registration/icp/Icp.h"
#include "obvision/registration/icp/assign/FlannPairAssignment.h"
#include "obvision/registration/icp/assign/filter/DistanceFilter.h"
#include "obvision/registration/icp/assign/filter/OutOfBoundsFilter2D.h"
#include "obvision/registration/icp/assign/filter/ReciprocalFilter.h"
#include "obvision/registration/ransacMatching/PDFMatching.h"
#include "obvision/registration/ransacMatching/RandomNormalMatching.h"
#include "obvision/registration/ransacMatching/TSD_PDFMatching.h"

// ---------------------------------------------------------------- deterministic libc RNG
static uint32_t g_lcg = 12345u;
static unsigned long long g_randCalls = 0;

extern "C" int rand(void) noexcept
{
  g_lcg = g_lcg * 1103515245u + 12345u;
  g_randCalls++;
  return (int)((g_lcg >> 8) & 0x7fffffu);
}

extern "C" void srand(unsigned int) noexcept
{
  // deliberately ignored: see header comment
}

extern "C" void ref_seed(unsigned int seed)
{
  g_lcg = seed;
  g_randCalls = 0;
}

extern "C" unsigned long long ref_rand_calls(void) { return g_randCalls; }

extern "C" void ref_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
extern "C" int ref_max_threads(void) { return omp_get_max_threads(); }

using namespace obvious;

extern "C" {

// ---------------------------------------------------------------- helpers
void ref_quiet(void)
{
  // src/slam.cpp:17 configures the logger with file_off|screen_off
  LOGMSG_CONF("", Logger::file_off | Logger::screen_off, DBG_ERROR, DBG_ERROR);
}

// obvious::Matrix::invert (gsl/Matrix.cpp:168-179) on an n x n row-major matrix
void ref_invert(int n, const double* in, double* out)
{
  Matrix M(n, n);
  M.setData(const_cast<double*>(in));
  M.invert();
  M.getData(out);
}

// operator* (gsl/Matrix.cpp:90-95)
void ref_matmul(int n, const double* a, const double* b, double* out)
{
  Matrix A(n, n), B(n, n);
  A.setData(const_cast<double*>(a));
  B.setData(const_cast<double*>(b));
  Matrix C = A * B;
  C.getData(out);
}

// ---------------------------------------------------------------- TsdGrid
void* ref_grid_create(double cellSize, int layoutPartition, int layoutGrid)
{
  return new TsdGrid(cellSize, (EnumTsdGridLayout)layoutPartition, (EnumTsdGridLayout)layoutGrid);
}

void ref_grid_destroy(void* g) { delete (TsdGrid*)g; }

void ref_grid_set_max_truncation(void* g, double v) { ((TsdGrid*)g)->setMaxTruncation(v); }
double ref_grid_get_max_truncation(void* g) { return ((TsdGrid*)g)->getMaxTruncation(); }
int ref_grid_cells_x(void* g) { return (int)((TsdGrid*)g)->getCellsX(); }
double ref_grid_min_x(void* g) { return ((TsdGrid*)g)->getMinX(); }
double ref_grid_max_x(void* g) { return ((TsdGrid*)g)->getMaxX(); }
double ref_grid_min_y(void* g) { return ((TsdGrid*)g)->getMinY(); }
double ref_grid_max_y(void* g) { return ((TsdGrid*)g)->getMaxY(); }

int ref_grid_free_footprint(void* g, double cx, double cy, double w, double h)
{
  obfloat c[2] = {cx, cy};
  return ((TsdGrid*)g)->freeFootprint(c, w, h) ? 1 : 0;
}

// TsdGrid::push (TsdGrid.cpp:217-284)
void ref_grid_push(void* g, void* s) { ((TsdGrid*)g)->push((SensorPolar2D*)s); }

int ref_grid_num_partitions(void* g)
{
  TsdGrid* grid = (TsdGrid*)g;
  const int per = grid->getCellsX() / grid->getPartitionSize();
  return per * per;
}

// state: 0 uninitialised, 1 empty (initWeight > 0, not initialised), 2 content  (TsdGrid.h:32-34)
int ref_grid_partition_state(void* g, int p, double* initWeight)
{
  TsdGridPartition* part = ((TsdGrid*)g)->getPartitions()[0][p];
  if(initWeight) *initWeight = part->_initWeight;
  if(part->isInitialized()) return 2;
  if(part->isEmpty()) return 1;
  return 0;
}

// all partitions at once: state[P], initWeight[P]
void ref_grid_partition_states(void* g, int* state, double* initWeight)
{
  TsdGrid* grid = (TsdGrid*)g;
  const int per = grid->getCellsX() / grid->getPartitionSize();
  for(int p = 0; p < per * per; p++) state[p] = ref_grid_partition_state(g, p, initWeight + p);
}

// Copies the (dim+1)x(dim+1) cell array of an initialised partition, row-major,
// border row/column included (struct TsdCell {tsd, weight}, TsdGridPartition.h:16-20).
int ref_grid_download_partition(void* g, int p, double* tsd, double* weight)
{
  TsdGridPartition* part = ((TsdGrid*)g)->getPartitions()[0][p];
  if(!part->isInitialized()) return 0;
  const unsigned int w = part->getWidth();
  const unsigned int h = part->getHeight();
  unsigned int i = 0;
  for(unsigned int y = 0; y <= h; y++)
    for(unsigned int x = 0; x <= w; x++, i++)
    {
      tsd[i] = part->_grid[y][x].tsd;
      weight[i] = part->_grid[y][x].weight;
    }
  return 1;
}

// Bench set-up aid (no hot-path arithmetic): allocate every partition through the public
// TsdGridPartition::init and set all (dim+1)^2 cells, so that a push runs in the dense regime
// (every in-range partition already allocated, BASELINE.md section 2 "observation").
void ref_grid_fill(void* g, double tsd, double weight, int only_uninitialized)
{
  TsdGrid* grid = (TsdGrid*)g;
  const int per = grid->getCellsX() / grid->getPartitionSize();
  for(int p = 0; p < per * per; p++)
  {
    TsdGridPartition* part = grid->getPartitions()[0][p];
    if(only_uninitialized && part->isInitialized()) continue;
    part->init(grid->getMaxTruncation());
    const unsigned int w = part->getWidth(), h = part->getHeight();
    for(unsigned int y = 0; y <= h; y++)
      for(unsigned int x = 0; x <= w; x++)
      {
        part->_grid[y][x].tsd = tsd;
        part->_grid[y][x].weight = weight;
      }
  }
}

// TsdGrid::storeGrid (TsdGrid.cpp:548-607)
int ref_grid_store(void* g, const char* path) { return ((TsdGrid*)g)->storeGrid(path) ? 1 : 0; }
// TsdGrid(const std::string&, FILE_SOURCE) (TsdGrid.cpp:25-110)
void* ref_grid_load(const char* path) { return new TsdGrid(std::string(path), FILE_SOURCE); }

// TsdGrid::interpolateBilinear (TsdGrid.h:284-304)
void ref_grid_interpolate_bilinear(void* g, int n, const double* xy, double* tsd, int* status)
{
  TsdGrid* grid = (TsdGrid*)g;
  for(int i = 0; i < n; i++)
  {
    obfloat c[2] = {xy[2 * i], xy[2 * i + 1]};
    obfloat v = NAN;
    status[i] = (int)grid->interpolateBilinear(c, &v);
    tsd[i] = v;
  }
}

// TsdGrid::interpolateNormal (TsdGrid.cpp:517-546)
void ref_grid_interpolate_normal(void* g, int n, const double* xy, double* normals, int* ok)
{
  TsdGrid* grid = (TsdGrid*)g;
  for(int i = 0; i < n; i++)
  {
    obfloat c[2] = {xy[2 * i], xy[2 * i + 1]};
    obfloat nn[2] = {NAN, NAN};
    ok[i] = grid->interpolateNormal(c, nn) ? 1 : 0;
    normals[2 * i] = nn[0];
    normals[2 * i + 1] = nn[1];
  }
}

// ---------------------------------------------------------------- SensorPolar2D
void* ref_sensor_create(int beams, double angularRes, double phiMin, double maxRange, double minRange, double lowReflectivityRange)
{
  return new SensorPolar2D(beams, angularRes, phiMin, maxRange, minRange, lowReflectivityRange);
}
void ref_sensor_destroy(void* s) { delete (SensorPolar2D*)s; }

// Sensor::setRealMeasurementData(double*, 1.0) (Sensor.cpp:125-134)
void ref_sensor_set_data(void* s, const double* ranges)
{
  ((SensorPolar2D*)s)->setRealMeasurementData(const_cast<double*>(ranges), 1.0);
}
// Sensor::setRealMeasurementData(vector<float>, 1.0) (Sensor.cpp:136-145), the node's path
void ref_sensor_set_data_f32(void* s, const float* ranges, int n)
{
  std::vector<float> v(ranges, ranges + n);
  ((SensorPolar2D*)s)->setRealMeasurementData(v, 1.0f);
}
void ref_sensor_set_standard_mask(void* s) { ((SensorPolar2D*)s)->setStandardMask(); }
void ref_sensor_set_mask(void* s, const unsigned char* mask)
{
  SensorPolar2D* sen = (SensorPolar2D*)s;
  bool* m = sen->getRealMeasurementMask();
  for(unsigned int i = 0; i < sen->getRealMeasurementSize(); i++) m[i] = mask[i] != 0;
}
void ref_sensor_get_data(void* s, double* out)
{
  SensorPolar2D* sen = (SensorPolar2D*)s;
  memcpy(out, sen->getRealMeasurementData(), sen->getRealMeasurementSize() * sizeof(double));
}
void ref_sensor_get_mask(void* s, unsigned char* out)
{
  SensorPolar2D* sen = (SensorPolar2D*)s;
  const bool* m = sen->getRealMeasurementMask();
  for(unsigned int i = 0; i < sen->getRealMeasurementSize(); i++) out[i] = m[i] ? 1 : 0;
}
// Sensor::transform (Sensor.cpp:50-60)
void ref_sensor_transform(void* s, const double* T9)
{
  Matrix T(3, 3);
  T.setData(const_cast<double*>(T9));
  ((SensorPolar2D*)s)->transform(&T);
}
void ref_sensor_set_pose(void* s, const double* T9)
{
  Matrix T(3, 3);
  T.setData(const_cast<double*>(T9));
  ((SensorPolar2D*)s)->setTransformation(T);
}
void ref_sensor_get_pose(void* s, double* T9)
{
  Matrix T = ((SensorPolar2D*)s)->getTransformation();
  T.getData(T9);
}
// Sensor::getNormalizedRayMap (Sensor.cpp:36-48); out is 2 x N row-major
void ref_sensor_get_normalized_rays(void* s, double norm, double* out)
{
  Matrix* R = ((SensorPolar2D*)s)->getNormalizedRayMap(norm);
  R->getData(out);
}
double ref_sensor_phi_lower(void* s) { return ((SensorPolar2D*)s)->getPhiLowerBound(); }
double ref_sensor_phi_upper(void* s) { return ((SensorPolar2D*)s)->getPhiUpperBound(); }

// SensorPolar2D::backProject(Matrix*, int*, Matrix*) (SensorPolar2D.cpp:117-135); xy is n x 2
void ref_sensor_back_project(void* s, int n, const double* xy, int* idx)
{
  Matrix M(n, 3);
  for(int i = 0; i < n; i++)
  {
    M(i, 0) = xy[2 * i];
    M(i, 1) = xy[2 * i + 1];
    M(i, 2) = 1.0;
  }
  ((SensorPolar2D*)s)->backProject(&M, idx);
}

// Sensor::dataToCartesianVectorMask (Sensor.cpp:168-190)
unsigned int ref_sensor_data_to_cartesian_mask(void* s, double* coords, unsigned char* mask)
{
  SensorPolar2D* sen = (SensorPolar2D*)s;
  const unsigned int n = sen->getRealMeasurementSize();
  bool* m = new bool[n];
  unsigned int valid = sen->dataToCartesianVectorMask(coords, m);
  for(unsigned int i = 0; i < n; i++) mask[i] = m[i] ? 1 : 0;
  delete[] m;
  return valid;
}

// ---------------------------------------------------------------- RayCastPolar2D
// RayCastPolar2D::calcCoordsFromCurrentViewMask (RayCastPolar2D.cpp:113-192)
unsigned int ref_raycast_mask(void* g, void* s, double* coords, double* normals, unsigned char* mask)
{
  SensorPolar2D* sen = (SensorPolar2D*)s;
  const unsigned int n = sen->getRealMeasurementSize();
  bool* m = new bool[n];
  RayCastPolar2D rc;
  unsigned int cnt = rc.calcCoordsFromCurrentViewMask((TsdGrid*)g, sen, coords, normals, m);
  for(unsigned int i = 0; i < n; i++) mask[i] = m[i] ? 1 : 0;
  delete[] m;
  return cnt;
}

// RayCastPolar2D::calcCoordsFromCurrentView (RayCastPolar2D.cpp:27-111), compacting variant
unsigned int ref_raycast_compact(void* g, void* s, double* coords, double* normals)
{
  RayCastPolar2D rc;
  unsigned int cnt = 0;
  rc.calcCoordsFromCurrentView((TsdGrid*)g, (SensorPolar2D*)s, coords, normals, &cnt);
  return cnt;
}

// ---------------------------------------------------------------- map publication (ThreadGrid.cpp:84,125)
// RayCastAxisAligned2D::calcCoords (RayCastAxisAligned2D.cpp:13-105); normals / occupied may be NULL.
// Returns the reference's cnt (number of doubles written to coords).
unsigned int ref_axis_map(void* g, double* coords, double* normals, signed char* occupied)
{
  RayCastAxisAligned2D rc;
  unsigned int cnt = 0;
  rc.calcCoords((TsdGrid*)g, coords, normals, &cnt, (char*)occupied);
  return cnt;
}

// TsdGrid::grid2ColorImage (TsdGrid.cpp:429-488)
void ref_color_image(void* g, unsigned char* image, unsigned int width, unsigned int height)
{
  ((TsdGrid*)g)->grid2ColorImage(image, width, height);
}

// ---------------------------------------------------------------- ICP (wired as ThreadLocalize.cpp:210-225)
struct RefIcp
{
  FlannPairAssignment* assigner;
  OutOfBoundsFilter2D* filterBounds;
  DistanceFilter* filterDist;
  ReciprocalFilter* filterReciprocal;
  ClosedFormEstimator2D* estimator;
  Icp* icp;
};

void* ref_icp_create(unsigned int maxIterations, double distMax, double distMin, unsigned int distIterations,
                     double xMin, double xMax, double yMin, double yMax)
{
  RefIcp* r = new RefIcp;
  r->assigner = new FlannPairAssignment(2);
  r->filterDist = new DistanceFilter(distMax, distMin, distIterations);
  r->filterReciprocal = new ReciprocalFilter();
  r->estimator = new ClosedFormEstimator2D();
  r->filterBounds = new OutOfBoundsFilter2D(xMin, xMax, yMin, yMax);
  r->assigner->addPreFilter(r->filterBounds);
  r->assigner->addPostFilter(r->filterDist);
  r->assigner->addPostFilter(r->filterReciprocal);
  r->icp = new Icp(r->assigner, r->estimator);
  r->icp->setMaxRMS(0.0);
  r->icp->setMaxIterations(maxIterations);
  r->icp->setConvergenceCounter(maxIterations);
  return r;
}

void ref_icp_destroy(void* h)
{
  RefIcp* r = (RefIcp*)h;
  delete r->icp;
  delete r->assigner;
  delete r->filterBounds;
  delete r->filterDist;
  delete r->filterReciprocal;
  delete r->estimator;
  delete r;
}

// ThreadLocalize::doRegistration (ThreadLocalize.cpp:571-581): reset, setPose, setModel,
// setScene, iterate, getFinalTransformation.  model/normals are nM x 2, scene nS x 2.
int ref_icp_run(void* h, const double* model, const double* normals, int nM, const double* scene, int nS,
                const double* pose9, const double* Tinit16, double* Tout9, double* rms, unsigned int* pairs,
                unsigned int* iterations)
{
  RefIcp* r = (RefIcp*)h;
  Matrix Mvalid(nM, 2, const_cast<double*>(model));
  Matrix Nvalid(nM, 2, const_cast<double*>(normals));
  Matrix Svalid(nS, 2, const_cast<double*>(scene));
  Matrix P(3, 3);
  P.setData(const_cast<double*>(pose9));
  Matrix T44(4, 4);
  T44.setData(const_cast<double*>(Tinit16));

  r->icp->reset();
  r->filterBounds->setPose(&P);
  r->icp->setModel(&Mvalid, &Nvalid);
  r->icp->setScene(&Svalid);
  *rms = 0.0;
  *pairs = 0;
  *iterations = 0;
  EnumIcpState st = r->icp->iterate(rms, pairs, iterations, &T44);
  Matrix T = r->icp->getFinalTransformation();
  T.getData(Tout9);
  return (int)st;
}

// Same set-up, but driven through the public Icp::step() so that the pair list of
// every iteration can be captured (Icp.cpp:410-462).  Icp::iterate applies Tinit
// through the private applyTransformation; here Tinit must be identity.
// pairModel/pairScene: maxIt x cap arrays, pairCount[maxIt]; Tfinal16 after each step: maxIt x 16.
int ref_icp_trace(void* h, const double* model, const double* normals, int nM, const double* scene, int nS,
                  const double* pose9, int maxIt, int cap, unsigned int* pairModel, unsigned int* pairScene,
                  int* pairCount, double* rmsOut, double* Tfinal16)
{
  RefIcp* r = (RefIcp*)h;
  Matrix Mvalid(nM, 2, const_cast<double*>(model));
  Matrix Nvalid(nM, 2, const_cast<double*>(normals));
  Matrix Svalid(nS, 2, const_cast<double*>(scene));
  Matrix P(3, 3);
  P.setData(const_cast<double*>(pose9));
  r->icp->reset();
  r->filterBounds->setPose(&P);
  r->icp->setModel(&Mvalid, &Nvalid);
  r->icp->setScene(&Svalid);
  r->icp->reset(); // copies scene into the working buffer and resets Tfinal (Icp.cpp:333-339)
  int it = 0;
  for(; it < maxIt; it++)
  {
    double rms = 0.0;
    unsigned int pairs = 0;
    EnumIcpState st = r->icp->step(&rms, &pairs);
    std::vector<StrCartesianIndexPair>* pv = r->assigner->getPairs();
    pairCount[it] = (int)pv->size();
    for(int k = 0; k < (int)pv->size() && k < cap; k++)
    {
      pairModel[(size_t)it * cap + k] = (*pv)[k].indexFirst;
      pairScene[(size_t)it * cap + k] = (*pv)[k].indexSecond;
    }
    rmsOut[it] = rms;
    Matrix T4 = r->icp->getFinalTransformation4x4();
    T4.getData(Tfinal16 + 16 * it);
    if(st != ICP_PROCESSING) { it++; break; }
  }
  return it;
}

// ---------------------------------------------------------------- matchers
// M, S: n x 2 row-major with validity masks (ray model preserved), as ThreadLocalize.cpp:369-377.
static void unpack(int n, const double* xy, const unsigned char* mask8, Matrix& M, bool* mask)
{
  for(int i = 0; i < n; i++)
  {
    M(i, 0) = xy[2 * i];
    M(i, 1) = xy[2 * i + 1];
    mask[i] = mask8[i] != 0;
  }
}

// TSD_PDFMatching::match (TSD_PDFMatching.cpp:31-294), ctor as ThreadLocalize.cpp:190
void ref_match_tsd(void* g, unsigned int trials, double epsThresh, unsigned int sizeControlSet, double zrand,
                   const double* TSensor9, int n, const double* M, const unsigned char* maskM, const double* S,
                   const unsigned char* maskS, double phiMax, double transMax, double resolution, double* Tout9)
{
  TSD_PDFMatching matcher(*(TsdGrid*)g, trials, epsThresh, sizeControlSet, zrand);
  Matrix Mm(n, 2), Sm(n, 2), TS(3, 3);
  bool* mM = new bool[n];
  bool* mS = new bool[n];
  unpack(n, M, maskM, Mm, mM);
  unpack(n, S, maskS, Sm, mS);
  TS.setData(const_cast<double*>(TSensor9));
  Matrix T = matcher.match(TS, &Mm, mM, NULL, &Sm, mS, phiMax, transMax, resolution);
  T.getData(Tout9);
  delete[] mM;
  delete[] mS;
}

// RandomNormalMatching::match (RandomNormalMatching.cpp:67-395), ctor as ThreadLocalize.cpp:183
void ref_match_rnm(unsigned int trials, double epsThresh, unsigned int sizeControlSet, int n, const double* M,
                   const unsigned char* maskM, const double* S, const unsigned char* maskS, double phiMax,
                   double transMax, double resolution, double* Tout9)
{
  RandomNormalMatching matcher(trials, epsThresh, sizeControlSet);
  Matrix Mm(n, 2), Sm(n, 2);
  bool* mM = new bool[n];
  bool* mS = new bool[n];
  unpack(n, M, maskM, Mm, mM);
  unpack(n, S, maskS, Sm, mS);
  Matrix T = matcher.match(&Mm, mM, NULL, &Sm, mS, phiMax, transMax, resolution);
  T.getData(Tout9);
  delete[] mM;
  delete[] mS;
}

// PDFMatching::match (PDFMatching.cpp:47-432), ctor as ThreadLocalize.cpp:186-187;
// params: zhit zphi zshort zmax zrand percentagePointsInC rangemax sigphi sighit lamshort maxAngleDiff maxAnglePenalty
void ref_match_pdf(unsigned int trials, double epsThresh, unsigned int sizeControlSet, const double* params12, int n,
                   const double* M, const unsigned char* maskM, const double* S, const unsigned char* maskS,
                   double phiMax, double transMax, double resolution, double* Tout9)
{
  const double* p = params12;
  PDFMatching matcher(trials, epsThresh, sizeControlSet, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9],
                      p[10], p[11]);
  Matrix Mm(n, 2), Sm(n, 2);
  bool* mM = new bool[n];
  bool* mS = new bool[n];
  unpack(n, M, maskM, Mm, mM);
  unpack(n, S, maskS, Sm, mS);
  Matrix T = matcher.match(&Mm, mM, NULL, &Sm, mS, phiMax, transMax, resolution);
  T.getData(Tout9);
  delete[] mM;
  delete[] mS;
}

// PDFMatching::probabilityOfTwoSingleScans (PDFMatching.cpp:435-487), public
double ref_pdf_probability(const double* params12, double m, double s, double phiDiff)
{
  const double* p = params12;
  PDFMatching matcher(1, 0.15, 1, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11]);
  return matcher.probabilityOfTwoSingleScans(m, s, phiDiff);
}

} // extern "C"
