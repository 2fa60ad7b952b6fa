// A ROS-free replay of what ThreadLocalize / ThreadMapping do with the obvious:: classes
// (reference src/ThreadLocalize.cpp:411-511 init, :310-409 eventLoop, :513-591 doRegistration with
// registration_mode ICP or TSD; src/ThreadMapping.cpp:32-62).  The SAME source builds against
//   (A) the reference's own headers + oracle/_ref/libohm_ref.so   (-I/root/reference/src -Ioracle/shim), and
//   (B) the B200 adapter headers + libtsdslam_b200.so             (-Iohm_tsd_slam_b200/obvious),
// which is the drop-in claim of INTEGRATION.md.  Input: a binary file of float32 scans; output: one pose per
// line.  usage: slam_loop <scans.bin> <layout_grid> <cell_size> <trunc_cells> <beams> <ang_res> <phi_min>
//                         <max_range> <min_range> <low_refl> <mode 0|3>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "obcore/base/Logger.h"
#include "obcore/math/mathbase.h"
#include "obvision/reconstruct/grid/RayCastAxisAligned2D.h"
#include "obvision/reconstruct/grid/RayCastPolar2D.h"
#include "obvision/reconstruct/grid/SensorPolar2D.h"
#include "obvision/reconstruct/grid/TsdGrid.h"
#include "obvision/registration/icp/icp_def.h"
#include "obvision/registration/ransacMatching/TSD_PDFMatching.h"

// ThreadLocalize::maskMatrix (ThreadLocalize.cpp:738-755)
static obvious::Matrix maskMatrix(obvious::Matrix* Mat, bool* mask, unsigned int maskSize, unsigned int validPoints)
{
  obvious::Matrix retMat(validPoints, 2);
  unsigned int cnt = 0;
  for(unsigned int i = 0; i < maskSize; i++)
    if(mask[i])
    {
      retMat(cnt, 0) = (*Mat)(i, 0);
      retMat(cnt, 1) = (*Mat)(i, 1);
      cnt++;
    }
  return retMat;
}

int main(int argc, char** argv)
{
  if(argc < 12) { fprintf(stderr, "usage: see the header comment\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if(!f) { perror("scans"); return 2; }
  const int layoutGrid = atoi(argv[2]);
  const double cellSize = atof(argv[3]), truncCells = atof(argv[4]);
  const unsigned int beams = (unsigned int)atoi(argv[5]);
  const double angRes = atof(argv[6]), phiMin = atof(argv[7]), maxRange = atof(argv[8]), minRange = atof(argv[9]), lowRefl = atof(argv[10]);
  const int mode = atoi(argv[11]);
  const int icpIterations = 30;
  LOGMSG_CONF("slam.log", obvious::Logger::file_off | obvious::Logger::screen_off, DBG_DEBUG, DBG_DEBUG);  // src/slam.cpp:17

  // SlamNode.cpp:77-78
  obvious::TsdGrid* grid = NULL;
  bool sharded = false;
#ifdef OBVIOUS_B200_H
  // adapter build only: SLAM_LOOP_BANDS=n runs the same loop on a grid sharded in n bands inside the library
  if(getenv("SLAM_LOOP_BANDS"))
  {
    grid = new obvious::TsdGrid(cellSize, obvious::LAYOUT_32x32, (obvious::EnumTsdGridLayout)layoutGrid, atoi(getenv("SLAM_LOOP_BANDS")), NULL);
    sharded = true;
  }
#endif
  if(!grid) grid = new obvious::TsdGrid(cellSize, obvious::LAYOUT_32x32, (obvious::EnumTsdGridLayout)layoutGrid);
  grid->setMaxTruncation(truncCells * cellSize);

  // ThreadLocalize ctor, ThreadLocalize.cpp:205-225
  obvious::RayCastPolar2D* rayCaster = new obvious::RayCastPolar2D();
  obvious::PairAssignment* assigner = new obvious::FlannPairAssignment(2);
  obvious::DistanceFilter* filterDist = new obvious::DistanceFilter(0.4, 0.02, icpIterations - 10);
  obvious::ReciprocalFilter* filterReciprocal = new obvious::ReciprocalFilter();
  obvious::IRigidEstimator* estimator = new obvious::ClosedFormEstimator2D();
  obvious::OutOfBoundsFilter2D* filterBounds = new obvious::OutOfBoundsFilter2D(grid->getMinX(), grid->getMaxX(), grid->getMinY(), grid->getMaxY());
  assigner->addPreFilter(filterBounds);
  assigner->addPostFilter(filterDist);
  assigner->addPostFilter(filterReciprocal);
  obvious::Icp* icp = new obvious::Icp(assigner, estimator);
  icp->setMaxRMS(0.0);
  icp->setMaxIterations(icpIterations);
  icp->setConvergenceCounter(icpIterations);
  obvious::TSD_PDFMatching* tsdMatcher = (mode == 3) ? new obvious::TSD_PDFMatching(*grid, 100, 0.15, 140, 0.25) : NULL;
#ifdef OBVIOUS_B200_H
  // adapter build only: SLAM_LOOP_DEVICE_PREP=1 moves the matcher's pre-processing to the device (match_prepare)
  if(tsdMatcher && getenv("SLAM_LOOP_DEVICE_PREP")) tsdMatcher->setDevicePreprocessing(true, 42);
#endif

  std::vector<float> ranges(beams);
  obvious::SensorPolar2D* sensor = NULL;
  double* scene = new double[beams * 2];
  bool* maskS = new bool[beams];
  double* modelCoords = new double[beams * 2]();
  double* modelNormals = new double[beams * 2]();
  bool* maskM = new bool[beams];
  int k = 0;
  while(fread(ranges.data(), sizeof(float), beams, f) == beams)
  {
    if(!sensor)
    {
      // ThreadLocalize::init, ThreadLocalize.cpp:466-507
      const double startX = grid->getCellsX() * grid->getCellSize() * 0.5, startY = grid->getCellsY() * grid->getCellSize() * 0.5;
      double tf[9] = {1, -0.0, startX, 0, 1, startY, 0, 0, 1};
      obvious::Matrix Tinit(3, 3);
      Tinit.setData(tf);
      sensor = new obvious::SensorPolar2D(beams, angRes, phiMin, maxRange, minRange, lowRefl);
      sensor->setRealMeasurementData(ranges, 1.0);
      sensor->setStandardMask();
      sensor->transform(&Tinit);
      obvious::obfloat t[2] = {startX, startY};
      if(!grid->freeFootprint(t, 0.6, 0.6)) fprintf(stderr, "footprint could not be freed\n");
      grid->push(sensor);  // ThreadMapping::initPush
    }
    else
    {
      // ThreadLocalize::eventLoop, ThreadLocalize.cpp:321-406
      sensor->setRealMeasurementData(ranges);
      sensor->setStandardMask();
      const unsigned int validModelPoints = rayCaster->calcCoordsFromCurrentViewMask(grid, sensor, modelCoords, modelNormals, maskM);
      if(validModelPoints == 0) { printf("%d no-model\n", k++); continue; }
      const unsigned int validScenePoints = sensor->dataToCartesianVectorMask(scene, maskS);
      obvious::Matrix M(beams, 2, modelCoords), N(beams, 2, modelNormals), S(beams, 2, scene);
      obvious::Matrix Mvalid = maskMatrix(&M, maskM, beams, validModelPoints);
      obvious::Matrix Nvalid = maskMatrix(&N, maskM, beams, validModelPoints);
      obvious::Matrix Svalid = maskMatrix(&S, maskS, beams, validScenePoints);
      // doRegistration, ThreadLocalize.cpp:513-581
      obvious::Matrix T44(4, 4);
      T44.setIdentity();
      obvious::Matrix T(3, 3);
      if(tsdMatcher)
      {
        T = tsdMatcher->match(sensor->getTransformation(), &M, maskM, NULL, &S, maskS, obvious::deg2rad(30.0), 0.25, sensor->getAngularResolution());
        T44(0, 0) = T(0, 0); T44(0, 1) = T(0, 1); T44(0, 3) = T(0, 2);
        T44(1, 0) = T(1, 0); T44(1, 1) = T(1, 1); T44(1, 3) = T(1, 2);
      }
      icp->reset();
      obvious::Matrix P = sensor->getTransformation();
      filterBounds->setPose(&P);
      icp->setModel(&Mvalid, &Nvalid);
      icp->setScene(&Svalid);
      double rms = 0.0;
      unsigned int pairs = 0, it = 0;
      icp->iterate(&rms, &pairs, &it, &T44);
      T = icp->getFinalTransformation();
      sensor->transform(&T);
      grid->push(sensor);  // every scan is pushed (stricter than the node's 5 cm / 0.03 rad gate)
      obvious::Matrix pose = sensor->getTransformation();
      printf("%d %.17g %.17g %.17g %u %u %u\n", k, pose(0, 2), pose(1, 2), atan2(pose(1, 0), pose(0, 0)), validModelPoints, pairs, it);
    }
    k++;
  }
  fclose(f);
  if(!sharded)  // the publisher's calls want an unsharded grid
  {
    // what ThreadGrid::eventLoop does with the map (ThreadGrid.cpp:17-28, :84, :125): crossings + occupancy grid,
    // colour image.  One summary line: -1 <crossings> <sum x> <sum y> <free cells> <image byte sum> 0
    const unsigned int cx = grid->getCellsX(), cy = grid->getCellsY();
    std::vector<char> occ((size_t)cx * cy, -1);
    std::vector<double> gridCoords((size_t)cx * cy);
    obvious::RayCastAxisAligned2D raycasterMap;
    unsigned int mapSize = 0;
    raycasterMap.calcCoords(grid, gridCoords.data(), NULL, &mapSize, occ.data());
    double sx = 0.0, sy = 0.0;
    for(unsigned int i = 0; i < mapSize / 2; i++) { sx += gridCoords[2 * i]; sy += gridCoords[2 * i + 1]; }
    unsigned long freeCells = 0;
    for(size_t i = 0; i < occ.size(); i++) freeCells += (occ[i] == 0);
    std::vector<unsigned char> img((size_t)3 * 160 * 120);
    grid->grid2ColorImage(img.data(), 160, 120);
    unsigned long imgSum = 0;
    for(size_t i = 0; i < img.size(); i++) imgSum += img[i];
    printf("-1 %u %.17g %.17g %lu %lu 0\n", mapSize / 2, sx, sy, freeCells, imgSum);
  }
  return 0;
}
