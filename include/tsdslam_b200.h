/*
 * tsdslam_b200 -- C ABI of the B200-native (sm_100a) mapping/localisation hot path of ohm_tsd_slam.
 *
 * This is the drop-in boundary.  The reference has no FFI: its extension points are the public
 * classes of namespace obvious, used in-process by ThreadMapping / ThreadLocalize.  Each entry point
 * below names the reference interface it replaces (file:line relative to the reference tree); the
 * header-compatible C++ adapter over this ABI lives in ohm_tsd_slam_b200/obvious/ (INTEGRATION.md).
 *
 * Conventions
 *   - every call returns 0 on success, a negative TSD_E_* code otherwise; tsd_last_error() gives text;
 *   - plain pointers and sizes only; the caller owns all host buffers, the library owns device memory;
 *   - all floating point is IEEE double, as in the reference (obfloat == double, obcore/base/types.h:28-31);
 *   - matrices are row-major; a pose is the 3x3 homogeneous matrix of Sensor::getTransformation();
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with TSD_E_NO_DEVICE.
 *   - one handle = one CUDA stream; calls on one handle are ordered, calls on different handles may
 *     overlap (the reference itself shares the grid between threads without a lock, ThreadMapping.cpp:46-61).
 */
#ifndef TSDSLAM_B200_H
#define TSDSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSD_OK 0
#define TSD_E_INVALID (-1)   /* bad argument */
#define TSD_E_NO_DEVICE (-2) /* no CUDA device / driver: the product has no CPU path */
#define TSD_E_CUDA (-3)      /* a CUDA runtime call failed */
#define TSD_E_NOMEM (-4)
#define TSD_E_RANGE (-5)     /* coordinates outside the grid (freeFootprint) */

/* EnumTsdGridInterpolate, reconstruct/grid/TsdGrid.h:28-31 */
#define TSD_INTERPOLATE_SUCCESS 0
#define TSD_INTERPOLATE_INVALIDINDEX 1
#define TSD_INTERPOLATE_EMPTYPARTITION 2
#define TSD_INTERPOLATE_ISNAN 3

/* EnumTsdGridPartitionIdentifier, reconstruct/grid/TsdGrid.h:32-34 */
#define TSD_PARTITION_UNINITIALIZED 0
#define TSD_PARTITION_EMPTY 1
#define TSD_PARTITION_CONTENT 2

/* EnumIcpState, registration/icp/Icp.h:25-32 */
#define TSD_ICP_IDLE 0
#define TSD_ICP_PROCESSING 1
#define TSD_ICP_NOTMATCHABLE 2
#define TSD_ICP_MAXITERATIONS 3
#define TSD_ICP_TIMEELAPSED 4
#define TSD_ICP_SUCCESS 5
#define TSD_ICP_CONVERGED 6
#define TSD_ICP_ERROR 7

const char* tsd_last_error(void);
/* number of CUDA devices visible, 0 when there is none */
int tsd_device_count(void);
/* total kernel launches issued by this library in this process (bench.py's gpu_launches) */
uint64_t tsd_kernel_launches(void);

/* Inverse of a 3x3 pose: LU with partial pivoting and column-wise solves, the routine behind
 * obvious::Matrix::invert (obcore/math/linalg/gsl/Matrix.cpp:168-179) as SensorPolar2D::backProject
 * (SensorPolar2D.cpp:120-121) and RayCastPolar2D (RayCastPolar2D.cpp:120-121) use it.  Host code. */
int tsd_invert3x3(const double in[9], double out[9]);

/* What the kernels read out of an obvious::SensorPolar2D (reconstruct/grid/SensorPolar2D.h,
 * reconstruct/Sensor.h): measurement data + mask, pose and the scalar sensor model. */
typedef struct tsd_scan
{
  int32_t n;             /* Sensor::getRealMeasurementSize() */
  int32_t _pad;
  const double* ranges;  /* Sensor::getRealMeasurementData(), host, n */
  const uint8_t* mask;   /* Sensor::getRealMeasurementMask(), host, n (bool) */
  double pose[9];        /* Sensor::getTransformation() */
  double pose_inv[9];    /* its inverse (tsd_invert3x3) */
  double phi_min;        /* SensorPolar2D::getPhiMin() */
  double angular_res;    /* SensorPolar2D::getAngularResolution() */
  double phi_lower;      /* SensorPolar2D::getPhiLowerBound() */
  double phi_upper;      /* SensorPolar2D::getPhiUpperBound() */
  double max_range;      /* Sensor::getMaximumRange() */
  double min_range;      /* Sensor::getMinimumRange() */
  double low_reflectivity_range; /* Sensor::getLowReflectivityRange() */
} tsd_scan_t;

/* ------------------------------------------------------------------------------------------------
 * TsdGrid  (reconstruct/grid/TsdGrid.h, TsdGrid.cpp; partitions: TsdGridPartition.h/.cpp,
 * classification: TsdGridComponent.cpp:43-124)
 * ---------------------------------------------------------------------------------------------- */
typedef struct tsd_grid tsd_grid_t;

/* TsdGrid::TsdGrid(cellSize, layoutPartition, layoutGrid)  (TsdGrid.cpp:20-23,112-169; SlamNode.cpp:77).
 * layoutPartition must be 5 (32x32 cells, the only layout the node uses); layoutGrid 5..16.
 * device: CUDA ordinal.  Cell state of the whole grid is allocated densely in HBM up front. */
int tsdg_create(double cell_size, int layout_partition, int layout_grid, int device, tsd_grid_t** out);

/* Same grid geometry, but this handle owns only the partition rows [row_begin, row_end) (a band of a
 * grid sharded over several GPUs, one process per GPU).  Cells outside the band are never touched;
 * one halo partition row above the band is kept for the replicated borders. */
int tsdg_create_band(double cell_size, int layout_partition, int layout_grid, int device, int part_row_begin,
                     int part_row_end, tsd_grid_t** out);
/* On a band, a push integrates the scan into the band's own rows and maintains every border strip that does
 * not depend on another band.  Pushes need no communication.  Before anything READS the map across a band
 * boundary (ray casting, interpolation, download), the caller brings the halos up to date:
 *   1. exchange boundary partition rows with the neighbouring bands (tsdg_band_row: which 0/1 = my lowest /
 *      highest row to send, 2/3 = the halo slots below / above to receive into; count doubles each of tsd and
 *      weight; NULL when there is no such neighbour),
 *   2. tsdg_band_push_finish: the band's top row takes its top / corner border strips from the halo above
 *      (TsdGrid::propagateBorders, TsdGrid.cpp:401-424, across the band boundary),
 *      and the halo row below the band gets the top / corner strips that mirror this band's first row.
 * Any number of pushes may precede one such synchronisation; the result equals the unsharded grid's. */
int tsdg_band_push_finish(tsd_grid_t* grid);
/* The same synchronisation in ONE kernel over peer memory (NVLink / NVSwitch P2P): each band stores its boundary
 * rows directly into its neighbours' halo rows, signals them, waits for theirs and completes the borders.
 * Set-up, once: every process exports its band (tsdg_band_export: CUDA IPC handles of its arrays, an opaque blob of
 * TSD_BAND_EXPORT_BYTES) and connects the bands below (side 0) and above (side 1) with the blobs it received;
 * bands living in one process connect with tsdg_band_connect_local.  tsdg_band_halo_sync is a collective among
 * neighbours: both sides of a boundary call it equally often, with the same dirty partition columns
 * [px0, px1] for that boundary (px1 < px0: nothing changed there, the boundary is skipped on both sides).  It only
 * enqueues work on the grid's stream. */
#define TSD_BAND_EXPORT_BYTES 256
int tsdg_band_export(tsd_grid_t* grid, void* blob);
int tsdg_band_connect(tsd_grid_t* grid, int side, const void* blob);
int tsdg_band_connect_local(tsd_grid_t* grid, int side, tsd_grid_t* neighbour);
int tsdg_band_halo_sync(tsd_grid_t* grid, int lo_px0, int lo_px1, int hi_px0, int hi_px1);
/* Allocation flags of ALL partitions of the grid (device memory, one byte each, 1 = allocated; row-major).  A band
 * keeps the flags of its own rows and of the row on either side current; ray casting walks rays through other
 * bands' rows too (the partition-skipping loop of RayCastPolar2D.cpp:223-235), so before a ray cast the caller
 * merges the bands' flags with an element-wise MAX (flags only ever go 0 -> 1). */
int tsdg_band_flags(tsd_grid_t* grid, uint8_t** flags, uint64_t* count);
/* Partition box {px0, py0, px1, py1} (inclusive) a scan can touch: every partition outside fails the range cull
 * of TsdGridComponent::isInRange (TsdGridComponent.cpp:50-58).  The classifier of a push looks at this box only;
 * a sharded caller uses it to skip scans that cannot reach its band and to bound the halo columns to exchange. */
int tsdg_scan_box(const tsd_grid_t* grid, const tsd_scan_t* scan, int32_t box[4]);
int tsdg_band_row(tsd_grid_t* grid, int which, double** tsd, double** weight, uint64_t* count);
int tsdg_destroy(tsd_grid_t* grid);

/* TsdGrid::setMaxTruncation (TsdGrid.cpp:206-215; SlamNode.cpp:78).  Must precede the first push. */
int tsdg_set_max_truncation(tsd_grid_t* grid, double val);
/* getters of TsdGrid.h:79-150 */
int tsdg_get_geometry(const tsd_grid_t* grid, int32_t* cells_x, int32_t* cells_y, int32_t* partition_size,
                      double* cell_size, double* min_x, double* max_x, double* min_y, double* max_y,
                      double* max_truncation);

/* TsdGrid::freeFootprint (TsdGrid.cpp:609-638; ThreadLocalize.cpp:504).  TSD_E_RANGE when out of bounds. */
int tsdg_free_footprint(tsd_grid_t* grid, double cx, double cy, double width, double height);

/* TsdGrid::push(SensorPolar2D*) (TsdGrid.cpp:217-284; ThreadMapping.cpp:38,55): classify partitions,
 * integrate the scan, refresh the replicated borders.  Blocking. */
int tsdg_push(tsd_grid_t* grid, const tsd_scan_t* scan);
/* Enqueue only (host buffers are copied before return); tsdg_sync() waits. */
int tsdg_push_async(tsd_grid_t* grid, const tsd_scan_t* scan);
int tsdg_sync(tsd_grid_t* grid);
/* The two halves of tsdg_push_async: copy a scan to the device once, integrate it (again) without any
 * host-to-device traffic.  Lets a caller keep the measurement resident (bench.py's device-resident leg). */
int tsdg_stage_scan(tsd_grid_t* grid, const tsd_scan_t* scan);
int tsdg_push_staged(tsd_grid_t* grid);
/* n scans in the order given, with the result of n TsdGrid::push calls (ThreadMapping::eventLoop drains its queue
 * of sensors one push after the other, ThreadMapping.cpp:43-62).  Consecutive scans of the same sensor model are
 * taken four or two at a time: one classification and one update launch for all of them, partitions seen by several
 * scans (the two lasers of a robot) are read and written once.  tsdg_push_batch blocks like tsdg_push; the statistics
 * reported afterwards are those of the last launch (its scans together).  tsdg_stage_batch (n = 1, 2 or 4, same
 * sensor model) + tsdg_push_staged is the device-resident variant. */
int tsdg_push_batch(tsd_grid_t* grid, const tsd_scan_t* scans, int32_t n);
int tsdg_push_batch_async(tsd_grid_t* grid, const tsd_scan_t* scans, int32_t n);
int tsdg_stage_batch(tsd_grid_t* grid, const tsd_scan_t* scans, int32_t n);
/* The handle's cudaStream_t, for callers that time or order work with CUDA events. */
void* tsdg_stream(tsd_grid_t* grid);
/* Orders the handle's stream against a caller's stream without blocking the host: direction 0 = `other`
 * waits for the handle's queued work, 1 = the handle waits for `other` (used around NCCL exchanges). */
int tsdg_stream_order(tsd_grid_t* grid, void* other_stream, int direction);
/* Measurement aid: with timing enabled every push records CUDA events around its kernels on the handle's
 * stream; ms = {tables + classify, update (K2+K3), borders (K4), whole push} of the last completed push. */
int tsdg_set_timing(tsd_grid_t* grid, int enable);
/* Measurement aid (bench.py's K2-only / K3-only roofline rows): the update kernel of the following pushes skips the
 * addTsd work of active partitions (mask bit 0) and / or the increaseEmptiness work of emptied partitions (bit 1).
 * The map is then NOT what the reference would compute; 0 restores the normal behaviour. */
int tsdg_set_update_filter(tsd_grid_t* grid, unsigned mask);
int tsdg_last_push_kernel_ms(tsd_grid_t* grid, float ms[4]);

/* Counters of the most recent completed push. */
typedef struct tsd_push_stats
{
  uint64_t cell_updates;     /* cells whose {tsd,weight} were rewritten (addTsd or increaseEmptiness) */
  uint64_t cell_visits;      /* cells of active partitions */
  uint32_t active_tiles;     /* partitions that passed isInRange */
  uint32_t emptied_tiles;    /* partitions on which increaseEmptiness ran (allocated or not) */
  uint32_t newly_initialized;/* partitions allocated by this push */
  uint32_t fallback_cells;   /* beam index resolved by the exact slow path */
} tsd_push_stats_t;
int tsdg_last_push_stats(tsd_grid_t* grid, tsd_push_stats_t* out);

/* TsdGrid::interpolateBilinear (TsdGrid.h:284-304; TSD_PDFMatching.cpp:241), batched: xy is n x 2. */
int tsdg_interpolate_bilinear(tsd_grid_t* grid, int32_t n, const double* xy, double* tsd, int32_t* status);
/* TsdGrid::interpolateNormal (TsdGrid.cpp:517-546), batched; ok[i] = 1 on success. */
int tsdg_interpolate_normal(tsd_grid_t* grid, int32_t n, const double* xy, double* normals, int32_t* ok);

/* Partition accessors (TsdGrid::getPartitions, TsdGridPartition::isInitialized/isEmpty and the state that
 * TsdGrid::storeGrid writes, TsdGrid.cpp:548-607).  state[p] is a TSD_PARTITION_* value. */
int tsdg_num_partitions(const tsd_grid_t* grid, int32_t* n);
int tsdg_partition_states(tsd_grid_t* grid, int32_t* state, double* init_weight);
/* (dim+1) x (dim+1) row-major arrays, border row/column included, like TsdGridPartition::_grid.
 * Returns TSD_E_INVALID for an uninitialised partition. */
int tsdg_download_partition(tsd_grid_t* grid, int32_t p, double* tsd, double* weight);
int tsdg_upload_partition(tsd_grid_t* grid, int32_t p, const double* tsd, const double* weight);
/* Allocate partitions with the given cell value: all of them, or only those not yet initialised
 * (bench: the dense regime, where every in-range partition of a push is already allocated). */
int tsdg_fill(tsd_grid_t* grid, double tsd, double weight, int only_uninitialized);

/* ------------------------------------------------------------------------------------------------
 * RayCastPolar2D  (reconstruct/grid/RayCastPolar2D.h/.cpp)
 * rays_world: 2 x n row-major = *Sensor::getNormalizedRayMap(cellSize) (Sensor.cpp:36-48), i.e. world-frame
 * beam directions of length cellSize.  Outputs are in the SENSOR frame (RayCastPolar2D.cpp:172-177).
 * ---------------------------------------------------------------------------------------------- */
/* calcCoordsFromCurrentViewMask (RayCastPolar2D.cpp:113-192; ThreadLocalize.cpp:353): coords/normals are
 * n x 2, written only where mask[i] != 0; *count = number of hits. */
int tsdg_raycast_mask(tsd_grid_t* grid, const tsd_scan_t* scan, const double* rays_world, double* coords,
                      double* normals, uint8_t* mask, uint32_t* count);
/* calcCoordsFromCurrentView (RayCastPolar2D.cpp:27-111): compacted, in beam order; *count = doubles written. */
int tsdg_raycast(tsd_grid_t* grid, const tsd_scan_t* scan, const double* rays_world, double* coords,
                 double* normals, uint32_t* count);

/* Sharded grid: per-beam first event among the steps whose sample lies in THIS band.
 * key[i] = 4*step + code (code 0: hit, 1: hit whose normal lookup failed, 2: abort; INT64_MAX: none);
 * payload[i] = {cx, cy, nx, ny} in the sensor frame for a hit, zeros otherwise.  Both stay on the device
 * (pointers valid until the next raycast on this handle).  The caller min-reduces the keys over all bands
 * (ncclAllReduce min), zeroes the payload of the beams it did not win and sum-reduces the payloads; a beam
 * is a hit iff the winning key has code 0 (ohm_tsd_slam_b200/sharded.py). */
int tsdg_raycast_band_keys(tsd_grid_t* grid, const tsd_scan_t* scan, const double* rays_world,
                           uint64_t** dev_keys, double** dev_payload);
/* The same merge done by the library over peer memory, one call per band (a COLLECTIVE: every band of the grid calls
 * it with the same scan): the marching kernel of each band stores its per-beam first events straight into every band's
 * exchange block (NVLink P2P stores / CUDA-IPC mappings) and signals; a second kernel waits for all bands' signals and
 * keeps, per beam, the earliest event.  Every band returns the full result -- what RayCastPolar2D::
 * calcCoordsFromCurrentViewMask returns on the unsharded grid (RayCastPolar2D.cpp:113-192).  Set-up, once: every band
 * exports its exchange block (tsdg_band_rcx_export, a blob of TSD_BAND_EXPORT_BYTES) and connects with the blobs of ALL
 * bands, ordered by rank (tsdg_band_rcx_connect; world <= 16, scans <= 2048 beams); bands living in one process
 * connect with tsdg_band_rcx_connect_local.  Halos and allocation flags must be current, as for tsdg_raycast_band_keys. */
int tsdg_band_rcx_export(tsd_grid_t* grid, void* blob);
int tsdg_band_rcx_connect(tsd_grid_t* grid, int rank, int world, const void* blobs);
int tsdg_band_rcx_connect_local(tsd_grid_t* grid, int rank, int world, tsd_grid_t** bands);
int tsdg_raycast_mask_sharded(tsd_grid_t* grid, const tsd_scan_t* scan, const double* rays_world, double* coords,
                              double* normals, uint8_t* mask, uint32_t* count);
/* Its two halves, for bands that live in one process (a blocking call per band would wait for bands not yet launched):
 * tsdg_raycast_sharded_launch on every band, then tsdg_raycast_sharded_collect on every band. */
int tsdg_raycast_sharded_launch(tsd_grid_t* grid, const tsd_scan_t* scan, const double* rays_world);
int tsdg_raycast_sharded_collect(tsd_grid_t* grid, int32_t n, double* coords, double* normals, uint8_t* mask, uint32_t* count);
/* step counters of the most recent raycast on this handle */
int tsdg_last_raycast_steps(tsd_grid_t* grid, uint64_t* fine_steps, uint64_t* coarse_steps);

/* ------------------------------------------------------------------------------------------------
 * Scan pre-processing on the device (what ThreadLocalize does to a LaserScan before the hot path starts):
 *   Sensor::setRealMeasurementData(vector<float>, scale)  (reconstruct/Sensor.cpp:136-145): data[i] = (double)(ranges[i] * scale)
 *   SensorPolar2D::setStandardMask (reconstruct/grid/SensorPolar2D.cpp:59-98; Sensor.cpp:246-272): zero depth, ranges beyond
 *       max_range -> inf, NaN -> inf + masked, depth discontinuities of less than 3 degrees masked
 *   Sensor::dataToCartesianVectorMask (Sensor.cpp:168-190): scene[i] = rays_local[:, i] * data[i] where valid (n x 2)
 *   ThreadLocalize::maskMatrix (src/ThreadLocalize.cpp:738-755): the valid scene points compacted, beam order (n_valid x 2)
 * rays_local: 2 x n row-major, SensorPolar2D::_raysLocal.  scene / scene_mask / scene_valid / n_valid may be NULL.  The
 * grid handle supplies device, stream and staging memory; the map is not touched.
 * ---------------------------------------------------------------------------------------------- */
int tsds_prepare_scan(tsd_grid_t* grid, int32_t n, const float* ranges, float scale, double max_range, double angular_res,
                      const double* rays_local, double* data, uint8_t* mask, double* scene, uint8_t* scene_mask,
                      double* scene_valid, uint32_t* n_valid);

/* ------------------------------------------------------------------------------------------------
 * Map publication (ThreadGrid.cpp:84,125), on the device: the map never travels to the host.
 * ------------------------------------------------------------------------------------------------ */
/* RayCastAxisAligned2D::calcCoords (RayCastAxisAligned2D.cpp:13-105): zero crossings of the TSD along the cell rows
 * and columns of every allocated inner partition, in the reference's order (partitions row-major; rows, then
 * columns).  coords: 2 doubles per crossing, room for cap_points crossings (TSD_E_RANGE if there are more: the
 * first cap_points are written); *count = number of DOUBLES written, as the reference's cnt.  normals (may be
 * NULL): as in the reference every normal is the one at the first crossing (the call passes the array base,
 * RayCastAxisAligned2D.cpp:52,73), and the array is left alone when that lookup fails.  occupied (may be NULL):
 * cells_x * cells_y bytes, in/out: 0 free, -1 occupied / unknown; cells calcCoords does not write keep the
 * caller's value (ThreadGrid initialises the array to -1 once). */
int tsdg_axis_aligned_map(tsd_grid_t* grid, double* coords, uint32_t cap_points, double* normals, uint32_t* count,
                          int8_t* occupied);
/* TsdGrid::grid2ColorImage (TsdGrid.cpp:429-488): 3 bytes per pixel, width x height. */
int tsdg_color_image(tsd_grid_t* grid, uint8_t* image, uint32_t width, uint32_t height);

/* The reference's checkpoint format: TsdGrid::storeGrid (TsdGrid.cpp:548-607) and TsdGrid(path, FILE_SOURCE)
 * (:25-110).  Text, 6 significant digits per value, borders not stored (the next push refreshes them).  The files
 * are interchangeable with the reference's. */
int tsdg_store(tsd_grid_t* grid, const char* path);
int tsdg_load(const char* path, int device, tsd_grid_t** out);

/* ------------------------------------------------------------------------------------------------
 * Icp + FlannPairAssignment + OutOfBoundsFilter2D + DistanceFilter + ReciprocalFilter +
 * ClosedFormEstimator2D, wired as ThreadLocalize.cpp:210-225 and run as ThreadLocalize.cpp:571-581.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tsd_icp tsd_icp_t;

/* bounds = {xMin, xMax, yMin, yMax} of OutOfBoundsFilter2D (OutOfBoundsFilter2D.cpp:8-15);
 * dist_iterations is DistanceFilter's third ctor argument (DistanceFilter.cpp:11-20), the node passes
 * icp_iterations - 10. */
int icp_create(uint32_t max_iterations, double dist_max, double dist_min, uint32_t dist_iterations,
               const double bounds[4], int device, tsd_icp_t** out);
int icp_destroy(tsd_icp_t* icp);
/* Icp::setMaxRMS / setConvergenceCounter (Icp.cpp:341-369).  icp_create sets 0.0 and max_iterations, the
 * node's values (ThreadLocalize.cpp:223-225). */
int icp_set_termination(tsd_icp_t* icp, double max_rms, uint32_t convergence_counter);
/* Icp::setMaxIterations (Icp.cpp:351-354); at most the value given to icp_create. */
int icp_set_max_iterations(tsd_icp_t* icp, uint32_t max_iterations);
/* Icp::reset + OutOfBoundsFilter2D::setPose + setModel + setScene + iterate + getFinalTransformation.
 * model/normals: n_model x 2, scene: n_scene x 2 (already compacted to valid points); t_init: 4x4 or NULL.
 * t_out 3x3; *mse is what the reference calls rms (mean squared pair distance, ClosedFormEstimator2D.cpp:59-62). */
int icp_run(tsd_icp_t* icp, const double* model, const double* normals, int32_t n_model, const double* scene,
            int32_t n_scene, const double pose[9], const double* t_init, double t_out[9], double* mse,
            uint32_t* pairs, uint32_t* iterations, int32_t* state);
/* One localisation step with the model kept on the device: RayCastPolar2D::calcCoordsFromCurrentViewMask from the scan's
 * pose, ThreadLocalize::maskMatrix and Icp::iterate (reference src/ThreadLocalize.cpp:333-361 and :571-581) -- the result
 * of tsdg_raycast_mask + icp_run(model = the hits in beam order, pose = scan->pose) without the two host round trips.
 * scene: n_scene x 2, the scan's valid points in the sensor frame (ThreadLocalize.cpp:341-352); *n_model: the number of
 * hits (the caller's validModelPoints).  grid and icp must live on the same device; at most 2048 beams. */
int tsdg_localize(tsd_grid_t* grid, tsd_icp_t* icp, const tsd_scan_t* scan, const double* rays_world, const double* scene,
                  int32_t n_scene, const double* t_init, double t_out[9], double* mse, uint32_t* pairs,
                  uint32_t* iterations, int32_t* state, uint32_t* n_model);
/* PairAssignment::determinePairs on its own (registration/icp/assign/PairAssignment.cpp:38-84 with the filters the node
 * wires, ThreadLocalize.cpp:210-221: OutOfBoundsFilter2D.cpp:27-37 before; FlannPairAssignment.cpp:64-92 exact 1-NN;
 * DistanceFilter.cpp:32-64 at its initial threshold and ReciprocalFilter.cpp:32-78 after).  Pairs in model-index
 * order: pair_model[i] = indexFirst, pair_scene[i] = indexSecond, dist_sqr[i] (may be NULL) their squared distance.
 * Arrays must hold min(n_model, n_scene) entries.  pose: the sensor pose the bounds filter transforms the scene with. */
int icp_pairs(tsd_icp_t* icp, const double* model, int32_t n_model, const double* scene, int32_t n_scene, const double pose[9],
              uint32_t* pair_model, uint32_t* pair_scene, double* dist_sqr, uint32_t* n_pairs);
/* Parity aid, off by default: record the pair list of every iteration (icp_get_trace). */
int icp_set_trace(tsd_icp_t* icp, int enable);
/* Parity aid: pair list, mse and accumulated 4x4 of every iteration of the last icp_run (tracing enabled).
 * pair_model/pair_scene: max_it x cap; returns iterations stored through *n_it. */
int icp_get_trace(tsd_icp_t* icp, int32_t max_it, int32_t cap, uint32_t* pair_model, uint32_t* pair_scene,
                  int32_t* pair_count, double* mse, double* t_final16, int32_t* n_it);

/* ------------------------------------------------------------------------------------------------
 * Hypothesis scoring of the RANSAC matchers (registration/ransacMatching/).  A hypothesis is the pair
 * (model index idx, scene index i) of the reference's trial loops; the kernels rebuild phi and T from
 * it exactly as TSD_PDFMatching.cpp:206-220 / RandomNormalMatching.cpp:251-263 / PDFMatching.cpp:235-250 do,
 * skip it when fabs(phi) >= phi_max, and score it.  m/s: n x 2 ray-model matrices, phi_m/phi_s: n.
 * control: 3 x n_control (homogeneous columns).  The winner is the first best in list order (the
 * reference's single-thread order: trial ascending, i ascending).
 * ---------------------------------------------------------------------------------------------- */
typedef struct tsd_hypothesis
{
  int32_t idx_model;
  int32_t idx_scene;
} tsd_hypothesis_t;

typedef struct tsd_matcher tsd_matcher_t;
int match_create(int device, tsd_matcher_t** out);
int match_destroy(tsd_matcher_t* m);

/* TSD_PDFMatching.cpp:213-260.  score[h] = product of likelihoods, or -1 when the hypothesis is skipped.
 * best = index of the first maximum (> 0), -1 when none; t_best 3x3 (identity when none). */
int match_score_tsd(tsd_matcher_t* m, tsd_grid_t* grid, int32_t n_hyp, const tsd_hypothesis_t* hyps, int32_t n,
                    const double* model, const double* scene, const double* phi_m, const double* phi_s,
                    double phi_max, int32_t n_control, const double* control, const double t_sensor[9],
                    double zrand, double* score, int32_t* best, double t_best[9]);

/* RandomNormalMatching.cpp:255-359.  model_valid: n_valid x 2 points of idxMValid with their orientations
 * phi_valid (phiM[idxMValid[k]]); phi_control: n_control.  Per hypothesis: cnt_match, max_cnt_match, err_sum
 * (cnt_match = -1 when skipped).  best follows the reference's ordered predicate (:344-359). */
int match_score_rnm(tsd_matcher_t* m, int32_t n_hyp, const tsd_hypothesis_t* hyps, int32_t n, const double* model,
                    const double* scene, const double* phi_m, const double* phi_s, double phi_max,
                    int32_t n_control, const double* control, const double* phi_control, int32_t n_valid,
                    const double* model_valid, const double* phi_valid, double theta_min, double theta_max,
                    double scale_distance, double scale_orientation, uint32_t cnt_match_thresh, int32_t* cnt_match,
                    int32_t* max_cnt_match, double* err_sum, int32_t* best, double t_best[9]);

/* PDFMatching.cpp:242-376 + probabilityOfTwoSingleScans (:435-487).  params: zhit zphi zshort zmax zrand
 * percentagePointsInC rangemax sigphi sighit lamshort maxAngleDiff maxAnglePenalty.  model_angles/model_dists:
 * n_valid (PDFMatching.cpp:200-204). */
int match_score_pdf(tsd_matcher_t* m, int32_t n_hyp, const tsd_hypothesis_t* hyps, int32_t n, const double* model,
                    const double* scene, const double* phi_m, const double* phi_s, double phi_max,
                    int32_t n_control, const double* control, int32_t n_valid, const double* model_angles,
                    const double* model_dists, const double params[12], double* prob, int32_t* fov_count,
                    int32_t* best, double t_best[9]);

/* ------------------------------------------------------------------------------------------------
 * Matcher pre-processing on the device: what TSD_PDFMatching.cpp:59-205 == RandomNormalMatching.cpp:94-247
 * == PDFMatching.cpp:67-233 do on the host before they score -- PCA normals and orientations of model and scene
 * (RandomMatching::calcNormals / calcPhi, RandomMatching.cpp:77-169), scene subsampling (:171-183), extractSamples (:41-50),
 * pickControlSet (:52-75), the trial order and the hypothesis list -- for 10^5-hypothesis relocalisation without host work
 * between the scan and the scores.  DELIBERATE departures from the reference, documented in DESIGN.md: random numbers come
 * from a counter-based generator, match_rng(seed, stream, index) (SplitMix64 finaliser), not from the sequential libc rand()
 * -- "keep scene point i" is match_rng(seed, 0, i) % 1000 >= threshold; "K of the valid indices without replacement" is
 * "the K smallest keys match_rng(seed, stream, index)": the trials (stream 2) in key order, i.e. in random order as in the
 * reference; the control set (stream 1) in scene order -- it is a set to every consumer, and neighbours help the scorers --
 * and pcaAnalysis' centroid is a double-precision running mean (the reference's is long double).
 * Every pointer of the result points into pinned host memory owned by the matcher, valid until the next match_prepare /
 * match_destroy on it; the device keeps a copy: handing these very pointers to match_score_tsd / _rnm / _pdf uploads
 * nothing.  n_hyp == 0: nothing to score (too few valid points, as RandomNormalMatching.cpp:160-170). */
typedef struct tsd_match_prep
{
  int32_t n, n_hyp, n_control, n_trials, n_valid_m, n_valid_s, span;
  double phi_max, theta_min, theta_max;
  const tsd_hypothesis_t* hyps;              /* n_hyp, in trial order, scene index ascending inside a trial */
  const double *model, *scene;               /* n x 2 (copies of the inputs) */
  const double *phi_m, *phi_s;               /* n; -1e6 where the mask is off */
  const uint8_t *mask_m_pca, *mask_s_pca;    /* n */
  const int32_t *idx_m_valid, *idx_s_valid;  /* n_valid_m, n_valid_s */
  const int32_t *idx_control, *idx_trials;   /* n_control, n_trials */
  const double *control, *phi_control;       /* 3 x n_control, n_control */
  const double *model_valid, *phi_valid;     /* n_valid_m x 2, n_valid_m (match_score_rnm) */
  const double *model_angles, *model_dists;  /* n_valid_m (match_score_pdf) */
} tsd_match_prep_t;
int match_prepare(tsd_matcher_t* m, int32_t n, const double* model, const uint8_t* mask_m, const double* scene,
                  const uint8_t* mask_s, int32_t pca_search_range, uint32_t size_control_set, uint32_t trials, double phi_max,
                  double resolution, uint64_t seed, tsd_match_prep_t* out);
uint64_t match_rng(uint64_t seed, uint32_t stream, uint32_t index);

/* ------------------------------------------------------------------------------------------------
 * One TsdGrid sharded over several devices INSIDE the library (one handle, one process; n_bands bands of whole
 * partition rows, band i on devices[i], or on device i mod #devices when devices is NULL; several bands may share a
 * device).  This is the handle behind obvious::TsdGrid(cellSize, layoutPartition, layoutGrid, nBands) in the adapter
 * (the node constructs ONE grid, SlamNode.cpp:77).  Pushes run on every band a scan reaches, concurrently, without
 * communication; the first read after pushes (ray cast, sampling, partition download) brings halo rows and allocation
 * flags up to date over peer memory; the ray cast is the collective of tsdg_raycast_mask_sharded.  Every result equals the
 * unsharded grid's bit for bit.  n_bands <= 16.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tsd_sharded tsd_sharded_t;
int tsdg_create_sharded(double cell_size, int layout_partition, int layout_grid, int n_bands, const int* devices,
                        tsd_sharded_t** out);
int tsdg_sharded_destroy(tsd_sharded_t* grid);
int tsdg_sharded_num_bands(const tsd_sharded_t* grid);
tsd_grid_t* tsdg_sharded_band(tsd_sharded_t* grid, int band);   /* a band's own handle (getters, per-band statistics) */
int tsdg_sharded_set_max_truncation(tsd_sharded_t* grid, double val);                      /* TsdGrid::setMaxTruncation */
int tsdg_sharded_free_footprint(tsd_sharded_t* grid, double cx, double cy, double w, double h);   /* TsdGrid::freeFootprint */
int tsdg_sharded_push(tsd_sharded_t* grid, const tsd_scan_t* scan);                        /* TsdGrid::push, blocking */
int tsdg_sharded_push_batch(tsd_sharded_t* grid, const tsd_scan_t* scans, int32_t n);
int tsdg_sharded_sync(tsd_sharded_t* grid);                  /* halos + flags now (otherwise done by the first read) */
int tsdg_sharded_last_push_stats(tsd_sharded_t* grid, tsd_push_stats_t* out);              /* summed over the bands */
int tsdg_sharded_raycast_mask(tsd_sharded_t* grid, const tsd_scan_t* scan, const double* rays_world, double* coords,
                              double* normals, uint8_t* mask, uint32_t* count);             /* calcCoordsFromCurrentViewMask */
int tsdg_sharded_interpolate_bilinear(tsd_sharded_t* grid, int32_t n, const double* xy, double* tsd, int32_t* status);
int tsdg_sharded_partition_states(tsd_sharded_t* grid, int32_t* state, double* init_weight);
int tsdg_sharded_download_partition(tsd_sharded_t* grid, int32_t p, double* tsd, double* weight);

#ifdef __cplusplus
}
#endif
#endif /* TSDSLAM_B200_H */
