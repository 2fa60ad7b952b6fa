/* TEST INFRASTRUCTURE ONLY -- part of oracle/ (see oracle/README.md).
 *
 * Plain-C restatement of the reference's hot-path algorithms (the "port" oracle).  It follows the
 * reference's own control flow and arithmetic order statement by statement (each function cites the
 * file:line it follows) and is PINNED against the reference itself: tests/test_oracle_pin.py compares it
 * bit for bit with oracle/_ref/libohm_ref.so (the reference's sources compiled unmodified) and with the
 * golden vectors under tests/golden/ that were generated from that library.
 *
 * The entry points mirror include/tsdslam_b200.h (prefix port_ instead of tsdg_/icp_/match_) so that the
 * same harness drives the port and the CUDA library.  Build: -O2 -ffp-contract=off, no -march.
 */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H
#include <stdint.h>
#include "../../include/tsdslam_b200.h" /* tsd_scan_t, tsd_hypothesis_t, status codes */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct port_grid port_grid_t;

port_grid_t* port_grid_create(double cell_size, int layout_partition, int layout_grid);
void port_grid_destroy(port_grid_t* g);
void port_grid_set_max_truncation(port_grid_t* g, double val);
void port_grid_get_geometry(const port_grid_t* g, int32_t* cells_x, int32_t* cells_y, int32_t* partition_size,
                            double* cell_size, double* min_x, double* max_x, double* min_y, double* max_y,
                            double* max_truncation);
int port_grid_free_footprint(port_grid_t* g, double cx, double cy, double width, double height);
void port_grid_push(port_grid_t* g, const tsd_scan_t* scan);
void port_grid_last_push_stats(port_grid_t* g, tsd_push_stats_t* out);
void port_grid_interpolate_bilinear(port_grid_t* g, int32_t n, const double* xy, double* tsd, int32_t* status);
void port_grid_interpolate_normal(port_grid_t* g, int32_t n, const double* xy, double* normals, int32_t* ok);
int32_t port_grid_num_partitions(const port_grid_t* g);
void port_grid_partition_states(port_grid_t* g, int32_t* state, double* init_weight);
int port_grid_download_partition(port_grid_t* g, int32_t p, double* tsd, double* weight);
int port_grid_upload_partition(port_grid_t* g, int32_t p, const double* tsd, const double* weight);
void port_grid_fill(port_grid_t* g, double tsd, double weight, int only_uninitialized);
void port_back_project(const tsd_scan_t* scan, int32_t n, const double* xy, int32_t* idx);
/* RayCastAxisAligned2D::calcCoords / TsdGrid::grid2ColorImage (map publication, ThreadGrid.cpp:84,125) */
void port_axis_map(port_grid_t* g, double* coords, double* normals, uint32_t* cnt, int8_t* occupied);
void port_color_image(port_grid_t* g, uint8_t* image, uint32_t width, uint32_t height);
/* TsdGrid::storeGrid / the file constructor (TsdGrid.cpp:548-607, :25-110): the reference's checkpoint format */
int port_grid_store(port_grid_t* g, const char* path);
port_grid_t* port_grid_load(const char* path);

int port_raycast_mask(port_grid_t* g, const tsd_scan_t* scan, const double* rays_world, double* coords,
                      double* normals, uint8_t* mask, uint32_t* count);
void port_raycast_steps(uint64_t* fine_steps, uint64_t* coarse_steps);
/* first event of every beam: key = 2*step + abort (UINT64_MAX: none), for the sharded-merge tests */
void port_raycast_keys(port_grid_t* g, const tsd_scan_t* scan, const double* rays_world, uint64_t* keys);

typedef struct port_icp port_icp_t;
port_icp_t* port_icp_create(uint32_t max_iterations, double dist_max, double dist_min, uint32_t dist_iterations,
                            const double bounds[4]);
void port_icp_destroy(port_icp_t* icp);
int port_icp_run(port_icp_t* icp, const double* model, const double* normals, int32_t n_model, const double* scene,
                 int32_t n_scene, const double pose[9], const double* t_init, double t_out[9], double* mse,
                 uint32_t* pairs, uint32_t* iterations, int32_t* state);
int port_icp_get_trace(port_icp_t* icp, int32_t max_it, int32_t cap, uint32_t* pair_model, uint32_t* pair_scene,
                       int32_t* pair_count, double* mse, double* t_final16, int32_t* n_it);

int port_match_score_tsd(port_grid_t* grid, int32_t n_hyp, const tsd_hypothesis_t* hyps, int32_t n,
                         const double* model, const double* scene, const double* phi_m, const double* phi_s,
                         double phi_max, int32_t n_control, const double* control, const double t_sensor[9],
                         double zrand, double* score, int32_t* best, double t_best[9]);
int port_match_score_rnm(int32_t n_hyp, const tsd_hypothesis_t* hyps, int32_t n, const double* model,
                         const double* scene, const double* phi_m, const double* phi_s, double phi_max,
                         int32_t n_control, const double* control, const double* phi_control, int32_t n_valid,
                         const double* model_valid, const double* phi_valid, double theta_min, double theta_max,
                         double scale_distance, double scale_orientation, uint32_t cnt_match_thresh,
                         int32_t* cnt_match, int32_t* max_cnt_match, double* err_sum, int32_t* best, double t_best[9]);
int port_match_score_pdf(int32_t n_hyp, const tsd_hypothesis_t* hyps, int32_t n, const double* model,
                         const double* scene, const double* phi_m, const double* phi_s, double phi_max,
                         int32_t n_control, const double* control, int32_t n_valid, const double* model_angles,
                         const double* model_dists, const double params[12], double* prob, int32_t* fov_count,
                         int32_t* best, double t_best[9]);

/* ---- matcher pre-processing + full match(), restated for pinning the scorers against the reference.
 * These draw from port_rand(), the same LCG as oracle/ref_capi.cpp. */
void port_seed(uint32_t seed);
int port_rand(void);

typedef struct port_match_prep
{
  int32_t n;
  int32_t n_control;
  int32_t n_valid_m;
  int32_t n_valid_s;
  int32_t n_hyp;
  int32_t span;
  double phi_max;
  double theta_min, theta_max;
  double* phi_m;        /* n */
  double* phi_s;        /* n */
  uint8_t* mask_m_pca;  /* n */
  uint8_t* mask_s_pca;  /* n */
  int32_t* idx_m_valid; /* n_valid_m */
  int32_t* idx_s_valid; /* n_valid_s */
  int32_t* idx_control; /* n_control */
  double* control;      /* 3 x n_control */
  double* phi_control;  /* n_control */
  tsd_hypothesis_t* hyps; /* n_hyp, canonical order */
} port_match_prep_t;

/* RandomMatching::{calcNormals,calcPhi,extractSamples,subsampleMask,pickControlSet} + the trial loop's
 * (idx, i) enumeration (TSD_PDFMatching.cpp:59-205 and the identical blocks of the other two matchers).
 * Returns NULL when the reference would return identity early. */
port_match_prep_t* port_match_prepare(int32_t n, const double* model, const uint8_t* mask_m, const double* scene,
                                      const uint8_t* mask_s, uint32_t trials, uint32_t size_control_set,
                                      double phi_max, double resolution);
void port_match_prep_free(port_match_prep_t* p);

void port_match_tsd(port_grid_t* grid, uint32_t trials, double eps_thresh, uint32_t size_control_set, double zrand,
                    const double t_sensor[9], int32_t n, const double* model, const uint8_t* mask_m,
                    const double* scene, const uint8_t* mask_s, double phi_max, double trans_max, double resolution,
                    double t_out[9]);
void port_match_rnm(uint32_t trials, double eps_thresh, uint32_t size_control_set, int32_t n, const double* model,
                    const uint8_t* mask_m, const double* scene, const uint8_t* mask_s, double phi_max,
                    double trans_max, double resolution, double t_out[9]);
void port_match_pdf(uint32_t trials, double eps_thresh, uint32_t size_control_set, const double params[12], int32_t n,
                    const double* model, const uint8_t* mask_m, const double* scene, const uint8_t* mask_s,
                    double phi_max, double trans_max, double resolution, double t_out[9]);

/* host-side sensor helpers that stay on the host in the product too */
void port_invert3x3(const double in[9], double out[9]);

#ifdef __cplusplus
}
#endif
#endif
