/* m3t_oracle.h — CPU oracle for the M3T pose-optimisation hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product path (3dobjecttracking_b200/, libm3t_b200.so)
 * includes, links or calls this. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * It is a dependency-free scalar float32 restatement of the reference's arithmetic
 * (DLR-RM/3DObjectTracking @ f0210618, M3T/src/{region_modality,depth_modality,color_histograms,
 * link,optimizer,region_model,depth_model,body}.cpp). The reference itself cannot be compiled in
 * this image (no Eigen / OpenCV C++ / GLFW headers), so the third-party pieces on the path are
 * restated from their published algorithms: Eigen >= 3.3.2 (M3T/CMakeLists.txt:29) LDLT<Lower>
 * with diagonal pivoting, MatrixBase::exp() (Pade 3/5/7 + scaling and squaring, Higham 2005),
 * Transform::inverse() for Affine (3x3 cofactor inverse) and Transform::rotation() (polar factor).
 *
 * Pinning status (see DESIGN.md "Oracle"): exact known answers of the reference's own tests are
 * reproduced for ColorHistograms (color_histograms_test.cpp:71-104) and the .bin model layout
 * (data/model_test/ .bin files); Jacobians are pinned by finite differences; the reference's modality /
 * optimizer golden matrices depend on an OpenGL-generated model that is not checked in, so they
 * are reproduced only approximately via a software-rasterised model (tests/golden/).
 */
#ifndef M3T_ORACLE_H_
#define M3T_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_SCHEDULE 8
#define ORC_N_DEPTH_OFFSETS 30
#define ORC_REGION_POINT_FLOATS 38 /* 152 B */
#define ORC_DEPTH_POINT_FLOATS 36  /* 144 B */

/* rotation_mode: how Transform3fA::rotation() (region_modality.cpp:1007, region_model.cpp:118)
 * is evaluated. 1 = faithful: polar factor of the linear block (what Eigen's SVD-based
 * rotation() returns for an Affine transform). 0 = linear block as is (what the CUDA path does;
 * differs by ~1e-7, used for bit-level comparisons). */
#define ORC_ROTATION_LINEAR 0
#define ORC_ROTATION_POLAR 1
/* exp_mode: 1 = Pade matrix exponential as Eigen's unsupported MatrixFunctions (link.cpp:224),
 * 0 = closed-form Rodrigues (what the CUDA path does). */
#define ORC_EXP_RODRIGUES 0
#define ORC_EXP_PADE 1

typedef struct orc_intrinsics {
  float fu, fv, ppu, ppv;
  int32_t width, height;
} orc_intrinsics;

/* Same field order as m3tb_region_params so that tests can fill both from one description. */
typedef struct orc_region_params {
  int32_t n_lines_max;
  int32_t use_adaptive_coverage;
  float reference_contour_length;
  float min_continuous_distance;
  int32_t function_length;
  int32_t distribution_length;
  float function_amplitude;
  float function_slope;
  float learning_rate;
  int32_t n_global_iterations;
  int32_t n_scales;
  int32_t scales[ORC_MAX_SCHEDULE];
  int32_t n_standard_deviations;
  float standard_deviations[ORC_MAX_SCHEDULE];
  int32_t n_histogram_bins;
  float learning_rate_f;
  float learning_rate_b;
  float unconsidered_line_length;
  float max_considered_line_length;
  int32_t measure_occlusions;
  float measured_depth_offset_radius;
  float measured_occlusion_radius;
  float measured_occlusion_threshold;
  int32_t n_unoccluded_iterations;
  int32_t min_n_unoccluded_lines;
} orc_region_params;

typedef struct orc_depth_params {
  int32_t n_points_max;
  int32_t use_adaptive_coverage;
  int32_t use_depth_scaling;
  float reference_surface_area;
  float stride_length;
  int32_t n_considered_distances;
  float considered_distances[ORC_MAX_SCHEDULE];
  int32_t n_standard_deviations;
  float standard_deviations[ORC_MAX_SCHEDULE];
  int32_t measure_occlusions;
  float measured_depth_offset_radius;
  float measured_occlusion_radius;
  float measured_occlusion_threshold;
  int32_t n_unoccluded_iterations;
  int32_t min_n_unoccluded_points;
} orc_depth_params;

typedef struct orc_region_line {
  int32_t model_index;
  int32_t valid;
  float center_f_body[3];
  float center_u, center_v;
  float normal_u, normal_v;
  float delta_r;
  float normal_component_to_scale;
  float distribution[12];
  float mean;
  float measured_variance;
} orc_region_line;

typedef struct orc_depth_point {
  int32_t model_index;
  int32_t valid;
  float center_f_body[3];
  float normal_f_body[3];
  float correspondence_center_f_camera[3];
} orc_depth_point;

typedef struct orc_model {
  int32_t n_views, n_points;
  const float* orientations;   /* [n_views][3] */
  const float* view_scalars;   /* contour_length (region) or surface_area (depth), [n_views] */
  const float* points;         /* [n_views][n_points][38 or 36] floats, .bin AoS */
  float stride_depth_offset, max_radius_depth_offset;
  float max_view_scalar;       /* max_contour_length_ / max_surface_area_ */
} orc_model;

typedef struct orc_color_frame {
  orc_intrinsics intrinsics;
  float world2camera[12];
  const uint8_t* bgr;
  size_t pitch;
} orc_color_frame;

typedef struct orc_depth_frame {
  orc_intrinsics intrinsics;
  float world2camera[12];
  const uint16_t* depth;
  size_t pitch;
  float depth_scale;
} orc_depth_frame;

/* One rigid body with its modalities, as the batch driver sees it. Null pointers switch a modality off. */
typedef struct orc_body {
  float body2world[12];           /* in/out */
  const orc_region_params* region;
  const orc_model* region_model;
  const orc_color_frame* color;
  const orc_depth_params* depth;
  const orc_model* depth_model;
  const orc_depth_frame* depth_frame;
  float* histogram_f;             /* n_bins^3, in/out */
  float* histogram_b;
  float tikhonov_rotation, tikhonov_translation;
  int32_t first_iteration;        /* set by orc_start_modality */
  /* scratch owned by the caller: capacity n_lines_max / n_points_max */
  orc_region_line* lines;
  orc_depth_point* points;
  int32_t n_lines, n_points;      /* out: processed model points */
  int32_t region_view, depth_view; /* out */
} orc_body;

void orc_region_params_default(orc_region_params* p);
void orc_depth_params_default(orc_depth_params* p);

/* -- small helpers exposed for unit tests -- */
void orc_pose_multiply(const float a[12], const float b[12], float out[12]);
void orc_pose_inverse(const float a[12], float out[12]);                 /* Affine inverse (cofactors) */
void orc_pose_rotation(const float a[12], int rotation_mode, float r[9]); /* Transform3fA::rotation() */
void orc_exp_skew(const float w[3], int exp_mode, float r[9]);           /* Vector2Skewsymmetric(w).exp() */
int orc_ldlt_solve(int n, const float* a_lower_rowmajor, const float* b, float* x); /* Eigen::LDLT<MatrixXf,Lower> */
void orc_function_lookup(const orc_region_params* p, float lookup_f[8], float lookup_b[8],
                         float* min_expected_variance);

/* -- ColorHistograms (color_histograms.cpp) -- */
void orc_hist_clear(int n_bins, float* memory_f, float* memory_b);
void orc_hist_add(int n_bins, float* memory, const uint8_t bgr[3]);            /* Add{Fore,Back}groundColor :60-70 */
void orc_hist_calculate(int n_bins, float learning_rate, const float* memory, float* histogram); /* :174-214 */
void orc_hist_get(int n_bins, const float* hist_f, const float* hist_b, const uint8_t bgr[3],
                  float* pf, float* pb);                                       /* GetProbabilities :94-102 */

/* -- GetClosestView (region_model.cpp:105-130, depth_model.cpp:81-106) -- */
int orc_closest_view(const orc_model* model, const float body2camera[12], int rotation_mode);

/* -- RegionModality -- */
/* StartModality / CalculateResults histogram collection: AddLinePixelColorsToTempHistograms
 * (region_modality.cpp:1025-1155) accumulating into memory_f / memory_b. */
void orc_region_add_line_pixels(const orc_region_params* p, const orc_model* model,
                                const orc_color_frame* color, const float body2world[12],
                                int rotation_mode, float* memory_f, float* memory_b);
/* CalculateCorrespondences (region_modality.cpp:390-465). Writes one record per processed model
 * point (valid flag set for survivors). Returns number of processed model points. */
int orc_region_correspondences(const orc_region_params* p, const orc_model* model,
                               const orc_color_frame* color, const orc_depth_frame* occlusion_depth,
                               const float* hist_f, const float* hist_b, const float body2world[12],
                               int iteration, int first_iteration, int corr_iteration,
                               int rotation_mode, orc_region_line* lines, int* view_index);
/* CalculateGradientAndHessian (region_modality.cpp:485-558). H is full symmetric 6x6. */
void orc_region_gradient_hessian(const orc_region_params* p, const orc_color_frame* color,
                                 const float body2world[12], const orc_region_line* lines, int n_lines,
                                 int corr_iteration, int opt_iteration, int rotation_mode,
                                 float g[6], float H[36]);

/* -- DepthModality -- */
int orc_depth_correspondences(const orc_depth_params* p, const orc_model* model,
                              const orc_depth_frame* frame, const float body2world[12],
                              int iteration, int first_iteration, int corr_iteration,
                              int rotation_mode, orc_depth_point* points, int* view_index);
void orc_depth_gradient_hessian(const orc_depth_params* p, const orc_depth_frame* frame,
                                const float body2world[12], const orc_depth_point* points, int n_points,
                                int corr_iteration, float g[6], float H[36]);

/* -- Optimizer::CalculateOptimization for one rigid body (root link, body2joint = I)
 * (optimizer.cpp:144-167, link.cpp:205-241). Returns 1 if the pose was updated, 0 if theta had NaN. */
int orc_optimize_rigid(const float g[6], const float H[36], float tikhonov_rotation,
                       float tikhonov_translation, int exp_mode, float body2world[12], float theta[6]);

/* -- batch drivers (Tracker::ExecuteTrackingStep etc.), OpenMP over bodies when n_threads > 1 -- */
void orc_start_modalities(orc_body* bodies, int n_bodies, int iteration, int rotation_mode, int n_threads);
/* phase_seconds[4] (may be NULL): correspondences / gradient+hessian / optimisation / results,
 * summed over threads, like examples/rbot_evaluator.cpp:354-414. */
void orc_tracking_step(orc_body* bodies, int n_bodies, int iteration, int corr_begin, int corr_end,
                       int n_update_iterations, int rotation_mode, int exp_mode, int n_threads,
                       double* phase_seconds);
void orc_calculate_results(orc_body* bodies, int n_bodies, int iteration, int rotation_mode, int n_threads);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
