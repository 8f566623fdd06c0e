/* m3t_synth.h — seeded, analytic generator of inputs for the pose-optimisation hot path
 * (SURVEY.md §7 step 2 / §8d): sparse viewpoint models in the reference's in-memory / .bin
 * DataPoint layout (region_model.h:89-110, depth_model.h:67-86), synthetic 640x480 BGR8 / U16
 * frames, ground-truth and perturbed start poses. No OpenGL: the body is a convex polytope (the
 * reference's own data/_body/triangle.obj prism), so silhouettes are convex hulls of projected
 * vertices and depth is an exact ray cast.
 *
 * This is a data tool, not part of the hot path and not the oracle; it has no CUDA dependency.
 */
#ifndef M3T_SYNTH_H_
#define M3T_SYNTH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct m3ts_intrinsics {
  float fu, fv, ppu, ppv;
  int32_t width, height;
} m3ts_intrinsics;

/* Number of geodesic views for n_divides: 10 * 4^n + 2 (model.cpp:386-454). */
int m3ts_n_views(int n_divides);

/* Region model: orientations[nv][3], contour_lengths[nv], points[nv][n_points] x 152 B
 * (center_f_body[3], normal_f_body[3], foreground_distance, background_distance, depth_offsets[30]).
 * Returns nv, or a negative value on error. */
int m3ts_generate_region_model(int n_divides, int n_points, float sphere_radius, uint64_t seed,
                               float* orientations, float* contour_lengths, void* points);
/* Depth model: points[nv][n_points] x 144 B (center_f_body[3], normal_f_body[3], depth_offsets[30]). */
int m3ts_generate_depth_model(int n_divides, int n_points, float sphere_radius, uint64_t seed,
                              float* orientations, float* surface_areas, void* points);

/* Ground-truth body2camera pose for body `index`: seeded rotation, translation that keeps the
 * whole body plus the longest correspondence lines (margin_px) inside the image at z in [z_min, z_max]. */
void m3ts_ground_truth_pose(uint64_t seed, int index, const m3ts_intrinsics* intr, float margin_px,
                            float z_min, float z_max, float body2camera[12]);
/* out = in * [R(axis, rot_deg) | trans_m * dir], axis/dir seeded. */
void m3ts_perturb_pose(uint64_t seed, int index, float rot_deg, float trans_m, const float in[12],
                       float out[12]);

/* Colour frame: foreground N(fg_mean, sigma) inside the silhouette at body2camera, background
 * N(bg_mean, sigma) elsewhere, per channel, clipped to u8. bgr rows are `pitch` bytes apart. */
void m3ts_render_color(const m3ts_intrinsics* intr, const float body2camera[12], uint64_t seed,
                       const uint8_t fg_mean[3], const uint8_t bg_mean[3], float sigma, uint8_t* bgr,
                       size_t pitch);
/* Depth frame: ray-cast prism, background plane at background_z, + N(0, noise_sigma) metres,
 * `invalid_fraction` of the pixels set to 0; stored as round(z / depth_scale). */
void m3ts_render_depth(const m3ts_intrinsics* intr, const float body2camera[12], uint64_t seed,
                       float background_z, float noise_sigma, float invalid_fraction, float depth_scale,
                       uint16_t* depth, size_t pitch);

#ifdef __cplusplus
}
#endif
#endif
