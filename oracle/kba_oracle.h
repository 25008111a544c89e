/*
 * kba_oracle.h -- CPU restatement ("oracle") of limo's keyframe_bundle_adjustment window solve.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under limo_b200/ may include, link or call this; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only as the
 * checker / CPU baseline, never as the thing shipped.
 *
 * Parity status: the residual formulas, losses, parameterisations, problem assembly and trimming follow the
 * reference sources cited at each function and are pinned against every known-answer test the reference
 * holds for this path (tests/test_oracle_reference_tests.py).  The Levenberg-Marquardt / Schur internals
 * follow ceres-solver 1.13.0 (the un-vendored dependency pinned by docker/src/Dockerfile:47) as restated in
 * SURVEY.md Appendix A.5-A.6; the reference holds NO golden vectors for the Ceres iterate sequence and Ceres
 * is not installable here, so that part is "parity unpinned" beyond the reference's geometric tests.
 *
 * It shares only the POD input/output structs of include/kba_b200.h with the product.
 */
#ifndef KBA_ORACLE_H
#define KBA_ORACLE_H

#include "kba_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* robust_optimization::solveTrimmed on the problem BundleAdjusterKeyframes::solve() builds
 * (bundle_adjuster_keyframes.cpp:629-767, robust_solving.cpp:140-248). num_threads <= 0: all cores. */
int kbo_solve_window(const kba_window* w, const kba_options* opt, kba_result* res, int num_threads);

/* ceres::Problem::Evaluate of the reprojection + depth blocks at the input state (robustified residuals
 * and local-parameterisation Jacobian blocks), same layout as kba_eval. */
int kbo_eval(const kba_window* w, const kba_options* opt, kba_eval_out* out);

void kbo_default_options(kba_options* opt);

/* Order of the parameter / function tolerance tests inside the trust-region loop: 0 (default) = ceres 1.13, the tests
 * look at the candidate and a firing test leaves x unchanged; 1 = the order SURVEY.md A.6 recalls for ceres <= 1.12, where
 * an acceptable candidate is applied before the solve ends.  Process-wide; used by scripts/ceres_order_sensitivity.py. */
void kbo_set_tolerance_order(int order);

/* ---- single-residual entry points, for the reference's known-answer tests and finite-difference checks ---- */
/* ReprojectionErrorWithQuaternions::operator() (cost_functors_ceres.hpp:91-155); returns 0 on failure (|z|<0.01). */
int kbo_reprojection(const double pose[7], const double cam_pose[7], const double intr[3], const double point[3],
                     double u, double v, double res[2], double jac_pose_local[12], double jac_point[6]);
/* LandmarkDepthError::operator() (cost_functors_ceres.hpp:193-212). */
void kbo_depth(const double pose[7], const double cam_pose[7], const double point[3], double d, double res[1],
               double jac_pose_local[6], double jac_point[3]);
/* GroundPlaneHeightRegularization::operator() (cost_functors_ceres.hpp:358-385); jac: pose(6 local), dir (3 local), dist(1), point(3). */
void kbo_gp_height(const double pose[7], const double dir[3], double dist, const double point[3], double res[1],
                   double jac_pose[6], double jac_dir[3], double jac_dist[1], double jac_point[3]);
/* GroundPlaneMotionRegularization::operator() (cost_functors_ceres.hpp:533-548). */
void kbo_gp_motion(const double pose0[7], const double pose1[7], const double dir0[3], double res[1],
                   double jac_pose0[6], double jac_pose1[6], double jac_dir0[3]);
/* PoseRegularization::operator() (cost_functors_ceres.hpp:229-242): |(T1*T0^-1).t| - scale. */
void kbo_scale_reg(const double pose1[7], const double pose0[7], double scale, double res[1], double jac_pose1[6],
                   double jac_pose0[6]);
/* SpeedRegularizationVector2::operator() (cost_functors_ceres.hpp:319-340). */
void kbo_speed_reg(const double pose[7], const double T_origin_before[7], double dt, const double v_before[3],
                   double res[3], double jac_pose[18]);
/* ceres::QuaternionParameterization x IdentityParameterization(3) Plus. */
void kbo_pose_plus(const double pose[7], const double delta[6], double out[7]);
/* FixScaleVectorPlus (local_parameterizations.hpp:146-162). */
void kbo_dir_plus(const double n[3], const double delta[3], double out[3]);
/* TrimmerQuantile::getOutliers (trimmer_quantile.hpp:40-63); ties broken by index; returns number rejected,
 * writes rejected[i] = 1. */
int kbo_trimmer_quantile(const double* values, int n, double quantile, unsigned char* rejected);
int kbo_trimmer_fix(const double* values, int n, double threshold, unsigned char* rejected); /* trimmer_fix.hpp:38-47 */
/* Triangulator::triangulate_rays (internal/triangulator.hpp:51-75): R_oc [n*9] row-major, t_oc [n*3], rays [n*3] (unit, camera frame). */
void kbo_triangulate_rays(int n, const double* R_oc, const double* t_oc, const double* rays, double out[3]);

/* push() landmark initialisation for a whole window: back-projection / ray triangulation / cheirality (batch form of
 * bundle_adjuster_keyframes.cpp:125-159,332-382 and landmark_selection_scheme_cheirality.cpp:22-60);
 * flags: bit 0 created, bit 1 in front of every observing camera. */
void kbo_init_landmarks(const kba_window* w, double* lm_pos_out, unsigned char* flags_out);

/* lidar depth extraction per the parameter file's specification (lidar_oracle.c; parity unpinned, see there) */
void kbo_lidar_default_options(kba_lidar_options* opt);
int kbo_lidar_depth(const float* cloud, int n_points, int stride, const double* T_cam_lidar, const double* intr,
                    const float* features_uv, int n_features, const kba_lidar_options* opt, float* depth_out);

#ifdef __cplusplus
}
#endif
#endif
