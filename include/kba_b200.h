/*
 * kba_b200.h -- C ABI of the B200-native keyframe bundle-adjustment hot path.
 *
 * This is the drop-in boundary for limo's `keyframe_bundle_adjustment` window solve.
 * The reference has no FFI: its "operator API" is the C++ class
 *   BundleAdjusterKeyframes            (keyframe_bundle_adjustment/include/keyframe_bundle_adjustment/bundle_adjuster_keyframes.hpp:40-335)
 * whose solve() / adjustPoseOnly() hand a ceres::Problem to
 *   robust_optimization::solveTrimmed  (robust_optimization/src/robust_solving.cpp:140-248).
 * Everything below the construction of that ceres::Problem -- residual/Jacobian evaluation
 * (cost_functors_ceres.hpp:53-222,224-250,355-438,507-555), robust losses, local parameterisations,
 * Schur elimination of the landmark blocks, the dense reduced solve, the Levenberg-Marquardt
 * loop and the quantile trimming -- is replaced by the entry points declared here.
 *
 * Plain C: pointers + sizes only, no torch / Eigen / ceres types.  All arrays are caller owned.
 * "host" entry points take host pointers and perform the host<->device copies themselves;
 * the kba_batch_* entry points keep a batch of windows resident in HBM.
 *
 * The library has NO CPU fallback: every entry point that computes returns KBA_ERR_CUDA if no
 * sm_100-class device is usable.
 */
#ifndef KBA_B200_H
#define KBA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KBA_VERSION_MAJOR 0
#define KBA_VERSION_MINOR 1

/* ---- status codes (reference: C++ exceptions / text report, bundle_adjuster_keyframes.cpp:630-632) ---- */
enum {
    KBA_OK = 0,
    KBA_ERR_BAD_ARG = 1,       /* null pointer, negative size, index out of range */
    KBA_ERR_CUDA = 2,          /* no device / CUDA runtime error (see kba_last_error) */
    KBA_ERR_NOT_ENOUGH_KF = 3, /* fewer than 3 keyframes: NotEnoughKeyframesException, cpp:630 */
    KBA_ERR_CAPACITY = 4,      /* window larger than the limits compiled into the kernels */
    KBA_ERR_NCCL = 5,
    KBA_ERR_TIMEOUT = 6        /* kba_result.status only: the host's safety cap ended the batch before this window finished */
};

/* ---- termination of one inner solve (mirrors ceres::TerminationType as used via Summary) ---- */
enum {
    KBA_TERM_CONVERGENCE = 0,    /* function / parameter / gradient tolerance or min radius */
    KBA_TERM_NO_CONVERGENCE = 1, /* max_num_iterations reached                            */
    KBA_TERM_FAILURE = 2         /* evaluation failed at the start (|z_cam| < 0.01) or 5 invalid steps */
};

/*
 * One optimisation window = what BundleAdjusterKeyframes::solve() (cpp:629-767) assembles
 * into a ceres::Problem.  Landmark-major CSR over the observations.
 *
 * Index conventions: keyframes 0..n_kf-1 are the ACTIVE keyframes in ascending id (= timestamp)
 * order (std::set iteration order, cpp:505); landmarks 0..n_lm-1 are the SELECTED landmarks in
 * ascending id order; observations of landmark j are obs[lm_obs_ptr[j] .. lm_obs_ptr[j+1]) sorted
 * by keyframe index, then camera index.
 */
typedef struct kba_window {
    int32_t n_kf, n_cam, n_lm, n_obs, n_gp;

    /* keyframes (keyframe.hpp:172-195) */
    const double* kf_pose;    /* [n_kf*7] quaternion (w,x,y,z) + translation; p_kf = R(q) p_origin + t (definitions.hpp:75-83) */
    const uint8_t* kf_fixed;  /* [n_kf]   1 = FixationStatus::Pose -> pose and plane blocks constant (cpp:198-219) */
    const double* kf_plane;   /* [n_kf*4] ground plane direction (3) + distance (definitions.hpp:27-34); may be NULL if n_gp == 0 */

    /* cameras (definitions.hpp:93-124) */
    const double* cam_intr;   /* [n_cam*3] focal length, principal point x, y */
    const double* cam_pose;   /* [n_cam*7] camera <- vehicle extrinsics, same 7-vector convention */

    /* landmarks (definitions.hpp:42-68) */
    const double* lm_pos;     /* [n_lm*3] position in the origin frame */
    const double* lm_weight;  /* [n_lm]   ScaledLoss weight (1.0, or shrubbery weight; cpp:589-591,616-618) */
    const int32_t* lm_obs_ptr;/* [n_lm+1] CSR row pointers */

    /* observations = FeaturePoint{u,v,d} (matches_msg_types/feature_point.hpp:24-26) */
    const int32_t* obs_kf;    /* [n_obs] keyframe index */
    const int32_t* obs_cam;   /* [n_obs] camera index, or NULL (all camera 0) */
    const float* obs_u;       /* [n_obs] */
    const float* obs_v;       /* [n_obs] */
    const float* obs_d;       /* [n_obs] lidar depth; a depth residual exists iff d > 0 (cpp:578) */

    /* ground-plane height residuals, already attached to their nearest keyframe and weighted by the
     * host exactly as addGroundPlaneResiduals does at problem-build time (cpp:517-562) */
    const int32_t* gp_lm;     /* [n_gp] landmark index */
    const int32_t* gp_kf;     /* [n_gp] keyframe index */
    const double* gp_weight;  /* [n_gp] ScaledLoss weight = 10 * (1 - dist/25) */

    /* scale regulariser PoseRegularization (cost_functors_ceres.hpp:224-250, cpp:890-904):
     * residual |(T_kf1 * T_kf0^-1).t| - scale_value, TrivialLoss scaled by scale_weight. */
    int32_t scale_kf0, scale_kf1;
    double scale_weight;      /* <= 0: no scale regulariser */
    double scale_value;

    /* ground-plane regularisation chain (cpp:769-818): weight > 0 adds, for consecutive keyframes,
     * normal difference (3w), distance difference (w), motion-in-plane (2w) and, for every keyframe,
     * the (0,0,1) normal prior (w).  The reference passes w = 10 and only if n_gp > 0 (cpp:717-719). */
    double plane_reg_weight;
    uint8_t plane_dist_fixed; /* 1: all plane distances constant (fewer than 10 depth residuals, cpp:722-728) */
    uint8_t landmarks_fixed;  /* 1: every landmark block constant = motion-only problem of adjustPoseOnly (cpp:862, 221-270) */
    uint8_t reserved_[6];

    /* SpeedRegularizationVector2 prior of adjustPoseOnly (cost_functors_ceres.hpp:300-353, cpp:835-853) on keyframe
     * speed_kf: residual(3) = (T_kf * speed_T_origin_before).t / speed_dt - speed_v_before, TrivialLoss * speed_weight. */
    int32_t speed_kf;         /* keyframe index the prior acts on */
    int32_t reserved2_;
    double speed_weight;      /* <= 0: none */
    double speed_dt;
    double speed_v_before[3];
    double speed_T_origin_before[7]; /* inverse of the newest window keyframe pose, frozen (7-vector convention) */
} kba_window;

/* Options = OutlierRejectionOptions (bundle_adjuster_keyframes.hpp:79-89) + the ceres / solveTrimmed
 * settings reachable from solve() (robust_solving.hpp:93-108, cpp:740-764). Fill with kba_default_options. */
typedef struct kba_options {
    double depth_thres;            /* 0.16  Cauchy scale of the depth residual            */
    double reprojection_thres;     /* 1.6   Cauchy scale of the reprojection residual     */
    double depth_quantile;         /* 0.95                                                 */
    double reprojection_quantile;  /* 0.95                                                 */
    double gp_quantile;            /* 1.0   (cpp:758)                                      */
    double gp_huber;               /* 0.1   Huber scale of the ground-plane residual (cpp:549) */
    int32_t num_trim_rounds;       /* entries of number_iterations (cpp:740-745); -1 = reference rule:
                                      num_iterations(1) rounds iff n_lm > min_landmarks_for_trimming */
    int32_t trim_solver_iterations;/* 2     max_num_iterations per trimming round (cpp:743)  */
    int32_t final_solver_iterations;/*100   max_num_iterations of the final solve (robust_solving.hpp:100) */
    int32_t min_landmarks_for_trimming; /* 100 for solve() (cpp:741), 30 for adjustPoseOnly (cpp:865) */
    int32_t min_residual_groups;   /* 30    (cpp:762, robust_solving.cpp:109)               */
    int32_t num_rounds_option;     /* outlier_rejection_options_.num_iterations, default 1  */
    double solver_time_sec;        /* max_solver_time_in_seconds of EVERY inner ceres::Solve (robust_solving.cpp:233-238), checked on
                                      the device between iterations: the solve ends NO_CONVERGENCE with its accepted iterate and
                                      solveTrimmed continues, so the final solve always runs.  <= 0: none; ignored by sharded
                                      solves (collective).  Parity runs use 20 s as the reference tests do
                                      (test/keyframe_bundle_adjustment.cpp:486) */
    /* ceres defaults, never overridden by the reference (robust_solving.hpp:101-103 are commented out) */
    double function_tolerance;     /* 1e-6  */
    double gradient_tolerance;     /* 1e-10 */
    double parameter_tolerance;    /* 1e-8  */
    double initial_trust_region_radius; /* 1e4 */
    double max_trust_region_radius;     /* 1e16 */
    double min_trust_region_radius;     /* 1e-32 */
    double min_relative_decrease;       /* 1e-3 */
    double min_lm_diagonal;             /* 1e-6 */
    double max_lm_diagonal;             /* 1e32 */
    int32_t max_consecutive_invalid_steps; /* 5 */
    int32_t precision;             /* 0: FP64 kernels; 1: FP32 Jacobian kernels with FP64 accumulation */
} kba_options;

/* One LM iteration record (subset of ceres::IterationSummary that the report prints). */
typedef struct kba_iteration {
    double cost;               /* cost at the END of the iteration (candidate cost if rejected) */
    double cost_change;
    double gradient_max_norm;
    double step_norm;
    double relative_decrease;
    double trust_region_radius;
    int32_t iteration;
    int32_t solve_index;       /* which inner ceres::Solve of solveTrimmed */
    int32_t step_is_valid, step_is_successful;
} kba_iteration;

#define KBA_MAX_SOLVES 8

/* Summary of one inner ceres::Solve call. */
typedef struct kba_solve_summary {
    double initial_cost, final_cost;
    int32_t num_iterations;    /* iterations attempted, excluding iteration 0 */
    int32_t num_successful_steps;
    int32_t termination;       /* KBA_TERM_* */
    int32_t num_landmarks;     /* landmark blocks in the program */
    int32_t num_residual_blocks;
    int32_t reserved_;
} kba_solve_summary;

/* Result of kba_solve_window = what the caller of solve() reads back from Keyframe::pose_,
 * Landmark::pos, Plane (mono_lidar.cpp:204,263-265,283) + robust_optimization::Summary. */
typedef struct kba_result {
    /* outputs, caller-allocated; any may be NULL */
    double* kf_pose;           /* [n_kf*7] */
    double* kf_plane;          /* [n_kf*4] */
    double* lm_pos;            /* [n_lm*3] */
    uint8_t* lm_rejected;      /* [n_lm] 1 if the trimming removed the landmark's residuals */
    kba_iteration* iterations; /* [iterations_capacity] optional per-iteration log */
    int32_t iterations_capacity;
    /* filled by the library */
    int32_t num_iteration_records;
    int32_t num_solves;
    int32_t status;            /* KBA_OK or error of this window */
    kba_solve_summary solves[KBA_MAX_SOLVES];
    double initial_cost;       /* Summary::initial_cost = first solve (robust_solving.hpp:71) */
    double final_cost;         /* Summary::final_cost  = last solve  (robust_solving.hpp:72) */
    double time_sec;           /* device time of the whole window solve (batch: of the whole batch) */
} kba_result;

/* What kba_eval materialises (parity / inspection entry point; evaluates at the window's input state). */
typedef struct kba_eval_out {
    double* residual;  /* [3*n_obs] rows (u, v, depth), robustified (sqrt(rho') applied); row 2 = 0 if no depth */
    double* jac_pose;  /* [18*n_obs] 3x6 row-major, d r~ / d (delta_rot, delta_trans); zeros for fixed keyframes */
    double* jac_lm;    /* [9*n_obs]  3x3 row-major, d r~ / d landmark */
    double* cost;      /* [1] 0.5 * sum rho over reprojection + depth blocks */
    int32_t* failed;   /* [1] 1 if any |z_cam| < 0.01 (evaluation failure, cost_functors_ceres.hpp:78-83) */
} kba_eval_out;

typedef struct kba_handle kba_handle;  /* one per host thread / GPU; not thread-safe */
typedef struct kba_batch kba_batch;    /* windows resident in HBM */

/* --- lifecycle --- */
int kba_version(void);                                   /* major*100 + minor */
const char* kba_last_error(void);                        /* message of the last failing call on this thread */
void kba_default_options(kba_options* opt);
int kba_create(kba_handle** out, int device);            /* replaces `new ceres::Problem` + thread pool */
void kba_destroy(kba_handle* h);
int kba_set_stream(kba_handle* h, void* cuda_stream);    /* cudaStream_t; NULL = default stream */

/* --- host-buffer entry points (the reference-facing calls) --- */
/* replaces robust_optimization::solveTrimmed(...) as called from solve() (cpp:765) */
int kba_solve_window(kba_handle* h, const kba_window* w, const kba_options* opt, kba_result* res);
/* many independent windows in one call ("BA windows/s") */
int kba_solve_batch(kba_handle* h, int32_t n_windows, const kba_window* w, const kba_options* opt, kba_result* res);
/* residuals + Jacobian blocks of the reprojection / depth residuals at the input state
 * (what ceres::Problem::Evaluate would return for those blocks, cf. robust_solving.cpp:44) */
int kba_eval(kba_handle* h, const kba_window* w, const kba_options* opt, kba_eval_out* out);

/* --- device-resident batch (inputs stay in HBM between solves) --- */
int kba_batch_create(kba_handle* h, int32_t n_windows, const kba_window* w, kba_batch** out);
int kba_batch_upload(kba_batch* b, int32_t n_windows, const kba_window* w); /* re-upload state, same shapes */
/* resets to the uploaded state and solves; returns when every window is done.  Issued as ONE CUDA graph launch (the
 * Levenberg-Marquardt pass is the body of a conditional WHILE node, the device decides when the batch is done) unless the
 * handle runs on the legacy default stream, which cannot be captured: give the handle a stream of its own (kba_create does,
 * kba_set_stream with a created stream keeps it).  INTEGRATION.md lists the switches (KBA_GRAPH, ...). */
int kba_batch_solve(kba_batch* b, const kba_options* opt);
int kba_batch_download(kba_batch* b, kba_result* res);
int kba_batch_transfer_bytes(kba_batch* b, int64_t* h2d_bytes, int64_t* d2h_bytes); /* of the last upload / download */
int kba_batch_jacobian_pass(kba_batch* b, const kba_options* opt, int32_t repeats, float* ms_out); /* residual/Jacobian kernel only */
void kba_batch_destroy(kba_batch* b);
/* counters for bench.py: kernels launched / device ms per kernel family since the last reset */
typedef struct kba_counters {
    int64_t launches_total;
    int64_t launches_jacobian, launches_prep, launches_schur, launches_solve, launches_backsub, launches_cost, launches_update, launches_trim;
    double ms_jacobian;  /* CUDA-event time of the residual/Jacobian kernel launches when timing is enabled */
    int64_t jacobian_obs;/* observations processed by those launches */
} kba_counters;
int kba_get_counters(kba_handle* h, kba_counters* out, int reset);
int kba_enable_kernel_timing(kba_handle* h, int on);

/* ---- ONE large window sharded over several GPUs by landmark blocks (BASELINE config 5) --------------------------------
 * Every rank (one process per GPU) holds ALL keyframes and a block of the landmarks with their observations
 * (limo_b200/parallel.py::shard_window shows the partition).  Per LM iteration the ranks exchange, with NCCL all-reduce
 * over NVLink, the reduced pose system [S | rhs] their landmarks contribute to, the per-keyframe J^T J blocks and the
 * cost / model-decrease scalars; the reduced solve and the LM controller then run replicated and bit-identically on
 * every rank (an NCCL all-reduce delivers the same bits everywhere).  Trimming quantiles are taken over all ranks'
 * landmarks.  There is no reference counterpart (the reference is single-process); north_star asks for it.
 * Restrictions: one window per batch, no ground-plane residuals (n_gp = 0), every free keyframe is in the program. */
typedef struct kba_shard_comm kba_shard_comm;
#define KBA_SHARD_ID_BYTES 128
int kba_shard_unique_id(void* id_out);  /* KBA_SHARD_ID_BYTES; rank 0 creates it, the host broadcasts it to all ranks */
int kba_shard_comm_create(kba_handle* h, int32_t rank, int32_t world, const void* id, kba_shard_comm** out); /* collective */
void kba_shard_comm_destroy(kba_shard_comm* c);
/* b holds this rank's shard; lm_begin = index of its first landmark in the whole window, lm_total = landmarks of the
 * whole window.  Afterwards kba_batch_solve is a collective call: every rank must make it. */
int kba_batch_set_shard(kba_batch* b, kba_shard_comm* comm, int32_t lm_begin, int32_t lm_total);

/* ---- persistent, device-resident sliding window ------------------------------------------------------------
 * The reference rebuilds the whole ceres::Problem for every solve() (bundle_adjuster_keyframes.cpp:635-637) and kba_solve_window
 * re-packs and re-uploads the whole window likewise.  A kba_track keeps what push() has seen on the device instead:
 *   - every pushed keyframe's pose, plane and measurements (landmark slot, camera, u, v, d) in one arena, uploaded ONCE at push;
 *   - landmark positions / weights by caller-assigned dense slot (the caller keeps LandmarkId -> slot), updated in place by solves.
 * kba_track_solve() then takes only the small per-solve lists (which keyframe slots are active, in ascending id order, with their
 * fixation; which landmark slots are selected, in ascending id order; the ground-plane attachments; the regulariser scalars),
 * gathers the window's CSR ON THE DEVICE (k_track_* in kba_pack.cu), packs and solves it exactly like kba_solve_window, writes
 * poses / planes / landmarks back into the store and returns them.  The window it builds is, array for array, the one the caller
 * would have passed to kba_solve_window, so the results are bit-identical (tests/test_track.py).
 * Capacities are fixed at creation; a window beyond `win_*` must go through kba_solve_window.  Not thread-safe (one handle). */
typedef struct kba_track kba_track;
typedef struct kba_track_caps {
    int32_t max_keyframes;     /* keyframe slots in the store (active or not)        */
    int32_t max_landmarks;     /* landmark slots                                      */
    int32_t max_measurements;  /* arena entries over all stored keyframes             */
    int32_t win_keyframes;     /* largest window: keyframes (<= 30: fused path)       */
    int32_t win_landmarks;     /*                 selected landmarks                  */
    int32_t win_observations;  /*                 observations                        */
    int32_t win_ground;        /*                 ground-plane residuals              */
} kba_track_caps;
int kba_track_create(kba_handle* h, const kba_track_caps* caps, int32_t n_cam, const double* cam_intr, const double* cam_pose,
                     kba_track** out);
void kba_track_destroy(kba_track* t);
/* push(): keyframe `kf_slot` (re-usable after kba_track_drop_keyframe) with pose, plane (4, may be NULL) and its measurements */
int kba_track_push_keyframe(kba_track* t, int32_t kf_slot, const double* pose7, const double* plane4, int32_t n_meas,
                            const int32_t* lm_slot, const int32_t* cam, const float* u, const float* v, const float* d);
int kba_track_drop_keyframe(kba_track* t, int32_t kf_slot);  /* its arena space is reclaimed by compaction when needed */
/* landmark state the host changes outside a solve: initial positions of new landmarks (push(), cpp:318-319), weights and
 * positions touched by the caller; any of pos / weight may be NULL */
int kba_track_set_landmarks(kba_track* t, int32_t n, const int32_t* lm_slot, const double* pos3, const double* weight);
int kba_track_set_keyframe_pose(kba_track* t, int32_t kf_slot, const double* pose7, const double* plane4);
/* the same for n keyframes with ONE copy (pose7s [n*7], plane4s [n*4] or NULL): what solve() sends for its active keyframes */
int kba_track_set_keyframe_poses(kba_track* t, int32_t n, const int32_t* kf_slot, const double* pose7s, const double* plane4s);
/* one solve() on the stored window; `sel` carries sizes, the scalar members and the ground-plane lists of kba_window (gp_lm =
 * index into lm_slot), its keyframe / landmark / observation arrays are ignored; scale_weight < 0 asks for the reference's own rule
 * (cpp:703-716: 1000, or 1000 / (depth + ground-plane residuals) beyond ten of them; plane_dist_fixed by cpp:722-728), evaluated
 * on the device from the gathered window so that the host need not visit a single observation.  Results: res->kf_pose / kf_plane [n_kf],
 * res->lm_pos / lm_rejected [n_lm] in selection order. */
int kba_track_solve(kba_track* t, int32_t n_kf, const int32_t* kf_slot, const uint8_t* kf_fixed, int32_t n_lm, const int32_t* lm_slot,
                    const kba_window* sel, const kba_options* opt, kba_result* res);
int kba_track_transfer_bytes(kba_track* t, int64_t* h2d_last_solve, int64_t* d2h_last_solve, int64_t* h2d_pushes_total);

/* ---- landmark initialisation of push() for a whole window ------------------------------------------
 * Replaces, for all landmarks of `w` at once, what BundleAdjusterKeyframes::push() does per new landmark on the host:
 * the first observation with a lidar depth (d >= 0) is back-projected (bundle_adjuster_keyframes.cpp:332-355); without
 * one, the point closest to all viewing rays is taken when there are at least two (cpp:125-159,358-382,
 * internal/triangulator.hpp:51-75); then the cheirality test of the landmark selector
 * (landmark_selection_scheme_cheirality.cpp:22-60).  Batch semantics: "first observation" is CSR order (keyframe, then
 * camera), where the incremental reference looks at the keyframe being pushed.  w->lm_pos is not read.
 * flags_out[j]: bit 0 = a position was computed, bit 1 = it lies in front of every observing camera. */
int kba_init_landmarks(kba_handle* h, const kba_window* w, double* lm_pos_out, uint8_t* flags_out, float* device_ms);

/* ---- lidar depth extraction (BASELINE config 4) --------------------------------------------------------------------
 * Replaces the un-vendored mono_lidar_depth::DepthEstimator call the limo front end makes per frame (install_repos.sh:9;
 * in-tree only its parameter file demo_keyframe_bundle_adjustment_meta/res/mono_lidar_fusion_parameters.yaml, whose
 * keys the fields below carry, defaults = the YAML values).  Per feature: lidar points projected into the pixel
 * rectangle around it -> depth-histogram segmentation (nearest local maximum) -> plane through the largest triangle
 * -> intersection with the view ray; -1 when any gate fails.  No reference code or tests exist for it:
 * the CPU restatement in oracle/ follows this specification, parity is "unpinned". */
typedef struct kba_lidar_options {
    int32_t image_width, image_height;  /* 1242 x 375 */
    double rect_width, rect_height;     /* pixelarea_search_witdh 6, pixelarea_search_height 9 (yaml:11-13) */
    double rect_offset_x, rect_offset_y;/* yaml:15-18 */
    double hist_bin_width;              /* histogram_segmentation_bin_witdh 0.3 m (yaml:58-61) */
    int32_t hist_min_count;             /* histogram_segmentation_min_pointcount 1 (yaml:63) */
    int32_t min_points;                 /* points needed for a plane: 3 */
    double depth_min, depth_max;        /* treshold_depth_min/max 0 / 100 (yaml:97-104) */
    double local_rel_tolerance;         /* treshold_depth_local_value 0.5, relative (yaml:106-114); < 0 disables */
    double triangle_crossnorm_min;      /* triangleplanar_crossnorm_treshold 0.1 (yaml:172-175) */
    double viewray_plane_min;           /* viewray_plane_orthoganality_treshold 0.1 (yaml:177-178) */
} kba_lidar_options;
void kba_lidar_default_options(kba_lidar_options* opt);
/* cloud: n_points x point_stride floats, xyz first (KITTI .bin layout x,y,z,intensity: apps/main_program/utility.h:28-39);
 * T_cam_lidar: 7-vector camera <- lidar; intr: f, cx, cy; features: n_features x 2 floats (u, v); depth_out: n_features floats. */
int kba_lidar_depth(kba_handle* h, const float* cloud, int32_t n_points, int32_t point_stride, const double* T_cam_lidar,
                    const double* intr, const float* features_uv, int32_t n_features, const kba_lidar_options* opt,
                    float* depth_out, float* device_ms);

#ifdef __cplusplus
}
#endif
#endif /* KBA_B200_H */
