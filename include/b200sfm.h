/* b200sfm.h -- C ABI of the B200-native global-SfM solver core.
 *
 * This is the drop-in boundary for the three numeric hot loops of
 * colmap/glomap's estimators.  The reference has no FFI layer: the seam is the
 * public surface of its estimator classes, and each entry point below replaces
 * the arithmetic behind one of them (paths relative to the reference root):
 *
 *   b200sfm_ba_*   <-  glomap::BundleAdjuster::Solve
 *                      glomap/estimators/bundle_adjustment.h:38-51, .cc:11-106
 *   b200sfm_gp_*   <-  glomap::GlobalPositioner::Solve
 *                      glomap/estimators/global_positioning.h:56-70, .cc:28-93
 *   b200sfm_ra_*   <-  glomap::RotationEstimator::EstimateRotations
 *                      glomap/estimators/global_rotation_averaging.h:77-87, .cc:40-85
 *
 * Conventions: plain pointers and sizes only (no C++/torch types), caller-owned
 * HOST buffers unless a parameter is documented as a device pointer, FP64
 * values, int32 indices (int64 CSR offsets), return 0 on success or a
 * b200sfm_status code; b200sfm_last_error() gives the message.  A context is
 * thread-compatible (one thread at a time), distinct contexts are independent
 * -- the same contract as the reference estimators.
 * There is no CPU fallback behind any of these calls.
 */
#ifndef B200SFM_H_
#define B200SFM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SFM_VERSION 100
/* doubles reserved per intrinsics block (colmap Camera::params, un-vendored) */
#define B200SFM_INTR_STRIDE 12
#define B200SFM_NCCL_ID_BYTES 128

typedef enum {
  B200SFM_OK = 0,
  B200SFM_ERR_INVALID_ARG = 1,
  B200SFM_ERR_CUDA = 2,
  B200SFM_ERR_NCCL = 3,
  B200SFM_ERR_EMPTY = 4,        /* reference returns false: no images / no tracks */
  B200SFM_ERR_UNSUPPORTED = 5,
  B200SFM_ERR_NUMERIC = 6       /* NaN encountered (reference: LOG(ERROR) + false) */
} b200sfm_status;

/* COLMAP camera model ids supported on the device (colmap/sensor/models.h). */
typedef enum {
  B200SFM_SIMPLE_PINHOLE = 0,   /* f, cx, cy */
  B200SFM_PINHOLE = 1,          /* fx, fy, cx, cy */
  B200SFM_SIMPLE_RADIAL = 2,    /* f, cx, cy, k */
  B200SFM_RADIAL = 3            /* f, cx, cy, k1, k2 */
} b200sfm_camera_model;

typedef enum {
  B200SFM_TERM_NONE = 0,
  B200SFM_TERM_FUNCTION_TOLERANCE = 1,
  B200SFM_TERM_PARAMETER_TOLERANCE = 2,
  B200SFM_TERM_GRADIENT_TOLERANCE = 3,
  B200SFM_TERM_MAX_ITERATIONS = 4,
  B200SFM_TERM_MIN_RADIUS = 5,
  B200SFM_TERM_INVALID_STEPS = 6
} b200sfm_termination;

typedef struct b200sfm_ctx b200sfm_ctx;

/* ---- context ------------------------------------------------------------ */
int b200sfm_version(void);
/* One context per process and GPU.  `device` is the CUDA ordinal
 * (reference: colmap::SetBestCudaDevice(gpu_indices[0]),
 * bundle_adjustment.cc:79-84). */
int b200sfm_create(int device, b200sfm_ctx** out);
/* Multi-GPU: one process per GPU; rank 0 obtains an id with
 * b200sfm_nccl_unique_id and the host distributes it (torch.distributed /
 * MPI / file).  Points (with all their observations) or edges are sharded
 * across ranks by the caller; camera-sized vectors are replicated and
 * all-reduced over NCCL inside the solver. */
int b200sfm_nccl_unique_id(void* out_id /* B200SFM_NCCL_ID_BYTES */);
int b200sfm_create_dist(int device, int rank, int world_size, const void* nccl_id, b200sfm_ctx** out);
void b200sfm_destroy(b200sfm_ctx* ctx);
const char* b200sfm_last_error(const b200sfm_ctx* ctx);
int b200sfm_rank(const b200sfm_ctx* ctx);
int b200sfm_world_size(const b200sfm_ctx* ctx);
/* the cudaStream_t every kernel of this context is launched on (for CUDA-event timing by the host) */
void* b200sfm_cuda_stream(const b200sfm_ctx* ctx);
/* kernels launched by this context so far */
int64_t b200sfm_kernel_launches(const b200sfm_ctx* ctx);

/* ---- statistics common to the LM-based solvers (BA, GP) ------------------ */
typedef struct {
  int32_t iterations;            /* LM iterations (successful + unsuccessful) == ceres summary.iterations - 1 */
  int32_t num_successful_steps;
  int32_t termination;           /* b200sfm_termination */
  int32_t usable;                /* summary.IsSolutionUsable() */
  double initial_cost;
  double final_cost;
  int64_t num_observations;      /* residual blocks actually used (this rank) */
  int64_t pcg_iterations;        /* total PCG iterations (mat-vecs) */
  int64_t kernel_launches;       /* kernels of this library launched by the call */
  double ms_total;               /* device time of the whole solve (CUDA events) */
  double ms_linearize;           /* accumulated time of the Jacobian+Schur kernel */
  int64_t n_linearize;
  double ms_matvec;              /* accumulated time of the implicit-Schur mat-vec kernel */
  int64_t n_matvec;
  double ms_h2d;                 /* host->device copies inside the call */
  double ms_d2h;
  int64_t h2d_bytes;
  int64_t d2h_bytes;
} b200sfm_lm_stats;

/* ---- (iii) bundle adjustment --------------------------------------------- */
/* Field-for-field mirror of BundleAdjusterOptions (bundle_adjustment.h:12-37)
 * + the inherited ceres::Solver::Options the reference sets
 * (optimization_base.h:18-23), + the PCG knobs of this implementation. */
typedef struct {
  int32_t optimize_rig_poses;        /* unknown cam_from_rig of the sensors marked with b200sfm_ba_problem_set_sensor_variable */
  int32_t optimize_rotations;        /* default 1 */
  int32_t optimize_translation;      /* default 1 */
  int32_t optimize_intrinsics;       /* default 1 in the reference; shared blocks, <= 12 variable parameters in total */
  int32_t optimize_principal_point;  /* default 0 */
  int32_t optimize_points;           /* default 1 */
  int32_t min_num_view_per_track;    /* default 3 */
  int32_t max_num_iterations;        /* default 200 */
  double thres_loss_function;        /* Huber threshold, default 1.0 px */
  double function_tolerance;         /* default 1e-5 */
  double gradient_tolerance;         /* Ceres default 1e-10 */
  double parameter_tolerance;        /* Ceres default 1e-8 */
  /* implementation knobs (no reference counterpart: the reference factors
   * the reduced camera system with CHOLMOD, bundle_adjustment.cc:94-96) */
  int32_t pcg_max_iterations;        /* default 500 */
  int32_t pcg_min_iterations;        /* default 0 */
  double pcg_rel_tolerance;          /* ||r_k|| <= tol * ||r_0||, default 1e-2 */
  int32_t preconditioner;            /* 0 = block-Jacobi on U, 1 = Schur-Jacobi (default) */
  int32_t profile_kernels;           /* 1: time linearize / mat-vec kernels with CUDA events */
  int32_t fixed_num_iterations;      /* >0: run exactly this many LM iterations (bench), ignore tolerances */
  int32_t design;                    /* data layout of the Schur passes: 0 = auto, 1 = v1 (stored 6x3 W blocks, atomics per
                                        observation), 2 = v2 (compact J rows in both orders, camera-order second pass).
                                        Identical arithmetic; auto picks v2 unless intrinsics are optimised. */
} b200sfm_ba_opts;

void b200sfm_ba_default_opts(b200sfm_ba_opts* opts);

/* One-shot solve with host buffers: uploads, solves, writes the results back
 * in place (the estimator mutates frames/tracks/cameras in place,
 * bundle_adjustment.cc:140-146).
 *   C cameras (frames with trivial rigs), P points (tracks), N observations,
 *   K intrinsics blocks.
 *   pt_obs_begin [P+1]  CSR by point over the observation arrays
 *   obs_cam      [N]    camera index of each observation
 *   obs_xy       [N][2] observed (distorted) pixel, Image::features
 *   cam_intr     [C]    intrinsics block of each camera
 *   intr_model   [K]    b200sfm_camera_model
 *   intr_params  [K][B200SFM_INTR_STRIDE]   in/out
 *   quat_xyzw    [C][4] cam_from_world rotation, Eigen coeffs order, in/out
 *   trans        [C][3] cam_from_world translation, in/out
 *   cam_const_mask [C]  bit0: rotation constant, bit1: translation constant.
 *                       The shim sets 3 on the first frame
 *                       (bundle_adjustment.cc:261-266).  May be NULL.
 *   points       [P][3] in/out
 * In a distributed context every rank passes its own shard of points and
 * observations and the same cameras/intrinsics. */
int b200sfm_ba_solve(b200sfm_ctx* ctx, const b200sfm_ba_opts* opts, int32_t C, int32_t P, int64_t N, int32_t K,
                     const int64_t* pt_obs_begin, const int32_t* obs_cam, const double* obs_xy,
                     const int32_t* cam_intr, const int32_t* intr_model, double* intr_params,
                     double* quat_xyzw, double* trans, const uint8_t* cam_const_mask, double* points,
                     b200sfm_lm_stats* stats);

/* Resident problem: upload once, solve many times (GlobalMapper re-solves the
 * same BundleAdjuster after flipping GetOptions().optimize_rotations,
 * controllers/global_mapper.cc:204-221). */
typedef struct b200sfm_ba_problem b200sfm_ba_problem;
int b200sfm_ba_problem_create(b200sfm_ctx* ctx, int32_t C, int32_t P, int64_t N, int32_t K,
                              const int64_t* pt_obs_begin, const int32_t* obs_cam, const double* obs_xy,
                              const int32_t* cam_intr, const int32_t* intr_model, const uint8_t* cam_const_mask,
                              int32_t min_num_view_per_track, b200sfm_ba_problem** out);
/* Known (constant) camera rigs -- the `!optimize_rig_poses` branch of
 * BundleAdjuster::AddPointToCameraConstraints (glomap/estimators/bundle_adjustment.cc:147-161,
 * colmap::RigReprojErrorConstantRigCostFunctor): the unknown pose blocks are the F FRAMES
 * (rig_from_world); every observation is made by an image = (frame, sensor) whose cam_from_rig is a
 * constant and whose intrinsics block belongs to the sensor:
 *     r = ImgFromCam(intr[sensor_intr[s]], R_cr[s] (R_f X + t_f) + t_cr[s]) - xy.
 * obs_frame[N] indexes the frames (state arrays quat/trans are [F]), obs_sensor[N] the S sensors
 * (S <= 65535; reference sensors carry the identity transform).  Everything else -- masks, state
 * calls, solve, filters -- is b200sfm_ba_problem_*; the angle filter's `cam_calibrated` is then [S]. */
int b200sfm_ba_problem_create_rig(b200sfm_ctx* ctx, int32_t F, int32_t P, int64_t N, int32_t K, int32_t S,
                                  const int64_t* pt_obs_begin, const int32_t* obs_frame, const uint16_t* obs_sensor,
                                  const double* obs_xy, const double* sensor_quat_xyzw /*[S][4] cam_from_rig*/,
                                  const double* sensor_trans /*[S][3]*/, const int32_t* sensor_intr /*[S]*/,
                                  const int32_t* intr_model, const uint8_t* frame_const_mask,
                                  int32_t min_num_view_per_track, b200sfm_ba_problem** out);
/* optimize_rig_poses (bundle_adjustment.cc:162-180,296-308, colmap::RigReprojErrorCostFunctor): mark the sensors whose
 * cam_from_rig is an UNKNOWN of the following solves (the reference: every non-reference camera sensor); takes effect
 * when b200sfm_ba_opts::optimize_rig_poses is set.  sensor_variable[S]: 1 = unknown.  The optimised poses are read
 * back with b200sfm_ba_problem_get_sensor_poses (quat_xyzw [S][4] / trans [S][3], either may be NULL). */
int b200sfm_ba_problem_set_sensor_variable(b200sfm_ba_problem* p, const uint8_t* sensor_variable);
int b200sfm_ba_problem_get_sensor_poses(b200sfm_ba_problem* p, double* sensor_quat_xyzw, double* sensor_trans);
int b200sfm_ba_problem_set_state(b200sfm_ba_problem* p, const double* intr_params, const double* quat_xyzw,
                                 const double* trans, const double* points);
int b200sfm_ba_problem_get_state(b200sfm_ba_problem* p, double* intr_params, double* quat_xyzw, double* trans,
                                 double* points);
/* device-side snapshot / restore of the state (benchmark loops) */
int b200sfm_ba_problem_save_state(b200sfm_ba_problem* p);
int b200sfm_ba_problem_restore_state(b200sfm_ba_problem* p);
int b200sfm_ba_problem_solve(b200sfm_ba_problem* p, const b200sfm_ba_opts* opts, b200sfm_lm_stats* stats);
/* robust cost 1/2 sum rho(|r|^2) of the current state (all ranks) */
int b200sfm_ba_problem_cost(b200sfm_ba_problem* p, const b200sfm_ba_opts* opts, double* cost);
/* Track filters on the resident problem: the mapper runs them between the
 * BA solves (controllers/global_mapper.cc:164-186,243-276,309-337) on the arrays the problem already
 * holds.  Reference: glomap/processors/track_filter.cc:7-52 (pixel reprojection), :54-90 (angle),
 * :92-127 (triangulation angle).  They evaluate the CURRENT state and return a keep-mask (1 = keep)
 * per observation / per track plus the reference's return value (number of tracks changed / removed);
 * the caller compacts Track::observations. */
int b200sfm_ba_problem_filter_reprojection(b200sfm_ba_problem* p, double max_reprojection_error, uint8_t* keep /*[N]*/,
                                           int64_t* num_tracks_changed);
int b200sfm_ba_problem_filter_angle(b200sfm_ba_problem* p, const double* bearings /*[N][3] features_undist, or NULL = resident*/,
                                    const uint8_t* cam_calibrated /*[C] or NULL*/, double max_angle_error_deg,
                                    uint8_t* keep /*[N]*/, int64_t* num_tracks_changed);
/* FilterTracksByReprojection with in_normalized_image = true (track_filter.cc:24-31) -- the variant the mapper
 * calls (controllers/global_mapper.cc:176-181,254-259,289-294): error = |X_c.xy / X_c.z - b.xy / (b.z + EPS)|
 * against the undistorted feature b (Image::features_undist), threshold 1e-2 by default (types.h:21). */
int b200sfm_ba_problem_filter_reprojection_normalized(b200sfm_ba_problem* p, const double* bearings /*[N][3] or NULL = resident*/,
                                                      double max_reprojection_error, uint8_t* keep /*[N]*/,
                                                      int64_t* num_tracks_changed);
/* The two per-element processors the mapper runs between the solvers, on the resident problem.
 * b200sfm_ba_problem_normalize -- glomap/processors/reconstruction_normalizer.cc:5-104 (NormalizeReconstruction): robust
 * p0..p1 percentile box and trimmed mean of the image centres (float coordinates, sorted per axis), similarity with
 * identity rotation X' = scale X + t applied to the frame poses, the cam_from_rig translations and the points of the
 * CURRENT state; returns the similarity (either pointer may be NULL).
 * b200sfm_ba_problem_undistort -- glomap/processors/image_undistorter.cc:7-53 (UndistortImages): unit bearing
 * CamFromImg(xy).homogeneous().normalized() of every observation from the current intrinsics, kept on the device;
 * bearings_out [N][3] may be NULL.  The two bearing-based filters below accept bearings == NULL and then use the resident
 * bearings (computing them first if needed), which saves the 24 N-byte upload per call. */
int b200sfm_ba_problem_normalize(b200sfm_ba_problem* p, int32_t fixed_scale, double extent, double p0, double p1,
                                 double* scale_out, double* translation_out /*[3]*/);
int b200sfm_ba_problem_undistort(b200sfm_ba_problem* p, double* bearings_out /*[N][3] or NULL*/);
int b200sfm_ba_problem_filter_triangulation_angle(b200sfm_ba_problem* p, double min_angle_deg, uint8_t* keep_track /*[P]*/,
                                                  int64_t* num_tracks_removed);
void b200sfm_ba_problem_free(b200sfm_ba_problem* p);

/* ---- track establishment ---------------------------------------------------------------------
 * TrackEngine::EstablishFullTracks (glomap/controllers/track_establishment.cc:5-17): the union-find over all inlier
 * matches of the valid image pairs (BlindConcatenation, :19-63) and the collection of the components into tracks with the
 * inconsistency rule (TrackCollection, :65-150) on the device.  Input: the M inlier matches as global feature ids
 * gid = image_id << 32 | feature_id (:48-53) with the pixel of both features (Image::features).  Result: tracks in
 * ascending track id (= smallest global id of the component, the reference's root rule :56-60), observations of a track
 * in ascending global id; a track with two features of ONE image further apart than thres_inconsistency pixels keeps its
 * id but loses its observations (:118-131) and is counted in num_discarded.  The order-dependent greedy selection
 * FindTracksForProblem (:153-234) stays on the host. */
typedef struct b200sfm_tracks b200sfm_tracks;
int b200sfm_tracks_establish(b200sfm_ctx* ctx, int64_t num_matches, const uint64_t* gid1, const uint64_t* gid2,
                             const double* xy1 /*[M][2]*/, const double* xy2 /*[M][2]*/, double thres_inconsistency,
                             b200sfm_tracks** out, int64_t* num_tracks, int64_t* num_observations, int64_t* num_discarded);
int b200sfm_tracks_get(b200sfm_tracks* t, uint64_t* track_ids /*[T]*/, int64_t* begin /*[T+1]*/, uint32_t* obs_image /*[n]*/,
                       uint32_t* obs_feature /*[n]*/);
void b200sfm_tracks_free(b200sfm_tracks* t);

/* ---- (ii) global positioning (BATA) ----------------------------------------- */
/* Mirror of GlobalPositionerOptions (global_positioning.h:9-54) + inherited
 * solver options (optimization_base.h:18-23) + PCG knobs.  Only the
 * ONLY_POINTS constraint type is implemented (the mapper enforces it,
 * controllers/global_mapper.cc:145-149); random initialisation
 * (generate_random_positions / points, seed) is done by the host shim. */
typedef struct {
  int32_t optimize_positions;        /* default 1 */
  int32_t optimize_points;           /* default 1 */
  int32_t optimize_scales;           /* default 1 */
  int32_t min_num_view_per_track;    /* default 3 */
  int32_t max_num_iterations;        /* default 100 */
  int32_t max_num_line_search_step_size_iterations;  /* Ceres default 20 (bounded problems) */
  double thres_loss_function;        /* Huber threshold, default 0.1 */
  double function_tolerance;         /* 1e-5 */
  double gradient_tolerance;         /* 1e-10 */
  double parameter_tolerance;        /* 1e-8 */
  int32_t pcg_max_iterations;        /* default 1000 */
  int32_t pcg_min_iterations;
  double pcg_rel_tolerance;          /* default 1e-2 */
  int32_t preconditioner;            /* 0 block-Jacobi, 1 Schur-Jacobi (default) */
  int32_t profile_kernels;
  int32_t fixed_num_iterations;
  int32_t reserved0;
} b200sfm_gp_opts;

void b200sfm_gp_default_opts(b200sfm_gp_opts* opts);

/* One-shot solve with host buffers (results written back in place):
 *   pt_obs_begin [P+1], obs_cam [N]  as for BA
 *   obs_dir      [N][3]  unit bearing rotated into the world frame,
 *                        R_cw^T * features_undist (global_positioning.cc:294-296)
 *   cam_calibrated [C]   1: Huber loss, 0: ScaledLoss(Huber, 0.5) (.cc:313-316); NULL = all calibrated
 *   cam_const_mask [C]   nonzero: centre held constant; may be NULL
 *   centers [C][3]       camera centres (the reference keeps them in
 *                        RigFromWorld().translation during the solve), in/out
 *   points  [P][3]       in/out
 *   scales  [N]          one per observation, in/out (initialise to 1, .cc:298);
 *                        lower bound 1e-5 (.cc:373); the first scale of rank 0 is constant (.cc:484-489) */
int b200sfm_gp_solve(b200sfm_ctx* ctx, const b200sfm_gp_opts* opts, int32_t C, int32_t P, int64_t N,
                     const int64_t* pt_obs_begin, const int32_t* obs_cam, const double* obs_dir,
                     const uint8_t* cam_calibrated, const uint8_t* cam_const_mask, double* centers, double* points,
                     double* scales, b200sfm_lm_stats* stats);

typedef struct b200sfm_gp_problem b200sfm_gp_problem;
int b200sfm_gp_problem_create(b200sfm_ctx* ctx, int32_t C, int32_t P, int64_t N, const int64_t* pt_obs_begin,
                              const int32_t* obs_cam, const double* obs_dir, const uint8_t* cam_calibrated,
                              const uint8_t* cam_const_mask, int32_t min_num_view_per_track, b200sfm_gp_problem** out);
/* Known rigs in global positioning -- RigBATAPairwiseDirectionError with the rig scale held at 1
 * (glomap/estimators/cost_function.h:49-82, global_positioning.cc:325-346,493-497):
 *     r = t_obs - s (X - c_frame + obs_offset),   obs_offset = R_cam_from_world^T t_cam_from_rig.
 * obs_cam then indexes FRAMES (c = rig centre).  obs_calibrated[N] (optional) is the prior-focal flag
 * of the observing camera and replaces the per-frame cam_calibrated (the loss is chosen per camera,
 * .cc:313-316).  NULL/NULL restores the trivial-frame behaviour. */
int b200sfm_gp_problem_set_rig_terms(b200sfm_gp_problem* p, const double* obs_offset /*[N][3]*/,
                                     const uint8_t* obs_calibrated /*[N] or NULL*/);
/* Unknown cam_from_rig in global positioning -- RigUnknownBATAPairwiseDirectionError (glomap/estimators/cost_function.h:
 * 90-136, call site global_positioning.cc:347-364): for the images of a sensor whose cam_from_rig translation is not
 * known yet,  r = t_obs - s (X - c_frame - R_rw^T u_s)  with u_s (3 doubles, "cam_from_rig_center") an unknown shared by
 * all images of the sensor.  obs_unknown_sensor[N]: index in [0, S_u) or -1 for observations of reference / calibrated
 * sensors; frame_rot[C][9]: rig_from_world rotations, row-major; centers[S_u][3]: initial values (the reference draws
 * U(-1,1)^3, global_positioning.cc:440-453) -- read back with b200sfm_gp_problem_get_rig_unknown. */
int b200sfm_gp_problem_set_rig_unknown(b200sfm_gp_problem* p, int32_t num_unknown_sensors, const int32_t* obs_unknown_sensor,
                                       const double* frame_rot, const double* centers);
int b200sfm_gp_problem_get_rig_unknown(b200sfm_gp_problem* p, double* centers);
int b200sfm_gp_problem_set_state(b200sfm_gp_problem* p, const double* centers, const double* points, const double* scales);
int b200sfm_gp_problem_get_state(b200sfm_gp_problem* p, double* centers, double* points, double* scales);
int b200sfm_gp_problem_save_state(b200sfm_gp_problem* p);
int b200sfm_gp_problem_restore_state(b200sfm_gp_problem* p);
int b200sfm_gp_problem_solve(b200sfm_gp_problem* p, const b200sfm_gp_opts* opts, b200sfm_lm_stats* stats);
void b200sfm_gp_problem_free(b200sfm_gp_problem* p);

/* ---- (i) rotation averaging --------------------------------------------------- */
/* Mirror of RotationEstimatorOptions (global_rotation_averaging.h:39-75) for
 * 3-DoF frames with trivial rigs.  skip_initialization / use_gravity / axis are
 * host-side concerns (the maximum-spanning-tree initialisation, math/tree.cc:78,
 * stays on the host: O(E log E)).  The l1_* fields are
 * the colmap::LeastAbsoluteDeviationSolver options the reference uses
 * (global_rotation_averaging.cc:483-486 + COLMAP defaults). */
typedef struct {
  int32_t max_num_l1_iterations;           /* 5 */
  int32_t max_num_irls_iterations;         /* 100 */
  int32_t weight_type;                     /* 0 GEMAN_MCCLURE (default), 1 HALF_NORM */
  int32_t use_weight;                      /* 0 */
  double l1_step_convergence_threshold;    /* 1e-3 */
  double irls_step_convergence_threshold;  /* 1e-3 */
  double irls_loss_parameter_sigma;        /* 5 degrees */
  int32_t l1_max_admm_iterations;          /* 10 (.cc:484) */
  int32_t reserved0;
  double l1_rho;                           /* 1.0 */
  double l1_absolute_tolerance;            /* 1e-4 */
  double l1_relative_tolerance;            /* 1e-2 */
  /* PCG replaces the reference's CHOLMOD factorisations (.cc:491,547-611) */
  int32_t pcg_max_iterations;              /* 5000 */
  int32_t reserved1;
  double pcg_rel_tolerance;                /* 1e-8 */
} b200sfm_ra_opts;

typedef struct {
  int32_t l1_iterations;
  int32_t irls_iterations;
  int32_t admm_iterations;
  int32_t usable;                 /* 0: NaN encountered -> EstimateRotations returns false */
  int64_t num_edges;
  int64_t pcg_iterations;         /* Laplacian mat-vecs */
  int64_t kernel_launches;
  double ms_total;
} b200sfm_ra_stats;

void b200sfm_ra_default_opts(b200sfm_ra_opts* opts);

/*   n_frames          registered frames (unknown blocks of 3)
 *   ei, ej [E]        image/frame indices of each valid pair (image_id1, image_id2)
 *   R_rel  [E][9]     row-major cam2_from_cam1 rotation (ImagePairTempInfo::R_rel, .h:29)
 *   edge_w [E]        ImagePair::weight (used when use_weight; <0 -> 1); may be NULL
 *   fixed_frame       gauge frame (fixed_camera_id_, .cc:248-256)
 *   theta  [n][3]     angle-axis of every frame: in = initial estimate, out = result.
 * In a distributed context every rank passes its own shard of edges and the
 * same theta. */
int b200sfm_ra_solve(b200sfm_ctx* ctx, const b200sfm_ra_opts* opts, int32_t n_frames, int64_t n_edges,
                     const int32_t* ei, const int32_t* ej, const double* R_rel, const double* edge_w,
                     int32_t fixed_frame, double* theta, b200sfm_ra_stats* stats);

/* use_gravity variant (global_rotation_averaging.cc:207-217,311-340,386-421): frames flagged in
 * frame_has_gravity carry ONE unknown, the angle about the gravity axis, passed as theta = (0, phi, 0)
 * with phi = RotUpToAngle(R_align^T R) (.cc:208-210); R_rel must already be gravity-aligned by the host
 * (R_align2^T R_rel R_align1, .cc:311-326); fixed_frame must be the first frame with gravity if any
 * (.cc:213-217).  Pairs of two gravity frames contribute one row; the rand() jitter of RelAngleError
 * near +-pi (.cc:28-33) is not reproduced.  frame_has_gravity == NULL is b200sfm_ra_solve. */
int b200sfm_ra_solve_gravity(b200sfm_ctx* ctx, const b200sfm_ra_opts* opts, int32_t n_frames, int64_t n_edges,
                             const int32_t* ei, const int32_t* ej, const double* R_rel, const double* edge_w,
                             const uint8_t* frame_has_gravity, int32_t fixed_frame, double* theta,
                             b200sfm_ra_stats* stats);

/* Rotation averaging with UNKNOWN cam_from_rig rotations (glomap/estimators/global_rotation_averaging.cc:173-245 the
 * unknown layout, :425-440 the extra -I / +I blocks, :646-693 the update with colmap::AverageQuaternions over the frames
 * of each camera, :726-736 residuals with R_k = R_cam R_frame, :805-813 the estimated rotations).
 *   theta [n_frames + n_cams][3]: frame rotations followed by the cam_from_rig rotations of the sensors that are not
 *   calibrated (initial values in, result out);  eci/ecj [E]: node index (>= n_frames) of the camera of image 1 / 2 or -1;
 *   R_rel carries the calibrated cam_from_rig factors (:305-309);  cam_frames (CSR over the n_cams cameras): the frames that
 *   hold an image of the camera.  A pair inside one frame is kept when a camera of it is unknown (:300-304). */
int b200sfm_ra_solve_rig(b200sfm_ctx* ctx, const b200sfm_ra_opts* opts, int32_t n_frames, int32_t n_cams, int64_t n_edges,
                         const int32_t* ei, const int32_t* ej, const int32_t* eci, const int32_t* ecj, const double* R_rel,
                         const double* edge_w, const int32_t* cam_frames_begin, const int32_t* cam_frames,
                         int32_t fixed_frame, double* theta, b200sfm_ra_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* B200SFM_H_ */
