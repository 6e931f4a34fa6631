/*
 * gf2_abi.h — C ABI of the B200-native sliding-window solver and KLT tracker that replace
 * the arithmetic of Ground-Fusion++'s Estimator::optimization() and
 * FeatureTracker::trackImage().
 *
 * Citations use the abbreviations of SURVEY.md:
 *   VE/  = Ground-Fusion++/vins_estimator/src/      LIO/ = Ground-Fusion++/lio/src/
 *
 * What this boundary replaces (the reference has no FFI; its "plug-in" interface is the
 * ceres::CostFunction::Evaluate API driven by ceres::Solve):
 *   - gf2_solver_*  : the ceres::Problem build + ceres::Solve call of
 *                     Estimator::optimization()            VE/estimator/estimator.cpp:2951-3392
 *   - gf2_set_*     : the wire layout of vector2double()   VE/estimator/estimator.cpp:2337-2414
 *   - gf2_get_*     : the inverse, double2vector()'s input VE/estimator/estimator.cpp:2501-2630
 *   - gf2_imu_preintegrate / gf2_wheel_preintegrate :
 *                     IntegrationBase::push_back chain     VE/factor/integration_base.h:39-167
 *                     WheelIntegrationBase::push_back      VE/factor/wheel_integration_base.h:41-178
 *   - gf2_marginalize: MarginalizationInfo::preMarginalize/marginalize
 *                                                          VE/factor/marginalization_factor.cpp:119-308
 *   - gf2_tracker_* : cv::calcOpticalFlowPyrLK call sites  VE/featureTracker/feature_tracker.cpp:122,132,135,141
 *
 * Conventions: plain C, opaque handles, caller-owned HOST buffers, int status (0 = OK,
 * negative = error), no exceptions cross the boundary, gf2_last_error() returns the message
 * of the last failure on the calling thread. Calls are thread-safe per handle, not across
 * handles sharing one. Every quaternion is stored [x y z w] (Eigen coeffs order, the order of
 * para_Pose[i][3..6], VE/estimator/estimator.cpp:2345-2348). Every matrix is row-major unless
 * stated. All batched arrays are window-major: element (w, ...) lives at w * stride + ....
 *
 * There is NO CPU fallback behind this ABI: every entry point that computes runs CUDA kernels
 * on the handle's device and fails with GF2_ERR_CUDA if no device is usable.
 */
#ifndef GF2_ABI_H_
#define GF2_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GF2_ABI_VERSION 1

/* WINDOW_SIZE + 1 and NUM_OF_F of VE/estimator/parameters.h:24-25 */
#define GF2_MAX_FRAMES 11
#define GF2_MAX_LANDMARKS 1000
#define GF2_MAX_PRIOR_DIM 96

enum {
  GF2_OK = 0,
  GF2_ERR_INVALID = -1,     /* bad argument / out of capacity */
  GF2_ERR_CUDA = -2,        /* CUDA runtime failure or no device */
  GF2_ERR_UNSUPPORTED = -3, /* feature named by the ABI but not built yet */
  GF2_ERR_NCCL = -4
};

/* ---------------------------------------------------------------- records */

/* One feature observation. The reference stores cv::Point2f values widened to double
 * (VE/featureTracker/feature_tracker.cpp:326-339; FeaturePerFrame VE/estimator/feature_manager.h:33-43):
 * x,y = undistorted normalised image point (z = 1 implied), vx,vy = its velocity. fp32 is lossless. */
typedef struct gf2_obs {
  float x, y, vx, vy;
} gf2_obs;

/* One raw IMU sample fed to IntegrationBase::push_back (VE/factor/integration_base.h:39). */
typedef struct gf2_imu_sample {
  double dt;
  double acc[3];
  double gyr[3];
} gf2_imu_sample;

/* State of an IntegrationBase after its last push_back (members at
 * VE/factor/integration_base.h:197-217). jacobian/covariance are 15x15 row-major in the order
 * O_P=0, O_R=3, O_V=6, O_BA=9, O_BG=12 (VE/estimator/parameters.h enum StateOrder). */
typedef struct gf2_imu_preint {
  double sum_dt;
  double delta_p[3];
  double delta_q[4]; /* x y z w */
  double delta_v[3];
  double lin_ba[3];
  double lin_bg[3];
  double jacobian[225];
  double covariance[225];
  int32_t valid; /* 0: factor skipped (the reference skips sum_dt > 10, estimator.cpp:3175) */
  int32_t pad_;
} gf2_imu_preint;

/* One raw wheel-odometry sample fed to WheelIntegrationBase::push_back
 * (VE/factor/wheel_integration_base.h:41). */
typedef struct gf2_wheel_sample {
  double dt;
  double vel[3];
  double gyr[3];
} gf2_wheel_sample;

/* State of a WheelIntegrationBase (members at VE/factor/wheel_integration_base.h:221-244).
 * jacobian is 6x3 row-major (rows O_P=0..2, O_R=3..5; cols sx, sy, sw); covariance 6x6. */
typedef struct gf2_wheel_preint {
  double sum_dt;
  double delta_p[3];
  double delta_q[4]; /* x y z w */
  double lin_sx, lin_sy, lin_sw, lin_td;
  double lin_vel[3], lin_gyr[3]; /* linearized_vel / linearized_gyr: first sample of the interval */
  double vel_1[3], gyr_1[3];     /* last sample of the interval */
  double jacobian[18];
  double covariance[36];
  int32_t valid;
  int32_t pad_;
} gf2_wheel_preint;

/* One LiDAR point-to-plane factor attached to window poses (synthetic composition of BASELINE.json config 4).
 * ct == 0: LidarPlaneNormFactor (LIO/liw/lidarFactor.cpp:13-50) on the pose of `frame`:
 *          residual = sqrt_info * weight * (normal . (R p_body + t) + offset).
 * ct == 1: CTLidarPlaneNormFactor (LIO/liw/lidarFactor.cpp:52-123, the factor of `icpmodel: CT_POINT_TO_PLANE`, every LIO config):
 *          begin pose = window pose `frame`, end pose = window pose `frame + 1`, the point is transformed by the pose
 *          interpolated at alpha_time (slerp / lerp); alpha_time comes from gf2_set_plane_alpha. The reference's rotation
 *          Jacobians of this factor are first-order in the begin-end rotation difference; they are reproduced as written.
 * p_body is expressed in the body frame (q_il / t_il of the factor's constructor already applied). */
typedef struct gf2_plane {
  double p_body[3];
  double normal[3];
  double offset;
  double weight;
  int32_t frame;
  int32_t ct;
} gf2_plane;

/* Parameter-block kinds a marginalization prior can keep
 * (last_marginalization_parameter_blocks, VE/estimator/estimator.cpp:3561-3595). */
enum {
  GF2_BLK_POSE = 0,      /* para_Pose[index]        size 7 / local 6 */
  GF2_BLK_SPEEDBIAS = 1, /* para_SpeedBias[index]   size 9 */
  GF2_BLK_EX_POSE = 2,   /* para_Ex_Pose[0]         size 7 / local 6 */
  GF2_BLK_TD = 3,        /* para_Td[0]              size 1 */
  GF2_BLK_EX_WHEEL = 4,  /* para_Ex_Pose_wheel[0]   size 7 / local 6 */
  GF2_BLK_SX = 5,
  GF2_BLK_SY = 6,
  GF2_BLK_SW = 7,
  GF2_BLK_TD_WHEEL = 8
};

/* One kept block of a MarginalizationInfo: keep_block_size / keep_block_idx / keep_block_data
 * (VE/factor/marginalization_factor.h:74-76). `offset` = keep_block_idx - m. */
typedef struct gf2_prior_block {
  int32_t kind;
  int32_t index;  /* frame index for POSE / SPEEDBIAS, else 0 */
  int32_t offset; /* first column of this block in linearized_jacobians */
  int32_t pad_;
  double x0[9]; /* linearisation point, global size (7, 9 or 1 used) */
} gf2_prior_block;

/* Bits of gf2_solve_opts.const_mask: blocks held constant by SetParameterBlockConstant
 * (VE/estimator/estimator.cpp:3051-3060, 3093-3117, 3158-3161). Frame poses and speed-biases are
 * always free (the "stationary" freeze at :3294-3307 makes the solve a no-op; callers skip it). */
enum {
  GF2_CONST_EX_POSE = 1,
  GF2_CONST_TD = 2,
  GF2_CONST_EX_WHEEL = 4,
  GF2_CONST_WHEEL_INTRINSIC = 8, /* sx, sy, sw together, as the reference does */
  GF2_CONST_TD_WHEEL = 16
};

typedef struct gf2_solver gf2_solver;

typedef struct gf2_solver_cfg {
  int32_t device;        /* CUDA device ordinal */
  int32_t max_windows;   /* batch capacity B */
  int32_t n_frames;      /* frame_count + 1, <= GF2_MAX_FRAMES */
  int32_t max_landmarks; /* per window, <= GF2_MAX_LANDMARKS */
  int32_t max_obs;       /* per window: sum over landmarks of track length (host obs included) */
  int32_t max_planes;    /* per window LiDAR plane factors (0 = none) */
  int32_t max_imu_samples;   /* per interval, for gf2_imu_preintegrate (0 = records only) */
  int32_t max_wheel_samples; /* per interval */
  int32_t use_wheel;     /* allocate wheel factor storage */
  int32_t max_prior_rows; /* row stride P of the prior arrays of gf2_set_prior; 0 selects GF2_MAX_PRIOR_DIM */
  int32_t sweep;         /* GF2_SWEEP_*: which Jacobian-sweep kernel gf2_solve / gf2_linearize launch (same results; 0 = by batch size) */
  int32_t reserved_[5];
} gf2_solver_cfg;

enum { /* gf2_solver_cfg.sweep */
  GF2_SWEEP_AUTO = 0,   /* calls of up to one window per SM take the window kernel, larger batches the batch kernel */
  GF2_SWEEP_BATCH = 1,  /* k_linearize: 4 warps per window, two windows per SM (throughput of a full batch) */
  GF2_SWEEP_WINDOW = 2  /* k_linearize_ws: one window per SM, 16 warp-specialised warps (latency of one robot's window) */
};

/* Options of one solve = the ceres::Solver::Options the reference sets (estimator.cpp:3364-3376)
 * plus the static members it configures elsewhere. */
typedef struct gf2_solve_opts {
  int32_t max_iterations;  /* NUM_ITERATIONS; Ceres counts attempted steps */
  uint32_t const_mask;     /* GF2_CONST_* */
  double huber_delta;      /* HuberLoss(1.0), estimator.cpp:2959 */
  double sqrt_info_px;     /* ProjectionTwoFrameOneCamFactor::sqrt_info = FOCAL_LENGTH/1.5 * I (estimator.cpp:193) */
  double g_norm;           /* G = (0,0,g_norm), parameters.cpp:223 */
  double lidar_sqrt_info;  /* LidarPlaneNormFactor::sqrt_info = sqrt(1/laser_point_cov) (lidarodom.cpp:10,13) */
  double max_time_s;       /* options.max_solver_time_in_seconds (estimator.cpp:3373-3376): 0 = no cap; > 0: checked after every iteration against the
                            * device clock since the start of the solve (termination NO_CONVERGENCE). Machine dependent, like the reference's. */
  /* Ceres 1.14 trust-region defaults; 0 selects the default */
  double initial_radius;       /* 1e4 */
  double function_tolerance;   /* 1e-6 */
  double gradient_tolerance;   /* 1e-10 */
  double parameter_tolerance;  /* 1e-8 */
  /* PoseSubsetParameterization of para_Ex_Pose_wheel (estimator.cpp:3069-3089, VE/factor/pose_subset_parameterization.cpp): bit k set =
   * tangent component k (0..2 translation, 3..5 rotation) is held by zeroing its delta in Plus while its column stays in the linear
   * system (the reference's ComputeJacobian is [I6; 0] regardless). extrinsic_type_wheel 0 (ALL) = 0, 1 (TRANSLATION) = 0x38,
   * 2 (ROTATION) = 0x07, 3 (NO_Z) = 0x04, 4 (NO_ROTATION_NO_Z) = 0x3c. */
  uint32_t wheel_ext_const_components;
  /* gf2_marginalize: how the kept system A is factorised into the prior (linearized_jacobians, linearized_residuals), marginalization_factor.cpp:293-303.
   * 0 = rank-revealing Cholesky, pivots <= eps dropped (J0 = L^T P^T; same J0^T J0 and J0^T r0 as the reference up to O(eps) = 1e-8 absolute);
   * 1 = the reference's eigen-decomposition with its eps truncation, literally (a 12-sweep Jacobi solver: ~20x slower). */
  uint32_t marg_eig;
  double reserved_[5];
} gf2_solve_opts;

enum { /* gf2_solve_summary.termination */
  GF2_TERM_NO_CONVERGENCE = 0, /* iteration budget exhausted */
  GF2_TERM_FUNCTION_TOL = 1,
  GF2_TERM_GRADIENT_TOL = 2,
  GF2_TERM_PARAMETER_TOL = 3,
  GF2_TERM_MIN_RADIUS = 4,
  GF2_TERM_FAILURE = 5
};

typedef struct gf2_solve_summary {
  double initial_cost;
  double final_cost;
  int32_t iterations;       /* attempted trust-region steps */
  int32_t successful_steps;
  int32_t termination;
  int32_t pad_;
} gf2_solve_summary;

/* ---------------------------------------------------------------- solver */

const char* gf2_last_error(void);
int gf2_abi_version(void);
/* number of usable CUDA devices (0 on a CPU-only box); never fails */
int gf2_device_count(void);

int gf2_solver_create(const gf2_solver_cfg* cfg, gf2_solver** out);
void gf2_solver_destroy(gf2_solver* h);
/* Run all work of this handle on the caller's cudaStream_t (NULL restores the handle's own stream). */
int gf2_solver_set_stream(gf2_solver* h, void* cuda_stream);
/* Device-side copy of the mutable state (frame states + inverse depths) of windows [first, first+n) and its
 * restore: lets a caller re-solve the same batch without re-uploading (benchmarks, what-if solves). */
int gf2_snapshot_states(gf2_solver* h, int first, int n);
int gf2_restore_states(gf2_solver* h, int first, int n);
/* Pinned host memory for full-speed transfers through this ABI (cudaHostAlloc / cudaFreeHost). */
void* gf2_host_alloc(size_t bytes);
void gf2_host_free(void* p);

/* Packed states of windows [first, first+n): para_Pose [n][F][7], para_SpeedBias [n][F][9],
 * para_Ex_Pose[0] [n][7], para_Td [n], and (wheel) para_Ex_Pose_wheel [n][7], sx/sy/sw [n][3],
 * para_Td_wheel [n]. Wheel pointers may be NULL when the solver was created with use_wheel = 0. */
int gf2_set_states(gf2_solver* h, int first, int n, const double* para_pose, const double* para_speedbias,
                   const double* ex_pose, const double* td, const double* ex_pose_wheel,
                   const double* sxsysw, const double* td_wheel);

/* Landmark table in f_manager.feature list order restricted to used_num >= 4
 * (getDepthVector order, VE/estimator/feature_manager.cpp:286-302):
 *   n_landmarks [n]; inv_depth [n][max_landmarks] (para_Feature); start_frame [n][max_landmarks];
 *   track_len [n][max_landmarks] (feature_per_frame.size(), >= 2 here; the reference requires >= 4);
 *   fixed [n][max_landmarks] (estimate_flag == 1 -> SetParameterBlockConstant, estimator.cpp:3352).
 * Observations [n][max_obs], landmark-major: landmark l owns records obs_begin(l) .. +track_len(l),
 * obs_begin = exclusive prefix sum of track_len; record k is the observation in frame
 * start_frame + k (k = 0 is the host observation pts_i). frame_td [n][F] is cur_td of the frame
 * (FeaturePerFrame::cur_td, VE/estimator/feature_manager.cpp:69). */
int gf2_set_landmarks(gf2_solver* h, int first, int n, const int32_t* n_landmarks, const double* inv_depth,
                      const int32_t* start_frame, const int32_t* track_len, const uint8_t* fixed,
                      const gf2_obs* obs, const double* frame_td);

/* The same observations without their velocities: xy [n][max_obs][2] floats, 8 B per observation over the bus
 * instead of 16 (the observations are 3/4 of the bytes a window uploads). The velocity enters the factor only as
 * (td - td_i) * velocity (VE/factor/projectionTwoFrameOneCamFactor.cpp:53-54), which vanishes when td is not
 * estimated and equals the cur_td of every frame (ESTIMATE_TD = 0, VE/estimator/estimator.cpp:3055-3061, the mode
 * gf2_solve supports); the device records are written with velocity 0, so the results are bit-equal to the full
 * records in that mode and differ as soon as td != frame_td. Pass obs = NULL to gf2_set_landmarks (table only,
 * resident observations untouched) and call this for the observations. */
int gf2_set_observations_xy(gf2_solver* h, int first, int n, const float* xy);

/* IMU factors: record k of a window links frames k and k+1 ([n][F-1]). */
int gf2_set_imu(gf2_solver* h, int first, int n, const gf2_imu_preint* preint);
/* Same from raw samples, preintegrated on the device: samples [n][F-1][max_imu_samples],
 * n_samples [n][F-1], first sample acc_0/gyr_0 [n][F-1][6] (acc then gyr), linearisation biases
 * lin_bias [n][F-1][6] (ba then bg), noise = {ACC_N, GYR_N, ACC_W, GYR_W}. */
int gf2_imu_preintegrate(gf2_solver* h, int first, int n, const gf2_imu_sample* samples,
                         const int32_t* n_samples, const double* first_sample, const double* lin_bias,
                         const double noise[4]);
/* Re-run the preintegration kernel on the samples already resident on the device (no H2D). */
int gf2_imu_preintegrate_resident(gf2_solver* h, int first, int n, const double noise[4]);
/* Read back the device-side preintegration records ([n][F-1]). */
int gf2_get_imu(gf2_solver* h, int first, int n, gf2_imu_preint* preint);

int gf2_set_wheel(gf2_solver* h, int first, int n, const gf2_wheel_preint* preint);
/* Same from raw samples, preintegrated on the device (WheelIntegrationBase::push_back chain,
 * VE/factor/wheel_integration_base.h:41-178): samples [n][F-1][max_wheel_samples], n_samples [n][F-1],
 * first sample vel_0/gyr_0 [n][F-1][6], lin [n][F-1][4] = linearized sx, sy, sw, td, noise = {VEL_N_wheel, GYR_N_wheel}. */
int gf2_wheel_preintegrate(gf2_solver* h, int first, int n, const gf2_wheel_sample* samples,
                           const int32_t* n_samples, const double* first_sample, const double* lin,
                           const double noise[2]);
int gf2_get_wheel(gf2_solver* h, int first, int n, gf2_wheel_preint* preint);

/* Marginalization prior per window: n_rows [n] (0 = no prior), J0 [n][P][P] with
 * P = cfg.max_prior_rows (GF2_MAX_PRIOR_DIM when 0; row r, column c at r*P + c; rows/cols >= n_rows ignored),
 * r0 [n][P], n_blocks [n], blocks [n][2*F+8]. */
int gf2_set_prior(gf2_solver* h, int first, int n, const int32_t* n_rows, const double* J0, const double* r0,
                  const int32_t* n_blocks, const gf2_prior_block* blocks);

/* Marginalization after the solve (MarginalizationInfo::preMarginalize / marginalize / getParameterBlocks,
 * VE/factor/marginalization_factor.cpp:119-330, as driven by VE/estimator/estimator.cpp:3394-3690).
 * Builds the prior of the NEXT window from what is resident on the device (states as gf2_solve left them, landmarks,
 * IMU records, the current prior) and replaces the resident prior with it; its kept blocks are already renamed by
 * addr_shift (frame f -> f-1 for GF2_MARGIN_OLD; newest -> second-newest for GF2_MARGIN_SECOND_NEW), so after
 * slideWindow the caller uploads the new states / landmarks and solves. status [n] (may be NULL):
 *   0 ok; GF2_MARG_INVALID: m == 0, prior cleared (valid = false, :205-210); GF2_MARG_UNCHANGED: SECOND_NEW without the
 *   second-newest pose in the old prior (estimator.cpp:3599), old prior kept; GF2_MARG_UNSUPPORTED: the old prior holds a
 *   block this build cannot keep; GF2_MARG_DEGENERATE: the dropped FRAME block (pose 0 + speed-bias 0, 15 x 15 after the landmarks) has an
 *   eigenvalue below eps = 1e-8, where the truncated pseudo-inverse of :281 differs from the inverse (landmarks whose own eigenvalue is
 *   below eps — zero baseline, e.g. a robot turning on the spot — are truncated as the reference's pseudo-inverse truncates them and
 *   do NOT raise this), old prior kept; GF2_MARG_TOO_LARGE: n > max_prior_rows.
 * m_dims [n] (may be NULL): the number of marginalized tangent dimensions m. */
enum { GF2_MARGIN_OLD = 0, GF2_MARGIN_SECOND_NEW = 1 };
enum { GF2_MARG_INVALID = -1, GF2_MARG_UNCHANGED = -2, GF2_MARG_UNSUPPORTED = -3, GF2_MARG_DEGENERATE = -4, GF2_MARG_TOO_LARGE = -5 };
int gf2_marginalize(gf2_solver* h, int first, int n, int32_t mode, const gf2_solve_opts* opts, int32_t* status,
                    int32_t* m_dims);
/* The same in two halves, for callers that do not need the new prior before the next frame (one robot: the reference's node spends the
 * time until the next image idle): gf2_marginalize_async launches the kernels on the handle's stream and returns; the new prior replaces
 * the resident one in stream order, so a following gf2_solve already sees it. gf2_marginalize_wait blocks until those kernels are done and
 * returns their status / m_dims (semantics as above). gf2_marginalize == async + wait. */
int gf2_marginalize_async(gf2_solver* h, int first, int n, int32_t mode, const gf2_solve_opts* opts);
int gf2_marginalize_wait(gf2_solver* h, int first, int n, int32_t* status, int32_t* m_dims);
/* The resident prior (same layout as gf2_set_prior). */
int gf2_get_prior(gf2_solver* h, int first, int n, int32_t* n_rows, double* J0, double* r0, int32_t* n_blocks,
                  gf2_prior_block* blocks);
/* Device time of the last gf2_marginalize (CUDA events on the handle's stream), ms. */
double gf2_last_marginalize_ms(gf2_solver* h);

/* LiDAR plane factors: n_planes [n], planes [n][max_planes] sorted or not by frame. */
int gf2_set_planes(gf2_solver* h, int first, int n, const int32_t* n_planes, const gf2_plane* planes);
/* alpha_time of the CTLidarPlaneNormFactor records (ct == 1): alpha [n][max_planes], entry q belongs to plane q of the window
 * (entries of ct == 0 planes are ignored). Call after gf2_set_planes; without it alpha_time is 0 (the begin pose). */
int gf2_set_plane_alpha(gf2_solver* h, int first, int n, const double* alpha);

/* Solve windows [first, first+n) (ceres::Solve with DENSE_SCHUR + traditional DOGLEG semantics).
 * Synchronous. summaries may be NULL. */
int gf2_solve(gf2_solver* h, int first, int n, const gf2_solve_opts* opts, gf2_solve_summary* summaries);

/* One linearisation only (residuals, Jacobians, Schur): fills the reduced system of each window.
 * Used by parity tests and by the roofline measurement. reduced_dim = gf2_reduced_dim(). */
int gf2_linearize(gf2_solver* h, int first, int n, const gf2_solve_opts* opts);
int gf2_reduced_dim(gf2_solver* h, const gf2_solve_opts* opts);
/* Reduced camera system of the last linearisation: S [n][D][D] (full symmetric, row-major, WITHOUT
 * the mu*diag regularisation), g [n][D] (reduced gradient J^T r after eliminating landmarks),
 * cost [n] = 0.5 * sum rho(|r|^2). D = gf2_reduced_dim(h, opts) of the options last linearised with.
 * Tangent order: per frame [pose 6 | speed-bias 9]; when a wheel calibration block is free
 * (const_mask lacks GF2_CONST_EX_WHEEL / _WHEEL_INTRINSIC / _TD_WHEEL) one more block of 15:
 * [ex-wheel 6 | sx sy sw | td-wheel | 5 unused], rows/columns of constant sub-blocks are zero. */
int gf2_get_reduced_system(gf2_solver* h, int first, int n, double* S, double* g, double* cost);

int gf2_get_states(gf2_solver* h, int first, int n, double* para_pose, double* para_speedbias, double* ex_pose,
                   double* td, double* ex_pose_wheel, double* sxsysw, double* td_wheel);
int gf2_get_landmarks(gf2_solver* h, int first, int n, double* inv_depth);

/* Multi-GPU, factor-sharded mode: every rank holds the same windows' states, a
 * disjoint subset of the landmarks/planes; the reduced system is summed with one ncclAllReduce per
 * linearisation. nccl_unique_id points at an ncclUniqueId (128 bytes) produced by rank 0. */
int gf2_comm_init(gf2_solver* h, int rank, int nranks, const void* nccl_unique_id);
int gf2_comm_unique_id(void* out_128_bytes);

/* Device-time of the phases of the last gf2_solve / gf2_linearize on this handle, in milliseconds,
 * measured with CUDA events on the handle's stream: [0] total, [1] linearise kernels (sum),
 * [2] reduced solve kernels, [3] back-substitution + candidate evaluation kernels, [4] number of
 * kernel launches, [5] number of linearise launches. */
int gf2_last_timing(gf2_solver* h, double out[8]);
/* Per-iteration record of the last solve: out [n][64][6] = {candidate cost, model cost change, relative decrease,
 * trust-region radius after the step, ambient step norm, decision (0 reject, 1 accept, 2 terminate/invalid)}. */
int gf2_get_trace(gf2_solver* h, int first, int n, double* out);

/* ---------------------------------------------------------------- tracker */

typedef struct gf2_tracker gf2_tracker;

typedef struct gf2_tracker_cfg {
  int32_t device;
  int32_t width, height; /* COL, ROW */
  int32_t max_pts;       /* MAX_CNT upper bound */
  int32_t win;           /* 21 */
  int32_t max_level;     /* capacity: 3 */
  int32_t max_iters;     /* 30 */
  int32_t max_streams;   /* independent image streams tracked per call (batch), >= 1 */
  double eps;            /* 0.01 */
  double min_eig;        /* 1e-4 */
} gf2_tracker_cfg;

enum { GF2_LK_USE_INITIAL_FLOW = 4 /* cv::OPTFLOW_USE_INITIAL_FLOW */ };

int gf2_tracker_create(const gf2_tracker_cfg* cfg, gf2_tracker** out);
void gf2_tracker_destroy(gf2_tracker* h);

/* cv::calcOpticalFlowPyrLK(prev, cur, prev_pts, cur_pts, status, err, Size(win,win), max_level,
 * TermCriteria(COUNT+EPS, max_iters, eps), flags, min_eig) for `n_streams` independent image pairs.
 * prev/cur: [n_streams] images of height*stride bytes (u8); prev may be NULL to reuse the pyramid
 * kept from the previous call's `cur` of the same stream (the reference's prev_img = cur_img,
 * feature_tracker.cpp:307). pts: [n_streams][max_pts][2] float; n_pts [n_streams]. */
int gf2_tracker_track(gf2_tracker* h, int n_streams, const uint8_t* prev, const uint8_t* cur, size_t stride,
                      const int32_t* n_pts, const float* prev_pts, float* cur_pts, uint8_t* status, float* err,
                      int flags, int max_level);

/* The forward + backward check of trackImage (feature_tracker.cpp:135-153) fused: forward LK at
 * max_level, reverse LK at level 1 with initial flow, status &= (reverse ok && round trip <= 0.5 px).
 * Outputs cur_pts and the combined status. */
int gf2_tracker_track_fb(gf2_tracker* h, int n_streams, const uint8_t* prev, const uint8_t* cur, size_t stride,
                         const int32_t* n_pts, const float* prev_pts, float* cur_pts, uint8_t* status,
                         int max_level);

/* The whole LK stage of FeatureTracker::trackImage (feature_tracker.cpp:113-153). predict_pts == NULL: forward LK at
 * max_level (:132-135). predict_pts != NULL (hasPrediction, set by setPrediction :851-871): forward LK at level 1 from the
 * predicted positions with OPTFLOW_USE_INITIAL_FLOW (:122-123); every stream with fewer than 10 successes is redone at
 * max_level from prev_pts without the prediction (:124-131; decided per stream on the device). flow_back != 0 adds the
 * reverse check of :137-153. Outputs cur_pts and the combined status. */
int gf2_tracker_track_image(gf2_tracker* h, int n_streams, const uint8_t* prev, const uint8_t* cur, size_t stride,
                            const int32_t* n_pts, const float* prev_pts, const float* predict_pts, int flow_back,
                            float* cur_pts, uint8_t* status, int max_level);

/* cv::goodFeaturesToTrack(img, corners, max_corners, quality_level, min_distance, mask) with the defaults trackImage uses
 * (feature_tracker.cpp:198: blockSize 3, Sobel aperture 3, min-eigenvalue score) for n_streams images.
 * img: [n_streams] images of height*stride bytes, or NULL to detect on the `cur` image of the last gf2_tracker_track* call
 * (the reference detects on cur_img right after tracking it). mask: [n_streams][height][width] bytes (0 = excluded, the
 * setMask() image :56-83) or NULL. max_corners[s] = MAX_CNT - tracked; 0 skips the stream (the reference's n_pts.clear()
 * branch), negative is rejected; at most cfg.max_pts corners are returned per stream.
 * out_xy: [n_streams][max_pts][2] in cv's order (descending score, ties by descending address, greedy min-distance
 * acceptance) - the order feature ids are assigned in (addPoints :85-93); out_n [n_streams].
 * The score map is bit-exact with cv::cornerMinEigenVal for widths that are multiples of 32. */
int gf2_tracker_detect(gf2_tracker* h, int n_streams, const uint8_t* img, size_t stride, const uint8_t* mask,
                       const int32_t* max_corners, double quality_level, double min_distance, float* out_xy,
                       int32_t* out_n);

/* The host half of gf2_tracker_detect on its own (needs no device): cv's corner selection (featureselect.cpp: sort by
 * descending score then descending address, greedy min-distance acceptance through a cvRound(min_distance) grid) over n
 * candidate keys = ordered_score << 32 | (y * width + x). Writes at most max_corners corners to out_xy and their number
 * to out_n. */
int gf2_detect_select(const uint64_t* keys, int n, int width, int height, int max_corners, double min_distance,
                      float* out_xy, int32_t* out_n);

/* cv::cornerMinEigenVal(img, eig, 3, 3): the detector's score map, eig [n_streams][height][width] float (parity checks). */
int gf2_tracker_min_eigen_map(gf2_tracker* h, int n_streams, const uint8_t* img, size_t stride, float* eig);

/* cv::createCLAHE(clip_limit, Size(tiles_x, tiles_y))->apply(img, out) for 8-bit images: what getImageFromMsg does when
 * EQUALIZE is set (VE/rosNodeTest.cpp:271-276; `equalize: 1` in config/realsense/m3dgr.yaml:16; cv defaults 40.0, 8x8).
 * Bit-exact with cv2. The tile grid must divide the image size (cv pads by reflection otherwise: GF2_ERR_UNSUPPORTED).
 * clip_limit <= 0 means no clipping, as in cv. */
int gf2_tracker_equalize(gf2_tracker* h, int n_streams, const uint8_t* img, size_t stride, double clip_limit,
                         int tiles_x, int tiles_y, uint8_t* out);

/* The same equalisation fused into the front end: after this call every image uploaded by gf2_tracker_track*,
 * gf2_tracker_detect and gf2_tracker_min_eigen_map is replaced ON THE DEVICE by its CLAHE before anything reads it, so the
 * node's cv::CLAHE call (and one host pass over the image) disappears. clip_limit <= 0 switches it off again. */
int gf2_tracker_set_equalize(gf2_tracker* h, double clip_limit, int tiles_x, int tiles_y);

/* The `cur` images of the last gf2_tracker_track* call as the device holds them (equalised if set): trackImage reads
 * cur_img on the host for its grey > 250 test (feature_tracker.cpp:160-167). out: [n_streams][height][width]. */
int gf2_tracker_get_image(gf2_tracker* h, int n_streams, uint8_t* out);

int gf2_tracker_last_timing(gf2_tracker* h, double out[8]);

/* ---------------------------------------------------------------- LIO factor construction */
/* lidarodom::addSurfCostFactor (LIO/liw/lio/lidarodom.cpp:929-1071) for one scan: per keypoint the voxel-hash nearest-neighbour
 * search searchNeighbors (:1087-1165), the neighbourhood statistics computeNeighborhoodDistribution (:887-927: barycentre,
 * covariance, normal = eigenvector of the smallest eigenvalue, planarity a2D), the weights and the point-to-plane gate, in
 * keypoint order with the max_num_residuals cap. The output records are the constructor arguments of CTLidarPlaneNormFactor /
 * LidarPlaneNormFactor (LIO/liw/lidarFactor.cpp:13-16, 52-56). The voxel map itself (addPointToMap, :1167-1213) stays with the
 * caller, who hands over a snapshot. */
typedef struct gf2_lio gf2_lio;

typedef struct gf2_lio_cfg {
  int32_t device;
  int32_t max_voxels;           /* capacity of the map snapshot */
  int32_t max_points_per_voxel; /* max_num_points_in_voxel (20) */
  int32_t max_keypoints;
} gf2_lio_cfg;

typedef struct gf2_lio_keypoint { /* point3D (LIO/common/cloudMap.hpp:17-32): the fields the path reads */
  double raw_point[3];
  double point[3];
  double alpha_time;
} gf2_lio_keypoint;

enum { GF2_ICP_CT_POINT_TO_PLANE = 0, GF2_ICP_POINT_TO_PLANE = 1 };

typedef struct gf2_lio_opts {         /* lidarodom options (lidarodom.h:33-52) + the frame state the function reads */
  double size_voxel_map;              /* 0.2 */
  double max_dist_to_plane_icp;       /* 0.3 */
  double power_planarity;             /* 2.0 */
  double weight_alpha;                /* 0.9 */
  double weight_neighborhood;         /* 0.1 */
  int32_t nb_voxels_visited;          /* frame_id < init_num_frames ? 2 : voxel_neighborhood (1) */
  int32_t threshold_voxel_capacity;   /* frame_id < init_num_frames ? 1 : threshold_voxel_occupancy */
  int32_t max_number_neighbors;       /* 20, <= 32 */
  int32_t min_number_neighbors;       /* 20 */
  int32_t num_closest_neighbors;      /* 1 */
  int32_t max_num_residuals;          /* 2000 */
  int32_t icp_model;                  /* GF2_ICP_* */
  int32_t pad_;
  double translation_begin[3];        /* p_frame->p_state->translation_begin: orients the normal (:939) */
  double rotation[4];                 /* p_state->rotation  [x y z w] (POINT_TO_PLANE: point_end, :1040-1042) */
  double translation[3];              /* p_state->translation */
  double R_IL[9];                     /* TIL_ = [R_IL | t_IL], row-major: location = TIL_ * raw_point (:988) */
  double t_IL[3];
} gf2_lio_opts;

int gf2_lio_create(const gf2_lio_cfg* cfg, gf2_lio** out);
void gf2_lio_destroy(gf2_lio* h);
/* Snapshot of the voxelHashMap (tsl::robin_map<voxel, voxelBlock>, cloudMap.hpp:34-83): n_voxels entries in any order,
 * keys [n][3] (short x, y, z), n_points [n], points [n][max_points_per_voxel][3] in insertion order. Duplicate keys are rejected. */
int gf2_lio_set_map(gf2_lio* h, int n_voxels, const int16_t* keys, const int32_t* n_points, const double* points);
/* Device-resident map maintenance: lidarodom::addPointToMap (LIO/liw/lio/lidarodom.cpp:1167-1213) for the n points of a scan IN ORDER
 * (map_incremental, :1226-1237): a point goes into the voxel short(p / size_voxel_map); an existing voxel takes it if it is not
 * full (max_points_per_voxel), the point is farther than min_distance_points from every point already there and the voxel holds at
 * least min_num_points points (when min_num_points > 0); a missing voxel is created only when min_num_points <= 0. Points interact
 * only inside a voxel, so the scan is sorted by voxel key (stable) and one thread walks each voxel's points in scan order: the
 * result equals the sequential loop. The map then stays on the device for gf2_lio_build_factors (no snapshot upload). */
int gf2_lio_add_points(gf2_lio* h, int n, const double* points /* [n][3] */, double size_voxel_map, double min_distance_points,
                       int min_num_points);
/* Number of voxels held / snapshot of the device map in the layout of gf2_lio_set_map (voxels in ascending key order). */
int gf2_lio_map_size(gf2_lio* h, int32_t* n_voxels);
int gf2_lio_get_map(gf2_lio* h, int16_t* keys, int32_t* n_points, double* points);

/* out_factors [max(max_num_residuals, 1)] (the reference tests the cap after pushing): p_body = raw_point (CT) or point_end (POINT_TO_PLANE), normal = nvec, offset =
 * -nvec . neighbour, weight, frame = index of the keypoint (the reference's valid_keypoints); out_alpha [same] =
 * kp.alpha_time; out_neighbors (nullable) [n_keypoints][max_number_neighbors][3] + out_n_neighbors [n_keypoints]: the sorted
 * neighbour lists (what searchNeighbors returns); *n_out residuals written. */
int gf2_lio_build_factors(gf2_lio* h, int n_keypoints, const gf2_lio_keypoint* keypoints, const gf2_lio_opts* opts,
                          gf2_plane* out_factors, double* out_alpha, int32_t* n_out, double* out_neighbors,
                          int32_t* out_n_neighbors);
int gf2_lio_last_timing(gf2_lio* h, double out[8]);

#ifdef __cplusplus
}
#endif
#endif /* GF2_ABI_H_ */
