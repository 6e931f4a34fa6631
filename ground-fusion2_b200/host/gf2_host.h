// gf2_host.h — host-side C++ mirror of the reference classes on the hot path, on top of the C ABI (include/gf2_abi.h).
//
// Same class / member / method names as the reference so that the calling code (rosNodeTest.cpp, processImage) reads the
// same; Eigen / OpenCV / Ceres / ROS do not exist in this build environment, so vectors are plain structs and images are raw
// uint8 pointers. What stays on the host is exactly what the reference keeps there: buffering of raw IMU samples per
// interval (IntegrationBase::push_back, VE/factor/integration_base.h:39-46), the landmark table and its index order
// (FeatureManager, VE/estimator/feature_manager.cpp:43-55,249-302), state packing (Estimator::vector2double / double2vector,
// VE/estimator/estimator.cpp:2337-2414 / 2501-2630) and the tracker glue (FeatureTracker::inBorder / setMask / addPoints /
// undistortedPts / ptsVelocity, VE/featureTracker/feature_tracker.cpp:14-93,797-847). All arithmetic of the hot path runs
// in the CUDA kernels behind gf2_solve / gf2_imu_preintegrate / gf2_tracker_track_fb; nothing here is a CPU fallback.
#pragma once
#include <cmath>
#include <cstdint>
#include <list>
#include <map>
#include <queue>
#include <set>
#include <string>
#include <utility>
#include <vector>
#include "../../include/gf2_abi.h"

namespace gf2host {

const double FOCAL_LENGTH = 600.0;  // VE/estimator/parameters.h:23-25
const int WINDOW_SIZE = 10;
const int NUM_OF_F = 1000;

struct Vector3d { double x = 0, y = 0, z = 0; };
struct Matrix3d { double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; };  // row-major
struct Quaterniond { double w = 1, x = 0, y = 0, z = 0; };
Quaterniond quatFromMatrix(const Matrix3d& R);   // Eigen Quaterniond(Matrix3d)
Matrix3d toRotationMatrix(const Quaterniond& q);  // Eigen toRotationMatrix (after normalisation where the reference normalises)
Vector3d R2ypr(const Matrix3d& R);               // Utility::R2ypr, VE/utility/utility.h:78-93 (degrees)
Matrix3d ypr2R(const Vector3d& ypr);             // Utility::ypr2R, :95-121
Matrix3d mul(const Matrix3d& a, const Matrix3d& b);
Vector3d mul(const Matrix3d& a, const Vector3d& v);
Matrix3d transpose(const Matrix3d& a);

// Parameters the path reads (subset of readParameters, VE/estimator/parameters.cpp:142-560)
struct Parameters {
  double ACC_N = 0.1, ACC_W = 0.001, GYR_N = 0.01, GYR_W = 0.0001, G_NORM = 9.81007;
  double SOLVER_TIME = 0.04; int NUM_ITERATIONS = 8;
  int USE_MCC = 0, DEPTH = 0;   // parameters.cpp:164,472
  int ESTIMATE_EXTRINSIC = 0, ESTIMATE_TD = 0, USE_IMU = 1, USE_WHEEL = 0, EQUALIZE = 0, MAX_CNT = 150, MIN_DIST = 30, FLOW_BACK = 1, ROW = 480, COL = 640;
  double TD = 0.0, F_THRESHOLD = 1.0, MIN_PARALLAX = 10.0 / FOCAL_LENGTH;
  Matrix3d RIC; Vector3d TIC;
  // wheel odometer (parameters.cpp:181-194, 234-335, 501-502)
  int ONLY_INITIAL_WITH_WHEEL = 0, ESTIMATE_EXTRINSIC_WHEEL = 0, ESTIMATE_INTRINSIC_WHEEL = 0, ESTIMATE_TD_WHEEL = 0, EXTRINSIC_TYPE_WHEEL = 0;
  double VEL_N_wheel = 0.01, GYR_N_wheel = 0.004, SX = 1.0, SY = 1.0, SW = 1.0, TD_WHEEL = 0.0;
  Matrix3d RIO; Vector3d TIO;
  double fx = 0, fy = 0, cx = 0, cy = 0, k1 = 0, k2 = 0, p1 = 0, p2 = 0;  // camodocal PINHOLE (GF/config/realsense/color.yaml)
};
// Reads the OpenCV-FileStorage YAML dialect of GF/config/realsense/m3dgr.yaml (+ the camera file named by cam0_calib);
// absent keys read as 0 like cv::FileNode. Returns false if the file cannot be opened.
bool readParameters(const std::string& config_file, Parameters& P);

// Raw-sample buffer of one preintegration interval; the integration itself is gf2_imu_preintegrate on the device.
class IntegrationBase {
 public:
  IntegrationBase(const Vector3d& _acc_0, const Vector3d& _gyr_0, const Vector3d& _linearized_ba, const Vector3d& _linearized_bg)
      : acc_0(_acc_0), gyr_0(_gyr_0), linearized_acc(_acc_0), linearized_gyr(_gyr_0), linearized_ba(_linearized_ba), linearized_bg(_linearized_bg) {}
  void push_back(double dt, const Vector3d& acc, const Vector3d& gyr) { dt_buf.push_back(dt); acc_buf.push_back(acc); gyr_buf.push_back(gyr); sum_dt += dt; acc_0 = acc; gyr_0 = gyr; }
  Vector3d acc_0, gyr_0;
  const Vector3d linearized_acc, linearized_gyr;
  Vector3d linearized_ba, linearized_bg;
  double sum_dt = 0.0;
  std::vector<double> dt_buf;
  std::vector<Vector3d> acc_buf, gyr_buf;
};

// Raw-sample buffer of one wheel preintegration interval (WheelIntegrationBase::push_back, VE/factor/wheel_integration_base.h:41-48);
// the integration itself is gf2_wheel_preintegrate on the device.
class WheelIntegrationBase {
 public:
  WheelIntegrationBase(const Vector3d& _vel_0, const Vector3d& _gyr_0, double _sx, double _sy, double _sw, double _td)
      : vel_0(_vel_0), gyr_0(_gyr_0), linearized_vel(_vel_0), linearized_gyr(_gyr_0), linearized_sx(_sx), linearized_sy(_sy), linearized_sw(_sw), linearized_td(_td) {}
  void push_back(double dt, const Vector3d& vel, const Vector3d& gyr) { dt_buf.push_back(dt); vel_buf.push_back(vel); gyr_buf.push_back(gyr); sum_dt += dt; vel_0 = vel; gyr_0 = gyr; }
  Vector3d vel_0, gyr_0;
  const Vector3d linearized_vel, linearized_gyr;
  double linearized_sx, linearized_sy, linearized_sw, linearized_td;
  double sum_dt = 0.0;
  std::vector<double> dt_buf;
  std::vector<Vector3d> vel_buf, gyr_buf;
};

class FeaturePerFrame {  // VE/estimator/feature_manager.h:28-62
 public:
  FeaturePerFrame(const double _point[8], double td) : cur_td(td) {
    point = {_point[0], _point[1], _point[2]}; uv[0] = _point[3]; uv[1] = _point[4]; velocity[0] = _point[5]; velocity[1] = _point[6]; depth = _point[7];
  }
  double cur_td;
  Vector3d point;
  double uv[2], velocity[2], depth;
};

class FeaturePerId {  // VE/estimator/feature_manager.h:64-90
 public:
  FeaturePerId(int _feature_id, int _start_frame) : feature_id(_feature_id), start_frame(_start_frame) {}
  const int feature_id;
  int start_frame;
  std::vector<FeaturePerFrame> feature_per_frame;
  int used_num = 0;
  double estimated_depth = -1.0;
  int estimate_flag = 0;  // 1: depth verified by the depth camera -> held constant (estimator.cpp:3352)
  int solve_flag = 0;     // 0 not solved, 1 ok, 2 failed
  int endFrame() const { return start_frame + (int)feature_per_frame.size() - 1; }
};

class FeatureManager {
 public:
  std::list<FeaturePerId> feature;
  int getFeatureCount();                         // feature_manager.cpp:43-55
  std::vector<double> getDepthVector();          // :286-302 (inverse depths, list order, used_num >= 4)
  void setDepth(const std::vector<double>& x);   // :249-267
  void removeFailures();                         // :269-278
  void clearDepth();                             // :280-284
  // keyframe decision + table update of one image (feature_manager.cpp:57-116); true = the second newest frame is a keyframe (MARGIN_OLD)
  bool addFeatureCheckParallax(int frame_count, const std::map<int, std::vector<std::pair<int, std::vector<double>>>>& image, double td);
  double compensatedParallax2(const FeaturePerId& it_per_id, int frame_count) const;   // :978-1011
  void removeOutlier(const std::set<int>& outlierIndex);                               // :801-816
  void removeBackShiftDepth(const Matrix3d& marg_R, const Vector3d& marg_P, const Matrix3d& new_R, const Vector3d& new_P);  // :818-856
  void removeBack();                                                                   // :858-874
  void removeFront(int frame_count);                                                   // :914-934
  // depth initialisation of landmarks without one (estimated_depth <= 0, >= 4 observations): :669-724 (SVD over the track) and
  // :726-799 (RGB-D: mean of the depth measurements that reproject within 10/460 into the other frames)
  void triangulate(int frameCnt, const Vector3d Ps[], const Matrix3d Rs[], const Vector3d tic[], const Matrix3d ric[]);
  void triangulateWithDepth(int frameCnt, const Vector3d Ps[], const Matrix3d Rs[], const Vector3d tic[], const Matrix3d ric[]);
  int last_track_num = 0, new_feature_num = 0, long_track_num = 0;
  double last_average_parallax = 0.0, MIN_PARALLAX = 10.0 / FOCAL_LENGTH, INIT_DEPTH = 5.0;
  int depth_threshold = 3;
  // append the observations of one image in the reference's container: map<id, vector<pair<cam, 8-vector>>> (cam 0 only)
  void addFeatures(int frame_count, const std::map<int, std::vector<std::pair<int, std::vector<double>>>>& image, double td);  // :67-88 (list insertion part)
};

// The kept part of a MarginalizationInfo (VE/factor/marginalization_factor.h:74-79)
struct MarginalizationPrior {
  bool valid = false;
  int n = 0;
  std::vector<double> linearized_jacobians;  // n x n row-major
  std::vector<double> linearized_residuals;
  std::vector<gf2_prior_block> blocks;       // keep_block_size / idx / data with the block each one maps to
};

class FeatureTracker;

// sync_process of the node for the RGB-D configuration (VE/rosNodeTest.cpp:305-598, the non-YOLO branch :395-428): colour and depth messages are
// paired when their stamps differ by at most 3 ms; the older unmatched head is thrown away.
class ImagePairSynchronizer {
 public:
  void push0(double t, int handle) { img0_buf.push({t, handle}); }
  void push1(double t, int handle) { img1_buf.push({t, handle}); }
  bool next(double* time, int* handle0, int* handle1);   // one pass of the loop body; false = nothing to pair yet
  int thrown0 = 0, thrown1 = 0;
 private:
  std::queue<std::pair<double, int>> img0_buf, img1_buf;
};

class Estimator {
 public:
  Estimator();
  ~Estimator();
  void setParameter(const Parameters& p);
  void clearState();
  void vector2double();   // estimator.cpp:2337-2414
  void double2vector();   // estimator.cpp:2501-2630 (yaw / position re-anchoring to frame 0, setDepth)
  // measurement queues and the consumer loop (estimator.cpp:324-372 inputIMU / inputWheel, :422-545 interval extraction, :554-763
  // processMeasurements with MULTIPLE_THREAD == 0: one pass, returns when a stream has to be waited for)
  typedef std::map<int, std::vector<std::pair<int, std::vector<double>>>> FeatureFrame;
  void inputIMU(double t, const Vector3d& linearAcceleration, const Vector3d& angularVelocity);
  void inputWheel(double t, const Vector3d& linearVelocity, const Vector3d& angularVelocity);
  void inputFeature(double t, const FeatureFrame& featureFrame);   // what inputImage pushes after trackImage (:232-236)
  // estimator.cpp:213-242 with MULTIPLE_THREAD == 0: trackImage on the device, push, processMeasurements; after the solve the tracker gets
  // removeOutliers + the prediction for the next frame (:1185-1189). _img: ROW x COL bytes, _img1: depth in millimetres or null.
  void inputImage(double t, const uint8_t* _img, const uint16_t* _img1 = nullptr);
  FeatureTracker* featureTracker = nullptr;   // owned; created by setParameter
  int inputImageCnt = 0;
  bool MULTIPLE_THREAD = false;
  bool IMUAvailable(double t) const;
  bool WheelAvailable(double t) const;
  bool getIMUInterval(double t0, double t1, std::vector<std::pair<double, Vector3d>>& accVector, std::vector<std::pair<double, Vector3d>>& gyrVector);
  bool getWheelInterval(double t0, double t1, std::vector<std::pair<double, Vector3d>>& velVector, std::vector<std::pair<double, Vector3d>>& gyrVector);
  int processMeasurements();   // returns the number of images consumed
  // IMU-rate pose output: fastPredictIMU (estimator.cpp:4076-4093, mid-point propagation of the latest state with every inputIMU sample) and
  // updateLatestStates (:4203-4228, re-anchor at the newest window state after each processImage and replay the queued samples)
  void fastPredictIMU(double t, const Vector3d& linear_acceleration, const Vector3d& angular_velocity);
  void updateLatestStates();
  double latest_time = 0.0; Vector3d latest_P, latest_V, latest_Ba, latest_Bg, latest_acc_0, latest_gyr_0; Matrix3d latest_Q;   // latest_Q kept as a rotation matrix
  bool latest_valid = false;
  std::queue<std::pair<double, Vector3d>> accBuf, gyrBuf, wheelVelBuf, wheelGyrBuf;
  std::queue<std::pair<double, FeatureFrame>> featureBuf;
  double prevTime = -1.0, curTime = 0.0, prevTime_wheel = -1.0, curTime_wheel = 0.0;
  bool solve_enabled = true;   // test hook: false = processImage keeps all bookkeeping but skips optimization() (CPU-only tests of the glue)
  // measurement processing around the solve (steady state, solver_flag == NON_LINEAR)
  void processIMU(double t, double dt, const Vector3d& linear_acceleration, const Vector3d& angular_velocity);   // :795-836 (sample buffering)
  void processWheel(double t, double dt, const Vector3d& linear_velocity, const Vector3d& angular_velocity);     // :837-896 (buffering + dead reckoning of the newest frame)
  void processImage(const std::map<int, std::vector<std::pair<int, std::vector<double>>>>& image, double header);  // :897-1216, the `else // not ini` branch
  void slideWindow();      // :3700-3857
  void slideWindowNew();   // :3859-3868
  void slideWindowOld();   // :3870-3899
  void outliersRejection(std::set<int>& removeIndex);          // :3971-4028
  void movingConsistencyCheckW(std::set<int>& removeIndex);    // :4030-4074
  void getPoseInWorldFrame(int index, double T[16]) const;     // :3901-3913 (row-major 4x4)
  std::map<int, Vector3d> predictPtsInNextFrame() const;       // :3915-3948 (the map handed to FeatureTracker::setPrediction)
  // pubOdometry's result line (VE/utility/visualization.cpp:315-387): "stamp x y z qx qy qz qw" of Ps / Rs[WINDOW_SIZE], fixed, 9 decimals (TUM format)
  std::string tumLine(double stamp) const;
  bool appendTum(const std::string& path, double stamp) const;
  // Test hook: when capture is set, optimization() keeps a copy of everything it hands to the C ABI and of what comes back, so that a
  // replay can be checked step by step against the CPU oracle on identical inputs (tests/test_gpu_replay.py).
  struct Capture {
    double pose[(WINDOW_SIZE + 1) * 7], sb[(WINDOW_SIZE + 1) * 9], ex[7], td, frame_td[WINDOW_SIZE + 1];
    int32_t n_lm = 0; std::vector<int32_t> start, len; std::vector<uint8_t> fixed; std::vector<double> invdep; std::vector<gf2_obs> obs;
    std::vector<gf2_imu_sample> imu_samples; std::vector<int32_t> imu_n; std::vector<double> imu_first, imu_bias;
    int32_t prior_rows = 0, prior_nblocks = 0; std::vector<double> prior_J0, prior_r0; std::vector<gf2_prior_block> prior_blocks;
    uint32_t const_mask = 0; int32_t marg_mode = 0;
    int32_t use_wheel = 0; std::vector<gf2_wheel_sample> wheel_samples; std::vector<int32_t> wheel_n; std::vector<double> wheel_first, wheel_lin;
    double exw[7], sxsysw[3], tdw, exw_out[7];
    double pose_out[(WINDOW_SIZE + 1) * 7], sb_out[(WINDOW_SIZE + 1) * 9]; std::vector<double> invdep_out;   // straight from gf2_get_states / gf2_get_landmarks
    std::vector<gf2_plane> planes;   // the LiDAR plane factors handed to gf2_set_planes
    double pose_marg[(WINDOW_SIZE + 1) * 7], sb_marg[(WINDOW_SIZE + 1) * 9]; std::vector<double> invdep_marg;  // the states gf2_marginalize ran at
  };
  bool capture = false; Capture cap;
  // LiDAR point-to-plane factors attached to window poses: the composition of BASELINE.json config 4 / 5 (in the reference the
  // LIO residuals live in lidarodom's own per-scan problem, LIO/liw/lio/lidarodom.cpp:534-661; here the factors addSurfCostFactor builds for
  // the scan of a frame, gf2_lio_build_factors, ride in the visual-inertial window on that frame's pose). inputLidarPlanes() attaches the
  // factors of the NEWEST frame before its processImage(); slideWindow() moves them with their frame; the factors of a marginalized frame
  // are dropped (they do not enter the prior).
  static const int kMaxPlanesPerFrame = 480;   // 11 frames x 480 = 5,280 planes: within the solver's plane-task capacity (BASELINE config 4: 500 per frame, 5,000 per window)
  std::vector<gf2_plane> lidar_planes[WINDOW_SIZE + 1];
  double lidar_sqrt_info = 31.622776601683793;   // sqrt(1 / laser_point_cov), LIO/liw/lio/lidarodom.cpp:10,13
  void inputLidarPlanes(const gf2_plane* planes, int n);
  bool planes_resident = false;
  double Headers[WINDOW_SIZE + 1] = {0};
  Matrix3d back_R0; Vector3d back_P0;
  int sum_of_back = 0, sum_of_front = 0;
  bool first_imu = false, first_wheel = false, DEPTH = false, USE_MCC = false;
  Vector3d acc_0, gyr_0, vel_0_wheel, gyr_0_wheel, latest_vel_wheel_0;
  void optimization();    // estimator.cpp:2951-3693: the ceres::Problem build + ceres::Solve (gf2_solve), then the
                          // marginalization of the oldest / second-newest frame (gf2_marginalize) into last_marginalization_info
  enum MarginalizationFlag { MARGIN_OLD = 0, MARGIN_SECOND_NEW = 1 };  // estimator.h:60-64
  MarginalizationFlag marginalization_flag = MARGIN_OLD;
  int last_marginalization_status = 0;   // gf2_marginalize status of the last optimization()
  // The prior lives on the device between frames: optimization() launches the marginalization asynchronously (gf2_marginalize_async) and
  // returns after the solve; the status is collected by the next optimization() (or by gf2h_get_prior / the capture hook, which also
  // download the prior into last_marginalization_info). async_marginalization = false restores the blocking behaviour.
  bool async_marginalization = true, marg_pending = false, prior_on_device = false, host_prior_current = true;
  void finishMarginalization(bool download);
  const char* lastError() const { return last_error.c_str(); }

  Parameters P;
  Vector3d Ps[WINDOW_SIZE + 1], Vs[WINDOW_SIZE + 1], Bas[WINDOW_SIZE + 1], Bgs[WINDOW_SIZE + 1];
  Matrix3d Rs[WINDOW_SIZE + 1];
  Vector3d tic[2]; Matrix3d ric[2];
  double td = 0.0;
  int frame_count = WINDOW_SIZE;
  bool openExEstimation = false, failure_occur = false;
  // wheel odometer states (estimator.h: tio, rio, sx, sy, sw, td_wheel; estimator.cpp:3063-3118, 3181-3212)
  Vector3d tio; Matrix3d rio; double sx = 1.0, sy = 1.0, sw = 1.0, td_wheel = 0.0;
  bool openExWheelEstimation = false, openIxEstimation = false, wdetect = false, wheelanomaly = false;
  WheelIntegrationBase* pre_integrations_wheel[WINDOW_SIZE + 1] = {nullptr};
  double para_Ex_Pose_wheel[1][7], para_Ix_sx_wheel[1][1], para_Ix_sy_wheel[1][1], para_Ix_sw_wheel[1][1], para_Td_wheel[1][1];
  Matrix3d last_R0; Vector3d last_P0;
  FeatureManager f_manager;
  IntegrationBase* pre_integrations[WINDOW_SIZE + 1] = {nullptr};
  MarginalizationPrior last_marginalization_info;

  double para_Pose[WINDOW_SIZE + 1][7];
  double para_SpeedBias[WINDOW_SIZE + 1][9];
  double para_Feature[NUM_OF_F][1];
  double para_Ex_Pose[2][7];
  double para_Td[1][1];
  gf2_solve_summary last_summary;

 private:
  gf2_solver* gf2 = nullptr;
  std::string last_error;
};

struct Point2f { float x = 0, y = 0; };

class FeatureTracker {
 public:
  FeatureTracker();
  ~FeatureTracker();
  void readIntrinsicParameter(const Parameters& p);
  // goodFeaturesToTrack(cur_img, n_pts, MAX_CNT - cur_pts.size(), 0.01, MIN_DIST, mask) (:198) runs on the device through
  // gf2_tracker_detect on the image trackImage has just uploaded. A hook can replace it for A/B checks (tests: cv2 through ctypes).
  typedef int (*Detector)(const uint8_t* img, int rows, int cols, const uint8_t* mask, int max_corners, int min_dist, float* out_xy, void* user);
  void setDetector(Detector d, void* user) { detector = d; detector_user = user; }
  void setPrediction(const std::map<int, Vector3d>& predictPts);   // :1006-1027 (spaceToPlane of the pinhole model)
  void removeOutliers(const std::set<int>& removePtsIds);          // :1029-1045
  // trackImage for the mono (+depth lookup) configuration with hasPrediction == false (feature_tracker.cpp:103-372)
  std::map<int, std::vector<std::pair<int, std::vector<double>>>> trackImage(double _cur_time, const uint8_t* _img, const uint16_t* depth = nullptr);
  bool inBorder(const Point2f& pt) const;        // :14-20
  void setMask();                                // :56-83
  void addPoints();                              // :85-93
  std::vector<Point2f> undistortedPts(const std::vector<Point2f>& pts) const;  // :797-808 + PinholeCamera::liftProjective
  std::vector<Point2f> ptsVelocity(const std::vector<int>& ids, const std::vector<Point2f>& pts, std::map<int, Point2f>& cur_id_pts, std::map<int, Point2f>& prev_id_pts);  // :810-847
  const char* lastError() const { return last_error.c_str(); }

  int row = 480, col = 640, MAX_CNT = 150, MIN_DIST = 30, FLOW_BACK = 1, EQUALIZE = 0;
  double fx = 0, fy = 0, cx = 0, cy = 0, k1 = 0, k2 = 0, p1 = 0, p2 = 0;
  std::vector<uint8_t> mask, cur_img;
  std::vector<Point2f> n_pts, prev_pts, cur_pts, cur_un_pts, prev_un_pts, pts_velocity, predict_pts;
  bool hasPrediction = false;
  std::vector<int> ids, track_cnt;
  std::map<int, Point2f> cur_un_pts_map, prev_un_pts_map;
  double cur_time = 0, prev_time = 0;
  int n_id = 0;
  bool have_prev = false;

 private:
  gf2_tracker* trk = nullptr;
  Detector detector = nullptr; void* detector_user = nullptr;
  std::string last_error;
};

}  // namespace gf2host
