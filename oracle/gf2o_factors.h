// TEST INFRASTRUCTURE — CPU oracle (see gf2o_linalg.h header). Restatement of the reference's factor
// library behind the same plug-in API the reference implements:
//   bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const
// (ceres::CostFunction; jacobians[k] row-major num_residuals x block_size in AMBIENT coordinates).
// parity: unpinned by reference tests except LidarPlaneNormFactor (LIO/apps/test_analytic_factor.cpp:54-109);
// validated by finite differences with the reference's own check() recipe (tests/test_oracle_factors.py).
#pragma once
#include "gf2o_linalg.h"
#include "../include/gf2_abi.h"

namespace gf2o {

enum { O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12 };  // VE/estimator/parameters.h StateOrder

struct CostFunction {
  std::vector<int> block_sizes;
  int num_residuals = 0;
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
};

// VE/factor/projectionTwoFrameOneCamFactor.{h,cpp} — SizedCostFunction<2, 7, 7, 7, 1, 1>
struct ProjectionTwoFrameOneCamFactor : CostFunction {
  V3 pts_i, pts_j, velocity_i, velocity_j;
  double td_i, td_j;
  static double sqrt_info;  // the reference's static Matrix2d = s * I (VE/estimator/estimator.cpp:193)
  ProjectionTwoFrameOneCamFactor(const V3& pi, const V3& pj, const double vi[2], const double vj[2], double tdi, double tdj);
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;
};

// VE/factor/integration_base.h
struct IntegrationBase {
  double dt;
  V3 acc_0, gyr_0, acc_1, gyr_1;
  V3 linearized_acc, linearized_gyr;
  V3 linearized_ba, linearized_bg;
  Mat<15, 15> jacobian, covariance;
  Mat<18, 18> noise;
  double sum_dt;
  V3 delta_p; Quat delta_q; V3 delta_v;
  static V3 G;  // VE/estimator/parameters.cpp:223
  IntegrationBase(const V3& acc0, const V3& gyr0, const V3& ba, const V3& bg, double acc_n, double gyr_n, double acc_w, double gyr_w);
  IntegrationBase(const gf2_imu_preint& rec);  // rebuild from a packed record (factor evaluation only)
  void push_back(double dt, const V3& acc, const V3& gyr) { propagate(dt, acc, gyr); }
  void propagate(double dt, const V3& acc_1, const V3& gyr_1);
  void midPointIntegration(double _dt, const V3& _acc_0, const V3& _gyr_0, const V3& _acc_1, const V3& _gyr_1,
                           const V3& delta_p, const Quat& delta_q, const V3& delta_v, const V3& linearized_ba,
                           const V3& linearized_bg, V3& result_delta_p, Quat& result_delta_q, V3& result_delta_v,
                           V3& result_linearized_ba, V3& result_linearized_bg, bool update_jacobian);
  Mat<15, 1> evaluate(const V3& Pi, const Quat& Qi, const V3& Vi, const V3& Bai, const V3& Bgi, const V3& Pj,
                      const Quat& Qj, const V3& Vj, const V3& Baj, const V3& Bgj) const;
  void pack(gf2_imu_preint* rec) const;
};

// VE/factor/imu_factor.h — SizedCostFunction<15, 7, 9, 7, 9>
struct IMUFactor : CostFunction {
  const IntegrationBase* pre_integration;
  explicit IMUFactor(const IntegrationBase* p);
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;
};

// VE/factor/wheel_integration_base.h
struct WheelIntegrationBase {
  double dt;
  V3 vel_0, gyr_0, vel_1, gyr_1;
  V3 linearized_vel, linearized_gyr;
  double linearized_sx, linearized_sy, linearized_sw, linearized_td;
  Mat<6, 3> jacobian;
  Mat<6, 6> covariance;
  Mat<12, 12> noise;
  double sum_dt;
  V3 delta_p; Quat delta_q;
  mutable V3 corrected_delta_p;       // evaluate() mutates these, Evaluate reads them afterwards
  mutable Quat corrected_delta_q;     // (wheel_integration_base.h:201-202, wheel_factor.h:199,223)
  WheelIntegrationBase(const V3& vel0, const V3& gyr0, double sx, double sy, double sw, double td, double vel_n, double gyr_n);
  WheelIntegrationBase(const gf2_wheel_preint& rec);
  void push_back(double dt, const V3& vel, const V3& gyr) { propagate(dt, vel, gyr); }
  void propagate(double dt, const V3& vel_1, const V3& gyr_1);
  Mat<6, 1> evaluate(const V3& Pi, const Quat& Qi, const Quat& qio, const V3& tio, double sx, double sy, double sw,
                     const V3& Pj, const Quat& Qj, double td) const;
  void pack(gf2_wheel_preint* rec) const;
};

// VE/factor/wheel_factor.h — SizedCostFunction<6, 7, 7, 7, 1, 1, 1, 1>
struct WheelFactor : CostFunction {
  const WheelIntegrationBase* pre_integration;
  explicit WheelFactor(const WheelIntegrationBase* p);
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;
};

// VE/factor/marginalization_factor.{h,cpp} — MarginalizationInfo (kept part) + MarginalizationFactor
struct MarginalizationInfo {
  int m = 0, n = 0;
  std::vector<int> keep_block_size, keep_block_idx;
  std::vector<std::vector<double>> keep_block_data;
  std::vector<double> linearized_jacobians;  // n x n row-major
  std::vector<double> linearized_residuals;  // n
  bool valid = true;
};
struct MarginalizationFactor : CostFunction {
  const MarginalizationInfo* info;
  explicit MarginalizationFactor(const MarginalizationInfo* i);
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;
};

// LIO/liw/lidarFactor.{h,cpp} — LidarPlaneNormFactor: SizedCostFunction<1, 3, 4>; q stored [x y z w]
struct LidarPlaneNormFactor : CostFunction {
  V3 point_body, norm_vector; double norm_offset, weight;
  static double sqrt_info;
  LidarPlaneNormFactor(const V3& pb, const V3& nv, double off, double w);
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;
};
// CTLidarPlaneNormFactor: SizedCostFunction<1, 3, 4, 3, 4>
struct CTLidarPlaneNormFactor : CostFunction {
  V3 raw_keypoint, norm_vector; double norm_offset, alpha_time, weight;
  static double sqrt_info;
  CTLidarPlaneNormFactor(const V3& kp, const V3& nv, double off, double alpha, double w);
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;
};
// The same LidarPlaneNormFactor residual attached to a VINS 7-block pose [p, qx qy qz qw]
// (BASELINE.json config 4's synthetic composition): one block of size 7, Jacobian 1x7 = [J_t | J_q(3) | 0].
struct LidarPlanePoseFactor : CostFunction {
  LidarPlaneNormFactor inner;
  LidarPlanePoseFactor(const V3& pb, const V3& nv, double off, double w);
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;
};

// CTLidarPlaneNormFactor attached to two VINS 7-block poses (begin = window pose f, end = window pose f + 1): blocks {7, 7},
// Jacobians 1x7 = [J_t | J_q(3) | 0] each.
struct CTLidarPlanePoseFactor : CostFunction {
  CTLidarPlaneNormFactor inner;
  CTLidarPlanePoseFactor(const V3& kp, const V3& nv, double off, double alpha, double w);
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;
};

}  // namespace gf2o
