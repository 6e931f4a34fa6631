// TEST INFRASTRUCTURE — C entry points around the REFERENCE's own factor classes, compiled unmodified from /root/reference
// (oracle/Makefile target `_ref`, output oracle/_ref/libgf2_ref.so; Eigen / Ceres / ROS / Sophus / OpenCV headers are the
// stand-ins under oracle/shim). Same signatures as the restated oracle's gf2o_* calls (gf2o_window.cpp), so that
// tests/test_oracle_vs_ref.py can compare the restatement with the reference code on identical inputs:
//   gf2r_factor_eval          ProjectionTwoFrameOneCamFactor / IMUFactor / WheelFactor / LidarPlaneNormFactor / CTLidarPlaneNormFactor ::Evaluate
//   gf2r_imu_preintegrate     IntegrationBase::push_back chain           (VE/factor/integration_base.h:39-167)
//   gf2r_wheel_preintegrate   WheelIntegrationBase::push_back chain      (VE/factor/wheel_integration_base.h:41-178)
//   gf2r_marginalize_window   MarginalizationInfo::{addResidualBlockInfo, preMarginalize, marginalize, getParameterBlocks}
//                             driven as Estimator::optimization() drives them (VE/estimator/estimator.cpp:3394-3690; that call
//                             sequence is restated here because estimator.cpp itself needs the whole ROS / OpenCV stack)
//   gf2r_pose_plus            PoseLocalParameterization::Plus / ComputeJacobian
// Only tests/ may load the library. It never travels into the product.
#include <cstring>
#include <map>
#include <memory>
#include <vector>
#include "factor/projectionTwoFrameOneCamFactor.h"
#include "factor/imu_factor.h"
#include "factor/wheel_factor.h"
#include "factor/marginalization_factor.h"
#include "factor/pose_local_parameterization.h"
#include "lidarFactor.h"
#include "../include/gf2_abi.h"

// ---- the globals of VE/estimator/parameters.cpp the factor sources read (parameters.cpp itself needs OpenCV's FileStorage)
double ACC_N, ACC_W, GYR_N, GYR_W;
double VEL_N_wheel, GYR_N_wheel, SX, SY, SW;
Eigen::Vector3d G{0.0, 0.0, 9.8};
double TD, TD_WHEEL;
int ESTIMATE_EXTRINSIC_WHEEL, ESTIMATE_INTRINSIC_WHEEL, ESTIMATE_TD_WHEEL, USE_WHEEL, USE_IMU = 1, USE_PLANE, ESTIMATE_EXTRINSIC, ESTIMATE_TD;
double ROLL_N, PITCH_N, ZPW_N, ROLL_N_INV, PITCH_N_INV, ZPW_N_INV;
Eigen::Matrix3d RIO;
Eigen::Vector3d TIO;
std::vector<Eigen::Matrix3d> RIC;
std::vector<Eigen::Vector3d> TIC;
CameraExtrinsicAdjustType CAM_EXT_ADJ_TYPE = ADJUST_CAM_ALL;
WheelExtrinsicAdjustType WHEEL_EXT_ADJ_TYPE = ADJUST_WHEEL_ALL;

namespace {
Eigen::Vector3d v3(const double* p) { return Eigen::Vector3d(p[0], p[1], p[2]); }

IntegrationBase* imu_from_record(const gf2_imu_preint& r) {
  IntegrationBase* ib = new IntegrationBase(Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), v3(r.lin_ba), v3(r.lin_bg));
  ib->sum_dt = r.sum_dt; ib->delta_p = v3(r.delta_p); ib->delta_v = v3(r.delta_v);
  ib->delta_q = Eigen::Quaterniond(r.delta_q[3], r.delta_q[0], r.delta_q[1], r.delta_q[2]);
  for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { ib->jacobian(i, j) = r.jacobian[i * 15 + j]; ib->covariance(i, j) = r.covariance[i * 15 + j]; }
  return ib;
}
void imu_to_record(const IntegrationBase& ib, gf2_imu_preint* o) {
  memset(o, 0, sizeof(*o));
  o->sum_dt = ib.sum_dt;
  for (int i = 0; i < 3; i++) { o->delta_p[i] = ib.delta_p(i); o->delta_v[i] = ib.delta_v(i); o->lin_ba[i] = ib.linearized_ba(i); o->lin_bg[i] = ib.linearized_bg(i); }
  o->delta_q[0] = ib.delta_q.x(); o->delta_q[1] = ib.delta_q.y(); o->delta_q[2] = ib.delta_q.z(); o->delta_q[3] = ib.delta_q.w();
  for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { o->jacobian[i * 15 + j] = ib.jacobian(i, j); o->covariance[i * 15 + j] = ib.covariance(i, j); }
  o->valid = 1;
}
WheelIntegrationBase* wheel_from_record(const gf2_wheel_preint& r) {
  WheelIntegrationBase* wb = new WheelIntegrationBase(v3(r.lin_vel), v3(r.lin_gyr), r.lin_sx, r.lin_sy, r.lin_sw, r.lin_td);
  wb->sum_dt = r.sum_dt; wb->delta_p = v3(r.delta_p);
  wb->delta_q = Eigen::Quaterniond(r.delta_q[3], r.delta_q[0], r.delta_q[1], r.delta_q[2]);
  wb->vel_1 = v3(r.vel_1); wb->gyr_1 = v3(r.gyr_1);
  for (int i = 0; i < 6; i++) { for (int j = 0; j < 3; j++) wb->jacobian(i, j) = r.jacobian[i * 3 + j]; for (int j = 0; j < 6; j++) wb->covariance(i, j) = r.covariance[i * 6 + j]; }
  return wb;
}
void wheel_to_record(const WheelIntegrationBase& wb, gf2_wheel_preint* o) {
  memset(o, 0, sizeof(*o));
  o->sum_dt = wb.sum_dt;
  for (int i = 0; i < 3; i++) { o->delta_p[i] = wb.delta_p(i); o->lin_vel[i] = wb.linearized_vel(i); o->lin_gyr[i] = wb.linearized_gyr(i); o->vel_1[i] = wb.vel_1(i); o->gyr_1[i] = wb.gyr_1(i); }
  o->delta_q[0] = wb.delta_q.x(); o->delta_q[1] = wb.delta_q.y(); o->delta_q[2] = wb.delta_q.z(); o->delta_q[3] = wb.delta_q.w();
  o->lin_sx = wb.linearized_sx; o->lin_sy = wb.linearized_sy; o->lin_sw = wb.linearized_sw; o->lin_td = wb.linearized_td;
  for (int i = 0; i < 6; i++) { for (int j = 0; j < 3; j++) o->jacobian[i * 3 + j] = wb.jacobian(i, j); for (int j = 0; j < 6; j++) o->covariance[i * 6 + j] = wb.covariance(i, j); }
  o->valid = 1;
}
}  // namespace

extern "C" {

// the window record of oracle/gf2o_window.cpp (same layout)
typedef struct gf2o_window {
  int32_t n_frames, n_landmarks, n_planes, prior_rows, prior_nblocks, use_wheel, prior_stride, pad_;
  double *para_pose, *para_speedbias, *ex_pose, *td, *ex_pose_wheel, *sxsysw, *td_wheel, *inv_depth;
  const int32_t *start_frame, *track_len;
  const uint8_t* fixed;
  const gf2_obs* obs;
  const double* frame_td;
  const gf2_imu_preint* imu;
  const gf2_wheel_preint* wheel;
  const double *prior_J0, *prior_r0;
  const gf2_prior_block* prior_blocks;
  const gf2_plane* planes;
  const double* plane_alpha;
} gf2o_window;

void gf2r_set_noise(const double imu_noise[4], const double wheel_noise[2]) {
  if (imu_noise) { ACC_N = imu_noise[0]; GYR_N = imu_noise[1]; ACC_W = imu_noise[2]; GYR_W = imu_noise[3]; }
  if (wheel_noise) { VEL_N_wheel = wheel_noise[0]; GYR_N_wheel = wheel_noise[1]; }
}

// kinds and argument layout: oracle/gf2o_window.cpp gf2o_factor_eval (0 projection, 1 IMU, 2 wheel, 3 LiDAR plane, 4 CT LiDAR plane)
int gf2r_factor_eval(int kind, const void* consts, const double* extra, const double* params_flat, double* residuals, double* jac_flat) {
  std::unique_ptr<ceres::CostFunction> f;
  std::unique_ptr<IntegrationBase> ib; std::unique_ptr<WheelIntegrationBase> wb;
  const double* c = (const double*)consts;
  switch (kind) {
    case 0:
      ProjectionTwoFrameOneCamFactor::sqrt_info = c[10] * Eigen::Matrix2d::Identity();   // estimator.cpp:… FOCAL_LENGTH / 1.5 * Matrix2d::Identity()
      f.reset(new ProjectionTwoFrameOneCamFactor(Eigen::Vector3d(c[0], c[1], 1), Eigen::Vector3d(c[5], c[6], 1), Eigen::Vector2d(c[2], c[3]), Eigen::Vector2d(c[7], c[8]), c[4], c[9]));
      break;
    case 1: ib.reset(imu_from_record(*(const gf2_imu_preint*)consts)); G = Eigen::Vector3d(0, 0, extra[0]); f.reset(new IMUFactor(ib.get())); break;
    case 2: wb.reset(wheel_from_record(*(const gf2_wheel_preint*)consts)); f.reset(new WheelFactor(wb.get())); break;
    case 3: CT_ICP::LidarPlaneNormFactor::sqrt_info = c[8]; f.reset(new CT_ICP::LidarPlaneNormFactor(v3(c), v3(c + 3), c[6], c[7])); break;
    case 4:
      CT_ICP::CTLidarPlaneNormFactor::sqrt_info = c[9]; CT_ICP::CTLidarPlaneNormFactor::t_il = Eigen::Vector3d::Zero(); CT_ICP::CTLidarPlaneNormFactor::q_il = Eigen::Quaterniond::Identity();
      f.reset(new CT_ICP::CTLidarPlaneNormFactor(v3(c), v3(c + 3), c[6], c[7], c[8]));
      break;
    default: return -1;
  }
  std::vector<const double*> pp; std::vector<double*> jj; size_t po = 0, jo = 0;
  for (int s : f->parameter_block_sizes()) { pp.push_back(params_flat + po); po += s; jj.push_back(jac_flat ? jac_flat + jo : nullptr); jo += (size_t)s * f->num_residuals(); }
  f->Evaluate(pp.data(), residuals, jac_flat ? jj.data() : nullptr);
  return f->num_residuals();
}

int gf2r_imu_preintegrate(const gf2_imu_sample* samples, int n, const double first[6], const double lin_bias[6], const double noise[4], gf2_imu_preint* out) {
  gf2r_set_noise(noise, nullptr);
  IntegrationBase ib(v3(first), v3(first + 3), v3(lin_bias), v3(lin_bias + 3));
  for (int i = 0; i < n; i++) ib.push_back(samples[i].dt, v3(samples[i].acc), v3(samples[i].gyr));
  imu_to_record(ib, out);
  return 0;
}
int gf2r_wheel_preintegrate(const gf2_wheel_sample* samples, int n, const double first[6], const double lin[4], const double noise[2], gf2_wheel_preint* out) {
  gf2r_set_noise(nullptr, noise);
  WheelIntegrationBase wb(v3(first), v3(first + 3), lin[0], lin[1], lin[2], lin[3]);
  for (int i = 0; i < n; i++) wb.push_back(samples[i].dt, v3(samples[i].vel), v3(samples[i].gyr));
  wheel_to_record(wb, out);
  return 0;
}

// PoseLocalParameterization::Plus (x [7], delta [6] -> x_plus_delta [7]) and ComputeJacobian (7x6 row-major)
void gf2r_pose_plus(const double* x, const double* delta, double* x_plus_delta, double* jacobian) {
  PoseLocalParameterization impl;
  const ceres::LocalParameterization& p = impl;   // the overrides are private in the reference class: call through the Ceres interface, as Ceres does
  p.Plus(x, delta, x_plus_delta);
  if (jacobian) p.ComputeJacobian(x, jacobian);
}

// Marginalization after the solve, see the header comment. Output as gf2o_marginalize_window: J0 [n][P], r0 [n], blocks renamed by addr_shift.
int gf2r_marginalize_window(const gf2o_window* w, const gf2_solve_opts* o, int mode, int P, double* J0, double* r0, int32_t* n_blocks, gf2_prior_block* blocks, int32_t* m_out) {
  const int F = w->n_frames;
  ProjectionTwoFrameOneCamFactor::sqrt_info = o->sqrt_info_px * Eigen::Matrix2d::Identity();
  G = Eigen::Vector3d(0, 0, o->g_norm);
  auto addrOf = [&](const gf2_prior_block& b) -> double* {
    switch (b.kind) {
      case GF2_BLK_POSE: return w->para_pose + 7 * b.index;
      case GF2_BLK_SPEEDBIAS: return w->para_speedbias + 9 * b.index;
      case GF2_BLK_EX_POSE: return w->ex_pose;
      case GF2_BLK_TD: return w->td;
      case GF2_BLK_EX_WHEEL: return w->ex_pose_wheel;
      case GF2_BLK_SX: return w->sxsysw + 0;
      case GF2_BLK_SY: return w->sxsysw + 1;
      case GF2_BLK_SW: return w->sxsysw + 2;
      case GF2_BLK_TD_WHEEL: return w->td_wheel;
    }
    return nullptr;
  };
  // last_marginalization_info: only what MarginalizationFactor reads (n, m, keep_block_*, linearized_*)
  std::unique_ptr<MarginalizationInfo> last;
  std::vector<double*> last_blocks;
  std::vector<std::vector<double>> last_x0;
  if (w->prior_rows > 0) {
    last.reset(new MarginalizationInfo());
    last->m = 0; last->n = w->prior_rows;
    last->linearized_jacobians.resize(last->n, last->n); last->linearized_residuals.resize(last->n);
    for (int r = 0; r < last->n; r++) { last->linearized_residuals(r) = w->prior_r0[r]; for (int c = 0; c < last->n; c++) last->linearized_jacobians(r, c) = w->prior_J0[(size_t)r * w->prior_stride + c]; }
    last_x0.resize(w->prior_nblocks);
    for (int b = 0; b < w->prior_nblocks; b++) {
      const gf2_prior_block& pb = w->prior_blocks[b];
      const int size = (pb.kind == GF2_BLK_POSE || pb.kind == GF2_BLK_EX_POSE || pb.kind == GF2_BLK_EX_WHEEL) ? 7 : (pb.kind == GF2_BLK_SPEEDBIAS ? 9 : 1);
      last_x0[b].assign(pb.x0, pb.x0 + size);
      last->keep_block_size.push_back(size); last->keep_block_idx.push_back(pb.offset); last->keep_block_data.push_back(last_x0[b].data());
      last_blocks.push_back(addrOf(pb));
    }
  }
  MarginalizationInfo* info = new MarginalizationInfo();   // owns its factors (deletes cost functions in its destructor, as in the reference)
  std::unique_ptr<IntegrationBase> imu0; std::unique_ptr<WheelIntegrationBase> wheel0;
  std::vector<std::unique_ptr<ceres::LossFunction>> losses;
  auto add = [&](ceres::CostFunction* f, bool loss, std::vector<double*> blocks_, std::vector<int> drop) {
    ceres::LossFunction* lf = nullptr;
    if (loss) { losses.emplace_back(new ceres::HuberLoss(o->huber_delta)); lf = losses.back().get(); }
    info->addResidualBlockInfo(new ResidualBlockInfo(f, lf, blocks_, drop));
  };
  std::map<double*, std::pair<int, int>> shift;  // address -> (kind, index) after slideWindow
  int rc = 0;
  if (mode == 0) {
    if (last) {  // estimator.cpp:3401-3415
      std::vector<int> drop;
      for (size_t i = 0; i < last_blocks.size(); i++) if (last_blocks[i] == w->para_pose || last_blocks[i] == w->para_speedbias) drop.push_back((int)i);
      add(new MarginalizationFactor(last.get()), false, last_blocks, drop);
    }
    if (w->imu && w->imu[0].valid && w->imu[0].sum_dt < 10.0) {  // :3416-3427
      imu0.reset(imu_from_record(w->imu[0]));
      add(new IMUFactor(imu0.get()), false, {w->para_pose, w->para_speedbias, w->para_pose + 7, w->para_speedbias + 9}, {0, 1});
    }
    if (w->use_wheel && w->wheel && w->wheel[0].valid && w->wheel[0].sum_dt < 10.0) {  // :3428-3439
      wheel0.reset(wheel_from_record(w->wheel[0]));
      add(new WheelFactor(wheel0.get()), false, {w->para_pose, w->para_pose + 7, w->ex_pose_wheel, w->sxsysw, w->sxsysw + 1, w->sxsysw + 2, w->td_wheel}, {0});
    }
    int ob = 0;  // :3495-3528
    for (int l = 0; l < w->n_landmarks; l++) {
      if (w->start_frame[l] == 0) {
        const gf2_obs& oi = w->obs[ob];
        for (int k = 1; k < w->track_len[l]; k++) {
          const gf2_obs& oj = w->obs[ob + k];
          add(new ProjectionTwoFrameOneCamFactor(Eigen::Vector3d(oi.x, oi.y, 1.0), Eigen::Vector3d(oj.x, oj.y, 1.0), Eigen::Vector2d(oi.vx, oi.vy), Eigen::Vector2d(oj.vx, oj.vy),
                                                 w->frame_td[0], w->frame_td[k]),
              true, {w->para_pose, w->para_pose + 7 * k, w->ex_pose, w->inv_depth + l, w->td}, {0, 3});
        }
      }
      ob += w->track_len[l];
    }
    for (int i = 1; i < F; i++) { shift[w->para_pose + 7 * i] = {GF2_BLK_POSE, i - 1}; shift[w->para_speedbias + 9 * i] = {GF2_BLK_SPEEDBIAS, i - 1}; }  // :3561-3570
  } else {
    const int sn = F - 2;  // WINDOW_SIZE - 1
    bool has = false; for (double* p : last_blocks) has |= (p == w->para_pose + 7 * sn);
    if (!last || !has) rc = -2;  // :3599-3600
    else {
      std::vector<int> drop;
      for (size_t i = 0; i < last_blocks.size(); i++) {
        if (last_blocks[i] == w->para_speedbias + 9 * sn) rc = -3;  // ROS_ASSERT, :3612
        if (last_blocks[i] == w->para_pose + 7 * sn) drop.push_back((int)i);
      }
      if (rc == 0) add(new MarginalizationFactor(last.get()), false, last_blocks, drop);
      for (int i = 0; i < F; i++) {  // :3653-3676
        if (i == sn) continue;
        const int to = (i == F - 1) ? i - 1 : i;
        shift[w->para_pose + 7 * i] = {GF2_BLK_POSE, to}; shift[w->para_speedbias + 9 * i] = {GF2_BLK_SPEEDBIAS, to};
      }
    }
  }
  if (rc != 0) { delete info; return rc; }
  shift[w->ex_pose] = {GF2_BLK_EX_POSE, 0}; shift[w->td] = {GF2_BLK_TD, 0};
  if (w->use_wheel) {
    shift[w->ex_pose_wheel] = {GF2_BLK_EX_WHEEL, 0}; shift[w->sxsysw] = {GF2_BLK_SX, 0}; shift[w->sxsysw + 1] = {GF2_BLK_SY, 0};
    shift[w->sxsysw + 2] = {GF2_BLK_SW, 0}; shift[w->td_wheel] = {GF2_BLK_TD_WHEEL, 0};
  }
  info->preMarginalize();
  info->marginalize();
  if (m_out) *m_out = info->m;
  if (!info->valid) { delete info; return -1; }
  if (info->n > P) { delete info; return -3; }
  std::unordered_map<long, double*> addr_shift;
  for (auto& kv : shift) addr_shift[reinterpret_cast<long>(kv.first)] = kv.first;   // identity: the renaming is applied below with (kind, index)
  std::vector<double*> kept = info->getParameterBlocks(addr_shift);
  const int n = info->n;
  for (int r = 0; r < n; r++) { r0[r] = info->linearized_residuals(r); for (int c = 0; c < n; c++) J0[(size_t)r * P + c] = info->linearized_jacobians(r, c); }
  *n_blocks = (int)kept.size();
  for (size_t b = 0; b < kept.size(); b++) {
    gf2_prior_block& pb = blocks[b]; memset(&pb, 0, sizeof(pb));
    auto it = shift.find(kept[b]);
    if (it == shift.end()) { delete info; return -3; }
    pb.kind = it->second.first; pb.index = it->second.second; pb.offset = info->keep_block_idx[b] - info->m;
    for (int k = 0; k < info->keep_block_size[b]; k++) pb.x0[k] = info->keep_block_data[b][k];
  }
  delete info;
  return n;
}

}  // extern "C"
