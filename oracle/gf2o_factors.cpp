// TEST INFRASTRUCTURE — CPU oracle. Line-by-line restatement of the reference factor files (cited per
// function) on top of gf2o_linalg.h instead of Eigen. Nothing here is used by the product path.
#include "gf2o_factors.h"

namespace gf2o {

double ProjectionTwoFrameOneCamFactor::sqrt_info = 400.0;  // FOCAL_LENGTH / 1.5, VE/estimator/estimator.cpp:193
V3 IntegrationBase::G = v3(0, 0, 9.8);
double LidarPlaneNormFactor::sqrt_info = 1.0;
double CTLidarPlaneNormFactor::sqrt_info = 1.0;

static inline V3 P3(const double* p) { return v3(p[0], p[1], p[2]); }
static inline Quat Q7(const double* p) { return Quat(p[6], p[3], p[4], p[5]); }  // Quaterniond(p[6], p[3], p[4], p[5])

template <int R, int C>
static void store(double* dst, const Mat<R, C>& m) { for (int i = 0; i < R * C; i++) dst[i] = m.a[i]; }

// ------------------------------------------------------------------ projection
// VE/factor/projectionTwoFrameOneCamFactor.cpp:16-41 (constructor), :43-151 (Evaluate), non-UNIT_SPHERE branch
ProjectionTwoFrameOneCamFactor::ProjectionTwoFrameOneCamFactor(const V3& pi, const V3& pj, const double vi[2],
                                                               const double vj[2], double tdi, double tdj)
    : pts_i(pi), pts_j(pj), td_i(tdi), td_j(tdj) {
  velocity_i = v3(vi[0], vi[1], 0);
  velocity_j = v3(vj[0], vj[1], 0);
  block_sizes = {7, 7, 7, 1, 1};
  num_residuals = 2;
}

bool ProjectionTwoFrameOneCamFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
  V3 Pi = P3(parameters[0]); Quat Qi = Q7(parameters[0]);
  V3 Pj = P3(parameters[1]); Quat Qj = Q7(parameters[1]);
  V3 tic = P3(parameters[2]); Quat qic = Q7(parameters[2]);
  double inv_dep_i = parameters[3][0];
  double td = parameters[4][0];

  V3 pts_i_td = pts_i - (td - td_i) * velocity_i;
  V3 pts_j_td = pts_j - (td - td_j) * velocity_j;
  V3 pts_camera_i = pts_i_td / inv_dep_i;
  V3 pts_imu_i = qic * pts_camera_i + tic;
  V3 pts_w = Qi * pts_imu_i + Pi;
  V3 pts_imu_j = Qj.inverse() * (pts_w - Pj);
  V3 pts_camera_j = qic.inverse() * (pts_imu_j - tic);

  double dep_j = pts_camera_j[2];
  residuals[0] = sqrt_info * (pts_camera_j[0] / dep_j - pts_j_td[0]);
  residuals[1] = sqrt_info * (pts_camera_j[1] / dep_j - pts_j_td[1]);

  if (jacobians) {
    M3 Ri = Qi.toRotationMatrix(), Rj = Qj.toRotationMatrix(), ric = qic.toRotationMatrix();
    Mat<2, 3> reduce;
    reduce(0, 0) = 1. / dep_j; reduce(0, 1) = 0; reduce(0, 2) = -pts_camera_j[0] / (dep_j * dep_j);
    reduce(1, 0) = 0; reduce(1, 1) = 1. / dep_j; reduce(1, 2) = -pts_camera_j[1] / (dep_j * dep_j);
    reduce = sqrt_info * reduce;

    if (jacobians[0]) {
      Mat<3, 6> jaco_i;
      jaco_i.setBlock<3, 3>(0, 0, ric.T() * Rj.T());
      jaco_i.setBlock<3, 3>(0, 3, ric.T() * Rj.T() * Ri * (-skew(pts_imu_i)));
      Mat<2, 6> j6 = reduce * jaco_i;
      Mat<2, 7> J; J.setBlock<2, 6>(0, 0, j6);
      store(jacobians[0], J);
    }
    if (jacobians[1]) {
      Mat<3, 6> jaco_j;
      jaco_j.setBlock<3, 3>(0, 0, ric.T() * (-Rj.T()));
      jaco_j.setBlock<3, 3>(0, 3, ric.T() * skew(pts_imu_j));
      Mat<2, 6> j6 = reduce * jaco_j;
      Mat<2, 7> J; J.setBlock<2, 6>(0, 0, j6);
      store(jacobians[1], J);
    }
    if (jacobians[2]) {
      Mat<3, 6> jaco_ex;
      jaco_ex.setBlock<3, 3>(0, 0, ric.T() * (Rj.T() * Ri - M3::Identity()));
      M3 tmp_r = ric.T() * Rj.T() * Ri * ric;
      jaco_ex.setBlock<3, 3>(0, 3, -(tmp_r * skew(pts_camera_i)) + skew(tmp_r * pts_camera_i) +
                                       skew(ric.T() * (Rj.T() * (Ri * tic + Pi - Pj) - tic)));
      Mat<2, 6> j6 = reduce * jaco_ex;
      Mat<2, 7> J; J.setBlock<2, 6>(0, 0, j6);
      store(jacobians[2], J);
    }
    if (jacobians[3]) {
      Mat<2, 1> jf = reduce * (ric.T() * Rj.T() * Ri * ric * pts_i_td) * (-1.0 / (inv_dep_i * inv_dep_i));
      store(jacobians[3], jf);
    }
    if (jacobians[4]) {
      Mat<2, 1> jt = reduce * (ric.T() * Rj.T() * Ri * ric * velocity_i) * (1.0 / inv_dep_i) * -1.0;
      jt[0] += sqrt_info * velocity_j[0];
      jt[1] += sqrt_info * velocity_j[1];
      store(jacobians[4], jt);
    }
  }
  return true;
}

// ------------------------------------------------------------------ IMU preintegration
// VE/factor/integration_base.h:22-37 (constructor, noise)
IntegrationBase::IntegrationBase(const V3& acc0, const V3& gyr0, const V3& ba, const V3& bg, double acc_n, double gyr_n,
                                 double acc_w, double gyr_w)
    : dt(0), acc_0(acc0), gyr_0(gyr0), linearized_acc(acc0), linearized_gyr(gyr0), linearized_ba(ba), linearized_bg(bg),
      jacobian(Mat<15, 15>::Identity()), sum_dt(0.0) {
  for (int i = 0; i < 3; i++) {
    noise(i, i) = acc_n * acc_n; noise(3 + i, 3 + i) = gyr_n * gyr_n;
    noise(6 + i, 6 + i) = acc_n * acc_n; noise(9 + i, 9 + i) = gyr_n * gyr_n;
    noise(12 + i, 12 + i) = acc_w * acc_w; noise(15 + i, 15 + i) = gyr_w * gyr_w;
  }
}
IntegrationBase::IntegrationBase(const gf2_imu_preint& r) : dt(0), sum_dt(r.sum_dt) {
  delta_p = P3(r.delta_p); delta_q = Quat::fromXYZW(r.delta_q); delta_v = P3(r.delta_v);
  linearized_ba = P3(r.lin_ba); linearized_bg = P3(r.lin_bg);
  for (int i = 0; i < 225; i++) { jacobian.a[i] = r.jacobian[i]; covariance.a[i] = r.covariance[i]; }
}
void IntegrationBase::pack(gf2_imu_preint* r) const {
  std::memset(r, 0, sizeof(*r));
  r->sum_dt = sum_dt;
  for (int i = 0; i < 3; i++) { r->delta_p[i] = delta_p[i]; r->delta_v[i] = delta_v[i]; r->lin_ba[i] = linearized_ba[i]; r->lin_bg[i] = linearized_bg[i]; }
  r->delta_q[0] = delta_q.x; r->delta_q[1] = delta_q.y; r->delta_q[2] = delta_q.z; r->delta_q[3] = delta_q.w;
  for (int i = 0; i < 225; i++) { r->jacobian[i] = jacobian.a[i]; r->covariance[i] = covariance.a[i]; }
  r->valid = 1;
}

// VE/factor/integration_base.h:63-137
void IntegrationBase::midPointIntegration(double _dt, const V3& _acc_0, const V3& _gyr_0, const V3& _acc_1, const V3& _gyr_1,
                                          const V3& delta_p, const Quat& delta_q, const V3& delta_v, const V3& linearized_ba,
                                          const V3& linearized_bg, V3& result_delta_p, Quat& result_delta_q,
                                          V3& result_delta_v, V3& result_linearized_ba, V3& result_linearized_bg,
                                          bool update_jacobian) {
  V3 un_acc_0 = delta_q * (_acc_0 - linearized_ba);
  V3 un_gyr = 0.5 * (_gyr_0 + _gyr_1) - linearized_bg;
  result_delta_q = delta_q * Quat(1, un_gyr[0] * _dt / 2, un_gyr[1] * _dt / 2, un_gyr[2] * _dt / 2);
  V3 un_acc_1 = result_delta_q * (_acc_1 - linearized_ba);
  V3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
  result_delta_p = delta_p + delta_v * _dt + 0.5 * un_acc * _dt * _dt;
  result_delta_v = delta_v + un_acc * _dt;
  result_linearized_ba = linearized_ba;
  result_linearized_bg = linearized_bg;

  if (update_jacobian) {
    V3 w_x = 0.5 * (_gyr_0 + _gyr_1) - linearized_bg;
    V3 a_0_x = _acc_0 - linearized_ba;
    V3 a_1_x = _acc_1 - linearized_ba;
    M3 R_w_x = skew(w_x), R_a_0_x = skew(a_0_x), R_a_1_x = skew(a_1_x);
    M3 I = M3::Identity();
    // NB: result_delta_q is NOT normalised here (toRotationMatrix of the un-normalised product), as in the reference.
    M3 dR = delta_q.toRotationMatrix(), rR = result_delta_q.toRotationMatrix();

    Mat<15, 15> F;
    F.setBlock<3, 3>(0, 0, I);
    F.setBlock<3, 3>(0, 3, -0.25 * dR * R_a_0_x * _dt * _dt + -0.25 * rR * R_a_1_x * (I - R_w_x * _dt) * _dt * _dt);
    F.setBlock<3, 3>(0, 6, I * _dt);
    F.setBlock<3, 3>(0, 9, -0.25 * (dR + rR) * _dt * _dt);
    F.setBlock<3, 3>(0, 12, -0.25 * rR * R_a_1_x * _dt * _dt * -_dt);
    F.setBlock<3, 3>(3, 3, I - R_w_x * _dt);
    F.setBlock<3, 3>(3, 12, -1.0 * I * _dt);
    F.setBlock<3, 3>(6, 3, -0.5 * dR * R_a_0_x * _dt + -0.5 * rR * R_a_1_x * (I - R_w_x * _dt) * _dt);
    F.setBlock<3, 3>(6, 6, I);
    F.setBlock<3, 3>(6, 9, -0.5 * (dR + rR) * _dt);
    F.setBlock<3, 3>(6, 12, -0.5 * rR * R_a_1_x * _dt * -_dt);
    F.setBlock<3, 3>(9, 9, I);
    F.setBlock<3, 3>(12, 12, I);

    Mat<15, 18> V;
    V.setBlock<3, 3>(0, 0, 0.25 * dR * _dt * _dt);
    M3 v03 = 0.25 * (-rR) * R_a_1_x * _dt * _dt * 0.5 * _dt;
    V.setBlock<3, 3>(0, 3, v03);
    V.setBlock<3, 3>(0, 6, 0.25 * rR * _dt * _dt);
    V.setBlock<3, 3>(0, 9, v03);
    V.setBlock<3, 3>(3, 3, 0.5 * I * _dt);
    V.setBlock<3, 3>(3, 9, 0.5 * I * _dt);
    V.setBlock<3, 3>(6, 0, 0.5 * dR * _dt);
    M3 v63 = 0.5 * (-rR) * R_a_1_x * _dt * 0.5 * _dt;
    V.setBlock<3, 3>(6, 3, v63);
    V.setBlock<3, 3>(6, 6, 0.5 * rR * _dt);
    V.setBlock<3, 3>(6, 9, v63);
    V.setBlock<3, 3>(9, 12, I * _dt);
    V.setBlock<3, 3>(12, 15, I * _dt);

    jacobian = F * jacobian;
    covariance = F * covariance * F.T() + V * noise * V.T();
  }
}

// VE/factor/integration_base.h:139-167
void IntegrationBase::propagate(double _dt, const V3& _acc_1, const V3& _gyr_1) {
  dt = _dt; acc_1 = _acc_1; gyr_1 = _gyr_1;
  V3 rp, rv, rba, rbg; Quat rq;
  midPointIntegration(_dt, acc_0, gyr_0, _acc_1, _gyr_1, delta_p, delta_q, delta_v, linearized_ba, linearized_bg, rp, rq, rv,
                      rba, rbg, true);
  delta_p = rp; delta_q = rq; delta_v = rv; linearized_ba = rba; linearized_bg = rbg;
  delta_q.normalize();
  sum_dt += dt;
  acc_0 = acc_1; gyr_0 = gyr_1;
}

// VE/factor/integration_base.h:169-195
Mat<15, 1> IntegrationBase::evaluate(const V3& Pi, const Quat& Qi, const V3& Vi, const V3& Bai, const V3& Bgi, const V3& Pj,
                                     const Quat& Qj, const V3& Vj, const V3& Baj, const V3& Bgj) const {
  Mat<15, 1> residuals;
  M3 dp_dba = jacobian.block<3, 3>(O_P, O_BA), dp_dbg = jacobian.block<3, 3>(O_P, O_BG);
  M3 dq_dbg = jacobian.block<3, 3>(O_R, O_BG);
  M3 dv_dba = jacobian.block<3, 3>(O_V, O_BA), dv_dbg = jacobian.block<3, 3>(O_V, O_BG);
  V3 dba = Bai - linearized_ba, dbg = Bgi - linearized_bg;
  Quat corrected_delta_q = delta_q * deltaQ(dq_dbg * dbg);
  V3 corrected_delta_v = delta_v + dv_dba * dba + dv_dbg * dbg;
  V3 corrected_delta_p = delta_p + dp_dba * dba + dp_dbg * dbg;
  V3 rp = Qi.inverse() * (0.5 * G * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt) - corrected_delta_p;
  V3 rq = 2.0 * (corrected_delta_q.inverse() * (Qi.inverse() * Qj)).vec();
  V3 rv = Qi.inverse() * (G * sum_dt + Vj - Vi) - corrected_delta_v;
  V3 rba = Baj - Bai, rbg = Bgj - Bgi;
  for (int i = 0; i < 3; i++) { residuals[O_P + i] = rp[i]; residuals[O_R + i] = rq[i]; residuals[O_V + i] = rv[i]; residuals[O_BA + i] = rba[i]; residuals[O_BG + i] = rbg[i]; }
  return residuals;
}

// ------------------------------------------------------------------ IMU factor
IMUFactor::IMUFactor(const IntegrationBase* p) : pre_integration(p) { block_sizes = {7, 9, 7, 9}; num_residuals = 15; }

// VE/factor/imu_factor.h:28-191
bool IMUFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
  V3 Pi = P3(parameters[0]); Quat Qi = Q7(parameters[0]);
  V3 Vi = P3(parameters[1]), Bai = P3(parameters[1] + 3), Bgi = P3(parameters[1] + 6);
  V3 Pj = P3(parameters[2]); Quat Qj = Q7(parameters[2]);
  V3 Vj = P3(parameters[3]), Baj = P3(parameters[3] + 3), Bgj = P3(parameters[3] + 6);
  const V3& G = IntegrationBase::G;

  Mat<15, 1> residual = pre_integration->evaluate(Pi, Qi, Vi, Bai, Bgi, Pj, Qj, Vj, Baj, Bgj);
  Mat<15, 15> sqrt_info;
  sqrtInfoFromCov<15>(pre_integration->covariance.a, sqrt_info);
  residual = sqrt_info * residual;
  store(residuals, residual);

  if (jacobians) {
    double sum_dt = pre_integration->sum_dt;
    M3 dp_dba = pre_integration->jacobian.block<3, 3>(O_P, O_BA), dp_dbg = pre_integration->jacobian.block<3, 3>(O_P, O_BG);
    M3 dq_dbg = pre_integration->jacobian.block<3, 3>(O_R, O_BG);
    M3 dv_dba = pre_integration->jacobian.block<3, 3>(O_V, O_BA), dv_dbg = pre_integration->jacobian.block<3, 3>(O_V, O_BG);
    M3 RiT = Qi.inverse().toRotationMatrix();

    if (jacobians[0]) {
      Mat<15, 7> J;
      J.setBlock<3, 3>(O_P, O_P, -RiT);
      J.setBlock<3, 3>(O_P, O_R, skew(Qi.inverse() * (0.5 * G * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt)));
      Quat corrected_delta_q = pre_integration->delta_q * deltaQ(dq_dbg * (Bgi - pre_integration->linearized_bg));
      J.setBlock<3, 3>(O_R, O_R, -((Qleft(Qj.inverse() * Qi) * Qright(corrected_delta_q)).block<3, 3>(1, 1)));
      J.setBlock<3, 3>(O_V, O_R, skew(Qi.inverse() * (G * sum_dt + Vj - Vi)));
      J = sqrt_info * J;
      store(jacobians[0], J);
    }
    if (jacobians[1]) {
      Mat<15, 9> J;
      J.setBlock<3, 3>(O_P, O_V - O_V, -RiT * sum_dt);
      J.setBlock<3, 3>(O_P, O_BA - O_V, -dp_dba);
      J.setBlock<3, 3>(O_P, O_BG - O_V, -dp_dbg);
      // uses the UNcorrected delta_q (imu_factor.h:137)
      J.setBlock<3, 3>(O_R, O_BG - O_V, -(Qleft(Qj.inverse() * Qi * pre_integration->delta_q).block<3, 3>(1, 1)) * dq_dbg);
      J.setBlock<3, 3>(O_V, O_V - O_V, -RiT);
      J.setBlock<3, 3>(O_V, O_BA - O_V, -dv_dba);
      J.setBlock<3, 3>(O_V, O_BG - O_V, -dv_dbg);
      J.setBlock<3, 3>(O_BA, O_BA - O_V, -M3::Identity());
      J.setBlock<3, 3>(O_BG, O_BG - O_V, -M3::Identity());
      J = sqrt_info * J;
      store(jacobians[1], J);
    }
    if (jacobians[2]) {
      Mat<15, 7> J;
      J.setBlock<3, 3>(O_P, O_P, RiT);
      Quat corrected_delta_q = pre_integration->delta_q * deltaQ(dq_dbg * (Bgi - pre_integration->linearized_bg));
      J.setBlock<3, 3>(O_R, O_R, Qleft(corrected_delta_q.inverse() * Qi.inverse() * Qj).block<3, 3>(1, 1));
      J = sqrt_info * J;
      store(jacobians[2], J);
    }
    if (jacobians[3]) {
      Mat<15, 9> J;
      J.setBlock<3, 3>(O_V, O_V - O_V, RiT);
      J.setBlock<3, 3>(O_BA, O_BA - O_V, M3::Identity());
      J.setBlock<3, 3>(O_BG, O_BG - O_V, M3::Identity());
      J = sqrt_info * J;
      store(jacobians[3], J);
    }
  }
  return true;
}

// ------------------------------------------------------------------ wheel preintegration
// VE/factor/wheel_integration_base.h:23-39
WheelIntegrationBase::WheelIntegrationBase(const V3& vel0, const V3& gyr0, double sx, double sy, double sw, double td,
                                           double vel_n, double gyr_n)
    : dt(0), vel_0(vel0), gyr_0(gyr0), linearized_vel(vel0), linearized_gyr(gyr0), linearized_sx(sx), linearized_sy(sy),
      linearized_sw(sw), linearized_td(td), sum_dt(0.0) {
  for (int i = 0; i < 3; i++) {
    noise(i, i) = vel_n * vel_n; noise(3 + i, 3 + i) = gyr_n * gyr_n;
    noise(6 + i, 6 + i) = vel_n * vel_n; noise(9 + i, 9 + i) = gyr_n * gyr_n;
  }
}
WheelIntegrationBase::WheelIntegrationBase(const gf2_wheel_preint& r) : dt(0), sum_dt(r.sum_dt) {
  delta_p = P3(r.delta_p); delta_q = Quat::fromXYZW(r.delta_q);
  linearized_sx = r.lin_sx; linearized_sy = r.lin_sy; linearized_sw = r.lin_sw; linearized_td = r.lin_td;
  linearized_vel = P3(r.lin_vel); linearized_gyr = P3(r.lin_gyr); vel_1 = P3(r.vel_1); gyr_1 = P3(r.gyr_1);
  for (int i = 0; i < 18; i++) jacobian.a[i] = r.jacobian[i];
  for (int i = 0; i < 36; i++) covariance.a[i] = r.covariance[i];
}
void WheelIntegrationBase::pack(gf2_wheel_preint* r) const {
  std::memset(r, 0, sizeof(*r));
  r->sum_dt = sum_dt;
  for (int i = 0; i < 3; i++) { r->delta_p[i] = delta_p[i]; r->lin_vel[i] = linearized_vel[i]; r->lin_gyr[i] = linearized_gyr[i]; r->vel_1[i] = vel_1[i]; r->gyr_1[i] = gyr_1[i]; }
  r->delta_q[0] = delta_q.x; r->delta_q[1] = delta_q.y; r->delta_q[2] = delta_q.z; r->delta_q[3] = delta_q.w;
  r->lin_sx = linearized_sx; r->lin_sy = linearized_sy; r->lin_sw = linearized_sw; r->lin_td = linearized_td;
  for (int i = 0; i < 18; i++) r->jacobian[i] = jacobian.a[i];
  for (int i = 0; i < 36; i++) r->covariance[i] = covariance.a[i];
  r->valid = 1;
}

// VE/factor/wheel_integration_base.h:67-146 (midPointIntegration) + :148-178 (propagate)
void WheelIntegrationBase::propagate(double _dt, const V3& _vel_1, const V3& _gyr_1) {
  dt = _dt; vel_1 = _vel_1; gyr_1 = _gyr_1;
  const V3 _vel_0 = vel_0, _gyr_0 = gyr_0;
  M3 sv = diag3(linearized_sx, linearized_sy, 1);
  V3 un_vel_0 = delta_q * (sv * _vel_0);
  V3 un_gyr = 0.5 * linearized_sw * (_gyr_0 + _gyr_1);
  Quat delta_delta_q(1, un_gyr[0] * _dt / 2, un_gyr[1] * _dt / 2, un_gyr[2] * _dt / 2);
  Quat result_delta_q = delta_q * delta_delta_q;
  V3 un_vel_1 = result_delta_q * (sv * _vel_1);
  V3 un_vel = 0.5 * (un_vel_0 + un_vel_1);
  V3 result_delta_p = delta_p + un_vel * _dt;

  {
    V3 vel_0_x = sv * _vel_0, vel_1_x = sv * _vel_1;
    M3 R_vel_0_x = skew(vel_0_x), R_vel_1_x = skew(vel_1_x);
    M3 dR = delta_q.toRotationMatrix(), rR = result_delta_q.toRotationMatrix(), ddR = delta_delta_q.toRotationMatrix();
    Mat<6, 6> F;
    F.setBlock<3, 3>(0, 0, M3::Identity());
    F.setBlock<3, 3>(0, 3, -0.5 * _dt * (dR * R_vel_0_x + rR * R_vel_1_x * ddR.T()));
    F.setBlock<3, 3>(3, 3, ddR.T());
    M3 Jr = rightJacobianSO3(un_gyr * _dt);
    Mat<6, 12> V;
    V.setBlock<3, 3>(0, 0, 0.5 * _dt * dR * sv);
    V.setBlock<3, 3>(0, 3, -0.25 * _dt * _dt * rR * R_vel_1_x * Jr);
    V.setBlock<3, 3>(0, 6, 0.5 * _dt * rR * sv);
    V.setBlock<3, 3>(0, 9, -0.25 * _dt * _dt * rR * R_vel_1_x * Jr);
    V.setBlock<3, 3>(3, 3, 0.5 * Jr * linearized_sw * _dt);
    V.setBlock<3, 3>(3, 9, 0.5 * Jr * linearized_sw * _dt);
    M3 I1 = diag3(1, 0, 0), I2 = diag3(0, 1, 0);
    V3 c0 = 0.5 * (dR * (I1 * _vel_0) + rR * (I1 * _vel_1)) * _dt;
    V3 c1 = 0.5 * (dR * (I2 * _vel_0) + rR * (I2 * _vel_1)) * _dt;
    V3 dr_dsw_last = v3(jacobian(3, 2), jacobian(4, 2), jacobian(5, 2));
    V3 dr_new = dr_dsw_last + Jr * (0.5 * (_gyr_0 + _gyr_1)) * _dt;
    V3 c2 = 0.5 * (dR * (skew(dr_dsw_last) * (sv * _vel_0)) + rR * (skew(dr_new) * (sv * _vel_1))) * _dt;
    for (int i = 0; i < 3; i++) { jacobian(i, 0) += c0[i]; jacobian(i, 1) += c1[i]; jacobian(3 + i, 2) = dr_new[i]; jacobian(i, 2) += c2[i]; }
    covariance = F * covariance * F.T() + V * noise * V.T();
  }
  delta_p = result_delta_p; delta_q = result_delta_q;
  delta_q.normalize();
  sum_dt += dt;
  vel_0 = vel_1; gyr_0 = gyr_1;
}

// VE/factor/wheel_integration_base.h:180-219
Mat<6, 1> WheelIntegrationBase::evaluate(const V3& Pi, const Quat& Qi, const Quat& qio, const V3& tio, double sx, double sy,
                                         double sw, const V3& Pj, const Quat& Qj, double td) const {
  Mat<6, 1> residuals;
  V3 dp_dsx = v3(jacobian(0, 0), jacobian(1, 0), jacobian(2, 0));
  V3 dp_dsy = v3(jacobian(0, 1), jacobian(1, 1), jacobian(2, 1));
  V3 dp_dsw = v3(jacobian(0, 2), jacobian(1, 2), jacobian(2, 2));
  V3 dq_dsw = v3(jacobian(3, 2), jacobian(4, 2), jacobian(5, 2));
  double dsx = sx - linearized_sx, dsy = sy - linearized_sy, dsw = sw - linearized_sw;
  M3 sv = diag3(sx, sy, 1);
  M3 Ri = Qi.toRotationMatrix(), Rj = Qj.toRotationMatrix(), rio = qio.toRotationMatrix();

  corrected_delta_p = delta_p + dp_dsx * dsx + dp_dsy * dsy + dp_dsw * dsw;
  corrected_delta_q = (delta_q.normalized() * so3Exp(dq_dsw * dsw)).normalized();
  double dtd = td - linearized_td;
  Quat delta_q_time = (so3Exp(sw * linearized_gyr * dtd) * corrected_delta_q.normalized() * so3Exp(-sw * gyr_1 * dtd)).normalized();
  V3 delta_p_time = so3Exp(sw * linearized_gyr * dtd).toRotationMatrix() *
                    (sv * linearized_vel * dtd + corrected_delta_p - corrected_delta_q * (sv * vel_1 * dtd));
  V3 rp = (Ri * rio).T() * (Rj * tio + Pj - Ri * tio - Pi) - delta_p_time;
  V3 rr = so3Log(delta_q_time.inverse() * (Qi * qio).inverse() * Qj * qio);
  for (int i = 0; i < 3; i++) { residuals[i] = rp[i]; residuals[3 + i] = rr[i]; }
  return residuals;
}

// ------------------------------------------------------------------ wheel factor
WheelFactor::WheelFactor(const WheelIntegrationBase* p) : pre_integration(p) { block_sizes = {7, 7, 7, 1, 1, 1, 1}; num_residuals = 6; }

// VE/factor/wheel_factor.h:28-247
bool WheelFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
  V3 Pi = P3(parameters[0]); Quat Qi = Q7(parameters[0]);
  V3 Pj = P3(parameters[1]); Quat Qj = Q7(parameters[1]);
  V3 tio = P3(parameters[2]); Quat qio = Q7(parameters[2]);
  double sx = parameters[3][0], sy = parameters[4][0], sw = parameters[5][0];
  M3 sv = diag3(sx, sy, 1);
  double td = parameters[6][0];
  const WheelIntegrationBase* pre = pre_integration;

  Mat<6, 1> residual = pre->evaluate(Pi, Qi, qio, tio, sx, sy, sw, Pj, Qj, td);
  Mat<6, 1> raw_residual = residual;
  Mat<6, 6> sqrt_info;
  sqrtInfoFromCov<6>(pre->covariance.a, sqrt_info);
  residual = sqrt_info * residual;
  store(residuals, residual);

  if (jacobians) {
    V3 dp_dsx = v3(pre->jacobian(0, 0), pre->jacobian(1, 0), pre->jacobian(2, 0));
    V3 dp_dsy = v3(pre->jacobian(0, 1), pre->jacobian(1, 1), pre->jacobian(2, 1));
    V3 dp_dsw = v3(pre->jacobian(0, 2), pre->jacobian(1, 2), pre->jacobian(2, 2));
    V3 dq_dsw = v3(pre->jacobian(3, 2), pre->jacobian(4, 2), pre->jacobian(5, 2));
    double dtd = td - pre->linearized_td;
    V3 raw_residual_r = v3(raw_residual[3], raw_residual[4], raw_residual[5]);
    M3 Jr_delta_q_inv = rightJacobianInvSO3(raw_residual_r);
    V3 drdsw = dq_dsw * (sw - pre->linearized_sw);
    M3 Jr_drdsw = rightJacobianSO3(drdsw);
    M3 ri = Qi.toRotationMatrix(), rj = Qj.toRotationMatrix(), rio = qio.toRotationMatrix();
    const Quat& cdq = pre->corrected_delta_q;
    const V3& cdp = pre->corrected_delta_p;

    if (jacobians[0]) {
      Mat<6, 7> J;
      J.setBlock<3, 3>(O_P, O_P, -((Qi * qio).inverse().toRotationMatrix()));
      J.setBlock<3, 3>(O_P, O_R, (ri * rio).T() * (ri * skew(tio)) + rio.T() * skew(ri.T() * (rj * tio + Pj - ri * tio - Pi)));
      J.setBlock<3, 3>(O_R, O_R, -(Jr_delta_q_inv * ((Qj * qio).inverse() * Qi).toRotationMatrix()));
      J = sqrt_info * J;
      store(jacobians[0], J);
    }
    if (jacobians[1]) {
      Mat<6, 7> J;
      J.setBlock<3, 3>(O_P, O_P, (Qi * qio).inverse().toRotationMatrix());
      J.setBlock<3, 3>(O_P, O_R, -(((Qi * qio).inverse() * Qj).toRotationMatrix() * skew(tio)));
      J.setBlock<3, 3>(O_R, O_R, Jr_delta_q_inv * qio.inverse().toRotationMatrix());
      J = sqrt_info * J;
      store(jacobians[1], J);
    }
    if (jacobians[2]) {
      Mat<6, 7> J;
      J.setBlock<3, 3>(O_P, O_P, (Qi * qio).inverse().toRotationMatrix() * (Qj.toRotationMatrix() - Qi.toRotationMatrix()));
      J.setBlock<3, 3>(O_P, O_R, skew((Qi * qio).inverse() * (Qj * tio + Pj - Qi * tio - Pi)));
      J.setBlock<3, 3>(O_R, O_R, Jr_delta_q_inv * (M3::Identity() - ((Qj * qio).inverse() * Qi * qio).toRotationMatrix()));
      J = sqrt_info * J;
      store(jacobians[2], J);
    }
    V3 forward_compensate_w = sw * pre->linearized_gyr * dtd;
    V3 forward_compensate_v = sv * pre->linearized_vel * dtd;
    V3 back_compensate_v = sv * pre->vel_1 * dtd;
    V3 back_compensate_w = sw * pre->gyr_1 * dtd;
    M3 Jrtd = rightJacobianSO3(forward_compensate_w);
    M3 Jr_minus_td = rightJacobianSO3(-forward_compensate_w);
    M3 I1 = diag3(1, 0, 0), I2 = diag3(0, 1, 0);
    M3 cdqR = cdq.toRotationMatrix();
    if (jacobians[3]) {
      Mat<6, 1> J;
      V3 t = -(so3Exp(forward_compensate_v).toRotationMatrix() *
               (I1 * pre->linearized_vel * dtd + dp_dsx - cdqR * (I1 * pre->vel_1) * dtd));
      for (int i = 0; i < 3; i++) J[i] = t[i];
      J = sqrt_info * J;
      store(jacobians[3], J);
    }
    if (jacobians[4]) {
      Mat<6, 1> J;
      V3 t = -(so3Exp(forward_compensate_v).toRotationMatrix() *
               (I2 * pre->linearized_vel * dtd + dp_dsy - cdqR * (I2 * pre->vel_1) * dtd));
      for (int i = 0; i < 3; i++) J[i] = t[i];
      J = sqrt_info * J;
      store(jacobians[4], J);
    }
    if (jacobians[5]) {
      Mat<6, 1> J;
      V3 tp = -(so3Exp(forward_compensate_w).toRotationMatrix() *
                (dp_dsw - cdqR * (skew(Jr_drdsw * dq_dsw) * (sv * pre->vel_1 * dtd)) +
                 skew(Jrtd * pre->linearized_gyr * dtd) * (forward_compensate_v + cdp - cdq * back_compensate_v)));
      V3 tr = -(Jr_delta_q_inv * so3Exp(-raw_residual_r).toRotationMatrix() * so3Exp(back_compensate_w).toRotationMatrix() *
                (cdq.inverse().toRotationMatrix() * (Jrtd * pre->linearized_gyr * dtd) + Jr_drdsw * dq_dsw));
      for (int i = 0; i < 3; i++) { J[i] = tp[i]; J[3 + i] = tr[i]; }
      J = sqrt_info * J;
      store(jacobians[5], J);
    }
    if (jacobians[6]) {
      Mat<6, 1> J;
      V3 tp = -(so3Exp(forward_compensate_w).toRotationMatrix() *
                (sv * pre->linearized_vel - cdqR * (sv * pre->vel_1) +
                 skew(Jrtd * (sw * pre->linearized_gyr)) * (forward_compensate_v + cdp - cdqR * back_compensate_v)));
      V3 tr = -(Jr_delta_q_inv * so3Exp(-raw_residual_r).toRotationMatrix() *
                (so3Exp(back_compensate_w).toRotationMatrix() * cdq.inverse().toRotationMatrix() * (Jrtd * (sw * pre->linearized_gyr)) -
                 Jr_minus_td * (sw * pre->gyr_1)));
      for (int i = 0; i < 3; i++) { J[i] = tp[i]; J[3 + i] = tr[i]; }
      J = sqrt_info * J;
      store(jacobians[6], J);
    }
  }
  return true;
}

// ------------------------------------------------------------------ marginalization prior
MarginalizationFactor::MarginalizationFactor(const MarginalizationInfo* i) : info(i) {
  for (int s : i->keep_block_size) block_sizes.push_back(s);
  num_residuals = i->n;
}

// VE/factor/marginalization_factor.cpp:344-392
bool MarginalizationFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
  int n = info->n, m = info->m;
  std::vector<double> dx(n, 0.0);
  for (size_t i = 0; i < info->keep_block_size.size(); i++) {
    int size = info->keep_block_size[i];
    int idx = info->keep_block_idx[i] - m;
    const double* x = parameters[i];
    const double* x0 = info->keep_block_data[i].data();
    if (size != 7) {
      for (int k = 0; k < size; k++) dx[idx + k] = x[k] - x0[k];
    } else {
      for (int k = 0; k < 3; k++) dx[idx + k] = x[k] - x0[k];
      Quat dq = Quat(x0[6], x0[3], x0[4], x0[5]).inverse() * Quat(x[6], x[3], x[4], x[5]);
      V3 v = 2.0 * dq.vec();
      if (!(dq.w >= 0)) v = -v;
      for (int k = 0; k < 3; k++) dx[idx + 3 + k] = v[k];
    }
  }
  for (int r = 0; r < n; r++) {
    double s = info->linearized_residuals[r];
    for (int c = 0; c < n; c++) s += info->linearized_jacobians[r * n + c] * dx[c];
    residuals[r] = s;
  }
  if (jacobians) {
    for (size_t i = 0; i < info->keep_block_size.size(); i++) {
      if (!jacobians[i]) continue;
      int size = info->keep_block_size[i], local_size = (size == 7 ? 6 : size);
      int idx = info->keep_block_idx[i] - m;
      for (int r = 0; r < n; r++) {
        for (int c = 0; c < size; c++) jacobians[i][r * size + c] = 0.0;
        for (int c = 0; c < local_size; c++) jacobians[i][r * size + c] = info->linearized_jacobians[r * n + idx + c];
      }
    }
  }
  return true;
}

// ------------------------------------------------------------------ LiDAR plane factors
LidarPlaneNormFactor::LidarPlaneNormFactor(const V3& pb, const V3& nv, double off, double w)
    : point_body(pb), norm_vector(nv), norm_offset(off), weight(w) { block_sizes = {3, 4}; num_residuals = 1; }

// LIO/liw/lidarFactor.cpp:18-50
bool LidarPlaneNormFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
  V3 translation = P3(parameters[0]);
  Quat rotation(parameters[1][3], parameters[1][0], parameters[1][1], parameters[1][2]);
  V3 point_world = rotation * point_body + translation;
  double distance = dot(norm_vector, point_world) + norm_offset;
  residuals[0] = sqrt_info * weight * distance;
  if (jacobians) {
    if (jacobians[0]) {
      for (int i = 0; i < 3; i++) jacobians[0][i] = sqrt_info * norm_vector[i] * weight;
    }
    if (jacobians[1]) {
      Mat<1, 3> jr = (-sqrt_info) * (norm_vector.T() * rotation.toRotationMatrix() * skew(point_body)) * weight;
      for (int i = 0; i < 3; i++) jacobians[1][i] = jr[i];
      jacobians[1][3] = 0.0;
    }
  }
  return true;
}

// Eigen::QuaternionBase::slerp
static Quat slerp(const Quat& a, double t, const Quat& b) {
  const double one = 1.0 - 2.220446049250313e-16;
  double d = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z;
  double absD = std::fabs(d);
  double scale0, scale1;
  if (absD >= one) { scale0 = 1.0 - t; scale1 = t; }
  else {
    double theta = std::acos(absD), sinTheta = std::sin(theta);
    scale0 = std::sin((1.0 - t) * theta) / sinTheta;
    scale1 = std::sin(t * theta) / sinTheta;
  }
  if (d < 0) scale1 = -scale1;
  return Quat(scale0 * a.w + scale1 * b.w, scale0 * a.x + scale1 * b.x, scale0 * a.y + scale1 * b.y, scale0 * a.z + scale1 * b.z);
}
static bool inv3(const M3& m, M3& out) { return invertN(3, m.a, out.a); }

CTLidarPlaneNormFactor::CTLidarPlaneNormFactor(const V3& kp, const V3& nv, double off, double alpha, double w)
    : raw_keypoint(kp), norm_vector(nv), norm_offset(off), alpha_time(alpha), weight(w) { block_sizes = {3, 4, 3, 4}; num_residuals = 1; }

// LIO/liw/lidarFactor.cpp:58-123 (t_il = 0, q_il = identity assumed applied by the caller)
bool CTLidarPlaneNormFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
  V3 tran_begin = P3(parameters[0]), tran_end = P3(parameters[2]);
  Quat rot_begin(parameters[1][3], parameters[1][0], parameters[1][1], parameters[1][2]);
  Quat rot_end(parameters[3][3], parameters[3][0], parameters[3][1], parameters[3][2]);
  Quat rot_slerp = slerp(rot_begin, alpha_time, rot_end);
  rot_slerp.normalize();
  V3 tran_slerp = tran_begin * (1 - alpha_time) + tran_end * alpha_time;
  V3 point_world = rot_slerp * raw_keypoint + tran_slerp;
  double distance = dot(norm_vector, point_world) + norm_offset;
  residuals[0] = sqrt_info * weight * distance;
  if (jacobians) {
    Mat<1, 3> jacobian_rot_slerp = -1.0 * (norm_vector.T() * rot_slerp.toRotationMatrix() * skew(raw_keypoint)) * weight;
    Quat rot_delta = rot_begin.inverse() * rot_end;
    Quat rot_identity;
    Quat rot_delta_slerp = slerp(rot_identity, alpha_time, rot_delta);
    if (jacobians[0]) for (int i = 0; i < 3; i++) jacobians[0][i] = sqrt_info * norm_vector[i] * weight * (1 - alpha_time);
    if (jacobians[1]) {
      M3 qld_inv; inv3(Qleft(rot_delta).block<3, 3>(1, 1), qld_inv);
      M3 js = rot_delta_slerp.toRotationMatrix().T() * (M3::Identity() - alpha_time * Qleft(rot_delta_slerp).block<3, 3>(1, 1) * qld_inv);
      Mat<1, 3> j = sqrt_info * (jacobian_rot_slerp * js);
      for (int i = 0; i < 3; i++) jacobians[1][i] = j[i];
      jacobians[1][3] = 0.0;
    }
    if (jacobians[2]) for (int i = 0; i < 3; i++) jacobians[2][i] = sqrt_info * norm_vector[i] * weight * alpha_time;
    if (jacobians[3]) {
      M3 qrd_inv; inv3(Qright(rot_delta).block<3, 3>(1, 1), qrd_inv);
      M3 js = alpha_time * Qright(rot_delta_slerp).block<3, 3>(1, 1) * qrd_inv;
      Mat<1, 3> j = sqrt_info * (jacobian_rot_slerp * js);
      for (int i = 0; i < 3; i++) jacobians[3][i] = j[i];
      jacobians[3][3] = 0.0;
    }
  }
  return true;
}

LidarPlanePoseFactor::LidarPlanePoseFactor(const V3& pb, const V3& nv, double off, double w) : inner(pb, nv, off, w) {
  block_sizes = {7}; num_residuals = 1;
}
bool LidarPlanePoseFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
  const double* p[2] = {parameters[0], parameters[0] + 3};
  double jt[3], jq[4];
  double* J[2] = {jt, jq};
  inner.Evaluate(p, residuals, (jacobians && jacobians[0]) ? J : nullptr);
  if (jacobians && jacobians[0]) {
    for (int i = 0; i < 3; i++) { jacobians[0][i] = jt[i]; jacobians[0][3 + i] = jq[i]; }
    jacobians[0][6] = 0.0;
  }
  return true;
}

CTLidarPlanePoseFactor::CTLidarPlanePoseFactor(const V3& kp, const V3& nv, double off, double alpha, double w) : inner(kp, nv, off, alpha, w) {
  block_sizes = {7, 7}; num_residuals = 1;
}
bool CTLidarPlanePoseFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
  const double* p[4] = {parameters[0], parameters[0] + 3, parameters[1], parameters[1] + 3};
  double jtb[3], jqb[4], jte[3], jqe[4];
  double* J[4] = {jtb, jqb, jte, jqe};
  inner.Evaluate(p, residuals, jacobians ? J : nullptr);
  if (jacobians) {
    if (jacobians[0]) { for (int i = 0; i < 3; i++) { jacobians[0][i] = jtb[i]; jacobians[0][3 + i] = jqb[i]; } jacobians[0][6] = 0.0; }
    if (jacobians[1]) { for (int i = 0; i < 3; i++) { jacobians[1][i] = jte[i]; jacobians[1][3 + i] = jqe[i]; } jacobians[1][6] = 0.0; }
  }
  return true;
}

}  // namespace gf2o
