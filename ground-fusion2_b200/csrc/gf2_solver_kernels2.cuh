// gf2_solver_kernels2.cuh — k_solve (assembly + Cholesky + Gauss-Newton step), k_backsub, k_candidate.
#pragma once
#include "gf2_solver_kernels.cuh"

namespace gf2 {

// ------------------------------------------------------------------------------------------------ IMU factor
struct ImuStates { V3 Pi, Vi, Bai, Bgi, Pj, Vj, Baj, Bgj; Q4 Qi, Qj; };

__device__ __forceinline__ ImuStates load_imu_states(const double* pose, const double* sb, int i) {
  ImuStates s;
  const double* pi = pose + 7 * i; const double* pj = pose + 7 * (i + 1);
  const double* si = sb + 9 * i; const double* sj = sb + 9 * (i + 1);
  s.Pi = ld3(pi); s.Qi = ldq(pi + 3); s.Pj = ld3(pj); s.Qj = ldq(pj + 3);
  s.Vi = ld3(si); s.Bai = ld3(si + 3); s.Bgi = ld3(si + 6); s.Vj = ld3(sj); s.Baj = ld3(sj + 3); s.Bgj = ld3(sj + 6);
  return s;
}
__device__ __forceinline__ M3 blk3(const double* J15, int r0, int c0) {
  M3 m;
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) m.m[r * 3 + c] = J15[(r0 + r) * 15 + c0 + c];
  return m;
}
__device__ __forceinline__ void put3(double* J, int ld, int r0, int c0, const M3& m, double sgn) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) J[(r0 + r) * ld + c0 + c] = sgn * m.m[r * 3 + c];
}

// IntegrationBase::evaluate (VE/factor/integration_base.h:169-195): raw 15-residual; optionally the raw 15x30 Jacobian of
// IMUFactor::Evaluate (VE/factor/imu_factor.h:98-187) in tangent columns [pose_i 6 | sb_i 9 | pose_j 6 | sb_j 9].
__device__ void imu_raw(const gf2_imu_preint& pre, const ImuStates& s, double g_norm, double* r /*15*/, double* J /*15x30 or null*/) {
  const V3 G = mk3(0, 0, g_norm);
  const double dt = pre.sum_dt;
  const M3 dp_dba = blk3(pre.jacobian, 0, 9), dp_dbg = blk3(pre.jacobian, 0, 12), dq_dbg = blk3(pre.jacobian, 3, 12);
  const M3 dv_dba = blk3(pre.jacobian, 6, 9), dv_dbg = blk3(pre.jacobian, 6, 12);
  const V3 dba = s.Bai - ld3(pre.lin_ba), dbg = s.Bgi - ld3(pre.lin_bg);
  const Q4 dq = ldq(pre.delta_q);
  const Q4 cdq = qmul(dq, deltaQ(mul(dq_dbg, dbg)));
  const V3 cdv = ld3(pre.delta_v) + mul(dv_dba, dba) + mul(dv_dbg, dbg);
  const V3 cdp = ld3(pre.delta_p) + mul(dp_dba, dba) + mul(dp_dbg, dbg);
  const Q4 QiInv = qinv(s.Qi);
  const V3 tp = 0.5 * dt * dt * G + s.Pj - s.Pi - dt * s.Vi;
  const V3 tv = dt * G + s.Vj - s.Vi;
  const V3 a_p = qrot(QiInv, tp), a_v = qrot(QiInv, tv);
  const V3 rp = a_p - cdp;
  const V3 rq = 2.0 * qvec(qmul(qinv(cdq), qmul(QiInv, s.Qj)));
  const V3 rv = a_v - cdv;
  const V3 rba = s.Baj - s.Bai, rbg = s.Bgj - s.Bgi;
  r[0] = rp.x; r[1] = rp.y; r[2] = rp.z; r[3] = rq.x; r[4] = rq.y; r[5] = rq.z; r[6] = rv.x; r[7] = rv.y; r[8] = rv.z;
  r[9] = rba.x; r[10] = rba.y; r[11] = rba.z; r[12] = rbg.x; r[13] = rbg.y; r[14] = rbg.z;
  if (!J) return;
  for (int i = 0; i < 450; i++) J[i] = 0.0;
  const M3 RiT = toR(QiInv);
  const Q4 QjInv = qinv(s.Qj);
  // pose_i (cols 0..5)
  put3(J, 30, 0, 0, RiT, -1.0);
  put3(J, 30, 0, 3, skew(a_p), 1.0);
  put3(J, 30, 3, 3, QleftQrightBR(qmul(QjInv, s.Qi), cdq), -1.0);
  put3(J, 30, 6, 3, skew(a_v), 1.0);
  // speed-bias_i (cols 6..14): V 6..8, BA 9..11, BG 12..14
  put3(J, 30, 0, 6, RiT, -dt);
  put3(J, 30, 0, 9, dp_dba, -1.0);
  put3(J, 30, 0, 12, dp_dbg, -1.0);
  put3(J, 30, 3, 12, mul(QleftBR(qmul(qmul(QjInv, s.Qi), dq)), dq_dbg), -1.0);  // uncorrected delta_q, imu_factor.h:137
  put3(J, 30, 6, 6, RiT, -1.0);
  put3(J, 30, 6, 9, dv_dba, -1.0);
  put3(J, 30, 6, 12, dv_dbg, -1.0);
  put3(J, 30, 9, 9, eye3(), -1.0);
  put3(J, 30, 12, 12, eye3(), -1.0);
  // pose_j (cols 15..20)
  put3(J, 30, 0, 15, RiT, 1.0);
  put3(J, 30, 3, 18, QleftBR(qmul(qinv(cdq), qmul(QiInv, s.Qj))), 1.0);
  // speed-bias_j (cols 21..29)
  put3(J, 30, 6, 21, RiT, 1.0);
  put3(J, 30, 9, 24, eye3(), 1.0);
  put3(J, 30, 12, 27, eye3(), 1.0);
}

// Warp-cooperative form of imu_raw for k_nonvis: every lane derives the (cheap) shared quantities itself, then lane b < 19
// forms the b-th non-zero 3x3 block of the Jacobian and lane 19 the residual — the serial version above costs one thread a
// few thousand dependent fp64 operations per factor. J (15x30) must have been zeroed by the caller.
__device__ void imu_raw_warp(const gf2_imu_preint& pre, const ImuStates& s, double g_norm, double* r /*15*/, double* J /*15x30*/, int lane) {
  const V3 G = mk3(0, 0, g_norm);
  const double dt = pre.sum_dt;
  const M3 dq_dbg = blk3(pre.jacobian, 3, 12);
  const V3 dba = s.Bai - ld3(pre.lin_ba), dbg = s.Bgi - ld3(pre.lin_bg);
  const Q4 dq = ldq(pre.delta_q);
  const Q4 cdq = qmul(dq, deltaQ(mul(dq_dbg, dbg)));
  const Q4 QiInv = qinv(s.Qi);
  const V3 tp = 0.5 * dt * dt * G + s.Pj - s.Pi - dt * s.Vi;
  const V3 tv = dt * G + s.Vj - s.Vi;
  const V3 a_p = qrot(QiInv, tp), a_v = qrot(QiInv, tv);
  if (lane == 19) {
    const M3 dp_dba = blk3(pre.jacobian, 0, 9), dp_dbg = blk3(pre.jacobian, 0, 12), dv_dba = blk3(pre.jacobian, 6, 9), dv_dbg = blk3(pre.jacobian, 6, 12);
    const V3 cdv = ld3(pre.delta_v) + mul(dv_dba, dba) + mul(dv_dbg, dbg);
    const V3 cdp = ld3(pre.delta_p) + mul(dp_dba, dba) + mul(dp_dbg, dbg);
    const V3 rp = a_p - cdp;
    const V3 rq = 2.0 * qvec(qmul(qinv(cdq), qmul(QiInv, s.Qj)));
    const V3 rv = a_v - cdv;
    const V3 rba = s.Baj - s.Bai, rbg = s.Bgj - s.Bgi;
    r[0] = rp.x; r[1] = rp.y; r[2] = rp.z; r[3] = rq.x; r[4] = rq.y; r[5] = rq.z; r[6] = rv.x; r[7] = rv.y; r[8] = rv.z;
    r[9] = rba.x; r[10] = rba.y; r[11] = rba.z; r[12] = rbg.x; r[13] = rbg.y; r[14] = rbg.z;
    return;
  }
  if (lane > 19) return;
  int r0 = 0, c0 = 0; double sgn = 1.0; M3 m;
  switch (lane) {
    case 0: r0 = 0; c0 = 0; m = toR(QiInv); sgn = -1.0; break;
    case 1: r0 = 0; c0 = 3; m = skew(a_p); break;
    case 2: r0 = 3; c0 = 3; m = QleftQrightBR(qmul(qinv(s.Qj), s.Qi), cdq); sgn = -1.0; break;
    case 3: r0 = 6; c0 = 3; m = skew(a_v); break;
    case 4: r0 = 0; c0 = 6; m = toR(QiInv); sgn = -dt; break;
    case 5: r0 = 0; c0 = 9; m = blk3(pre.jacobian, 0, 9); sgn = -1.0; break;
    case 6: r0 = 0; c0 = 12; m = blk3(pre.jacobian, 0, 12); sgn = -1.0; break;
    case 7: r0 = 3; c0 = 12; m = mul(QleftBR(qmul(qmul(qinv(s.Qj), s.Qi), dq)), dq_dbg); sgn = -1.0; break;  // uncorrected delta_q, imu_factor.h:137
    case 8: r0 = 6; c0 = 6; m = toR(QiInv); sgn = -1.0; break;
    case 9: r0 = 6; c0 = 9; m = blk3(pre.jacobian, 6, 9); sgn = -1.0; break;
    case 10: r0 = 6; c0 = 12; m = blk3(pre.jacobian, 6, 12); sgn = -1.0; break;
    case 11: r0 = 9; c0 = 9; m = eye3(); sgn = -1.0; break;
    case 12: r0 = 12; c0 = 12; m = eye3(); sgn = -1.0; break;
    case 13: r0 = 0; c0 = 15; m = toR(QiInv); break;
    case 14: r0 = 3; c0 = 18; m = QleftBR(qmul(qinv(cdq), qmul(QiInv, s.Qj))); break;
    case 15: r0 = 6; c0 = 21; m = toR(QiInv); break;
    case 16: r0 = 9; c0 = 24; m = eye3(); break;
    case 17: r0 = 12; c0 = 27; m = eye3(); break;
    default: return;
  }
  put3(J, 30, r0, c0, m, sgn);
}

// ------------------------------------------------------------------------------------------------ wheel factor
// WheelIntegrationBase::evaluate (VE/factor/wheel_integration_base.h:180-219) and the pose Jacobians of WheelFactor::Evaluate
// (VE/factor/wheel_factor.h:117-156). Calibration blocks (extrinsic, sx sy sw, td) are constant in this build, so only
// jacobians[0] (pose_i) and jacobians[1] (pose_j) are formed: J is 6 x 12 in tangent columns [pose_i 6 | pose_j 6].
__device__ void wheel_raw(const gf2_wheel_preint& pre, const double* pose_i, const double* pose_j, const double* exw, const double* sxw, double tdw,
                          double* r /*6*/, double* J /*6x12 or null*/) {
  const V3 Pi = ld3(pose_i), Pj = ld3(pose_j), tio = ld3(exw);
  const Q4 Qi = ldq(pose_i + 3), Qj = ldq(pose_j + 3), qio = ldq(exw + 3);
  const double sx = sxw[0], sy = sxw[1], sw = sxw[2];
  const V3 dp_dsx = mk3(pre.jacobian[0], pre.jacobian[3], pre.jacobian[6]), dp_dsy = mk3(pre.jacobian[1], pre.jacobian[4], pre.jacobian[7]);
  const V3 dp_dsw = mk3(pre.jacobian[2], pre.jacobian[5], pre.jacobian[8]), dq_dsw = mk3(pre.jacobian[11], pre.jacobian[14], pre.jacobian[17]);
  const double dsx = sx - pre.lin_sx, dsy = sy - pre.lin_sy, dsw = sw - pre.lin_sw;
  const M3 Ri = toR(Qi), Rj = toR(Qj), rio = toR(qio);
  const V3 cdp = ld3(pre.delta_p) + dsx * dp_dsx + dsy * dp_dsy + dsw * dp_dsw;
  const Q4 cdq = qnormalized(qmul(qnormalized(ldq(pre.delta_q)), so3Exp(dsw * dq_dsw)));
  const double dtd = tdw - pre.lin_td;
  const V3 lg = ld3(pre.lin_gyr), lv = ld3(pre.lin_vel), g1 = ld3(pre.gyr_1), v1 = ld3(pre.vel_1);
  const Q4 ef = so3Exp((sw * dtd) * lg);
  const Q4 dqt = qnormalized(qmul(qmul(ef, cdq), so3Exp((-sw * dtd) * g1)));
  const V3 svlv = mk3(sx * lv.x, sy * lv.y, lv.z), svv1 = mk3(sx * v1.x, sy * v1.y, v1.z);
  const V3 dpt = mul(toR(ef), dtd * svlv + cdp - qrot(cdq, dtd * svv1));
  const V3 dpw = mul(Rj, tio) + Pj - mul(Ri, tio) - Pi;
  const M3 Rio = mul(Ri, rio);
  const V3 rp = mulT(Rio, dpw) - dpt;
  const Q4 Qio = qmul(Qi, qio);
  const V3 rr = so3Log(qmul(qmul(qmul(qinv(dqt), qinv(Qio)), Qj), qio));
  r[0] = rp.x; r[1] = rp.y; r[2] = rp.z; r[3] = rr.x; r[4] = rr.y; r[5] = rr.z;
  if (!J) return;
  for (int i = 0; i < 72; i++) J[i] = 0.0;
  const M3 Jrinv = rightJacobianInvSO3(rr);
  const M3 RioInv = toR(qinv(Qio));
  put3(J, 12, 0, 0, RioInv, -1.0);
  put3(J, 12, 0, 3, add(mulT(Rio, mul(Ri, skew(tio))), mulT(rio, skew(mulT(Ri, dpw)))), 1.0);
  put3(J, 12, 3, 3, mul(Jrinv, toR(qmul(qinv(qmul(Qj, qio)), Qi))), -1.0);
  put3(J, 12, 0, 6, RioInv, 1.0);
  put3(J, 12, 0, 9, mul(toR(qmul(qinv(Qio), Qj)), skew(tio)), -1.0);
  put3(J, 12, 3, 9, mul(Jrinv, toR(qinv(qio))), 1.0);
}
// Calibration-block Jacobians of WheelFactor::Evaluate (VE/factor/wheel_factor.h:157-247): extrinsic body_T_wheel (6),
// sx, sy, sw, td — constant in the solve of this build, but kept variables of the marginalization prior. Jc is 6 x 10 in
// tangent columns [ex_wheel 6 | sx | sy | sw | td]. Quirks of the reference are kept (e.g. Exp(forward_compensate_v) as a
// rotation in the sx / sy columns).
__device__ void wheel_calib_jacobians(const gf2_wheel_preint& pre, const double* pose_i, const double* pose_j, const double* exw, const double* sxw, double tdw,
                                      double* Jc /*6x10*/) {
  const V3 Pi = ld3(pose_i), Pj = ld3(pose_j), tio = ld3(exw);
  const Q4 Qi = ldq(pose_i + 3), Qj = ldq(pose_j + 3), qio = ldq(exw + 3);
  const double sx = sxw[0], sy = sxw[1], sw = sxw[2];
  const V3 dp_dsx = mk3(pre.jacobian[0], pre.jacobian[3], pre.jacobian[6]), dp_dsy = mk3(pre.jacobian[1], pre.jacobian[4], pre.jacobian[7]);
  const V3 dp_dsw = mk3(pre.jacobian[2], pre.jacobian[5], pre.jacobian[8]), dq_dsw = mk3(pre.jacobian[11], pre.jacobian[14], pre.jacobian[17]);
  const double dsx = sx - pre.lin_sx, dsy = sy - pre.lin_sy, dsw = sw - pre.lin_sw;
  const M3 Ri = toR(Qi), Rj = toR(Qj);
  const V3 cdp = ld3(pre.delta_p) + dsx * dp_dsx + dsy * dp_dsy + dsw * dp_dsw;
  const Q4 cdq = qnormalized(qmul(qnormalized(ldq(pre.delta_q)), so3Exp(dsw * dq_dsw)));
  const double dtd = tdw - pre.lin_td;
  const V3 lg = ld3(pre.lin_gyr), lv = ld3(pre.lin_vel), g1 = ld3(pre.gyr_1), v1 = ld3(pre.vel_1);
  const Q4 ef = so3Exp((sw * dtd) * lg);
  const Q4 dqt = qnormalized(qmul(qmul(ef, cdq), so3Exp((-sw * dtd) * g1)));
  const Q4 Qio = qmul(Qi, qio);
  const V3 rr = so3Log(qmul(qmul(qmul(qinv(dqt), qinv(Qio)), Qj), qio));   // raw rotation residual
  const M3 Jrinv = rightJacobianInvSO3(rr);
  const M3 Jr_drdsw = rightJacobianSO3(dsw * dq_dsw);
  const V3 fcw = (sw * dtd) * lg, bcw = (sw * dtd) * g1;
  const V3 fcv = dtd * mk3(sx * lv.x, sy * lv.y, lv.z), bcv = dtd * mk3(sx * v1.x, sy * v1.y, v1.z);
  const M3 Jrtd = rightJacobianSO3(fcw), Jr_minus_td = rightJacobianSO3(-fcw);
  const M3 cdqR = toR(cdq), RioInv = toR(qinv(Qio));
  for (int i = 0; i < 60; i++) Jc[i] = 0.0;
  // extrinsic, wheel_factor.h:157-171
  put3(Jc, 10, 0, 0, mul(RioInv, sub(Rj, Ri)), 1.0);
  put3(Jc, 10, 0, 3, skew(qrot(qinv(Qio), qrot(Qj, tio) + Pj - qrot(Qi, tio) - Pi)), 1.0);
  put3(Jc, 10, 3, 3, mul(Jrinv, sub(eye3(), toR(qmul(qmul(qinv(qmul(Qj, qio)), Qi), qio)))), 1.0);
  // sx, sy
  const M3 Efv = toR(so3Exp(fcv));
  const V3 tsx = -mul(Efv, mk3(lv.x * dtd, 0, 0) + dp_dsx - mul(cdqR, mk3(v1.x, 0, 0)) * dtd);
  const V3 tsy = -mul(Efv, mk3(0, lv.y * dtd, 0) + dp_dsy - mul(cdqR, mk3(0, v1.y, 0)) * dtd);
  Jc[0 * 10 + 6] = tsx.x; Jc[1 * 10 + 6] = tsx.y; Jc[2 * 10 + 6] = tsx.z;
  Jc[0 * 10 + 7] = tsy.x; Jc[1 * 10 + 7] = tsy.y; Jc[2 * 10 + 7] = tsy.z;
  // sw
  const M3 Efw = toR(so3Exp(fcw));
  const V3 inner = fcv + cdp - qrot(cdq, bcv);
  const V3 tpw = -mul(Efw, dp_dsw - mul(cdqR, cross(mul(Jr_drdsw, dq_dsw), mk3(sx * v1.x, sy * v1.y, v1.z) * dtd)) + cross(mul(Jrtd, lg * dtd), inner));
  const M3 Emr = toR(so3Exp(-rr)), Ebw = toR(so3Exp(bcw)), cdqInvR = toR(qinv(cdq));
  const V3 trw = -mul(Jrinv, mul(Emr, mul(Ebw, mul(cdqInvR, mul(Jrtd, lg * dtd)) + mul(Jr_drdsw, dq_dsw))));
  Jc[0 * 10 + 8] = tpw.x; Jc[1 * 10 + 8] = tpw.y; Jc[2 * 10 + 8] = tpw.z; Jc[3 * 10 + 8] = trw.x; Jc[4 * 10 + 8] = trw.y; Jc[5 * 10 + 8] = trw.z;
  // td
  const V3 svlv = mk3(sx * lv.x, sy * lv.y, lv.z), svv1 = mk3(sx * v1.x, sy * v1.y, v1.z);
  const V3 tpt = -mul(Efw, svlv - mul(cdqR, svv1) + cross(mul(Jrtd, sw * lg), fcv + cdp - mul(cdqR, bcv)));
  const V3 trt = -mul(Jrinv, mul(Emr, mul(Ebw, mul(cdqInvR, mul(Jrtd, sw * lg))) - mul(Jr_minus_td, sw * g1)));
  Jc[0 * 10 + 9] = tpt.x; Jc[1 * 10 + 9] = tpt.y; Jc[2 * 10 + 9] = tpt.z; Jc[3 * 10 + 9] = trt.x; Jc[4 * 10 + 9] = trt.y; Jc[5 * 10 + 9] = trt.z;
}
// out-of-line copy for k_nonvis: inlined there it quadruples the spills of the IMU path, which never calls it
__device__ __noinline__ void wheel_calib_jacobians_masked(const gf2_wheel_preint& pre, const double* pose_i, const double* pose_j, const double* exw, const double* sxw, double tdw,
                                                          int wcal, double* Jc /*6x10*/) {
  wheel_calib_jacobians(pre, pose_i, pose_j, exw, sxw, tdw, Jc);
  for (int a = 0; a < 6; a++) {  // constant sub-blocks have no column
    if (!(wcal & 1)) for (int c = 0; c < 6; c++) Jc[a * 10 + c] = 0.0;
    if (!(wcal & 2)) for (int c = 6; c < 9; c++) Jc[a * 10 + c] = 0.0;
    if (!(wcal & 4)) Jc[a * 10 + 9] = 0.0;
  }
}
__device__ __forceinline__ double wheel_cost(const double* sq /*6x6 upper*/, const double* r) {
  double c = 0;
  for (int a = 0; a < 6; a++) { double s = 0; for (int k = a; k < 6; k++) s += sq[a * 6 + k] * r[k]; c += s * s; }
  return 0.5 * c;
}

// LidarPlaneNormFactor (LIO/liw/lidarFactor.cpp:18-50) attached to the window pose of its frame: residual and the 1x6 tangent
// Jacobian [sqrt_info w n^T | -sqrt_info w n^T R [p]x]
__device__ __forceinline__ double plane_residual(const gf2_plane& pl, const FrameCtx& f, double sqrt_info, double* J /*6 or null*/) {
  const V3 p = ld3(pl.p_body), n = ld3(pl.normal);
  M3 R; for (int i = 0; i < 9; i++) R.m[i] = f.R[i];
  const V3 pw = mul(R, p) + mk3(f.P[0], f.P[1], f.P[2]);
  const double sw = sqrt_info * pl.weight;
  if (J) {
    const V3 a = mulT(R, n);        // R^T n;  (a^T [p]x) = (a x p)^T
    const V3 c = cross(a, p);
    J[0] = sw * n.x; J[1] = sw * n.y; J[2] = sw * n.z; J[3] = -sw * c.x; J[4] = -sw * c.y; J[5] = -sw * c.z;
  }
  return sw * (dot(n, pw) + pl.offset);
}

// Eigen's QuaternionBase::slerp (the call of LIO/liw/lidarFactor.cpp:67,81)
__device__ __forceinline__ Q4 eigen_slerp(Q4 a, double t, Q4 b) {
  const double one = 1.0 - 2.220446049250313e-16;
  const double d = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z, absD = fabs(d);
  double s0, s1;
  if (absD >= one) { s0 = 1.0 - t; s1 = t; }
  else { const double theta = acos(absD), st = sin(theta); s0 = sin((1.0 - t) * theta) / st; s1 = sin(t * theta) / st; }
  if (d < 0) s1 = -s1;
  Q4 r; r.w = s0 * a.w + s1 * b.w; r.x = s0 * a.x + s1 * b.x; r.y = s0 * a.y + s1 * b.y; r.z = s0 * a.z + s1 * b.z;
  return r;
}
__device__ __forceinline__ M3 inv3x3(const M3& a) {
  const double* m = a.m;
  const double c0 = m[4] * m[8] - m[5] * m[7], c1 = m[5] * m[6] - m[3] * m[8], c2 = m[3] * m[7] - m[4] * m[6];
  const double id = 1.0 / (m[0] * c0 + m[1] * c1 + m[2] * c2);
  M3 r;
  r.m[0] = c0 * id; r.m[1] = (m[2] * m[7] - m[1] * m[8]) * id; r.m[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  r.m[3] = c1 * id; r.m[4] = (m[0] * m[8] - m[2] * m[6]) * id; r.m[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  r.m[6] = c2 * id; r.m[7] = (m[1] * m[6] - m[0] * m[7]) * id; r.m[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return r;
}
// CTLidarPlaneNormFactor (LIO/liw/lidarFactor.cpp:58-123) attached to the window poses pose_b (begin) and pose_e (end): residual and
// the 1x12 tangent Jacobian [t_begin 3 | q_begin 3 | t_end 3 | q_end 3]. The rotation blocks are the reference's formulas verbatim
// (first order in the begin-end rotation difference).
__device__ double ct_plane_residual(const gf2_plane& pl, double alpha, const double* pose_b, const double* pose_e, double sqrt_info, double* J /*12 or null*/) {
  const V3 p = ld3(pl.p_body), n = ld3(pl.normal);
  const V3 tb = ld3(pose_b), te = ld3(pose_e);
  const Q4 qb = ldq(pose_b + 3), qe = ldq(pose_e + 3);
  const Q4 qs = qnormalized(eigen_slerp(qb, alpha, qe));
  const V3 ts = tb * (1.0 - alpha) + te * alpha;
  const M3 Rs = toR(qs);
  const V3 pw = mul(Rs, p) + ts;
  const double sw = sqrt_info * pl.weight;
  if (J) {
    const V3 jrs = -1.0 * mulT(mul(Rs, skew(p)), n) * pl.weight;   // -(n^T R_slerp [p]x) * weight as a column
    const Q4 rd = qmul(qinv(qb), qe);
    Q4 id; id.w = 1.0; id.x = 0.0; id.y = 0.0; id.z = 0.0;
    const Q4 rds = eigen_slerp(id, alpha, rd);
    const M3 jsb = mulT(toR(rds), sub(eye3(), scale(mul(QleftBR(rds), inv3x3(QleftBR(rd))), alpha)));
    const M3 jse = scale(mul(QrightBR(rds), inv3x3(QrightBR(rd))), alpha);
    const V3 jb = mulT(jsb, jrs), je = mulT(jse, jrs);           // (jrs^T M)^T = M^T jrs
    J[0] = sw * n.x * (1.0 - alpha); J[1] = sw * n.y * (1.0 - alpha); J[2] = sw * n.z * (1.0 - alpha);
    J[3] = sqrt_info * jb.x; J[4] = sqrt_info * jb.y; J[5] = sqrt_info * jb.z;
    J[6] = sw * n.x * alpha; J[7] = sw * n.y * alpha; J[8] = sw * n.z * alpha;
    J[9] = sqrt_info * je.x; J[10] = sqrt_info * je.y; J[11] = sqrt_info * je.z;
  }
  return sw * (dot(n, pw) + pl.offset);
}

// 0.5 * |sqrt_info * r|^2 by one thread
__device__ __forceinline__ double imu_cost(const double* sq /*15x15 upper*/, const double* r) {
  double c = 0;
  for (int a = 0; a < 15; a++) { double s = 0; for (int k = a; k < 15; k++) s += sq[a * 15 + k] * r[k]; c += s * s; }
  return 0.5 * c;
}

// MarginalizationFactor::Evaluate's dx (VE/factor/marginalization_factor.cpp:356-374) for one kept block. cal = {ex_wheel[7],
// sxsysw[3], td_wheel} of the state the prior is evaluated at (null without wheel). The camera extrinsic and td never move in
// this build (x == x0), so their dx is 0.
struct CalibPtr { const double *exw, *sxw, *tdw; };
__device__ __forceinline__ void prior_pose_dx(const double* x, const double* x0, double* dx) {
  for (int k = 0; k < 3; k++) dx[k] = x[k] - x0[k];
  Q4 q0 = ldq(x0 + 3), q = ldq(x + 3);
  Q4 dq = qmul(qinv(q0), q);
  V3 v = 2.0 * qvec(dq);
  if (!(dq.w >= 0)) v = -v;
  dx[3] = v.x; dx[4] = v.y; dx[5] = v.z;
}
__device__ __forceinline__ void prior_block_dx(const gf2_prior_block& b, const double* pose, const double* sb, const CalibPtr& cal, double* dx /* at offset */) {
  if (b.kind == GF2_BLK_POSE) prior_pose_dx(pose + 7 * b.index, b.x0, dx + b.offset);
  else if (b.kind == GF2_BLK_SPEEDBIAS) {
    const double* x = sb + 9 * b.index;
    for (int k = 0; k < 9; k++) dx[b.offset + k] = x[k] - b.x0[k];
  } else if (b.kind == GF2_BLK_EX_WHEEL && cal.exw) prior_pose_dx(cal.exw, b.x0, dx + b.offset);
  else if (b.kind >= GF2_BLK_SX && b.kind <= GF2_BLK_SW && cal.sxw) dx[b.offset] = cal.sxw[b.kind - GF2_BLK_SX] - b.x0[0];
  else if (b.kind == GF2_BLK_TD_WHEEL && cal.tdw) dx[b.offset] = cal.tdw[0] - b.x0[0];
  else {
    const int ls = (b.kind == GF2_BLK_EX_POSE || b.kind == GF2_BLK_EX_WHEEL) ? 6 : 1;
    for (int k = 0; k < ls; k++) dx[b.offset + k] = 0.0;
  }
}
__device__ __forceinline__ CalibPtr calib_of(const KP& p, int w, bool candidate) {
  CalibPtr c = {nullptr, nullptr, nullptr};
  if (!p.use_wheel) return c;
  const bool cc = candidate && p.wcal;
  c.exw = (cc ? p.exw_c : p.exw) + (size_t)w * 7; c.sxw = (cc ? p.sxw_c : p.sxw) + (size_t)w * 3; c.tdw = (cc ? p.tdw_c : p.tdw) + w;
  return c;
}

__device__ __forceinline__ int pidx(int i, int j) { return i * (i + 1) / 2 + j; }  // packed lower, j <= i

// ------------------------------------------------------------------------------------------------ k_backsub
// sweep 2: z_l = (g_l - w_l^T z_x) / v'_l and the landmark parts of the dogleg dot products.
//
// w_l^T x for a pose-space vector x never needs the 2x6 Jacobians: with Jx = d r / d pts_w (gf2_solver_lin.cuh, header)
//     J_pose_i x_i + J_pose_j x_j = Jx ( x_i^p - [Xw - Pi]x (Ri x_i^th)  -  x_j^p + [Xw - Pj]x (Rj x_j^th) ),
// so per frame the two rotated vectors Rf x_f^th are formed once and every observation costs one cross product and one
// 2x3 product per vector.
struct StepShared {
  double rimu[kMaxF][15];   // raw IMU residuals of the candidate (k_cand_eval)
  FrameCtx fr[kMaxF];
  CamCtx cam;
  double zx[kNVMax], ux[kNVMax];
  double rz[kMaxF][3], ru[kMaxF][3];  // Rf z_f^theta, Rf u_f^theta
  double red[8 * 32];
  double dx[kP];
  int decision;
};

// body of sweep 2 for one window: threads t < 256 walk the landmarks; the six sums are valid in thread 0 afterwards
__device__ __forceinline__ void backsub_body(const KP& p, int w, const WinState& st, StepShared& S, double (&sums)[6]) {
  const int t = threadIdx.x, F = p.F, NV = 6 * F;
  const int nlm = p.nlm[w];
  const int32_t* start = p.start + (size_t)w * p.Lm; const int32_t* tlen = p.tlen + (size_t)w * p.Lm; const int32_t* obeg = p.obeg + (size_t)w * p.Lm;
  const float4* obs = p.obs + (size_t)w * p.Om;
  build_frames(p.pose + (size_t)w * F * 7, p.ex + (size_t)w * 7, p.td[w], F, S.fr, &S.cam);
  if (t < NV) { const int d = 15 * (t / 6) + t % 6; S.zx[t] = p.zx[(size_t)w * p.Ds + d]; S.ux[t] = p.ux[(size_t)w * p.Ds + d]; }
  __syncthreads();
  if (t < 2 * F) {
    const int f = t >> 1; const double* x = (t & 1) ? &S.ux[6 * f + 3] : &S.zx[6 * f + 3]; const double* R = S.fr[f].R;
    double* o = (t & 1) ? S.ru[f] : S.rz[f];
    o[0] = R[0] * x[0] + R[1] * x[1] + R[2] * x[2]; o[1] = R[3] * x[0] + R[4] * x[1] + R[5] * x[2]; o[2] = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
  }
  __syncthreads();
  const double* ftd = p.frame_td + (size_t)w * F;
  const double mu = st.mu, sqi = p.sqrt_info_px;
#pragma unroll
  for (int q = 0; q < 6; q++) sums[q] = 0.0;  // dlg2, gn2, gz, zEz, uEz, uHu
  for (int l = t; l < nlm && t < 256; l += 256) {
   {
    const size_t o = (size_t)w * p.Lm + l;
    const double v = p.lm_v[o];
    if (!(v > 0.0)) { p.lm_z[o] = 0.0; continue; }  // fixed or unobserved landmark
    const int i = start[l], L = tlen[l], ob = obeg[l];
    const double gl = p.lm_g[o], s_l = p.lm_s[o];
    LmCtx lc; landmark_ctx(S.fr[i], S.cam, obs[ob], ftd[i], p.invdep[o], lc);
    float4 lo1 = obs[ob + (L > 1 ? 1 : 0)];
    // per observation only m = Jx^T j_lambda (Huber-weighted) and c = m x d are formed:
    //   w_l^T x = (sum m) . (x_i^p - e_i x (Ri x_i^th))  +  sum_k ( -m_k . x_j^p + c_k . (Rj x_j^th) )
    double az = 0.0, cu = 0.0;  // w_l^T z_x, w_l^T u_x
    V3 ms = mk3(0, 0, 0);
    float4 oj = lo1;
    for (int k = 1; k < L; k++) {
      const int j = i + k;
      const FrameCtx& fj = S.fr[j];
      const float4 o = oj;
      if (k + 1 < L) oj = obs[ob + k + 1];
      const V3 d = mk3(lc.Xw.x - fj.P[0], lc.Xw.y - fj.P[1], lc.Xw.z - fj.P[2]);
      const double px = fj.A[0] * d.x + fj.A[1] * d.y + fj.A[2] * d.z - S.cam.rtt[0];
      const double py = fj.A[3] * d.x + fj.A[4] * d.y + fj.A[5] * d.z - S.cam.rtt[1];
      const double pz = fj.A[6] * d.x + fj.A[7] * d.y + fj.A[8] * d.z - S.cam.rtt[2];
      const double dt = S.cam.td - ftd[j];
      const double iz = fast_rcp(pz);
      const double r0 = sqi * (px * iz - ((double)o.x - dt * (double)o.z)), r1 = sqi * (py * iz - ((double)o.y - dt * (double)o.w));
      // Jx enters twice (through j_lambda and through the pose Jacobians): the Huber scale appears squared, rho' = delta / |r|
      const double sq = r0 * r0 + r1 * r1, hb = p.huber * p.huber;
      const double sc2 = sq > hb ? p.huber * fast_rsqrt(sq) : 1.0;
      const double a = sc2 * sqi * sqi * iz * iz;
      const double qx = -px * iz, qy = -py * iz;   // rows of Jx / (sqrt_info / z): (A0 + qx A2), (A1 + qy A2)
      const double j00 = fj.A[0] + qx * fj.A[6], j01 = fj.A[1] + qx * fj.A[7], j02 = fj.A[2] + qx * fj.A[8];
      const double j10 = fj.A[3] + qy * fj.A[6], j11 = fj.A[4] + qy * fj.A[7], j12 = fj.A[5] + qy * fj.A[8];
      const double jl0 = a * (j00 * lc.dXdl.x + j01 * lc.dXdl.y + j02 * lc.dXdl.z), jl1 = a * (j10 * lc.dXdl.x + j11 * lc.dXdl.y + j12 * lc.dXdl.z);
      const V3 m = mk3(jl0 * j00 + jl1 * j10, jl0 * j01 + jl1 * j11, jl0 * j02 + jl1 * j12);
      const V3 c = cross(m, d);
      ms = ms + m;
      az += c.x * S.rz[j][0] + c.y * S.rz[j][1] + c.z * S.rz[j][2] - (m.x * S.zx[6 * j] + m.y * S.zx[6 * j + 1] + m.z * S.zx[6 * j + 2]);
      cu += c.x * S.ru[j][0] + c.y * S.ru[j][1] + c.z * S.ru[j][2] - (m.x * S.ux[6 * j] + m.y * S.ux[6 * j + 1] + m.z * S.ux[6 * j + 2]);
    }
    {  // host-frame part
      const V3 ei = mk3(lc.Xw.x - S.fr[i].P[0], lc.Xw.y - S.fr[i].P[1], lc.Xw.z - S.fr[i].P[2]);
      const V3 hz = mk3(S.zx[6 * i], S.zx[6 * i + 1], S.zx[6 * i + 2]) - cross(ei, ld3(S.rz[i]));
      const V3 hu = mk3(S.ux[6 * i], S.ux[6 * i + 1], S.ux[6 * i + 2]) - cross(ei, ld3(S.ru[i]));
      az += dot(ms, hz); cu += dot(ms, hu);
    }
    const double d2 = fmin(fmax(s_l * s_l * v, 1e-6), 1e32);
    const double e = d2 / (s_l * s_l);
    const double vp = v + mu * e;
    const double u = s_l * s_l * gl / d2;
    const double z = (gl - az) / vp;
    p.lm_z[(size_t)w * p.Lm + l] = z;
    sums[0] += s_l * s_l * gl * gl / d2;
    sums[1] += d2 * z * z / (s_l * s_l);
    sums[2] += gl * z;
    sums[3] += az * az / vp + 2.0 * az * z + v * z * z;             // landmark part of z^T H z
    sums[4] += cu * az / vp + cu * z + u * az + v * u * z;          // landmark part of u^T H z
    sums[5] += cu * cu / vp + 2.0 * u * cu + v * u * u;             // landmark part of u^T H u
   }
  }
  block_sum<6>(sums, S.red);
}
__global__ void __launch_bounds__(256, 3) k_backsub(KP p, int w0) {
  const int w = w0 + blockIdx.x;
  WinState& st = p.st[w];
  if (!st.active || st.reuse || !st.lin_valid) return;
  __shared__ StepShared S;
  double sums[6];
  backsub_body(p, w, st, S, sums);
  if (threadIdx.x == 0) { double* cs = p.c_sums + (size_t)w * 8; for (int i = 0; i < 6; i++) cs[i] = sums[i]; }
}

// ------------------------------------------------------------------------------------------------ k_candidate
// DoglegStrategy::ComputeTraditionalDoglegStep + candidate evaluation + step acceptance.
__device__ void dogleg_coefficients(WinState& st, const double* cs) {
  st.dlg2_l = cs[0]; st.gn2_l = cs[1]; st.gz_l = cs[2]; st.zHz_l = cs[3]; st.uHz_l = cs[4]; st.uHu_l = cs[5];
  const double dlg2 = st.dlg2_x + st.dlg2_l, gn2 = st.gn2_x + st.gn2_l, gz = st.gz_x + st.gz_l;
  const double uHu = st.uSu - st.mu * st.uEu_x + st.uHu_l;
  const double alpha = dlg2 / uHu;
  const double gradient_norm = sqrt(dlg2), gn_norm = sqrt(gn2), radius = st.radius;
  double a, b, nrm;
  if (gn_norm <= radius) { a = 0.0; b = 1.0; nrm = gn_norm; }
  else if (gradient_norm * alpha >= radius) { a = radius / gradient_norm; b = 0.0; nrm = radius; }
  else {
    const double b_dot_a = alpha * gz;  // -alpha * dot(gradient_, gauss_newton_step_), dot = -g^T z
    const double a_sq = (alpha * gradient_norm) * (alpha * gradient_norm);
    const double bma = a_sq - 2 * b_dot_a + gn2;
    const double c = b_dot_a - a_sq;
    const double d = sqrt(c * c + bma * (radius * radius - a_sq));
    const double beta = (c <= 0) ? (d - c) / bma : (radius * radius - a_sq) / (d + c);
    a = alpha * (1.0 - beta); b = beta;
    nrm = sqrt(a * a * dlg2 + 2 * a * b * gz + b * b * gn2);
  }
  st.coef_a = a; st.coef_b = b; st.dogleg_step_norm = nrm;
  // model_cost_change = -(J delta)^T (r + J delta / 2), delta = -(a u + b z)
  // quadratic forms of the UNregularised H evaluated explicitly (x-part through the Cholesky factor, landmark part in
  // k_backsub): the identity H z = g - mu E z is not used because it amplifies the linear-solve error
  const double uHz = st.uSz - st.mu * st.uEz_x + st.uHz_l, zHz = st.zSz - st.mu * st.zEz_x + st.zHz_l;
  st.model_cost_change = a * dlg2 + b * gz - 0.5 * (a * a * uHu + 2 * a * b * uHz + b * b * zHz);
}

// body of sweep 3 for one window (288 threads: 8 landmark warps + warp 8 for the non-visual factors); st.coef_a / coef_b are set and visible.
// Afterwards thread 0 holds acc = this rank's landmark / plane sums and accx = the frame-state / non-visual sums.
__device__ __forceinline__ void cand_body(const KP& p, int w, const WinState& st, StepShared& S, double (&acc)[3], double (&accx)[3]) {
  const int t = threadIdx.x, F = p.F;
  const double a = st.coef_a, b = st.coef_b;
  const double* pose = p.pose + (size_t)w * F * 7; const double* sb = p.sb + (size_t)w * F * 9;
  double* pose_c = p.pose_c + (size_t)w * F * 7; double* sb_c = p.sb_c + (size_t)w * F * 9;
  // acc: this rank's landmarks / planes: candidate cost, |dl|^2, |l|^2 (all-reduced in sharded mode); accx: frame states and non-visual
  // factors (every rank computes the same): cost, |dx|^2, |x|^2
#pragma unroll
  for (int q = 0; q < 3; q++) { acc[q] = 0.0; accx[q] = 0.0; }
  // retraction of frame states: PoseLocalParameterization::Plus (VE/factor/pose_local_parameterization.cpp:12-28)
  if (t < F) {
    const double* zx = p.zx + (size_t)w * p.Ds + 15 * t; const double* ux = p.ux + (size_t)w * p.Ds + 15 * t;
    double dl[15];
    for (int k = 0; k < 15; k++) dl[k] = -(a * ux[k] + b * zx[k]);
    const double* x = pose + 7 * t; double* xc = pose_c + 7 * t;
    for (int k = 0; k < 3; k++) { xc[k] = x[k] + dl[k]; accx[1] += dl[k] * dl[k]; accx[2] += x[k] * x[k]; }
    Q4 q = ldq(x + 3); Q4 qn = qnormalized(qmul(q, deltaQ(mk3(dl[3], dl[4], dl[5]))));
    xc[3] = qn.x; xc[4] = qn.y; xc[5] = qn.z; xc[6] = qn.w;
    accx[1] += (q.x - qn.x) * (q.x - qn.x) + (q.y - qn.y) * (q.y - qn.y) + (q.z - qn.z) * (q.z - qn.z) + (q.w - qn.w) * (q.w - qn.w);
    accx[2] += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    const double* v = sb + 9 * t; double* vc = sb_c + 9 * t;
    for (int k = 0; k < 9; k++) { vc[k] = v[k] + dl[6 + k]; accx[1] += dl[6 + k] * dl[6 + k]; accx[2] += v[k] * v[k]; }
  } else if (t == F && p.wcal) {
    // free wheel calibration blocks (block row F of the reduced system): body_T_wheel by PoseLocalParameterization::Plus, the
    // scalars additively; constant sub-blocks are copied (their step entries are zero) and stay out of the norms, as in Ceres
    const double* zx = p.zx + (size_t)w * p.Ds + 15 * F; const double* ux = p.ux + (size_t)w * p.Ds + 15 * F;
    double dl[10];
    for (int k = 0; k < 10; k++) dl[k] = -(a * ux[k] + b * zx[k]);
    for (int k = 0; k < 6; k++) if ((p.wsub >> k) & 1u) dl[k] = 0.0;   // PoseSubsetParameterization::Plus (pose_subset_parameterization.cpp:29-33)
    const double* x = p.exw + (size_t)w * 7; double* xc = p.exw_c + (size_t)w * 7;
    if (p.wcal & 1) {
      for (int k = 0; k < 3; k++) { xc[k] = x[k] + dl[k]; accx[1] += dl[k] * dl[k]; accx[2] += x[k] * x[k]; }
      Q4 q = ldq(x + 3); Q4 qn = qnormalized(qmul(q, deltaQ(mk3(dl[3], dl[4], dl[5]))));
      xc[3] = qn.x; xc[4] = qn.y; xc[5] = qn.z; xc[6] = qn.w;
      accx[1] += (q.x - qn.x) * (q.x - qn.x) + (q.y - qn.y) * (q.y - qn.y) + (q.z - qn.z) * (q.z - qn.z) + (q.w - qn.w) * (q.w - qn.w);
      accx[2] += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    } else for (int k = 0; k < 7; k++) xc[k] = x[k];
    const double* sv = p.sxw + (size_t)w * 3; double* svc = p.sxw_c + (size_t)w * 3;
    for (int k = 0; k < 3; k++) { const double d = (p.wcal & 2) ? dl[6 + k] : 0.0; svc[k] = sv[k] + d; if (p.wcal & 2) { accx[1] += d * d; accx[2] += sv[k] * sv[k]; } }
    { const double d = (p.wcal & 4) ? dl[9] : 0.0; p.tdw_c[w] = p.tdw[w] + d; if (p.wcal & 4) { accx[1] += d * d; accx[2] += p.tdw[w] * p.tdw[w]; } }
  }
  __syncthreads();
  build_frames(pose_c, p.ex + (size_t)w * 7, p.td[w], F, S.fr, &S.cam);
  __syncthreads();
  const int nlm = p.nlm[w];
  const int32_t* start = p.start + (size_t)w * p.Lm; const int32_t* tlen = p.tlen + (size_t)w * p.Lm; const int32_t* obeg = p.obeg + (size_t)w * p.Lm;
  const float4* obs = p.obs + (size_t)w * p.Om;
  const double* ftd = p.frame_td + (size_t)w * F;
  const double sqi = p.sqrt_info_px;
  for (int l = t; l < nlm && t < 256; l += 256) {
    const size_t o = (size_t)w * p.Lm + l;
    const int i = start[l], L = tlen[l], ob = obeg[l];
    const double v = p.lm_v[o], lam = p.invdep[o];
    const float4 oi = obs[ob];
    float4 oj = obs[ob + (L > 1 ? 1 : 0)];
    double lam_c = lam;
    if (v > 0.0) {
      const double s_l = p.lm_s[o];
      const double d2 = fmin(fmax(s_l * s_l * v, 1e-6), 1e32);
      const double u = s_l * s_l * p.lm_g[o] / d2;
      const double dl = -(a * u + b * p.lm_z[o]);
      lam_c = lam + dl;
      acc[1] += dl * dl; acc[2] += lam * lam;
    }
    p.invdep_c[o] = lam_c;
    LmCtx lc; landmark_ctx(S.fr[i], S.cam, oi, ftd[i], lam_c, lc);
    for (int k = 1; k < L; k++) {
      const FrameCtx& fj = S.fr[i + k];
      const float4 ok = oj;
      if (k + 1 < L) oj = obs[ob + k + 1];
      const double dx = lc.Xw.x - fj.P[0], dy = lc.Xw.y - fj.P[1], dz = lc.Xw.z - fj.P[2];
      const double px = fj.A[0] * dx + fj.A[1] * dy + fj.A[2] * dz - S.cam.rtt[0];
      const double py = fj.A[3] * dx + fj.A[4] * dy + fj.A[5] * dz - S.cam.rtt[1];
      const double pz = fj.A[6] * dx + fj.A[7] * dy + fj.A[8] * dz - S.cam.rtt[2];
      const double dt = S.cam.td - ftd[i + k];
      const double iz = fast_rcp(pz);
      const double r0 = sqi * (px * iz - ((double)ok.x - dt * (double)ok.z)), r1 = sqi * (py * iz - ((double)ok.y - dt * (double)ok.w));
      // 0.5 * rho(s) of ceres::HuberLoss
      const double sq = r0 * r0 + r1 * r1, hb = p.huber * p.huber;
      acc[0] += sq > hb ? 0.5 * (2.0 * p.huber * (sq * fast_rsqrt(sq)) - hb) : 0.5 * sq;
    }
  }
  if (p.planes && t < 256) {  // LiDAR plane residuals at the candidate
    const int np = p.n_planes[w];
    const gf2_plane* pls = p.planes + (size_t)w * p.Pm;
    const double* palpha = p.plane_alpha ? p.plane_alpha + (size_t)w * p.Pm : nullptr;
    for (int q = t; q < np; q += 256) {
      const int f = pls[q].frame;
      const double r = pls[q].ct ? ct_plane_residual(pls[q], palpha ? palpha[q] : 0.0, pose_c + 7 * f, pose_c + 7 * (f + 1), p.lidar_sqrt_info, nullptr)
                                 : plane_residual(pls[q], S.fr[f], p.lidar_sqrt_info, nullptr);
      acc[0] += 0.5 * r * r;
    }
  }
  // IMU / wheel factors and the prior at the candidate: warp 8 alone, concurrently with the landmark warps
  if (t >= 256) {
    const int ln = t - 256;
    if (p.imu) {
      // lane k < F-1 evaluates the raw residual of interval k; the weighted cost 0.5 |sqrt_info r|^2 of all intervals is then
      // spread over the warp as (interval, row) items so that no lane walks a whole 15x15 matrix out of global memory alone
      bool ok = false;
      if (ln < F - 1) {
        const gf2_imu_preint& pre = p.imu[(size_t)w * (F - 1) + ln];
        ok = pre.valid && pre.sum_dt <= 10.0;
        if (ok) { double r[15]; ImuStates s2 = load_imu_states(pose_c, sb_c, ln); imu_raw(pre, s2, p.g_norm, r, nullptr); for (int k = 0; k < 15; k++) S.rimu[ln][k] = r[k]; }
      }
      const unsigned okmask = __ballot_sync(0xffffffffu, ok);
      __syncwarp();
      for (int it = ln; it < 15 * (F - 1); it += 32) {
        const int k = it / 15, a = it % 15;
        if (!((okmask >> k) & 1u)) continue;
        const double* sq = p.imu_sqrt + ((size_t)w * (F - 1) + k) * 225 + a * 15;
        double s2 = 0; for (int c = a; c < 15; c++) s2 += sq[c] * S.rimu[k][c];
        accx[0] += 0.5 * s2 * s2;
      }
    }
    if (p.wheel && ln >= 16 && ln - 16 < F - 1) {
      const int k = ln - 16;
      const gf2_wheel_preint& pre = p.wheel[(size_t)w * (F - 1) + k];
      if (pre.valid && pre.sum_dt <= 10.0) {
        double r[6];
        const CalibPtr cal = calib_of(p, w, true);
        wheel_raw(pre, pose_c + 7 * k, pose_c + 7 * (k + 1), cal.exw, cal.sxw, cal.tdw[0], r, nullptr);
        accx[0] += wheel_cost(p.wheel_sqrt + ((size_t)w * (F - 1) + k) * 36, r);
      }
    }
    const int n = p.prior_rows ? p.prior_rows[w] : 0;
    if (n > 0) {
      const gf2_prior_block* blk = p.prior_blocks + (size_t)w * (2 * F + 8);
      if (ln < p.prior_nblocks[w]) prior_block_dx(blk[ln], pose_c, sb_c, calib_of(p, w, true), S.dx);
      __syncwarp();
      const double* J0 = p.prior_J0 + (size_t)w * p.Pr * p.Pr; const double* r0p = p.prior_r0 + (size_t)w * p.Pr;
      for (int row = ln; row < n; row += 32) { double s2 = r0p[row]; for (int c = 0; c < n; c++) s2 += J0[row * p.Pr + c] * S.dx[c]; accx[0] += 0.5 * s2 * s2; }
    }
  }
  block_sum<3>(acc, S.red);
  block_sum<3>(accx, S.red);
}
__global__ void __launch_bounds__(288, 2) k_cand_eval(KP p, int w0) {
  const int w = w0 + blockIdx.x;
  WinState& st = p.st[w];
  if (!st.active) return;
  if (!st.lin_valid) return;  // invalid linear solve already handled in k_solve
  __shared__ StepShared S;
  if (threadIdx.x == 0) dogleg_coefficients(st, p.c_sums + (size_t)w * 8);
  __syncthreads();
  double acc[3], accx[3];
  cand_body(p, w, st, S, acc, accx);
  if (threadIdx.x == 0) {
    double* cc = p.c_cand + (size_t)w * 4;
    cc[0] = acc[0]; cc[1] = acc[1]; cc[2] = acc[2];
    st.cand_nv = accx[0]; st.step2_x = accx[1]; st.xnorm2_x = accx[2];
  }
}

// TrustRegionMinimizer: tolerance checks, step acceptance, radius update (after the candidate sums are complete)
// acc (read by thread 0 only) = {candidate cost, |step|^2, |x|^2} over all factors / parameter blocks
__device__ __forceinline__ void decide_body(const KP& p, int w, WinState& st, int& s_decision, const double (&acc)[3]) {
  const int t = threadIdx.x, F = p.F;
  const int nlm = p.nlm[w];
  if (t == 0) {
    int decision = 0;  // 0 reject, 1 accept, 2 terminated (no accept)
    st.cand_cost = acc[0];
    st.x_norm2 = acc[2];
    st.x_cost_prev = st.x_cost;
    st.iteration++;
    if (!(st.model_cost_change > 0.0)) {  // HandleInvalidStep
      st.invalid_count++;
      st.mu *= 10.0; st.reuse = 0;
      if (st.invalid_count >= 5) { st.active = 0; st.termination = GF2_TERM_FAILURE; }
      decision = 2;
    } else {
      st.invalid_count = 0;
      const double step_norm = sqrt(acc[1]), x_norm = sqrt(acc[2]);
      const double cost_change = st.x_cost - st.cand_cost;
      if (step_norm <= p.ptol * (x_norm + p.ptol)) { st.active = 0; st.termination = GF2_TERM_PARAMETER_TOL; decision = 2; }
      else if (fabs(cost_change) <= p.ftol * st.x_cost) { st.active = 0; st.termination = GF2_TERM_FUNCTION_TOL; decision = 2; }
      else {
        const double rho = cost_change / st.model_cost_change;
        if (rho > 1e-3) {
          decision = 1; st.successful++;
          if (rho < 0.25) st.radius *= 0.5;
          if (rho > 0.75) st.radius = fmax(st.radius, 3.0 * st.dogleg_step_norm);
          st.mu = fmax(1e-8, 2.0 * st.mu / 10.0);
          st.reuse = 0; st.x_cost = st.cand_cost;
        } else { st.radius *= 0.5; st.reuse = 1; }
      }
    }
    if (st.active) {
      // TrustRegionMinimizer checks the iteration budget and then the wall clock after every iteration (MaxSolverTimeReached: "Maximum solver
      // time reached", NO_CONVERGENCE). The clock is the device's %globaltimer since k_prepare of this solve: like the reference's, the number
      // of iterations a capped solve runs depends on the machine.
      unsigned long long now = 0;
      if (p.max_time_ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (st.iteration >= p.max_iterations) { st.active = 0; st.termination = GF2_TERM_NO_CONVERGENCE; }
      else if (p.max_time_ns && now - st.t_start_ns >= p.max_time_ns) { st.active = 0; st.termination = GF2_TERM_NO_CONVERGENCE; }
      else if (st.radius <= 1e-32) { st.active = 0; st.termination = GF2_TERM_MIN_RADIUS; }
    }
    s_decision = decision;
    if (p.trace && st.iteration <= 64) {
      double* tr = p.trace + ((size_t)w * 64 + (st.iteration - 1)) * 6;
      tr[0] = st.cand_cost; tr[1] = st.model_cost_change; tr[2] = (st.x_cost_prev - st.cand_cost) / st.model_cost_change; tr[3] = st.radius; tr[4] = sqrt(acc[1]); tr[5] = decision;
    }
  }
  __syncthreads();
  if (s_decision == 1) {  // x = candidate
    double* poseW = p.pose + (size_t)w * F * 7; double* sbW = p.sb + (size_t)w * F * 9;
    const double* pose_c = p.pose_c + (size_t)w * F * 7; const double* sb_c = p.sb_c + (size_t)w * F * 9;
    for (int i = t; i < F * 7; i += blockDim.x) poseW[i] = pose_c[i];
    for (int i = t; i < F * 9; i += blockDim.x) sbW[i] = sb_c[i];
    for (int l = t; l < nlm; l += blockDim.x) p.invdep[(size_t)w * p.Lm + l] = p.invdep_c[(size_t)w * p.Lm + l];
    if (p.wcal) {
      if (t < 7) p.exw[(size_t)w * 7 + t] = p.exw_c[(size_t)w * 7 + t];
      else if (t < 10) p.sxw[(size_t)w * 3 + t - 7] = p.sxw_c[(size_t)w * 3 + t - 7];
      else if (t == 10) p.tdw[w] = p.tdw_c[w];
    }
  }
}
__global__ void __launch_bounds__(256) k_decide(KP p, int w0) {
  const int w = w0 + blockIdx.x;
  WinState& st = p.st[w];
  if (!st.active) return;
  if (!st.lin_valid) return;
  __shared__ int s_decision;
  const double* cc = p.c_cand + (size_t)w * 4;
  const double acc[3] = {cc[0] + st.cand_nv, cc[1] + st.step2_x, cc[2] + st.xnorm2_x};
  decide_body(p, w, st, s_decision, acc);
}

// ------------------------------------------------------------------------------------------------ k_step
// Single-GPU path: sweep 2, the dogleg coefficients, sweep 3 and the trust-region decision of one window in ONE launch (the sums the three
// stages exchange stay in the CTA; the factor-sharded mode all-reduces them between k_backsub, k_cand_eval and k_decide instead). The
// second sweep over the window's observation records (120 KB) finds them in L1 / L2.
__global__ void __launch_bounds__(288, 2) k_step(KP p, int w0) {
  const int w = w0 + blockIdx.x;
  WinState& st = p.st[w];
  if (!st.active || !st.lin_valid) return;   // invalid linear solve already handled in k_solve2
  __shared__ StepShared S;
  __shared__ double cs[8];
  if (!st.reuse) {   // a rejected step keeps the linearisation and its landmark sums (st.dlg2_l ..): only the radius changed
    double sums[6];
    backsub_body(p, w, st, S, sums);
    if (threadIdx.x == 0) for (int i = 0; i < 6; i++) cs[i] = sums[i];
  } else if (threadIdx.x == 0) {
    cs[0] = st.dlg2_l; cs[1] = st.gn2_l; cs[2] = st.gz_l; cs[3] = st.zHz_l; cs[4] = st.uHz_l; cs[5] = st.uHu_l;
  }
  __syncthreads();
  if (threadIdx.x == 0) dogleg_coefficients(st, cs);
  __syncthreads();
  double acc[3], accx[3];
  cand_body(p, w, st, S, acc, accx);
  const double tot[3] = {acc[0] + accx[0], acc[1] + accx[1], acc[2] + accx[2]};
  decide_body(p, w, st, S.decision, tot);
}

}  // namespace gf2
