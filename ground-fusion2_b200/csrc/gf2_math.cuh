// gf2_math.cuh — small fixed-size fp64 math for the sm_100a kernels (device + host inline).
// Quaternions are [x y z w] (the wire order of para_Pose[i][3..6], VE/estimator/estimator.cpp:2345-2348).
// The formulas restate Eigen's Quaterniond ops and VE/utility/utility.h:23-76 so that the kernels evaluate the
// same expressions as the reference factors; written from scratch for CUDA (no Eigen).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define GF2_HD __host__ __device__ __forceinline__

namespace gf2 {

struct V3 { double x, y, z; };
struct Q4 { double x, y, z, w; };
struct M3 { double m[9]; };  // row-major

GF2_HD V3 mk3(double x, double y, double z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
GF2_HD V3 ld3(const double* p) { return mk3(p[0], p[1], p[2]); }
GF2_HD Q4 ldq(const double* p) { Q4 q; q.x = p[0]; q.y = p[1]; q.z = p[2]; q.w = p[3]; return q; }
GF2_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
GF2_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
GF2_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
GF2_HD V3 operator*(double s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
GF2_HD V3 operator*(V3 a, double s) { return mk3(s * a.x, s * a.y, s * a.z); }
GF2_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GF2_HD V3 cross(V3 a, V3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
GF2_HD double norm2(V3 a) { return dot(a, a); }
GF2_HD double get(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

GF2_HD M3 eye3() { M3 r; for (int i = 0; i < 9; i++) r.m[i] = 0; r.m[0] = r.m[4] = r.m[8] = 1; return r; }
GF2_HD M3 zero3() { M3 r; for (int i = 0; i < 9; i++) r.m[i] = 0; return r; }
GF2_HD M3 skew(V3 q) { M3 r; r.m[0] = 0; r.m[1] = -q.z; r.m[2] = q.y; r.m[3] = q.z; r.m[4] = 0; r.m[5] = -q.x; r.m[6] = -q.y; r.m[7] = q.x; r.m[8] = 0; return r; }
GF2_HD M3 mul(const M3& a, const M3& b) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
  return r;
}
GF2_HD M3 mulT(const M3& a, const M3& b) {  // a^T b
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i * 3 + j] = a.m[i] * b.m[j] + a.m[3 + i] * b.m[3 + j] + a.m[6 + i] * b.m[6 + j];
  return r;
}
GF2_HD M3 mulBT(const M3& a, const M3& b) {  // a b^T
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i * 3 + j] = a.m[i * 3] * b.m[j * 3] + a.m[i * 3 + 1] * b.m[j * 3 + 1] + a.m[i * 3 + 2] * b.m[j * 3 + 2];
  return r;
}
GF2_HD M3 transpose(const M3& a) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i * 3 + j] = a.m[j * 3 + i]; return r; }
GF2_HD M3 add(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] + b.m[i]; return r; }
GF2_HD M3 sub(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] - b.m[i]; return r; }
GF2_HD M3 scale(const M3& a, double s) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] * s; return r; }
GF2_HD V3 mul(const M3& a, V3 v) { return mk3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z); }
GF2_HD V3 mulT(const M3& a, V3 v) { return mk3(a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z, a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z); }

// Eigen QuaternionBase::toRotationMatrix
GF2_HD M3 toR(Q4 q) {
  M3 r;
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r.m[0] = 1 - (tyy + tzz); r.m[1] = txy - twz; r.m[2] = txz + twy;
  r.m[3] = txy + twz; r.m[4] = 1 - (txx + tzz); r.m[5] = tyz - twx;
  r.m[6] = txz - twy; r.m[7] = tyz + twx; r.m[8] = 1 - (txx + tyy);
  return r;
}
GF2_HD Q4 qmul(Q4 a, Q4 b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
GF2_HD Q4 qinv(Q4 q) { double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w; Q4 r; r.x = -q.x / n2; r.y = -q.y / n2; r.z = -q.z / n2; r.w = q.w / n2; return r; }
GF2_HD Q4 qnormalized(Q4 q) { double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w); Q4 r; r.x = q.x / n; r.y = q.y / n; r.z = q.z / n; r.w = q.w / n; return r; }
GF2_HD V3 qvec(Q4 q) { return mk3(q.x, q.y, q.z); }
// Eigen _transformVector
GF2_HD V3 qrot(Q4 q, V3 v) { V3 u = qvec(q); V3 uv = cross(u, v); uv = uv + uv; return v + q.w * uv + cross(u, uv); }
// Utility::deltaQ, VE/utility/utility.h:23-36
GF2_HD Q4 deltaQ(V3 th) { Q4 q; q.w = 1.0; q.x = th.x / 2.0; q.y = th.y / 2.0; q.z = th.z / 2.0; return qnormalized(q); }
// bottom-right 3x3 of Utility::Qleft(q) / Qright(q) (VE/utility/utility.h:58-76): w I +/- skew(vec)
GF2_HD M3 QleftBR(Q4 q) { M3 s = skew(qvec(q)); M3 r = s; r.m[0] += q.w; r.m[4] += q.w; r.m[8] += q.w; return r; }
GF2_HD M3 QrightBR(Q4 q) { M3 s = skew(qvec(q)); M3 r = scale(s, -1.0); r.m[0] += q.w; r.m[4] += q.w; r.m[8] += q.w; return r; }
// bottom-right 3x3 of Qleft(a) * Qright(b):  (4x4 product, rows/cols 1..3)  = a_v b_v^T ... computed explicitly
GF2_HD M3 QleftQrightBR(Q4 a, Q4 b) {
  // L = [[aw, -av^T],[av, aw I + [av]x]], R = [[bw, -bv^T],[bv, bw I - [bv]x]]; BR = -av bv^T... wait sign: row block 2 of L times col block 2 of R
  // = av * (-bv^T) + (aw I + [av]x)(bw I - [bv]x)
  V3 av = qvec(a), bv = qvec(b);
  M3 A = QleftBR(a), B = QrightBR(b);
  M3 r = mul(A, B);
  r.m[0] -= av.x * bv.x; r.m[1] -= av.x * bv.y; r.m[2] -= av.x * bv.z;
  r.m[3] -= av.y * bv.x; r.m[4] -= av.y * bv.y; r.m[5] -= av.y * bv.z;
  r.m[6] -= av.z * bv.x; r.m[7] -= av.z * bv.y; r.m[8] -= av.z * bv.z;
  return r;
}

#ifdef __CUDACC__
// Branch-free reciprocal / reciprocal square root for normal, positive arguments (depths, squared norms): the hardware
// approximation (measured 1e-6 relative, scripts/ubench/approx_accuracy.cu) refined by TWO Newton steps: 1e-6 -> 1e-12 -> 1.1e-16
// (rcp) / 2.2e-16 (rsqrt), i.e. already at the rounding level, a third step changes nothing (profiles/ubench_approx_accuracy_r2.txt).
// The library versions carry special-case branches that split the basic block the kernel wants to software-pipeline.
__device__ __forceinline__ double fast_rcp(double x) {
  double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#pragma unroll
  for (int it = 0; it < 2; it++) { const double e = fma(-x, r, 1.0); r = fma(r, e, r); }
  return r;
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#pragma unroll
  for (int it = 0; it < 2; it++) { const double e = fma(-x * r, r, 1.0); r = fma(0.5 * r, e, r); }
  return r;
}
#endif

// Sophus SO3::exp / log and the SO(3) right Jacobians (VE/utility/sophus_utils.hpp:155-236), used by the wheel factor.
GF2_HD Q4 so3Exp(V3 om) {
  const double eps = 1e-10;
  double t2 = norm2(om), imag, real;
  if (t2 < eps * eps) { double t4 = t2 * t2; imag = 0.5 - (1.0 / 48.0) * t2 + (1.0 / 3840.0) * t4; real = 1.0 - (1.0 / 8.0) * t2 + (1.0 / 384.0) * t4; }
  else { double t = sqrt(t2), h = 0.5 * t; imag = sin(h) / t; real = cos(h); }
  Q4 q; q.w = real; q.x = imag * om.x; q.y = imag * om.y; q.z = imag * om.z; return q;
}
GF2_HD V3 so3Log(Q4 qin) {
  const double eps = 1e-10;
  Q4 q = qnormalized(qin);
  double sn = q.x * q.x + q.y * q.y + q.z * q.z, w = q.w, f;
  if (sn < eps * eps) { double w2 = w * w; f = 2.0 / w - (2.0 / 3.0) * sn / (w * w2); }
  else { double n = sqrt(sn); if (fabs(w) < eps) f = (w > 0 ? M_PI : -M_PI) / n; else f = 2.0 * atan(n / w) / n; }
  return f * qvec(q);
}
GF2_HD M3 rightJacobianSO3(V3 phi) {
  double n2 = norm2(phi); M3 h = skew(phi), h2 = mul(h, h), J = eye3();
  if (n2 > 1e-10) { double n = sqrt(n2), n3 = n2 * n; J = sub(J, scale(h, (1 - cos(n)) / n2)); J = add(J, scale(h2, (n - sin(n)) / n3)); }
  else { J = sub(J, scale(h, 0.5)); J = add(J, scale(h2, 1.0 / 6.0)); }
  return J;
}
GF2_HD M3 rightJacobianInvSO3(V3 phi) {
  double n2 = norm2(phi); M3 h = skew(phi), h2 = mul(h, h), J = eye3();
  J = add(J, scale(h, 0.5));
  if (n2 > 1e-10) {
    double n = sqrt(n2);
    if (n < M_PI - 1e-5) J = add(J, scale(h2, 1.0 / n2 - (1.0 + cos(n)) / (2.0 * n * sin(n))));
    else J = add(J, scale(h2, 1.0 / (M_PI * M_PI)));
  } else J = add(J, scale(h2, 1.0 / 12.0));
  return J;
}

}  // namespace gf2
