// TEST INFRASTRUCTURE — CPU oracle for the Ground-Fusion++ hot path. Not product code: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use anything in oracle/.
//
// gf2o_linalg.h — minimal fixed-size dense linear algebra standing in for the Eigen 3.3 types the
// reference uses (no Eigen exists in this container). Semantics follow Eigen where they matter for
// parity: Quaterniond(w,x,y,z) constructor order, toRotationMatrix(), q*v, q.inverse() = conj/|q|^2,
// normalized(), LLT lower factor, matrix inverse via partial-pivot LU.
#pragma once
#include <cmath>
#include <cstring>
#include <cstdio>
#include <vector>
#include <algorithm>

namespace gf2o {

template <int R, int C>
struct Mat {
  double a[R * C];
  Mat() { for (int i = 0; i < R * C; i++) a[i] = 0.0; }
  double& operator()(int r, int c) { return a[r * C + c]; }
  double operator()(int r, int c) const { return a[r * C + c]; }
  double& operator[](int i) { return a[i]; }
  double operator[](int i) const { return a[i]; }
  static Mat Identity() { Mat m; for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = 1.0; return m; }
  Mat<C, R> T() const { Mat<C, R> t; for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) t(c, r) = (*this)(r, c); return t; }
  template <int BR, int BC> Mat<BR, BC> block(int r0, int c0) const {
    Mat<BR, BC> b; for (int r = 0; r < BR; r++) for (int c = 0; c < BC; c++) b(r, c) = (*this)(r0 + r, c0 + c); return b; }
  template <int BR, int BC> void setBlock(int r0, int c0, const Mat<BR, BC>& b) {
    for (int r = 0; r < BR; r++) for (int c = 0; c < BC; c++) (*this)(r0 + r, c0 + c) = b(r, c); }
  double squaredNorm() const { double s = 0; for (int i = 0; i < R * C; i++) s += a[i] * a[i]; return s; }
  double norm() const { return std::sqrt(squaredNorm()); }
  double maxCoeff() const { double m = a[0]; for (int i = 1; i < R * C; i++) m = std::max(m, a[i]); return m; }
  double minCoeff() const { double m = a[0]; for (int i = 1; i < R * C; i++) m = std::min(m, a[i]); return m; }
};

template <int R, int C> Mat<R, C> operator+(const Mat<R, C>& x, const Mat<R, C>& y) { Mat<R, C> z; for (int i = 0; i < R * C; i++) z.a[i] = x.a[i] + y.a[i]; return z; }
template <int R, int C> Mat<R, C> operator-(const Mat<R, C>& x, const Mat<R, C>& y) { Mat<R, C> z; for (int i = 0; i < R * C; i++) z.a[i] = x.a[i] - y.a[i]; return z; }
template <int R, int C> Mat<R, C> operator-(const Mat<R, C>& x) { Mat<R, C> z; for (int i = 0; i < R * C; i++) z.a[i] = -x.a[i]; return z; }
template <int R, int C> Mat<R, C> operator*(const Mat<R, C>& x, double s) { Mat<R, C> z; for (int i = 0; i < R * C; i++) z.a[i] = x.a[i] * s; return z; }
template <int R, int C> Mat<R, C> operator*(double s, const Mat<R, C>& x) { return x * s; }
template <int R, int C> Mat<R, C> operator/(const Mat<R, C>& x, double s) { Mat<R, C> z; for (int i = 0; i < R * C; i++) z.a[i] = x.a[i] / s; return z; }
template <int R, int K, int C> Mat<R, C> operator*(const Mat<R, K>& x, const Mat<K, C>& y) {
  Mat<R, C> z;
  for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) { double s = 0; for (int k = 0; k < K; k++) s += x(r, k) * y(k, c); z(r, c) = s; }
  return z;
}
template <int R, int C> double dot(const Mat<R, C>& x, const Mat<R, C>& y) { double s = 0; for (int i = 0; i < R * C; i++) s += x.a[i] * y.a[i]; return s; }

typedef Mat<3, 1> V3;
typedef Mat<3, 3> M3;

inline V3 v3(double x, double y, double z) { V3 v; v[0] = x; v[1] = y; v[2] = z; return v; }
inline V3 cross(const V3& a, const V3& b) { return v3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]); }
// Utility::skewSymmetric, VE/utility/utility.h:38-46
inline M3 skew(const V3& q) { M3 m; m(0, 1) = -q[2]; m(0, 2) = q[1]; m(1, 0) = q[2]; m(1, 2) = -q[0]; m(2, 0) = -q[1]; m(2, 1) = q[0]; return m; }
inline M3 diag3(double a, double b, double c) { M3 m; m(0, 0) = a; m(1, 1) = b; m(2, 2) = c; return m; }

// Eigen::Quaterniond stand-in. Constructor order (w, x, y, z) as Eigen's.
struct Quat {
  double w, x, y, z;
  Quat() : w(1), x(0), y(0), z(0) {}
  Quat(double w_, double x_, double y_, double z_) : w(w_), x(x_), y(y_), z(z_) {}
  static Quat fromXYZW(const double* p) { return Quat(p[3], p[0], p[1], p[2]); }
  V3 vec() const { return v3(x, y, z); }
  double squaredNorm() const { return w * w + x * x + y * y + z * z; }
  Quat conjugate() const { return Quat(w, -x, -y, -z); }
  // Eigen: inverse() = conjugate() / squaredNorm() (zero quaternion not handled, as in unit use)
  Quat inverse() const { double n2 = squaredNorm(); return Quat(w / n2, -x / n2, -y / n2, -z / n2); }
  Quat normalized() const { double n = std::sqrt(squaredNorm()); return Quat(w / n, x / n, y / n, z / n); }
  void normalize() { *this = normalized(); }
  // Eigen QuaternionBase::toRotationMatrix()
  M3 toRotationMatrix() const {
    M3 r;
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
    r(1, 0) = txy + twz; r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = 1 - (txx + tyy);
    return r;
  }
};
inline Quat operator*(const Quat& a, const Quat& b) {
  return Quat(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
              a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
              a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x);
}
// Eigen QuaternionBase::_transformVector: v + 2w(u x v) + 2 u x (u x v)
inline V3 operator*(const Quat& q, const V3& v) {
  V3 u = q.vec();
  V3 uv = cross(u, v);
  uv = uv + uv;
  return v + q.w * uv + cross(u, uv);
}
// Eigen Quaternion(Matrix3) constructor (Shoemake), used by the generator and double2vector.
inline Quat quatFromMatrix(const M3& m) {
  Quat q;
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0) {
    t = std::sqrt(t + 1.0); q.w = 0.5 * t; t = 0.5 / t;
    q.x = (m(2, 1) - m(1, 2)) * t; q.y = (m(0, 2) - m(2, 0)) * t; q.z = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0; if (m(1, 1) > m(0, 0)) i = 1; if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    double qv[3]; qv[i] = 0.5 * t; t = 0.5 / t;
    q.w = (m(k, j) - m(j, k)) * t; qv[j] = (m(j, i) + m(i, j)) * t; qv[k] = (m(k, i) + m(i, k)) * t;
    q.x = qv[0]; q.y = qv[1]; q.z = qv[2];
  }
  return q;
}

// Utility::deltaQ, VE/utility/utility.h:23-36
inline Quat deltaQ(const V3& theta) { Quat dq(1.0, theta[0] / 2.0, theta[1] / 2.0, theta[2] / 2.0); dq.normalize(); return dq; }
// Utility::Qleft / Qright, VE/utility/utility.h:58-76 — 4x4 in [w x y z] order; positify is the identity (:49-56)
inline Mat<4, 4> Qleft(const Quat& q) {
  Mat<4, 4> a; V3 v = q.vec(); M3 b = q.w * M3::Identity() + skew(v);
  a(0, 0) = q.w; for (int i = 0; i < 3; i++) { a(0, 1 + i) = -v[i]; a(1 + i, 0) = v[i]; }
  a.setBlock<3, 3>(1, 1, b); return a;
}
inline Mat<4, 4> Qright(const Quat& p) {
  Mat<4, 4> a; V3 v = p.vec(); M3 b = p.w * M3::Identity() - skew(v);
  a(0, 0) = p.w; for (int i = 0; i < 3; i++) { a(0, 1 + i) = -v[i]; a(1 + i, 0) = v[i]; }
  a.setBlock<3, 3>(1, 1, b); return a;
}

// ---- Sophus SO3 pieces used by the wheel factor (Sophus is un-vendored for VINS; the LIO copy at
// Ground-Fusion++/lio/thirdparty/sophus/so3.hpp:260-311 (log) and :638-668 (exp) is the restated source;
// Constants<double>::epsilon() = 1e-10, common.hpp:117).
inline Quat so3Exp(const V3& omega) {
  const double eps = 1e-10;
  double theta_sq = omega.squaredNorm();
  double imag, real;
  if (theta_sq < eps * eps) {
    double theta_po4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    double theta = std::sqrt(theta_sq);
    double half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  return Quat(real, imag * omega[0], imag * omega[1], imag * omega[2]);
}
// SO3(quaternion) normalises its input (so3.hpp:539-544); log() is atan-based.
inline V3 so3Log(const Quat& q_in) {
  const double eps = 1e-10;
  Quat q = q_in.normalized();
  double squared_n = q.x * q.x + q.y * q.y + q.z * q.z;
  double w = q.w;
  double two_atan_nbyw_by_n;
  if (squared_n < eps * eps) {
    double squared_w = w * w;
    two_atan_nbyw_by_n = 2.0 / w - (2.0 / 3.0) * squared_n / (w * squared_w);
  } else {
    double n = std::sqrt(squared_n);
    if (std::fabs(w) < eps) {
      two_atan_nbyw_by_n = (w > 0 ? M_PI : -M_PI) / n;
    } else {
      two_atan_nbyw_by_n = 2.0 * std::atan(n / w) / n;
    }
  }
  return two_atan_nbyw_by_n * q.vec();
}
// Sophus::rightJacobianSO3 / rightJacobianInvSO3, VE/utility/sophus_utils.hpp:155-236
inline M3 rightJacobianSO3(const V3& phi) {
  double n2 = phi.squaredNorm();
  M3 hat = skew(phi), hat2 = hat * hat;
  M3 J = M3::Identity();
  if (n2 > 1e-10) {
    double n = std::sqrt(n2), n3 = n2 * n;
    J = J - hat * ((1 - std::cos(n)) / n2);
    J = J + hat2 * ((n - std::sin(n)) / n3);
  } else {
    J = J - hat / 2.0;
    J = J + hat2 / 6.0;
  }
  return J;
}
inline M3 rightJacobianInvSO3(const V3& phi) {
  double n2 = phi.squaredNorm();
  M3 hat = skew(phi), hat2 = hat * hat;
  M3 J = M3::Identity();
  J = J + hat / 2.0;
  if (n2 > 1e-10) {
    double n = std::sqrt(n2);
    if (n < M_PI - std::sqrt(1e-10)) {
      J = J + hat2 * (1.0 / n2 - (1.0 + std::cos(n)) / (2.0 * n * std::sin(n)));
    } else {
      J = J + hat2 / (M_PI * M_PI);
    }
  } else {
    J = J + hat2 / 12.0;
  }
  return J;
}

// ---- dynamic helpers (row-major n x n in std::vector) --------------------------------------
// Inverse by partial-pivot Gauss-Jordan (Eigen's generic inverse() for N > 4 is PartialPivLU-based).
inline bool invertN(int n, const double* A, double* Ainv) {
  std::vector<double> m(n * 2 * n);
  for (int r = 0; r < n; r++) { for (int c = 0; c < n; c++) { m[r * 2 * n + c] = A[r * n + c]; m[r * 2 * n + n + c] = (r == c); } }
  for (int col = 0; col < n; col++) {
    int piv = col; double best = std::fabs(m[col * 2 * n + col]);
    for (int r = col + 1; r < n; r++) { double v = std::fabs(m[r * 2 * n + col]); if (v > best) { best = v; piv = r; } }
    if (best == 0.0) return false;
    if (piv != col) for (int c = 0; c < 2 * n; c++) std::swap(m[piv * 2 * n + c], m[col * 2 * n + c]);
    double d = m[col * 2 * n + col];
    for (int c = 0; c < 2 * n; c++) m[col * 2 * n + c] /= d;
    for (int r = 0; r < n; r++) if (r != col) {
      double f = m[r * 2 * n + col]; if (f == 0.0) continue;
      for (int c = 0; c < 2 * n; c++) m[r * 2 * n + c] -= f * m[col * 2 * n + c];
    }
  }
  for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) Ainv[r * n + c] = m[r * 2 * n + n + c];
  return true;
}
// Lower Cholesky A = L L^T in place on the lower triangle (upper left untouched). false if not PD.
inline bool choleskyLower(int n, double* A) {
  for (int j = 0; j < n; j++) {
    double d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d); A[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  return true;
}
inline void choleskySolve(int n, const double* L, double* b) {
  for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[i * n + k] * b[k]; b[i] = s / L[i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < n; k++) s -= L[k * n + i] * b[k]; b[i] = s / L[i * n + i]; }
}
// sqrt_info = LLT(cov.inverse()).matrixL().transpose()  (VE/factor/imu_factor.h:73, wheel_factor.h:85)
template <int N> bool sqrtInfoFromCov(const double* cov, Mat<N, N>& sqrt_info) {
  double inv[N * N];
  if (!invertN(N, cov, inv)) return false;
  if (!choleskyLower(N, inv)) return false;
  for (int r = 0; r < N; r++) for (int c = 0; c < N; c++) sqrt_info(r, c) = (c >= r) ? inv[c * N + r] : 0.0;  // L^T
  return true;
}

// Symmetric eigen-decomposition by cyclic Jacobi (stand-in for Eigen::SelfAdjointEigenSolver in
// VE/factor/marginalization_factor.cpp:279,294). A row-major n x n (symmetric); on return evals ascending,
// evecs column k (row-major n x n: evecs[r*n+k]) the k-th eigenvector.
inline void symEigen(int n, const double* A_in, double* evals, double* evecs) {
  std::vector<double> A(A_in, A_in + n * n);
  for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) evecs[r * n + c] = (r == c);
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0; for (int r = 0; r < n; r++) for (int c = r + 1; c < n; c++) off += A[r * n + c] * A[r * n + c];
    double dg = 0; for (int r = 0; r < n; r++) dg += A[r * n + r] * A[r * n + r];
    if (off <= 1e-32 * (dg + 1e-300)) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      double apq = A[p * n + q]; if (apq == 0.0) continue;
      double app = A[p * n + p], aqq = A[q * n + q];
      double tau = (aqq - app) / (2.0 * apq);
      double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
      double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
      for (int k = 0; k < n; k++) { double akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq; }
      for (int k = 0; k < n; k++) { double apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk; }
      for (int k = 0; k < n; k++) { double vkp = evecs[k * n + p], vkq = evecs[k * n + q]; evecs[k * n + p] = c * vkp - s * vkq; evecs[k * n + q] = s * vkp + c * vkq; }
    }
  }
  std::vector<int> idx(n); for (int i = 0; i < n; i++) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return A[a * n + a] < A[b * n + b]; });
  std::vector<double> ev(n * n);
  for (int k = 0; k < n; k++) { evals[k] = A[idx[k] * n + idx[k]]; for (int r = 0; r < n; r++) ev[r * n + k] = evecs[r * n + idx[k]]; }
  std::memcpy(evecs, ev.data(), sizeof(double) * n * n);
}

}  // namespace gf2o
