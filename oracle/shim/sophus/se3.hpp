// TEST INFRASTRUCTURE — the part of Sophus the reference's wheel factor and utility/sophus_utils.hpp touch, so that those
// sources compile unmodified: SO3<double> with exp / log / hat / matrix / unit_quaternion / group product, formulas as in the
// Sophus copy vendored by the reference (Ground-Fusion++/lio/thirdparty/sophus/so3.hpp:246-256 hat, :260-311 logAndTheta,
// :349-362 product, :638-667 expAndTheta; common.hpp:117 epsilon = 1e-10). SE3 / Sim3 / RxSO3 are declared only (the function
// templates of sophus_utils.hpp that use them are never instantiated on this path).
#pragma once
#include <cmath>
#include <Eigen/Dense>
#define EIGEN_STATIC_ASSERT_FIXED_SIZE(T)
namespace Sophus {
template <class Scalar> struct Constants {
  static Scalar epsilon() { return Scalar(1e-10); }
  static Scalar epsilonSqrt() { return std::sqrt(epsilon()); }
  static Scalar pi() { return Scalar(3.141592653589793238462643383279502884); }
};
template <class Scalar> using Vector3 = Eigen::Matrix<Scalar, 3, 1>;
template <class Scalar> using Matrix3 = Eigen::Matrix<Scalar, 3, 3>;
template <class Scalar_>
class SO3 {
 public:
  typedef Scalar_ Scalar;
  typedef Eigen::Matrix<Scalar, 3, 1> Tangent;
  typedef Eigen::Matrix<Scalar, 3, 1> Point;
  typedef Eigen::Matrix<Scalar, 3, 3> Transformation;
  typedef Eigen::Matrix<Scalar, 3, 3> Adjoint;
  SO3() : q_(1, 0, 0, 0) {}
  template <class D> explicit SO3(const Eigen::QuaternionBase<D>& q) : q_(q) { q_.normalize(); }   // so3.hpp: the constructor re-normalises
  template <class D> explicit SO3(const Eigen::MatrixBase<D>& R) : q_(R) { q_.normalize(); }
  const Eigen::Quaternion<Scalar>& unit_quaternion() const { return q_; }
  Transformation matrix() const { return q_.toRotationMatrix(); }
  SO3 inverse() const { SO3 r; r.q_ = q_.conjugate(); return r; }
  SO3 operator*(const SO3& o) const {
    const Eigen::Quaternion<Scalar>&a = q_, &b = o.q_;
    SO3 r;
    r.q_ = Eigen::Quaternion<Scalar>(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(), a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                                     a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(), a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
    return r;
  }
  template <class D> Point operator*(const Eigen::MatrixBase<D>& p) const {
    Point uv = q_.vec().cross(p);
    uv += uv;
    return p + q_.w() * uv + q_.vec().cross(uv);
  }
  static Transformation hat(const Tangent& omega) {
    Transformation Omega;
    Omega << Scalar(0), -omega(2), omega(1), omega(2), Scalar(0), -omega(0), -omega(1), omega(0), Scalar(0);
    return Omega;
  }
  template <class D> static SO3 exp(const Eigen::MatrixBase<D>& omega_) {
    const Tangent omega = omega_;
    const Scalar theta_sq = omega.squaredNorm();
    Scalar imag_factor, real_factor;
    if (theta_sq < Constants<Scalar>::epsilon() * Constants<Scalar>::epsilon()) {
      const Scalar theta_po4 = theta_sq * theta_sq;
      imag_factor = Scalar(0.5) - Scalar(1.0 / 48.0) * theta_sq + Scalar(1.0 / 3840.0) * theta_po4;
      real_factor = Scalar(1) - Scalar(1.0 / 8.0) * theta_sq + Scalar(1.0 / 384.0) * theta_po4;
    } else {
      const Scalar theta = std::sqrt(theta_sq), half_theta = Scalar(0.5) * theta;
      imag_factor = std::sin(half_theta) / theta;
      real_factor = std::cos(half_theta);
    }
    SO3 q;
    q.q_ = Eigen::Quaternion<Scalar>(real_factor, imag_factor * omega.x(), imag_factor * omega.y(), imag_factor * omega.z());
    return q;
  }
  Tangent log() const {
    const Scalar squared_n = q_.vec().squaredNorm(), w = q_.w();
    Scalar two_atan_nbyw_by_n;
    if (squared_n < Constants<Scalar>::epsilon() * Constants<Scalar>::epsilon()) {
      const Scalar squared_w = w * w;
      two_atan_nbyw_by_n = Scalar(2) / w - Scalar(2.0 / 3.0) * (squared_n) / (w * squared_w);
    } else {
      const Scalar n = std::sqrt(squared_n);
      if (std::abs(w) < Constants<Scalar>::epsilon()) two_atan_nbyw_by_n = (w > Scalar(0) ? Constants<Scalar>::pi() : -Constants<Scalar>::pi()) / n;
      else two_atan_nbyw_by_n = Scalar(2) * std::atan(n / w) / n;
    }
    Tangent t = two_atan_nbyw_by_n * q_.vec();
    return t;
  }
 private:
  Eigen::Quaternion<Scalar> q_;
};
typedef SO3<double> SO3d;
template <class Scalar> class RxSO3 {
 public:
  typedef Eigen::Matrix<Scalar, 4, 1> Tangent;
  template <class D> static RxSO3 exp(const Eigen::MatrixBase<D>&) { return RxSO3(); }
  Tangent log() const { return Tangent(); }
};
template <class Scalar> class SE3 {
 public:
  typedef Eigen::Matrix<Scalar, 6, 1> Tangent;
  SE3() {}
  template <class D> SE3(const SO3<Scalar>& r, const Eigen::MatrixBase<D>& t) : so3_(r), t_(t) {}
  const SO3<Scalar>& so3() const { return so3_; }
  const Eigen::Matrix<Scalar, 3, 1>& translation() const { return t_; }
 private:
  SO3<Scalar> so3_; Eigen::Matrix<Scalar, 3, 1> t_;
};
template <class Scalar> class Sim3 {
 public:
  typedef Eigen::Matrix<Scalar, 7, 1> Tangent;
  Sim3() {}
  template <class D> Sim3(const RxSO3<Scalar>& r, const Eigen::MatrixBase<D>& t) : r_(r), t_(t) {}
  const RxSO3<Scalar>& rxso3() const { return r_; }
  const Eigen::Matrix<Scalar, 3, 1>& translation() const { return t_; }
 private:
  RxSO3<Scalar> r_; Eigen::Matrix<Scalar, 3, 1> t_;
};
typedef SE3<double> SE3d;
}  // namespace Sophus
