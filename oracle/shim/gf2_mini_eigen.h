// TEST INFRASTRUCTURE — a small, eagerly evaluated subset of the Eigen 3 API, just large enough that the reference's own
// factor sources (Ground-Fusion++/vins_estimator/src/factor/*.{h,cpp}, utility/utility.{h,cpp}, lio/src/liw/lidarFactor.cpp)
// compile UNMODIFIED from /root/reference (oracle/Makefile, target _ref). Eigen itself is not installed in this image
// (SURVEY.md 8(c)); with this shim the formulas that are executed are the reference's, only the dense linear-algebra
// primitives underneath (products, LU inverse, LLT, symmetric eigen-decomposition) are restated here. Where Eigen's
// result depends on its formula (Quaternion::toRotationMatrix, _transformVector, quaternion product / slerp /
// FromTwoVectors / matrix -> quaternion) the same formula is used, so results agree with a real Eigen build to rounding.
// Every arithmetic expression returns a dynamically sized matrix by value (no expression templates, no aliasing issues).
// Only `double` scalars. Not a product file: nothing outside oracle/ and tests/ may include it.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>
#include <numeric>
#include <cstdint>
#include <string>
#include <map>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_STATIC_ASSERT_VECTOR_SPECIFIC_SIZE(T, N)
#define EIGEN_STATIC_ASSERT_MATRIX_SPECIFIC_SIZE(T, R, C)
#define EIGEN_DEVICE_FUNC
#define EIGEN_STRONG_INLINE inline

namespace Eigen {

const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1 };
enum { ComputeThinU = 4, ComputeThinV = 8, ComputeFullU = 16, ComputeFullV = 32 };
typedef std::ptrdiff_t Index;

template <class Derived> class MatrixBase;
template <class S, int R, int C, int O = 0, int MR = R, int MC = C> class Matrix;
class View;
class ArrayXd;
class BoolArray;
class DiagonalWrapper;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, 1, Dynamic> RowVectorXd;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 3, 1> Vector3f;   // declared only (never instantiated with arithmetic)
typedef MatrixXd Dyn;

namespace internal {
inline void fail(const char* what) { std::fprintf(stderr, "mini-eigen: %s\n", what); std::abort(); }
}

// ------------------------------------------------------------------------------------------------ comma initialiser
template <class M>
class CommaInitializer {
 public:
  CommaInitializer(M& m, double first) : m_(m), k_(0) { put(first); }
  CommaInitializer& operator,(double v) { put(v); return *this; }
  template <class O> CommaInitializer& operator,(const MatrixBase<O>& o) {   // a block of coefficients (vectors only appear in the sources)
    for (Index j = 0; j < o.cols(); j++) for (Index i = 0; i < o.rows(); i++) put(o(i, j));
    return *this;
  }
 private:
  void put(double v) { const Index c = m_.cols(); if (k_ >= m_.rows() * c) internal::fail("too many coefficients in a comma initialiser"); m_(k_ / c, k_ % c) = v; k_++; }
  M& m_; Index k_;
};

// ------------------------------------------------------------------------------------------------ MatrixBase
template <class Derived>
class MatrixBase {
 public:
  typedef double Scalar;
  typedef double RealScalar;
  const Derived& derived() const { return *static_cast<const Derived*>(this); }
  Derived& derived() { return *static_cast<Derived*>(this); }
  Index rows() const { return derived().rows_(); }
  Index cols() const { return derived().cols_(); }
  Index size() const { return rows() * cols(); }
  double operator()(Index i, Index j) const { return derived().at(i, j); }
  double& operator()(Index i, Index j) { return derived().at(i, j); }
  double operator()(Index i) const { return cols() == 1 ? derived().at(i, 0) : derived().at(0, i); }
  double& operator()(Index i) { return cols() == 1 ? derived().at(i, 0) : derived().at(0, i); }
  double operator[](Index i) const { return (*this)(i); }
  double& operator[](Index i) { return (*this)(i); }
  double coeff(Index i, Index j) const { return (*this)(i, j); }
  double& coeffRef(Index i, Index j) { return (*this)(i, j); }
  double x() const { return (*this)(0); } double y() const { return (*this)(1); } double z() const { return (*this)(2); } double w() const { return (*this)(3); }
  double& x() { return (*this)(0); } double& y() { return (*this)(1); } double& z() { return (*this)(2); } double& w() { return (*this)(3); }

  // ---- windows (always mutable views; constness is not tracked by this shim)
  inline View block(Index i, Index j, Index r, Index c) const;
  template <int BR, int BC> View block(Index i, Index j) const;
  template <int N> View head() const; View head(Index n) const;
  template <int N> View tail() const; View tail(Index n) const;
  template <int N> View segment(Index i) const; View segment(Index i, Index n) const;
  View col(Index j) const; View row(Index i) const;
  template <int N> View leftCols() const; View leftCols(Index n) const;
  template <int N> View rightCols() const; View rightCols(Index n) const;
  template <int N> View middleCols(Index j) const; View middleCols(Index j, Index n) const;
  template <int N> View topRows() const; View topRows(Index n) const;
  template <int N> View bottomRows() const; View bottomRows(Index n) const;
  template <int N> View middleRows(Index i) const; View middleRows(Index i, Index n) const;
  template <int BR, int BC> View topLeftCorner() const; template <int BR, int BC> View topRightCorner() const;
  template <int BR, int BC> View bottomLeftCorner() const; template <int BR, int BC> View bottomRightCorner() const;
  View topLeftCorner(Index r, Index c) const; View bottomRightCorner(Index r, Index c) const;
  View diagonal() const;

  // ---- value-returning operations
  inline Dyn eval() const;
  inline Dyn transpose() const;
  inline Dyn adjoint() const;
  inline Dyn inverse() const;
  inline Dyn normalized() const;
  inline Dyn cwiseSqrt() const;
  inline Dyn cwiseAbs() const;
  inline Dyn cwiseInverse() const;
  template <class O> Dyn cwiseProduct(const MatrixBase<O>& o) const;
  template <class O> Dyn cross(const MatrixBase<O>& o) const;
  template <class O> double dot(const MatrixBase<O>& o) const { double s = 0; for (Index i = 0; i < size(); i++) s += (*this)(i) * o(i); return s; }
  template <class T> Dyn cast() const;
  inline ArrayXd array() const;
  inline DiagonalWrapper asDiagonal() const;
  double squaredNorm() const { double s = 0; for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) s += (*this)(i, j) * (*this)(i, j); return s; }
  double norm() const { return std::sqrt(squaredNorm()); }
  double sum() const { double s = 0; for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) s += (*this)(i, j); return s; }
  double trace() const { double s = 0; for (Index i = 0; i < std::min(rows(), cols()); i++) s += (*this)(i, i); return s; }
  double minCoeff() const { double s = (*this)(0, 0); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) s = std::min(s, (*this)(i, j)); return s; }
  double maxCoeff() const { double s = (*this)(0, 0); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) s = std::max(s, (*this)(i, j)); return s; }
  double determinant() const;
  bool allFinite() const { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) if (!std::isfinite((*this)(i, j))) return false; return true; }
  bool hasNaN() const { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) if ((*this)(i, j) != (*this)(i, j)) return true; return false; }
  template <class O> bool isApprox(const MatrixBase<O>& o, double prec = 1e-12) const;

  // ---- in-place
  Derived& setZero() { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) (*this)(i, j) = 0.0; return derived(); }
  Derived& setOnes() { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) (*this)(i, j) = 1.0; return derived(); }
  Derived& setConstant(double v) { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) (*this)(i, j) = v; return derived(); }
  Derived& setIdentity() { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) (*this)(i, j) = i == j ? 1.0 : 0.0; return derived(); }
  void normalize() { const double n = norm(); if (n > 0) *this /= n; }
  inline void transposeInPlace();
  template <class O> Derived& operator+=(const MatrixBase<O>& o);
  template <class O> Derived& operator-=(const MatrixBase<O>& o);
  template <class O> Derived& operator*=(const MatrixBase<O>& o);
  Derived& operator*=(double s) { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) (*this)(i, j) *= s; return derived(); }
  Derived& operator/=(double s) { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) (*this)(i, j) /= s; return derived(); }
  CommaInitializer<Derived> operator<<(double v) { return CommaInitializer<Derived>(derived(), v); }
  template <class O> CommaInitializer<Derived> operator<<(const MatrixBase<O>& o) { CommaInitializer<Derived> c(derived(), o(0, 0)); bool first = true;
    for (Index j = 0; j < o.cols(); j++) for (Index i = 0; i < o.rows(); i++) { if (first) { first = false; continue; } c, o(i, j); } return c; }
};

// ------------------------------------------------------------------------------------------------ Matrix (owning; storage always dynamic)
template <class S, int R, int C, int O, int MR, int MC>
class Matrix : public MatrixBase<Matrix<S, R, C, O, MR, MC>> {
 public:
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C, Options = O, IsRowMajor = (O & RowMajor) ? 1 : 0 };
  typedef MatrixBase<Matrix> Base;
  typedef S Scalar;
  Matrix() : r_(R > 0 ? R : 0), c_(C > 0 ? C : 0), d_((size_t)r_ * c_, 0.0) {}
  // (n): size of a dynamic vector; ignored for fixed sizes (Eigen: `Vector3d ypr(3)`)
  explicit Matrix(int n) : Matrix() { if (R == Dynamic && C == 1) resize(n, 1); else if (C == Dynamic && R == 1) resize(1, n); else if (R == Dynamic && C == Dynamic) resize(n, n); }
  explicit Matrix(Index n) : Matrix((int)n) {}
  explicit Matrix(size_t n) : Matrix((int)n) {}
  // (a, b): the two coefficients of a fixed 2-vector, else the sizes
  template <class T0, class T1> Matrix(const T0& a, const T1& b) : Matrix() {
    if (R * C == 2 && R > 0 && C > 0) { d_[0] = (double)a; d_[1] = (double)b; } else resize((Index)a, (Index)b);
  }
  Matrix(double x, double y, double z) : Matrix() { need(3); d_[0] = x; d_[1] = y; d_[2] = z; }
  Matrix(double x, double y, double z, double w) : Matrix() { need(4); d_[0] = x; d_[1] = y; d_[2] = z; d_[3] = w; }
  Matrix(const Matrix&) = default;
  Matrix(Matrix&&) = default;
  template <class D> Matrix(const MatrixBase<D>& o) : Matrix() { assign(o); }
  Matrix(const ArrayXd& a);           // defined after ArrayXd
  Matrix(const DiagonalWrapper& d);   // `Matrix3d sv = Vector3d(sx, sy, 1).asDiagonal();`
  Matrix& operator=(const Matrix& o) { assign(o); return *this; }
  Matrix& operator=(Matrix&& o) { if ((R == Dynamic || R == o.r_) && (C == Dynamic || C == o.c_)) { r_ = o.r_; c_ = o.c_; d_ = std::move(o.d_); } else assign(o); return *this; }
  template <class D> Matrix& operator=(const MatrixBase<D>& o) { assign(o); return *this; }
  Matrix& operator=(const ArrayXd& a);
  Matrix& operator=(const DiagonalWrapper& d);

  static Matrix Zero() { return Matrix(); }
  static Matrix Zero(Index r, Index c) { Matrix m; m.resize(r, c); return m; }
  static Matrix Zero(Index n) { Matrix m((int)n); return m; }
  static Matrix Ones() { Matrix m; m.setOnes(); return m; }
  static Matrix Ones(Index r, Index c) { Matrix m; m.resize(r, c); m.setOnes(); return m; }
  static Matrix Ones(Index n) { Matrix m((int)n); m.setOnes(); return m; }
  static Matrix Constant(double v) { Matrix m; m.setConstant(v); return m; }
  static Matrix Constant(Index r, Index c, double v) { Matrix m; m.resize(r, c); m.setConstant(v); return m; }
  static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
  static Matrix Identity(Index r, Index c) { Matrix m; m.resize(r, c); m.setIdentity(); return m; }
  static Matrix UnitX() { Matrix m; m(0) = 1; return m; }
  static Matrix UnitY() { Matrix m; m(1) = 1; return m; }
  static Matrix UnitZ() { Matrix m; m(2) = 1; return m; }

  void resize(Index r, Index c) {
    if ((R != Dynamic && r != R) || (C != Dynamic && c != C)) internal::fail("resize of a fixed dimension");
    r_ = r; c_ = c; d_.assign((size_t)r * c, 0.0);
  }
  void resize(Index n) { if (C == 1) resize(n, 1); else if (R == 1) resize(1, n); else internal::fail("resize(n) of a matrix"); }
  void conservativeResize(Index r, Index c) { Matrix t; t.resize(r, c); for (Index j = 0; j < std::min(c, c_); j++) for (Index i = 0; i < std::min(r, r_); i++) t.at(i, j) = at(i, j); *this = std::move(t); }
  void conservativeResize(Index n) { if (C == 1) conservativeResize(n, 1); else conservativeResize(1, n); }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }
  Index rows_() const { return r_; }
  Index cols_() const { return c_; }
  Index outerStride() const { return IsRowMajor ? c_ : r_; }
  double at(Index i, Index j) const { chk(i, j); return d_[IsRowMajor ? i * c_ + j : j * r_ + i]; }
  double& at(Index i, Index j) { chk(i, j); return d_[IsRowMajor ? i * c_ + j : j * r_ + i]; }
  using Base::operator();

 private:
  void chk(Index i, Index j) const { if (i < 0 || j < 0 || i >= r_ || j >= c_) internal::fail("index out of range"); }
  void need(Index n) { if (R == Dynamic || C == Dynamic) { if (C == 1 || (R == Dynamic && C == Dynamic)) resize(n, 1); else resize(1, n); } if (r_ * c_ != n) internal::fail("wrong number of coefficients"); }
  template <class D> void assign(const MatrixBase<D>& o) {
    Index r = o.rows(), c = o.cols();
    bool tr = false;
    // Eigen lets a row vector initialise a column vector (and vice versa)
    if (((R != Dynamic && r != R) || (C != Dynamic && c != C)) && ((R == 1 && c == 1) || (C == 1 && r == 1)) && (R == Dynamic || R == c) && (C == Dynamic || C == r)) { tr = true; std::swap(r, c); }
    if ((R != Dynamic && r != R) || (C != Dynamic && c != C)) { std::fprintf(stderr, "mini-eigen: assigning %ldx%ld to a fixed %dx%d matrix\n", (long)o.rows(), (long)o.cols(), R, C); std::abort(); }
    std::vector<double> t((size_t)r * c);
    for (Index j = 0; j < c; j++) for (Index i = 0; i < r; i++) t[IsRowMajor ? i * c + j : j * r + i] = tr ? o(j, i) : o(i, j);
    r_ = r; c_ = c; d_.swap(t);
  }
  Index r_, c_;
  std::vector<double> d_;
};

// ------------------------------------------------------------------------------------------------ View: a strided window on foreign storage
class View : public MatrixBase<View> {
 public:
  View(double* p, Index r, Index c, Index rs, Index cs) : p_(p), r_(r), c_(c), rs_(rs), cs_(cs) {}
  View(const View&) = default;
  View& operator=(const View& o) { return assignFrom(o); }   // copies coefficients (never re-seats the window)
  template <class D> View& operator=(const MatrixBase<D>& o) { return assignFrom(o); }
  inline View& operator=(const ArrayXd& a);
  Index rows_() const { return r_; }
  Index cols_() const { return c_; }
  double at(Index i, Index j) const { chk(i, j); return p_[i * rs_ + j * cs_]; }
  double& at(Index i, Index j) { chk(i, j); return p_[i * rs_ + j * cs_]; }
  double* data() { return p_; }
  const double* data() const { return p_; }
  using MatrixBase<View>::operator();
 protected:
  template <class D> View& assignFrom(const MatrixBase<D>& o) {
    Index r = o.rows(), c = o.cols();
    const bool tr = (r != r_ || c != c_) && r == c_ && c == r_ && (r == 1 || c == 1);
    if (!tr && (r != r_ || c != c_)) { std::fprintf(stderr, "mini-eigen: assigning %ldx%ld to a %ldx%ld block\n", (long)r, (long)c, (long)r_, (long)c_); std::abort(); }
    std::vector<double> t((size_t)r_ * c_);   // through a temporary: the source may overlap the window
    for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) t[(size_t)j * r_ + i] = tr ? o(j, i) : o(i, j);
    for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) p_[i * rs_ + j * cs_] = t[(size_t)j * r_ + i];
    return *this;
  }
  void chk(Index i, Index j) const { if (i < 0 || j < 0 || i >= r_ || j >= c_) internal::fail("block index out of range"); }
  double* p_; Index r_, c_, rs_, cs_;
};

// Map<Matrix<...>> / Map<const Matrix<...>>
template <class M> struct map_traits;
template <class S, int R, int C, int O, int MR, int MC> struct map_traits<Matrix<S, R, C, O, MR, MC>> { enum { Rows = R, Cols = C, RowMaj = (O & RowMajor) ? 1 : 0 }; };
template <class S, int R, int C, int O, int MR, int MC> struct map_traits<const Matrix<S, R, C, O, MR, MC>> { enum { Rows = R, Cols = C, RowMaj = (O & RowMajor) ? 1 : 0 }; };
template <class M>
class Map : public View {
  typedef map_traits<M> T;
 public:
  explicit Map(const double* p) : View(const_cast<double*>(p), T::Rows, T::Cols, T::RowMaj ? T::Cols : 1, T::RowMaj ? 1 : T::Rows) { static_assert(T::Rows > 0 && T::Cols > 0, "sizes needed"); }
  Map(const double* p, Index n) : View(const_cast<double*>(p), T::Cols == 1 ? n : 1, T::Cols == 1 ? 1 : n, 1, 1) {}
  Map(const double* p, Index r, Index c) : View(const_cast<double*>(p), r, c, T::RowMaj ? c : 1, T::RowMaj ? 1 : r) {}
  Map(const Map&) = default;
  Map& operator=(const Map& o) { View::operator=(static_cast<const View&>(o)); return *this; }
  template <class D> Map& operator=(const MatrixBase<D>& o) { View::operator=(o); return *this; }
};

// ------------------------------------------------------------------------------------------------ MatrixBase: windows
namespace internal {
template <class D> View window(const MatrixBase<D>& m, Index i, Index j, Index r, Index c) {
  if (i < 0 || j < 0 || r < 0 || c < 0 || i + r > m.rows() || j + c > m.cols()) fail("block outside the matrix");
  MatrixBase<D>& mm = const_cast<MatrixBase<D>&>(m);
  if (r == 0 || c == 0) return View(nullptr, r, c, 0, 0);
  double* p = &mm(i, j);
  const Index rs = m.rows() > i + 1 ? &mm(i + 1, j) - p : 0, cs = m.cols() > j + 1 ? &mm(i, j + 1) - p : 0;
  return View(p, r, c, rs, cs);
}
}
template <class D> View MatrixBase<D>::block(Index i, Index j, Index r, Index c) const { return internal::window(*this, i, j, r, c); }
template <class D> template <int BR, int BC> View MatrixBase<D>::block(Index i, Index j) const { return internal::window(*this, i, j, BR, BC); }
template <class D> View MatrixBase<D>::segment(Index i, Index n) const { return cols() == 1 ? block(i, 0, n, 1) : block(0, i, 1, n); }
template <class D> template <int N> View MatrixBase<D>::segment(Index i) const { return segment(i, N); }
template <class D> View MatrixBase<D>::head(Index n) const { return segment(0, n); }
template <class D> template <int N> View MatrixBase<D>::head() const { return segment(0, N); }
template <class D> View MatrixBase<D>::tail(Index n) const { return segment(size() - n, n); }
template <class D> template <int N> View MatrixBase<D>::tail() const { return segment(size() - N, N); }
template <class D> View MatrixBase<D>::col(Index j) const { return block(0, j, rows(), 1); }
template <class D> View MatrixBase<D>::row(Index i) const { return block(i, 0, 1, cols()); }
template <class D> View MatrixBase<D>::leftCols(Index n) const { return block(0, 0, rows(), n); }
template <class D> template <int N> View MatrixBase<D>::leftCols() const { return leftCols(N); }
template <class D> View MatrixBase<D>::rightCols(Index n) const { return block(0, cols() - n, rows(), n); }
template <class D> template <int N> View MatrixBase<D>::rightCols() const { return rightCols(N); }
template <class D> View MatrixBase<D>::middleCols(Index j, Index n) const { return block(0, j, rows(), n); }
template <class D> template <int N> View MatrixBase<D>::middleCols(Index j) const { return middleCols(j, N); }
template <class D> View MatrixBase<D>::topRows(Index n) const { return block(0, 0, n, cols()); }
template <class D> template <int N> View MatrixBase<D>::topRows() const { return topRows(N); }
template <class D> View MatrixBase<D>::bottomRows(Index n) const { return block(rows() - n, 0, n, cols()); }
template <class D> template <int N> View MatrixBase<D>::bottomRows() const { return bottomRows(N); }
template <class D> View MatrixBase<D>::middleRows(Index i, Index n) const { return block(i, 0, n, cols()); }
template <class D> template <int N> View MatrixBase<D>::middleRows(Index i) const { return middleRows(i, N); }
template <class D> template <int BR, int BC> View MatrixBase<D>::topLeftCorner() const { return block(0, 0, BR, BC); }
template <class D> template <int BR, int BC> View MatrixBase<D>::topRightCorner() const { return block(0, cols() - BC, BR, BC); }
template <class D> template <int BR, int BC> View MatrixBase<D>::bottomLeftCorner() const { return block(rows() - BR, 0, BR, BC); }
template <class D> template <int BR, int BC> View MatrixBase<D>::bottomRightCorner() const { return block(rows() - BR, cols() - BC, BR, BC); }
template <class D> View MatrixBase<D>::topLeftCorner(Index r, Index c) const { return block(0, 0, r, c); }
template <class D> View MatrixBase<D>::bottomRightCorner(Index r, Index c) const { return block(rows() - r, cols() - c, r, c); }
template <class D> View MatrixBase<D>::diagonal() const {
  MatrixBase<D>& mm = const_cast<MatrixBase<D>&>(*this);
  const Index n = std::min(rows(), cols());
  return View(&mm(0, 0), n, 1, n > 1 ? &mm(1, 1) - &mm(0, 0) : 0, 0);
}

// ------------------------------------------------------------------------------------------------ arithmetic (eager)
template <class D> Dyn MatrixBase<D>::eval() const { Dyn m; m.resize(rows(), cols()); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) m(i, j) = (*this)(i, j); return m; }
template <class D> Dyn MatrixBase<D>::transpose() const { Dyn m; m.resize(cols(), rows()); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) m(j, i) = (*this)(i, j); return m; }
template <class D> Dyn MatrixBase<D>::adjoint() const { return transpose(); }
template <class D> void MatrixBase<D>::transposeInPlace() { Dyn t = transpose(); derived() = t; }
template <class D> Dyn MatrixBase<D>::normalized() const { Dyn m = eval(); const double n = norm(); if (n > 0) m /= n; return m; }
template <class D> Dyn MatrixBase<D>::cwiseSqrt() const { Dyn m = eval(); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) m(i, j) = std::sqrt(m(i, j)); return m; }
template <class D> Dyn MatrixBase<D>::cwiseAbs() const { Dyn m = eval(); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) m(i, j) = std::fabs(m(i, j)); return m; }
template <class D> Dyn MatrixBase<D>::cwiseInverse() const { Dyn m = eval(); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) m(i, j) = 1.0 / m(i, j); return m; }
template <class D> template <class O> Dyn MatrixBase<D>::cwiseProduct(const MatrixBase<O>& o) const { Dyn m = eval(); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) m(i, j) *= o(i, j); return m; }
template <class D> template <class T> Dyn MatrixBase<D>::cast() const { return eval(); }
template <class D> template <class O> Dyn MatrixBase<D>::cross(const MatrixBase<O>& o) const {
  if (size() != 3 || o.size() != 3) internal::fail("cross of non 3-vectors");
  const MatrixBase<D>& a = *this;
  Dyn m; m.resize(rows(), cols());
  m(0) = a(1) * o(2) - a(2) * o(1); m(1) = a(2) * o(0) - a(0) * o(2); m(2) = a(0) * o(1) - a(1) * o(0);
  return m;
}
template <class D> template <class O> bool MatrixBase<D>::isApprox(const MatrixBase<O>& o, double prec) const {
  double d2 = 0; for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) { const double d = (*this)(i, j) - o(i, j); d2 += d * d; }
  return d2 <= prec * prec * std::min(squaredNorm(), o.squaredNorm());
}
template <class D> template <class O> D& MatrixBase<D>::operator+=(const MatrixBase<O>& o) {
  if (o.rows() != rows() || o.cols() != cols()) internal::fail("+= size mismatch");
  Dyn t = o.eval(); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) (*this)(i, j) += t(i, j); return derived();
}
template <class D> template <class O> D& MatrixBase<D>::operator-=(const MatrixBase<O>& o) {
  if (o.rows() != rows() || o.cols() != cols()) internal::fail("-= size mismatch");
  Dyn t = o.eval(); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) (*this)(i, j) -= t(i, j); return derived();
}

template <class A, class B> Dyn operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  if (a.rows() != b.rows() || a.cols() != b.cols()) internal::fail("+ size mismatch");
  Dyn m; m.resize(a.rows(), a.cols()); for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) m(i, j) = a(i, j) + b(i, j); return m;
}
template <class A, class B> Dyn operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  if (a.rows() != b.rows() || a.cols() != b.cols()) internal::fail("- size mismatch");
  Dyn m; m.resize(a.rows(), a.cols()); for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) m(i, j) = a(i, j) - b(i, j); return m;
}
template <class A> Dyn operator-(const MatrixBase<A>& a) { Dyn m; m.resize(a.rows(), a.cols()); for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) m(i, j) = -a(i, j); return m; }
template <class A> Dyn operator*(const MatrixBase<A>& a, double s) { Dyn m; m.resize(a.rows(), a.cols()); for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) m(i, j) = a(i, j) * s; return m; }
template <class A> Dyn operator*(double s, const MatrixBase<A>& a) { Dyn m; m.resize(a.rows(), a.cols()); for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) m(i, j) = s * a(i, j); return m; }
template <class A> Dyn operator/(const MatrixBase<A>& a, double s) { Dyn m; m.resize(a.rows(), a.cols()); for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) m(i, j) = a(i, j) / s; return m; }
template <class A, class B> Dyn operator*(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  if (a.cols() != b.rows()) { std::fprintf(stderr, "mini-eigen: product of %ldx%ld and %ldx%ld\n", (long)a.rows(), (long)a.cols(), (long)b.rows(), (long)b.cols()); std::abort(); }
  const Index n = a.rows(), m_ = b.cols(), k_ = a.cols();
  Dyn m; m.resize(n, m_);
  for (Index j = 0; j < m_; j++) for (Index i = 0; i < n; i++) { double s = 0; for (Index k = 0; k < k_; k++) s += a(i, k) * b(k, j); m(i, j) = s; }
  return m;
}
template <class D> template <class O> D& MatrixBase<D>::operator*=(const MatrixBase<O>& o) { Dyn t = (*this) * o; derived() = t; return derived(); }
template <class A, class B> bool operator==(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  if (a.rows() != b.rows() || a.cols() != b.cols()) return false;
  for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) if (a(i, j) != b(i, j)) return false;
  return true;
}
template <class A, class B> bool operator!=(const MatrixBase<A>& a, const MatrixBase<B>& b) { return !(a == b); }
template <class A> std::ostream& operator<<(std::ostream& os, const MatrixBase<A>& a) {
  for (Index i = 0; i < a.rows(); i++) { for (Index j = 0; j < a.cols(); j++) os << (j ? " " : "") << a(i, j); if (i + 1 < a.rows()) os << "\n"; }
  return os;
}
// 1x1 results used as scalars (`double d = (a - b).transpose() * n;`)
template <class S, int R, int C, int O, int MR, int MC> struct scalar_conv;

// ------------------------------------------------------------------------------------------------ LU inverse / determinant (partial pivoting, as Eigen's PartialPivLU for sizes > 4)
namespace internal {
inline bool lu_factor(std::vector<double>& a, Index n, std::vector<Index>& piv, int& sign) {
  piv.resize(n); sign = 1;
  for (Index k = 0; k < n; k++) {
    Index p = k; double best = std::fabs(a[k * n + k]);
    for (Index i = k + 1; i < n; i++) if (std::fabs(a[i * n + k]) > best) { best = std::fabs(a[i * n + k]); p = i; }
    piv[k] = p;
    if (p != k) { for (Index j = 0; j < n; j++) std::swap(a[k * n + j], a[p * n + j]); sign = -sign; }
    const double d = a[k * n + k];
    if (d == 0.0) continue;
    for (Index i = k + 1; i < n; i++) { const double f = a[i * n + k] /= d; if (f != 0.0) for (Index j = k + 1; j < n; j++) a[i * n + j] -= f * a[k * n + j]; }
  }
  return true;
}
}
template <class D> Dyn MatrixBase<D>::inverse() const {
  const Index n = rows();
  if (n != cols()) internal::fail("inverse of a non-square matrix");
  std::vector<double> a((size_t)n * n); for (Index i = 0; i < n; i++) for (Index j = 0; j < n; j++) a[i * n + j] = (*this)(i, j);
  std::vector<Index> piv; int sign; internal::lu_factor(a, n, piv, sign);
  Dyn inv; inv.resize(n, n);
  std::vector<double> col(n);
  for (Index c = 0; c < n; c++) {
    for (Index i = 0; i < n; i++) col[i] = i == c ? 1.0 : 0.0;
    for (Index k = 0; k < n; k++) std::swap(col[k], col[piv[k]]);
    for (Index i = 0; i < n; i++) { double s = col[i]; for (Index j = 0; j < i; j++) s -= a[i * n + j] * col[j]; col[i] = s; }
    for (Index i = n - 1; i >= 0; i--) { double s = col[i]; for (Index j = i + 1; j < n; j++) s -= a[i * n + j] * col[j]; col[i] = s / a[i * n + i]; }
    for (Index i = 0; i < n; i++) inv(i, c) = col[i];
  }
  return inv;
}
template <class D> double MatrixBase<D>::determinant() const {
  const Index n = rows();
  std::vector<double> a((size_t)n * n); for (Index i = 0; i < n; i++) for (Index j = 0; j < n; j++) a[i * n + j] = (*this)(i, j);
  std::vector<Index> piv; int sign; internal::lu_factor(a, n, piv, sign);
  double d = sign; for (Index i = 0; i < n; i++) d *= a[i * n + i];
  return d;
}

// ------------------------------------------------------------------------------------------------ LLT
template <class M>
class LLT {
 public:
  LLT() {}
  template <class D> explicit LLT(const MatrixBase<D>& a) { compute(a); }
  template <class D> LLT& compute(const MatrixBase<D>& a) {
    const Index n = a.rows(); L_.resize(n, n); ok_ = true;
    for (Index j = 0; j < n; j++) {
      double d = a(j, j); for (Index k = 0; k < j; k++) d -= L_(j, k) * L_(j, k);
      if (!(d > 0.0)) ok_ = false;
      d = std::sqrt(d); L_(j, j) = d;
      for (Index i = j + 1; i < n; i++) { double s = a(i, j); for (Index k = 0; k < j; k++) s -= L_(i, k) * L_(j, k); L_(i, j) = s / d; }
    }
    return *this;
  }
  Dyn matrixL() const { return L_; }
  Dyn matrixU() const { return L_.transpose(); }
  template <class D> Dyn solve(const MatrixBase<D>& b) const {
    const Index n = L_.rows(); Dyn x = b.eval();
    for (Index c = 0; c < x.cols(); c++) {
      for (Index i = 0; i < n; i++) { double s = x(i, c); for (Index j = 0; j < i; j++) s -= L_(i, j) * x(j, c); x(i, c) = s / L_(i, i); }
      for (Index i = n - 1; i >= 0; i--) { double s = x(i, c); for (Index j = i + 1; j < n; j++) s -= L_(j, i) * x(j, c); x(i, c) = s / L_(i, i); }
    }
    return x;
  }
  int info() const { return ok_ ? 0 : 1; }
 private:
  Dyn L_; bool ok_ = false;
};
enum { Success = 0, NumericalIssue = 1 };

// ------------------------------------------------------------------------------------------------ SelfAdjointEigenSolver (cyclic Jacobi; eigenvalues ascending like Eigen)
template <class M>
class SelfAdjointEigenSolver {
 public:
  SelfAdjointEigenSolver() {}
  template <class D> explicit SelfAdjointEigenSolver(const MatrixBase<D>& a) { compute(a); }
  template <class D> SelfAdjointEigenSolver& compute(const MatrixBase<D>& a_) {
    const Index n = a_.rows();
    std::vector<double> a((size_t)n * n), v((size_t)n * n, 0.0);
    for (Index i = 0; i < n; i++) for (Index j = 0; j < n; j++) a[i * n + j] = a_(std::max(i, j), std::min(i, j));   // lower triangle, as Eigen
    for (Index i = 0; i < n; i++) v[i * n + i] = 1.0;
    for (int sweep = 0; sweep < 100; sweep++) {
      double off = 0, diag = 0;
      for (Index i = 0; i < n; i++) { diag += a[i * n + i] * a[i * n + i]; for (Index j = i + 1; j < n; j++) off += a[i * n + j] * a[i * n + j]; }
      if (off <= 1e-32 * diag || off == 0.0) break;
      for (Index p = 0; p < n - 1; p++) for (Index q = p + 1; q < n; q++) {
        const double apq = a[p * n + q];
        if (apq == 0.0) continue;
        const double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (Index k = 0; k < n; k++) { const double akp = a[k * n + p], akq = a[k * n + q]; a[k * n + p] = c * akp - s * akq; a[k * n + q] = s * akp + c * akq; }
        for (Index k = 0; k < n; k++) { const double apk = a[p * n + k], aqk = a[q * n + k]; a[p * n + k] = c * apk - s * aqk; a[q * n + k] = s * apk + c * aqk; }
        for (Index k = 0; k < n; k++) { const double vkp = v[k * n + p], vkq = v[k * n + q]; v[k * n + p] = c * vkp - s * vkq; v[k * n + q] = s * vkp + c * vkq; }
      }
    }
    std::vector<Index> order(n); for (Index i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](Index x, Index y) { return a[x * n + x] < a[y * n + y]; });
    vals_.resize(n, 1); vecs_.resize(n, n);
    for (Index k = 0; k < n; k++) { vals_(k) = a[order[k] * n + order[k]]; for (Index i = 0; i < n; i++) vecs_(i, k) = v[i * n + order[k]]; }
    return *this;
  }
  const VectorXd& eigenvalues() const { return vals_; }
  const MatrixXd& eigenvectors() const { return vecs_; }
  int info() const { return 0; }
 private:
  VectorXd vals_; MatrixXd vecs_;
};

// declared so that uninstantiated templates of the sources parse (Utility::quaternionAverage)
template <class M> class JacobiSVD {
 public:
  template <class D> JacobiSVD(const MatrixBase<D>&, unsigned = 0) { internal::fail("JacobiSVD is not part of this shim"); }
  VectorXd singularValues() const { return VectorXd(); }
  MatrixXd matrixU() const { return MatrixXd(); }
  MatrixXd matrixV() const { return MatrixXd(); }
};

// ------------------------------------------------------------------------------------------------ arrays: (v.array() > eps).select(v.array().inverse(), 0)
class ArrayXd {
 public:
  ArrayXd() {}
  explicit ArrayXd(const Dyn& m) : m_(m) {}
  ArrayXd inverse() const { ArrayXd r(m_); for (Index j = 0; j < m_.cols(); j++) for (Index i = 0; i < m_.rows(); i++) r.m_(i, j) = 1.0 / m_(i, j); return r; }
  ArrayXd sqrt() const { ArrayXd r(m_); for (Index j = 0; j < m_.cols(); j++) for (Index i = 0; i < m_.rows(); i++) r.m_(i, j) = std::sqrt(m_(i, j)); return r; }
  ArrayXd abs() const { ArrayXd r(m_); for (Index j = 0; j < m_.cols(); j++) for (Index i = 0; i < m_.rows(); i++) r.m_(i, j) = std::fabs(m_(i, j)); return r; }
  ArrayXd square() const { ArrayXd r(m_); for (Index j = 0; j < m_.cols(); j++) for (Index i = 0; i < m_.rows(); i++) r.m_(i, j) = m_(i, j) * m_(i, j); return r; }
  inline BoolArray operator>(double t) const;
  inline BoolArray operator<(double t) const;
  inline BoolArray operator>=(double t) const;
  const Dyn& matrix() const { return m_; }
  double sum() const { return m_.sum(); }
  double maxCoeff() const { return m_.maxCoeff(); }
  double minCoeff() const { return m_.minCoeff(); }
  Dyn m_;
};
class BoolArray {
 public:
  BoolArray(Index r, Index c) : r_(r), c_(c), b_((size_t)r * c) {}
  ArrayXd select(const ArrayXd& a, double other) const { ArrayXd r(a.m_); for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) if (!b_[j * r_ + i]) r.m_(i, j) = other; return r; }
  ArrayXd select(const ArrayXd& a, const ArrayXd& o) const { ArrayXd r(a.m_); for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) if (!b_[j * r_ + i]) r.m_(i, j) = o.m_(i, j); return r; }
  bool all() const { for (char v : b_) if (!v) return false; return true; }
  bool any() const { for (char v : b_) if (v) return true; return false; }
  Index count() const { Index n = 0; for (char v : b_) n += v != 0; return n; }
  Index r_, c_; std::vector<char> b_;
};
BoolArray ArrayXd::operator>(double t) const { BoolArray b(m_.rows(), m_.cols()); for (Index j = 0; j < m_.cols(); j++) for (Index i = 0; i < m_.rows(); i++) b.b_[j * m_.rows() + i] = m_(i, j) > t; return b; }
BoolArray ArrayXd::operator<(double t) const { BoolArray b(m_.rows(), m_.cols()); for (Index j = 0; j < m_.cols(); j++) for (Index i = 0; i < m_.rows(); i++) b.b_[j * m_.rows() + i] = m_(i, j) < t; return b; }
BoolArray ArrayXd::operator>=(double t) const { BoolArray b(m_.rows(), m_.cols()); for (Index j = 0; j < m_.cols(); j++) for (Index i = 0; i < m_.rows(); i++) b.b_[j * m_.rows() + i] = m_(i, j) >= t; return b; }
template <class D> ArrayXd MatrixBase<D>::array() const { return ArrayXd(eval()); }
template <class S, int R, int C, int O, int MR, int MC> Matrix<S, R, C, O, MR, MC>::Matrix(const ArrayXd& a) : Matrix() { assign(a.m_); }
template <class S, int R, int C, int O, int MR, int MC> Matrix<S, R, C, O, MR, MC>& Matrix<S, R, C, O, MR, MC>::operator=(const ArrayXd& a) { assign(a.m_); return *this; }
View& View::operator=(const ArrayXd& a) { return assignFrom(a.m_); }

class DiagonalWrapper {
 public:
  explicit DiagonalWrapper(const Dyn& v) : v_(v) {}
  Dyn toDenseMatrix() const { const Index n = v_.size(); Dyn m; m.resize(n, n); for (Index i = 0; i < n; i++) m(i, i) = v_(i); return m; }
  operator Dyn() const { return toDenseMatrix(); }
  Dyn v_;
};
template <class D> DiagonalWrapper MatrixBase<D>::asDiagonal() const { return DiagonalWrapper(eval()); }
template <class S, int R, int C, int O, int MR, int MC> Matrix<S, R, C, O, MR, MC>::Matrix(const DiagonalWrapper& d) : Matrix() { assign(d.toDenseMatrix()); }
template <class S, int R, int C, int O, int MR, int MC> Matrix<S, R, C, O, MR, MC>& Matrix<S, R, C, O, MR, MC>::operator=(const DiagonalWrapper& d) { assign(d.toDenseMatrix()); return *this; }
template <class A> Dyn operator*(const DiagonalWrapper& d, const MatrixBase<A>& a) {
  if (d.v_.size() != a.rows()) internal::fail("diagonal product size mismatch");
  Dyn m; m.resize(a.rows(), a.cols()); for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) m(i, j) = d.v_(i) * a(i, j); return m;
}
template <class A> Dyn operator*(const MatrixBase<A>& a, const DiagonalWrapper& d) {
  if (d.v_.size() != a.cols()) internal::fail("diagonal product size mismatch");
  Dyn m; m.resize(a.rows(), a.cols()); for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) m(i, j) = a(i, j) * d.v_(j); return m;
}

// ------------------------------------------------------------------------------------------------ Quaternion (coefficients stored x y z w, as Eigen)
template <class S> class Quaternion;
template <class Derived>
class QuaternionBase {
 public:
  typedef double Scalar;
  const double* q() const { return static_cast<const Derived*>(this)->qdata(); }
  double* q() { return static_cast<Derived*>(this)->qdata(); }
  double x() const { return q()[0]; } double y() const { return q()[1]; } double z() const { return q()[2]; } double w() const { return q()[3]; }
  double& x() { return q()[0]; } double& y() { return q()[1]; } double& z() { return q()[2]; } double& w() { return q()[3]; }
  View vec() const { return View(const_cast<double*>(q()), 3, 1, 1, 1); }
  View coeffs() const { return View(const_cast<double*>(q()), 4, 1, 1, 1); }
  double squaredNorm() const { return x() * x() + y() * y() + z() * z() + w() * w(); }
  double norm() const { return std::sqrt(squaredNorm()); }
  void normalize() { const double n = norm(); for (int i = 0; i < 4; i++) q()[i] /= n; }
  inline Quaternion<double> normalized() const;
  inline Quaternion<double> conjugate() const;
  inline Quaternion<double> inverse() const;
  inline Matrix3d toRotationMatrix() const;
  inline Matrix3d matrix() const { return toRotationMatrix(); }
  Derived& setIdentity() { x() = y() = z() = 0.0; w() = 1.0; return *static_cast<Derived*>(this); }
  template <class O> double dot(const QuaternionBase<O>& o) const { return x() * o.x() + y() * o.y() + z() * o.z() + w() * o.w(); }
  template <class O> inline Quaternion<double> slerp(double t, const QuaternionBase<O>& other) const;
  template <class O> double angularDistance(const QuaternionBase<O>& o) const;
  template <class T> Quaternion<double> cast() const;
  // Eigen's QuaternionBase::_transformVector
  // (RotationBase::operator*: a 3x3 matrix operand is multiplied by toRotationMatrix(), a 3-vector goes through _transformVector)
  template <class V> Dyn operator*(const MatrixBase<V>& v) const {
    if (v.rows() == 3 && v.cols() == 3) return toRotationMatrix() * v;
    if (v.size() != 3) internal::fail("quaternion * non 3-vector");
    const double ux = y() * v(2) - z() * v(1), uy = z() * v(0) - x() * v(2), uz = x() * v(1) - y() * v(0);   // vec x v
    const double uvx = 2.0 * ux, uvy = 2.0 * uy, uvz = 2.0 * uz;
    Dyn r; r.resize(3, 1);
    r(0) = v(0) + w() * uvx + (y() * uvz - z() * uvy);
    r(1) = v(1) + w() * uvy + (z() * uvx - x() * uvz);
    r(2) = v(2) + w() * uvz + (x() * uvy - y() * uvx);
    return r;
  }
};
template <class S>
class Quaternion : public QuaternionBase<Quaternion<S>> {
 public:
  typedef S Scalar;
  Quaternion() { c_[0] = c_[1] = c_[2] = 0; c_[3] = 1; }   // Eigen leaves it uninitialised; identity is a safe superset
  Quaternion(double w, double x, double y, double z) { c_[0] = x; c_[1] = y; c_[2] = z; c_[3] = w; }
  explicit Quaternion(const double* xyzw) { for (int i = 0; i < 4; i++) c_[i] = xyzw[i]; }
  template <class O> Quaternion(const QuaternionBase<O>& o) { for (int i = 0; i < 4; i++) c_[i] = o.q()[i]; }
  template <class D> explicit Quaternion(const MatrixBase<D>& m) { *this = m; }
  template <class O> Quaternion& operator=(const QuaternionBase<O>& o) { for (int i = 0; i < 4; i++) c_[i] = o.q()[i]; return *this; }
  // rotation matrix -> quaternion (Eigen's quaternionbase_assign_impl<Other, 3, 3>) or 4-vector of coefficients
  template <class D> Quaternion& operator=(const MatrixBase<D>& m) {
    if (m.rows() == 4 && m.cols() == 1) { for (int i = 0; i < 4; i++) c_[i] = m(i); return *this; }
    if (m.rows() != 3 || m.cols() != 3) internal::fail("quaternion from a non 3x3 matrix");
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.0) {
      t = std::sqrt(t + 1.0); c_[3] = 0.5 * t; t = 0.5 / t;
      c_[0] = (m(2, 1) - m(1, 2)) * t; c_[1] = (m(0, 2) - m(2, 0)) * t; c_[2] = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0; if (m(1, 1) > m(0, 0)) i = 1; if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
      c_[i] = 0.5 * t; t = 0.5 / t;
      c_[3] = (m(k, j) - m(j, k)) * t; c_[j] = (m(j, i) + m(i, j)) * t; c_[k] = (m(k, i) + m(i, k)) * t;
    }
    return *this;
  }
  static Quaternion Identity() { return Quaternion(1, 0, 0, 0); }
  // Eigen's setFromTwoVectors (the nearly-opposite branch needs an SVD and is not restated)
  template <class A, class B> static Quaternion FromTwoVectors(const MatrixBase<A>& a, const MatrixBase<B>& b) {
    Dyn v0 = a.normalized(), v1 = b.normalized();
    const double c = v1.dot(v0);
    if (c < -1.0 + 1e-12) internal::fail("FromTwoVectors of opposite vectors is not part of this shim");
    Dyn axis = v0.cross(v1);
    const double s = std::sqrt((1.0 + c) * 2.0), invs = 1.0 / s;
    return Quaternion(s * 0.5, axis(0) * invs, axis(1) * invs, axis(2) * invs);
  }
  const double* qdata() const { return c_; }
  double* qdata() { return c_; }
 private:
  double c_[4];
};
typedef Quaternion<double> Quaterniond;
template <> class Map<Quaterniond> : public QuaternionBase<Map<Quaterniond>> {
 public:
  explicit Map(double* p) : p_(p) {}
  template <class O> Map& operator=(const QuaternionBase<O>& o) { for (int i = 0; i < 4; i++) p_[i] = o.q()[i]; return *this; }
  Map& operator=(const Map& o) { for (int i = 0; i < 4; i++) p_[i] = o.p_[i]; return *this; }
  const double* qdata() const { return p_; }
  double* qdata() { return p_; }
 private:
  double* p_;
};
template <> class Map<const Quaterniond> : public QuaternionBase<Map<const Quaterniond>> {
 public:
  explicit Map(const double* p) : p_(const_cast<double*>(p)) {}
  const double* qdata() const { return p_; }
  double* qdata() { return p_; }
 private:
  double* p_;
};
template <class D> Quaterniond QuaternionBase<D>::normalized() const { const double n = norm(); return Quaterniond(w() / n, x() / n, y() / n, z() / n); }
template <class D> Quaterniond QuaternionBase<D>::conjugate() const { return Quaterniond(w(), -x(), -y(), -z()); }
template <class D> Quaterniond QuaternionBase<D>::inverse() const {
  const double n2 = squaredNorm();
  if (n2 > 0.0) return Quaterniond(w() / n2, -x() / n2, -y() / n2, -z() / n2);
  return Quaterniond(0, 0, 0, 0);
}
template <class D> template <class T> Quaterniond QuaternionBase<D>::cast() const { return Quaterniond(w(), x(), y(), z()); }
template <class D> Matrix3d QuaternionBase<D>::toRotationMatrix() const {
  Matrix3d res;
  const double tx = 2.0 * x(), ty = 2.0 * y(), tz = 2.0 * z();
  const double twx = tx * w(), twy = ty * w(), twz = tz * w();
  const double txx = tx * x(), txy = ty * x(), txz = tz * x();
  const double tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
  res(0, 0) = 1.0 - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
  res(1, 0) = txy + twz; res(1, 1) = 1.0 - (txx + tzz); res(1, 2) = tyz - twx;
  res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = 1.0 - (txx + tyy);
  return res;
}
template <class A, class B> Quaterniond operator*(const QuaternionBase<A>& a, const QuaternionBase<B>& b) {
  return Quaterniond(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                     a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                     a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                     a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
}
// Eigen's QuaternionBase::slerp
template <class D> template <class O> Quaterniond QuaternionBase<D>::slerp(double t, const QuaternionBase<O>& other) const {
  const double one = 1.0 - 2.220446049250313e-16;
  const double d = this->dot(other), absD = std::fabs(d);
  double scale0, scale1;
  if (absD >= one) { scale0 = 1.0 - t; scale1 = t; }
  else {
    const double theta = std::acos(absD), sinTheta = std::sin(theta);
    scale0 = std::sin((1.0 - t) * theta) / sinTheta; scale1 = std::sin(t * theta) / sinTheta;
  }
  if (d < 0.0) scale1 = -scale1;
  return Quaterniond(scale0 * w() + scale1 * other.w(), scale0 * x() + scale1 * other.x(), scale0 * y() + scale1 * other.y(), scale0 * z() + scale1 * other.z());
}
template <class D> template <class O> double QuaternionBase<D>::angularDistance(const QuaternionBase<O>& o) const {
  Quaterniond d = (*this) * o.conjugate();
  return 2.0 * std::atan2(d.vec().norm(), std::fabs(d.w()));
}

// AngleAxis (Rodrigues), enough for `AngleAxisd(angle, axis).toRotationMatrix()` / `Quaterniond(AngleAxisd)`
class AngleAxisd {
 public:
  AngleAxisd() : angle_(0), axis_(1, 0, 0) {}
  template <class D> AngleAxisd(double angle, const MatrixBase<D>& axis) : angle_(angle), axis_(axis) {}
  double angle() const { return angle_; }
  const Vector3d& axis() const { return axis_; }
  Matrix3d toRotationMatrix() const {
    Matrix3d res; const double s = std::sin(angle_), c = std::cos(angle_);
    const double cx = (1.0 - c) * axis_(0), cy = (1.0 - c) * axis_(1), cz = (1.0 - c) * axis_(2);
    double tmp = cx * axis_(1); res(0, 1) = tmp - s * axis_(2); res(1, 0) = tmp + s * axis_(2);
    tmp = cx * axis_(2); res(0, 2) = tmp + s * axis_(1); res(2, 0) = tmp - s * axis_(1);
    tmp = cy * axis_(2); res(1, 2) = tmp - s * axis_(0); res(2, 1) = tmp + s * axis_(0);
    res(0, 0) = cx * axis_(0) + c; res(1, 1) = cy * axis_(1) + c; res(2, 2) = cz * axis_(2) + c;
    return res;
  }
  Matrix3d matrix() const { return toRotationMatrix(); }
 private:
  double angle_; Vector3d axis_;
};

}  // namespace Eigen
