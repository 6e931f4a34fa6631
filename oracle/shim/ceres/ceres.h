// TEST INFRASTRUCTURE — the plug-in interfaces of Ceres 1.14 the reference's factor classes derive from (cost_function.h,
// sized_cost_function.h, loss_function.h, local_parameterization.h), declarations only + HuberLoss / CauchyLoss (loss_function.cc).
// No solver: the reference's factor code is EVALUATED through these interfaces by oracle/ref_bridge.cpp.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include <limits>
#include <algorithm>
namespace ceres {
typedef int int32;
class CostFunction {
 public:
  CostFunction() : num_residuals_(0) {}
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int32>& parameter_block_sizes() const { return parameter_block_sizes_; }
  int num_residuals() const { return num_residuals_; }
 protected:
  std::vector<int32>* mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
  void set_num_residuals(int n) { num_residuals_ = n; }
 private:
  std::vector<int32> parameter_block_sizes_;
  int num_residuals_;
};
template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() { set_num_residuals(kNumResiduals); *mutable_parameter_block_sizes() = std::vector<int32>{Ns...}; }
  virtual ~SizedCostFunction() {}
};
// declared so that the reference's autodiff factories parse; never evaluated by the bridge
template <class Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public SizedCostFunction<kNumResiduals, Ns...> {
 public:
  explicit AutoDiffCostFunction(Functor* f) : functor_(f) {}
  virtual ~AutoDiffCostFunction() { delete functor_; }
  virtual bool Evaluate(double const* const*, double*, double**) const { return false; }
 private:
  Functor* functor_;
};
class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};
class HuberLoss : public LossFunction {   // loss_function.cc: rho(s) = s for s <= a^2, 2 a sqrt(s) - a^2 otherwise
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  virtual void Evaluate(double s, double rho[3]) const {
    if (s > b_) { const double r = std::sqrt(s); rho[0] = 2.0 * a_ * r - b_; rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r); rho[2] = -rho[1] / (2.0 * s); }
    else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  }
 private:
  const double a_, b_;
};
class CauchyLoss : public LossFunction {
 public:
  explicit CauchyLoss(double a) : b_(a * a), c_(1.0 / b_) {}
  virtual void Evaluate(double s, double rho[3]) const {
    const double sum = 1.0 + s * c_, inv = 1.0 / sum;
    rho[0] = b_ * std::log(sum); rho[1] = std::max(std::numeric_limits<double>::min(), inv); rho[2] = -c_ * (inv * inv);
  }
 private:
  const double b_, c_;
};
class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};
}  // namespace ceres
