// TEST INFRASTRUCTURE — CPU oracle. Restatement of the part of Ceres Solver 1.14 the reference drives through
// ceres::Solve (VE/estimator/estimator.cpp:3364-3379): Problem bookkeeping, HuberLoss + Corrector, local
// parameterizations [I6;0], Jacobi scaling, DENSE_SCHUR elimination of the 1-dim landmark blocks, Eigen-LLT
// reduced solve, TRADITIONAL_DOGLEG trust region, and TrustRegionMinimizer's accept/terminate logic.
// Ceres is an un-vendored dependency (README.md:78 pins 1.14): this is restated from its published algorithm
// (trust_region_minimizer.cc, dogleg_strategy.cc, corrector.cc, schur_eliminator_impl.h) — "parity unpinned":
// no reference test stores solver outputs. It is cross-checked against scipy.optimize.least_squares minima in
// tests/test_oracle_solver.py.
#pragma once
#include <vector>
#include "gf2o_factors.h"

namespace gf2o {

struct ParameterBlock {
  double* data = nullptr;
  int size = 0;
  int local_size = 0;
  bool pose_manifold = false;  // PoseLocalParameterization / PoseSubsetParameterization (VE/factor/pose_local_parameterization.cpp)
  bool subset_mask[6] = {false, false, false, false, false, false};  // PoseSubsetParameterization::Plus zeroes these deltas
  bool constant = false;
  bool eliminate = false;  // e-block of the Schur ordering (free landmark)
  int col = -1;            // first column in the (f | e) tangent vector, assigned by Solve
};

struct ResidualBlock {
  const CostFunction* cost = nullptr;
  bool huber = false;
  std::vector<int> params;
};

struct Problem {
  std::vector<ParameterBlock> blocks;
  std::vector<ResidualBlock> residuals;
  int AddParameterBlock(double* data, int size, bool pose_manifold = false);
  void SetParameterBlockConstant(int id) { blocks[id].constant = true; }
  void AddResidualBlock(const CostFunction* f, bool huber, const std::vector<int>& params);
};

struct SolverOptions {
  int max_num_iterations = 8;
  double huber_delta = 1.0;
  double initial_trust_region_radius = 1e4;
  double max_trust_region_radius = 1e16;
  double min_trust_region_radius = 1e-32;
  double min_relative_decrease = 1e-3;
  double function_tolerance = 1e-6;
  double gradient_tolerance = 1e-10;
  double parameter_tolerance = 1e-8;
  bool jacobi_scaling = true;
};

struct IterationRecord { double cost, model_cost_change, relative_decrease, radius, step_norm; bool successful; };

struct SolverSummary {
  double initial_cost = 0, final_cost = 0;
  int iterations = 0, successful_steps = 0, termination = 0;
  std::vector<IterationRecord> trace;
};

// One linearisation in the unscaled tangent space, for parity tests of the GPU sweep:
// reduced system S (D x D, row-major), g (D), cost; columns follow the f-block order of the problem.
struct Linearization {
  int D = 0, E = 0;
  std::vector<double> S, g;      // Schur complement WITHOUT regularisation, reduced gradient
  std::vector<double> H_ff_diag; // diag of F^T F (unreduced)
  std::vector<double> ete, etr;  // per e-block
  double cost = 0;
};

void Solve(const SolverOptions& opt, Problem* problem, SolverSummary* summary);
void Linearize(const SolverOptions& opt, Problem* problem, Linearization* out);

}  // namespace gf2o
