// TEST INFRASTRUCTURE — CPU oracle. Restatement of the marginalization step that follows the solve in
// Estimator::optimization(): ResidualBlockInfo::Evaluate (VE/factor/marginalization_factor.cpp:12-78),
// MarginalizationInfo::addResidualBlockInfo / preMarginalize / marginalize / getParameterBlocks (:98-330) and the two
// call sites (VE/estimator/estimator.cpp:3394-3560 MARGIN_OLD with the addr_shift remap :3561-3595,
// :3597-3690 MARGIN_SECOND_NEW). "parity unpinned": the reference stores no expected prior anywhere; tests pin this
// file against an independent numpy restatement (tests/test_oracle_marg.py).
//
// Deviation that cannot be avoided: the reference keys its block tables by parameter ADDRESS in std::unordered_map,
// so the order of blocks inside the marginalized / kept sets is whatever the hash iteration gives on that run. Here the
// order is first insertion. The prior it defines, || r0 + J0 dx ||^2 as a function of the kept blocks, does not depend on
// that order (columns of J0 move with keep_block_idx; eigenvalues are unchanged), so comparisons are made on
// J0^T J0 and J0^T r0 scattered by block, never on J0 element by element.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include "gf2o_ceres.h"

namespace gf2o_marg {
using namespace gf2o;

// ceres::HuberLoss::Evaluate
inline void huberRho(double a, double s, double rho[3]) {
  const double b = a * a;
  if (s > b) { const double r = std::sqrt(s); rho[0] = 2 * a * r - b; rho[1] = std::max(1e-300, a / r); rho[2] = -rho[1] / (2 * s); }
  else { rho[0] = s; rho[1] = 1; rho[2] = 0; }
}

struct ResidualBlockInfo {  // marginalization_factor.h:24-43
  const CostFunction* cost_function;
  bool loss;
  std::vector<double*> parameter_blocks;
  std::vector<int> drop_set;
  std::vector<std::vector<double>> jacobians;  // row-major num_residuals x size
  std::vector<double> residuals;
  void Evaluate(double huber_delta) {  // marginalization_factor.cpp:12-78
    const int R = cost_function->num_residuals;
    residuals.assign(R, 0.0);
    jacobians.resize(cost_function->block_sizes.size());
    std::vector<double*> raw;
    for (size_t i = 0; i < jacobians.size(); i++) { jacobians[i].assign((size_t)R * cost_function->block_sizes[i], 0.0); raw.push_back(jacobians[i].data()); }
    std::vector<const double*> pp(parameter_blocks.begin(), parameter_blocks.end());
    cost_function->Evaluate(pp.data(), residuals.data(), raw.data());
    if (loss) {
      double sq_norm = 0; for (double r : residuals) sq_norm += r * r;
      double rho[3]; huberRho(huber_delta, sq_norm, rho);
      const double sqrt_rho1 = std::sqrt(rho[1]);
      double residual_scaling, alpha_sq_norm;
      if (sq_norm == 0.0 || rho[2] <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
      else {
        const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
        const double alpha = 1.0 - std::sqrt(D);
        residual_scaling = sqrt_rho1 / (1 - alpha);
        alpha_sq_norm = alpha / sq_norm;
      }
      for (size_t i = 0; i < jacobians.size(); i++) {
        const int S = cost_function->block_sizes[i];
        std::vector<double> rtJ(S, 0.0);
        for (int r = 0; r < R; r++) for (int c = 0; c < S; c++) rtJ[c] += residuals[r] * jacobians[i][r * S + c];
        for (int r = 0; r < R; r++) for (int c = 0; c < S; c++) jacobians[i][r * S + c] = sqrt_rho1 * (jacobians[i][r * S + c] - alpha_sq_norm * residuals[r] * rtJ[c]);
      }
      for (double& r : residuals) r *= residual_scaling;
    }
  }
};

struct Info {  // MarginalizationInfo, marginalization_factor.h:56-88
  std::vector<std::unique_ptr<ResidualBlockInfo>> factors;
  int m = 0, n = 0;
  // insertion-ordered stand-ins for the unordered_map<long, ...> tables
  std::vector<double*> order;                 // every block, first appearance
  std::map<double*, int> size;                // parameter_block_size
  std::map<double*, int> idx;                 // parameter_block_idx
  std::vector<double*> drop_order;            // keys of parameter_block_idx before marginalize(), first appearance
  std::map<double*, std::vector<double>> data;  // parameter_block_data
  std::vector<double> linearized_jacobians, linearized_residuals, A_full, b_full;
  bool valid = true;
  static int localSize(int s) { return s == 7 ? 6 : s; }

  void addResidualBlockInfo(ResidualBlockInfo* r) {  // :98-117
    factors.emplace_back(r);
    for (size_t i = 0; i < r->parameter_blocks.size(); i++) {
      double* a = r->parameter_blocks[i];
      if (!size.count(a)) order.push_back(a);
      size[a] = r->cost_function->block_sizes[i];
    }
    for (int d : r->drop_set) { double* a = r->parameter_blocks[d]; if (!idx.count(a)) drop_order.push_back(a); idx[a] = 0; }
  }
  void preMarginalize(double huber_delta) {  // :119-138
    for (auto& f : factors) {
      f->Evaluate(huber_delta);
      for (size_t i = 0; i < f->parameter_blocks.size(); i++) {
        double* a = f->parameter_blocks[i];
        if (!data.count(a)) data[a] = std::vector<double>(a, a + f->cost_function->block_sizes[i]);
      }
    }
  }
  void marginalize(double eps) {  // :183-308
    int pos = 0;
    for (double* a : drop_order) { idx[a] = pos; pos += localSize(size[a]); }
    m = pos;
    for (double* a : order) if (!idx.count(a)) { idx[a] = pos; pos += localSize(size[a]); }
    n = pos - m;
    if (m == 0) { valid = false; return; }
    std::vector<double> A((size_t)pos * pos, 0.0), b(pos, 0.0);
    for (auto& f : factors) {  // ThreadsConstructA, :150-181 (the thread split only changes the summation order)
      const int R = f->cost_function->num_residuals;
      for (size_t i = 0; i < f->parameter_blocks.size(); i++) {
        const int idx_i = idx[f->parameter_blocks[i]], gi = f->cost_function->block_sizes[i], si = localSize(gi);
        for (size_t j = i; j < f->parameter_blocks.size(); j++) {
          const int idx_j = idx[f->parameter_blocks[j]], gj = f->cost_function->block_sizes[j], sj = localSize(gj);
          for (int a = 0; a < si; a++) for (int c = 0; c < sj; c++) {
            double s = 0; for (int r = 0; r < R; r++) s += f->jacobians[i][r * gi + a] * f->jacobians[j][r * gj + c];
            A[(size_t)(idx_i + a) * pos + idx_j + c] += s;
          }
          if (i != j) for (int a = 0; a < si; a++) for (int c = 0; c < sj; c++) A[(size_t)(idx_j + c) * pos + idx_i + a] = A[(size_t)(idx_i + a) * pos + idx_j + c];
        }
        for (int a = 0; a < si; a++) { double s = 0; for (int r = 0; r < R; r++) s += f->jacobians[i][r * gi + a] * f->residuals[r]; b[idx_i + a] += s; }
      }
    }
    A_full = A; b_full = b;
    // Amm = 0.5 (Amm + Amm^T); pseudo-inverse by eigenvalue truncation, :277-283
    std::vector<double> Amm((size_t)m * m), ev(m), V((size_t)m * m), Ainv((size_t)m * m, 0.0);
    for (int r = 0; r < m; r++) for (int c = 0; c < m; c++) Amm[(size_t)r * m + c] = 0.5 * (A[(size_t)r * pos + c] + A[(size_t)c * pos + r]);
    symEigen(m, Amm.data(), ev.data(), V.data());
    for (int k = 0; k < m; k++) {
      if (!(ev[k] > eps)) continue;
      const double inv = 1.0 / ev[k];
      for (int r = 0; r < m; r++) { const double vr = V[(size_t)r * m + k] * inv; for (int c = 0; c < m; c++) Ainv[(size_t)r * m + c] += vr * V[(size_t)c * m + k]; }
    }
    // A = Arr - Arm Amm_inv Amr ; b = brr - Arm Amm_inv bmm, :285-291
    std::vector<double> T((size_t)n * m, 0.0);  // Arm * Amm_inv
    for (int r = 0; r < n; r++) for (int k = 0; k < m; k++) { const double a = A[(size_t)(m + r) * pos + k]; if (a != 0.0) for (int c = 0; c < m; c++) T[(size_t)r * m + c] += a * Ainv[(size_t)k * m + c]; }
    std::vector<double> Ar((size_t)n * n), br(n);
    for (int r = 0; r < n; r++) {
      for (int c = 0; c < n; c++) { double s = A[(size_t)(m + r) * pos + m + c]; for (int k = 0; k < m; k++) s -= T[(size_t)r * m + k] * A[(size_t)k * pos + m + c]; Ar[(size_t)r * n + c] = s; }
      double s = b[m + r]; for (int k = 0; k < m; k++) s -= T[(size_t)r * m + k] * b[k]; br[r] = s;
    }
    // :293-303
    std::vector<double> S(n), V2((size_t)n * n);
    symEigen(n, Ar.data(), S.data(), V2.data());
    linearized_jacobians.assign((size_t)n * n, 0.0); linearized_residuals.assign(n, 0.0);
    for (int k = 0; k < n; k++) {
      const double s = S[k] > eps ? S[k] : 0.0, si = S[k] > eps ? 1.0 / S[k] : 0.0;
      const double ss = std::sqrt(s), sis = std::sqrt(si);
      double vb = 0;
      for (int c = 0; c < n; c++) { linearized_jacobians[(size_t)k * n + c] = ss * V2[(size_t)c * n + k]; vb += V2[(size_t)c * n + k] * br[c]; }
      linearized_residuals[k] = sis * vb;
    }
  }
};

}  // namespace gf2o_marg
