// TEST INFRASTRUCTURE — CPU oracle. See gf2o_ceres.h for what this restates and its pinning status.
#include "gf2o_ceres.h"
#include <limits>

namespace gf2o {

int Problem::AddParameterBlock(double* data, int size, bool pose_manifold) {
  ParameterBlock b; b.data = data; b.size = size; b.pose_manifold = pose_manifold;
  b.local_size = pose_manifold ? 6 : size;
  blocks.push_back(b);
  return (int)blocks.size() - 1;
}
void Problem::AddResidualBlock(const CostFunction* f, bool huber, const std::vector<int>& params) {
  ResidualBlock r; r.cost = f; r.huber = huber; r.params = params; residuals.push_back(r);
}

namespace {

// ceres::HuberLoss::Evaluate (loss_function.cc)
inline void huberRho(double a, double s, double rho[3]) {
  const double b = a * a;
  if (s > b) {
    const double r = std::sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

struct Program {
  Problem* p;
  SolverOptions opt;
  int D = 0, E = 0, num_rows = 0, num_global = 0;
  std::vector<int> goff;                 // per block: offset in the state vector (non-const only), else -1
  std::vector<int> row_off;              // per residual block
  std::vector<std::vector<int>> jac_off; // per residual block per slot: offset in jac (or -1 if const)
  std::vector<int> e_of_rb;              // per residual block: slot index of its e-block param, or -1
  std::vector<std::vector<int>> rbs_of_e; // per e-block index: residual blocks touching it
  std::vector<int> rbs_no_e;
  std::vector<double> jac;               // local Jacobians, row-major r x local_size per (rb, slot)
  size_t jac_size = 0;

  void build() {
    goff.assign(p->blocks.size(), -1);
    int col_f = 0, ne = 0;
    for (auto& b : p->blocks) if (!b.constant && !b.eliminate) { b.col = col_f; col_f += b.local_size; }
    D = col_f;
    for (auto& b : p->blocks) if (!b.constant && b.eliminate) { b.col = D + ne; ne++; }
    E = ne;
    int g = 0;
    for (size_t i = 0; i < p->blocks.size(); i++) if (!p->blocks[i].constant) { goff[i] = g; g += p->blocks[i].size; }
    num_global = g;
    rbs_of_e.assign(E, {});
    row_off.resize(p->residuals.size()); jac_off.resize(p->residuals.size()); e_of_rb.assign(p->residuals.size(), -1);
    int rows = 0; size_t jo = 0;
    for (size_t k = 0; k < p->residuals.size(); k++) {
      const ResidualBlock& rb = p->residuals[k];
      row_off[k] = rows;
      int r = rb.cost->num_residuals;
      jac_off[k].assign(rb.params.size(), -1);
      for (size_t j = 0; j < rb.params.size(); j++) {
        const ParameterBlock& b = p->blocks[rb.params[j]];
        if (b.constant) continue;
        jac_off[k][j] = (int)jo; jo += (size_t)r * b.local_size;
        if (b.eliminate) { e_of_rb[k] = (int)j; rbs_of_e[b.col - D].push_back((int)k); }
      }
      if (e_of_rb[k] < 0) rbs_no_e.push_back((int)k);
      rows += r;
    }
    num_rows = rows; jac_size = jo; jac.assign(jo, 0.0);
  }

  void gather(std::vector<double>& x) const {
    x.assign(num_global, 0.0);
    for (size_t i = 0; i < p->blocks.size(); i++) if (goff[i] >= 0) std::memcpy(&x[goff[i]], p->blocks[i].data, sizeof(double) * p->blocks[i].size);
  }
  void scatter(const std::vector<double>& x) const {
    for (size_t i = 0; i < p->blocks.size(); i++) if (goff[i] >= 0) std::memcpy(p->blocks[i].data, &x[goff[i]], sizeof(double) * p->blocks[i].size);
  }
  // Plus over all free blocks. Pose blocks: VE/factor/pose_local_parameterization.cpp:12-28 /
  // pose_subset_parameterization.cpp:28-58 (masked deltas zeroed).
  void plus(const std::vector<double>& x, const std::vector<double>& delta, std::vector<double>& out) const {
    out = x;
    for (size_t i = 0; i < p->blocks.size(); i++) {
      const ParameterBlock& b = p->blocks[i];
      if (goff[i] < 0) continue;
      const double* xx = &x[goff[i]]; const double* d = &delta[b.col]; double* o = &out[goff[i]];
      if (b.pose_manifold) {
        double dl[6]; for (int k = 0; k < 6; k++) dl[k] = b.subset_mask[k] ? 0.0 : d[k];
        for (int k = 0; k < 3; k++) o[k] = xx[k] + dl[k];
        Quat q(xx[6], xx[3], xx[4], xx[5]);
        Quat dq = deltaQ(v3(dl[3], dl[4], dl[5]));
        Quat r = (q * dq).normalized();
        o[3] = r.x; o[4] = r.y; o[5] = r.z; o[6] = r.w;
      } else {
        for (int k = 0; k < b.size; k++) o[k] = xx[k] + d[k];
      }
    }
  }

  // Evaluator::Evaluate: cost = sum 0.5 rho(|r|^2); residuals/Jacobians loss-corrected (corrector.cc) and
  // projected to the tangent space (J_global * [I6;0] = first 6 columns for pose blocks).
  bool evaluate(const std::vector<double>& x, double* cost, double* residuals, bool want_jac) {
    double total = 0.0;
    std::vector<const double*> params;
    std::vector<double> gj;            // global jacobian scratch
    std::vector<double*> gjp;
    double rloc[GF2_MAX_PRIOR_DIM + 16];
    for (size_t k = 0; k < p->residuals.size(); k++) {
      const ResidualBlock& rb = p->residuals[k];
      const int r = rb.cost->num_residuals;
      params.resize(rb.params.size()); gjp.assign(rb.params.size(), nullptr);
      size_t need = 0;
      for (size_t j = 0; j < rb.params.size(); j++) {
        const int id = rb.params[j];
        params[j] = goff[id] >= 0 ? &x[goff[id]] : p->blocks[id].data;
        if (want_jac && jac_off[k][j] >= 0) need += (size_t)r * p->blocks[id].size;
      }
      if (want_jac) {
        gj.assign(need, 0.0); size_t o = 0;
        for (size_t j = 0; j < rb.params.size(); j++) if (jac_off[k][j] >= 0) { gjp[j] = &gj[o]; o += (size_t)r * p->blocks[rb.params[j]].size; }
      }
      double* rr = residuals ? residuals + row_off[k] : rloc;
      if (!rb.cost->Evaluate(params.data(), rr, want_jac ? gjp.data() : nullptr)) return false;
      double sq = 0; for (int i = 0; i < r; i++) sq += rr[i] * rr[i];
      double sqrt_rho1 = 1.0, residual_scaling = 1.0, alpha_sq_norm = 0.0;
      if (rb.huber) {
        double rho[3]; huberRho(opt.huber_delta, sq, rho);
        total += 0.5 * rho[0];
        sqrt_rho1 = std::sqrt(rho[1]);
        if (sq == 0.0 || rho[2] <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
        else {
          const double Dd = 1.0 + 2.0 * sq * rho[2] / rho[1];
          const double alpha = 1.0 - std::sqrt(Dd);
          residual_scaling = sqrt_rho1 / (1 - alpha); alpha_sq_norm = alpha / sq;
        }
      } else total += 0.5 * sq;
      if (want_jac) {
        for (size_t j = 0; j < rb.params.size(); j++) {
          if (jac_off[k][j] < 0) continue;
          const ParameterBlock& b = p->blocks[rb.params[j]];
          double* dst = &jac[jac_off[k][j]];
          const double* src = gjp[j];
          for (int c = 0; c < b.local_size; c++) {
            double rtj = 0; if (alpha_sq_norm != 0.0) for (int i = 0; i < r; i++) rtj += rr[i] * src[i * b.size + c];
            for (int i = 0; i < r; i++) dst[i * b.local_size + c] = sqrt_rho1 * (src[i * b.size + c] - alpha_sq_norm * rr[i] * rtj);
          }
        }
      }
      if (rb.huber && residuals) for (int i = 0; i < r; i++) rr[i] *= residual_scaling;
    }
    *cost = total;
    return true;
  }

  int cols(size_t k, size_t j) const { return p->blocks[p->residuals[k].params[j]].col; }
  int lsz(size_t k, size_t j) const { return p->blocks[p->residuals[k].params[j]].local_size; }

  void squaredColumnNorm(std::vector<double>& out) const {
    out.assign(D + E, 0.0);
    for (size_t k = 0; k < p->residuals.size(); k++) { int r = p->residuals[k].cost->num_residuals;
      for (size_t j = 0; j < jac_off[k].size(); j++) { if (jac_off[k][j] < 0) continue; int c0 = cols(k, j), ls = lsz(k, j); const double* J = &jac[jac_off[k][j]];
        for (int i = 0; i < r; i++) for (int c = 0; c < ls; c++) out[c0 + c] += J[i * ls + c] * J[i * ls + c]; } }
  }
  void scaleColumns(const std::vector<double>& s) {
    for (size_t k = 0; k < p->residuals.size(); k++) { int r = p->residuals[k].cost->num_residuals;
      for (size_t j = 0; j < jac_off[k].size(); j++) { if (jac_off[k][j] < 0) continue; int c0 = cols(k, j), ls = lsz(k, j); double* J = &jac[jac_off[k][j]];
        for (int i = 0; i < r; i++) for (int c = 0; c < ls; c++) J[i * ls + c] *= s[c0 + c]; } }
  }
  void leftMultiply(const double* res, std::vector<double>& y) const {  // y = J^T res
    y.assign(D + E, 0.0);
    for (size_t k = 0; k < p->residuals.size(); k++) { int r = p->residuals[k].cost->num_residuals; const double* rr = res + row_off[k];
      for (size_t j = 0; j < jac_off[k].size(); j++) { if (jac_off[k][j] < 0) continue; int c0 = cols(k, j), ls = lsz(k, j); const double* J = &jac[jac_off[k][j]];
        for (int i = 0; i < r; i++) for (int c = 0; c < ls; c++) y[c0 + c] += J[i * ls + c] * rr[i]; } }
  }
  void rightMultiply(const double* x, std::vector<double>& y) const {  // y = J x
    y.assign(num_rows, 0.0);
    for (size_t k = 0; k < p->residuals.size(); k++) { int r = p->residuals[k].cost->num_residuals; double* yy = &y[row_off[k]];
      for (size_t j = 0; j < jac_off[k].size(); j++) { if (jac_off[k][j] < 0) continue; int c0 = cols(k, j), ls = lsz(k, j); const double* J = &jac[jac_off[k][j]];
        for (int i = 0; i < r; i++) { double s = 0; for (int c = 0; c < ls; c++) s += J[i * ls + c] * x[c0 + c]; yy[i] += s; } } }
  }

  // Accumulate F^T F and F^T r of one residual block into lhs/rhs (f-block params only)
  void addFtF(size_t k, const double* res, std::vector<double>& lhs, std::vector<double>& rhs) const {
    int r = p->residuals[k].cost->num_residuals; const double* rr = res + row_off[k];
    for (size_t a = 0; a < jac_off[k].size(); a++) {
      if (jac_off[k][a] < 0 || (int)a == e_of_rb[k]) continue;
      int ca = cols(k, a), la = lsz(k, a); const double* Ja = &jac[jac_off[k][a]];
      for (int i = 0; i < r; i++) for (int c = 0; c < la; c++) rhs[ca + c] += Ja[i * la + c] * rr[i];
      for (size_t b = 0; b < jac_off[k].size(); b++) {
        if (jac_off[k][b] < 0 || (int)b == e_of_rb[k]) continue;
        int cb = cols(k, b), lb = lsz(k, b); const double* Jb = &jac[jac_off[k][b]];
        for (int x = 0; x < la; x++) for (int y = 0; y < lb; y++) { double s = 0; for (int i = 0; i < r; i++) s += Ja[i * la + x] * Jb[i * lb + y]; lhs[(size_t)(ca + x) * D + cb + y] += s; }
      }
    }
  }

  // DENSE_SCHUR: (J^T J + diag(Dlm)^2) y = J^T res, eliminating the e-blocks (schur_eliminator_impl.h), reduced
  // system solved by Cholesky (DenseSchurComplementSolver::SolveReducedLinearSystem, Eigen LLT). If `keep` is set
  // the unregularised pieces are exported for Linearize().
  bool schurSolve(const double* res, const double* Dlm, std::vector<double>& y, Linearization* keep = nullptr) const {
    std::vector<double> lhs((size_t)D * D, 0.0), rhs(D, 0.0), w(D, 0.0), ete(E), etr(E);
    std::vector<int> touched;
    std::vector<char> mark(D, 0);
    for (int k : rbs_no_e) addFtF(k, res, lhs, rhs);
    if (keep) { keep->H_ff_diag.assign(D, 0.0); }
    for (int l = 0; l < E; l++) {
      double dl = Dlm ? Dlm[D + l] : 0.0;
      double ee = dl * dl, er = 0.0;
      touched.clear();
      for (int k : rbs_of_e[l]) {
        int r = p->residuals[k].cost->num_residuals; const double* rr = res + row_off[k];
        const double* Je = &jac[jac_off[k][e_of_rb[k]]];
        for (int i = 0; i < r; i++) { ee += Je[i] * Je[i]; er += Je[i] * rr[i]; }
        for (size_t a = 0; a < jac_off[k].size(); a++) {
          if (jac_off[k][a] < 0 || (int)a == e_of_rb[k]) continue;
          int ca = cols(k, a), la = lsz(k, a); const double* Ja = &jac[jac_off[k][a]];
          for (int c = 0; c < la; c++) { double s = 0; for (int i = 0; i < r; i++) s += Ja[i * la + c] * Je[i];
            if (!mark[ca + c]) { mark[ca + c] = 1; touched.push_back(ca + c); }
            w[ca + c] += s; }
        }
        addFtF(k, res, lhs, rhs);
      }
      ete[l] = ee; etr[l] = er;
      if (ee > 0.0) {
        double inv = 1.0 / ee;
        for (int i : touched) { double wi = w[i] * inv; rhs[i] -= wi * er; for (int j : touched) lhs[(size_t)i * D + j] -= wi * w[j]; }
      }
      for (int i : touched) { w[i] = 0.0; mark[i] = 0; }
    }
    if (keep) {
      keep->D = D; keep->E = E; keep->S = lhs; keep->g = rhs; keep->ete = ete; keep->etr = etr;
      return true;
    }
    if (Dlm) for (int i = 0; i < D; i++) lhs[(size_t)i * D + i] += Dlm[i] * Dlm[i];
    if (!choleskyLower(D, lhs.data())) return false;
    choleskySolve(D, lhs.data(), rhs.data());
    y.assign(D + E, 0.0);
    for (int i = 0; i < D; i++) y[i] = rhs[i];
    // back substitution: y_e = (e^T r - e^T F y_f) / (e^T e + D_e^2)
    for (int l = 0; l < E; l++) {
      double s = etr[l];
      for (int k : rbs_of_e[l]) {
        int r = p->residuals[k].cost->num_residuals; const double* Je = &jac[jac_off[k][e_of_rb[k]]];
        for (size_t a = 0; a < jac_off[k].size(); a++) {
          if (jac_off[k][a] < 0 || (int)a == e_of_rb[k]) continue;
          int ca = cols(k, a), la = lsz(k, a); const double* Ja = &jac[jac_off[k][a]];
          for (int i = 0; i < r; i++) { double t = 0; for (int c = 0; c < la; c++) t += Ja[i * la + c] * y[ca + c]; s -= Je[i] * t; }
        }
      }
      y[D + l] = ete[l] > 0.0 ? s / ete[l] : 0.0;
    }
    for (double v : y) if (!std::isfinite(v)) return false;
    return true;
  }
};

}  // namespace

void Linearize(const SolverOptions& opt, Problem* problem, Linearization* out) {
  Program prog; prog.p = problem; prog.opt = opt; prog.build();
  std::vector<double> x; prog.gather(x);
  std::vector<double> res(prog.num_rows);
  double cost = 0;
  prog.evaluate(x, &cost, res.data(), true);
  std::vector<double> y;
  prog.schurSolve(res.data(), nullptr, y, out);
  out->cost = cost;
  std::vector<double> cn; prog.squaredColumnNorm(cn);
  for (int i = 0; i < prog.D; i++) out->H_ff_diag[i] = cn[i];
}

// TrustRegionMinimizer::Minimize (trust_region_minimizer.cc, Ceres 1.14) with DoglegStrategy (dogleg_strategy.cc,
// TRADITIONAL_DOGLEG) and monotonic steps.
void Solve(const SolverOptions& opt, Problem* problem, SolverSummary* summary) {
  Program prog; prog.p = problem; prog.opt = opt; prog.build();
  const int n = prog.D + prog.E;
  *summary = SolverSummary();
  if (n == 0) return;
  std::vector<double> x, cand, best;
  prog.gather(x); best = x;
  double x_norm = 0; for (double v : x) x_norm += v * v; x_norm = std::sqrt(x_norm);
  std::vector<double> residuals(prog.num_rows), gradient, scale(n, 1.0), tmp;
  double x_cost = 0, minimum_cost = std::numeric_limits<double>::max();
  double gradient_max_norm = 0;
  int iteration = 0;

  auto evalGradJac = [&]() -> bool {  // EvaluateGradientAndJacobian
    if (!prog.evaluate(x, &x_cost, residuals.data(), true)) return false;
    prog.leftMultiply(residuals.data(), gradient);  // unscaled gradient (evaluator)
    if (opt.jacobi_scaling) {
      if (iteration == 0) { prog.squaredColumnNorm(scale); for (int i = 0; i < n; i++) scale[i] = 1.0 / (1.0 + std::sqrt(scale[i])); }
      prog.scaleColumns(scale);
    }
    std::vector<double> neg(n), proj; for (int i = 0; i < n; i++) neg[i] = -gradient[i];
    prog.plus(x, neg, proj);
    gradient_max_norm = 0; for (size_t i = 0; i < x.size(); i++) gradient_max_norm = std::max(gradient_max_norm, std::fabs(x[i] - proj[i]));
    return true;
  };

  // Dogleg strategy state
  double radius = opt.initial_trust_region_radius;
  const double min_mu = 1e-8, max_mu = 1.0, mu_increase = 10.0, min_diag = 1e-6, max_diag = 1e32;
  double mu = min_mu, alpha = 0, dogleg_step_norm = 0;
  bool reuse = false;
  std::vector<double> diagonal, dl_gradient, gauss_newton, lm_diag(n), step(n), delta(n);

  auto doglegInterp = [&]() {  // ComputeTraditionalDoglegStep
    double gradient_norm = 0, gn_norm = 0;
    for (int i = 0; i < n; i++) { gradient_norm += dl_gradient[i] * dl_gradient[i]; gn_norm += gauss_newton[i] * gauss_newton[i]; }
    gradient_norm = std::sqrt(gradient_norm); gn_norm = std::sqrt(gn_norm);
    if (gn_norm <= radius) {
      for (int i = 0; i < n; i++) step[i] = gauss_newton[i] / diagonal[i];
      dogleg_step_norm = gn_norm; return;
    }
    if (gradient_norm * alpha >= radius) {
      for (int i = 0; i < n; i++) step[i] = -(radius / gradient_norm) * dl_gradient[i] / diagonal[i];
      dogleg_step_norm = radius; return;
    }
    double gdot = 0; for (int i = 0; i < n; i++) gdot += dl_gradient[i] * gauss_newton[i];
    const double b_dot_a = -alpha * gdot;
    const double a_squared_norm = std::pow(alpha * gradient_norm, 2.0);
    const double b_minus_a_squared_norm = a_squared_norm - 2 * b_dot_a + std::pow(gn_norm, 2);
    const double c = b_dot_a - a_squared_norm;
    const double d = std::sqrt(c * c + b_minus_a_squared_norm * (std::pow(radius, 2.0) - a_squared_norm));
    double beta = (c <= 0) ? (d - c) / b_minus_a_squared_norm : (radius * radius - a_squared_norm) / (d + c);
    double sn = 0;
    for (int i = 0; i < n; i++) { double v = (-alpha * (1.0 - beta)) * dl_gradient[i] + beta * gauss_newton[i]; sn += v * v; step[i] = v / diagonal[i]; }
    dogleg_step_norm = std::sqrt(sn);
  };
  // returns 0 ok, 1 linear-solver failure
  auto computeStep = [&]() -> int {  // DoglegStrategy::ComputeStep
    if (reuse) { doglegInterp(); return 0; }
    reuse = true;
    prog.squaredColumnNorm(diagonal);
    for (int i = 0; i < n; i++) diagonal[i] = std::sqrt(std::min(std::max(diagonal[i], min_diag), max_diag));
    prog.leftMultiply(residuals.data(), dl_gradient);
    for (int i = 0; i < n; i++) dl_gradient[i] /= diagonal[i];
    // Cauchy point
    std::vector<double> sg(n), Jg; for (int i = 0; i < n; i++) sg[i] = dl_gradient[i] / diagonal[i];
    prog.rightMultiply(sg.data(), Jg);
    double g2 = 0, jg2 = 0; for (int i = 0; i < n; i++) g2 += dl_gradient[i] * dl_gradient[i]; for (double v : Jg) jg2 += v * v;
    alpha = g2 / jg2;
    // Gauss-Newton step
    bool ok = false;
    while (mu < max_mu) {
      for (int i = 0; i < n; i++) lm_diag[i] = diagonal[i] * std::sqrt(mu);
      if (prog.schurSolve(residuals.data(), lm_diag.data(), gauss_newton)) { ok = true; break; }
      mu *= mu_increase;
    }
    if (!ok) return 1;
    for (int i = 0; i < n; i++) gauss_newton[i] *= -diagonal[i];
    doglegInterp();
    return 0;
  };

  // IterationZero
  if (!evalGradJac()) { summary->termination = GF2_TERM_FAILURE; return; }
  summary->initial_cost = x_cost;
  bool step_is_successful = true;
  int num_consecutive_invalid = 0;
  double candidate_cost = x_cost, model_cost_change = 0;
  std::vector<double> model_residuals;

  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (step_is_successful) {
      summary->successful_steps++;
      if (x_cost < minimum_cost) { minimum_cost = x_cost; best = x; }
    }
    summary->iterations = iteration;
    if (iteration >= opt.max_num_iterations) { summary->termination = GF2_TERM_NO_CONVERGENCE; break; }
    if (step_is_successful && gradient_max_norm <= opt.gradient_tolerance) { summary->termination = GF2_TERM_GRADIENT_TOL; break; }
    if (radius <= opt.min_trust_region_radius) { summary->termination = GF2_TERM_MIN_RADIUS; break; }

    iteration++;
    step_is_successful = false;
    // ComputeTrustRegionStep
    int st = computeStep();
    bool step_valid = false;
    if (st == 0) {
      prog.rightMultiply(step.data(), model_residuals);
      double mc = 0; for (int i = 0; i < prog.num_rows; i++) mc += model_residuals[i] * (residuals[i] + model_residuals[i] / 2.0);
      model_cost_change = -mc;
      step_valid = model_cost_change > 0.0;
    }
    if (!step_valid) {  // HandleInvalidStep
      if (++num_consecutive_invalid >= 5) { summary->termination = GF2_TERM_FAILURE; summary->iterations = iteration; break; }
      mu *= mu_increase; reuse = false;  // StepIsInvalid
      summary->trace.push_back({x_cost, model_cost_change, 0, radius, 0, false});
      continue;
    }
    num_consecutive_invalid = 0;
    for (int i = 0; i < n; i++) delta[i] = step[i] * scale[i];
    // ComputeCandidatePointAndEvaluateCost
    prog.plus(x, delta, cand);
    if (!prog.evaluate(cand, &candidate_cost, nullptr, false)) candidate_cost = std::numeric_limits<double>::max();
    // ParameterToleranceReached
    double step_norm = 0; for (size_t i = 0; i < x.size(); i++) step_norm += (x[i] - cand[i]) * (x[i] - cand[i]); step_norm = std::sqrt(step_norm);
    if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { summary->termination = GF2_TERM_PARAMETER_TOL; summary->iterations = iteration; break; }
    // FunctionToleranceReached
    double cost_change = x_cost - candidate_cost;
    if (std::fabs(cost_change) <= opt.function_tolerance * x_cost) { summary->termination = GF2_TERM_FUNCTION_TOL; summary->iterations = iteration; break; }
    // IsStepSuccessful
    double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > opt.min_relative_decrease) {  // HandleSuccessfulStep
      x = cand; x_norm = 0; for (double v : x) x_norm += v * v; x_norm = std::sqrt(x_norm);
      if (!evalGradJac()) { summary->termination = GF2_TERM_FAILURE; break; }
      step_is_successful = true;
      if (relative_decrease < 0.25) radius *= 0.5;                                  // StepAccepted
      if (relative_decrease > 0.75) radius = std::max(radius, 3.0 * dogleg_step_norm);
      radius = std::min(radius, opt.max_trust_region_radius);
      mu = std::max(min_mu, 2.0 * mu / mu_increase);
      reuse = false;
    } else {  // HandleUnsuccessfulStep -> StepRejected
      radius *= 0.5; reuse = true;
    }
    summary->trace.push_back({step_is_successful ? x_cost : candidate_cost, model_cost_change, relative_decrease, radius, step_norm, step_is_successful});
  }
  summary->final_cost = minimum_cost == std::numeric_limits<double>::max() ? x_cost : minimum_cost;
  prog.scatter(best);
}

}  // namespace gf2o
