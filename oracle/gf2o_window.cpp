// TEST INFRASTRUCTURE — CPU oracle. Builds, for one window given in the gf2_abi.h wire layout, the same
// ceres::Problem that Estimator::optimization() builds (VE/estimator/estimator.cpp:3014-3358: parameter blocks
// :3014-3161, prior :3163-3169, IMU :3170-3180, wheel :3181-3212, projection :3330-3358) and solves it with the
// restated Ceres (gf2o_ceres.cpp). Exposes a C API for ctypes (tests, bench cpu_baseline / --impl reference).
#include <thread>
#include <atomic>
#include <memory>
#include <map>
#include <algorithm>
#include "gf2o_ceres.h"
#include "gf2o_marg.h"

using namespace gf2o;

extern "C" {

// One window, pointers into caller arrays (same element layout as the gf2_set_* calls of include/gf2_abi.h).
typedef struct gf2o_window {
  int32_t n_frames;
  int32_t n_landmarks;
  int32_t n_planes;
  int32_t prior_rows;
  int32_t prior_nblocks;
  int32_t use_wheel;
  int32_t prior_stride;   /* row stride of prior_J0 */
  int32_t pad_;
  double* para_pose;      /* [F][7] in/out */
  double* para_speedbias; /* [F][9] in/out */
  double* ex_pose;        /* [7] in/out */
  double* td;             /* [1] in/out */
  double* ex_pose_wheel;  /* [7] */
  double* sxsysw;         /* [3] */
  double* td_wheel;       /* [1] */
  double* inv_depth;      /* [n_landmarks] in/out */
  const int32_t* start_frame;
  const int32_t* track_len;
  const uint8_t* fixed;
  const gf2_obs* obs;     /* landmark-major */
  const double* frame_td; /* [F] */
  const gf2_imu_preint* imu;     /* [F-1] or NULL */
  const gf2_wheel_preint* wheel; /* [F-1] or NULL */
  const double* prior_J0;        /* row stride GF2_MAX_PRIOR_DIM */
  const double* prior_r0;
  const gf2_prior_block* prior_blocks;
  const gf2_plane* planes;
  const double* plane_alpha;     /* [n_planes] alpha_time of the ct == 1 planes, or NULL */
} gf2o_window;

}  // extern "C"

namespace {

struct BuiltWindow {
  Problem problem;
  std::vector<std::unique_ptr<CostFunction>> factors;
  std::vector<std::unique_ptr<IntegrationBase>> imus;
  std::vector<std::unique_ptr<WheelIntegrationBase>> wheels;
  MarginalizationInfo prior;
};

void buildWindow(const gf2o_window& w, const gf2_solve_opts& o, BuiltWindow& bw) {
  Problem& P = bw.problem;
  const int F = w.n_frames;
  ProjectionTwoFrameOneCamFactor::sqrt_info = o.sqrt_info_px;
  IntegrationBase::G = v3(0, 0, o.g_norm);
  LidarPlaneNormFactor::sqrt_info = o.lidar_sqrt_info;
  std::vector<int> id_pose(F), id_sb(F);
  for (int i = 0; i < F; i++) {  // estimator.cpp:3014-3020
    id_pose[i] = P.AddParameterBlock(w.para_pose + 7 * i, 7, true);
    id_sb[i] = P.AddParameterBlock(w.para_speedbias + 9 * i, 9, false);
  }
  int id_ex = P.AddParameterBlock(w.ex_pose, 7, true);  // :3024-3061
  if (o.const_mask & GF2_CONST_EX_POSE) P.SetParameterBlockConstant(id_ex);
  int id_td = P.AddParameterBlock(w.td, 1);  // :3155-3161
  if (o.const_mask & GF2_CONST_TD) P.SetParameterBlockConstant(id_td);
  int id_exw = -1, id_sx = -1, id_sy = -1, id_sw = -1, id_tdw = -1;
  if (w.use_wheel) {  // :3063-3118
    id_exw = P.AddParameterBlock(w.ex_pose_wheel, 7, true);
    if (o.const_mask & GF2_CONST_EX_WHEEL) P.SetParameterBlockConstant(id_exw);
    for (int k = 0; k < 6; k++) P.blocks[id_exw].subset_mask[k] = ((o.wheel_ext_const_components >> k) & 1u) != 0;   // PoseSubsetParameterization
    id_sx = P.AddParameterBlock(w.sxsysw + 0, 1); id_sy = P.AddParameterBlock(w.sxsysw + 1, 1); id_sw = P.AddParameterBlock(w.sxsysw + 2, 1);
    if (o.const_mask & GF2_CONST_WHEEL_INTRINSIC) { P.SetParameterBlockConstant(id_sx); P.SetParameterBlockConstant(id_sy); P.SetParameterBlockConstant(id_sw); }
    id_tdw = P.AddParameterBlock(w.td_wheel, 1);
    if (o.const_mask & GF2_CONST_TD_WHEEL) P.SetParameterBlockConstant(id_tdw);
  }
  auto blockId = [&](const gf2_prior_block& b) -> int {
    switch (b.kind) {
      case GF2_BLK_POSE: return id_pose[b.index];
      case GF2_BLK_SPEEDBIAS: return id_sb[b.index];
      case GF2_BLK_EX_POSE: return id_ex;
      case GF2_BLK_TD: return id_td;
      case GF2_BLK_EX_WHEEL: return id_exw;
      case GF2_BLK_SX: return id_sx;
      case GF2_BLK_SY: return id_sy;
      case GF2_BLK_SW: return id_sw;
      case GF2_BLK_TD_WHEEL: return id_tdw;
    }
    return -1;
  };
  // prior, :3163-3169
  if (w.prior_rows > 0) {
    MarginalizationInfo& mi = bw.prior;
    mi.m = 0; mi.n = w.prior_rows;
    mi.linearized_jacobians.resize((size_t)mi.n * mi.n); mi.linearized_residuals.resize(mi.n);
    for (int r = 0; r < mi.n; r++) { mi.linearized_residuals[r] = w.prior_r0[r]; for (int c = 0; c < mi.n; c++) mi.linearized_jacobians[(size_t)r * mi.n + c] = w.prior_J0[(size_t)r * w.prior_stride + c]; }
    std::vector<int> ids;
    for (int b = 0; b < w.prior_nblocks; b++) {
      const gf2_prior_block& pb = w.prior_blocks[b];
      int size = (pb.kind == GF2_BLK_POSE || pb.kind == GF2_BLK_EX_POSE || pb.kind == GF2_BLK_EX_WHEEL) ? 7 : (pb.kind == GF2_BLK_SPEEDBIAS ? 9 : 1);
      mi.keep_block_size.push_back(size); mi.keep_block_idx.push_back(pb.offset);
      mi.keep_block_data.emplace_back(pb.x0, pb.x0 + size);
      ids.push_back(blockId(pb));
    }
    bw.factors.emplace_back(new MarginalizationFactor(&mi));
    P.AddResidualBlock(bw.factors.back().get(), false, ids);
  }
  // IMU, :3170-3180
  if (w.imu) for (int i = 0; i + 1 < F; i++) {
    if (!w.imu[i].valid || w.imu[i].sum_dt > 10.0) continue;
    bw.imus.emplace_back(new IntegrationBase(w.imu[i]));
    bw.factors.emplace_back(new IMUFactor(bw.imus.back().get()));
    P.AddResidualBlock(bw.factors.back().get(), false, {id_pose[i], id_sb[i], id_pose[i + 1], id_sb[i + 1]});
  }
  // wheel, :3181-3212
  if (w.use_wheel && w.wheel) for (int i = 0; i + 1 < F; i++) {
    if (!w.wheel[i].valid || w.wheel[i].sum_dt > 10.0) continue;
    bw.wheels.emplace_back(new WheelIntegrationBase(w.wheel[i]));
    bw.factors.emplace_back(new WheelFactor(bw.wheels.back().get()));
    P.AddResidualBlock(bw.factors.back().get(), false, {id_pose[i], id_pose[i + 1], id_exw, id_sx, id_sy, id_sw, id_tdw});
  }
  // LiDAR planes (BASELINE.json config 4 composition)
  for (int k = 0; k < w.n_planes; k++) {
    const gf2_plane& pl = w.planes[k];
    if (pl.ct) {  // CTLidarPlaneNormFactor between the window poses frame (begin) and frame + 1 (end)
      CTLidarPlaneNormFactor::sqrt_info = o.lidar_sqrt_info;
      bw.factors.emplace_back(new CTLidarPlanePoseFactor(v3(pl.p_body[0], pl.p_body[1], pl.p_body[2]), v3(pl.normal[0], pl.normal[1], pl.normal[2]), pl.offset,
                                                         w.plane_alpha ? w.plane_alpha[k] : 0.0, pl.weight));
      P.AddResidualBlock(bw.factors.back().get(), false, {id_pose[pl.frame], id_pose[pl.frame + 1]});
      continue;
    }
    bw.factors.emplace_back(new LidarPlanePoseFactor(v3(pl.p_body[0], pl.p_body[1], pl.p_body[2]), v3(pl.normal[0], pl.normal[1], pl.normal[2]), pl.offset, pl.weight));
    P.AddResidualBlock(bw.factors.back().get(), false, {id_pose[pl.frame]});
  }
  // projection, :3330-3358
  int ob = 0;
  for (int l = 0; l < w.n_landmarks; l++) {
    int id_l = P.AddParameterBlock(w.inv_depth + l, 1);
    if (w.fixed && w.fixed[l]) P.SetParameterBlockConstant(id_l); else P.blocks[id_l].eliminate = true;
    const int imu_i = w.start_frame[l];
    const gf2_obs& oi = w.obs[ob];
    V3 pts_i = v3(oi.x, oi.y, 1.0); double vi[2] = {oi.vx, oi.vy};
    for (int k = 1; k < w.track_len[l]; k++) {
      const int imu_j = imu_i + k;
      const gf2_obs& oj = w.obs[ob + k];
      V3 pts_j = v3(oj.x, oj.y, 1.0); double vj[2] = {oj.vx, oj.vy};
      bw.factors.emplace_back(new ProjectionTwoFrameOneCamFactor(pts_i, pts_j, vi, vj, w.frame_td[imu_i], w.frame_td[imu_j]));
      P.AddResidualBlock(bw.factors.back().get(), true, {id_pose[imu_i], id_pose[imu_j], id_ex, id_l, id_td});
    }
    ob += w.track_len[l];
  }
}

SolverOptions toOptions(const gf2_solve_opts& o) {
  SolverOptions s;
  s.max_num_iterations = o.max_iterations;
  s.huber_delta = o.huber_delta;
  if (o.initial_radius > 0) s.initial_trust_region_radius = o.initial_radius;
  if (o.function_tolerance > 0) s.function_tolerance = o.function_tolerance;
  if (o.gradient_tolerance > 0) s.gradient_tolerance = o.gradient_tolerance;
  if (o.parameter_tolerance > 0) s.parameter_tolerance = o.parameter_tolerance;
  return s;
}

}  // namespace

extern "C" {

int gf2o_solve_window(const gf2o_window* w, const gf2_solve_opts* opts, gf2_solve_summary* out, double* trace /* [max_it][6] or NULL */) {
  BuiltWindow bw;
  buildWindow(*w, *opts, bw);
  SolverSummary s;
  Solve(toOptions(*opts), &bw.problem, &s);
  if (out) { out->initial_cost = s.initial_cost; out->final_cost = s.final_cost; out->iterations = s.iterations; out->successful_steps = s.successful_steps; out->termination = s.termination; out->pad_ = 0; }
  if (trace) for (size_t i = 0; i < s.trace.size() && (int)i < opts->max_iterations; i++) {
    trace[6 * i + 0] = s.trace[i].cost; trace[6 * i + 1] = s.trace[i].model_cost_change; trace[6 * i + 2] = s.trace[i].relative_decrease;
    trace[6 * i + 3] = s.trace[i].radius; trace[6 * i + 4] = s.trace[i].step_norm; trace[6 * i + 5] = s.trace[i].successful;
  }
  return 0;
}

// One linearisation: S [D*D], g [D], cost; returns D (or -1 if D > max_dim)
int gf2o_linearize_window(const gf2o_window* w, const gf2_solve_opts* opts, int max_dim, double* S, double* g, double* cost,
                          double* ete, double* etr) {
  BuiltWindow bw;
  buildWindow(*w, *opts, bw);
  Linearization lin;
  Linearize(toOptions(*opts), &bw.problem, &lin);
  if (lin.D > max_dim) return -1;
  std::memcpy(S, lin.S.data(), sizeof(double) * lin.D * lin.D);
  std::memcpy(g, lin.g.data(), sizeof(double) * lin.D);
  *cost = lin.cost;
  if (ete) std::memcpy(ete, lin.ete.data(), sizeof(double) * lin.E);
  if (etr) std::memcpy(etr, lin.etr.data(), sizeof(double) * lin.E);
  return lin.D;
}

// Batched solve over independent windows with a thread pool: the CPU baseline ("port") of bench.py.
// Arrays use the same window-major strides as the gf2_set_* calls.
typedef struct gf2o_batch {
  int32_t n_windows, n_frames, max_landmarks, max_obs, max_planes, use_wheel, prior_stride, pad_;
  double *para_pose, *para_speedbias, *ex_pose, *td, *ex_pose_wheel, *sxsysw, *td_wheel, *inv_depth;
  const int32_t *n_landmarks, *start_frame, *track_len;
  const uint8_t* fixed;
  const gf2_obs* obs;
  const double* frame_td;
  const gf2_imu_preint* imu;
  const gf2_wheel_preint* wheel;
  const int32_t* prior_rows; const double *prior_J0, *prior_r0; const int32_t* prior_nblocks; const gf2_prior_block* prior_blocks;
  const int32_t* n_planes; const gf2_plane* planes;
  const double* plane_alpha;   /* [n_windows][max_planes] or NULL */
} gf2o_batch;

static void windowOf(const gf2o_batch& b, int i, gf2o_window& w) {
  const int F = b.n_frames;
  std::memset(&w, 0, sizeof(w));
  w.n_frames = F; w.use_wheel = b.use_wheel;
  w.n_landmarks = b.n_landmarks[i];
  w.para_pose = b.para_pose + (size_t)i * F * 7; w.para_speedbias = b.para_speedbias + (size_t)i * F * 9;
  w.ex_pose = b.ex_pose + (size_t)i * 7; w.td = b.td + i;
  if (b.use_wheel) { w.ex_pose_wheel = b.ex_pose_wheel + (size_t)i * 7; w.sxsysw = b.sxsysw + (size_t)i * 3; w.td_wheel = b.td_wheel + i; }
  w.inv_depth = b.inv_depth + (size_t)i * b.max_landmarks;
  w.start_frame = b.start_frame + (size_t)i * b.max_landmarks; w.track_len = b.track_len + (size_t)i * b.max_landmarks;
  w.fixed = b.fixed ? b.fixed + (size_t)i * b.max_landmarks : nullptr;
  w.obs = b.obs + (size_t)i * b.max_obs; w.frame_td = b.frame_td + (size_t)i * F;
  w.imu = b.imu ? b.imu + (size_t)i * (F - 1) : nullptr;
  w.wheel = (b.use_wheel && b.wheel) ? b.wheel + (size_t)i * (F - 1) : nullptr;
  if (b.prior_rows) {
    w.prior_rows = b.prior_rows[i]; w.prior_nblocks = b.prior_nblocks[i];
    w.prior_stride = b.prior_stride; w.prior_J0 = b.prior_J0 + (size_t)i * b.prior_stride * b.prior_stride; w.prior_r0 = b.prior_r0 + (size_t)i * b.prior_stride;
    w.prior_blocks = b.prior_blocks + (size_t)i * (2 * F + 8);
  }
  if (b.n_planes && b.max_planes > 0) { w.n_planes = b.n_planes[i]; w.planes = b.planes + (size_t)i * b.max_planes; w.plane_alpha = b.plane_alpha ? b.plane_alpha + (size_t)i * b.max_planes : nullptr; }
}

int gf2o_batch_window(const gf2o_batch* b, int i, gf2o_window* w) { windowOf(*b, i, *w); return 0; }

int gf2o_solve_batch(const gf2o_batch* b, const gf2_solve_opts* opts, gf2_solve_summary* summaries, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  std::atomic<int> next(0);
  // the factor statics are process-wide in the reference too; set once before the threads start
  ProjectionTwoFrameOneCamFactor::sqrt_info = opts->sqrt_info_px;
  auto worker = [&]() {
    for (;;) {
      int i = next.fetch_add(1);
      if (i >= b->n_windows) break;
      gf2o_window w; windowOf(*b, i, w);
      gf2o_solve_window(&w, opts, summaries ? summaries + i : nullptr, nullptr);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; t++) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
  return 0;
}

// ---- single-factor evaluation for the finite-difference / golden tests -------------------------
// kind: 0 projection  consts = [xi yi vxi vyi tdi xj yj vxj vyj tdj sqrt_info]     params 7,7,7,1,1
//       1 imu         consts = gf2_imu_preint* (as bytes), extra[0] = g_norm       params 7,9,7,9
//       2 wheel       consts = gf2_wheel_preint*                                   params 7,7,7,1,1,1,1
//       3 lidar plane consts = [p(3) n(3) offset weight sqrt_info]                 params 3,4
//       4 ct plane    consts = [kp(3) n(3) offset alpha weight sqrt_info]          params 3,4,3,4
//       5 plane@pose  consts as 3                                                  params 7
// params_flat: the blocks concatenated; jac_flat: the row-major Jacobian blocks concatenated (ambient sizes).
int gf2o_factor_eval(int kind, const void* consts, const double* extra, const double* params_flat, double* residuals,
                     double* jac_flat) {
  std::unique_ptr<CostFunction> f;
  std::unique_ptr<IntegrationBase> ib; std::unique_ptr<WheelIntegrationBase> wb;
  const double* c = (const double*)consts;
  switch (kind) {
    case 0: { double vi[2] = {c[2], c[3]}, vj[2] = {c[7], c[8]};
      ProjectionTwoFrameOneCamFactor::sqrt_info = c[10];
      f.reset(new ProjectionTwoFrameOneCamFactor(v3(c[0], c[1], 1), v3(c[5], c[6], 1), vi, vj, c[4], c[9])); break; }
    case 1: ib.reset(new IntegrationBase(*(const gf2_imu_preint*)consts)); IntegrationBase::G = v3(0, 0, extra[0]); f.reset(new IMUFactor(ib.get())); break;
    case 2: wb.reset(new WheelIntegrationBase(*(const gf2_wheel_preint*)consts)); f.reset(new WheelFactor(wb.get())); break;
    case 3: LidarPlaneNormFactor::sqrt_info = c[8]; f.reset(new LidarPlaneNormFactor(v3(c[0], c[1], c[2]), v3(c[3], c[4], c[5]), c[6], c[7])); break;
    case 4: CTLidarPlaneNormFactor::sqrt_info = c[9]; f.reset(new CTLidarPlaneNormFactor(v3(c[0], c[1], c[2]), v3(c[3], c[4], c[5]), c[6], c[7], c[8])); break;
    case 5: LidarPlaneNormFactor::sqrt_info = c[8]; f.reset(new LidarPlanePoseFactor(v3(c[0], c[1], c[2]), v3(c[3], c[4], c[5]), c[6], c[7])); break;
    case 6: CTLidarPlaneNormFactor::sqrt_info = c[9]; f.reset(new CTLidarPlanePoseFactor(v3(c[0], c[1], c[2]), v3(c[3], c[4], c[5]), c[6], c[7], c[8])); break;
    default: return -1;
  }
  std::vector<const double*> pp; std::vector<double*> jj; size_t po = 0, jo = 0;
  for (int s : f->block_sizes) { pp.push_back(params_flat + po); po += s; jj.push_back(jac_flat ? jac_flat + jo : nullptr); jo += (size_t)s * f->num_residuals; }
  f->Evaluate(pp.data(), residuals, jac_flat ? jj.data() : nullptr);
  return f->num_residuals;
}

// IntegrationBase chain: first sample (acc_0, gyr_0), linearisation biases, then push_back per sample.
int gf2o_imu_preintegrate(const gf2_imu_sample* samples, int n, const double first[6], const double lin_bias[6],
                          const double noise[4], gf2_imu_preint* out) {
  IntegrationBase ib(v3(first[0], first[1], first[2]), v3(first[3], first[4], first[5]), v3(lin_bias[0], lin_bias[1], lin_bias[2]),
                     v3(lin_bias[3], lin_bias[4], lin_bias[5]), noise[0], noise[1], noise[2], noise[3]);
  for (int i = 0; i < n; i++) ib.push_back(samples[i].dt, v3(samples[i].acc[0], samples[i].acc[1], samples[i].acc[2]), v3(samples[i].gyr[0], samples[i].gyr[1], samples[i].gyr[2]));
  ib.pack(out);
  return 0;
}
int gf2o_wheel_preintegrate(const gf2_wheel_sample* samples, int n, const double first[6], const double lin[4] /* sx sy sw td */,
                            const double noise[2] /* vel_n gyr_n */, gf2_wheel_preint* out) {
  WheelIntegrationBase wb(v3(first[0], first[1], first[2]), v3(first[3], first[4], first[5]), lin[0], lin[1], lin[2], lin[3], noise[0], noise[1]);
  for (int i = 0; i < n; i++) wb.push_back(samples[i].dt, v3(samples[i].vel[0], samples[i].vel[1], samples[i].vel[2]), v3(samples[i].gyr[0], samples[i].gyr[1], samples[i].gyr[2]));
  wb.pack(out);
  return 0;
}

// Marginalization after the solve (VE/estimator/estimator.cpp:3394-3690). mode 0 = MARGIN_OLD, 1 = MARGIN_SECOND_NEW.
// Output = the kept part of the new MarginalizationInfo with the blocks already renamed by addr_shift (:3561-3595 /
// :3653-3681), i.e. in the indexing of the window AFTER slideWindow: J0 [n][P] row stride P, r0 [n], blocks.
// Returns n (>= 0), -1 if the new prior is invalid (m == 0), -2 if nothing is marginalized (SECOND_NEW without the
// second-newest pose in the old prior: the old prior stays), -3 on an unsupported input.
int gf2o_marginalize_window(const gf2o_window* w, const gf2_solve_opts* o, int mode, int P, double* J0, double* r0, int32_t* n_blocks,
                            gf2_prior_block* blocks, int32_t* m_out) {
  const int F = w->n_frames;
  ProjectionTwoFrameOneCamFactor::sqrt_info = o->sqrt_info_px;
  IntegrationBase::G = v3(0, 0, o->g_norm);
  gf2o_marg::Info info;
  std::vector<std::unique_ptr<CostFunction>> owned;
  MarginalizationInfo last;  // the kept part of last_marginalization_info
  std::unique_ptr<IntegrationBase> imu0; std::unique_ptr<WheelIntegrationBase> wheel0;
  auto addrOf = [&](const gf2_prior_block& b) -> double* {
    switch (b.kind) {
      case GF2_BLK_POSE: return w->para_pose + 7 * b.index;
      case GF2_BLK_SPEEDBIAS: return w->para_speedbias + 9 * b.index;
      case GF2_BLK_EX_POSE: return w->ex_pose;
      case GF2_BLK_TD: return w->td;
      case GF2_BLK_EX_WHEEL: return w->ex_pose_wheel;
      case GF2_BLK_SX: return w->sxsysw + 0;
      case GF2_BLK_SY: return w->sxsysw + 1;
      case GF2_BLK_SW: return w->sxsysw + 2;
      case GF2_BLK_TD_WHEEL: return w->td_wheel;
    }
    return nullptr;
  };
  std::vector<double*> last_blocks;
  if (w->prior_rows > 0) {
    last.m = 0; last.n = w->prior_rows;
    last.linearized_jacobians.resize((size_t)last.n * last.n); last.linearized_residuals.resize(last.n);
    for (int r = 0; r < last.n; r++) { last.linearized_residuals[r] = w->prior_r0[r]; for (int c = 0; c < last.n; c++) last.linearized_jacobians[(size_t)r * last.n + c] = w->prior_J0[(size_t)r * w->prior_stride + c]; }
    for (int b = 0; b < w->prior_nblocks; b++) {
      const gf2_prior_block& pb = w->prior_blocks[b];
      int size = (pb.kind == GF2_BLK_POSE || pb.kind == GF2_BLK_EX_POSE || pb.kind == GF2_BLK_EX_WHEEL) ? 7 : (pb.kind == GF2_BLK_SPEEDBIAS ? 9 : 1);
      last.keep_block_size.push_back(size); last.keep_block_idx.push_back(pb.offset); last.keep_block_data.emplace_back(pb.x0, pb.x0 + size);
      last_blocks.push_back(addrOf(pb));
    }
  }
  auto add = [&](CostFunction* f, bool loss, std::vector<double*> blocks_, std::vector<int> drop) {
    owned.emplace_back(f);
    auto* r = new gf2o_marg::ResidualBlockInfo(); r->cost_function = f; r->loss = loss; r->parameter_blocks = blocks_; r->drop_set = drop;
    info.addResidualBlockInfo(r);
  };
  std::map<double*, std::pair<int, int>> shift;  // address -> (kind, index) after slideWindow
  if (mode == 0) {
    if (w->prior_rows > 0) {  // :3401-3415
      std::vector<int> drop;
      for (size_t i = 0; i < last_blocks.size(); i++) if (last_blocks[i] == w->para_pose || last_blocks[i] == w->para_speedbias) drop.push_back((int)i);
      add(new MarginalizationFactor(&last), false, last_blocks, drop);
    }
    if (w->imu && w->imu[0].valid && w->imu[0].sum_dt < 10.0) {  // :3416-3427
      imu0.reset(new IntegrationBase(w->imu[0]));
      add(new IMUFactor(imu0.get()), false, {w->para_pose, w->para_speedbias, w->para_pose + 7, w->para_speedbias + 9}, {0, 1});
    }
    if (w->use_wheel && w->wheel && w->wheel[0].valid && w->wheel[0].sum_dt < 10.0) {  // :3428-3439
      wheel0.reset(new WheelIntegrationBase(w->wheel[0]));
      add(new WheelFactor(wheel0.get()), false, {w->para_pose, w->para_pose + 7, w->ex_pose_wheel, w->sxsysw, w->sxsysw + 1, w->sxsysw + 2, w->td_wheel}, {0});
    }
    int ob = 0;  // :3495-3528
    for (int l = 0; l < w->n_landmarks; l++) {
      const int imu_i = w->start_frame[l];
      if (imu_i == 0) {
        const gf2_obs& oi = w->obs[ob];
        V3 pts_i = v3(oi.x, oi.y, 1.0); double vi[2] = {oi.vx, oi.vy};
        for (int k = 1; k < w->track_len[l]; k++) {
          const gf2_obs& oj = w->obs[ob + k];
          V3 pts_j = v3(oj.x, oj.y, 1.0); double vj[2] = {oj.vx, oj.vy};
          add(new ProjectionTwoFrameOneCamFactor(pts_i, pts_j, vi, vj, w->frame_td[0], w->frame_td[k]), true,
              {w->para_pose, w->para_pose + 7 * k, w->ex_pose, w->inv_depth + l, w->td}, {0, 3});
        }
      }
      ob += w->track_len[l];
    }
    for (int i = 1; i < F; i++) { shift[w->para_pose + 7 * i] = {GF2_BLK_POSE, i - 1}; shift[w->para_speedbias + 9 * i] = {GF2_BLK_SPEEDBIAS, i - 1}; }  // :3561-3570
  } else {
    const int sn = F - 2;  // WINDOW_SIZE - 1
    if (!(w->prior_rows > 0) || !std::count(last_blocks.begin(), last_blocks.end(), w->para_pose + 7 * sn)) return -2;  // :3599-3600
    std::vector<int> drop;
    for (size_t i = 0; i < last_blocks.size(); i++) {
      if (last_blocks[i] == w->para_speedbias + 9 * sn) return -3;  // ROS_ASSERT, :3612
      if (last_blocks[i] == w->para_pose + 7 * sn) drop.push_back((int)i);
    }
    add(new MarginalizationFactor(&last), false, last_blocks, drop);
    for (int i = 0; i < F; i++) {  // :3653-3676
      if (i == sn) continue;
      const int to = (i == F - 1) ? i - 1 : i;
      shift[w->para_pose + 7 * i] = {GF2_BLK_POSE, to}; shift[w->para_speedbias + 9 * i] = {GF2_BLK_SPEEDBIAS, to};
    }
  }
  shift[w->ex_pose] = {GF2_BLK_EX_POSE, 0}; shift[w->td] = {GF2_BLK_TD, 0};
  if (w->use_wheel) {
    shift[w->ex_pose_wheel] = {GF2_BLK_EX_WHEEL, 0}; shift[w->sxsysw] = {GF2_BLK_SX, 0}; shift[w->sxsysw + 1] = {GF2_BLK_SY, 0};
    shift[w->sxsysw + 2] = {GF2_BLK_SW, 0}; shift[w->td_wheel] = {GF2_BLK_TD_WHEEL, 0};
  }
  info.preMarginalize(o->huber_delta);
  info.marginalize(1e-8);  // eps, marginalization_factor.h:83
  if (m_out) *m_out = info.m;
  if (!info.valid) return -1;
  if (info.n > P) return -3;
  // getParameterBlocks, :310-330
  int nb = 0;
  for (double* a : info.order) {
    if (info.idx[a] < info.m) continue;
    auto it = shift.find(a);
    if (it == shift.end()) return -3;
    gf2_prior_block& pb = blocks[nb++];
    std::memset(&pb, 0, sizeof(pb));
    pb.kind = it->second.first; pb.index = it->second.second; pb.offset = info.idx[a] - info.m;
    const std::vector<double>& d = info.data[a];
    for (size_t k = 0; k < d.size(); k++) pb.x0[k] = d[k];
  }
  *n_blocks = nb;
  for (int r = 0; r < info.n; r++) { r0[r] = info.linearized_residuals[r]; for (int c = 0; c < info.n; c++) J0[(size_t)r * P + c] = info.linearized_jacobians[(size_t)r * info.n + c]; }
  return info.n;
}

int gf2o_sym_eigen(int n, const double* A, double* evals, double* evecs) { symEigen(n, A, evals, evecs); return 0; }
int gf2o_sizeof(int what) {
  switch (what) { case 0: return sizeof(gf2_imu_preint); case 1: return sizeof(gf2_wheel_preint); case 2: return sizeof(gf2_plane);
    case 3: return sizeof(gf2_prior_block); case 4: return sizeof(gf2o_window); case 5: return sizeof(gf2o_batch); case 6: return sizeof(gf2_solve_opts); case 7: return sizeof(gf2_solve_summary); }
  return -1;
}

}  // extern "C"
