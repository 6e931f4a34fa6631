// gf2_solver_kernels.cuh — device data layout and kernels of the batched sliding-window solver.
//
// One window = the problem Estimator::optimization() hands to ceres::Solve (VE/estimator/estimator.cpp:2951-3392).
// A batch of windows is solved in lock-step, one CTA per window per kernel, four kernels per trust-region iteration:
//
//   k_linearize   sweep 1: every ProjectionTwoFrameOneCamFactor (+ LiDAR plane factor) evaluated once, Huber-corrected,
//                 pose-block J^T J accumulated by warp-shuffle reductions, landmarks Schur-eliminated with an fp64
//                 tensor-core SYRK (mma.sync m8n8k4.f64); writes the 6F x 6F visual reduced system (17 KB/window).
//   k_solve       assembles the full reduced camera system in shared memory (visual part + IMU + wheel + prior),
//                 Jacobi-scaling / dogleg diagonal, Cholesky, Gauss-Newton step.
//   k_backsub     sweep 2: landmark back-substitution and the dot products the dogleg step needs.
//   k_candidate   sweep 3: dogleg interpolation, retraction (PoseLocalParameterization::Plus), candidate cost,
//                 accept / reject / terminate exactly as Ceres 1.14's TrustRegionMinimizer + DoglegStrategy.
//
// Tangent order of the reduced system (D = 15 F): per frame [pose 6 | speed-bias 9].
#pragma once
#include <stdint.h>
#include "gf2_math.cuh"
#include "../../include/gf2_abi.h"

namespace gf2 {

constexpr int kMaxF = GF2_MAX_FRAMES;
constexpr int kNVMax = 6 * kMaxF;          // 66 visual tangent dims
constexpr int kNVP = 72;                   // padded to 9 mma tiles of 8; column 66 carries the landmark gradient
constexpr int kLinThreads = 256;
constexpr int kWTStride = kLinThreads + 4; // transposed W tile: [kNVP][kWTStride] doubles
constexpr int kSolveThreads = 256;
constexpr int kP = GF2_MAX_PRIOR_DIM;

struct WinState {  // per-window trust-region state (DoglegStrategy + TrustRegionMinimizer members)
  double radius, mu, x_cost, cand_cost, model_cost_change, dogleg_step_norm, x_norm2;
  double coef_a, coef_b;            // delta = -(a*u + b*z)
  double initial_cost;
  // x-part sums from k_solve, landmark-part sums from k_backsub
  double dlg2_x, gn2_x, gz_x, zEz_x, uEz_x, uSu, uEu_x, gmax_x;
  double dlg2_l, gn2_l, gz_l, zEz_l, uEz_l, uHu_l, gmax_l;
  double cost_vis;                  // from k_linearize
  int32_t iteration, successful, termination, active, reuse, invalid_count, lin_valid, pad_;
};

struct KP {  // kernel parameters (device pointers are window-major with the strides below)
  int nW, F, Lm, Om, Pm, D, use_wheel;
  uint32_t const_mask;
  double huber, sqrt_info_px, g_norm, lidar_sqrt_info;
  double ftol, gtol, ptol;
  int max_iterations;
  // states (current / candidate)
  double *pose, *sb, *ex, *td, *exw, *sxw, *tdw, *invdep;
  double *pose_c, *sb_c, *invdep_c;
  // landmarks
  const int32_t *nlm, *start, *tlen, *obeg;
  const uint8_t* fixed;
  const float4* obs;
  const double* frame_td;
  // IMU
  const gf2_imu_preint* imu;
  double* imu_sqrt;  // [nW][F-1][225] upper-triangular sqrt_info
  // wheel
  const gf2_wheel_preint* wheel;
  double* wheel_sqrt;  // [nW][F-1][36]
  // prior
  const int32_t *prior_rows, *prior_nblocks;
  const double *prior_J0, *prior_r0;
  const gf2_prior_block* prior_blocks;
  double* prior_H;    // [nW][P][P] = J0^T J0
  int32_t* prior_map; // [nW][P] column -> tangent index (or -1)
  // planes
  const int32_t* n_planes;
  const gf2_plane* planes;
  // work
  double *Svis, *gvis, *Udiag;     // [nW][66*66], [nW][72], [nW][66]
  double *lm_v, *lm_g, *lm_s, *lm_z; // [nW][Lm]
  double *sx, *zx, *ux, *ex_diag;  // [nW][D] jacobi scale, GN step, u, e
  double *Sfull, *gfull;           // optional dump of the assembled reduced system [nW][D*D], [nW][D]
  WinState* st;
};

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum of up to NV values per thread; result valid in thread 0 (red must hold NV * 32 doubles)
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) { double s = warp_sum(v[i]); if (lane == 0) red[i * 32 + wid] = s; }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) { double s = lane < nw ? red[i * 32 + lane] : 0.0; s = warp_sum(s); v[i] = s; }
  }
  __syncthreads();
}

struct FrameCtx {  // per-frame quantities shared by all factors of a window
  double P[3];
  double R[9];  // R_wb
  double A[9];  // ric^T R^T
};
struct CamCtx {
  double ric[9], tic[3], rtt[3];  // rtt = ric^T tic
  double td;
};

__device__ __forceinline__ void build_frames(const double* pose, const double* ex, double td, int F, FrameCtx* fr, CamCtx* cam) {
  const int t = threadIdx.x;
  if (t == 0) {
    M3 ric = toR(ldq(ex + 3)); V3 tic = ld3(ex);
    for (int i = 0; i < 9; i++) cam->ric[i] = ric.m[i];
    cam->tic[0] = tic.x; cam->tic[1] = tic.y; cam->tic[2] = tic.z;
    V3 r = mulT(ric, tic); cam->rtt[0] = r.x; cam->rtt[1] = r.y; cam->rtt[2] = r.z;
    cam->td = td;
  }
  if (t < F) {
    const double* p = pose + 7 * t;
    M3 R = toR(ldq(p + 3)); M3 ric = toR(ldq(ex + 3));
    M3 A = mulBT(transpose(ric), R);  // ric^T R^T
    for (int i = 0; i < 9; i++) { fr[t].R[i] = R.m[i]; fr[t].A[i] = A.m[i]; }
    fr[t].P[0] = p[0]; fr[t].P[1] = p[1]; fr[t].P[2] = p[2];
  }
}

struct LmCtx {  // per-landmark quantities (host frame side of ProjectionTwoFrameOneCamFactor)
  V3 Xw;        // pts_w
  V3 dXdl;      // d pts_w / d inv_dep = Ri ric pts_i_td * (-1/inv_dep^2)
  M3 Gi;        // d pts_w / d theta_i = -Ri skew(pts_imu_i)
};

// host-frame part of VE/factor/projectionTwoFrameOneCamFactor.cpp:59-64 and the pieces of :102-140 that do not depend on j
__device__ __forceinline__ void landmark_ctx(const FrameCtx& fi, const CamCtx& cam, float4 oi, double td_i, double inv_dep, LmCtx& lc) {
  const double dt = cam.td - td_i;
  V3 pts_i_td = mk3((double)oi.x - dt * (double)oi.z, (double)oi.y - dt * (double)oi.w, 1.0);
  V3 pc = mk3(pts_i_td.x / inv_dep, pts_i_td.y / inv_dep, pts_i_td.z / inv_dep);
  M3 ric; for (int i = 0; i < 9; i++) ric.m[i] = cam.ric[i];
  M3 Ri; for (int i = 0; i < 9; i++) Ri.m[i] = fi.R[i];
  V3 p_imu = mul(ric, pc) + mk3(cam.tic[0], cam.tic[1], cam.tic[2]);
  lc.Xw = mul(Ri, p_imu) + mk3(fi.P[0], fi.P[1], fi.P[2]);
  lc.Gi = scale(mul(Ri, skew(p_imu)), -1.0);
  lc.dXdl = mul(Ri, mul(ric, pts_i_td)) * (-1.0 / (inv_dep * inv_dep));
}

// residual of one observation in frame j (projectionTwoFrameOneCamFactor.cpp:59-76), Huber weight applied by caller
__device__ __forceinline__ void obs_residual(const FrameCtx& fj, const CamCtx& cam, const LmCtx& lc, float4 oj, double td_j,
                                             double sqrt_info, double& r0, double& r1, V3& pcj) {
  const double dt = cam.td - td_j;
  V3 d = lc.Xw - mk3(fj.P[0], fj.P[1], fj.P[2]);
  pcj = mk3(fj.A[0] * d.x + fj.A[1] * d.y + fj.A[2] * d.z - cam.rtt[0], fj.A[3] * d.x + fj.A[4] * d.y + fj.A[5] * d.z - cam.rtt[1],
            fj.A[6] * d.x + fj.A[7] * d.y + fj.A[8] * d.z - cam.rtt[2]);
  const double ptx = (double)oj.x - dt * (double)oj.z, pty = (double)oj.y - dt * (double)oj.w;
  r0 = sqrt_info * (pcj.x / pcj.z - ptx);
  r1 = sqrt_info * (pcj.y / pcj.z - pty);
}

// ceres::HuberLoss + Corrector (rho'' <= 0 -> both r and J scaled by sqrt(rho')), restated in-tree at
// VE/factor/marginalization_factor.cpp:46-77. Returns rho(s)/... : cost contribution 0.5*rho0 and the scale.
__device__ __forceinline__ void huber(double delta, double sq, double& half_rho, double& scl) {
  const double b = delta * delta;
  if (sq > b) { const double r = sqrt(sq); half_rho = 0.5 * (2.0 * delta * r - b); scl = sqrt(fmax(2.2250738585072014e-308, delta / r)); }
  else { half_rho = 0.5 * sq; scl = 1.0; }
}

// Jacobians of one observation wrt d pts_w (Jx, 2x3) and pose_j (Jj, 2x6), projectionTwoFrameOneCamFactor.cpp:83-124
__device__ __forceinline__ void obs_jacobians(const FrameCtx& fj, const CamCtx& cam, const LmCtx& lc, V3 pcj, double sqrt_info,
                                              double (&Jx)[6], double (&Jj)[12]) {
  const double iz = 1.0 / pcj.z;
  const double a = sqrt_info * iz, bx = -sqrt_info * pcj.x * iz * iz, by = -sqrt_info * pcj.y * iz * iz;
  // reduce = sqrt_info * [[1/z, 0, -x/z^2], [0, 1/z, -y/z^2]];  Jx = reduce * A_j
#pragma unroll
  for (int c = 0; c < 3; c++) { Jx[c] = a * fj.A[c] + bx * fj.A[6 + c]; Jx[3 + c] = a * fj.A[3 + c] + by * fj.A[6 + c]; }
  // B = reduce * ric^T  (2x3)
  double B[6];
#pragma unroll
  for (int c = 0; c < 3; c++) { B[c] = a * cam.ric[c * 3 + 0] + bx * cam.ric[c * 3 + 2]; B[3 + c] = a * cam.ric[c * 3 + 1] + by * cam.ric[c * 3 + 2]; }
  // pts_imu_j = ric * pcj + tic
  V3 pj = mk3(cam.ric[0] * pcj.x + cam.ric[1] * pcj.y + cam.ric[2] * pcj.z + cam.tic[0], cam.ric[3] * pcj.x + cam.ric[4] * pcj.y + cam.ric[5] * pcj.z + cam.tic[1],
              cam.ric[6] * pcj.x + cam.ric[7] * pcj.y + cam.ric[8] * pcj.z + cam.tic[2]);
  // Jj = [ -Jx | B * skew(pj) ];  row * skew(p) = (row x ... ) : (b^T [p]x)_c
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const double b0 = B[r * 3], b1 = B[r * 3 + 1], b2 = B[r * 3 + 2];
    Jj[r * 6 + 0] = -Jx[r * 3 + 0]; Jj[r * 6 + 1] = -Jx[r * 3 + 1]; Jj[r * 6 + 2] = -Jx[r * 3 + 2];
    Jj[r * 6 + 3] = b1 * pj.z - b2 * pj.y;
    Jj[r * 6 + 4] = b2 * pj.x - b0 * pj.z;
    Jj[r * 6 + 5] = b0 * pj.y - b1 * pj.x;
  }
}

__device__ __forceinline__ void mma_f64(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------ k_prepare
// Once per solve: IMU/wheel sqrt_info = LLT(cov^-1).L^T (VE/factor/imu_factor.h:73, wheel_factor.h:85), prior
// H = J0^T J0 and its column map, trust-region state reset, ambient norm of x.
template <int N>
__device__ void sqrt_info_from_cov(const double* cov, double* out /* N*N row-major, upper triangular */, double* scratch /* 2*N*N */) {
  double* m = scratch;  // [N][2N] Gauss-Jordan with partial pivoting
  for (int r = 0; r < N; r++) for (int c = 0; c < N; c++) { m[r * 2 * N + c] = cov[r * N + c]; m[r * 2 * N + N + c] = (r == c) ? 1.0 : 0.0; }
  for (int col = 0; col < N; col++) {
    int piv = col; double best = fabs(m[col * 2 * N + col]);
    for (int r = col + 1; r < N; r++) { double v = fabs(m[r * 2 * N + col]); if (v > best) { best = v; piv = r; } }
    if (piv != col) for (int c = 0; c < 2 * N; c++) { double t = m[piv * 2 * N + c]; m[piv * 2 * N + c] = m[col * 2 * N + c]; m[col * 2 * N + c] = t; }
    double d = m[col * 2 * N + col];
    for (int c = 0; c < 2 * N; c++) m[col * 2 * N + c] /= d;
    for (int r = 0; r < N; r++) if (r != col) { double f = m[r * 2 * N + col]; if (f != 0.0) for (int c = 0; c < 2 * N; c++) m[r * 2 * N + c] -= f * m[col * 2 * N + c]; }
  }
  // Cholesky of the inverse (lower), in place in the right half
  for (int j = 0; j < N; j++) {
    double d = m[j * 2 * N + N + j];
    for (int k = 0; k < j; k++) d -= m[j * 2 * N + N + k] * m[j * 2 * N + N + k];
    d = sqrt(d); m[j * 2 * N + N + j] = d;
    for (int i = j + 1; i < N; i++) { double s = m[i * 2 * N + N + j]; for (int k = 0; k < j; k++) s -= m[i * 2 * N + N + k] * m[j * 2 * N + N + k]; m[i * 2 * N + N + j] = s / d; }
  }
  for (int r = 0; r < N; r++) for (int c = 0; c < N; c++) out[r * N + c] = (c >= r) ? m[c * 2 * N + N + r] : 0.0;
}

__global__ void k_prepare(KP p, int w0) {
  const int w = w0 + blockIdx.x;
  const int F = p.F, t = threadIdx.x;
  extern __shared__ double sh[];
  // IMU sqrt_info: one thread per factor, scratch in global-backed local arrays would spill; use shared: 2*225 doubles per factor
  if (p.imu && t < F - 1) {
    const gf2_imu_preint& rec = p.imu[(size_t)w * (F - 1) + t];
    double* out = p.imu_sqrt + ((size_t)w * (F - 1) + t) * 225;
    if (rec.valid && rec.sum_dt <= 10.0) sqrt_info_from_cov<15>(rec.covariance, out, sh + t * 450);
  }
  __syncthreads();
  if (p.use_wheel && p.wheel && t < F - 1) {
    const gf2_wheel_preint& rec = p.wheel[(size_t)w * (F - 1) + t];
    double* out = p.wheel_sqrt + ((size_t)w * (F - 1) + t) * 36;
    if (rec.valid && rec.sum_dt <= 10.0) sqrt_info_from_cov<6>(rec.covariance, out, sh + t * 72);
  }
  // prior: column map and H = J0^T J0
  const int n = p.prior_rows ? p.prior_rows[w] : 0;
  if (n > 0) {
    int32_t* map = p.prior_map + (size_t)w * kP;
    const gf2_prior_block* blk = p.prior_blocks + (size_t)w * (2 * F + 8);
    if (t == 0) {
      for (int c = 0; c < kP; c++) map[c] = -1;
      for (int b = 0; b < p.prior_nblocks[w]; b++) {
        int base = -1, ls = 0;
        if (blk[b].kind == GF2_BLK_POSE) { base = 15 * blk[b].index; ls = 6; }
        else if (blk[b].kind == GF2_BLK_SPEEDBIAS) { base = 15 * blk[b].index + 6; ls = 9; }
        else { ls = (blk[b].kind == GF2_BLK_EX_POSE || blk[b].kind == GF2_BLK_EX_WHEEL) ? 6 : 1; }  // constant calibration blocks: no column
        for (int c = 0; c < ls; c++) map[blk[b].offset + c] = base < 0 ? -1 : base + c;
      }
    }
    const double* J0 = p.prior_J0 + (size_t)w * kP * kP;
    double* H = p.prior_H + (size_t)w * kP * kP;
    for (int idx = t; idx < n * n; idx += blockDim.x) {
      const int a = idx / n, b = idx % n;
      double s = 0; for (int r = 0; r < n; r++) s += J0[r * kP + a] * J0[r * kP + b];
      H[a * kP + b] = s;
    }
  }
  if (t == 0) {
    WinState& s = p.st[w];
    s.radius = 1e4; s.mu = 1e-8; s.x_cost = 0; s.cand_cost = 0; s.model_cost_change = 0; s.dogleg_step_norm = 0;
    s.iteration = 0; s.successful = 0; s.termination = GF2_TERM_NO_CONVERGENCE; s.active = 1; s.reuse = 0; s.invalid_count = 0; s.lin_valid = 0;
    s.initial_cost = 0; s.coef_a = 0; s.coef_b = 0; s.x_norm2 = 0;
  }
}

// ------------------------------------------------------------------------------------------------ k_linearize
struct LinShared {
  FrameCtx fr[kMaxF];
  CamCtx cam;
  double U[kNVMax * kNVMax];  // pose-block Hessian of the visual factors, upper block triangle filled
  double g[kNVP];
  double invv[kLinThreads];
  double red[8 * 32];
  double WT[kNVP * kWTStride];
};

// Add the 63 products of one observation (blocks (i,j), (j,j) upper, g_j) for all lanes of a group sharing (i, j).
__device__ __forceinline__ void accumulate_pair(double* U, double* g, bool mine, int lane, int leader, int i, int j,
                                                const double (&Ji)[12], const double (&Jj)[12], double r0, double r1) {
  const int NV = kNVMax;
#pragma unroll
  for (int a = 0; a < 6; a++) {
#pragma unroll
    for (int b = 0; b < 6; b++) {
      double v = mine ? (Ji[a] * Jj[b] + Ji[6 + a] * Jj[6 + b]) : 0.0;
      v = warp_sum(v);
      if (lane == leader) atomicAdd(&U[(6 * i + a) * NV + 6 * j + b], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 6; a++) {
#pragma unroll
    for (int b = a; b < 6; b++) {
      double v = mine ? (Jj[a] * Jj[b] + Jj[6 + a] * Jj[6 + b]) : 0.0;
      v = warp_sum(v);
      if (lane == leader) atomicAdd(&U[(6 * j + a) * NV + 6 * j + b], v);
    }
    double v = mine ? (Jj[a] * r0 + Jj[6 + a] * r1) : 0.0;
    v = warp_sum(v);
    if (lane == leader) atomicAdd(&g[6 * j + a], v);
  }
}

__global__ void __launch_bounds__(kLinThreads, 1) k_linearize(KP p, int w0) {
  const int w = w0 + blockIdx.x;
  WinState& st = p.st[w];
  if (!st.active || st.reuse) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LinShared& S = *reinterpret_cast<LinShared*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int F = p.F, NV = 6 * F;
  const double* pose = p.pose + (size_t)w * F * 7;
  build_frames(pose, p.ex + (size_t)w * 7, p.td[w], F, S.fr, &S.cam);
  for (int i = t; i < kNVMax * kNVMax; i += kLinThreads) S.U[i] = 0.0;
  if (t < kNVP) S.g[t] = 0.0;
  __syncthreads();

  const int nlm = p.nlm[w];
  const int32_t* start = p.start + (size_t)w * p.Lm; const int32_t* tlen = p.tlen + (size_t)w * p.Lm; const int32_t* obeg = p.obeg + (size_t)w * p.Lm;
  const float4* obs = p.obs + (size_t)w * p.Om;
  const double* ftd = p.frame_td + (size_t)w * F;
  const double mu = st.mu;
  const bool it0 = (st.iteration == 0);
  double cost_acc = 0.0, gmax = 0.0;
  // Schur accumulators: sym tiles (a <= b) of the 9x9 tile grid, tile index q -> warp q % 8
  double C[6][2];
#pragma unroll
  for (int q = 0; q < 6; q++) { C[q][0] = 0.0; C[q][1] = 0.0; }

  for (int base = 0; base < nlm; base += kLinThreads) {
    const int l = base + t;
    const bool have = l < nlm;
    // zero my column of WT
    for (int c = 0; c < kNVP; c++) S.WT[c * kWTStride + t] = 0.0;
    int i = 0, L = 0, ob = 0; bool fx = false; double lam = 1.0;
    LmCtx lc; float4 oi = make_float4(0, 0, 0, 0);
    if (have) {
      i = start[l]; L = tlen[l]; ob = obeg[l]; fx = p.fixed[(size_t)w * p.Lm + l] != 0; lam = p.invdep[(size_t)w * p.Lm + l];
      oi = obs[ob];
      landmark_ctx(S.fr[i], S.cam, oi, ftd[i], lam, lc);
    }
    double M[6] = {0, 0, 0, 0, 0, 0};  // sum Jx^T Jx (xx xy xz yy yz zz)
    double m3[3] = {0, 0, 0};          // sum Jx^T jl
    double n3[3] = {0, 0, 0};          // sum Jx^T r
    double v = 0.0, gl = 0.0;
    const int Lmax = __reduce_max_sync(0xffffffffu, L);
    for (int k = 1; k < Lmax; k++) {
      const bool valid = have && k < L;
      const int j = i + k;
      double Jx[6], Jj[12], Ji[12], r0 = 0, r1 = 0, jl0 = 0, jl1 = 0;
      if (valid) {
        V3 pcj; const float4 oj = obs[ob + k];
        obs_residual(S.fr[j], S.cam, lc, oj, ftd[j], p.sqrt_info_px, r0, r1, pcj);
        obs_jacobians(S.fr[j], S.cam, lc, pcj, p.sqrt_info_px, Jx, Jj);
        double hr, sc; huber(p.huber, r0 * r0 + r1 * r1, hr, sc);
        cost_acc += hr;
        r0 *= sc; r1 *= sc;
#pragma unroll
        for (int c = 0; c < 6; c++) Jx[c] *= sc;
#pragma unroll
        for (int c = 0; c < 12; c++) Jj[c] *= sc;
        // Ji = [Jx | Jx * Gi]
#pragma unroll
        for (int r = 0; r < 2; r++) {
          Ji[r * 6 + 0] = Jx[r * 3]; Ji[r * 6 + 1] = Jx[r * 3 + 1]; Ji[r * 6 + 2] = Jx[r * 3 + 2];
#pragma unroll
          for (int c = 0; c < 3; c++) Ji[r * 6 + 3 + c] = Jx[r * 3] * lc.Gi.m[c] + Jx[r * 3 + 1] * lc.Gi.m[3 + c] + Jx[r * 3 + 2] * lc.Gi.m[6 + c];
        }
        if (!fx) {
          jl0 = Jx[0] * lc.dXdl.x + Jx[1] * lc.dXdl.y + Jx[2] * lc.dXdl.z;
          jl1 = Jx[3] * lc.dXdl.x + Jx[4] * lc.dXdl.y + Jx[5] * lc.dXdl.z;
        }
        M[0] += Jx[0] * Jx[0] + Jx[3] * Jx[3]; M[1] += Jx[0] * Jx[1] + Jx[3] * Jx[4]; M[2] += Jx[0] * Jx[2] + Jx[3] * Jx[5];
        M[3] += Jx[1] * Jx[1] + Jx[4] * Jx[4]; M[4] += Jx[1] * Jx[2] + Jx[4] * Jx[5]; M[5] += Jx[2] * Jx[2] + Jx[5] * Jx[5];
#pragma unroll
        for (int c = 0; c < 3; c++) { m3[c] += Jx[c] * jl0 + Jx[3 + c] * jl1; n3[c] += Jx[c] * r0 + Jx[3 + c] * r1; }
        v += jl0 * jl0 + jl1 * jl1; gl += jl0 * r0 + jl1 * r1;
        // w_j = Jj^T jl  -> WT rows 6j..6j+5
#pragma unroll
        for (int c = 0; c < 6; c++) S.WT[(6 * j + c) * kWTStride + t] = Jj[c] * jl0 + Jj[6 + c] * jl1;
      } else {
#pragma unroll
        for (int c = 0; c < 12; c++) { Ji[c] = 0; Jj[c] = 0; }
      }
      // pose-block accumulation, grouped by (i, j) inside the warp
      const int key = valid ? (i * 16 + j) : -1;
      unsigned todo = __ballot_sync(0xffffffffu, valid);
      while (todo) {
        const int leader = __ffs(todo) - 1;
        const int k0 = __shfl_sync(0xffffffffu, key, leader);
        const bool mine = valid && key == k0;
        const unsigned grp = __ballot_sync(0xffffffffu, mine);
        accumulate_pair(S.U, S.g, mine, lane, leader, k0 >> 4, k0 & 15, Ji, Jj, r0, r1);
        todo &= ~grp;
      }
    }
    // host-frame block: U_ii += [M, M Gi; Gi^T M, Gi^T M Gi], g_i += [n3; Gi^T n3], w_i = [m3; Gi^T m3]
    {
      double Hii[21], gi[6], wi[6];
      if (have) {
        const double Mf[9] = {M[0], M[1], M[2], M[1], M[3], M[4], M[2], M[4], M[5]};
        double MG[9];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int c = 0; c < 3; c++) MG[r * 3 + c] = Mf[r * 3] * lc.Gi.m[c] + Mf[r * 3 + 1] * lc.Gi.m[3 + c] + Mf[r * 3 + 2] * lc.Gi.m[6 + c];
        int q = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
          for (int b = a; b < 6; b++) {
            double val;
            if (a < 3 && b < 3) val = Mf[a * 3 + b];
            else if (a < 3) val = MG[a * 3 + (b - 3)];
            else val = lc.Gi.m[(a - 3)] * MG[(b - 3)] + lc.Gi.m[3 + (a - 3)] * MG[3 + (b - 3)] + lc.Gi.m[6 + (a - 3)] * MG[6 + (b - 3)];
            Hii[q++] = val;
          }
#pragma unroll
        for (int c = 0; c < 3; c++) {
          gi[c] = n3[c]; wi[c] = m3[c];
          gi[3 + c] = lc.Gi.m[c] * n3[0] + lc.Gi.m[3 + c] * n3[1] + lc.Gi.m[6 + c] * n3[2];
          wi[3 + c] = lc.Gi.m[c] * m3[0] + lc.Gi.m[3 + c] * m3[1] + lc.Gi.m[6 + c] * m3[2];
        }
#pragma unroll
        for (int c = 0; c < 6; c++) S.WT[(6 * i + c) * kWTStride + t] = wi[c];
      } else {
#pragma unroll
        for (int q = 0; q < 21; q++) Hii[q] = 0;
#pragma unroll
        for (int c = 0; c < 6; c++) gi[c] = 0;
      }
      const int key = have ? i : -1;
      unsigned todo = __ballot_sync(0xffffffffu, have);
      while (todo) {
        const int leader = __ffs(todo) - 1;
        const int k0 = __shfl_sync(0xffffffffu, key, leader);
        const bool mine = have && key == k0;
        const unsigned grp = __ballot_sync(0xffffffffu, mine);
        int q = 0;
#pragma unroll
        for (int a = 0; a < 6; a++) {
#pragma unroll
          for (int b = a; b < 6; b++) { double s = warp_sum(mine ? Hii[q] : 0.0); q++; if (lane == leader) atomicAdd(&S.U[(6 * k0 + a) * kNVMax + 6 * k0 + b], s); }
          double s = warp_sum(mine ? gi[a] : 0.0);
          if (lane == leader) atomicAdd(&S.g[6 * k0 + a], s);
        }
        todo &= ~grp;
      }
    }
    // landmark scalars: jacobi scale (iteration 0), regularised v' = v + mu * e, 1/v'
    double inv = 0.0;
    if (have) {
      double s_l;
      if (it0) { s_l = 1.0 / (1.0 + sqrt(v)); p.lm_s[(size_t)w * p.Lm + l] = s_l; } else s_l = p.lm_s[(size_t)w * p.Lm + l];
      const double d2 = fmin(fmax(s_l * s_l * v, 1e-6), 1e32);
      const double e = d2 / (s_l * s_l);
      const double vp = v + mu * e;
      inv = (!fx && v > 0.0) ? 1.0 / vp : 0.0;
      p.lm_v[(size_t)w * p.Lm + l] = fx ? 0.0 : v;
      p.lm_g[(size_t)w * p.Lm + l] = fx ? 0.0 : gl;
      S.WT[66 * kWTStride + t] = fx ? 0.0 : gl;
      if (!fx) gmax = fmax(gmax, fabs(gl));
    }
    S.invv[t] = inv;
    __syncthreads();
    // Schur SYRK on the fp64 tensor cores: C[a-tile][b-tile] += sum_l (W[l][a] / v'_l) * W[l][b]
    {
      const int kr = lane & 3, mc = lane >> 2;
      int q = 0;
      for (int ta = 0; ta < 9; ta++)
        for (int tb = ta; tb < 9; tb++, q++) {
          if ((q & 7) != wid) continue;
          const int slot = q >> 3;
          double c0 = 0, c1 = 0;
          const double* wa = &S.WT[(8 * ta + mc) * kWTStride + kr];
          const double* wb = &S.WT[(8 * tb + mc) * kWTStride + kr];
#pragma unroll 4
          for (int k0 = 0; k0 < kLinThreads; k0 += 4) {
            const double a = wa[k0] * S.invv[k0 + kr];
            const double b = wb[k0];
            mma_f64(c0, c1, a, b);
          }
          // static indexing of the register array
#pragma unroll
          for (int s2 = 0; s2 < 6; s2++) if (s2 == slot) { C[s2][0] += c0; C[s2][1] += c1; }
        }
    }
    __syncthreads();
  }

  // S_vis = U - C (upper block triangle + tiles), g_vis = g - C[:,66]; mirror and write out
  {
    const int mc = lane >> 2, kr = lane & 3;
    int q = 0;
    for (int ta = 0; ta < 9; ta++)
      for (int tb = ta; tb < 9; tb++, q++) {
        if ((q & 7) != wid) continue;
        const int slot = q >> 3;
        double c0 = 0, c1 = 0;
#pragma unroll
        for (int s2 = 0; s2 < 6; s2++) if (s2 == slot) { c0 = C[s2][0]; c1 = C[s2][1]; }
        const int row = 8 * ta + mc, col = 8 * tb + 2 * kr;
        // stash the Schur tile into WT (free now) as a dense 72x72 matrix: reuse WT[row*72 + col]
        S.WT[row * kNVP + col] = c0; S.WT[row * kNVP + col + 1] = c1;
      }
  }
  __syncthreads();
  double* Svis = p.Svis + (size_t)w * kNVMax * kNVMax;
  for (int idx = t; idx < NV * NV; idx += kLinThreads) {
    const int r = idx / NV, c = idx % NV;
    const int a = r <= c ? r : c, b = r <= c ? c : r;  // upper element (a <= b)
    // U upper: blocks with frame(a) <= frame(b); inside diagonal blocks only a<=b stored -> (a,b) with a<=b always stored
    const double u = S.U[a * kNVMax + b];
    const int ta = a >> 3, tb = b >> 3;
    const double cs = (ta <= tb) ? S.WT[a * kNVP + b] : S.WT[b * kNVP + a];
    Svis[r * kNVMax + c] = u - cs;
  }
  if (t < NV) {
    p.gvis[(size_t)w * kNVP + t] = S.g[t] - S.WT[t * kNVP + 66];
    p.Udiag[(size_t)w * kNVMax + t] = S.U[t * kNVMax + t];
  }
  double red2[2] = {cost_acc, 0.0};
  block_sum<2>(red2, S.red);
  gmax = warp_max(gmax);
  if (lane == 0) S.red[wid] = gmax;
  __syncthreads();
  if (t == 0) {
    double gm = 0; for (int i2 = 0; i2 < kLinThreads / 32; i2++) gm = fmax(gm, S.red[i2]);
    st.cost_vis = red2[0]; st.gmax_l = gm;
  }
}

}  // namespace gf2
