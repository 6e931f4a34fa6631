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
constexpr int kSolveThreads = 256;
// k_linearize's output record of one window, contiguous so that the factor-sharded mode all-reduces ONE buffer per linearisation
// (SURVEY 8(e)): [Svis 66*66 | gvis 72 | gschur 72 | Udiag 66 | c_lin 4 | pad 6]; KP::Svis / gvis / gschur / Udiag / c_lin point at their
// field of window 0 and are indexed with the stride kVisRec.
constexpr int kVisRec = kNVMax * kNVMax + 2 * kNVP + kNVMax + 4 + 6;   // 4576 doubles = 36,608 B
static_assert(kVisRec % 2 == 0, "record keeps 16-byte alignment");
constexpr int kP = GF2_MAX_PRIOR_DIM;

struct WinState {  // per-window trust-region state (DoglegStrategy + TrustRegionMinimizer members)
  double radius, mu, x_cost, cand_cost, model_cost_change, dogleg_step_norm, x_norm2;
  double coef_a, coef_b;            // delta = -(a*u + b*z)
  double initial_cost, x_cost_prev;
  // x-part sums from k_solve, landmark-part sums from k_backsub
  double dlg2_x, gn2_x, gz_x, zEz_x, uEz_x, uSu, uEu_x, gmax_x;
  double dlg2_l, gn2_l, gz_l, zHz_l, uHz_l, uHu_l, gmax_l;
  double cand_nv, step2_x, xnorm2_x; // candidate: non-visual cost, |dx|^2 and |x|^2 of the frame states (identical on all ranks)
  double uSz, zSz;                  // u^T S' z and z^T S' z through the Cholesky factor (k_solve2)
  double cost_vis;                  // from k_linearize
  int32_t iteration, successful, termination, active, reuse, invalid_count, lin_valid, pad_;
  unsigned long long t_start_ns;    // %globaltimer at k_prepare: the solve's clock for max_solver_time_in_seconds
};

struct KP {  // kernel parameters (device pointers are window-major with the strides below)
  int nW, F, Lm, Om, Pm, D, use_wheel, Pr;  // Pr: row stride of the prior arrays (<= GF2_MAX_PRIOR_DIM)
  // Free wheel calibration blocks (estimate_wheel_extrinsic / _intrinsic / td_wheel, VE/estimator/estimator.cpp:3063-3118,3160):
  // wcal bit 0: body_T_wheel free, bit 1: sx sy sw free, bit 2: td_wheel free. They form one more block row "frame F" of the
  // reduced system with tangent layout [ex_wheel 6 | sx sy sw | td_wheel | 5 unused]. Ds = 15 (F + 1): row stride of sx/zx/ux/ex_diag.
  int wcal, Ds;
  uint32_t wsub;   // PoseSubsetParameterization of the wheel extrinsic: tangent components whose delta Plus zeroes (their columns stay in the system)
  uint32_t const_mask;
  double huber, sqrt_info_px, g_norm, lidar_sqrt_info;
  double ftol, gtol, ptol;
  int max_iterations;
  unsigned long long max_time_ns;   // ceres::Solver::Options::max_solver_time_in_seconds (0: no cap)
  // states (current / candidate)
  double *pose, *sb, *ex, *td, *exw, *sxw, *tdw, *invdep;
  double *pose_c, *sb_c, *invdep_c;
  double *exw_c, *sxw_c, *tdw_c;   // candidate wheel calibration (wcal != 0)
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
  double *imu_H, *imu_g;   // [nW][F-1][465] packed lower J^T J of each IMU factor, [nW][F-1][30] J^T r  (k_nonvis)
  double *prior_g, *cost_nv; // [nW][P] J0^T r, [nW] cost of the non-visual factors                  (k_nonvis)
  double* prior_H;    // [nW][P][P] = J0^T J0
  int32_t* prior_map; // [nW][P] column -> tangent index (or -1)
  // planes
  const int32_t* n_planes;
  const gf2_plane* planes;
  const double* plane_alpha;   // [nW][Pm] alpha_time of the ct == 1 planes (CTLidarPlaneNormFactor)
  // work
  double *Svis, *gvis, *gschur, *Udiag;  // fields of the per-window record of kVisRec doubles (see kVisRec)
  double *lm_v, *lm_g, *lm_s, *lm_z; // [nW][Lm]
  double *sx, *zx, *ux, *ex_diag;  // [nW][D] jacobi scale, GN step, u, e
  double *Sfull, *gfull;           // optional dump of the assembled reduced system [nW][D*D], [nW][D]
  int32_t *pperm, *ptask_first, *ptask_cnt, *ptask_frame, *nptasks;  // k_tasks: plane permutation by frame, warp tasks
  double *wheel_H, *wheel_g;       // [nW][F-1][3*36] pose blocks (i,i),(j,i),(j,j) of each wheel factor, [nW][F-1][12]  (k_nonvis)
  double *wheel_Hc, *wheel_gc;     // wcal: [nW][F-1][220] blocks (calib,pose_i) 10x6, (calib,pose_j) 10x6, (calib,calib) 10x10; [nW][F-1][10]
  int4* lminfo;                    // k_tasks: landmarks sorted by start frame, packed (index, track length, first observation, fixed)
  int32_t *task_first, *task_cnt, *task_start, *ntasks;  // k_tasks: warp tasks over that order
  // cross-rank scalars (factor-sharded mode all-reduces them; single GPU reads them straight back)
  double *c_lin, *c_gmax, *c_sums, *c_cand;  // record field {visual cost} (stride kVisRec), [nW] max|g_l|, [nW][8] k_backsub sums, [nW][4] {cand visual cost, |dl|^2, |l|^2}
  double* trace;                   // [nW][64][6]: candidate cost, model change, rho, radius, step norm, decision
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
  const double il = 1.0 / inv_dep;  // one reciprocal (fp64 division is a long software sequence); pts / inv_dep -> pts * il
  V3 pc = mk3(pts_i_td.x * il, pts_i_td.y * il, il);
  M3 ric; for (int i = 0; i < 9; i++) ric.m[i] = cam.ric[i];
  M3 Ri; for (int i = 0; i < 9; i++) Ri.m[i] = fi.R[i];
  V3 p_imu = mul(ric, pc) + mk3(cam.tic[0], cam.tic[1], cam.tic[2]);
  lc.Xw = mul(Ri, p_imu) + mk3(fi.P[0], fi.P[1], fi.P[2]);
  lc.Gi = scale(mul(Ri, skew(p_imu)), -1.0);
  lc.dXdl = mul(Ri, mul(ric, pts_i_td)) * (-(il * il));
}

// residual of one observation in frame j (projectionTwoFrameOneCamFactor.cpp:59-76), Huber weight applied by caller
__device__ __forceinline__ void obs_residual(const FrameCtx& fj, const CamCtx& cam, const LmCtx& lc, float4 oj, double td_j,
                                             double sqrt_info, double& r0, double& r1, V3& pcj) {
  const double dt = cam.td - td_j;
  V3 d = lc.Xw - mk3(fj.P[0], fj.P[1], fj.P[2]);
  pcj = mk3(fj.A[0] * d.x + fj.A[1] * d.y + fj.A[2] * d.z - cam.rtt[0], fj.A[3] * d.x + fj.A[4] * d.y + fj.A[5] * d.z - cam.rtt[1],
            fj.A[6] * d.x + fj.A[7] * d.y + fj.A[8] * d.z - cam.rtt[2]);
  const double ptx = (double)oj.x - dt * (double)oj.z, pty = (double)oj.y - dt * (double)oj.w;
  const double iz = 1.0 / pcj.z;  // one reciprocal per observation (fp64 division is a long software sequence)
  r0 = sqrt_info * (pcj.x * iz - ptx);
  r1 = sqrt_info * (pcj.y * iz - pty);
}

// ceres::HuberLoss + Corrector (rho'' <= 0 -> both r and J scaled by sqrt(rho')), restated in-tree at
// VE/factor/marginalization_factor.cpp:46-77. Returns rho(s)/... : cost contribution 0.5*rho0 and the scale.
__device__ __forceinline__ void huber(double delta, double sq, double& half_rho, double& scl) {
  const double b = delta * delta;
  if (sq > b) {
    // r = sqrt(sq) and scl = sqrt(delta / r) through two reciprocal square roots (fp64 sqrt and division are long software
    // sequences): r = sq * rsqrt(sq), delta / r = delta * rsqrt(sq), sqrt(y) = y * rsqrt(y)
    const double ir = rsqrt(sq), r = sq * ir, y = fmax(2.2250738585072014e-308, delta * ir);
    half_rho = 0.5 * (2.0 * delta * r - b); scl = y * rsqrt(y);
  } else { half_rho = 0.5 * sq; scl = 1.0; }
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
// sqrt_info = LLT(cov^-1).matrixL()^T (VE/factor/imu_factor.h:73, wheel_factor.h:85): Gauss-Jordan inverse with partial pivoting on the
// augmented [cov | I] followed by the Cholesky factor of the inverse, ONE WARP per factor: lane c owns column c of the N x 2N augmented
// matrix in shared memory. Every element goes through the same operations in the same order as a one-thread elimination would apply, so
// the result does not depend on the lane count (a single thread per factor took 0.45 ms per 15 x 15 factor: the longest kernel of a
// one-window solve).
template <int N>
__device__ void sqrt_info_from_cov_warp(const double* cov, double* out /* N*N row-major, upper triangular */, double* m /* shared scratch [N][2N] */) {
  const int lane = threadIdx.x & 31;
  constexpr int W = 2 * N;
  static_assert(W <= 32, "one lane per column of the augmented matrix");
  if (lane < W) for (int r = 0; r < N; r++) m[r * W + lane] = lane < N ? cov[r * N + lane] : (r == lane - N ? 1.0 : 0.0);
  __syncwarp();
  for (int col = 0; col < N; col++) {
    // pivot: the row >= col with the largest |m[r][col]| (first one on ties, as a sequential scan with `>` finds)
    double best = (lane >= col && lane < N) ? fabs(m[lane * W + col]) : -1.0; int piv = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int op = __shfl_xor_sync(0xffffffffu, piv, o);
      if (ob > best || (ob == best && op < piv)) { best = ob; piv = op; }
    }
    if (lane < W) {
      if (piv != col) { const double t = m[piv * W + lane]; m[piv * W + lane] = m[col * W + lane]; m[col * W + lane] = t; }
    }
    __syncwarp();
    const double d = m[col * W + col];
    __syncwarp();
    if (lane < W) m[col * W + lane] /= d;
    __syncwarp();
    double f[N];
#pragma unroll
    for (int r = 0; r < N; r++) f[r] = m[r * W + col];   // read before lane `col` rewrites its own column
    __syncwarp();
    if (lane < W) {
      const double pr = m[col * W + lane];
#pragma unroll
      for (int r = 0; r < N; r++) if (r != col && f[r] != 0.0) m[r * W + lane] -= f[r] * pr;
    }
    __syncwarp();
  }
  // Cholesky of the inverse (lower), in place in the right half: column j by lanes i = j .. N-1
  for (int j = 0; j < N; j++) {
    double dj = m[j * W + N + j];
    for (int k = 0; k < j; k++) dj -= m[j * W + N + k] * m[j * W + N + k];
    dj = sqrt(dj);
    __syncwarp();
    if (lane == j) m[j * W + N + j] = dj;
    else if (lane > j && lane < N) { double sacc = m[lane * W + N + j]; for (int k = 0; k < j; k++) sacc -= m[lane * W + N + k] * m[j * W + N + k]; m[lane * W + N + j] = sacc / dj; }
    __syncwarp();
  }
  if (lane < N) for (int r = 0; r < N; r++) out[r * N + lane] = (lane >= r) ? m[lane * W + N + r] : 0.0;
  __syncwarp();
}

__global__ void k_prepare(KP p, int w0) {
  const int w = w0 + blockIdx.x;
  const int F = p.F, t = threadIdx.x;
  extern __shared__ double sh[];
  // IMU / wheel sqrt_info: one warp per factor (warp q handles the factors q, q + nwarps, ...), 450 doubles of shared scratch per warp
  {
    const int wid = t >> 5, nwarp = blockDim.x >> 5;
    if (p.imu) for (int k = wid; k < F - 1; k += nwarp) {
      const gf2_imu_preint& rec = p.imu[(size_t)w * (F - 1) + k];
      if (rec.valid && rec.sum_dt <= 10.0) sqrt_info_from_cov_warp<15>(rec.covariance, p.imu_sqrt + ((size_t)w * (F - 1) + k) * 225, sh + wid * 450);
    }
    if (p.use_wheel && p.wheel) for (int k = wid; k < F - 1; k += nwarp) {
      const gf2_wheel_preint& rec = p.wheel[(size_t)w * (F - 1) + k];
      if (rec.valid && rec.sum_dt <= 10.0) sqrt_info_from_cov_warp<6>(rec.covariance, p.wheel_sqrt + ((size_t)w * (F - 1) + k) * 36, sh + wid * 450);
    }
  }
  // prior: column map and H = J0^T J0
  const int n = p.prior_rows ? p.prior_rows[w] : 0;
  if (n > 0) {
    int32_t* map = p.prior_map + (size_t)w * p.Pr;
    const gf2_prior_block* blk = p.prior_blocks + (size_t)w * (2 * F + 8);
    if (t == 0) {
      for (int c = 0; c < p.Pr; c++) map[c] = -1;
      for (int b = 0; b < p.prior_nblocks[w]; b++) {
        int base = -1, ls = 0;
        if (blk[b].kind == GF2_BLK_POSE) { base = 15 * blk[b].index; ls = 6; }
        else if (blk[b].kind == GF2_BLK_SPEEDBIAS) { base = 15 * blk[b].index + 6; ls = 9; }
        else {  // calibration blocks: a column only when the block is free in this solve (wheel calibration, block row F)
          ls = (blk[b].kind == GF2_BLK_EX_POSE || blk[b].kind == GF2_BLK_EX_WHEEL) ? 6 : 1;
          const int k = blk[b].kind;
          if (k == GF2_BLK_EX_WHEEL && (p.wcal & 1)) base = 15 * F;
          else if (k >= GF2_BLK_SX && k <= GF2_BLK_SW && (p.wcal & 2)) base = 15 * F + 6 + (k - GF2_BLK_SX);
          else if (k == GF2_BLK_TD_WHEEL && (p.wcal & 4)) base = 15 * F + 9;
        }
        for (int c = 0; c < ls; c++) map[blk[b].offset + c] = base < 0 ? -1 : base + c;
      }
    }
    const double* J0 = p.prior_J0 + (size_t)w * p.Pr * p.Pr;
    double* H = p.prior_H + (size_t)w * p.Pr * p.Pr;
    for (int idx = t; idx < n * n; idx += blockDim.x) {
      const int a = idx / n, b = idx % n;
      double s = 0; for (int r = 0; r < n; r++) s += J0[r * p.Pr + a] * J0[r * p.Pr + b];
      H[a * p.Pr + b] = s;
    }
  }
  if (t == 0) {
    WinState& s = p.st[w];
    s.radius = 1e4; s.mu = 1e-8; s.x_cost = 0; s.cand_cost = 0; s.model_cost_change = 0; s.dogleg_step_norm = 0;
    s.iteration = 0; s.successful = 0; s.termination = GF2_TERM_NO_CONVERGENCE; s.active = 1; s.reuse = 0; s.invalid_count = 0; s.lin_valid = 0;
    s.initial_cost = 0; s.coef_a = 0; s.coef_b = 0; s.x_norm2 = 0;
    unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    s.t_start_ns = now;
  }
}

}  // namespace gf2
