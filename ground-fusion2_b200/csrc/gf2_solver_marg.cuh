// gf2_solver_marg.cuh — marginalization after the solve (the prior of the next window).
//
// Replaces MarginalizationInfo::addResidualBlockInfo / preMarginalize / marginalize / getParameterBlocks
// (VE/factor/marginalization_factor.cpp:98-330) as driven by Estimator::optimization() (VE/estimator/estimator.cpp:3394-3595
// MARGIN_OLD, :3597-3690 MARGIN_SECOND_NEW). One CTA per window.
//
//   k_marg_build   evaluates the factors that touch the dropped blocks at the current states (old prior, IMU factor 0, the
//                  projection factors of the landmarks hosted in frame 0 with the Huber corrector of :46-77), assembles
//                  A = sum J^T J, b = sum J^T r in shared memory and eliminates the dropped blocks: the landmarks one at a time
//                  (each is a 1x1 diagonal block of Amm coupled to pose 0 only), then the frame-0 block by Cholesky.
//                  This equals Arr - Arm pinv(Amm) Amr of :285-291 whenever no eigenvalue of Amm is below eps = 1e-8, which is
//                  tested exactly (Amm - eps I positive definite <=> every v_l > eps and the eps-shifted 15x15 Schur
//                  complement has a Cholesky factor). Windows that fail the test are reported (GF2_MARG_DEGENERATE) and
//                  keep their old prior — see DESIGN.md.
//   k_marg_eig     compacts the kept system to the blocks the factors touched, runs a cyclic Jacobi eigensolver in shared
//                  memory (round-robin parallel ordering, n/2 disjoint rotations per step) and writes
//                  linearized_jacobians = sqrt(S) V^T, linearized_residuals = sqrt(1/S) V^T b (:293-303, eps truncation)
//                  plus the kept blocks renamed by addr_shift (:3561-3595 / :3653-3681) straight into the solver's prior
//                  storage, so the next gf2_solve uses it without a host round trip.
//
// Block order: the reference iterates an unordered_map keyed by address, so its order is unspecified; here kept blocks are
// ordered pose 0..F-2, speed-bias 0, ex-pose, td (new-window indices). The prior as a function of the states is the same.
#pragma once
#include "gf2_solver_kernels2.cuh"

namespace gf2 {

constexpr int kMargM = 15;                              // frame-0 part of the dropped set: pose 6 + speed-bias 9
constexpr int kMargKMax = 6 * (kMaxF - 1) + 9 + 6 + 1 + 10;  // kept tangent dims, canonical: poses, sb0, ex-pose, td, wheel calibration (86)
constexpr int kMargTMax = kMargM + kMargKMax;           // 101
constexpr int kMargLD = kMargTMax | 1;                  // odd leading dimension: conflict-free column walks
constexpr int kMargThreads = 256;
constexpr int kEigThreads = 512;
constexpr int kMargBlocksMax = kMaxF + 8;               // kept blocks: F-1 poses, sb0, ex, td, ex-wheel, sx, sy, sw, td-wheel
constexpr int kMargChunk = 32;                          // landmarks eliminated per rank-k update
constexpr int kMargNCMax = 13 + 6 * (kMaxF - 1);        // columns a frame-0 landmark can touch: pose0, poses 1..F-1, ex, td  (73)
constexpr int kMargOwn = (kMargNCMax * (kMargNCMax + 1) / 2 + kMargNCMax + kMargThreads - 1) / kMargThreads;  // matrix entries per thread (11)
constexpr double kMargEps = 1e-8;                       // MarginalizationInfo::eps, VE/factor/marginalization_factor.h:83
constexpr int kMargStageSum = kMargTMax * kMargTMax + kMargTMax + 36 + 1;   // A (T x T, stride kMargTMax) | b | C6 | n_lm0
constexpr int kMargStageMax = kMargBlocksMax + 2 + 1 + 1;                    // touched | mtouched | bad | record written

enum { GF2_MARG_OK_ = 0, GF2_MARG_INVALID_ = -1, GF2_MARG_UNCHANGED_ = -2, GF2_MARG_UNSUPPORTED_ = -3, GF2_MARG_DEGENERATE_ = -4, GF2_MARG_TOO_LARGE_ = -5 };

struct MargP {
  int mode;            // 0 MARGIN_OLD, 1 MARGIN_SECOND_NEW
  int eig;             // 0: rank-revealing Cholesky factor of the kept system (default), 1: the reference's eigen-decomposition, literally
  // factor-sharded mode: the frame-0 landmarks of a window live on different ranks. k_marg_build stops after the landmark
  // elimination and leaves the window's system [A | b | C6 | n_lm0] (SUM over the ranks) and flags (MAX) in the staging buffers; the ranks
  // all-reduce them and k_marg_finish eliminates the frame block on the sums. The prior / IMU / wheel factors are replicated: rank 0 alone
  // contributes their values (rank > 0 zeroes the system before its landmarks), every rank their flags.
  int nranks, rank;
  double* stage_sum;   // [nW][kMargStageSum]
  double* stage_max;   // [nW][kMargStageMax]
  double* A;           // [nW][kMargKMax][kMargKMax] kept system after elimination
  double* b;           // [nW][kMargKMax]
  int32_t* touched;    // [nW][kMargBlocksMax] kept block touched by a factor
  int32_t* status;     // [nW]
  int32_t* mdim;       // [nW] m (dropped tangent dims) for reporting
  // outputs of k_marg_eig (may alias the solver's prior storage)
  int32_t *out_rows, *out_nblocks;
  double *out_J0, *out_r0;
  gf2_prior_block* out_blocks;
};

struct MargShared {
  FrameCtx fr[kMaxF];
  CamCtx cam;
  double b[kMargTMax];
  double W[kMargChunk][kMargNCMax + 2];   // scaled landmark rows of one chunk (+ g_l column); odd-ish stride
  int16_t lm0[GF2_MAX_LANDMARKS];         // landmarks hosted in frame 0
  double Z[kMargM][kMargKMax + 1];
  double Y[kMargM][kMargM + 1];
  double C6[6][6];
  double dx[kP], pr[kP];
  int cmap[kP];
  int touched[kMargBlocksMax];
  int mtouched[2];   // pose0 / sb0 (MARGIN_OLD) present in the problem
  int n_lm0, bad, maxlen;
};

// kept-layout column of (kind, new index); -1 if the block cannot be kept in this build
__device__ __forceinline__ int marg_kept_col(int F, int kind, int index) {
  const int NP = 6 * (F - 1);
  switch (kind) {
    case GF2_BLK_POSE: return (index >= 0 && index < F - 1) ? 6 * index : -1;
    case GF2_BLK_SPEEDBIAS: return index == 0 ? NP : -1;
    case GF2_BLK_EX_POSE: return NP + 9;
    case GF2_BLK_TD: return NP + 15;
    case GF2_BLK_EX_WHEEL: return NP + 16;
    case GF2_BLK_SX: return NP + 22;
    case GF2_BLK_SY: return NP + 23;
    case GF2_BLK_SW: return NP + 24;
    case GF2_BLK_TD_WHEEL: return NP + 25;
  }
  return -1;
}
__device__ __forceinline__ int marg_kept_block(int F, int kind, int index) {  // slot in touched[]
  switch (kind) {
    case GF2_BLK_POSE: return index;
    case GF2_BLK_SPEEDBIAS: return F - 1;
    case GF2_BLK_EX_POSE: return F;
    case GF2_BLK_TD: return F + 1;
    case GF2_BLK_EX_WHEEL: return F + 2;
    case GF2_BLK_SX: return F + 3;
    case GF2_BLK_SY: return F + 4;
    case GF2_BLK_SW: return F + 5;
    case GF2_BLK_TD_WHEEL: return F + 6;
  }
  return -1;
}
// kept block b (slot of touched[]) -> kind, local size, canonical column
__device__ __forceinline__ void marg_block_info(int F, int b, int& kind, int& ls, int& c0) {
  const int NP = 6 * (F - 1);
  if (b < F - 1) { kind = GF2_BLK_POSE; ls = 6; c0 = 6 * b; }
  else if (b == F - 1) { kind = GF2_BLK_SPEEDBIAS; ls = 9; c0 = NP; }
  else if (b == F) { kind = GF2_BLK_EX_POSE; ls = 6; c0 = NP + 9; }
  else if (b == F + 1) { kind = GF2_BLK_TD; ls = 1; c0 = NP + 15; }
  else if (b == F + 2) { kind = GF2_BLK_EX_WHEEL; ls = 6; c0 = NP + 16; }
  else { kind = GF2_BLK_SX + (b - (F + 3)); ls = 1; c0 = NP + 22 + (b - (F + 3)); }
}

// Jacobians of one projection factor wrt the camera extrinsic (2x6) and td (2x1),
// VE/factor/projectionTwoFrameOneCamFactor.cpp:125-146, expressed with the quantities obs_jacobians already has:
//   reduce * ric^T (Rj^T Ri - I)                 = Jx Ri - B
//   reduce * (-tmp_r [pc_i]x + [pc_j]x)          with reduce * tmp_r = Jx Ri ric =: Jc
//   td: -Jc velocity_i / inv_dep + sqrt_info * velocity_j
__device__ __forceinline__ void obs_jacobians_calib(const FrameCtx& fi, const CamCtx& cam, V3 pci, V3 pcj, double sqrt_info, const double (&Jx)[6],
                                                    float4 oi, float4 oj, double inv_dep, double (&Jex)[12], double (&Jtd)[2]) {
  const double iz = 1.0 / pcj.z;
  const double a = sqrt_info * iz, bx = -sqrt_info * pcj.x * iz * iz, by = -sqrt_info * pcj.y * iz * iz;
  double B[6];
#pragma unroll
  for (int c = 0; c < 3; c++) { B[c] = a * cam.ric[c * 3 + 0] + bx * cam.ric[c * 3 + 2]; B[3 + c] = a * cam.ric[c * 3 + 1] + by * cam.ric[c * 3 + 2]; }
  double JR[6], Jc[6];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) JR[r * 3 + c] = Jx[r * 3] * fi.R[c] + Jx[r * 3 + 1] * fi.R[3 + c] + Jx[r * 3 + 2] * fi.R[6 + c];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) Jc[r * 3 + c] = JR[r * 3] * cam.ric[c] + JR[r * 3 + 1] * cam.ric[3 + c] + JR[r * 3 + 2] * cam.ric[6 + c];
  const double red[6] = {a, 0.0, bx, 0.0, a, by};
#pragma unroll
  for (int r = 0; r < 2; r++) {
    Jex[r * 6 + 0] = JR[r * 3 + 0] - B[r * 3 + 0]; Jex[r * 6 + 1] = JR[r * 3 + 1] - B[r * 3 + 1]; Jex[r * 6 + 2] = JR[r * 3 + 2] - B[r * 3 + 2];
    // row * [p]x = (row_1 p_z - row_2 p_y, row_2 p_x - row_0 p_z, row_0 p_y - row_1 p_x)
    const double c0 = Jc[r * 3], c1 = Jc[r * 3 + 1], c2 = Jc[r * 3 + 2];
    const double d0 = red[r * 3], d1 = red[r * 3 + 1], d2 = red[r * 3 + 2];
    Jex[r * 6 + 3] = -(c1 * pci.z - c2 * pci.y) + (d1 * pcj.z - d2 * pcj.y);
    Jex[r * 6 + 4] = -(c2 * pci.x - c0 * pci.z) + (d2 * pcj.x - d0 * pcj.z);
    Jex[r * 6 + 5] = -(c0 * pci.y - c1 * pci.x) + (d0 * pcj.y - d1 * pcj.x);
  }
  Jtd[0] = -(Jc[0] * (double)oi.z + Jc[1] * (double)oi.w) / inv_dep + sqrt_info * (double)oj.z;
  Jtd[1] = -(Jc[3] * (double)oi.z + Jc[4] * (double)oi.w) / inv_dep + sqrt_info * (double)oj.w;
}

// Second half of the marginalization of one window: the system [A | b] of the dropped frame block and the kept blocks is complete in
// shared memory (upper triangle); symmetrise, eliminate the frame block, write the kept system for k_marg_eig.
__device__ __forceinline__ void marg_finish_block(const KP& p, int w, const MargP& mp, MargShared& s, double* A, int T, int K) {
  const int t = threadIdx.x;
  // ---- symmetrise, then eliminate the frame-0 block (MARGIN_OLD: pose0 + sb0, MARGIN_SECOND_NEW: pose F-2)
  for (int e = t; e < T * T; e += blockDim.x) { const int a = e / T, c = e % T; if (a > c) A[a * kMargLD + c] = A[c * kMargLD + a]; }
  __syncthreads();
  const int md = (mp.mode == 0 ? (s.mtouched[0] ? 6 : 0) + (s.mtouched[1] ? 9 : 0) : 6);
  if (t == 0) {
    int status = GF2_MARG_OK_;
    if (s.bad == 1) status = GF2_MARG_UNSUPPORTED_;
    else if (s.bad == 2) status = GF2_MARG_DEGENERATE_;
    else if (md + s.n_lm0 == 0) status = GF2_MARG_INVALID_;
    else {
      // Cholesky of Y (the dropped frame block after the landmarks) and of its eps-shifted twin
      for (int pass = 0; pass < 2 && status == GF2_MARG_OK_; pass++) {
        for (int i = 0; i < kMargM; i++) for (int j = 0; j <= i; j++) {
          const bool pi = mp.mode == 1 ? i < 6 : (i < 6 ? s.mtouched[0] : s.mtouched[1]) != 0;
          const bool pj = mp.mode == 1 ? j < 6 : (j < 6 ? s.mtouched[0] : s.mtouched[1]) != 0;
          double a = (pi && pj) ? A[i * kMargLD + j] : (i == j ? 1.0 : 0.0);   // absent dims: identity pivot, zero coupling
          if (pass == 0 && pi && pj) { if (i == j) a -= kMargEps; if (i < 6) a -= s.C6[j][i]; }
          s.Y[i][j] = a;
        }
        for (int j = 0; j < kMargM && status == GF2_MARG_OK_; j++) {
          double d = s.Y[j][j]; for (int k = 0; k < j; k++) d -= s.Y[j][k] * s.Y[j][k];
          if (!(d > 0.0)) { status = GF2_MARG_DEGENERATE_; break; }
          d = sqrt(d); s.Y[j][j] = d;
          for (int i = j + 1; i < kMargM; i++) { double v = s.Y[i][j]; for (int k = 0; k < j; k++) v -= s.Y[i][k] * s.Y[j][k]; s.Y[i][j] = v / d; }
        }
      }
    }
    mp.status[w] = status; mp.mdim[w] = md + s.n_lm0;
    s.bad = status;
  }
  __syncthreads();
  if (s.bad != GF2_MARG_OK_) return;
  // Z = L^-1 [X | b_m], X = A[0:15, 15:T]; absent dropped dims have zero rows in X and an identity pivot
  if (t <= K) {
    double z[kMargM];
    for (int i = 0; i < kMargM; i++) {
      double v = (t < K) ? A[i * kMargLD + kMargM + t] : s.b[i];
      for (int k = 0; k < i; k++) v -= s.Y[i][k] * z[k];
      z[i] = v / s.Y[i][i];
    }
    for (int i = 0; i < kMargM; i++) s.Z[i][t] = z[i];
  }
  __syncthreads();
  double* Ao = mp.A + (size_t)w * kMargKMax * kMargKMax;
  for (int e = t; e < K * K; e += blockDim.x) {
    const int a = e / K, c = e % K;
    double v = A[(kMargM + a) * kMargLD + kMargM + c];
    for (int k = 0; k < kMargM; k++) v -= s.Z[k][a] * s.Z[k][c];
    Ao[a * kMargKMax + c] = v;
  }
  if (t < K) { double v = s.b[kMargM + t]; for (int k = 0; k < kMargM; k++) v -= s.Z[k][t] * s.Z[k][K]; mp.b[(size_t)w * kMargKMax + t] = v; }
  if (t < kMargBlocksMax) mp.touched[(size_t)w * kMargBlocksMax + t] = s.touched[t];
}

__global__ void __launch_bounds__(kMargThreads) k_marg_build(KP p, int w0, MargP mp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MargShared& s = *reinterpret_cast<MargShared*>(smem_raw);
  double* A = reinterpret_cast<double*>(smem_raw + ((sizeof(MargShared) + 15) & ~size_t(15)));  // [kMargTMax][kMargLD], upper triangle accumulated
  const int w = w0 + blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5, nw = blockDim.x >> 5;
  const int F = p.F, NP = 6 * (F - 1), K = NP + 16 + (p.use_wheel ? 10 : 0), T = kMargM + K;
  const double* pose = p.pose + (size_t)w * F * 7;
  const double* sb = p.sb + (size_t)w * F * 9;
  const int sn = F - 2;  // second-newest frame (WINDOW_SIZE - 1)

  for (int i = t; i < kMargTMax * kMargLD; i += blockDim.x) A[i] = 0.0;
  for (int i = t; i < kMargTMax; i += blockDim.x) s.b[i] = 0.0;
  if (t < kMargBlocksMax) s.touched[t] = 0;
  if (t < 36) (&s.C6[0][0])[t] = 0.0;
  if (t == 0) { s.mtouched[0] = s.mtouched[1] = 0; s.n_lm0 = 0; s.bad = 0; s.maxlen = 0; }
  build_frames(pose, p.ex + (size_t)w * 7, p.td[w], F, s.fr, &s.cam);
  __syncthreads();

  // ---- old prior: r = r0 + J0 dx at the current states; A += J0^T J0, b += J0^T r in the marginalization layout
  const int n0 = p.prior_rows ? p.prior_rows[w] : 0;
  if (n0 > 0) {
    const gf2_prior_block* blk = p.prior_blocks + (size_t)w * (2 * F + 8);
    const int nb = p.prior_nblocks[w];
    if (t < nb) {
      const gf2_prior_block& b = blk[t];
      prior_block_dx(b, pose, sb, calib_of(p, w, false), s.dx);
      int col = -1, ls = (b.kind == GF2_BLK_POSE || b.kind == GF2_BLK_EX_POSE || b.kind == GF2_BLK_EX_WHEEL) ? 6 : (b.kind == GF2_BLK_SPEEDBIAS ? 9 : 1);
      if (mp.mode == 0) {
        if (b.kind == GF2_BLK_POSE && b.index == 0) { col = 0; s.mtouched[0] = 1; }
        else if (b.kind == GF2_BLK_SPEEDBIAS && b.index == 0) { col = 6; s.mtouched[1] = 1; }
        else {
          const int ni = (b.kind == GF2_BLK_POSE || b.kind == GF2_BLK_SPEEDBIAS) ? b.index - 1 : b.index;
          const int kc = marg_kept_col(F, b.kind, ni);
          if (kc >= 0) { col = kMargM + kc; s.touched[marg_kept_block(F, b.kind, ni)] = 1; }
        }
      } else {
        if (b.kind == GF2_BLK_POSE && b.index == sn) { col = 0; s.mtouched[0] = 1; }
        else {
          const int ni = (b.kind == GF2_BLK_POSE || b.kind == GF2_BLK_SPEEDBIAS) ? (b.index == F - 1 ? F - 2 : b.index) : b.index;
          const int kc = (b.kind == GF2_BLK_SPEEDBIAS && b.index == sn) ? -1 : marg_kept_col(F, b.kind, ni);
          if (kc >= 0) { col = kMargM + kc; s.touched[marg_kept_block(F, b.kind, ni)] = 1; }
        }
      }
      if (col < 0) s.bad = 1;
      for (int k = 0; k < ls; k++) s.cmap[b.offset + k] = col < 0 ? 0 : col + k;
    }
    __syncthreads();
    const double* J0 = p.prior_J0 + (size_t)w * p.Pr * p.Pr;
    const double* r0 = p.prior_r0 + (size_t)w * p.Pr;
    if (t < n0) { double v = r0[t]; for (int c = 0; c < n0; c++) v += J0[t * p.Pr + c] * s.dx[c]; s.pr[t] = v; }
    __syncthreads();
    const double* H = p.prior_H + (size_t)w * p.Pr * p.Pr;  // J0^T J0 (k_prepare)
    for (int e = t; e < n0 * n0; e += blockDim.x) {
      const int a = e / n0, c = e % n0;
      const int ia = s.cmap[a], ic = s.cmap[c];
      if (ia <= ic) A[ia * kMargLD + ic] += H[a * p.Pr + c];   // cmap is injective: no two (a, c) share a target
    }
    if (t < n0) { double v = 0; for (int r = 0; r < n0; r++) v += J0[r * p.Pr + t] * s.pr[r]; s.b[s.cmap[t]] += v; }
  }
  __syncthreads();
  if (mp.mode == 1 && !(n0 > 0 && s.mtouched[0])) {  // estimator.cpp:3599-3600: nothing to do, the old prior stays
    if (t == 0) { mp.status[w] = GF2_MARG_UNCHANGED_; mp.mdim[w] = 0; }
    if (mp.nranks > 1) {   // (the same decision on every rank: the prior is replicated) no record for k_marg_finish
      for (int i = t; i < kMargStageSum; i += blockDim.x) mp.stage_sum[(size_t)w * kMargStageSum + i] = 0.0;
      for (int i = t; i < kMargStageMax; i += blockDim.x) mp.stage_max[(size_t)w * kMargStageMax + i] = 0.0;
    }
    return;
  }
  if (mp.mode == 1 && mp.nranks > 1 && mp.rank > 0) {   // MARGIN_SECOND_NEW has replicated factors only: rank 0's values, everybody's flags
    for (int i = t; i < kMargTMax * kMargLD; i += blockDim.x) A[i] = 0.0;
    for (int i = t; i < kMargTMax; i += blockDim.x) s.b[i] = 0.0;
    __syncthreads();
  }

  if (mp.mode == 0) {
    // ---- IMU factor 0 (estimator.cpp:3416-3427): blocks pose0, sb0 dropped; pose1, sb1 kept as pose 0 / sb 0 of the next window
    if (wid == 0 && p.imu) {
      const gf2_imu_preint& pre = p.imu[(size_t)w * (F - 1)];
      if (pre.valid && pre.sum_dt < 10.0) {
        double* J = &s.Z[0][0];   // 15x30 + 15 scratch: Z is free until the elimination
        double* r = J + 450;
        static_assert(sizeof(s.Z) >= sizeof(double) * 465, "scratch");
        if (lane == 0) { ImuStates s2 = load_imu_states(pose, sb, 0); imu_raw(pre, s2, p.g_norm, r, J); }
        __syncwarp();
        const double* sq = p.imu_sqrt + ((size_t)w * (F - 1)) * 225;
        if (lane < 30) { for (int a = 0; a < 15; a++) { double acc = 0; for (int kk = a; kk < 15; kk++) acc += sq[a * 15 + kk] * J[kk * 30 + lane]; J[a * 30 + lane] = acc; } }
        else if (lane == 30) { for (int a = 0; a < 15; a++) { double acc = 0; for (int kk = a; kk < 15; kk++) acc += sq[a * 15 + kk] * r[kk]; r[a] = acc; } }
        __syncwarp();
        // factor columns [pose_i 6 | sb_i 9 | pose_j 6 | sb_j 9] -> layout columns
        for (int e = lane; e < 900; e += 32) {
          const int a = e / 30, c = e % 30;
          const int ia = a < 15 ? a : (a < 21 ? kMargM + (a - 15) : kMargM + NP + (a - 21));
          const int ic = c < 15 ? c : (c < 21 ? kMargM + (c - 15) : kMargM + NP + (c - 21));
          if (ia <= ic) { double acc = 0; for (int rr = 0; rr < 15; rr++) acc += J[rr * 30 + a] * J[rr * 30 + c]; A[ia * kMargLD + ic] += acc; }
        }
        if (lane < 30) {
          const int a = lane; const int ia = a < 15 ? a : (a < 21 ? kMargM + (a - 15) : kMargM + NP + (a - 21));
          double acc = 0; for (int rr = 0; rr < 15; rr++) acc += J[rr * 30 + a] * r[rr];
          s.b[ia] += acc;
        }
        if (lane == 0) { s.mtouched[0] = s.mtouched[1] = 1; s.touched[0] = 1; s.touched[F - 1] = 1; }
      }
    }
    __syncthreads();

    // ---- wheel factor 0 (estimator.cpp:3428-3439): pose0 dropped; pose1, body_T_wheel, sx, sy, sw, td_wheel kept
    if (wid == 0 && p.wheel) {
      const gf2_wheel_preint& pre = p.wheel[(size_t)w * (F - 1)];
      if (pre.valid && pre.sum_dt < 10.0) {
        double* J = &s.Z[0][0];   // 6 x 22 [pose_i 6 | pose_j 6 | ex_wheel 6 | sx sy sw td] + 6 residuals
        double* r = J + 132;
        if (lane == 0) {
          double Jp[72], Jc[60];
          wheel_raw(pre, pose, pose + 7, p.exw + (size_t)w * 7, p.sxw + (size_t)w * 3, p.tdw[w], r, Jp);
          wheel_calib_jacobians(pre, pose, pose + 7, p.exw + (size_t)w * 7, p.sxw + (size_t)w * 3, p.tdw[w], Jc);
          for (int a = 0; a < 6; a++) { for (int c = 0; c < 12; c++) J[a * 22 + c] = Jp[a * 12 + c]; for (int c = 0; c < 10; c++) J[a * 22 + 12 + c] = Jc[a * 10 + c]; }
        }
        __syncwarp();
        const double* sq = p.wheel_sqrt + ((size_t)w * (F - 1)) * 36;
        if (lane < 22) { for (int a = 0; a < 6; a++) { double acc = 0; for (int kk = a; kk < 6; kk++) acc += sq[a * 6 + kk] * J[kk * 22 + lane]; J[a * 22 + lane] = acc; } }
        else if (lane == 22) { for (int a = 0; a < 6; a++) { double acc = 0; for (int kk = a; kk < 6; kk++) acc += sq[a * 6 + kk] * r[kk]; r[a] = acc; } }
        __syncwarp();
        for (int e = lane; e < 484; e += 32) {
          const int a = e / 22, c = e % 22;
          const int ia = a < 6 ? a : (a < 12 ? kMargM + (a - 6) : kMargM + NP + 16 + (a - 12));
          const int ic = c < 6 ? c : (c < 12 ? kMargM + (c - 6) : kMargM + NP + 16 + (c - 12));
          if (ia <= ic) { double acc = 0; for (int rr = 0; rr < 6; rr++) acc += J[rr * 22 + a] * J[rr * 22 + c]; A[ia * kMargLD + ic] += acc; }
        }
        if (lane < 22) {
          const int a = lane; const int ia = a < 6 ? a : (a < 12 ? kMargM + (a - 6) : kMargM + NP + 16 + (a - 12));
          double acc = 0; for (int rr = 0; rr < 6; rr++) acc += J[rr * 22 + a] * r[rr];
          s.b[ia] += acc;
        }
        if (lane == 0) { s.mtouched[0] = 1; s.touched[0] = 1; for (int b2 = F + 2; b2 <= F + 6; b2++) s.touched[b2] = 1; }
      }
    }
    __syncthreads();

    if (mp.nranks > 1 && mp.rank > 0) {   // the replicated factors count once (rank 0); their flags stay
      for (int i = t; i < kMargTMax * kMargLD; i += blockDim.x) A[i] = 0.0;
      for (int i = t; i < kMargTMax; i += blockDim.x) s.b[i] = 0.0;
      __syncthreads();
    }
    // ---- projection factors of the landmarks hosted in frame 0 (estimator.cpp:3495-3528). One warp per landmark, one lane
    //      per observation. Each landmark is a 1x1 block of Amm coupled to pose 0 only, so it is eliminated on the spot:
    //      its scaled row w / sqrt(v) over the touched columns (and g_l / sqrt(v)) is staged in shared memory, and after
    //      every chunk of 32 landmarks each thread subtracts the rank-32 update from the matrix entries it owns — no atomics.
    const int nl = p.nlm[w];
    const int32_t* start = p.start + (size_t)w * p.Lm;
    const int32_t* tlen = p.tlen + (size_t)w * p.Lm;
    const int32_t* obeg = p.obeg + (size_t)w * p.Lm;
    const float4* obs = p.obs + (size_t)w * p.Om;
    const double* ftd = p.frame_td + (size_t)w * F;
    const int cEX = kMargM + NP + 9, cTD = kMargM + NP + 15;
    const int NC = 13 + NP;                      // compact columns of a landmark row: pose0 6 | poses 1..F-1 | ex 6 | td 1
    // frame-0 landmark list (deterministic order)
    if (wid == 0) {
      int cnt = 0;
      for (int base = 0; base < nl; base += 32) {
        const int l = base + lane;
        const bool is0 = l < nl && start[l] == 0;
        const unsigned bal = __ballot_sync(0xffffffffu, is0);
        if (is0) s.lm0[cnt + __popc(bal & ((1u << lane) - 1u))] = (int16_t)l;
        cnt += __popc(bal);
      }
      if (lane == 0) s.n_lm0 = cnt;
    }
    __syncthreads();
    const int n_lm0 = s.n_lm0;
    // entries of the compact upper triangle (+ the b column) owned by this thread
    const int NE = NC * (NC + 1) / 2 + NC;
    int ea[kMargOwn], ec[kMargOwn]; double eacc[kMargOwn];
#pragma unroll
    for (int i = 0; i < kMargOwn; i++) {
      const int e = t + i * kMargThreads;
      eacc[i] = 0.0; ea[i] = -1; ec[i] = 0;
      if (e < NE) {
        if (e >= NE - NC) { ea[i] = e - (NE - NC); ec[i] = NC; }         // b column: pairs (a, NC)
        else { int a = 0, r = e; while (r >= NC - a) { r -= NC - a; a++; } ea[i] = a; ec[i] = a + r; }
      }
    }
    double cacc[3] = {0.0, 0.0, 0.0}, cb = 0.0;   // per-warp accumulators of the common 13x13 block (lane e & 31 owns entry e) and b
    int maxlen = 0;
    for (int base = 0; base < n_lm0; base += kMargChunk) {
      for (int rr4 = 0; rr4 < kMargChunk / (kMargThreads / 32); rr4++) {
        const int row = wid * (kMargChunk / (kMargThreads / 32)) + rr4;
        double* wv = s.W[row];
        for (int c = lane; c <= NC; c += 32) wv[c] = 0.0;
        __syncwarp();
        if (base + row >= n_lm0) continue;
        const int l = s.lm0[base + row];
        const int len = tlen[l];
        maxlen = max(maxlen, len);
        const float4 oi = obs[obeg[l]];
        const double lam = p.invdep[(size_t)w * p.Lm + l];
        LmCtx lc; landmark_ctx(s.fr[0], s.cam, oi, ftd[0], lam, lc);
        const double dti = s.cam.td - ftd[0];
        const V3 pci = mk3(((double)oi.x - dti * (double)oi.z) / lam, ((double)oi.y - dti * (double)oi.w) / lam, 1.0 / lam);
        const int k = lane + 1;           // this lane's observation: frame k
        const bool on = k < len;
        double Jc13[2][13];               // common columns [pose0 6 | ex 6 | td 1]
        double Jj[12], Jl[2], r[2];
#pragma unroll
        for (int i = 0; i < 13; i++) Jc13[0][i] = Jc13[1][i] = 0.0;
#pragma unroll
        for (int i = 0; i < 12; i++) Jj[i] = 0.0;
        Jl[0] = Jl[1] = r[0] = r[1] = 0.0;
        if (on) {
          const float4 oj = obs[obeg[l] + k];
          V3 pcj; double r0v, r1v;
          obs_residual(s.fr[k], s.cam, lc, oj, ftd[k], p.sqrt_info_px, r0v, r1v, pcj);
          double Jx[6];
          obs_jacobians(s.fr[k], s.cam, lc, pcj, p.sqrt_info_px, Jx, Jj);
          double Jex[12], Jtd[2];
          obs_jacobians_calib(s.fr[0], s.cam, pci, pcj, p.sqrt_info_px, Jx, oi, oj, lam, Jex, Jtd);
          double half_rho, scl; huber(p.huber, r0v * r0v + r1v * r1v, half_rho, scl);
#pragma unroll
          for (int rr = 0; rr < 2; rr++) {
            const double jx0 = Jx[rr * 3], jx1 = Jx[rr * 3 + 1], jx2 = Jx[rr * 3 + 2];
            Jc13[rr][0] = scl * jx0; Jc13[rr][1] = scl * jx1; Jc13[rr][2] = scl * jx2;
#pragma unroll
            for (int c = 0; c < 3; c++) Jc13[rr][3 + c] = scl * (jx0 * lc.Gi.m[c] + jx1 * lc.Gi.m[3 + c] + jx2 * lc.Gi.m[6 + c]);
#pragma unroll
            for (int c = 0; c < 6; c++) { Jc13[rr][6 + c] = scl * Jex[rr * 6 + c]; Jj[rr * 6 + c] *= scl; }
            Jc13[rr][12] = scl * Jtd[rr];
            Jl[rr] = scl * (jx0 * lc.dXdl.x + jx1 * lc.dXdl.y + jx2 * lc.dXdl.z);
          }
          r[0] = scl * r0v; r[1] = scl * r1v;
        }
        // landmark row: v, g_l, w over the touched columns, scaled by 1/sqrt(v)
        const double v = warp_sum(Jl[0] * Jl[0] + Jl[1] * Jl[1]);
        const double gl = warp_sum(Jl[0] * r[0] + Jl[1] * r[1]);
        // v <= eps: the landmark's eigenvalue of Amm is truncated by the reference's pseudo-inverse (marginalization_factor.cpp:278-283).
        // Its coupling to the frame block is bounded by |B_l|^2 <= v C_jj <= 1e-8 C_jj (Amm is positive semi-definite), so the landmark is an
        // (almost) decoupled null direction: truncating it == not subtracting its Schur term; its factors' direct J^T J terms stay, as
        // in the reference. (A robot turning on the spot gives exactly this: zero baseline, d r / d lambda = 0.)
        const bool okv = v > kMargEps;
        const double isv = okv ? rsqrt(v) : 0.0;
        double q6[6];
#pragma unroll
        for (int i = 0; i < 13; i++) {
          const double wi = warp_sum(Jc13[0][i] * Jl[0] + Jc13[1][i] * Jl[1]);
          if (i < 6) q6[i] = wi;
          if (lane == 0) wv[i < 6 ? i : 6 + NP + (i - 6)] = wi * isv;
        }
        if (on) {
#pragma unroll
          for (int c = 0; c < 6; c++) wv[6 + 6 * (k - 1) + c] = (Jj[c] * Jl[0] + Jj[6 + c] * Jl[1]) * isv;
        }
        if (lane == 0) wv[NC] = gl * isv;
        // eps-shifted test matrix: Y_eps = Y - eps I - sum q q^T eps / (v (v - eps))
        if (okv && lane < 21) {
          int a = 0, c = lane; while (c >= 6 - a) { c -= 6 - a; a++; } c += a;
          double qa = 0, qc = 0;
#pragma unroll
          for (int i = 0; i < 6; i++) { if (i == a) qa = q6[i]; if (i == c) qc = q6[i]; }
          atomicAdd(&s.C6[a][c], qa * qc * (kMargEps / (v * (v - kMargEps))));
        }
        // direct J^T J, J^T r: common x common (reduced over the lanes, kept in registers across landmarks),
        // common x pose_k and pose_k x pose_k (lane local, spread over the pose blocks -> shared atomics)
        {
          int e = 0;
#pragma unroll
          for (int a = 0; a < 13; a++) {
#pragma unroll
            for (int c = a; c < 13; c++, e++) {
              const double val = warp_sum(Jc13[0][a] * Jc13[0][c] + Jc13[1][a] * Jc13[1][c]);
              if (lane == (e & 31)) cacc[e >> 5] += val;
            }
            const double gb = warp_sum(Jc13[0][a] * r[0] + Jc13[1][a] * r[1]);
            if (lane == a) cb += gb;
          }
        }
        if (on) {
          const int cj = kMargM + 6 * (k - 1);
#pragma unroll
          for (int a = 0; a < 13; a++) {
            const int ia = a < 6 ? a : (a < 12 ? cEX + (a - 6) : cTD);
#pragma unroll
            for (int c = 0; c < 6; c++) {
              const double val = Jc13[0][a] * Jj[c] + Jc13[1][a] * Jj[6 + c];
              if (ia <= cj) atomicAdd(&A[ia * kMargLD + cj + c], val); else atomicAdd(&A[(cj + c) * kMargLD + ia], val);
            }
          }
#pragma unroll
          for (int a = 0; a < 6; a++) {
#pragma unroll
            for (int c = a; c < 6; c++) atomicAdd(&A[(cj + a) * kMargLD + cj + c], Jj[a] * Jj[c] + Jj[6 + a] * Jj[6 + c]);
            atomicAdd(&s.b[cj + a], Jj[a] * r[0] + Jj[6 + a] * r[1]);
          }
        }
      }
      __syncthreads();
      // rank-kMargChunk update of the owned entries: acc += sum_r W[r][a] W[r][c]
#pragma unroll
      for (int i = 0; i < kMargOwn; i++) {
        if (ea[i] < 0) continue;
        double acc = eacc[i];
#pragma unroll 8
        for (int r = 0; r < kMargChunk; r++) acc += s.W[r][ea[i]] * s.W[r][ec[i]];
        eacc[i] = acc;
      }
      __syncthreads();
    }
    // flush: per-warp common block (atomics, 8 warps) then the owned elimination entries (after a barrier, plain stores)
    {
      int e = 0;
      for (int a = 0; a < 13; a++) for (int c = a; c < 13; c++, e++) {
        if (lane == (e & 31)) {
          const int ia = a < 6 ? a : (a < 12 ? cEX + (a - 6) : cTD), ic = c < 6 ? c : (c < 12 ? cEX + (c - 6) : cTD);
          atomicAdd(&A[ia * kMargLD + ic], cacc[e >> 5]);
        }
      }
      if (lane < 13) atomicAdd(&s.b[lane < 6 ? lane : (lane < 12 ? cEX + (lane - 6) : cTD)], cb);
    }
    if (lane == 0) atomicMax(&s.maxlen, maxlen);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kMargOwn; i++) {
      if (ea[i] < 0) continue;
      const int a = ea[i], c = ec[i];
      const int ia = a < 6 ? a : (a < 6 + NP ? kMargM + (a - 6) : (a < 12 + NP ? cEX + (a - 6 - NP) : cTD));
      if (c == NC) s.b[ia] -= eacc[i];
      else {
        const int ic = c < 6 ? c : (c < 6 + NP ? kMargM + (c - 6) : (c < 12 + NP ? cEX + (c - 6 - NP) : cTD));
        A[ia * kMargLD + ic] -= eacc[i];
      }
    }
    __syncthreads();
    if (t == 0 && s.n_lm0 > 0) { s.mtouched[0] = 1; s.touched[F] = 1; s.touched[F + 1] = 1; }
    if (t >= 1 && t < s.maxlen) s.touched[t - 1] = 1;   // poses 1..maxlen-1 observed a frame-0 landmark -> new index t-1
    __syncthreads();
  }

  if (mp.nranks > 1) {   // factor-sharded mode: hand the partial system to the all-reduce, k_marg_finish takes it from there
    double* ss = mp.stage_sum + (size_t)w * kMargStageSum; double* sm = mp.stage_max + (size_t)w * kMargStageMax;
    for (int e = t; e < kMargTMax * kMargTMax; e += blockDim.x) { const int a = e / kMargTMax, c = e % kMargTMax; ss[e] = (a <= c && c < T) ? A[a * kMargLD + c] : 0.0; }
    for (int i = t; i < kMargTMax; i += blockDim.x) ss[kMargTMax * kMargTMax + i] = s.b[i];
    if (t < 36) ss[kMargTMax * kMargTMax + kMargTMax + t] = (&s.C6[0][0])[t];
    if (t == 0) ss[kMargTMax * kMargTMax + kMargTMax + 36] = (double)s.n_lm0;
    if (t < kMargBlocksMax) sm[t] = (double)s.touched[t];
    if (t < 2) sm[kMargBlocksMax + t] = (double)s.mtouched[t];
    if (t == 0) { sm[kMargBlocksMax + 2] = (double)s.bad; sm[kMargBlocksMax + 3] = 1.0; }
    return;
  }
  marg_finish_block(p, w, mp, s, A, T, K);
}

// second half of k_marg_build in the factor-sharded mode: the all-reduced system of a window back into shared memory, then the elimination
__global__ void __launch_bounds__(kMargThreads) k_marg_finish(KP p, int w0, MargP mp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MargShared& s = *reinterpret_cast<MargShared*>(smem_raw);
  double* A = reinterpret_cast<double*>(smem_raw + ((sizeof(MargShared) + 15) & ~size_t(15)));
  const int w = w0 + blockIdx.x, t = threadIdx.x;
  const int F = p.F, NP = 6 * (F - 1), K = NP + 16 + (p.use_wheel ? 10 : 0), T = kMargM + K;
  const double* ss = mp.stage_sum + (size_t)w * kMargStageSum; const double* sm = mp.stage_max + (size_t)w * kMargStageMax;
  if (sm[kMargBlocksMax + 3] == 0.0) return;   // GF2_MARG_UNCHANGED on every rank (status written by k_marg_build)
  for (int e = t; e < kMargTMax * kMargLD; e += blockDim.x) A[e] = 0.0;
  __syncthreads();
  for (int e = t; e < kMargTMax * kMargTMax; e += blockDim.x) { const int a = e / kMargTMax, c = e % kMargTMax; if (a <= c) A[a * kMargLD + c] = ss[e]; }
  for (int i = t; i < kMargTMax; i += blockDim.x) s.b[i] = ss[kMargTMax * kMargTMax + i];
  if (t < 36) (&s.C6[0][0])[t] = ss[kMargTMax * kMargTMax + kMargTMax + t];
  if (t < kMargBlocksMax) s.touched[t] = sm[t] != 0.0;
  if (t < 2) s.mtouched[t] = sm[kMargBlocksMax + t] != 0.0;
  if (t == 0) { s.n_lm0 = (int)(ss[kMargTMax * kMargTMax + kMargTMax + 36] + 0.5); s.bad = (int)sm[kMargBlocksMax + 2]; s.maxlen = 0; }
  __syncthreads();
  marg_finish_block(p, w, mp, s, A, T, K);
}

// --------------------------------------------------------------------------------------------------------------------
// Cyclic Jacobi (two-sided, round-robin parallel ordering) on an n x n symmetric matrix in shared memory; V accumulates
// the rotations. Stand-in for Eigen::SelfAdjointEigenSolver at VE/factor/marginalization_factor.cpp:293.
struct EigShared {
  double c[kMargKMax / 2 + 1], s[kMargKMax / 2 + 1];
  int pp[kMargKMax / 2 + 1], qq[kMargKMax / 2 + 1];
  double red[2][kEigThreads / 32];
  double bvec[kMargKMax];
  int col[kMargKMax];          // compact column -> canonical kept column
  int col2[kMargKMax], pos[kMargKMax];   // pivot order of the rank-revealing Cholesky and its inverse
  double y[kMargKMax];
  int boff[kMargBlocksMax];    // compact offset of each kept block (-1 if absent)
  int n, done;
};

// Rank-revealing Cholesky (diagonal pivoting, right-looking) of the symmetric n x n matrix M in shared memory by the whole CTA:
// P^T M P = L L^T + (remainder whose diagonal is <= tol). perm[j] = original index of the j-th pivot; returns the rank (number of
// pivots above tol). Rows / columns are swapped physically, so afterwards M[i][j], i >= j, j < rank holds L in pivoted order.
__device__ __forceinline__ int marg_pivoted_cholesky(double* M, int n, int LD, double tol, int* perm, double* red_val, int* red_idx) {
  const int t = threadIdx.x, nt = blockDim.x, lane = t & 31, wid = t >> 5, nw = nt >> 5;
  if (t < n) perm[t] = t;
  for (int j = 0; j < n; j++) {
    __syncthreads();                       // the trailing update of column j - 1 (and perm) is complete
    // pivot: largest remaining diagonal entry (first index on ties)
    double best = -1.0; int bi = j;
    for (int i = j + t; i < n; i += nt) { const double d = M[i * LD + i]; if (d > best) { best = d; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { red_val[wid] = best; red_idx[wid] = bi; }
    __syncthreads();
    best = red_val[0]; bi = red_idx[0];
    for (int q = 1; q < nw; q++) if (red_val[q] > best || (red_val[q] == best && red_idx[q] < bi)) { best = red_val[q]; bi = red_idx[q]; }
    if (!(best > tol)) return j;           // uniform: every thread reduced the same values
    __syncthreads();                       // red_* are read
    if (bi != j) {                         // symmetric swap j <-> bi of the full square (both triangles are kept consistent)
      for (int c = t; c < n; c += nt) { const double x = M[j * LD + c]; M[j * LD + c] = M[bi * LD + c]; M[bi * LD + c] = x; }
      __syncthreads();
      for (int r = t; r < n; r += nt) { const double x = M[r * LD + j]; M[r * LD + j] = M[r * LD + bi]; M[r * LD + bi] = x; }
      if (t == 0) { const int x = perm[j]; perm[j] = perm[bi]; perm[bi] = x; }
      __syncthreads();
    }
    const double ljj = sqrt(M[j * LD + j]), inv = 1.0 / ljj;
    __syncthreads();                       // every thread has read the pivot
    for (int i = j + 1 + t; i < n; i += nt) M[i * LD + j] *= inv;
    if (t == 0) M[j * LD + j] = ljj;
    __syncthreads();
    const int m = n - j - 1;
    for (int idx = t; idx < m * m; idx += nt) {   // trailing update of the full remaining square (symmetric)
      const int a = idx / m, c = idx - a * m;
      M[(j + 1 + a) * LD + j + 1 + c] -= M[(j + 1 + a) * LD + j] * M[(j + 1 + c) * LD + j];
    }
  }
  __syncthreads();
  return n;
}

__global__ void __launch_bounds__(kEigThreads, 2) k_marg_eig(KP p, int w0, MargP mp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EigShared& s = *reinterpret_cast<EigShared*>(smem_raw);
  double* A = reinterpret_cast<double*>(smem_raw + ((sizeof(EigShared) + 15) & ~size_t(15)));
  const int Kc = 6 * (p.F - 1) + 16 + (p.use_wheel ? 10 : 0);   // capacity of this handle's kept system
  const int LD = Kc | 1;
  double* V = A + Kc * LD;
  const int w = w0 + blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5, nw = blockDim.x >> 5;
  const int F = p.F, NP = 6 * (F - 1);
  const int st = mp.status[w];
  if (st == GF2_MARG_UNCHANGED_ || st == GF2_MARG_UNSUPPORTED_ || st == GF2_MARG_DEGENERATE_) return;   // old prior stays
  if (st == GF2_MARG_INVALID_) { if (t == 0) { mp.out_rows[w] = 0; mp.out_nblocks[w] = 0; } return; }       // valid = false, :205-210
  if (t == 0) {
    int n = 0;
    const int nblk = F + 2 + (p.use_wheel ? 5 : 0);
    for (int b = 0; b < kMargBlocksMax; b++) {
      int kind, ls, c0; marg_block_info(F, b, kind, ls, c0);
      if (b < nblk && mp.touched[(size_t)w * kMargBlocksMax + b]) { s.boff[b] = n; for (int k = 0; k < ls; k++) s.col[n + k] = c0 + k; n += ls; }
      else s.boff[b] = -1;
    }
    s.n = n; s.done = 0;
  }
  __syncthreads();
  const int n = s.n;
  if (n > p.Pr) { if (t == 0) mp.status[w] = GF2_MARG_TOO_LARGE_; return; }
  const double* Ai = mp.A + (size_t)w * kMargKMax * kMargKMax;
  for (int e = t; e < n * n; e += blockDim.x) { const int a = e / n, c = e % n; A[a * LD + c] = Ai[s.col[a] * kMargKMax + s.col[c]]; V[a * LD + c] = (a == c) ? 1.0 : 0.0; }
  if (t < n) s.bvec[t] = mp.b[(size_t)w * kMargKMax + s.col[t]];
  __syncthreads();
  double* J0 = mp.out_J0 + (size_t)w * p.Pr * p.Pr;
  double* r0 = mp.out_r0 + (size_t)w * p.Pr;
  // ---- default factorisation. The reference takes the eigen-decomposition of the kept system, drops the eigenvalues <= eps = 1e-8 (:294-303)
  // and sets J0 = sqrt(S) V^T, r0 = sqrt(1/S) V^T b. Every consumer (MarginalizationFactor::Evaluate, the next marginalization) sees J0
  // only through r0 + J0 dx, J0^T J0 and J0^T r0, so any factor with the same two products is the same prior. A rank-revealing Cholesky
  // factorisation, P^T A P = L L^T with the pivots <= eps dropped, gives J0 = L^T P^T, r0 = L^-1 P^T b: J0^T J0 differs from the
  // eigenvalue-truncated A by O(eps) in absolute terms (entries of A reach 1e8: relative 1e-16) at n^3 / 3 operations instead of ~12 Jacobi
  // sweeps of 3 n^3 (2.4 ms -> 0.1 ms for one 76-dim window). mp.eig != 0 selects the literal eigen-decomposition below instead; the
  // tests run both and compare them with the restated / compiled reference.
  if (!mp.eig) {
    for (int e = t; e < n * n; e += blockDim.x) { const int a = e / n, c = e % n; V[a * LD + c] = A[a * LD + c]; }
    __syncthreads();
    const int rank = marg_pivoted_cholesky(V, n, LD, kMargEps, s.col2, &s.red[0][0], s.pp);
    __syncthreads();
    // inverse permutation: position of original column c in the pivoted order
    if (t < n) s.pos[s.col2[t]] = t;
    // r0 = L^-1 P^T b over the kept pivots (column-oriented forward substitution), zero for the dropped ones
    if (t < n) s.y[t] = s.bvec[s.col2[t]];
    __syncthreads();
    for (int j = 0; j < rank; j++) {
      if (t == 0) s.y[j] /= V[j * LD + j];
      __syncthreads();
      const double yj = s.y[j];
      for (int i = j + 1 + t; i < rank; i += blockDim.x) s.y[i] -= V[i * LD + j] * yj;
      __syncthreads();
    }
    for (int e = t; e < n * n; e += blockDim.x) { const int k = e / n, c = e % n; const int pc = s.pos[c]; J0[k * p.Pr + c] = (k < rank && pc >= k) ? V[pc * LD + k] : 0.0; }   // J0 = L^T P^T
    if (t < n) r0[t] = t < rank ? s.y[t] : 0.0;
    if (t == 0) s.done = 2;
    __syncthreads();
  }
  const bool fast_path = s.done == 2;
  // Round-robin ordering: np - 1 steps per sweep, each with np / 2 disjoint pairs. Every pair is served by a group of tpp
  // threads that all derive the rotation from (app, aqq, apq) themselves; per step: rotate rows of A and columns of V, then
  // columns of A.
  const int np = n + (n & 1), half = np / 2;
  const int tpp = blockDim.x / half;
  const int k = t / tpp, j0 = t % tpp;
  const bool act = k < half;
  for (int sweep = 0; sweep < 60 && !fast_path; sweep++) {
    // convergence: |off| <= 1e-15 |diag| (Frobenius). The rounding floor of the rotations sits near 1e-16, so the oracle's 1e-16
    // test can stall for the full sweep budget here; the quadratic phase jumps from ~1e-13 straight to the floor.
    double off = 0, dg = 0;
    for (int a = wid; a < n; a += nw) for (int c = lane; c < n; c += 32) { const double v = A[a * LD + c]; if (a == c) dg += v * v; else if (a < c) off += v * v; }
    off = warp_sum(off); dg = warp_sum(dg);
    if (lane == 0) { s.red[0][wid] = off; s.red[1][wid] = dg; }
    __syncthreads();
    if (t == 0) { double o = 0, d = 0; for (int i = 0; i < nw; i++) { o += s.red[0][i]; d += s.red[1][i]; } s.done = (o <= 1e-30 * (d + 1e-300)); }
    __syncthreads();
    if (s.done) break;
    for (int step = 0; step < np - 1; step++) {
      // one thread per pair derives the rotation (fp64 sqrt / divide are long software sequences on the fp64 pipe: doing this
      // redundantly in every thread of the group saturates that pipe) and publishes it through shared memory
      if (act && j0 == 0) {
        int a, b;
        if (k == 0) { a = np - 1; b = step; }
        else { a = step + k; if (a >= np - 1) a -= np - 1; b = step - k; if (b < 0) b += np - 1; }
        const int pi_ = a < b ? a : b; int qi_ = a < b ? b : a;
        double c_ = 1.0, s_ = 0.0;
        if (qi_ >= n) qi_ = -1;
        else {
          const double apq = A[pi_ * LD + qi_];
          if (apq != 0.0) {
            // t = sgn(tau) / (|tau| + sqrt(1 + tau^2)), tau = (aqq - app) / (2 apq)  ==  sgn(d) e / (|d| + sqrt(d^2 + e^2))
            const double d = A[qi_ * LD + qi_] - A[pi_ * LD + pi_], e = 2.0 * apq;
            const double tt = (d >= 0 ? e : -e) / (fabs(d) + sqrt(d * d + e * e));
            c_ = rsqrt(1.0 + tt * tt); s_ = tt * c_;
          } else qi_ = -1;
        }
        s.pp[k] = pi_; s.qq[k] = qi_; s.c[k] = c_; s.s[k] = s_;
      }
      __syncthreads();
      int pi = 0, qi = -1; double c = 1.0, sn = 0.0;
      if (act) { pi = s.pp[k]; qi = s.qq[k]; c = s.c[k]; sn = s.s[k]; }
      if (qi >= 0) {
        for (int j = j0; j < n; j += tpp) {   // rows: A <- J^T A ; columns: V <- V J
          const double x = A[pi * LD + j], y = A[qi * LD + j];
          A[pi * LD + j] = c * x - sn * y; A[qi * LD + j] = sn * x + c * y;
          const double vx = V[j * LD + pi], vy = V[j * LD + qi];
          V[j * LD + pi] = c * vx - sn * vy; V[j * LD + qi] = sn * vx + c * vy;
        }
      }
      __syncthreads();
      if (qi >= 0) {
        for (int i = j0; i < n; i += tpp) {   // columns: A <- A J
          const double x = A[i * LD + pi], y = A[i * LD + qi];
          A[i * LD + pi] = c * x - sn * y; A[i * LD + qi] = sn * x + c * y;
        }
      }
      __syncthreads();
    }
  }
  // linearized_jacobians = sqrt(S) V^T, linearized_residuals = sqrt(1/S) V^T b with the eps truncation of :294-303
  for (int e = t; e < n * n && !fast_path; e += blockDim.x) {
    const int k = e / n, c = e % n;
    const double S = A[k * LD + k];
    J0[k * p.Pr + c] = S > kMargEps ? sqrt(S) * V[c * LD + k] : 0.0;
  }
  if (t < n && !fast_path) {
    const double S = A[t * LD + t];
    double vb = 0; for (int c = 0; c < n; c++) vb += V[c * LD + t] * s.bvec[c];
    r0[t] = S > kMargEps ? sqrt(1.0 / S) * vb : 0.0;
  }
  // kept blocks renamed by addr_shift; keep_block_data = the states at marginalization time (preMarginalize, :119-138)
  if (t < kMargBlocksMax) {
    const int b = t;
    if (s.boff[b] >= 0) {
      int slot = 0; for (int i = 0; i < b; i++) slot += (s.boff[i] >= 0);
      gf2_prior_block& o = mp.out_blocks[(size_t)w * (2 * F + 8) + slot];
      int kind, ls, c0; marg_block_info(F, b, kind, ls, c0);
      o.kind = kind; o.index = 0; o.offset = s.boff[b]; o.pad_ = 0;
      for (int k = 0; k < 9; k++) o.x0[k] = 0.0;
      if (kind == GF2_BLK_POSE) {
        const int old = mp.mode == 0 ? b + 1 : (b == F - 2 ? F - 1 : b);
        o.index = b;
        for (int k = 0; k < 7; k++) o.x0[k] = p.pose[(size_t)w * F * 7 + 7 * old + k];
      } else if (kind == GF2_BLK_SPEEDBIAS) {
        const int old = mp.mode == 0 ? 1 : 0;
        for (int k = 0; k < 9; k++) o.x0[k] = p.sb[(size_t)w * F * 9 + 9 * old + k];
      } else if (kind == GF2_BLK_EX_POSE) { for (int k = 0; k < 7; k++) o.x0[k] = p.ex[(size_t)w * 7 + k]; }
      else if (kind == GF2_BLK_TD) o.x0[0] = p.td[w];
      else if (kind == GF2_BLK_EX_WHEEL) { for (int k = 0; k < 7; k++) o.x0[k] = p.exw[(size_t)w * 7 + k]; }
      else if (kind == GF2_BLK_TD_WHEEL) o.x0[0] = p.tdw[w];
      else o.x0[0] = p.sxw[(size_t)w * 3 + (kind - GF2_BLK_SX)];
    }
  }
  if (t == 0) {
    int nb = 0; for (int b = 0; b < kMargBlocksMax; b++) nb += (s.boff[b] >= 0);
    mp.out_rows[w] = n; mp.out_nblocks[w] = nb;
  }
}

}  // namespace gf2
