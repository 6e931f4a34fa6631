// gf2_solver_solve.cuh — k_nonvis (IMU / prior linearisation) and k_solve2 (reduced camera system: assembly,
// blocked Cholesky on the fp64 tensor cores, Gauss-Newton step).
//
// The reduced system S' (D = 15 F, <= 165) is kept in shared memory as the lower block triangle of 15x15 frame
// blocks, each padded to 16 rows x 20 doubles (row stride 20 keeps the mma.m8n8k4 fragment loads of a half-warp on
// distinct banks). Right-looking blocked Cholesky: per block column K, warp 0 factors the diagonal block and inverts
// it, the panel blocks become A_IK * L_KK^-T and the trailing blocks A_IJ -= L_IK L_JK^T, both as DMMA products.
#pragma once
#include "gf2_solver_kernels2.cuh"

namespace gf2 {

constexpr int kBS = 20;            // row stride of a frame block
constexpr int kBlk = 16 * kBS;     // doubles per frame block
constexpr int kNonvisThreads = 320;

__device__ __forceinline__ int bidx(int I, int J) { return (I * (I + 1) / 2 + J) * kBlk; }  // J <= I

// ------------------------------------------------------------------------------------------------ k_nonvis
// IMUFactor linearisation (VE/factor/imu_factor.h:28-191): per factor the 30x30 J^T J (packed lower), J^T r and cost.
// Marginalization prior (VE/factor/marginalization_factor.cpp:344-392): r = r0 + J0 dx, gradient J0^T r, cost.
__global__ void __launch_bounds__(kNonvisThreads, 2) k_nonvis(KP p, int w0) {
  const int w = w0 + blockIdx.x;
  WinState& st = p.st[w];
  if (!st.active || st.reuse) return;
  __shared__ double scratch[kNonvisThreads / 32][472];
  __shared__ double dx[kP], pr[kP], red[32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5, nw = blockDim.x >> 5;
  const int F = p.F;
  const double* pose = p.pose + (size_t)w * F * 7;
  const double* sb = p.sb + (size_t)w * F * 9;
  double cost = 0.0;
  if (p.imu) {
    for (int k = wid; k < F - 1; k += nw) {
      const gf2_imu_preint& pre = p.imu[(size_t)w * (F - 1) + k];
      double* Hout = p.imu_H + ((size_t)w * (F - 1) + k) * 675;
      double* gout = p.imu_g + ((size_t)w * (F - 1) + k) * 30;
      if (!pre.valid || pre.sum_dt > 10.0) {
        for (int i = lane; i < 675; i += 32) Hout[i] = 0.0;
        if (lane < 30) gout[lane] = 0.0;
        continue;
      }
      double* J = scratch[wid];
      double* r = J + 450;
      for (int i = lane; i < 450; i += 32) J[i] = 0.0;
      __syncwarp();
      { ImuStates s2 = load_imu_states(pose, sb, k); imu_raw_warp(pre, s2, p.g_norm, r, J, lane); }
      __syncwarp();
      const double* sq = p.imu_sqrt + ((size_t)w * (F - 1) + k) * 225;
      const int mr = lane >> 2, mq = lane & 3;
      {  // J <- sqrt_info * J (15x15 upper triangular times 15x30) on the fp64 tensor cores; residual by lane 30 alongside
        double jv[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
#pragma unroll
          for (int ks = 0; ks < 4; ks++) { const int kk = 4 * ks + mq, c = 8 * nt + mr; jv[nt][ks] = (kk < 15 && c < 30) ? J[kk * 30 + c] : 0.0; }
        double rs = 0.0;
        if (lane < 15) { for (int kk = lane; kk < 15; kk++) rs += sq[lane * 15 + kk] * r[kk]; }
        double acc[2][4][2];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
          double av[4];
#pragma unroll
          for (int ks = 0; ks < 4; ks++) { const int a = 8 * mt + mr, kk = 4 * ks + mq; av[ks] = (a < 15 && kk < 15 && kk >= a) ? sq[a * 15 + kk] : 0.0; }
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < 4; ks++) mma_f64(acc[mt][nt][0], acc[mt][nt][1], av[ks], jv[nt][ks]);
          }
        }
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            const int a = 8 * mt + mr, c = 8 * nt + 2 * mq;
            if (a < 15 && c < 30) { J[a * 30 + c] = acc[mt][nt][0]; J[a * 30 + c + 1] = acc[mt][nt][1]; }
          }
        if (lane < 15) r[lane] = rs;
      }
      __syncwarp();
      if (lane == 0) { double c = 0; for (int a = 0; a < 15; a++) c += r[a] * r[a]; cost += 0.5 * c; }
      {  // J^T J (30x30, K = 15): lower tiles only; blocks (i,i), (j,i), (j,j) in the layout k_solve2 reads (diagonal blocks: lower part)
        double jv[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
#pragma unroll
          for (int ks = 0; ks < 4; ks++) { const int kk = 4 * ks + mq, c = 8 * nt + mr; jv[nt][ks] = (kk < 15 && c < 30) ? J[kk * 30 + c] : 0.0; }
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
          for (int nt = 0; nt <= mt; nt++) {
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 4; ks++) mma_f64(d0, d1, jv[mt][ks], jv[nt][ks]);
            const int a = 8 * mt + mr;
#pragma unroll
            for (int h2 = 0; h2 < 2; h2++) {
              const int b = 8 * nt + 2 * mq + h2; const double v = h2 ? d1 : d0;
              if (a >= 30 || b >= 30) continue;
              if (a < 15) { if (b < 15) Hout[a * 15 + b] = v; }
              else if (b < 15) Hout[225 + (a - 15) * 15 + b] = v;
              else Hout[450 + (a - 15) * 15 + (b - 15)] = v;
            }
          }
      }
      if (lane < 30) { double acc = 0; for (int rr = 0; rr < 15; rr++) acc += J[rr * 30 + lane] * r[rr]; gout[lane] = acc; }
      __syncwarp();
    }
  }
  if (p.wheel) {  // WheelFactor: 6 residuals, pose_i / pose_j blocks; with free calibration blocks (p.wcal) also their columns
    for (int k = wid; k < F - 1; k += nw) {
      const gf2_wheel_preint& pre = p.wheel[(size_t)w * (F - 1) + k];
      double* Hout = p.wheel_H + ((size_t)w * (F - 1) + k) * 108;
      double* gout = p.wheel_g + ((size_t)w * (F - 1) + k) * 12;
      double* Hc = p.wcal ? p.wheel_Hc + ((size_t)w * (F - 1) + k) * 220 : nullptr;
      double* gc = p.wcal ? p.wheel_gc + ((size_t)w * (F - 1) + k) * 10 : nullptr;
      if (!pre.valid || pre.sum_dt > 10.0) {
        for (int i = lane; i < 108; i += 32) Hout[i] = 0.0;
        if (lane < 12) gout[lane] = 0.0;
        if (Hc) { for (int i = lane; i < 220; i += 32) Hc[i] = 0.0; if (lane < 10) gc[lane] = 0.0; }
        continue;
      }
      double* J = scratch[wid];   // 6 x 12 [pose_i | pose_j]
      double* r = J + 72;
      double* Jc = J + 80;        // 6 x 10 [ex_wheel 6 | sx | sy | sw | td_wheel]
      __syncwarp();
      if (lane == 0) wheel_raw(pre, pose + 7 * k, pose + 7 * (k + 1), p.exw + (size_t)w * 7, p.sxw + (size_t)w * 3, p.tdw[w], r, J);
      else if (lane == 1 && p.wcal) wheel_calib_jacobians_masked(pre, pose + 7 * k, pose + 7 * (k + 1), p.exw + (size_t)w * 7, p.sxw + (size_t)w * 3, p.tdw[w], p.wcal, Jc);
      __syncwarp();
      const double* sq = p.wheel_sqrt + ((size_t)w * (F - 1) + k) * 36;
      if (lane < 12) { for (int a = 0; a < 6; a++) { double acc = 0; for (int kk = a; kk < 6; kk++) acc += sq[a * 6 + kk] * J[kk * 12 + lane]; J[a * 12 + lane] = acc; } }
      else if (lane == 12) { for (int a = 0; a < 6; a++) { double acc = 0; for (int kk = a; kk < 6; kk++) acc += sq[a * 6 + kk] * r[kk]; r[a] = acc; } }
      else if (lane >= 16 && lane < 26 && p.wcal) { const int c = lane - 16; for (int a = 0; a < 6; a++) { double acc = 0; for (int kk = a; kk < 6; kk++) acc += sq[a * 6 + kk] * Jc[kk * 10 + c]; Jc[a * 10 + c] = acc; } }
      __syncwarp();
      if (lane == 0) { double c = 0; for (int a = 0; a < 6; a++) c += r[a] * r[a]; cost += 0.5 * c; }
      for (int idx = lane; idx < 108; idx += 32) {  // 6x6 blocks (i,i), (j,i), (j,j)
        const int blk = idx / 36, e = idx % 36;
        const int a = e / 6 + (blk >= 1 ? 6 : 0), b = e % 6 + (blk == 2 ? 6 : 0);
        double acc = 0; for (int rr = 0; rr < 6; rr++) acc += J[rr * 12 + a] * J[rr * 12 + b];
        Hout[idx] = acc;
      }
      if (lane < 12) { double acc = 0; for (int rr = 0; rr < 6; rr++) acc += J[rr * 12 + lane] * r[rr]; gout[lane] = acc; }
      if (Hc) {
        for (int idx = lane; idx < 220; idx += 32) {
          double acc = 0;
          if (idx < 120) { const int a = (idx % 60) / 6, b = idx % 6 + (idx >= 60 ? 6 : 0); for (int rr = 0; rr < 6; rr++) acc += Jc[rr * 10 + a] * J[rr * 12 + b]; }
          else { const int a = (idx - 120) / 10, b = (idx - 120) % 10; for (int rr = 0; rr < 6; rr++) acc += Jc[rr * 10 + a] * Jc[rr * 10 + b]; }
          Hc[idx] = acc;
        }
        if (lane < 10) { double acc = 0; for (int rr = 0; rr < 6; rr++) acc += Jc[rr * 10 + lane] * r[rr]; gc[lane] = acc; }
      }
      __syncwarp();
    }
  }
  const int n = p.prior_rows ? p.prior_rows[w] : 0;
  if (n > 0) {
    const gf2_prior_block* blk = p.prior_blocks + (size_t)w * (2 * F + 8);
    if (t < p.prior_nblocks[w]) prior_block_dx(blk[t], pose, sb, calib_of(p, w, false), dx);
    __syncthreads();
    const double* J0 = p.prior_J0 + (size_t)w * p.Pr * p.Pr;
    const double* r0 = p.prior_r0 + (size_t)w * p.Pr;
    if (t < n) { double s2 = r0[t]; for (int c = 0; c < n; c++) s2 += J0[t * p.Pr + c] * dx[c]; pr[t] = s2; cost += 0.5 * s2 * s2; }
    __syncthreads();
    if (t < n) { double s2 = 0; for (int r = 0; r < n; r++) s2 += J0[r * p.Pr + t] * pr[r]; p.prior_g[(size_t)w * p.Pr + t] = s2; }
  }
  cost = warp_sum(cost);
  if (lane == 0) red[wid] = cost;
  __syncthreads();
  if (t == 0) { double c = 0; for (int i = 0; i < nw; i++) c += red[i]; p.cost_nv[w] = c; }
}

// ------------------------------------------------------------------------------------------------ k_solve2
// Storage of the reduced camera matrix / its Cholesky factor in shared memory. In frame-major order [pose 6 | speed-bias 9]
// the speed-bias rows of frame I are structurally zero against every frame J <= I - 2 — in S (an IMU factor links
// adjacent frames only) and in L (eliminating frame J connects its later neighbours: every later pose and the speed-bias of
// frame J + 1, nothing further). So blocks (I, J) with I - J >= 2 ("far") keep only their 6 pose rows. 15 x 20 doubles for a
// near block, 6 x 20 for a far one: 21 * 300 + 45 * 120 = 11,700 doubles instead of 66 * 320 = 21,120, which lets two
// windows share one SM (the kernel is a chain of short dependent phases; a second resident CTA fills the bubbles).
constexpr int kNearBlk = 15 * kBS, kFarBlk = 6 * kBS;

constexpr int kNBlkPairsS = kMaxF * (kMaxF + 1) / 2;
// Free wheel calibration blocks (KP::wcal) add ONE block row after the frames: "frame" F with tangent layout [ex_wheel 6 | sx sy sw |
// td_wheel | 5 unused]. A wheel factor couples it to the pose rows of two frames, the factorisation fills the rest, so all its
// blocks (F, J) are stored with 15 rows (12 more near blocks: 139 KB of shared memory, one window per SM in that mode). Constant
// sub-blocks and the unused entries are decoupled unit diagonals.
constexpr int kMaxFx = kMaxF + 1;
static_assert(kSolveThreads >= 15 * kMaxFx, "one thread per tangent dimension in the assembly");

struct Solve2Shared {
  double g[kMaxFx * 16], gs[kMaxFx * 16], Hd[kMaxFx * 16], s[kMaxFx * 16], e[kMaxFx * 16], u[kMaxFx * 16], z[kMaxFx * 16];
  double yu[kMaxFx * 16], fw[kMaxFx * 16];  // L^T u and the forward-substitution result (= L^T z)
  double red[8 * 32];
  double dinv[kMaxFx * 16];  // reciprocal diagonal of every factored diagonal block
  int boff[kMaxFx * kMaxFx];  // offset of block (I, J), J <= I
  int blin[kMaxF * (kMaxF + 1) / 2];  // offset of block number I (I + 1) / 2 + J (the order k_linearize writes Svis in)
  int flag;
  int pad_;
  double A[1];  // dynamic tail
};

__host__ __device__ __forceinline__ int solve2_matrix_doubles(int F) { return (2 * F - 1) * kNearBlk + ((F - 1) * (F - 2) / 2) * kFarBlk; }
__host__ __device__ __forceinline__ int solve2_matrix_doubles(int F, int wcal) { return solve2_matrix_doubles(F) + (wcal ? (F + 1) * kNearBlk : 0); }

// C (rc rows) -= X (rx rows) * Y (ry rows)^T over k = 0..15 on the fp64 tensor cores, one warp; rows beyond a block's own
// are never read (predicated to zero) or written. c may alias x.
__device__ __forceinline__ void block_mma(double* c, const double* x, const double* y, int rc, int rx, int ry, int lane) {
  const int r = lane >> 2, q = lane & 3;
  const int rmin = rc < rx ? rc : rx;
  const bool m1 = rmin > 8, n1 = ry > 8;
  double acc[2][2][2];
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nt = 0; nt < 2; nt++) {
      const bool ok = (8 * mt + r) < rc && (mt == 0 || m1) && (nt == 0 || n1);
      acc[mt][nt][0] = ok ? c[(8 * mt + r) * kBS + 8 * nt + 2 * q] : 0.0;
      acc[mt][nt][1] = ok ? c[(8 * mt + r) * kBS + 8 * nt + 2 * q + 1] : 0.0;
    }
#pragma unroll
  for (int ks = 0; ks < 4; ks++) {
    double a[2], b[2];
    a[0] = (r < rx) ? -x[r * kBS + 4 * ks + q] : 0.0;
    a[1] = (m1 && 8 + r < rx) ? -x[(8 + r) * kBS + 4 * ks + q] : 0.0;
    b[0] = (r < ry) ? y[r * kBS + 4 * ks + q] : 0.0;
    b[1] = (n1 && 8 + r < ry) ? y[(8 + r) * kBS + 4 * ks + q] : 0.0;
    mma_f64(acc[0][0][0], acc[0][0][1], a[0], b[0]);
    if (n1) mma_f64(acc[0][1][0], acc[0][1][1], a[0], b[1]);
    if (m1) { mma_f64(acc[1][0][0], acc[1][0][1], a[1], b[0]); if (n1) mma_f64(acc[1][1][0], acc[1][1][1], a[1], b[1]); }
  }
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nt = 0; nt < 2; nt++) {
      const bool ok = (8 * mt + r) < rc && (mt == 0 || m1) && (nt == 0 || n1);
      if (ok) { c[(8 * mt + r) * kBS + 8 * nt + 2 * q] = acc[mt][nt][0]; c[(8 * mt + r) * kBS + 8 * nt + 2 * q + 1] = acc[mt][nt][1]; }
    }
}

// One row of a panel: x L_KK^T = a (15 entries) by one thread, x -> out (may alias in). Out of line for the same reason as
// factor_diag_block.
__device__ __noinline__ void panel_row(const double* in, const double* Akk, const double* dinv, double* out) {
  double x[15];
#pragma unroll
  for (int c = 0; c < 15; c++) x[c] = in[c];
#pragma unroll
  for (int c = 0; c < 15; c++) {
    double acc = x[c];
#pragma unroll
    for (int k = 0; k < c; k++) acc -= x[k] * Akk[c * kBS + k];
    x[c] = acc * dinv[c];
  }
#pragma unroll
  for (int c = 0; c < 15; c++) out[c] = x[c];
}

// Cholesky factor of one 15x15 diagonal block in place by one warp: lane = row, the row lives in registers, pivots and
// column entries travel by shuffle; the reciprocal pivots go to dinv. Returns true when a pivot is not positive. Kept out of
// line so that the fully unrolled pivot chain is register-allocated on its own (inlined twice into k_solve2 it fell back to
// local memory, on the kernel's critical path).
template <int J>
__device__ __forceinline__ void chol_pivot(double (&a)[15], double* dinv, int lane, bool& bad) {
  if constexpr (J < 15) {
    const double d = __shfl_sync(0xffffffffu, a[J], J);
    if (!(d > 0.0)) bad = true;
    const double rs = fast_rsqrt(fmax(d, 1e-300));  // fp64 sqrt/div are long dependent software sequences: one rsqrt + multiplies
    const double lj = (lane == J) ? d * rs : a[J] * rs;
    if (lane == J) dinv[J] = rs;
    if (lane >= J) a[J] = lj;
#pragma unroll
    for (int k = J + 1; k < 15; k++) { const double lkj = __shfl_sync(0xffffffffu, lj, k); if (lane >= k) a[k] -= lj * lkj; }
    chol_pivot<J + 1>(a, dinv, lane, bad);  // compile-time recursion: the pivot loop must be fully unrolled for a[] to stay in registers
  }
}
__device__ __noinline__ bool factor_diag_block(double* Akk, double* dinv, int lane) {
  double a[15];
#pragma unroll
  for (int c = 0; c < 15; c++) a[c] = (lane < 15 && c <= lane) ? Akk[lane * kBS + c] : 0.0;
  bool bad = false;
  chol_pivot<0>(a, dinv, lane, bad);
  if (lane < 15) {
#pragma unroll
    for (int c = 0; c < 15; c++) Akk[lane * kBS + c] = (c <= lane) ? a[c] : 0.0;
  }
  return bad;
}

__global__ void __launch_bounds__(kSolveThreads, 2) k_solve2(KP p, int w0) {
  const int w = w0 + blockIdx.x;
  WinState& st = p.st[w];
  if (!st.active || st.reuse) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Solve2Shared& S = *reinterpret_cast<Solve2Shared*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5, nt = blockDim.x, nwarp = nt >> 5;
  const int F = p.F, D = p.D, NV = 6 * F;
  const int Fc = p.wcal ? F : -1;            // block row of the free wheel calibration (none: -1)
  const int Fx = F + (p.wcal ? 1 : 0), Dx = 15 * Fx;
  const int NB = F * (F + 1) / 2, NBx = Fx * (Fx + 1) / 2;
  const int NA = solve2_matrix_doubles(F, p.wcal);
  auto brows = [&](int I, int J) -> int { return (I != Fc && I - J >= 2) ? 6 : 15; };
  auto calib_live = [&](int c) -> bool { return c < 6 ? (p.wcal & 1) : c < 9 ? (p.wcal & 2) : c == 9 ? (p.wcal & 4) : false; };
  double* A = S.A;
#ifdef GF2_PHASE_CLOCKS
  long long pc[16]; int npc = 0;
#define GF2_PC() do { __syncthreads(); pc[npc++] = clock64(); } while (0)
#else
#define GF2_PC() do {} while (0)
#endif
  GF2_PC();
  if (t < NBx) {  // block offsets, row-major over the lower triangle (the calibration row, if any, comes last: all near blocks)
    int I = 0; while ((I + 1) * (I + 2) / 2 <= t) I++;
    const int J = t - I * (I + 1) / 2;
    int off = 0;
    for (int i = 0; i < I; i++) off += (i >= 2 ? (i - 1) * kFarBlk : 0) + (i >= 1 ? 2 : 1) * kNearBlk;
    if (I == Fc) off += J * kNearBlk;
    else off += (J <= I - 2) ? J * kFarBlk : ((I >= 2 ? (I - 1) * kFarBlk : 0) + (J - (I >= 1 ? I - 1 : 0)) * kNearBlk);
    S.boff[I * kMaxFx + J] = off; if (t < NB) S.blin[t] = off;
  }
  // ---- assembly. Every global load of the visual and IMU parts is issued before anything waits on one (the per-element
  // load -> shared-memory update loops this replaces cost 40k cycles of exposed DRAM latency per window)
  // and never summed in flight (an accumulate-as-you-load chain waits one DRAM latency per element). Thread t < 225 owns
  // element e = t of every 15x15 IMU block; the visual blocks are spread over all threads.
  constexpr int kVisPT = (kNBlkPairsS * 36 + kSolveThreads - 1) / kSolveThreads;          // 10
  const double* Svis = p.Svis + (size_t)w * kVisRec;
  const double* Hw = p.imu ? p.imu_H + (size_t)w * (F - 1) * 675 : nullptr;
  const bool imu_on = Hw != nullptr && t < 225;
  double vis[kVisPT], hii[kMaxF - 1], hji[kMaxF - 1], hjj[kMaxF - 1];  // blocks (i,i), (j,i), (j,j) of IMU factor K
#pragma unroll
  for (int u = 0; u < kVisPT; u++) { const int idx = t + u * kSolveThreads; vis[u] = idx < NB * 36 ? Svis[idx] : 0.0; }
#pragma unroll
  for (int K = 0; K < kMaxF - 1; K++) {
    const bool on = imu_on && K < F - 1;
    hii[K] = on ? Hw[K * 675 + t] : 0.0; hji[K] = on ? Hw[K * 675 + 225 + t] : 0.0; hjj[K] = on ? Hw[K * 675 + 450 + t] : 0.0;
  }
  double gv = 0.0, gsv = 0.0, udv = 0.0, gimu = 0.0;
  if (t < NV) { gv = p.gvis[(size_t)w * kVisRec + t]; gsv = p.gschur[(size_t)w * kVisRec + t]; udv = p.Udiag[(size_t)w * kVisRec + t]; }
  if (p.imu && t < D) {
    const double* gw = p.imu_g + (size_t)w * (F - 1) * 30;
    const int K = t / 15, r = t % 15;
    if (K > 0) gimu += gw[(K - 1) * 30 + 15 + r];
    if (K < F - 1) gimu += gw[K * 30 + r];
  }
  GF2_PC();
  for (int i = t; i < NA; i += nt) A[i] = 0.0;
  for (int i = t; i < Fx * 16; i += nt) { S.g[i] = 0.0; S.gs[i] = 0.0; S.Hd[i] = 0.0; S.z[i] = 0.0; S.u[i] = 0.0; }
  __syncthreads();
  GF2_PC();
  auto bidx2 = [&](int I, int J) -> int { return S.boff[I * kMaxFx + J]; };
  // element (i, j), j <= i, global tangent indices; rows a far block does not store read as zero / are never written
  auto stored = [&](int i, int j) -> bool { const int I = i / 15, J = j / 15; return (I - J < 2) || (i - 15 * I) < 6 || I == Fc; };
  auto Ael2 = [&](int i, int j) -> double& { const int I = i / 15, J = j / 15; return A[bidx2(I, J) + (i - 15 * I) * kBS + (j - 15 * J)]; };
  // visual Schur complement (blocked 6x6 layout written by k_linearize: coalesced read, pose part of each block)
#pragma unroll
  for (int u = 0; u < kVisPT; u++) {
    const int idx = t + u * kSolveThreads;
    if (idx < NB * 36) { const int blk = idx / 36, e = idx % 36; A[S.blin[blk] + (e / 6) * kBS + e % 6] = vis[u]; }
  }
  if (t < NV) { const int da = 15 * (t / 6) + t % 6; S.g[da] = gv; S.gs[da] = gsv; S.Hd[da] = udv; }
  __syncthreads();
  GF2_PC();
  // ---- prior (H = J0^T J0 precomputed by k_prepare, gradient by k_nonvis)
  const int n = p.prior_rows ? p.prior_rows[w] : 0;
  if (n > 0) {
    const int32_t* map = p.prior_map + (size_t)w * p.Pr;
    const double* H = p.prior_H + (size_t)w * p.Pr * p.Pr;
    if (t < n && map[t] >= 0) { S.g[map[t]] += p.prior_g[(size_t)w * p.Pr + t]; S.Hd[map[t]] += H[t * p.Pr + t]; }
    for (int idx = t; idx < n * n; idx += nt) {
      const int a = idx / n, b = idx % n;
      const int ma = map[a], mb = map[b];
      if (ma < 0 || mb < 0 || mb > ma) continue;
      if (stored(ma, mb)) Ael2(ma, mb) += H[a * p.Pr + b];   // gf2_set_prior rejects priors that would put mass elsewhere
    }
    __syncthreads();
  }
  GF2_PC();
  // ---- IMU blocks, by destination: A_KK += H_{K-1}[(j,j)] + H_K[(i,i)],  A_{K+1,K} += H_K[(j,i)]  (values loaded above)
  if (p.imu) {
    if (t < 225) {
      const int r = t / 15, c = t % 15;
#pragma unroll
      for (int K = 0; K < kMaxF; K++) {
        if (K >= F) break;
        const double v = (K > 0 ? hjj[K > 0 ? K - 1 : 0] : 0.0) + (K < kMaxF - 1 ? hii[K < kMaxF - 1 ? K : 0] : 0.0);
        if (c <= r) A[bidx2(K, K) + r * kBS + c] += v;
        if (r == c) S.Hd[15 * K + r] += v;
        if (K < F - 1) A[bidx2(K + 1, K) + r * kBS + c] += hji[K < kMaxF - 1 ? K : 0];
      }
    }
    if (t < D) S.g[t] += gimu;
    __syncthreads();
  }
  if (p.wheel) {  // wheel pose blocks, by destination like the IMU blocks (pose part = rows/cols 0..5 of a frame block)
    const double* Hw = p.wheel_H + (size_t)w * (F - 1) * 108;
    const double* gw = p.wheel_g + (size_t)w * (F - 1) * 12;
    for (int idx = t; idx < (2 * F - 1) * 36; idx += nt) {
      const int e = idx % 36, r = e / 6, c = e % 6;
      if (idx < F * 36) {
        const int K = idx / 36;
        double v = 0.0;
        if (K > 0) v += Hw[(K - 1) * 108 + 72 + e];
        if (K < F - 1) v += Hw[K * 108 + e];
        if (c <= r) A[bidx2(K, K) + r * kBS + c] += v;
        if (r == c) S.Hd[15 * K + r] += v;
      } else {
        const int K = idx / 36 - F;
        A[bidx2(K + 1, K) + r * kBS + c] += Hw[K * 108 + 36 + e];
      }
    }
    for (int i = t; i < 6 * F; i += nt) {
      const int K = i / 6, r = i % 6;
      double v = 0.0;
      if (K > 0) v += gw[(K - 1) * 12 + 6 + r];
      if (K < F - 1) v += gw[K * 12 + r];
      S.g[15 * K + r] += v;
    }
    if (p.wcal) {  // calibration block row: (calib, pose_K) from the factors K (as pose_i) and K-1 (as pose_j); (calib, calib) summed
      const double* Hc = p.wheel_Hc + (size_t)w * (F - 1) * 220;
      const double* gc = p.wheel_gc + (size_t)w * (F - 1) * 10;
      for (int idx = t; idx < F * 60; idx += nt) {
        const int K = idx / 60, e = idx % 60;
        double v = 0.0;
        if (K < F - 1) v += Hc[K * 220 + e];
        if (K > 0) v += Hc[(K - 1) * 220 + 60 + e];
        A[bidx2(F, K) + (e / 6) * kBS + e % 6] += v;
      }
      for (int idx = t; idx < 100; idx += nt) {
        const int a = idx / 10, b = idx % 10;
        double v = 0.0;
        for (int K = 0; K < F - 1; K++) v += Hc[K * 220 + 120 + idx];
        if (b <= a) A[bidx2(F, F) + a * kBS + b] += v;
        if (a == b) S.Hd[15 * F + a] += v;
      }
      if (t < 10) { double v = 0.0; for (int K = 0; K < F - 1; K++) v += gc[K * 10 + t]; S.g[15 * F + t] += v; }
    }
    __syncthreads();
  }
  if (t == 0) { st.cost_vis = p.c_lin[(size_t)w * kVisRec]; st.gmax_l = p.c_gmax[w]; st.x_cost = st.cost_vis + p.cost_nv[w]; if (st.iteration == 0) st.initial_cost = st.x_cost; }
  if (p.Sfull) {
    double* Sf = p.Sfull + (size_t)w * p.Ds * p.Ds;   // Dx x Dx, packed at the front of the window's slot
    for (int idx = t; idx < Dx * Dx; idx += nt) {
      const int a = idx / Dx, b = idx % Dx; const int hi = a >= b ? a : b, lo = a >= b ? b : a;
      Sf[idx] = stored(hi, lo) ? Ael2(hi, lo) : 0.0;
    }
    for (int i = t; i < Dx; i += nt) p.gfull[(size_t)w * p.Ds + i] = S.g[i] - S.gs[i];  // reduced gradient (rhs)
  }
  GF2_PC();
  // ---- Jacobi scaling (iteration 0), dogleg diagonal, gradient quantities
  const double mu = st.mu;
  double sums[3] = {0, 0, 0};  // dlg2, uEu, -
  double gmax = 0.0;
  for (int i = t; i < Dx; i += nt) {
    // constant / unused calibration entry: S.g[i] is zero by construction (no Jacobian column, no prior column) and is NOT written here —
    // thread F reads the calibration gradient below without a barrier in between
    if (i >= D && !calib_live(i - D)) { S.s[i] = 1.0; S.e[i] = 0.0; S.u[i] = 0.0; continue; }
    double sc;
    if (st.iteration == 0) { sc = 1.0 / (1.0 + sqrt(S.Hd[i])); p.sx[(size_t)w * p.Ds + i] = sc; } else sc = p.sx[(size_t)w * p.Ds + i];
    const double d2 = fmin(fmax(sc * sc * S.Hd[i], 1e-6), 1e32);
    const double e = d2 / (sc * sc);
    const double u = sc * sc * S.g[i] / d2;
    S.s[i] = sc; S.e[i] = e; S.u[i] = u;
    sums[0] += sc * sc * S.g[i] * S.g[i] / d2;
    sums[1] += e * u * u;
  }
  if (t < F) {  // gradient_max_norm = |x - Plus(x, -g)|_inf
    const double* x = p.pose + (size_t)w * F * 7 + 7 * t; const double* gg = &S.g[15 * t];
    for (int k = 0; k < 3; k++) gmax = fmax(gmax, fabs(gg[k]));
    Q4 q = ldq(x + 3); Q4 qn = qnormalized(qmul(q, deltaQ(mk3(-gg[3], -gg[4], -gg[5]))));
    gmax = fmax(gmax, fmax(fmax(fabs(q.x - qn.x), fabs(q.y - qn.y)), fmax(fabs(q.z - qn.z), fabs(q.w - qn.w))));
    for (int k = 6; k < 15; k++) gmax = fmax(gmax, fabs(gg[k]));
  } else if (t == F && p.wcal) {
    double gg[10];
    for (int k = 0; k < 10; k++) gg[k] = (k < 6 && ((p.wsub >> k) & 1u)) ? 0.0 : S.g[15 * F + k];   // Plus of the subset parameterization ignores these
    if (p.wcal & 1) {
      const double* x = p.exw + (size_t)w * 7;
      for (int k = 0; k < 3; k++) gmax = fmax(gmax, fabs(gg[k]));
      Q4 q = ldq(x + 3); Q4 qn = qnormalized(qmul(q, deltaQ(mk3(-gg[3], -gg[4], -gg[5]))));
      gmax = fmax(gmax, fmax(fmax(fabs(q.x - qn.x), fabs(q.y - qn.y)), fmax(fabs(q.z - qn.z), fabs(q.w - qn.w))));
    }
    for (int k = 6; k < 10; k++) if (calib_live(k)) gmax = fmax(gmax, fabs(gg[k]));
  }
  __syncthreads();
  for (int i = t; i < Dx; i += nt) { if (i >= D && !calib_live(i - D)) Ael2(i, i) = 1.0; else Ael2(i, i) += mu * S.e[i]; }
  __syncthreads();
  block_sum<3>(sums, S.red);
  gmax = warp_max(gmax);
  if (lane == 0) S.red[wid] = gmax;
  __syncthreads();
  if (t == 0) {
    double gm = 0; for (int i = 0; i < nwarp; i++) gm = fmax(gm, S.red[i]);
    st.dlg2_x = sums[0]; st.uEu_x = sums[1]; st.gmax_x = gm;
    S.flag = 0;
    if (fmax(gm, p.c_gmax[w]) <= p.gtol) { st.active = 0; st.termination = GF2_TERM_GRADIENT_TOL; S.flag = 2; }
  }
  __syncthreads();
  if (S.flag == 2) return;

  // ---- blocked Cholesky (right-looking over the frame blocks) with one-step look-ahead: warp 0 factors diagonal block
  // K+1 as soon as its own trailing pair (K+1, K+1) is done, while warps 1..7 finish the rest of trailing update K
  auto factor_diag = [&](int K) { if (factor_diag_block(A + bidx2(K, K), &S.dinv[16 * K], lane) && lane == 0) S.flag = 1; };  // by warp 0
  // right-hand side g - g_schur, kept per block with stride 16; it rides along the factorisation as one more row of every
  // panel (forward substitution for free): after block step K, S.fw[15 K ..] = (L^-1 rhs)_K
  for (int i = t; i < Dx; i += nt) S.z[16 * (i / 15) + i % 15] = S.g[i] - S.gs[i];
  if (wid == 0) factor_diag(0);
  __syncthreads();
#ifdef GF2_PHASE_CLOCKS
  long long cp_panel = 0, cp_w0 = 0, cp_trail = 0, cq0, cq1, cq2, cq3;
#endif
  for (int K = 0; K < Fx && !S.flag; K++) {
#ifdef GF2_PHASE_CLOCKS
    cq0 = clock64();
#endif
    const double* Akk = A + bidx2(K, K);
    // panel: X L_KK^T = A_IK, one thread per stored row of the block column (15 rows of the near block, 6 of each far one)
    // plus the 15 rows of the calibration block row, if there is one below
    const int nNear = K + 1 < F ? 15 : 0, nFar = 6 * (F - 2 - K > 0 ? F - 2 - K : 0), nCal = (Fc >= 0 && K < Fc) ? 15 : 0;
    const int prow = nNear + nFar + nCal;
    const int pI = t < nNear ? K + 1 : t < nNear + nFar ? K + 2 + (t - nNear) / 6 : Fc;
    const int prr = t < nNear ? t : t < nNear + nFar ? (t - nNear) % 6 : t - nNear - nFar;
    if (t < prow) { double* row = A + bidx2(pI, K) + prr * kBS; panel_row(row, Akk, &S.dinv[16 * K], row); }
    else if (t == prow) panel_row(&S.z[16 * K], Akk, &S.dinv[16 * K], &S.fw[15 * K]);  // the right-hand-side row
    __syncthreads();
    if (t < prow) {  // right-hand side of the rows below: rhs_I[rr] -= L_IK[rr, :] . fw_K
      const double* row = A + bidx2(pI, K) + prr * kBS;
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < 15; c++) acc += row[c] * S.fw[15 * K + c];
      S.z[16 * pI + prr] -= acc;
    }
    __syncthreads();
#ifdef GF2_PHASE_CLOCKS
    cq1 = clock64();
#endif
    // trailing update on the fp64 tensor cores: A_IJ -= L_IK L_JK^T for K < J <= I
    {
      const int m = Fx - 1 - K;
      const int npairs = m * (m + 1) / 2;
      if (wid == 0) {
        if (npairs > 0) { block_mma(A + bidx2(K + 1, K + 1), A + bidx2(K + 1, K), A + bidx2(K + 1, K), 15, 15, 15, lane); __syncwarp(); factor_diag(K + 1); }
      } else {
        for (int q = wid; q < npairs; q += nwarp - 1) {  // pairs 1 .. npairs-1 over warps 1..7 (pair 0 = (K+1, K+1) is warp 0's)
          int a = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);  // row of pair q in the lower triangle
          if ((a + 1) * (a + 2) / 2 <= q) a++; else if (a * (a + 1) / 2 > q) a--;
          const int b = q - a * (a + 1) / 2;
          const int I = K + 1 + a, J = K + 1 + b;
          block_mma(A + bidx2(I, J), A + bidx2(I, K), A + bidx2(J, K), brows(I, J), brows(I, K), brows(J, K), lane);
        }
      }
    }
#ifdef GF2_PHASE_CLOCKS
    cq2 = clock64();
#endif
    __syncthreads();
#ifdef GF2_PHASE_CLOCKS
    cq3 = clock64(); cp_panel += cq1 - cq0; cp_w0 += cq2 - cq1; cp_trail += cq3 - cq1;
#endif
  }

  if (S.flag == 1) {  // LINEAR_SOLVER_FAILURE -> invalid step (mu *= 10, re-linearise); DESIGN.md "deviations"
    if (t == 0) {
      st.mu *= 10.0; st.reuse = 0; st.iteration++; st.invalid_count++;
      if (st.invalid_count >= 5 || st.mu >= 1.0) { st.active = 0; st.termination = GF2_TERM_FAILURE; }
      else if (st.iteration >= p.max_iterations) { st.active = 0; st.termination = GF2_TERM_NO_CONVERGENCE; }
      st.lin_valid = 0;
    }
    return;
  }
  GF2_PC();
  // ---- yu = L^T u (needs the diagonal blocks as factored, before they are overwritten by their inverses)
  for (int col = t; col < Dx; col += nt) {
    const int J = col / 15, c = col % 15;
    double yu = 0;
    for (int I = J; I < Fx; I++) {
      const double* L = A + bidx2(I, J) + c;
      const int nr = brows(I, J);
      for (int r = 0; r < nr; r++) yu += L[r * kBS] * S.u[15 * I + r];
    }
    S.yu[col] = yu;
  }
  __syncthreads();
  GF2_PC();
  // ---- inverses of the diagonal blocks IN PLACE, all K in parallel (lane = column of L^-1, forward substitution down the rows)
  for (int K = wid; K < Fx; K += nwarp) {
    double* Akk = A + bidx2(K, K);
    double x[15];
    if (lane < 15) {
#pragma unroll
      for (int r = 0; r < 15; r++) {
        double acc = (r == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < r; k++) acc -= Akk[r * kBS + k] * x[k];
        x[r] = (r >= lane) ? acc * S.dinv[16 * K + r] : 0.0;
      }
    }
    __syncwarp();
    if (lane < 15) {
#pragma unroll
      for (int r = 0; r < 15; r++) Akk[r * kBS + lane] = (r >= lane) ? x[r] : 0.0;
    }
  }
  GF2_PC();
  // ---- z = L^-T fw: backward block substitution (the forward half was done inside the factorisation); the block solve by
  // warp 0, the column updates by one thread per column. z kept per block with stride 16.
  for (int i = t; i < Dx; i += nt) S.z[16 * (i / 15) + i % 15] = S.fw[i];
  __syncthreads();
  for (int K = Fx - 1; K >= 0; K--) {
    if (wid == 0) {  // z_K = L_KK^-T z_K (explicit inverse)
      const double* Li = A + bidx2(K, K);
      double v = 0;
      if (lane < 15) {
#pragma unroll
        for (int r = 0; r < 15; r++) v += (r >= lane) ? Li[r * kBS + lane] * S.z[16 * K + r] : 0.0;
      }
      __syncwarp();
      if (lane < 15) S.z[16 * K + lane] = v;
    }
    __syncthreads();
    if (t < 15 * K) {  // z_J -= L_KJ^T z_K for all J < K, one column per thread
      const int J = t / 15, c = t % 15;
      const double* Lc = A + bidx2(K, J) + c;
      double acc = 0.0;
      if (brows(K, J) == 6) {
#pragma unroll
        for (int r = 0; r < 6; r++) acc += Lc[r * kBS] * S.z[16 * K + r];
      } else {
#pragma unroll
        for (int r = 0; r < 15; r++) acc += Lc[r * kBS] * S.z[16 * K + r];
      }
      S.z[16 * J + c] -= acc;
    }
    __syncthreads();
  }
  GF2_PC();
  // quadratic forms of the model cost change through the factor: u^T S' u = |L^T u|^2, u^T S' z = (L^T u) . (L^T z),
  // z^T S' z = |L^T z|^2 with L^T z = the forward-substitution result
  double s3[3] = {0, 0, 0};  // uSu, uSz, zSz
  for (int col = t; col < Dx; col += nt) { const double yu = S.yu[col], yz = S.fw[col]; s3[0] += yu * yu; s3[1] += yu * yz; s3[2] += yz * yz; }
  block_sum<3>(s3, S.red);
  if (t == 0) { st.uSu = s3[0]; st.uSz = s3[1]; st.zSz = s3[2]; }
  double s2[4] = {0, 0, 0, 0};  // gn2, gz, zEz, uEz
  for (int i = t; i < Dx; i += nt) {
    const double z = S.z[16 * (i / 15) + i % 15], sc = S.s[i], e = S.e[i];
    const double d2 = e * sc * sc;
    s2[0] += d2 * z * z / (sc * sc);
    s2[1] += S.g[i] * z;
    s2[2] += e * z * z;
    s2[3] += e * S.u[i] * z;
    p.zx[(size_t)w * p.Ds + i] = z; p.ux[(size_t)w * p.Ds + i] = S.u[i]; p.ex_diag[(size_t)w * p.Ds + i] = e;
  }
  block_sum<4>(s2, S.red);
  if (t == 0) { st.gn2_x = s2[0]; st.gz_x = s2[1]; st.zEz_x = s2[2]; st.uEz_x = s2[3]; st.lin_valid = 1; }
  GF2_PC();
#ifdef GF2_PHASE_CLOCKS
  if (t == 0 && blockIdx.x == 300) { printf("k_solve2 phases:"); for (int i = 1; i < npc; i++) printf(" %lld", pc[i] - pc[i - 1]); printf(" total %lld\n", pc[npc - 1] - pc[0]); }
#endif
}

}  // namespace gf2
