// gf2_solver_lin.cuh — k_tasks (start-frame-uniform warp tasks) and k_linearize (sweep 1).
//
// Work decomposition of the Jacobian sweep: landmarks are grouped by host frame (start_frame); every group is cut into
// warp tasks of <= 32 landmarks, so all lanes of a warp evaluate, at step k, observations of the SAME frame pair
// (i, i + k).
//
// Moment form of the pose-block Hessian. Every Jacobian of ProjectionTwoFrameOneCamFactor
// (VE/factor/projectionTwoFrameOneCamFactor.cpp:102-140) factors through Jx = d r / d pts_w (2x3):
//     J_pose_i = Jx [ I | -[Xw - Pi]x Ri ],   J_pose_j = Jx [ -I | [d]x Rj ],   j_lambda = Jx dXw/dlambda,   d = Xw - Pj,
// and [Xw - Pi]x = [d]x + [Pj - Pi]x. With N = Jx^T Jx, h = Jx^T r and D = [d]x, the three 6x6 blocks (i,i), (i,j), (j,j)
// and both gradients of ALL factors of one frame pair follow from five sums over those factors,
//     M0 = sum N (6),  M1 = sum N D (9),  M2 = sum D^T N D (6),  h0 = sum h (3),  h1 = sum D^T h (3)      -- 27 doubles,
// and frame-only quantities (Ri, Rj, Pj - Pi). A lane therefore forms 27 numbers per observation (no 2x6 Jacobians, no
// 6x6 products); the warp reduces them with a transposed butterfly (31 shuffle-adds for 32 values, lane q ends up with
// total q) and adds them to the pair's moment record in shared memory; the blocks are expanded once per pair at the end
// of the kernel. Landmark columns w_l = J_pose^T j_lambda = [ -n ; Rj^T (n x d) ], n = Jx^T j_lambda, go, transposed, into
// a shared tile consumed by the fp64 tensor-core SYRK (mma.sync m8n8k4) that forms the Schur complement.
#pragma once
#include "gf2_solver_kernels.cuh"

namespace gf2 {

constexpr int kMaxTasks = 48;  // ceil(1000 / 32) + 11 partial tasks + slack
constexpr int kMaxPlaneTasks = 192;  // planes: up to ~5.8k per window in tasks of 32 of the same frame

// ------------------------------------------------------------------------------------------------ k_tasks
// Deterministic counting sort of the landmark table by start frame -> perm, and the warp-task list. One warp per start
// frame scans the table in order, 32 entries per step (the reference table is already sorted, VE/estimator/feature_manager.cpp:67-88;
// any order is accepted).
constexpr int kTaskThreads = 128;
__global__ void __launch_bounds__(kTaskThreads) k_tasks(KP p, int w0) {
  const int w = w0 + blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5, nw = kTaskThreads / 32, F = p.F;
  __shared__ int cnt[2 * kMaxF], ofs[2 * kMaxF + 1];
  __shared__ uint8_t key[kMaxPlaneTasks * 32];   // sort key of every table entry (start frame / 2 frame + ct), read from HBM once by the whole CTA
  static_assert(kMaxPlaneTasks * 32 >= GF2_MAX_LANDMARKS, "key buffer holds the landmark table");
  const int nlm = p.nlm[w];
  const int32_t* start = p.start + (size_t)w * p.Lm;
  for (int l = t; l < nlm; l += kTaskThreads) key[l] = (uint8_t)start[l];
  __syncthreads();
  // one warp per start frame, 32 table entries per step: counts and (below) ranks by ballot, in table order
  for (int f = wid; f < F; f += nw) {
    int c = 0;
    for (int base = 0; base < nlm; base += 32) { const int l = base + lane; c += __popc(__ballot_sync(0xffffffffu, l < nlm && key[l] == f)); }
    if (lane == 0) cnt[f] = c;
  }
  __syncthreads();
  if (t == 0) {
    int o = 0, nt = 0;
    int32_t* tf = p.task_first + (size_t)w * kMaxTasks; int32_t* tc = p.task_cnt + (size_t)w * kMaxTasks; int32_t* ts = p.task_start + (size_t)w * kMaxTasks;
    for (int s = 0; s < F; s++) {
      ofs[s] = o;
      for (int c = 0; c < cnt[s]; c += 32) { tf[nt] = o + c; tc[nt] = min(32, cnt[s] - c); ts[nt] = s; nt++; }
      o += cnt[s];
    }
    ofs[F] = o;
    p.ntasks[w] = nt;
  }
  __syncthreads();
  // packed record (landmark, track length, first observation, fixed) in sorted order: one load gives k_linearize everything
  // it needs to issue the dependent loads of a landmark
  {
    int4* info = p.lminfo + (size_t)w * p.Lm; const int32_t* tlen = p.tlen + (size_t)w * p.Lm; const int32_t* obeg = p.obeg + (size_t)w * p.Lm;
    const uint8_t* fixed = p.fixed + (size_t)w * p.Lm;
    for (int f = wid; f < F; f += nw) {
      int o = ofs[f];
      for (int base = 0; base < nlm; base += 32) {
        const int l = base + lane;
        const bool is = l < nlm && key[l] == f;
        const unsigned bal = __ballot_sync(0xffffffffu, is);
        if (is) info[o + __popc(bal & ((1u << lane) - 1u))] = make_int4(l, tlen[l], obeg[l], fixed[l] != 0);
        o += __popc(bal);
      }
    }
  }
  if (!p.planes) return;
  // LiDAR plane factors grouped by key = 2 * frame + ct (LidarPlaneNormFactor / CTLidarPlaneNormFactor), tasks of <= 32 planes
  __syncthreads();
  const int np = min(p.n_planes[w], kMaxPlaneTasks * 32);
  const gf2_plane* pls = p.planes + (size_t)w * p.Pm;
  for (int q = t; q < np; q += kTaskThreads) key[q] = (uint8_t)(2 * pls[q].frame + (pls[q].ct ? 1 : 0));
  __syncthreads();
  auto key_of = [&](int q) { return (int)key[q]; };
  for (int s = wid; s < 2 * F; s += nw) {
    int c = 0;
    for (int base = 0; base < np; base += 32) { const int q = base + lane; c += __popc(__ballot_sync(0xffffffffu, q < np && key_of(q) == s)); }
    if (lane == 0) cnt[s] = c;
  }
  __syncthreads();
  if (t == 0) {
    int o = 0, nt = 0;
    int32_t* tf = p.ptask_first + (size_t)w * kMaxPlaneTasks; int32_t* tc = p.ptask_cnt + (size_t)w * kMaxPlaneTasks; int32_t* ts = p.ptask_frame + (size_t)w * kMaxPlaneTasks;
    for (int s = 0; s < 2 * F; s++) {
      ofs[s] = o;
      for (int c = 0; c < cnt[s] && nt < kMaxPlaneTasks; c += 32) { tf[nt] = o + c; tc[nt] = min(32, cnt[s] - c); ts[nt] = s; nt++; }
      o += cnt[s];
    }
    p.nptasks[w] = nt;
  }
  __syncthreads();
  {
    int32_t* perm = p.pperm + (size_t)w * p.Pm;
    for (int s = wid; s < 2 * F; s += nw) {
      int o = ofs[s];
      for (int base = 0; base < np; base += 32) {
        const int q = base + lane;
        const bool is = q < np && key_of(q) == s;
        const unsigned bal = __ballot_sync(0xffffffffu, is);
        if (is) perm[o + __popc(bal & ((1u << lane) - 1u))] = q;
        o += __popc(bal);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ k_linearize
constexpr int kLinThreads = 128;   // 4 warps; 110 KB of shared memory -> two windows resident per SM
constexpr int kLinWarps = kLinThreads / 32;
constexpr int kWTRows = 68;        // 66 tangent rows + the landmark-gradient row 66 + one always-zero row 67 (tile padding)
constexpr int kWTStride = kLinThreads + 4;  // transposed W tile: [kWTRows][kWTStride] doubles
constexpr int kNBlkPairs = kMaxF * (kMaxF + 1) / 2;
constexpr int kNPairs = kMaxF * (kMaxF - 1) / 2;  // frame pairs i < j
constexpr int kMomStride = 28;     // upper part (m <= n, m < 6, n < 7) of the 7x7 Gram matrix of Y = [Jx | Jx [d]x | r]
constexpr int kYStride = 68;       // staging of Y^T per warp: [8][kYStride], rows 0..63 = the warp's residual rows, column 7 stays zero
constexpr int kSchurSlots = 12;    // Schur tiles owned by a warp (11, 11, 11, 12)
// SYRK tile grid: tile row a covers the rows r' = 8a .. 8a+7 of the END-ALIGNED index r' = r + kRowShift (r = tangent index 0..65, landmark
// gradient at r = 66; r = 67 and r' < kRowShift are zero padding). Every landmark's support ends at the last frame, so a round whose
// first host frame is i touches the tile rows a >= (6 i + kRowShift) / 8 only: 193 instead of 221 tile-rounds per W10-F1000 window.
constexpr int kRowShift = 4;

struct LinShared {
  FrameCtx fr[kMaxF];
  CamCtx cam;
  double g[kNVP];
  double Mom[kNPairs * kMomStride];  // per frame pair (i < j): the 27 moment sums
  double red[2 * 32];
  int task_first[kMaxTasks], task_cnt[kMaxTasks], task_start[kMaxTasks];
  union {
    double Y[kLinWarps][8 * kYStride];  // during the sweep
    double U[kNBlkPairs * 36];          // afterwards: pose-block Hessian of the visual factors, 6x6 blocks (bi <= bj), row-major inside
  };
  double WT[kWTRows * kWTStride];      // transposed, pre-scaled landmark columns  w_l / sqrt(v'_l); afterwards scratch
};
constexpr int kSumStride = 32;
constexpr int kSumsOfs = kNVP * kNVP;  // scratch in WT after the sweep: dense 72x72 Schur tiles, then the per-frame un-rotated sums
constexpr int kTermsOfs = kSumsOfs + kMaxF * kSumStride;   // then the F (F - 1) per-frame-pair terms of those sums
static_assert(kWTRows * kWTStride >= kTermsOfs + kMaxF * (kMaxF - 1) * 30, "scratch fits in WT");

__device__ __forceinline__ int ublk(int bi, int bj, int F) { return (bi * F - bi * (bi - 1) / 2 + (bj - bi)) * 36; }  // bi <= bj
__device__ __forceinline__ int pidx(int i, int j, int F) { return i * (2 * F - i - 1) / 2 + (j - i - 1); }             // i < j
__device__ __forceinline__ int momidx(int m, int n) { return 7 * m - m * (m - 1) / 2 + (n - m); }                      // m <= n

// Transposed butterfly: every lane brings 32 values; afterwards lane q holds the warp-wide sum of value q.
__device__ __forceinline__ double butterfly32(double (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int q = 0; q < off; q++) {
      const double send = up ? v[q] : v[q + off];
      const double keep = up ? v[q + off] : v[q];
      v[q] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ---- Schur SYRK tile ownership. The 45 tiles (a <= b) of the 9x9 grid are owned by diagonals d = b - a:
// warp 0: d = 0, 7; warp 1: d = 1, 6; warp 2: d = 2, 5; warp 3: d = 3, 4, 8  (11 / 11 / 11 / 12 tiles). Every warp then
// owns about a quarter of EVERY tile row a, so the suffix a >= a_min that a round touches stays balanced, and per k-step
// a warp loads each row fragment once (the A and B fragments of m8n8k4 have the same lane layout) for all its tiles.
struct TileAB { int a, b; };
__host__ __device__ constexpr TileAB tile_of(int W, int slot) {
  const int dg[4][3] = {{0, 7, -1}, {1, 6, -1}, {2, 5, -1}, {3, 4, 8}};
  int s = 0;
  for (int q = 0; q < 3; q++) {
    const int d = dg[W][q];
    if (d < 0) break;
    for (int a = 0; a + d < 9; a++) { if (s == slot) return TileAB{a, a + d}; s++; }
  }
  return TileAB{-1, -1};
}
template <int W, int M, int S>
__device__ __forceinline__ void syrk_tile(const double (&f)[9], double (&C)[kSchurSlots][2]) {
  constexpr TileAB tl = tile_of(W, S);
  if constexpr (tl.a >= 0 && tl.a >= M) mma_f64(C[S][0], C[S][1], f[tl.a], f[tl.b]);
}
// row fragment of tile row I: WT[8 I + fq - kRowShift][k0 + fk]; the first kRowShift rows of tile row 0 are padding
template <int M, int I, int STRIDE>
__device__ __forceinline__ void syrk_frag(double (&f)[9], const double* wb, bool lo_ok) {
  if constexpr (I >= M) {
    if constexpr (I == 0) f[0] = lo_ok ? wb[0] : 0.0;
    else f[I] = wb[I * 8 * STRIDE];
  }
}
// one round: C[tile] += sum over the COLS landmark columns of a W tile with row stride STRIDE, tiles with a >= M only
template <int W, int M, int STRIDE, int COLS>
__device__ __forceinline__ void syrk_round(const double* wb /* &WT[fq - kRowShift][fk] */, bool lo_ok, double (&C)[kSchurSlots][2]) {
#pragma unroll (COLS >= 128 ? 4 : 2)
  for (int k0 = 0; k0 < COLS; k0 += 4) {
    double f[9];
    syrk_frag<M, 0, STRIDE>(f, wb + k0, lo_ok); syrk_frag<M, 1, STRIDE>(f, wb + k0, lo_ok); syrk_frag<M, 2, STRIDE>(f, wb + k0, lo_ok);
    syrk_frag<M, 3, STRIDE>(f, wb + k0, lo_ok); syrk_frag<M, 4, STRIDE>(f, wb + k0, lo_ok); syrk_frag<M, 5, STRIDE>(f, wb + k0, lo_ok);
    syrk_frag<M, 6, STRIDE>(f, wb + k0, lo_ok); syrk_frag<M, 7, STRIDE>(f, wb + k0, lo_ok); syrk_frag<M, 8, STRIDE>(f, wb + k0, lo_ok);
    syrk_tile<W, M, 0>(f, C); syrk_tile<W, M, 1>(f, C); syrk_tile<W, M, 2>(f, C); syrk_tile<W, M, 3>(f, C);
    syrk_tile<W, M, 4>(f, C); syrk_tile<W, M, 5>(f, C); syrk_tile<W, M, 6>(f, C); syrk_tile<W, M, 7>(f, C);
    syrk_tile<W, M, 8>(f, C); syrk_tile<W, M, 9>(f, C); syrk_tile<W, M, 10>(f, C); syrk_tile<W, M, 11>(f, C);
  }
}
template <int W, int STRIDE, int COLS>
__device__ __forceinline__ void syrk_warp(int a_min, const double* wb, bool lo_ok, double (&C)[kSchurSlots][2]) {
  switch (a_min) {
    case 0: syrk_round<W, 0, STRIDE, COLS>(wb, lo_ok, C); break;
    case 1: syrk_round<W, 1, STRIDE, COLS>(wb, lo_ok, C); break;
    case 2: syrk_round<W, 2, STRIDE, COLS>(wb, lo_ok, C); break;
    case 3: syrk_round<W, 3, STRIDE, COLS>(wb, lo_ok, C); break;
    case 4: syrk_round<W, 4, STRIDE, COLS>(wb, lo_ok, C); break;
    case 5: syrk_round<W, 5, STRIDE, COLS>(wb, lo_ok, C); break;
    case 6: syrk_round<W, 6, STRIDE, COLS>(wb, lo_ok, C); break;
    case 7: syrk_round<W, 7, STRIDE, COLS>(wb, lo_ok, C); break;
    default: syrk_round<W, 8, STRIDE, COLS>(wb, lo_ok, C); break;
  }
}
// accumulator tiles -> dense 72x72 (upper tiles, end-aligned index) in shared memory
template <int W, int S>
__device__ __forceinline__ void syrk_store_tile(double* dense, const double (&C)[kSchurSlots][2], int fq, int fk) {
  constexpr TileAB tl = tile_of(W, S);
  if constexpr (tl.a >= 0) { double* d = dense + (8 * tl.a + fq) * kNVP + 8 * tl.b + 2 * fk; d[0] = C[S][0]; d[1] = C[S][1]; }
}
template <int W>
__device__ __forceinline__ void syrk_store(double* dense, const double (&C)[kSchurSlots][2], int fq, int fk) {
  syrk_store_tile<W, 0>(dense, C, fq, fk); syrk_store_tile<W, 1>(dense, C, fq, fk); syrk_store_tile<W, 2>(dense, C, fq, fk);
  syrk_store_tile<W, 3>(dense, C, fq, fk); syrk_store_tile<W, 4>(dense, C, fq, fk); syrk_store_tile<W, 5>(dense, C, fq, fk);
  syrk_store_tile<W, 6>(dense, C, fq, fk); syrk_store_tile<W, 7>(dense, C, fq, fk); syrk_store_tile<W, 8>(dense, C, fq, fk);
  syrk_store_tile<W, 9>(dense, C, fq, fk); syrk_store_tile<W, 10>(dense, C, fq, fk); syrk_store_tile<W, 11>(dense, C, fq, fk);
}

// ---- expansion of the frame-pair moments (header comment)
struct PairMom { M3 M0, M1, M2; V3 h0, h1; };
__device__ __forceinline__ PairMom load_mom(const double* m) {
  PairMom q;
  q.M0.m[0] = m[momidx(0, 0)]; q.M0.m[1] = q.M0.m[3] = m[momidx(0, 1)]; q.M0.m[2] = q.M0.m[6] = m[momidx(0, 2)];
  q.M0.m[4] = m[momidx(1, 1)]; q.M0.m[5] = q.M0.m[7] = m[momidx(1, 2)]; q.M0.m[8] = m[momidx(2, 2)];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) q.M1.m[a * 3 + b] = m[momidx(a, 3 + b)];
  q.M2.m[0] = m[momidx(3, 3)]; q.M2.m[1] = q.M2.m[3] = m[momidx(3, 4)]; q.M2.m[2] = q.M2.m[6] = m[momidx(3, 5)];
  q.M2.m[4] = m[momidx(4, 4)]; q.M2.m[5] = q.M2.m[7] = m[momidx(4, 5)]; q.M2.m[8] = m[momidx(5, 5)];
  q.h0 = mk3(m[momidx(0, 6)], m[momidx(1, 6)], m[momidx(2, 6)]);
  q.h1 = mk3(m[momidx(3, 6)], m[momidx(4, 6)], m[momidx(5, 6)]);
  return q;
}
__device__ __forceinline__ M3 ldm3(const double* p) { M3 m; for (int i = 0; i < 9; i++) m.m[i] = p[i]; return m; }
__device__ __forceinline__ void put3s(double* blk, int r0, int c0, const M3& m, double sgn) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) blk[(r0 + r) * 6 + c0 + c] = sgn * m.m[r * 3 + c];
}
// off-diagonal block (i, j), i < j, sub-block sub = 0: pp, 1: p-theta, 2: theta-p, 3: theta-theta
__device__ __forceinline__ void expand_offdiag(const double* mom, const FrameCtx& fi, const FrameCtx& fj, int sub, double* B) {
  const PairMom q = load_mom(mom);
  const M3 Ri = ldm3(fi.R), Rj = ldm3(fj.R);
  if (sub == 0) { put3s(B, 0, 0, q.M0, -1.0); return; }
  if (sub == 1) { put3s(B, 0, 3, mul(q.M1, Rj), 1.0); return; }
  const M3 C = skew(mk3(fj.P[0] - fi.P[0], fj.P[1] - fi.P[1], fj.P[2] - fi.P[2]));
  if (sub == 2) { put3s(B, 3, 0, mulT(Ri, add(transpose(q.M1), mulT(C, q.M0))), 1.0); return; }
  put3s(B, 3, 3, mulT(Ri, mul(add(q.M2, mulT(C, q.M1)), Rj)), -1.0);
}
// un-rotated sums of frame f over all pairs it takes part in; grp 0: PP (sym 6) | GP (3) | GT (3), grp 1: PT (9), grp 2: TT (9)
// so that block (f,f) = [[PP, -PT Rf], [., Rf^T TT Rf]] and g_f = [GP ; Rf^T GT]; out: PP 0..5 | GP 6..8 | GT 9..11 | PT 12..20 | TT 21..29
__device__ __forceinline__ void expand_diag_sums(const double* Mom, const FrameCtx* fr, int f, int F, int grp, double* out /* kSumStride */) {
  M3 acc = zero3(); V3 gp = mk3(0, 0, 0), gt = mk3(0, 0, 0);
  for (int o = 0; o < F; o++) {
    if (o == f) continue;
    const bool irole = f < o;  // f is the host frame of the pair
    const PairMom q = load_mom(Mom + (irole ? pidx(f, o, F) : pidx(o, f, F)) * kMomStride);
    if (!irole) {
      if (grp == 0) { acc = add(acc, q.M0); gp = gp - q.h0; gt = gt + q.h1; }
      else if (grp == 1) acc = add(acc, q.M1);
      else acc = add(acc, q.M2);
    } else {
      const M3 C = skew(mk3(fr[o].P[0] - fr[f].P[0], fr[o].P[1] - fr[f].P[1], fr[o].P[2] - fr[f].P[2]));
      if (grp == 0) { acc = add(acc, q.M0); gp = gp + q.h0; gt = gt - (q.h1 + mulT(C, q.h0)); }
      else if (grp == 1) acc = add(acc, add(q.M1, mul(q.M0, C)));
      else { const M3 CtM1 = mulT(C, q.M1); acc = add(acc, add(add(q.M2, CtM1), add(transpose(CtM1), mulT(C, mul(q.M0, C))))); }
    }
  }
  if (grp == 0) {
    out[0] = acc.m[0]; out[1] = acc.m[1]; out[2] = acc.m[2]; out[3] = acc.m[4]; out[4] = acc.m[5]; out[5] = acc.m[8];
    out[6] = gp.x; out[7] = gp.y; out[8] = gp.z; out[9] = gt.x; out[10] = gt.y; out[11] = gt.z;
  } else {
    for (int c = 0; c < 9; c++) out[(grp == 1 ? 12 : 21) + c] = acc.m[c];
  }
}

// The same sums split by the other frame o, so that the F (F - 1) (frame, other frame) terms spread over the CTA instead of 3 F threads walking
// F - 1 frame pairs each (those 33 threads were most of the kernel's tail): term of frame o in the un-rotated sums of frame f, all three
// groups, in the layout of expand_diag_sums (30 doubles). expand_diag_reduce adds the terms of a frame in ascending o, the order of the
// loop above: the result is the same to the last bit, and does not depend on the number of threads.
constexpr int kDiagTerm = 30;
__device__ __forceinline__ void expand_diag_term(const double* Mom, const FrameCtx* fr, int f, int o, int F, double* out /* kDiagTerm */) {
  const bool irole = f < o;  // f is the host frame of the pair
  const PairMom q = load_mom(Mom + (irole ? pidx(f, o, F) : pidx(o, f, F)) * kMomStride);
  M3 pp, pt, tt; V3 gp, gt;
  if (!irole) { pp = q.M0; gp = mk3(0, 0, 0) - q.h0; gt = q.h1; pt = q.M1; tt = q.M2; }
  else {
    const M3 C = skew(mk3(fr[o].P[0] - fr[f].P[0], fr[o].P[1] - fr[f].P[1], fr[o].P[2] - fr[f].P[2]));
    pp = q.M0; gp = q.h0; gt = mk3(0, 0, 0) - (q.h1 + mulT(C, q.h0));
    pt = add(q.M1, mul(q.M0, C));
    const M3 CtM1 = mulT(C, q.M1); tt = add(add(q.M2, CtM1), add(transpose(CtM1), mulT(C, mul(q.M0, C))));
  }
  out[0] = pp.m[0]; out[1] = pp.m[1]; out[2] = pp.m[2]; out[3] = pp.m[4]; out[4] = pp.m[5]; out[5] = pp.m[8];
  out[6] = gp.x; out[7] = gp.y; out[8] = gp.z; out[9] = gt.x; out[10] = gt.y; out[11] = gt.z;
#pragma unroll
  for (int c = 0; c < 9; c++) { out[12 + c] = pt.m[c]; out[21 + c] = tt.m[c]; }
}
// term index of (f, o), o != f
__device__ __forceinline__ int diag_term_idx(int f, int o, int F) { return (f * (F - 1) + (o < f ? o : o - 1)) * kDiagTerm; }
// entry e of the sums of frame f = the terms of all other frames, ascending
__device__ __forceinline__ double expand_diag_reduce(const double* terms, int f, int e, int F) {
  double acc = 0.0;
  for (int o = 0; o < F; o++) if (o != f) acc += terms[diag_term_idx(f, o, F) + e];
  return acc;
}

__global__ void __launch_bounds__(kLinThreads, 2) k_linearize(KP p, int w0) {
  const int w = w0 + blockIdx.x;
  WinState& st = p.st[w];
  if (!st.active || st.reuse) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LinShared& S = *reinterpret_cast<LinShared*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int F = p.F, NV = 6 * F;
  const double* pose = p.pose + (size_t)w * F * 7;
  build_frames(pose, p.ex + (size_t)w * 7, p.td[w], F, S.fr, &S.cam);
  for (int i = t; i < kNPairs * kMomStride; i += kLinThreads) S.Mom[i] = 0.0;
  for (int i = t; i < kLinWarps * 8 * kYStride; i += kLinThreads) (&S.Y[0][0])[i] = 0.0;
  if (t < kNVP) S.g[t] = 0.0;
  const int ntasks = p.ntasks[w];
  if (t < kMaxTasks) { S.task_first[t] = p.task_first[(size_t)w * kMaxTasks + t]; S.task_cnt[t] = p.task_cnt[(size_t)w * kMaxTasks + t]; S.task_start[t] = p.task_start[(size_t)w * kMaxTasks + t]; }
  for (int i = t; i < kWTRows * kWTStride; i += kLinThreads) S.WT[i] = 0.0;
  __syncthreads();

  const int4* lminfo = p.lminfo + (size_t)w * p.Lm;
  const float4* obs = p.obs + (size_t)w * p.Om;
  const double* ftd = p.frame_td + (size_t)w * F;
  const double mu = st.mu;
  const bool it0 = (st.iteration == 0);
  const double sqi = p.sqrt_info_px;
  double cost_acc = 0.0, gmax = 0.0;
  double C[kSchurSlots][2];  // Schur accumulators of the tiles this warp owns (tile_of)
#pragma unroll
  for (int q = 0; q < kSchurSlots; q++) { C[q][0] = 0.0; C[q][1] = 0.0; }
  const int fq = lane >> 2, fk = lane & 3;  // fragment coordinates: row/col index 0..7, k index 0..3
  double* Yw = S.Y[wid];
  const double* wb = &S.WT[(fq - kRowShift) * kWTStride + fk];
  const bool lo_ok = fq >= kRowShift;

  // landmark record of a lane for one round, fetched ahead in two levels so that the dependent global loads are in flight
  // during the SYRK of earlier rounds: the packed table entry two rounds ahead, the landmark's state and first observations
  // one round ahead
  struct LaneLm { int l, L, ob; bool have, fx; double lam, s_l; float4 oi, o1; };
  auto fetch_info = [&](int tb) {
    const int task = tb + wid;
    return (task < ntasks && lane < S.task_cnt[task]) ? lminfo[S.task_first[task] + lane] : make_int4(0, 0, 0, -1);  // w = -1: idle lane
  };
  auto fetch_lm = [&](int4 info) {
    LaneLm q; q.l = info.x; q.L = info.y; q.ob = info.z; q.fx = info.w > 0; q.have = info.w >= 0; q.lam = 1.0; q.s_l = 1.0; q.oi = make_float4(0, 0, 0, 0); q.o1 = q.oi;
    if (q.have) {
      q.lam = p.invdep[(size_t)w * p.Lm + q.l];
      if (!it0) q.s_l = p.lm_s[(size_t)w * p.Lm + q.l];
      q.oi = obs[q.ob];
    }
    // observation of the first step this warp will take (the step order is rotated per warp, see below)
    const int nk = __reduce_max_sync(0xffffffffu, q.L) - 1;
    if (q.have && q.L > 1) q.o1 = obs[q.ob + min(1 + (wid * nk) / kLinWarps, q.L - 1)];
    return q;
  };
#ifdef GF2_PHASE_CLOCKS
  long long lc_t0 = clock64(), lc_step = 0, lc_syrk = 0, lc_a, lc_b, lc_c;
#endif
  LaneLm nxt = fetch_lm(fetch_info(0));
  int4 info_n = fetch_info(kLinWarps);
  for (int tbase = 0; tbase < ntasks; tbase += kLinWarps) {
#ifdef GF2_PHASE_CLOCKS
    lc_a = clock64();
#endif
    const int task = tbase + wid;
    const bool have_task = task < ntasks;
    const int i = have_task ? S.task_start[task] : 0;
    const LaneLm cur = nxt;
    const bool have = cur.have;
    const int a_min = (6 * S.task_start[tbase] + kRowShift) >> 3;  // tile rows above the round's first host frame are skipped by the SYRK
    const int row0 = max(8 * a_min - kRowShift, 0);
    const int l = cur.l, L = cur.L, ob = cur.ob; const bool fx = cur.fx;
    LmCtx lc; lc.Xw = mk3(0, 0, 0); lc.dXdl = mk3(0, 0, 0);
    if (have) landmark_ctx(S.fr[i], S.cam, cur.oi, ftd[i], cur.lam, lc);
    // rows of this thread's column that the step loop below will not write: everything for an idle lane, else the rows
    // left of the host frame and right of the track's last frame
    {
      const int z0 = have ? 6 * i : kWTRows - 1, z1 = have ? 6 * (i + L) : kWTRows - 1;
      for (int c = row0; c < z0; c++) S.WT[c * kWTStride + t] = 0.0;
      for (int c = z1; c < kWTRows - 1; c++) S.WT[c * kWTStride + t] = 0.0;
    }
    V3 ns = mk3(0, 0, 0);  // sum_k n_k  ->  w_i = [ ns ; Ri^T ((Xw - Pi) x ns) ]
    double v = 0.0, gl = 0.0;
    const int Lmax = __reduce_max_sync(0xffffffffu, L);
    float4 oj_n = cur.o1;
    const double* yf = Yw + fq * kYStride + fk;
    // One observation per lane, branch-free (idle lanes compute on dummy data and contribute zeros) so that it shares a basic
    // block with the tensor-core Gram of the PREVIOUS step and the scheduler can interleave the two: this lane's two rows of
    // Y = [Jx | Jx [d]x | r], the landmark sums and the column w_j.
    auto obs_math = [&](int k, double (&y0)[7], double (&y1)[7]) {
      const int j = i + k;  // uniform across the warp
      const bool valid = have && k < L;
      const float4 oj = oj_n;
      oj_n = obs[ob + (have ? min(k == Lmax - 1 ? 1 : k + 1, L - 1) : 0)];  // next step's observation: in flight during this step (index clamped, no branch)
      const FrameCtx& fj = S.fr[j];
      const double dx = lc.Xw.x - fj.P[0], dy = lc.Xw.y - fj.P[1], dz = lc.Xw.z - fj.P[2];
      const double px = fj.A[0] * dx + fj.A[1] * dy + fj.A[2] * dz - S.cam.rtt[0];
      const double py = fj.A[3] * dx + fj.A[4] * dy + fj.A[5] * dz - S.cam.rtt[1];
      const double pzr = fj.A[6] * dx + fj.A[7] * dy + fj.A[8] * dz - S.cam.rtt[2];
      const double dt = S.cam.td - ftd[j];
      const double iz = fast_rcp(valid ? pzr : 1.0);
      double r0 = sqi * (px * iz - ((double)oj.x - dt * (double)oj.z));
      double r1 = sqi * (py * iz - ((double)oj.y - dt * (double)oj.w));
      // ceres::HuberLoss + Corrector as in huber(), branch-free: outside the inlier region r = sqrt(sq), scale = sqrt(delta / r)
      const double sq = r0 * r0 + r1 * r1, hb = p.huber * p.huber;
      const bool outl = sq > hb;
      const double ir = fast_rsqrt(fmax(sq, hb)), yy = fmax(2.2250738585072014e-308, p.huber * ir);
      const double sc = outl ? yy * fast_rsqrt(yy) : 1.0;
      const double hr = outl ? 0.5 * (2.0 * p.huber * (sq * ir) - hb) : 0.5 * sq;
      const double msk = valid ? 1.0 : 0.0;
      cost_acc += msk * hr;
      r0 *= sc * msk; r1 *= sc * msk;
      // Jx = sc * sqrt_info * [[1/z, 0, -x/z^2], [0, 1/z, -y/z^2]] * A_j
      const double a = msk * sc * sqi * iz, bx = -a * px * iz, by = -a * py * iz;
      const double j00 = a * fj.A[0] + bx * fj.A[6], j01 = a * fj.A[1] + bx * fj.A[7], j02 = a * fj.A[2] + bx * fj.A[8];
      const double j10 = a * fj.A[3] + by * fj.A[6], j11 = a * fj.A[4] + by * fj.A[7], j12 = a * fj.A[5] + by * fj.A[8];
      y0[0] = j00; y0[1] = j01; y0[2] = j02; y1[0] = j10; y1[1] = j11; y1[2] = j12;
      // (row [d]x)_c: (b1 dz - b2 dy, b2 dx - b0 dz, b0 dy - b1 dx)
      y0[3] = j01 * dz - j02 * dy; y0[4] = j02 * dx - j00 * dz; y0[5] = j00 * dy - j01 * dx;
      y1[3] = j11 * dz - j12 * dy; y1[4] = j12 * dx - j10 * dz; y1[5] = j10 * dy - j11 * dx;
      y0[6] = r0; y1[6] = r1;
      // landmark column: j_lambda = Jx dXw/dlambda (zero for a fixed landmark), n = Jx^T j_lambda, w_j = [ -n ; Rj^T (n x d) ]
      const double fm = fx ? 0.0 : 1.0;
      const double jl0 = fm * (j00 * lc.dXdl.x + j01 * lc.dXdl.y + j02 * lc.dXdl.z);
      const double jl1 = fm * (j10 * lc.dXdl.x + j11 * lc.dXdl.y + j12 * lc.dXdl.z);
      v += jl0 * jl0 + jl1 * jl1; gl += jl0 * r0 + jl1 * r1;
      const double nx = j00 * jl0 + j10 * jl1, ny = j01 * jl0 + j11 * jl1, nz = j02 * jl0 + j12 * jl1;
      ns.x += nx; ns.y += ny; ns.z += nz;
      const double qx = ny * dz - nz * dy, qy = nz * dx - nx * dz, qz = nx * dy - ny * dx;
      // unconditional: an idle lane stores the zeros these rows of its column hold anyway
      double* wt = &S.WT[(6 * j) * kWTStride + t];
      wt[0] = -nx; wt[kWTStride] = -ny; wt[2 * kWTStride] = -nz;
      wt[3 * kWTStride] = fj.R[0] * qx + fj.R[3] * qy + fj.R[6] * qz;
      wt[4 * kWTStride] = fj.R[1] * qx + fj.R[4] * qy + fj.R[7] * qz;
      wt[5 * kWTStride] = fj.R[2] * qx + fj.R[5] * qy + fj.R[8] * qz;
    };
    // staging of the warp's 64 rows (transposed) and the fragments of the Gram product: the A and B fragment of a k-step are
    // the same element of Y^T (column 7 of Y is never written and stays zero)
    auto stage = [&](const double (&y0)[7], const double (&y1)[7], double (&fa)[16]) {
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 7; c++) { Yw[c * kYStride + lane] = y0[c]; Yw[c * kYStride + 32 + lane] = y1[c]; }
      __syncwarp();
#pragma unroll
      for (int s = 0; s < 16; s++) fa[s] = yf[4 * s];
    };
    // Gram matrix Y^T Y on the tensor cores: one 8x8 tile, K = 64, four independent accumulator chains
    auto gram = [&](const double (&fa)[16], double& g0, double& g1) {
      double c0 = 0, c1 = 0, e0 = 0, e1 = 0, c2 = 0, c3 = 0, e2 = 0, e3 = 0;
#pragma unroll
      for (int s = 0; s < 16; s += 4) { mma_f64(c0, c1, fa[s], fa[s]); mma_f64(e0, e1, fa[s + 1], fa[s + 1]); mma_f64(c2, c3, fa[s + 2], fa[s + 2]); mma_f64(e2, e3, fa[s + 3], fa[s + 3]); }
      g0 = (c0 + e0) + (c2 + e2); g1 = (c1 + e1) + (c3 + e3);
    };
    auto flush = [&](int j, double g0, double g1) {
      if (fq < 6) {
        double* mom = &S.Mom[pidx(i, j, F) * kMomStride + momidx(fq, fq)] - fq;  // entry (fq, n) at mom[n]
        if (2 * fk >= fq) atomicAdd(&mom[2 * fk], g0);
        if (2 * fk + 1 >= fq && fk < 3) atomicAdd(&mom[2 * fk + 1], g1);
      }
    };
    if (have_task && Lmax > 1) {
      double y0[7], y1[7], fa[16], g0, g1;
      // the warps of a round usually share the host frame: each starts at a different step so that their atomic flushes
      // hit different frame pairs
      const int nk = Lmax - 1;
      int k = 1 + (wid * nk) / kLinWarps, kp;
      obs_math(k, y0, y1);
      stage(y0, y1, fa);
      for (int q = 1; q < nk; q++) {
        kp = k; k = k == nk ? 1 : k + 1;
        gram(fa, g0, g1);      // the tensor cores reduce the previous step ...
        obs_math(k, y0, y1);   // ... while this step runs on the fp64 FMA pipe
        flush(i + kp, g0, g1);
        stage(y0, y1, fa);
      }
      gram(fa, g0, g1);
      flush(i + k, g0, g1);
    }
    // landmark scalars: jacobi scale (iteration 0), regularised v' = v + mu * e; the landmark's column of W is scaled by
    // 1/sqrt(v') so that the SYRK below needs no per-element multiply
    if (have) {
      double s_l;
      if (it0) { s_l = 1.0 / (1.0 + sqrt(v)); p.lm_s[(size_t)w * p.Lm + l] = s_l; } else s_l = cur.s_l;
      const double d2 = fmin(fmax(s_l * s_l * v, 1e-6), 1e32);
      const double e = d2 / (s_l * s_l);
      const double vp = v + mu * e;
      const double rs = (!fx && v > 0.0) ? rsqrt(vp) : 0.0;
      p.lm_v[(size_t)w * p.Lm + l] = fx ? 0.0 : v;
      p.lm_g[(size_t)w * p.Lm + l] = fx ? 0.0 : gl;
      if (!fx) gmax = fmax(gmax, fabs(gl));
      const FrameCtx& fi = S.fr[i];
      const V3 e_i = mk3(lc.Xw.x - fi.P[0], lc.Xw.y - fi.P[1], lc.Xw.z - fi.P[2]);
      const V3 q = cross(e_i, ns);  // Gi^T ns = Ri^T ((Xw - Pi) x ns)
      double* wt = &S.WT[(6 * i) * kWTStride + t];
      wt[0] = ns.x * rs; wt[kWTStride] = ns.y * rs; wt[2 * kWTStride] = ns.z * rs;
      wt[3 * kWTStride] = (fi.R[0] * q.x + fi.R[3] * q.y + fi.R[6] * q.z) * rs;
      wt[4 * kWTStride] = (fi.R[1] * q.x + fi.R[4] * q.y + fi.R[7] * q.z) * rs;
      wt[5 * kWTStride] = (fi.R[2] * q.x + fi.R[5] * q.y + fi.R[8] * q.z) * rs;
      for (int c = 6 * (i + 1); c < 6 * (i + L); c++) S.WT[c * kWTStride + t] *= rs;
      S.WT[66 * kWTStride + t] = fx ? 0.0 : gl * rs;
    } else {
      S.WT[66 * kWTStride + t] = 0.0;
    }
    nxt = fetch_lm(info_n);
    info_n = fetch_info(tbase + 2 * kLinWarps);
    __syncthreads();
#ifdef GF2_PHASE_CLOCKS
    lc_b = clock64();
#endif
    // Schur SYRK on the fp64 tensor cores: C[a-tile][b-tile] += sum_l Ws[l][a] * Ws[l][b]
    {
      switch (wid) {
        case 0: syrk_warp<0, kWTStride, kLinThreads>(a_min, wb, lo_ok, C); break;
        case 1: syrk_warp<1, kWTStride, kLinThreads>(a_min, wb, lo_ok, C); break;
        case 2: syrk_warp<2, kWTStride, kLinThreads>(a_min, wb, lo_ok, C); break;
        default: syrk_warp<3, kWTStride, kLinThreads>(a_min, wb, lo_ok, C); break;
      }
    }
    __syncthreads();
#ifdef GF2_PHASE_CLOCKS
    lc_c = clock64(); lc_step += lc_b - lc_a; lc_syrk += lc_c - lc_b;
#endif
  }
#ifdef GF2_PHASE_CLOCKS
  const long long lc_t1 = clock64();
#endif

  // Schur tiles -> dense 72x72 (upper tiles) in WT, which is free now
  switch (wid) {
    case 0: syrk_store<0>(S.WT, C, fq, fk); break;
    case 1: syrk_store<1>(S.WT, C, fq, fk); break;
    case 2: syrk_store<2>(S.WT, C, fq, fk); break;
    default: syrk_store<3>(S.WT, C, fq, fk); break;
  }
  // frame-pair moments -> off-diagonal blocks (i, j) and the un-rotated per-frame sums
  {
    const int npairs = F * (F - 1) / 2, nterms = F * (F - 1);
    double* terms = S.WT + kTermsOfs;
    for (int q = t; q < nterms + 4 * npairs; q += kLinThreads) {
      if (q < nterms) { const int f = q / (F - 1), r = q - f * (F - 1), o = r < f ? r : r + 1; expand_diag_term(S.Mom, S.fr, f, o, F, terms + diag_term_idx(f, o, F)); continue; }
      const int q2 = q - nterms, pr = q2 >> 2, sub = q2 & 3;
      int i = 0, rem = pr;
      while (rem >= F - 1 - i) { rem -= F - 1 - i; i++; }
      const int j = i + 1 + rem;
      expand_offdiag(&S.Mom[pr * kMomStride], S.fr[i], S.fr[j], sub, &S.U[ublk(i, j, F)]);
    }
  }
  __syncthreads();
  {
    double* sums = S.WT + kSumsOfs;
    for (int q = t; q < F * kDiagTerm; q += kLinThreads) { const int f = q / kDiagTerm, e = q - f * kDiagTerm; sums[f * kSumStride + e] = expand_diag_reduce(S.WT + kTermsOfs, f, e, F); }
  }
  __syncthreads();
  // diagonal blocks (f, f) = [[PP, -PT Rf], [., Rf^T TT Rf]] and the gradient g_f = [GP ; Rf^T GT]
  {
    const double* sums = S.WT + kSumsOfs;
    for (int q = t; q < 42 * F; q += kLinThreads) {
      const int f = q / 42, e = q - 42 * f;
      const double* sf = sums + f * kSumStride;
      const double* R = S.fr[f].R;
      if (e >= 36) {
        const int c = e - 36;
        S.g[6 * f + c] = c < 3 ? sf[6 + c] : R[c - 3] * sf[9] + R[3 + c - 3] * sf[10] + R[6 + c - 3] * sf[11];
        continue;
      }
      int r = e / 6, c = e - 6 * r;
      if (r > c) { const int x = r; r = c; c = x; }  // symmetric: evaluate the upper element
      double val;
      if (c < 3) { const int lo = r, hi = c; val = sf[lo == 0 ? hi : (lo == 1 ? 2 + hi : 5)]; }  // PP sym: (00 01 02 11 12 22)
      else if (r < 3) { const double* pt = sf + 12 + 3 * r; const int cc = c - 3; val = -(pt[0] * R[cc] + pt[1] * R[3 + cc] + pt[2] * R[6 + cc]); }
      else {
        const double* tt = sf + 21; const int rr = r - 3, cc = c - 3;
        double acc = 0;
#pragma unroll
        for (int a = 0; a < 3; a++) acc += R[3 * a + rr] * (tt[3 * a] * R[cc] + tt[3 * a + 1] * R[3 + cc] + tt[3 * a + 2] * R[6 + cc]);
        val = acc;
      }
      S.U[ublk(f, f, F) + e] = val;
    }
  }
  __syncthreads();

  // LiDAR plane factors: tasks of <= 32 planes of one key (frame, CT or not). The sums [J | r]^T [J | r] over a task's planes are a Gram matrix
  // with K = 32: the lanes stage their row [J | r] (7 columns; 13 for a CT factor: begin pose, end pose, residual) transposed in shared memory
  // and the warp forms the 8x8 tile(s) on the fp64 tensor cores (8 mma per tile), accumulating over CONSECUTIVE tasks of the same key in
  // registers; the tile entries go to the pose blocks (f, f), (f, f + 1), (f + 1, f + 1) and the gradient by shared atomics only when the
  // key changes. (A 27-value butterfly per block and task cost ~4x the instructions: 34 % of the sweep of a 5,000-plane window.)
  if (p.planes) {
    const int npt = p.nptasks[w];
    const gf2_plane* pls = p.planes + (size_t)w * p.Pm;
    const int32_t* pperm = p.pperm + (size_t)w * p.Pm;
    constexpr int kPS = 36;                                    // staging row stride (4 mod 16: conflict-free fragment loads)
    double* stg = S.WT + kTermsOfs + wid * 13 * kPS;           // the frame-pair terms are summed: their space stages the planes
    static_assert(kLinWarps * 13 * kPS <= kMaxF * (kMaxF - 1) * 30, "plane staging fits in the terms area");
    const double* poseW = p.pose + (size_t)w * F * 7;
    const double* palpha = p.plane_alpha ? p.plane_alpha + (size_t)w * p.Pm : nullptr;
    double t00[2] = {0, 0}, t01[2] = {0, 0}, t11[2] = {0, 0};   // tiles (rows 0..7 x cols 0..7), (0..7 x 8..15), (8..15 x 8..15) of the Gram matrix
    int cur_key = -1;
    auto flush = [&]() {
      if (cur_key < 0) return;
      const int f = cur_key >> 1; const bool ct = cur_key & 1;
      auto put = [&](int r, int c, double val) {   // entry (r, c), r <= c, of the Gram matrix of [J_f (6) | r] or [J_f (6) | J_f+1 (6) | r]
        const int rc = ct ? 12 : 6;                // residual column
        if (r >= rc || r > c) return;
        if (c == rc) { atomicAdd(&S.g[6 * (f + r / 6) + r % 6], val); return; }
        if (c > rc) return;
        const int br = r / 6, bc = c / 6;          // block row / column: 0 = frame f, 1 = frame f + 1
        double* B = &S.U[ublk(f + br, f + bc, F)];
        atomicAdd(&B[(r % 6) * 6 + c % 6], val);
        if (br == bc && r != c) atomicAdd(&B[(c % 6) * 6 + r % 6], val);   // diagonal blocks are kept full
      };
      put(fq, 2 * fk, t00[0]); put(fq, 2 * fk + 1, t00[1]);
      if (ct) { put(fq, 8 + 2 * fk, t01[0]); put(fq, 9 + 2 * fk, t01[1]); put(8 + fq, 8 + 2 * fk, t11[0]); put(8 + fq, 9 + 2 * fk, t11[1]); }
      t00[0] = t00[1] = t01[0] = t01[1] = t11[0] = t11[1] = 0.0;
    };
    for (int q = wid; q < npt; q += kLinWarps) {
      const int key = p.ptask_frame[(size_t)w * kMaxPlaneTasks + q], cnt = p.ptask_cnt[(size_t)w * kMaxPlaneTasks + q], first = p.ptask_first[(size_t)w * kMaxPlaneTasks + q];
      const int f = key >> 1; const bool ct = key & 1;
      if (key != cur_key) { flush(); cur_key = key; }
      double Jc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, r = 0.0;
      if (lane < cnt) {
        const int pi = pperm[first + lane];
        if (!ct) r = plane_residual(pls[pi], S.fr[f], p.lidar_sqrt_info, Jc);   // LidarPlaneNormFactor on the pose of frame f
        else r = ct_plane_residual(pls[pi], palpha ? palpha[pi] : 0.0, poseW + 7 * f, poseW + 7 * (f + 1), p.lidar_sqrt_info, Jc);   // CTLidarPlaneNormFactor: poses f (begin), f + 1 (end)
        cost_acc += 0.5 * r * r;
      }
      const int nc = ct ? 12 : 6;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 12; c++) if (c < nc) stg[c * kPS + lane] = Jc[c];
      stg[nc * kPS + lane] = r;
      __syncwarp();
      const double* s0 = stg + fq * kPS + fk;
      const bool ok0 = fq <= nc, ok1 = ct && 8 + fq <= nc;   // rows that exist (row nc = the residual column)
#pragma unroll
      for (int ks = 0; ks < 8; ks++) {
        const double a0 = ok0 ? s0[4 * ks] : 0.0;
        mma_f64(t00[0], t00[1], a0, a0);
        if (ct) {   // warp-uniform
          const double a1 = ok1 ? s0[8 * kPS + 4 * ks] : 0.0;
          mma_f64(t01[0], t01[1], a0, a1); mma_f64(t11[0], t11[1], a1, a1);
        }
      }
    }
    flush();
    __syncthreads();
  }
  // S_vis = U - Schur, g_schur = Schur[:,66]
  double* Svis = p.Svis + (size_t)w * kVisRec;
  // blocked output: lower block pairs (bi >= bj) in the order bi (bi + 1) / 2 + bj, each a row-major 6x6 block (the layout
  // k_solve2 assembles from); diagonal blocks are written symmetric from their upper triangle
  for (int idx = t; idx < (F * (F + 1) / 2) * 36; idx += kLinThreads) {
    const int blk = idx / 36, e = idx % 36;
    int bi = (int)((sqrtf(8.0f * (float)blk + 1.0f) - 1.0f) * 0.5f);  // row of the lower block triangle
    if ((bi + 1) * (bi + 2) / 2 <= blk) bi++; else if (bi * (bi + 1) / 2 > blk) bi--;
    const int bj = blk - bi * (bi + 1) / 2;
    const int r = 6 * bi + e / 6, c = 6 * bj + e % 6;
    const int a = r <= c ? r : c, b = r <= c ? c : r;  // upper element (a <= b) of U and of the Schur tiles
    Svis[idx] = S.U[ublk(a / 6, b / 6, F) + (a % 6) * 6 + (b % 6)] - S.WT[(a + kRowShift) * kNVP + b + kRowShift];
  }
  for (int q = t; q < NV; q += kLinThreads) {
    p.gvis[(size_t)w * kVisRec + q] = S.g[q];                       // full visual gradient J^T r (pose part)
    p.gschur[(size_t)w * kVisRec + q] = S.WT[(q + kRowShift) * kNVP + 66 + kRowShift];         // sum_l w_l g_l / v'_l, subtracted to form the reduced rhs
    p.Udiag[(size_t)w * kVisRec + q] = S.U[ublk(q / 6, q / 6, F) + (q % 6) * 7];
  }
  double red2[2] = {cost_acc, 0.0};
  block_sum<2>(red2, S.red);
  gmax = warp_max(gmax);
  if (lane == 0) S.red[wid] = gmax;
  __syncthreads();
  if (t == 0) {
    double gm = 0; for (int i2 = 0; i2 < kLinWarps; i2++) gm = fmax(gm, S.red[i2]);
    p.c_lin[(size_t)w * kVisRec] = red2[0]; p.c_gmax[w] = gm;
  }
#ifdef GF2_PHASE_CLOCKS
  if (t == 0 && blockIdx.x == 300) printf("k_linearize clocks: init %lld step %lld syrk %lld tail %lld total %lld\n", lc_t1 - lc_t0 - lc_step - lc_syrk, lc_step, lc_syrk, clock64() - lc_t1, clock64() - lc_t0);
#endif
}

}  // namespace gf2
