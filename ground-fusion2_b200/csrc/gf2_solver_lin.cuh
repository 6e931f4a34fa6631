// gf2_solver_lin.cuh — k_tasks (start-frame-uniform warp tasks) and k_linearize (sweep 1).
//
// Work decomposition of the Jacobian sweep: landmarks are grouped by host frame (start_frame); every group is cut into
// warp tasks of <= 32 landmarks, so all lanes of a warp evaluate, at step k, observations of the SAME frame pair
// (i, i + k). Their 63 Hessian/gradient products (block (i,j) 36, block (j,j) upper 21, g_j 6) are then reduced across
// the warp with a transposed butterfly (31 shuffle-adds per 32 values instead of 5 per value) and each lane flushes
// one reduced value per batch into the CTA's shared pose-block matrix. Landmark columns w_l go, transposed, into a
// shared tile consumed by the fp64 tensor-core SYRK (mma.sync m8n8k4) that forms the Schur complement.
#pragma once
#include "gf2_solver_kernels.cuh"

namespace gf2 {

constexpr int kMaxTasks = 48;  // ceil(1000 / 32) + 11 partial tasks + slack
constexpr int kMaxPlaneTasks = 192;  // planes: up to ~5.8k per window in tasks of 32 of the same frame

// ------------------------------------------------------------------------------------------------ k_tasks
// Deterministic counting sort of the landmark table by start frame -> perm, and the warp-task list. One thread per start
// frame scans the table in order (the reference table is already sorted, VE/estimator/feature_manager.cpp:67-88; any
// order is accepted).
__global__ void k_tasks(KP p, int w0) {
  const int w = w0 + blockIdx.x;
  const int t = threadIdx.x, F = p.F;
  __shared__ int cnt[kMaxF], ofs[kMaxF + 1];
  const int nlm = p.nlm[w];
  const int32_t* start = p.start + (size_t)w * p.Lm;
  if (t < F) { int c = 0; for (int l = 0; l < nlm; l++) c += (start[l] == t); cnt[t] = c; }
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
  if (t < F) { int32_t* perm = p.perm + (size_t)w * p.Lm; int o = ofs[t]; for (int l = 0; l < nlm; l++) if (start[l] == t) perm[o++] = l; }
  if (!p.planes) return;
  // LiDAR plane factors grouped by frame, tasks of <= 32 planes
  __syncthreads();
  const int np = p.n_planes[w];
  const gf2_plane* pls = p.planes + (size_t)w * p.Pm;
  if (t < F) { int c = 0; for (int q = 0; q < np; q++) c += (pls[q].frame == t); cnt[t] = c; }
  __syncthreads();
  if (t == 0) {
    int o = 0, nt = 0;
    int32_t* tf = p.ptask_first + (size_t)w * kMaxPlaneTasks; int32_t* tc = p.ptask_cnt + (size_t)w * kMaxPlaneTasks; int32_t* ts = p.ptask_frame + (size_t)w * kMaxPlaneTasks;
    for (int s = 0; s < F; s++) {
      ofs[s] = o;
      for (int c = 0; c < cnt[s] && nt < kMaxPlaneTasks; c += 32) { tf[nt] = o + c; tc[nt] = min(32, cnt[s] - c); ts[nt] = s; nt++; }
      o += cnt[s];
    }
    p.nptasks[w] = nt;
  }
  __syncthreads();
  if (t < F) { int32_t* perm = p.pperm + (size_t)w * p.Pm; int o = ofs[t]; for (int q = 0; q < np; q++) if (pls[q].frame == t) perm[o++] = q; }
}

// ------------------------------------------------------------------------------------------------ k_linearize
constexpr int kWTRows = 68;        // 66 tangent rows + the landmark-gradient row 66 + one always-zero row 67 (tile padding)
constexpr int kStageStride = 13;   // per residual row: Ji (6) | Jj (6) | r
constexpr int kNBlkPairs = kMaxF * (kMaxF + 1) / 2;

struct LinShared {
  FrameCtx fr[kMaxF];
  CamCtx cam;
  double U[kNBlkPairs * 36];  // pose-block Hessian of the visual factors: 6x6 blocks (bi <= bj), row-major inside
  double g[kNVP];
  double red[8 * 32];
  int task_first[kMaxTasks], task_cnt[kMaxTasks], task_start[kMaxTasks];
  double stage[8][64 * kStageStride];  // per warp: the 64 residual rows of the current step
  double WT[kWTRows * kWTStride];      // transposed, pre-scaled landmark columns  w_l / sqrt(v'_l)
};

__device__ __forceinline__ int ublk(int bi, int bj, int F) { return (bi * F - bi * (bi - 1) / 2 + (bj - bi)) * 36; }  // bi <= bj

// flush one 8x8 accumulator tile (m = lane/4, n = 2*(lane%4)+{0,1}) into a 6x6 block (+ optional gradient column n = 6)
__device__ __forceinline__ void flush_tile(double* blk, double* gvec, double c0, double c1, int lane) {
  const int m = lane >> 2, n = (lane & 3) * 2;
  if (m < 6) {
    if (n < 6) { atomicAdd(&blk[m * 6 + n], c0); atomicAdd(&blk[m * 6 + n + 1], c1); }
    else if (gvec) atomicAdd(&gvec[m], c0);  // n == 6
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
  for (int i = t; i < kNBlkPairs * 36; i += kLinThreads) S.U[i] = 0.0;
  if (t < kNVP) S.g[t] = 0.0;
  const int ntasks = p.ntasks[w];
  if (t < kMaxTasks) { S.task_first[t] = p.task_first[(size_t)w * kMaxTasks + t]; S.task_cnt[t] = p.task_cnt[(size_t)w * kMaxTasks + t]; S.task_start[t] = p.task_start[(size_t)w * kMaxTasks + t]; }
  for (int i = t; i < kWTRows * kWTStride; i += kLinThreads) S.WT[i] = 0.0;
  __syncthreads();

  const int32_t* tlen = p.tlen + (size_t)w * p.Lm; const int32_t* obeg = p.obeg + (size_t)w * p.Lm; const int32_t* perm = p.perm + (size_t)w * p.Lm;
  const float4* obs = p.obs + (size_t)w * p.Om;
  const double* ftd = p.frame_td + (size_t)w * F;
  const double mu = st.mu;
  const bool it0 = (st.iteration == 0);
  double cost_acc = 0.0, gmax = 0.0;
  double C[6][2];  // Schur accumulators: sym tiles (a <= b) of the 9x9 tile grid, tile q -> warp q % 8, slot q / 8
#pragma unroll
  for (int q = 0; q < 6; q++) { C[q][0] = 0.0; C[q][1] = 0.0; }
  int tile_a[6], tile_b[6];  // tiles owned by this warp: q = wid + 8 * slot in row-major order of the upper 9x9 tile triangle
#pragma unroll
  for (int s2 = 0; s2 < 6; s2++) {
    int q = wid + 8 * s2, a = 0;
    if (q >= 45) { tile_a[s2] = -1; tile_b[s2] = 0; continue; }
    while (q >= 9 - a) { q -= 9 - a; a++; }
    tile_a[s2] = a; tile_b[s2] = a + q;
  }
  double* stg = S.stage[wid];
  const int fq = lane >> 2, fk = lane & 3;  // fragment coordinates: row/col index 0..7, k index 0..3

  for (int tbase = 0; tbase < ntasks; tbase += 8) {
    const int task = tbase + wid;
    const bool have_task = task < ntasks;
    const int i = have_task ? S.task_start[task] : 0;
    const bool have = have_task && lane < S.task_cnt[task];
    const int row0 = 8 * ((6 * S.task_start[tbase]) >> 3);  // tiles left of the round's first host frame are skipped by the SYRK
    for (int c = row0; c < kWTRows - 1; c++) S.WT[c * kWTStride + t] = 0.0;
    int l = 0, L = 0, ob = 0; bool fx = false; double lam = 1.0;
    LmCtx lc;
    if (have) {
      l = perm[S.task_first[task] + lane]; L = tlen[l]; ob = obeg[l]; fx = p.fixed[(size_t)w * p.Lm + l] != 0; lam = p.invdep[(size_t)w * p.Lm + l];
      landmark_ctx(S.fr[i], S.cam, obs[ob], ftd[i], lam, lc);
    }
    double wi[6] = {0, 0, 0, 0, 0, 0};  // w_i = sum_k Ji^T jl
    double v = 0.0, gl = 0.0;
    double cii0 = 0.0, cii1 = 0.0;      // Ji^T [Ji | r] accumulated over the steps of the task (host-frame block + g_i)
    const int Lmax = __reduce_max_sync(0xffffffffu, L);
    for (int k = 1; k < Lmax; k++) {
      const bool valid = have && k < L;
      const int j = i + k;  // uniform across the warp
      double Jx[6], Jj[12], Ji[12], r0 = 0, r1 = 0;
#pragma unroll
      for (int c = 0; c < 12; c++) { Ji[c] = 0; Jj[c] = 0; }
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
#pragma unroll
        for (int r = 0; r < 2; r++) {  // Ji = [Jx | Jx * Gi]
          Ji[r * 6 + 0] = Jx[r * 3]; Ji[r * 6 + 1] = Jx[r * 3 + 1]; Ji[r * 6 + 2] = Jx[r * 3 + 2];
#pragma unroll
          for (int c = 0; c < 3; c++) Ji[r * 6 + 3 + c] = Jx[r * 3] * lc.Gi.m[c] + Jx[r * 3 + 1] * lc.Gi.m[3 + c] + Jx[r * 3 + 2] * lc.Gi.m[6 + c];
        }
        double jl0 = 0, jl1 = 0;
        if (!fx) {
          jl0 = Jx[0] * lc.dXdl.x + Jx[1] * lc.dXdl.y + Jx[2] * lc.dXdl.z;
          jl1 = Jx[3] * lc.dXdl.x + Jx[4] * lc.dXdl.y + Jx[5] * lc.dXdl.z;
        }
        v += jl0 * jl0 + jl1 * jl1; gl += jl0 * r0 + jl1 * r1;
#pragma unroll
        for (int c = 0; c < 6; c++) {
          wi[c] += Ji[c] * jl0 + Ji[6 + c] * jl1;
          S.WT[(6 * j + c) * kWTStride + t] = Jj[c] * jl0 + Jj[6 + c] * jl1;  // w_j = Jj^T jl
        }
      }
      if (have_task) {
        // stage the 64 residual rows of this step, then reduce across the warp on the tensor cores:
        //   C1 = Ji^T Jj -> block (i, j);  C2 = Jj^T [Jj | r] -> block (j, j), g_j;  Cii += Ji^T [Ji | r]
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 2; r++) {
          double* row = stg + (2 * lane + r) * kStageStride;
#pragma unroll
          for (int c = 0; c < 6; c++) { row[c] = Ji[r * 6 + c]; row[6 + c] = Jj[r * 6 + c]; }
          row[12] = r ? r1 : r0;
        }
        __syncwarp();
        double c10 = 0, c11 = 0, c20 = 0, c21 = 0, d10 = 0, d11 = 0, d20 = 0, d21 = 0, dii0 = 0, dii1 = 0;
        const int colx = fq < 6 ? fq : 12;  // Ji column, or the residual for fragment index 6 (index 7 reads it too, masked below)
#pragma unroll 4
        for (int s = 0; s < 16; s++) {
          const double* row = stg + (4 * s + fk) * kStageStride;
          const double x = row[colx];
          const double y = row[6 + (fq < 6 ? fq : 0)];
          const double a_i = fq < 6 ? x : 0.0;              // Ji^T as A, also Ji as B (n < 6)
          const double a_j = fq < 6 ? y : 0.0;              // Jj^T as A, also Jj as B (n < 6)
          const double b_ir = fq < 7 ? x : 0.0;             // [Ji | r]
          const double b_jr = fq < 6 ? y : (fq == 6 ? x : 0.0);  // [Jj | r]
          if (s & 1) { mma_f64(d10, d11, a_i, a_j); mma_f64(d20, d21, a_j, b_jr); mma_f64(dii0, dii1, a_i, b_ir); }
          else { mma_f64(c10, c11, a_i, a_j); mma_f64(c20, c21, a_j, b_jr); mma_f64(cii0, cii1, a_i, b_ir); }
        }
        c10 += d10; c11 += d11; c20 += d20; c21 += d21; cii0 += dii0; cii1 += dii1;
        flush_tile(&S.U[ublk(i, j, F)], nullptr, c10, c11, lane);
        flush_tile(&S.U[ublk(j, j, F)], &S.g[6 * j], c20, c21, lane);
      }
    }
    if (have_task) flush_tile(&S.U[ublk(i, i, F)], &S.g[6 * i], cii0, cii1, lane);
    // landmark scalars: jacobi scale (iteration 0), regularised v' = v + mu * e; the landmark's column of W is scaled by
    // 1/sqrt(v') so that the SYRK below needs no per-element multiply
    if (have) {
      double s_l;
      if (it0) { s_l = 1.0 / (1.0 + sqrt(v)); p.lm_s[(size_t)w * p.Lm + l] = s_l; } else s_l = p.lm_s[(size_t)w * p.Lm + l];
      const double d2 = fmin(fmax(s_l * s_l * v, 1e-6), 1e32);
      const double e = d2 / (s_l * s_l);
      const double vp = v + mu * e;
      const double rs = (!fx && v > 0.0) ? rsqrt(vp) : 0.0;
      p.lm_v[(size_t)w * p.Lm + l] = fx ? 0.0 : v;
      p.lm_g[(size_t)w * p.Lm + l] = fx ? 0.0 : gl;
      if (!fx) gmax = fmax(gmax, fabs(gl));
#pragma unroll
      for (int c = 0; c < 6; c++) S.WT[(6 * i + c) * kWTStride + t] = wi[c] * rs;
      for (int c = 6 * (i + 1); c < 6 * (i + L); c++) S.WT[c * kWTStride + t] *= rs;
      S.WT[66 * kWTStride + t] = fx ? 0.0 : gl * rs;
    }
    __syncthreads();
    // Schur SYRK on the fp64 tensor cores: C[a-tile][b-tile] += sum_l Ws[l][a] * Ws[l][b]
    {
      const int ta_min = row0 >> 3;
      const double* wa[6]; const double* wb[6]; bool on[6];
#pragma unroll
      for (int s2 = 0; s2 < 6; s2++) {
        on[s2] = tile_a[s2] >= ta_min;  // (also false for unowned slots: tile_a = -1)
        wa[s2] = &S.WT[min(8 * max(tile_a[s2], 0) + fq, kWTRows - 1) * kWTStride + fk];
        wb[s2] = &S.WT[min(8 * tile_b[s2] + fq, kWTRows - 1) * kWTStride + fk];
      }
      // all owned tiles advance together: six independent accumulator chains hide the mma latency
#pragma unroll 2
      for (int k0 = 0; k0 < kLinThreads; k0 += 4) {
#pragma unroll
        for (int s2 = 0; s2 < 6; s2++) if (on[s2]) mma_f64(C[s2][0], C[s2][1], wa[s2][k0], wb[s2][k0]);
      }
    }
    __syncthreads();
  }

  // LiDAR plane factors: tasks of <= 32 planes of one frame; [J | r]^T [J | r] reduced on the tensor cores into block (f, f), g_f
  if (p.planes) {
    const int npt = p.nptasks[w];
    const gf2_plane* pls = p.planes + (size_t)w * p.Pm;
    const int32_t* pperm = p.pperm + (size_t)w * p.Pm;
    for (int q = wid; q < npt; q += 8) {
      const int f = p.ptask_frame[(size_t)w * kMaxPlaneTasks + q], cnt = p.ptask_cnt[(size_t)w * kMaxPlaneTasks + q], first = p.ptask_first[(size_t)w * kMaxPlaneTasks + q];
      double Jp[6] = {0, 0, 0, 0, 0, 0}, r = 0.0;
      if (lane < cnt) { r = plane_residual(pls[pperm[first + lane]], S.fr[f], p.lidar_sqrt_info, Jp); cost_acc += 0.5 * r * r; }
      __syncwarp();
      double* row = stg + lane * kStageStride;
#pragma unroll
      for (int c = 0; c < 6; c++) row[c] = Jp[c];
      row[6] = r;
      __syncwarp();
      double c0 = 0, c1 = 0;
#pragma unroll
      for (int s = 0; s < 8; s++) {
        const double x = stg[(4 * s + fk) * kStageStride + (fq < 7 ? fq : 0)];
        mma_f64(c0, c1, fq < 6 ? x : 0.0, fq < 7 ? x : 0.0);
      }
      flush_tile(&S.U[ublk(f, f, F)], &S.g[6 * f], c0, c1, lane);
    }
    __syncthreads();
  }
  // S_vis = U - C, g_schur = C[:,66]; the Schur tiles are staged in WT (free now) as a dense 72x72 matrix
#pragma unroll
  for (int s2 = 0; s2 < 6; s2++) if (tile_a[s2] >= 0) {
    const int row = 8 * tile_a[s2] + fq, col = 8 * tile_b[s2] + 2 * fk;
    S.WT[row * kNVP + col] = C[s2][0]; S.WT[row * kNVP + col + 1] = C[s2][1];
  }
  __syncthreads();
  double* Svis = p.Svis + (size_t)w * kNVMax * kNVMax;
  // blocked output: lower block pairs (bi >= bj) in the order bi (bi + 1) / 2 + bj, each a row-major 6x6 block (the layout
  // k_solve2 assembles from); diagonal blocks are written symmetric from their upper triangle
  for (int idx = t; idx < (F * (F + 1) / 2) * 36; idx += kLinThreads) {
    const int blk = idx / 36, e = idx % 36;
    int bi = 0; while ((bi + 1) * (bi + 2) / 2 <= blk) bi++;
    const int bj = blk - bi * (bi + 1) / 2;
    const int r = 6 * bi + e / 6, c = 6 * bj + e % 6;
    const int a = r <= c ? r : c, b = r <= c ? c : r;  // upper element (a <= b) of U and of the Schur tiles
    Svis[idx] = S.U[ublk(a / 6, b / 6, F) + (a % 6) * 6 + (b % 6)] - S.WT[a * kNVP + b];
  }
  (void)NV;
  if (t < NV) {
    p.gvis[(size_t)w * kNVP + t] = S.g[t];                       // full visual gradient J^T r (pose part)
    p.gschur[(size_t)w * kNVP + t] = S.WT[t * kNVP + 66];         // sum_l w_l g_l / v'_l, subtracted to form the reduced rhs
    p.Udiag[(size_t)w * kNVMax + t] = S.U[ublk(t / 6, t / 6, F) + (t % 6) * 7];
  }
  double red2[2] = {cost_acc, 0.0};
  block_sum<2>(red2, S.red);
  gmax = warp_max(gmax);
  if (lane == 0) S.red[wid] = gmax;
  __syncthreads();
  if (t == 0) {
    double gm = 0; for (int i2 = 0; i2 < kLinThreads / 32; i2++) gm = fmax(gm, S.red[i2]);
    p.c_lin[(size_t)w * 4] = red2[0]; p.c_gmax[w] = gm;
  }
}

}  // namespace gf2
