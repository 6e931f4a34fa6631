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
}

// ------------------------------------------------------------------------------------------------ k_linearize
struct LinShared {
  FrameCtx fr[kMaxF];
  CamCtx cam;
  double U[kNVMax * kNVMax];  // pose-block Hessian of the visual factors, upper block triangle filled
  double g[kNVP];
  double invv[kLinThreads];
  double red[8 * 32];
  int task_first[kMaxTasks], task_cnt[kMaxTasks], task_start[kMaxTasks];
  double WT[kNVP * kWTStride];
};

// transposed butterfly over 32 values: on return lane L holds sum over all lanes of v[L] (in v[0])
__device__ __forceinline__ double butterfly32(double (&v)[32], int lane) {
#pragma unroll
  for (int bit = 16, h = 16; bit > 0; bit >>= 1, h >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int t = 0; t < h; t++) {
      const double send = up ? v[t] : v[t + h];
      const double keep = up ? v[t + h] : v[t];
      v[t] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  return v[0];
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
  const int ntasks = p.ntasks[w];
  if (t < kMaxTasks) { S.task_first[t] = p.task_first[(size_t)w * kMaxTasks + t]; S.task_cnt[t] = p.task_cnt[(size_t)w * kMaxTasks + t]; S.task_start[t] = p.task_start[(size_t)w * kMaxTasks + t]; }
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

  for (int tbase = 0; tbase < ntasks; tbase += 8) {
    const int task = tbase + wid;
    const bool have_task = task < ntasks;
    const int i = have_task ? S.task_start[task] : 0;
    const bool have = have_task && lane < S.task_cnt[task];
    for (int c = 8 * ((6 * S.task_start[tbase]) >> 3); c < kNVP; c++) S.WT[c * kWTStride + t] = 0.0;  // tiles left of the round first host frame are skipped below
    int l = 0, L = 0, ob = 0; bool fx = false; double lam = 1.0;
    LmCtx lc;
    if (have) {
      l = perm[S.task_first[task] + lane]; L = tlen[l]; ob = obeg[l]; fx = p.fixed[(size_t)w * p.Lm + l] != 0; lam = p.invdep[(size_t)w * p.Lm + l];
      landmark_ctx(S.fr[i], S.cam, obs[ob], ftd[i], lam, lc);
    }
    double M[6] = {0, 0, 0, 0, 0, 0}, m3[3] = {0, 0, 0}, n3[3] = {0, 0, 0};
    double v = 0.0, gl = 0.0;
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
        M[0] += Jx[0] * Jx[0] + Jx[3] * Jx[3]; M[1] += Jx[0] * Jx[1] + Jx[3] * Jx[4]; M[2] += Jx[0] * Jx[2] + Jx[3] * Jx[5];
        M[3] += Jx[1] * Jx[1] + Jx[4] * Jx[4]; M[4] += Jx[1] * Jx[2] + Jx[4] * Jx[5]; M[5] += Jx[2] * Jx[2] + Jx[5] * Jx[5];
#pragma unroll
        for (int c = 0; c < 3; c++) { m3[c] += Jx[c] * jl0 + Jx[3 + c] * jl1; n3[c] += Jx[c] * r0 + Jx[3 + c] * r1; }
        v += jl0 * jl0 + jl1 * jl1; gl += jl0 * r0 + jl1 * r1;
#pragma unroll
        for (int c = 0; c < 6; c++) S.WT[(6 * j + c) * kWTStride + t] = Jj[c] * jl0 + Jj[6 + c] * jl1;  // w_j = Jj^T jl
      }
      if (have_task && j < F) {
        // batch A: entries 0..31 of the (i, j) block
        double b[32];
#pragma unroll
        for (int e = 0; e < 32; e++) b[e] = Ji[e / 6] * Jj[e % 6] + Ji[6 + e / 6] * Jj[6 + e % 6];
        double s = butterfly32(b, lane);
        atomicAdd(&S.U[(6 * i + lane / 6) * kNVMax + 6 * j + lane % 6], s);
        // batch B: entries 32..35 of (i, j), the 21 upper entries of (j, j), g_j (6), one spare
#pragma unroll
        for (int e = 0; e < 4; e++) b[e] = Ji[5] * Jj[2 + e] + Ji[11] * Jj[8 + e];
        {
          int q = 4;
#pragma unroll
          for (int a = 0; a < 6; a++)
#pragma unroll
            for (int c = a; c < 6; c++) b[q++] = Jj[a] * Jj[c] + Jj[6 + a] * Jj[6 + c];
#pragma unroll
          for (int a = 0; a < 6; a++) b[q++] = Jj[a] * r0 + Jj[6 + a] * r1;
          b[31] = 0.0;
        }
        s = butterfly32(b, lane);
        if (lane < 4) atomicAdd(&S.U[(6 * i + 5) * kNVMax + 6 * j + 2 + lane], s);
        else if (lane < 25) {
          // unrank the upper-triangular index lane - 4 -> (a, c), a <= c < 6
          int q = lane - 4, a = 0;
          while (q >= 6 - a) { q -= 6 - a; a++; }
          atomicAdd(&S.U[(6 * j + a) * kNVMax + 6 * j + a + q], s);
        } else if (lane < 31) atomicAdd(&S.g[6 * j + lane - 25], s);
      }
    }
    // host-frame block: U_ii += [M, M Gi; Gi^T M, Gi^T M Gi], g_i += [n3; Gi^T n3], w_i = [m3; Gi^T m3]
    if (have_task) {
      double b[32];
#pragma unroll
      for (int e = 0; e < 32; e++) b[e] = 0.0;
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
          for (int c = a; c < 6; c++) {
            double val;
            if (a < 3 && c < 3) val = Mf[a * 3 + c];
            else if (a < 3) val = MG[a * 3 + (c - 3)];
            else val = lc.Gi.m[(a - 3)] * MG[(c - 3)] + lc.Gi.m[3 + (a - 3)] * MG[3 + (c - 3)] + lc.Gi.m[6 + (a - 3)] * MG[6 + (c - 3)];
            b[q++] = val;
          }
#pragma unroll
        for (int c = 0; c < 3; c++) {
          b[21 + c] = n3[c];
          b[24 + c] = lc.Gi.m[c] * n3[0] + lc.Gi.m[3 + c] * n3[1] + lc.Gi.m[6 + c] * n3[2];
          S.WT[(6 * i + c) * kWTStride + t] = m3[c];
          S.WT[(6 * i + 3 + c) * kWTStride + t] = lc.Gi.m[c] * m3[0] + lc.Gi.m[3 + c] * m3[1] + lc.Gi.m[6 + c] * m3[2];
        }
      }
      const double s = butterfly32(b, lane);
      if (lane < 21) {
        int q = lane, a = 0;
        while (q >= 6 - a) { q -= 6 - a; a++; }
        atomicAdd(&S.U[(6 * i + a) * kNVMax + 6 * i + a + q], s);
      } else if (lane < 27) atomicAdd(&S.g[6 * i + lane - 21], s);
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
    // Schur SYRK on the fp64 tensor cores: C[a-tile][b-tile] += sum_l (W[l][a] / v'_l) * W[l][b]; tiles whose rows are all
    // left of the round's first host frame are identically zero and skipped
    {
      const int kr = lane & 3, mc = lane >> 2;
      const int ta_min = (6 * S.task_start[tbase]) >> 3;
      int q = 0;
      for (int ta = 0; ta < 9; ta++)
        for (int tb = ta; tb < 9; tb++, q++) {
          if ((q & 7) != wid || ta < ta_min) continue;
          const int slot = q >> 3;
          double c0 = 0, c1 = 0;
          const double* wa = &S.WT[(8 * ta + mc) * kWTStride + kr];
          const double* wb = &S.WT[(8 * tb + mc) * kWTStride + kr];
#pragma unroll 4
          for (int k0 = 0; k0 < kLinThreads; k0 += 4) mma_f64(c0, c1, wa[k0] * S.invv[k0 + kr], wb[k0]);
#pragma unroll
          for (int s2 = 0; s2 < 6; s2++) if (s2 == slot) { C[s2][0] += c0; C[s2][1] += c1; }
        }
    }
    __syncthreads();
  }

  // S_vis = U - C, g_vis = g - C[:,66]; the Schur tiles are staged in WT (free now) as a dense 72x72 matrix
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
        S.WT[row * kNVP + col] = c0; S.WT[row * kNVP + col + 1] = c1;
      }
  }
  __syncthreads();
  double* Svis = p.Svis + (size_t)w * kNVMax * kNVMax;
  for (int idx = t; idx < NV * NV; idx += kLinThreads) {
    const int r = idx / NV, c = idx % NV;
    const int a = r <= c ? r : c, b = r <= c ? c : r;  // stored element (a <= b)
    Svis[r * kNVMax + c] = S.U[a * kNVMax + b] - S.WT[a * kNVP + b];
  }
  if (t < NV) {
    p.gvis[(size_t)w * kNVP + t] = S.g[t];                       // full visual gradient J^T r (pose part)
    p.gschur[(size_t)w * kNVP + t] = S.WT[t * kNVP + 66];         // sum_l w_l g_l / v'_l, subtracted to form the reduced rhs
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
