// gf2_solver_lin_ws.cuh — k_linearize_ws: the warp-specialised Jacobian sweep for SMALL batches (one robot = one window per call).
//
// k_linearize (gf2_solver_lin.cuh) gets its occupancy from the batch: four warps per window, two windows per SM. With fewer windows than SMs
// that leaves the SM at four warps, and a window's sweep is a chain of 8 x (observation phase -> barrier -> SYRK -> barrier). This kernel
// spends a whole SM on one window instead: 16 warps, split by role, connected by mbarriers (no CTA-wide barrier inside the sweep). Same
// arithmetic per factor as k_linearize (moment form, Gram on the fp64 tensor cores, end-aligned SYRK tiles), same outputs; measured on
// B200: one W10-F1000 window in 0.062 ms instead of 0.091 ms, a full 4096-window batch 12 % slower (profiles/sweep_variants_r2.txt) - the launcher picks
// by batch size.
#pragma once
#include "gf2_solver_lin.cuh"

namespace gf2 {
namespace ws {

// Warp-specialised, one window per CTA, one CTA (16 warps) per SM:
//   producer warps 0..11 (six pairs)   evaluate the projection factors. A pair owns one warp task (<= 32 landmarks of one host frame) at a
//       time; its two warps take the odd / even steps k of the task (frame pair (i, i + k)), so a landmark's column of W is written by two
//       warps and its scalars (v, g_l, sum n) are combined through shared memory behind two named barriers. The task's observation records
//       (contiguous in HBM when the landmark table is ordered by start frame, as the reference's is) are staged in shared memory by ONE bulk
//       asynchronous copy (cp.async.bulk global -> shared, completion on the pair's mbarrier, SASS UBLKCP) issued a whole task ahead.
//   consumer warps 12..15              own the 45 Schur tiles in registers and run the fp64 tensor-core SYRK over the tiles of W the producers
//       hand over through a ring of kStages shared-memory stages (mbarrier full / empty per stage).
// The fp64 FMA chains of the producers (latency bound) and the DMMA stream of the consumers share the SM's fp64 datapath; splitting them by
// warp keeps both fed at the same time instead of alternating phases behind CTA-wide barriers.
constexpr int kLinThreads = 512;
constexpr int kProdWarps = 12, kConsWarps = 4, kPairs = kProdWarps / 2;
constexpr int kProdThreads = kProdWarps * 32, kConsThreads = kConsWarps * 32;
constexpr int kTileCols = 32;      // one warp task per tile of W
constexpr int kWTRows = 68;        // 66 tangent rows + the landmark-gradient row 66 + one always-zero row 67 (tile padding)
constexpr int kWTStride = kTileCols + 4;  // transposed W tile: [kWTRows][kWTStride] doubles
constexpr int kStages = 6;
constexpr int kObsStage = 32 * kMaxF;     // observation records of one task
constexpr int kYStride = 68;       // staging of Y^T per producer warp: [7][kYStride], rows 0..63 = the warp's residual rows (row 7 is not stored)
// SYRK tile grid: tile row a covers the rows r' = 8a .. 8a+7 of the END-ALIGNED index r' = r + kRowShift (r = tangent index 0..65, landmark
// gradient at r = 66; r = 67 and r' < kRowShift are zero padding). Every landmark's support ends at the last frame, so a task whose
// host frame is i touches the tile rows a >= (6 i + kRowShift) / 8 only.
// named barriers: 0 = __syncthreads, 1 + 2 pair / 2 + 2 pair = the two hand-overs inside a producer pair, 13 = producers, 14 = consumers
constexpr int kBarProd = 13, kBarCons = 14;

struct LinShared {
  FrameCtx fr[kMaxF];
  CamCtx cam;
  double g[kNVP];
  double Mom[kNPairs * kMomStride];  // per frame pair (i < j): the 27 moment sums
  double red[2 * 32];
  int task_first[kMaxTasks], task_cnt[kMaxTasks], task_start[kMaxTasks];
  unsigned long long full[kStages], empty[kStages], obsbar[kPairs];
  double part[kPairs][6][32];        // pair hand-over: v, g_l, sum n (3) of the even-step warp; 1 / sqrt(v') back
  union {
    double Y[kProdWarps][7 * kYStride];  // during the sweep
    struct {
      double U[kNBlkPairs * 36];         // afterwards: pose-block Hessian of the visual factors, 6x6 blocks (bi <= bj), row-major inside
      double sums[kMaxF * kSumStride];   // un-rotated per-frame sums
    };
  };
  alignas(16) float4 sobs[kPairs][kObsStage];  // bulk-copied observation records of the pair's current task
  alignas(16) double WT[kStages][kWTRows * kWTStride];  // transposed, pre-scaled landmark columns  w_l / sqrt(v'_l); afterwards the dense Schur tiles
};
static_assert(kStages == kPairs, "stage s is producer pair s's");
static_assert(kStages * kWTRows * kWTStride >= kNVP * kNVP, "dense 72x72 Schur tiles fit in the stage ring");
static_assert(sizeof(LinShared) <= 227 * 1024, "one CTA per SM");
static_assert(sizeof(float4) * kPairs * kObsStage >= sizeof(double) * kMaxF * (kMaxF - 1) * kDiagTerm, "the per-frame-pair terms of the tail fit in the observation staging");

// ---- synchronisation primitives (PTX): named barriers between warp groups, mbarriers for the stage ring and the bulk copies
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"   // suspend-time hint: the warp sleeps in hardware instead of polling the shared-memory pipe
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// the consumer side of k_linearize for consumer warp W
template <int W>
__device__ __forceinline__ void lin_consume(LinShared& S, int ntasks, int lane) {
  double C[kSchurSlots][2];  // Schur accumulators of the tiles this warp owns (tile_of)
#pragma unroll
  for (int q = 0; q < kSchurSlots; q++) { C[q][0] = 0.0; C[q][1] = 0.0; }
  const int fq = lane >> 2, fk = lane & 3;  // fragment coordinates: row/col index 0..7, k index 0..3
  const bool lo_ok = fq >= kRowShift;
  // tiles are consumed in task order (stage n % kStages belongs to producer pair n % kPairs): the accumulation order is fixed, results do
  // not depend on which pair finishes first
  for (int n = 0; n < ntasks; n++) {
    const int s = n % kStages;
    const int a_min = (6 * S.task_start[n] + kRowShift) >> 3;  // tile rows above the task's host frame hold zeros: skipped
    mbar_wait(&S.full[s], (n / kStages) & 1);
    syrk_warp<W, kWTStride, kTileCols>(a_min, &S.WT[s][(fq - kRowShift) * kWTStride + fk], lo_ok, C);
    mbar_arrive(&S.empty[s]);
  }
  bar_sync(kBarCons, kConsThreads);  // every consumer warp has read the last stage: the ring becomes the dense tile store
  syrk_store<W>(&S.WT[0][0], C, fq, fk);
}

__global__ void __launch_bounds__(kLinThreads, 1) k_linearize_ws(KP p, int w0) {
  const int w = w0 + blockIdx.x;
  WinState& st = p.st[w];
  if (!st.active || st.reuse) return;
  extern __shared__ __align__(128) unsigned char smem_lin_ws[];
  LinShared& S = *reinterpret_cast<LinShared*>(smem_lin_ws);
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int F = p.F, NV = 6 * F;
  const double* pose = p.pose + (size_t)w * F * 7;
  build_frames(pose, p.ex + (size_t)w * 7, p.td[w], F, S.fr, &S.cam);
  for (int i = t; i < kNPairs * kMomStride; i += kLinThreads) S.Mom[i] = 0.0;
  if (t < kNVP) S.g[t] = 0.0;
  const int ntasks = p.ntasks[w];
  if (t < kMaxTasks) { S.task_first[t] = p.task_first[(size_t)w * kMaxTasks + t]; S.task_cnt[t] = p.task_cnt[(size_t)w * kMaxTasks + t]; S.task_start[t] = p.task_start[(size_t)w * kMaxTasks + t]; }
  for (int i = t; i < kStages * kWTRows * kWTStride; i += kLinThreads) (&S.WT[0][0])[i] = 0.0;
  if (t == 0) {
    for (int s = 0; s < kStages; s++) { mbar_init(&S.full[s], 64); mbar_init(&S.empty[s], kConsThreads); }
    for (int q = 0; q < kPairs; q++) mbar_init(&S.obsbar[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double cost_acc = 0.0, gmax = 0.0;
  if (wid >= kProdWarps) {
    // ------------------------------------------------------------------------------------------ consumers: Schur SYRK
    switch (wid - kProdWarps) {
      case 0: lin_consume<0>(S, ntasks, lane); break;
      case 1: lin_consume<1>(S, ntasks, lane); break;
      case 2: lin_consume<2>(S, ntasks, lane); break;
      default: lin_consume<3>(S, ntasks, lane); break;
    }
  } else {
    // ------------------------------------------------------------------------------------------ producers: factor sweep
    const int pr = wid >> 1, hf = wid & 1;   // pair, half: hf = 0 takes the steps k = 1, 3, .., hf = 1 the steps k = 2, 4, ..
    const int4* lminfo = p.lminfo + (size_t)w * p.Lm;
    const float4* obs = p.obs + (size_t)w * p.Om;
    const double* ftd = p.frame_td + (size_t)w * F;
    const double mu = st.mu;
    const bool it0 = (st.iteration == 0);
    const double sqi = p.sqrt_info_px;
    const int fq = lane >> 2, fk = lane & 3;
    double* Yw = S.Y[wid];
    float4* sob = S.sobs[pr];
    double (*part)[32] = S.part[pr];
    const int barA = 1 + 2 * pr, barB = 2 + 2 * pr;

    // a lane's landmark of one task: packed table entry, state, and where the task's observation records are
    struct LaneLm { int l, L, ob; bool have, fx; double lam, s_l; bool contig; int ob0, nrec; };
    auto fetch_info = [&](int task) {   // packed table entry of this lane's landmark; w = -1: idle lane
      return (task < ntasks && lane < S.task_cnt[task]) ? lminfo[S.task_first[task] + lane] : make_int4(0, 1, 0, -1);
    };
    auto fetch_lm = [&](int task, int4 info) {
      LaneLm q;
      const int cnt = task < ntasks ? S.task_cnt[task] : 0;
      q.l = info.x; q.L = info.y; q.ob = info.z; q.fx = info.w > 0; q.have = info.w >= 0; q.lam = 1.0; q.s_l = 1.0;
      if (q.have) {
        q.lam = p.invdep[(size_t)w * p.Lm + q.l];
        if (!it0) q.s_l = p.lm_s[(size_t)w * p.Lm + q.l];
      }
      // contiguous run of observation records (landmark table in start-frame order) -> one bulk copy
      const int end = q.ob + q.L;
      const int prev_end = __shfl_up_sync(0xffffffffu, end, 1);
      const bool ok = !q.have || lane == 0 || q.ob == prev_end;
      q.ob0 = __shfl_sync(0xffffffffu, q.ob, 0);
      q.nrec = __shfl_sync(0xffffffffu, end, cnt > 0 ? cnt - 1 : 0) - q.ob0;
      q.contig = __all_sync(0xffffffffu, ok) && cnt > 0 && q.nrec > 0 && q.nrec <= kObsStage;
      return q;
    };
    auto stage_obs = [&](const LaneLm& q) {  // one elected lane of the pair
      if (q.contig && hf == 0 && lane == 0) {
        mbar_expect_tx(&S.obsbar[pr], (uint32_t)q.nrec * 16u);
        bulk_g2s(sob, obs + q.ob0, (uint32_t)q.nrec * 16u, &S.obsbar[pr]);
      }
    };
    // fetched ahead in two levels so that no dependent global load is waited for: the table entry two tasks ahead, the landmark's state
    // (and the bulk copy of its observations) one task ahead
    LaneLm nxt = fetch_lm(pr, fetch_info(pr));
    stage_obs(nxt);
    int4 info_n = fetch_info(pr + kPairs);
    uint32_t obs_par = 0;
    for (int task = pr; task < ntasks; task += kPairs) {
      const int stage = pr, use = task / kPairs;   // the pair's own stage, used for the use-th time
      double* WTs = S.WT[stage];
      const int i = S.task_start[task];
      const LaneLm cur = nxt;
      const bool have = cur.have;
      const int l = cur.l, L = cur.L; const bool fx = cur.fx;
      const int Lmax = __reduce_max_sync(0xffffffffu, L);
      const int orel = have ? cur.ob - cur.ob0 : 0;
      const float4* osrc = cur.contig ? (const float4*)sob + orel : obs + (have ? cur.ob : 0);
      if (cur.contig) { mbar_wait(&S.obsbar[pr], obs_par); obs_par ^= 1; }
      LmCtx lc; lc.Xw = mk3(0, 0, 0); lc.dXdl = mk3(0, 0, 0);
      if (have) landmark_ctx(S.fr[i], S.cam, osrc[0], ftd[i], cur.lam, lc);
      if (use > 0) mbar_wait(&S.empty[stage], (use - 1) & 1);  // the consumers are done with the stage's previous tile
      // rows of this lane's column that no step writes: left of the host frame (hf 0) and right of the longest track (hf 1)
      {
        const int a_min = (6 * i + kRowShift) >> 3;
        const int row0 = max(8 * a_min - kRowShift, 0);
        if (hf == 0) for (int c = row0; c < 6 * i; c++) WTs[c * kWTStride + lane] = 0.0;
        else for (int c = 6 * (i + Lmax); c < kWTRows - 2; c++) WTs[c * kWTStride + lane] = 0.0;
      }
      V3 ns = mk3(0, 0, 0);  // sum_k n_k  ->  w_i = [ ns ; Ri^T ((Xw - Pi) x ns) ]
      double v = 0.0, gl = 0.0;
      const int nh = (Lmax - hf) >> 1;         // steps of this half: k = 1 + hf + 2 m, m < nh
      int m = (pr * nh) / kPairs;              // pairs that share a host frame start at different steps: their atomic flushes hit different frame pairs
      for (int q = 0; q < nh; q++) {
        const int k = 1 + hf + 2 * m;
        m = m + 1 == nh ? 0 : m + 1;
        const int j = i + k;  // uniform across the warp
        const bool valid = have && k < L;
        const float4 oj = osrc[valid ? k : 0];
        const FrameCtx& fj = S.fr[j];
        const double dx = lc.Xw.x - fj.P[0], dy = lc.Xw.y - fj.P[1], dz = lc.Xw.z - fj.P[2];
        const double px = fj.A[0] * dx + fj.A[1] * dy + fj.A[2] * dz - S.cam.rtt[0];
        const double py = fj.A[3] * dx + fj.A[4] * dy + fj.A[5] * dz - S.cam.rtt[1];
        const double pzr = fj.A[6] * dx + fj.A[7] * dy + fj.A[8] * dz - S.cam.rtt[2];
        const double dt = S.cam.td - ftd[j];
        const double iz = fast_rcp(valid ? pzr : 1.0);
        double r0 = sqi * (px * iz - ((double)oj.x - dt * (double)oj.z));
        double r1 = sqi * (py * iz - ((double)oj.y - dt * (double)oj.w));
        // ceres::HuberLoss + Corrector as in huber(): outside the inlier region r = sqrt(sq), scale = sqrt(delta / r). The two reciprocal square
        // roots sit on the step's dependency chain, so they are evaluated only when some lane of the warp has an outlier (warp-uniform branch)
        const double sq = r0 * r0 + r1 * r1, hb = p.huber * p.huber;
        const bool outl = valid && sq > hb;
        double sc = 1.0, hr = 0.5 * sq;
        if (__any_sync(0xffffffffu, outl)) {
          const double ir = fast_rsqrt(fmax(sq, hb)), yy = fmax(2.2250738585072014e-308, p.huber * ir);
          sc = outl ? yy * fast_rsqrt(yy) : 1.0;
          hr = outl ? 0.5 * (2.0 * p.huber * (sq * ir) - hb) : 0.5 * sq;
        }
        const double msk = valid ? 1.0 : 0.0;
        cost_acc += msk * hr;
        r0 *= sc * msk; r1 *= sc * msk;
        // Jx = sc * sqrt_info * [[1/z, 0, -x/z^2], [0, 1/z, -y/z^2]] * A_j
        const double a = msk * sc * sqi * iz, bx = -a * px * iz, by = -a * py * iz;
        const double j00 = a * fj.A[0] + bx * fj.A[6], j01 = a * fj.A[1] + bx * fj.A[7], j02 = a * fj.A[2] + bx * fj.A[8];
        const double j10 = a * fj.A[3] + by * fj.A[6], j11 = a * fj.A[4] + by * fj.A[7], j12 = a * fj.A[5] + by * fj.A[8];
        // this lane's two rows of Y = [Jx | Jx [d]x | r], staged transposed for the Gram product;  (row [d]x)_c: (b1 dz - b2 dy, b2 dx - b0 dz, b0 dy - b1 dx)
        __syncwarp();
        Yw[0 * kYStride + lane] = j00; Yw[1 * kYStride + lane] = j01; Yw[2 * kYStride + lane] = j02;
        Yw[0 * kYStride + 32 + lane] = j10; Yw[1 * kYStride + 32 + lane] = j11; Yw[2 * kYStride + 32 + lane] = j12;
        Yw[3 * kYStride + lane] = j01 * dz - j02 * dy; Yw[4 * kYStride + lane] = j02 * dx - j00 * dz; Yw[5 * kYStride + lane] = j00 * dy - j01 * dx;
        Yw[3 * kYStride + 32 + lane] = j11 * dz - j12 * dy; Yw[4 * kYStride + 32 + lane] = j12 * dx - j10 * dz; Yw[5 * kYStride + 32 + lane] = j10 * dy - j11 * dx;
        Yw[6 * kYStride + lane] = r0; Yw[6 * kYStride + 32 + lane] = r1;
        __syncwarp();
        // landmark column: j_lambda = Jx dXw/dlambda (zero for a fixed landmark), n = Jx^T j_lambda, w_j = [ -n ; Rj^T (n x d) ]
        const double fm = fx ? 0.0 : 1.0;
        const double jl0 = fm * (j00 * lc.dXdl.x + j01 * lc.dXdl.y + j02 * lc.dXdl.z);
        const double jl1 = fm * (j10 * lc.dXdl.x + j11 * lc.dXdl.y + j12 * lc.dXdl.z);
        v += jl0 * jl0 + jl1 * jl1; gl += jl0 * r0 + jl1 * r1;
        const double nx = j00 * jl0 + j10 * jl1, ny = j01 * jl0 + j11 * jl1, nz = j02 * jl0 + j12 * jl1;
        ns.x += nx; ns.y += ny; ns.z += nz;
        const double qx = ny * dz - nz * dy, qy = nz * dx - nx * dz, qz = nx * dy - ny * dx;
        // unconditional: an idle lane stores the zeros these rows of its column hold anyway
        double* wt = &WTs[(6 * j) * kWTStride + lane];
        wt[0] = -nx; wt[kWTStride] = -ny; wt[2 * kWTStride] = -nz;
        wt[3 * kWTStride] = fj.R[0] * qx + fj.R[3] * qy + fj.R[6] * qz;
        wt[4 * kWTStride] = fj.R[1] * qx + fj.R[4] * qy + fj.R[7] * qz;
        wt[5 * kWTStride] = fj.R[2] * qx + fj.R[5] * qy + fj.R[8] * qz;
        // Gram matrix Y^T Y on the tensor cores: one 8x8 tile over the warp's 64 rows (K = 64), four independent accumulator chains; the A
        // and B fragment of a k-step are the same element of Y^T (row 7 of Y^T does not exist: zero)
        {
          const double* yf = Yw + fq * kYStride + fk;
          double c0 = 0, c1 = 0, e0 = 0, e1 = 0, c2 = 0, c3 = 0, e2 = 0, e3 = 0;
#pragma unroll
          for (int s = 0; s < 16; s += 4) {
            const double f0 = fq < 7 ? yf[4 * s] : 0.0, f1 = fq < 7 ? yf[4 * s + 4] : 0.0, f2 = fq < 7 ? yf[4 * s + 8] : 0.0, f3 = fq < 7 ? yf[4 * s + 12] : 0.0;
            mma_f64(c0, c1, f0, f0); mma_f64(e0, e1, f1, f1); mma_f64(c2, c3, f2, f2); mma_f64(e2, e3, f3, f3);
          }
          const double g0 = (c0 + e0) + (c2 + e2), g1 = (c1 + e1) + (c3 + e3);
          if (fq < 6) {
            double* mom = &S.Mom[pidx(i, j, F) * kMomStride + momidx(fq, fq)] - fq;  // entry (fq, n) at mom[n]
            if (2 * fk >= fq) atomicAdd(&mom[2 * fk], g0);
            if (2 * fk + 1 >= fq && fk < 3) atomicAdd(&mom[2 * fk + 1], g1);
          }
        }
      }
      // the pair's two halves of the landmark scalars meet: hf 1 hands its sums over, hf 0 forms v' and the host-frame rows and hands 1 / sqrt(v') back
      if (hf == 1) {
        part[0][lane] = v; part[1][lane] = gl; part[2][lane] = ns.x; part[3][lane] = ns.y; part[4][lane] = ns.z;
        bar_arrive(barA, 64);
      } else {
        bar_sync(barA, 64);
        v += part[0][lane]; gl += part[1][lane]; ns.x += part[2][lane]; ns.y += part[3][lane]; ns.z += part[4][lane];
      }
      // both halves have read the staged observations: the next task's records can land (issued a whole task ahead of their use)
      nxt = fetch_lm(task + kPairs, info_n);
      info_n = fetch_info(task + 2 * kPairs);
      if (hf == 0) {
        stage_obs(nxt);
        // landmark scalars: jacobi scale (iteration 0), regularised v' = v + mu * e; the landmark's column of W is scaled by 1/sqrt(v') so that
        // the SYRK needs no per-element multiply
        double rs = 0.0;
        double* wt = &WTs[(6 * i) * kWTStride + lane];
        if (have) {
          double s_l;
          if (it0) { s_l = 1.0 / (1.0 + sqrt(v)); p.lm_s[(size_t)w * p.Lm + l] = s_l; } else s_l = cur.s_l;
          const double d2 = fmin(fmax(s_l * s_l * v, 1e-6), 1e32);
          const double e = d2 / (s_l * s_l);
          const double vp = v + mu * e;
          rs = (!fx && v > 0.0) ? rsqrt(vp) : 0.0;
          p.lm_v[(size_t)w * p.Lm + l] = fx ? 0.0 : v;
          p.lm_g[(size_t)w * p.Lm + l] = fx ? 0.0 : gl;
          if (!fx) gmax = fmax(gmax, fabs(gl));
          const FrameCtx& fi = S.fr[i];
          const V3 e_i = mk3(lc.Xw.x - fi.P[0], lc.Xw.y - fi.P[1], lc.Xw.z - fi.P[2]);
          const V3 q = cross(e_i, ns);  // Gi^T ns = Ri^T ((Xw - Pi) x ns)
          wt[0] = ns.x * rs; wt[kWTStride] = ns.y * rs; wt[2 * kWTStride] = ns.z * rs;
          wt[3 * kWTStride] = (fi.R[0] * q.x + fi.R[3] * q.y + fi.R[6] * q.z) * rs;
          wt[4 * kWTStride] = (fi.R[1] * q.x + fi.R[4] * q.y + fi.R[7] * q.z) * rs;
          wt[5 * kWTStride] = (fi.R[2] * q.x + fi.R[5] * q.y + fi.R[8] * q.z) * rs;
          WTs[66 * kWTStride + lane] = fx ? 0.0 : gl * rs;
        } else {
#pragma unroll
          for (int c = 0; c < 6; c++) wt[c * kWTStride] = 0.0;
          WTs[66 * kWTStride + lane] = 0.0;
        }
        part[5][lane] = rs;
        bar_arrive(barB, 64);
        for (int k = 1; k < Lmax; k += 2) {
          double* ws = &WTs[(6 * (i + k)) * kWTStride + lane];
#pragma unroll
          for (int c = 0; c < 6; c++) ws[c * kWTStride] *= rs;
        }
      } else {
        bar_sync(barB, 64);
        const double rs = part[5][lane];
        for (int k = 2; k < Lmax; k += 2) {
          double* ws = &WTs[(6 * (i + k)) * kWTStride + lane];
#pragma unroll
          for (int c = 0; c < 6; c++) ws[c * kWTStride] *= rs;
        }
      }
      mbar_arrive(&S.full[stage]);
    }
    bar_sync(kBarProd, kProdThreads);   // the moments are complete, the Y staging is free -> U / sums

    // frame-pair moments -> off-diagonal blocks (i, j) and the un-rotated per-frame sums
    {
      const int npairs = F * (F - 1) / 2, nterms = F * (F - 1);
      double* terms = reinterpret_cast<double*>(&S.sobs[0][0]);   // the observation staging is free: every producer is past its last task
      for (int q = t; q < nterms + 4 * npairs; q += kProdThreads) {
        if (q < nterms) { const int f = q / (F - 1), r = q - f * (F - 1), o = r < f ? r : r + 1; expand_diag_term(S.Mom, S.fr, f, o, F, terms + diag_term_idx(f, o, F)); continue; }
        const int q2 = q - nterms, pq = q2 >> 2, sub = q2 & 3;
        int i = 0, rem = pq;
        while (rem >= F - 1 - i) { rem -= F - 1 - i; i++; }
        const int j = i + 1 + rem;
        expand_offdiag(&S.Mom[pq * kMomStride], S.fr[i], S.fr[j], sub, &S.U[ublk(i, j, F)]);
      }
    }
    bar_sync(kBarProd, kProdThreads);
    for (int q = t; q < F * kDiagTerm; q += kProdThreads) { const int f = q / kDiagTerm, e = q - f * kDiagTerm; S.sums[f * kSumStride + e] = expand_diag_reduce(reinterpret_cast<const double*>(&S.sobs[0][0]), f, e, F); }
    bar_sync(kBarProd, kProdThreads);
    // diagonal blocks (f, f) = [[PP, -PT Rf], [., Rf^T TT Rf]] and the gradient g_f = [GP ; Rf^T GT]
    for (int q = t; q < 42 * F; q += kProdThreads) {
      const int f = q / 42, e = q - 42 * f;
      const double* sf = S.sums + f * kSumStride;
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
    // LiDAR plane factors: tasks of <= 32 planes of one frame; [J | r]^T [J | r] (upper 21 + J^T r 6 = 27 sums) by a warp butterfly
    if (p.planes) {
      bar_sync(kBarProd, kProdThreads);
      const int npt = p.nptasks[w];
      const gf2_plane* pls = p.planes + (size_t)w * p.Pm;
      const int32_t* pperm = p.pperm + (size_t)w * p.Pm;
      // [J | r]^T [J | r] of one 6-dim pose block summed over the warp's planes -> diagonal block (f, f) and gradient
      auto accum_diag = [&](int f, const double* Jp, double r) {
        double m[32];
        {
          int c = 0;
#pragma unroll
          for (int a = 0; a < 6; a++)
#pragma unroll
            for (int b = a; b < 6; b++) m[c++] = Jp[a] * Jp[b];
#pragma unroll
          for (int a = 0; a < 6; a++) m[21 + a] = Jp[a] * r;
#pragma unroll
          for (int a = 27; a < 32; a++) m[a] = 0.0;
        }
        const double tot = butterfly32(m, lane);
        double* B = &S.U[ublk(f, f, F)];
        if (lane < 21) {
          int a = 0, rem = lane;
          while (rem >= 6 - a) { rem -= 6 - a; a++; }
          const int b = a + rem;
          atomicAdd(&B[a * 6 + b], tot);
          if (a != b) atomicAdd(&B[b * 6 + a], tot);
        } else if (lane < 27) atomicAdd(&S.g[6 * f + lane - 21], tot);
      };
      const double* palpha = p.plane_alpha ? p.plane_alpha + (size_t)w * p.Pm : nullptr;
      for (int q = wid; q < npt; q += kProdWarps) {
        const int key = p.ptask_frame[(size_t)w * kMaxPlaneTasks + q], cnt = p.ptask_cnt[(size_t)w * kMaxPlaneTasks + q], first = p.ptask_first[(size_t)w * kMaxPlaneTasks + q];
        const int f = key >> 1;
        if (!(key & 1)) {  // LidarPlaneNormFactor on the pose of frame f
          double Jp[6] = {0, 0, 0, 0, 0, 0}, r = 0.0;
          if (lane < cnt) { r = plane_residual(pls[pperm[first + lane]], S.fr[f], p.lidar_sqrt_info, Jp); cost_acc += 0.5 * r * r; }
          accum_diag(f, Jp, r);
        } else {           // CTLidarPlaneNormFactor between the poses of frames f (begin) and f + 1 (end): blocks (f,f), (f,f+1), (f+1,f+1)
          double Jc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, r = 0.0;
          if (lane < cnt) {
            const int pi = pperm[first + lane];
            r = ct_plane_residual(pls[pi], palpha ? palpha[pi] : 0.0, pose + 7 * f, pose + 7 * (f + 1), p.lidar_sqrt_info, Jc);
            cost_acc += 0.5 * r * r;
          }
          accum_diag(f, Jc, r);
          accum_diag(f + 1, Jc + 6, r);
          double* B = &S.U[ublk(f, f + 1, F)];   // rows: frame f, columns: frame f + 1
#pragma unroll
          for (int pass = 0; pass < 2; pass++) {
            double m[32];
#pragma unroll
            for (int e = 0; e < 32; e++) { const int idx = 32 * pass + e; m[e] = idx < 36 ? Jc[idx / 6] * Jc[6 + idx % 6] : 0.0; }
            const double tot = butterfly32(m, lane);
            const int idx = 32 * pass + lane;
            if (idx < 36) atomicAdd(&B[idx], tot);
          }
        }
      }
    }
  }
  __syncthreads();
  // S_vis = U - Schur, g_schur = Schur[:,66]
  const double* dense = &S.WT[0][0];
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
    Svis[idx] = S.U[ublk(a / 6, b / 6, F) + (a % 6) * 6 + (b % 6)] - dense[(a + kRowShift) * kNVP + b + kRowShift];
  }
  for (int q = t; q < NV; q += kLinThreads) {
    p.gvis[(size_t)w * kVisRec + q] = S.g[q];                       // full visual gradient J^T r (pose part)
    p.gschur[(size_t)w * kVisRec + q] = dense[(q + kRowShift) * kNVP + 66 + kRowShift];         // sum_l w_l g_l / v'_l, subtracted to form the reduced rhs
    p.Udiag[(size_t)w * kVisRec + q] = S.U[ublk(q / 6, q / 6, F) + (q % 6) * 7];
  }
  double red2[2] = {cost_acc, 0.0};
  block_sum<2>(red2, S.red);
  gmax = warp_max(gmax);
  if (lane == 0) S.red[wid] = gmax;
  __syncthreads();
  if (t == 0) {
    double gm = 0; for (int i2 = 0; i2 < kLinThreads / 32; i2++) gm = fmax(gm, S.red[i2]);
    p.c_lin[(size_t)w * kVisRec] = red2[0]; p.c_gmax[w] = gm;
  }
}

}  // namespace ws
}  // namespace gf2
