// gf2_solver.cu — C-ABI entry points of the batched sliding-window solver (include/gf2_abi.h) and the host-side
// orchestration: device memory, H2D/D2H of the reference's packed arrays, the trust-region kernel sequence and its
// CUDA-event timing. No CPU fallback: every compute call launches kernels on the handle's device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "gf2_solver_solve.cuh"
#include "gf2_solver_lin.cuh"
#include "gf2_solver_lin_ws.cuh"
#include "gf2_solver_marg.cuh"
#include "gf2_common.h"

using namespace gf2;

namespace gf2 {

// ------------------------------------------------------------------------------------------------ preintegration
// IntegrationBase::push_back chain (VE/factor/integration_base.h:39-167): mid-point integration with the 15x15 jacobian
// and covariance propagation  jacobian = F jacobian,  covariance = F covariance F^T + V noise V^T.
// One HALF-warp per interval (two intervals per warp: the per-sample quantities every lane of an interval forms - mid-point integration,
// the F and V blocks - cost a warp instruction each whether 16 or 32 lanes want them, so the second half-warp's interval rides for free;
// round 2: 3.0 -> 1.6 ms per 40,960 intervals). Lane c (< 15) of the half keeps column c of the jacobian and of the covariance in
// registers; F is block sparse (identity + seven 3x3 blocks), so F x is eight 3x3 mat-vecs per column; F C F^T = F (F C)^T uses one
// transpose through shared memory; V noise V^T comes from the 15x18 V staged in shared memory.
struct FBlocks { M3 f01, f03, f04, f11, f21, f23, f24; double dt; };

__device__ __forceinline__ void apply_F(const FBlocks& F, double (&x)[15]) {
  const V3 x0 = mk3(x[0], x[1], x[2]), x1 = mk3(x[3], x[4], x[5]), x2 = mk3(x[6], x[7], x[8]), x3 = mk3(x[9], x[10], x[11]), x4 = mk3(x[12], x[13], x[14]);
  const V3 n0 = x0 + mul(F.f01, x1) + F.dt * x2 + mul(F.f03, x3) + mul(F.f04, x4);
  const V3 n1 = mul(F.f11, x1) - F.dt * x4;
  const V3 n2 = mul(F.f21, x1) + x2 + mul(F.f23, x3) + mul(F.f24, x4);
  x[0] = n0.x; x[1] = n0.y; x[2] = n0.z; x[3] = n1.x; x[4] = n1.y; x[5] = n1.z; x[6] = n2.x; x[7] = n2.y; x[8] = n2.z;
}

constexpr int kPreWarps = 4;

__global__ void __launch_bounds__(32 * kPreWarps) k_imu_preintegrate(int n_intervals, int max_samples, const gf2_imu_sample* samples, const int32_t* n_samples,
                                                                      const double* first, const double* lin_bias, double acc_n, double gyr_n, double acc_w, double gyr_w,
                                                                      gf2_imu_preint* out) {
  __shared__ double T[2 * kPreWarps][15 * 16];
  __shared__ double Vs[2 * kPreWarps][15 * 18];
  const int hw = threadIdx.x >> 4, lane = threadIdx.x & 15;   // half-warp of the CTA, lane inside the half ("lane" below = column index)
  const int idx_raw = blockIdx.x * 2 * kPreWarps + hw;
  if (blockIdx.x * 2 * kPreWarps + (hw & ~1) >= n_intervals) return;   // the whole warp is past the end
  const bool live = idx_raw < n_intervals;                              // the odd half of the last warp may have no interval: it shadows its neighbour
  const int idx = live ? idx_raw : idx_raw - 1;
  const gf2_imu_sample* smp = samples + (size_t)idx * max_samples;
  V3 acc_0 = ld3(first + 6 * idx), gyr_0 = ld3(first + 6 * idx + 3);
  const V3 ba = ld3(lin_bias + 6 * idx), bg = ld3(lin_bias + 6 * idx + 3);
  V3 dp = mk3(0, 0, 0), dv = mk3(0, 0, 0); Q4 dq; dq.x = dq.y = dq.z = 0; dq.w = 1;
  double jc[15], cc[15];  // column `lane` of jacobian and covariance
#pragma unroll
  for (int r = 0; r < 15; r++) { jc[r] = (r == lane) ? 1.0 : 0.0; cc[r] = 0.0; }
  double sum_dt = 0;
  const double nz[6] = {acc_n * acc_n, gyr_n * gyr_n, acc_n * acc_n, gyr_n * gyr_n, acc_w * acc_w, gyr_w * gyr_w};
  double* Tw = T[hw]; double* Vw = Vs[hw];
  const int ns = n_samples[idx];
  const int ns_warp = max(ns, __shfl_xor_sync(0xffffffffu, ns, 16));   // the two halves step together; a half that has run out of samples idles
  for (int s = 0; s < ns_warp; s++) {
    const bool on = s < ns;
    if (!on) { __syncwarp(); __syncwarp(); __syncwarp(); continue; }   // (keeps the warp-level barriers below matched)
    const double dt = smp[s].dt;
    const V3 acc_1 = ld3(smp[s].acc), gyr_1 = ld3(smp[s].gyr);
    // midPointIntegration, integration_base.h:72-81 (every lane, uniform)
    const V3 un_acc_0 = qrot(dq, acc_0 - ba);
    const V3 un_gyr = 0.5 * (gyr_0 + gyr_1) - bg;
    Q4 hq; hq.w = 1; hq.x = un_gyr.x * dt / 2; hq.y = un_gyr.y * dt / 2; hq.z = un_gyr.z * dt / 2;
    const Q4 rq = qmul(dq, hq);
    const V3 un_acc_1 = qrot(rq, acc_1 - ba);
    const V3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
    const V3 rp = dp + dt * dv + (0.5 * dt * dt) * un_acc;
    const V3 rv = dv + dt * un_acc;
    // F and V blocks, integration_base.h:83-131 (result_delta_q not normalised here, as in the reference)
    const M3 R_w_x = skew(un_gyr), R_a_0_x = skew(acc_0 - ba), R_a_1_x = skew(acc_1 - ba);
    const M3 dR = toR(dq), rR = toR(rq), I = eye3();
    const M3 IwR = sub(I, scale(R_w_x, dt));
    const M3 rRa1 = mul(rR, R_a_1_x);
    FBlocks F;
    F.dt = dt;
    F.f01 = add(scale(mul(dR, R_a_0_x), -0.25 * dt * dt), scale(mul(rRa1, IwR), -0.25 * dt * dt));
    F.f03 = scale(add(dR, rR), -0.25 * dt * dt);
    F.f04 = scale(rRa1, -0.25 * dt * dt * -dt);
    F.f11 = IwR;
    F.f21 = add(scale(mul(dR, R_a_0_x), -0.5 * dt), scale(mul(rRa1, IwR), -0.5 * dt));
    F.f23 = scale(add(dR, rR), -0.5 * dt);
    F.f24 = scale(rRa1, -0.5 * dt * -dt);
    // V staged in shared memory (lanes 0..8 write the nine distinct 3x3 blocks' rows): [15][18]
    __syncwarp();
    for (int i = lane; i < 270; i += 16) Vw[i] = 0.0;
    __syncwarp();
    if (lane == 0) {
      put3(Vw, 18, 0, 0, dR, 0.25 * dt * dt);
      put3(Vw, 18, 0, 3, rRa1, -0.25 * dt * dt * 0.5 * dt);
      put3(Vw, 18, 0, 6, rR, 0.25 * dt * dt);
      put3(Vw, 18, 0, 9, rRa1, -0.25 * dt * dt * 0.5 * dt);
      put3(Vw, 18, 3, 3, I, 0.5 * dt);
      put3(Vw, 18, 3, 9, I, 0.5 * dt);
    } else if (lane == 1) {
      put3(Vw, 18, 6, 0, dR, 0.5 * dt);
      put3(Vw, 18, 6, 3, rRa1, -0.5 * dt * 0.5 * dt);
      put3(Vw, 18, 6, 6, rR, 0.5 * dt);
      put3(Vw, 18, 6, 9, rRa1, -0.5 * dt * 0.5 * dt);
      put3(Vw, 18, 9, 12, I, dt);
      put3(Vw, 18, 12, 15, I, dt);
    }
    if (lane < 15) {
      apply_F(F, jc);               // jacobian = F * jacobian
      apply_F(F, cc);               // column of F * covariance
#pragma unroll
      for (int r = 0; r < 15; r++) Tw[r * 16 + lane] = cc[r];
    }
    __syncwarp();
    if (lane < 15) {
#pragma unroll
      for (int k = 0; k < 15; k++) cc[k] = Tw[lane * 16 + k];   // column `lane` of (F C)^T
      apply_F(F, cc);                                          // column of F (F C)^T = F C F^T
#pragma unroll
      for (int r = 0; r < 15; r++) {                           // + V noise V^T
        double q = 0;
#pragma unroll
        for (int k = 0; k < 18; k++) q += nz[k / 3] * Vw[r * 18 + k] * Vw[lane * 18 + k];
        cc[r] += q;
      }
    }
    dp = rp; dv = rv; dq = qnormalized(rq);
    sum_dt += dt; acc_0 = acc_1; gyr_0 = gyr_1;
  }
  if (!live) return;
  gf2_imu_preint& o = out[idx];
  if (lane == 0) {
    o.sum_dt = sum_dt;
    o.delta_p[0] = dp.x; o.delta_p[1] = dp.y; o.delta_p[2] = dp.z;
    o.delta_q[0] = dq.x; o.delta_q[1] = dq.y; o.delta_q[2] = dq.z; o.delta_q[3] = dq.w;
    o.delta_v[0] = dv.x; o.delta_v[1] = dv.y; o.delta_v[2] = dv.z;
    o.lin_ba[0] = ba.x; o.lin_ba[1] = ba.y; o.lin_ba[2] = ba.z; o.lin_bg[0] = bg.x; o.lin_bg[1] = bg.y; o.lin_bg[2] = bg.z;
    o.valid = 1; o.pad_ = 0;
  }
  if (lane < 15) {
#pragma unroll
    for (int r = 0; r < 15; r++) { o.jacobian[r * 15 + lane] = jc[r]; o.covariance[r * 15 + lane] = cc[r]; }
  }
}

}  // namespace gf2

// ------------------------------------------------------------------------------------------------ NCCL (loaded at run time)
// libgf2_b200.so has no link-time NCCL dependency: the factor-sharded mode dlopen()s libnccl.so.2 (the copy torch already
// loaded when the caller is a torch.distributed process), so CPU-only boxes and single-GPU users never need it.
namespace {
typedef struct { char internal[128]; } gf2_nccl_uid;
// ---------------------------------------------------------------------------------------------------------------------
// Wheel-odometry preintegration (WheelIntegrationBase::push_back -> propagate -> midPointIntegration,
// VE/factor/wheel_integration_base.h:41-178). The chain is 6x6 and 5 samples long at 50 Hz / 10 Hz frames, so one thread
// per interval is enough (B * (F-1) threads); F = [[I, A], [0, ddR^T]] and the 6x12 V are used in their block form.
__global__ void __launch_bounds__(128) k_wheel_preintegrate(int total, int max_samples, const gf2_wheel_sample* __restrict__ samples, const int32_t* __restrict__ n_samples,
                                                             const double* __restrict__ first_sample, const double* __restrict__ lin, double vel_n, double gyr_n,
                                                             gf2_wheel_preint* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const gf2_wheel_sample* sm = samples + (size_t)t * max_samples;
  const int ns = n_samples[t];
  V3 vel_0 = ld3(first_sample + 6 * t), gyr_0 = ld3(first_sample + 6 * t + 3);
  const V3 lin_vel = vel_0, lin_gyr = gyr_0;
  const double sx = lin[4 * t], sy = lin[4 * t + 1], sw = lin[4 * t + 2], ltd = lin[4 * t + 3];
  V3 delta_p = mk3(0, 0, 0); Q4 delta_q; delta_q.x = delta_q.y = delta_q.z = 0; delta_q.w = 1;
  double jac[18], cov[36];
  for (int i = 0; i < 18; i++) jac[i] = 0;
  for (int i = 0; i < 36; i++) cov[i] = 0;
  double sum_dt = 0;
  const double nv = vel_n * vel_n, ng = gyr_n * gyr_n;
  for (int k = 0; k < ns; k++) {
    const double dt = sm[k].dt; const V3 vel_1 = ld3(sm[k].vel), gyr_1 = ld3(sm[k].gyr);
    const V3 sv0 = mk3(sx * vel_0.x, sy * vel_0.y, vel_0.z), sv1 = mk3(sx * vel_1.x, sy * vel_1.y, vel_1.z);
    const V3 un_vel_0 = qrot(delta_q, sv0);
    const V3 gsum = gyr_0 + gyr_1;
    const V3 un_gyr = (0.5 * sw) * gsum;
    Q4 ddq; ddq.w = 1; ddq.x = un_gyr.x * dt / 2; ddq.y = un_gyr.y * dt / 2; ddq.z = un_gyr.z * dt / 2;
    const Q4 rq = qmul(delta_q, ddq);
    const V3 un_vel_1 = qrot(rq, sv1);
    const V3 un_vel = 0.5 * (un_vel_0 + un_vel_1);
    const V3 result_delta_p = delta_p + un_vel * dt;
    const M3 dR = toR(delta_q), rR = toR(rq), ddR = toR(ddq);
    const M3 Rv0 = skew(sv0), Rv1 = skew(sv1);
    const M3 rRv1 = mul(rR, Rv1);
    const M3 A = scale(add(mul(dR, Rv0), mulBT(rRv1, ddR)), -0.5 * dt);   // F(0:3, 3:6)
    const M3 Bt = transpose(ddR);                                          // F(3:6, 3:6)
    const M3 Jr = rightJacobianSO3(un_gyr * dt);
    M3 Ssv = zero3(); Ssv.m[0] = sx; Ssv.m[4] = sy; Ssv.m[8] = 1;
    const M3 V00 = scale(mul(dR, Ssv), 0.5 * dt), V03 = scale(mul(rRv1, Jr), -0.25 * dt * dt), V06 = scale(mul(rR, Ssv), 0.5 * dt);
    const M3 V33 = scale(Jr, 0.5 * sw * dt);
    // jacobian wrt (sx, sy, sw)
    const V3 c0 = (0.5 * dt) * (mul(dR, mk3(vel_0.x, 0, 0)) + mul(rR, mk3(vel_1.x, 0, 0)));
    const V3 c1 = (0.5 * dt) * (mul(dR, mk3(0, vel_0.y, 0)) + mul(rR, mk3(0, vel_1.y, 0)));
    const V3 dr_last = mk3(jac[3 * 3 + 2], jac[4 * 3 + 2], jac[5 * 3 + 2]);
    const V3 dr_new = dr_last + mul(Jr, 0.5 * gsum) * dt;
    const V3 c2 = (0.5 * dt) * (mul(dR, cross(dr_last, sv0)) + mul(rR, cross(dr_new, sv1)));
    for (int i = 0; i < 3; i++) { jac[i * 3 + 0] += get(c0, i); jac[i * 3 + 1] += get(c1, i); jac[(3 + i) * 3 + 2] = get(dr_new, i); jac[i * 3 + 2] += get(c2, i); }
    // covariance = F cov F^T + V noise V^T with cov = [[P, Q], [Q^T, R]]
    M3 P, Q, QT, R;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { P.m[i * 3 + j] = cov[i * 6 + j]; Q.m[i * 3 + j] = cov[i * 6 + 3 + j]; QT.m[i * 3 + j] = cov[(3 + i) * 6 + j]; R.m[i * 3 + j] = cov[(3 + i) * 6 + 3 + j]; }
    // F cov = [[P + A Q^T, Q + A R], [B Q^T, B R]];  (F cov) F^T = [[(P + A QT) + (Q + A R) A^T, (Q + A R) B^T], [B QT + B R A^T, B R B^T]]
    const M3 X0 = add(P, mul(A, QT)), X1 = add(Q, mul(A, R)), Y0 = mul(Bt, QT), Y1 = mul(Bt, R);
    M3 N00 = add(X0, mulBT(X1, A)), N01 = mulBT(X1, Bt), N10 = add(Y0, mulBT(Y1, A)), N11 = mulBT(Y1, Bt);
    // V noise V^T: noise = diag(nv I, ng I, nv I, ng I); V = [[V00, V03, V06, V03], [0, V33, 0, V33]]
    N00 = add(N00, add(scale(add(mulBT(V00, V00), mulBT(V06, V06)), nv), scale(mulBT(V03, V03), 2.0 * ng)));
    const M3 c03 = scale(mulBT(V03, V33), 2.0 * ng);
    N01 = add(N01, c03); N10 = add(N10, transpose(c03));
    N11 = add(N11, scale(mulBT(V33, V33), 2.0 * ng));
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { cov[i * 6 + j] = N00.m[i * 3 + j]; cov[i * 6 + 3 + j] = N01.m[i * 3 + j]; cov[(3 + i) * 6 + j] = N10.m[i * 3 + j]; cov[(3 + i) * 6 + 3 + j] = N11.m[i * 3 + j]; }
    delta_p = result_delta_p; delta_q = qnormalized(rq);
    sum_dt += dt;
    vel_0 = vel_1; gyr_0 = gyr_1;
  }
  gf2_wheel_preint& r = out[t];
  r.sum_dt = sum_dt;
  r.delta_p[0] = delta_p.x; r.delta_p[1] = delta_p.y; r.delta_p[2] = delta_p.z;
  r.delta_q[0] = delta_q.x; r.delta_q[1] = delta_q.y; r.delta_q[2] = delta_q.z; r.delta_q[3] = delta_q.w;
  r.lin_sx = sx; r.lin_sy = sy; r.lin_sw = sw; r.lin_td = ltd;
  r.lin_vel[0] = lin_vel.x; r.lin_vel[1] = lin_vel.y; r.lin_vel[2] = lin_vel.z; r.lin_gyr[0] = lin_gyr.x; r.lin_gyr[1] = lin_gyr.y; r.lin_gyr[2] = lin_gyr.z;
  r.vel_1[0] = vel_0.x; r.vel_1[1] = vel_0.y; r.vel_1[2] = vel_0.z; r.gyr_1[0] = gyr_0.x; r.gyr_1[1] = gyr_0.y; r.gyr_1[2] = gyr_0.z;
  for (int i = 0; i < 18; i++) r.jacobian[i] = jac[i];
  for (int i = 0; i < 36; i++) r.covariance[i] = cov[i];
  r.valid = 1; r.pad_ = 0;
}

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(gf2_nccl_uid*) = nullptr;
  int (*CommInitRank)(void**, int, gf2_nccl_uid, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load() {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return false;
    GetUniqueId = (int (*)(gf2_nccl_uid*))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (int (*)(void**, int, gf2_nccl_uid, int))dlsym(lib, "ncclCommInitRank");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    ReduceScatter = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclReduceScatter");
    AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
    GroupStart = (int (*)())dlsym(lib, "ncclGroupStart"); GroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
    CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && AllReduce && ReduceScatter && AllGather && GroupStart && GroupEnd && CommDestroy;
  }
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8, kNcclChar = 0, kNcclSum = 0, kNcclMax = 2;  // ncclDataType_t / ncclRedOp_t values of nccl.h
}  // namespace

// ------------------------------------------------------------------------------------------------ handle
struct gf2_solver {
  gf2_solver_cfg cfg;
  int D;
  int sm_count = 148;
  int last_Dx = 0;   // reduced dimension of the last run (D, or 15 (F + 1) with free wheel calibration blocks)
  cudaStream_t stream, own_stream;
  double *snap_pose = nullptr, *snap_sb = nullptr, *snap_invdep = nullptr;
  KP kp;  // device pointers for window 0
  std::vector<void*> allocs;
  // raw-sample staging for preintegration
  gf2_imu_sample* d_imu_samples = nullptr; int32_t* d_imu_n = nullptr; double *d_imu_first = nullptr, *d_imu_bias = nullptr;
  gf2_imu_preint* d_imu = nullptr;
  gf2_wheel_preint* d_wheel = nullptr;
  gf2_wheel_sample* d_wheel_samples = nullptr; int32_t* d_wheel_n = nullptr; double *d_wheel_first = nullptr, *d_wheel_lin = nullptr;
  MargP marg = {}; int32_t* h_marg = nullptr; double marg_ms = 0;  // marginalization work buffers (allocated on first gf2_marginalize)
  void* nccl_comm = nullptr; int comm_rank = 0, comm_size = 1;
  bool has_imu = false, has_wheel = false, has_prior = false, has_planes = false, dump_full = false;
  std::vector<int32_t> h_obeg;
  float2* d_xy = nullptr;   // staging of gf2_set_observations_xy (allocated on its first call)
  cudaEvent_t ev[4 * 64 + 3];
  cudaEvent_t ev_marg[3] = {nullptr, nullptr, nullptr}; bool marg_pending = false;   // gf2_marginalize_async / _wait
  cudaEvent_t ev_nccl[8 * 64];   // pairs around the collectives of the factor-sharded mode (created by gf2_comm_init)
  double timing[8];
  WinState* h_state = nullptr;
};

template <typename T>
static int dalloc(gf2_solver* h, T** p, size_t count) {
  void* q = nullptr;
  if (count == 0) count = 1;
  if (cudaMalloc(&q, count * sizeof(T)) != cudaSuccess) return gf2::fail(GF2_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(cudaGetLastError()));
  cudaMemset(q, 0, count * sizeof(T));
  h->allocs.push_back(q);
  *p = (T*)q;
  return GF2_OK;
}

#define GF2_TRY(x) do { int rc_ = (x); if (rc_ != GF2_OK) return rc_; } while (0)
#define GF2_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return gf2::fail(GF2_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } while (0)

extern "C" {

const char* gf2_last_error(void) { return gf2::last_error(); }
int gf2_abi_version(void) { return GF2_ABI_VERSION; }
int gf2_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

void* gf2_host_alloc(size_t bytes) { void* p = nullptr; if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void gf2_host_free(void* p) { if (p) cudaFreeHost(p); }

int gf2_solver_create(const gf2_solver_cfg* cfg, gf2_solver** out) {
  if (!cfg || !out) return gf2::fail(GF2_ERR_INVALID, "null argument");
  if (cfg->n_frames < 2 || cfg->n_frames > GF2_MAX_FRAMES) return gf2::fail(GF2_ERR_INVALID, "n_frames %d out of range [2, %d]", cfg->n_frames, GF2_MAX_FRAMES);
  if (cfg->sweep < GF2_SWEEP_AUTO || cfg->sweep > GF2_SWEEP_WINDOW) return gf2::fail(GF2_ERR_INVALID, "sweep %d is not a GF2_SWEEP_* value", cfg->sweep);
  if (cfg->max_landmarks < 1 || cfg->max_landmarks > GF2_MAX_LANDMARKS) return gf2::fail(GF2_ERR_INVALID, "max_landmarks %d out of range", cfg->max_landmarks);
  if (cfg->max_windows < 1 || cfg->max_obs < 1) return gf2::fail(GF2_ERR_INVALID, "max_windows / max_obs must be positive");
  if (gf2_device_count() <= cfg->device) return gf2::fail(GF2_ERR_CUDA, "CUDA device %d not available (no CPU fallback exists)", cfg->device);
  GF2_CUDA(cudaSetDevice(cfg->device));
  gf2_solver* h = new gf2_solver();
  h->cfg = *cfg;
  const int B = cfg->max_windows, F = cfg->n_frames, Lm = cfg->max_landmarks, Om = cfg->max_obs, Pm = cfg->max_planes;
  h->D = 15 * F;
  memset(&h->kp, 0, sizeof(KP));
  KP& k = h->kp;
  k.nW = B; k.F = F; k.Lm = Lm; k.Om = Om; k.Pm = Pm; k.D = h->D; k.Ds = 15 * (F + 1); k.wcal = 0; k.use_wheel = cfg->use_wheel;
  k.Pr = cfg->max_prior_rows > 0 ? cfg->max_prior_rows : GF2_MAX_PRIOR_DIM;
  if (k.Pr > GF2_MAX_PRIOR_DIM) { delete h; return gf2::fail(GF2_ERR_INVALID, "max_prior_rows %d exceeds %d", k.Pr, GF2_MAX_PRIOR_DIM); }
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return gf2::fail(GF2_ERR_CUDA, "stream creation failed"); }
  h->stream = h->own_stream;
  int rc = GF2_OK;
#define A(ptr, type, count) if (rc == GF2_OK) rc = dalloc<type>(h, (type**)&(ptr), (size_t)(count))
  A(k.pose, double, (size_t)B * F * 7); A(k.sb, double, (size_t)B * F * 9); A(k.ex, double, (size_t)B * 7); A(k.td, double, B);
  A(k.exw, double, (size_t)B * 7); A(k.sxw, double, (size_t)B * 3); A(k.tdw, double, B);
  A(k.invdep, double, (size_t)B * Lm); A(k.pose_c, double, (size_t)B * F * 7); A(k.sb_c, double, (size_t)B * F * 9); A(k.invdep_c, double, (size_t)B * Lm);
  A(k.nlm, int32_t, B); A(k.start, int32_t, (size_t)B * Lm); A(k.tlen, int32_t, (size_t)B * Lm); A(k.obeg, int32_t, (size_t)B * Lm);
  A(k.fixed, uint8_t, (size_t)B * Lm); A(k.obs, float4, (size_t)B * Om); A(k.frame_td, double, (size_t)B * F);
  A(h->d_imu, gf2_imu_preint, (size_t)B * (F - 1)); A(k.imu_sqrt, double, (size_t)B * (F - 1) * 225);
  if (cfg->use_wheel) { A(h->d_wheel, gf2_wheel_preint, (size_t)B * (F - 1)); A(k.wheel_sqrt, double, (size_t)B * (F - 1) * 36); }
  if (cfg->use_wheel && cfg->max_wheel_samples > 0) {
    A(h->d_wheel_samples, gf2_wheel_sample, (size_t)B * (F - 1) * cfg->max_wheel_samples); A(h->d_wheel_n, int32_t, (size_t)B * (F - 1));
    A(h->d_wheel_first, double, (size_t)B * (F - 1) * 6); A(h->d_wheel_lin, double, (size_t)B * (F - 1) * 4);
  }
  A(k.prior_rows, int32_t, B); A(k.prior_nblocks, int32_t, B); A(k.prior_J0, double, (size_t)B * k.Pr * k.Pr); A(k.prior_r0, double, (size_t)B * k.Pr);
  A(k.prior_blocks, gf2_prior_block, (size_t)B * (2 * F + 8)); A(k.prior_H, double, (size_t)B * k.Pr * k.Pr); A(k.prior_map, int32_t, (size_t)B * k.Pr);
  if (Pm > 0) { A(k.n_planes, int32_t, B); A(k.planes, gf2_plane, (size_t)B * Pm); A(k.plane_alpha, double, (size_t)B * Pm); }
  A(k.Svis, double, (size_t)B * kVisRec);   // one record per window: Svis | gvis | gschur | Udiag | c_lin
  A(k.lm_v, double, (size_t)B * Lm); A(k.lm_g, double, (size_t)B * Lm); A(k.lm_s, double, (size_t)B * Lm); A(k.lm_z, double, (size_t)B * Lm);
  A(k.sx, double, (size_t)B * k.Ds); A(k.zx, double, (size_t)B * k.Ds); A(k.ux, double, (size_t)B * k.Ds); A(k.ex_diag, double, (size_t)B * k.Ds);
  A(k.imu_H, double, (size_t)B * (F - 1) * 675); A(k.imu_g, double, (size_t)B * (F - 1) * 30);
  A(k.prior_g, double, (size_t)B * k.Pr); A(k.cost_nv, double, B);
  A(k.trace, double, (size_t)B * 64 * 6);
  A(k.c_gmax, double, B); A(k.c_sums, double, (size_t)B * 8); A(k.c_cand, double, (size_t)B * 4);
  if (Pm > 0) { A(k.pperm, int32_t, (size_t)B * Pm); A(k.ptask_first, int32_t, (size_t)B * kMaxPlaneTasks); A(k.ptask_cnt, int32_t, (size_t)B * kMaxPlaneTasks); A(k.ptask_frame, int32_t, (size_t)B * kMaxPlaneTasks); A(k.nptasks, int32_t, B); }
  if (cfg->use_wheel) {
    A(k.wheel_H, double, (size_t)B * (F - 1) * 108); A(k.wheel_g, double, (size_t)B * (F - 1) * 12);
    A(k.wheel_Hc, double, (size_t)B * (F - 1) * 220); A(k.wheel_gc, double, (size_t)B * (F - 1) * 10);
    A(k.exw_c, double, (size_t)B * 7); A(k.sxw_c, double, (size_t)B * 3); A(k.tdw_c, double, B);
  }
  A(k.lminfo, int4, (size_t)B * Lm); A(k.task_first, int32_t, (size_t)B * kMaxTasks); A(k.task_cnt, int32_t, (size_t)B * kMaxTasks);
  A(k.task_start, int32_t, (size_t)B * kMaxTasks); A(k.ntasks, int32_t, B);
  A(k.st, WinState, B);
  if (cfg->max_imu_samples > 0) {
    A(h->d_imu_samples, gf2_imu_sample, (size_t)B * (F - 1) * cfg->max_imu_samples); A(h->d_imu_n, int32_t, (size_t)B * (F - 1));
    A(h->d_imu_first, double, (size_t)B * (F - 1) * 6); A(h->d_imu_bias, double, (size_t)B * (F - 1) * 6);
  }
#undef A
  if (rc != GF2_OK) { gf2_solver_destroy(h); return rc; }
  k.gvis = k.Svis + kNVMax * kNVMax; k.gschur = k.gvis + kNVP; k.Udiag = k.gschur + kNVP; k.c_lin = k.Udiag + kNVMax;
  cudaMemset((void*)k.prior_rows, 0, sizeof(int32_t) * B); cudaMemset((void*)k.prior_nblocks, 0, sizeof(int32_t) * B);
  for (auto& e : h->ev) cudaEventCreate(&e);
  for (auto& e : h->ev_marg) cudaEventCreate(&e);
  cudaHostAlloc((void**)&h->h_state, sizeof(WinState) * B, cudaHostAllocDefault);
  // opt in to large dynamic shared memory
  cudaFuncSetAttribute(k_linearize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LinShared));
  cudaFuncSetAttribute(ws::k_linearize_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ws::LinShared));
  cudaFuncSetAttribute(k_solve2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Solve2Shared) + sizeof(double) * solve2_matrix_doubles(F, 1)));
  cudaFuncSetAttribute(k_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 450 * GF2_MAX_FRAMES));
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
  // function attributes are per device: set here, after cudaSetDevice, for every handle (not once per process)
  cudaFuncSetAttribute(k_marg_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((sizeof(MargShared) + 15) & ~size_t(15)) + sizeof(double) * kMargTMax * kMargLD));
  cudaFuncSetAttribute(k_marg_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((sizeof(MargShared) + 15) & ~size_t(15)) + sizeof(double) * kMargTMax * kMargLD));
  cudaFuncSetAttribute(k_marg_eig, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((sizeof(EigShared) + 15) & ~size_t(15)) + sizeof(double) * 2 * kMargKMax * (kMargKMax | 1)));
  if (cudaGetLastError() != cudaSuccess) { gf2_solver_destroy(h); return gf2::fail(GF2_ERR_CUDA, "cudaFuncSetAttribute failed (is this an sm_100a device?)"); }
  *out = h;
  return GF2_OK;
}

void gf2_solver_destroy(gf2_solver* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs) cudaFree(p);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_nccl) if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_marg) if (e) cudaEventDestroy(e);
  if (h->h_state) cudaFreeHost(h->h_state);
  if (h->h_marg) cudaFreeHost(h->h_marg);
  if (h->nccl_comm && g_nccl.lib) g_nccl.CommDestroy(h->nccl_comm);
  cudaStreamDestroy(h->own_stream);
  delete h;
}

static int check_range(gf2_solver* h, int first, int n) {
  if (!h) return gf2::fail(GF2_ERR_INVALID, "null handle");
  if (first < 0 || n < 0 || first + n > h->cfg.max_windows) return gf2::fail(GF2_ERR_INVALID, "window range [%d, %d) outside capacity %d", first, first + n, h->cfg.max_windows);
  cudaSetDevice(h->cfg.device);
  return GF2_OK;
}
#define H2D(dst, src, bytes) do { if ((src) && (bytes) > 0) GF2_CUDA(cudaMemcpyAsync((void*)(dst), (src), (bytes), cudaMemcpyHostToDevice, h->stream)); } while (0)
#define D2H(dst, src, bytes) do { if ((dst) && (bytes) > 0) GF2_CUDA(cudaMemcpyAsync((dst), (const void*)(src), (bytes), cudaMemcpyDeviceToHost, h->stream)); } while (0)

int gf2_set_states(gf2_solver* h, int first, int n, const double* para_pose, const double* para_speedbias, const double* ex_pose,
                   const double* td, const double* ex_pose_wheel, const double* sxsysw, const double* td_wheel) {
  GF2_TRY(check_range(h, first, n));
  const KP& k = h->kp; const int F = k.F;
  H2D(k.pose + (size_t)first * F * 7, para_pose, sizeof(double) * n * F * 7);
  H2D(k.sb + (size_t)first * F * 9, para_speedbias, sizeof(double) * n * F * 9);
  H2D(k.ex + (size_t)first * 7, ex_pose, sizeof(double) * n * 7);
  H2D(k.td + first, td, sizeof(double) * n);
  if (h->cfg.use_wheel) {
    H2D(k.exw + (size_t)first * 7, ex_pose_wheel, sizeof(double) * n * 7);
    H2D(k.sxw + (size_t)first * 3, sxsysw, sizeof(double) * n * 3);
    H2D(k.tdw + first, td_wheel, sizeof(double) * n);
  }
  return GF2_OK;
}

int gf2_set_landmarks(gf2_solver* h, int first, int n, const int32_t* n_landmarks, const double* inv_depth, const int32_t* start_frame,
                      const int32_t* track_len, const uint8_t* fixed, const gf2_obs* obs, const double* frame_td) {
  GF2_TRY(check_range(h, first, n));
  const KP& k = h->kp; const int F = k.F, Lm = k.Lm, Om = k.Om;
  if (!n_landmarks || !inv_depth || !start_frame || !track_len || !frame_td) return gf2::fail(GF2_ERR_INVALID, "null landmark array");
  // validate + exclusive prefix sum of track_len on the host (index bookkeeping, bit-exact by construction)
  h->h_obeg.assign((size_t)n * Lm, 0);
  for (int w = 0; w < n; w++) {
    const int nl = n_landmarks[w];
    if (nl < 0 || nl > Lm) return gf2::fail(GF2_ERR_INVALID, "window %d: n_landmarks %d exceeds capacity %d", first + w, nl, Lm);
    int acc = 0;
    for (int l = 0; l < nl; l++) {
      const int s = start_frame[(size_t)w * Lm + l], L = track_len[(size_t)w * Lm + l];
      if (s < 0 || L < 1 || s + L > F) return gf2::fail(GF2_ERR_INVALID, "window %d landmark %d: track [%d, %d) outside the %d frames", first + w, l, s, s + L, F);
      h->h_obeg[(size_t)w * Lm + l] = acc; acc += L;
    }
    if (acc > Om) return gf2::fail(GF2_ERR_INVALID, "window %d: %d observations exceed capacity %d", first + w, acc, Om);
  }
  H2D(k.nlm + first, n_landmarks, sizeof(int32_t) * n);
  H2D(k.invdep + (size_t)first * Lm, inv_depth, sizeof(double) * n * Lm);
  H2D(k.start + (size_t)first * Lm, start_frame, sizeof(int32_t) * n * Lm);
  H2D(k.tlen + (size_t)first * Lm, track_len, sizeof(int32_t) * n * Lm);
  H2D(k.obeg + (size_t)first * Lm, h->h_obeg.data(), sizeof(int32_t) * n * Lm);
  if (fixed) H2D(k.fixed + (size_t)first * Lm, fixed, sizeof(uint8_t) * n * Lm);
  else GF2_CUDA(cudaMemsetAsync((void*)(k.fixed + (size_t)first * Lm), 0, (size_t)n * Lm, h->stream));
  H2D(k.obs + (size_t)first * Om, obs, sizeof(gf2_obs) * n * Om);
  H2D(k.frame_td + (size_t)first * F, frame_td, sizeof(double) * n * F);
  GF2_CUDA(cudaStreamSynchronize(h->stream));  // h_obeg is reused by the next call
  return GF2_OK;
}

// Positions only: 8 B per observation over the bus instead of 16; the records on the device keep their layout (velocity 0).
__global__ void k_expand_xy(const float2* __restrict__ xy, float4* __restrict__ obs, size_t count) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    const float2 p = xy[i];
    obs[i] = make_float4(p.x, p.y, 0.f, 0.f);
  }
}

int gf2_set_observations_xy(gf2_solver* h, int first, int n, const float* xy) {
  GF2_TRY(check_range(h, first, n));
  if (!xy) return gf2::fail(GF2_ERR_INVALID, "null observation array");
  const KP& k = h->kp; const int Om = k.Om;
  if (!h->d_xy) GF2_TRY(dalloc(h, &h->d_xy, (size_t)h->cfg.max_windows * Om));
  const size_t count = (size_t)n * Om;
  H2D(h->d_xy, xy, sizeof(float2) * count);
  const int blocks = (int)std::min<size_t>((count + 255) / 256, (size_t)8 * h->sm_count);
  k_expand_xy<<<blocks, 256, 0, h->stream>>>(h->d_xy, const_cast<float4*>(k.obs) + (size_t)first * Om, count);
  GF2_CUDA(cudaGetLastError());
  GF2_CUDA(cudaStreamSynchronize(h->stream));  // the staging area is reused by the next call
  return GF2_OK;
}

int gf2_set_imu(gf2_solver* h, int first, int n, const gf2_imu_preint* preint) {
  GF2_TRY(check_range(h, first, n));
  H2D(h->d_imu + (size_t)first * (h->kp.F - 1), preint, sizeof(gf2_imu_preint) * n * (h->kp.F - 1));
  h->has_imu = preint != nullptr;
  return GF2_OK;
}

int gf2_imu_preintegrate(gf2_solver* h, int first, int n, const gf2_imu_sample* samples, const int32_t* n_samples, const double* first_sample,
                         const double* lin_bias, const double noise[4]) {
  GF2_TRY(check_range(h, first, n));
  const int ms = h->cfg.max_imu_samples, Fm1 = h->kp.F - 1;
  if (ms <= 0) return gf2::fail(GF2_ERR_INVALID, "solver created with max_imu_samples = 0");
  if (!samples || !n_samples || !first_sample || !lin_bias || !noise) return gf2::fail(GF2_ERR_INVALID, "null argument");
  for (size_t i = 0; i < (size_t)n * Fm1; i++) if (n_samples[i] < 0 || n_samples[i] > ms) return gf2::fail(GF2_ERR_INVALID, "n_samples[%zu] = %d exceeds max_imu_samples %d", i, n_samples[i], ms);
  const size_t off = (size_t)first * Fm1;
  H2D(h->d_imu_samples + off * ms, samples, sizeof(gf2_imu_sample) * n * Fm1 * ms);
  H2D(h->d_imu_n + off, n_samples, sizeof(int32_t) * n * Fm1);
  H2D(h->d_imu_first + off * 6, first_sample, sizeof(double) * n * Fm1 * 6);
  H2D(h->d_imu_bias + off * 6, lin_bias, sizeof(double) * n * Fm1 * 6);
  const int total = n * Fm1;
  k_imu_preintegrate<<<(total + 2 * kPreWarps - 1) / (2 * kPreWarps), 32 * kPreWarps, 0, h->stream>>>(total, ms, h->d_imu_samples + off * ms, h->d_imu_n + off, h->d_imu_first + off * 6,
                                                               h->d_imu_bias + off * 6, noise[0], noise[1], noise[2], noise[3], h->d_imu + off);
  GF2_CUDA(cudaGetLastError());
  h->has_imu = true;
  return GF2_OK;
}

int gf2_imu_preintegrate_resident(gf2_solver* h, int first, int n, const double noise[4]) {
  GF2_TRY(check_range(h, first, n));
  const int ms = h->cfg.max_imu_samples, Fm1 = h->kp.F - 1;
  if (ms <= 0) return gf2::fail(GF2_ERR_INVALID, "solver created with max_imu_samples = 0");
  const size_t off = (size_t)first * Fm1; const int total = n * Fm1;
  k_imu_preintegrate<<<(total + 2 * kPreWarps - 1) / (2 * kPreWarps), 32 * kPreWarps, 0, h->stream>>>(total, ms, h->d_imu_samples + off * ms, h->d_imu_n + off, h->d_imu_first + off * 6,
                                                               h->d_imu_bias + off * 6, noise[0], noise[1], noise[2], noise[3], h->d_imu + off);
  GF2_CUDA(cudaGetLastError());
  h->has_imu = true;
  return GF2_OK;
}

int gf2_solver_set_stream(gf2_solver* h, void* cuda_stream) {
  if (!h) return gf2::fail(GF2_ERR_INVALID, "null handle");
  cudaStreamSynchronize(h->stream);
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return GF2_OK;
}

int gf2_snapshot_states(gf2_solver* h, int first, int n) {
  GF2_TRY(check_range(h, first, n));
  const KP& k = h->kp; const int B = h->cfg.max_windows;
  if (!h->snap_pose) { GF2_TRY(dalloc<double>(h, &h->snap_pose, (size_t)B * k.F * 7)); GF2_TRY(dalloc<double>(h, &h->snap_sb, (size_t)B * k.F * 9)); GF2_TRY(dalloc<double>(h, &h->snap_invdep, (size_t)B * k.Lm)); }
  GF2_CUDA(cudaMemcpyAsync(h->snap_pose + (size_t)first * k.F * 7, k.pose + (size_t)first * k.F * 7, sizeof(double) * n * k.F * 7, cudaMemcpyDeviceToDevice, h->stream));
  GF2_CUDA(cudaMemcpyAsync(h->snap_sb + (size_t)first * k.F * 9, k.sb + (size_t)first * k.F * 9, sizeof(double) * n * k.F * 9, cudaMemcpyDeviceToDevice, h->stream));
  GF2_CUDA(cudaMemcpyAsync(h->snap_invdep + (size_t)first * k.Lm, k.invdep + (size_t)first * k.Lm, sizeof(double) * n * k.Lm, cudaMemcpyDeviceToDevice, h->stream));
  return GF2_OK;
}
int gf2_restore_states(gf2_solver* h, int first, int n) {
  GF2_TRY(check_range(h, first, n));
  const KP& k = h->kp;
  if (!h->snap_pose) return gf2::fail(GF2_ERR_INVALID, "no snapshot taken");
  GF2_CUDA(cudaMemcpyAsync(k.pose + (size_t)first * k.F * 7, h->snap_pose + (size_t)first * k.F * 7, sizeof(double) * n * k.F * 7, cudaMemcpyDeviceToDevice, h->stream));
  GF2_CUDA(cudaMemcpyAsync(k.sb + (size_t)first * k.F * 9, h->snap_sb + (size_t)first * k.F * 9, sizeof(double) * n * k.F * 9, cudaMemcpyDeviceToDevice, h->stream));
  GF2_CUDA(cudaMemcpyAsync(k.invdep + (size_t)first * k.Lm, h->snap_invdep + (size_t)first * k.Lm, sizeof(double) * n * k.Lm, cudaMemcpyDeviceToDevice, h->stream));
  return GF2_OK;
}

int gf2_get_imu(gf2_solver* h, int first, int n, gf2_imu_preint* preint) {
  GF2_TRY(check_range(h, first, n));
  D2H(preint, h->d_imu + (size_t)first * (h->kp.F - 1), sizeof(gf2_imu_preint) * n * (h->kp.F - 1));
  GF2_CUDA(cudaStreamSynchronize(h->stream));
  return GF2_OK;
}

int gf2_set_wheel(gf2_solver* h, int first, int n, const gf2_wheel_preint* preint) {
  GF2_TRY(check_range(h, first, n));
  if (!h->cfg.use_wheel) return gf2::fail(GF2_ERR_INVALID, "solver created with use_wheel = 0");
  H2D(h->d_wheel + (size_t)first * (h->kp.F - 1), preint, sizeof(gf2_wheel_preint) * n * (h->kp.F - 1));
  h->has_wheel = preint != nullptr;
  return GF2_OK;
}

int gf2_wheel_preintegrate(gf2_solver* h, int first, int n, const gf2_wheel_sample* samples, const int32_t* n_samples, const double* first_sample,
                           const double* lin, const double noise[2]) {
  GF2_TRY(check_range(h, first, n));
  const int ms = h->cfg.max_wheel_samples, Fm1 = h->kp.F - 1;
  if (!h->cfg.use_wheel || ms <= 0) return gf2::fail(GF2_ERR_INVALID, "solver created with use_wheel = 0 or max_wheel_samples = 0");
  if (!samples || !n_samples || !first_sample || !lin || !noise) return gf2::fail(GF2_ERR_INVALID, "null argument");
  for (size_t i = 0; i < (size_t)n * Fm1; i++) if (n_samples[i] < 0 || n_samples[i] > ms) return gf2::fail(GF2_ERR_INVALID, "n_samples[%zu] = %d exceeds max_wheel_samples %d", i, n_samples[i], ms);
  const size_t off = (size_t)first * Fm1;
  H2D(h->d_wheel_samples + off * ms, samples, sizeof(gf2_wheel_sample) * n * Fm1 * ms);
  H2D(h->d_wheel_n + off, n_samples, sizeof(int32_t) * n * Fm1);
  H2D(h->d_wheel_first + off * 6, first_sample, sizeof(double) * n * Fm1 * 6);
  H2D(h->d_wheel_lin + off * 4, lin, sizeof(double) * n * Fm1 * 4);
  const int total = n * Fm1;
  k_wheel_preintegrate<<<(total + 127) / 128, 128, 0, h->stream>>>(total, ms, h->d_wheel_samples + off * ms, h->d_wheel_n + off, h->d_wheel_first + off * 6,
                                                                  h->d_wheel_lin + off * 4, noise[0], noise[1], h->d_wheel + off);
  GF2_CUDA(cudaGetLastError());
  h->has_wheel = true;
  return GF2_OK;
}

int gf2_get_wheel(gf2_solver* h, int first, int n, gf2_wheel_preint* preint) {
  GF2_TRY(check_range(h, first, n));
  if (!h->cfg.use_wheel) return gf2::fail(GF2_ERR_INVALID, "solver created with use_wheel = 0");
  D2H(preint, h->d_wheel + (size_t)first * (h->kp.F - 1), sizeof(gf2_wheel_preint) * n * (h->kp.F - 1));
  GF2_CUDA(cudaStreamSynchronize(h->stream));
  return GF2_OK;
}

int gf2_set_prior(gf2_solver* h, int first, int n, const int32_t* n_rows, const double* J0, const double* r0, const int32_t* n_blocks,
                  const gf2_prior_block* blocks) {
  GF2_TRY(check_range(h, first, n));
  const KP& k = h->kp;
  if (!n_rows) { h->has_prior = false; return GF2_OK; }
  for (int w = 0; w < n; w++) {
    if (n_rows[w] < 0 || n_rows[w] > k.Pr) return gf2::fail(GF2_ERR_INVALID, "prior of window %d has %d rows (max %d)", first + w, n_rows[w], k.Pr);
    if (n_rows[w] > 0 && (!n_blocks || !blocks || !J0 || !r0)) return gf2::fail(GF2_ERR_INVALID, "prior of window %d has rows but a null J0 / r0 / n_blocks / blocks array", first + w);
    if (n_rows[w] > 0 && (n_blocks[w] < 1 || n_blocks[w] > 2 * k.F + 8)) return gf2::fail(GF2_ERR_INVALID, "prior of window %d has %d blocks", first + w, n_blocks[w]);
    if (n_blocks && (n_blocks[w] < 0 || n_blocks[w] > 2 * k.F + 8)) return gf2::fail(GF2_ERR_INVALID, "prior of window %d has %d blocks", first + w, n_blocks[w]);
    // k_solve2 stores the speed-bias rows of frame I only against frames I-1 and I (structural zeros of S and of its Cholesky
    // factor): a prior that couples speed-bias k >= 2 to far frames would break that; the reference's priors keep speed-bias 0 only
    if (blocks && n_blocks) for (int b = 0; b < n_blocks[w]; b++) {
      const gf2_prior_block& pb = blocks[(size_t)w * (2 * k.F + 8) + b];
      if (pb.kind == GF2_BLK_SPEEDBIAS && pb.index >= 2) return gf2::fail(GF2_ERR_UNSUPPORTED, "prior of window %d keeps speed-bias %d (only 0 and 1 are supported)", first + w, pb.index);
    }
  }
  H2D(k.prior_rows + first, n_rows, sizeof(int32_t) * n);
  H2D(k.prior_nblocks + first, n_blocks, sizeof(int32_t) * n);
  H2D(k.prior_J0 + (size_t)first * k.Pr * k.Pr, J0, sizeof(double) * n * k.Pr * k.Pr);
  H2D(k.prior_r0 + (size_t)first * k.Pr, r0, sizeof(double) * n * k.Pr);
  H2D(k.prior_blocks + (size_t)first * (2 * k.F + 8), blocks, sizeof(gf2_prior_block) * n * (2 * k.F + 8));
  h->has_prior = true;
  return GF2_OK;
}

int gf2_set_planes(gf2_solver* h, int first, int n, const int32_t* n_planes, const gf2_plane* planes) {
  GF2_TRY(check_range(h, first, n));
  const KP& k = h->kp;
  if (k.Pm <= 0) return gf2::fail(GF2_ERR_INVALID, "solver created with max_planes = 0");
  if (!n_planes || !planes) return gf2::fail(GF2_ERR_INVALID, "null argument");
  if ((k.Pm + 31) / 32 + 2 * k.F > kMaxPlaneTasks) return gf2::fail(GF2_ERR_INVALID, "max_planes %d exceeds the task capacity", k.Pm);
  for (int w = 0; w < n; w++) if (n_planes[w] < 0 || n_planes[w] > k.Pm) return gf2::fail(GF2_ERR_INVALID, "n_planes[%d] = %d exceeds capacity %d", w, n_planes[w], k.Pm);
  for (int w = 0; w < n; w++) for (int q = 0; q < n_planes[w]; q++) {
    const int f = planes[(size_t)w * k.Pm + q].frame;
    if (f < 0 || f >= k.F) return gf2::fail(GF2_ERR_INVALID, "window %d plane %d: frame %d outside [0, %d)", first + w, q, f, k.F);
    if (planes[(size_t)w * k.Pm + q].ct && f + 1 >= k.F) return gf2::fail(GF2_ERR_INVALID, "window %d plane %d: a CT plane of frame %d needs the end pose %d", first + w, q, f, f + 1);
  }
  H2D(k.n_planes + first, n_planes, sizeof(int32_t) * n);   // nothing is uploaded unless every record is valid
  H2D(k.planes + (size_t)first * k.Pm, planes, sizeof(gf2_plane) * n * k.Pm);
  h->has_planes = true;
  return GF2_OK;
}

int gf2_set_plane_alpha(gf2_solver* h, int first, int n, const double* alpha) {
  GF2_TRY(check_range(h, first, n));
  const KP& k = h->kp;
  if (k.Pm <= 0) return gf2::fail(GF2_ERR_INVALID, "solver created with max_planes = 0");
  if (!alpha) return gf2::fail(GF2_ERR_INVALID, "null argument");
  H2D((double*)k.plane_alpha + (size_t)first * k.Pm, alpha, sizeof(double) * n * k.Pm);
  return GF2_OK;
}

static int fill_kp(gf2_solver* h, const gf2_solve_opts* o, KP& k) {
  if (!o) return gf2::fail(GF2_ERR_INVALID, "null options");
  if (!(o->max_time_s >= 0.0)) return gf2::fail(GF2_ERR_INVALID, "max_time_s must be >= 0 (0: no wall-clock cap)");
  if (o->initial_radius > 0 && o->initial_radius != 1e4) return gf2::fail(GF2_ERR_UNSUPPORTED, "initial_radius other than the Ceres default 1e4");
  if (o->max_iterations < 0 || o->max_iterations > 64) return gf2::fail(GF2_ERR_INVALID, "max_iterations %d out of range [0, 64]", o->max_iterations);
  const uint32_t need = GF2_CONST_EX_POSE | GF2_CONST_TD;
  if ((o->const_mask & need) != need) return gf2::fail(GF2_ERR_UNSUPPORTED, "free camera extrinsic / td blocks are not built yet (const_mask must hold EX_POSE|TD)");
  k = h->kp;
  // free wheel calibration blocks (estimate_wheel_extrinsic: 1 in the shipped wheel configs): one more block row of the reduced system
  k.wcal = 0;
  if (h->cfg.use_wheel && h->has_wheel) k.wcal = ((o->const_mask & GF2_CONST_EX_WHEEL) ? 0 : 1) | ((o->const_mask & GF2_CONST_WHEEL_INTRINSIC) ? 0 : 2) | ((o->const_mask & GF2_CONST_TD_WHEEL) ? 0 : 4);
  k.wsub = o->wheel_ext_const_components & 0x3fu;
  // (factor-sharded mode: the calibration block row is fed by the wheel factors and the prior only, which every rank holds; it rides in the
  //  all-gathered steps [z | u] of stride Ds like the frame rows)
  k.const_mask = o->const_mask; k.huber = o->huber_delta; k.sqrt_info_px = o->sqrt_info_px; k.g_norm = o->g_norm; k.lidar_sqrt_info = o->lidar_sqrt_info;
  k.ftol = o->function_tolerance > 0 ? o->function_tolerance : 1e-6;
  k.gtol = o->gradient_tolerance > 0 ? o->gradient_tolerance : 1e-10;
  k.ptol = o->parameter_tolerance > 0 ? o->parameter_tolerance : 1e-8;
  k.max_iterations = o->max_iterations;
  k.max_time_ns = o->max_time_s > 0.0 ? (unsigned long long)(o->max_time_s * 1e9 + 0.5) : 0ull;
  k.imu = h->has_imu ? h->d_imu : nullptr;
  k.wheel = (h->cfg.use_wheel && h->has_wheel) ? h->d_wheel : nullptr;
  if (!h->has_prior) { k.prior_rows = nullptr; }
  if (!h->has_planes) { k.planes = nullptr; }
  return GF2_OK;
}

static int run(gf2_solver* h, int first, int n, const gf2_solve_opts* opts, gf2_solve_summary* summaries, int iterations, bool only_linearize) {
  GF2_TRY(check_range(h, first, n));
  if (n == 0) return GF2_OK;
  KP k;
  GF2_TRY(fill_kp(h, opts, k));
  const int D = h->D;
  h->last_Dx = k.wcal ? k.Ds : h->D;
  const size_t sh_solve = sizeof(Solve2Shared) + sizeof(double) * solve2_matrix_doubles(k.F, k.wcal);
  int ne = 0, nn = 0;
  cudaEventRecord(h->ev[ne++], h->stream);
  k_prepare<<<n, 128, sizeof(double) * 450 * GF2_MAX_FRAMES, h->stream>>>(k, first);
  k_tasks<<<n, kTaskThreads, 0, h->stream>>>(k, first);
  cudaEventRecord(h->ev[ne++], h->stream);
  const int iters = only_linearize ? 1 : iterations;
  // One robot / small batches (up to two windows per SM) are launch- and latency-bound: the sweep runs one window per SM on 16 warp-specialised
  // warps (k_linearize_ws) and back-substitution + candidate + decision share one launch (k_step). A full batch gets its occupancy from the
  // batch: k_linearize (two windows per SM) and the three step kernels at their own occupancy. cfg.sweep overrides the choice of the sweep.
  const bool small = n <= 2 * h->sm_count;
  // (the window kernel holds one window per SM: up to one window per SM it runs a single wave, 0.062 vs 0.091 ms; beyond that the batch
  //  kernel's two windows per SM win)
  const bool sweep_ws = h->cfg.sweep == GF2_SWEEP_WINDOW || (h->cfg.sweep == GF2_SWEEP_AUTO && n <= h->sm_count);
  const bool fused_step = !h->nccl_comm && small;
  for (int it = 0; it < iters; it++) {
    if (sweep_ws) ws::k_linearize_ws<<<n, ws::kLinThreads, sizeof(ws::LinShared), h->stream>>>(k, first);
    else k_linearize<<<n, kLinThreads, sizeof(LinShared), h->stream>>>(k, first);
    // Factor-sharded mode, SURVEY 8(e). The sweep ran on this rank's landmarks / planes of ALL n windows. When n divides by the ranks the
    // windows' records [Svis | gvis | gschur | Udiag | visual cost] (36.6 KB each) are REDUCE-SCATTERED: rank r receives the summed records of
    // its n / N windows, assembles + factorises only those (k_nonvis, k_solve2: the reduced solve is sharded by window instead of being
    // replicated), and the steps [zx | ux] + trust-region states are ALL-GATHERED: same bytes on the wire as one all-reduce, 1 / N of the solve
    // work. Otherwise (and for the gf2_linearize dump): one all-reduce, every rank solves every window.
    const bool rs = h->nccl_comm && !only_linearize && n % h->comm_size == 0;
    const int n_own = rs ? n / h->comm_size : n, first_own = rs ? first + h->comm_rank * n_own : first;
    if (h->nccl_comm) {
      cudaEventRecord(h->ev_nccl[nn++], h->stream);
      g_nccl.GroupStart();
      if (rs) {
        g_nccl.ReduceScatter(k.Svis + (size_t)first * kVisRec, k.Svis + (size_t)first_own * kVisRec, (size_t)n_own * kVisRec, kNcclFloat64, kNcclSum, h->nccl_comm, h->stream);
        g_nccl.ReduceScatter(k.c_gmax + first, k.c_gmax + first_own, (size_t)n_own, kNcclFloat64, kNcclMax, h->nccl_comm, h->stream);
      } else {
        g_nccl.AllReduce(k.Svis + (size_t)first * kVisRec, k.Svis + (size_t)first * kVisRec, (size_t)n * kVisRec, kNcclFloat64, kNcclSum, h->nccl_comm, h->stream);
        g_nccl.AllReduce(k.c_gmax + first, k.c_gmax + first, (size_t)n, kNcclFloat64, kNcclMax, h->nccl_comm, h->stream);   // a MAX cannot ride in the SUM buffer
      }
      g_nccl.GroupEnd();
      cudaEventRecord(h->ev_nccl[nn++], h->stream);
    }
    cudaEventRecord(h->ev[ne++], h->stream);
    k_nonvis<<<n_own, kNonvisThreads, 0, h->stream>>>(k, first_own);
    k_solve2<<<n_own, kSolveThreads, sh_solve, h->stream>>>(k, first_own);
    if (rs) {
      cudaEventRecord(h->ev_nccl[nn++], h->stream);
      g_nccl.GroupStart();
      g_nccl.AllGather(k.zx + (size_t)first_own * k.Ds, k.zx + (size_t)first * k.Ds, (size_t)n_own * k.Ds, kNcclFloat64, h->nccl_comm, h->stream);
      g_nccl.AllGather(k.ux + (size_t)first_own * k.Ds, k.ux + (size_t)first * k.Ds, (size_t)n_own * k.Ds, kNcclFloat64, h->nccl_comm, h->stream);
      g_nccl.AllGather(k.st + first_own, k.st + first, (size_t)n_own * sizeof(WinState), kNcclChar, h->nccl_comm, h->stream);
      g_nccl.GroupEnd();
      cudaEventRecord(h->ev_nccl[nn++], h->stream);
    }
    cudaEventRecord(h->ev[ne++], h->stream);
    if (!only_linearize && fused_step) {
      // one robot / small batches: back-substitution, dogleg step, candidate evaluation and the accept / reject decision in one launch
      // (launch-latency bound there; a full batch runs the three kernels at their own, higher occupancy: measured 0.82 vs 0.87 ms per 4096 windows)
      cudaEventRecord(h->ev[ne++], h->stream);
      k_step<<<n, 288, 0, h->stream>>>(k, first);
      cudaEventRecord(h->ev[ne++], h->stream);
    } else if (!only_linearize) {
      k_backsub<<<n, 256, 0, h->stream>>>(k, first);
      if (h->nccl_comm) {
        cudaEventRecord(h->ev_nccl[nn++], h->stream);
        g_nccl.AllReduce(k.c_sums + (size_t)first * 8, k.c_sums + (size_t)first * 8, (size_t)n * 8, kNcclFloat64, kNcclSum, h->nccl_comm, h->stream);
        cudaEventRecord(h->ev_nccl[nn++], h->stream);
      }
      cudaEventRecord(h->ev[ne++], h->stream);
      k_cand_eval<<<n, 288, 0, h->stream>>>(k, first);
      if (h->nccl_comm) {
        cudaEventRecord(h->ev_nccl[nn++], h->stream);
        g_nccl.AllReduce(k.c_cand + (size_t)first * 4, k.c_cand + (size_t)first * 4, (size_t)n * 4, kNcclFloat64, kNcclSum, h->nccl_comm, h->stream);
        cudaEventRecord(h->ev_nccl[nn++], h->stream);
      }
      k_decide<<<n, 256, 0, h->stream>>>(k, first);
      cudaEventRecord(h->ev[ne++], h->stream);
    }
  }
  GF2_CUDA(cudaGetLastError());
  GF2_CUDA(cudaMemcpyAsync(h->h_state, k.st + first, sizeof(WinState) * n, cudaMemcpyDeviceToHost, h->stream));
  GF2_CUDA(cudaStreamSynchronize(h->stream));
  // timing
  memset(h->timing, 0, sizeof(h->timing));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[ne - 1]); h->timing[0] = ms;
  const int per = only_linearize ? 2 : 4;
  for (int it = 0; it < iters; it++) {
    const int b = 2 + it * per;
    cudaEventElapsedTime(&ms, h->ev[b - 1], h->ev[b]); h->timing[1] += ms;       // linearise (iteration 0 includes k_prepare)
    cudaEventElapsedTime(&ms, h->ev[b], h->ev[b + 1]); h->timing[2] += ms;       // solve
    if (!only_linearize) { cudaEventElapsedTime(&ms, h->ev[b + 1], h->ev[b + 3]); h->timing[3] += ms; }
  }
  h->timing[4] = 2 + iters * (only_linearize ? 3 : (fused_step ? 4 : 6)); h->timing[5] = iters;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->timing[6] = ms;  // k_prepare
  for (int q = 0; q + 1 < nn; q += 2) { cudaEventElapsedTime(&ms, h->ev_nccl[q], h->ev_nccl[q + 1]); h->timing[7] += ms; }   // time inside the NCCL collectives (factor-sharded mode)
  if (summaries) for (int w = 0; w < n; w++) {
    const WinState& s = h->h_state[w];
    summaries[w].initial_cost = s.initial_cost; summaries[w].final_cost = s.x_cost; summaries[w].iterations = s.iteration;
    summaries[w].successful_steps = s.successful + 1; summaries[w].termination = s.termination; summaries[w].pad_ = 0;
  }
  return GF2_OK;
}

int gf2_solve(gf2_solver* h, int first, int n, const gf2_solve_opts* opts, gf2_solve_summary* summaries) {
  if (!opts) return gf2::fail(GF2_ERR_INVALID, "null options");
  return run(h, first, n, opts, summaries, opts->max_iterations, false);
}

int gf2_linearize(gf2_solver* h, int first, int n, const gf2_solve_opts* opts) {
  if (!h) return gf2::fail(GF2_ERR_INVALID, "null handle");
  if (!h->kp.Sfull) {
    int rc = dalloc<double>(h, &h->kp.Sfull, (size_t)h->cfg.max_windows * h->kp.Ds * h->kp.Ds); if (rc) return rc;
    rc = dalloc<double>(h, &h->kp.gfull, (size_t)h->cfg.max_windows * h->kp.Ds); if (rc) return rc;
  }
  return run(h, first, n, opts, nullptr, 1, true);
}

// D = 15 F; with free wheel calibration blocks (opts->const_mask) one more block row [ex_wheel 6 | sx sy sw | td_wheel | 5 unused]
int gf2_reduced_dim(gf2_solver* h, const gf2_solve_opts* opts) {
  if (!h) return gf2::fail(GF2_ERR_INVALID, "null handle");
  KP k;
  if (opts && fill_kp(h, opts, k) == GF2_OK && k.wcal) return h->kp.Ds;
  return h->D;
}

int gf2_get_reduced_system(gf2_solver* h, int first, int n, double* S, double* g, double* cost) {
  GF2_TRY(check_range(h, first, n));
  if (!h->kp.Sfull) return gf2::fail(GF2_ERR_INVALID, "gf2_linearize has not been called");
  const int D = h->last_Dx, Ds = h->kp.Ds;   // the dimension of the last linearisation (gf2_reduced_dim for the same options)
  GF2_CUDA(cudaMemcpy2DAsync(S, sizeof(double) * D * D, h->kp.Sfull + (size_t)first * Ds * Ds, sizeof(double) * Ds * Ds, sizeof(double) * D * D, n, cudaMemcpyDeviceToHost, h->stream));
  GF2_CUDA(cudaMemcpy2DAsync(g, sizeof(double) * D, h->kp.gfull + (size_t)first * Ds, sizeof(double) * Ds, sizeof(double) * D, n, cudaMemcpyDeviceToHost, h->stream));
  GF2_CUDA(cudaMemcpyAsync(h->h_state, h->kp.st + first, sizeof(WinState) * n, cudaMemcpyDeviceToHost, h->stream));
  GF2_CUDA(cudaStreamSynchronize(h->stream));
  if (cost) for (int w = 0; w < n; w++) cost[w] = h->h_state[w].x_cost;
  return GF2_OK;
}

int gf2_get_states(gf2_solver* h, int first, int n, double* para_pose, double* para_speedbias, double* ex_pose, double* td,
                   double* ex_pose_wheel, double* sxsysw, double* td_wheel) {
  GF2_TRY(check_range(h, first, n));
  const KP& k = h->kp; const int F = k.F;
  D2H(para_pose, k.pose + (size_t)first * F * 7, sizeof(double) * n * F * 7);
  D2H(para_speedbias, k.sb + (size_t)first * F * 9, sizeof(double) * n * F * 9);
  D2H(ex_pose, k.ex + (size_t)first * 7, sizeof(double) * n * 7);
  D2H(td, k.td + first, sizeof(double) * n);
  if (h->cfg.use_wheel) {
    D2H(ex_pose_wheel, k.exw + (size_t)first * 7, sizeof(double) * n * 7);
    D2H(sxsysw, k.sxw + (size_t)first * 3, sizeof(double) * n * 3);
    D2H(td_wheel, k.tdw + first, sizeof(double) * n);
  }
  GF2_CUDA(cudaStreamSynchronize(h->stream));
  return GF2_OK;
}

int gf2_get_landmarks(gf2_solver* h, int first, int n, double* inv_depth) {
  GF2_TRY(check_range(h, first, n));
  D2H(inv_depth, h->kp.invdep + (size_t)first * h->kp.Lm, sizeof(double) * n * h->kp.Lm);
  GF2_CUDA(cudaStreamSynchronize(h->stream));
  return GF2_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int gf2_marginalize(gf2_solver* h, int first, int n, int32_t mode, const gf2_solve_opts* opts, int32_t* status, int32_t* m_dims) {
  GF2_TRY(gf2_marginalize_async(h, first, n, mode, opts));
  return gf2_marginalize_wait(h, first, n, status, m_dims);
}

int gf2_marginalize_async(gf2_solver* h, int first, int n, int32_t mode, const gf2_solve_opts* opts) {
  GF2_TRY(check_range(h, first, n));
  if (!opts) return gf2::fail(GF2_ERR_INVALID, "null options");
  if (mode != GF2_MARGIN_OLD && mode != GF2_MARGIN_SECOND_NEW) return gf2::fail(GF2_ERR_INVALID, "mode %d", mode);
  if (n == 0) return GF2_OK;
  KP k;
  GF2_TRY(fill_kp(h, opts, k));
  const int B = h->cfg.max_windows;
  if (!h->marg.A) {
    GF2_TRY(dalloc<double>(h, &h->marg.A, (size_t)B * kMargKMax * kMargKMax)); GF2_TRY(dalloc<double>(h, &h->marg.b, (size_t)B * kMargKMax));
    GF2_TRY(dalloc<int32_t>(h, &h->marg.touched, (size_t)B * kMargBlocksMax)); GF2_TRY(dalloc<int32_t>(h, &h->marg.status, B)); GF2_TRY(dalloc<int32_t>(h, &h->marg.mdim, B));
    GF2_CUDA(cudaMallocHost((void**)&h->h_marg, sizeof(int32_t) * 2 * B));
  }
  if (h->nccl_comm && !h->marg.stage_sum) {   // factor-sharded mode: the partial systems of the ranks meet in these buffers
    GF2_TRY(dalloc<double>(h, &h->marg.stage_sum, (size_t)B * kMargStageSum)); GF2_TRY(dalloc<double>(h, &h->marg.stage_max, (size_t)B * kMargStageMax));
  }
  MargP mp = h->marg;
  mp.mode = mode; mp.eig = opts->marg_eig ? 1 : 0;
  mp.nranks = h->nccl_comm ? h->comm_size : 1; mp.rank = h->nccl_comm ? h->comm_rank : 0;
  mp.out_rows = const_cast<int32_t*>(h->kp.prior_rows); mp.out_nblocks = const_cast<int32_t*>(h->kp.prior_nblocks);
  mp.out_J0 = const_cast<double*>(h->kp.prior_J0); mp.out_r0 = const_cast<double*>(h->kp.prior_r0);
  mp.out_blocks = const_cast<gf2_prior_block*>(h->kp.prior_blocks);
  const size_t sh_build = ((sizeof(MargShared) + 15) & ~size_t(15)) + sizeof(double) * kMargTMax * kMargLD;
  const int Kc = 6 * (k.F - 1) + 16 + (k.use_wheel ? 10 : 0);
  const size_t sh_eig = ((sizeof(EigShared) + 15) & ~size_t(15)) + sizeof(double) * 2 * Kc * (Kc | 1);
  cudaEventRecord(h->ev_marg[0], h->stream);
  k_prepare<<<n, 128, sizeof(double) * 450 * GF2_MAX_FRAMES, h->stream>>>(k, first);   // IMU sqrt_info, J0^T J0 of the old prior
  k_marg_build<<<n, kMargThreads, sh_build, h->stream>>>(k, first, mp);
  if (h->nccl_comm) {
    // every rank eliminated its own frame-0 landmarks (l mod N) into the window's [frame block | kept blocks] system: sum the partial
    // systems (82.7 KB per window) and take the maximum of the flags, then every rank eliminates the frame block and factorises the same
    // kept system: the new prior is replicated like the old one
    g_nccl.GroupStart();
    g_nccl.AllReduce(mp.stage_sum + (size_t)first * kMargStageSum, mp.stage_sum + (size_t)first * kMargStageSum, (size_t)n * kMargStageSum, kNcclFloat64, kNcclSum, h->nccl_comm, h->stream);
    g_nccl.AllReduce(mp.stage_max + (size_t)first * kMargStageMax, mp.stage_max + (size_t)first * kMargStageMax, (size_t)n * kMargStageMax, kNcclFloat64, kNcclMax, h->nccl_comm, h->stream);
    g_nccl.GroupEnd();
    k_marg_finish<<<n, kMargThreads, sh_build, h->stream>>>(k, first, mp);
  }
  k_marg_eig<<<n, kEigThreads, sh_eig, h->stream>>>(k, first, mp);
  cudaEventRecord(h->ev_marg[1], h->stream);
  GF2_CUDA(cudaGetLastError());
  GF2_CUDA(cudaMemcpyAsync(h->h_marg, mp.status + first, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->stream));
  GF2_CUDA(cudaMemcpyAsync(h->h_marg + B, mp.mdim + first, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->stream));
  cudaEventRecord(h->ev_marg[2], h->stream);
  h->has_prior = true;   // windows outside [first, first + n) keep prior_rows = 0 from creation unless gf2_set_prior filled them
  h->marg_pending = true;
  return GF2_OK;
}

int gf2_marginalize_wait(gf2_solver* h, int first, int n, int32_t* status, int32_t* m_dims) {
  GF2_TRY(check_range(h, first, n));
  if (n == 0) return GF2_OK;
  if (!h->marg_pending || !h->h_marg) return gf2::fail(GF2_ERR_INVALID, "gf2_marginalize_wait without a pending gf2_marginalize_async");
  const int B = h->cfg.max_windows;
  GF2_CUDA(cudaEventSynchronize(h->ev_marg[2]));
  h->marg_pending = false;
  float ms = 0; cudaEventElapsedTime(&ms, h->ev_marg[0], h->ev_marg[1]);
  h->marg_ms = ms;
  for (int i = 0; i < n; i++) { if (status) status[i] = h->h_marg[i]; if (m_dims) m_dims[i] = h->h_marg[B + i]; }
  return GF2_OK;
}

int gf2_get_prior(gf2_solver* h, int first, int n, int32_t* n_rows, double* J0, double* r0, int32_t* n_blocks, gf2_prior_block* blocks) {
  GF2_TRY(check_range(h, first, n));
  const KP& k = h->kp;
  D2H(n_rows, k.prior_rows + first, sizeof(int32_t) * n);
  D2H(J0, k.prior_J0 + (size_t)first * k.Pr * k.Pr, sizeof(double) * n * k.Pr * k.Pr);
  D2H(r0, k.prior_r0 + (size_t)first * k.Pr, sizeof(double) * n * k.Pr);
  D2H(n_blocks, k.prior_nblocks + first, sizeof(int32_t) * n);
  D2H(blocks, k.prior_blocks + (size_t)first * (2 * k.F + 8), sizeof(gf2_prior_block) * n * (2 * k.F + 8));
  GF2_CUDA(cudaStreamSynchronize(h->stream));
  return GF2_OK;
}

double gf2_last_marginalize_ms(gf2_solver* h) { return h ? h->marg_ms : 0.0; }

int gf2_get_trace(gf2_solver* h, int first, int n, double* out) {
  GF2_TRY(check_range(h, first, n));
  D2H(out, h->kp.trace + (size_t)first * 64 * 6, sizeof(double) * n * 64 * 6);
  GF2_CUDA(cudaStreamSynchronize(h->stream));
  return GF2_OK;
}

int gf2_comm_init(gf2_solver* h, int rank, int nranks, const void* nccl_unique_id) {
  if (!h || !nccl_unique_id) return gf2::fail(GF2_ERR_INVALID, "null argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) return gf2::fail(GF2_ERR_INVALID, "bad rank %d of %d", rank, nranks);
  if (!g_nccl.load()) return gf2::fail(GF2_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror());
  cudaSetDevice(h->cfg.device);
  gf2_nccl_uid uid; memcpy(&uid, nccl_unique_id, sizeof(uid));
  void* comm = nullptr;
  const int rc = g_nccl.CommInitRank(&comm, nranks, uid, rank);
  if (rc != 0) return gf2::fail(GF2_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
  if (h->nccl_comm) g_nccl.CommDestroy(h->nccl_comm);
  h->nccl_comm = nranks > 1 ? comm : nullptr;
  if (nranks == 1) g_nccl.CommDestroy(comm);
  h->comm_rank = rank; h->comm_size = nranks;
  if (h->nccl_comm && !h->ev_nccl[0]) for (auto& e : h->ev_nccl) cudaEventCreate(&e);
  return GF2_OK;
}
int gf2_comm_unique_id(void* out) {
  if (!out) return gf2::fail(GF2_ERR_INVALID, "null argument");
  if (!g_nccl.load()) return gf2::fail(GF2_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror());
  gf2_nccl_uid uid;
  const int rc = g_nccl.GetUniqueId(&uid);
  if (rc != 0) return gf2::fail(GF2_ERR_NCCL, "ncclGetUniqueId failed (%d)", rc);
  memcpy(out, &uid, sizeof(uid));
  return GF2_OK;
}

int gf2_last_timing(gf2_solver* h, double out[8]) {
  if (!h || !out) return gf2::fail(GF2_ERR_INVALID, "null argument");
  memcpy(out, h->timing, sizeof(h->timing));
  return GF2_OK;
}

}  // extern "C"
