// gf2_lio.cu — LIO factor construction on sm_100a: lidarodom::addSurfCostFactor (LIO/liw/lio/lidarodom.cpp:929-1071) for one scan.
//
//   k_lio_factors   one warp per keypoint:
//                   searchNeighbors (:1087-1165): the (2 nb + 1)^3 voxels around the keypoint (first 200 in the reference's x, y, z loop
//                   order) are looked up by binary search in the sorted key table, one voxel per lane; the max_number_neighbors nearest
//                   points are then extracted in ascending (distance, visit order) by repeated warp-wide arg-min — the same set and
//                   order as the reference's bounded max-heap with its strict `<` replacement test;
//                   computeNeighborhoodDistribution (:887-927): barycentre and covariance summed in neighbour order, cyclic Jacobi
//                   eigen-decomposition of the 3x3 covariance, normal = eigenvector of the smallest eigenvalue, planarity a2D;
//                   normal orientation, weight and the point-to-plane gate (:936-1008).
// The residual records are compacted in keypoint order with the max_num_residuals cap on the host (integer bookkeeping, like the
// reference's sequential loop). The voxel map itself (addPointToMap) stays with the caller; gf2_lio_set_map takes a snapshot.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "gf2_common.h"

namespace gf2 {

constexpr int kLioMaxNb = 32;      // max_number_neighbors capacity
constexpr int kLioMaxClosest = 4;  // num_closest_neighbors capacity
constexpr int kLioWarps = 4;

__host__ __device__ __forceinline__ unsigned long long lio_key(int x, int y, int z) {
  return ((unsigned long long)(unsigned)(x + 32768) << 32) | ((unsigned long long)(unsigned)(y + 32768) << 16) | (unsigned long long)(unsigned)(z + 32768);
}

struct LioRec { double normal[3], offset, weight; };

struct LioArgs {
  int n_keypoints, n_voxels, M;
  const unsigned long long* keys;  // sorted
  const int32_t* vox;              // sorted position -> voxel index of the snapshot
  const int32_t* n_points;         // [n_voxels] snapshot order
  const double* points;            // [n_voxels][M][3] snapshot order
  const gf2_lio_keypoint* kp;
  gf2_lio_opts o;
  int32_t* cnt;                    // [n_keypoints] residuals of the keypoint (-1: a2D is NaN)
  LioRec* rec;                     // [n_keypoints][kLioMaxClosest]
  double* neighbors;               // [n_keypoints][max_number_neighbors][3] or null
  int32_t* n_neighbors;            // [n_keypoints]
};

// cyclic Jacobi for a symmetric 3x3 (same sweep order and stopping rule as the test oracle's solver); evals ascending, evecs columns
__device__ void lio_eigen3(const double* Cin, double* evals, double* V) {
  double A[9];
  for (int i = 0; i < 9; i++) { A[i] = Cin[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 100; sweep++) {
    const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    const double dg = A[0] * A[0] + A[4] * A[4] + A[8] * A[8];
    if (off <= 1e-32 * (dg + 1e-300)) break;
    for (int p = 0; p < 3; p++) for (int q = p + 1; q < 3; q++) {
      const double apq = A[p * 3 + q]; if (apq == 0.0) continue;
      const double app = A[p * 3 + p], aqq = A[q * 3 + q];
      const double tau = (aqq - app) / (2.0 * apq);
      const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
      const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
      for (int k = 0; k < 3; k++) { const double akp = A[k * 3 + p], akq = A[k * 3 + q]; A[k * 3 + p] = c * akp - s * akq; A[k * 3 + q] = s * akp + c * akq; }
      for (int k = 0; k < 3; k++) { const double apk = A[p * 3 + k], aqk = A[q * 3 + k]; A[p * 3 + k] = c * apk - s * aqk; A[q * 3 + k] = s * apk + c * aqk; }
      for (int k = 0; k < 3; k++) { const double vkp = V[k * 3 + p], vkq = V[k * 3 + q]; V[k * 3 + p] = c * vkp - s * vkq; V[k * 3 + q] = s * vkp + c * vkq; }
    }
  }
  int idx[3] = {0, 1, 2};   // stable ascending sort of the diagonal
  for (int i = 1; i < 3; i++) for (int j = i; j > 0 && A[idx[j] * 4] < A[idx[j - 1] * 4]; j--) { const int t = idx[j]; idx[j] = idx[j - 1]; idx[j - 1] = t; }
  double W[9];
  for (int k = 0; k < 3; k++) { evals[k] = A[idx[k] * 4]; for (int r = 0; r < 3; r++) W[r * 3 + k] = V[r * 3 + idx[k]]; }
  for (int i = 0; i < 9; i++) V[i] = W[i];
}

__global__ void __launch_bounds__(32 * kLioWarps) k_lio_factors(LioArgs a) {
  __shared__ double s_nb[kLioWarps][kLioMaxNb][3];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * kLioWarps + wid;
  if (k >= a.n_keypoints) return;
  const gf2_lio_opts& o = a.o;
  const gf2_lio_keypoint kp = a.kp[k];
  const double px = kp.point[0], py = kp.point[1], pz = kp.point[2];
  const int kx = (short)(int)(px / o.size_voxel_map), ky = (short)(int)(py / o.size_voxel_map), kz = (short)(int)(pz / o.size_voxel_map);
  const int nb = o.nb_voxels_visited, side = 2 * nb + 1;
  const int nvox = min(side * side * side, 200);   // max_iterations = 200 (:1100-1112)
  // ---- voxel lookup: lane owns the voxels lane, lane + 32, ... of the visit order (at most 7 for 200 voxels)
  constexpr int kSlots = 7;
  const double* vp[kSlots]; int vc[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; s++) {
    vp[s] = nullptr; vc[s] = 0;
    const int v = lane + 32 * s;
    if (v < nvox) {
      const int ix = v / (side * side), iy = (v / side) % side, iz = v % side;
      const int X = kx - nb + ix, Y = ky - nb + iy, Z = kz - nb + iz;
      if (X >= -32768 && X <= 32767 && Y >= -32768 && Y <= 32767 && Z >= -32768 && Z <= 32767) {
        const unsigned long long key = lio_key(X, Y, Z);
        int lo = 0, hi = a.n_voxels;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.keys[mid] < key) lo = mid + 1; else hi = mid; }
        if (lo < a.n_voxels && a.keys[lo] == key) {
          const int vi = a.vox[lo], c = a.n_points[vi];
          if (c >= o.threshold_voxel_capacity) { vp[s] = a.points + (size_t)vi * a.M * 3; vc[s] = c; }
        }
      }
    }
  }
  // ---- the max_number_neighbors nearest candidates in ascending (distance, visit index) order
  const int want = min(o.max_number_neighbors, kLioMaxNb);
  double last_d = -1.0; int last_v = -1;
  int found = 0;
  for (int r = 0; r < want; r++) {
    double best_d = INFINITY; int best_v = 0x7fffffff; double bx = 0, by = 0, bz = 0;
#pragma unroll
    for (int s = 0; s < kSlots; s++) {
      for (int i = 0; i < vc[s]; i++) {
        const double qx = vp[s][3 * i], qy = vp[s][3 * i + 1], qz = vp[s][3 * i + 2];
        const double dx = qx - px, dy = qy - py, dz = qz - pz;
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        const int v = (lane + 32 * s) * a.M + i;
        const bool after = d > last_d || (d == last_d && v > last_v);
        if (after && (d < best_d || (d == best_d && v < best_v))) { best_d = d; best_v = v; bx = qx; by = qy; bz = qz; }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, best_d, off); const int ov = __shfl_xor_sync(0xffffffffu, best_v, off);
      const double ox = __shfl_xor_sync(0xffffffffu, bx, off), oy = __shfl_xor_sync(0xffffffffu, by, off), oz = __shfl_xor_sync(0xffffffffu, bz, off);
      if (od < best_d || (od == best_d && ov < best_v)) { best_d = od; best_v = ov; bx = ox; by = oy; bz = oz; }
    }
    if (best_v == 0x7fffffff) break;   // fewer candidates than wanted
    if (lane == 0) { s_nb[wid][r][0] = bx; s_nb[wid][r][1] = by; s_nb[wid][r][2] = bz; }
    last_d = best_d; last_v = best_v; found = r + 1;
  }
  __syncwarp();
  if (a.neighbors) for (int i = lane; i < found * 3; i += 32) a.neighbors[((size_t)k * o.max_number_neighbors) * 3 + i] = s_nb[wid][i / 3][i % 3];
  if (lane != 0) return;
  if (a.n_neighbors) a.n_neighbors[k] = found;
  a.cnt[k] = 0;
  if (found < o.min_number_neighbors) return;
  // ---- computeNeighborhoodDistribution (:887-927)
  const double (*N)[3] = s_nb[wid];
  double bc[3] = {0, 0, 0};
  for (int i = 0; i < found; i++) { bc[0] += N[i][0]; bc[1] += N[i][1]; bc[2] += N[i][2]; }
  for (int c = 0; c < 3; c++) bc[c] /= (double)found;
  double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < found; i++) {
    const double d[3] = {N[i][0] - bc[0], N[i][1] - bc[1], N[i][2] - bc[2]};
    for (int r = 0; r < 3; r++) for (int c = r; c < 3; c++) C[r * 3 + c] += d[r] * d[c];
  }
  C[3] = C[1]; C[6] = C[2]; C[7] = C[5];
  double ev[3], V[9];
  lio_eigen3(C, ev, V);
  double n[3] = {V[0], V[3], V[6]};
  { const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]); n[0] /= nn; n[1] /= nn; n[2] /= nn; }
  const double sigma_1 = sqrt(fabs(ev[2])), sigma_2 = sqrt(fabs(ev[1])), sigma_3 = sqrt(fabs(ev[0]));
  const double a2D = (sigma_2 - sigma_3) / sigma_1;
  if (a2D != a2D) { a.cnt[k] = -1; return; }   // the reference throws std::runtime_error("error")
  const double planarity_w = pow(a2D, o.power_planarity);
  const double* rp = kp.raw_point;
  const double loc[3] = {o.R_IL[0] * rp[0] + o.R_IL[1] * rp[1] + o.R_IL[2] * rp[2] + o.t_IL[0], o.R_IL[3] * rp[0] + o.R_IL[4] * rp[1] + o.R_IL[5] * rp[2] + o.t_IL[1],
                         o.R_IL[6] * rp[0] + o.R_IL[7] * rp[1] + o.R_IL[8] * rp[2] + o.t_IL[2]};
  if (n[0] * (o.translation_begin[0] - loc[0]) + n[1] * (o.translation_begin[1] - loc[1]) + n[2] * (o.translation_begin[2] - loc[2]) < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
  double lw = fabs(o.weight_alpha), ln = fabs(o.weight_neighborhood);
  { const double sum = lw + ln; lw /= sum; ln /= sum; }
  const double d0x = N[0][0] - px, d0y = N[0][1] - py, d0z = N[0][2] - pz;
  const double weight = lw * planarity_w + ln * exp(-sqrt(d0x * d0x + d0y * d0y + d0z * d0z) / (o.max_dist_to_plane_icp * o.min_number_neighbors));
  int c = 0;
  for (int i = 0; i < o.num_closest_neighbors && i < found && i < kLioMaxClosest; i++) {
    const double dist = fabs((px - N[i][0]) * n[0] + (py - N[i][1]) * n[1] + (pz - N[i][2]) * n[2]);
    if (dist >= o.max_dist_to_plane_icp) continue;
    double nv[3] = {n[0], n[1], n[2]};
    { const double nn = sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]); nv[0] /= nn; nv[1] /= nn; nv[2] /= nn; }
    LioRec& R = a.rec[(size_t)k * kLioMaxClosest + c];
    R.normal[0] = nv[0]; R.normal[1] = nv[1]; R.normal[2] = nv[2];
    R.offset = -(nv[0] * N[i][0] + nv[1] * N[i][1] + nv[2] * N[i][2]);
    R.weight = weight;
    c++;
  }
  a.cnt[k] = c;
}

// ---- map maintenance (addPointToMap, :1167-1213)
__global__ void k_lio_point_keys(int n, const double* __restrict__ pts, double size, unsigned long long* __restrict__ keys, int32_t* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int kx = (short)(int)(pts[3 * i] / size), ky = (short)(int)(pts[3 * i + 1] / size), kz = (short)(int)(pts[3 * i + 2] / size);
  keys[i] = lio_key(kx, ky, kz); idx[i] = i;
}

// sorted scan positions; the thread at the head of a run of equal keys inserts the run's points in scan order
__global__ void k_lio_insert(int n, const unsigned long long* __restrict__ skeys, const int32_t* __restrict__ sidx, const double* __restrict__ pts, int n_voxels,
                             const unsigned long long* __restrict__ tkeys, const int32_t* __restrict__ tvox, int max_voxels, int M, double size, double min_dist,
                             int min_num_points, int32_t* __restrict__ n_points, double* __restrict__ vpoints, unsigned long long* __restrict__ new_keys,
                             int32_t* __restrict__ n_new /* [0]: new voxels, [1]: overflow flag */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = skeys[i];
  if (i > 0 && skeys[i - 1] == key) return;   // not a run head
  int slot = -1;
  { int lo = 0, hi = n_voxels;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (tkeys[mid] < key) lo = mid + 1; else hi = mid; }
    if (lo < n_voxels && tkeys[lo] == key) slot = tvox[lo]; }
  int count = slot >= 0 ? n_points[slot] : 0;
  for (int j = i; j < n && skeys[j] == key; j++) {
    const double* p = pts + 3 * (size_t)sidx[j];
    if (slot < 0) {
      if (min_num_points > 0) return;          // voxel missing and new voxels are not created in this mode: the whole run is dropped
      const int k = atomicAdd(&n_new[0], 1);
      if (n_voxels + k >= max_voxels) { n_new[1] = 1; return; }
      slot = n_voxels + k; new_keys[k] = key; count = 0;
      double* q = vpoints + ((size_t)slot * M) * 3;
      q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; count = 1;   // voxelBlock block(max); block.AddPoint(point)
      continue;
    }
    if (count >= M) break;                      // IsFull: nothing else of this run can enter
    double dmin = 10.0 * size * size;
    const double* q = vpoints + ((size_t)slot * M) * 3;
    for (int c = 0; c < count; c++) {
      const double dx = q[3 * c] - p[0], dy = q[3 * c + 1] - p[1], dz = q[3 * c + 2] - p[2];
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (d2 < dmin) dmin = d2;
    }
    if (dmin > min_dist * min_dist && (min_num_points <= 0 || count >= min_num_points)) {
      double* w = vpoints + ((size_t)slot * M + count) * 3;
      w[0] = p[0]; w[1] = p[1]; w[2] = p[2]; count++;
    }
  }
  if (slot >= 0) n_points[slot] = count;
}

__global__ void k_lio_append_table(int n_old, int n_new, const unsigned long long* __restrict__ new_keys, unsigned long long* __restrict__ tkeys, int32_t* __restrict__ tvox) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_new) { tkeys[n_old + i] = new_keys[i]; tvox[n_old + i] = n_old + i; }
}

}  // namespace gf2

using namespace gf2;

struct gf2_lio {
  gf2_lio_cfg cfg;
  cudaStream_t stream;
  unsigned long long* d_keys; int32_t *d_vox, *d_npts; double* d_points;
  gf2_lio_keypoint* d_kp; int32_t* d_cnt; LioRec* d_rec; double* d_nb; int32_t* d_nnb;
  int n_voxels = 0;
  // map maintenance scratch (allocated by the first gf2_lio_add_points)
  int scan_cap = 0; double* d_scan = nullptr; unsigned long long *d_pk = nullptr, *d_pk2 = nullptr, *d_newk = nullptr, *d_tk2 = nullptr;
  int32_t *d_pi = nullptr, *d_pi2 = nullptr, *d_tv2 = nullptr, *d_nnew = nullptr; void* d_tmp = nullptr; size_t tmp_bytes = 0;
  std::vector<void*> allocs;
  std::vector<int32_t> h_cnt; std::vector<LioRec> h_rec;
  cudaEvent_t ev[3];
  double timing[8];
};

#define GF2L_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return gf2::fail(GF2_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } while (0)

extern "C" {

int gf2_lio_create(const gf2_lio_cfg* cfg, gf2_lio** out) {
  if (!cfg || !out) return gf2::fail(GF2_ERR_INVALID, "null argument");
  if (cfg->max_voxels < 1 || cfg->max_points_per_voxel < 1 || cfg->max_keypoints < 1) return gf2::fail(GF2_ERR_INVALID, "bad LIO capacities");
  int ndev = 0; if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device) { cudaGetLastError(); return gf2::fail(GF2_ERR_CUDA, "CUDA device %d not available (no CPU fallback exists)", cfg->device); }
  GF2L_CUDA(cudaSetDevice(cfg->device));
  gf2_lio* h = new gf2_lio();
  h->cfg = *cfg;
  cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  auto alloc = [&](void** p, size_t bytes) { if (cudaMalloc(p, bytes) != cudaSuccess) return false; h->allocs.push_back(*p); return true; };
  const size_t V = cfg->max_voxels, K = cfg->max_keypoints;
  const bool ok = alloc((void**)&h->d_keys, 8 * V) && alloc((void**)&h->d_vox, 4 * V) && alloc((void**)&h->d_npts, 4 * V) && alloc((void**)&h->d_points, sizeof(double) * 3 * V * cfg->max_points_per_voxel) &&
                  alloc((void**)&h->d_kp, sizeof(gf2_lio_keypoint) * K) && alloc((void**)&h->d_cnt, 4 * K) && alloc((void**)&h->d_rec, sizeof(LioRec) * K * kLioMaxClosest) &&
                  alloc((void**)&h->d_nb, sizeof(double) * 3 * K * kLioMaxNb) && alloc((void**)&h->d_nnb, 4 * K);
  if (!ok) { gf2_lio_destroy(h); return gf2::fail(GF2_ERR_CUDA, "LIO allocation failed"); }
  for (auto& e : h->ev) cudaEventCreate(&e);
  h->h_cnt.resize(K); h->h_rec.resize(K * kLioMaxClosest);
  memset(h->timing, 0, sizeof(h->timing));
  *out = h;
  return GF2_OK;
}

void gf2_lio_destroy(gf2_lio* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs) cudaFree(p);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  cudaStreamDestroy(h->stream);
  delete h;
}

int gf2_lio_set_map(gf2_lio* h, int n_voxels, const int16_t* keys, const int32_t* n_points, const double* points) {
  if (!h || (n_voxels > 0 && (!keys || !n_points || !points))) return gf2::fail(GF2_ERR_INVALID, "null argument");
  if (n_voxels < 0 || n_voxels > h->cfg.max_voxels) return gf2::fail(GF2_ERR_INVALID, "n_voxels %d outside [0, %d]", n_voxels, h->cfg.max_voxels);
  for (int v = 0; v < n_voxels; v++) if (n_points[v] < 0 || n_points[v] > h->cfg.max_points_per_voxel) return gf2::fail(GF2_ERR_INVALID, "voxel %d holds %d points (capacity %d)", v, n_points[v], h->cfg.max_points_per_voxel);
  // sorted key table for the device-side binary search (the hash order of the reference's robin_map carries no meaning)
  std::vector<std::pair<unsigned long long, int32_t>> kv(n_voxels);
  for (int v = 0; v < n_voxels; v++) kv[v] = {lio_key(keys[3 * v], keys[3 * v + 1], keys[3 * v + 2]), v};
  std::sort(kv.begin(), kv.end());
  for (int v = 1; v < n_voxels; v++) if (kv[v].first == kv[v - 1].first) return gf2::fail(GF2_ERR_INVALID, "voxel key (%d, %d, %d) appears twice", keys[3 * kv[v].second], keys[3 * kv[v].second + 1], keys[3 * kv[v].second + 2]);
  std::vector<unsigned long long> sk(n_voxels); std::vector<int32_t> sv(n_voxels);
  for (int v = 0; v < n_voxels; v++) { sk[v] = kv[v].first; sv[v] = kv[v].second; }
  cudaSetDevice(h->cfg.device);
  if (n_voxels > 0) {
    GF2L_CUDA(cudaMemcpyAsync(h->d_keys, sk.data(), 8 * (size_t)n_voxels, cudaMemcpyHostToDevice, h->stream));
    GF2L_CUDA(cudaMemcpyAsync(h->d_vox, sv.data(), 4 * (size_t)n_voxels, cudaMemcpyHostToDevice, h->stream));
    GF2L_CUDA(cudaMemcpyAsync(h->d_npts, n_points, 4 * (size_t)n_voxels, cudaMemcpyHostToDevice, h->stream));
    GF2L_CUDA(cudaMemcpyAsync(h->d_points, points, sizeof(double) * 3 * (size_t)n_voxels * h->cfg.max_points_per_voxel, cudaMemcpyHostToDevice, h->stream));
  }
  GF2L_CUDA(cudaStreamSynchronize(h->stream));   // sk / sv are locals
  h->n_voxels = n_voxels;
  return GF2_OK;
}

int gf2_lio_add_points(gf2_lio* h, int n, const double* points, double size_voxel_map, double min_distance_points, int min_num_points) {
  if (!h || (n > 0 && !points)) return gf2::fail(GF2_ERR_INVALID, "null argument");
  if (n < 0 || !(size_voxel_map > 0.0) || min_distance_points < 0.0) return gf2::fail(GF2_ERR_INVALID, "bad argument");
  if (n == 0) return GF2_OK;
  cudaSetDevice(h->cfg.device);
  const int V = h->cfg.max_voxels;
  if (n > h->scan_cap) {   // scratch grows to the largest scan seen
    auto alloc = [&](void** p, size_t bytes) { if (cudaMalloc(p, bytes) != cudaSuccess) return false; h->allocs.push_back(*p); return true; };
    const size_t cap = (size_t)n + n / 4 + 1024;
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int)cap, 0, 48, h->stream);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, V, 0, 48, h->stream);
    const size_t tb = t1 > t2 ? t1 : t2;
    bool ok = alloc((void**)&h->d_scan, sizeof(double) * 3 * cap) && alloc((void**)&h->d_pk, 8 * cap) && alloc((void**)&h->d_pk2, 8 * cap) && alloc((void**)&h->d_pi, 4 * cap) && alloc((void**)&h->d_pi2, 4 * cap) &&
              alloc(&h->d_tmp, tb);
    if (ok && !h->d_newk) ok = alloc((void**)&h->d_newk, 8 * (size_t)V) && alloc((void**)&h->d_tk2, 8 * (size_t)V) && alloc((void**)&h->d_tv2, 4 * (size_t)V) && alloc((void**)&h->d_nnew, 8);
    if (!ok) { h->scan_cap = 0; return gf2::fail(GF2_ERR_CUDA, "LIO map scratch allocation failed"); }
    h->scan_cap = (int)cap; h->tmp_bytes = tb;
  }
  cudaEventRecord(h->ev[0], h->stream);
  GF2L_CUDA(cudaMemcpyAsync(h->d_scan, points, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  GF2L_CUDA(cudaMemsetAsync(h->d_nnew, 0, 8, h->stream));
  cudaEventRecord(h->ev[1], h->stream);
  k_lio_point_keys<<<(n + 255) / 256, 256, 0, h->stream>>>(n, h->d_scan, size_voxel_map, h->d_pk, h->d_pi);
  size_t tb = h->tmp_bytes;
  GF2L_CUDA(cub::DeviceRadixSort::SortPairs(h->d_tmp, tb, h->d_pk, h->d_pk2, h->d_pi, h->d_pi2, n, 0, 48, h->stream));   // LSD radix sort: stable, scan order kept inside a voxel
  k_lio_insert<<<(n + 127) / 128, 128, 0, h->stream>>>(n, h->d_pk2, h->d_pi2, h->d_scan, h->n_voxels, h->d_keys, h->d_vox, V, h->cfg.max_points_per_voxel, size_voxel_map,
                                                       min_distance_points, min_num_points, h->d_npts, h->d_points, h->d_newk, h->d_nnew);
  GF2L_CUDA(cudaGetLastError());
  int32_t nn[2] = {0, 0};
  GF2L_CUDA(cudaMemcpyAsync(nn, h->d_nnew, 8, cudaMemcpyDeviceToHost, h->stream));
  GF2L_CUDA(cudaStreamSynchronize(h->stream));
  if (nn[1] || h->n_voxels + nn[0] > V) return gf2::fail(GF2_ERR_INVALID, "voxel map capacity %d exceeded (%d voxels + %d new)", V, h->n_voxels, nn[0]);
  if (nn[0] > 0) {   // the new voxels join the sorted key table: append, then sort the table again (48-bit keys)
    k_lio_append_table<<<(nn[0] + 255) / 256, 256, 0, h->stream>>>(h->n_voxels, nn[0], h->d_newk, h->d_keys, h->d_vox);
    const int total = h->n_voxels + nn[0];
    tb = h->tmp_bytes;
    GF2L_CUDA(cub::DeviceRadixSort::SortPairs(h->d_tmp, tb, h->d_keys, h->d_tk2, h->d_vox, h->d_tv2, total, 0, 48, h->stream));
    GF2L_CUDA(cudaMemcpyAsync(h->d_keys, h->d_tk2, 8 * (size_t)total, cudaMemcpyDeviceToDevice, h->stream));
    GF2L_CUDA(cudaMemcpyAsync(h->d_vox, h->d_tv2, 4 * (size_t)total, cudaMemcpyDeviceToDevice, h->stream));
    h->n_voxels = total;
  }
  cudaEventRecord(h->ev[2], h->stream);
  GF2L_CUDA(cudaStreamSynchronize(h->stream));
  float ms; memset(h->timing, 0, sizeof(h->timing));
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[2]); h->timing[0] = ms;
  cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); h->timing[1] = ms;
  h->timing[2] = h->n_voxels; h->timing[3] = n; h->timing[4] = nn[0];
  return GF2_OK;
}

int gf2_lio_map_size(gf2_lio* h, int32_t* n_voxels) {
  if (!h || !n_voxels) return gf2::fail(GF2_ERR_INVALID, "null argument");
  *n_voxels = h->n_voxels;
  return GF2_OK;
}

int gf2_lio_get_map(gf2_lio* h, int16_t* keys, int32_t* n_points, double* points) {
  if (!h || !keys || !n_points || !points) return gf2::fail(GF2_ERR_INVALID, "null argument");
  const int n = h->n_voxels, M = h->cfg.max_points_per_voxel;
  if (n == 0) return GF2_OK;
  cudaSetDevice(h->cfg.device);
  std::vector<unsigned long long> tk(n); std::vector<int32_t> tv(n), np(n); std::vector<double> pts((size_t)n * M * 3);
  GF2L_CUDA(cudaMemcpyAsync(tk.data(), h->d_keys, 8 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  GF2L_CUDA(cudaMemcpyAsync(tv.data(), h->d_vox, 4 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  GF2L_CUDA(cudaMemcpyAsync(np.data(), h->d_npts, 4 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  GF2L_CUDA(cudaMemcpyAsync(pts.data(), h->d_points, sizeof(double) * 3 * (size_t)n * M, cudaMemcpyDeviceToHost, h->stream));
  GF2L_CUDA(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < n; i++) {   // ascending key order
    const int s = tv[i];
    keys[3 * i] = (int16_t)((int)((tk[i] >> 32) & 0xffff) - 32768); keys[3 * i + 1] = (int16_t)((int)((tk[i] >> 16) & 0xffff) - 32768); keys[3 * i + 2] = (int16_t)((int)(tk[i] & 0xffff) - 32768);
    n_points[i] = np[s];
    memcpy(points + (size_t)i * M * 3, pts.data() + (size_t)s * M * 3, sizeof(double) * 3 * M);
  }
  return GF2_OK;
}

int gf2_lio_build_factors(gf2_lio* h, int n_keypoints, const gf2_lio_keypoint* keypoints, const gf2_lio_opts* o, gf2_plane* out_factors, double* out_alpha,
                          int32_t* n_out, double* out_neighbors, int32_t* out_n_neighbors) {
  if (!h || !o || !n_out || (n_keypoints > 0 && (!keypoints || !out_factors || !out_alpha))) return gf2::fail(GF2_ERR_INVALID, "null argument");
  if (n_keypoints < 0 || n_keypoints > h->cfg.max_keypoints) return gf2::fail(GF2_ERR_INVALID, "n_keypoints %d outside [0, %d]", n_keypoints, h->cfg.max_keypoints);
  if (o->max_number_neighbors < 1 || o->max_number_neighbors > kLioMaxNb) return gf2::fail(GF2_ERR_INVALID, "max_number_neighbors %d outside [1, %d]", o->max_number_neighbors, kLioMaxNb);
  if (o->num_closest_neighbors < 0 || o->num_closest_neighbors > kLioMaxClosest) return gf2::fail(GF2_ERR_INVALID, "num_closest_neighbors %d outside [0, %d]", o->num_closest_neighbors, kLioMaxClosest);
  if (o->nb_voxels_visited < 0 || !(o->size_voxel_map > 0.0) || o->max_num_residuals < 0) return gf2::fail(GF2_ERR_INVALID, "bad LIO options");
  if (o->icp_model != GF2_ICP_CT_POINT_TO_PLANE && o->icp_model != GF2_ICP_POINT_TO_PLANE) return gf2::fail(GF2_ERR_INVALID, "unknown icp_model %d", o->icp_model);
  *n_out = 0;
  if (n_keypoints == 0) return GF2_OK;
  cudaSetDevice(h->cfg.device);
  cudaEventRecord(h->ev[0], h->stream);
  GF2L_CUDA(cudaMemcpyAsync(h->d_kp, keypoints, sizeof(gf2_lio_keypoint) * n_keypoints, cudaMemcpyHostToDevice, h->stream));
  LioArgs a;
  a.n_keypoints = n_keypoints; a.n_voxels = h->n_voxels; a.M = h->cfg.max_points_per_voxel;
  a.keys = h->d_keys; a.vox = h->d_vox; a.n_points = h->d_npts; a.points = h->d_points; a.kp = h->d_kp; a.o = *o;
  a.cnt = h->d_cnt; a.rec = h->d_rec; a.neighbors = out_neighbors ? h->d_nb : nullptr; a.n_neighbors = out_n_neighbors ? h->d_nnb : nullptr;
  cudaEventRecord(h->ev[1], h->stream);
  k_lio_factors<<<(n_keypoints + kLioWarps - 1) / kLioWarps, 32 * kLioWarps, 0, h->stream>>>(a);
  GF2L_CUDA(cudaGetLastError());
  cudaEventRecord(h->ev[2], h->stream);
  GF2L_CUDA(cudaMemcpyAsync(h->h_cnt.data(), h->d_cnt, 4 * (size_t)n_keypoints, cudaMemcpyDeviceToHost, h->stream));
  GF2L_CUDA(cudaMemcpyAsync(h->h_rec.data(), h->d_rec, sizeof(LioRec) * (size_t)n_keypoints * kLioMaxClosest, cudaMemcpyDeviceToHost, h->stream));
  if (out_neighbors) GF2L_CUDA(cudaMemcpyAsync(out_neighbors, h->d_nb, sizeof(double) * 3 * (size_t)n_keypoints * o->max_number_neighbors, cudaMemcpyDeviceToHost, h->stream));
  if (out_n_neighbors) GF2L_CUDA(cudaMemcpyAsync(out_n_neighbors, h->d_nnb, 4 * (size_t)n_keypoints, cudaMemcpyDeviceToHost, h->stream));
  GF2L_CUDA(cudaStreamSynchronize(h->stream));
  // the reference's sequential loop (:967-1062): residuals in keypoint order, hard stop at max_num_residuals
  // p_state->rotation.inverse() for POINT_TO_PLANE (:1040-1042)
  const double qx = o->rotation[0], qy = o->rotation[1], qz = o->rotation[2], qw = o->rotation[3];
  const double R[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw), 2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw),
                       2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)};
  int num = 0;
  bool full = false;   // the reference tests the cap AFTER pushing a residual (:1058-1061): at least one gets through
  for (int k = 0; k < n_keypoints && !full; k++) {
    if (h->h_cnt[k] < 0) return gf2::fail(GF2_ERR_INVALID, "keypoint %d: planarity a2D is NaN (the reference throws here, lidarodom.cpp:921-924)", k);
    for (int c = 0; c < h->h_cnt[k] && !full; c++) {
      const LioRec& r = h->h_rec[(size_t)k * kLioMaxClosest + c];
      gf2_plane& f = out_factors[num];
      memset(&f, 0, sizeof(f));
      f.normal[0] = r.normal[0]; f.normal[1] = r.normal[1]; f.normal[2] = r.normal[2]; f.offset = r.offset; f.weight = r.weight; f.frame = k; f.ct = (o->icp_model == GF2_ICP_CT_POINT_TO_PLANE) ? 1 : 0;
      const gf2_lio_keypoint& kp = keypoints[k];
      if (o->icp_model == GF2_ICP_CT_POINT_TO_PLANE) { f.p_body[0] = kp.raw_point[0]; f.p_body[1] = kp.raw_point[1]; f.p_body[2] = kp.raw_point[2]; }
      else {
        const double* p = kp.point; const double* t = o->translation;
        const double d[3] = {p[0], p[1], p[2]};
        const double a3[3] = {R[0] * d[0] + R[3] * d[1] + R[6] * d[2], R[1] * d[0] + R[4] * d[1] + R[7] * d[2], R[2] * d[0] + R[5] * d[1] + R[8] * d[2]};
        const double b3[3] = {R[0] * t[0] + R[3] * t[1] + R[6] * t[2], R[1] * t[0] + R[4] * t[1] + R[7] * t[2], R[2] * t[0] + R[5] * t[1] + R[8] * t[2]};
        f.p_body[0] = a3[0] - b3[0]; f.p_body[1] = a3[1] - b3[1]; f.p_body[2] = a3[2] - b3[2];
      }
      out_alpha[num] = kp.alpha_time;
      num++;
      if (num >= o->max_num_residuals) full = true;
    }
  }
  *n_out = num;
  float ms; memset(h->timing, 0, sizeof(h->timing));
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[2]); h->timing[0] = ms;
  cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); h->timing[1] = ms;
  h->timing[2] = h->n_voxels; h->timing[3] = n_keypoints;
  return GF2_OK;
}

int gf2_lio_last_timing(gf2_lio* h, double out[8]) {
  if (!h || !out) return gf2::fail(GF2_ERR_INVALID, "null argument");
  memcpy(out, h->timing, sizeof(h->timing));
  return GF2_OK;
}

}  // extern "C"
