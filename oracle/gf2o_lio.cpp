// TEST INFRASTRUCTURE — CPU restatement of the LIO factor construction of the reference, used only by tests/ and bench.py's CPU
// baseline leg: lidarodom::searchNeighbors (LIO/liw/lio/lidarodom.cpp:1087-1165), computeNeighborhoodDistribution (:887-927) and
// addSurfCostFactor (:929-1071). Eigen / tsl::robin_map are un-vendored dependencies: the voxel map is a std::unordered_map over the
// same key (cloudMap.hpp:34-58), Eigen::SelfAdjointEigenSolver<Matrix3d> is replaced by the cyclic Jacobi solver of gf2o_linalg.h
// (eigenvalues ascending; the eigenvector of a simple eigenvalue is unique up to sign and the sign is fixed by :939-940).
// Parity unpinned (the reference holds no known-answer test for this function); pinned instead against an independent numpy
// restatement (brute-force neighbour search + numpy.linalg.eigh) in tests/test_oracle_lio.py.
#include <chrono>
#include <cmath>
#include <cstring>
#include <queue>
#include <tuple>
#include <unordered_map>
#include <vector>
#include "../include/gf2_abi.h"
#include "gf2o_linalg.h"

namespace {
struct P3 { double x, y, z; };
struct Vox { short x, y, z; bool operator==(const Vox& o) const { return x == o.x && y == o.y && z == o.z; } };
struct VoxHash { size_t operator()(const Vox& v) const { return v.x * (size_t)73856093 + v.y * (size_t)19349669 + v.z * (size_t)83492791; } };  // cloudMap.hpp:88-99
using Map = std::unordered_map<Vox, std::vector<P3>, VoxHash>;
using Item = std::tuple<double, P3>;
struct Cmp { bool operator()(const Item& l, const Item& r) const { return std::get<0>(l) < std::get<0>(r); } };

// lidarodom::searchNeighbors, :1087-1165
std::vector<P3> searchNeighbors(const Map& map, const P3& point, int nb_voxels_visited, double size_voxel_map, int max_num_neighbors, int threshold_voxel_capacity) {
  short kx = static_cast<short>(point.x / size_voxel_map), ky = static_cast<short>(point.y / size_voxel_map), kz = static_cast<short>(point.z / size_voxel_map);
  std::priority_queue<Item, std::vector<Item>, Cmp> pq;
  int max_iterations = 200, iteration_count = 0;
  bool stop = false;
  for (short kxx = kx - nb_voxels_visited; kxx < kx + nb_voxels_visited + 1 && !stop; ++kxx)
    for (short kyy = ky - nb_voxels_visited; kyy < ky + nb_voxels_visited + 1 && !stop; ++kyy)
      for (short kzz = kz - nb_voxels_visited; kzz < kz + nb_voxels_visited + 1; ++kzz) {
        if (++iteration_count > max_iterations) { stop = true; break; }
        auto search = map.find(Vox{kxx, kyy, kzz});
        if (search == map.end()) continue;
        const std::vector<P3>& blk = search->second;
        if ((int)blk.size() < threshold_voxel_capacity) continue;
        for (const P3& nb : blk) {
          const double dx = nb.x - point.x, dy = nb.y - point.y, dz = nb.z - point.z;
          const double distance = std::sqrt(dx * dx + dy * dy + dz * dz);
          if ((int)pq.size() == max_num_neighbors) { if (distance < std::get<0>(pq.top())) { pq.pop(); pq.emplace(distance, nb); } }
          else pq.emplace(distance, nb);
        }
      }
  const size_t size = pq.size();
  std::vector<P3> out(size);
  for (size_t i = 0; i < size; ++i) { out[size - 1 - i] = std::get<1>(pq.top()); pq.pop(); }
  return out;
}
}  // namespace

extern "C" {

// returns the number of residuals; arrays as in gf2_lio_build_factors (out_neighbors / out_n_neighbors nullable)
int gf2o_lio_build_factors(int n_voxels, const int16_t* keys, const int32_t* n_points, const double* points, int max_points_per_voxel, int n_keypoints,
                           const gf2_lio_keypoint* keypoints, const gf2_lio_opts* o, gf2_plane* out_factors, double* out_alpha, double* out_neighbors,
                           int32_t* out_n_neighbors, double* loop_seconds /* nullable: time of the keypoint loop alone (map build excluded) */) {
  Map map;
  for (int v = 0; v < n_voxels; v++) {
    std::vector<P3>& blk = map[Vox{keys[3 * v], keys[3 * v + 1], keys[3 * v + 2]}];
    for (int i = 0; i < n_points[v]; i++) { const double* p = points + ((size_t)v * max_points_per_voxel + i) * 3; blk.push_back({p[0], p[1], p[2]}); }
  }
  double lambda_weight = std::fabs(o->weight_alpha), lambda_neighborhood = std::fabs(o->weight_neighborhood);   // :944-948
  const double sum = lambda_weight + lambda_neighborhood;
  lambda_weight /= sum; lambda_neighborhood /= sum;
  const double kMaxPointToPlane = o->max_dist_to_plane_icp;
  // p_state->rotation.inverse() for POINT_TO_PLANE (:1040-1042)
  const double qx = o->rotation[0], qy = o->rotation[1], qz = o->rotation[2], qw = o->rotation[3];
  const double R[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw), 2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw),
                       2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)};
  int num_residuals = 0;
  const auto t_loop0 = std::chrono::steady_clock::now();
  struct Stop { std::chrono::steady_clock::time_point t0; double* out; ~Stop() { if (out) *out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); } } stop{t_loop0, loop_seconds};
  for (int k = 0; k < n_keypoints; ++k) {
    const gf2_lio_keypoint& kp = keypoints[k];
    const P3 pt = {kp.point[0], kp.point[1], kp.point[2]};
    std::vector<P3> nbs = searchNeighbors(map, pt, o->nb_voxels_visited, o->size_voxel_map, o->max_number_neighbors, o->threshold_voxel_capacity);
    if (out_n_neighbors) out_n_neighbors[k] = (int32_t)nbs.size();
    if (out_neighbors) for (size_t i = 0; i < nbs.size(); i++) { double* q = out_neighbors + ((size_t)k * o->max_number_neighbors + i) * 3; q[0] = nbs[i].x; q[1] = nbs[i].y; q[2] = nbs[i].z; }
    if ((int)nbs.size() < o->min_number_neighbors) continue;
    // computeNeighborhoodDistribution, :887-927
    double bc[3] = {0, 0, 0};
    for (const P3& p : nbs) { bc[0] += p.x; bc[1] += p.y; bc[2] += p.z; }
    for (double& b : bc) b /= (double)nbs.size();
    double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (const P3& p : nbs) {
      const double d[3] = {p.x - bc[0], p.y - bc[1], p.z - bc[2]};
      for (int a = 0; a < 3; ++a) for (int b = a; b < 3; ++b) C[a * 3 + b] += d[a] * d[b];
    }
    C[3] = C[1]; C[6] = C[2]; C[7] = C[5];
    double ev[3], V[9];
    gf2o::symEigen(3, C, ev, V);
    double n[3] = {V[0], V[3], V[6]};   // eigenvectors().col(0).normalized()
    { const double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]); for (double& c : n) c /= nn; }
    const double sigma_1 = std::sqrt(std::fabs(ev[2])), sigma_2 = std::sqrt(std::fabs(ev[1])), sigma_3 = std::sqrt(std::fabs(ev[0]));
    const double a2D = (sigma_2 - sigma_3) / sigma_1;
    if (a2D != a2D) return -1;   // the reference throws std::runtime_error("error")
    const double planarity_w = std::pow(a2D, o->power_planarity);   // :938
    const double loc[3] = {o->R_IL[0] * kp.raw_point[0] + o->R_IL[1] * kp.raw_point[1] + o->R_IL[2] * kp.raw_point[2] + o->t_IL[0],
                           o->R_IL[3] * kp.raw_point[0] + o->R_IL[4] * kp.raw_point[1] + o->R_IL[5] * kp.raw_point[2] + o->t_IL[1],
                           o->R_IL[6] * kp.raw_point[0] + o->R_IL[7] * kp.raw_point[1] + o->R_IL[8] * kp.raw_point[2] + o->t_IL[2]};
    if (n[0] * (o->translation_begin[0] - loc[0]) + n[1] * (o->translation_begin[1] - loc[1]) + n[2] * (o->translation_begin[2] - loc[2]) < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    const double d0[3] = {nbs[0].x - pt.x, nbs[0].y - pt.y, nbs[0].z - pt.z};
    const double weight = lambda_weight * planarity_w + lambda_neighborhood * std::exp(-std::sqrt(d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2]) / (kMaxPointToPlane * o->min_number_neighbors));
    for (int i = 0; i < o->num_closest_neighbors && (size_t)i < nbs.size(); ++i) {
      const double dist = std::fabs((pt.x - nbs[i].x) * n[0] + (pt.y - nbs[i].y) * n[1] + (pt.z - nbs[i].z) * n[2]);
      if (dist >= o->max_dist_to_plane_icp) continue;
      gf2_plane& f = out_factors[num_residuals];
      std::memset(&f, 0, sizeof(f));
      double nv[3] = {n[0], n[1], n[2]};   // neighborhood.normal.normalized() (:1012)
      { const double nn = std::sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]); for (double& c : nv) c /= nn; }
      f.normal[0] = nv[0]; f.normal[1] = nv[1]; f.normal[2] = nv[2];
      f.offset = -(nv[0] * nbs[i].x + nv[1] * nbs[i].y + nv[2] * nbs[i].z);
      f.weight = weight; f.frame = k; f.ct = (o->icp_model == GF2_ICP_CT_POINT_TO_PLANE) ? 1 : 0;
      if (o->icp_model == GF2_ICP_CT_POINT_TO_PLANE) { f.p_body[0] = kp.raw_point[0]; f.p_body[1] = kp.raw_point[1]; f.p_body[2] = kp.raw_point[2]; }
      else {  // point_end = R^-1 kp.point - R^-1 translation
        const double a[3] = {R[0] * pt.x + R[3] * pt.y + R[6] * pt.z, R[1] * pt.x + R[4] * pt.y + R[7] * pt.z, R[2] * pt.x + R[5] * pt.y + R[8] * pt.z};
        const double* t = o->translation;
        const double b[3] = {R[0] * t[0] + R[3] * t[1] + R[6] * t[2], R[1] * t[0] + R[4] * t[1] + R[7] * t[2], R[2] * t[0] + R[5] * t[1] + R[8] * t[2]};
        f.p_body[0] = a[0] - b[0]; f.p_body[1] = a[1] - b[1]; f.p_body[2] = a[2] - b[2];
      }
      out_alpha[num_residuals] = kp.alpha_time;
      ++num_residuals;
      if (num_residuals >= o->max_num_residuals) break;
    }
    if (num_residuals >= o->max_num_residuals) break;
  }
  return num_residuals;
}

}  // extern "C"
