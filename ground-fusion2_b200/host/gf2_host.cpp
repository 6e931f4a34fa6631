// gf2_host.cpp — see gf2_host.h. Host-side mirror of Estimator / FeatureManager / FeatureTracker glue over the C ABI.
#include "gf2_host.h"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>

namespace gf2host {
constexpr int kMaxWheelSamples = 64;  // per wheel preintegration interval
constexpr int kMaxImuSamples = 256;   // per preintegration interval (MARGIN_SECOND_NEW merges intervals)

// ------------------------------------------------------------------------------------------------ small math
Matrix3d mul(const Matrix3d& a, const Matrix3d& b) {
  Matrix3d r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
  return r;
}
Vector3d mul(const Matrix3d& a, const Vector3d& v) {
  return {a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z};
}
Matrix3d transpose(const Matrix3d& a) { Matrix3d r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i * 3 + j] = a.m[j * 3 + i]; return r; }
Quaterniond quatFromMatrix(const Matrix3d& R) {  // Eigen::Quaterniond(Matrix3d)
  Quaterniond q; const double* m = R.m;
  double t = m[0] + m[4] + m[8];
  if (t > 0) { t = std::sqrt(t + 1.0); q.w = 0.5 * t; t = 0.5 / t; q.x = (m[7] - m[5]) * t; q.y = (m[2] - m[6]) * t; q.z = (m[3] - m[1]) * t; }
  else {
    int i = 0; if (m[4] > m[0]) i = 1; if (m[8] > m[i * 4]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i * 4] - m[j * 4] - m[k * 4] + 1.0);
    double v[3]; v[i] = 0.5 * t; t = 0.5 / t;
    q.w = (m[k * 3 + j] - m[j * 3 + k]) * t; v[j] = (m[j * 3 + i] + m[i * 3 + j]) * t; v[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  return q;
}
Matrix3d toRotationMatrix(const Quaterniond& q) {
  Matrix3d r; const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r.m[0] = 1 - (tyy + tzz); r.m[1] = txy - twz; r.m[2] = txz + twy; r.m[3] = txy + twz; r.m[4] = 1 - (txx + tzz); r.m[5] = tyz - twx;
  r.m[6] = txz - twy; r.m[7] = tyz + twx; r.m[8] = 1 - (txx + tyy);
  return r;
}
static Quaterniond normalized(Quaterniond q) { double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z); return {q.w / n, q.x / n, q.y / n, q.z / n}; }
Vector3d R2ypr(const Matrix3d& R) {  // VE/utility/utility.h:78-93
  const double n0 = R.m[0], n1 = R.m[3], n2 = R.m[6], o0 = R.m[1], o1 = R.m[4], a0 = R.m[2], a1 = R.m[5];
  const double y = std::atan2(n1, n0);
  const double p = std::atan2(-n2, n0 * std::cos(y) + n1 * std::sin(y));
  const double r = std::atan2(a0 * std::sin(y) - a1 * std::cos(y), -o0 * std::sin(y) + o1 * std::cos(y));
  return {y / M_PI * 180.0, p / M_PI * 180.0, r / M_PI * 180.0};
}
Matrix3d ypr2R(const Vector3d& ypr) {  // VE/utility/utility.h:95-121
  const double y = ypr.x / 180.0 * M_PI, p = ypr.y / 180.0 * M_PI, r = ypr.z / 180.0 * M_PI;
  Matrix3d Rz, Ry, Rx;
  Rz.m[0] = std::cos(y); Rz.m[1] = -std::sin(y); Rz.m[3] = std::sin(y); Rz.m[4] = std::cos(y);
  Ry.m[0] = std::cos(p); Ry.m[2] = std::sin(p); Ry.m[6] = -std::sin(p); Ry.m[8] = std::cos(p);
  Rx.m[4] = std::cos(r); Rx.m[5] = -std::sin(r); Rx.m[7] = std::sin(r); Rx.m[8] = std::cos(r);
  return mul(mul(Rz, Ry), Rx);
}

// ------------------------------------------------------------------------------------------------ YAML (OpenCV FileStorage dialect)
namespace {
struct Yaml {
  std::map<std::string, std::string> scalars;
  std::map<std::string, std::vector<double>> matrices;
  // C stdio + strtod only: this library is loaded into Python processes next to other C++ runtimes, iostream/locale
  // state must not be touched
  bool load(const std::string& path) {
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    std::string cur_key; bool in_data = false; std::string data;
    auto trim = [](std::string s) { size_t a = s.find_first_not_of(" \t\r\n\""); size_t b = s.find_last_not_of(" \t\r\n\""); return a == std::string::npos ? std::string() : s.substr(a, b - a + 1); };
    auto flush = [&]() {
      if (cur_key.empty() || data.empty()) return;
      std::vector<double> v;
      const char* c = data.c_str();
      while (*c) {
        while (*c && (*c == ',' || *c == '[' || *c == ']' || *c == ' ' || *c == '\t' || *c == '\r' || *c == '\n')) c++;
        if (!*c) break;
        char* end = nullptr; const double x = strtod(c, &end);
        if (end == c) break;
        v.push_back(x); c = end;
      }
      matrices[cur_key] = v; data.clear();
    };
    char buf[4096];
    while (fgets(buf, sizeof(buf), f)) {
      std::string line(buf);
      while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
      size_t hash = line.find('#'); if (hash != std::string::npos) line = line.substr(0, hash);
      if (line.empty() || line[0] == '%' || line.compare(0, 3, "---") == 0) continue;
      if (in_data) { data += " " + line; if (line.find(']') != std::string::npos) { in_data = false; flush(); } continue; }
      size_t colon = line.find(':');
      if (colon == std::string::npos) continue;
      std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
      const bool nested = line[0] == ' ' || line[0] == '\t';
      if (!nested) {
        if (val.find("!!opencv-matrix") != std::string::npos) { cur_key = key; continue; }
        if (val.empty()) { cur_key = key; continue; }  // mapping node (distortion_parameters: ...)
        scalars[key] = val; cur_key.clear();
      } else {
        if (key == "data") { data = val; if (val.find(']') == std::string::npos) in_data = true; else flush(); }
        else if (key != "rows" && key != "cols" && key != "dt") scalars[cur_key.empty() ? key : cur_key + "." + key] = val;
      }
    }
    fclose(f);
    return true;
  }
  double num(const std::string& k) const { auto it = scalars.find(k); return it == scalars.end() ? 0.0 : std::atof(it->second.c_str()); }  // absent -> 0 (cv::FileNode)
  std::string str(const std::string& k) const { auto it = scalars.find(k); return it == scalars.end() ? std::string() : it->second; }
};
}  // namespace

bool readParameters(const std::string& config_file, Parameters& P) {  // keys and derived values of parameters.cpp:160-556
  Yaml y;
  if (!y.load(config_file)) return false;
  P.USE_IMU = (int)y.num("imu"); P.USE_WHEEL = (int)y.num("wheel");
  P.MAX_CNT = (int)y.num("max_cnt"); P.MIN_DIST = (int)y.num("min_dist"); P.F_THRESHOLD = y.num("F_threshold"); P.FLOW_BACK = (int)y.num("flow_back"); P.EQUALIZE = (int)y.num("equalize");
  P.ACC_N = y.num("acc_n"); P.ACC_W = y.num("acc_w"); P.GYR_N = y.num("gyr_n"); P.GYR_W = y.num("gyr_w"); P.G_NORM = y.num("g_norm");
  P.SOLVER_TIME = y.num("max_solver_time"); P.NUM_ITERATIONS = (int)y.num("max_num_iterations");
  P.MIN_PARALLAX = y.num("keyframe_parallax") / FOCAL_LENGTH;  // :351-352
  P.ESTIMATE_EXTRINSIC = (int)y.num("estimate_extrinsic"); P.ESTIMATE_TD = (int)y.num("estimate_td"); P.TD = y.num("td");
  P.ROW = (int)y.num("image_height"); P.COL = (int)y.num("image_width");
  P.USE_MCC = (int)y.num("use_mcc"); P.DEPTH = (int)y.num("depth");   // parameters.cpp:164,472 (absent key reads as 0, cv::FileNode semantics)
  P.ONLY_INITIAL_WITH_WHEEL = (int)y.num("only_initial_with_wheel");
  if (P.USE_WHEEL) {  // parameters.cpp:234-335
    P.VEL_N_wheel = y.num("wheel_velocity_noise_sigma"); P.GYR_N_wheel = y.num("wheel_gyro_noise_sigma");
    P.ESTIMATE_EXTRINSIC_WHEEL = (int)y.num("estimate_wheel_extrinsic"); P.EXTRINSIC_TYPE_WHEEL = (int)y.num("extrinsic_type_wheel");
    P.ESTIMATE_INTRINSIC_WHEEL = (int)y.num("estimate_wheel_intrinsic");
    if (y.scalars.count("sx")) P.SX = y.num("sx");
    if (y.scalars.count("sy")) P.SY = y.num("sy");
    if (y.scalars.count("sw")) P.SW = y.num("sw");
    auto iw = y.matrices.find("body_T_wheel");
    if (iw != y.matrices.end() && iw->second.size() == 16) {
      const std::vector<double>& T = iw->second;
      Matrix3d R; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R.m[r * 3 + c] = T[r * 4 + c];
      P.RIO = toRotationMatrix(normalized(quatFromMatrix(R)));  // QIO.normalize() (:272-275)
      P.TIO = {T[3], T[7], T[11]};
    }
  }
  P.TD_WHEEL = y.num("td_wheel"); P.ESTIMATE_TD_WHEEL = (int)y.num("estimate_td_wheel");
  auto it = y.matrices.find("body_T_cam0");
  if (it != y.matrices.end() && it->second.size() == 16) {
    const std::vector<double>& T = it->second;
    Matrix3d R; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R.m[r * 3 + c] = T[r * 4 + c];
    P.RIC = toRotationMatrix(normalized(quatFromMatrix(R)));  // re-orthonormalised through a quaternion (:387-395)
    P.TIC = {T[3], T[7], T[11]};
  }
  const std::string cam = y.str("cam0_calib");
  if (!cam.empty()) {
    const size_t slash = config_file.find_last_of('/');
    Yaml c;
    if (c.load((slash == std::string::npos ? std::string() : config_file.substr(0, slash + 1)) + cam)) {
      P.fx = c.num("projection_parameters.fx"); P.fy = c.num("projection_parameters.fy"); P.cx = c.num("projection_parameters.cx"); P.cy = c.num("projection_parameters.cy");
      P.k1 = c.num("distortion_parameters.k1"); P.k2 = c.num("distortion_parameters.k2"); P.p1 = c.num("distortion_parameters.p1"); P.p2 = c.num("distortion_parameters.p2");
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------------ FeatureManager
int FeatureManager::getFeatureCount() {
  int cnt = 0;
  for (auto& it : feature) { it.used_num = (int)it.feature_per_frame.size(); if (it.used_num >= 4) cnt++; }
  return cnt;
}
std::vector<double> FeatureManager::getDepthVector() {
  std::vector<double> dep_vec(getFeatureCount());
  int feature_index = -1;
  for (auto& it_per_id : feature) {
    it_per_id.used_num = (int)it_per_id.feature_per_frame.size();
    if (it_per_id.used_num < 4) continue;
    dep_vec[++feature_index] = 1. / it_per_id.estimated_depth;
  }
  return dep_vec;
}
void FeatureManager::setDepth(const std::vector<double>& x) {
  int feature_index = -1;
  for (auto& it_per_id : feature) {
    it_per_id.used_num = (int)it_per_id.feature_per_frame.size();
    if (it_per_id.used_num < 4) continue;
    it_per_id.estimated_depth = 1.0 / x[++feature_index];
    it_per_id.solve_flag = it_per_id.estimated_depth < 0 ? 2 : 1;
  }
}
void FeatureManager::removeFailures() {
  for (auto it = feature.begin(), it_next = feature.begin(); it != feature.end(); it = it_next) { it_next++; if (it->solve_flag == 2) feature.erase(it); }
}
void FeatureManager::clearDepth() { for (auto& it : feature) it.estimated_depth = -1; }
void FeatureManager::addFeatures(int frame_count, const std::map<int, std::vector<std::pair<int, std::vector<double>>>>& image, double td) {
  for (auto& id_pts : image) {  // std::map iteration = ascending feature id, new features appended at the list tail
    FeaturePerFrame f_per_fra(id_pts.second[0].second.data(), td);
    const int feature_id = id_pts.first;
    auto it = std::find_if(feature.begin(), feature.end(), [feature_id](const FeaturePerId& f) { return f.feature_id == feature_id; });
    if (it == feature.end()) { feature.push_back(FeaturePerId(feature_id, frame_count)); feature.back().feature_per_frame.push_back(f_per_fra); }
    else it->feature_per_frame.push_back(f_per_fra);
  }
}

// feature_manager.cpp:57-116
bool FeatureManager::addFeatureCheckParallax(int frame_count, const std::map<int, std::vector<std::pair<int, std::vector<double>>>>& image, double td) {
  double parallax_sum = 0; int parallax_num = 0;
  last_track_num = 0; last_average_parallax = 0; new_feature_num = 0; long_track_num = 0;
  for (auto& id_pts : image) {
    FeaturePerFrame f_per_fra(id_pts.second[0].second.data(), td);
    const int feature_id = id_pts.first;
    auto it = std::find_if(feature.begin(), feature.end(), [feature_id](const FeaturePerId& f) { return f.feature_id == feature_id; });
    if (it == feature.end()) { feature.push_back(FeaturePerId(feature_id, frame_count)); feature.back().feature_per_frame.push_back(f_per_fra); new_feature_num++; }
    else { it->feature_per_frame.push_back(f_per_fra); last_track_num++; if (it->feature_per_frame.size() >= 4) long_track_num++; }
  }
  if (frame_count < 2 || last_track_num < 20 || long_track_num < 40 || new_feature_num > 0.5 * last_track_num) return true;
  for (auto& it_per_id : feature)
    if (it_per_id.start_frame <= frame_count - 2 && it_per_id.start_frame + int(it_per_id.feature_per_frame.size()) - 1 >= frame_count - 1) {
      parallax_sum += compensatedParallax2(it_per_id, frame_count); parallax_num++;
    }
  if (parallax_num == 0) return true;
  last_average_parallax = parallax_sum / parallax_num * FOCAL_LENGTH;
  return parallax_sum / parallax_num >= MIN_PARALLAX;
}
// :978-1011 (the rotation compensation is commented out in the reference: p_i_comp = p_i)
double FeatureManager::compensatedParallax2(const FeaturePerId& it_per_id, int frame_count) const {
  const FeaturePerFrame& frame_i = it_per_id.feature_per_frame[frame_count - 2 - it_per_id.start_frame];
  const FeaturePerFrame& frame_j = it_per_id.feature_per_frame[frame_count - 1 - it_per_id.start_frame];
  double ans = 0;
  const double u_j = frame_j.point.x, v_j = frame_j.point.y;
  const Vector3d p_i = frame_i.point, p_i_comp = p_i;
  const double dep_i = p_i.z, u_i = p_i.x / dep_i, v_i = p_i.y / dep_i, du = u_i - u_j, dv = v_i - v_j;
  const double dep_i_comp = p_i_comp.z, u_i_comp = p_i_comp.x / dep_i_comp, v_i_comp = p_i_comp.y / dep_i_comp, du_comp = u_i_comp - u_j, dv_comp = v_i_comp - v_j;
  ans = std::max(ans, std::sqrt(std::min(du * du + dv * dv, du_comp * du_comp + dv_comp * dv_comp)));
  return ans;
}
void FeatureManager::removeOutlier(const std::set<int>& outlierIndex) {   // :801-816
  for (auto it = feature.begin(), it_next = feature.begin(); it != feature.end(); it = it_next) {
    it_next++;
    if (outlierIndex.find(it->feature_id) != outlierIndex.end()) feature.erase(it);
  }
}
void FeatureManager::removeBackShiftDepth(const Matrix3d& marg_R, const Vector3d& marg_P, const Matrix3d& new_R, const Vector3d& new_P) {   // :818-856
  for (auto it = feature.begin(), it_next = feature.begin(); it != feature.end(); it = it_next) {
    it_next++;
    if (it->start_frame != 0) it->start_frame--;
    else {
      const Vector3d uv_i = it->feature_per_frame[0].point;
      it->feature_per_frame.erase(it->feature_per_frame.begin());
      if (it->feature_per_frame.size() < 2) { feature.erase(it); continue; }
      const Vector3d pts_i = {uv_i.x * it->estimated_depth, uv_i.y * it->estimated_depth, uv_i.z * it->estimated_depth};
      const Vector3d r = mul(marg_R, pts_i), w_pts_i = {r.x + marg_P.x, r.y + marg_P.y, r.z + marg_P.z};
      const Vector3d pts_j = mul(transpose(new_R), Vector3d{w_pts_i.x - new_P.x, w_pts_i.y - new_P.y, w_pts_i.z - new_P.z});
      const double dep_j = pts_j.z;
      it->estimated_depth = dep_j > 0 ? dep_j : INIT_DEPTH;
    }
  }
}
void FeatureManager::removeBack() {   // :858-874
  for (auto it = feature.begin(), it_next = feature.begin(); it != feature.end(); it = it_next) {
    it_next++;
    if (it->start_frame != 0) it->start_frame--;
    else { it->feature_per_frame.erase(it->feature_per_frame.begin()); if (it->feature_per_frame.size() == 0) feature.erase(it); }
  }
}
void FeatureManager::removeFront(int frame_count) {   // :914-934
  for (auto it = feature.begin(), it_next = feature.begin(); it != feature.end(); it = it_next) {
    it_next++;
    if (it->start_frame == frame_count) it->start_frame--;
    else {
      const int j = WINDOW_SIZE - 1 - it->start_frame;
      if (it->endFrame() < frame_count - 1) continue;
      it->feature_per_frame.erase(it->feature_per_frame.begin() + j);
      if (it->feature_per_frame.size() == 0) feature.erase(it);
    }
  }
}
namespace {
// right singular vector of the smallest singular value of an m x 4 matrix = eigenvector of the smallest eigenvalue of A^T A
// (Eigen::JacobiSVD(...).matrixV().rightCols<1>() in the reference, up to the sign, which the caller divides out)
void smallestRightSingularVector4(const std::vector<double>& A /* m x 4 row-major */, double v[4]) {
  double M[16] = {0}, V[16];
  const size_t m = A.size() / 4;
  for (size_t r = 0; r < m; r++) for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) M[a * 4 + b] += A[r * 4 + a] * A[r * 4 + b];
  for (int i = 0; i < 16; i++) V[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0, dg = 0;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) (r == c ? dg : off) += M[r * 4 + c] * M[r * 4 + c];
    if (off <= 1e-32 * (dg + 1e-300)) break;
    for (int p = 0; p < 4; p++) for (int q = p + 1; q < 4; q++) {
      const double apq = M[p * 4 + q]; if (apq == 0.0) continue;
      const double tau = (M[q * 4 + q] - M[p * 4 + p]) / (2.0 * apq);
      const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau)), c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
      for (int k = 0; k < 4; k++) { const double a = M[k * 4 + p], b = M[k * 4 + q]; M[k * 4 + p] = c * a - s * b; M[k * 4 + q] = s * a + c * b; }
      for (int k = 0; k < 4; k++) { const double a = M[p * 4 + k], b = M[q * 4 + k]; M[p * 4 + k] = c * a - s * b; M[q * 4 + k] = s * a + c * b; }
      for (int k = 0; k < 4; k++) { const double a = V[k * 4 + p], b = V[k * 4 + q]; V[k * 4 + p] = c * a - s * b; V[k * 4 + q] = s * a + c * b; }
    }
  }
  int best = 0; for (int i = 1; i < 4; i++) if (M[i * 5] < M[best * 5]) best = i;
  for (int k = 0; k < 4; k++) v[k] = V[k * 4 + best];
}
inline Vector3d add3(const Vector3d& a, const Vector3d& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vector3d sub3(const Vector3d& a, const Vector3d& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
}  // namespace
void FeatureManager::triangulate(int, const Vector3d Ps[], const Matrix3d Rs[], const Vector3d tic[], const Matrix3d ric[]) {   // :669-724
  for (auto& it_per_id : feature) {
    if (it_per_id.estimated_depth > 0) continue;
    it_per_id.used_num = (int)it_per_id.feature_per_frame.size();
    if (it_per_id.used_num < 4) continue;
    const int imu_i = it_per_id.start_frame; int imu_j = imu_i - 1;
    std::vector<double> svd_A;
    const Vector3d t0 = add3(Ps[imu_i], mul(Rs[imu_i], tic[0])); const Matrix3d R0 = mul(Rs[imu_i], ric[0]);
    for (auto& it_per_frame : it_per_id.feature_per_frame) {
      imu_j++;
      const Vector3d t1 = add3(Ps[imu_j], mul(Rs[imu_j], tic[0])); const Matrix3d R1 = mul(Rs[imu_j], ric[0]);
      const Vector3d t = mul(transpose(R0), sub3(t1, t0)); const Matrix3d R = mul(transpose(R0), R1);
      const Matrix3d Rt = transpose(R); const Vector3d mt = mul(Rt, t);
      const double P[12] = {Rt.m[0], Rt.m[1], Rt.m[2], -mt.x, Rt.m[3], Rt.m[4], Rt.m[5], -mt.y, Rt.m[6], Rt.m[7], Rt.m[8], -mt.z};
      const Vector3d p = it_per_frame.point; const double n = std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
      const double f[3] = {p.x / n, p.y / n, p.z / n};
      for (int c = 0; c < 4; c++) svd_A.push_back(f[0] * P[8 + c] - f[2] * P[c]);
      for (int c = 0; c < 4; c++) svd_A.push_back(f[1] * P[8 + c] - f[2] * P[4 + c]);
    }
    double v[4]; smallestRightSingularVector4(svd_A, v);
    it_per_id.estimated_depth = v[2] / v[3];
    it_per_id.estimate_flag = 2;
    if (it_per_id.estimated_depth < 0.1) { it_per_id.estimated_depth = INIT_DEPTH; it_per_id.estimate_flag = 0; }
  }
}
void FeatureManager::triangulateWithDepth(int, const Vector3d Ps[], const Matrix3d Rs[], const Vector3d tic[], const Matrix3d ric[]) {   // :726-799
  for (auto& it_per_id : feature) {
    it_per_id.used_num = (int)it_per_id.feature_per_frame.size();
    if (it_per_id.used_num < 4) continue;
    if (it_per_id.estimated_depth > 0) continue;
    const int start_frame = it_per_id.start_frame;
    std::vector<double> verified_depths;
    const Vector3d tr = add3(Ps[start_frame], mul(Rs[start_frame], tic[0])); const Matrix3d Rr = mul(Rs[start_frame], ric[0]);
    const int n = (int)it_per_id.feature_per_frame.size();
    for (int i = 0; i < n; i++) {
      const Vector3d t0 = add3(Ps[start_frame + i], mul(Rs[start_frame + i], tic[0])); const Matrix3d R0 = mul(Rs[start_frame + i], ric[0]);
      const FeaturePerFrame& fi = it_per_id.feature_per_frame[i];
      if (fi.depth < 0.1 || fi.depth > depth_threshold) continue;
      const Vector3d point0 = {fi.point.x * fi.depth, fi.point.y * fi.depth, fi.point.z * fi.depth};
      const Vector3d t2r = mul(transpose(Rr), sub3(t0, tr)); const Matrix3d R2r = mul(transpose(Rr), R0);
      for (int j = 0; j < n; j++) {
        if (i == j) continue;
        const Vector3d t1 = add3(Ps[start_frame + j], mul(Rs[start_frame + j], tic[0])); const Matrix3d R1 = mul(Rs[start_frame + j], ric[0]);
        const Vector3d t20 = mul(transpose(R0), sub3(t1, t0)); const Matrix3d R20 = mul(transpose(R0), R1);
        const Vector3d pp = sub3(mul(transpose(R20), point0), mul(transpose(R20), t20));
        const FeaturePerFrame& fj = it_per_id.feature_per_frame[j];
        const double rx = fj.point.x - pp.x / pp.z, ry = fj.point.y - pp.y / pp.z;
        if (std::sqrt(rx * rx + ry * ry) < 10.0 / 460) { const Vector3d point_r = add3(mul(R2r, point0), t2r); verified_depths.push_back(point_r.z); }
      }
    }
    if (verified_depths.empty()) continue;
    double depth_sum = 0.0; for (double d : verified_depths) depth_sum += d;
    it_per_id.estimated_depth = depth_sum / verified_depths.size();
    it_per_id.estimate_flag = 1;
    if (it_per_id.estimated_depth < 0.1) { it_per_id.estimated_depth = INIT_DEPTH; it_per_id.estimate_flag = 0; }
  }
}

// ------------------------------------------------------------------------------------------------ Estimator
Estimator::Estimator() { clearState(); memset(&last_summary, 0, sizeof(last_summary)); }
bool ImagePairSynchronizer::next(double* time, int* handle0, int* handle1) {   // rosNodeTest.cpp:395-428
  while (!img0_buf.empty() && !img1_buf.empty()) {
    const double time0 = img0_buf.front().first, time1 = img1_buf.front().first;
    if (time0 < time1 - 0.003) { img0_buf.pop(); thrown0++; }
    else if (time0 > time1 + 0.003) { img1_buf.pop(); thrown1++; }
    else { *time = time0; *handle0 = img0_buf.front().second; *handle1 = img1_buf.front().second; img0_buf.pop(); img1_buf.pop(); return true; }
  }
  return false;
}

Estimator::~Estimator() {
  delete featureTracker;
  if (gf2) gf2_solver_destroy(gf2);
  for (auto& p : pre_integrations) { delete p; p = nullptr; }
  for (auto& p : pre_integrations_wheel) { delete p; p = nullptr; }
}
void Estimator::setParameter(const Parameters& p) {
  P = p; tic[0] = p.TIC; ric[0] = p.RIC; td = p.TD;
  tio = p.TIO; rio = p.RIO; sx = p.SX; sy = p.SY; sw = p.SW; td_wheel = p.TD_WHEEL;
  f_manager.MIN_PARALLAX = p.MIN_PARALLAX;
  USE_MCC = p.USE_MCC != 0; DEPTH = p.DEPTH != 0;
  if (!featureTracker) featureTracker = new FeatureTracker();
  featureTracker->readIntrinsicParameter(p);
}
void Estimator::inputImage(double t, const uint8_t* _img, const uint16_t* _img1) {   // estimator.cpp:213-242
  inputImageCnt++;
  if (!featureTracker) { last_error = "setParameter has not been called"; return; }
  FeatureFrame featureFrame = featureTracker->trackImage(t, _img, _img1);
  if (featureTracker->lastError()[0]) { last_error = featureTracker->lastError(); return; }
  if (MULTIPLE_THREAD) { if (inputImageCnt % 2 == 0) featureBuf.push({t, featureFrame}); }
  else { featureBuf.push({t, featureFrame}); processMeasurements(); }
}
void Estimator::clearState() {
  for (int i = 0; i <= WINDOW_SIZE; i++) { Rs[i] = Matrix3d(); Ps[i] = Vector3d(); Vs[i] = Vector3d(); Bas[i] = Vector3d(); Bgs[i] = Vector3d(); delete pre_integrations[i]; pre_integrations[i] = nullptr; }
  for (auto& p : pre_integrations_wheel) { delete p; p = nullptr; }
  if (marg_pending && gf2) { int32_t st_ = 0, m_ = 0; gf2_marginalize_wait(gf2, 0, 1, &st_, &m_); }
  marg_pending = false; prior_on_device = false; host_prior_current = true;
  f_manager.feature.clear(); last_marginalization_info = MarginalizationPrior(); failure_occur = false; openExEstimation = false;
  openExWheelEstimation = false; openIxEstimation = false;
}

void Estimator::vector2double() {
  for (int i = 0; i <= WINDOW_SIZE; i++) {
    para_Pose[i][0] = Ps[i].x; para_Pose[i][1] = Ps[i].y; para_Pose[i][2] = Ps[i].z;
    const Quaterniond q = quatFromMatrix(Rs[i]);
    para_Pose[i][3] = q.x; para_Pose[i][4] = q.y; para_Pose[i][5] = q.z; para_Pose[i][6] = q.w;
    if (P.USE_IMU) {
      para_SpeedBias[i][0] = Vs[i].x; para_SpeedBias[i][1] = Vs[i].y; para_SpeedBias[i][2] = Vs[i].z;
      para_SpeedBias[i][3] = Bas[i].x; para_SpeedBias[i][4] = Bas[i].y; para_SpeedBias[i][5] = Bas[i].z;
      para_SpeedBias[i][6] = Bgs[i].x; para_SpeedBias[i][7] = Bgs[i].y; para_SpeedBias[i][8] = Bgs[i].z;
    }
  }
  para_Ex_Pose[0][0] = tic[0].x; para_Ex_Pose[0][1] = tic[0].y; para_Ex_Pose[0][2] = tic[0].z;
  const Quaterniond q = quatFromMatrix(ric[0]);
  para_Ex_Pose[0][3] = q.x; para_Ex_Pose[0][4] = q.y; para_Ex_Pose[0][5] = q.z; para_Ex_Pose[0][6] = q.w;
  const std::vector<double> dep = f_manager.getDepthVector();
  for (int i = 0; i < f_manager.getFeatureCount(); i++) para_Feature[i][0] = dep[i];
  para_Td[0][0] = td;
  // wheel blocks (estimator.cpp:2378-2390)
  para_Ex_Pose_wheel[0][0] = tio.x; para_Ex_Pose_wheel[0][1] = tio.y; para_Ex_Pose_wheel[0][2] = tio.z;
  const Quaterniond qw = quatFromMatrix(rio);
  para_Ex_Pose_wheel[0][3] = qw.x; para_Ex_Pose_wheel[0][4] = qw.y; para_Ex_Pose_wheel[0][5] = qw.z; para_Ex_Pose_wheel[0][6] = qw.w;
  para_Ix_sx_wheel[0][0] = sx; para_Ix_sy_wheel[0][0] = sy; para_Ix_sw_wheel[0][0] = sw;
  para_Td_wheel[0][0] = td_wheel;
}

void Estimator::double2vector() {
  Vector3d origin_R0 = R2ypr(Rs[0]);
  Vector3d origin_P0 = Ps[0];
  if (failure_occur) { origin_R0 = R2ypr(last_R0); origin_P0 = last_P0; failure_occur = false; }
  if (P.USE_IMU) {
    const Matrix3d R00 = toRotationMatrix({para_Pose[0][6], para_Pose[0][3], para_Pose[0][4], para_Pose[0][5]});
    const Vector3d origin_R00 = R2ypr(R00);
    const double y_diff = origin_R0.x - origin_R00.x;
    Matrix3d rot_diff = ypr2R({y_diff, 0, 0});
    if (std::fabs(std::fabs(origin_R0.y) - 90) < 1.0 || std::fabs(std::fabs(origin_R00.y) - 90) < 1.0) rot_diff = mul(Rs[0], transpose(R00));  // euler singular point
    for (int i = 0; i <= WINDOW_SIZE; i++) {
      Rs[i] = mul(rot_diff, toRotationMatrix(normalized({para_Pose[i][6], para_Pose[i][3], para_Pose[i][4], para_Pose[i][5]})));
      const Vector3d d = mul(rot_diff, Vector3d{para_Pose[i][0] - para_Pose[0][0], para_Pose[i][1] - para_Pose[0][1], para_Pose[i][2] - para_Pose[0][2]});
      Ps[i] = {d.x + origin_P0.x, d.y + origin_P0.y, d.z + origin_P0.z};
      Vs[i] = mul(rot_diff, Vector3d{para_SpeedBias[i][0], para_SpeedBias[i][1], para_SpeedBias[i][2]});
      Bas[i] = {para_SpeedBias[i][3], para_SpeedBias[i][4], para_SpeedBias[i][5]};
      Bgs[i] = {para_SpeedBias[i][6], para_SpeedBias[i][7], para_SpeedBias[i][8]};
    }
    tic[0] = {para_Ex_Pose[0][0], para_Ex_Pose[0][1], para_Ex_Pose[0][2]};
    ric[0] = toRotationMatrix({para_Ex_Pose[0][6], para_Ex_Pose[0][3], para_Ex_Pose[0][4], para_Ex_Pose[0][5]});
  } else {
    for (int i = 0; i <= WINDOW_SIZE; i++) {
      Rs[i] = toRotationMatrix(normalized({para_Pose[i][6], para_Pose[i][3], para_Pose[i][4], para_Pose[i][5]}));
      Ps[i] = {para_Pose[i][0], para_Pose[i][1], para_Pose[i][2]};
    }
  }
  std::vector<double> dep = f_manager.getDepthVector();
  for (int i = 0; i < f_manager.getFeatureCount(); i++) dep[i] = para_Feature[i][0];
  f_manager.setDepth(dep);
  if (P.USE_IMU) td = para_Td[0][0];
  if (P.USE_WHEEL) {  // estimator.cpp:2584-2606
    tio = {para_Ex_Pose_wheel[0][0], para_Ex_Pose_wheel[0][1], para_Ex_Pose_wheel[0][2]};
    rio = toRotationMatrix(normalized({para_Ex_Pose_wheel[0][6], para_Ex_Pose_wheel[0][3], para_Ex_Pose_wheel[0][4], para_Ex_Pose_wheel[0][5]}));
    sx = para_Ix_sx_wheel[0][0]; sy = para_Ix_sy_wheel[0][0]; sw = para_Ix_sw_wheel[0][0];
    td_wheel = para_Td_wheel[0][0];
  }
}

void Estimator::optimization() {
  last_error.clear();
  if (marg_pending) { finishMarginalization(capture); if (!last_error.empty()) return; }
  vector2double();
  const int F = frame_count + 1;
  if (!gf2) {
    gf2_solver_cfg cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.device = 0; cfg.max_windows = 1; cfg.n_frames = WINDOW_SIZE + 1; cfg.max_landmarks = NUM_OF_F; cfg.max_obs = NUM_OF_F * (WINDOW_SIZE + 1);
    cfg.max_imu_samples = kMaxImuSamples; cfg.max_prior_rows = GF2_MAX_PRIOR_DIM; cfg.max_planes = (WINDOW_SIZE + 1) * kMaxPlanesPerFrame;
    cfg.use_wheel = (P.USE_WHEEL && !P.ONLY_INITIAL_WITH_WHEEL) ? 1 : 0; cfg.max_wheel_samples = cfg.use_wheel ? kMaxWheelSamples : 0;
    if (gf2_solver_create(&cfg, &gf2) != GF2_OK) { last_error = gf2_last_error(); gf2 = nullptr; return; }
  }
  if (F != WINDOW_SIZE + 1) { last_error = "gf2host::Estimator::optimization handles the steady state frame_count == WINDOW_SIZE"; return; }
  const bool wheel_on = P.USE_WHEEL && !P.ONLY_INITIAL_WITH_WHEEL;   // estimator.cpp:3063, 3181
  double sxsysw[3] = {para_Ix_sx_wheel[0][0], para_Ix_sy_wheel[0][0], para_Ix_sw_wheel[0][0]};
  // landmark table in getDepthVector order (used_num >= 4, estimator.cpp:3330-3358)
  std::vector<int32_t> start, len; std::vector<uint8_t> fixed; std::vector<gf2_obs> obs; std::vector<double> frame_td(F, td);
  for (auto& it_per_id : f_manager.feature) {
    it_per_id.used_num = (int)it_per_id.feature_per_frame.size();
    if (it_per_id.used_num < 4) continue;
    start.push_back(it_per_id.start_frame); len.push_back(it_per_id.used_num); fixed.push_back(it_per_id.estimate_flag == 1);
    int f = it_per_id.start_frame;
    for (auto& pf : it_per_id.feature_per_frame) {
      obs.push_back({(float)pf.point.x, (float)pf.point.y, (float)pf.velocity[0], (float)pf.velocity[1]});
      if (f < F) frame_td[f] = pf.cur_td;
      f++;
    }
  }
  int32_t n_lm = (int32_t)start.size();
  std::vector<double> invdep(NUM_OF_F, 1.0); for (int i = 0; i < n_lm; i++) invdep[i] = para_Feature[i][0];
  start.resize(NUM_OF_F, 0); len.resize(NUM_OF_F, 0); fixed.resize(NUM_OF_F, 0); obs.resize((size_t)NUM_OF_F * (WINDOW_SIZE + 1));
  if (capture) {
    memcpy(cap.pose, para_Pose, sizeof(cap.pose)); memcpy(cap.sb, para_SpeedBias, sizeof(cap.sb)); memcpy(cap.ex, para_Ex_Pose[0], sizeof(cap.ex)); cap.td = para_Td[0][0];
    for (int i = 0; i < F; i++) cap.frame_td[i] = frame_td[i];
    cap.n_lm = n_lm; cap.start.assign(start.begin(), start.begin() + n_lm); cap.len.assign(len.begin(), len.begin() + n_lm); cap.fixed.assign(fixed.begin(), fixed.begin() + n_lm);
    cap.invdep.assign(invdep.begin(), invdep.begin() + n_lm);
    size_t no = 0; for (int i = 0; i < n_lm; i++) no += (size_t)len[i];
    cap.obs.assign(obs.begin(), obs.begin() + no);
  }
  int rc = gf2_set_states(gf2, 0, 1, &para_Pose[0][0], &para_SpeedBias[0][0], para_Ex_Pose[0], para_Td[0], wheel_on ? para_Ex_Pose_wheel[0] : nullptr, wheel_on ? sxsysw : nullptr, wheel_on ? para_Td_wheel[0] : nullptr);
  if (rc == GF2_OK) rc = gf2_set_landmarks(gf2, 0, 1, &n_lm, invdep.data(), start.data(), len.data(), fixed.data(), obs.data(), frame_td.data());
  // raw IMU samples of every interval -> device preintegration (IntegrationBase::push_back chain)
  if (rc == GF2_OK && P.USE_IMU) {
    std::vector<gf2_imu_sample> smp((size_t)(F - 1) * kMaxImuSamples); std::vector<int32_t> ns(F - 1, 0); std::vector<double> first((F - 1) * 6, 0.0), bias((F - 1) * 6, 0.0);
    for (int j = 1; j < F; j++) {
      const IntegrationBase* pi = pre_integrations[j];
      if (!pi) { last_error = "pre_integrations[j] missing"; return; }
      if (pi->dt_buf.size() > (size_t)kMaxImuSamples) { last_error = "an IMU interval holds more samples than the device buffer (kMaxImuSamples)"; return; }
      const int n = (int)pi->dt_buf.size();
      ns[j - 1] = n;
      for (int s = 0; s < n; s++) { gf2_imu_sample& o = smp[(size_t)(j - 1) * kMaxImuSamples + s]; o.dt = pi->dt_buf[s]; o.acc[0] = pi->acc_buf[s].x; o.acc[1] = pi->acc_buf[s].y; o.acc[2] = pi->acc_buf[s].z; o.gyr[0] = pi->gyr_buf[s].x; o.gyr[1] = pi->gyr_buf[s].y; o.gyr[2] = pi->gyr_buf[s].z; }
      double* f6 = &first[(j - 1) * 6]; f6[0] = pi->linearized_acc.x; f6[1] = pi->linearized_acc.y; f6[2] = pi->linearized_acc.z; f6[3] = pi->linearized_gyr.x; f6[4] = pi->linearized_gyr.y; f6[5] = pi->linearized_gyr.z;
      double* b6 = &bias[(j - 1) * 6]; b6[0] = pi->linearized_ba.x; b6[1] = pi->linearized_ba.y; b6[2] = pi->linearized_ba.z; b6[3] = pi->linearized_bg.x; b6[4] = pi->linearized_bg.y; b6[5] = pi->linearized_bg.z;
    }
    const double noise[4] = {P.ACC_N, P.GYR_N, P.ACC_W, P.GYR_W};
    if (capture) { cap.imu_samples = smp; cap.imu_n = ns; cap.imu_first = first; cap.imu_bias = bias; }
    rc = gf2_imu_preintegrate(gf2, 0, 1, smp.data(), ns.data(), first.data(), bias.data(), noise);
  }
  // raw wheel samples of every interval -> device preintegration (WheelIntegrationBase::push_back chain), estimator.cpp:3181-3212
  if (rc == GF2_OK && wheel_on) {
    std::vector<gf2_wheel_sample> smp((size_t)(F - 1) * kMaxWheelSamples); std::vector<int32_t> ns(F - 1, 0); std::vector<double> first((F - 1) * 6, 0.0), lin((F - 1) * 4, 0.0);
    for (int j = 1; j < F; j++) {
      const WheelIntegrationBase* pi = pre_integrations_wheel[j];
      if (!pi) { last_error = "pre_integrations_wheel[j] missing"; return; }
      if (pi->dt_buf.size() > (size_t)kMaxWheelSamples) { last_error = "a wheel interval holds more samples than the device buffer (kMaxWheelSamples)"; return; }
      const int n = (int)pi->dt_buf.size();
      ns[j - 1] = n;
      for (int k = 0; k < n; k++) { gf2_wheel_sample& o = smp[(size_t)(j - 1) * kMaxWheelSamples + k]; o.dt = pi->dt_buf[k]; o.vel[0] = pi->vel_buf[k].x; o.vel[1] = pi->vel_buf[k].y; o.vel[2] = pi->vel_buf[k].z; o.gyr[0] = pi->gyr_buf[k].x; o.gyr[1] = pi->gyr_buf[k].y; o.gyr[2] = pi->gyr_buf[k].z; }
      double* f6 = &first[(j - 1) * 6]; f6[0] = pi->linearized_vel.x; f6[1] = pi->linearized_vel.y; f6[2] = pi->linearized_vel.z; f6[3] = pi->linearized_gyr.x; f6[4] = pi->linearized_gyr.y; f6[5] = pi->linearized_gyr.z;
      double* l4 = &lin[(j - 1) * 4]; l4[0] = pi->linearized_sx; l4[1] = pi->linearized_sy; l4[2] = pi->linearized_sw; l4[3] = pi->linearized_td;
    }
    const double noise[2] = {P.VEL_N_wheel, P.GYR_N_wheel};
    if (capture) { cap.wheel_samples = smp; cap.wheel_n = ns; cap.wheel_first = first; cap.wheel_lin = lin; }
    rc = gf2_wheel_preintegrate(gf2, 0, 1, smp.data(), ns.data(), first.data(), lin.data(), noise);
    if (rc == GF2_OK && wdetect && wheelanomaly) {  // "wheel anomaly, skip optimization" (:3195-3199): no wheel factor enters the problem
      std::vector<gf2_wheel_preint> rec(F - 1);
      rc = gf2_get_wheel(gf2, 0, 1, rec.data());
      for (auto& r : rec) r.valid = 0;
      if (rc == GF2_OK) rc = gf2_set_wheel(gf2, 0, 1, rec.data());
    }
  }
  // LiDAR plane factors of the frames in the window (config 4 / 5 composition), frame index = window slot
  bool any_planes = false;
  if (rc == GF2_OK) {
    std::vector<gf2_plane> pl;
    for (int f = 0; f < F; f++) for (gf2_plane q : lidar_planes[f]) { q.frame = f; q.ct = 0; pl.push_back(q); }
    int32_t np = (int32_t)pl.size();
    any_planes = np > 0 || planes_resident;
    if (any_planes) {
      pl.resize((size_t)(WINDOW_SIZE + 1) * kMaxPlanesPerFrame);
      rc = gf2_set_planes(gf2, 0, 1, &np, pl.data());
      planes_resident = np > 0;
    }
    if (capture) { cap.planes.assign(pl.begin(), pl.begin() + np); }
  }
  if (rc == GF2_OK && capture && !host_prior_current) { finishMarginalization(true); if (!last_error.empty()) return; }
  if (rc == GF2_OK && (!prior_on_device || capture)) {   // otherwise the prior the last marginalization left on the device is the current one
    const MarginalizationPrior& mp = last_marginalization_info;
    int32_t rows = (mp.valid ? mp.n : 0), nb = (int32_t)mp.blocks.size();
    std::vector<double> J0((size_t)GF2_MAX_PRIOR_DIM * GF2_MAX_PRIOR_DIM, 0.0), r0(GF2_MAX_PRIOR_DIM, 0.0);
    std::vector<gf2_prior_block> blocks(2 * F + 8); memset(blocks.data(), 0, sizeof(gf2_prior_block) * blocks.size());
    for (int r = 0; r < rows; r++) { r0[r] = mp.linearized_residuals[r]; for (int c = 0; c < rows; c++) J0[(size_t)r * GF2_MAX_PRIOR_DIM + c] = mp.linearized_jacobians[(size_t)r * rows + c]; }
    for (int b = 0; b < nb && b < (int)blocks.size(); b++) blocks[b] = mp.blocks[b];
    if (capture) { cap.prior_rows = rows; cap.prior_nblocks = rows > 0 ? nb : 0; cap.prior_J0 = J0; cap.prior_r0 = r0; cap.prior_blocks = blocks; }
    rc = gf2_set_prior(gf2, 0, 1, &rows, J0.data(), r0.data(), &nb, blocks.data());
    if (rc == GF2_OK) prior_on_device = true;
  }
  if (rc == GF2_OK) {
    gf2_solve_opts o; memset(&o, 0, sizeof(o));
    o.max_iterations = P.NUM_ITERATIONS; o.huber_delta = 1.0; o.sqrt_info_px = FOCAL_LENGTH / 1.5; o.g_norm = P.G_NORM; o.lidar_sqrt_info = lidar_sqrt_info;
    // blocks held constant exactly when the reference calls SetParameterBlockConstant (estimator.cpp:3051-3060, 3158-3161)
    const double v0 = std::sqrt(Vs[0].x * Vs[0].x + Vs[0].y * Vs[0].y + Vs[0].z * Vs[0].z);
    if ((P.ESTIMATE_EXTRINSIC && frame_count == WINDOW_SIZE && v0 > 0.2) || openExEstimation) openExEstimation = true; else o.const_mask |= GF2_CONST_EX_POSE;
    if (!P.ESTIMATE_TD || v0 < 0.2) o.const_mask |= GF2_CONST_TD;
    // wheel calibration blocks (estimator.cpp:3063-3110, 3160-3161); extrinsic_type_wheel 0 = ADJUST_WHEEL_ALL (every shipped config)
    { static const uint32_t kSubset[5] = {0x00u, 0x38u, 0x07u, 0x04u, 0x3cu};   // ALL, TRANSLATION, ROTATION, NO_Z, NO_ROTATION_NO_Z (:3069-3089)
      o.wheel_ext_const_components = (wheel_on && P.ESTIMATE_EXTRINSIC_WHEEL && P.EXTRINSIC_TYPE_WHEEL >= 0 && P.EXTRINSIC_TYPE_WHEEL <= 4) ? kSubset[P.EXTRINSIC_TYPE_WHEEL] : 0u; }
    if ((wheel_on && P.ESTIMATE_EXTRINSIC_WHEEL && frame_count == WINDOW_SIZE && v0 > 0.2) || openExWheelEstimation) openExWheelEstimation = true; else o.const_mask |= GF2_CONST_EX_WHEEL;
    if ((wheel_on && P.ESTIMATE_INTRINSIC_WHEEL && frame_count == WINDOW_SIZE && v0 > 0.2) || openIxEstimation) openIxEstimation = true; else o.const_mask |= GF2_CONST_WHEEL_INTRINSIC;
    if (!P.ESTIMATE_TD_WHEEL || v0 < 0.2) o.const_mask |= GF2_CONST_TD_WHEEL;
    // options.max_solver_time_in_seconds = SOLVER_TIME (estimator.cpp:3373-3376; 0.04 s in the shipped configs): honoured on the device's clock.
    // The parity-test hook (capture) runs without the cap: a capped solve is machine dependent by construction.
    o.max_time_s = capture ? 0.0 : (marginalization_flag == MARGIN_OLD ? P.SOLVER_TIME * 4.0 / 5.0 : P.SOLVER_TIME);
    if (capture) {
      cap.const_mask = o.const_mask; cap.marg_mode = marginalization_flag == MARGIN_OLD ? 0 : 1; cap.use_wheel = wheel_on ? 1 : 0;
      memcpy(cap.exw, para_Ex_Pose_wheel[0], sizeof(cap.exw)); memcpy(cap.sxsysw, sxsysw, sizeof(cap.sxsysw)); cap.tdw = para_Td_wheel[0][0];
    }
    rc = gf2_solve(gf2, 0, 1, &o, &last_summary);
  }
  if (rc == GF2_OK) rc = gf2_get_states(gf2, 0, 1, &para_Pose[0][0], &para_SpeedBias[0][0], para_Ex_Pose[0], para_Td[0], wheel_on ? para_Ex_Pose_wheel[0] : nullptr, wheel_on ? sxsysw : nullptr, wheel_on ? para_Td_wheel[0] : nullptr);
  if (rc == GF2_OK) { rc = gf2_get_landmarks(gf2, 0, 1, invdep.data()); for (int i = 0; i < n_lm; i++) para_Feature[i][0] = invdep[i]; }
  if (rc != GF2_OK) { last_error = gf2_last_error(); return; }  // the reference logs and carries on (no exceptions)
  if (wheel_on) { para_Ix_sx_wheel[0][0] = sxsysw[0]; para_Ix_sy_wheel[0][0] = sxsysw[1]; para_Ix_sw_wheel[0][0] = sxsysw[2]; }
  if (capture) { memcpy(cap.exw_out, para_Ex_Pose_wheel[0], sizeof(cap.exw_out)); memcpy(cap.pose_out, para_Pose, sizeof(cap.pose_out)); memcpy(cap.sb_out, para_SpeedBias, sizeof(cap.sb_out)); cap.invdep_out.assign(invdep.begin(), invdep.begin() + n_lm); }
  double2vector();

  // ---- marginalization (estimator.cpp:3394-3690): runs at the states double2vector() left (yaw / position re-anchored),
  // re-packed by vector2double() exactly as the reference does at :3399 / :3603
  vector2double();
  for (int i = 0; i < n_lm; i++) invdep[i] = para_Feature[i][0];
  sxsysw[0] = para_Ix_sx_wheel[0][0]; sxsysw[1] = para_Ix_sy_wheel[0][0]; sxsysw[2] = para_Ix_sw_wheel[0][0];
  if (capture) { memcpy(cap.pose_marg, para_Pose, sizeof(cap.pose_marg)); memcpy(cap.sb_marg, para_SpeedBias, sizeof(cap.sb_marg)); cap.invdep_marg.assign(invdep.begin(), invdep.begin() + n_lm); }
  rc = gf2_set_states(gf2, 0, 1, &para_Pose[0][0], &para_SpeedBias[0][0], para_Ex_Pose[0], para_Td[0], wheel_on ? para_Ex_Pose_wheel[0] : nullptr, wheel_on ? sxsysw : nullptr, wheel_on ? para_Td_wheel[0] : nullptr);
  if (rc == GF2_OK) rc = gf2_set_landmarks(gf2, 0, 1, &n_lm, invdep.data(), start.data(), len.data(), fixed.data(), obs.data(), frame_td.data());
  gf2_solve_opts o; memset(&o, 0, sizeof(o));
  o.max_iterations = P.NUM_ITERATIONS; o.huber_delta = 1.0; o.sqrt_info_px = FOCAL_LENGTH / 1.5; o.g_norm = P.G_NORM; o.lidar_sqrt_info = 1.0;
  o.const_mask = GF2_CONST_EX_POSE | GF2_CONST_TD | GF2_CONST_EX_WHEEL | GF2_CONST_WHEEL_INTRINSIC | GF2_CONST_TD_WHEEL;
  if (rc == GF2_OK) rc = gf2_marginalize_async(gf2, 0, 1, marginalization_flag == MARGIN_OLD ? GF2_MARGIN_OLD : GF2_MARGIN_SECOND_NEW, &o);
  if (rc != GF2_OK) { last_error = gf2_last_error(); return; }
  // The new prior replaces the device-resident one in stream order and is only needed by the NEXT frame's solve: with
  // async_marginalization the call returns here (the kernels run while the node waits for the next image) and the prior never crosses
  // PCIe; the status is collected at the start of the next optimization() or by whoever asks for the prior (gf2h_get_prior, capture).
  marg_pending = true; prior_on_device = true; host_prior_current = false;
  if (capture || !async_marginalization) finishMarginalization(true);
}

// Collects the status of the pending gf2_marginalize_async and (optionally) downloads the new prior into last_marginalization_info.
void Estimator::finishMarginalization(bool download) {
  const int F = WINDOW_SIZE + 1;
  if (marg_pending) {
    int32_t st = 0, m = 0;
    marg_pending = false;
    if (gf2_marginalize_wait(gf2, 0, 1, &st, &m) != GF2_OK) { last_error = gf2_last_error(); return; }
    last_marginalization_status = st;
    if (st == GF2_MARG_INVALID) { last_marginalization_info = MarginalizationPrior(); host_prior_current = true; }   // valid = false (the device prior was cleared too)
    else if (st == GF2_MARG_UNCHANGED) { /* the previous prior stays on both sides (block indices untouched, as at :3599) */ }
    else if (st != 0) {
      // DEGENERATE / UNSUPPORTED / TOO_LARGE have no counterpart in the reference (it always produces a prior). Keeping the old prior while
      // slideWindow() shifts the frames would attach pose k's prior to the former frame k + 1: stop instead of corrupting the estimate.
      last_error = "gf2_marginalize: window status " + std::to_string(st) + " (no prior produced); estimation stopped";
      last_marginalization_info = MarginalizationPrior(); host_prior_current = true; prior_on_device = false;
      return;
    }
  }
  if (download && !host_prior_current && prior_on_device && gf2) {
    int32_t rows = 0, nb = 0;
    std::vector<double> J0((size_t)GF2_MAX_PRIOR_DIM * GF2_MAX_PRIOR_DIM), r0(GF2_MAX_PRIOR_DIM);
    std::vector<gf2_prior_block> blocks(2 * F + 8);
    if (gf2_get_prior(gf2, 0, 1, &rows, J0.data(), r0.data(), &nb, blocks.data()) != GF2_OK) { last_error = gf2_last_error(); return; }
    MarginalizationPrior& mp = last_marginalization_info;
    mp.valid = rows > 0; mp.n = rows; mp.linearized_jacobians.resize((size_t)rows * rows); mp.linearized_residuals.assign(r0.begin(), r0.begin() + rows);
    for (int r = 0; r < rows; r++) for (int c = 0; c < rows; c++) mp.linearized_jacobians[(size_t)r * rows + c] = J0[(size_t)r * GF2_MAX_PRIOR_DIM + c];
    mp.blocks.assign(blocks.begin(), blocks.begin() + (rows > 0 ? nb : 0));
    host_prior_current = true;
  }
}

// ---- measurement queues (estimator.cpp:324-372, 422-545, 554-763) ---------------------------------------------------------------------
void Estimator::inputIMU(double t, const Vector3d& a, const Vector3d& g) {   // :324-352 (the publishers are outside this build)
  accBuf.push({t, a}); gyrBuf.push({t, g});
  if (latest_valid) fastPredictIMU(t, a, g);
}
void Estimator::fastPredictIMU(double t, const Vector3d& linear_acceleration, const Vector3d& angular_velocity) {   // :4076-4093
  const double dt = t - latest_time; latest_time = t;
  const Vector3d g = {0.0, 0.0, P.G_NORM};
  auto sub = [](const Vector3d& a, const Vector3d& b) { return Vector3d{a.x - b.x, a.y - b.y, a.z - b.z}; };
  const Vector3d un_acc_0 = sub(mul(latest_Q, sub(latest_acc_0, latest_Ba)), g);
  const Vector3d un_gyr = sub(Vector3d{0.5 * (latest_gyr_0.x + angular_velocity.x), 0.5 * (latest_gyr_0.y + angular_velocity.y), 0.5 * (latest_gyr_0.z + angular_velocity.z)}, latest_Bg);
  latest_Q = mul(latest_Q, toRotationMatrix(normalized({1.0, un_gyr.x * dt / 2.0, un_gyr.y * dt / 2.0, un_gyr.z * dt / 2.0})));   // latest_Q * Utility::deltaQ(un_gyr * dt)
  const Vector3d un_acc_1 = sub(mul(latest_Q, sub(linear_acceleration, latest_Ba)), g);
  const Vector3d un_acc = {0.5 * (un_acc_0.x + un_acc_1.x), 0.5 * (un_acc_0.y + un_acc_1.y), 0.5 * (un_acc_0.z + un_acc_1.z)};
  latest_P = {latest_P.x + dt * latest_V.x + 0.5 * dt * dt * un_acc.x, latest_P.y + dt * latest_V.y + 0.5 * dt * dt * un_acc.y, latest_P.z + dt * latest_V.z + 0.5 * dt * dt * un_acc.z};
  latest_V = {latest_V.x + dt * un_acc.x, latest_V.y + dt * un_acc.y, latest_V.z + dt * un_acc.z};
  latest_acc_0 = linear_acceleration; latest_gyr_0 = angular_velocity;
}
void Estimator::updateLatestStates() {   // :4203-4228
  latest_time = Headers[frame_count] + td;
  latest_P = Ps[frame_count]; latest_Q = Rs[frame_count]; latest_V = Vs[frame_count]; latest_Ba = Bas[frame_count]; latest_Bg = Bgs[frame_count];
  latest_acc_0 = acc_0; latest_gyr_0 = gyr_0; latest_valid = true;
  std::queue<std::pair<double, Vector3d>> tmp_accBuf = accBuf, tmp_gyrBuf = gyrBuf;
  while (!tmp_accBuf.empty()) { fastPredictIMU(tmp_accBuf.front().first, tmp_accBuf.front().second, tmp_gyrBuf.front().second); tmp_accBuf.pop(); tmp_gyrBuf.pop(); }
}
void Estimator::inputWheel(double t, const Vector3d& v, const Vector3d& g) { wheelVelBuf.push({t, v}); wheelGyrBuf.push({t, g}); }
void Estimator::inputFeature(double t, const FeatureFrame& f) { featureBuf.push({t, f}); }
bool Estimator::IMUAvailable(double t) const { return !accBuf.empty() && t <= accBuf.back().first; }
bool Estimator::WheelAvailable(double t) const { return !wheelVelBuf.empty() && t <= wheelVelBuf.back().first; }
static bool takeInterval(std::queue<std::pair<double, Vector3d>>& A, std::queue<std::pair<double, Vector3d>>& B, double t0, double t1,
                         std::vector<std::pair<double, Vector3d>>& av, std::vector<std::pair<double, Vector3d>>& bv) {   // :422-455 / :456-526
  if (A.empty()) return false;
  if (!(t1 <= A.back().first)) return false;                      // "wait for imu"
  while (A.front().first <= t0) { A.pop(); B.pop(); }
  while (A.front().first < t1) { av.push_back(A.front()); A.pop(); bv.push_back(B.front()); B.pop(); }
  av.push_back(A.front()); bv.push_back(B.front());              // the first sample at or after t1 stays queued for the next interval
  return true;
}
bool Estimator::getIMUInterval(double t0, double t1, std::vector<std::pair<double, Vector3d>>& av, std::vector<std::pair<double, Vector3d>>& gv) { return takeInterval(accBuf, gyrBuf, t0, t1, av, gv); }
bool Estimator::getWheelInterval(double t0, double t1, std::vector<std::pair<double, Vector3d>>& vv, std::vector<std::pair<double, Vector3d>>& gv) { return takeInterval(wheelVelBuf, wheelGyrBuf, t0, t1, vv, gv); }
int Estimator::processMeasurements() {
  int consumed = 0;
  while (!featureBuf.empty()) {
    const std::pair<double, FeatureFrame>& feature = featureBuf.front();
    curTime = feature.first + td; curTime_wheel = curTime - td_wheel;
    const bool wheel_on = P.USE_WHEEL != 0;
    if (P.USE_IMU && !IMUAvailable(feature.first + td)) break;                      // MULTIPLE_THREAD == 0: return and wait
    if (wheel_on && !WheelAvailable(feature.first + td - td_wheel)) break;
    std::vector<std::pair<double, Vector3d>> accVector, gyrVector, velWheelVector, gyrWheelVector;
    if (P.USE_IMU) getIMUInterval(prevTime, curTime, accVector, gyrVector);
    const FeatureFrame image = feature.second; const double header = feature.first;
    featureBuf.pop();
    if (wheel_on) getWheelInterval(prevTime_wheel, curTime_wheel, velWheelVector, gyrWheelVector);
    for (size_t i = 0; i < accVector.size(); i++) {   // :640-651: the first and the last sample are cut at the image times
      double dt;
      if (i == 0) dt = accVector[i].first - prevTime;
      else if (i == accVector.size() - 1) dt = curTime - accVector[i - 1].first;
      else dt = accVector[i].first - accVector[i - 1].first;
      processIMU(accVector[i].first, dt, accVector[i].second, gyrVector[i].second);
    }
    for (size_t i = 0; i < velWheelVector.size(); i++) {
      double dt;
      if (i == 0) dt = velWheelVector[i].first - prevTime_wheel;
      else if (i == velWheelVector.size() - 1) dt = curTime_wheel - velWheelVector[i - 1].first;
      else dt = velWheelVector[i].first - velWheelVector[i - 1].first;
      processWheel(velWheelVector[i].first, dt, velWheelVector[i].second, gyrWheelVector[i].second);
    }
    processImage(image, header);
    prevTime = curTime; prevTime_wheel = curTime_wheel;
    consumed++;
    if (!last_error.empty()) break;
  }
  return consumed;
}

// ---- measurement processing around the solve (steady state) -------------------------------------------------------------------------
void Estimator::processIMU(double, double dt, const Vector3d& linear_acceleration, const Vector3d& angular_velocity) {   // estimator.cpp:795-836
  if (!first_imu) { first_imu = true; acc_0 = linear_acceleration; gyr_0 = angular_velocity; }
  if (!pre_integrations[frame_count]) pre_integrations[frame_count] = new IntegrationBase(acc_0, gyr_0, Bas[frame_count], Bgs[frame_count]);
  if (frame_count != 0) pre_integrations[frame_count]->push_back(dt, linear_acceleration, angular_velocity);   // dt_buf / *_buf live inside the buffer class
  acc_0 = linear_acceleration; gyr_0 = angular_velocity;
}
void Estimator::processWheel(double, double dt, const Vector3d& linear_velocity, const Vector3d& angular_velocity) {   // :837-896
  if (!first_wheel) { first_wheel = true; vel_0_wheel = linear_velocity; gyr_0_wheel = angular_velocity; }
  if (!pre_integrations_wheel[frame_count]) pre_integrations_wheel[frame_count] = new WheelIntegrationBase(vel_0_wheel, gyr_0_wheel, sx, sy, sw, td_wheel);
  if (frame_count != 0) {
    pre_integrations_wheel[frame_count]->push_back(dt, linear_velocity, angular_velocity);
    const int j = frame_count;
    const Vector3d un_gyr = {0.5 * (gyr_0_wheel.x + angular_velocity.x), 0.5 * (gyr_0_wheel.y + angular_velocity.y), 0.5 * (gyr_0_wheel.z + angular_velocity.z)};
    const Vector3d un_vel_0 = mul(Rs[j], latest_vel_wheel_0);
    // dead reckoning of the newest frame (systemstationary == false): Rs[j] *= deltaQ(un_gyr dt), Vs[j] = (Rs[j] v + un_vel_0) / 2, Ps[j] += dt Vs[j]
    Rs[j] = mul(Rs[j], toRotationMatrix(normalized({1.0, un_gyr.x * dt / 2.0, un_gyr.y * dt / 2.0, un_gyr.z * dt / 2.0})));
    const Vector3d rv = mul(Rs[j], linear_velocity);
    Vs[j] = {0.5 * (rv.x + un_vel_0.x), 0.5 * (rv.y + un_vel_0.y), 0.5 * (rv.z + un_vel_0.z)};
    Ps[j] = {Ps[j].x + dt * Vs[j].x, Ps[j].y + dt * Vs[j].y, Ps[j].z + dt * Vs[j].z};
    latest_vel_wheel_0 = linear_velocity;
  }
  vel_0_wheel = linear_velocity; gyr_0_wheel = angular_velocity;
}

static double reprojectionError(const Matrix3d& Ri, const Vector3d& Pi, const Matrix3d& rici, const Vector3d& tici, const Matrix3d& Rj, const Vector3d& Pj,
                                const Matrix3d& ricj, const Vector3d& ticj, double depth, const Vector3d& uvi, const Vector3d& uvj, double* err3d) {   // :3950-3969
  const Vector3d a = mul(rici, Vector3d{depth * uvi.x, depth * uvi.y, depth * uvi.z});
  const Vector3d b = mul(Ri, Vector3d{a.x + tici.x, a.y + tici.y, a.z + tici.z});
  const Vector3d pts_w = {b.x + Pi.x, b.y + Pi.y, b.z + Pi.z};
  const Vector3d c = mul(transpose(Rj), Vector3d{pts_w.x - Pj.x, pts_w.y - Pj.y, pts_w.z - Pj.z});
  const Vector3d pts_cj = mul(transpose(ricj), Vector3d{c.x - ticj.x, c.y - ticj.y, c.z - ticj.z});
  if (err3d) { const double dx = pts_cj.x - uvj.x, dy = pts_cj.y - uvj.y, dz = pts_cj.z - uvj.z; *err3d = std::sqrt(dx * dx + dy * dy + dz * dz) / depth; }
  const double rx = pts_cj.x / pts_cj.z - uvj.x, ry = pts_cj.y / pts_cj.z - uvj.y;
  return std::sqrt(rx * rx + ry * ry);
}
void Estimator::outliersRejection(std::set<int>& removeIndex) {   // :3971-4028
  for (auto& it_per_id : f_manager.feature) {
    double err = 0; int errCnt = 0;
    it_per_id.used_num = (int)it_per_id.feature_per_frame.size();
    if (it_per_id.used_num < 4) continue;
    const int imu_i = it_per_id.start_frame; int imu_j = imu_i - 1;
    const Vector3d pts_i = it_per_id.feature_per_frame[0].point; const double depth = it_per_id.estimated_depth;
    for (auto& it_per_frame : it_per_id.feature_per_frame) {
      imu_j++;
      if (imu_i != imu_j) { err += reprojectionError(Rs[imu_i], Ps[imu_i], ric[0], tic[0], Rs[imu_j], Ps[imu_j], ric[0], tic[0], depth, pts_i, it_per_frame.point, nullptr); errCnt++; }
    }
    const double ave_err = err / errCnt;
    if (ave_err * FOCAL_LENGTH > 3) removeIndex.insert(it_per_id.feature_id);
  }
}
void Estimator::movingConsistencyCheckW(std::set<int>& removeIndex) {   // :4030-4074
  for (auto& it_per_id : f_manager.feature) {
    it_per_id.used_num = (int)it_per_id.feature_per_frame.size();
    if (!(it_per_id.used_num >= 2 && it_per_id.start_frame < WINDOW_SIZE - 2)) continue;
    const double depth = it_per_id.estimated_depth;
    if (depth < 0) continue;
    double err = 0, err3D = 0; int errCnt = 0;
    const int wheel_i = it_per_id.start_frame; int wheel_j = wheel_i - 1;
    const Vector3d pts_i = it_per_id.feature_per_frame[0].point;
    for (auto& it_per_frame : it_per_id.feature_per_frame) {
      wheel_j++;
      if (wheel_i != wheel_j) {
        double e3 = 0;
        err += reprojectionError(Rs[wheel_i], Ps[wheel_i], ric[0], tic[0], Rs[wheel_j], Ps[wheel_j], ric[0], tic[0], depth, pts_i, it_per_frame.point, &e3);
        err3D += e3; errCnt++;
      }
    }
    if (errCnt > 0 && (FOCAL_LENGTH * err / errCnt > 10 || err3D / errCnt > 2.0)) removeIndex.insert(it_per_id.feature_id);
  }
}
void Estimator::getPoseInWorldFrame(int index, double T[16]) const {   // :3901-3913
  for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) T[r * 4 + c] = Rs[index].m[r * 3 + c];
  T[3] = Ps[index].x; T[7] = Ps[index].y; T[11] = Ps[index].z;
}
std::map<int, Vector3d> Estimator::predictPtsInNextFrame() const {   // :3915-3948: constant-velocity pose, landmarks of the newest frame into its camera
  std::map<int, Vector3d> predictPts;
  if (frame_count < 2) return predictPts;
  // nextT = curT * (prevT^-1 * curT)
  const Matrix3d Rc = Rs[frame_count], Rp = Rs[frame_count - 1]; const Vector3d Pc = Ps[frame_count], Pp = Ps[frame_count - 1];
  const Matrix3d Rrel = mul(transpose(Rp), Rc); const Vector3d trel = mul(transpose(Rp), Vector3d{Pc.x - Pp.x, Pc.y - Pp.y, Pc.z - Pp.z});
  const Matrix3d Rn = mul(Rc, Rrel); const Vector3d rn = mul(Rc, trel), Pn = {rn.x + Pc.x, rn.y + Pc.y, rn.z + Pc.z};
  for (auto& it_per_id : f_manager.feature) {
    if (!(it_per_id.estimated_depth > 0)) continue;
    const int firstIndex = it_per_id.start_frame, lastIndex = it_per_id.start_frame + (int)it_per_id.feature_per_frame.size() - 1;
    if ((int)it_per_id.feature_per_frame.size() >= 2 && lastIndex == frame_count) {
      const double depth = it_per_id.estimated_depth; const Vector3d p0 = it_per_id.feature_per_frame[0].point;
      const Vector3d a = mul(ric[0], Vector3d{depth * p0.x, depth * p0.y, depth * p0.z}), pts_j = {a.x + tic[0].x, a.y + tic[0].y, a.z + tic[0].z};
      const Vector3d b = mul(Rs[firstIndex], pts_j), pts_w = {b.x + Ps[firstIndex].x, b.y + Ps[firstIndex].y, b.z + Ps[firstIndex].z};
      const Vector3d pts_local = mul(transpose(Rn), Vector3d{pts_w.x - Pn.x, pts_w.y - Pn.y, pts_w.z - Pn.z});
      predictPts[it_per_id.feature_id] = mul(transpose(ric[0]), Vector3d{pts_local.x - tic[0].x, pts_local.y - tic[0].y, pts_local.z - tic[0].z});
    }
  }
  return predictPts;
}
std::string Estimator::tumLine(double stamp) const {   // visualization.cpp:371-385
  const Quaterniond q = quatFromMatrix(Rs[WINDOW_SIZE]);
  char buf[256];
  snprintf(buf, sizeof(buf), "%.9f %.9f %.9f %.9f %.9f %.9f %.9f %.9f\n", stamp, Ps[WINDOW_SIZE].x, Ps[WINDOW_SIZE].y, Ps[WINDOW_SIZE].z, q.x, q.y, q.z, q.w);
  return buf;
}
bool Estimator::appendTum(const std::string& path, double stamp) const {
  FILE* f = fopen(path.c_str(), "a");
  if (!f) return false;
  const std::string l = tumLine(stamp); fputs(l.c_str(), f); fclose(f);
  return true;
}
void Estimator::slideWindow() {   // :3700-3857 (GNSS buffers and all_image_frame belong to subsystems outside this build)
  if (marginalization_flag == MARGIN_OLD) {
    back_R0 = Rs[0]; back_P0 = Ps[0];
    if (frame_count == WINDOW_SIZE) {
      for (int i = 0; i < WINDOW_SIZE; i++) {
        Headers[i] = Headers[i + 1];
        std::swap(Rs[i], Rs[i + 1]); std::swap(Ps[i], Ps[i + 1]);
        if (P.USE_IMU) { std::swap(pre_integrations[i], pre_integrations[i + 1]); std::swap(Vs[i], Vs[i + 1]); std::swap(Bas[i], Bas[i + 1]); std::swap(Bgs[i], Bgs[i + 1]); }
        if (P.USE_WHEEL) std::swap(pre_integrations_wheel[i], pre_integrations_wheel[i + 1]);
        std::swap(lidar_planes[i], lidar_planes[i + 1]);
      }
      lidar_planes[WINDOW_SIZE].clear();   // the factors of the marginalized frame (now in the last slot) are dropped
      Headers[WINDOW_SIZE] = Headers[WINDOW_SIZE - 1]; Ps[WINDOW_SIZE] = Ps[WINDOW_SIZE - 1]; Rs[WINDOW_SIZE] = Rs[WINDOW_SIZE - 1];
      if (P.USE_IMU) {
        Vs[WINDOW_SIZE] = Vs[WINDOW_SIZE - 1]; Bas[WINDOW_SIZE] = Bas[WINDOW_SIZE - 1]; Bgs[WINDOW_SIZE] = Bgs[WINDOW_SIZE - 1];
        delete pre_integrations[WINDOW_SIZE];
        pre_integrations[WINDOW_SIZE] = new IntegrationBase(acc_0, gyr_0, Bas[WINDOW_SIZE], Bgs[WINDOW_SIZE]);
      }
      if (P.USE_WHEEL) { delete pre_integrations_wheel[WINDOW_SIZE]; pre_integrations_wheel[WINDOW_SIZE] = new WheelIntegrationBase(vel_0_wheel, gyr_0_wheel, sx, sy, sw, td_wheel); }
      slideWindowOld();
    }
  } else if (frame_count == WINDOW_SIZE) {
    Headers[frame_count - 1] = Headers[frame_count]; Ps[frame_count - 1] = Ps[frame_count]; Rs[frame_count - 1] = Rs[frame_count];
    if (P.USE_IMU) {
      const IntegrationBase* src = pre_integrations[frame_count];
      for (size_t i = 0; src && i < src->dt_buf.size(); i++) pre_integrations[frame_count - 1]->push_back(src->dt_buf[i], src->acc_buf[i], src->gyr_buf[i]);
      Vs[frame_count - 1] = Vs[frame_count]; Bas[frame_count - 1] = Bas[frame_count]; Bgs[frame_count - 1] = Bgs[frame_count];
      delete pre_integrations[WINDOW_SIZE];
      pre_integrations[WINDOW_SIZE] = new IntegrationBase(acc_0, gyr_0, Bas[WINDOW_SIZE], Bgs[WINDOW_SIZE]);
    }
    if (P.USE_WHEEL) {
      const WheelIntegrationBase* src = pre_integrations_wheel[frame_count];
      for (size_t i = 0; src && i < src->dt_buf.size(); i++) pre_integrations_wheel[frame_count - 1]->push_back(src->dt_buf[i], src->vel_buf[i], src->gyr_buf[i]);
      delete pre_integrations_wheel[WINDOW_SIZE];
      pre_integrations_wheel[WINDOW_SIZE] = new WheelIntegrationBase(vel_0_wheel, gyr_0_wheel, sx, sy, sw, td_wheel);
    }
    lidar_planes[frame_count - 1] = std::move(lidar_planes[frame_count]); lidar_planes[frame_count].clear();   // the newest scan takes the second-newest slot
    slideWindowNew();
  }
}
void Estimator::inputLidarPlanes(const gf2_plane* planes, int n) {
  n = std::min(n, kMaxPlanesPerFrame);
  lidar_planes[frame_count].assign(planes, planes + std::max(n, 0));
}
void Estimator::slideWindowNew() { sum_of_front++; f_manager.removeFront(frame_count); }   // :3859-3868
void Estimator::slideWindowOld() {   // :3870-3899 with solver_flag == NON_LINEAR (shift_depth)
  sum_of_back++;
  const Matrix3d R0 = mul(back_R0, ric[0]), R1 = mul(Rs[0], ric[0]);
  const Vector3d a = mul(back_R0, tic[0]), b = mul(Rs[0], tic[0]);
  f_manager.removeBackShiftDepth(R0, {back_P0.x + a.x, back_P0.y + a.y, back_P0.z + a.z}, R1, {Ps[0].x + b.x, Ps[0].y + b.y, Ps[0].z + b.z});
}
// processImage, the `else // not ini` branch (:1133-1215): keyframe decision, depth initialisation, solve + marginalization, moving-consistency
// outlier removal, window slide. failureDetection() returns false on its first line in the reference (:2910), so nothing is lost by not calling it;
// GNSS / line features / the image-frame map of the initialiser are outside this build.
void Estimator::processImage(const std::map<int, std::vector<std::pair<int, std::vector<double>>>>& image, double header) {
  marginalization_flag = f_manager.addFeatureCheckParallax(frame_count, image, td) ? MARGIN_OLD : MARGIN_SECOND_NEW;   // :900-911
  Headers[frame_count] = header;
  if (DEPTH) f_manager.triangulateWithDepth(frame_count, Ps, Rs, tic, ric);
  f_manager.triangulate(frame_count, Ps, Rs, tic, ric);
  std::set<int> removeIndex;
  if (USE_MCC) { movingConsistencyCheckW(removeIndex); f_manager.removeOutlier(removeIndex); }
  if (solve_enabled) {
    optimization();
    if (!last_error.empty()) return;
    // :1175-1182: the reference declares a SECOND, inner `set<int> removeIndex` in this branch, so with use_mcc: 0 (m3dgr / m2dgr / HILTI22
    // yaml) the tracker below receives the OUTER set, which is still empty: the feature manager drops the outliers, the tracker keeps them.
    if (!USE_MCC) { std::set<int> removeIndexInner; movingConsistencyCheckW(removeIndexInner); f_manager.removeOutlier(removeIndexInner); }
    if (!MULTIPLE_THREAD && featureTracker) {   // :1185-1189
      featureTracker->removeOutliers(removeIndex);
      featureTracker->setPrediction(predictPtsInNextFrame());
    }
  }
  slideWindow();
  f_manager.removeFailures();
  last_R0 = Rs[0]; last_P0 = Ps[0];
  updateLatestStates();   // :1214
}

// ------------------------------------------------------------------------------------------------ FeatureTracker
FeatureTracker::FeatureTracker() {}
FeatureTracker::~FeatureTracker() { if (trk) gf2_tracker_destroy(trk); }
void FeatureTracker::readIntrinsicParameter(const Parameters& p) {
  row = p.ROW; col = p.COL; MAX_CNT = p.MAX_CNT; MIN_DIST = p.MIN_DIST; FLOW_BACK = p.FLOW_BACK; EQUALIZE = p.EQUALIZE;
  fx = p.fx; fy = p.fy; cx = p.cx; cy = p.cy; k1 = p.k1; k2 = p.k2; p1 = p.p1; p2 = p.p2;
}
static inline int cvRound(double v) { return (int)std::nearbyint(v); }
bool FeatureTracker::inBorder(const Point2f& pt) const {
  const int BORDER_SIZE = 1;
  const int img_x = cvRound(pt.x), img_y = cvRound(pt.y);
  return BORDER_SIZE <= img_x && img_x < col - BORDER_SIZE && BORDER_SIZE <= img_y && img_y < row - BORDER_SIZE;
}
// cv::circle(mask, center, radius, 0, -1) with LINE_8 / shift 0: OpenCV's Bresenham Circle() in fill mode
static void fillCircle(std::vector<uint8_t>& img, int rows, int cols, int cxi, int cyi, int radius) {
  int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
  auto hline = [&](int y, int x0, int x1) { if (y < 0 || y >= rows) return; x0 = std::max(x0, 0); x1 = std::min(x1, cols - 1); for (int x = x0; x <= x1; x++) img[(size_t)y * cols + x] = 0; };
  while (dx >= dy) {
    hline(cyi - dy, cxi - dx, cxi + dx); hline(cyi + dy, cxi - dx, cxi + dx);
    hline(cyi - dx, cxi - dy, cxi + dy); hline(cyi + dx, cxi - dy, cxi + dy);
    dy++; err += plus; plus += 2;
    const int mask = (err <= 0) - 1;
    err -= minus & mask; dx += mask; minus -= mask & 2;
  }
}
void FeatureTracker::setMask() {
  mask.assign((size_t)row * col, 255);
  std::vector<std::pair<int, std::pair<Point2f, int>>> cnt_pts_id;
  for (unsigned int i = 0; i < cur_pts.size(); i++) cnt_pts_id.push_back(std::make_pair(track_cnt[i], std::make_pair(cur_pts[i], ids[i])));
  std::sort(cnt_pts_id.begin(), cnt_pts_id.end(), [](const std::pair<int, std::pair<Point2f, int>>& a, const std::pair<int, std::pair<Point2f, int>>& b) { return a.first > b.first; });
  cur_pts.clear(); ids.clear(); track_cnt.clear();
  for (auto& it : cnt_pts_id) {
    const int px = cvRound(it.second.first.x), py = cvRound(it.second.first.y);  // mask.at<uchar>(Point2f) -> Point via saturate_cast
    if (px >= 0 && px < col && py >= 0 && py < row && mask[(size_t)py * col + px] == 255) {
      cur_pts.push_back(it.second.first); ids.push_back(it.second.second); track_cnt.push_back(it.first);
      fillCircle(mask, row, col, px, py, MIN_DIST);
    }
  }
}
void FeatureTracker::addPoints() {
  for (auto& p : n_pts) { cur_pts.push_back(p); ids.push_back(n_id++); track_cnt.push_back(1); }
}
std::vector<Point2f> FeatureTracker::undistortedPts(const std::vector<Point2f>& pts) const {
  std::vector<Point2f> un_pts;
  const bool noDistortion = (k1 == 0.0 && k2 == 0.0 && p1 == 0.0 && p2 == 0.0);
  auto distortion = [&](double x, double y, double& dxo, double& dyo) {  // PinholeCamera::distortion
    const double mx2 = x * x, my2 = y * y, mxy = x * y, rho2 = mx2 + my2, rad = k1 * rho2 + k2 * rho2 * rho2;
    dxo = x * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2); dyo = y * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2);
  };
  for (auto& p : pts) {  // PinholeCamera::liftProjective (PinholeCamera.cc:450-510), recursive distortion model, n = 8
    const double mx_d = (1.0 / fx) * p.x + (-cx / fx), my_d = (1.0 / fy) * p.y + (-cy / fy);
    double mx_u = mx_d, my_u = my_d;
    if (!noDistortion) {
      double dx, dy; distortion(mx_d, my_d, dx, dy); mx_u = mx_d - dx; my_u = my_d - dy;
      for (int i = 1; i < 8; ++i) { distortion(mx_u, my_u, dx, dy); mx_u = mx_d - dx; my_u = my_d - dy; }
    }
    un_pts.push_back({(float)(mx_u / 1.0), (float)(my_u / 1.0)});
  }
  return un_pts;
}
std::vector<Point2f> FeatureTracker::ptsVelocity(const std::vector<int>& ids_, const std::vector<Point2f>& pts, std::map<int, Point2f>& cur_id_pts, std::map<int, Point2f>& prev_id_pts) {
  std::vector<Point2f> v;
  cur_id_pts.clear();
  for (unsigned int i = 0; i < ids_.size(); i++) cur_id_pts.insert(std::make_pair(ids_[i], pts[i]));
  if (!prev_id_pts.empty()) {
    const double dt = cur_time - prev_time;
    for (unsigned int i = 0; i < pts.size(); i++) {
      auto it = prev_id_pts.find(ids_[i]);
      if (it != prev_id_pts.end()) v.push_back({(float)((pts[i].x - it->second.x) / dt), (float)((pts[i].y - it->second.y) / dt)});
      else v.push_back({0, 0});
    }
  } else for (unsigned int i = 0; i < cur_pts.size(); i++) v.push_back({0, 0});
  return v;
}

void FeatureTracker::setPrediction(const std::map<int, Vector3d>& predictPts) {
  hasPrediction = true;
  predict_pts.clear();
  const bool noDistortion = (k1 == 0.0 && k2 == 0.0 && p1 == 0.0 && p2 == 0.0);
  for (size_t i = 0; i < ids.size(); i++) {
    auto it = predictPts.find(ids[i]);
    if (it != predictPts.end()) {  // PinholeCamera::spaceToPlane (PinholeCamera.cc:520-542)
      const double x = it->second.x / it->second.z, y = it->second.y / it->second.z;
      double xd = x, yd = y;
      if (!noDistortion) {
        const double mx2 = x * x, my2 = y * y, mxy = x * y, rho2 = mx2 + my2, rad = k1 * rho2 + k2 * rho2 * rho2;
        xd = x + (x * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2)); yd = y + (y * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2));
      }
      predict_pts.push_back({(float)(fx * xd + cx), (float)(fy * yd + cy)});
    } else predict_pts.push_back(prev_pts[i]);
  }
}
void FeatureTracker::removeOutliers(const std::set<int>& removePtsIds) {
  size_t j = 0;
  for (size_t i = 0; i < ids.size(); i++)
    if (removePtsIds.find(ids[i]) == removePtsIds.end()) { prev_pts[j] = prev_pts[i]; ids[j] = ids[i]; track_cnt[j] = track_cnt[i]; j++; }
  prev_pts.resize(j); ids.resize(j); track_cnt.resize(j);
}

std::map<int, std::vector<std::pair<int, std::vector<double>>>> FeatureTracker::trackImage(double _cur_time, const uint8_t* _img, const uint16_t* depth) {
  std::map<int, std::vector<std::pair<int, std::vector<double>>>> featureFrame;
  last_error.clear();
  cur_time = _cur_time;
  cur_img.assign(_img, _img + (size_t)row * col);
  cur_pts.clear();
  if (!trk) {
    gf2_tracker_cfg c; memset(&c, 0, sizeof(c));
    c.device = 0; c.width = col; c.height = row; c.max_pts = std::max(MAX_CNT, 8); c.win = 21; c.max_level = 3; c.max_iters = 30; c.max_streams = 1; c.eps = 0.01; c.min_eig = 1e-4;
    if (gf2_tracker_create(&c, &trk) != GF2_OK) { last_error = gf2_last_error(); trk = nullptr; return featureFrame; }
    // EQUALIZE: the node's cv::createCLAHE()->apply (VE/rosNodeTest.cpp:271-276) moves onto the device, fused into the upload
    if (EQUALIZE && gf2_tracker_set_equalize(trk, 40.0, 8, 8) != GF2_OK) { last_error = gf2_last_error(); return featureFrame; }
  }
  if (prev_pts.size() > 0) {
    int32_t n = (int32_t)prev_pts.size();
    cur_pts.resize(n);
    std::vector<uint8_t> status(std::max(MAX_CNT, 8), 0);
    std::vector<float> pin((size_t)std::max(MAX_CNT, 8) * 2, 0.f), pout((size_t)std::max(MAX_CNT, 8) * 2, 0.f);
    for (int i = 0; i < n; i++) { pin[2 * i] = prev_pts[i].x; pin[2 * i + 1] = prev_pts[i].y; }
    // forward LK (maxLevel 3) + reverse check (maxLevel 1, initial flow, <= 0.5 px): feature_tracker.cpp:135-153. The pyramid of the
    // previous image is still on the device (prev_img = cur_img, :307), so only the new image is uploaded.
    // With a prediction (:118-131) the forward pass starts from predict_pts at level 1 and falls back to level 3 when fewer than 10
    // points succeed; the decision is taken on the device inside gf2_tracker_track_image.
    std::vector<float> ppred;
    if (hasPrediction && (int)predict_pts.size() == n) {
      ppred.assign((size_t)std::max(MAX_CNT, 8) * 2, 0.f);
      for (int i = 0; i < n; i++) { ppred[2 * i] = predict_pts[i].x; ppred[2 * i + 1] = predict_pts[i].y; }
    }
    int rc = gf2_tracker_track_image(trk, 1, nullptr, cur_img.data(), (size_t)col, &n, pin.data(), ppred.empty() ? nullptr : ppred.data(), FLOW_BACK, pout.data(), status.data(), 3);
    if (rc != GF2_OK) { last_error = gf2_last_error(); return featureFrame; }
    if (EQUALIZE && gf2_tracker_get_image(trk, 1, cur_img.data()) != GF2_OK) { last_error = gf2_last_error(); return featureFrame; }  // cur_img as tracked
    for (int i = 0; i < n; i++) cur_pts[i] = {pout[2 * i], pout[2 * i + 1]};
    for (int i = 0; i < n; i++) {
      if (status[i] && !inBorder(cur_pts[i])) status[i] = 0;
      const int p_u = (int)cur_pts[i].x, p_v = (int)cur_pts[i].y;  // truncation, not rounding (:160-163)
      if (status[i] && p_u >= 0 && p_u < col && p_v >= 0 && p_v < row && cur_img[(size_t)p_v * col + p_u] > 250) status[i] = 0;
    }
    auto reduce_pts = [&](std::vector<Point2f>& v) { int j = 0; for (int i = 0; i < (int)v.size(); i++) if (status[i]) v[j++] = v[i]; v.resize(j); };
    auto reduce_int = [&](std::vector<int>& v) { int j = 0; for (int i = 0; i < (int)v.size(); i++) if (status[i]) v[j++] = v[i]; v.resize(j); };
    reduce_pts(prev_pts); reduce_pts(cur_pts); reduce_int(ids); reduce_int(track_cnt);
  } else {
    // first image: upload it so that its pyramid is the cached "prev" of the next call
    int32_t n0 = 0; const size_t cap = (size_t)std::max(MAX_CNT, 8);  // the ABI moves max_pts-sized point arrays
    std::vector<float> dummy_in(cap * 2, 0.f), dummy_out(cap * 2, 0.f); std::vector<uint8_t> st(cap, 0);
    if (gf2_tracker_track(trk, 1, cur_img.data(), cur_img.data(), (size_t)col, &n0, dummy_in.data(), dummy_out.data(), st.data(), nullptr, 0, 3) != GF2_OK) { last_error = gf2_last_error(); return featureFrame; }
    if (EQUALIZE && gf2_tracker_get_image(trk, 1, cur_img.data()) != GF2_OK) { last_error = gf2_last_error(); return featureFrame; }
  }
  for (auto& n : track_cnt) n++;
  setMask();
  const int n_max_cnt = MAX_CNT - (int)cur_pts.size();
  n_pts.clear();
  if (n_max_cnt > 0 && detector) {
    std::vector<float> xy((size_t)n_max_cnt * 2);
    const int got = detector(cur_img.data(), row, col, mask.data(), n_max_cnt, MIN_DIST, xy.data(), detector_user);
    for (int i = 0; i < got && i < n_max_cnt; i++) n_pts.push_back({xy[2 * i], xy[2 * i + 1]});
  } else if (n_max_cnt > 0) {
    // cv::goodFeaturesToTrack(cur_img, n_pts, MAX_CNT - cur_pts.size(), 0.01, MIN_DIST, mask) (:198) on the image already resident
    std::vector<float> xy((size_t)std::max(MAX_CNT, 8) * 2); int32_t want = n_max_cnt, got = 0;
    if (gf2_tracker_detect(trk, 1, nullptr, (size_t)col, mask.data(), &want, 0.01, (double)MIN_DIST, xy.data(), &got) != GF2_OK) { last_error = gf2_last_error(); return featureFrame; }
    for (int i = 0; i < got; i++) n_pts.push_back({xy[2 * i], xy[2 * i + 1]});
  }
  addPoints();
  cur_un_pts = undistortedPts(cur_pts);
  pts_velocity = ptsVelocity(ids, cur_un_pts, cur_un_pts_map, prev_un_pts_map);
  prev_pts = cur_pts; prev_un_pts = cur_un_pts; prev_un_pts_map = cur_un_pts_map; prev_time = cur_time; have_prev = true;
  hasPrediction = false;  // :312
  for (size_t i = 0; i < ids.size(); i++) {
    double depth_value = -2.4;  // "depthmono" of the mono branch (:331)
    if (depth) depth_value = (double)(int)depth[(size_t)std::lround(cur_pts[i].y) * col + std::lround(cur_pts[i].x)] / 1000;  // :360-361
    featureFrame[ids[i]].emplace_back(0, std::vector<double>{cur_un_pts[i].x, cur_un_pts[i].y, 1.0, cur_pts[i].x, cur_pts[i].y, pts_velocity[i].x, pts_velocity[i].y, depth_value});
  }
  return featureFrame;
}

}  // namespace gf2host

// ------------------------------------------------------------------------------------------------ flat C test API (ctypes)
using namespace gf2host;
extern "C" {
void* gf2h_estimator_create() { return new Estimator(); }
void gf2h_estimator_destroy(void* e) { delete (Estimator*)e; }
int gf2h_read_parameters(void* e, const char* path) { Parameters p; if (!readParameters(path, p)) return -1; ((Estimator*)e)->setParameter(p); return 0; }
void gf2h_get_parameters(void* e, double* out /*24*/) {
  const Parameters& p = ((Estimator*)e)->P;
  const double v[24] = {p.ACC_N, p.ACC_W, p.GYR_N, p.GYR_W, p.G_NORM, p.SOLVER_TIME, (double)p.NUM_ITERATIONS, (double)p.MAX_CNT, (double)p.MIN_DIST, (double)p.ROW, (double)p.COL,
                        p.fx, p.fy, p.cx, p.cy, p.TIC.x, p.TIC.y, p.TIC.z, p.RIC.m[0], p.RIC.m[4], p.RIC.m[8], p.MIN_PARALLAX, (double)p.ESTIMATE_EXTRINSIC, (double)p.FLOW_BACK};
  memcpy(out, v, sizeof(v));
}
// states as rows of [P(3) R(9 row-major) V(3) Ba(3) Bg(3)] = 21 doubles per frame
void gf2h_set_frame_states(void* e, const double* s) {
  Estimator* E = (Estimator*)e;
  for (int i = 0; i <= WINDOW_SIZE; i++) { const double* r = s + 21 * i; E->Ps[i] = {r[0], r[1], r[2]}; memcpy(E->Rs[i].m, r + 3, 72); E->Vs[i] = {r[12], r[13], r[14]}; E->Bas[i] = {r[15], r[16], r[17]}; E->Bgs[i] = {r[18], r[19], r[20]}; }
}
void gf2h_get_frame_states(void* e, double* s) {
  Estimator* E = (Estimator*)e;
  for (int i = 0; i <= WINDOW_SIZE; i++) { double* r = s + 21 * i; r[0] = E->Ps[i].x; r[1] = E->Ps[i].y; r[2] = E->Ps[i].z; memcpy(r + 3, E->Rs[i].m, 72); r[12] = E->Vs[i].x; r[13] = E->Vs[i].y; r[14] = E->Vs[i].z;
    r[15] = E->Bas[i].x; r[16] = E->Bas[i].y; r[17] = E->Bas[i].z; r[18] = E->Bgs[i].x; r[19] = E->Bgs[i].y; r[20] = E->Bgs[i].z; }
}
void gf2h_set_extrinsic(void* e, const double* tic3, const double* ric9, double td, double g_norm, const double noise[4]) {
  Estimator* E = (Estimator*)e; E->tic[0] = {tic3[0], tic3[1], tic3[2]}; memcpy(E->ric[0].m, ric9, 72); E->td = td; E->P.G_NORM = g_norm;
  E->P.ACC_N = noise[0]; E->P.GYR_N = noise[1]; E->P.ACC_W = noise[2]; E->P.GYR_W = noise[3];
}
// one image worth of features: ids[n], 8-vectors[n][8]; appended in std::map (ascending id) order like processImage does
void gf2h_add_image(void* e, int frame_count, int n, const int* ids, const double* pts8, double td) {
  std::map<int, std::vector<std::pair<int, std::vector<double>>>> image;
  for (int i = 0; i < n; i++) image[ids[i]].emplace_back(0, std::vector<double>(pts8 + 8 * i, pts8 + 8 * i + 8));
  ((Estimator*)e)->f_manager.addFeatures(frame_count, image, td);
}
void gf2h_set_depths(void* e, int n, const int* ids, const double* depth, const int* estimate_flag) {
  Estimator* E = (Estimator*)e;
  for (int i = 0; i < n; i++) for (auto& f : E->f_manager.feature) if (f.feature_id == ids[i]) { f.estimated_depth = depth[i]; if (estimate_flag) f.estimate_flag = estimate_flag[i]; }
}
int gf2h_feature_table(void* e, int max_n, int* ids, int* start, int* len, double* depth, int* solve_flag) {
  Estimator* E = (Estimator*)e; int k = 0;
  for (auto& f : E->f_manager.feature) { if (k >= max_n) break; ids[k] = f.feature_id; start[k] = f.start_frame; len[k] = (int)f.feature_per_frame.size(); depth[k] = f.estimated_depth; solve_flag[k] = f.solve_flag; k++; }
  return k;
}
int gf2h_depth_vector(void* e, double* out) { auto v = ((Estimator*)e)->f_manager.getDepthVector(); for (size_t i = 0; i < v.size(); i++) out[i] = v[i]; return (int)v.size(); }
void gf2h_new_interval(void* e, int j, const double* acc0, const double* gyr0, const double* ba, const double* bg) {
  Estimator* E = (Estimator*)e; delete E->pre_integrations[j];
  E->pre_integrations[j] = new IntegrationBase({acc0[0], acc0[1], acc0[2]}, {gyr0[0], gyr0[1], gyr0[2]}, {ba[0], ba[1], ba[2]}, {bg[0], bg[1], bg[2]});
}
void gf2h_push_imu(void* e, int j, double dt, const double* acc, const double* gyr) { ((Estimator*)e)->pre_integrations[j]->push_back(dt, {acc[0], acc[1], acc[2]}, {gyr[0], gyr[1], gyr[2]}); }
// wheel: calibration {tio 3, rio 9 row-major, sx, sy, sw, td_wheel} and flags {use_wheel, estimate_extrinsic, estimate_intrinsic, estimate_td, vel_n, gyr_n}
void gf2h_set_wheel_parameters(void* e, const double* calib16, const double* flags6) {
  Estimator* E = (Estimator*)e; Parameters p = E->P;
  p.TIO = {calib16[0], calib16[1], calib16[2]}; for (int i = 0; i < 9; i++) p.RIO.m[i] = calib16[3 + i];
  p.SX = calib16[12]; p.SY = calib16[13]; p.SW = calib16[14]; p.TD_WHEEL = calib16[15];
  p.USE_WHEEL = (int)flags6[0]; p.ESTIMATE_EXTRINSIC_WHEEL = (int)flags6[1]; p.ESTIMATE_INTRINSIC_WHEEL = (int)flags6[2]; p.ESTIMATE_TD_WHEEL = (int)flags6[3];
  p.VEL_N_wheel = flags6[4]; p.GYR_N_wheel = flags6[5];
  E->P = p; E->tio = p.TIO; E->rio = p.RIO; E->sx = p.SX; E->sy = p.SY; E->sw = p.SW; E->td_wheel = p.TD_WHEEL;
}
void gf2h_get_wheel_states(void* e, double* calib16, int* open_flags2) {
  Estimator* E = (Estimator*)e;
  calib16[0] = E->tio.x; calib16[1] = E->tio.y; calib16[2] = E->tio.z; for (int i = 0; i < 9; i++) calib16[3 + i] = E->rio.m[i];
  calib16[12] = E->sx; calib16[13] = E->sy; calib16[14] = E->sw; calib16[15] = E->td_wheel;
  open_flags2[0] = E->openExWheelEstimation; open_flags2[1] = E->openIxEstimation;
}
void gf2h_new_wheel_interval(void* e, int j, const double* vel0, const double* gyr0) {
  Estimator* E = (Estimator*)e; delete E->pre_integrations_wheel[j];
  E->pre_integrations_wheel[j] = new WheelIntegrationBase({vel0[0], vel0[1], vel0[2]}, {gyr0[0], gyr0[1], gyr0[2]}, E->sx, E->sy, E->sw, E->td_wheel);
}
void gf2h_push_wheel(void* e, int j, double dt, const double* vel, const double* gyr) { ((Estimator*)e)->pre_integrations_wheel[j]->push_back(dt, {vel[0], vel[1], vel[2]}, {gyr[0], gyr[1], gyr[2]}); }
// ---- FeatureManager bookkeeping / window glue (CPU-only logic: tests/test_host_cpu.py drives it without a device)
static std::map<int, std::vector<std::pair<int, std::vector<double>>>> image_of(int n, const int* ids, const double* pts8) {
  std::map<int, std::vector<std::pair<int, std::vector<double>>>> image;
  for (int i = 0; i < n; i++) image[ids[i]].emplace_back(0, std::vector<double>(pts8 + 8 * i, pts8 + 8 * i + 8));
  return image;
}
int gf2h_add_feature_check_parallax(void* e, int frame_count, int n, const int* ids, const double* pts8, double td, double* stats4) {
  FeatureManager& fm = ((Estimator*)e)->f_manager;
  const bool key = fm.addFeatureCheckParallax(frame_count, image_of(n, ids, pts8), td);
  if (stats4) { stats4[0] = fm.last_track_num; stats4[1] = fm.new_feature_num; stats4[2] = fm.long_track_num; stats4[3] = fm.last_average_parallax; }
  return key ? 1 : 0;
}
void gf2h_set_min_parallax(void* e, double v) { ((Estimator*)e)->f_manager.MIN_PARALLAX = v; }
void gf2h_remove_back(void* e) { ((Estimator*)e)->f_manager.removeBack(); }
void gf2h_remove_front(void* e, int frame_count) { ((Estimator*)e)->f_manager.removeFront(frame_count); }
void gf2h_remove_failures(void* e) { ((Estimator*)e)->f_manager.removeFailures(); }
void gf2h_remove_outlier(void* e, int n, const int* ids) { ((Estimator*)e)->f_manager.removeOutlier(std::set<int>(ids, ids + n)); }
void gf2h_remove_back_shift_depth(void* e, const double* mR, const double* mP, const double* nR, const double* nP) {
  Matrix3d a, b; for (int i = 0; i < 9; i++) { a.m[i] = mR[i]; b.m[i] = nR[i]; }
  ((Estimator*)e)->f_manager.removeBackShiftDepth(a, {mP[0], mP[1], mP[2]}, b, {nP[0], nP[1], nP[2]});
}
void gf2h_triangulate(void* e, int with_depth) {
  Estimator* E = (Estimator*)e;
  if (with_depth) E->f_manager.triangulateWithDepth(E->frame_count, E->Ps, E->Rs, E->tic, E->ric); else E->f_manager.triangulate(E->frame_count, E->Ps, E->Rs, E->tic, E->ric);
}
int gf2h_feature_obs(void* e, int feature_id, int max_n, double* out8, int* estimate_flag) {   // rows [x y z u v vx vy depth] of one landmark's track
  for (auto& f : ((Estimator*)e)->f_manager.feature) if (f.feature_id == feature_id) {
    int k = 0;
    for (auto& pf : f.feature_per_frame) { if (k >= max_n) break; double* o = out8 + 8 * k++; o[0] = pf.point.x; o[1] = pf.point.y; o[2] = pf.point.z; o[3] = pf.uv[0]; o[4] = pf.uv[1]; o[5] = pf.velocity[0]; o[6] = pf.velocity[1]; o[7] = pf.depth; }
    if (estimate_flag) *estimate_flag = f.estimate_flag;
    return k;
  }
  return -1;
}
int gf2h_check_outliers(void* e, int which /* 0 outliersRejection, 1 movingConsistencyCheckW */, int max_n, int* ids) {
  std::set<int> idx; if (which) ((Estimator*)e)->movingConsistencyCheckW(idx); else ((Estimator*)e)->outliersRejection(idx);
  int k = 0; for (int i : idx) if (k < max_n) ids[k++] = i;
  return (int)idx.size();
}
int gf2h_predict_pts(void* e, int max_n, int* ids, double* xyz) {
  auto m = ((Estimator*)e)->predictPtsInNextFrame(); int k = 0;
  for (auto& kv : m) if (k < max_n) { ids[k] = kv.first; xyz[3 * k] = kv.second.x; xyz[3 * k + 1] = kv.second.y; xyz[3 * k + 2] = kv.second.z; k++; }
  return (int)m.size();
}
void gf2h_set_flags(void* e, int use_imu, int use_wheel, int depth, int use_mcc) { Estimator* E = (Estimator*)e; E->P.USE_IMU = use_imu; E->P.USE_WHEEL = use_wheel; E->DEPTH = depth != 0; E->USE_MCC = use_mcc != 0; }
void gf2h_process_imu(void* e, double t, double dt, const double* acc, const double* gyr) { ((Estimator*)e)->processIMU(t, dt, {acc[0], acc[1], acc[2]}, {gyr[0], gyr[1], gyr[2]}); }
void gf2h_process_wheel(void* e, double t, double dt, const double* vel, const double* gyr) { ((Estimator*)e)->processWheel(t, dt, {vel[0], vel[1], vel[2]}, {gyr[0], gyr[1], gyr[2]}); }
void gf2h_slide_window(void* e) { ((Estimator*)e)->slideWindow(); }
void gf2h_set_headers(void* e, const double* h11) { for (int i = 0; i <= WINDOW_SIZE; i++) ((Estimator*)e)->Headers[i] = h11[i]; }
void gf2h_get_headers(void* e, double* h11, int* counters2) { Estimator* E = (Estimator*)e; for (int i = 0; i <= WINDOW_SIZE; i++) h11[i] = E->Headers[i]; counters2[0] = E->sum_of_back; counters2[1] = E->sum_of_front; }
int gf2h_interval_samples(void* e, int j, int wheel, int max_n, double* dt) {   // dt_buf of pre_integrations[j] / pre_integrations_wheel[j]
  Estimator* E = (Estimator*)e; const std::vector<double>* b = nullptr;
  if (wheel) { if (E->pre_integrations_wheel[j]) b = &E->pre_integrations_wheel[j]->dt_buf; } else if (E->pre_integrations[j]) b = &E->pre_integrations[j]->dt_buf;
  if (!b) return -1;
  for (size_t i = 0; i < b->size() && (int)i < max_n; i++) dt[i] = (*b)[i];
  return (int)b->size();
}
int gf2h_append_tum(void* e, const char* path, double stamp) { return ((Estimator*)e)->appendTum(path, stamp) ? 0 : -1; }
void gf2h_set_imu0(void* e, const double* acc, const double* gyr) {   // the sample the next interval starts from (acc_0 / gyr_0 of processIMU)
  Estimator* E = (Estimator*)e; E->first_imu = true; E->acc_0 = {acc[0], acc[1], acc[2]}; E->gyr_0 = {gyr[0], gyr[1], gyr[2]};
}
// single-threaded public API of the reference: setParameter -> inputIMU / inputImage
void gf2h_set_tracker_parameters(void* e, int rows, int cols, int max_cnt, int min_dist, int equalize, const double* intr8) {
  Estimator* E = (Estimator*)e; Parameters p = E->P;
  p.ROW = rows; p.COL = cols; p.MAX_CNT = max_cnt; p.MIN_DIST = min_dist; p.FLOW_BACK = 1; p.EQUALIZE = equalize;
  p.fx = intr8[0]; p.fy = intr8[1]; p.cx = intr8[2]; p.cy = intr8[3]; p.k1 = intr8[4]; p.k2 = intr8[5]; p.p1 = intr8[6]; p.p2 = intr8[7];
  p.TIC = E->tic[0]; p.RIC = E->ric[0]; p.TD = E->td; p.TIO = E->tio; p.RIO = E->rio; p.SX = E->sx; p.SY = E->sy; p.SW = E->sw; p.TD_WHEEL = E->td_wheel;
  p.MIN_PARALLAX = E->f_manager.MIN_PARALLAX;
  E->setParameter(p);
}
int gf2h_input_image(void* e, double t, const uint8_t* img, const uint16_t* depth) {
  Estimator* E = (Estimator*)e; E->inputImage(t, img, depth);
  return E->lastError()[0] ? -1 : (E->marginalization_flag == Estimator::MARGIN_OLD ? 0 : 1);
}
// the tracker's current feature table rows [id, x, y, 1, u, v, vx, vy, depth, track_cnt] (used to fill the first window in the replay test)
int gf2h_estimator_track_only(void* e, double t, const uint8_t* img, const uint16_t* depth, int max_n, double* out10);
void* gf2h_sync_create() { return new ImagePairSynchronizer(); }
void gf2h_sync_destroy(void* s) { delete (ImagePairSynchronizer*)s; }
void gf2h_sync_push(void* s, int which, double t, int handle) { if (which) ((ImagePairSynchronizer*)s)->push1(t, handle); else ((ImagePairSynchronizer*)s)->push0(t, handle); }
int gf2h_sync_next(void* s, double* t, int* h0, int* h1, int* thrown2) {
  ImagePairSynchronizer* S = (ImagePairSynchronizer*)s; const bool ok = S->next(t, h0, h1); thrown2[0] = S->thrown0; thrown2[1] = S->thrown1; return ok ? 1 : 0;
}
void gf2h_update_latest_states(void* e) { ((Estimator*)e)->updateLatestStates(); }
void gf2h_get_latest(void* e, double* out16 /* time, P 3, R 9 row-major, V 3 */) {
  Estimator* E = (Estimator*)e; out16[0] = E->latest_time; out16[1] = E->latest_P.x; out16[2] = E->latest_P.y; out16[3] = E->latest_P.z;
  for (int i = 0; i < 9; i++) out16[4 + i] = E->latest_Q.m[i];
  out16[13] = E->latest_V.x; out16[14] = E->latest_V.y; out16[15] = E->latest_V.z;
}
void gf2h_set_solve_enabled(void* e, int on) { ((Estimator*)e)->solve_enabled = on != 0; }
void gf2h_input_imu(void* e, double t, const double* acc, const double* gyr) { ((Estimator*)e)->inputIMU(t, {acc[0], acc[1], acc[2]}, {gyr[0], gyr[1], gyr[2]}); }
void gf2h_input_wheel(void* e, double t, const double* vel, const double* gyr) { ((Estimator*)e)->inputWheel(t, {vel[0], vel[1], vel[2]}, {gyr[0], gyr[1], gyr[2]}); }
void gf2h_input_feature(void* e, double t, int n, const int* ids, const double* pts8) { ((Estimator*)e)->inputFeature(t, image_of(n, ids, pts8)); }
void gf2h_set_prev_time(void* e, double prev, double prev_wheel) { ((Estimator*)e)->prevTime = prev; ((Estimator*)e)->prevTime_wheel = prev_wheel; }
int gf2h_process_measurements(void* e) { return ((Estimator*)e)->processMeasurements(); }
int gf2h_queue_sizes(void* e, int* s3) { Estimator* E = (Estimator*)e; s3[0] = (int)E->accBuf.size(); s3[1] = (int)E->wheelVelBuf.size(); s3[2] = (int)E->featureBuf.size(); return 0; }
void gf2h_set_capture(void* e, int on) { ((Estimator*)e)->capture = on != 0; }
// sizes: [n_lm, n_obs, prior_rows, prior_nblocks, const_mask, marg_mode]
void gf2h_capture_sizes(void* e, int* s6) {
  const Estimator::Capture& c = ((Estimator*)e)->cap;
  s6[0] = c.n_lm; s6[1] = (int)c.obs.size(); s6[2] = c.prior_rows; s6[3] = c.prior_nblocks; s6[4] = (int)c.const_mask; s6[5] = c.marg_mode;
}
void gf2h_capture_get(void* e, double* pose, double* sb, double* ex_td8, double* frame_td, int32_t* start, int32_t* len, uint8_t* fixed, double* invdep, gf2_obs* obs,
                      gf2_imu_sample* imu_samples /*[10][kMaxImuSamples]*/, int32_t* imu_n, double* imu_first, double* imu_bias, double* prior_J0 /* [96][96] */, double* prior_r0,
                      gf2_prior_block* prior_blocks, double* pose_out, double* sb_out, double* invdep_out, double* pose_marg, double* sb_marg, double* invdep_marg) {
  const Estimator::Capture& c = ((Estimator*)e)->cap;
  memcpy(pose, c.pose, sizeof(c.pose)); memcpy(sb, c.sb, sizeof(c.sb)); memcpy(ex_td8, c.ex, sizeof(c.ex)); ex_td8[7] = c.td; memcpy(frame_td, c.frame_td, sizeof(c.frame_td));
  for (int i = 0; i < c.n_lm; i++) { start[i] = c.start[i]; len[i] = c.len[i]; fixed[i] = c.fixed[i]; invdep[i] = c.invdep[i]; invdep_out[i] = c.invdep_out[i]; invdep_marg[i] = c.invdep_marg[i]; }
  for (size_t i = 0; i < c.obs.size(); i++) obs[i] = c.obs[i];
  for (size_t i = 0; i < c.imu_samples.size(); i++) imu_samples[i] = c.imu_samples[i];
  for (size_t i = 0; i < c.imu_n.size(); i++) imu_n[i] = c.imu_n[i];
  for (size_t i = 0; i < c.imu_first.size(); i++) { imu_first[i] = c.imu_first[i]; imu_bias[i] = c.imu_bias[i]; }
  for (size_t i = 0; i < c.prior_J0.size(); i++) prior_J0[i] = c.prior_J0[i];
  for (size_t i = 0; i < c.prior_r0.size(); i++) prior_r0[i] = c.prior_r0[i];
  for (size_t i = 0; i < c.prior_blocks.size(); i++) prior_blocks[i] = c.prior_blocks[i];
  memcpy(pose_out, c.pose_out, sizeof(c.pose_out)); memcpy(sb_out, c.sb_out, sizeof(c.sb_out)); memcpy(pose_marg, c.pose_marg, sizeof(c.pose_marg)); memcpy(sb_marg, c.sb_marg, sizeof(c.sb_marg));
}
// wheel part of the capture: returns use_wheel; calib12 = [ex_wheel 7 | sx sy sw | td_wheel | -], exw_out 7 = para_Ex_Pose_wheel after the solve
int gf2h_capture_get_wheel(void* e, gf2_wheel_sample* samples /*[10][kMaxWheelSamples = 64]*/, int32_t* n, double* first, double* lin, double* calib12, double* exw_out) {
  const Estimator::Capture& c = ((Estimator*)e)->cap;
  for (size_t i = 0; i < c.wheel_samples.size(); i++) samples[i] = c.wheel_samples[i];
  for (size_t i = 0; i < c.wheel_n.size(); i++) n[i] = c.wheel_n[i];
  for (size_t i = 0; i < c.wheel_first.size(); i++) first[i] = c.wheel_first[i];
  for (size_t i = 0; i < c.wheel_lin.size(); i++) lin[i] = c.wheel_lin[i];
  memcpy(calib12, c.exw, sizeof(c.exw)); calib12[7] = c.sxsysw[0]; calib12[8] = c.sxsysw[1]; calib12[9] = c.sxsysw[2]; calib12[10] = c.tdw; calib12[11] = 0.0;
  memcpy(exw_out, c.exw_out, sizeof(c.exw_out));
  return c.use_wheel;
}
void gf2h_input_lidar_planes(void* e, int n, const gf2_plane* planes, double sqrt_info) { Estimator* E = (Estimator*)e; if (sqrt_info > 0) E->lidar_sqrt_info = sqrt_info; E->inputLidarPlanes(planes, n); }
int gf2h_capture_get_planes(void* e, gf2_plane* out, int cap_n) {
  const Estimator::Capture& c = ((Estimator*)e)->cap;
  const int n = std::min((int)c.planes.size(), cap_n);
  for (int i = 0; i < n; i++) out[i] = c.planes[i];
  return (int)c.planes.size();
}
int gf2h_process_image(void* e, int n, const int* ids, const double* pts8, double header) {
  Estimator* E = (Estimator*)e; E->processImage(image_of(n, ids, pts8), header);
  return E->lastError()[0] ? -1 : (E->marginalization_flag == Estimator::MARGIN_OLD ? 0 : 1);
}
void gf2h_set_prior(void* e, int n, const double* J0, const double* r0, int nblocks, const gf2_prior_block* blocks) {
  Estimator* E_ = (Estimator*)e; if (E_->marg_pending) E_->finishMarginalization(false);
  E_->prior_on_device = false; E_->host_prior_current = true;   // the host copy is the truth now: uploaded by the next optimization()
  MarginalizationPrior& mp = E_->last_marginalization_info;
  mp.valid = n > 0; mp.n = n; mp.linearized_jacobians.assign(J0, J0 + (size_t)n * n); mp.linearized_residuals.assign(r0, r0 + n); mp.blocks.assign(blocks, blocks + nblocks);
}
void gf2h_vector2double(void* e, double* pose /*11x7*/, double* sb /*11x9*/, double* ex /*7*/) {
  Estimator* E = (Estimator*)e; E->vector2double(); memcpy(pose, E->para_Pose, sizeof(E->para_Pose)); memcpy(sb, E->para_SpeedBias, sizeof(E->para_SpeedBias)); memcpy(ex, E->para_Ex_Pose[0], 56);
}
void gf2h_double2vector(void* e, const double* pose, const double* sb, int n_feat, const double* feat) {
  Estimator* E = (Estimator*)e; memcpy(E->para_Pose, pose, sizeof(E->para_Pose)); memcpy(E->para_SpeedBias, sb, sizeof(E->para_SpeedBias));
  for (int i = 0; i < n_feat; i++) E->para_Feature[i][0] = feat[i];
  E->double2vector();
}
// Blocks until the marginalization launched by the last processImage() is done (what the inter-frame gap of a live stream provides).
int gf2h_finish_marginalization(void* e) { Estimator* E = (Estimator*)e; E->finishMarginalization(false); return E->lastError()[0] ? -1 : 0; }
void gf2h_set_async_marginalization(void* e, int on) { ((Estimator*)e)->async_marginalization = on != 0; }
void gf2h_set_marginalization_flag(void* e, int flag) { ((Estimator*)e)->marginalization_flag = flag ? Estimator::MARGIN_SECOND_NEW : Estimator::MARGIN_OLD; }
int gf2h_get_prior(void* e, int* n, double* J0 /* n*n */, double* r0, int* nblocks, gf2_prior_block* blocks, int* status) {
  Estimator* E = (Estimator*)e; E->finishMarginalization(true);   // collects a pending marginalization and fetches the device-resident prior
  const MarginalizationPrior& mp = E->last_marginalization_info;
  *n = mp.valid ? mp.n : 0; *nblocks = (int)mp.blocks.size(); *status = E->last_marginalization_status;
  for (size_t i = 0; i < mp.linearized_jacobians.size(); i++) J0[i] = mp.linearized_jacobians[i];
  for (size_t i = 0; i < mp.linearized_residuals.size(); i++) r0[i] = mp.linearized_residuals[i];
  for (size_t i = 0; i < mp.blocks.size(); i++) blocks[i] = mp.blocks[i];
  return mp.valid ? 1 : 0;
}
void gf2h_get_para(void* e, double* pose /*11x7*/, double* sb /*11x9*/, double* feat /*NUM_OF_F*/) {
  Estimator* E = (Estimator*)e; memcpy(pose, E->para_Pose, sizeof(E->para_Pose)); memcpy(sb, E->para_SpeedBias, sizeof(E->para_SpeedBias));
  for (int i = 0; i < NUM_OF_F; i++) feat[i] = E->para_Feature[i][0];
}
int gf2h_optimization(void* e, gf2_solve_summary* s) { Estimator* E = (Estimator*)e; E->optimization(); if (s) *s = E->last_summary; return E->lastError()[0] ? -1 : 0; }
const char* gf2h_last_error(void* e) { return ((Estimator*)e)->lastError(); }

void* gf2h_tracker_create(int rows, int cols, int max_cnt, int min_dist, const double* intr8) {
  FeatureTracker* t = new FeatureTracker(); Parameters p; p.ROW = rows; p.COL = cols; p.MAX_CNT = max_cnt; p.MIN_DIST = min_dist; p.FLOW_BACK = 1;
  p.fx = intr8[0]; p.fy = intr8[1]; p.cx = intr8[2]; p.cy = intr8[3]; p.k1 = intr8[4]; p.k2 = intr8[5]; p.p1 = intr8[6]; p.p2 = intr8[7];
  t->readIntrinsicParameter(p); return t;
}
void gf2h_tracker_destroy(void* t) { delete (FeatureTracker*)t; }
void gf2h_tracker_set_equalize(void* t, int on) { ((FeatureTracker*)t)->EQUALIZE = on; }   // before the first image
void gf2h_tracker_set_detector(void* t, FeatureTracker::Detector d, void* user) { ((FeatureTracker*)t)->setDetector(d, user); }
void gf2h_tracker_set_prediction(void* t, int n, const int* ids, const double* xyz) {
  std::map<int, Vector3d> m; for (int i = 0; i < n; i++) m[ids[i]] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  ((FeatureTracker*)t)->setPrediction(m);
}
void gf2h_tracker_remove_outliers(void* t, int n, const int* ids) { ((FeatureTracker*)t)->removeOutliers(std::set<int>(ids, ids + n)); }
// returns n; out rows = [id, x, y, 1, u, v, vx, vy, depth, track_cnt]
int gf2h_tracker_track(void* t, double time, const uint8_t* img, const uint16_t* depth, int max_n, double* out10) {
  FeatureTracker* T = (FeatureTracker*)t;
  auto ff = T->trackImage(time, img, depth);
  if (T->lastError()[0]) return -1;
  int k = 0;
  for (size_t i = 0; i < T->ids.size() && k < max_n; i++, k++) {
    const auto& v = ff[T->ids[i]][0].second; double* o = out10 + 10 * k;
    o[0] = T->ids[i]; for (int c = 0; c < 8; c++) o[1 + c] = v[c]; o[9] = T->track_cnt[i];
  }
  return k;
}
int gf2h_estimator_track_only(void* e, double t, const uint8_t* img, const uint16_t* depth, int max_n, double* out10) {
  Estimator* E = (Estimator*)e; if (!E->featureTracker) return -1;
  return gf2h_tracker_track(E->featureTracker, t, img, depth, max_n, out10);
}
int gf2h_tracker_mask(void* t, uint8_t* out) { FeatureTracker* T = (FeatureTracker*)t; memcpy(out, T->mask.data(), T->mask.size()); return (int)T->mask.size(); }
const char* gf2h_tracker_last_error(void* t) { return ((FeatureTracker*)t)->lastError(); }
}
