// vloam_kitti_io.hpp -- KITTI on-disk formats either side of the path, for C++ callers of the C ABI.
//   read_velodyne_bin   : float32 x y z reflectance sweeps (the layout fed to ScanRegistration::input)
//   KittiPoseWriter     : the per-frame evaluation line of VloamTF::{VO,LO,MO}2Cam0StartFrame
//                         (vloam_tf.cpp:84-160): cam0_start_T_cam0_last, cast to float, 12 x "%f"
// Header-only, no dependencies beyond the C++ standard library (the reference uses tf2 / Eigen for this).
#pragma once
#include <array>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace vloam {

inline std::vector<float> read_velodyne_bin(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open " + path);
  fseek(f, 0, SEEK_END);
  const long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (bytes < 0 || bytes % 16) { fclose(f); throw std::runtime_error(path + ": size is not a multiple of 4 floats"); }
  std::vector<float> v((size_t)bytes / 4);
  const size_t got = fread(v.data(), 4, v.size(), f);
  fclose(f);
  if (got != v.size()) throw std::runtime_error(path + ": short read");
  return v;  // n = size() / 4 points, stride 4
}

using Mat4 = std::array<double, 16>;  // row-major
inline Mat4 mat4_identity() { return {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }
inline Mat4 mat4_mul(const Mat4& a, const Mat4& b) {
  Mat4 c{};
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j]; c[i * 4 + j] = s; }
  return c;
}
inline Mat4 mat4_rigid_inverse(const Mat4& m) {  // [R t; 0 1]^-1 = [R^T  -R^T t; 0 1]
  Mat4 r = mat4_identity();
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[i * 4 + j] = m[j * 4 + i];
  for (int i = 0; i < 3; ++i) r[i * 4 + 3] = -(r[i * 4] * m[3] + r[i * 4 + 1] * m[7] + r[i * 4 + 2] * m[11]);
  return r;
}
// Eigen::Quaterniond(x, y, z, w)::toRotationMatrix + translation
inline Mat4 pose_matrix(const double q[4], const double t[3]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  return {1 - (tyy + tzz), txy - twz, txz + twy, t[0], txy + twz, 1 - (txx + tzz), tyz - twx, t[1], txz - twy, tyz + twx, 1 - (txx + tyy), t[2], 0, 0, 0, 1};
}

class KittiPoseWriter {
 public:
  explicit KittiPoseWriter(const std::string& path, const Mat4& base_T_cam0 = mat4_identity())
      : f_(path.empty() ? nullptr : fopen(path.c_str(), "w")), base_T_cam0_(base_T_cam0), cam0_T_base_(mat4_rigid_inverse(base_T_cam0)) {}
  ~KittiPoseWriter() { if (f_) fclose(f_); }
  // world_T_base_last = (q xyzw, t): e.g. the mapped pose of vloam_b200_process_frame (pose_out + 7)
  std::string write(const double q[4], const double t[3]) {
    const Mat4 last = mat4_mul(mat4_mul(cam0_T_base_, pose_matrix(q, t)), base_T_cam0_);
    if (!started_) { start_inv_ = mat4_rigid_inverse(last); started_ = true; }  // count == 0
    const Mat4 m = mat4_mul(start_inv_, last);
    char line[512];
    int o = 0;
    for (int k = 0; k < 12; ++k) o += snprintf(line + o, sizeof line - o, k ? " %f" : "%f", (double)(float)m[k]);
    if (f_) fprintf(f_, "%s\n", line);
    return line;
  }
 private:
  FILE* f_;
  Mat4 base_T_cam0_, cam0_T_base_, start_inv_{};
  bool started_ = false;
};

}  // namespace vloam
